// ubench.cu -- pipe-rate microbenchmarks that decide the kernel design (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench scripts/experiments/ubench.cu && /tmp/ubench
// Every test runs 148*k CTAs of 256 threads (2 per SM) and reports cycles per warp-instruction per SM,
// i.e. the reciprocal throughput of the SM-wide pipe (1.0 = one warp instruction per clock per SM).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int THREADS = 256;
constexpr int ITER = 2048;
constexpr int UNR = 16;

// mode: 0 conflict-free (lane -> bank lane), 1 two lanes per bank at different addresses (2-way conflict),
//       2 pairs of lanes on the SAME address, 3 random-ish spread (stride 33 -> conflict free, other rows)
__device__ __forceinline__ int lane_index(int mode, int lane, int warp)
{
    switch (mode) {
    case 0: return lane + warp * 64;
    case 1: return (lane & 15) + (lane >> 4) * 32 + warp * 64;
    case 2: return (lane >> 1) + warp * 64;
    default: return (lane * 33) % 2048 + warp * 7;
    }
}

__global__ void k_lds(int mode, int* out, long long* cyc)
{
    __shared__ int s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* p = s + lane_index(mode, lane, warp);
    int acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += *(volatile const int*)(p + u * 128);
    }
    const long long t1 = clock64();
    if (acc == 0x7fffffff) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds64(int* out, long long* cyc)
{
    __shared__ __align__(16) int s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int2* p = reinterpret_cast<const int2*>(s) + lane + warp * 32;
    int acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            int vx, vy;
            asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(vx), "=r"(vy) : "r"((unsigned)__cvta_generic_to_shared(p + u * 64)));
            acc += vx + vy;
        }
    }
    const long long t1 = clock64();
    if (acc == 0x7fffffff) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_atoms(int mode, int* out, long long* cyc)
{
    __shared__ int s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* p = s + lane_index(mode, lane, warp);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) atomicAdd(p + u * 128, it + u);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (s[threadIdx.x] == 0x7fffffff) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// float shared atomics (CAS loop on sm_100?) for the record
__global__ void k_atoms_f32(int* out, long long* cyc)
{
    __shared__ float s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* p = s + lane + warp * 64;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) atomicAdd(p + u * 128, 1.0f);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (s[threadIdx.x] == -1.f) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) * 4;
}

__global__ void k_sts(int* out, long long* cyc)
{
    __shared__ int s[8192];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    volatile int* p = s + lane + warp * 64;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) p[u * 128] = it + u;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (s[threadIdx.x] == 0x7fffffff) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// non-atomic read-modify-write (LDS + IADD + STS) for comparison with ATOMS
__global__ void k_rmw(int* out, long long* cyc)
{
    __shared__ int s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    volatile int* p = s + lane + warp * 64;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) p[u * 128] = p[u * 128] + it;
    }
    const long long t1 = clock64();
    __syncthreads();
    if (s[threadIdx.x] == 0x7fffffff) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl(int* out, long long* cyc)
{
    int v = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) v += __shfl_down_sync(0xffffffffu, v, 1 + (u & 3));
    }
    const long long t1 = clock64();
    if (v == 0x7fffffff) out[0] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// fp64: 8 independent DFMA chains per thread
__global__ void k_dfma(int* out, long long* cyc, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) x[u & 7] = fma(x[u & 7], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// fp64 dependent chain latency (1 warp per CTA)
__global__ void k_dfma_lat(int* out, long long* cyc, double a, double b)
{
    double x = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) x = fma(x, a, b);
    }
    const long long t1 = clock64();
    if (x == 12345.678) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_f2f(int* out, long long* cyc, double a)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i + a;
    float acc = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) { acc += (float)x[u & 7]; x[u & 7] = __longlong_as_double(__double_as_longlong(x[u & 7]) ^ (long long)it); }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma(int* out, long long* cyc, float a, float b)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) x[u & 7] = fmaf(x[u & 7], a, b);
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678f) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: 1 LDS + 1 FFMA alternating (the order-3 gather core)
__global__ void k_lds_ffma(int* out, long long* cyc, float w)
{
    __shared__ float s[8192];
    for (int i = threadIdx.x; i < 8192; i += THREADS) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* p = s + lane + warp * 64;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; u += 4) {
            acc0 = fmaf(*(volatile const float*)(p + u * 128), w, acc0);
            acc1 = fmaf(*(volatile const float*)(p + u * 128 + 128), w, acc1);
            acc2 = fmaf(*(volatile const float*)(p + u * 128 + 256), w, acc2);
            acc3 = fmaf(*(volatile const float*)(p + u * 128 + 384), w, acc3);
        }
    }
    const long long t1 = clock64();
    if (acc0 + acc1 + acc2 + acc3 == 12345.678f) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static int run(const char* name, int warps_per_cta, F launch)
{
    int* out;
    long long* cyc;
    const int grid = 148 * 2;
    CHECK(cudaMalloc(&out, 4));
    CHECK(cudaMalloc(&cyc, grid * sizeof(long long)));
    launch(grid, out, cyc);                               // warm
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch(grid, out, cyc);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    static long long h[148 * 2];
    CHECK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    // per SM: 2 CTAs x warps_per_cta warps, each ITER*UNR instructions of the tested kind
    const double winstr = 2.0 * warps_per_cta * ITER * UNR;
    printf("%-34s %8.3f cyc per warp-instr per SM   (kernel %.3f ms, %.0f cyc per CTA)\n", name, mean / winstr, ms, mean);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main()
{
    cudaDeviceProp pr;
    CHECK(cudaGetDeviceProperties(&pr, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", pr.name, pr.multiProcessorCount, pr.major, pr.minor);
    run("LDS.32 conflict-free", 8, [](int g, int* o, long long* c) { k_lds<<<g, THREADS>>>(0, o, c); });
    run("LDS.32 2-way bank conflict", 8, [](int g, int* o, long long* c) { k_lds<<<g, THREADS>>>(1, o, c); });
    run("LDS.32 pairs same address", 8, [](int g, int* o, long long* c) { k_lds<<<g, THREADS>>>(2, o, c); });
    run("LDS.32 stride 33", 8, [](int g, int* o, long long* c) { k_lds<<<g, THREADS>>>(3, o, c); });
    run("LDS.64 conflict-free", 8, [](int g, int* o, long long* c) { k_lds64<<<g, THREADS>>>(o, c); });
    run("STS.32 conflict-free", 8, [](int g, int* o, long long* c) { k_sts<<<g, THREADS>>>(o, c); });
    run("ATOMS.ADD conflict-free", 8, [](int g, int* o, long long* c) { k_atoms<<<g, THREADS>>>(0, o, c); });
    run("ATOMS.ADD 2-way bank conflict", 8, [](int g, int* o, long long* c) { k_atoms<<<g, THREADS>>>(1, o, c); });
    run("ATOMS.ADD pairs same address", 8, [](int g, int* o, long long* c) { k_atoms<<<g, THREADS>>>(2, o, c); });
    run("ATOMS.ADD stride 33", 8, [](int g, int* o, long long* c) { k_atoms<<<g, THREADS>>>(3, o, c); });
    run("atomicAdd(float) shared", 8, [](int g, int* o, long long* c) { k_atoms_f32<<<g, THREADS>>>(o, c); });
    run("LDS+IADD+STS (non-atomic rmw)", 8, [](int g, int* o, long long* c) { k_rmw<<<g, THREADS>>>(o, c); });
    run("SHFL.DOWN", 8, [](int g, int* o, long long* c) { k_shfl<<<g, THREADS>>>(o, c); });
    run("DFMA 8 chains", 8, [](int g, int* o, long long* c) { k_dfma<<<g, THREADS>>>(o, c, 1.0000001, 1e-9); });
    run("FFMA 8 chains", 8, [](int g, int* o, long long* c) { k_ffma<<<g, THREADS>>>(o, c, 1.0000001f, 1e-9f); });
    run("F2F.F32.F64", 8, [](int g, int* o, long long* c) { k_f2f<<<g, THREADS>>>(o, c, 0.5); });
    run("LDS + FFMA pairs (per pair)", 8, [](int g, int* o, long long* c) { k_lds_ffma<<<g, THREADS>>>(o, c, 0.5f); });
    {   // latency: one warp per CTA; cycles per instruction of the dependent chain
        int* out; long long* cyc;
        cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
        k_dfma_lat<<<1, 32>>>(out, cyc, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s %8.3f cyc (dependent chain)\n", "DFMA latency", (double)h / (ITER * UNR));
    }
    return 0;
}
