// ubench2.cu -- shared-memory load width / occupancy sweep and fp64 conversion rates (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench2 scripts/experiments/ubench2.cu && /tmp/ubench2
// Reports warp-instructions per clock per SM and bytes per clock per SM, from the kernel's CUDA-event time at
// the measured SM clock (clock64 over the loop), for 1, 2 and 4 CTAs of 256 threads per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int THREADS = 256;
constexpr int ITER = 4096;
constexpr int UNR = 16;

template <int W>   // W = 1, 2, 4 words per lane
__device__ __forceinline__ int lds_w(unsigned addr)
{
    int a, b, c, d;
    if (W == 1) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(addr)); return a; }
    if (W == 2) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr)); return a + b; }
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
    return a + b + c + d;
}

// PAT 0: lane -> consecutive W-word groups (conflict free)
// PAT 1: "gather" pattern: lane l reads the aligned W-word group that contains word (5 * l / 4 + row_skew), i.e.
//        neighbouring lanes often share a group (stretch 1.25, realistic for the interpolation windows)
template <int W, int PAT>
__global__ void k_lds(int* out, long long* cyc)
{
    extern __shared__ __align__(16) int s[];
    for (int i = threadIdx.x; i < 12288; i += THREADS) s[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int word = (PAT == 0) ? lane * W : ((5 * lane / 4) & ~(W - 1));
    word += warp * 256;
    const unsigned base = (unsigned)__cvta_generic_to_shared(s + word);
    int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; u += 4) {
            acc0 += lds_w<W>(base + (u + 0) * 1024 + (it & 1) * 512);
            acc1 += lds_w<W>(base + (u + 1) * 1024 + (it & 1) * 512);
            acc2 += lds_w<W>(base + (u + 2) * 1024 + (it & 1) * 512);
            acc3 += lds_w<W>(base + (u + 3) * 1024 + (it & 1) * 512);
        }
    }
    const long long t1 = clock64();
    if (acc0 + acc1 + acc2 + acc3 == 0x7fffffff) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PAT>
__global__ void k_atoms(int* out, long long* cyc)
{
    extern __shared__ __align__(16) int s[];
    for (int i = threadIdx.x; i < 12288; i += THREADS) s[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int word = (PAT == 0) ? lane : (5 * lane / 4);
    word += warp * 256;
    const unsigned base = (unsigned)__cvta_generic_to_shared(s + word);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(base + u * 1024 + (it & 1) * 512), "r"(it + u) : "memory");
    }
    const long long t1 = clock64();
    __syncthreads();
    if (s[threadIdx.x] == 0x7fffffff) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// KIND 0: floor(double) (FRND.F64), 1: (int)double (F2I.S32.F64), 2: (float)double (F2F.F32.F64),
// 3: DADD, 4: DSETP + select, 5: magic-number floor split (3 DADD), 6: FADD
template <int KIND>
__global__ void k_cvt(int* out, long long* cyc, double a)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1.25e-3 + i * 7.5 + a;
    double accd = 0.0;
    float accf = 0.f;
    int acci = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            double& v = x[u & 7];
            if (KIND == 0) v = floor(v) + 0.3;                       // FRND + DADD
            if (KIND == 1) { acci += (int)v; v = __longlong_as_double(__double_as_longlong(v) ^ (long long)acci); }
            if (KIND == 2) { accf += (float)v; v = __longlong_as_double(__double_as_longlong(v) ^ (long long)it); }
            if (KIND == 3) v = __dadd_rn(v, a);
            if (KIND == 4) { v = (v > a) ? a : __longlong_as_double(__double_as_longlong(v) + 1); }
            if (KIND == 5) { const double M = 6755399441055744.0; const double t = __dadd_rd(v, M); const double fl = __dsub_rn(t, M);
                             acci += __double2loint(t); v = __dsub_rn(v, fl) + 1.7; }
            if (KIND == 6) { accf = accf + (float)u; }
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) accd += x[i];
    if (accd + accf + acci == 12345.678) out[0] = 1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kernel, int ctas_per_sm, double instr_per_iter, double bytes_per_instr, size_t smem)
{
    int* out; long long* cyc;
    const int grid = 148 * ctas_per_sm;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, grid * sizeof(long long));
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
    kernel<<<grid, THREADS, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    kernel<<<grid, THREADS, smem>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 8];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += (double)h[i];
    mean /= grid;
    const double winstr = (double)ctas_per_sm * (THREADS / 32) * ITER * instr_per_iter;   // per SM
    printf("%-40s %d CTA/SM: %6.3f warp-instr/clk/SM", name, ctas_per_sm, winstr / mean);
    if (bytes_per_instr > 0) printf("  %7.1f B/clk/SM", winstr / mean * bytes_per_instr);
    printf("  (%s)\n", cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

template <int KIND>
__global__ void k_cvt_w(int* out, long long* cyc) { }

int main()
{
    const size_t smem = 12288 * 4;
    for (int c = 1; c <= 4; c *= 2) {
        run("LDS.32  consecutive", k_lds<1, 0>, c, UNR, 128, smem);
        run("LDS.64  consecutive", k_lds<2, 0>, c, UNR, 256, smem);
        run("LDS.128 consecutive", k_lds<4, 0>, c, UNR, 512, smem);
        run("LDS.32  gather pattern (stretch 1.25)", k_lds<1, 1>, c, UNR, 0, smem);
        run("LDS.64  gather pattern (aligned pairs)", k_lds<2, 1>, c, UNR, 0, smem);
        run("LDS.128 gather pattern (aligned quads)", k_lds<4, 1>, c, UNR, 0, smem);
        run("RED.shared.add.u32 consecutive", k_atoms<0>, c, UNR, 0, smem);
        run("RED.shared.add.u32 gather pattern", k_atoms<1>, c, UNR, 0, smem);
    }
    // fp64 conversions etc.: 2 CTAs per SM
    {
        int* out; long long* cyc;
        cudaMalloc(&out, 4); cudaMalloc(&cyc, 296 * 8);
        static long long h[296];
        const char* names[7] = {"FRND.F64 (floor) + DADD", "F2I.S32.F64 + LOP", "F2F.F32.F64 + FADD + LOP", "DADD", "DSETP + SEL",
                                "floor split (3 DADD)", "FADD"};
        for (int k = 0; k < 7; ++k) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (k) {
                case 0: k_cvt<0><<<296, THREADS>>>(out, cyc, 0.5); break;
                case 1: k_cvt<1><<<296, THREADS>>>(out, cyc, 0.5); break;
                case 2: k_cvt<2><<<296, THREADS>>>(out, cyc, 0.5); break;
                case 3: k_cvt<3><<<296, THREADS>>>(out, cyc, 0.5); break;
                case 4: k_cvt<4><<<296, THREADS>>>(out, cyc, 0.5); break;
                case 5: k_cvt<5><<<296, THREADS>>>(out, cyc, 0.5); break;
                default: k_cvt<6><<<296, THREADS>>>(out, cyc, 0.5); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < 296; ++i) mean += (double)h[i];
            mean /= 296;
            printf("%-40s %8.3f clk per (warp x source-level op) per SM\n", names[k], mean / (2.0 * 8 * ITER * UNR));
        }
    }
    return 0;
}
