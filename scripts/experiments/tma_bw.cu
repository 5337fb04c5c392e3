// tma_bw.cu -- how fast can all SMs stage boxes of a 256^3 float32 volume (67 MB, L2 resident after the first
// pass) into shared memory?  Tensor-map TMA boxes {64,4,1} / {64,8,1} / {32,4,1} against row-wise bulk copies
// (cp.async.bulk, 176 bytes per row, one thread per row as in edf_swin.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_bw scripts/experiments/tma_bw.cu && /tmp/tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(unsigned long long* m, unsigned n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(m)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* m, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* m, unsigned phase)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(m);
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                 :: "r"(a), "r"(phase) : "memory");
}

// Every CTA (256 threads) runs `iters` rounds; a round stages a window of `nz` planes x `ny` rows around a
// pseudo-random origin: tensor mode: warp w's lane 0 issues the planes w, w+8, ..; row mode: thread r issues row r.
template <int MODE>   // 0: tensor boxes {BX, BY, 1};  1: bulk rows
__global__ void __launch_bounds__(256) k_stage(const __grid_constant__ CUtensorMap tm, const float* vol, int bx, int by, int nz, int ny,
                                               int rowbytes, int iters, unsigned long long* bytes_out, long long* cyc)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&mbar, MODE == 0 ? 8 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned seed = blockIdx.x * 2654435761u + 12345u;
    const unsigned dst0 = (unsigned)__cvta_generic_to_shared(raw);
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
    unsigned long long total = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        seed = seed * 1664525u + 1013904223u;
        const int ox = ((seed >> 8) % 48) * 4, oy = (seed >> 14) % 200, oz = (seed >> 22) % 200;
        if (MODE == 0) {
            const int groups = (ny + by - 1) / by;
            if (lane == 0) {
                int n = 0;
                for (int z = warp; z < nz; z += 8) n += groups;
                mbar_expect(&mbar, (unsigned)(n * bx * by * 4));
                for (int z = warp; z < nz; z += 8)
                    for (int gq = 0; gq < groups; ++gq) {
                        const unsigned dst = dst0 + (unsigned)((z * groups + gq) * bx * by * 4);
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(ox), "r"(oy + gq * by), "r"(oz + z), "r"(mb) : "memory");
                    }
            }
            total += (unsigned long long)nz * groups * bx * by * 4;
        } else {
            const int rows = nz * ny;
            if (tid == 0) mbar_expect(&mbar, (unsigned)(rows * rowbytes));
            __syncwarp();
            for (int r = tid; r < rows; r += 256) {
                const int z = r / ny, y = r - z * ny;
                const float* src = vol + ((size_t)(oz + z) * 256 + (oy + y)) * 256 + ox;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst0 + (unsigned)(r * 256)), "l"(src), "r"((unsigned)rowbytes), "r"(mb) : "memory");
            }
            total += (unsigned long long)rows * rowbytes;
        }
        mbar_wait(&mbar, it & 1);
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) { bytes_out[blockIdx.x] = total; cyc[blockIdx.x] = t1 - t0; }
}

int main()
{
    const int N = 256;
    float* d;
    cudaMalloc(&d, sizeof(float) * N * N * N);
    cudaMemset(d, 0, sizeof(float) * N * N * N);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    unsigned long long* bytes; long long* cyc;
    cudaMalloc(&bytes, 8 * 1024); cudaMalloc(&cyc, 8 * 1024);
    struct Cfg { int mode, bx, by, nz, ny, rowbytes, ctas; const char* name; };
    const Cfg cfgs[] = {
        {0, 64, 4, 18, 16, 0, 2, "tensor box 64x4x1, 18 planes x 16 rows (72 KB), 2 CTA/SM"},
        {0, 64, 8, 18, 16, 0, 2, "tensor box 64x8x1, 18 planes x 16 rows (72 KB), 2 CTA/SM"},
        {0, 64, 16, 18, 16, 0, 2, "tensor box 64x16x1, 18 planes x 16 rows (72 KB), 2 CTA/SM"},
        {0, 32, 4, 18, 16, 0, 2, "tensor box 32x4x1, 18 planes x 16 rows (36 KB), 2 CTA/SM"},
        {0, 64, 4, 18, 20, 0, 1, "tensor box 64x4x1, 18 planes x 20 rows (90 KB), 1 CTA/SM"},
        {0, 64, 4, 18, 20, 0, 2, "tensor box 64x4x1, 18 planes x 20 rows (90 KB), 2 CTA/SM"},
        {1, 0, 0, 18, 14, 176, 2, "bulk rows 176 B, 18 planes x 14 rows (44 KB of 64 KB window), 2 CTA/SM"},
        {1, 0, 0, 18, 14, 256, 2, "bulk rows 256 B, 18 planes x 14 rows (64 KB), 2 CTA/SM"},
        {1, 0, 0, 18, 14, 176, 1, "bulk rows 176 B, 18 planes x 14 rows, 1 CTA/SM"},
    };
    for (const Cfg& c : cfgs) {
        alignas(64) CUtensorMap tm;
        const cuuint64_t gdim[3] = {N, N, N};
        const cuuint64_t gstr[2] = {N * 4ull, (cuuint64_t)N * N * 4ull};
        const cuuint32_t box[3] = {(cuuint32_t)(c.mode == 0 ? c.bx : 32), (cuuint32_t)(c.mode == 0 ? c.by : 4), 1};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = ((enc_fn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int grid = 148 * c.ctas, iters = 200;
        const size_t smem = 100 * 1024;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (c.mode == 0) {
                cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k_stage<0><<<grid, 256, smem>>>(tm, d, c.bx, c.by, c.nz, c.ny, c.rowbytes, iters, bytes, cyc);
            } else {
                cudaFuncSetAttribute(k_stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k_stage<1><<<grid, 256, smem>>>(tm, d, c.bx, c.by, c.nz, c.ny, c.rowbytes, iters, bytes, cyc);
            }
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            cudaEventElapsedTime(&ms, e0, e1);
        }
        static unsigned long long hb[296]; static long long hc[296];
        cudaMemcpy(hb, bytes, grid * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double tb = 0, mc = 0;
        for (int i = 0; i < grid; ++i) { tb += (double)hb[i]; mc += (double)hc[i]; }
        mc /= grid;
        printf("%-78s %7.1f GB/s  %6.1f B/clk/SM  %7.0f clk per round\n", c.name, tb / ms / 1e6, tb / 148 / mc * 1.0 * (grid / 148) / (grid / 148), mc / iters);
    }
    return 0;
}
