// Stand-alone check of cp.reduce.async.bulk.tensor.3d (.add, f32) with a [19][23][52] box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_test scripts/experiments/tma_reduce_test.cu && /tmp/tma_test
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#ifndef WX
#define WX 52
#define WY 23
#define WZ 19
#endif
#ifndef MODE
#define MODE 0
#endif
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tmap, int cx, int cy, int cz, float* gptr)
{
    extern __shared__ __align__(128) unsigned char raw[];
    float* win = reinterpret_cast<float*>(raw);
    for (int i = threadIdx.x; i < WX * WY * WZ; i += blockDim.x) win[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(win);
#if MODE == 0
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     :: "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(src) : "memory");
#elif MODE == 1
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     :: "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(src) : "memory");
#else
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                     :: "l"(gptr), "r"(src), "r"((unsigned)(WX * 4)) : "memory");
#endif
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncthreads();
}

int main()
{
    const int NX = 256, NY = 64, NZ = 32;
    float* d;
    cudaMalloc(&d, sizeof(float) * NX * NY * NZ);
    cudaMemset(d, 0, sizeof(float) * NX * NY * NZ);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    printf("entry point: %s q=%d ptr=%p\n", cudaGetErrorString(e), (int)q, ptr);
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {NX, NY, NZ};
    const cuuint64_t gstr[2] = {NX * 4ull, (cuuint64_t)NX * NY * 4ull};
    const cuuint32_t box[3] = {WX, WY, WZ};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((enc_fn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    const size_t smem = sizeof(float) * WX * WY * WZ;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("attr: %s\n", cudaGetErrorString(e));
    k<<<1, 256, smem>>>(tm, -4, 5, 30, d);      // clipped at x<0 and z>=NZ
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    k<<<1, 256, smem>>>(tm, 8, 5, 3, d + 1024);
    e = cudaDeviceSynchronize();
    printf("kernel2: %s\n", cudaGetErrorString(e));
    std::vector<float> h(NX * NY * NZ);
    cudaMemcpy(h.data(), d, sizeof(float) * h.size(), cudaMemcpyDeviceToHost);
    double sum = 0;
    for (float v : h) sum += v;
    const double expect = (double)(WX - 4) * WY * 2 + (double)WX * WY * WZ;   // first box clipped to x>=0, z in {30,31}
    printf("sum=%.1f expect=%.1f  h[3,5,8]=%.1f h[31,5,0]=%.1f\n", sum, expect, h[(3 * NY + 5) * NX + 8], h[(31 * NY + 5) * NX + 0]);
    return 0;
}
