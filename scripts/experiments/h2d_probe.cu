// Host->device bandwidth on this box: copy engine (1..4 concurrent streams) against a copy KERNEL that
// reads pinned host memory directly (UVA), for buffers first-touched on each NUMA node.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o h2d_probe h2d_probe.cu && ./h2d_probe
#include <cuda_runtime.h>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void bind_node(int node)
{
    cpu_set_t set; CPU_ZERO(&set);
    // nodes of the pool's hosts: node0 = cpus 0-31,64-95; node1 = 32-63,96-127
    for (int c = 0; c < 128; ++c) { const int n = ((c % 64) >= 32) ? 1 : 0; if (node < 0 || n == node) CPU_SET(c, &set); }
    sched_setaffinity(0, sizeof(set), &set);
}

int main()
{
    const size_t N = 64ull << 20;
    char* d; CK(cudaMalloc(&d, N));
    cudaStream_t st[4]; for (auto& s : st) CK(cudaStreamCreate(&s));
    for (int node = 0; node < 2; ++node) {
        bind_node(node);
        char* h; CK(cudaHostAlloc(&h, N, cudaHostAllocDefault));
        memset(h, 1, N);
        for (int ns = 1; ns <= 4; ns *= 2) {
            double best = 0;
            for (int rep = 0; rep < 6; ++rep) {
                CK(cudaDeviceSynchronize());
                const double t0 = now();
                for (int it = 0; it < 5; ++it)
                    for (int s = 0; s < ns; ++s)
                        CK(cudaMemcpyAsync(d + s * (N / ns), h + s * (N / ns), N / ns, cudaMemcpyHostToDevice, st[s]));
                CK(cudaDeviceSynchronize());
                const double bw = 5.0 * N / (now() - t0) / 1e9;
                if (bw > best) best = bw;
            }
            printf("node %d  copy engine, %d stream(s): %.1f GB/s\n", node, ns, best);
        }
        for (int chunk_mb = 1; chunk_mb <= 8; chunk_mb *= 8) {       // many small copies on one stream (slab pipeline)
            CK(cudaDeviceSynchronize());
            const size_t cs = (size_t)chunk_mb << 20;
            const double t0 = now();
            for (int it = 0; it < 3; ++it)
                for (size_t o = 0; o < N; o += cs) CK(cudaMemcpyAsync(d + o, h + o, cs, cudaMemcpyHostToDevice, st[0]));
            CK(cudaDeviceSynchronize());
            printf("node %d  copy engine, %d MB pieces: %.1f GB/s\n", node, chunk_mb, 3.0 * N / (now() - t0) / 1e9);
        }
        for (int blocks = 37; blocks <= 1184; blocks *= 2) {
            double best = 0;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaDeviceSynchronize());
                const double t0 = now();
                for (int it = 0; it < 3; ++it) copy_kernel<<<blocks, 512, 0, st[0]>>>((const int4*)h, (int4*)d, N / 16);
                CK(cudaDeviceSynchronize());
                const double bw = 3.0 * N / (now() - t0) / 1e9;
                if (bw > best) best = bw;
            }
            printf("node %d  copy kernel from pinned host memory, %d CTAs: %.1f GB/s\n", node, blocks, best);
        }
        // device -> host for reference
        {
            CK(cudaDeviceSynchronize());
            const double t0 = now();
            for (int it = 0; it < 5; ++it) CK(cudaMemcpyAsync(h, d, N, cudaMemcpyDeviceToHost, st[0]));
            CK(cudaDeviceSynchronize());
            printf("node %d  copy engine D2H: %.1f GB/s\n", node, 5.0 * N / (now() - t0) / 1e9);
            const double t1 = now();
            for (int it = 0; it < 3; ++it) copy_kernel<<<296, 512, 0, st[0]>>>((const int4*)d, (int4*)h, N / 16);
            CK(cudaDeviceSynchronize());
            printf("node %d  copy kernel D2H (writes to pinned host memory): %.1f GB/s\n", node, 3.0 * N / (now() - t1) / 1e9);
        }
        CK(cudaFreeHost(h));
    }
    return 0;
}
