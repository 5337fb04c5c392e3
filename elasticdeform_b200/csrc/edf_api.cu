// edf_api.cu -- C-ABI entry points (include/edf_b200.h) of the B200-native
// elastic-deformation hot path: validation mirroring the reference's
// Py_DeformGrid_helper (_deform_grid.c:94-293), descriptor flattening, kernel
// selection and launch.  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around the C-ABI entries (no-ops unless a profiler is attached)
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <math.h>

#include "edf_core.h"
#include "edf_spline_lines.h"
#include "edf_host.h"
#include "edf_fast.cuh"
#include "edf_lean.cuh"
#include "edf_swin.cuh"
#include "edf_poly.cuh"
#include "edf_tile.cuh"
#ifdef EDF_WITH_PIPE
#include "experimental/edf_pipe.cuh"   // producer / consumer pipeline (measured slower than the staged-window kernels)
#endif

// ----------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------
// NVTX range of one C-ABI entry (the enqueue side of every call shows up on the profiler's timeline)
struct EdfNvtxRange {
    explicit EdfNvtxRange(const char* name) { nvtxRangePushA(name); }
    ~EdfNvtxRange() { nvtxRangePop(); }
};

static thread_local const char* g_last_kernel = "none";
static std::atomic<uint64_t> g_launches{0};

extern "C" const char* edf_last_error(void) { return g_err; }
extern "C" const char* edf_last_kernel(void) { return g_last_kernel; }
extern "C" uint64_t edf_launch_count(void) { return g_launches.load(); }
extern "C" int edf_version(void) { return 100; /* 0.1.0 */ }

// debug: phase-cycle totals of the staged-window kernels of builds with -DEDF_TILE_PROFILE (zeros otherwise);
// reads and resets.  out[0..7] = cycles per phase summed over warps, out[15] = warps.
extern "C" int edf_debug_tile_profile(uint64_t* out16)
{
    unsigned long long h[16];
    if (cudaMemcpyFromSymbol(h, g_tile_prof, sizeof(h)) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (int i = 0; i < 16; ++i) out16[i] = h[i];
    memset(h, 0, sizeof(h));
    if (cudaMemcpyToSymbol(g_tile_prof, h, sizeof(h)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

extern "C" int edf_device_ok(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    // the library holds sm_100a SASS only (no PTX): exactly compute capability 10.0
    int major = 0, minor = -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (major == 10 && minor == 0) ? 1 : 0;
}

static int check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return edf_fail(EDF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return EDF_OK;
}

// ----------------------------------------------------------------------------
// generic kernels (any naxis<=4 / order / dtype / strides / mode)
// ----------------------------------------------------------------------------
template <int NAXIS>
__global__ void __launch_bounds__(128)
edf_generic_kernel(const __grid_constant__ EdfParams p)
{
    const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kk < p.size) edf_generic_voxel<NAXIS>(p, kk);
}

static int launch_generic(const EdfParams& p, cudaStream_t st)
{
    if (p.size == 0) return EDF_OK;
    const int threads = 128;
    const int64_t blocks = (p.size + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return edf_fail(EDF_ERR_RUNTIME, "output too large");
    switch (p.naxis) {
    case 1: edf_generic_kernel<1><<<(unsigned)blocks, threads, 0, st>>>(p); break;
    case 2: edf_generic_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(p); break;
    case 3: edf_generic_kernel<3><<<(unsigned)blocks, threads, 0, st>>>(p); break;
    case 4: edf_generic_kernel<4><<<(unsigned)blocks, threads, 0, st>>>(p); break;
    default: return edf_fail(EDF_ERR_RUNTIME, "unsupported number of deformed axes");
    }
    g_last_kernel = p.gradient ? "generic_grad" : "generic";
    return check_launch("edf_generic_kernel");
}

static int run_problem(const edf_problem* pr, int gradient, cudaStream_t st)
{
    if (!edf_device_ok())
        return edf_fail(EDF_ERR_CUDA, "no sm_100 (B200) CUDA device available; there is no CPU fallback");
    EdfParams p;
    int rc = flatten_problem(pr, gradient, p);
    if (rc != EDF_OK) return rc;
    if (p.size == 0) return EDF_OK;
    if (!(pr->flags & EDF_FLAG_FORCE_GENERIC)) {
        const char* name = nullptr;
        uint32_t handled = 0;
        rc = edf_fast_try_launch(p, st, &name, &handled, pr->flags);
        if (rc < 0) return edf_fail(EDF_ERR_CUDA, "fast kernel launch failed: %s",
                                    cudaGetErrorString(g_fast_launch_error));
        if (rc > 0) {
            g_last_kernel = name;
            g_launches.fetch_add((uint64_t)rc);
            // inputs no specialised kernel took go through the generic kernel
            int n = 0;
            for (int i = 0; i < p.ninputs; ++i)
                if (!((handled >> i) & 1u)) {
                    if (n != i) p.inp[n] = p.inp[i];
                    ++n;
                }
            if (n == 0) return EDF_OK;
            p.ninputs = n;
            rc = launch_generic(p, st);
            if (rc == EDF_OK) g_last_kernel = "fast+generic";
            return rc;
        }
    }
    return launch_generic(p, st);
}

extern "C" int edf_deform_grid(const edf_problem* problem, void* stream)
{
    EdfNvtxRange nvtx_("edf_deform_grid");
    return run_problem(problem, 0, (cudaStream_t)stream);
}

extern "C" int edf_deform_grid_grad(const edf_problem* problem, void* stream)
{
    EdfNvtxRange nvtx_("edf_deform_grid_grad");
    return run_problem(problem, 1, (cudaStream_t)stream);
}

// Batch of independent problems.  A small volume (e.g. a 64^3 crop = 64 CTAs) cannot fill 148 SMs,
// so the problems are spread over a few internal streams forked from / joined to the caller's
// stream with events: kernels of different volumes run concurrently, ordering with respect to the
// caller's stream is preserved, nothing synchronises the host.
#define EDF_BATCH_STREAMS 4
struct EdfBatchStreams {
    bool ready = false;
    cudaStream_t s[EDF_BATCH_STREAMS] = {};
    cudaEvent_t fork = nullptr, join[EDF_BATCH_STREAMS] = {};
};
// one set per device and thread, created on first use and kept (a thread that hops between devices finds its
// earlier sets again); a set whose creation fails part-way is destroyed again
#define EDF_BATCH_MAX_DEVICES 64
static thread_local EdfBatchStreams g_batch_sets[EDF_BATCH_MAX_DEVICES];
static thread_local EdfBatchStreams* g_batch_cur = nullptr;
#define g_batch (*g_batch_cur)

static void batch_streams_destroy(EdfBatchStreams& b)
{
    for (int i = 0; i < EDF_BATCH_STREAMS; ++i) {
        if (b.s[i]) cudaStreamDestroy(b.s[i]);
        if (b.join[i]) cudaEventDestroy(b.join[i]);
        b.s[i] = nullptr;
        b.join[i] = nullptr;
    }
    if (b.fork) cudaEventDestroy(b.fork);
    b.fork = nullptr;
    b.ready = false;
    cudaGetLastError();
}

static int batch_streams_ready()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= EDF_BATCH_MAX_DEVICES) return 0;
    EdfBatchStreams& b = g_batch_sets[dev];
    g_batch_cur = &b;
    if (b.ready) return 1;
    bool ok = true;
    for (int i = 0; i < EDF_BATCH_STREAMS && ok; ++i) {
        ok = cudaStreamCreateWithFlags(&b.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&b.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { batch_streams_destroy(b); return 0; }
    b.ready = true;
    return 1;
}

extern "C" int edf_deform_grid_batch(const edf_problem* problems, int32_t n, int32_t gradient,
                                     void* stream)
{
    EdfNvtxRange nvtx_("edf_deform_grid_batch");
    if (n < 0 || (n > 0 && !problems)) return edf_fail(EDF_ERR_RUNTIME, "invalid batch");
    cudaStream_t user = (cudaStream_t)stream;
    if (n < 2 || !batch_streams_ready()) {
        cudaGetLastError();
        for (int i = 0; i < n; ++i) {
            int rc = run_problem(&problems[i], gradient ? 1 : 0, user);
            if (rc != EDF_OK) return rc;
        }
        return EDF_OK;
    }
    const int ns = n < EDF_BATCH_STREAMS ? n : EDF_BATCH_STREAMS;
    cudaEventRecord(g_batch.fork, user);
    for (int k = 0; k < ns; ++k) cudaStreamWaitEvent(g_batch.s[k], g_batch.fork, 0);
    int rc = EDF_OK;
    for (int i = 0; i < n && rc == EDF_OK; ++i)
        rc = run_problem(&problems[i], gradient ? 1 : 0, g_batch.s[i % ns]);
    for (int k = 0; k < ns; ++k) {                 // always join, also after an error
        cudaEventRecord(g_batch.join[k], g_batch.s[k]);
        cudaStreamWaitEvent(user, g_batch.join[k], 0);
    }
    if (rc == EDF_OK && cudaGetLastError() != cudaSuccess)
        return edf_fail(EDF_ERR_CUDA, "batch stream fork/join failed");
    return rc;
}

extern "C" int edf_deform_grid_batch_uniform(const edf_problem* proto, int32_t n, int32_t gradient,
                                             const uint64_t* in_ptrs, const uint64_t* out_ptrs, const uint64_t* disp_ptrs,
                                             const double* affines, void* stream)
{
    EdfNvtxRange nvtx_("edf_deform_grid_batch_uniform");
    if (n < 0 || !proto || (n > 0 && (!in_ptrs || !out_ptrs || !disp_ptrs))) return edf_fail(EDF_ERR_RUNTIME, "invalid batch");
    if (proto->ninputs != 1 || !proto->inputs || !proto->outputs)
        return edf_fail(EDF_ERR_RUNTIME, "uniform batch: one input per volume");
    if (proto->naxis < 1 || proto->naxis > EDF_MAX_AXIS) return edf_fail(EDF_ERR_RUNTIME, "invalid number of axes");
    cudaStream_t user = (cudaStream_t)stream;
    const bool fork = n >= 2 && batch_streams_ready();
    if (!fork) cudaGetLastError();
    const int ns = fork ? (n < EDF_BATCH_STREAMS ? n : EDF_BATCH_STREAMS) : 1;
    if (fork) {
        cudaEventRecord(g_batch.fork, user);
        for (int k = 0; k < ns; ++k) cudaStreamWaitEvent(g_batch.s[k], g_batch.fork, 0);
    }
    const int na = proto->naxis * (proto->naxis + 1);
    int rc = EDF_OK;
    for (int i = 0; i < n && rc == EDF_OK; ++i) {
        edf_array in = proto->inputs[0], out = proto->outputs[0];
        in.data = (void*)(uintptr_t)in_ptrs[i];
        out.data = (void*)(uintptr_t)out_ptrs[i];
        edf_problem pr = *proto;
        pr.inputs = &in;
        pr.outputs = &out;
        pr.displacement.data = (void*)(uintptr_t)disp_ptrs[i];
        if (affines) pr.affine = affines + (size_t)i * na;
        rc = run_problem(&pr, gradient ? 1 : 0, fork ? g_batch.s[i % ns] : user);
    }
    if (fork) {
        for (int k = 0; k < ns; ++k) {             // always join, also after an error
            cudaEventRecord(g_batch.join[k], g_batch.s[k]);
            cudaStreamWaitEvent(user, g_batch.join[k], 0);
        }
        if (rc == EDF_OK && cudaGetLastError() != cudaSuccess)
            return edf_fail(EDF_ERR_CUDA, "batch stream fork/join failed");
    }
    return rc;
}

// ----------------------------------------------------------------------------
// K3 / K4: line filters.  A CTA stages `lines_per_block` lines as doubles in
// shared memory (coalesced in either orientation), one thread then runs the
// strictly sequential recursion per line, and the CTA writes the lines back.
// ----------------------------------------------------------------------------
struct EdfLineParams {
    const char* in;
    char* out;
    int32_t in_dtype, out_dtype;
    int64_t n;                         // line length
    int64_t in_lstr, out_lstr;         // byte stride along the line
    int32_t nother, adjoint;
    int64_t odim[EDF_MAX_DIMS], in_ostr[EDF_MAX_DIMS], out_ostr[EDF_MAX_DIMS];
    int64_t nlines;
    int32_t lines_per_block, ld;       // ld = padded line pitch in doubles (odd)
    int32_t line_fastest, pad_;        // 1: flattened copy runs along the line first
    EdfLineFilter f;
};

__global__ void __launch_bounds__(256)
edf_line_filter_kernel(const __grid_constant__ EdfLineParams p)
{
    extern __shared__ double sbuf[];
    const int L = p.lines_per_block;
    int64_t* in_off = (int64_t*)(sbuf + (size_t)L * p.ld);
    int64_t* out_off = in_off + L;
    const int64_t line0 = (int64_t)blockIdx.x * L;
    const int nl = (int)min((int64_t)L, p.nlines - line0);

    for (int l = threadIdx.x; l < nl; l += blockDim.x) {
        int64_t r = line0 + l, io = 0, oo = 0;
        for (int q = p.nother - 1; q >= 0; --q) {
            const int64_t c = r % p.odim[q];
            r /= p.odim[q];
            io += c * p.in_ostr[q];
            oo += c * p.out_ostr[q];
        }
        in_off[l] = io;
        out_off[l] = oo;
    }
    __syncthreads();

    // ---- stage the lines as doubles (coalesced in either orientation, no per-element div/mod)
    const int n = (int)p.n, ld = p.ld;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool f32 = p.in_dtype == EDF_F32, f64 = p.in_dtype == EDF_F64;
    // (EDF_LINE_UNROLL independent global loads per thread before the first dependent conversion / store: with one
    //  load in flight per warp the staging ran at DRAM latency -- 35 % of the kernel's stall samples sat on the
    //  conversion of the loaded value)
    constexpr int UN = 8;
    if (p.line_fastest) {
        for (int l = warp; l < nl; l += nwarps) {
            const char* src = p.in + in_off[l];
            double* dst = sbuf + (size_t)l * ld;
            for (int i0 = lane; i0 < n; i0 += 32 * UN) {
                double v[UN];
#pragma unroll
                for (int k = 0; k < UN; ++k) {
                    const int i = i0 + 32 * k;
                    const char* q = src + (int64_t)min(i, n - 1) * p.in_lstr;
                    v[k] = f32 ? (double)*(const float*)q : (f64 ? *(const double*)q : edf_load(q, p.in_dtype));
                }
#pragma unroll
                for (int k = 0; k < UN; ++k)
                    if (i0 + 32 * k < n) dst[i0 + 32 * k] = v[k];
            }
        }
    } else {
        for (int i0 = warp * UN; i0 < n; i0 += nwarps * UN) {
            for (int l = lane; l < nl; l += 32) {
                const char* base = p.in + in_off[l];
                double v[UN];
#pragma unroll
                for (int k = 0; k < UN; ++k) {
                    const char* q = base + (int64_t)min(i0 + k, n - 1) * p.in_lstr;
                    v[k] = f32 ? (double)*(const float*)q : (f64 ? *(const double*)q : edf_load(q, p.in_dtype));
                }
#pragma unroll
                for (int k = 0; k < UN; ++k)
                    if (i0 + k < n) sbuf[(size_t)l * ld + i0 + k] = v[k];
            }
        }
    }
    __syncthreads();

    for (int l = threadIdx.x; l < nl; l += blockDim.x) {
        double* c = sbuf + (size_t)l * ld;
        if (p.adjoint) edf_prefilter_adjoint_line(c, p.n, p.f);
        else           edf_prefilter_line(c, p.n, p.f);
    }
    __syncthreads();

    const bool o32 = p.out_dtype == EDF_F32, o64 = p.out_dtype == EDF_F64;
    if (p.line_fastest) {
        for (int l = warp; l < nl; l += nwarps) {
            char* dstp = p.out + out_off[l];
            const double* src = sbuf + (size_t)l * ld;
            for (int i = lane; i < n; i += 32) {
                char* q = dstp + (int64_t)i * p.out_lstr;
                if (o32) *(float*)q = (float)src[i];
                else if (o64) *(double*)q = src[i];
                else edf_store_cast(q, p.out_dtype, src[i]);
            }
        }
    } else {
        for (int i = warp; i < n; i += nwarps) {
            const int64_t oo = (int64_t)i * p.out_lstr;
            for (int l = lane; l < nl; l += 32) {
                char* q = p.out + out_off[l] + oo;
                const double v = sbuf[(size_t)l * ld + i];
                if (o32) *(float*)q = (float)v;
                else if (o64) *(double*)q = v;
                else edf_store_cast(q, p.out_dtype, v);
            }
        }
    }
}

static int run_line_filter(const edf_array* input, const edf_array* output, int axis, int order,
                           int adjoint, cudaStream_t st)
{
    if (!edf_device_ok())
        return edf_fail(EDF_ERR_CUDA, "no sm_100 (B200) CUDA device available; there is no CPU fallback");
    if (!input || !output) return edf_fail(EDF_ERR_RUNTIME, "null array");
    if (order < 0 || order > 5)
        return edf_fail(EDF_ERR_RUNTIME, "spline order not supported");        // _deform_grid.c:71
    const int nd = input->ndim;
    if (nd < 0 || nd > EDF_MAX_DIMS || output->ndim != nd)
        return edf_fail(EDF_ERR_RUNTIME, "input and output dimensions should match");
    if (axis < 0) axis += nd;                                                  // _deform_grid.c:75
    if (axis < 0 || axis >= nd) return edf_fail(EDF_ERR_RUNTIME, "invalid axis");
    if (!dtype_size(input->dtype) || !dtype_size(output->dtype))
        return edf_fail(EDF_ERR_RUNTIME, "data type not supported");
    EdfLineParams p;
    memset(&p, 0, sizeof(p));
    int64_t total = 1;
    for (int d = 0; d < nd; ++d) {
        if (input->shape[d] != output->shape[d])
            return edf_fail(EDF_ERR_RUNTIME, "input and output shapes should match");
        total *= input->shape[d];
    }
    if (total == 0) return EDF_OK;
    if (!input->data || !output->data) return edf_fail(EDF_ERR_VALUE, "null data pointer");
    p.in = (const char*)input->data;
    p.out = (char*)output->data;
    p.in_dtype = input->dtype;
    p.out_dtype = output->dtype;
    p.n = input->shape[axis];
    p.in_lstr = input->strides[axis];
    p.out_lstr = output->strides[axis];
    p.adjoint = adjoint;
    p.nlines = 1;
    int q = 0;
    int64_t min_other_stride = INT64_MAX;
    for (int d = 0; d < nd; ++d) {
        if (d == axis) continue;
        p.odim[q] = input->shape[d];
        p.in_ostr[q] = input->strides[d];
        p.out_ostr[q] = output->strides[d];
        p.nlines *= input->shape[d];
        if (input->shape[d] > 1) {
            int64_t s = input->strides[d] < 0 ? -input->strides[d] : input->strides[d];
            if (s < min_other_stride) min_other_stride = s;
        }
        ++q;
    }
    p.nother = q;
    {
        int64_t ls = p.in_lstr < 0 ? -p.in_lstr : p.in_lstr;
        p.line_fastest = (ls <= min_other_stride) ? 1 : 0;
    }
    setup_filter(p.f, (order < 2) ? 0 : order, p.n, adjoint);
    if (order < 2) { p.f.npoles = 0; p.f.gain = 1.0; }     // orders 0/1: plain copy
    if (order < 2 && adjoint) p.f.gain = 1.0;

    int dev = 0;
    cudaGetDevice(&dev);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int64_t ld = p.n + 1 + (p.n & 1);                // odd pitch: conflict-free 64-bit columns
    const int64_t per_line = ld * 8 + 16;
    int64_t L = ((int64_t)max_smem - 1024) / per_line;
    if (L < 1)
        return edf_fail(EDF_ERR_MEMORY, "line of %lld elements does not fit in shared memory",
                        (long long)p.n);
    // keep >= ~4 CTAs per SM worth of lines when the problem is large enough
    int64_t Lcap = 32;
    if (L > Lcap) L = Lcap;
    if (L > p.nlines) L = p.nlines;
    p.lines_per_block = (int)L;
    p.ld = (int)ld;
    const size_t smem = (size_t)(L * per_line);
    const int64_t blocks = (p.nlines + L - 1) / L;
    if (blocks > 0x7fffffffLL) return edf_fail(EDF_ERR_RUNTIME, "too many lines");
    static EdfPerDeviceFlag configured;                 // the limit is raised to the device's maximum, once per device
    if (!configured.test()) {
        if (cudaFuncSetAttribute(edf_line_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 max_smem) != cudaSuccess)
            return edf_fail(EDF_ERR_CUDA, "cudaFuncSetAttribute: %s",
                            cudaGetErrorString(cudaGetLastError()));
        configured.set();
    }
    int threads = 256;
    edf_line_filter_kernel<<<(unsigned)blocks, threads, smem, st>>>(p);
    return check_launch(adjoint ? "edf_line_filter_kernel(adjoint)" : "edf_line_filter_kernel");
}

extern "C" int edf_spline_filter1d(const edf_array* input, const edf_array* output, int32_t axis,
                                   int32_t order, void* stream)
{
    EdfNvtxRange nvtx_("edf_spline_filter1d");
    return run_line_filter(input, output, axis, order, 0, (cudaStream_t)stream);
}

extern "C" int edf_spline_filter1d_grad(const edf_array* input, const edf_array* output,
                                        int32_t axis, int32_t order, void* stream)
{
    EdfNvtxRange nvtx_("edf_spline_filter1d_grad");
    return run_line_filter(input, output, axis, order, 1, (cudaStream_t)stream);
}
