// edf_fast.cuh -- specialised sm_100a kernels for the hot configurations
// (2-D / 3-D deformations of float32 volumes at spline orders 0..5, and order-0
// "label" volumes of any 1/2/4/8-byte type), forward gather and gradient scatter.
//
// Work decomposition: a CTA of 256 threads owns a tile of 64 (x) x 8 (rows) x 4
// (slabs) output voxels.  The coarse-grid B-spline displacement is contracted
// separably per tile (edf_fast_core.h) so a voxel needs only 4*naxis fp64 FMAs; all
// discrete decisions are kept bit-identical to the reference by re-evaluating the
// rare voxels that sit within 2e-8 of a rounding / boundary threshold in the exact
// reference order.  Interpolation weights and the (order+1)^naxis-tap gather run in
// fp32 (float32 data); label copies move raw bits.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "edf_core.h"
#include "edf_fast_core.h"

// "Configured on this device" flag for cudaFuncSetAttribute: function attributes are PER DEVICE, so a process
// that drives several GPUs has to raise the dynamic shared-memory limit of a kernel on each of them.  test() /
// set() refer to the calling thread's current device; two threads may both configure (harmless), none launches
// before the attributes are in place.  Devices beyond 127 are configured on every launch.
struct EdfPerDeviceFlag {
    std::atomic<uint64_t> mask[2];
    static int device() { int d = 0; return cudaGetDevice(&d) == cudaSuccess ? d : -1; }
    bool test() const
    {
        const int d = device();
        return d >= 0 && d < 128 && ((mask[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1ull);
    }
    void set()
    {
        const int d = device();
        if (d >= 0 && d < 128) mask[d >> 6].fetch_or(1ull << (d & 63), std::memory_order_release);
    }
};

#define EDF_FAST_G 4               // thread groups (slabs in 3-D, row blocks in 2-D)
#define EDF_FAST_M 8               // rows per group and chunk
#define EDF_FAST_RY 64             // rows (second-last axis) a CTA walks through, chunk by chunk
#define EDF_FAST_THREADS (EDF_FAST_TX * EDF_FAST_G)

// Tile schedule of the staged-window kernels (edf_swin.cuh): a 1-D grid whose CTAs are dispatched in
// block-index order, cut into up to three segments of z-tiles with decreasing rows per CTA, so that
// the CTAs that run last -- the tail during which SMs go idle one by one -- are short ones.
struct EdfTileSched {
    uint32_t nseg, gx;             // segments in use; tiles along x
    uint32_t cta_begin[4];         // first block index of segment s (cta_begin[nseg] = grid size)
    uint32_t z_begin[4];           // first z-tile of segment s
    uint32_t ry[3], gy[3];         // rows per CTA and tiles along y of segment s
};

struct EdfFastLaunch {
    uint32_t input_mask;           // which p.inp[] entries this launch processes
    uint32_t rows_per_cta;         // lean kernels: rows of the second-last axis per CTA (multiple of 8, <= 64)
    EdfTileSched sched;            // staged-window kernels only
    uint64_t cval_bits[EDF_MAX_INPUTS];   // constant value converted to the output dtype
    int32_t  istr_e[EDF_MAX_INPUTS][EDF_MAX_AXIS];   // element strides (deformed axes)
    int32_t  ostr_e[EDF_MAX_INPUTS][EDF_MAX_AXIS];
};

// rows of the second-last axis processed per chunk: 3-D: M rows for each of the G slabs,
// 2-D: G*M consecutive rows
template <int NAXIS>
struct EdfFastGeom {
    static constexpr int CHUNK_ROWS = (NAXIS == 3) ? EDF_FAST_M : EDF_FAST_G * EDF_FAST_M;
    static constexpr int NCHUNK = EDF_FAST_RY / CHUNK_ROWS;
    static constexpr int NSLAB = (NAXIS == 3) ? EDF_FAST_G : 1;
};

template <int NAXIS>
struct EdfFastSmem {
    double wz[EDF_FAST_G][4];
    double wy[EDF_FAST_RY][4];
    double wx[EDF_FAST_TX][4];
    int    sz[EDF_FAST_G];
    int    sy[EDF_FAST_RY];
    int    sx[EDF_FAST_TX];
    int    ny, nx, nonzero, pad_;
    double A[NAXIS][EdfFastGeom<NAXIS>::NSLAB][EDF_FAST_NC][EDF_FAST_NC];
    double B[2][NAXIS][EDF_FAST_G][EDF_FAST_M][EDF_FAST_NC];     // double buffered per chunk
};

__device__ __forceinline__ int edf_mirror_index32(int idx, int len)
{
    if ((unsigned)idx < (unsigned)len) return idx;           // interior: no division
    return (int)edf_mirror_index((int64_t)idx, (int64_t)len);
}

// ---------------------------------------------------------------------------------------
// CTA prologue: per-axis control tables for the whole tile, then the z-contraction A of
// the displacement coefficients restricted to the control points the tile touches
// ---------------------------------------------------------------------------------------
template <int NAXIS>
__device__ __forceinline__ void edf_fast_tile_setup(const EdfParams& p, EdfFastSmem<NAXIS>& s,
                                                    int64_t z0, int64_t y0, int64_t x0)
{
    constexpr int AX = NAXIS - 1, AY = NAXIS - 2;
    constexpr int NSLAB = EdfFastGeom<NAXIS>::NSLAB;
    const int tid = threadIdx.x;
    if (tid == 0) s.nonzero = 0;
    if (tid < EDF_FAST_TX) {
        const int64_t o = min(x0 + tid, p.odim[AX] - 1);
        edf_fast_ctrl_entry(p, AX, o, s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_FAST_TX + EDF_FAST_RY) {
        const int t = tid - EDF_FAST_TX;
        const int64_t o = min(y0 + t, p.odim[AY] - 1);
        edf_fast_ctrl_entry(p, AY, o, s.wy[t], &s.sy[t]);
    } else if (NAXIS == 3 && tid < EDF_FAST_TX + EDF_FAST_RY + EDF_FAST_G) {
        const int t = tid - EDF_FAST_TX - EDF_FAST_RY;
        const int64_t o = min(z0 + t, p.odim[0] - 1);
        edf_fast_ctrl_entry(p, 0, o, s.wz[t], &s.sz[t]);
    }
    __syncthreads();
    const int sy_min = s.sy[0], sx_min = s.sx[0];
    const int ny = s.sy[EDF_FAST_RY - 1] - sy_min + 4;        // control rows / columns touched
    const int nx = s.sx[EDF_FAST_TX - 1] - sx_min + 4;        // (<= EDF_FAST_NC, host-checked)
    if (tid == 0) { s.ny = ny; s.nx = nx; }
    bool nz = false;
    const int na = NAXIS * NSLAB * ny * nx;
    for (int e = tid; e < na; e += EDF_FAST_THREADS) {
        const int jx = e % nx;
        const int jy = (e / nx) % ny;
        const int t = (e / (nx * ny)) % NSLAB;
        const int h = e / (nx * ny * NSLAB);
        const int my = edf_mirror_index32(sy_min + jy, (int)p.ncp[AY]);
        const int mx = edf_mirror_index32(sx_min + jx, (int)p.ncp[AX]);
        double a = 0.0;
        if (NAXIS == 3) {
            const char* base = p.disp + p.dstr[0] * h + my * p.dstr[2] + mx * p.dstr[3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int mz = edf_mirror_index32(s.sz[t] + i, (int)p.ncp[0]);
                const double c = (p.ddtype == EDF_F64) ? *(const double*)(base + mz * p.dstr[1])
                                                       : (double)*(const float*)(base + mz * p.dstr[1]);
                nz |= (c != 0.0);
                a = fma(c, s.wz[t][i], a);
            }
        } else {
            const char* q = p.disp + p.dstr[0] * h + my * p.dstr[1] + mx * p.dstr[2];
            a = (p.ddtype == EDF_F64) ? *(const double*)q : (double)*(const float*)q;
            nz |= (a != 0.0);
        }
        s.A[h][t][jy][jx] = a;
    }
    if (nz) s.nonzero = 1;                      // benign race: all writers store 1
    __syncthreads();
}

// y-contraction B of chunk c (rows c*CHUNK_ROWS ...) into buffer `buf`
template <int NAXIS>
__device__ __forceinline__ void edf_fast_chunk_setup(EdfFastSmem<NAXIS>& s, int c, int buf)
{
    const int nx = s.nx;
    const int sy_min = s.sy[0];
    const int nb = NAXIS * EDF_FAST_G * EDF_FAST_M * nx;
    for (int e = threadIdx.x; e < nb; e += EDF_FAST_THREADS) {
        const int jx = e % nx;
        const int m = (e / nx) % EDF_FAST_M;
        const int g = (e / (nx * EDF_FAST_M)) % EDF_FAST_G;
        const int h = e / (nx * EDF_FAST_M * EDF_FAST_G);
        const int row = c * EdfFastGeom<NAXIS>::CHUNK_ROWS + ((NAXIS == 3) ? m : g * EDF_FAST_M + m);
        const int t = (NAXIS == 3) ? g : 0;
        const int r0 = s.sy[row] - sy_min;
        double b = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) b = fma(s.A[h][t][r0 + j][jx], s.wy[row][j], b);
        s.B[buf][h][g][m][jx] = b;
    }
}

// Exact reference-order displacement (deform.c:650-758) of one voxel, out of line: it runs for the ~1 voxel in 10^5
// that sits next to a rounding threshold -- but a warp that hits one waits for it, and its CTA with it, so the
// latency of this routine is on the critical path of every fast kernel (ncu, round 2: 400 warp-level calls of the
// generic form -- a 64-iteration loop with local-memory index tables and a 13-way dtype switch per tap, ~10^5
// cycles each -- accounted for a third of the stall samples of the order-0 forward kernel).  For float64 / float32
// control points (what the fast kernels accept) the taps run four at a time with independent loads and no dtype
// switch; the products ((D * w_0) * w_1) * w_2 and the running sum keep the reference's order and rounding (no FMA).
template <int NAXIS, typename TD>
__device__ __forceinline__ void edf_displacement_exact_unrolled(const EdfParams& p, const int* o, double* dd)
{
    double dw[NAXIS][4];
    int64_t doff[NAXIS][4];
#pragma unroll
    for (int a = 0; a < NAXIS; ++a) {
        const double cp = edf_control_pos(p, a, (int64_t)o[a]);
        const int64_t start = (int64_t)floor(cp) - 1;
        edf_bspline_weights(cp, 3, dw[a]);
        const bool edge = start < 0 || start + 3 >= p.ncp[a];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            int64_t idx = start + l;
            if (edge) idx = edf_mirror_index(idx, p.ncp[a]);
            doff[a][l] = idx * p.dstr[a + 1];
        }
    }
    // One row of four taps per iteration (independent loads), rows and components in rolled loops with the weight /
    // offset tables in local memory: the routine must stay SMALL IN REGISTERS -- it is called from inside the voxel
    // loops of the table kernels, and every register it uses is one the caller has to spill around the call (a fully
    // unrolled form, 90 registers, cost edf_fast_real_kernel<3,3,double> 640 bytes of extra spills and half its speed).
#pragma unroll 1
    for (int h = 0; h < NAXIS; ++h) {
        const char* bh = p.disp + p.dstr[0] * h;
        double sum = 0.0;
        constexpr int NROW = (NAXIS == 3) ? 16 : 4;
#pragma unroll 1
        for (int r = 0; r < NROW; ++r) {
            const int i = (NAXIS == 3) ? (r >> 2) : 0, j = (NAXIS == 3) ? (r & 3) : r;
            const char* row = (NAXIS == 3) ? bh + doff[0][i] + doff[1][j] : bh + doff[0][j];
            const double wij = (NAXIS == 3) ? 0.0 : 0.0;
            (void)wij;
            double c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = (double)*(const TD*)(row + doff[NAXIS - 1][k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double t = c[k];
                if (NAXIS == 3) t = xmul(xmul(xmul(t, dw[0][i]), dw[1][j]), dw[2][k]);
                else            t = xmul(xmul(t, dw[0][j]), dw[1][k]);
                sum = xadd(sum, t);
            }
        }
        dd[h] = sum;
    }
}

template <int NAXIS>
__device__ __noinline__ void edf_displacement_exact_cold(const EdfParams& p, const int* o, double* dd)
{
    static_assert(NAXIS == 2 || NAXIS == 3, "fast kernels: 2-D and 3-D");
    if (p.ddtype == EDF_F64) {
        edf_displacement_exact_unrolled<NAXIS, double>(p, o, dd);
    } else if (p.ddtype == EDF_F32) {
        edf_displacement_exact_unrolled<NAXIS, float>(p, o, dd);
    } else {
        int64_t o64[NAXIS];
#pragma unroll
        for (int h = 0; h < NAXIS; ++h) o64[h] = o[h];
        edf_displacement_exact<NAXIS>(p, o64, dd);
    }
}

// un-mapped source coordinates in[h] of one voxel, with the exact-order re-evaluation
// of the voxels that sit next to a discontinuity
template <int NAXIS>
__device__ __forceinline__ void edf_fast_voxel_coords(const EdfParams& p, const EdfFastSmem<NAXIS>& s,
                                                      int buf, const int* o, int g, int m,
                                                      const double* wx, int sxrel, double* in)
{
    bool danger = false;
#pragma unroll
    for (int h = 0; h < NAXIS; ++h) {
        double d = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) d = fma(s.B[buf][h][g][m][sxrel + k], wx[k], d);
        if (p.has_affine) in[h] = edf_source_coordinate<NAXIS, int>(p, o, h, d);
        else              in[h] = xadd(xadd((double)o[h], p.ooff_d[h]), d);     // deform.c:778-781
        danger |= edf_near_half_integer(in[h]);
    }
    if (danger && s.nonzero) {
        double dd[NAXIS];
        edf_displacement_exact_cold<NAXIS>(p, o, dd);
#pragma unroll
        for (int h = 0; h < NAXIS; ++h) in[h] = edf_source_coordinate<NAXIS, int>(p, o, h, dd[h]);
    }
}

// per-axis tap offsets (element units) with the reference's mirror edge mapping
template <int ORDER>
__device__ __forceinline__ bool edf_fast_tap_offsets(int start, int len, int stride_e, int* off)
{
    const bool edge = start < 0 || start + ORDER >= len;
#pragma unroll
    for (int l = 0; l <= ORDER; ++l) {
        int idx = start + l;
        if (edge) idx = edf_mirror_index32(idx, len);
        off[l] = idx * stride_e;
    }
    return edge;
}

// Tile walker shared by all fast kernels: calls body(o, in) for every output voxel of the CTA's
// tile, `in` being the un-mapped source coordinates.
template <int NAXIS, typename Body>
__device__ __forceinline__ void edf_fast_walk_tile(const EdfParams& p, EdfFastSmem<NAXIS>& s, Body& body)
{
    constexpr int AX = NAXIS - 1, AY = NAXIS - 2;
    constexpr int NCHUNK = EdfFastGeom<NAXIS>::NCHUNK;
    constexpr int CHUNK_ROWS = EdfFastGeom<NAXIS>::CHUNK_ROWS;
    const int64_t x0 = (int64_t)blockIdx.x * EDF_FAST_TX;
    const int64_t y0 = (int64_t)blockIdx.y * EDF_FAST_RY;
    const int64_t z0 = (int64_t)blockIdx.z * EDF_FAST_G;
    edf_fast_tile_setup<NAXIS>(p, s, z0, y0, x0);

    const int tx = threadIdx.x & (EDF_FAST_TX - 1);
    const int g = threadIdx.x >> 6;
    const int64_t x = x0 + tx;
    const bool xok = x < p.odim[AX];
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[tx][k];
    const int sxrel = s.sx[tx] - s.sx[0];
    const int nchunk = (int)min((int64_t)NCHUNK, (p.odim[AY] - y0 + CHUNK_ROWS - 1) / CHUNK_ROWS);

    edf_fast_chunk_setup<NAXIS>(s, 0, 0);
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) edf_fast_chunk_setup<NAXIS>(s, c + 1, buf ^ 1);   // overlaps with the voxel loop
        if (xok) {
            for (int m = 0; m < EDF_FAST_M; ++m) {
                int o[NAXIS];
                o[AX] = (int)x;
                if (NAXIS == 3) {
                    o[0] = (int)z0 + g;
                    o[AY] = (int)y0 + c * CHUNK_ROWS + m;
                    if (o[0] >= (int)p.odim[0] || o[AY] >= (int)p.odim[AY]) continue;
                } else {
                    o[AY] = (int)y0 + c * CHUNK_ROWS + g * EDF_FAST_M + m;
                    if (o[AY] >= (int)p.odim[AY]) continue;
                }
                double in[NAXIS];
                edf_fast_voxel_coords<NAXIS>(p, s, buf, o, g, m, wx, sxrel, in);
                body(o, in);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// float32 kernel: forward gather (GRAD=false) or gradient scatter (GRAD=true)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float  edf_fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double edf_fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ void edf_bits_to(uint64_t bits, float* v) { *v = __uint_as_float((uint32_t)bits); }
__device__ __forceinline__ void edf_bits_to(uint64_t bits, double* v) { *v = __longlong_as_double((long long)bits); }

// all the work for one float32 / float64 input at one output voxel (general path: any mode, edges,
// strides, non-deformed "step" axes); weights and accumulation in the array's own precision
template <int NAXIS, int ORDER, bool GRAD, typename T>
__device__ __forceinline__ void edf_fast_real_one_input(const EdfParams& p, const EdfFastLaunch& L, int ii,
                                                       const int* o, const double* in)
{
    constexpr int NT = ORDER + 1;
    const EdfInputDesc& d = p.inp[ii];
    bool constant = false, edge = false;
    T w[NAXIS][NT];
    int off[NAXIS][NT];
#pragma unroll
    for (int h = 0; h < NAXIS; ++h) {
        int st = 0;
        T fr = (T)0;
        if (!constant && !edf_fast_finish(p, d.mode, ORDER, h, in[h], &st, &fr)) constant = true;
        if (!constant) {
            edge |= edf_fast_tap_offsets<ORDER>(st, (int)p.idim[h], L.istr_e[ii][h], off[h]);
            if (ORDER > 0) edf_bspline_weights_t<ORDER, T>(fr, w[h]);
        }
    }
    int64_t obase = 0;
#pragma unroll
    for (int h = 0; h < NAXIS; ++h) obase += (int64_t)o[h] * L.ostr_e[ii][h];
    // interior voxels on a unit-stride last axis: taps are base + i*sz + j*sy + k
    const bool dense = !edge && L.istr_e[ii][NAXIS - 1] == 1;
    const int sz_e = L.istr_e[ii][0], sy_e = (NAXIS == 3) ? L.istr_e[ii][1] : 0;

    const int64_t nsteps = d.nsteps;
    for (int64_t ss = 0; ss < nsteps; ++ss) {
        int64_t istep = 0, ostep = 0;
        if (d.nstep_rank == 1) {
            istep = d.in_step_str[0] * ss;
            ostep = d.out_step_str[0] * ss;
        } else if (d.nstep_rank > 1) {
            int64_t r = ss;
            for (int q = 0; q < d.nstep_rank; ++q) {
                const int64_t c = r % d.step_dim[q];
                r /= d.step_dim[q];
                istep += d.in_step_str[q] * c;
                ostep += d.out_step_str[q] * c;
            }
        }
        T* po = (T*)(d.out + ostep) + obase;
        if (!GRAD) {
            const T* __restrict__ pi = (const T*)(d.in + istep);
            T t;
            if (constant) {
                edf_bits_to(L.cval_bits[ii], &t);
            } else if (ORDER == 0) {
                int e = 0;
#pragma unroll
                for (int h = 0; h < NAXIS; ++h) e += off[h][0];
                t = __ldg(pi + e);
            } else if (dense) {
                int e0 = 0;
#pragma unroll
                for (int h = 0; h < NAXIS; ++h) e0 += off[h][0];
                const T* base = pi + e0;
                t = (T)0;
                if (NAXIS == 3) {
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const T* pz = base + (int64_t)i * sz_e;
                        T ti = (T)0;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const T* row = pz + (int64_t)j * sy_e;
                            T tj = (T)0;
#pragma unroll
                            for (int k = 0; k < NT; ++k) tj = edf_fma_t(__ldg(row + k), w[2][k], tj);
                            ti = edf_fma_t(tj, w[1][j], ti);
                        }
                        t = edf_fma_t(ti, w[0][i], t);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const T* row = base + (int64_t)j * sz_e;
                        T tj = (T)0;
#pragma unroll
                        for (int k = 0; k < NT; ++k) tj = edf_fma_t(__ldg(row + k), w[1][k], tj);
                        t = edf_fma_t(tj, w[0][j], t);
                    }
                }
            } else if (NAXIS == 3) {
                t = (T)0;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    T ti = (T)0;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const T* row = pi + (off[0][i] + off[1][j]);
                        T tj = (T)0;
#pragma unroll
                        for (int k = 0; k < NT; ++k) tj = edf_fma_t(__ldg(row + off[2][k]), w[2][k], tj);
                        ti = edf_fma_t(tj, w[1][j], ti);
                    }
                    t = edf_fma_t(ti, w[0][i], t);
                }
            } else {
                t = (T)0;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const T* row = pi + off[0][j];
                    T tj = (T)0;
#pragma unroll
                    for (int k = 0; k < NT; ++k) tj = edf_fma_t(__ldg(row + off[1][k]), w[1][k], tj);
                    t = edf_fma_t(tj, w[0][j], t);
                }
            }
            *po = t;
        } else if (!constant) {
            T* pi = (T*)(d.in + istep);
            const T gval = *po;
            if (ORDER == 0) {
                int e = 0;
#pragma unroll
                for (int h = 0; h < NAXIS; ++h) e += off[h][0];
                atomicAdd(pi + e, gval);
            } else if (NAXIS == 3) {
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const T gi = gval * w[0][i];
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const T gj = gi * w[1][j];
                        T* row = pi + (off[0][i] + off[1][j]);
#pragma unroll
                        for (int k = 0; k < NT; ++k) atomicAdd(row + off[2][k], gj * w[2][k]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const T gj = gval * w[0][j];
                    T* row = pi + off[0][j];
#pragma unroll
                    for (int k = 0; k < NT; ++k) atomicAdd(row + off[1][k], gj * w[1][k]);
                }
            }
        }
    }
}

template <int NAXIS, int ORDER, bool GRAD>
__device__ __forceinline__ void edf_fast_f32_one_input(const EdfParams& p, const EdfFastLaunch& L, int ii,
                                                       const int* o, const double* in)
{
    edf_fast_real_one_input<NAXIS, ORDER, GRAD, float>(p, L, ii, o, in);
}

template <int NAXIS, int ORDER, bool GRAD, typename T>
struct EdfFastRealBody {
    const EdfParams& p;
    const EdfFastLaunch& L;

    __device__ __forceinline__ void operator()(const int* o, const double* in) const
    {
        for (int ii = 0; ii < p.ninputs; ++ii) {
            if (!((L.input_mask >> ii) & 1u)) continue;
            edf_fast_real_one_input<NAXIS, ORDER, GRAD, T>(p, L, ii, o, in);
        }
    }
};

template <int NAXIS, int ORDER, bool GRAD, typename T>
__global__ void __launch_bounds__(EDF_FAST_THREADS, 2)
edf_fast_real_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L)
{
    __shared__ EdfFastSmem<NAXIS> s;
    EdfFastRealBody<NAXIS, ORDER, GRAD, T> body{p, L};
    edf_fast_walk_tile<NAXIS>(p, s, body);
}

// ---------------------------------------------------------------------------------------
// order-0 forward for any element type: the output is a bit copy of the selected input
// voxel (label volumes, BASELINE config 3's int32 input) or the converted cval
// ---------------------------------------------------------------------------------------
template <int NAXIS, typename T>
struct EdfFastCopyBody {
    const EdfParams& p;
    const EdfFastLaunch& L;

    __device__ __forceinline__ void operator()(const int* o, const double* in) const
    {
        for (int ii = 0; ii < p.ninputs; ++ii) {
            if (!((L.input_mask >> ii) & 1u)) continue;
            const EdfInputDesc& d = p.inp[ii];
            bool constant = false;
            int64_t e = 0;
#pragma unroll
            for (int h = 0; h < NAXIS; ++h) {
                int st = 0, off0;
                float fr;
                if (!constant && !edf_fast_finish(p, d.mode, 0, h, in[h], &st, &fr)) constant = true;
                if (!constant) {
                    edf_fast_tap_offsets<0>(st, (int)p.idim[h], L.istr_e[ii][h], &off0);
                    e += off0;
                }
            }
            int64_t obase = 0;
#pragma unroll
            for (int h = 0; h < NAXIS; ++h) obase += (int64_t)o[h] * L.ostr_e[ii][h];
            for (int64_t ss = 0; ss < d.nsteps; ++ss) {
                int64_t istep = 0, ostep = 0, r = ss;
                for (int q = 0; q < d.nstep_rank; ++q) {
                    const int64_t c = r % d.step_dim[q];
                    r /= d.step_dim[q];
                    istep += d.in_step_str[q] * c;
                    ostep += d.out_step_str[q] * c;
                }
                T v;
                if (constant) v = (T)L.cval_bits[ii];
                else          v = __ldg((const T*)(d.in + istep) + e);
                *((T*)(d.out + ostep) + obase) = v;
            }
        }
    }
};

template <int NAXIS, typename T>
__global__ void __launch_bounds__(EDF_FAST_THREADS, 2)
edf_fast_copy_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L)
{
    __shared__ EdfFastSmem<NAXIS> s;
    EdfFastCopyBody<NAXIS, T> body{p, L};
    edf_fast_walk_tile<NAXIS>(p, s, body);
}

// ---------------------------------------------------------------------------------------
// host-side selection
// ---------------------------------------------------------------------------------------
static thread_local cudaError_t g_fast_launch_error = cudaSuccess;

enum { EDF_CLASS_NONE = 0, EDF_CLASS_F32 = 1, EDF_CLASS_COPY = 2, EDF_CLASS_F64 = 3 };

static inline int edf_elem_size(int dt)
{
    switch (dt) {
    case EDF_BOOL: case EDF_U8: case EDF_I8: return 1;
    case EDF_U16: case EDF_I16: return 2;
    case EDF_U32: case EDF_I32: case EDF_F32: return 4;
    default: return 8;
    }
}

// class of input ii, or NONE when it has to go through the generic kernel
static int edf_fast_input_class(const EdfParams& p, int ii, EdfFastLaunch& L)
{
    const EdfInputDesc& d = p.inp[ii];
    if (d.in_dtype != d.out_dtype) return EDF_CLASS_NONE;
    const int es = edf_elem_size(d.in_dtype);
    int cls;
    if (d.in_dtype == EDF_F32) cls = EDF_CLASS_F32;
    else if (d.in_dtype == EDF_F64) cls = EDF_CLASS_F64;
    else if (d.order == 0 && !p.gradient && d.in_dtype != EDF_BOOL) cls = EDF_CLASS_COPY;
    else return EDF_CLASS_NONE;
    int64_t max_in = 0, max_out = 0;
    for (int h = 0; h < p.naxis; ++h) {
        if (d.istr[h] % es || d.ostr[h] % es) return EDF_CLASS_NONE;
        const int64_t se = d.istr[h] / es, so = d.ostr[h] / es;
        if (se < 0 || so < 0) return EDF_CLASS_NONE;
        max_in += se * (p.idim[h] - 1);
        max_out += so * (p.odim[h] - 1);
        if (se > 0x7fffffffLL || so > 0x7fffffffLL) return EDF_CLASS_NONE;
        L.istr_e[ii][h] = (int32_t)se;
        L.ostr_e[ii][h] = (int32_t)so;
    }
    if (max_in > 0x7fffffffLL || max_out > 0x7fffffffLL) return EDF_CLASS_NONE;
    if (((uintptr_t)d.in % es) || ((uintptr_t)d.out % es)) return EDF_CLASS_NONE;
    for (int q = 0; q < d.nstep_rank; ++q)
        if (d.in_step_str[q] % es || d.out_step_str[q] % es) return EDF_CLASS_NONE;
    uint64_t bits = 0;
    edf_store((char*)&bits, d.out_dtype, d.cval);       // reference conversion rule (deform.c:906-919)
    L.cval_bits[ii] = bits;
    return cls;
}

template <int NAXIS, bool GRAD, typename T>
static void edf_fast_launch_real(int order, dim3 grid, cudaStream_t st, const EdfParams& p,
                                 const EdfFastLaunch& L)
{
    switch (order) {
    case 0: edf_fast_real_kernel<NAXIS, 0, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 1: edf_fast_real_kernel<NAXIS, 1, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 2: edf_fast_real_kernel<NAXIS, 2, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 3: edf_fast_real_kernel<NAXIS, 3, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 4: edf_fast_real_kernel<NAXIS, 4, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    default: edf_fast_real_kernel<NAXIS, 5, GRAD, T><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    }
}

template <int NAXIS>
static void edf_fast_launch_copy(int es, dim3 grid, cudaStream_t st, const EdfParams& p,
                                 const EdfFastLaunch& L)
{
    switch (es) {
    case 1: edf_fast_copy_kernel<NAXIS, uint8_t><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 2: edf_fast_copy_kernel<NAXIS, uint16_t><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    case 4: edf_fast_copy_kernel<NAXIS, uint32_t><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    default: edf_fast_copy_kernel<NAXIS, unsigned long long><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L); break;
    }
}

// straight-line 3-D float32 kernels (edf_lean.cuh)
static bool edf_lean_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii);
static void edf_lean_launch(int order, int gradient, dim3 grid, cudaStream_t st, const EdfParams& p,
                            const EdfFastLaunch& L, int ii);
static int edf_lean_launch_gradwin(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& L, int ii);
static bool edf_gradwin_eligible(const EdfParams& p);
#ifndef EDF_SWIN_FWD_MIN_VOXELS
#define EDF_SWIN_FWD_MIN_VOXELS (3ull << 20)   // output voxels from which the staged-window forward gather is the default
#endif
static bool edf_swin_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii);
static bool edf_swin_fwd_env();
static int edf_swin_max_fwd_order();
static bool edf_swin_grad_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii, bool all_orders);
static int edf_swin_launch(int order, int gradient, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& L, int ii);
// polynomial-coordinate kernels (edf_poly.cuh)
static bool edf_poly_direct_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii);
static int edf_poly_direct_launch(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& L, int ii);
// staged-window kernels of round 2 (edf_tile.cuh): tensor-map TMA staging, polynomial coordinates
static bool edf_tile_fwd_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii);
static int edf_tile_launch_fwd(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& L, int ii);
static bool edf_tile_env()                              // EDF_TILE=1: the tensor-map TMA forward kernel instead of the bulk-copy one
{                                                       //   (A/B runs: 0.56 vs 0.47 ms at order 3, 0.34 vs 0.32 ms at order 2 on 256^3)
    static std::atomic<int> v{-1};
    int x = v.load(std::memory_order_relaxed);
    if (x < 0) { const char* e = getenv("EDF_TILE"); x = (e && *e && *e != '0') ? 1 : 0; v.store(x, std::memory_order_relaxed); }
    return x != 0;
}
// producer / consumer pipeline of staged windows (edf_pipe.cuh): forward gather, orders 0-3, 'constant' mode
static bool edf_pipe_fwd_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii);
static int edf_pipe_launch_fwd(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& L, int ii);

// Tries to run (part of) the problem on the specialised kernels.
//   *handled_mask receives the inputs that were processed (the caller runs the generic
//   kernel for the others); returns the number of kernels launched, or <0 on a launch error.
static int edf_fast_try_launch(const EdfParams& p, cudaStream_t st, const char** name,
                               uint32_t* handled_mask, uint32_t flags)
{
    const bool windows = !(flags & EDF_FLAG_NO_WINDOW);
    *handled_mask = 0;
    if (p.naxis != 2 && p.naxis != 3) return 0;
    const int AX = p.naxis - 1, AY = p.naxis - 2;
    for (int a = 0; a < p.naxis; ++a) {
        if (p.idim[a] < 2 || p.idim[a] > 0x3fffffff || p.odim[a] > 0x3fffffff) return 0;
        if (p.ncp[a] > 0x3fffffff) return 0;
    }
    if (p.ddtype != EDF_F64 && p.ddtype != EDF_F32) return 0;
    if (!edf_fast_ctrl_span_ok(p, AX, EDF_FAST_TX)) return 0;
    if (!edf_fast_ctrl_span_ok(p, AY, EDF_FAST_RY)) return 0;

    EdfFastLaunch L;
    memset(&L, 0, sizeof(L));
    int cls[EDF_MAX_INPUTS];
    for (int ii = 0; ii < p.ninputs; ++ii) cls[ii] = edf_fast_input_class(p, ii, L);

    dim3 grid;
    grid.x = (unsigned)((p.odim[AX] + EDF_FAST_TX - 1) / EDF_FAST_TX);
    grid.y = (unsigned)((p.odim[AY] + EDF_FAST_RY - 1) / EDF_FAST_RY);
    grid.z = (p.naxis == 3) ? (unsigned)((p.odim[0] + EDF_FAST_G - 1) / EDF_FAST_G) : 1u;
    if (grid.y > 65535u || grid.z > 65535u) return 0;

    int launches = 0;
    bool done[EDF_MAX_INPUTS] = {false};
    for (int ii = 0; ii < p.ninputs; ++ii) {
        if (done[ii] || cls[ii] == EDF_CLASS_NONE) continue;
        // orders 0 / 1 forward, 'constant' mode, 4-byte elements -- float32 volumes with or without a channel axis, and
        // 4-byte label volumes at order 0 (bit copy): the polynomial-coordinate direct kernel, channel loop inside
        if (!p.gradient && (cls[ii] == EDF_CLASS_F32 || cls[ii] == EDF_CLASS_COPY) && p.inp[ii].order <= 1 &&
            !(flags & (EDF_FLAG_STAGED_FWD | EDF_FLAG_NO_WINDOW)) &&
            (p.inp[ii].nstep_rank == 1 || cls[ii] == EDF_CLASS_COPY) && edf_poly_direct_eligible(p, L, ii)) {
            L.input_mask = 1u << ii;
            if (edf_poly_direct_launch(p.inp[ii].order, st, p, L, ii) == 0) {
                g_fast_launch_error = cudaGetLastError();
                if (g_fast_launch_error != cudaSuccess) return -1;
                *name = "poly3d_f32_direct";
                done[ii] = true;
                ++launches;
                *handled_mask |= 1u << ii;
                continue;
            }
        }
        if (cls[ii] == EDF_CLASS_F32 && edf_lean_eligible(p, L, ii)) {
            L.input_mask = 1u << ii;
            // small volumes: fewer rows per CTA until the grid fills the 148 SMs at least twice over
            unsigned ry = EDF_FAST_RY;
            while (ry > EDF_FAST_M &&
                   (uint64_t)grid.x * ((p.odim[AY] + ry - 1) / ry) * grid.z < 4ull * 148) ry >>= 1;
            L.rows_per_cta = ry;
            dim3 lgrid = grid;
            lgrid.y = (unsigned)((p.odim[AY] + ry - 1) / ry);
            // staged-window kernels (edf_swin.cuh): forward at orders 2 / 3, gradient at orders 0-3, unless the
            // caller hints a steep field (boxes that outgrow the window -> the round-1 kernels are faster there)
            int rcs = -2;
            const bool steep = (flags & EDF_FLAG_STEEP) != 0;
            // An affine map that magnifies by more than ~3.6x per axis sends > 48 output voxels to every input cell: the
            // 32-bit fixed-point cells of the gradient windows (headroom ~128 voxels' worth of mass) are not used then
            // (the staged-window kernel also checks the multiplicity per chunk on the device, for what the displacement adds)
            bool dense_affine = false;
            if (p.gradient && p.has_affine) {
                const double* A = p.affine;
                const double det = A[0] * (A[5] * A[10] - A[6] * A[9]) - A[1] * (A[4] * A[10] - A[6] * A[8]) + A[2] * (A[4] * A[9] - A[5] * A[8]);
                dense_affine = !(fabs(det) >= 1.0 / 48.0);
            }
            const int ord = p.inp[ii].order;
            bool poly = false, tile = false, pipe = false;
#ifdef EDF_WITH_PIPE
            if (!p.gradient && windows && !steep && !(flags & EDF_FLAG_STAGED_FWD) && edf_pipe_fwd_eligible(p, L, ii)) {
                rcs = edf_pipe_launch_fwd(ord, st, p, L, ii);
                pipe = rcs == 0;
            }
#endif
            if (rcs != -2) {
                // launched (or failed) above
            } else if (!p.gradient && ord <= 1 && !(flags & EDF_FLAG_STAGED_FWD) && edf_poly_direct_eligible(p, L, ii)) {
                // orders 0 / 1, 'constant' mode: polynomial coordinates, direct gather
                rcs = edf_poly_direct_launch(ord, st, p, L, ii);
                poly = rcs == 0;
            } else if (windows && !p.gradient) {
                // (and unless the volume is small: below ~two waves of CTAs the per-CTA prologue and the two barriers
                //  per chunk weigh more than the cheaper taps -- 128^3, order 3: 0.169 ms staged against 0.137 ms direct)
                const bool big = (uint64_t)p.odim[0] * (uint64_t)p.odim[1] * (uint64_t)p.odim[2] >= EDF_SWIN_FWD_MIN_VOXELS;
                const bool want = (flags & EDF_FLAG_STAGED_FWD) || edf_swin_fwd_env() ||
                                  (!steep && big && ord >= 2 && ord <= edf_swin_max_fwd_order());
                if (want && (edf_tile_env() || (flags & EDF_FLAG_STAGED_FWD)) && edf_tile_fwd_eligible(p, L, ii)) {
                    rcs = edf_tile_launch_fwd(ord, st, p, L, ii);
                    tile = rcs == 0;
                }
                if (rcs == -2 && want && edf_swin_eligible(p, L, ii)) rcs = edf_swin_launch(ord, 0, st, p, L, ii);
            } else if (windows && !dense_affine && p.gradient && !(flags & EDF_FLAG_FIXED_WINDOW) &&
                       ((flags & EDF_FLAG_STAGED_ALL) || !(steep && ord >= 2))) {
                if (edf_swin_grad_eligible(p, L, ii, (flags & EDF_FLAG_STAGED_ALL) != 0))
                    rcs = edf_swin_launch(ord, 1, st, p, L, ii);
            }
            if (rcs == -1) return -1;
            if (rcs == 0 && pipe) {
                *name = "pipe3d_f32";
            } else if (rcs == 0 && poly) {
                *name = "poly3d_f32_direct";
            } else if (rcs == 0 && tile) {
                *name = p.gradient ? "tile3d_f32_grad" : "tile3d_f32";
            } else if (rcs == 0) {
                *name = p.gradient ? "swin3d_f32_grad" : "swin3d_f32";
            } else if (windows && !dense_affine && p.gradient && edf_gradwin_eligible(p)) {
                const int rcw = edf_lean_launch_gradwin(p.inp[ii].order, st, p, L, ii);
                if (rcw < 0) return -1;
                *name = rcw == 2 ? "lean3d_f32_gradwin_tma" : "lean3d_f32_gradwin";
            } else {
                edf_lean_launch(p.inp[ii].order, p.gradient, lgrid, st, p, L, ii);
                *name = p.gradient ? "lean3d_f32_grad" : "lean3d_f32";
            }
            g_fast_launch_error = cudaGetLastError();
            if (g_fast_launch_error != cudaSuccess) return -1;
            done[ii] = true;
            ++launches;
            *handled_mask |= 1u << ii;
            continue;
        }
        // group every later input with the same class / order / element size
        uint32_t mask = 0;
        const int es = edf_elem_size(p.inp[ii].in_dtype);
        for (int jj = ii; jj < p.ninputs; ++jj) {
            if (done[jj] || cls[jj] != cls[ii]) continue;
            if (p.inp[jj].order != p.inp[ii].order) continue;
            if (edf_elem_size(p.inp[jj].in_dtype) != es) continue;
            mask |= 1u << jj;
            done[jj] = true;
        }
        L.input_mask = mask;
        if (cls[ii] == EDF_CLASS_F32) {
            const int order = p.inp[ii].order;
            if (p.naxis == 3) {
                if (p.gradient) edf_fast_launch_real<3, true, float>(order, grid, st, p, L);
                else            edf_fast_launch_real<3, false, float>(order, grid, st, p, L);
            } else {
                if (p.gradient) edf_fast_launch_real<2, true, float>(order, grid, st, p, L);
                else            edf_fast_launch_real<2, false, float>(order, grid, st, p, L);
            }
            *name = p.gradient ? "fast_f32_grad" : "fast_f32";
        } else if (cls[ii] == EDF_CLASS_F64) {
            const int order = p.inp[ii].order;
            if (p.naxis == 3) {
                if (p.gradient) edf_fast_launch_real<3, true, double>(order, grid, st, p, L);
                else            edf_fast_launch_real<3, false, double>(order, grid, st, p, L);
            } else {
                if (p.gradient) edf_fast_launch_real<2, true, double>(order, grid, st, p, L);
                else            edf_fast_launch_real<2, false, double>(order, grid, st, p, L);
            }
            *name = p.gradient ? "fast_f64_grad" : "fast_f64";
        } else {
            if (p.naxis == 3) edf_fast_launch_copy<3>(es, grid, st, p, L);
            else              edf_fast_launch_copy<2>(es, grid, st, p, L);
            *name = "fast_copy";
        }
        g_fast_launch_error = cudaGetLastError();
        if (g_fast_launch_error != cudaSuccess) return -1;
        ++launches;
        *handled_mask |= mask;
    }
    return launches;
}
