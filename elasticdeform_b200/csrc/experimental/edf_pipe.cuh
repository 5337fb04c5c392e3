// edf_pipe.cuh -- K1p: forward gather through a PRODUCER / CONSUMER pipeline of staged windows (round 2).
//
// The staged-window kernels of round 1 (edf_swin.cuh) and the first tensor-map kernel (edf_tile.cuh) run
// coordinates -> box -> copy -> wait -> gather in sequence per chunk with CTA barriers in between; 70 % of their
// time is not the gather.  This kernel decouples the two halves:
//
//   * one CTA per SM, persistent over tiles of 32 (x) x ry (y) x 16 (z) output voxels; 16 CONSUMER warps
//     (warp = z-slab, lane = x, a thread walks along y) and one PRODUCER warp;
//   * the producer bounds the source box of the next chunk (32 x up to 8 rows x 16 slabs) WITHOUT the voxels'
//     own coordinates: it evaluates the displacement exactly (deform.c:650-758, the cold reference-order routine)
//     at a 4 x 2 x 4 lattice of the chunk's voxels, one per lane; the field is a cubic spline with 64-voxel knots,
//     so the lattice extrema plus a margin bound the chunk.  It picks the largest row count (8/4/2/1) whose box
//     fits a stage, publishes the box and fills the stage with tensor-map TMA boxes of 64 x 4 x 1 floats
//     (cp.async.bulk.tensor.3d, SASS UTMALDG; out-of-volume cells arrive as zeros) that complete on the stage's
//     `full` transaction mbarrier.  It runs ahead of the consumers by one stage (two stages of 104 KB);
//   * consumers wait on `full`, gather, and release the stage through its `empty` mbarrier: no CTA-wide barrier
//     in the steady state.  A voxel whose taps are not inside the published box (the margin was too small) reads
//     its taps from global memory instead -- the box is a performance hint, never a correctness assumption;
//   * coordinates: the displacement of a thread's column is one cubic polynomial per component in the row's
//     fractional control position u (edf_poly.cuh).  Here the row's own index, the crop offset and the affine map
//     are folded into the polynomial as well, and so is the constant 1.5 * 2^29: the last FMA of the Horner
//     form then rounds the SOURCE COORDINATE to a multiple of 2^-23, and floor / fractional offset / range test
//     are integer operations on the two words of the result (no conversion instructions, no fp64 compares):
//         T = c + 1.5*2^29  ->  floor(c) = bits [23,55) of T - const,   frac(c) = (lo & 0x7fffff) * 2^-23;
//   * border chunks ('constant' mode still mirrors the taps of in-range voxels that cross the border,
//     deform.c:791-813): only the cells -1 and len can be read; the consumers copy them from cells 1 and len-2
//     in three sweeps (x, y, z) before gathering.
//
// Exactness: as in the other float32 kernels every discrete decision equals the reference's because voxels
// within 2^-19 of a threshold (integer coordinates for odd orders, integer and half-integer ones for even
// orders) are re-evaluated in the reference order (edf_poly_slow_voxel); the quantisation to 2^-23 moves a
// coordinate by < 1.2e-7, well inside that zone.  Interpolation weights are float32 as before.
#pragma once
#include "../edf_tile.cuh"

#define EDF_PP_TX 32               // x positions per warp / tile
#define EDF_PP_G 8                 // z-slabs per tile
#define EDF_PP_NW 16               // consumer warps: warp w owns slab w % 8 and the rows of half w / 8 of a chunk
#define EDF_PP_MR 8                // rows per chunk (at most)
#define EDF_PP_CONSUMERS (EDF_PP_TX * EDF_PP_NW)
#define EDF_PP_THREADS (EDF_PP_CONSUMERS + 32)
#define EDF_PP_PITCH 64            // floats between window rows (= the TMA box width)
#define EDF_PP_BY 4                // rows per TMA box
#ifndef EDF_PP_STAGE_ROWS
#define EDF_PP_STAGE_ROWS 416      // stage capacity in rows of 256 bytes (104 KB)
#endif
#define EDF_PP_STAGES 2
#define EDF_PP_RY 64               // rows per tile (table capacity)
#define EDF_PP_NC 8
#ifndef EDF_PP_U
#define EDF_PP_U 1                 // rows per iteration of a consumer thread (2: no faster on B200, measured)
#endif

#ifdef EDF_PIPE_STATS
#define EDF_PS_DECL long long ps_t = clock64(); long long ps_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define EDF_PS_MARK(k) { const long long t_ = clock64(); ps_acc[k] += t_ - ps_t; ps_t = t_; }
#define EDF_PS_FLUSH if ((threadIdx.x & 31) == 0) { for (int k_ = 0; k_ < 8; ++k_) if (ps_acc[k_]) atomicAdd(&g_tile_prof[k_], (unsigned long long)ps_acc[k_]); }
#else
#define EDF_PS_DECL
#define EDF_PS_MARK(k)
#define EDF_PS_FLUSH
#endif

static_assert(EDF_PP_BY == EDF_TL_BY && EDF_PP_PITCH == EDF_TL_PITCH, "tensor map of edf_tile_make_map");

struct EdfPipeBox {
    int wz0, wy0, wx0;             // first cell of the box
    int nz, ny, nyal, nx;          // planes, rows (used / allocated), columns (used); nz == 0: nothing staged
    int nb;                        // rows of the chunk
    int border;                    // bit 0 / 1 / 2: the box holds cell -1 or len along x / y / z
    int pad_[3];
};

struct EdfPipeSmem {
    double u[EDF_PP_RY];           // fractional control position of each row of the tile
    double wz[EDF_PP_G][4];
    double wx[EDF_PP_TX][4];
    double T[EDF_PP_NW][3][4][EDF_PP_NC];
    int    jy[EDF_PP_RY];
    int    sz[EDF_PP_G];
    int    sx[EDF_PP_TX];
    double rrat;                   // (I_y - 1) / (P_y - 1): rows per control interval
    unsigned rare[EDF_PP_G][EDF_PP_RY];   // per slab and row: lanes whose voxel goes to the single-voxel routine at the end of the tile
    EdfPipeBox box[EDF_PP_STAGES];
    unsigned long long full[EDF_PP_STAGES], empty[EDF_PP_STAGES];
};

__device__ __forceinline__ void edf_mbar_arrive(unsigned long long* mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
__device__ __forceinline__ void edf_consumer_sync()
{
    asm volatile("bar.sync 1, %0;" :: "n"(EDF_PP_CONSUMERS) : "memory");
}

// consumer prologue of a tile: control tables (ends with a consumer barrier)
__device__ __forceinline__ void edf_pipe_tables(const EdfParams& p, EdfPipeSmem& s, int z0, int y0, int x0, int nrow, int tid)
{
    if (tid < EDF_PP_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_PP_TX + EDF_PP_RY) {
        const int t = tid - EDF_PP_TX;
        if (t < nrow) {
            const double cp = edf_control_pos(p, 1, (int64_t)(y0 + t));
            const double fl = floor(cp);
            s.u[t] = xsub(cp, fl);
            s.jy[t] = (int)fl - 1;
        }
    } else if (tid < EDF_PP_TX + EDF_PP_RY + EDF_PP_G) {
        const int t = tid - EDF_PP_TX - EDF_PP_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    } else if (tid == EDF_PP_TX + EDF_PP_RY + EDF_PP_G) {
        s.rrat = xdiv(p.idim_m1[1], (double)(p.ncp[1] - 1));
    }
    edf_consumer_sync();
}

// Polynomial of this thread's column for the control interval whose window starts at control row j0, with the row
// index, crop offset, affine map and the fixed-point constant folded in (see the header).  Warp-collective.
// out[h*4 + k]: coefficient of u^k of the source coordinate along axis h (+ 0.5 for even orders) + 1.5*2^29.
template <int ORDER>
__device__ __forceinline__ bool edf_pipe_poly(const EdfParams& p, EdfPipeSmem& s, int g, int warp, int lane, int j0, int z, int x, double* out)
{
    double a[12];
    const bool nz = edf_poly_build(p, s, g, lane, j0, a, warp);
    edf_poly_fold<ORDER>(p, a, s.rrat, j0, z, x, out);
    return nz | (p.has_affine != 0);
}

// in-range voxel whose taps are not inside the staged box: straight from global memory (mirror map of edge taps)
template <int ORDER>
__device__ __noinline__ float edf_pipe_direct_voxel(const float* __restrict__ pin, int stz, int sty, int stx, float ez, float ey, float ex,
                                                    int lenz, int leny, int lenx, int isz, int isy)
{
    const float h = (ORDER & 1) ? 0.5f : 0.0f;
    return edf_swin_direct_gather<ORDER>(pin, stz, sty, stx, ez + h, ey + h, ex + h, lenz, leny, lenx, isz, isy);
}

// cells -1 and len of a border box from cells 1 and len-2 (deform.c:796-810), one sweep per axis.  All 512 consumers.
__device__ __forceinline__ void edf_pipe_patch(float* win, const EdfPipeBox& b, int lenz, int leny, int lenx, int tid)
{
    const int rows = b.nz * b.nyal;
    if (b.border & 1) {
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int c = (side ? lenx : -1) - b.wx0, cm = (side ? lenx - 2 : 1) - b.wx0;
            if ((unsigned)c < (unsigned)EDF_PP_PITCH && (unsigned)cm < (unsigned)EDF_PP_PITCH)
                for (int r = tid; r < rows; r += EDF_PP_CONSUMERS) win[r * EDF_PP_PITCH + c] = win[r * EDF_PP_PITCH + cm];
        }
        edf_consumer_sync();
    }
    if (b.border & 2) {
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int yr = (side ? leny : -1) - b.wy0, ym = (side ? leny - 2 : 1) - b.wy0;
            if ((unsigned)yr < (unsigned)b.nyal && (unsigned)ym < (unsigned)b.nyal)
                for (int e = tid; e < b.nz * EDF_PP_PITCH; e += EDF_PP_CONSUMERS) {
                    const int zr = e >> 6, c = e & 63;
                    win[(zr * b.nyal + yr) * EDF_PP_PITCH + c] = win[(zr * b.nyal + ym) * EDF_PP_PITCH + c];
                }
        }
        edf_consumer_sync();
    }
    if (b.border & 4) {
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const int zr = (side ? lenz : -1) - b.wz0, zm = (side ? lenz - 2 : 1) - b.wz0;
            if ((unsigned)zr < (unsigned)b.nz && (unsigned)zm < (unsigned)b.nz)
                for (int e = tid; e < b.nyal * EDF_PP_PITCH; e += EDF_PP_CONSUMERS)
                    win[zr * b.nyal * EDF_PP_PITCH + e] = win[zm * b.nyal * EDF_PP_PITCH + e];
        }
        edf_consumer_sync();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes before the next TMA fill of these cells
}

// Producer's lattice evaluation.  It only bounds a chunk's box -- never a result -- so it need not follow the
// reference's operation order: a lane keeps the displacement of its lattice column (z, x) as a cubic in the row's
// fractional control position (as the consumers do for their columns) and rebuilds it when the row enters the next
// control interval.  Uniform cubic B-spline weights, mirrored control indices as deform.c:664-686.
__device__ __forceinline__ void edf_pipe_axis_taps(const EdfParams& p, int a, double cp, int64_t stride, double* w, int64_t* off)
{
    const double fl = floor(cp);
    const double u = cp - fl, v = 1.0 - u;
    const int j = (int)fl - 1;
    w[0] = v * v * v * (1.0 / 6.0);
    w[3] = u * u * u * (1.0 / 6.0);
    w[1] = (u * u * (u - 2.0) * 3.0 + 4.0) * (1.0 / 6.0);
    w[2] = (v * v * (v - 2.0) * 3.0 + 4.0) * (1.0 / 6.0);
#pragma unroll
    for (int l = 0; l < 4; ++l) off[l] = (int64_t)edf_mirror_index32(j + l, (int)p.ncp[a]) * stride;
}
__device__ __noinline__ void edf_pipe_column_poly(const EdfParams& p, const double* wz, const double* wx, const int64_t* oz, const int64_t* ox,
                                                  int j0, double* pa /*[12]*/)
{
    int64_t oy[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) oy[jj] = (int64_t)edf_mirror_index32(j0 + jj, (int)p.ncp[1]) * p.dstr[2];
    const bool f64 = p.ddtype == EDF_F64;
#pragma unroll 1
    for (int h = 0; h < 3; ++h) {
        const char* bh = p.disp + p.dstr[0] * h;
        double E[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            double cf[16];
            if (f64) {
#pragma unroll
                for (int q = 0; q < 16; ++q) cf[q] = *(const double*)(bh + oz[q >> 2] + oy[jj] + ox[q & 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) cf[q] = (double)*(const float*)(bh + oz[q >> 2] + oy[jj] + ox[q & 3]);
            }
            double e = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) t = fma(cf[i * 4 + k], wx[k], t);
                e = fma(t, wz[i], e);
            }
            E[jj] = e;
        }
        pa[h * 4 + 0] = (E[0] + 4.0 * E[1] + E[2]) * (1.0 / 6.0);
        pa[h * 4 + 1] = (E[2] - E[0]) * 0.5;
        pa[h * 4 + 2] = (E[0] - 2.0 * E[1] + E[2]) * 0.5;
        pa[h * 4 + 3] = ((E[3] - E[0]) + 3.0 * (E[1] - E[2])) * (1.0 / 6.0);
    }
}

// one axis of the producer's box: window range of the lattice's floor coordinates, clamped to the cells an
// in-range voxel can read (-1 .. len for orders 2 / 3, 0 .. len-1 below)
template <int ORDER>
__device__ __forceinline__ void edf_pipe_box_axis(double cmin, double cmax, int len, int& lo, int& hi, bool& brd)
{
    // floor of (c [+ 0.5 for even orders]); half a cell of margin on either side for the curvature between lattice points
    const double hadj = (ORDER & 1) ? 0.0 : 0.5;
    const double a = fmin(fmax(cmin + hadj - 0.5, -4.0), (double)len + 3.0);     // NaN -> -4
    const double b = fmin(fmax(cmax + hadj + 0.5, -4.0), (double)len + 3.0);
    lo = (int)floor(a) - ORDER / 2;
    hi = (int)floor(b) - ORDER / 2 + ORDER;
    const int cl = (ORDER >= 2) ? -1 : 0, ch = (ORDER >= 2) ? len : len - 1;
    lo = max(lo, cl);
    hi = min(hi, ch);
    if (lo <= hi) {
        if (lo < 0) hi = max(hi, 1);                               // the mirror sources of the border cells
        if (hi >= len) lo = min(lo, len - 2);
        brd = (lo < 0) | (hi >= len);
    } else {
        brd = false;
    }
}

template <int ORDER>
__global__ void __launch_bounds__(EDF_PP_THREADS, 1)
edf_pipe3d_fwd_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L,
                      const __grid_constant__ CUtensorMap tmap, const int ii)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    EdfPipeSmem& s = *reinterpret_cast<EdfPipeSmem*>(smem_raw);
    constexpr int SMEM_HDR = (int)((sizeof(EdfPipeSmem) + 1023) & ~(size_t)1023);
    constexpr int STAGE_FLOATS = EDF_PP_STAGE_ROWS * EDF_PP_PITCH;
    float* const win0 = reinterpret_cast<float*>(smem_raw + SMEM_HDR);
    constexpr int NT = ORDER + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < EDF_PP_STAGES; ++q) {
            edf_mbar_init(&s.full[q], 1);
            edf_mbar_init(&s.empty[q], EDF_PP_NW);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const unsigned gx = L.sched.gx, gy = L.sched.gy[0];
    const int ry = (int)L.rows_per_cta;
    const unsigned ntiles = L.sched.z_begin[0];                    // total number of tiles
    unsigned cc = 0;                                               // chunk counter (same sequence on both sides)

    if (warp == EDF_PP_NW) {
        // ============================== producer ==============================
        const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win0);
        const int xoff = (lane & 3) == 0 ? 0 : (lane & 3) == 1 ? 10 : (lane & 3) == 2 ? 21 : 31;
        const int zq = (lane >> 2) & 3;
        const int zoff = zq == 0 ? 0 : zq == 1 ? 2 : zq == 2 ? 5 : 7;
        double sc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) sc[a] = (double)(p.ncp[a] - 1) / p.idim_m1[a];
        EDF_PS_DECL
        for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int x0 = (int)(t % gx) * EDF_PP_TX, y0 = (int)((t / gx) % gy) * ry, z0 = (int)(t / (gx * gy)) * EDF_PP_G;
            const int nrow = min(ry, ody - y0);
            int o[3];
            o[0] = min(z0 + zoff, odz - 1);
            o[2] = min(x0 + xoff, odx - 1);
            double cwz[4], cwx[4], pa[12];
            int64_t coz[4], cox[4];
            edf_pipe_axis_taps(p, 0, ((double)o[0] + p.ooff_d[0]) * sc[0], p.dstr[1], cwz, coz);
            edf_pipe_axis_taps(p, 2, ((double)o[2] + p.ooff_d[2]) * sc[2], p.dstr[3], cwx, cox);
            int pj = INT_MIN;
            int m = 0;
            while (m < nrow) {
                int nb = min(EDF_PP_MR, nrow - m);
                EdfPipeBox b;
                for (;;) {
                    o[1] = y0 + m + ((lane >> 4) ? nb - 1 : 0);
                    double dd[3];
                    {
                        const double cp = ((double)o[1] + p.ooff_d[1]) * sc[1];
                        const double fl = floor(cp);
                        const double u = cp - fl;
                        const int j0 = (int)fl - 1;
                        if (j0 != pj) {
                            edf_pipe_column_poly(p, cwz, cwx, coz, cox, j0, pa);
                            pj = j0;
                        }
#pragma unroll
                        for (int h = 0; h < 3; ++h) dd[h] = fma(fma(fma(pa[h * 4 + 3], u, pa[h * 4 + 2]), u, pa[h * 4 + 1]), u, pa[h * 4 + 0]);
                    }
                    int lo[3], hi[3];
                    bool brd[3];
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        const double c = edf_source_coordinate<3, int>(p, o, h, dd[h]);
                        // warp min / max of a double through its order-preserving integer image would need 64-bit
                        // reductions; the clamped coordinate fits a float with room to spare for a box bound
                        float cf = (float)fmin(fmax(c, -1.0e6), 1.0e6);
                        if (!(cf == cf)) cf = -1.0e6f;
                        float mn = cf, mx = cf;
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) {
                            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
                            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
                        }
                        // float rounding of the bound: widen by one ulp-scale step (1e-3 of a cell at 1e4)
                        const double wd = 1.0e-6 * fmax(fabs((double)mn), fabs((double)mx)) + 1.0e-6;
                        edf_pipe_box_axis<ORDER>((double)mn - wd, (double)mx + wd, h == 0 ? lenz : h == 1 ? leny : lenx, lo[h], hi[h], brd[h]);
                    }
                    b.nb = nb;
                    b.wz0 = lo[0]; b.wy0 = lo[1]; b.wx0 = lo[2] & ~3;
                    b.nz = hi[0] - lo[0] + 1;
                    b.ny = hi[1] - lo[1] + 1;
                    b.nyal = (b.ny + EDF_PP_BY - 1) & ~(EDF_PP_BY - 1);
                    b.nx = hi[2] - b.wx0 + 1;
                    b.border = (brd[2] ? 1 : 0) | (brd[1] ? 2 : 0) | (brd[0] ? 4 : 0);
                    const bool empty = (b.nz <= 0) | (b.ny <= 0) | (hi[2] < lo[2]);
                    if (empty) { b.nz = 0; b.ny = 0; b.nyal = 0; b.nx = 0; b.border = 0; break; }
                    const bool fit = (b.nx <= EDF_PP_PITCH) & (b.nz * b.nyal <= EDF_PP_STAGE_ROWS);
                    if (fit) break;
                    if (nb == 1) { b.nz = 0; b.ny = 0; b.nyal = 0; b.nx = 0; b.border = 0; break; }   // steep: consumers gather directly
                    nb = nb > 4 ? 4 : nb >> 1;
                }
                const unsigned st = cc & 1u;
                EDF_PS_MARK(4)
                if (cc >= EDF_PP_STAGES) edf_mbar_wait(&s.empty[st], ((cc >> 1) - 1u) & 1u);
                EDF_PS_MARK(5)
                if (lane == 0) s.box[st] = b;
                __syncwarp();
                const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s.full[st]);
                if (b.nz > 0) {
                    const int groups = b.nyal / EDF_PP_BY;
                    if (lane == 0) {
                        const unsigned bytes = (unsigned)(b.nz * b.nyal) * (EDF_PP_PITCH * 4u);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(bytes) : "memory");
                    }
                    __syncwarp();
                    const uint32_t dst0 = win_s + st * (STAGE_FLOATS * 4u);
                    for (int zr = lane; zr < b.nz; zr += 32) {
                        uint32_t dst = dst0 + (uint32_t)(zr * b.nyal) * (EDF_PP_PITCH * 4u);
                        for (int q = 0; q < groups; ++q, dst += EDF_PP_PITCH * EDF_PP_BY * 4u)
                            edf_tma_box3d(dst, &tmap, b.wx0, b.wy0 + q * EDF_PP_BY, b.wz0 + zr, mb);
                    }
                } else if (lane == 0) {
                    edf_mbar_arrive(&s.full[st]);
                }
                EDF_PS_MARK(6)
#ifdef EDF_PIPE_STATS
                if (lane == 0) {
                    atomicAdd(&g_tile_prof[10], 1ull);
                    atomicAdd(&g_tile_prof[11], (unsigned long long)nb);
                    if (b.nz == 0) atomicAdd(&g_tile_prof[9], 1ull);
                    atomicAdd(&g_tile_prof[12], (unsigned long long)(b.nz * b.nyal));
                }
#endif
                m += nb;
                ++cc;
            }
        }
        EDF_PS_FLUSH
        return;
    }

    // ============================== consumers ==============================
    const int g = warp & (EDF_PP_G - 1), half = warp >> 3;
    const uint32_t win_sa = (uint32_t)__cvta_generic_to_shared(win0);
    const EdfInputDesc& d = p.inp[ii];
    const float* __restrict__ pin = (const float*)d.in;
    float* __restrict__ pout = (float*)d.out;
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osz = L.ostr_e[ii][0], osy = L.ostr_e[ii][1], osx = L.ostr_e[ii][2];
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    // strict in-range tests on the floor coordinate (see the header); even orders test floor(2c + 1)
    const unsigned rngz = (ORDER & 1) ? (unsigned)(lenz - 2) : (unsigned)(2 * lenz - 3);
    const unsigned rngy = (ORDER & 1) ? (unsigned)(leny - 2) : (unsigned)(2 * leny - 3);
    const unsigned rngx = (ORDER & 1) ? (unsigned)(lenx - 2) : (unsigned)(2 * lenx - 3);

    EDF_PS_DECL
#ifdef EDF_PIPE_STATS
    const long long ps_t0 = clock64();
#endif
    for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int x0 = (int)(t % gx) * EDF_PP_TX, y0 = (int)((t / gx) % gy) * ry, z0 = (int)(t / (gx * gy)) * EDF_PP_G;
        const int nrow = min(ry, ody - y0);
        edf_consumer_sync();                                       // every warp is done with the previous tile's tables
        edf_pipe_tables(p, s, z0, y0, x0, nrow, tid);
        const int x = x0 + lane, z = z0 + g;
        const bool tok = (x < odx) & (z < odz);
        const int xc = min(x, odx - 1), zc = min(z, odz - 1);
        const int obase = zc * osz + xc * osx;                     // element offsets fit 32 bits (host-checked)
        double a[12];
        int jcur = INT_MIN;
        bool gate = false;
        int m = 0;
        while (m < nrow) {
            const unsigned st = cc & 1u;
            EDF_PS_MARK(2)
            edf_mbar_wait(&s.full[st], (cc >> 1) & 1u);
            EDF_PS_MARK(0)
            const EdfPipeBox b = s.box[st];
            float* const win = win0 + st * STAGE_FLOATS;
            if (b.border) edf_pipe_patch(win, b, lenz, leny, lenx, tid);
            EDF_PS_MARK(7)
            const int slab = b.nyal * EDF_PP_PITCH;
            const int limz = b.nz - 1 - ORDER, limy = b.ny - 1 - ORDER, limx = b.nx - 1 - ORDER;
            const bool usable = (b.nz > 0) & (limz >= 0) & (limy >= 0) & (limx >= 0);
            const int nh = (b.nb + 1) >> 1;                        // rows of this warp: its half of the chunk
            const int wend = min(b.nb, (half + 1) * nh);
            int r = half * nh;
            while (r < wend) {
                // rows [r, rend): one control interval of the y axis (warp-uniform); the polynomial is rebuilt between
                // the segments, outside the row loop
                const int jr = s.jy[m + r];
                int rend = wend;
                if (s.jy[m + wend - 1] != jr) {
                    rend = r + 1;
                    while (s.jy[m + rend] == jr) ++rend;
                }
                if (jr != jcur) {
                    EDF_PS_MARK(1)
                    gate = edf_pipe_poly<ORDER>(p, s, g, warp, lane, jr, zc, xc, a);
                    jcur = jr;
                    EDF_PS_MARK(2)
                }
                unsigned dmask = 0;                                // rows of this segment whose voxel gathers from global memory
                const int rseg = r;
#pragma unroll 1
                for (; r < rend; r += EDF_PP_U) {
                    // EDF_PP_U rows per iteration: coordinates and classification of all of them first, then ONE straight-line
                    // block with the taps of all of them, so that the loads of independent voxels overlap (one voxel at a
                    // time, ptxas keeps three loads in flight and the warp stalls on every FMA)
                    int off[EDF_PP_U];
                    float ez[EDF_PP_U], ey[EDF_PP_U], ex[EDF_PP_U];
                    bool gat[EDF_PP_U], wr[EDF_PP_U];
                    bool anyg = false;
#pragma unroll
                    for (int q = 0; q < EDF_PP_U; ++q) {
                        const bool live = r + q < rend;
                        const int mm = m + min(r + q, rend - 1);
                        EdfPipeVoxel v;
                        edf_pipe_coords<ORDER>(a, s.u[mm], gate, lenz, leny, lenx, rngz, rngy, rngx, v);
                        const int rz = v.stz - b.wz0, ryw = v.sty - b.wy0, rx = v.stx - b.wx0;
                        const bool cont = usable & ((unsigned)rz <= (unsigned)limz) & ((unsigned)ryw <= (unsigned)limy) & ((unsigned)rx <= (unsigned)limx);
                        const bool act = tok & live & v.inr & !v.slow;
                        // voxels next to a threshold are left to the single-voxel routine at the end of the tile (no calls
                        // inside this loop); in-range voxels whose taps are not all inside the staged box follow after the segment
                        const bool rare = tok & live & v.slow;
                        const unsigned rmask = __ballot_sync(0xffffffffu, rare);
                        if ((lane == 0) & live) s.rare[g][mm] = rmask;
                        if (act & !cont) dmask |= 1u << (r + q - rseg);
                        gat[q] = act & cont;
                        wr[q] = tok & live & !rare & (cont | !act);
                        off[q] = gat[q] ? (rz * b.nyal + ryw) * EDF_PP_PITCH + rx : 0;
                        ez[q] = v.ez; ey[q] = v.ey; ex[q] = v.ex;
                        anyg |= gat[q];
                    }
                    float res[EDF_PP_U];
#pragma unroll
                    for (int q = 0; q < EDF_PP_U; ++q) res[q] = cvalf;
                    if (__any_sync(0xffffffffu, anyg)) {
                        if (ORDER == 0) {
#pragma unroll
                            for (int q = 0; q < EDF_PP_U; ++q) {
                                const float t = win[off[q]];
                                if (gat[q]) res[q] = t;
                            }
                        } else {
                            float wzf[EDF_PP_U][NT], wyf[EDF_PP_U][NT], wxf[EDF_PP_U][NT], acc[EDF_PP_U];
#pragma unroll
                            for (int q = 0; q < EDF_PP_U; ++q) {
                                edf_pipe_weights<ORDER>(ez[q], wzf[q]);
                                edf_pipe_weights<ORDER>(ey[q], wyf[q]);
                                edf_pipe_weights<ORDER>(ex[q], wxf[q]);
                                acc[q] = 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                // one z-plane of taps of every voxel of the iteration: all loads first, a warp-level fence, then
                                // the FMAs (ptxas does not move shared-memory loads across the fence)
                                float tv[EDF_PP_U][NT * NT];
#pragma unroll
                                for (int q = 0; q < EDF_PP_U; ++q) {
                                    const float* qi = win + off[q] + i * slab;
#pragma unroll
                                    for (int j = 0; j < NT; ++j)
#pragma unroll
                                        for (int k = 0; k < NT; ++k) tv[q][j * NT + k] = qi[j * EDF_PP_PITCH + k];
                                }
                                __syncwarp();
#pragma unroll
                                for (int q = 0; q < EDF_PP_U; ++q) {
                                    float ti = 0.f;
#pragma unroll
                                    for (int j = 0; j < NT; ++j) {
                                        float tj = tv[q][j * NT] * wxf[q][0];
#pragma unroll
                                        for (int k = 1; k < NT; ++k) tj = fmaf(tv[q][j * NT + k], wxf[q][k], tj);
                                        ti = (j == 0) ? tj * wyf[q][0] : fmaf(tj, wyf[q][j], ti);
                                    }
                                    acc[q] = (i == 0) ? ti * wzf[q][0] : fmaf(ti, wzf[q][i], acc[q]);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < EDF_PP_U; ++q)
                                if (gat[q]) res[q] = acc[q];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < EDF_PP_U; ++q)
                        if (wr[q]) pout[obase + (y0 + m + r + q) * osy] = res[q];
                }
                r = rend;
                if (dmask) {
                    // the box missed these voxels (margin too small, or a chunk that does not fit a stage at all): same
                    // coordinates, taps straight from global memory with the mirror map of edge taps
#pragma unroll 1
                    for (int q = 0; q < EDF_PP_MR; ++q) {
                        if (!((dmask >> q) & 1u)) continue;
                        EdfPipeVoxel v;
                        edf_pipe_coords<ORDER>(a, s.u[m + rseg + q], gate, lenz, leny, lenx, rngz, rngy, rngx, v);
                        const float hh = (ORDER & 1) ? 0.5f : 0.0f;
                        float val;
                        if (ORDER == 0) val = __ldg(pin + (edf_mirror1(v.stz, lenz) * isz + edf_mirror1(v.sty, leny) * isy + edf_mirror1(v.stx, lenx)));
                        else val = edf_swin_direct_gather<ORDER>(pin, v.stz, v.sty, v.stx, v.ez + hh, v.ey + hh, v.ex + hh, lenz, leny, lenx, isz, isy);
                        pout[obase + (y0 + m + rseg + q) * osy] = val;
                    }
                }
            }
            __syncwarp();
            EDF_PS_MARK(1)
            if (lane == 0) edf_mbar_arrive(&s.empty[st]);
            m += b.nb;
            ++cc;
        }
        // rare voxels of the tile (~1 in 10^5 for a smooth field): exact reference-order coordinates and the general
        // single-voxel gather (any position, mirrored edge taps)
        edf_consumer_sync();                                       // the masks of a slab come from two warps
        EDF_PS_MARK(2)
#pragma unroll 1
        for (int q = half; q < nrow; q += 2) {
            const unsigned rm = s.rare[g][q];
#ifdef EDF_PIPE_STATS
            if (lane == 0 && rm) atomicAdd(&g_tile_prof[8], (unsigned long long)__popc(rm));
#endif
            if ((rm >> lane) & 1u) edf_poly_slow_voxel<ORDER, false>(p, L, ii, z, y0 + q, x);
        }
        EDF_PS_MARK(3)
    }
    EDF_PS_FLUSH
#ifdef EDF_PIPE_STATS
    if (tid == 0) {
        const unsigned long long tot = (unsigned long long)(clock64() - ps_t0);
        atomicMax(&g_tile_prof[13], tot);
        atomicAdd(&g_tile_prof[14], tot);
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static EdfPerDeviceFlag g_pipe_configured;

static bool edf_pipe_fwd_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    const EdfInputDesc& d = p.inp[ii];
    if (d.order < 0 || d.order > 3 || d.mode != EDF_MODE_CONSTANT) return false;
    if (!edf_tile_common_ok(p, L, ii)) return false;               // 3-D float32, unit x stride, 16-byte rows, control spans
    if (p.ncp[1] < 2 || p.idim[2] < EDF_PP_PITCH) return false;
    for (int a = 0; a < 3; ++a)
        if (p.idim[a] > (1 << 26) || p.odim[a] > (1 << 26)) return false;      // fixed-point coordinates: |c| < 2^28
    static int off = -1;                                           // EDF_NO_PIPE=1: previous kernels (A/B runs)
    if (off < 0) { const char* e = getenv("EDF_NO_PIPE"); off = (e && *e && *e != '0') ? 1 : 0; }
    return !off;
}

static int edf_pipe_sm_count()
{
    static std::atomic<int> n{0};
    int v = n.load(std::memory_order_relaxed);
    if (v <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n.store(v, std::memory_order_relaxed);
    }
    return v;
}

// returns 0 = launched, -2 = not applicable, -1 = CUDA error
static int edf_pipe_launch_fwd(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    EdfFastLaunch L = Lin;
    const uint64_t gx = (uint64_t)((p.odim[2] + EDF_PP_TX - 1) / EDF_PP_TX);
    const uint64_t gz = (uint64_t)((p.odim[0] + EDF_PP_G - 1) / EDF_PP_G);
    const int nsm = edf_pipe_sm_count();
    unsigned ry = EDF_PP_RY;
    static int env_ry = -1;                                        // EDF_PIPE_ROWS=8/16/32/64: rows per tile (A/B runs)
    if (env_ry < 0) { const char* e = getenv("EDF_PIPE_ROWS"); env_ry = (e && *e) ? atoi(e) : 0; }
    if (env_ry == 8 || env_ry == 16 || env_ry == 32 || env_ry == 64) ry = (unsigned)env_ry;
    else while (ry > EDF_PP_MR && gx * ((p.odim[1] + ry - 1) / ry) * gz < 6ull * (uint64_t)nsm) ry >>= 1;
    const uint64_t gy = (uint64_t)((p.odim[1] + ry - 1) / ry);
    const uint64_t ntiles = gx * gy * gz;
    if (ntiles == 0 || ntiles >= (1ull << 31)) return -2;
    L.rows_per_cta = ry;
    L.sched.nseg = 1;
    L.sched.gx = (unsigned)gx;
    L.sched.gy[0] = (unsigned)gy;
    L.sched.ry[0] = ry;
    L.sched.z_begin[0] = (unsigned)ntiles;
    alignas(64) CUtensorMap tm;
    if (!edf_tile_make_map(&tm, p.inp[ii].in, p, L.istr_e[ii][0], L.istr_e[ii][1])) return -2;
    const size_t smem = ((sizeof(EdfPipeSmem) + 1023) & ~(size_t)1023) + (size_t)EDF_PP_STAGES * EDF_PP_STAGE_ROWS * EDF_PP_PITCH * 4;
    if (!g_pipe_configured.test()) {
        cudaFuncSetAttribute(edf_pipe3d_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_pipe3d_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_pipe3d_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_pipe3d_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaGetLastError() != cudaSuccess) return -1;
        g_pipe_configured.set();
    }
    const unsigned grid = (unsigned)(ntiles < (uint64_t)nsm ? ntiles : (uint64_t)nsm);
    switch (order) {
    case 0: edf_pipe3d_fwd_kernel<0><<<grid, EDF_PP_THREADS, smem, st>>>(p, L, tm, ii); break;
    case 1: edf_pipe3d_fwd_kernel<1><<<grid, EDF_PP_THREADS, smem, st>>>(p, L, tm, ii); break;
    case 2: edf_pipe3d_fwd_kernel<2><<<grid, EDF_PP_THREADS, smem, st>>>(p, L, tm, ii); break;
    default: edf_pipe3d_fwd_kernel<3><<<grid, EDF_PP_THREADS, smem, st>>>(p, L, tm, ii); break;
    }
    return 0;
}
