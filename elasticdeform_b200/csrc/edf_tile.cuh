// edf_tile.cuh -- staged-window kernels of round 2 (K1t forward gather, K2t gradient scatter).
//
// Same idea as edf_swin.cuh (gather / scatter through a window of the volume held in shared memory, exact
// bounding box per chunk, rows 64 floats apart so that the bank of a tap is its x index mod 32), rebuilt
// around three changes:
//   * coordinates from the per-thread polynomial form of the displacement (edf_poly.cuh): 9 fp64 FMAs and one
//     broadcast load per voxel, no per-chunk table contraction;
//   * the window is filled by TENSOR-MAP TMA (cp.async.bulk.tensor.3d, SASS UTMALDG): boxes of 64 x 4 x 1
//     floats, one elected lane per warp issues the planes of its slab, completion through one transaction
//     mbarrier.  Rows / columns of a box that lie outside the volume arrive as zeros (TMA out-of-bounds fill)
//     and are then overwritten with the mirror images the reference's edge taps read (deform.c:796-810) by a
//     short patch pass that only border chunks run; planes outside the volume are fetched from their mirror
//     plane directly.  (The innermost box coordinate must be a multiple of 4 floats: an unaligned one traps
//     with cudaErrorIllegalInstruction -- scripts/experiments/tma_tensor_test.cu, profiles/r2/.)
//   * cubic weights in 11 operations per axis (power form) instead of 17.
// Chunks of 8 slabs x 4 rows x 32 columns as in edf_swin.cuh (8-row chunks halve the staged bytes per voxel, but
// their boxes outgrow a 100 KB window in a third of the chunks of the headline field; measured slower).  A chunk
// whose box outgrows the window is processed as two 2-row halves, then falls back to the direct gather.
// Two CTAs per SM: one CTA's coordinate phase and TMA wait overlap the other's gather.  (A warp-specialised
// producer / consumer pipeline over the same pieces is kept under csrc/experimental/: correct, but latency bound
// at the two stages that fit in shared memory -- see DESIGN.md.)
#pragma once
#include <cuda.h>
#include "edf_poly.cuh"

#define EDF_TL_MR 4                // rows per chunk (voxels per thread and chunk)
#define EDF_TL_PITCH 64            // floats between window rows
#define EDF_TL_BY 4                // rows per TMA box
#ifndef EDF_TL_ROWS
#define EDF_TL_ROWS 380            // window capacity in rows (95 KB); 2 CTAs per SM
#endif
#define EDF_TL_MAXQ (EDF_TL_PITCH / 4)

// Phase timing (debug builds with -DEDF_TILE_PROFILE): lane 0 of every warp accumulates the cycles it spends in
// each phase; edf_debug_tile_profile() reads and resets the totals.
__device__ unsigned long long g_tile_prof[16];
#ifdef EDF_TILE_PROFILE
#define EDF_TP_DECL long long tp_t = clock64(); long long tp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define EDF_TP_MARK(k) { const long long t_ = clock64(); tp_acc[k] += t_ - tp_t; tp_t = t_; }
#define EDF_TP_FLUSH if ((threadIdx.x & 31) == 0) { for (int k_ = 0; k_ < 8; ++k_) atomicAdd(&g_tile_prof[k_], (unsigned long long)tp_acc[k_]); atomicAdd(&g_tile_prof[15], 1ull); }
#else
#define EDF_TP_DECL
#define EDF_TP_MARK(k)
#define EDF_TP_FLUSH
#endif

struct EdfTileSmem {
    EdfPolyTables t;
    int bb[3][8];                  // [chunk % 3]: min z,y,x start, max z,y,x start, (gradient: max |dY| bits) of the chunk's active voxels
    int hb[2][8];                  // the same per 2-row half, only when the full box does not fit the window
    unsigned long long mbar;
};

__device__ __forceinline__ void edf_tma_box3d(uint32_t smem_dst, const CUtensorMap* tm, int cx, int cy, int cz, uint32_t mbar_s)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(cx), "r"(cy), "r"(cz), "r"(mbar_s) : "memory");
}

// Classification of one voxel from its un-mapped source coordinates.  CMODE ('constant'): the window start and
// fractional offsets come straight from the coordinate; a voxel next to ANY integer (odd orders) / half-integer
// or integer (even orders) -- which covers the floor thresholds and the range limits 0 and len-1 -- is redone in
// the reference order (`slow`).  Other modes: the logic of edf_swin_voxel (out-of-range coordinates mapped out of line).
template <int ORDER, bool CMODE>
__device__ __forceinline__ void edf_tile_classify(int mode, double inz, double iny, double inx, double limz, double limy, double limx,
                                                  int lenz, int leny, int lenx, bool gate,
                                                  int& stz, int& sty, int& stx, float& fz, float& fy, float& fx,
                                                  bool& slow, bool& cst, bool& oob)
{
    const bool inr = (inz >= 0.0) & (inz <= limz) & (iny >= 0.0) & (iny <= limy) & (inx >= 0.0) & (inx <= limx);
    if (CMODE) {
        edf_floor_split<ORDER>(inz, stz, fz);
        edf_floor_split<ORDER>(iny, sty, fy);
        edf_floor_split<ORDER>(inx, stx, fx);
        float dmax;
        if (ORDER & 1) {
            dmax = fmaxf(fmaxf(fabsf(fz - 0.5f), fabsf(fy - 0.5f)), fabsf(fx - 0.5f));
        } else {
            const float q0 = fabsf(fabsf(fz) - 0.25f), q1 = fabsf(fabsf(fy) - 0.25f), q2 = fabsf(fabsf(fx) - 0.25f);
            dmax = 2.0f * fmaxf(fmaxf(q0, q1), q2);
        }
        slow = gate & !(dmax < 0.5f - EDF_LEAN_EPSF);              // NaN -> slow
        oob = !inr;
        cst = !inr;
        return;
    }
    const bool loz = !(inz >= 0.0), hiz = inz > limz;             // NaN counts as "low"
    const bool loy = !(iny >= 0.0), hiy = iny > limy;
    const bool lox = !(inx >= 0.0), hix = inx > limx;
    double cz = loz ? 0.0 : (hiz ? limz : inz);
    double cy = loy ? 0.0 : (hiy ? limy : iny);
    double cx = lox ? 0.0 : (hix ? limx : inx);
    bool mapped_danger = false, nanflag = false;
    if (!inr) {
        if (loz | hiz) { mapped_danger |= edf_near_half_integer(inz); cz = edf_map_coordinate_cold(inz, lenz, mode); }
        if (loy | hiy) { mapped_danger |= edf_near_half_integer(iny); cy = edf_map_coordinate_cold(iny, leny, mode); }
        if (lox | hix) { mapped_danger |= edf_near_half_integer(inx); cx = edf_map_coordinate_cold(inx, lenx, mode); }
        if (!((cz > -1.0) & (cy > -1.0) & (cx > -1.0))) { nanflag = true; cz = cy = cx = 0.0; }   // NaN
    }
    edf_floor_split<ORDER>(cz, stz, fz);
    edf_floor_split<ORDER>(cy, sty, fy);
    edf_floor_split<ORDER>(cx, stx, fx);
    bool danger;
    if (ORDER & 1)
        danger = (fz < EDF_LEAN_EPSF) | (fz > 1.0f - EDF_LEAN_EPSF) | (fy < EDF_LEAN_EPSF) |
                 (fy > 1.0f - EDF_LEAN_EPSF) | (fx < EDF_LEAN_EPSF) | (fx > 1.0f - EDF_LEAN_EPSF);
    else
        danger = (fabsf(fz) < EDF_LEAN_EPSF) | (fabsf(fz) > 0.5f - EDF_LEAN_EPSF) |
                 (fabsf(fy) < EDF_LEAN_EPSF) | (fabsf(fy) > 0.5f - EDF_LEAN_EPSF) |
                 (fabsf(fx) < EDF_LEAN_EPSF) | (fabsf(fx) > 0.5f - EDF_LEAN_EPSF);
    danger |= mapped_danger;
    oob = !inr;
    cst = false;
    slow = (gate & danger) | nanflag;
}

// tile of this CTA from the 1-D block index: x fastest, then y, then z
__device__ __forceinline__ void edf_tile_origin(const EdfFastLaunch& L, int& x0, int& y0, int& z0)
{
    const unsigned gx = L.sched.gx, gy = L.sched.gy[0];
    const unsigned bid = blockIdx.x;
    const unsigned tx = bid % gx, t = bid / gx;
    x0 = (int)tx * EDF_PL_TX;
    y0 = (int)(t % gy) * (int)L.rows_per_cta;
    z0 = (int)(t / gy) * EDF_PL_G;
}

// Window geometry of one pass (CTA-uniform, from a bounding box of window starts)
struct EdfTileBox {
    int wz0, wy0, wx0, nzw, nyal, nq;
    bool empty, fit;
};
template <int ORDER>
__device__ __forceinline__ EdfTileBox edf_tile_box(int mnz, int mny, int mnx, int mxz, int mxy, int mxx)
{
    constexpr int NT = ORDER + 1;
    EdfTileBox b;
    b.empty = mnz > mxz;
    b.wz0 = mnz; b.wy0 = mny; b.wx0 = mnx & ~3;
    b.nzw = mxz - mnz + NT;
    const int nyw = mxy - mny + NT;
    b.nyal = (nyw + EDF_TL_BY - 1) & ~(EDF_TL_BY - 1);
    b.nq = ((mxx + NT - 1 - b.wx0) >> 2) + 1;
    b.fit = !b.empty && b.nq <= EDF_TL_MAXQ && b.nzw <= EDF_TL_ROWS && b.nyal <= EDF_TL_ROWS && b.nzw * b.nyal <= EDF_TL_ROWS;
    return b;
}

// Issue the TMA boxes of a window (lane 0 of every warp: the planes zr = warp, warp + 8, ...) and arrive on the
// transaction barrier with the byte count (the barrier expects one arrival per warp).
__device__ __forceinline__ void edf_tile_stage(const CUtensorMap* tm, const EdfTileBox& b, uint32_t win_s, uint32_t mbar_s,
                                               int warp, int lane, int lenz)
{
    if (lane == 0) {
        const int groups = b.nyal / EDF_TL_BY;
        int nplanes = 0;
        for (int zr = warp; zr < b.nzw; zr += EDF_PL_G) ++nplanes;
        const unsigned bytes = (unsigned)(nplanes * groups) * (EDF_TL_PITCH * EDF_TL_BY * 4u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(bytes) : "memory");
        for (int zr = warp; zr < b.nzw; zr += EDF_PL_G) {
            const int gz = edf_mirror1(b.wz0 + zr, lenz);
            uint32_t dst = win_s + (uint32_t)(zr * b.nyal) * (EDF_TL_PITCH * 4u);
            for (int q = 0; q < groups; ++q, dst += EDF_TL_PITCH * EDF_TL_BY * 4u)
                edf_tma_box3d(dst, tm, b.wx0, b.wy0 + q * EDF_TL_BY, gz, mbar_s);
        }
    }
}

// Border chunks: cells of the window that lie outside the volume along y or x (TMA delivered zeros) take the value
// of their mirror cell (deform.c:796-810; single reflection, extents >= 8).  All sources are in-volume cells of the
// same plane, which the pass never writes, so no ordering is needed inside it.  CTA-collective; ends with a barrier.
template <int ORDER>
__device__ __forceinline__ void edf_tile_patch(float* win, const EdfTileBox& b, int leny, int lenx, int tid)
{
    const int rows = b.nzw * b.nyal;
    const int ncol = 4 * b.nq;
    const bool ylo = b.wy0 < 0, yhi = b.wy0 + b.nyal > leny;
    const bool xlo = b.wx0 < 0, xhi = b.wx0 + ncol > lenx;
    if (xlo | xhi) {
        // in-volume rows: the out-of-volume columns (at most `order` cells deep on either side matter; all of them are patched)
        const int nlo = xlo ? -b.wx0 : 0;                         // columns [0, nlo) are left of the volume
        const int chi = xhi ? lenx - b.wx0 : ncol;                // columns [chi, ncol) are right of it
        const int nout = nlo + (ncol - chi);
        for (int e = tid; e < rows * nout; e += EDF_PL_THREADS) {
            const int r = e / nout, k = e - r * nout;
            const int c = k < nlo ? k : chi + (k - nlo);
            const int yr = r % b.nyal;
            const int gy = b.wy0 + yr;
            if ((unsigned)gy >= (unsigned)leny) continue;         // out-of-volume rows: below
            const int cm = edf_mirror1(b.wx0 + c, lenx) - b.wx0;
            if ((unsigned)cm < (unsigned)ncol) win[r * EDF_TL_PITCH + c] = win[r * EDF_TL_PITCH + cm];
        }
    }
    if (ylo | yhi) {
        // out-of-volume rows: every column, from the mirror row (and mirror column where that is outside, too)
        for (int e = tid; e < rows * ncol; e += EDF_PL_THREADS) {
            const int r = e / ncol, c = e - r * ncol;
            const int zr = r / b.nyal, yr = r - zr * b.nyal;
            const int gy = b.wy0 + yr;
            if ((unsigned)gy < (unsigned)leny) continue;
            const int ym = edf_mirror1(gy, leny) - b.wy0;
            const int cm = edf_mirror1(b.wx0 + c, lenx) - b.wx0;
            if ((unsigned)ym < (unsigned)b.nyal && (unsigned)cm < (unsigned)ncol)
                win[r * EDF_TL_PITCH + c] = win[(zr * b.nyal + ym) * EDF_TL_PITCH + cm];
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes before the next TMA fill of these cells
    __syncthreads();
}

// cubic B-spline weights from the fractional offset, 11 operations per axis (the reference's closed forms,
// deform.c:171-177, expanded in powers of t; differences to the reference's float evaluation are ~1 ulp)
__device__ __forceinline__ void edf_tile_weights3(float t, float* w)
{
    const float t2 = t * t, t3 = t2 * t;
    w[3] = t3 * (1.0f / 6.0f);
    w[0] = fmaf(-t3, 1.0f / 6.0f, fmaf(t2, 0.5f, fmaf(t, -0.5f, 1.0f / 6.0f)));
    w[1] = fmaf(t3, 0.5f, fmaf(t2, -1.0f, 2.0f / 3.0f));
    w[2] = fmaf(t3, -0.5f, fmaf(t2, 0.5f, fmaf(t, 0.5f, 1.0f / 6.0f)));
}
template <int ORDER>
__device__ __forceinline__ void edf_tile_weights(float t, float* w)
{
    if (ORDER == 3) edf_tile_weights3(t, w);
    else edf_bspline_weights_f32<ORDER>(t, w);
}

// voxel whose chunk does not fit the window at all (very steep field): taps straight from global memory.  Out of
// line with scalar arguments, so that the per-voxel state arrays of the caller keep static indices (registers).
template <int ORDER>
__device__ __noinline__ float edf_tile_direct_voxel(const float* __restrict__ pin, unsigned pk, float fz, float fy, float fx,
                                                    int z, int y, int x, int lenz, int leny, int lenx, int isz, int isy)
{
    const int stz = z - EDF_SW_PK_BIAS + (int)(pk >> 20), sty = y - EDF_SW_PK_BIAS + (int)((pk >> 10) & 1023u);
    const int stx = x - EDF_SW_PK_BIAS + (int)(pk & 1023u);
    return edf_swin_direct_gather<ORDER>(pin, stz, sty, stx, fz, fy, fx, lenz, leny, lenx, isz, isy);
}

template <int ORDER, bool CMODE>
__global__ void __launch_bounds__(EDF_PL_THREADS, 2)
edf_tile3d_fwd_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L,
                      const __grid_constant__ CUtensorMap tmap, const int ii)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    EdfTileSmem& s = *reinterpret_cast<EdfTileSmem*>(smem_raw);
    float* const win = reinterpret_cast<float*>(smem_raw + ((sizeof(EdfTileSmem) + 1023) & ~(size_t)1023));
    constexpr int NT = ORDER + 1;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;     // warp = slab
    int x0, y0, z0;
    edf_tile_origin(L, x0, y0, z0);
    const int ry = (int)L.rows_per_cta;

    if (tid == 0) edf_mbar_init(&s.mbar, EDF_PL_G);
    if (tid < 3 * 8) (&s.bb[0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
    if (tid >= 32 && tid < 48) (&s.hb[0][0])[tid - 32] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
    EDF_TP_DECL
    edf_poly_tables(p, s.t, z0, y0, x0, ry);                      // ends with a CTA barrier
    EDF_TP_MARK(0)

    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    const int nrow = min(ry, ody - y0);

    const EdfInputDesc& d = p.inp[ii];
    const float* __restrict__ pin = (const float*)d.in;
    float* __restrict__ pout = (float*)d.out;
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int obase_zx = z * L.ostr_e[ii][0] + x * L.ostr_e[ii][2];   // element offsets fit 32 bits (host-checked)
    const bool affine = p.has_affine != 0;
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const double bz = xadd((double)z, p.ooff_d[0]);
    const double bx = xadd((double)x, p.ooff_d[2]);
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&s.mbar);

    double a[12];
    int jcur = INT_MIN;
    bool gate = false;
    int par = 0;
    unsigned mphase = 0;
    int nb = EDF_TL_MR;
    double by = xadd((double)y0, p.ooff_d[1]);                     // exact: integers

    for (int m0 = 0; m0 < nrow; m0 += nb) {
        const int yc0 = y0 + m0;
        // rows of this chunk: up to 4, all inside one control interval of the y axis (CTA-uniform), so that one
        // polynomial serves the chunk; the rebuild below is its only call site
        {
            const int jr = s.t.jy[m0];
            if (jr != jcur) {
                gate = edf_poly_build(p, s.t, g, lane, jr, a);
                jcur = jr;
            }
            nb = min(EDF_TL_MR, nrow - m0);
#pragma unroll
            for (int u = EDF_TL_MR - 1; u >= 1; --u)
                if (u < nb && s.t.jy[m0 + u] != jr) nb = u;
        }
        // ---- phase A: coordinates and classification of this thread's voxels
        unsigned pk[EDF_TL_MR];
        float fz[EDF_TL_MR], fy[EDF_TL_MR], fx[EDF_TL_MR];
        unsigned actm = 0, cstm = 0, slowm = 0;
        int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
        double yd = by;
#pragma unroll
        for (int u = 0; u < EDF_TL_MR; ++u) {
            const int m = m0 + min(u, nb - 1);
            const int y = y0 + m;
            const bool valid = tok && (u < nb);
            double dz, dy, dx;
            edf_poly_eval(a, s.t.u[m], dz, dy, dx);
            double inz, iny, inx;
            if (!affine) {
                inz = xadd(bz, dz);
                iny = xadd(yd, dy);
                inx = xadd(bx, dx);
            } else {
                const int o[3] = {z, y, x};
                inz = edf_source_coordinate<3, int>(p, o, 0, dz);
                iny = edf_source_coordinate<3, int>(p, o, 1, dy);
                inx = edf_source_coordinate<3, int>(p, o, 2, dx);
            }
            if (u + 1 < nb) yd = xadd(yd, 1.0);
            int stz, sty, stx;
            bool slow, cst, oob;
            edf_tile_classify<ORDER, CMODE>(d.mode, inz, iny, inx, limz, limy, limx, lenz, leny, lenx, gate,
                                            stz, sty, stx, fz[u], fy[u], fx[u], slow, cst, oob);
            const bool packed = edf_swin_pack(stz - z, sty - y, stx - x, pk[u]);
            slow = valid & (slow | (!cst & !packed));
            if (slow) slowm |= 1u << u;
            if (valid & !slow & cst) cstm |= 1u << u;
            if (valid & !slow & !cst) {
                actm |= 1u << u;
                mn[0] = min(mn[0], stz); mn[1] = min(mn[1], sty); mn[2] = min(mn[2], stx);
                mx[0] = max(mx[0], stz); mx[1] = max(mx[1], sty); mx[2] = max(mx[2], stx);
            }
        }
        by = xadd(by, (double)nb);
        // ---- phase B: exact bounding box of the chunk's tap windows
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            mn[q] = __reduce_min_sync(0xffffffffu, mn[q]);
            mx[q] = __reduce_max_sync(0xffffffffu, mx[q]);
        }
        if (lane == 0 && mn[0] != INT_MAX) {
            int* b = s.bb[par];
            atomicMin(b + 0, mn[0]); atomicMin(b + 1, mn[1]); atomicMin(b + 2, mn[2]);
            atomicMax(b + 3, mx[0]); atomicMax(b + 4, mx[1]); atomicMax(b + 5, mx[2]);
        }
        EDF_TP_MARK(1)
        __syncthreads();                                           // box complete; previous chunk's gathers done
        EDF_TP_MARK(2)
        // the box of chunk c+2 (= chunk c-1's, no longer read) is reset here: chunk c+2's atomics come after
        // the next chunk's barrier
        if (tid < 8) s.bb[par == 0 ? 2 : par - 1][tid] = (tid < 3) ? INT_MAX : INT_MIN;
        const int* b0 = s.bb[par];
        const EdfTileBox full = edf_tile_box<ORDER>(b0[0], b0[1], b0[2], b0[3], b0[4], b0[5]);
        const int npass = (full.fit || full.empty) ? 1 : 2;
        if (npass == 2) {
            // the box outgrows the window (steep field): boxes of the two 2-row halves, reduced from the packed starts
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int hn[3] = {INT_MAX, INT_MAX, INT_MAX}, hx[3] = {INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
                for (int u = 2 * h; u < 2 * h + 2; ++u)
                    if ((actm >> u) & 1u) {
                        const int stz = z - EDF_SW_PK_BIAS + (int)(pk[u] >> 20), sty = yc0 + u - EDF_SW_PK_BIAS + (int)((pk[u] >> 10) & 1023u);
                        const int stx = x - EDF_SW_PK_BIAS + (int)(pk[u] & 1023u);
                        hn[0] = min(hn[0], stz); hn[1] = min(hn[1], sty); hn[2] = min(hn[2], stx);
                        hx[0] = max(hx[0], stz); hx[1] = max(hx[1], sty); hx[2] = max(hx[2], stx);
                    }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    hn[q] = __reduce_min_sync(0xffffffffu, hn[q]);
                    hx[q] = __reduce_max_sync(0xffffffffu, hx[q]);
                }
                if (lane == 0 && hn[0] != INT_MAX) {
                    int* b = s.hb[h];
                    atomicMin(b + 0, hn[0]); atomicMin(b + 1, hn[1]); atomicMin(b + 2, hn[2]);
                    atomicMax(b + 3, hx[0]); atomicMax(b + 4, hx[1]); atomicMax(b + 5, hx[2]);
                }
            }
            __syncthreads();
        }
        for (int ps = 0; ps < npass; ++ps) {
            EdfTileBox bx_ = full;
            unsigned rowmask = 0xfu;
            if (npass == 2) {
                const int* bh = s.hb[ps];
                bx_ = edf_tile_box<ORDER>(bh[0], bh[1], bh[2], bh[3], bh[4], bh[5]);
                rowmask = ps ? 0xcu : 0x3u;
                if (ps) __syncthreads();                           // first pass's gathers done before the window is refilled
            }
            const unsigned am = actm & rowmask;
            if (bx_.fit) {
                // ---- phase C: fill the window (tensor-map TMA), patch the border cells
                edf_tile_stage(&tmap, bx_, win_s, mbar_s, g, lane, lenz);
                edf_mbar_wait(&s.mbar, mphase);
                mphase ^= 1u;
                const bool border = (bx_.wy0 < 0) | (bx_.wy0 + bx_.nyal > leny) | (bx_.wx0 < 0) | (bx_.wx0 + 4 * bx_.nq > lenx);
                if (border) edf_tile_patch<ORDER>(win, bx_, leny, lenx, tid);
                EDF_TP_MARK(3)
                // ---- phase D: gather from the window (inactive lanes read cell 0 and discard)
                const int slab = bx_.nyal * EDF_TL_PITCH;
                const int lin0 = ((z - EDF_SW_PK_BIAS - bx_.wz0) * bx_.nyal + (yc0 - EDF_SW_PK_BIAS - bx_.wy0)) * EDF_TL_PITCH +
                                 (x - EDF_SW_PK_BIAS - bx_.wx0);
#pragma unroll
                for (int u = 0; u < EDF_TL_MR; ++u) {
                    const bool act = (am >> u) & 1u;
                    if (!__any_sync(0xffffffffu, act)) continue;
                    const int rz = (int)(pk[u] >> 20), ryw = (int)((pk[u] >> 10) & 1023u), rx = (int)(pk[u] & 1023u);
                    const int off = act ? lin0 + u * EDF_TL_PITCH + (rz * bx_.nyal + ryw) * EDF_TL_PITCH + rx : 0;
                    const float* q0 = win + off;
                    float wzf[NT], wyf[NT], wxf[NT];
                    edf_tile_weights<ORDER>(fz[u], wzf);
                    edf_tile_weights<ORDER>(fy[u], wyf);
                    edf_tile_weights<ORDER>(fx[u], wxf);
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const float* qi = q0 + i * slab;
                        float ti = 0.f;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const float* r = qi + j * EDF_TL_PITCH;
                            float tj = r[0] * wxf[0];
#pragma unroll
                            for (int k = 1; k < NT; ++k) tj = fmaf(r[k], wxf[k], tj);
                            ti = (j == 0) ? tj * wyf[0] : fmaf(tj, wyf[j], ti);
                        }
                        acc = (i == 0) ? ti * wzf[0] : fmaf(ti, wzf[i], acc);
                    }
                    if (act) pout[obase_zx + (yc0 + u) * osy] = acc;
                }
                EDF_TP_MARK(4)
            } else if (!bx_.empty) {
                // not even a 2-row half fits the window (very steep field): straight from global memory
#pragma unroll
                for (int u = 0; u < EDF_TL_MR; ++u)
                    if ((am >> u) & 1u)
                        pout[obase_zx + (yc0 + u) * osy] =
                            edf_tile_direct_voxel<ORDER>(pin, pk[u], fz[u], fy[u], fx[u], z, yc0 + u, x, lenz, leny, lenx, isz, isy);
            }
        }
        if (npass == 2) {
            __syncthreads();                                       // half boxes read by every thread: reset for the next user
            if (tid < 16) (&s.hb[0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
        }
#pragma unroll
        for (int u = 0; u < EDF_TL_MR; ++u)
            if ((cstm >> u) & 1u) pout[obase_zx + (yc0 + u) * osy] = cvalf;            // deform.c:903
        // ---- rare voxels (next to a rounding / boundary threshold, huge displacements): reference order
        if (slowm) {
#pragma unroll 1
            for (int u = 0; u < EDF_TL_MR; ++u)
                if ((slowm >> u) & 1u) edf_poly_slow_voxel<ORDER, false>(p, L, ii, z, yc0 + u, x);
        }
        EDF_TP_MARK(5)
        par = par == 2 ? 0 : par + 1;
    }
    EDF_TP_FLUSH
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EdfTensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EdfTensorMapEncodeFn edf_tensor_map_encoder()
{
    static std::atomic<void*> fn{nullptr};
    void* f = fn.load(std::memory_order_acquire);
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        void* ptr = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) {
            cudaGetLastError();
            return nullptr;
        }
        fn.store(ptr, std::memory_order_release);
        f = ptr;
    }
    return (EdfTensorMapEncodeFn)f;
}

// tensor map of a float32 / int32 volume [lenz][leny][lenx] (element strides isz, isy, 1) with boxes of 64 x 4 x 1
static bool edf_tile_make_map(CUtensorMap* tm, const void* base, const EdfParams& p, int isz, int isy)
{
    EdfTensorMapEncodeFn enc = edf_tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.idim[2], (cuuint64_t)p.idim[1], (cuuint64_t)p.idim[0]};
    const cuuint64_t gstr[2] = {(cuuint64_t)isy * 4ull, (cuuint64_t)isz * 4ull};
    const cuuint32_t box[3] = {EDF_TL_PITCH, EDF_TL_BY, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static EdfPerDeviceFlag g_tile_configured;

static bool edf_tile_common_ok(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    if (!edf_swin_common_ok(p, L, ii)) return false;               // lean-eligible, not 'wrap', 16-byte aligned rows, lenx % 4 == 0
    if (!edf_fast_ctrl_span_ok(p, 2, EDF_PL_TX, EDF_PL_NC)) return false;
    if ((p.idim[1] - 1) < 8 * (p.ncp[1] - 1)) return false;        // a control interval spans several rows
    // tensor-map limits: strides < 2^40 bytes and multiples of 16 (checked above), extents < 2^32
    static int off = -1;                                           // EDF_NO_TILE=1: round-1 kernels (A/B runs)
    if (off < 0) { const char* e = getenv("EDF_NO_TILE"); off = (e && *e && *e != '0') ? 1 : 0; }
    return !off;
}

static bool edf_tile_fwd_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    const EdfInputDesc& d = p.inp[ii];
    if (d.order < 2 || d.order > 3) return false;
    return edf_tile_common_ok(p, L, ii);
}

static void edf_tile_grid(const EdfParams& p, EdfFastLaunch& L, unsigned& ncta)
{
    const uint64_t gx = (uint64_t)((p.odim[2] + EDF_PL_TX - 1) / EDF_PL_TX);
    const uint64_t gz = (uint64_t)((p.odim[0] + EDF_PL_G - 1) / EDF_PL_G);
    unsigned ry = 16;
    static int env_ry = -1;                                        // EDF_TILE_ROWS=8/16/32/64: rows per CTA (A/B runs)
    if (env_ry < 0) { const char* e = getenv("EDF_TILE_ROWS"); env_ry = (e && *e) ? atoi(e) : 0; }
    if (env_ry == 8 || env_ry == 16 || env_ry == 32 || env_ry == 64) ry = (unsigned)env_ry;
    while (ry > EDF_TL_MR && gx * ((p.odim[1] + ry - 1) / ry) * gz < 4ull * 148) ry >>= 1;
    const uint64_t gy = (uint64_t)((p.odim[1] + ry - 1) / ry);
    L.rows_per_cta = ry;
    L.sched.nseg = 1;
    L.sched.gx = (unsigned)gx;
    L.sched.gy[0] = (unsigned)gy;
    L.sched.ry[0] = ry;
    ncta = (unsigned)(gx * gy * gz);
}

// returns 0 = launched, -2 = not applicable, -1 = CUDA error
static int edf_tile_launch_fwd(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    EdfFastLaunch L = Lin;
    unsigned grid = 0;
    edf_tile_grid(p, L, grid);
    if (grid == 0 || (uint64_t)grid >= (1ull << 31)) return -2;
    alignas(64) CUtensorMap tm;
    if (!edf_tile_make_map(&tm, p.inp[ii].in, p, L.istr_e[ii][0], L.istr_e[ii][1])) return -2;
    const size_t smem = ((sizeof(EdfTileSmem) + 1023) & ~(size_t)1023) + (size_t)EDF_TL_ROWS * EDF_TL_PITCH * 4;
    if (!g_tile_configured.test()) {
        cudaFuncSetAttribute(edf_tile3d_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_tile3d_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_tile3d_fwd_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_tile3d_fwd_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaGetLastError() != cudaSuccess) return -1;
        g_tile_configured.set();
    }
    const bool cm = p.inp[ii].mode == EDF_MODE_CONSTANT;
    if (order == 2) {
        if (cm) edf_tile3d_fwd_kernel<2, true><<<grid, EDF_PL_THREADS, smem, st>>>(p, L, tm, ii);
        else    edf_tile3d_fwd_kernel<2, false><<<grid, EDF_PL_THREADS, smem, st>>>(p, L, tm, ii);
    } else {
        if (cm) edf_tile3d_fwd_kernel<3, true><<<grid, EDF_PL_THREADS, smem, st>>>(p, L, tm, ii);
        else    edf_tile3d_fwd_kernel<3, false><<<grid, EDF_PL_THREADS, smem, st>>>(p, L, tm, ii);
    }
    return 0;
}
