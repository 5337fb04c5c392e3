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
//   * WARP SPECIALISATION.  One CTA of 24 warps per SM.  Eight PRODUCER warps (one per slab) run ahead: they
//     evaluate the coordinates of a chunk of up to 8 rows x 8 slabs x 32 columns, write one 16-byte record per
//     voxel (packed window start + three fractional offsets) to shared memory, reduce the chunk's bounding box
//     among themselves (named barrier), re-cut the chunk with half as many rows while its box outgrows a
//     window, and issue the TMA fill.  Sixteen CONSUMER warps wait on the stage's "full" transaction barrier
//     (producer arrivals + TMA bytes), gather their rows of the chunk in a plain loop over the records (one
//     64-tap body in the instruction cache instead of 4-8 unrolled copies) and arrive on the stage's "empty"
//     barrier.  Two stages (records + window each): coordinates, box logic, TMA latency and the gather of
//     different chunks overlap, and no CTA-wide barrier is left in the steady state.
#pragma once
#include <cuda.h>
#include "edf_poly.cuh"

#define EDF_TL_MR 8                // rows per chunk (at most)
#ifndef EDF_TL_CW
#define EDF_TL_CW 8                // consumer warps (multiple of the 8 slabs)
#endif
#define EDF_TL_PW 8                // producer warps = slabs per CTA
#define EDF_TL_CTHREADS (EDF_TL_CW * 32)
#define EDF_TL_PTHREADS (EDF_TL_PW * 32)
#define EDF_TL_THREADS (EDF_TL_CTHREADS + EDF_TL_PTHREADS)
#define EDF_TL_STAGES 2
#define EDF_TL_PITCH 64            // floats between window rows
#define EDF_TL_BY 4                // rows per TMA box
#ifndef EDF_TL_ROWS
#define EDF_TL_ROWS 288            // capacity of one window in rows (72 KB); two stages per CTA, one CTA per SM
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

// chunk descriptor written by the producers, read by the consumers
struct EdfTileMeta {
    int m0, nb;                    // rows [m0, m0 + nb) of the tile
    int wz0, wy0, wx0, nzw, nyal, nq;
    int staged, last;              // window filled by TMA (else: no active voxel, or nothing fits); last chunk of the tile
    int pad_[6];
};

struct EdfTileSmem {
    EdfPolyTables t;
    int bb[3][4][8];               // [chunk % 3][attempt]: min z,y,x start, max z,y,x start, (gradient: max |dY| bits) of the active voxels
    EdfTileMeta meta[EDF_TL_STAGES];
    unsigned long long full[EDF_TL_STAGES];    // producers' arrivals + TMA bytes of the stage
    unsigned long long empty[EDF_TL_STAGES];   // consumers' arrivals: records and window of the stage are free again
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
    // columns: at most the window's 64; voxels whose taps lie beyond (a warp stretched over more than ~60 columns)
    // are caught at gather time and take the single-voxel routine
    b.nq = min(((mxx + NT - 1 - b.wx0) >> 2) + 1, EDF_TL_MAXQ);
    b.fit = !b.empty && b.nzw <= EDF_TL_ROWS && b.nyal <= EDF_TL_ROWS && b.nzw * b.nyal <= EDF_TL_ROWS;
    return b;
}

// Issue the TMA boxes of a window (lane 0 of every producer warp: the planes zr = warp, warp + 8, ...) and arrive on
// the stage's transaction barrier with the byte count (the barrier expects one arrival per producer warp).
__device__ __forceinline__ void edf_tile_stage(const CUtensorMap* tm, const EdfTileBox& b, uint32_t win_s, uint32_t mbar_s,
                                               int warp, int lane, int lenz)
{
    if (lane == 0) {
        const int groups = b.nyal / EDF_TL_BY;
        int nplanes = 0;
        for (int zr = warp; zr < b.nzw; zr += EDF_TL_PW) ++nplanes;
        const unsigned bytes = (unsigned)(nplanes * groups) * (EDF_TL_PITCH * EDF_TL_BY * 4u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(bytes) : "memory");
        for (int zr = warp; zr < b.nzw; zr += EDF_TL_PW) {
            const int gz = edf_mirror1(b.wz0 + zr, lenz);
            uint32_t dst = win_s + (uint32_t)(zr * b.nyal) * (EDF_TL_PITCH * 4u);
            for (int q = 0; q < groups; ++q, dst += EDF_TL_PITCH * EDF_TL_BY * 4u)
                edf_tma_box3d(dst, tm, b.wx0, b.wy0 + q * EDF_TL_BY, gz, mbar_s);
        }
    }
}

// Border chunks: cells of the window that lie outside the volume along y or x (TMA delivered zeros) take the value
// of their mirror cell (deform.c:796-810; single reflection, extents >= 8).  All sources are in-volume cells of the
// same plane, which the pass never writes, so no ordering is needed inside it.  Collective over the consumer warps
// (tid < EDF_TL_CTHREADS); ends with their named barrier.
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
        for (int e = tid; e < rows * nout; e += EDF_TL_CTHREADS) {
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
        for (int e = tid; e < rows * ncol; e += EDF_TL_CTHREADS) {
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
    asm volatile("bar.sync 1, %0;" :: "n"(EDF_TL_CTHREADS) : "memory");   // consumer warps only
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

// mbarrier wait with a watchdog in debug builds (-DEDF_TILE_DEBUG): reports and traps instead of hanging
__device__ __forceinline__ void edf_tile_wait(unsigned long long* mbar, unsigned parity, int what, int stage)
{
#ifdef EDF_TILE_DEBUG
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    for (long long it = 0;; ++it) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (it > 2000000) {
            if ((threadIdx.x & 31) == 0)
                printf("edf_tile_wait timeout: block %d warp %d what %d stage %d parity %u\n", blockIdx.x, threadIdx.x >> 5, what, stage, parity);
            __trap();
        }
    }
#else
    edf_mbar_wait(mbar, parity);
#endif
}

__device__ __forceinline__ void edf_mbar_arrive(uint32_t mbar_s)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar_s) : "memory");
}

// record flags (top two bits of the packed window start)
#define EDF_TL_F_NONE 0u
#define EDF_TL_F_ACT 1u
#define EDF_TL_F_CST 2u
#define EDF_TL_F_SLOW 3u

template <int ORDER, bool CMODE>
__global__ void __launch_bounds__(EDF_TL_THREADS, 1)
edf_tile3d_fwd_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L,
                      const __grid_constant__ CUtensorMap tmap, const int ii)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    EdfTileSmem& s = *reinterpret_cast<EdfTileSmem*>(smem_raw);
    constexpr size_t REC_OFF = (sizeof(EdfTileSmem) + 1023) & ~(size_t)1023;
    constexpr int RECS = EDF_TL_MR * EDF_TL_PW * 32;               // records per stage
    constexpr size_t WIN_OFF = REC_OFF + (size_t)EDF_TL_STAGES * RECS * 16;
    constexpr int WINF = EDF_TL_ROWS * EDF_TL_PITCH;               // floats per window
    uint4* const rec0 = reinterpret_cast<uint4*>(smem_raw + REC_OFF);
    float* const win0 = reinterpret_cast<float*>(smem_raw + WIN_OFF);
    constexpr int NT = ORDER + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int x0, y0, z0;
    edf_tile_origin(L, x0, y0, z0);
    const int ry = (int)L.rows_per_cta;

    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < EDF_TL_STAGES; ++q) { edf_mbar_init(&s.full[q], EDF_TL_PW); edf_mbar_init(&s.empty[q], EDF_TL_CW); }
    }
    if (tid < 3 * 4 * 8) (&s.bb[0][0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
    edf_poly_tables(p, s.t, z0, y0, x0, ry);                      // ends with a CTA barrier

    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const int nrow = min(ry, ody - y0);
    const EdfInputDesc& d = p.inp[ii];
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const int x = x0 + lane;

    if (warp >= EDF_TL_CW) {
        // =================================== producers ===================================
        const int pw = warp - EDF_TL_CW;                           // slab of this warp
        const int ptid = tid - EDF_TL_CTHREADS;
        const int z = z0 + pw;
        const bool tok = (x < odx) && (z < odz);
        const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
        const bool affine = p.has_affine != 0;
        const double bz = xadd((double)z, p.ooff_d[0]);
        const double bx = xadd((double)x, p.ooff_d[2]);
        const double offy = p.ooff_d[1];
        const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win0);
        const uint32_t full_s = (uint32_t)__cvta_generic_to_shared(&s.full[0]);
        double a[12];
        int jcur = INT_MIN;
        bool gate = false;
        int nb_pref = EDF_TL_MR;
        int par = 0;
        unsigned ephase = 0;                                       // bit q: parity of the next wait on empty[q]
        int stage = 0;
        for (int m0 = 0; m0 < nrow;) {
            // the stage's records and window are free once the consumers have arrived (first use: nothing to wait for)
            if ((ephase >> (stage + 8)) & 1u) {                    // the stage has been used before
                edf_tile_wait(&s.empty[stage], (ephase >> stage) & 1u, 1, stage);
                ephase ^= 1u << stage;
            }
            ephase |= 1u << (stage + 8);
            const int jr = s.t.jy[m0];
            if (jr != jcur) {                                      // the row entered another control interval
                gate = edf_poly_build(p, s.t, pw, lane, jr, a, pw);
                jcur = jr;
            }
            int nb = min(nb_pref, nrow - m0);
#pragma unroll
            for (int u = EDF_TL_MR - 1; u >= 1; --u)
                if (u < nb && s.t.jy[m0 + u] != jr) nb = u;
            uint4* const rec = rec0 + stage * RECS + pw * 32 + lane;          // + r * (PW * 32)
            int att = 0;
            EdfTileBox box;
            for (;;) {
                // ---- coordinates, classification, records of this thread's column; bounding box of the active voxels
                int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
                double yd = xadd((double)(y0 + m0), offy);
#pragma unroll 2
                for (int r = 0; r < nb; ++r) {
                    const int y = y0 + m0 + r;
                    double dz, dy, dx;
                    edf_poly_eval(a, s.t.u[m0 + r], dz, dy, dx);
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
                    yd = xadd(yd, 1.0);
                    int stz, sty, stx;
                    float fz, fy, fx;
                    bool slow, cst, oob;
                    edf_tile_classify<ORDER, CMODE>(d.mode, inz, iny, inx, limz, limy, limx, lenz, leny, lenx, gate,
                                                    stz, sty, stx, fz, fy, fx, slow, cst, oob);
                    unsigned pk;
                    const bool packed = edf_swin_pack(stz - z, sty - y, stx - x, pk);
                    slow = slow | (!cst & !packed);
                    unsigned flag = slow ? EDF_TL_F_SLOW : (cst ? EDF_TL_F_CST : EDF_TL_F_ACT);
                    if (!tok) flag = EDF_TL_F_NONE;
                    if (flag == EDF_TL_F_ACT) {
                        mn[0] = min(mn[0], stz); mn[1] = min(mn[1], sty); mn[2] = min(mn[2], stx);
                        mx[0] = max(mx[0], stz); mx[1] = max(mx[1], sty); mx[2] = max(mx[2], stx);
                    }
                    rec[r * (EDF_TL_PW * 32)] = make_uint4((pk & 0x3fffffffu) | (flag << 30), __float_as_uint(fz), __float_as_uint(fy), __float_as_uint(fx));
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    mn[q] = __reduce_min_sync(0xffffffffu, mn[q]);
                    mx[q] = __reduce_max_sync(0xffffffffu, mx[q]);
                }
                int* b = s.bb[par][att];
                if (lane == 0 && mn[0] != INT_MAX) {
                    atomicMin(b + 0, mn[0]); atomicMin(b + 1, mn[1]); atomicMin(b + 2, mn[2]);
                    atomicMax(b + 3, mx[0]); atomicMax(b + 4, mx[1]); atomicMax(b + 5, mx[2]);
                }
                asm volatile("bar.sync 2, %0;" :: "n"(EDF_TL_PTHREADS) : "memory");     // producers only: box complete
                box = edf_tile_box<ORDER>(b[0], b[1], b[2], b[3], b[4], b[5]);
                if (box.fit || box.empty || nb == 1 || att == 3) break;
                nb = (nb + 1) >> 1;                                // re-cut with half the rows
                nb_pref = nb;
                ++att;
            }
            // box slots of the chunk after next: last read two chunks ago, next written after the next chunk's barrier
            if (ptid < 32) (&s.bb[par == 0 ? 2 : par - 1][0][0])[ptid] = ((ptid & 7) < 3) ? INT_MAX : INT_MIN;
            if (box.fit) {
                // rows of the next chunk: two more if this box would still fit with them, two fewer if it is nearly full
                const int rows = box.nzw * box.nyal;
                const int grown = box.nzw * ((box.nyal + 3 + EDF_TL_BY - 1) & ~(EDF_TL_BY - 1));
                if (grown <= EDF_TL_ROWS - EDF_TL_ROWS / 16) nb_pref = min(EDF_TL_MR, nb + 2);
                else if (rows > EDF_TL_ROWS - EDF_TL_ROWS / 8) nb_pref = max(2, nb - 2);
                else nb_pref = nb;
            }
            const bool last = m0 + nb >= nrow;
            if (ptid == 0) {
                EdfTileMeta& mt = s.meta[stage];
                mt.m0 = m0; mt.nb = nb;
                mt.wz0 = box.wz0; mt.wy0 = box.wy0; mt.wx0 = box.wx0; mt.nzw = box.nzw; mt.nyal = box.nyal; mt.nq = box.nq;
                mt.staged = box.fit ? 1 : 0;
                mt.last = last ? 1 : 0;
            }
            __syncwarp();                                          // this warp's records (and warp 0's descriptor) before the arrival
            if (box.fit) {
                edf_tile_stage(&tmap, box, win_s + (uint32_t)stage * (WINF * 4u), full_s + (uint32_t)stage * 8u, pw, lane, lenz);
            } else if (lane == 0) {
                edf_mbar_arrive(full_s + (uint32_t)stage * 8u);
            }
            m0 += nb;
            par = par == 2 ? 0 : par + 1;
            stage ^= 1;
        }
        return;
    }

    // =================================== consumers ===================================
    const int g = warp & (EDF_TL_PW - 1), rpar = warp >> 3;        // slab of this warp; its rows: rpar, rpar + CW/8, ...
    const int z = z0 + g;
    float* __restrict__ pout = (float*)d.out;
    const float* __restrict__ pin = (const float*)d.in;
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int obase_zx = z * L.ostr_e[ii][0] + x * L.ostr_e[ii][2];   // element offsets fit 32 bits (host-checked)
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const uint32_t empty_s = (uint32_t)__cvta_generic_to_shared(&s.empty[0]);
    unsigned fphase = 0;
    int stage = 0;
    for (;;) {
        edf_tile_wait(&s.full[stage], (fphase >> stage) & 1u, 0, stage);
        fphase ^= 1u << stage;
        const EdfTileMeta& mt = s.meta[stage];
        const int m0 = mt.m0, nb = mt.nb;
        const bool last = mt.last != 0;
        EdfTileBox box;
        box.wz0 = mt.wz0; box.wy0 = mt.wy0; box.wx0 = mt.wx0; box.nzw = mt.nzw; box.nyal = mt.nyal; box.nq = mt.nq;
        const bool staged = mt.staged != 0;
        float* const win = win0 + stage * WINF;
        if (staged) {
            const bool border = (box.wy0 < 0) | (box.wy0 + box.nyal > leny) | (box.wx0 < 0) | (box.wx0 + 4 * box.nq > lenx);
            if (border) edf_tile_patch<ORDER>(win, box, leny, lenx, tid);
        }
        const uint4* const rec = rec0 + stage * RECS + g * 32 + lane;
        const int slab = box.nyal * EDF_TL_PITCH;
        const int yc0 = y0 + m0;
        const int lin0 = ((z - EDF_SW_PK_BIAS - box.wz0) * box.nyal + (yc0 - EDF_SW_PK_BIAS - box.wy0)) * EDF_TL_PITCH +
                         (x - EDF_SW_PK_BIAS - box.wx0);
        const int rxmax = 4 * box.nq - NT + EDF_SW_PK_BIAS;        // largest packed x start whose taps lie inside the window columns
#pragma unroll 1
        for (int r = rpar; r < nb; r += EDF_TL_CW / EDF_TL_PW) {
            const uint4 sv = rec[r * (EDF_TL_PW * 32)];
            unsigned flag = sv.x >> 30;
            const int rz = (int)((sv.x >> 20) & 1023u), ryw = (int)((sv.x >> 10) & 1023u), rx = (int)(sv.x & 1023u);
            // taps outside the window (a warp stretched over more than the window's 64 columns, or a chunk of which not
            // even a single row fits): those voxels gather straight from global memory
            const bool direct = flag == EDF_TL_F_ACT && (!staged || rx > rxmax - (x - box.wx0));
            const bool act = flag == EDF_TL_F_ACT && !direct;
            const int y = yc0 + r;
            if (__any_sync(0xffffffffu, act)) {
                const int off = act ? lin0 + r * EDF_TL_PITCH + (rz * box.nyal + ryw) * EDF_TL_PITCH + rx : 0;
                const float* q0 = win + off;
                float wzf[NT], wyf[NT], wxf[NT];
                edf_tile_weights<ORDER>(__uint_as_float(sv.y), wzf);
                edf_tile_weights<ORDER>(__uint_as_float(sv.z), wyf);
                edf_tile_weights<ORDER>(__uint_as_float(sv.w), wxf);
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float* qi = q0 + i * slab;
                    float ti = 0.f;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const float* rr = qi + j * EDF_TL_PITCH;
                        float tj = rr[0] * wxf[0];
#pragma unroll
                        for (int k = 1; k < NT; ++k) tj = fmaf(rr[k], wxf[k], tj);
                        ti = (j == 0) ? tj * wyf[0] : fmaf(tj, wyf[j], ti);
                    }
                    acc = (i == 0) ? ti * wzf[0] : fmaf(ti, wzf[i], acc);
                }
                if (act) pout[obase_zx + y * osy] = acc;
            }
            if (direct) {
                const int stz = z - EDF_SW_PK_BIAS + rz, sty = y - EDF_SW_PK_BIAS + ryw, stx = x - EDF_SW_PK_BIAS + rx;
                pout[obase_zx + y * osy] = edf_swin_direct_gather<ORDER>(pin, stz, sty, stx, __uint_as_float(sv.y), __uint_as_float(sv.z),
                                                                         __uint_as_float(sv.w), lenz, leny, lenx, isz, isy);
            }
            if (flag == EDF_TL_F_CST) pout[obase_zx + y * osy] = cvalf;              // deform.c:903
            if (flag == EDF_TL_F_SLOW) edf_poly_slow_voxel<ORDER, false>(p, L, ii, z, y, x);
        }
        __syncwarp();
        if (lane == 0) edf_mbar_arrive(empty_s + (uint32_t)stage * 8u);
        if (last) break;
        stage ^= 1;
    }
    (void)odz;
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
    unsigned ry = 32;
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
    const size_t smem = ((sizeof(EdfTileSmem) + 1023) & ~(size_t)1023) + (size_t)EDF_TL_STAGES * (EDF_TL_MR * EDF_TL_PW * 32 * 16 + (size_t)EDF_TL_ROWS * EDF_TL_PITCH * 4);
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
        if (cm) edf_tile3d_fwd_kernel<2, true><<<grid, EDF_TL_THREADS, smem, st>>>(p, L, tm, ii);
        else    edf_tile3d_fwd_kernel<2, false><<<grid, EDF_TL_THREADS, smem, st>>>(p, L, tm, ii);
    } else {
        if (cm) edf_tile3d_fwd_kernel<3, true><<<grid, EDF_TL_THREADS, smem, st>>>(p, L, tm, ii);
        else    edf_tile3d_fwd_kernel<3, false><<<grid, EDF_TL_THREADS, smem, st>>>(p, L, tm, ii);
    }
    return 0;
}
