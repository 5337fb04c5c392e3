// edf_swin.cuh -- forward gather through a STAGED shared-memory window (K1 for orders >= 2).
//
// Why: the direct gather (edf_lean3d_fwd_kernel) issues (order+1)^3 scalar global loads per voxel;
// the 32 windows of a warp straddle 3.4 cache lines per load instruction and the L1 replays a load
// once per line at ~2 cycles per replay, which is what bounds it (0.64 ms for 256^3 at order 3:
// 114 M wavefronts / 148 SMs x 1.8 cycles).  Shared memory serves a warp load in ONE cycle as long
// as the 32 lanes hit 32 different banks, so this kernel first copies the part of the input volume
// a chunk of output voxels can reach into shared memory and gathers from there:
//
//   CTA = 8 warps = 8 z-slabs x 32 x positions, walking along y in chunks of 4 rows (1024 voxels).
//   Per chunk: (A) every thread runs the coordinate pipeline of its 4 voxels (fp64 separable
//   B-spline of the control grid, boundary classification, window start + fractional offset; rare
//   voxels next to a rounding threshold are redone in the reference order, as in the lean kernel);
//   (B) warp REDUX + shared atomics give the EXACT bounding box of the chunk's tap windows;
//   (C) the box is copied from the volume with 16-byte cp.async (LDGSTS, L2 -> shared memory, no
//   registers; 16 lanes per window row), rows outside the volume resolved through the reference's
//   mirror map of edge taps (deform.c:791-813) at staging time, so the gather itself has no border
//   case; (D) every voxel reads its (order+1)^3 taps with LDS at immediate offsets from one base
//   address per z-tap plane.  The window rows are 64 floats apart: the bank of a tap is its x index
//   mod 32, and neighbouring lanes have neighbouring x indices whatever their y / z rows are, so
//   the loads do not collide under shear (what remains: a warp whose 32 windows span more than 32
//   columns -- a locally stretching field -- needs two wavefronts per load).
//   The window has a fixed capacity (EDF_SW_ROWS rows of 64 floats) and dynamic extents.  A chunk
//   whose box does not fit (very steep field) is gathered straight from global memory by the
//   single-voxel routine.  Between the barriers a thread keeps 4 registers per voxel (the three
//   window starts packed into one word + three fractional offsets).
//   Two CTAs per SM: one CTA's staging / coordinate phase overlaps the other's gather.
//
// Results are bit-identical to edf_lean3d_fwd_kernel (same weights, same FMA order).
#pragma once
#include "edf_lean.cuh"
#include <limits.h>

#define EDF_SW_TX 32               // x positions per warp / CTA
#define EDF_SW_G 8                 // z-slabs per CTA (one warp each)
#define EDF_SW_MR 4                // rows per chunk
#define EDF_SW_THREADS (EDF_SW_TX * EDF_SW_G)
#define EDF_SW_RY 32               // rows a CTA walks through (table capacity)
#define EDF_SW_NC 8                // control-point span capacity of the tables
#define EDF_SW_PITCH 64            // floats between window rows: multiple of 32 -> bank = x index mod 32
#ifndef EDF_SW_ROWS
#define EDF_SW_ROWS 364            // window capacity in rows; 2 CTAs x 112 KB per SM
#endif
#ifndef EDF_SW_MINBLOCKS
#define EDF_SW_MINBLOCKS 2         // resident CTAs per SM the register allocation aims at (3 needs EDF_SW_ROWS <= 208
#endif                             //   and 85 registers per thread: see DESIGN.md (f) for what that costs in spills)
#define EDF_SW_MAXQ (EDF_SW_PITCH / 4)
#define EDF_SW_DENSE_MAX 16        // active voxels of a chunk per window start of its box beyond which the gradient window is not used
#ifndef EDF_SWIN_GRAD_MAXORDER
#define EDF_SWIN_GRAD_MAXORDER 3      // highest spline order the staged-window gradient kernel takes over by default
#endif                                //   (measured on B200, 256^3: 0.24 / 0.27 / 0.47 / 0.77 ms at orders 0-3 against
                                      //   0.35 / 0.39 / 0.57 / 0.78 ms for the fixed window; 7 % slower at order 5)

static_assert(EDF_SW_MR == EDF_GW_MR && EDF_SW_NC == EDF_GW_NC, "edf_gw_coords is shared with the window gradient");

struct EdfSwinSmem {
    double wz[EDF_SW_G][4];
    double wy[EDF_SW_RY][4];
    double wx[EDF_SW_TX][4];
    int    sz[EDF_SW_G];
    int    sy[EDF_SW_RY];
    int    sx[EDF_SW_TX];
    int    ny, nx, nonzero, pad_;
    int    bb[3][8];                // [chunk % 3]: min z,y,x start, max z,y,x start of the chunk's active voxels
    int    hb[2][8];                // gradient: the same per 2-row half, only when the full box does not fit the window
    unsigned long long mbar;        // transaction barrier of the TMA bulk copies that fill the window
    double A[3][EDF_SW_G][EDF_SW_NC][EDF_SW_NC];
    double Bw[EDF_SW_G][3][EDF_SW_MR][EDF_SW_NC];
    __align__(128) float win[EDF_SW_ROWS * EDF_SW_PITCH];
};

__device__ __forceinline__ void edf_cp_async16(uint32_t smem_dst, const float* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void edf_cp_async4(uint32_t smem_dst, const float* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): one instruction moves a whole window row global -> shared and
//      reports its bytes to a transaction mbarrier.  (The tensor-map form -- UTMALDG boxes, edf_tile.cuh -- was
//      measured 17 % slower in this kernel structure; the row form needs no descriptor and suits the
//      data-dependent box.)
#ifndef EDF_SW_BULK
#define EDF_SW_BULK 1              // 1: stage the window with TMA bulk copies; 0: with 16-byte cp.async (LDGSTS)
#endif
__device__ __forceinline__ void edf_mbar_init(unsigned long long* mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void edf_mbar_expect_tx(unsigned long long* mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void edf_mbar_wait(unsigned long long* mbar, unsigned phase)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EDF_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EDF_MBAR_DONE_%=;\n"
        "bra EDF_MBAR_WAIT_%=;\n"
        "EDF_MBAR_DONE_%=:\n"
        "}\n" :: "r"(a), "r"(phase) : "memory");
}
__device__ __forceinline__ void edf_bulk_g2s(uint32_t smem_dst, const float* gsrc, unsigned bytes, unsigned long long* mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}

// exact floor / window start / fractional offset without conversion instructions for the floor:
// RD(v + 1.5 * 2^52) has floor(v) in its low mantissa bits (|v| < 2^31)
template <int ORDER>
__device__ __forceinline__ void edf_floor_split(double c, int& start, float& frac)
{
    const double M = 6755399441055744.0;
    const double v = (ORDER & 1) ? c : xadd(c, 0.5);
    const double t = __dadd_rd(v, M);
    const double fl = __dsub_rn(t, M);
    start = __double2loint(t) - ORDER / 2;
    frac = (float)xsub(c, fl);
}

// per-voxel state kept across the staging barrier: the three window starts relative to the voxel's own
// index, 10 bits each (displacements beyond +-511 voxels take the single-voxel routine)
#define EDF_SW_PK_BIAS 512
__device__ __forceinline__ bool edf_swin_pack(int dz, int dy, int dx, unsigned& pk)
{
    const unsigned a = (unsigned)(dz + EDF_SW_PK_BIAS), b = (unsigned)(dy + EDF_SW_PK_BIAS), c = (unsigned)(dx + EDF_SW_PK_BIAS);
    pk = (a << 20) | ((b & 1023u) << 10) | (c & 1023u);
    return (a < 1024u) & (b < 1024u) & (c < 1024u);
}

// Coordinate pipeline of one voxel (shared by the forward and the gradient kernel): table coordinates, range
// classification, boundary map (non-constant modes, out of line), window starts and fractional offsets.
// `slow`: the voxel sits next to a rounding / boundary threshold (or is NaN) and must be redone by the
// single-voxel routine in the reference order; `cst`: it takes the constant value (deform.c:782, :819-823).
template <int ORDER, bool CMODE>
__device__ __forceinline__ void edf_swin_voxel(const EdfParams& p, int mode, const double (*Bw)[EDF_SW_MR][EDF_SW_NC],
                                               int u, int sxrel, const double* wx, bool affine, int z, int y, int x,
                                               double bz, double bx, double offy, double limz, double limy, double limx,
                                               int lenz, int leny, int lenx, bool gate,
                                               int& stz, int& sty, int& stx, float& fz, float& fy, float& fx,
                                               bool& slow, bool& cst, bool& oob)
{
    double inz, iny, inx;
    edf_gw_coords(p, Bw, u, sxrel, wx, affine, z, y, x, bz, bx, offy, inz, iny, inx);
    const bool loz = !(inz >= 0.0), hiz = inz > limz;         // NaN counts as "low"
    const bool loy = !(iny >= 0.0), hiy = iny > limy;
    const bool lox = !(inx >= 0.0), hix = inx > limx;
    double cz = loz ? 0.0 : (hiz ? limz : inz);
    double cy = loy ? 0.0 : (hiy ? limy : iny);
    double cx = lox ? 0.0 : (hix ? limx : inx);
    const bool inr = !(loz | hiz | loy | hiy | lox | hix);
    bool mapped_danger = false, nanflag = false;
    if (!CMODE && !inr) {
        // boundary map of the out-of-range axes, out of line (deform.c:47-128)
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
    bool nearmiss = false;
    if (CMODE) {
        const double qz = fabs(xsub(inz, cz)), qy = fabs(xsub(iny, cy)), qx = fabs(xsub(inx, cx));
        nearmiss = ((qz > 0.0) & (qz < EDF_FAST_EPS)) | ((qy > 0.0) & (qy < EDF_FAST_EPS)) |
                   ((qx > 0.0) & (qx < EDF_FAST_EPS));
    }
    oob = !inr;
    cst = !inr & CMODE;
    slow = (gate & ((inr | !CMODE) ? danger : nearmiss)) | nanflag;
}

// Direct forms for the chunks whose box does not fit the window (very steep fields): the taps straight
// from / to global memory with the reference's mirror map of edge taps, same weights and FMA order as the
// window forms (and as edf_lean3d_fwd_kernel).
template <int ORDER>
__device__ __forceinline__ float edf_swin_direct_gather(const float* __restrict__ pin, int stz, int sty, int stx,
                                                        float fz, float fy, float fx, int lenz, int leny, int lenx,
                                                        int isz, int isy)
{
    constexpr int NT = ORDER + 1;
    int oz[NT], oy[NT], ox[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        oz[i] = edf_mirror1(stz + i, lenz) * isz;
        oy[i] = edf_mirror1(sty + i, leny) * isy;
        ox[i] = edf_mirror1(stx + i, lenx);
    }
    float wzf[NT], wyf[NT], wxf[NT];
    edf_bspline_weights_f32<ORDER>(fz, wzf);
    edf_bspline_weights_f32<ORDER>(fy, wyf);
    edf_bspline_weights_f32<ORDER>(fx, wxf);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        float ti = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float* r = pin + (oz[i] + oy[j]);
            float tj = __ldg(r + ox[0]) * wxf[0];
#pragma unroll
            for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + ox[k]), wxf[k], tj);
            ti = (j == 0) ? tj * wyf[0] : fmaf(tj, wyf[j], ti);
        }
        acc = (i == 0) ? ti * wzf[0] : fmaf(ti, wzf[i], acc);
    }
    return acc;
}

template <int ORDER>
__device__ __forceinline__ void edf_swin_direct_scatter(float* __restrict__ pdx, float g, int stz, int sty, int stx,
                                                        float fz, float fy, float fx, int lenz, int leny, int lenx,
                                                        int isz, int isy)
{
    constexpr int NT = ORDER + 1;
    int oz[NT], oy[NT], ox[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        oz[i] = edf_mirror1(stz + i, lenz) * isz;
        oy[i] = edf_mirror1(sty + i, leny) * isy;
        ox[i] = edf_mirror1(stx + i, lenx);
    }
    float wzf[NT], wyf[NT], wxf[NT];
    if (ORDER > 0) {
        edf_bspline_weights_f32<ORDER>(fz, wzf);
        edf_bspline_weights_f32<ORDER>(fy, wyf);
        edf_bspline_weights_f32<ORDER>(fx, wxf);
    }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const float gi = (ORDER > 0) ? g * wzf[i] : g;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float gj = (ORDER > 0) ? gi * wyf[j] : gi;
            float* r = pdx + (oz[i] + oy[j]);
#pragma unroll
            for (int k = 0; k < NT; ++k) atomicAdd(r + ox[k], (ORDER > 0) ? gj * wxf[k] : gj);
        }
    }
}

// z-contraction A of the displacement coefficients over the control points the CTA touches (prologue of both
// kernels).  Thread -> (component h, slab t, control column jx), loop over the control rows jy: no index
// arithmetic in the loop, the four z-taps of an entry are independent loads.  Same operation order per entry
// as the fast kernels (fma chain over the z-taps, starting from 0).
__device__ __forceinline__ void edf_swin_ztable(const EdfParams& p, EdfSwinSmem& s, int tid, int sy_min0, int sx_min0,
                                                int ny, int nxx)
{
    static_assert(EDF_SW_NC == 8 && 3 * EDF_SW_G * EDF_SW_NC <= EDF_SW_THREADS, "thread mapping of edf_swin_ztable");
    const int jx = tid & 7, ht = tid >> 3;
    if (ht >= 3 * EDF_SW_G || jx >= nxx) return;
    const int h = ht / EDF_SW_G, t = ht % EDF_SW_G;
    const char* base = p.disp + p.dstr[0] * h + (int64_t)edf_mirror_index32(sx_min0 + jx, (int)p.ncp[2]) * p.dstr[3];
    int64_t oz[4];
    double w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        oz[i] = (int64_t)edf_mirror_index32(s.sz[t] + i, (int)p.ncp[0]) * p.dstr[1];
        w[i] = s.wz[t][i];
    }
    const bool f64 = p.ddtype == EDF_F64;
    bool nz = false;
    for (int jy = 0; jy < ny; ++jy) {
        const char* row = base + (int64_t)edf_mirror_index32(sy_min0 + jy, (int)p.ncp[1]) * p.dstr[2];
        double cf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            cf[i] = f64 ? *(const double*)(row + oz[i]) : (double)*(const float*)(row + oz[i]);
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            nz |= (cf[i] != 0.0);
            a = fma(cf[i], w[i], a);
        }
        s.A[h][t][jy][jx] = a;
    }
    if (nz) s.nonzero = 1;
}

// warp-private y-contraction of the displacement tables for the 4 rows of chunk c (lane -> row m, control
// column jx): Bw[h][m][jx] = sum_j A[h][g][r0 + j][jx] * wy[row][j]
__device__ __forceinline__ void edf_swin_ycontract(const EdfSwinSmem& s, double (*Bw)[EDF_SW_MR][EDF_SW_NC], int g, int lane,
                                                   int c, int nx, int sy_min)
{
    static_assert(EDF_SW_MR == 4 && EDF_SW_NC == 8, "lane mapping of the y-contraction");
    const int m = lane & 3, jx = lane >> 2;
    if (jx < nx) {
        const int row = c * EDF_SW_MR + m;
        const int r0 = s.sy[row] - sy_min;
        double wyr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wyr[j] = s.wy[row][j];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            double b = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) b = fma(s.A[h][g][r0 + j][jx], wyr[j], b);
            Bw[h][m][jx] = b;
        }
    }
}

// tile of this CTA from the 1-D block index (EdfTileSched)
__device__ __forceinline__ void edf_swin_tile(const EdfTileSched& T, int& x0, int& y0, int& z0, int& ry)
{
    unsigned bid = blockIdx.x;
    unsigned sgm = 0;
    while (sgm + 1 < T.nseg && bid >= T.cta_begin[sgm + 1]) ++sgm;
    bid -= T.cta_begin[sgm];
    ry = (int)T.ry[sgm];
    const unsigned tx = bid % T.gx, t = bid / T.gx;
    const unsigned gy = T.gy[sgm];
    x0 = (int)tx * EDF_SW_TX;
    y0 = (int)(t % gy) * ry;
    z0 = (int)(t / gy + T.z_begin[sgm]) * EDF_SW_G;
}

// CMODE: boundary mode 'constant' (out-of-range voxels take cval: no coordinate map, and no call inside
// the coordinate phase, which keeps its register allocation free of call-crossing live ranges)
template <int ORDER, bool CMODE>
__global__ void __launch_bounds__(EDF_SW_THREADS, EDF_SW_MINBLOCKS)
edf_swin3d_fwd_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    EdfSwinSmem& s = *reinterpret_cast<EdfSwinSmem*>(smem_raw);
    constexpr int NT = ORDER + 1;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;     // warp = slab
    int x0, y0, z0, ry;
    edf_swin_tile(L.sched, x0, y0, z0, ry);

    // ---- prologue: control tables, z-contraction A of the displacement coefficients, empty boxes
    if (tid == 0) {
        s.nonzero = 0;
        if (EDF_SW_BULK) edf_mbar_init(&s.mbar, 1);
    }
    if (tid < 3 * 8) (&s.bb[0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : ((tid & 7) == 7 ? 0 : INT_MIN);
    if (tid < EDF_SW_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_SW_TX + EDF_SW_RY) {
        const int t = tid - EDF_SW_TX;
        edf_fast_ctrl_entry(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1), s.wy[t], &s.sy[t]);
    } else if (tid < EDF_SW_TX + EDF_SW_RY + EDF_SW_G) {
        const int t = tid - EDF_SW_TX - EDF_SW_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    }
    __syncthreads();
    {
        const int sy_min0 = s.sy[0], sx_min0 = s.sx[0];
        const int ny = s.sy[ry - 1] - sy_min0 + 4;                  // control rows the CTA's ry rows touch
        const int nxx = s.sx[EDF_SW_TX - 1] - sx_min0 + 4;
        if (tid == 0) { s.ny = ny; s.nx = nxx; }
        edf_swin_ztable(p, s, tid, sy_min0, sx_min0, ny, nxx);
    }
    __syncthreads();

    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[lane][k];
    const int sxrel = s.sx[lane] - s.sx[0];
    const int nchunk = min(ry / EDF_SW_MR, (ody - y0 + EDF_SW_MR - 1) / EDF_SW_MR);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_SW_MR][EDF_SW_NC] = s.Bw[g];

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
    const double offy = p.ooff_d[1];
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(s.win);
#if !EDF_SW_BULK
    // staging role of this thread: 16 lanes per window row, one 16-byte group each
    const int sq = tid & 15, srow = tid >> 4;
#endif

    int par = 0;                                                   // c % 3
    unsigned mphase = 0;                                           // parity of the window barrier's current phase
    for (int c = 0; c < nchunk; ++c) {
        const int yc0 = y0 + c * EDF_SW_MR;
        edf_swin_ycontract(s, Bw, g, lane, c, nx, sy_min);
        __syncwarp();

        // ---- phase A: coordinates and classification of this thread's 4 voxels (branch-free, two at a time)
        unsigned pk[EDF_SW_MR];
        float fz[EDF_SW_MR], fy[EDF_SW_MR], fx[EDF_SW_MR];
        unsigned actm = 0, cstm = 0, slowm = 0;
        int mnz = INT_MAX, mny = INT_MAX, mnx = INT_MAX, mxz = INT_MIN, mxy = INT_MIN, mxx = INT_MIN;
#pragma unroll
        for (int u = 0; u < EDF_SW_MR; ++u) {
            const int y = yc0 + u;
            const bool valid = tok && (y < ody);
            int stz, sty, stx;
            bool slow, cst, oob;
            edf_swin_voxel<ORDER, CMODE>(p, d.mode, Bw, u, sxrel, wx, affine, z, y, x, bz, bx, offy, limz, limy, limx,
                                         lenz, leny, lenx, gate, stz, sty, stx, fz[u], fy[u], fx[u], slow, cst, oob);
            const bool packed = edf_swin_pack(stz - z, sty - y, stx - x, pk[u]);
            slow = valid & (slow | (!cst & !packed));
            if (slow) slowm |= 1u << u;
            if (valid & !slow & cst) cstm |= 1u << u;
            if (valid & !slow & !cst) {
                actm |= 1u << u;
                mnz = min(mnz, stz); mny = min(mny, sty); mnx = min(mnx, stx);
                mxz = max(mxz, stz); mxy = max(mxy, sty); mxx = max(mxx, stx);
            }
        }

        // ---- phase B: exact bounding box of the chunk's tap windows
        mnz = __reduce_min_sync(0xffffffffu, mnz);
        mny = __reduce_min_sync(0xffffffffu, mny);
        mnx = __reduce_min_sync(0xffffffffu, mnx);
        mxz = __reduce_max_sync(0xffffffffu, mxz);
        mxy = __reduce_max_sync(0xffffffffu, mxy);
        mxx = __reduce_max_sync(0xffffffffu, mxx);
        if (lane == 0 && mnz != INT_MAX) {
            int* b = s.bb[par];
            atomicMin(b + 0, mnz); atomicMin(b + 1, mny); atomicMin(b + 2, mnx);
            atomicMax(b + 3, mxz); atomicMax(b + 4, mxy); atomicMax(b + 5, mxx);
        }
        __syncthreads();                                           // box complete; previous chunk's gathers done
        // the box of chunk c+2 (= chunk c-1's, no longer read) is reset here: chunk c+2's atomics come after
        // the next chunk's barrier
        if (tid < 8) s.bb[par == 0 ? 2 : par - 1][tid] = (tid < 3) ? INT_MAX : (tid == 7 ? 0 : INT_MIN);
        const int wz0 = s.bb[par][0], wy0 = s.bb[par][1], wx0 = s.bb[par][2] & ~3;
        const int nzw = s.bb[par][3] - wz0 + NT, nyw = s.bb[par][4] - wy0 + NT;
        const int nq = ((s.bb[par][5] + NT - 1 - wx0) >> 2) + 1;
        const bool empty = s.bb[par][0] > s.bb[par][3];
        const bool fit = !empty && nq <= EDF_SW_MAXQ && nzw <= EDF_SW_ROWS && nyw <= EDF_SW_ROWS && nzw * nyw <= EDF_SW_ROWS;
        if (fit) {
            // ---- phase C: copy the box into the window.  Rows / planes outside the volume are the mirror images
            //      the reference's edge taps read (deform.c:796-810); 16-byte groups left or right of the volume
            //      (lenx % 4 == 0: inside or outside as a whole) likewise, element by element.
            const int rows = nzw * nyw;
#if EDF_SW_BULK
            {
                // one bulk copy per window row (the part of it inside the volume along x), issued by one thread
                // each; thread 0 announces the byte total to the transaction barrier, every thread then waits
                // for the phase to complete.  Groups left / right of the volume: element-wise, mirrored.
                const int xs = max(wx0, 0), xe = min(wx0 + 4 * nq, lenx);
                const unsigned rbytes = (unsigned)(xe - xs) * 4u;
                if (tid == 0) edf_mbar_expect_tx(&s.mbar, rbytes * (unsigned)rows);
                const unsigned inv_y = 0xffffffffu / (unsigned)nyw + 1u;                    // nyw >= 3
                for (int r = tid; r < rows; r += EDF_SW_THREADS) {
                    const int zr = (int)__umulhi((unsigned)r, inv_y);
                    const int yr = r - zr * nyw;
                    const int gz = edf_mirror1(wz0 + zr, lenz), gy = edf_mirror1(wy0 + yr, leny);
                    edf_bulk_g2s(win_s + (uint32_t)(r * EDF_SW_PITCH + (xs - wx0)) * 4u, pin + (gz * isz + gy * isy + xs), rbytes, &s.mbar);
                }
                const bool xborder = (wx0 < 0) | (wx0 + 4 * nq > lenx);                     // CTA-uniform
                if (xborder) {
                    // at most one group on either side (the window reaches at most `order` cells past the border)
                    const int side = tid & 1;                                              // 0: left group, 1: right group
                    const int gx = side ? lenx : -4;
                    const bool need = side ? (wx0 + 4 * nq > lenx) : (wx0 < 0);
                    if (need) {
                        const int m0 = edf_mirror1(gx, lenx), m1 = edf_mirror1(gx + 1, lenx);
                        const int m2 = edf_mirror1(gx + 2, lenx), m3 = edf_mirror1(gx + 3, lenx);
                        for (int r = tid >> 1; r < rows; r += EDF_SW_THREADS / 2) {
                            const int zr = (int)__umulhi((unsigned)r, inv_y);
                            const int yr = r - zr * nyw;
                            const float* src = pin + (edf_mirror1(wz0 + zr, lenz) * isz + edf_mirror1(wy0 + yr, leny) * isy);
                            const uint32_t dst = win_s + (uint32_t)(r * EDF_SW_PITCH + (gx - wx0)) * 4u;
                            edf_cp_async4(dst, src + m0);
                            edf_cp_async4(dst + 4, src + m1);
                            edf_cp_async4(dst + 8, src + m2);
                            edf_cp_async4(dst + 12, src + m3);
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncthreads();
                }
                edf_mbar_wait(&s.mbar, mphase);
                mphase ^= 1u;
            }
#else
            if (sq < nq) {
                const int gx = wx0 + 4 * sq;
                const bool xin = (unsigned)gx < (unsigned)lenx;
                int r = srow;
                int zr = (int)__umulhi((unsigned)r, 0xffffffffu / (unsigned)nyw + 1u);      // r / nyw (nyw >= 2)
                int yr = r - zr * nyw;
                uint32_t dst = win_s + (uint32_t)(r * EDF_SW_PITCH + 4 * sq) * 4u;
                if (xin) {
                    const float* srcx = pin + gx;
                    for (; r < rows; r += EDF_SW_THREADS / 16) {
                        const int gz = edf_mirror1(wz0 + zr, lenz), gy = edf_mirror1(wy0 + yr, leny);
                        edf_cp_async16(dst, srcx + (gz * isz + gy * isy));
                        dst += (EDF_SW_THREADS / 16) * EDF_SW_PITCH * 4;
                        yr += EDF_SW_THREADS / 16;
                        while (yr >= nyw) { yr -= nyw; ++zr; }
                    }
                } else {
                    const int m0 = edf_mirror1(gx, lenx), m1 = edf_mirror1(gx + 1, lenx);
                    const int m2 = edf_mirror1(gx + 2, lenx), m3 = edf_mirror1(gx + 3, lenx);
                    for (; r < rows; r += EDF_SW_THREADS / 16) {
                        const int gz = edf_mirror1(wz0 + zr, lenz), gy = edf_mirror1(wy0 + yr, leny);
                        const float* src = pin + (gz * isz + gy * isy);
                        edf_cp_async4(dst, src + m0);
                        edf_cp_async4(dst + 4, src + m1);
                        edf_cp_async4(dst + 8, src + m2);
                        edf_cp_async4(dst + 12, src + m3);
                        dst += (EDF_SW_THREADS / 16) * EDF_SW_PITCH * 4;
                        yr += EDF_SW_THREADS / 16;
                        while (yr >= nyw) { yr -= nyw; ++zr; }
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
#endif
            // ---- phase D: gather from the window (inactive lanes read cell 0 and discard)
            const int slab = nyw * EDF_SW_PITCH;
            const int lin0 = ((z - EDF_SW_PK_BIAS - wz0) * nyw + (yc0 - EDF_SW_PK_BIAS - wy0)) * EDF_SW_PITCH +
                             (x - EDF_SW_PK_BIAS - wx0);
#pragma unroll
            for (int u = 0; u < EDF_SW_MR; ++u) {
                const bool a = (actm >> u) & 1u;
                if (!__any_sync(0xffffffffu, a)) continue;
                const int rz = (int)(pk[u] >> 20), ryw = (int)((pk[u] >> 10) & 1023u), rx = (int)(pk[u] & 1023u);
                const int off = a ? lin0 + u * EDF_SW_PITCH + (rz * nyw + ryw) * EDF_SW_PITCH + rx : 0;
                const float* b0 = s.win + off;
                float wzf[NT], wyf[NT], wxf[NT];
                edf_bspline_weights_f32<ORDER>(fz[u], wzf);
                edf_bspline_weights_f32<ORDER>(fy[u], wyf);
                edf_bspline_weights_f32<ORDER>(fx[u], wxf);
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float* bi = b0 + i * slab;
                    float ti = 0.f;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const float* r = bi + j * EDF_SW_PITCH;
                        float tj = r[0] * wxf[0];
#pragma unroll
                        for (int k = 1; k < NT; ++k) tj = fmaf(r[k], wxf[k], tj);
                        ti = (j == 0) ? tj * wyf[0] : fmaf(tj, wyf[j], ti);
                    }
                    acc = (i == 0) ? ti * wzf[0] : fmaf(ti, wzf[i], acc);
                }
                if (a) pout[obase_zx + (yc0 + u) * osy] = acc;
            }
        } else if (!empty) {
            // the chunk's tap windows span more than the window holds (very steep field): straight from global memory
#pragma unroll
            for (int u = 0; u < EDF_SW_MR; ++u) {
                if (!((actm >> u) & 1u)) continue;
                const int stz = z - EDF_SW_PK_BIAS + (int)(pk[u] >> 20), sty = yc0 + u - EDF_SW_PK_BIAS + (int)((pk[u] >> 10) & 1023u);
                const int stx = x - EDF_SW_PK_BIAS + (int)(pk[u] & 1023u);
                pout[obase_zx + (yc0 + u) * osy] =
                    edf_swin_direct_gather<ORDER>(pin, stz, sty, stx, fz[u], fy[u], fx[u], lenz, leny, lenx, isz, isy);
            }
        }
#pragma unroll
        for (int u = 0; u < EDF_SW_MR; ++u)
            if ((cstm >> u) & 1u) pout[obase_zx + (yc0 + u) * osy] = cvalf;            // deform.c:903
        // ---- rare voxels (next to a rounding / boundary threshold, huge displacements, chunks that do not fit
        //      the window): the single-voxel routine, from the same table coordinates.  The only call of the loop
        //      body sits here, where no per-chunk state is live.  (Moving this block, the constant stores and the
        //      next chunk's y-contraction between the issue of the bulk copies and the wait for them -- to hide
        //      the copy latency -- measured 5-12 % SLOWER, 0.518 vs 0.495 ms at order 3: the state kept live
        //      across the call costs more than the wait.)
        if (slowm) {
#pragma unroll 1
            for (int u = 0; u < EDF_SW_MR; ++u) {
                if (!((slowm >> u) & 1u)) continue;
                double inz, iny, inx;
                edf_gw_coords(p, Bw, u, sxrel, wx, affine, z, yc0 + u, x, bz, bx, offy, inz, iny, inx);
                edf_lean_forward_slow<ORDER>(p, L, ii, z, yc0 + u, x, inz, iny, inx, gate);
            }
        }
        __syncwarp();                                              // lanes in the rare-voxel loop still read this chunk's Bw
        par = par == 2 ? 0 : par + 1;
    }
}

// =======================================================================================
// Gradient scatter through the same window (K2): the adjoint of the kernel above.
//
// Same chunks, same coordinate phase and the same exact bounding box, but instead of being filled from
// the volume the window starts out zero, every active voxel ADDS its (order+1)^3 contributions
// dY * wz * wy * wx to it with native shared-memory integer atomics (32-bit fixed point: float atomics
// are CAS loops on sm_100; the largest possible contribution of the chunk maps to just under 2^22, so
// a cell holds 512 of them and the resolution is max|dY_chunk| * wmax^3 * 2^-22), and the box is then added to dX
// once with 16-byte vector atomics (RED.128), row by row, and re-zeroed.  Against edf_lean3d_gradwin_kernel
// (fixed 19x23x52 window): rows are 64 cells apart, so the ATOMS of a warp collide only where the field
// stretches (2.8 -> ~1.6 wavefronts each); the box is exact, so no voxel misses the window and only the
// box -- not the whole window -- is flushed; taps across the border of the volume accumulate in virtual
// cells whose row / column is folded back by the reference's mirror map (deform.c:796-810) at flush time.
// =======================================================================================
template <int ORDER, bool CMODE>
__global__ void __launch_bounds__(EDF_SW_THREADS, EDF_SW_MINBLOCKS)
edf_swin3d_grad_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    EdfSwinSmem& s = *reinterpret_cast<EdfSwinSmem*>(smem_raw);
    constexpr int NT = ORDER + 1;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;     // warp = slab
    int x0, y0, z0, ry;
    edf_swin_tile(L.sched, x0, y0, z0, ry);
    int* const win = reinterpret_cast<int*>(s.win);

    // ---- prologue: control tables, z-contraction A, empty boxes, zero window
    if (tid == 0) s.nonzero = 0;
    if (tid < 3 * 8) (&s.bb[0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : ((tid & 7) == 7 ? 0 : INT_MIN);
    if (tid >= 32 && tid < 48) (&s.hb[0][0])[tid - 32] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
    if (tid < EDF_SW_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_SW_TX + EDF_SW_RY) {
        const int t = tid - EDF_SW_TX;
        edf_fast_ctrl_entry(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1), s.wy[t], &s.sy[t]);
    } else if (tid < EDF_SW_TX + EDF_SW_RY + EDF_SW_G) {
        const int t = tid - EDF_SW_TX - EDF_SW_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    }
    for (int e = tid; e < EDF_SW_ROWS * EDF_SW_PITCH / 4; e += EDF_SW_THREADS)
        reinterpret_cast<int4*>(win)[e] = make_int4(0, 0, 0, 0);
    __syncthreads();
    {
        const int sy_min0 = s.sy[0], sx_min0 = s.sx[0];
        const int ny = s.sy[ry - 1] - sy_min0 + 4;                  // control rows the CTA's ry rows touch
        const int nxx = s.sx[EDF_SW_TX - 1] - sx_min0 + 4;
        if (tid == 0) { s.ny = ny; s.nx = nxx; }
        edf_swin_ztable(p, s, tid, sy_min0, sx_min0, ny, nxx);
    }
    __syncthreads();

    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[lane][k];
    const int sxrel = s.sx[lane] - s.sx[0];
    const int nchunk = min(ry / EDF_SW_MR, (ody - y0 + EDF_SW_MR - 1) / EDF_SW_MR);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_SW_MR][EDF_SW_NC] = s.Bw[g];

    const EdfInputDesc& d = p.inp[ii];
    float* __restrict__ pdx = (float*)d.in;                       // dX accumulator
    const float* __restrict__ pdy = (const float*)d.out;          // upstream gradient dY
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int obase_zx = z * L.ostr_e[ii][0] + x * L.ostr_e[ii][2];
    const bool affine = p.has_affine != 0;
    const double bz = xadd((double)z, p.ooff_d[0]);
    const double bx = xadd((double)x, p.ooff_d[2]);
    const double offy = p.ooff_d[1];
    const int sq = tid & 15, srow = tid >> 4;                     // flush role: 16 lanes per window row

    // dY of the next chunk is loaded one chunk ahead (its latency hides behind the scatter / flush phases)
    float gvn[EDF_SW_MR];
#pragma unroll
    for (int u = 0; u < EDF_SW_MR; ++u)
        gvn[u] = (tok && nchunk > 0 && y0 + u < ody) ? __ldg(pdy + (obase_zx + (y0 + u) * osy)) : 0.f;

    int par = 0;
    for (int c = 0; c < nchunk; ++c) {
        const int yc0 = y0 + c * EDF_SW_MR;
        edf_swin_ycontract(s, Bw, g, lane, c, nx, sy_min);
        __syncwarp();

        // ---- phase A: dY, coordinates and classification of this thread's 4 voxels
        unsigned pk[EDF_SW_MR];
        float fz[EDF_SW_MR], fy[EDF_SW_MR], fx[EDF_SW_MR], gv[EDF_SW_MR];
        unsigned actm = 0, slowm = 0, dirm = 0;
        int mnz = INT_MAX, mny = INT_MAX, mnx = INT_MAX, mxz = INT_MIN, mxy = INT_MIN, mxx = INT_MIN;
        float gmax = 0.f;
#pragma unroll
        for (int u = 0; u < EDF_SW_MR; ++u) {
            const int y = yc0 + u;
            const bool valid = tok && (y < ody);
            gv[u] = gvn[u];                                       // 0 for voxels outside the output
            const bool live = valid & (gv[u] != 0.f);             // zero gradients contribute nothing
            int stz, sty, stx;
            bool slow, cst, oob;
            edf_swin_voxel<ORDER, CMODE>(p, d.mode, Bw, u, sxrel, wx, affine, z, y, x, bz, bx, offy, limz, limy, limx,
                                         lenz, leny, lenx, gate, stz, sty, stx, fz[u], fy[u], fx[u], slow, cst, oob);
            const bool packed = edf_swin_pack(stz - z, sty - y, stx - x, pk[u]);
            slow = live & (slow | (!cst & !packed));
            if (slow) slowm |= 1u << u;
            // Voxels that a non-constant boundary mode folded back into the volume pile up on the border cells
            // (all of a chunk's out-of-range voxels can land on ONE cell in 'nearest' mode), which the 32-bit
            // fixed-point cells of the window cannot hold: they scatter straight to dX in float.
            if (!CMODE && live & !slow & oob) dirm |= 1u << u;
            else if (live & !slow & !cst) {                       // constant voxels pass no gradient (deform.c:928)
                actm |= 1u << u;
                mnz = min(mnz, stz); mny = min(mny, sty); mnx = min(mnx, stx);
                mxz = max(mxz, stz); mxy = max(mxy, sty); mxx = max(mxx, stx);
                gmax = fmaxf(gmax, fabsf(gv[u]));
            }
        }

        // ---- phase B: exact bounding box of the chunk's tap windows, max |dY| of the active voxels
        mnz = __reduce_min_sync(0xffffffffu, mnz);
        mny = __reduce_min_sync(0xffffffffu, mny);
        mnx = __reduce_min_sync(0xffffffffu, mnx);
        mxz = __reduce_max_sync(0xffffffffu, mxz);
        mxy = __reduce_max_sync(0xffffffffu, mxy);
        mxx = __reduce_max_sync(0xffffffffu, mxx);
        const int gbits = __reduce_max_sync(0xffffffffu, __float_as_int(gmax));    // non-negative floats order as ints
        // Density guard: the 32-bit fixed-point cells hold ~128 (orders 0 / 1) to ~150 (orders >= 2) voxels' worth of
        // mass.  Where the map collapses (a fold of the displacement field; affine magnification is excluded on the
        // host, edf_fast_try_launch) that many voxels of one chunk can land on one cell and the accumulator would
        // wrap.  The chunk's active voxels are counted against the window starts its box can hold; a chunk with
        // more than EDF_SW_DENSE_MAX voxels per start cell on average scatters straight to dX in float instead.
#ifndef EDF_SW_NO_GUARD
        const int nact = __reduce_add_sync(0xffffffffu, __popc(actm));
#else
        const int nact = 0;
#endif
        if (lane == 0 && mnz != INT_MAX) {
            int* b = s.bb[par];
            atomicMin(b + 0, mnz); atomicMin(b + 1, mny); atomicMin(b + 2, mnx);
            atomicMax(b + 3, mxz); atomicMax(b + 4, mxy); atomicMax(b + 5, mxx);
            atomicMax(b + 6, gbits);
            atomicAdd(b + 7, nact);
        }
        __syncthreads();                                           // box complete; previous chunk's flush done
        if (tid < 8) s.bb[par == 0 ? 2 : par - 1][tid] = (tid < 3) ? INT_MAX : (tid == 7 ? 0 : INT_MIN);
#pragma unroll
        for (int u = 0; u < EDF_SW_MR; ++u) {
            const int yn = yc0 + EDF_SW_MR + u;
            gvn[u] = (tok && c + 1 < nchunk && yn < ody) ? __ldg(pdy + (obase_zx + yn * osy)) : 0.f;
        }
        const bool empty = s.bb[par][0] > s.bb[par][3];
        const long long startcells = (long long)(s.bb[par][3] - s.bb[par][0] + 1) * (s.bb[par][4] - s.bb[par][1] + 1) * (s.bb[par][5] - s.bb[par][2] + 1);
        const bool dense = !empty && (long long)s.bb[par][7] > EDF_SW_DENSE_MAX * startcells;
        const bool usewin = !empty && !dense && !(L.input_mask >> 31);
        bool fitfull;
        {
            const int nzw_ = s.bb[par][3] - s.bb[par][0] + NT, nyw_ = s.bb[par][4] - s.bb[par][1] + NT;
            const int nq_ = ((s.bb[par][5] + NT - 1 - (s.bb[par][2] & ~3)) >> 2) + 1;
            fitfull = nq_ <= EDF_SW_MAXQ && nzw_ <= EDF_SW_ROWS && nyw_ <= EDF_SW_ROWS && nzw_ * nyw_ <= EDF_SW_ROWS;
        }
        // A chunk whose box outgrows the window (a stretch of steep field: ~5 % of the chunks of the headline volume) is
        // scattered as two 2-row halves, each through its own box, before anything falls back to float atomics on dX
        // (which cost ~5x the window path per voxel).
        const int npass = (usewin && !fitfull) ? 2 : 1;
        if (npass == 2) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                int hn0 = INT_MAX, hn1 = INT_MAX, hn2 = INT_MAX, hx0 = INT_MIN, hx1 = INT_MIN, hx2 = INT_MIN;
#pragma unroll
                for (int u = 2 * hf; u < 2 * hf + 2; ++u)
                    if ((actm >> u) & 1u) {
                        const int stz = z - EDF_SW_PK_BIAS + (int)(pk[u] >> 20), sty = yc0 + u - EDF_SW_PK_BIAS + (int)((pk[u] >> 10) & 1023u);
                        const int stx = x - EDF_SW_PK_BIAS + (int)(pk[u] & 1023u);
                        hn0 = min(hn0, stz); hn1 = min(hn1, sty); hn2 = min(hn2, stx);
                        hx0 = max(hx0, stz); hx1 = max(hx1, sty); hx2 = max(hx2, stx);
                    }
                hn0 = __reduce_min_sync(0xffffffffu, hn0); hn1 = __reduce_min_sync(0xffffffffu, hn1); hn2 = __reduce_min_sync(0xffffffffu, hn2);
                hx0 = __reduce_max_sync(0xffffffffu, hx0); hx1 = __reduce_max_sync(0xffffffffu, hx1); hx2 = __reduce_max_sync(0xffffffffu, hx2);
                if (lane == 0 && hn0 != INT_MAX) {
                    int* hbp = s.hb[hf];
                    atomicMin(hbp + 0, hn0); atomicMin(hbp + 1, hn1); atomicMin(hbp + 2, hn2);
                    atomicMax(hbp + 3, hx0); atomicMax(hbp + 4, hx1); atomicMax(hbp + 5, hx2);
                }
            }
            __syncthreads();
        }
        for (int ps = 0; ps < npass; ++ps) {
        const int* bx_ = (npass == 2) ? s.hb[ps] : s.bb[par];
        const unsigned rowmask = (npass == 2) ? (ps ? 0xcu : 0x3u) : 0xfu;
        const int wz0 = bx_[0], wy0 = bx_[1], wx0 = bx_[2] & ~3;
        const int nzw = bx_[3] - wz0 + NT, nyw = bx_[4] - wy0 + NT;
        const int nq = ((bx_[5] + NT - 1 - wx0) >> 2) + 1;
        const bool pempty = bx_[0] > bx_[3];
        const bool fit = usewin && !pempty && nq <= EDF_SW_MAXQ && nzw <= EDF_SW_ROWS && nyw <= EDF_SW_ROWS && nzw * nyw <= EDF_SW_ROWS;
        if (ps) __syncthreads();                                   // the first half's flush has re-zeroed the window
        if (fit) {
            // fixed-point scale: the largest possible contribution, max|dY| of the chunk times the largest weight
            // product of this order, maps to just under 2^22 -- the range of the magic-number rounding that
            // orders >= 2 use (FFMA + IADD; F2I runs at a quarter of the rate) -- so a cell holds 512 such
            // contributions; resolution max|dY_chunk| * wmax^3 * 2^-22 (7e-8 max|dY| at order 3).
            const float gmax_c = __int_as_float(s.bb[par][6]);
            constexpr float WMAX = (ORDER <= 1) ? 1.0f : (ORDER == 2) ? 0.75f : (ORDER == 3) ? (2.0f / 3.0f)
                                 : (ORDER == 4) ? (115.0f / 192.0f) : 0.55f;
            // (orders 0 / 1 convert with F2I and have no such range limit: 2^24, 128 contributions per cell)
            const float scale = ((ORDER <= 1 ? 16777216.0f : 4194304.0f) * 0.999f) /
                                (fmaxf(gmax_c, 1e-30f) * (WMAX * WMAX * WMAX));
            const float inv_scale = 1.0f / scale;
            // ---- phase C: scatter into the window
            const int slab = nyw * EDF_SW_PITCH;
            const int lin0 = ((z - EDF_SW_PK_BIAS - wz0) * nyw + (yc0 - EDF_SW_PK_BIAS - wy0)) * EDF_SW_PITCH +
                             (x - EDF_SW_PK_BIAS - wx0);
#pragma unroll
            for (int u = 0; u < EDF_SW_MR; ++u) {
                if (!(((actm & rowmask) >> u) & 1u)) continue;
                const int rz = (int)(pk[u] >> 20), ryw = (int)((pk[u] >> 10) & 1023u), rx = (int)(pk[u] & 1023u);
                int* b0 = win + (lin0 + u * EDF_SW_PITCH + (rz * nyw + ryw) * EDF_SW_PITCH + rx);
                float wzf[NT], wyf[NT], wxf[NT];
                if (ORDER > 0) {
                    edf_bspline_weights_f32<ORDER>(fz[u], wzf);
                    edf_bspline_weights_f32<ORDER>(fy[u], wyf);
                    edf_bspline_weights_f32<ORDER>(fx[u], wxf);
                }
                const float gs = gv[u] * scale;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float gi = (ORDER > 0) ? gs * wzf[i] : gs;
                    int* bi = b0 + i * slab;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const float gj = (ORDER > 0) ? gi * wyf[j] : gi;
                        int* r = bi + j * EDF_SW_PITCH;
#pragma unroll
                        for (int k = 0; k < NT; ++k) {
                            if (ORDER >= 2)      atomicAdd(r + k, edf_gw_round(gj, wxf[k]));
                            else if (ORDER == 1) atomicAdd(r + k, __float2int_rn(gj * wxf[k]));
                            else                 atomicAdd(r + k, __float2int_rn(gj));
                        }
                    }
                }
            }
            __syncthreads();
            // ---- phase D: add the box to dX (each touched 16-byte group once) and re-zero it.  Rows / planes
            //      outside the volume fold back through the mirror map, groups left / right of it element-wise.
            if (sq < nq) {
                // thread -> (row yr of every plane, 16-byte group sq); the planes are the inner loop.  (Walking the
                // window rows in memory order instead -- every thread ~rows/16 (plane, row) pairs, all 16 row groups
                // busy whatever the box's shape -- measured 8-15 % slower: 0.751 vs 0.695 ms at order 3.)
                const int gx = wx0 + 4 * sq;
                const int pstep = nyw * EDF_SW_PITCH / 4;                                  // int4 units between planes
                const bool interior = (wz0 >= 0) & (wz0 + nzw <= lenz) & (wy0 >= 0) & (wy0 + nyw <= leny) &
                                      (wx0 >= 0) & (wx0 + 4 * nq <= lenx);                 // CTA-uniform
                const bool xin = (unsigned)gx < (unsigned)lenx;
                for (int yr = srow; yr < nyw; yr += EDF_SW_THREADS / 16) {
                    int4* src = reinterpret_cast<int4*>(win + (yr * EDF_SW_PITCH + 4 * sq));
                    if (interior) {
                        // the box lies inside the volume: no mirror map
                        float* dst = pdx + (wz0 * isz + (wy0 + yr) * isy + gx);
                        for (int zr = 0; zr < nzw; ++zr, src += pstep, dst += isz) {
                            const int4 v = *src;
                            if ((v.x | v.y | v.z | v.w) != 0) {
                                *src = make_int4(0, 0, 0, 0);
                                atomicAdd(reinterpret_cast<float4*>(dst),
                                          make_float4((float)v.x * inv_scale, (float)v.y * inv_scale,
                                                      (float)v.z * inv_scale, (float)v.w * inv_scale));
                            }
                        }
                    } else {
                        const int gy = edf_mirror1(wy0 + yr, leny);
                        for (int zr = 0; zr < nzw; ++zr, src += pstep) {
                            const int4 v = *src;
                            if ((v.x | v.y | v.z | v.w) != 0) {
                                *src = make_int4(0, 0, 0, 0);
                                float* dst = pdx + (edf_mirror1(wz0 + zr, lenz) * isz + gy * isy);
                                const float4 f = make_float4((float)v.x * inv_scale, (float)v.y * inv_scale,
                                                             (float)v.z * inv_scale, (float)v.w * inv_scale);
                                if (xin) {
                                    atomicAdd(reinterpret_cast<float4*>(dst + gx), f);
                                } else {
                                    if (v.x) atomicAdd(dst + edf_mirror1(gx, lenx), f.x);
                                    if (v.y) atomicAdd(dst + edf_mirror1(gx + 1, lenx), f.y);
                                    if (v.z) atomicAdd(dst + edf_mirror1(gx + 2, lenx), f.z);
                                    if (v.w) atomicAdd(dst + edf_mirror1(gx + 3, lenx), f.w);
                                }
                            }
                        }
                    }
                }
            }
        } else if (!pempty) {
            dirm |= actm & rowmask;                                // does not fit the window (very steep field)
        }
        }
        if (npass == 2) {
            __syncthreads();                                       // half boxes read by every thread: reset for the next user
            if (tid < 16) (&s.hb[0][0])[tid] = ((tid & 7) < 3) ? INT_MAX : INT_MIN;
        }
        // ---- voxels that bypass the window: direct global atomics in float
        if (dirm) {
#pragma unroll
            for (int u = 0; u < EDF_SW_MR; ++u) {
                if (!((dirm >> u) & 1u)) continue;
                const int stz = z - EDF_SW_PK_BIAS + (int)(pk[u] >> 20), sty = yc0 + u - EDF_SW_PK_BIAS + (int)((pk[u] >> 10) & 1023u);
                const int stx = x - EDF_SW_PK_BIAS + (int)(pk[u] & 1023u);
                edf_swin_direct_scatter<ORDER>(pdx, gv[u], stz, sty, stx, fz[u], fy[u], fx[u], lenz, leny, lenx, isz, isy);
            }
        }
        // ---- rare voxels: the single-voxel routine (reference-order coordinates, global atomics)
        if (slowm) {
#pragma unroll 1
            for (int u = 0; u < EDF_SW_MR; ++u) {
                if (!((slowm >> u) & 1u)) continue;
                double inz, iny, inx;
                edf_gw_coords(p, Bw, u, sxrel, wx, affine, z, yc0 + u, x, bz, bx, offy, inz, iny, inx);
                edf_gradwin_slow_voxel<ORDER>(p, L, ii, z, yc0 + u, x, inz, iny, inx, gate);
            }
        }
        __syncwarp();                                              // lanes in the rare-voxel loop still read this chunk's Bw
        par = par == 2 ? 0 : par + 1;
    }
}

static EdfPerDeviceFlag g_swin_configured;

// shared conditions of the two staged-window kernels
static bool edf_swin_common_ok(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    if (!edf_lean_eligible(p, L, ii)) return false;
    const EdfInputDesc& d = p.inp[ii];
    if (d.mode == EDF_MODE_WRAP) return false;         // wrapped voxels land on the far side: no compact box
    if (p.idim[2] % 4) return false;                   // 16-byte groups are inside or outside the volume as a whole
    if (((uintptr_t)d.in % 16) || (L.istr_e[ii][0] % 4) || (L.istr_e[ii][1] % 4)) return false;
    if (!edf_fast_ctrl_span_ok(p, 2, EDF_SW_TX, EDF_SW_NC)) return false;
    if (!edf_fast_ctrl_span_ok(p, 1, EDF_SW_RY, EDF_SW_NC)) return false;
    return true;
}

static bool edf_swin_fwd_env()                          // EDF_STAGED_FWD=1: staged forward gather for every eligible call (A/B runs)
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("EDF_STAGED_FWD"); v = (e && *e && *e != '0') ? 1 : 0; }
    return v != 0;
}

#ifndef EDF_SWIN_FWD_MAXORDER
#define EDF_SWIN_FWD_MAXORDER 3       // staged forward gather by default at orders 2 .. this (B200, 256^3, sigma 8:
#endif                                //   0.36 / 0.54 ms at orders 2 / 3 against 0.37 / 0.66 ms direct; slower at order 5)
static int edf_swin_max_fwd_order()
{
    static int v = -2;                                  // EDF_SWIN_FWD_MAXORDER=-1: direct gather everywhere (A/B runs)
    if (v < -1) { const char* e = getenv("EDF_SWIN_FWD_MAXORDER"); v = (e && *e) ? atoi(e) : EDF_SWIN_FWD_MAXORDER; }
    return v;
}

// forward: orders 2-5 are implemented; which of them take this kernel by default is decided in edf_fast_try_launch
static bool edf_swin_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    const EdfInputDesc& d = p.inp[ii];
    if (d.order < 2 || d.order > 5) return false;
    return edf_swin_common_ok(p, L, ii);
}

static int edf_swin_max_grad_order()
{
    static int v = -2;                                  // EDF_SWIN_GRAD_MAXORDER=-1 disables the kernel (A/B runs)
    if (v < -1) { const char* e = getenv("EDF_SWIN_GRAD_MAXORDER"); v = (e && *e) ? atoi(e) : EDF_SWIN_GRAD_MAXORDER; }
    return v;
}

static bool edf_swin_grad_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii, bool all_orders)
{
    const EdfInputDesc& d = p.inp[ii];
    if (d.order < 0 || d.order > (all_orders ? 5 : edf_swin_max_grad_order())) return false;
    return edf_swin_common_ok(p, L, ii);
}

#ifndef EDF_SWIN_TAIL
#define EDF_SWIN_TAIL 1            // 1: graded tail (the last z-tiles run with ry/2 and ry/4 rows per CTA); 0: uniform grid
#endif
// Measured on B200 (256^3 float32, 5^3 grid, CUDA events, same box; rows per CTA / graded tail):
//   gradient, order 3, constant:   32 rows 0.778 ms | 16 rows 0.744 | 32 rows + tail 0.694   (orders 1 / 2: 1-4 % faster with the tail)
//   forward,  order 3, constant:   32 rows 0.542 ms | 16 rows 0.516 | 32 rows + tail 0.566   (order 2: 32 rows best, 0.358 vs 0.376)
//   forward / gradient, order 3, nearest (no cheap out-of-range CTAs at the end of the grid): the tail saves 9 % / 8-11 %
// -> the gradient always takes the graded tail; the forward gather takes it in the non-constant modes, and in
//    'constant' mode (where the last z-tiles are cheap anyway) runs 16 rows per CTA at orders >= 3.
static bool edf_swin_grid(const EdfParams& p, int order, bool gradient, bool cmode, unsigned& ncta, EdfTileSched& T)
{
    const uint64_t gx = (uint64_t)((p.odim[2] + EDF_SW_TX - 1) / EDF_SW_TX);
    const uint64_t gz = (uint64_t)((p.odim[0] + EDF_SW_G - 1) / EDF_SW_G);
    unsigned ry = (!gradient && cmode && order >= 3) ? EDF_SW_RY / 2 : EDF_SW_RY;
    const bool want_tail = gradient || !cmode;
    static int env_ry = -1, env_tail = -2;              // EDF_SWIN_ROWS=4/8/16/32: rows per CTA; EDF_SWIN_TAIL=n: z-tiles in the
    if (env_ry < 0) { const char* e = getenv("EDF_SWIN_ROWS"); env_ry = (e && *e) ? atoi(e) : 0; }     // graded tail, 0 = none
    if (env_tail < -1) { const char* e = getenv("EDF_SWIN_TAIL"); env_tail = (e && *e) ? atoi(e) : -1; }  // (A/B runs)
    if (env_ry == 4 || env_ry == 8 || env_ry == 16 || env_ry == 32) ry = (unsigned)env_ry;
    while (ry > EDF_SW_MR && gx * ((p.odim[1] + ry - 1) / ry) * gz < 4ull * 148) ry >>= 1;
    const uint64_t per_z = gx * (uint64_t)((p.odim[1] + ry - 1) / ry);       // CTAs of one z-tile in the main segment
    // graded tail: about one wave (2 CTAs on each of 148 SMs) of main-segment CTAs is replaced by CTAs of a
    // half and a quarter of the rows, so that the last CTAs to finish are short.  Only for grids of several
    // waves, and never more than a third of the volume.
    uint64_t nt = 0;
    if (((EDF_SWIN_TAIL && want_tail) || env_tail > 0) && env_tail != 0 && ry >= 2 * EDF_SW_MR && per_z * gz >= 3ull * 296) {
        nt = env_tail > 0 ? (uint64_t)env_tail : (296 + per_z - 1) / per_z;
        if (nt > gz / 3) nt = gz / 3;
    }
    const uint64_t nt2 = (ry >= 4 * EDF_SW_MR) ? nt / 2 : 0;                 // z-tiles at a quarter of the rows
    const uint64_t nt1 = nt - nt2;                                           // z-tiles at half the rows
    T.gx = (unsigned)gx;
    uint64_t cta = 0, zt = 0;
    unsigned sgm = 0;
    const uint64_t zcount[3] = {gz - nt, nt1, nt2};
    const unsigned rys[3] = {ry, ry / 2, ry / 4};
    for (int k = 0; k < 3; ++k) {
        if (!zcount[k]) continue;
        const uint64_t gy = (uint64_t)((p.odim[1] + rys[k] - 1) / rys[k]);
        T.cta_begin[sgm] = (unsigned)cta;
        T.z_begin[sgm] = (unsigned)zt;
        T.ry[sgm] = rys[k];
        T.gy[sgm] = (unsigned)gy;
        cta += gx * gy * zcount[k];
        zt += zcount[k];
        ++sgm;
    }
    T.nseg = sgm;
    for (unsigned k = sgm; k < 4; ++k) { T.cta_begin[k] = (unsigned)cta; T.z_begin[k] = (unsigned)zt; }
    ncta = (unsigned)cta;
    return cta > 0 && cta < (1ull << 31);
}

// returns 0 = launched, -2 = not applicable (caller takes the direct kernel), -1 = CUDA error
static int edf_swin_launch(int order, int gradient, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    EdfFastLaunch L = Lin;
    unsigned grid;
    if (!edf_swin_grid(p, order, gradient != 0, p.inp[ii].mode == EDF_MODE_CONSTANT, grid, L.sched)) return -2;
    L.rows_per_cta = L.sched.ry[0];
    {
        static int dbg = -1;                            // debug: EDF_SWIN_DEBUG_UNFIT=1 sends every chunk of the gradient
        if (dbg < 0) { const char* e = getenv("EDF_SWIN_DEBUG_UNFIT"); dbg = (e && *e && *e != '0') ? 1 : 0; }   // kernel down
        if (dbg) L.input_mask |= 0x80000000u;           // the "does not fit the window" path
    }
    const size_t smem = sizeof(EdfSwinSmem);
    if (!g_swin_configured.test()) {
#define EDF_SW_ATTR(O)                                                                                              \
    cudaFuncSetAttribute(edf_swin3d_fwd_kernel<O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    cudaFuncSetAttribute(edf_swin3d_fwd_kernel<O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        EDF_SW_ATTR(2); EDF_SW_ATTR(3); EDF_SW_ATTR(4); EDF_SW_ATTR(5);
#undef EDF_SW_ATTR
#define EDF_SW_ATTR(O)                                                                                              \
    cudaFuncSetAttribute(edf_swin3d_grad_kernel<O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    cudaFuncSetAttribute(edf_swin3d_grad_kernel<O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        EDF_SW_ATTR(0); EDF_SW_ATTR(1); EDF_SW_ATTR(2); EDF_SW_ATTR(3); EDF_SW_ATTR(4); EDF_SW_ATTR(5);
#undef EDF_SW_ATTR
        if (cudaGetLastError() != cudaSuccess) return -1;
        g_swin_configured.set();
    }
    const bool cm = p.inp[ii].mode == EDF_MODE_CONSTANT;
#define EDF_SW_CASE(K, O)                                                                \
    if (cm) K<O, true><<<grid, EDF_SW_THREADS, smem, st>>>(p, L, ii);                     \
    else    K<O, false><<<grid, EDF_SW_THREADS, smem, st>>>(p, L, ii);                    \
    break;
    if (!gradient) {
        switch (order) {
        case 2: EDF_SW_CASE(edf_swin3d_fwd_kernel, 2)
        case 3: EDF_SW_CASE(edf_swin3d_fwd_kernel, 3)
        case 4: EDF_SW_CASE(edf_swin3d_fwd_kernel, 4)
        default: EDF_SW_CASE(edf_swin3d_fwd_kernel, 5)
        }
    } else {
        switch (order) {
        case 0: EDF_SW_CASE(edf_swin3d_grad_kernel, 0)
        case 1: EDF_SW_CASE(edf_swin3d_grad_kernel, 1)
        case 2: EDF_SW_CASE(edf_swin3d_grad_kernel, 2)
        case 3: EDF_SW_CASE(edf_swin3d_grad_kernel, 3)
        case 4: EDF_SW_CASE(edf_swin3d_grad_kernel, 4)
        default: EDF_SW_CASE(edf_swin3d_grad_kernel, 5)
        }
    }
#undef EDF_SW_CASE
    return 0;
}
