// edf_lean.cuh -- straight-line 3-D float32 kernels for the headline configuration:
// one float32 volume, unit-stride last axis, no non-deformed axes, any order/mode.
//
// Same tile geometry and coordinate pipeline as edf_fast.cuh (tables A/B in shared memory, 12
// fp64 FMAs per voxel), but the common case -- a voxel whose whole tap window lies inside the
// volume and whose coordinates are not next to a rounding threshold -- runs branch-free:
//   3 x (floor, frac, start)  ->  fp32 weights  ->  16 row pointers by 64-bit adds  ->
//   (order+1)^3 loads with immediate offsets  ->  fp32 FMAs  ->  one coalesced store.
// Everything else (edges, out-of-range coordinates in a non-constant mode, voxels within 1e-6 of
// a threshold, which are re-evaluated in the reference order) goes through the out-of-line
// general routine of edf_fast.cuh, so results are identical to the general fast kernel.
#pragma once
#include "edf_fast.cuh"
#include <stdlib.h>

#define EDF_LEAN_EPSF 1e-6f        // float-side threshold test; superset of EDF_FAST_EPS (2e-8)

// shared-memory tables of the lean kernels: like EdfFastSmem, but the y-contraction B is
// private to each warp (a warp owns one slab g and 32 x positions), so the main loop needs no
// CTA-wide barrier and warps with little work (outside the volume) do not hold the others up.
struct EdfLeanSmem {
    double wz[EDF_FAST_G][4];
    double wy[EDF_FAST_RY][4];
    double wx[EDF_FAST_TX][4];
    int    sz[EDF_FAST_G];
    int    sy[EDF_FAST_RY];
    int    sx[EDF_FAST_TX];
    int    ny, nx, nonzero, pad_;
    double A[3][EDF_FAST_G][EDF_FAST_NC][EDF_FAST_NC];
    double Bw[EDF_FAST_THREADS / 32][3][EDF_FAST_M][EDF_FAST_NC];
};

__device__ __forceinline__ void edf_lean_tile_setup(const EdfParams& p, EdfLeanSmem& s, int z0, int y0, int x0)
{
    const int tid = threadIdx.x;
    if (tid == 0) s.nonzero = 0;
    if (tid < EDF_FAST_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_FAST_TX + EDF_FAST_RY) {
        const int t = tid - EDF_FAST_TX;
        edf_fast_ctrl_entry(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1), s.wy[t], &s.sy[t]);
    } else if (tid < EDF_FAST_TX + EDF_FAST_RY + EDF_FAST_G) {
        const int t = tid - EDF_FAST_TX - EDF_FAST_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    }
    __syncthreads();
    const int sy_min = s.sy[0], sx_min = s.sx[0];
    const int ny = s.sy[EDF_FAST_RY - 1] - sy_min + 4;
    const int nx = s.sx[EDF_FAST_TX - 1] - sx_min + 4;
    if (tid == 0) { s.ny = ny; s.nx = nx; }
    bool nz = false;
    const int na = 3 * EDF_FAST_G * ny * nx;
    for (int e = tid; e < na; e += EDF_FAST_THREADS) {
        const int jx = e % nx;
        const int jy = (e / nx) % ny;
        const int t = (e / (nx * ny)) % EDF_FAST_G;
        const int h = e / (nx * ny * EDF_FAST_G);
        const int my = edf_mirror_index32(sy_min + jy, (int)p.ncp[1]);
        const int mx = edf_mirror_index32(sx_min + jx, (int)p.ncp[2]);
        const char* base = p.disp + p.dstr[0] * h + my * p.dstr[2] + mx * p.dstr[3];
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int mz = edf_mirror_index32(s.sz[t] + i, (int)p.ncp[0]);
            const double c = (p.ddtype == EDF_F64) ? *(const double*)(base + mz * p.dstr[1])
                                                   : (double)*(const float*)(base + mz * p.dstr[1]);
            nz |= (c != 0.0);
            a = fma(c, s.wz[t][i], a);
        }
        s.A[h][t][jy][jx] = a;
    }
    if (nz) s.nonzero = 1;
    __syncthreads();
}

// exact reference-order re-evaluation of one voxel's un-mapped coordinates (cold)
__device__ __noinline__ void edf_lean_exact_coords(const EdfParams& p, int z, int y, int x,
                                                   double* inz, double* iny, double* inx)
{
    int o[3] = {z, y, x};
    double dd[3];
    edf_displacement_exact_cold<3>(p, o, dd);
    *inz = edf_source_coordinate<3, int>(p, o, 0, dd[0]);
    *iny = edf_source_coordinate<3, int>(p, o, 1, dd[1]);
    *inx = edf_source_coordinate<3, int>(p, o, 2, dd[2]);
}

// One axis of one voxel: boundary map (only out-of-range coordinates leave the inline path),
// floor / fractional offset / window start, and the "next to a threshold" tests.
// Returns true when the voxel takes the constant value (deform.c:782, :819-823).
template <int ORDER>
__device__ __forceinline__ bool edf_lean_axis(const EdfParams& p, int h, int mode, double in, double lim,
                                              bool gate, int& start, float& frac, bool& danger)
{
    double cc = in;
    if (!((in >= 0.0) & (in <= lim))) {
        if (mode == EDF_MODE_CONSTANT) {
            // misses the volume: constant, unless it misses by less than the re-evaluation threshold
            const double q = fabs(xsub(in, fmin(fmax(in, 0.0), lim)));
            danger |= gate & (q > 0.0) & (q < EDF_FAST_EPS);
            return true;
        }
        danger |= gate & edf_near_half_integer(in);
        cc = edf_map_coordinate_cold(in, p.idim[h], mode);
        if (!(cc > -1.0)) return true;
    }
    const double fl = (ORDER & 1) ? floor(cc) : floor(xadd(cc, 0.5));
    frac = (float)xsub(cc, fl);
    start = (int)fl - ORDER / 2;
    // float-side test, a superset of |2*in - rint(2*in)| < 2*EDF_FAST_EPS: odd orders have their
    // thresholds at the integers, even orders at the half-integers (and the boundary tests at the
    // integers 0 and len-1)
    if (ORDER & 1) danger |= gate & ((frac < EDF_LEAN_EPSF) | (frac > 1.0f - EDF_LEAN_EPSF));
    else           danger |= gate & ((fabsf(frac) < EDF_LEAN_EPSF) | (fabsf(frac) > 0.5f - EDF_LEAN_EPSF));
    return false;
}

template <int ORDER, bool GRAD>
__global__ void __launch_bounds__(EDF_FAST_THREADS, 2)
edf_lean3d_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    __shared__ EdfLeanSmem s;
    constexpr int NT = ORDER + 1;
    const int x0 = blockIdx.x * EDF_FAST_TX;
    const int ry = (int)L.rows_per_cta;
    const int y0 = blockIdx.y * ry;
    const int z0 = blockIdx.z * EDF_FAST_G;
    edf_lean_tile_setup(p, s, z0, y0, x0);

    const int tx = threadIdx.x & (EDF_FAST_TX - 1);
    const int g = threadIdx.x >> 6;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = x0 + tx, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[tx][k];
    const int sxrel = s.sx[tx] - s.sx[0];
    const int nchunk = min(ry / EDF_FAST_M, (ody - y0 + EDF_FAST_M - 1) / EDF_FAST_M);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_FAST_M][EDF_FAST_NC] = s.Bw[warp];

    // per-thread constants of this input
    const EdfInputDesc& d = p.inp[ii];
    float* __restrict__ pin = (float*)d.in;                     // forward: read; gradient: accumulated
    float* __restrict__ pout = (float*)d.out;                   // forward: written; gradient: dY
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int64_t obase_zx = (int64_t)z * L.ostr_e[ii][0] + (int64_t)x * L.ostr_e[ii][2];
    const bool affine = p.has_affine != 0;
    const int mode = d.mode;
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const double bz = xadd((double)z, p.ooff_d[0]);             // deform.c:778-781: (o + offset) + displacement
    const double bx = xadd((double)x, p.ooff_d[2]);
    const double offy = p.ooff_d[1];

    for (int c = 0; c < nchunk; ++c) {
        // ---- y-contraction of this warp's slab for the 8 rows of the chunk (warp-private)
        for (int e = lane; e < 3 * EDF_FAST_M * nx; e += 32) {
            const int jx = e % nx;
            const int m = (e / nx) % EDF_FAST_M;
            const int h = e / (nx * EDF_FAST_M);
            const int row = c * EDF_FAST_M + m;
            const int r0 = s.sy[row] - sy_min;
            double b = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) b = fma(s.A[h][g][r0 + j][jx], s.wy[row][j], b);
            Bw[h][m][jx] = b;
        }
        __syncwarp();
        if (tok) {
#pragma unroll 1
            for (int m = 0; m < EDF_FAST_M; ++m) {
                const int y = y0 + c * EDF_FAST_M + m;
                if (y >= ody) break;
                // ---- displacement: x-contraction of the tile tables (12 fp64 FMAs)
                double dz = 0.0, dy = 0.0, dx = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    dz = fma(Bw[0][m][sxrel + k], wx[k], dz);
                    dy = fma(Bw[1][m][sxrel + k], wx[k], dy);
                    dx = fma(Bw[2][m][sxrel + k], wx[k], dx);
                }
                double inz, iny, inx;
                if (!affine) {
                    inz = xadd(bz, dz);
                    iny = xadd(xadd((double)y, offy), dy);
                    inx = xadd(bx, dx);
                } else {
                    const int o[3] = {z, y, x};
                    inz = edf_source_coordinate<3, int>(p, o, 0, dz);
                    iny = edf_source_coordinate<3, int>(p, o, 1, dy);
                    inx = edf_source_coordinate<3, int>(p, o, 2, dx);
                }
                // ---- boundary map, window start, fractional offsets; voxels next to a rounding /
                //      boundary threshold are re-evaluated once in the exact reference order
                int stz = 0, sty = 0, stx = 0;
                float fz = 0.f, fy = 0.f, fx = 0.f;
                bool constant, danger;
#pragma unroll 1
                for (int pass = 0;; ++pass) {
                    danger = false;
                    constant = edf_lean_axis<ORDER>(p, 0, mode, inz, limz, gate, stz, fz, danger);
                    if (!constant) constant = edf_lean_axis<ORDER>(p, 1, mode, iny, limy, gate, sty, fy, danger);
                    if (!constant) constant = edf_lean_axis<ORDER>(p, 2, mode, inx, limx, gate, stx, fx, danger);
                    if (!danger || pass) break;
                    edf_lean_exact_coords(p, z, y, x, &inz, &iny, &inx);
                }
                float* po = pout + (obase_zx + (int64_t)y * osy);
                if (constant) {
                    if (!GRAD) *po = cvalf;                                 // deform.c:903
                    continue;
                }
                // ---- tap rows: element offsets along z and y with the reference's mirror mapping at
                //      the edges (deform.c:791-813); the x taps are consecutive unless the x window
                //      itself crosses the border somewhere in the warp
                int oz[NT], oy[NT];
                {
                    const bool ez = (stz < 0) | (stz + ORDER >= lenz);
                    const bool ey = (sty < 0) | (sty + ORDER >= leny);
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        int iz = stz + i, iy = sty + i;
                        if (ez) iz = edf_mirror_index32(iz, lenz);
                        if (ey) iy = edf_mirror_index32(iy, leny);
                        oz[i] = iz * isz;
                        oy[i] = iy * isy;
                    }
                }
                const bool ex = (stx < 0) | (stx + ORDER >= lenx);
                float wz[NT], wy[NT], wxf[NT];
                if (ORDER > 0) {
                    edf_bspline_weights_f32<ORDER>(fz, wz);
                    edf_bspline_weights_f32<ORDER>(fy, wy);
                    edf_bspline_weights_f32<ORDER>(fx, wxf);
                }
                const bool warp_ex = __any_sync(__activemask(), ex);
                if (!GRAD) {
                    float t = 0.f;
                    if (!warp_ex) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            float ti = 0.f;
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const float* r = pin + (oz[i] + oy[j] + stx);
                                float tj = (ORDER > 0) ? __ldg(r) * wxf[0] : __ldg(r);
#pragma unroll
                                for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + k), wxf[k], tj);
                                if (ORDER > 0) ti = (j == 0) ? tj * wy[0] : fmaf(tj, wy[j], ti);
                                else ti = tj;
                            }
                            if (ORDER > 0) t = (i == 0) ? ti * wz[0] : fmaf(ti, wz[i], t);
                            else t = ti;
                        }
                    } else {
                        int oxk[NT];
#pragma unroll
                        for (int k = 0; k < NT; ++k) oxk[k] = ex ? edf_mirror_index32(stx + k, lenx) : stx + k;
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            float ti = 0.f;
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const float* r = pin + (oz[i] + oy[j]);
                                float tj = (ORDER > 0) ? __ldg(r + oxk[0]) * wxf[0] : __ldg(r + oxk[0]);
#pragma unroll
                                for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + oxk[k]), wxf[k], tj);
                                if (ORDER > 0) ti = (j == 0) ? tj * wy[0] : fmaf(tj, wy[j], ti);
                                else ti = tj;
                            }
                            if (ORDER > 0) t = (i == 0) ? ti * wz[0] : fmaf(ti, wz[i], t);
                            else t = ti;
                        }
                    }
                    *po = t;
                } else {
                    const float gval = *po;
                    int oxk[NT];
#pragma unroll
                    for (int k = 0; k < NT; ++k) oxk[k] = ex ? edf_mirror_index32(stx + k, lenx) : stx + k;
#pragma unroll
                    for (int i = 0; i < NT; ++i) {
                        const float gi = (ORDER > 0) ? gval * wz[i] : gval;
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            const float gj = (ORDER > 0) ? gi * wy[j] : gi;
                            float* r = pin + (oz[i] + oy[j]);
#pragma unroll
                            for (int k = 0; k < NT; ++k)
                                atomicAdd(r + oxk[k], (ORDER > 0) ? gj * wxf[k] : gj);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// =======================================================================================
// Forward kernel, batched: U rows of a thread are processed together, phase by phase
// (coordinates -> classification -> offsets / weights -> all loads -> FMAs -> stores), written
// branch-free so that the U independent dependency chains interleave.  The per-voxel latency
// chain (LDS -> 4 DFMA -> floor -> convert -> address -> LDG -> FFMA -> STG) is ~10x longer than
// its issue time; with 16 warps per SM one voxel at a time left the issue slots 57-66 % idle.
// Rare voxels (next to a rounding threshold, or out of range in a non-constant mode) are flagged
// and redone afterwards by the single-voxel routine, which overwrites the output.
// Mirror mapping of edge taps uses the single-reflection form, valid for extents >= 8.
// =======================================================================================
template <int ORDER>
__device__ __noinline__ void edf_lean_forward_slow(const EdfParams& p, const EdfFastLaunch& L, int ii,
                                                   int z, int y, int x, double inz, double iny, double inx,
                                                   bool gate)
{
    int o[3] = {z, y, x};
    double in[3] = {inz, iny, inx};
    if (gate && (edf_near_half_integer(in[0]) || edf_near_half_integer(in[1]) || edf_near_half_integer(in[2]))) {
        double dd[3];
        edf_displacement_exact_cold<3>(p, o, dd);
#pragma unroll
        for (int h = 0; h < 3; ++h) in[h] = edf_source_coordinate<3, int>(p, o, h, dd[h]);
    }
    edf_fast_f32_one_input<3, ORDER, false>(p, L, ii, o, in);
}

// single-reflection mirror map (deform.c:796-810 for indices within one period of the border)
__device__ __forceinline__ int edf_mirror1(int idx, int len)
{
    idx = idx < 0 ? -idx : idx;
    return idx >= len ? 2 * len - 2 - idx : idx;
}

#ifndef EDF_LEAN_FWD_MINB
#define EDF_LEAN_FWD_MINB 2        // CTAs per SM the batched forward kernel is compiled for
#endif
#ifndef EDF_LEAN_FWD_U3
#define EDF_LEAN_FWD_U3 2          // rows per iteration at orders 2 and 3
#endif
template <int ORDER, int U>
__global__ void __launch_bounds__(EDF_FAST_THREADS, EDF_LEAN_FWD_MINB)
edf_lean3d_fwd_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    __shared__ EdfLeanSmem s;
    constexpr int NT = ORDER + 1;
    const int x0 = blockIdx.x * EDF_FAST_TX;
    const int ry = (int)L.rows_per_cta;
    const int y0 = blockIdx.y * ry;
    const int z0 = blockIdx.z * EDF_FAST_G;
    edf_lean_tile_setup(p, s, z0, y0, x0);

    const int tx = threadIdx.x & (EDF_FAST_TX - 1);
    const int g = threadIdx.x >> 6;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = x0 + tx, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[tx][k];
    const int sxrel = s.sx[tx] - s.sx[0];
    const int nchunk = min(ry / EDF_FAST_M, (ody - y0 + EDF_FAST_M - 1) / EDF_FAST_M);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_FAST_M][EDF_FAST_NC] = s.Bw[warp];

    const EdfInputDesc& d = p.inp[ii];
    const float* __restrict__ pin = (const float*)d.in;
    float* __restrict__ pout = (float*)d.out;
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int64_t obase_zx = (int64_t)z * L.ostr_e[ii][0] + (int64_t)x * L.ostr_e[ii][2];
    const bool affine = p.has_affine != 0;
    const bool cmode = d.mode == EDF_MODE_CONSTANT;
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const double bz = xadd((double)z, p.ooff_d[0]);
    const double bx = xadd((double)x, p.ooff_d[2]);
    const double offy = p.ooff_d[1];

    for (int c = 0; c < nchunk; ++c) {
        {
            // lane -> (row m, column phase q): no integer division in the table loop
            static_assert(EDF_FAST_M == 8, "lane mapping of the B table assumes 8 rows per chunk");
            const int m = lane & 7, q = lane >> 3;
            const int row = c * EDF_FAST_M + m;
            const int r0 = s.sy[row] - sy_min;
            double wyr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wyr[j] = s.wy[row][j];
            // the 3 * nx (component, column) pairs go round the 4 phases; nx >= 4, so one wrap per step
            for (int h = 0, jx = q; h < 3;) {
                double b = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) b = fma(s.A[h][g][r0 + j][jx], wyr[j], b);
                Bw[h][m][jx] = b;
                jx += 4;
                if (jx >= nx) { jx -= nx; ++h; }
            }
        }
        __syncwarp();
        if (tok) {
#pragma unroll 1
            for (int m0 = 0; m0 < EDF_FAST_M; m0 += U) {
                const int yb = y0 + c * EDF_FAST_M + m0;
                if (yb >= ody) break;
                double inz[U], iny[U], inx[U];
                int stz[U], sty[U], stx[U];
                float fz[U], fy[U], fx[U];
                bool valid[U], cst[U], slow[U];
                bool any_ex = false;
                // ---- phase 1: coordinates and classification (branch-free)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int m = m0 + u;
                    const int y = yb + u;
                    valid[u] = y < ody;
                    double dz = 0.0, dy = 0.0, dx = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        dz = fma(Bw[0][m][sxrel + k], wx[k], dz);
                        dy = fma(Bw[1][m][sxrel + k], wx[k], dy);
                        dx = fma(Bw[2][m][sxrel + k], wx[k], dx);
                    }
                    if (!affine) {
                        inz[u] = xadd(bz, dz);
                        iny[u] = xadd(xadd((double)y, offy), dy);
                        inx[u] = xadd(bx, dx);
                    } else {
                        const int o[3] = {z, y, x};
                        inz[u] = edf_source_coordinate<3, int>(p, o, 0, dz);
                        iny[u] = edf_source_coordinate<3, int>(p, o, 1, dy);
                        inx[u] = edf_source_coordinate<3, int>(p, o, 2, dx);
                    }
                    // clamp into the volume: out-of-range voxels get a harmless in-range address
                    // (comparisons + selects: fmin/fmax on doubles cost three times as many instructions)
                    const bool loz = !(inz[u] >= 0.0), hiz = inz[u] > limz;         // NaN counts as "low"
                    const bool loy = !(iny[u] >= 0.0), hiy = iny[u] > limy;
                    const bool lox = !(inx[u] >= 0.0), hix = inx[u] > limx;
                    double cz = loz ? 0.0 : (hiz ? limz : inz[u]);
                    double cy = loy ? 0.0 : (hiy ? limy : iny[u]);
                    double cx = lox ? 0.0 : (hix ? limx : inx[u]);
                    const bool inr = !(loz | hiz | loy | hiy | lox | hix);
                    bool mapped_danger = false, nanflag = false;
                    if (!cmode && !inr) {
                        // boundary map of the out-of-range axes, out of line (deform.c:47-128); the mapped
                        // coordinate is in [0, len-1] (reflect: possibly in (-1, 0), handled by the mirror taps)
                        if (loz | hiz) { mapped_danger |= edf_near_half_integer(inz[u]); cz = edf_map_coordinate_cold(inz[u], lenz, d.mode); }
                        if (loy | hiy) { mapped_danger |= edf_near_half_integer(iny[u]); cy = edf_map_coordinate_cold(iny[u], leny, d.mode); }
                        if (lox | hix) { mapped_danger |= edf_near_half_integer(inx[u]); cx = edf_map_coordinate_cold(inx[u], lenx, d.mode); }
                        if (!((cz > -1.0) & (cy > -1.0) & (cx > -1.0))) { nanflag = true; cz = cy = cx = 0.0; }   // NaN
                    }
                    const double flz = (ORDER & 1) ? floor(cz) : floor(xadd(cz, 0.5));
                    const double fly = (ORDER & 1) ? floor(cy) : floor(xadd(cy, 0.5));
                    const double flx = (ORDER & 1) ? floor(cx) : floor(xadd(cx, 0.5));
                    fz[u] = (float)xsub(cz, flz);
                    fy[u] = (float)xsub(cy, fly);
                    fx[u] = (float)xsub(cx, flx);
                    stz[u] = (int)flz - ORDER / 2;
                    sty[u] = (int)fly - ORDER / 2;
                    stx[u] = (int)flx - ORDER / 2;
                    bool danger;
                    if (ORDER & 1)
                        danger = (fz[u] < EDF_LEAN_EPSF) | (fz[u] > 1.0f - EDF_LEAN_EPSF) | (fy[u] < EDF_LEAN_EPSF) |
                                 (fy[u] > 1.0f - EDF_LEAN_EPSF) | (fx[u] < EDF_LEAN_EPSF) | (fx[u] > 1.0f - EDF_LEAN_EPSF);
                    else
                        danger = (fabsf(fz[u]) < EDF_LEAN_EPSF) | (fabsf(fz[u]) > 0.5f - EDF_LEAN_EPSF) |
                                 (fabsf(fy[u]) < EDF_LEAN_EPSF) | (fabsf(fy[u]) > 0.5f - EDF_LEAN_EPSF) |
                                 (fabsf(fx[u]) < EDF_LEAN_EPSF) | (fabsf(fx[u]) > 0.5f - EDF_LEAN_EPSF);
                    danger |= mapped_danger;
                    // in range (or mapped into range): only thresholds matter.  Out of range in 'constant' mode:
                    // cval, unless the voxel misses the volume by less than the re-evaluation threshold.
                    bool nearmiss = false;
                    if (cmode) {
                        const double qz = fabs(xsub(inz[u], cz)), qy = fabs(xsub(iny[u], cy)), qx = fabs(xsub(inx[u], cx));
                        nearmiss = ((qz > 0.0) & (qz < EDF_FAST_EPS)) | ((qy > 0.0) & (qy < EDF_FAST_EPS)) |
                                   ((qx > 0.0) & (qx < EDF_FAST_EPS));
                    }
                    slow[u] = valid[u] & ((gate & ((inr | !cmode) ? danger : nearmiss)) | nanflag);
                    cst[u] = valid[u] & !inr & cmode & !slow[u];
                    any_ex |= (stx[u] < 0) | (stx[u] + ORDER >= lenx) | (sty[u] < 0) | (sty[u] + ORDER >= leny) |
                              (stz[u] < 0) | (stz[u] + ORDER >= lenz);
                }
                // ---- phase 2: tap offsets (single-reflection mirror at the edges) and weights
                float t[U];
                const bool warp_ex = __any_sync(__activemask(), any_ex);
                if (ORDER == 0) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        t[u] = __ldg(pin + (edf_mirror1(stz[u], lenz) * isz + edf_mirror1(sty[u], leny) * isy +
                                            edf_mirror1(stx[u], lenx)));
                } else {
                    float wz[U][NT], wy[U][NT], wxf[U][NT];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        edf_bspline_weights_f32<ORDER>(fz[u], wz[u]);
                        edf_bspline_weights_f32<ORDER>(fy[u], wy[u]);
                        edf_bspline_weights_f32<ORDER>(fx[u], wxf[u]);
                    }
                    if (!warp_ex) {
                        // every window of the warp lies inside the volume: one base pointer per voxel, rows at
                        // warp-uniform distances (i * isz + j * isy), x taps at immediate offsets
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const float* b0 = pin + (stz[u] * isz + sty[u] * isy + stx[u]);
                            float acc = 0.f;
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                float ti = 0.f;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float* r = b0 + (i * isz + j * isy);
                                    float tj = __ldg(r) * wxf[u][0];
#pragma unroll
                                    for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + k), wxf[u][k], tj);
                                    ti = (j == 0) ? tj * wy[u][0] : fmaf(tj, wy[u][j], ti);
                                }
                                acc = (i == 0) ? ti * wz[u][0] : fmaf(ti, wz[u][i], acc);
                            }
                            t[u] = acc;
                        }
                    } else {
                        // some window of the warp crosses the border of the volume: mirrored taps on all axes
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            int oz[NT], oy[NT], oxk[NT];
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                oz[i] = edf_mirror1(stz[u] + i, lenz) * isz;
                                oy[i] = edf_mirror1(sty[u] + i, leny) * isy;
                                oxk[i] = edf_mirror1(stx[u] + i, lenx);
                            }
                            float acc = 0.f;
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                float ti = 0.f;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float* r = pin + (oz[i] + oy[j]);
                                    float tj = __ldg(r + oxk[0]) * wxf[u][0];
#pragma unroll
                                    for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + oxk[k]), wxf[u][k], tj);
                                    ti = (j == 0) ? tj * wy[u][0] : fmaf(tj, wy[u][j], ti);
                                }
                                acc = (i == 0) ? ti * wz[u][0] : fmaf(ti, wz[u][i], acc);
                            }
                            t[u] = acc;
                        }
                    }
                }
                // ---- phase 3: stores; flagged voxels are redone by the single-voxel routine
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (valid[u]) {
                        float* po = pout + (obase_zx + (int64_t)(yb + u) * osy);
                        *po = cst[u] ? cvalf : t[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (slow[u]) edf_lean_forward_slow<ORDER>(p, L, ii, z, yb + u, x, inz[u], iny[u], inx[u], gate);
            }
        }
        __syncwarp();
    }
}

static bool edf_lean_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    if (p.naxis != 3) return false;
    const EdfInputDesc& d = p.inp[ii];
    if (d.in_dtype != EDF_F32 || d.out_dtype != EDF_F32) return false;
    if (d.nstep_rank != 0) return false;
    if (L.istr_e[ii][2] != 1) return false;
    for (int a = 0; a < 3; ++a)
        if (p.idim[a] < 8) return false;      // single-reflection mirror map of the batched forward kernel
    return true;
}

static void edf_lean_launch(int order, int gradient, dim3 grid, cudaStream_t st, const EdfParams& p,
                            const EdfFastLaunch& L, int ii)
{
    if (!gradient) {                                  // batched forward kernel
        switch (order) {
        case 0: edf_lean3d_fwd_kernel<0, 4><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        case 1: edf_lean3d_fwd_kernel<1, 4><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        case 2: edf_lean3d_fwd_kernel<2, EDF_LEAN_FWD_U3><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        case 3: edf_lean3d_fwd_kernel<3, EDF_LEAN_FWD_U3><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        case 4: edf_lean3d_fwd_kernel<4, 1><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        default: edf_lean3d_fwd_kernel<5, 1><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
        }
        return;
    }
    switch (order) {                                  // plain scatter (no accumulation window)
    case 0: edf_lean3d_kernel<0, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    case 1: edf_lean3d_kernel<1, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    case 2: edf_lean3d_kernel<2, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    case 3: edf_lean3d_kernel<3, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    case 4: edf_lean3d_kernel<4, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    default: edf_lean3d_kernel<5, true><<<grid, EDF_FAST_THREADS, 0, st>>>(p, L, ii); break;
    }
}

// =======================================================================================
// Gradient scatter with a shared-memory accumulation window (K2).
//
// The plain scatter issues (order+1)^3 global float atomics per voxel whose addresses coalesce
// badly (~10 L2 atomic requests per warp instruction): it is bound by L2 atomic throughput.
// Here a CTA (4 warps = 4 z-slabs x 32 x positions) accumulates the contributions of a chunk of
// 8 rows into a window of the dX volume held in shared memory and then adds the window to dX
// with perfectly coalesced atomics (each touched cell once per chunk, ~10 per voxel instead of
// 64).  Shared-memory float atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN), so the window
// accumulates in 32-bit FIXED POINT with native ATOMS.ADD: contributions are scaled so that
// the chunk's max |dY| maps to 2^22 (absolute resolution max|dY_chunk| * 2^-23, headroom 512x);
// the flush converts back and adds in float.  Voxels whose tap window touches the volume border
// or does not fit the accumulation window fall back to direct global atomics.
// =======================================================================================
#define EDF_GW_TX 32               // x positions per warp / CTA
#ifndef EDF_GW_G
#define EDF_GW_G 4                 // z-slabs per CTA
#endif
#define EDF_GW_RG 2                // row groups: warp w owns slab (w % G) and rows rg*MR .. rg*MR+MR-1 of a chunk
#define EDF_GW_MR (EDF_FAST_M / EDF_GW_RG)
#define EDF_GW_WARPS (EDF_GW_G * EDF_GW_RG)
#define EDF_GW_THREADS (EDF_GW_TX * EDF_GW_WARPS)
#ifndef EDF_GW_MINB
#define EDF_GW_MINB 2              // CTAs per SM the kernel is compiled for
#endif
#define EDF_GW_NC 8                // control-point span capacity of this kernel's tables
#ifndef EDF_GW_WZ
#define EDF_GW_WZ 19               // accumulation window (cells of dX) around the chunk's footprint:
#define EDF_GW_WY 23               //   tile extent + taps + ~0.2 x (other extents) + margins
#define EDF_GW_WX 52               // multiple of 4: the flush moves 16-byte groups
#endif
#define EDF_GW_MARGIN 2
#define EDF_GW_FIX 24              // chunk max |dY| maps into [2^(FIX-1), 2^FIX]

struct EdfGradWinSmem {
    double wz[EDF_GW_G][4];
    double wy[EDF_FAST_RY][4];
    double wx[EDF_GW_TX][4];
    int    sz[EDF_GW_G];
    int    sy[EDF_FAST_RY];
    int    sx[EDF_GW_TX];
    int    ny, nx, nonzero, gmax_bits;
    int    wmin[3], pad_;
    int    wmax[3], pad2_;          // density guard: extent of the chunk's corner coordinates
    unsigned rowmask[(EDF_GW_WZ * EDF_GW_WY + 31) / 32 + 1];   // window rows holding contributions (TMA flush)
    double A[3][EDF_GW_G][EDF_GW_NC][EDF_GW_NC];
    double Bw[EDF_GW_WARPS][3][EDF_GW_MR][EDF_GW_NC];
    __align__(128) int win[EDF_GW_WZ * EDF_GW_WY * EDF_GW_WX];   // 128-byte aligned: TMA source
};

// un-mapped source coordinates of voxel (z, y, x) from the warp's B table (same arithmetic as the
// forward lean kernel)
__device__ __forceinline__ void edf_gw_coords(const EdfParams& p, const double (*Bw)[EDF_GW_MR][EDF_GW_NC],
                                              int m, int sxrel, const double* wx, bool affine, int z, int y, int x,
                                              double bz, double bx, double offy, double& inz, double& iny, double& inx)
{
    double dz = 0.0, dy = 0.0, dx = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dz = fma(Bw[0][m][sxrel + k], wx[k], dz);
        dy = fma(Bw[1][m][sxrel + k], wx[k], dy);
        dx = fma(Bw[2][m][sxrel + k], wx[k], dx);
    }
    if (!affine) {
        inz = xadd(bz, dz);
        iny = xadd(xadd((double)y, offy), dy);
        inx = xadd(bx, dx);
    } else {
        const int o[3] = {z, y, x};
        inz = edf_source_coordinate<3, int>(p, o, 0, dz);
        iny = edf_source_coordinate<3, int>(p, o, 1, dy);
        inx = edf_source_coordinate<3, int>(p, o, 2, dx);
    }
}

// fixed-point rounding of g * w for the window accumulators: round-to-nearest-even of the exact product
// through the 1.5 * 2^23 magic add (|g * w| < 2^22)
__device__ __forceinline__ int edf_gw_round(float g, float w)
{
    return __float_as_int(fmaf(g, w, 12582912.0f)) - 0x4B400000;
}

// rare voxel of the window gradient (next to a rounding / boundary threshold): reference-order
// re-evaluation, then the general scatter with global atomics
template <int ORDER>
__device__ __noinline__ void edf_gradwin_slow_voxel(const EdfParams& p, const EdfFastLaunch& L, int ii,
                                                    int z, int y, int x, double inz, double iny, double inx, bool gate)
{
    int o[3] = {z, y, x};
    double in[3] = {inz, iny, inx};
    if (gate && (edf_near_half_integer(in[0]) || edf_near_half_integer(in[1]) || edf_near_half_integer(in[2]))) {
        double dd[3];
        edf_displacement_exact_cold<3>(p, o, dd);
#pragma unroll
        for (int h = 0; h < 3; ++h) in[h] = edf_source_coordinate<3, int>(p, o, h, dd[h]);
    }
    edf_fast_f32_one_input<3, ORDER, true>(p, L, ii, o, in);
}

// FLUSH: 0 = scalar atomics, 1 = 16-byte vector atomics, 2 = TMA bulk reduce-add per window row
// (cp.reduce.async.bulk ... .add.f32, SASS UBLKRED).  A single tensor-map reduce of the whole box
// (cp.reduce.async.bulk.tensor.3d, UTMAREDG) would be the natural form; it needs the window converted to
// float in place and a 16-byte aligned innermost box coordinate (round 1 took the traps of unaligned
// coordinates for a defect of the pool: scripts/experiments/tma_tensor_test.cu, DESIGN.md "TMA").
template <int ORDER, int FLUSH>
__global__ void __launch_bounds__(EDF_GW_THREADS, EDF_GW_MINB)
edf_lean3d_gradwin_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    EdfGradWinSmem& s = *reinterpret_cast<EdfGradWinSmem*>(smem_raw);
    constexpr int NT = ORDER + 1;
    constexpr int NWIN = EDF_GW_WZ * EDF_GW_WY * EDF_GW_WX;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * EDF_GW_TX;
    const int ry = (int)L.rows_per_cta;
    const int y0 = blockIdx.y * ry;
    const int z0 = blockIdx.z * EDF_GW_G;

    // ---- prologue: control tables, z-contraction A, zeroed window
    if (tid == 0) s.nonzero = 0;
    if (tid < EDF_GW_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_GW_TX + EDF_FAST_RY) {
        const int t = tid - EDF_GW_TX;
        edf_fast_ctrl_entry(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1), s.wy[t], &s.sy[t]);
    } else if (tid < EDF_GW_TX + EDF_FAST_RY + EDF_GW_G) {
        const int t = tid - EDF_GW_TX - EDF_FAST_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    }
    for (int e = tid; e < NWIN; e += EDF_GW_THREADS) s.win[e] = 0;
    if (tid < (EDF_GW_WZ * EDF_GW_WY + 31) / 32 + 1) s.rowmask[tid] = 0;
    __syncthreads();
    {
        const int sy_min0 = s.sy[0], sx_min0 = s.sx[0];
        const int ny = s.sy[EDF_FAST_RY - 1] - sy_min0 + 4;
        const int nxx = s.sx[EDF_GW_TX - 1] - sx_min0 + 4;
        if (tid == 0) { s.ny = ny; s.nx = nxx; }
        bool nz = false;
        const int na = 3 * EDF_GW_G * ny * nxx;
        for (int e = tid; e < na; e += EDF_GW_THREADS) {
            const int jx = e % nxx;
            const int jy = (e / nxx) % ny;
            const int t = (e / (nxx * ny)) % EDF_GW_G;
            const int h = e / (nxx * ny * EDF_GW_G);
            const int my = edf_mirror_index32(sy_min0 + jy, (int)p.ncp[1]);
            const int mx = edf_mirror_index32(sx_min0 + jx, (int)p.ncp[2]);
            const char* base = p.disp + p.dstr[0] * h + my * p.dstr[2] + mx * p.dstr[3];
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int mz = edf_mirror_index32(s.sz[t] + i, (int)p.ncp[0]);
                const double cf = (p.ddtype == EDF_F64) ? *(const double*)(base + mz * p.dstr[1])
                                                        : (double)*(const float*)(base + mz * p.dstr[1]);
                nz |= (cf != 0.0);
                a = fma(cf, s.wz[t][i], a);
            }
            s.A[h][t][jy][jx] = a;
        }
        if (nz) s.nonzero = 1;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    const int g = warp % EDF_GW_G, rg = warp / EDF_GW_G;          // slab and row group of this warp
    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[lane][k];
    const int sxrel = s.sx[lane] - s.sx[0];
    const int nchunk = min(ry / EDF_FAST_M, (ody - y0 + EDF_FAST_M - 1) / EDF_FAST_M);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_GW_MR][EDF_GW_NC] = s.Bw[warp];

    const EdfInputDesc& d = p.inp[ii];
    float* __restrict__ pdx = (float*)d.in;                      // dX accumulator
    const float* __restrict__ pdy = (const float*)d.out;         // upstream gradient dY
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int64_t obase_zx = (int64_t)z * L.ostr_e[ii][0] + (int64_t)x * L.ostr_e[ii][2];
    const bool affine = p.has_affine != 0;
    const int mode = d.mode;
    const double bz = xadd((double)z, p.ooff_d[0]);
    const double bx = xadd((double)x, p.ooff_d[2]);
    const double offy = p.ooff_d[1];
    const int ozmax = odz - 1 - z0 < EDF_GW_G - 1 ? odz - 1 - z0 : EDF_GW_G - 1;   // last valid slab
    const int oxmax = odx - 1 - x0 < EDF_GW_TX - 1 ? odx - 1 - x0 : EDF_GW_TX - 1;     // last valid lane

    for (int c = 0; c < nchunk; ++c) {
        const int yc0 = y0 + c * EDF_FAST_M;
        const int mlast = min(EDF_FAST_M - 1, ody - 1 - yc0);
        // ---- warp-private y-contraction for this warp's rows of the chunk
        {
            // lane -> (row m, phase q); the 3 * nx (component, column) pairs go round the phases (no division)
            static_assert(EDF_GW_MR == 4, "lane mapping of the B table assumes 4 rows per warp");
            const int m = lane & 3, q = lane >> 2;
            const int row = c * EDF_FAST_M + rg * EDF_GW_MR + m;
            const int r0 = s.sy[row] - sy_min;
            double wyr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) wyr[j] = s.wy[row][j];
            int h = 0, jx = q;
            while (jx >= nx) { jx -= nx; ++h; }
            while (h < 3) {
                double b = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) b = fma(s.A[h][g][r0 + j][jx], wyr[j], b);
                Bw[h][m][jx] = b;
                jx += 8;
                while (jx >= nx) { jx -= nx; ++h; }
            }
        }
        if (tid == 0) { s.wmin[0] = s.wmin[1] = s.wmin[2] = 0x7fffffff; s.wmax[0] = s.wmax[1] = s.wmax[2] = (int)0x80000000; s.gmax_bits = 0; }
        __syncthreads();
        // ---- chunk statistics: max |dY| (fixed-point scale) and the window origin from the
        //      corner voxels of the chunk box
        const int mr0 = rg * EDF_GW_MR;                            // first row of this warp in the chunk
        const int mrlast = min(EDF_GW_MR - 1, mlast - mr0);         // last valid local row (may be < 0)
        float gmax = 0.f;
#pragma unroll
        for (int m = 0; m < EDF_GW_MR; ++m) {
            const float gq = (tok && m <= mrlast) ? __ldg(pdy + (obase_zx + (int64_t)(yc0 + mr0 + m) * osy)) : 0.f;
            gmax = fmaxf(gmax, fabsf(gq));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        if (lane == 0 && gmax > 0.f) atomicMax(&s.gmax_bits, __float_as_int(gmax));
        if ((g == 0 || g == ozmax) && (lane == 0 || lane == oxmax) && tok && mrlast >= 0) {
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                const int m = q ? mrlast : 0;
                double inz, iny, inx;
                edf_gw_coords(p, Bw, m, sxrel, wx, affine, z, yc0 + mr0 + m, x, bz, bx, offy, inz, iny, inx);
                // floor of the clamped coordinate: out-of-volume corners must not drag the window away
                const int fz_ = (int)floor(fmin(fmax(inz, 0.0), limz));
                const int fy_ = (int)floor(fmin(fmax(iny, 0.0), limy));
                const int fx_ = (int)floor(fmin(fmax(inx, 0.0), limx));
                atomicMin(&s.wmin[0], fz_);
                atomicMin(&s.wmin[1], fy_);
                atomicMin(&s.wmin[2], fx_);
                atomicMax(&s.wmax[0], fz_);
                atomicMax(&s.wmax[1], fy_);
                atomicMax(&s.wmax[2], fx_);
            }
        }
        __syncthreads();
        const float gmax_c = __int_as_float(s.gmax_bits);
        if (gmax_c > 0.f) {                                       // an all-zero chunk contributes nothing
            // power-of-two scale: max|dY| of the chunk -> [2^(FIX-1), 2^FIX]
            int ex;
            frexpf(gmax_c, &ex);                                  // gmax = f * 2^ex, f in [0.5, 1)
            // orders >= 2 convert with the magic-number add (FFMA + IADD; F2I runs at a quarter of the rate on
            // the XU pipe), which holds |v| < 2^22: one bit less of scale there (largest weight product 0.42)
            constexpr int FIX = (ORDER >= 2) ? EDF_GW_FIX - 1 : EDF_GW_FIX;
            // orders >= 2: the largest possible contribution (max|dY| times the largest weight product of the
            // order) maps to just under 2^22, the range of the magic-number rounding: 1.2-6x finer than the
            // power-of-two scale, which matters once the prefilter adjoint amplifies the rounding noise
            constexpr float WMAX = (ORDER <= 1) ? 1.0f : (ORDER == 2) ? 0.75f : (ORDER == 3) ? (2.0f / 3.0f)
                                 : (ORDER == 4) ? (115.0f / 192.0f) : 0.55f;
            const float scale = (ORDER >= 2) ? (4194304.0f * 0.999f) / (fmaxf(gmax_c, 1e-30f) * (WMAX * WMAX * WMAX))
                                             : ldexpf(1.0f, FIX - ex);
            const float inv_scale = (ORDER >= 2) ? 1.0f / scale : ldexpf(1.0f, ex - FIX);
            // density guard (see edf_swin.cuh): where the map collapses -- the chunk's 1024 voxels land on a handful of
            // cells -- the 32-bit fixed-point cells would wrap; such a chunk gets a window no voxel can hit and
            // scatters straight to dX in float
            const long long gcells = (long long)(s.wmax[0] - s.wmin[0] + 1) * (s.wmax[1] - s.wmin[1] + 1) * (s.wmax[2] - s.wmin[2] + 1);
            const bool gdense = 1024 > 16 * gcells;
            const int wz0 = gdense ? -0x20000000 : s.wmin[0] - (ORDER + 1) / 2 - EDF_GW_MARGIN;
            const int wy0 = s.wmin[1] - (ORDER + 1) / 2 - EDF_GW_MARGIN;
            const int wx0 = (s.wmin[2] - (ORDER + 1) / 2 - EDF_GW_MARGIN) & ~3;   // 16-byte aligned columns
            if (tok && ORDER <= 1) {
                // rows of this warp, U at a time: branch-free coordinates / classification for all U, then the
                // scatters.  Rare voxels (next to a threshold) go out of line; voxels at the volume border or
                // outside the window use direct global atomics.
                constexpr int U = 4;      // orders 0/1 only: at order 3 the batched form spills (0.92 vs 0.87 ms)
                const bool cmode = mode == EDF_MODE_CONSTANT;
#pragma unroll 1
                for (int m0 = 0; m0 <= mrlast; m0 += U) {
                    double inz[U], iny[U], inx[U];
                    int stz[U], sty[U], stx[U];
                    float fz[U], fy[U], fx[U], gv[U];
                    bool act[U], slow[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int m = min(m0 + u, mrlast);
                        const int y = yc0 + mr0 + m;
                        const bool valid = (m0 + u) <= mrlast;
                        gv[u] = valid ? __ldg(pdy + (obase_zx + (int64_t)y * osy)) : 0.f;   // L1 hit (read above)
                        edf_gw_coords(p, Bw, m, sxrel, wx, affine, z, y, x, bz, bx, offy, inz[u], iny[u], inx[u]);
                        double cz = fmin(fmax(inz[u], 0.0), limz);
                        double cy = fmin(fmax(iny[u], 0.0), limy);
                        double cx = fmin(fmax(inx[u], 0.0), limx);
                        const bool inr = (cz == inz[u]) & (cy == iny[u]) & (cx == inx[u]);
                        bool mapped_danger = false, nanflag = false;
                        if (!cmode && !inr) {
                            if (cz != inz[u]) { mapped_danger |= edf_near_half_integer(inz[u]); cz = edf_map_coordinate_cold(inz[u], lenz, mode); }
                            if (cy != iny[u]) { mapped_danger |= edf_near_half_integer(iny[u]); cy = edf_map_coordinate_cold(iny[u], leny, mode); }
                            if (cx != inx[u]) { mapped_danger |= edf_near_half_integer(inx[u]); cx = edf_map_coordinate_cold(inx[u], lenx, mode); }
                            if (!((cz > -1.0) & (cy > -1.0) & (cx > -1.0))) { nanflag = true; cz = cy = cx = 0.0; }
                        }
                        const double flz = (ORDER & 1) ? floor(cz) : floor(xadd(cz, 0.5));
                        const double fly = (ORDER & 1) ? floor(cy) : floor(xadd(cy, 0.5));
                        const double flx = (ORDER & 1) ? floor(cx) : floor(xadd(cx, 0.5));
                        fz[u] = (float)xsub(cz, flz);
                        fy[u] = (float)xsub(cy, fly);
                        fx[u] = (float)xsub(cx, flx);
                        stz[u] = (int)flz - ORDER / 2;
                        sty[u] = (int)fly - ORDER / 2;
                        stx[u] = (int)flx - ORDER / 2;
                        bool danger;
                        if (ORDER & 1)
                            danger = (fz[u] < EDF_LEAN_EPSF) | (fz[u] > 1.0f - EDF_LEAN_EPSF) | (fy[u] < EDF_LEAN_EPSF) |
                                     (fy[u] > 1.0f - EDF_LEAN_EPSF) | (fx[u] < EDF_LEAN_EPSF) | (fx[u] > 1.0f - EDF_LEAN_EPSF);
                        else
                            danger = (fabsf(fz[u]) < EDF_LEAN_EPSF) | (fabsf(fz[u]) > 0.5f - EDF_LEAN_EPSF) |
                                     (fabsf(fy[u]) < EDF_LEAN_EPSF) | (fabsf(fy[u]) > 0.5f - EDF_LEAN_EPSF) |
                                     (fabsf(fx[u]) < EDF_LEAN_EPSF) | (fabsf(fx[u]) > 0.5f - EDF_LEAN_EPSF);
                        danger |= mapped_danger;
                        bool nearmiss = false;
                        if (cmode) {
                            const double qz = fabs(xsub(inz[u], cz)), qy = fabs(xsub(iny[u], cy)), qx = fabs(xsub(inx[u], cx));
                            nearmiss = ((qz > 0.0) & (qz < EDF_FAST_EPS)) | ((qy > 0.0) & (qy < EDF_FAST_EPS)) |
                                       ((qx > 0.0) & (qx < EDF_FAST_EPS));
                        }
                        const bool live = valid & (gv[u] != 0.f);
                        slow[u] = live & ((gate & ((inr | !cmode) ? danger : nearmiss)) | nanflag);
                        act[u] = live & !slow[u] & (inr | !cmode);          // constant voxels pass no gradient (deform.c:928)
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (!act[u]) continue;
                        float wzf[NT], wyf[NT], wxf[NT];
                        if (ORDER > 0) {
                            edf_bspline_weights_f32<ORDER>(fz[u], wzf);
                            edf_bspline_weights_f32<ORDER>(fy[u], wyf);
                            edf_bspline_weights_f32<ORDER>(fx[u], wxf);
                        }
                        const bool edge = (stz[u] < 0) | (stz[u] + ORDER >= lenz) | (sty[u] < 0) | (sty[u] + ORDER >= leny) |
                                          (stx[u] < 0) | (stx[u] + ORDER >= lenx);
                        const int rz = stz[u] - wz0, ry_ = sty[u] - wy0, rx = stx[u] - wx0;
                        const bool inwin = (rz >= 0) & (rz + ORDER < EDF_GW_WZ) & (ry_ >= 0) & (ry_ + ORDER < EDF_GW_WY) &
                                           (rx >= 0) & (rx + ORDER < EDF_GW_WX);
                        if (!edge && inwin) {
                            int* wbase = s.win + ((rz * EDF_GW_WY + ry_) * EDF_GW_WX + rx);
                            const float gs = gv[u] * scale;
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const float gi = (ORDER > 0) ? gs * wzf[i] : gs;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float gj = (ORDER > 0) ? gi * wyf[j] : gi;
                                    int* r = wbase + (i * EDF_GW_WY + j) * EDF_GW_WX;
#pragma unroll
                                    for (int k = 0; k < NT; ++k)
                                        atomicAdd(r + k, __float2int_rn((ORDER > 0) ? gj * wxf[k] : gj));
                                }
                            }
                        } else {
                            // border of the volume / outside the accumulation window: direct global atomics
                            int ozt[NT], oyt[NT], oxt[NT];
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                ozt[i] = edf_mirror_index32(stz[u] + i, lenz) * isz;
                                oyt[i] = edf_mirror_index32(sty[u] + i, leny) * isy;
                                oxt[i] = edf_mirror_index32(stx[u] + i, lenx);
                            }
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const float gi = (ORDER > 0) ? gv[u] * wzf[i] : gv[u];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float gj = (ORDER > 0) ? gi * wyf[j] : gi;
                                    float* r = pdx + (ozt[i] + oyt[j]);
#pragma unroll
                                    for (int k = 0; k < NT; ++k)
                                        atomicAdd(r + oxt[k], (ORDER > 0) ? gj * wxf[k] : gj);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (slow[u])
                            edf_gradwin_slow_voxel<ORDER>(p, L, ii, z, yc0 + mr0 + m0 + u, x, inz[u], iny[u], inx[u], gate);
                }
            }
            if (tok && ORDER >= 2) {
                // one voxel at a time (64+ taps keep the registers full)
#pragma unroll 1
                for (int m = 0; m <= mrlast; ++m) {
                    const int y = yc0 + mr0 + m;
                    const float gval = __ldg(pdy + (obase_zx + (int64_t)y * osy));   // L1 hit (read above)
                    if (gval == 0.f) continue;
                    double inz, iny, inx;
                    edf_gw_coords(p, Bw, m, sxrel, wx, affine, z, y, x, bz, bx, offy, inz, iny, inx);
                    int stz = 0, sty = 0, stx = 0;
                    float fz = 0.f, fy = 0.f, fx = 0.f;
                    bool constant, danger;
#pragma unroll 1
                    for (int pass = 0;; ++pass) {
                        danger = false;
                        constant = edf_lean_axis<ORDER>(p, 0, mode, inz, limz, gate, stz, fz, danger);
                        if (!constant) constant = edf_lean_axis<ORDER>(p, 1, mode, iny, limy, gate, sty, fy, danger);
                        if (!constant) constant = edf_lean_axis<ORDER>(p, 2, mode, inx, limx, gate, stx, fx, danger);
                        if (!danger || pass) break;
                        edf_lean_exact_coords(p, z, y, x, &inz, &iny, &inx);
                    }
                    if (constant) continue;                        // deform.c:928: no gradient through cval
                    float wzf[NT], wyf[NT], wxf[NT];
                    if (ORDER > 0) {
                        edf_bspline_weights_f32<ORDER>(fz, wzf);
                        edf_bspline_weights_f32<ORDER>(fy, wyf);
                        edf_bspline_weights_f32<ORDER>(fx, wxf);
                    }
                    const bool edge = (stz < 0) | (stz + ORDER >= lenz) | (sty < 0) | (sty + ORDER >= leny) |
                                      (stx < 0) | (stx + ORDER >= lenx);
                    // a warp with a border voxel takes the mirrored-index form for all its lanes (identical
                    // results: the mirror of an inside index is the index itself)
                    const bool wedge = __any_sync(__activemask(), edge);
                    const float gs = gval * scale;
                    bool done = false;
                    if (!wedge) {
                        const int rz = stz - wz0, ry = sty - wy0, rx = stx - wx0;
                        const bool inwin = (rz >= 0) & (rz + ORDER < EDF_GW_WZ) & (ry >= 0) & (ry + ORDER < EDF_GW_WY) &
                                           (rx >= 0) & (rx + ORDER < EDF_GW_WX);
                        if (inwin) {
                            int* wbase = s.win + ((rz * EDF_GW_WY + ry) * EDF_GW_WX + rx);
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const float gi = gs * wzf[i];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float gj = gi * wyf[j];
                                    int* r = wbase + (i * EDF_GW_WY + j) * EDF_GW_WX;
#pragma unroll
                                    for (int k = 0; k < NT; ++k)
                                        atomicAdd(r + k, edf_gw_round(gj, wxf[k]));
                                }
                            }
                            done = true;
                        }
                    } else {
                        int rzi[NT], ryi[NT], rxi[NT];
                        bool inw = true;
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            rzi[i] = edf_mirror1(stz + i, lenz) - wz0;
                            ryi[i] = edf_mirror1(sty + i, leny) - wy0;
                            rxi[i] = edf_mirror1(stx + i, lenx) - wx0;
                            inw &= ((unsigned)rzi[i] < (unsigned)EDF_GW_WZ) & ((unsigned)ryi[i] < (unsigned)EDF_GW_WY) &
                                   ((unsigned)rxi[i] < (unsigned)EDF_GW_WX);
                            rzi[i] *= EDF_GW_WY * EDF_GW_WX;
                            ryi[i] *= EDF_GW_WX;
                        }
                        // single reflection suffices only while the window start is within one length of the volume
                        inw &= (stz >= -lenz) & (stz + ORDER < 2 * lenz - 1) & (sty >= -leny) & (sty + ORDER < 2 * leny - 1) &
                               (stx >= -lenx) & (stx + ORDER < 2 * lenx - 1);
                        if (inw) {
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                const float gi = gs * wzf[i];
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float gj = gi * wyf[j];
                                    int* r = s.win + (rzi[i] + ryi[j]);
#pragma unroll
                                    for (int k = 0; k < NT; ++k)
                                        atomicAdd(r + rxi[k], edf_gw_round(gj, wxf[k]));
                                }
                            }
                            done = true;
                        }
                    }
                    if (!done) {
                        // outside the accumulation window: direct global atomics
                        int ozt[NT], oyt[NT], oxt[NT];
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            ozt[i] = edf_mirror_index32(stz + i, lenz) * isz;
                            oyt[i] = edf_mirror_index32(sty + i, leny) * isy;
                            oxt[i] = edf_mirror_index32(stx + i, lenx);
                        }
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            const float gi = (ORDER > 0) ? gval * wzf[i] : gval;
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const float gj = (ORDER > 0) ? gi * wyf[j] : gi;
                                float* r = pdx + (ozt[i] + oyt[j]);
#pragma unroll
                                for (int k = 0; k < NT; ++k)
                                    atomicAdd(r + oxt[k], (ORDER > 0) ? gj * wxf[k] : gj);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- flush: every touched window cell once, coalesced along x, and re-zero
            if (FLUSH == 2) {
                // TMA flush: convert the fixed-point window to float in place, then one bulk reduce-add
                // (cp.reduce.async.bulk ... .add.f32, SASS UBLKRED) per window ROW that holds contributions:
                // a row is 52 contiguous floats both in shared memory and in dX.
                constexpr int GPR = EDF_GW_WX / 4;                 // 16-byte groups per row
                for (int q = tid; q < NWIN / 4; q += EDF_GW_THREADS) {
                    const int4 v = reinterpret_cast<int4*>(s.win)[q];
                    if ((v.x | v.y | v.z | v.w) != 0) {
                        const int row = q / GPR;
                        atomicOr(&s.rowmask[row >> 5], 1u << (row & 31));
                        reinterpret_cast<float4*>(s.win)[q] = make_float4((float)v.x * inv_scale, (float)v.y * inv_scale,
                                                                         (float)v.z * inv_scale, (float)v.w * inv_scale);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                bool issued = false;
                for (int row = tid; row < EDF_GW_WZ * EDF_GW_WY; row += EDF_GW_THREADS) {
                    if (!((s.rowmask[row >> 5] >> (row & 31)) & 1u)) continue;
                    const int iz = row / EDF_GW_WY, iy = row % EDF_GW_WY;
                    const int xs = max(wx0, 0), xe = min(wx0 + EDF_GW_WX, lenx);   // lenx % 4 == 0 (host-checked)
                    // rows outside the volume never receive contributions (their voxels take the fallback)
                    float* dst = pdx + ((wz0 + iz) * isz + (wy0 + iy) * isy + xs);
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(s.win + row * EDF_GW_WX + (xs - wx0));
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 :: "l"(dst), "r"(src), "r"((uint32_t)((xe - xs) * 4)) : "memory");
                    issued = true;
                }
                if (issued) {
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                __syncthreads();
                for (int q = tid; q < NWIN / 4; q += EDF_GW_THREADS) {
                    const int row = q / GPR;
                    if ((s.rowmask[row >> 5] >> (row & 31)) & 1u) reinterpret_cast<int4*>(s.win)[q] = make_int4(0, 0, 0, 0);
                }
                __syncthreads();
                if (tid < (EDF_GW_WZ * EDF_GW_WY + 31) / 32 + 1) s.rowmask[tid] = 0;
            } else if (FLUSH == 1) {
                // 16-byte group q = tid, tid + THREADS, ...: (plane, row, group-in-row) advance incrementally
                constexpr int GPR = EDF_GW_WX / 4;                                  // groups per row
                constexpr int DG = EDF_GW_THREADS % GPR, DR = EDF_GW_THREADS / GPR;  // step in groups / rows
                static_assert(DR + 1 < 2 * EDF_GW_WY, "at most two row wraps per step");
                int cg = tid % GPR, iy = (tid / GPR) % EDF_GW_WY, iz = tid / (GPR * EDF_GW_WY);
                for (int q = tid; q < NWIN / 4; q += EDF_GW_THREADS) {
                    int4 v = reinterpret_cast<int4*>(s.win)[q];
                    if ((v.x | v.y | v.z | v.w) != 0) {
                        reinterpret_cast<int4*>(s.win)[q] = make_int4(0, 0, 0, 0);
                        float4* dst = reinterpret_cast<float4*>(pdx + ((wz0 + iz) * isz + (wy0 + iy) * isy + (wx0 + 4 * cg)));
                        atomicAdd(dst, make_float4((float)v.x * inv_scale, (float)v.y * inv_scale,
                                                   (float)v.z * inv_scale, (float)v.w * inv_scale));
                    }
                    cg += DG;
                    const int carry = cg >= GPR;
                    cg -= carry ? GPR : 0;
                    iy += DR + carry;
                    int wrap = iy >= EDF_GW_WY;
                    iy -= wrap ? EDF_GW_WY : 0;
                    iz += wrap;
                    if (DR + 1 >= EDF_GW_WY) {
                        wrap = iy >= EDF_GW_WY;
                        iy -= wrap ? EDF_GW_WY : 0;
                        iz += wrap;
                    }
                }
            } else {
                for (int e = tid; e < NWIN; e += EDF_GW_THREADS) {
                    const int v = s.win[e];
                    if (v != 0) {
                        s.win[e] = 0;
                        const int ix = e % EDF_GW_WX;
                        const int iy = (e / EDF_GW_WX) % EDF_GW_WY;
                        const int iz = e / (EDF_GW_WX * EDF_GW_WY);
                        atomicAdd(pdx + ((wz0 + iz) * isz + (wy0 + iy) * isy + (wx0 + ix)), (float)v * inv_scale);
                    }
                }
            }
        }
        __syncthreads();
    }
}

static EdfPerDeviceFlag g_gradwin_configured;

// The window kernel uses 32-lane x tiles; its control-point tables must hold a 32-wide span, and
// dX element offsets must fit 32 bits (already guaranteed by edf_fast_input_class).
static bool edf_gradwin_eligible(const EdfParams& p)
{
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("EDF_NO_GRADWIN"); disabled = (e && *e && *e != '0') ? 1 : 0; }
    if (disabled) return false;
    return p.naxis == 3 && edf_fast_ctrl_span_ok(p, 2, EDF_GW_TX, EDF_GW_NC) && edf_fast_ctrl_span_ok(p, 1, EDF_FAST_RY, EDF_GW_NC);
}

static int edf_lean_launch_gradwin(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    dim3 grid;
    grid.x = (unsigned)((p.odim[2] + EDF_GW_TX - 1) / EDF_GW_TX);
    grid.z = (unsigned)((p.odim[0] + EDF_GW_G - 1) / EDF_GW_G);
    unsigned ry = EDF_FAST_RY;                       // fewer rows per CTA for small volumes (fill 148 SMs x 2 CTAs twice)
    while (ry > EDF_FAST_M && (uint64_t)grid.x * ((p.odim[1] + ry - 1) / ry) * grid.z < 4ull * 148) ry >>= 1;
    grid.y = (unsigned)((p.odim[1] + ry - 1) / ry);
    EdfFastLaunch L = Lin;
    L.rows_per_cta = ry;
    const size_t smem = sizeof(EdfGradWinSmem);
    if (!g_gradwin_configured.test()) {
#define EDF_GW_ATTR(O)                                                                                           \
    cudaFuncSetAttribute(edf_lean3d_gradwin_kernel<O, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    cudaFuncSetAttribute(edf_lean3d_gradwin_kernel<O, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    cudaFuncSetAttribute(edf_lean3d_gradwin_kernel<O, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        EDF_GW_ATTR(0); EDF_GW_ATTR(1); EDF_GW_ATTR(2); EDF_GW_ATTR(3); EDF_GW_ATTR(4); EDF_GW_ATTR(5);
#undef EDF_GW_ATTR
        if (cudaGetLastError() != cudaSuccess) return -1;
        g_gradwin_configured.set();
    }
    // 16-byte vector flush / TMA need 16-byte aligned rows of dX
    const bool vec = ((uintptr_t)p.inp[ii].in % 16 == 0) && (L.istr_e[ii][0] % 4 == 0) && (L.istr_e[ii][1] % 4 == 0);
    static int want_tma = -1;
    if (want_tma < 0) { const char* e = getenv("EDF_GRADWIN_TMA"); want_tma = (e && *e && *e != '0') ? 1 : 0; }
    int flush = vec ? 1 : 0;
    if (vec && want_tma && p.idim[2] % 4 == 0) flush = 2;
#define EDF_GW_CASE(O)                                                                                         \
    case O:                                                                                                    \
        if (flush == 2)      edf_lean3d_gradwin_kernel<O, 2><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii); \
        else if (flush == 1) edf_lean3d_gradwin_kernel<O, 1><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii); \
        else                 edf_lean3d_gradwin_kernel<O, 0><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii); \
        break;
    switch (order) {
        EDF_GW_CASE(0) EDF_GW_CASE(1) EDF_GW_CASE(2) EDF_GW_CASE(3) EDF_GW_CASE(4)
    default:
        if (flush == 2)      edf_lean3d_gradwin_kernel<5, 2><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii);
        else if (flush == 1) edf_lean3d_gradwin_kernel<5, 1><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii);
        else                 edf_lean3d_gradwin_kernel<5, 0><<<grid, EDF_GW_THREADS, smem, st>>>(p, L, ii);
        break;
    }
#undef EDF_GW_CASE
    return flush == 2 ? 2 : 0;
}
