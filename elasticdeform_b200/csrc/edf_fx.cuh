// edf_fx.cuh -- polynomial form of the displacement along a thread's column and fixed-point source coordinates
// (shared by the staged-window kernels, edf_swin.cuh, and the direct / tile kernels, edf_poly.cuh, edf_tile.cuh).
#pragma once
#include "edf_lean.cuh"
#include <limits.h>

#define EDF_PL_TX 32               // x positions per warp / CTA
#define EDF_PL_NC 8                // control columns the 32 lanes of a warp can touch (span + 4)

// Rebuild the polynomial coefficients of this thread's column for the control interval whose window starts at
// control row j0.  Warp-collective (all 32 lanes).  Returns the warp's gate: false when every control
// coefficient the warp touches is zero (then a == 0 exactly).
template <class Tab>
__device__ __forceinline__ bool edf_poly_build(const EdfParams& p, Tab& s, int g, int lane, int j0, double* a /*[3][4]*/, int tw = -1)
{
    static_assert(EDF_PL_NC == 8, "lane -> (control row, control column) mapping");
    if (tw < 0) tw = g;                                            // table slot of this warp
    const int j = lane >> 3, kx = lane & 7;
    const int sx0 = s.sx[0];
    const int nxw = s.sx[EDF_PL_TX - 1] - sx0 + 4;
    bool nz = false;
    __syncwarp();
    if (kx < nxw) {
        const int my = edf_mirror_index32(j0 + j, (int)p.ncp[1]);
        const int mx = edf_mirror_index32(sx0 + kx, (int)p.ncp[2]);
        const bool f64 = p.ddtype == EDF_F64;
        int64_t oz[4];
        double w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            oz[i] = (int64_t)edf_mirror_index32(s.sz[g] + i, (int)p.ncp[0]) * p.dstr[1];
            w[i] = s.wz[g][i];
        }
        const char* base = p.disp + (int64_t)my * p.dstr[2] + (int64_t)mx * p.dstr[3];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const char* bh = base + p.dstr[0] * h;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double cf = f64 ? *(const double*)(bh + oz[i]) : (double)*(const float*)(bh + oz[i]);
                nz |= (cf != 0.0);
                acc = fma(cf, w[i], acc);
            }
            s.T[tw][h][j][kx] = acc;
        }
    }
    const bool gate = __any_sync(0xffffffffu, nz);
    // (the __any_sync above orders the table writes before the reads below)
    const int sxrel = s.sx[lane] - sx0;
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[lane][k];
#pragma unroll
    for (int h = 0; h < 3; ++h) {
        double E[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            double e = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) e = fma(s.T[tw][h][jj][sxrel + k], wx[k], e);
            E[jj] = e;
        }
        // uniform cubic B-spline segment -> power basis in u (weights as in deform.c:171-177)
        a[h * 4 + 0] = (E[0] + 4.0 * E[1] + E[2]) * (1.0 / 6.0);
        a[h * 4 + 1] = (E[2] - E[0]) * 0.5;
        a[h * 4 + 2] = (E[0] - 2.0 * E[1] + E[2]) * 0.5;
        a[h * 4 + 3] = ((E[3] - E[0]) + 3.0 * (E[1] - E[2])) * (1.0 / 6.0);
    }
    __syncwarp();
    return gate;
}


// ---------------------------------------------------------------------------------------------------------
// Fixed-point source coordinates.  The row's own index, the crop offset and the affine map (deform.c:771-781)
// are folded into the column's polynomial, and so is the constant 1.5 * 2^29: the last FMA of the Horner form
// then rounds the SOURCE COORDINATE c to a multiple of 2^-23, and floor / fractional offset / range tests are
// integer operations on the two words of the result -- no conversion instructions, no fp64 compares:
//     T = c + 1.5*2^29   ->   floor(c) = bits [23,55) of T - const,    frac(c) = (lo & 0x7fffff) * 2^-23.
// Even orders fold another 0.5 in (their window start is floor(c + 0.5), deform.c:784-788).  The quantisation
// moves a coordinate by < 1.2e-7; voxels within 2^-21 of a threshold (integer coordinates for odd orders, integer
// and half-integer ones for even orders) are flagged `slow` and redone in the reference order, so every discrete
// decision still equals the reference's.
// ---------------------------------------------------------------------------------------------------------
#define EDF_PP_FBITS 23
#define EDF_PP_MAGIC 805306368.0   // 1.5 * 2^29: ulp 2^-23
#define EDF_PP_HI0 0x41C00000u     // high word of 2^29 (exponent 1052)
#define EDF_PP_FLBIAS 0x90000000u  // bits [23,55) of the pattern of EDF_PP_MAGIC
#define EDF_PP_NEAR 4.7683716e-7f  // 2^-21: four steps of the 2^-23 grid (quantisation: half a step each for the fold and the last FMA)

// a[h*4 + k] (displacement along axis h as a cubic in u, edf_poly_build) -> out[h*4 + k]: source coordinate
// (+ 0.5 for even orders) + 1.5*2^29 of the column (z, x) as a cubic in u, for the control interval whose window
// starts at control row j0.  r = (I_y - 1) / (P_y - 1): y + off_y = (j0 + 1 + u) * r  (cp = (P-1)(y+off)/(I-1), deform.c:655)
template <int ORDER>
__device__ __forceinline__ void edf_poly_fold(const EdfParams& p, const double* a, double r, int j0, int z, int x, double* out)
{
    const double yj = xmul((double)(j0 + 1), r);
    const double half = (ORDER & 1) ? 0.0 : 0.5;
    if (p.has_affine) {
        const double yo = xsub(yj, p.ooff_d[1]);                 // output row index at u = 0
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const double* A = p.affine + h * 4;
            const double c = fma(A[0], (double)z, fma(A[1], yo, fma(A[2], (double)x, A[3]))) + p.ooff_d[h];
            out[h * 4 + 0] = ((a[h * 4 + 0] + c) + half) + EDF_PP_MAGIC;
            out[h * 4 + 1] = fma(A[1], r, a[h * 4 + 1]);
            out[h * 4 + 2] = a[h * 4 + 2];
            out[h * 4 + 3] = a[h * 4 + 3];
        }
    } else {
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const double c = (h == 0) ? xadd((double)z, p.ooff_d[0]) : (h == 1) ? yj : xadd((double)x, p.ooff_d[2]);
            out[h * 4 + 0] = ((a[h * 4 + 0] + c) + half) + EDF_PP_MAGIC;
            out[h * 4 + 1] = (h == 1) ? a[h * 4 + 1] + r : a[h * 4 + 1];
            out[h * 4 + 2] = a[h * 4 + 2];
            out[h * 4 + 3] = a[h * 4 + 3];
        }
    }
}

// Source coordinates of one voxel from the folded polynomial (see the header): window starts, centred fractional
// offsets e = frac - 0.5, strict in-range flag and `slow` (redo in the reference order).
struct EdfPipeVoxel {
    int stz, sty, stx;
    float ez, ey, ex;
    bool inr, slow;
};
template <int ORDER>
__device__ __forceinline__ void edf_pipe_coords(const double* a, double u, bool gate, int lenz, int leny, int lenx,
                                                unsigned rngz, unsigned rngy, unsigned rngx, EdfPipeVoxel& v)
{
    const double Tz = fma(fma(fma(a[3], u, a[2]), u, a[1]), u, a[0]);
    const double Ty = fma(fma(fma(a[7], u, a[6]), u, a[5]), u, a[4]);
    const double Tx = fma(fma(fma(a[11], u, a[10]), u, a[9]), u, a[8]);
    const unsigned loz = (unsigned)__double2loint(Tz), hiz = (unsigned)__double2hiint(Tz);
    const unsigned loy = (unsigned)__double2loint(Ty), hiy = (unsigned)__double2hiint(Ty);
    const unsigned lox = (unsigned)__double2loint(Tx), hix = (unsigned)__double2hiint(Tx);
    // |c| < 2^28 (and not NaN): the exponent field of all three results is that of 2^29
    const bool expok = (((hiz - EDF_PP_HI0) | (hiy - EDF_PP_HI0) | (hix - EDF_PP_HI0)) < 0x00100000u);
    const int flz = (int)(__funnelshift_r(loz, hiz, EDF_PP_FBITS) - EDF_PP_FLBIAS);
    const int fly = (int)(__funnelshift_r(loy, hiy, EDF_PP_FBITS) - EDF_PP_FLBIAS);
    const int flx = (int)(__funnelshift_r(lox, hix, EDF_PP_FBITS) - EDF_PP_FLBIAS);
    const unsigned gqz = loz & 0x7fffffu, gqy = loy & 0x7fffffu, gqx = lox & 0x7fffffu;
    v.ez = __uint_as_float(gqz | 0x3f800000u) - 1.5f;             // exact
    v.ey = __uint_as_float(gqy | 0x3f800000u) - 1.5f;
    v.ex = __uint_as_float(gqx | 0x3f800000u) - 1.5f;
    bool near;
    if (ORDER & 1) {
        v.inr = ((unsigned)flz <= rngz) & ((unsigned)fly <= rngy) & ((unsigned)flx <= rngx);
        near = !(fmaxf(fmaxf(fabsf(v.ez), fabsf(v.ey)), fabsf(v.ex)) < 0.5f - EDF_PP_NEAR);
    } else {
        // T holds c + 0.5: floor(2c + 1) = 2 * floor + (frac >= 0.5); c in [0, len-1] <=> 1 <= that <= 2 len - 1
        const unsigned hz = 2u * (unsigned)flz + (gqz >> 22), hy = 2u * (unsigned)fly + (gqy >> 22), hx = 2u * (unsigned)flx + (gqx >> 22);
        v.inr = (hz - 1u <= rngz) & (hy - 1u <= rngy) & (hx - 1u <= rngx);
        const float qz = fabsf(fabsf(v.ez) - 0.25f), qy = fabsf(fabsf(v.ey) - 0.25f), qx = fabsf(fabsf(v.ex) - 0.25f);
        near = !(fmaxf(fmaxf(qz, qy), qx) < 0.25f - EDF_PP_NEAR);
    }
    v.slow = (gate & near) | !expok;
    if (!v.inr & !v.slow) {
        // exactly on the upper limit (un-gated integer coordinates: identity maps): in range in the reference
        const unsigned k = (ORDER & 1) ? 0u : 0x400000u;
        v.slow = ((flz == lenz - 1) & (gqz == k)) | ((fly == leny - 1) & (gqy == k)) | ((flx == lenx - 1) & (gqx == k));
    }
    v.stz = flz - ORDER / 2; v.sty = fly - ORDER / 2; v.stx = flx - ORDER / 2;
}

