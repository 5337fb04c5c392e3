// edf_fast_core.h -- coordinate pipeline of the specialised ("fast") kernels.
//
// The reference evaluates the displacement of every output voxel as a 4^naxis-tap
// sum in a fixed order (deform.c:693-758) -- ~85 % of its run time.  The fast
// kernels use the tensor-product structure instead:
//     d_h(z,y,x) = sum_k [ sum_j ( sum_i D[h,i,j,k] wz_i(z) ) wy_j(y) ] wx_k(x)
// with the inner contractions hoisted per CTA tile (tables A, B in shared memory),
// so a voxel costs 4*naxis fp64 FMAs.  That re-association changes d_h by a few
// ulp (~1e-13), which can only matter where the result depends DIScontinuously on
// the coordinate: the floor()/rounding of the window start, the in<0 / in>len-1
// boundary tests and the branch points of the boundary maps -- all of which sit
// at integer or half-integer values of the un-mapped coordinate `in`.  Voxels
// whose `in` lies within EDF_FAST_EPS of a multiple of 0.5 on any axis are
// therefore re-evaluated with the exact reference order (edf_displacement_exact),
// which makes every discrete decision (window start, edge mapping, constant flag,
// nearest-neighbour voxel choice) bit-identical to the reference, while the
// continuous part (interpolation weights) carries only the ~1e-13 difference.
#pragma once
#include "edf_core.h"

#define EDF_FAST_EPS 2e-8          // >> separable-evaluation error for |coefficients| < 1e6

// tile geometry of the fast kernels (also used by the host simulator)
#define EDF_FAST_TX 64             // voxels along the last deformed axis per CTA
#define EDF_FAST_TY 8              // rows along the second-last axis per CTA
#define EDF_FAST_TZ 4              // slabs along the first axis per CTA (3-D only)
#define EDF_FAST_NC 12             // max control-point span per tile along y and x

EDF_HD bool edf_near_half_integer(double v)
{
    // distance of 2v to the nearest integer via the 1.5*2^52 rounding trick (two fp64 adds
    // instead of a conversion-unit rint); valid for |v| < 2^50
    const double t = xadd(v, v);
    const double r = xsub(xadd(t, 6755399441055744.0), 6755399441055744.0);
    return fabs(xsub(t, r)) < 2.0 * EDF_FAST_EPS;
}

// control-grid table entry of output index o on axis a: window start + 4 weights
EDF_HD void edf_fast_ctrl_entry(const EdfParams& p, int a, int64_t o, double* w4, int* start)
{
    const double cp = edf_control_pos(p, a, o);
    *start = (int)floor(cp) - 1;
    edf_bspline_weights(cp, 3, w4);
}

// Can the fast path's fixed-size tables hold the control points a tile touches?
// (span of window starts over T consecutive outputs, plus the 4-tap window)
EDF_HD bool edf_fast_ctrl_span_ok(const EdfParams& p, int a, int T, int NC = EDF_FAST_NC)
{
    if (p.idim[a] < 2) return false;                        // cp = x/0 in the reference
    // starts differ by at most ceil((T-1)*(P-1)/(I-1)) over a tile
    const int64_t num = (int64_t)(T - 1) * (p.ncp[a] - 1);
    const int64_t den = p.idim[a] - 1;
    const int64_t span = (num + den - 1) / den + 1;
    return span + 4 <= NC;
}

// Finish one axis for one input: boundary map, window start, fractional offset.
// Returns false when the voxel takes the constant value (deform.c:782, :819-823).
template <typename F>
EDF_HD bool edf_fast_finish(const EdfParams& p, int mode, int order, int h, double in,
                            int* start, F* frac)
{
    double cc = in;
    if (!(in >= 0.0 && in <= p.idim_m1[h])) {               // same map as edf_map_coordinate
        if (mode == EDF_MODE_CONSTANT) cc = (in < 0 || in > p.idim_m1[h]) ? -1.0 : in;
        else cc = edf_map_coordinate_cold(in, p.idim[h], mode);
    }
    if (!(cc > -1.0)) return false;
    const double fl = (order & 1) ? floor(cc) : floor(xadd(cc, 0.5));
    *start = (int)fl - order / 2;
    *frac = (F)(cc - fl);
    return true;
}

// B-spline basis weights in float from the fractional offset x (= delta to the middle
// knot, as produced by edf_fast_finish).  Same closed forms as deform.c:171-265.
template <int ORDER, typename F>
EDF_HD void edf_bspline_weights_t(F x, F* w)
{
    const F y = x, z = (F)1.0 - x;
    if (ORDER == 1) {
        w[0] = (F)1.0 - x;
    } else if (ORDER == 2) {
        w[1] = (F)0.75 - x * x;
        const F yy = (F)0.5 - x;
        w[0] = (F)0.5 * yy * yy;
    } else if (ORDER == 3) {
        w[1] = (y * y * (y - (F)2.0) * (F)3.0 + (F)4.0) * ((F)1.0 / (F)6.0);
        w[2] = (z * z * (z - (F)2.0) * (F)3.0 + (F)4.0) * ((F)1.0 / (F)6.0);
        w[0] = z * z * z * ((F)1.0 / (F)6.0);
    } else if (ORDER == 4) {
        F t = x * x;
        w[2] = t * (t * (F)0.25 - (F)0.625) + (F)115.0 / (F)192.0;
        F yy = (F)1.0 + x;
        w[1] = yy * (yy * (yy * ((F)5.0 - yy) * ((F)1.0 / (F)6.0) - (F)1.25) + (F)5.0 / (F)24.0) + (F)55.0 / (F)96.0;
        w[3] = z * (z * (z * ((F)5.0 - z) * ((F)1.0 / (F)6.0) - (F)1.25) + (F)5.0 / (F)24.0) + (F)55.0 / (F)96.0;
        yy = (F)0.5 - x;
        t = yy * yy;
        w[0] = t * t * ((F)1.0 / (F)24.0);
    } else if (ORDER == 5) {
        F t = y * y;
        w[2] = t * (t * ((F)0.25 - y * ((F)1.0 / (F)12.0)) - (F)0.5) + (F)0.55;
        t = z * z;
        w[3] = t * (t * ((F)0.25 - z * ((F)1.0 / (F)12.0)) - (F)0.5) + (F)0.55;
        F yy = y + (F)1.0;
        w[1] = yy * (yy * (yy * (yy * (yy * ((F)1.0 / (F)24.0) - (F)0.375) + (F)1.25) - (F)1.75) + (F)0.625) + (F)0.425;
        F zz = z + (F)1.0;
        w[4] = zz * (zz * (zz * (zz * (zz * ((F)1.0 / (F)24.0) - (F)0.375) + (F)1.25) - (F)1.75) + (F)0.625) + (F)0.425;
        yy = (F)1.0 - x;
        t = yy * yy;
        w[0] = yy * t * t * ((F)1.0 / (F)120.0);
    }
    if (ORDER >= 1) {
        F last = (F)1.0;
#pragma unroll
        for (int i = 0; i < ORDER; ++i) last -= w[i];
        w[ORDER] = last;
    }
}

template <int ORDER>
EDF_HD void edf_bspline_weights_f32(float x, float* w)
{
    edf_bspline_weights_t<ORDER, float>(x, w);
}

// ---------------------------------------------------------------------------------------
// Host model of the fast coordinate pipeline (used ONLY by tests/_hostsim): walks the
// output tile by tile exactly like the kernels do and reports, per voxel and axis, the
// window start / fractional offset / constant flag for input `ii`, plus how many voxels
// needed the exact re-evaluation.
// ---------------------------------------------------------------------------------------
#if !defined(__CUDACC__)
template <int NAXIS>
static int edf_fast_coords_host_n(const EdfParams& p, int ii, int64_t* starts, float* fracs,
                                  uint8_t* constant, int64_t* n_exact)
{
    const int TX = EDF_FAST_TX, TY = EDF_FAST_TY, TZ = (NAXIS == 3) ? EDF_FAST_TZ : 1;
    const int AX = NAXIS - 1, AY = NAXIS - 2, AZ = 0;
    const int64_t ox = p.odim[AX], oy = p.odim[AY], oz = (NAXIS == 3) ? p.odim[AZ] : 1;
    const EdfInputDesc& d = p.inp[ii];
    *n_exact = 0;
    for (int64_t z0 = 0; z0 < oz; z0 += TZ)
    for (int64_t y0 = 0; y0 < oy; y0 += TY)
    for (int64_t x0 = 0; x0 < ox; x0 += TX) {
        double wz[EDF_FAST_TZ][4], wy[EDF_FAST_TY][4], wx[EDF_FAST_TX][4];
        int sz[EDF_FAST_TZ], sy[EDF_FAST_TY], sx[EDF_FAST_TX];
        for (int t = 0; t < TZ; ++t) {
            if (NAXIS == 3) edf_fast_ctrl_entry(p, AZ, z0 + t < oz ? z0 + t : oz - 1, wz[t], &sz[t]);
        }
        for (int t = 0; t < TY; ++t) edf_fast_ctrl_entry(p, AY, y0 + t < oy ? y0 + t : oy - 1, wy[t], &sy[t]);
        for (int t = 0; t < TX; ++t) edf_fast_ctrl_entry(p, AX, x0 + t < ox ? x0 + t : ox - 1, wx[t], &sx[t]);
        const int sy_min = sy[0], sx_min = sx[0];
        static double A[3][EDF_FAST_TZ][EDF_FAST_NC][EDF_FAST_NC];
        static double B[3][EDF_FAST_TZ][EDF_FAST_TY][EDF_FAST_NC];
        bool allzero = true;
        for (int h = 0; h < NAXIS; ++h)
        for (int t = 0; t < TZ; ++t)
        for (int jy = 0; jy < EDF_FAST_NC; ++jy)
        for (int jx = 0; jx < EDF_FAST_NC; ++jx) {
            const int64_t my = edf_mirror_index(sy_min + jy, p.ncp[AY]);
            const int64_t mx = edf_mirror_index(sx_min + jx, p.ncp[AX]);
            double a = 0.0;
            if (NAXIS == 3) {
                for (int i = 0; i < 4; ++i) {
                    const int64_t mz = edf_mirror_index(sz[t] + i, p.ncp[AZ]);
                    const double c = edf_load(p.disp + p.dstr[0] * h + mz * p.dstr[1] + my * p.dstr[2] + mx * p.dstr[3], p.ddtype);
                    if (c != 0.0) allzero = false;
                    a = fma(c, wz[t][i], a);
                }
            } else {
                a = edf_load(p.disp + p.dstr[0] * h + my * p.dstr[1] + mx * p.dstr[2], p.ddtype);
                if (a != 0.0) allzero = false;
            }
            A[h][t][jy][jx] = a;
        }
        for (int h = 0; h < NAXIS; ++h)
        for (int t = 0; t < TZ; ++t)
        for (int ty = 0; ty < TY; ++ty)
        for (int jx = 0; jx < EDF_FAST_NC; ++jx) {
            double b = 0.0;
            for (int j = 0; j < 4; ++j) b = fma(A[h][t][sy[ty] - sy_min + j][jx], wy[ty][j], b);
            B[h][t][ty][jx] = b;
        }
        for (int t = 0; t < TZ && z0 + t < oz; ++t)
        for (int ty = 0; ty < TY && y0 + ty < oy; ++ty)
        for (int tx = 0; tx < TX && x0 + tx < ox; ++tx) {
            int64_t o[NAXIS];
            if (NAXIS == 3) o[AZ] = z0 + t;
            o[AY] = y0 + ty;
            o[AX] = x0 + tx;
            double dd[NAXIS], in[NAXIS];
            bool danger = false;
            for (int h = 0; h < NAXIS; ++h) {
                double s = 0.0;
                for (int k = 0; k < 4; ++k) s = fma(B[h][t][ty][sx[tx] - sx_min + k], wx[tx][k], s);
                dd[h] = s;
                in[h] = edf_source_coordinate<NAXIS, int64_t>(p, o, h, dd[h]);
                if (edf_near_half_integer(in[h])) danger = true;
            }
            if (danger && !allzero) {
                edf_displacement_exact<NAXIS>(p, o, dd);
                for (int h = 0; h < NAXIS; ++h) in[h] = edf_source_coordinate<NAXIS, int64_t>(p, o, h, dd[h]);
                ++*n_exact;
            }
            int64_t kk = 0;
            for (int h = 0; h < NAXIS; ++h) kk = kk * p.odim[h] + o[h];
            bool cst = false;
            for (int h = 0; h < NAXIS; ++h) {
                int st = 0; float fr = 0.f;
                if (!cst && !edf_fast_finish(p, d.mode, d.order, h, in[h], &st, &fr)) cst = true;
                starts[kk * NAXIS + h] = cst ? 0 : st;
                fracs[kk * NAXIS + h] = cst ? 0.f : fr;
            }
            constant[kk] = cst ? 1 : 0;
        }
    }
    return 0;
}

static int edf_fast_coords_host(const EdfParams& p, int ii, int64_t* starts, float* fracs,
                                uint8_t* constant, int64_t* n_exact)
{
    if (p.naxis == 3) return edf_fast_coords_host_n<3>(p, ii, starts, fracs, constant, n_exact);
    if (p.naxis == 2) return edf_fast_coords_host_n<2>(p, ii, starts, fracs, constant, n_exact);
    return -1;
}
#endif
