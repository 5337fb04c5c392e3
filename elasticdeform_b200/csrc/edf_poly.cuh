// edf_poly.cuh -- per-thread POLYNOMIAL form of the displacement field (round 2).
//
// The kernels of round 1 evaluate the coarse-grid B-spline displacement of every output voxel from two
// shared-memory tables (A: z-contraction per CTA, B: y-contraction per 4-row chunk) with 12 fp64 FMAs and
// 12 shared loads per voxel, plus a warp-wide table contraction every chunk.  Here a thread (fixed z, x; it
// walks along y) keeps the displacement of its column as ONE CUBIC POLYNOMIAL per component in the
// fractional control position u of the row,
//     d_h(y) = a_h0 + a_h1 u + a_h2 u^2 + a_h3 u^3,        u = cp(y) - floor(cp(y)),   cp as deform.c:655
// valid while the row stays inside one control interval (64 rows for the headline configuration): per voxel
// 9 fp64 FMAs (Horner) and one broadcast load of u; the coefficients (12 doubles in registers) are rebuilt
// by the warp when the row enters the next interval: the warp contracts the control coefficients along z into
// a 3 x 4 x NC table (its z weights are warp-uniform), every lane contracts that along x with its own weights,
// and converts the four B-spline coefficients E_0..E_3 of the segment to the power basis.
//
// Accuracy: same contract as the table form (edf_fast_core.h).  The polynomial differs from the reference's
// 64-term sum by ~1e-13 (re-association); every discrete decision is kept bit-identical by re-evaluating the
// voxels whose coordinate lies within 1e-6 of a threshold in the exact reference order.  A warp whose control
// coefficients are all zero gets a == 0 exactly (identity transform: no re-evaluation needed).
#pragma once
#include "edf_swin.cuh"
#include "edf_fx.cuh"     // polynomial build, fold and fixed-point coordinates

#define EDF_PL_G 8                 // z-slabs per CTA (one warp each)
#define EDF_PL_THREADS (EDF_PL_TX * EDF_PL_G)
#define EDF_PL_RY 64               // rows a CTA walks through (table capacity)
#define EDF_PL_MAXWARPS 16         // warps per CTA of the largest kernel using these tables

struct EdfPolyTables {
    double u[EDF_PL_RY];           // fractional control position of each row of the tile
    double wz[EDF_PL_G][4];        // z weights of each slab (deform.c:160-268, order 3)
    double wx[EDF_PL_TX][4];       // x weights of each lane
    double T[EDF_PL_MAXWARPS][3][4][EDF_PL_NC];   // warp-private: z-contracted control coefficients of the current y interval
    int    jy[EDF_PL_RY];          // first control row of each row's window (floor(cp) - 1)
    int    sz[EDF_PL_G];
    int    sx[EDF_PL_TX];
    double rrat;                   // (I_y - 1) / (P_y - 1): rows per control interval
};

// CTA prologue: control tables of the tile (x0.., y0.., z0..); ends with a CTA barrier
__device__ __forceinline__ void edf_poly_tables(const EdfParams& p, EdfPolyTables& s, int z0, int y0, int x0, int ry)
{
    const int tid = threadIdx.x;
    if (tid < EDF_PL_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_PL_TX + EDF_PL_RY) {
        const int t = tid - EDF_PL_TX;
        if (t < ry) {
            const double cp = edf_control_pos(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1));
            const double fl = floor(cp);
            s.u[t] = xsub(cp, fl);
            s.jy[t] = (int)fl - 1;
        }
    } else if (tid < EDF_PL_TX + EDF_PL_RY + EDF_PL_G) {
        const int t = tid - EDF_PL_TX - EDF_PL_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    } else if (tid == EDF_PL_TX + EDF_PL_RY + EDF_PL_G) {
        s.rrat = xdiv(p.idim_m1[1], (double)(p.ncp[1] > 1 ? p.ncp[1] - 1 : 1));
    }
    __syncthreads();
}

// out-of-line form for kernels that check the interval row by row (several call sites): the coefficients come
// back through a local array that the caller copies into its registers
__device__ __noinline__ bool edf_poly_build_nl(const EdfParams& p, EdfPolyTables& s, int g, int lane, int j0, double* out)
{
    return edf_poly_build(p, s, g, lane, j0, out);
}

// displacement of the thread's column at fractional control position u (9 fp64 FMAs)
__device__ __forceinline__ void edf_poly_eval(const double* a, double u, double& dz, double& dy, double& dx)
{
    dz = fma(fma(fma(a[3], u, a[2]), u, a[1]), u, a[0]);
    dy = fma(fma(fma(a[7], u, a[6]), u, a[5]), u, a[4]);
    dx = fma(fma(fma(a[11], u, a[10]), u, a[9]), u, a[8]);
}


// Lean form for the direct kernels: the same fixed-point coordinates with the flags folded into one code,
//     >= 0 : element offset of the first tap (in range, taps inside the volume -- orders 0 / 1, 'constant' mode)
//     -1   : takes the constant value        -2 : redo in the reference order (`slow`)
// limlo*: low word of the pattern of (len - 1 [+ 0.5]) + 1.5*2^29 -- a coordinate exactly on the upper limit is in
// range in the reference (deform.c:84-86) although its floor fails the strict test; a chance match of the low word
// alone only sends a voxel to the exact routine.
template <int ORDER>
__device__ __forceinline__ int edf_fx_code(const double* a, double u, bool gate, unsigned rngz, unsigned rngy, unsigned rngx,
                                           unsigned limloz, unsigned limloy, unsigned limlox, int isz, int isy,
                                           float& ez, float& ey, float& ex)
{
    static_assert(ORDER == 0 || ORDER == 1, "taps inside the volume whenever the coordinate is strictly in range");
    const double Tz = fma(fma(fma(a[3], u, a[2]), u, a[1]), u, a[0]);
    const double Ty = fma(fma(fma(a[7], u, a[6]), u, a[5]), u, a[4]);
    const double Tx = fma(fma(fma(a[11], u, a[10]), u, a[9]), u, a[8]);
    const unsigned loz = (unsigned)__double2loint(Tz), hiz = (unsigned)__double2hiint(Tz);
    const unsigned loy = (unsigned)__double2loint(Ty), hiy = (unsigned)__double2hiint(Ty);
    const unsigned lox = (unsigned)__double2loint(Tx), hix = (unsigned)__double2hiint(Tx);
    const unsigned bad = ((hiz - EDF_PP_HI0) | (hiy - EDF_PP_HI0) | (hix - EDF_PP_HI0)) >> 20;     // != 0: |c| >= 2^28 or NaN
    const unsigned flz = __funnelshift_r(loz, hiz, EDF_PP_FBITS) - EDF_PP_FLBIAS;
    const unsigned fly = __funnelshift_r(loy, hiy, EDF_PP_FBITS) - EDF_PP_FLBIAS;
    const unsigned flx = __funnelshift_r(lox, hix, EDF_PP_FBITS) - EDF_PP_FLBIAS;
    ez = __uint_as_float((loz & 0x7fffffu) | 0x3f800000u) - 1.5f;
    ey = __uint_as_float((loy & 0x7fffffu) | 0x3f800000u) - 1.5f;
    ex = __uint_as_float((lox & 0x7fffffu) | 0x3f800000u) - 1.5f;
    bool inr, near;
    if (ORDER & 1) {
        inr = (flz <= rngz) & (fly <= rngy) & (flx <= rngx);
        near = !(fmaxf(fmaxf(fabsf(ez), fabsf(ey)), fabsf(ex)) < 0.5f - EDF_PP_NEAR);
    } else {
        const unsigned hz = __funnelshift_r(loz, hiz, EDF_PP_FBITS - 1) - 2u * EDF_PP_FLBIAS - 1u;      // floor(2c + 1) - 1
        const unsigned hy = __funnelshift_r(loy, hiy, EDF_PP_FBITS - 1) - 2u * EDF_PP_FLBIAS - 1u;
        const unsigned hx = __funnelshift_r(lox, hix, EDF_PP_FBITS - 1) - 2u * EDF_PP_FLBIAS - 1u;
        inr = (hz <= rngz) & (hy <= rngy) & (hx <= rngx);
        const float qz = fabsf(fabsf(ez) - 0.25f), qy = fabsf(fabsf(ey) - 0.25f), qx = fabsf(fabsf(ex) - 0.25f);
        near = !(fmaxf(fmaxf(qz, qy), qx) < 0.25f - EDF_PP_NEAR);
    }
    const bool slow = (gate & near) | (bad != 0u) | (loz == limloz) | (loy == limloy) | (lox == limlox);
    const int off = (int)flz * isz + (int)fly * isy + (int)flx;
    return slow ? -2 : (inr ? off : -1);
}

// interpolation weights from e = frac - 0.5 (odd orders, frac in [0,1)) or e = frac (even orders, in [-0.5,0.5))
template <int ORDER>
__device__ __forceinline__ void edf_pipe_weights(float e, float* w)
{
    if (ORDER == 1) {
        w[0] = 0.5f - e;
        w[1] = 0.5f + e;
    } else if (ORDER == 2) {
        const float a = 0.5f - e, b = 0.5f + e;
        w[0] = 0.5f * a * a;
        w[1] = fmaf(-e, e, 0.75f);
        w[2] = 0.5f * b * b;
    } else if (ORDER == 3) {
        // (0.5 -+ e)^3 / 6 and 2/3 - t^2 + t^3/2 at t = 0.5 +- e, split into even and odd parts of e
        const float e2 = e * e;
        const float A = fmaf(e2, 0.25f, 1.0f / 48.0f);
        const float B = e * fmaf(e2, 1.0f / 6.0f, 0.125f);
        const float C = fmaf(e2, -0.25f, 23.0f / 48.0f);
        const float D = e * fmaf(e2, -0.5f, 0.625f);
        w[0] = A - B; w[3] = A + B;
        w[1] = C - D; w[2] = C + D;
    }
}

// rare voxel (next to a rounding / boundary threshold, edge of the volume): exact reference-order coordinates,
// then the general single-voxel routine (any mode, mirrored edge taps)
template <int ORDER, bool GRAD>
__device__ __noinline__ void edf_poly_slow_voxel(const EdfParams& p, const EdfFastLaunch& L, int ii, int z, int y, int x)
{
    int o[3] = {z, y, x};
    double in[3];
    edf_lean_exact_coords(p, z, y, x, &in[0], &in[1], &in[2]);
    edf_fast_f32_one_input<3, ORDER, GRAD>(p, L, ii, o, in);
}

// ---------------------------------------------------------------------------------------------------------
// Forward gather, orders 0 / 1, boundary mode 'constant', straight from global memory (K1d).
// One tap (order 0) or eight (order 1) per voxel: the kernel is bound by the coordinate pipeline, so it runs
// the polynomial form with many resident warps (<= 64 registers) and U rows per iteration in branch-free
// phases (coordinates -> classification -> loads -> FMAs -> stores) so that the dependency chains overlap.
// ---------------------------------------------------------------------------------------------------------
#ifndef EDF_PL_DIRECT_MINB
#define EDF_PL_DIRECT_MINB 3
#endif
template <int ORDER, int U>
__global__ void __launch_bounds__(EDF_PL_THREADS, EDF_PL_DIRECT_MINB)
edf_poly3d_fwd_direct_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    static_assert(ORDER == 0 || ORDER == 1, "direct polynomial kernel: orders 0 and 1");
    __shared__ EdfPolyTables s;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const int ry = (int)L.rows_per_cta;
    const int x0 = blockIdx.x * EDF_PL_TX, y0 = blockIdx.y * ry, z0 = blockIdx.z * EDF_PL_G;
    edf_poly_tables(p, s, z0, y0, x0, ry);

    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    if (z >= odz) return;                                          // whole warp (no CTA barrier below)
    const bool tok = x < odx;
    const int xc = min(x, odx - 1);
    const EdfInputDesc& d = p.inp[ii];
    const float* __restrict__ pin = (const float*)d.in;
    float* __restrict__ pout = (float*)d.out;
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    // strict in-range tests on the fixed-point floor (edf_pipe_coords); in 'constant' mode an in-range voxel that is
    // not next to a threshold has all its taps inside the volume at orders 0 and 1: no edge case
    const unsigned rngz = (ORDER & 1) ? (unsigned)(lenz - 2) : (unsigned)(2 * lenz - 3);
    const unsigned rngy = (ORDER & 1) ? (unsigned)(leny - 2) : (unsigned)(2 * leny - 3);
    const unsigned rngx = (ORDER & 1) ? (unsigned)(lenx - 2) : (unsigned)(2 * lenx - 3);
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int obase_zx = z * L.ostr_e[ii][0] + xc * L.ostr_e[ii][2];     // element offsets fit 32 bits (host-checked)
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const int nrow = min(ry, ody - y0);
    const double rrat = s.rrat;
    // non-deformed "step" axis (channels sharing the displacement): element strides, 1 step when there is none
    const int64_t nsteps = d.nstep_rank == 1 ? d.nsteps : 1;
    const int64_t in_step = d.nstep_rank == 1 ? d.in_step_str[0] / 4 : 0, out_step = d.nstep_rank == 1 ? d.out_step_str[0] / 4 : 0;
    const unsigned hlf = (ORDER & 1) ? 0u : 0x400000u;
    const unsigned limloz = ((unsigned)(lenz - 1) << EDF_PP_FBITS) + hlf, limloy = ((unsigned)(leny - 1) << EDF_PP_FBITS) + hlf,
                   limlox = ((unsigned)(lenx - 1) << EDF_PP_FBITS) + hlf;

    double a[12];
    int jcur = INT_MIN;
    bool gate = false;

    int nb = U;
#pragma unroll 1
    for (int m0 = 0; m0 < nrow; m0 += nb) {
        // rows of this iteration: up to U, all inside one control interval (the only call site of the rebuild)
        const int jr = s.jy[m0];
        if (jr != jcur) {                                          // warp-uniform
            double a0[12];
            gate = edf_poly_build(p, s, g, lane, jr, a0) | (p.has_affine != 0);
            edf_poly_fold<ORDER>(p, a0, rrat, jr, z, xc, a);
            jcur = jr;
        }
        nb = min(U, nrow - m0);
#pragma unroll
        for (int u = U - 1; u >= 1; --u)
            if (u < nb && s.jy[m0 + u] != jr) nb = u;
        int e[U];
        float fz[U], fy[U], fx[U];
        int emin = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + min(u, nb - 1);
            e[u] = edf_fx_code<ORDER>(a, s.u[m], gate, rngz, rngy, rngx, limloz, limloy, limlox, isz, isy, fz[u], fy[u], fx[u]);
            emin = min(emin, e[u]);
        }
        // taps and stores, once per non-deformed step (channel): the coordinates above are shared by all of them
        // (deform.c:828-838 loops the steps inside the voxel loop as well)
        const float* pin_c = pin;
        float* pout_c = pout;
#pragma unroll 1
        for (int64_t ss = 0; ss < nsteps; ++ss, pin_c += in_step, pout_c += out_step) {
            float t[U];
            if (ORDER == 0) {
#pragma unroll
                for (int u = 0; u < U; ++u) t[u] = __ldg(pin_c + max(e[u], 0));
            } else {
                float v[U][8];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float* b0 = pin_c + max(e[u], 0);
                    const float* b1 = b0 + isy;
                    const float* b2 = b0 + isz;
                    const float* b3 = b2 + isy;
                    v[u][0] = __ldg(b0); v[u][1] = __ldg(b0 + 1);
                    v[u][2] = __ldg(b1); v[u][3] = __ldg(b1 + 1);
                    v[u][4] = __ldg(b2); v[u][5] = __ldg(b2 + 1);
                    v[u][6] = __ldg(b3); v[u][7] = __ldg(b3 + 1);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    // weights 0.5 -+ e from the centred offsets; x, then y, then z as in the other float32 kernels
                    const float wx0 = 0.5f - fx[u], wx1 = 0.5f + fx[u];
                    const float wy0 = 0.5f - fy[u], wy1 = 0.5f + fy[u];
                    const float wz0 = 0.5f - fz[u], wz1 = 0.5f + fz[u];
                    const float r00 = fmaf(v[u][1], wx1, v[u][0] * wx0);
                    const float r01 = fmaf(v[u][3], wx1, v[u][2] * wx0);
                    const float r10 = fmaf(v[u][5], wx1, v[u][4] * wx0);
                    const float r11 = fmaf(v[u][7], wx1, v[u][6] * wx0);
                    const float p0 = fmaf(r01, wy1, r00 * wy0);
                    const float p1 = fmaf(r11, wy1, r10 * wy0);
                    t[u] = fmaf(p1, wz1, p0 * wz0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (tok & (u < nb) & (e[u] != -2)) pout_c[obase_zx + (y0 + m0 + u) * osy] = e[u] >= 0 ? t[u] : cvalf;
        }
        if ((emin == -2) & tok) {
#pragma unroll 1
            for (int u = 0; u < U; ++u)
                if ((e[u] == -2) & (u < nb)) edf_poly_slow_voxel<ORDER, false>(p, L, ii, z, y0 + m0 + u, x);
        }
    }
}

// order 0/1 direct kernel: conditions on top of edf_lean_eligible
static bool edf_poly_direct_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    // 3-D, 4-byte elements (float32 at orders 0 / 1; any 4-byte type at order 0: the output is a bit copy), unit
    // stride along x, at most one non-deformed step axis (channels), 'constant' mode
    if (p.naxis != 3) return false;
    const EdfInputDesc& d = p.inp[ii];
    if (d.in_dtype != d.out_dtype || edf_elem_size(d.in_dtype) != 4) return false;
    if (d.in_dtype != EDF_F32 && d.order != 0) return false;
    if (d.nstep_rank > 1) return false;
    if (d.nstep_rank == 1 && ((d.in_step_str[0] % 4) || (d.out_step_str[0] % 4))) return false;
    if (L.istr_e[ii][2] != 1) return false;
    for (int a = 0; a < 3; ++a)
        if (p.idim[a] < 8) return false;
    if (d.mode != EDF_MODE_CONSTANT || d.order > 1) return false;
    if (p.ncp[1] < 2) return false;                                // the row index is folded into the polynomial in u
    for (int a = 0; a < 3; ++a)
        if (p.idim[a] > (1 << 26) || p.odim[a] > (1 << 26)) return false;   // fixed-point coordinates: |c| < 2^28
    if (!edf_fast_ctrl_span_ok(p, 2, EDF_PL_TX, EDF_PL_NC)) return false;
    // a control interval should span several rows, or the polynomial is rebuilt all the time
    if ((p.idim[1] - 1) < 8 * (p.ncp[1] - 1)) return false;
    if (p.odim[0] > 0x3fffffff / 2 || p.odim[1] > 0x3fffffff / 2) return false;
    static int off = -1;                                    // EDF_NO_POLY=1: round-1 kernels (A/B runs)
    if (off < 0) { const char* e = getenv("EDF_NO_POLY"); off = (e && *e && *e != '0') ? 1 : 0; }
    return !off;
}

static int edf_poly_direct_launch(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    EdfFastLaunch L = Lin;
    dim3 grid;
    grid.x = (unsigned)((p.odim[2] + EDF_PL_TX - 1) / EDF_PL_TX);
    grid.z = (unsigned)((p.odim[0] + EDF_PL_G - 1) / EDF_PL_G);
    unsigned ry = EDF_PL_RY;
    while (ry > 8 && (uint64_t)grid.x * ((p.odim[1] + ry - 1) / ry) * grid.z < 8ull * 148) ry >>= 1;
    grid.y = (unsigned)((p.odim[1] + ry - 1) / ry);
    if (grid.y > 65535u || grid.z > 65535u) return -2;
    L.rows_per_cta = ry;
    if (order == 0) edf_poly3d_fwd_direct_kernel<0, 4><<<grid, EDF_PL_THREADS, 0, st>>>(p, L, ii);
    else            edf_poly3d_fwd_direct_kernel<1, 4><<<grid, EDF_PL_THREADS, 0, st>>>(p, L, ii);
    return 0;
}
