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

#define EDF_PL_TX 32               // x positions per warp / CTA
#define EDF_PL_G 8                 // z-slabs per CTA (one warp each)
#define EDF_PL_THREADS (EDF_PL_TX * EDF_PL_G)
#define EDF_PL_RY 64               // rows a CTA walks through (table capacity)
#define EDF_PL_NC 8                // control columns the 32 lanes of a warp can touch (span + 4)
#define EDF_PL_MAXWARPS 16         // warps per CTA of the largest kernel using these tables

struct EdfPolyTables {
    double u[EDF_PL_RY];           // fractional control position of each row of the tile
    double wz[EDF_PL_G][4];        // z weights of each slab (deform.c:160-268, order 3)
    double wx[EDF_PL_TX][4];       // x weights of each lane
    double T[EDF_PL_MAXWARPS][3][4][EDF_PL_NC];   // warp-private: z-contracted control coefficients of the current y interval
    int    jy[EDF_PL_RY];          // first control row of each row's window (floor(cp) - 1)
    int    sz[EDF_PL_G];
    int    sx[EDF_PL_TX];
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
    }
    __syncthreads();
}

// Rebuild the polynomial coefficients of this thread's column for the control interval whose window starts at
// control row j0.  Warp-collective (all 32 lanes).  Returns the warp's gate: false when every control
// coefficient the warp touches is zero (then a == 0 exactly).
__device__ __forceinline__ bool edf_poly_build(const EdfParams& p, EdfPolyTables& s, int g, int lane, int j0, double* a /*[3][4]*/, int tw = -1)
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
    const EdfInputDesc& d = p.inp[ii];
    const float* __restrict__ pin = (const float*)d.in;
    float* __restrict__ pout = (float*)d.out;
    const int lenz = (int)p.idim[0], leny = (int)p.idim[1], lenx = (int)p.idim[2];
    const double limz = p.idim_m1[0], limy = p.idim_m1[1], limx = p.idim_m1[2];
    const int isz = L.istr_e[ii][0], isy = L.istr_e[ii][1];
    const int osy = L.ostr_e[ii][1];
    const int obase_zx = z * L.ostr_e[ii][0] + x * L.ostr_e[ii][2];      // element offsets fit 32 bits (host-checked)
    const float cvalf = __uint_as_float((uint32_t)L.cval_bits[ii]);
    const double bz = xadd((double)z, p.ooff_d[0]);
    const double bx = xadd((double)x, p.ooff_d[2]);
    const int nrow = min(ry, ody - y0);

    double a[12];
    int jcur = INT_MIN;
    bool gate = false;
    double by = xadd((double)y0, p.ooff_d[1]);                     // exact: integers

    int nb = U;
#pragma unroll 1
    for (int m0 = 0; m0 < nrow; m0 += nb) {
        // rows of this iteration: up to U, all inside one control interval (the only call site of the rebuild)
        const int jr = s.jy[m0];
        if (jr != jcur) {                                          // warp-uniform
            gate = edf_poly_build(p, s, g, lane, jr, a);
            jcur = jr;
        }
        nb = min(U, nrow - m0);
#pragma unroll
        for (int u = U - 1; u >= 1; --u)
            if (u < nb && s.jy[m0 + u] != jr) nb = u;
        int st[U][3];
        float fr[U][3];
        bool inr[U], slow[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int m = m0 + min(u, nb - 1);
            double dz, dy, dx;
            edf_poly_eval(a, s.u[m], dz, dy, dx);
            const double cz = xadd(bz, dz);
            const double cy = xadd(xadd(by, (double)(m - m0)), dy);
            const double cx = xadd(bx, dx);
            inr[u] = (cz >= 0.0) & (cz <= limz) & (cy >= 0.0) & (cy <= limy) & (cx >= 0.0) & (cx <= limx);
            edf_floor_split<ORDER>(cz, st[u][0], fr[u][0]);
            edf_floor_split<ORDER>(cy, st[u][1], fr[u][1]);
            edf_floor_split<ORDER>(cx, st[u][2], fr[u][2]);
            // next to a threshold (odd orders: the integers; even orders: the half-integers, and the integers 0 and
            // len-1 of the range test): redone in the reference order.  |c| >= 2^31 never passes the range test.
            float dmax;
            if (ORDER & 1) {
                dmax = fmaxf(fmaxf(fabsf(fr[u][0] - 0.5f), fabsf(fr[u][1] - 0.5f)), fabsf(fr[u][2] - 0.5f));
            } else {
                const float q0 = fabsf(fabsf(fr[u][0]) - 0.25f), q1 = fabsf(fabsf(fr[u][1]) - 0.25f), q2 = fabsf(fabsf(fr[u][2]) - 0.25f);
                dmax = 2.0f * fmaxf(fmaxf(q0, q1), q2);            // |fr| near 0 or near 0.5  <=>  | |fr| - 0.25 | near 0.25
            }
            const bool danger = gate & !(dmax < 0.5f - EDF_LEAN_EPSF);      // NaN -> danger
            // taps across the border of the volume (only exact-integer coordinates get here in 'constant' mode)
            const bool edge = ((unsigned)st[u][0] > (unsigned)(lenz - 1 - ORDER)) | ((unsigned)st[u][1] > (unsigned)(leny - 1 - ORDER)) |
                              ((unsigned)st[u][2] > (unsigned)(lenx - 1 - ORDER));
            const bool valid = tok & (u < nb);
            slow[u] = valid & (danger | (inr[u] & edge));
            inr[u] = inr[u] & !edge;
        }
        float t[U];
        if (ORDER == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = inr[u] ? st[u][0] * isz + st[u][1] * isy + st[u][2] : 0;
                t[u] = __ldg(pin + e);
            }
        } else {
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = inr[u] ? st[u][0] * isz + st[u][1] * isy + st[u][2] : 0;
                const float* b0 = pin + e;
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
                float wz[2], wy[2], wx[2];
                edf_bspline_weights_f32<1>(fr[u][0], wz);
                edf_bspline_weights_f32<1>(fr[u][1], wy);
                edf_bspline_weights_f32<1>(fr[u][2], wx);
                // same association as the other float32 kernels: x, then y, then z
                const float r00 = fmaf(v[u][1], wx[1], v[u][0] * wx[0]);
                const float r01 = fmaf(v[u][3], wx[1], v[u][2] * wx[0]);
                const float r10 = fmaf(v[u][5], wx[1], v[u][4] * wx[0]);
                const float r11 = fmaf(v[u][7], wx[1], v[u][6] * wx[0]);
                const float p0 = fmaf(r01, wy[1], r00 * wy[0]);
                const float p1 = fmaf(r11, wy[1], r10 * wy[0]);
                t[u] = fmaf(p1, wz[1], p0 * wz[0]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (tok & (u < nb)) pout[obase_zx + (y0 + m0 + u) * osy] = inr[u] ? t[u] : cvalf;
        bool anyslow = false;
#pragma unroll
        for (int u = 0; u < U; ++u) anyslow |= slow[u];
        if (anyslow) {
#pragma unroll 1
            for (int u = 0; u < U; ++u)
                if (slow[u]) edf_poly_slow_voxel<ORDER, false>(p, L, ii, z, y0 + m0 + u, x);
        }
        by = xadd(by, (double)nb);
    }
}

// order 0/1 direct kernel: conditions on top of edf_lean_eligible
static bool edf_poly_direct_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    if (!edf_lean_eligible(p, L, ii)) return false;
    const EdfInputDesc& d = p.inp[ii];
    if (d.mode != EDF_MODE_CONSTANT || d.order > 1 || p.has_affine) return false;
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
