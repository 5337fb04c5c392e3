// EXPERIMENT (round 1, rejected) -- not compiled into the library.
//
// Forward gather through a shared-memory window: one CTA per SM stages the source box of a chunk of
// 8 x 8 x 32 output voxels (24 x 24 x 64 cells, row pitch 64 words so that bank = x mod 32 in every row)
// and gathers the taps with LDS instead of LDG.  Bit-identical to the direct gather on the whole GPU
// parity suite, but slower on B200 for the headline field (256^3 float32, 5^3 grid):
//
//     order   sigma   direct gather   window gather
//       2       8       0.366 ms        0.550 ms
//       3       8       0.652 ms        0.787 ms   (first geometry, 8x16x32 chunks / 22x30x64 window: 0.994 ms)
//       5       8       1.832 ms        1.739 ms
//       3      16       0.892 ms        1.056 ms
//       5      16       2.726 ms        2.706 ms
//
// ncu (profiles/r1_fwdwin_experiment.md): 646 instructions per 32 voxels against 487, the staging loop and the
// three CTA barriers per chunk do not overlap with the gather at one CTA per SM, 16 % of the warps still need
// the mirrored-index form and 10 % of the taps fall back to global memory; the L1/shared pipe sits at 48 % --
// fewer wavefronts per tap (1.6 against 3.45) did not turn into time.  To use it, append this file to
// csrc/edf_lean.cuh and route edf_fast_try_launch to edf_lean_launch_fwdwin (see git history of round 1).

// =======================================================================================
// Forward gather through a shared-memory window (K1w, orders >= 2).
//
// The direct gather (edf_lean3d_fwd_kernel) is bound by L1 wavefronts: the 32 tap addresses of a warp
// load span 3.4 cache lines on average, because neighbouring voxels' y/z source indices differ under
// shear.  Here one CTA per SM (16 warps = 8 z-slabs x 2 row groups, 32 x positions) stages the source
// box of a chunk of 8 x 16 x 32 output voxels in shared memory with coalesced 16-byte loads
// (~10 staged cells per output voxel), laid out with a row pitch of 64 words so that the bank of a
// cell is x mod 32 in every row: a warp's tap load then costs ~1.4 wavefronts whatever rows its
// lanes are in.  The window origin comes from the chunk's corner voxels; voxels whose taps are not
// all inside the staged box (steeper fields than the window allows, or a volume border the mirror
// map sends elsewhere) load from global memory like the direct kernel, so the result never depends
// on the window.  Arithmetic, thresholds and re-evaluation rules are those of the direct kernel:
// the two produce bit-identical outputs.
// =======================================================================================
#define EDF_FW_TX 32
#define EDF_FW_G 8
#define EDF_FW_RG 2
#define EDF_FW_MR 4                                   // rows per warp and chunk
#define EDF_FW_CH (EDF_FW_RG * EDF_FW_MR)             // rows per chunk
#define EDF_FW_WARPS (EDF_FW_G * EDF_FW_RG)
#define EDF_FW_THREADS (EDF_FW_TX * EDF_FW_WARPS)
#define EDF_FW_NC 8
#define EDF_FW_WZ 24
#define EDF_FW_WY 24
#define EDF_FW_WX 64                                  // row pitch = window width: multiple of 32 words

struct EdfFwdWinSmem {
    double wz[EDF_FW_G][4];
    double wy[EDF_FAST_RY][4];
    double wx[EDF_FW_TX][4];
    int    sz[EDF_FW_G];
    int    sy[EDF_FAST_RY];
    int    sx[EDF_FW_TX];
    int    ny, nx, nonzero, pad_;
    int    wmin[4], wmax[4];
    double A[3][EDF_FW_G][EDF_FW_NC][EDF_FW_NC];
    double Bw[EDF_FW_WARPS][3][EDF_FW_MR][EDF_FW_NC];
    __align__(16) float win[EDF_FW_WZ * EDF_FW_WY * EDF_FW_WX];
};

template <int ORDER>
__global__ void __launch_bounds__(EDF_FW_THREADS, 1)
edf_lean3d_fwdwin_kernel(const __grid_constant__ EdfParams p, const __grid_constant__ EdfFastLaunch L, const int ii)
{
    extern __shared__ __align__(128) unsigned char edf_fw_smem_raw[];
    EdfFwdWinSmem& s = *reinterpret_cast<EdfFwdWinSmem*>(edf_fw_smem_raw);
    constexpr int NT = ORDER + 1;
    constexpr int U = 2;
    constexpr int NWIN = EDF_FW_WZ * EDF_FW_WY * EDF_FW_WX;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * EDF_FW_TX;
    const int ry = (int)L.rows_per_cta;
    const int y0 = blockIdx.y * ry;
    const int z0 = blockIdx.z * EDF_FW_G;

    // ---- prologue: control tables and the z-contraction A
    if (tid == 0) s.nonzero = 0;
    if (tid < EDF_FW_TX) {
        edf_fast_ctrl_entry(p, 2, min((int64_t)(x0 + tid), p.odim[2] - 1), s.wx[tid], &s.sx[tid]);
    } else if (tid < EDF_FW_TX + EDF_FAST_RY) {
        const int t = tid - EDF_FW_TX;
        edf_fast_ctrl_entry(p, 1, min((int64_t)(y0 + t), p.odim[1] - 1), s.wy[t], &s.sy[t]);
    } else if (tid < EDF_FW_TX + EDF_FAST_RY + EDF_FW_G) {
        const int t = tid - EDF_FW_TX - EDF_FAST_RY;
        edf_fast_ctrl_entry(p, 0, min((int64_t)(z0 + t), p.odim[0] - 1), s.wz[t], &s.sz[t]);
    }
    __syncthreads();
    {
        const int sy_min0 = s.sy[0], sx_min0 = s.sx[0];
        const int ny = s.sy[EDF_FAST_RY - 1] - sy_min0 + 4;
        const int nxx = s.sx[EDF_FW_TX - 1] - sx_min0 + 4;
        if (tid == 0) { s.ny = ny; s.nx = nxx; }
        bool nz = false;
        const int na = 3 * EDF_FW_G * ny * nxx;
        for (int e = tid; e < na; e += EDF_FW_THREADS) {
            const int jx = e % nxx;
            const int jy = (e / nxx) % ny;
            const int t = (e / (nxx * ny)) % EDF_FW_G;
            const int h = e / (nxx * ny * EDF_FW_G);
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
    const int g = warp % EDF_FW_G, rg = warp / EDF_FW_G;
    const int x = x0 + lane, z = z0 + g;
    const int odz = (int)p.odim[0], ody = (int)p.odim[1], odx = (int)p.odim[2];
    const bool tok = (x < odx) && (z < odz);
    double wx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx[k] = s.wx[lane][k];
    const int sxrel = s.sx[lane] - s.sx[0];
    const int nchunk = min(ry / EDF_FW_CH, (ody - y0 + EDF_FW_CH - 1) / EDF_FW_CH);
    const bool gate = s.nonzero != 0;
    const int nx = s.nx, sy_min = s.sy[0];
    double (*Bw)[EDF_FW_MR][EDF_FW_NC] = s.Bw[warp];

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
    const int ozmax = odz - 1 - z0 < EDF_FW_G - 1 ? odz - 1 - z0 : EDF_FW_G - 1;     // last valid slab
    const int oxmax = odx - 1 - x0 < EDF_FW_TX - 1 ? odx - 1 - x0 : EDF_FW_TX - 1;   // last valid lane

    for (int c = 0; c < nchunk; ++c) {
        const int yc0 = y0 + c * EDF_FW_CH;
        const int mlast = min(EDF_FW_CH - 1, ody - 1 - yc0);
        const int mr0 = rg * EDF_FW_MR;                            // first row of this warp in the chunk
        const int mrlast = min(EDF_FW_MR - 1, mlast - mr0);         // last valid local row (may be < 0)
        // ---- warp-private y-contraction for this warp's rows of the chunk
        {
            static_assert(EDF_FW_MR == 4, "lane mapping of the B table assumes 4 rows per warp");
            const int m = lane & 3, q = lane >> 2;
            const int row = c * EDF_FW_CH + mr0 + m;
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
        if (tid < 3) { s.wmin[tid] = 0x7fffffff; s.wmax[tid] = -0x7fffffff; }
        __syncthreads();
        // ---- source box of the chunk from its corner voxels (clamped into the volume)
        if ((g == 0 || g == ozmax) && (lane == 0 || lane == oxmax) && tok && mrlast >= 0) {
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {
                const int m = q ? mrlast : 0;
                double inz, iny, inx;
                {
                    double dz = 0.0, dy = 0.0, dx = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        dz = fma(Bw[0][m][sxrel + k], wx[k], dz);
                        dy = fma(Bw[1][m][sxrel + k], wx[k], dy);
                        dx = fma(Bw[2][m][sxrel + k], wx[k], dx);
                    }
                    const int y = yc0 + mr0 + m;
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
                const int fz_ = (int)floor(fmin(fmax(inz, 0.0), limz));
                const int fy_ = (int)floor(fmin(fmax(iny, 0.0), limy));
                const int fx_ = (int)floor(fmin(fmax(inx, 0.0), limx));
                atomicMin(&s.wmin[0], fz_); atomicMax(&s.wmax[0], fz_);
                atomicMin(&s.wmin[1], fy_); atomicMax(&s.wmax[1], fy_);
                atomicMin(&s.wmin[2], fx_); atomicMax(&s.wmax[2], fx_);
            }
        }
        __syncthreads();
        // ---- window origin: taps of the box span [min - ORDER/2, max + 1 + (ORDER+1)/2]; the slack of the
        //      window is split evenly on both sides (x: origin rounded down to a 16-byte group)
        int wz0, wy0, wx0;
        bool usewin;
        {
            const int loz = s.wmin[0] - ORDER / 2, ez = s.wmax[0] + 2 + (ORDER + 1) / 2 - loz;
            const int loy = s.wmin[1] - ORDER / 2, ey = s.wmax[1] + 2 + (ORDER + 1) / 2 - loy;
            const int lox = s.wmin[2] - ORDER / 2, ex = s.wmax[2] + 2 + (ORDER + 1) / 2 - lox;
            usewin = (s.wmin[0] <= s.wmax[0]) & (ez <= EDF_FW_WZ) & (ey <= EDF_FW_WY) & (ex + 3 <= EDF_FW_WX);
            wz0 = loz - (EDF_FW_WZ - ez) / 2;
            wy0 = loy - (EDF_FW_WY - ey) / 2;
            wx0 = (lox - (EDF_FW_WX - 3 - ex) / 2) & ~3;
        }
        // ---- stage the box: 16-byte groups, coalesced along x; cells outside the volume are not read
        if (usewin) {
            constexpr int GPR = EDF_FW_WX / 4;
            constexpr int DG = EDF_FW_THREADS % GPR, DR = EDF_FW_THREADS / GPR;
            static_assert(DG == 0 && DR + 1 < 2 * EDF_FW_WY, "group / row stepping of the staging loop");
            const int cg = tid % GPR;
            int iy = (tid / GPR) % EDF_FW_WY, iz = tid / (GPR * EDF_FW_WY);
            const int gx = wx0 + 4 * cg;
            const bool xin = (gx >= 0) & (gx + 3 < lenx);
            for (int q = tid; q < NWIN / 4; q += EDF_FW_THREADS) {
                const int gz = wz0 + iz, gy = wy0 + iy;
                if (xin & (gz >= 0) & (gz < lenz) & (gy >= 0) & (gy < leny))
                    reinterpret_cast<float4*>(s.win)[q] = __ldg(reinterpret_cast<const float4*>(pin + (gz * isz + gy * isy + gx)));
                iy += DR;
                int wrap = iy >= EDF_FW_WY;
                iy -= wrap ? EDF_FW_WY : 0;
                iz += wrap;
                wrap = iy >= EDF_FW_WY;
                iy -= wrap ? EDF_FW_WY : 0;
                iz += wrap;
            }
        }
        __syncthreads();
        // ---- gather
        if (tok) {
#pragma unroll 1
            for (int m0 = 0; m0 <= mrlast; m0 += U) {
                const int yb = yc0 + mr0 + m0;
                double inz[U], iny[U], inx[U];
                int stz[U], sty[U], stx[U];
                float fz[U], fy[U], fx[U];
                bool valid[U], cst[U], slow[U];
                bool any_ex = false;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int m = min(m0 + u, EDF_FW_MR - 1);
                    const int y = yb + u;
                    valid[u] = (m0 + u) <= mrlast;
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
                    const bool loz = !(inz[u] >= 0.0), hiz = inz[u] > limz;         // NaN counts as "low"
                    const bool loy = !(iny[u] >= 0.0), hiy = iny[u] > limy;
                    const bool lox = !(inx[u] >= 0.0), hix = inx[u] > limx;
                    double cz = loz ? 0.0 : (hiz ? limz : inz[u]);
                    double cy = loy ? 0.0 : (hiy ? limy : iny[u]);
                    double cx = lox ? 0.0 : (hix ? limx : inx[u]);
                    const bool inr = !(loz | hiz | loy | hiy | lox | hix);
                    bool mapped_danger = false, nanflag = false;
                    if (!cmode && !inr) {
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
                float t[U];
                const bool warp_ex = __any_sync(__activemask(), any_ex);
                float wzf[U][NT], wyf[U][NT], wxf[U][NT];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    edf_bspline_weights_f32<ORDER>(fz[u], wzf[u]);
                    edf_bspline_weights_f32<ORDER>(fy[u], wyf[u]);
                    edf_bspline_weights_f32<ORDER>(fx[u], wxf[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    t[u] = 0.f;
                    const bool need = valid[u] & !cst[u] & !slow[u];     // else: nothing to gather for this voxel
                    // the whole warp takes one form of the tap loop: staged box or global memory
                    if (!warp_ex) {
                        const int rz = stz[u] - wz0, ry_ = sty[u] - wy0, rx = stx[u] - wx0;
                        const bool inwin = usewin & (rz >= 0) & (rz + ORDER < EDF_FW_WZ) & (ry_ >= 0) & (ry_ + ORDER < EDF_FW_WY) &
                                           (rx >= 0) & (rx + ORDER < EDF_FW_WX);
                        if (__all_sync(__activemask(), inwin | !need)) {
                            if (need) {
                                const float* wb = s.win + ((rz * EDF_FW_WY + ry_) * EDF_FW_WX + rx);
                                float acc = 0.f;
#pragma unroll
                                for (int i = 0; i < NT; ++i) {
                                    float ti = 0.f;
#pragma unroll
                                    for (int j = 0; j < NT; ++j) {
                                        const float* r = wb + (i * EDF_FW_WY + j) * EDF_FW_WX;
                                        float tj = r[0] * wxf[u][0];
#pragma unroll
                                        for (int k = 1; k < NT; ++k) tj = fmaf(r[k], wxf[u][k], tj);
                                        ti = (j == 0) ? tj * wyf[u][0] : fmaf(tj, wyf[u][j], ti);
                                    }
                                    acc = (i == 0) ? ti * wzf[u][0] : fmaf(ti, wzf[u][i], acc);
                                }
                                t[u] = acc;
                            }
                        } else if (need) {
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
                                    ti = (j == 0) ? tj * wyf[u][0] : fmaf(tj, wyf[u][j], ti);
                                }
                                acc = (i == 0) ? ti * wzf[u][0] : fmaf(ti, wzf[u][i], acc);
                            }
                            t[u] = acc;
                        }
                    } else {
                        // a window of the warp crosses the volume border: mirrored taps
                        int rzi[NT], ryi[NT], rxi[NT];
                        bool inw = usewin;
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            rzi[i] = edf_mirror1(stz[u] + i, lenz);
                            ryi[i] = edf_mirror1(sty[u] + i, leny);
                            rxi[i] = edf_mirror1(stx[u] + i, lenx);
                            inw &= ((unsigned)(rzi[i] - wz0) < (unsigned)EDF_FW_WZ) & ((unsigned)(ryi[i] - wy0) < (unsigned)EDF_FW_WY) &
                                   ((unsigned)(rxi[i] - wx0) < (unsigned)EDF_FW_WX);
                        }
                        if (__all_sync(__activemask(), inw | !need)) {
                            if (need) {
#pragma unroll
                                for (int i = 0; i < NT; ++i) {
                                    rzi[i] = (rzi[i] - wz0) * (EDF_FW_WY * EDF_FW_WX);
                                    ryi[i] = (ryi[i] - wy0) * EDF_FW_WX;
                                    rxi[i] -= wx0;
                                }
                                float acc = 0.f;
#pragma unroll
                                for (int i = 0; i < NT; ++i) {
                                    float ti = 0.f;
#pragma unroll
                                    for (int j = 0; j < NT; ++j) {
                                        const float* r = s.win + (rzi[i] + ryi[j]);
                                        float tj = r[rxi[0]] * wxf[u][0];
#pragma unroll
                                        for (int k = 1; k < NT; ++k) tj = fmaf(r[rxi[k]], wxf[u][k], tj);
                                        ti = (j == 0) ? tj * wyf[u][0] : fmaf(tj, wyf[u][j], ti);
                                    }
                                    acc = (i == 0) ? ti * wzf[u][0] : fmaf(ti, wzf[u][i], acc);
                                }
                                t[u] = acc;
                            }
                        } else if (need) {
#pragma unroll
                            for (int i = 0; i < NT; ++i) { rzi[i] *= isz; ryi[i] *= isy; }
                            float acc = 0.f;
#pragma unroll
                            for (int i = 0; i < NT; ++i) {
                                float ti = 0.f;
#pragma unroll
                                for (int j = 0; j < NT; ++j) {
                                    const float* r = pin + (rzi[i] + ryi[j]);
                                    float tj = __ldg(r + rxi[0]) * wxf[u][0];
#pragma unroll
                                    for (int k = 1; k < NT; ++k) tj = fmaf(__ldg(r + rxi[k]), wxf[u][k], tj);
                                    ti = (j == 0) ? tj * wyf[u][0] : fmaf(tj, wyf[u][j], ti);
                                }
                                acc = (i == 0) ? ti * wzf[u][0] : fmaf(ti, wzf[u][i], acc);
                            }
                            t[u] = acc;
                        }
                    }
                }
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
        __syncthreads();                                           // the next chunk restages the window
    }
}

static bool g_fwdwin_configured = false;

// The window forward kernel needs 16-byte aligned rows of the input (vector staging), control tables
// that hold a 32-wide x span and a 64-row y span, and a grid that fills the 148 SMs at least twice.
static bool edf_fwdwin_eligible(const EdfParams& p, const EdfFastLaunch& L, int ii)
{
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("EDF_NO_FWDWIN"); disabled = (e && *e && *e != '0') ? 1 : 0; }
    if (disabled || p.gradient || p.naxis != 3) return false;
    if (p.inp[ii].order < 2) return false;
    if (!edf_fast_ctrl_span_ok(p, 2, EDF_FW_TX, EDF_FW_NC) || !edf_fast_ctrl_span_ok(p, 1, EDF_FAST_RY, EDF_FW_NC)) return false;
    if ((uintptr_t)p.inp[ii].in % 16 != 0 || L.istr_e[ii][0] % 4 != 0 || L.istr_e[ii][1] % 4 != 0 || p.idim[2] % 4 != 0) return false;
    const uint64_t gx = (uint64_t)((p.odim[2] + EDF_FW_TX - 1) / EDF_FW_TX), gz = (uint64_t)((p.odim[0] + EDF_FW_G - 1) / EDF_FW_G);
    return gx * ((p.odim[1] + EDF_FW_CH - 1) / EDF_FW_CH) * gz >= 2ull * 148;
}

static int edf_lean_launch_fwdwin(int order, cudaStream_t st, const EdfParams& p, const EdfFastLaunch& Lin, int ii)
{
    dim3 grid;
    grid.x = (unsigned)((p.odim[2] + EDF_FW_TX - 1) / EDF_FW_TX);
    grid.z = (unsigned)((p.odim[0] + EDF_FW_G - 1) / EDF_FW_G);
    unsigned ry = EDF_FAST_RY;                       // fewer rows per CTA for small volumes
    while (ry > EDF_FW_CH && (uint64_t)grid.x * ((p.odim[1] + ry - 1) / ry) * grid.z < 4ull * 148) ry >>= 1;
    grid.y = (unsigned)((p.odim[1] + ry - 1) / ry);
    EdfFastLaunch L = Lin;
    L.rows_per_cta = ry;
    const size_t smem = sizeof(EdfFwdWinSmem);
    if (!g_fwdwin_configured) {
        cudaFuncSetAttribute(edf_lean3d_fwdwin_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_lean3d_fwdwin_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_lean3d_fwdwin_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(edf_lean3d_fwdwin_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaGetLastError() != cudaSuccess) return -1;
        g_fwdwin_configured = true;
    }
    switch (order) {
    case 2: edf_lean3d_fwdwin_kernel<2><<<grid, EDF_FW_THREADS, smem, st>>>(p, L, ii); break;
    case 3: edf_lean3d_fwdwin_kernel<3><<<grid, EDF_FW_THREADS, smem, st>>>(p, L, ii); break;
    case 4: edf_lean3d_fwdwin_kernel<4><<<grid, EDF_FW_THREADS, smem, st>>>(p, L, ii); break;
    default: edf_lean3d_fwdwin_kernel<5><<<grid, EDF_FW_THREADS, smem, st>>>(p, L, ii); break;
    }
    return 0;
}
