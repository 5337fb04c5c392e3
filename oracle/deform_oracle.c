/*
 * deform_oracle.c -- CPU restatement ("port") of the reference algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under elasticdeform_b200/ may import,
 * link or call this file; it exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg have an independent checker that also runs
 * where /root/reference does not exist.  Parity status: PINNED -- bit-for-bit
 * against the compiled, unmodified reference (oracle/_ref, built by
 * oracle/build_ref.py) and against the committed fixtures tests/golden/*.npz
 * generated from the reference (tests/golden/make_golden.py); see
 * tests/test_oracle.py.
 *
 * It restates, in plain C without NumPy, the three C entry points of the
 * reference path (all citations: /root/reference/elasticdeform/...):
 *   orc_deform_grid           <- DeformGrid(),            deform.c:340-1043
 *   orc_spline_filter1d_grad  <- NI_SplineFilter1DGrad(), deform.c:1049-1168
 *   orc_spline_filter1d       <- scipy.ndimage.spline_filter1d(mode='mirror'),
 *                                third-party (SciPy 1.18.1 installed here;
 *                                call sites deform_grid.py:160, :168, :271);
 *                                algorithm restated from SciPy's published
 *                                ni_splines.c and pinned against the installed
 *                                SciPy bit for bit.
 * Arithmetic is double throughout in the reference's operation order; build
 * with -O2 -ffp-contract=off (x86-64 gcc emits no FMA for the reference either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXDIMS 16

enum { ORC_BOOL = 0, ORC_U8, ORC_U16, ORC_U32, ORC_U64, ORC_I8, ORC_I16, ORC_I32, ORC_I64,
       ORC_F32, ORC_F64 };
enum { ORC_NEAREST = 0, ORC_WRAP = 1, ORC_REFLECT = 2, ORC_MIRROR = 3, ORC_CONSTANT = 4 };

typedef struct {
    void*   data;
    int32_t dtype;
    int32_t ndim;
    int64_t shape[ORC_MAXDIMS];
    int64_t strides[ORC_MAXDIMS];   /* bytes */
} orc_array;

/* ---- element access (deform.c:282-338) ------------------------------------------------ */
static double get_elem(const char* p, int dt)
{
    switch (dt) {
    case ORC_BOOL: case ORC_U8: return (double)*(const unsigned char*)p;
    case ORC_U16: return (double)*(const unsigned short*)p;
    case ORC_U32: return (double)*(const unsigned int*)p;
    case ORC_U64: return (double)*(const unsigned long*)p;
    case ORC_I8:  return (double)*(const signed char*)p;
    case ORC_I16: return (double)*(const short*)p;
    case ORC_I32: return (double)*(const int*)p;
    case ORC_I64: return (double)*(const long*)p;
    case ORC_F32: return (double)*(const float*)p;
    default:      return *(const double*)p;
    }
}

#define PUT_UINT(T, MAXV) do { t = t > 0 ? t + 0.5 : 0; t = t > (MAXV) ? (MAXV) : t; \
                               t = t < 0 ? 0 : t; *(T*)p = (T)t; } while (0)
#define PUT_INT(T, MINV, MAXV) do { t = t > 0 ? t + 0.5 : t - 0.5; t = t > (MAXV) ? (MAXV) : t; \
                                    t = t < (MINV) ? (MINV) : t; *(T*)p = (T)t; } while (0)

static void put_elem(char* p, int dt, double t)        /* deform.c:906-919 */
{
    switch (dt) {
    case ORC_BOOL: *(unsigned char*)p = (unsigned char)t; break;
    case ORC_U8:  PUT_UINT(unsigned char, 255.0); break;
    case ORC_U16: PUT_UINT(unsigned short, 65535.0); break;
    case ORC_U32: PUT_UINT(unsigned int, 4294967295.0); break;
    case ORC_U64: PUT_UINT(unsigned long, 18446744073709551615.0); break;
    case ORC_I8:  PUT_INT(signed char, -128.0, 127.0); break;
    case ORC_I16: PUT_INT(short, -32768.0, 32767.0); break;
    case ORC_I32: PUT_INT(int, -2147483648.0, 2147483647.0); break;
    case ORC_I64: PUT_INT(long, -9223372036854775808.0, 9223372036854775807.0); break;
    case ORC_F32: *(float*)p = (float)t; break;
    default:      *(double*)p = t; break;
    }
}

static void add_elem(char* p, int dt, double c)        /* deform.c:309-312, :975-987 */
{
    switch (dt) {
    case ORC_BOOL: case ORC_U8: *(unsigned char*)p += (unsigned char)c; break;
    case ORC_U16: *(unsigned short*)p += (unsigned short)c; break;
    case ORC_U32: *(unsigned int*)p += (unsigned int)c; break;
    case ORC_U64: *(unsigned long*)p += (unsigned long)c; break;
    case ORC_I8:  *(signed char*)p += (signed char)c; break;
    case ORC_I16: *(short*)p += (short)c; break;
    case ORC_I32: *(int*)p += (int)c; break;
    case ORC_I64: *(long*)p += (long)c; break;
    case ORC_F32: *(float*)p += (float)c; break;
    default:      *(double*)p += c; break;
    }
}

/* ---- coordinate boundary map (deform.c:47-128) ------------------------------------------ */
static double map_coord(double in, int64_t len, int mode)
{
    if (in < 0) {
        if (mode == ORC_MIRROR) {
            if (len <= 1) in = 0;
            else {
                int64_t sz2 = 2 * len - 2;
                in = sz2 * (int64_t)(-in / sz2) + in;
                in = in <= 1 - len ? in + sz2 : -in;
            }
        } else if (mode == ORC_REFLECT) {
            if (len <= 1) in = 0;
            else {
                int64_t sz2 = 2 * len;
                if (in < -sz2) in = sz2 * (int64_t)(-in / sz2) + in;
                in = in < -len ? in + sz2 : -in - 1;
            }
        } else if (mode == ORC_WRAP) {
            if (len <= 1) in = 0;
            else {
                int64_t sz = len - 1;
                in += sz * ((int64_t)(-in / sz) + 1);
            }
        } else if (mode == ORC_NEAREST) {
            in = 0;
        } else if (mode == ORC_CONSTANT) {
            in = -1;
        }
    } else if (in > len - 1) {
        if (mode == ORC_MIRROR) {
            if (len <= 1) in = 0;
            else {
                int64_t sz2 = 2 * len - 2;
                in -= sz2 * (int64_t)(in / sz2);
                if (in >= len) in = sz2 - in;
            }
        } else if (mode == ORC_REFLECT) {
            if (len <= 1) in = 0;
            else {
                int64_t sz2 = 2 * len;
                in -= sz2 * (int64_t)(in / sz2);
                if (in >= len) in = sz2 - in - 1;
            }
        } else if (mode == ORC_WRAP) {
            if (len <= 1) in = 0;
            else {
                int64_t sz = len - 1;
                in -= sz * (int64_t)(in / sz);
            }
        } else if (mode == ORC_NEAREST) {
            in = len - 1;
        } else if (mode == ORC_CONSTANT) {
            in = -1;
        }
    }
    return in;
}

/* mirror map of a tap index (deform.c:669-683, :796-810) */
static int64_t mirror_idx(int64_t idx, int64_t len)
{
    if (len <= 1) return 0;
    int64_t s2 = 2 * len - 2;
    if (idx < 0) {
        idx = s2 * (int)(-idx / s2) + idx;
        idx = idx <= 1 - len ? idx + s2 : -idx;
    } else if (idx >= len) {
        idx -= s2 * (int)(idx / s2);
        if (idx >= len) idx = s2 - idx;
    }
    return idx;
}

/* ---- B-spline basis weights (deform.c:160-268) ----------------------------------------- */
static void bspline_weights(double x, int order, double* w)
{
    double y, z, t;
    int i;
    x -= floor(order & 1 ? x : x + 0.5);
    y = x;
    z = 1.0 - x;
    switch (order) {
    case 1:
        w[0] = 1.0 - x;
        break;
    case 2:
        w[1] = 0.75 - x * x;
        y = 0.5 - x;
        w[0] = 0.5 * y * y;
        break;
    case 3:
        w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
        w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
        w[0] = z * z * z / 6.0;
        break;
    case 4:
        t = x * x;
        w[2] = t * (t * 0.25 - 0.625) + 115.0 / 192.0;
        y = 1.0 + x;
        w[1] = y * (y * (y * (5.0 - y) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0;
        w[3] = z * (z * (z * (5.0 - z) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0;
        y = 0.5 - x;
        t = y * y;
        w[0] = t * t / 24.0;
        break;
    case 5:
        t = y * y;
        w[2] = t * (t * (0.25 - y / 12.0) - 0.5) + 0.55;
        t = z * z;
        w[3] = t * (t * (0.25 - z / 12.0) - 0.5) + 0.55;
        y += 1.0;
        w[1] = y * (y * (y * (y * (y / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425;
        z += 1.0;
        w[4] = z * (z * (z * (z * (z / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425;
        y = 1.0 - x;
        t = y * y;
        w[0] = y * t * t / 120.0;
        break;
    default:
        return;
    }
    w[order] = 1.0;
    for (i = 0; i < order; ++i) w[order] -= w[i];
}

/* ---- the per-voxel loop (deform.c:340-1043) ----------------------------------------------
 * inputs/outputs: forward  -> inputs read, outputs written
 *                 gradient -> inputs are the dX accumulators, outputs hold dY
 * axis: [ninputs*naxis]; affine: [naxis*(naxis+1)] or NULL; output_offset: [naxis] or NULL.
 * Returns 0 on success, 1 for an unsupported argument.                                      */
int orc_deform_grid(int gradient, int ninputs, const orc_array* inputs,
                    const orc_array* displacement, const int64_t* output_offset,
                    const orc_array* outputs, int naxis, const int* axis, const int* orders,
                    const int* modes, const double* cvals, const double* affine)
{
    int64_t idim[ORC_MAXDIMS], odim[ORC_MAXDIMS], ooff[ORC_MAXDIMS], ncp[ORC_MAXDIMS];
    int64_t o[ORC_MAXDIMS];
    int64_t size = 1, kk;
    int a, h, l, ii;
    double dw[ORC_MAXDIMS][4];
    int64_t dtap[ORC_MAXDIMS][4];
    double displ[ORC_MAXDIMS];
    double w[ORC_MAXDIMS][6];
    int64_t tap[ORC_MAXDIMS][6];
    int tc[ORC_MAXDIMS];

    if (naxis < 1 || naxis > ORC_MAXDIMS - 1) return 1;
    for (a = 0; a < naxis; ++a) {
        idim[a] = inputs[0].shape[axis[a]];          /* deform.c:383 */
        odim[a] = outputs[0].shape[axis[a]];         /* deform.c:384 */
        ooff[a] = output_offset ? output_offset[a] : 0;
        ncp[a] = displacement->shape[a + 1];
        size *= odim[a];
        o[a] = 0;
    }
    int64_t ndtaps = 1;
    for (a = 0; a < naxis; ++a) ndtaps *= 4;

    for (kk = 0; kk < size; ++kk) {
        /* -- displacement stage (deform.c:650-758): cubic B-spline of the control grid -- */
        for (a = 0; a < naxis; ++a) {
            double cp = (double)(ncp[a] - 1) * (double)(o[a] + ooff[a]) / (double)(idim[a] - 1);
            int64_t start = (int64_t)floor(cp) - 1;
            int edge = (start < 0 || start + 3 >= ncp[a]);
            bspline_weights(cp, 3, dw[a]);
            for (l = 0; l < 4; ++l) {
                int64_t idx = start + l;
                if (edge) idx = mirror_idx(idx, ncp[a]);
                dtap[a][l] = idx * displacement->strides[a + 1];
            }
        }
        for (h = 0; h < naxis; ++h) {
            int64_t j;
            displ[h] = 0.0;
            for (a = 0; a < naxis; ++a) tc[a] = 0;
            for (j = 0; j < ndtaps; ++j) {
                int64_t off = displacement->strides[0] * h;
                double c;
                for (a = 0; a < naxis; ++a) off += dtap[a][tc[a]];
                c = get_elem((const char*)displacement->data + off, displacement->dtype);
                for (a = 0; a < naxis; ++a) c *= dw[a][tc[a]];
                displ[h] += c;
                for (a = naxis - 1; a >= 0; --a) {
                    if (tc[a] < 3) { tc[a]++; break; }
                    tc[a] = 0;
                }
            }
        }

        /* -- per input: coordinate stage + gather / scatter (deform.c:762-1000) -- */
        for (ii = 0; ii < ninputs; ++ii) {
            const orc_array* in = &inputs[ii];
            const orc_array* out = &outputs[ii];
            const int* ax = axis + ii * naxis;
            const int order = orders[ii];
            int constant = 0;
            int64_t ntaps = 1, obase = 0, nsteps = 1, ss;
            int steprank = 0, used[ORC_MAXDIMS];
            int64_t sdim[ORC_MAXDIMS], sis[ORC_MAXDIMS], sos[ORC_MAXDIMS];

            for (h = 0; h < naxis; ++h) {
                double cc;
                if (affine) {
                    cc = 0.0;
                    for (l = 0; l < naxis; ++l) cc += affine[h * (naxis + 1) + l] * (double)o[l];
                    cc += affine[h * (naxis + 1) + naxis];
                } else {
                    cc = (double)o[h];
                }
                cc = map_coord(cc + ooff[h] + displ[h], idim[h], modes[ii]);
                if (cc > -1.0) {
                    int64_t start;
                    int edge;
                    if (order & 1) start = (int64_t)floor(cc) - order / 2;
                    else           start = (int64_t)floor(cc + 0.5) - order / 2;
                    edge = (start < 0 || start + order >= idim[h]);
                    for (l = 0; l <= order; ++l) {
                        int64_t idx = start + l;
                        if (edge) idx = mirror_idx(idx, idim[h]);
                        tap[h][l] = idx * in->strides[ax[h]];
                    }
                    bspline_weights(cc, order, w[h]);
                } else {
                    constant = 1;
                    break;
                }
            }
            for (h = 0; h < naxis; ++h) {
                ntaps *= order + 1;
                obase += o[h] * out->strides[ax[h]];
            }
            /* non-deformed axes ("steps", deform.c:405-436, :828-838) */
            for (l = 0; l < in->ndim; ++l) used[l] = 0;
            for (h = 0; h < naxis; ++h) used[ax[h]] = 1;
            for (l = 0; l < in->ndim; ++l) {
                if (used[l]) continue;
                sdim[steprank] = in->shape[l];
                sis[steprank] = in->strides[l];
                sos[steprank] = out->strides[l];
                nsteps *= in->shape[l];
                steprank++;
            }
            for (ss = 0; ss < nsteps; ++ss) {
                int64_t istep = 0, ostep = 0, r = ss, j;
                char* po;
                for (l = 0; l < steprank; ++l) {
                    istep += sis[l] * (r % sdim[l]);
                    ostep += sos[l] * (r % sdim[l]);
                    r /= sdim[l];
                }
                po = (char*)out->data + obase + ostep;
                if (!gradient) {
                    double t = 0.0;
                    if (!constant) {
                        for (h = 0; h < naxis; ++h) tc[h] = 0;
                        for (j = 0; j < ntaps; ++j) {
                            int64_t off = istep;
                            double c;
                            for (h = 0; h < naxis; ++h) off += tap[h][tc[h]];
                            c = get_elem((const char*)in->data + off, in->dtype);
                            if (order > 0)
                                for (h = 0; h < naxis; ++h) c *= w[h][tc[h]];
                            t += c;
                            for (h = naxis - 1; h >= 0; --h) {
                                if (tc[h] < order) { tc[h]++; break; }
                                tc[h] = 0;
                            }
                        }
                    } else {
                        t = cvals[ii];
                    }
                    put_elem(po, out->dtype, t);
                } else if (!constant) {
                    double g = get_elem(po, out->dtype);
                    for (h = 0; h < naxis; ++h) tc[h] = 0;
                    for (j = 0; j < ntaps; ++j) {
                        int64_t off = istep;
                        double c = g;
                        if (order > 0)
                            for (h = 0; h < naxis; ++h) c *= w[h][tc[h]];
                        for (h = 0; h < naxis; ++h) off += tap[h][tc[h]];
                        add_elem((char*)in->data + off, in->dtype, c);
                        for (h = naxis - 1; h >= 0; --h) {
                            if (tc[h] < order) { tc[h]++; break; }
                            tc[h] = 0;
                        }
                    }
                }
            }
        }
        /* next output voxel, C order over the deformed axes (NI_ITERATOR_NEXT, from_scipy.h:67) */
        for (a = naxis - 1; a >= 0; --a) {
            if (o[a] < odim[a] - 1) { o[a]++; break; }
            o[a] = 0;
        }
    }
    return 0;
}

/* ---- line filters ---------------------------------------------------------------------- */
static int filter_poles(int order, double* pole, int scipy_table)
{
    if (scipy_table) {            /* SciPy >= 1.6 ni_splines.c tabulates the poles */
        switch (order) {
        case 2: pole[0] = -0.171572875253809902396622551580603843; return 1;
        case 3: pole[0] = -0.267949192431122706472553658494127633; return 1;
        case 4: pole[0] = -0.361341225900220177092212841325675255;
                pole[1] = -0.013725429297339121360331226939128204; return 2;
        case 5: pole[0] = -0.430575347099973791851434783493520110;
                pole[1] = -0.043096288203264653822712376822550182; return 2;
        default: return 0;
        }
    }
    switch (order) {              /* deform.c:1063-1084 */
    case 2: pole[0] = sqrt(8.0) - 3.0; return 1;
    case 3: pole[0] = sqrt(3.0) - 2.0; return 1;
    case 4: pole[0] = sqrt(664.0 - sqrt(438976.0)) + sqrt(304.0) - 19.0;
            pole[1] = sqrt(664.0 + sqrt(438976.0)) - sqrt(304.0) - 19.0; return 2;
    case 5: pole[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
            pole[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5; return 2;
    default: return 0;
    }
}

static void forward_line(double* c, int64_t n, int order)
{
    double pole[2], gain = 1.0;
    int np = filter_poles(order, pole, 1), h;
    int64_t i;
    if (n <= 1 || np == 0) return;
    for (h = 0; h < np; ++h) gain *= (1.0 - pole[h]) * (1.0 - 1.0 / pole[h]);
    for (i = 0; i < n; ++i) c[i] *= gain;
    for (h = 0; h < np; ++h) {
        const double z = pole[h];
        const double zn1 = pow(z, (double)(n - 1));
        double zi = z;
        c[0] = c[0] + zn1 * c[n - 1];
        for (i = 1; i < n - 1; ++i) {
            c[0] += zi * (c[i] + zn1 * c[n - 1 - i]);
            zi *= z;
        }
        c[0] /= 1 - zn1 * zn1;
        for (i = 1; i < n; ++i) c[i] += z * c[i - 1];
        c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1);
        for (i = n - 2; i >= 0; --i) c[i] = z * (c[i + 1] - c[i]);
    }
}

static void adjoint_line(double* ln, int64_t len, int order)       /* deform.c:1116-1156 */
{
    double pole[2], weight = 1.0;
    int np = filter_poles(order, pole, 0), hh;
    int64_t ll;
    if (len <= 1) return;
    for (hh = 0; hh < np; ++hh) weight *= (1.0 - pole[hh]) * (1.0 - 1.0 / pole[hh]);
    for (hh = 0; hh < np; ++hh) {
        double p = pole[hh];
        int max = (int)ceil(log(1e-15) / log(fabs(p)));
        double sum = p * ln[0];
        ln[0] = -p * ln[0];
        for (ll = 1; ll < len - 1; ++ll) {
            sum = p * (sum + ln[ll]);
            ln[ll] = p * (ln[ll - 1] - ln[ll]);
        }
        sum = (p / (p * p - 1.0)) * (sum + ln[len - 1]);
        ln[len - 2] += p * sum;
        ln[len - 1] = sum;
        for (ll = len - 2; ll >= 0; --ll) ln[ll] += p * ln[ll + 1];
        if (max < len) {
            double zn = p;
            for (ll = 1; ll < len; ++ll) {
                ln[ll] += zn * ln[0];
                zn *= p;
            }
        } else {
            double zn = p, iz = 1.0 / p, z2n = pow(p, (double)(len - 1));
            ln[0] = ln[0] / (1.0 - z2n * z2n);
            ln[len - 1] += z2n * ln[0];
            z2n *= z2n * iz;
            for (ll = 1; ll <= len - 2; ++ll) {
                ln[ll] += (zn + z2n) * ln[0];
                zn *= p;
                z2n *= iz;
            }
        }
    }
    for (ll = 0; ll < len; ++ll) ln[ll] *= weight;
}

static void cast_elem(char* p, int dt, double v)   /* line buffer -> array: plain C cast */
{
    switch (dt) {
    case ORC_BOOL: *(unsigned char*)p = (unsigned char)v; break;
    case ORC_U8:  *(unsigned char*)p = (unsigned char)v; break;
    case ORC_U16: *(unsigned short*)p = (unsigned short)v; break;
    case ORC_U32: *(unsigned int*)p = (unsigned int)v; break;
    case ORC_U64: *(unsigned long*)p = (unsigned long)v; break;
    case ORC_I8:  *(signed char*)p = (signed char)v; break;
    case ORC_I16: *(short*)p = (short)v; break;
    case ORC_I32: *(int*)p = (int)v; break;
    case ORC_I64: *(long*)p = (long)v; break;
    case ORC_F32: *(float*)p = (float)v; break;
    default:      *(double*)p = v; break;
    }
}

static int filter_lines(const orc_array* in, const orc_array* out, int axis, int order, int adjoint)
{
    int nd = in->ndim, d;
    int64_t n, nlines = 1, line, i;
    int64_t odims[ORC_MAXDIMS], ois[ORC_MAXDIMS], oos[ORC_MAXDIMS];
    int q = 0;
    double* buf;
    if (axis < 0) axis += nd;
    if (axis < 0 || axis >= nd || order < 0 || order > 5) return 1;
    n = in->shape[axis];
    for (d = 0; d < nd; ++d) {
        if (d == axis) continue;
        odims[q] = in->shape[d]; ois[q] = in->strides[d]; oos[q] = out->strides[d];
        nlines *= in->shape[d];
        ++q;
    }
    if (n < 1 || nlines < 1) return 0;
    buf = (double*)malloc(sizeof(double) * (size_t)n);
    if (!buf) return 2;
    for (line = 0; line < nlines; ++line) {
        int64_t r = line, io = 0, oo = 0;
        for (d = q - 1; d >= 0; --d) {
            io += (r % odims[d]) * ois[d];
            oo += (r % odims[d]) * oos[d];
            r /= odims[d];
        }
        for (i = 0; i < n; ++i)
            buf[i] = get_elem((const char*)in->data + io + i * in->strides[axis], in->dtype);
        if (adjoint) adjoint_line(buf, n, order);
        else if (order > 1) forward_line(buf, n, order);
        for (i = 0; i < n; ++i)
            cast_elem((char*)out->data + oo + i * out->strides[axis], out->dtype, buf[i]);
    }
    free(buf);
    return 0;
}

int orc_spline_filter1d(const orc_array* in, const orc_array* out, int axis, int order)
{
    return filter_lines(in, out, axis, order, 0);
}

int orc_spline_filter1d_grad(const orc_array* in, const orc_array* out, int axis, int order)
{
    return filter_lines(in, out, axis, order, 1);
}
