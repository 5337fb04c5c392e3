// edf_spline_lines.h -- one-line B-spline prefilter (K3) and its adjoint (K4).
//
// Both are strictly sequential IIR recursions evaluated in double in a fixed
// operation order so that the results are bit-identical to
//   K3: scipy.ndimage.spline_filter1d(mode='mirror')   (SciPy 1.18.1, the
//       third-party dependency the reference calls at deform_grid.py:160, :168,
//       :271; algorithm restated from its published ni_splines.c and pinned by
//       tests/test_oracle.py against the installed SciPy, bit for bit);
//   K4: NI_SplineFilter1DGrad, reference deform.c:1049-1168.
#pragma once
#include "edf_core.h"

struct EdfLineFilter {
    int32_t order, npoles;
    double  pole[2];
    double  pole_pow[2];      // pole^(n-1), computed on the host with libm pow()
    double  gain;             // prod (1-z)(1-1/z)
    int32_t trunc_max[2];     // K4 only: ceil(log(1e-15)/log|z|)   (deform.c:1119)
};

// Both recursions run on blocks of EDF_LINE_BLK elements held in registers: the loads of a block are independent of
// the recursion (issued together, ahead of it), the stores follow it, so that the dependent chain per element is
// the one or two fp64 operations of the filter itself and not a shared-memory round trip (the line lives in shared
// memory; a load after a store to the same array cannot be hoisted by the compiler).  Operation order and rounding per
// element are unchanged -- same results bit for bit.
#define EDF_LINE_BLK 8

// K3: in-place on a contiguous double line c[0..n)
EDF_HD void edf_prefilter_line(double* c, int64_t n, const EdfLineFilter& f)
{
    if (n <= 1 || f.npoles == 0) return;
    for (int64_t i = 0; i < n; ++i) c[i] = xmul(c[i], f.gain);          // gain first
    for (int h = 0; h < f.npoles; ++h) {
        const double z = f.pole[h];
        const double zn1 = f.pole_pow[h];
        // causal initialisation, mirror boundary (exact finite sum)
        double c0 = xadd(c[0], xmul(zn1, c[n - 1]));
        double zi = z;
        {
            int64_t i = 1;
            for (; i + EDF_LINE_BLK <= n - 1; i += EDF_LINE_BLK) {
                double a[EDF_LINE_BLK], b[EDF_LINE_BLK];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) { a[k] = c[i + k]; b[k] = c[n - 1 - i - k]; }
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) {
                    c0 = xadd(c0, xmul(zi, xadd(a[k], xmul(zn1, b[k]))));
                    zi = xmul(zi, z);
                }
            }
            for (; i < n - 1; ++i) {
                c0 = xadd(c0, xmul(zi, xadd(c[i], xmul(zn1, c[n - 1 - i]))));
                zi = xmul(zi, z);
            }
        }
        c[0] = xdiv(c0, xsub(1.0, xmul(zn1, zn1)));
        {
            double prev = c[0];
            int64_t i = 1;
            for (; i + EDF_LINE_BLK <= n; i += EDF_LINE_BLK) {
                double a[EDF_LINE_BLK];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) a[k] = c[i + k];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) { prev = xadd(a[k], xmul(z, prev)); a[k] = prev; }
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) c[i + k] = a[k];
            }
            for (; i < n; ++i) { prev = xadd(c[i], xmul(z, prev)); c[i] = prev; }
        }
        // anti-causal initialisation, mirror boundary
        c[n - 1] = xdiv(xmul(xadd(xmul(z, c[n - 2]), c[n - 1]), z), xsub(xmul(z, z), 1.0));
        {
            double next = c[n - 1];
            int64_t i = n - 2;
            for (; i - (EDF_LINE_BLK - 1) >= 0; i -= EDF_LINE_BLK) {
                double a[EDF_LINE_BLK];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) a[k] = c[i - k];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) { next = xmul(z, xsub(next, a[k])); a[k] = next; }
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) c[i - k] = a[k];
            }
            for (; i >= 0; --i) { next = xmul(z, xsub(next, c[i])); c[i] = next; }
        }
    }
}

// K4: transpose of the (pre-1.6 SciPy) prefilter, reference deform.c:1116-1156
EDF_HD void edf_prefilter_adjoint_line(double* ln, int64_t n, const EdfLineFilter& f)
{
    if (n <= 1) return;
    for (int h = 0; h < f.npoles; ++h) {
        const double p = f.pole[h];
        double sum = xmul(p, ln[0]);
        {
            // ln[l] = p * (ln[l-1] - ln[l]) reads the UPDATED ln[l-1] (in-place loop, deform.c:1123-1126): carried in a register
            double before = xmul(-p, ln[0]);
            ln[0] = before;
            int64_t l = 1;
            for (; l + EDF_LINE_BLK <= n - 1; l += EDF_LINE_BLK) {
                double a[EDF_LINE_BLK];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) a[k] = ln[l + k];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) {
                    sum = xmul(p, xadd(sum, a[k]));
                    before = xmul(p, xsub(before, a[k]));
                    a[k] = before;
                }
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) ln[l + k] = a[k];
            }
            for (; l < n - 1; ++l) {
                const double cur = ln[l];
                sum = xmul(p, xadd(sum, cur));
                before = xmul(p, xsub(before, cur));
                ln[l] = before;
            }
        }
        sum = xmul(xdiv(p, xsub(xmul(p, p), 1.0)), xadd(sum, ln[n - 1]));
        ln[n - 2] = xadd(ln[n - 2], xmul(p, sum));
        ln[n - 1] = sum;
        {
            double next = ln[n - 1];
            int64_t l = n - 2;
            for (; l - (EDF_LINE_BLK - 1) >= 0; l -= EDF_LINE_BLK) {
                double a[EDF_LINE_BLK];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) a[k] = ln[l - k];
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) { next = xadd(a[k], xmul(p, next)); a[k] = next; }
#pragma unroll
                for (int k = 0; k < EDF_LINE_BLK; ++k) ln[l - k] = a[k];
            }
            for (; l >= 0; --l) { next = xadd(ln[l], xmul(p, next)); ln[l] = next; }
        }
        if ((int64_t)f.trunc_max[h] < n) {
            double zn = p;
            const double l0 = ln[0];
            for (int64_t l = 1; l < n; ++l) {
                ln[l] = xadd(ln[l], xmul(zn, l0));
                zn = xmul(zn, p);
            }
        } else {
            double zn = p;
            const double iz = xdiv(1.0, p);
            double z2n = f.pole_pow[h];
            ln[0] = xdiv(ln[0], xsub(1.0, xmul(z2n, z2n)));
            ln[n - 1] = xadd(ln[n - 1], xmul(z2n, ln[0]));
            z2n = xmul(z2n, xmul(z2n, iz));
            const double l0 = ln[0];
            for (int64_t l = 1; l <= n - 2; ++l) {
                ln[l] = xadd(ln[l], xmul(xadd(zn, z2n), l0));
                zn = xmul(zn, p);
                z2n = xmul(z2n, iz);
            }
        }
    }
    for (int64_t l = 0; l < n; ++l) ln[l] = xmul(ln[l], f.gain);
}

// Line buffer -> array element, C cast semantics of SciPy's NI_LineBufferToArray
// (and of the reference's copy of it, from_nd_image.c:~440-487).
EDF_HD void edf_store_cast(char* p, int dtype, double v)
{
    switch (dtype) {
    case EDF_BOOL: *(uint8_t*)p  = (uint8_t)(int32_t)v; break;
    case EDF_U8:   *(uint8_t*)p  = (uint8_t)(int32_t)v; break;
    case EDF_U16:  *(uint16_t*)p = (uint16_t)(int32_t)v; break;
    case EDF_U32:  *(uint32_t*)p = (uint32_t)(long long)v; break;
    case EDF_U64:  *(unsigned long long*)p = (unsigned long long)v; break;
    case EDF_I8:   *(int8_t*)p   = (int8_t)(int32_t)v; break;
    case EDF_I16:  *(int16_t*)p  = (int16_t)(int32_t)v; break;
    case EDF_I32:  *(int32_t*)p  = (int32_t)v; break;
    case EDF_I64:  *(long long*)p = (long long)v; break;
    case EDF_F32:  *(float*)p  = (float)v; break;
    default:       *(double*)p = v; break;
    }
}
