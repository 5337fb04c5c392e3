// edf_core.h -- per-voxel arithmetic of the elastic-deformation hot path.
//
// Everything in this header is __host__ __device__ so that the very same code
// that runs in the sm_100a kernels can be compiled by g++ into the test-only
// host simulator (tests/_hostsim) and checked against the oracle without a GPU.
// The product library never executes these functions on the host.
//
// Arithmetic contract: the reference does ALL coordinate and interpolation
// arithmetic in double, in a fixed operation order, compiled without FMA
// contraction (gcc -O2, x86-64).  The "exact" routines below reproduce that
// order with explicitly rounded operations (__dmul_rn/__dadd_rn are never
// contracted by nvcc), so their results are bit-identical to the reference.
#pragma once

#include <stdint.h>
#include <math.h>

#include "../../include/edf_b200.h"

#if defined(__CUDACC__)
#define EDF_HD __host__ __device__ __forceinline__
#define EDF_D __device__ __forceinline__
// out-of-line: keeps rarely taken, division-heavy paths from being if-converted /
// speculated into the hot loop (ptxas hoists pure fp64 divisions above branches)
#define EDF_COLD __host__ __device__ __noinline__
#else
#define EDF_HD static inline
#define EDF_COLD static __attribute__((noinline))
#endif

#define EDF_MAX_STEP (EDF_MAX_DIMS - 1)

// ----------------------------------------------------------------------------
// exactly rounded fp64 ops that the compiler may not fuse
// ----------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
EDF_HD double xmul(double a, double b) { return __dmul_rn(a, b); }
EDF_HD double xadd(double a, double b) { return __dadd_rn(a, b); }
EDF_HD double xsub(double a, double b) { return __dsub_rn(a, b); }
EDF_HD double xdiv(double a, double b) { return __ddiv_rn(a, b); }
#else
// host build: compiled with -ffp-contract=off
EDF_HD double xmul(double a, double b) { return a * b; }
EDF_HD double xadd(double a, double b) { return a + b; }
EDF_HD double xsub(double a, double b) { return a - b; }
EDF_HD double xdiv(double a, double b) { return a / b; }
#endif

// ----------------------------------------------------------------------------
// kernel parameter block (flattened edf_problem; passed by value)
// ----------------------------------------------------------------------------
struct EdfInputDesc {
    char*   in;                       // forward: coefficients; gradient: dX accumulator
    char*   out;                      // forward: result;       gradient: dY
    int32_t in_dtype, out_dtype;
    int32_t order, mode;
    double  cval;
    int64_t istr[EDF_MAX_AXIS];       // byte strides of the deformed axes of `in`
    int64_t ostr[EDF_MAX_AXIS];       // byte strides of the deformed axes of `out`
    int32_t nstep_rank, pad_;
    int64_t nsteps;                   // product of the non-deformed dims
    int64_t step_dim[EDF_MAX_STEP];
    int64_t in_step_str[EDF_MAX_STEP];
    int64_t out_step_str[EDF_MAX_STEP];
};

struct EdfParams {
    int32_t naxis, ninputs, gradient, has_affine;
    int64_t size;                     // number of output voxels over the deformed axes
    int64_t idim[EDF_MAX_AXIS];       // input extent per deformed axis   (deform.c:383)
    int64_t odim[EDF_MAX_AXIS];       // output extent per deformed axis  (deform.c:384)
    int64_t ooff[EDF_MAX_AXIS];       // crop offset                      (deform.c:439-446)
    int64_t ncp[EDF_MAX_AXIS];        // control points per axis          (deform.c:449-451)
    int64_t dstr[EDF_MAX_AXIS + 1];   // byte strides of the displacement array
    const char* disp;                 // device pointer, prefiltered coefficients
    int32_t ddtype, pad_;
    double  affine[EDF_MAX_AXIS * (EDF_MAX_AXIS + 1)];
    double  idim_m1[EDF_MAX_AXIS];    // (double)(idim - 1), hoisted int64 -> double conversions
    double  ooff_d[EDF_MAX_AXIS];     // (double)ooff
    EdfInputDesc inp[EDF_MAX_INPUTS];
};

// ----------------------------------------------------------------------------
// B-spline basis weights, orders 1..5.  Same closed forms and the same
// operation order as reference deform.c:160-268 (itself SciPy's).
// ----------------------------------------------------------------------------
EDF_HD void edf_bspline_weights(double x, int order, double* w)
{
    const double fl = floor((order & 1) ? x : xadd(x, 0.5));
    x = xsub(x, fl);                                  // delta to the middle knot
    double y = x, z = xsub(1.0, x), t;
    switch (order) {
    case 1:
        w[0] = xsub(1.0, x);
        break;
    case 2:
        w[1] = xsub(0.75, xmul(x, x));
        y = xsub(0.5, x);
        w[0] = xmul(xmul(0.5, y), y);
        break;
    case 3:
        w[1] = xdiv(xadd(xmul(xmul(xmul(y, y), xsub(y, 2.0)), 3.0), 4.0), 6.0);
        w[2] = xdiv(xadd(xmul(xmul(xmul(z, z), xsub(z, 2.0)), 3.0), 4.0), 6.0);
        w[0] = xdiv(xmul(xmul(z, z), z), 6.0);
        break;
    case 4:
        t = xmul(x, x);
        w[2] = xadd(xmul(t, xsub(xmul(t, 0.25), 0.625)), 115.0 / 192.0);
        y = xadd(1.0, x);
        w[1] = xadd(xmul(y, xadd(xmul(y, xsub(xdiv(xmul(y, xsub(5.0, y)), 6.0), 1.25)),
                                 5.0 / 24.0)),
                    55.0 / 96.0);
        w[3] = xadd(xmul(z, xadd(xmul(z, xsub(xdiv(xmul(z, xsub(5.0, z)), 6.0), 1.25)),
                                 5.0 / 24.0)),
                    55.0 / 96.0);
        y = xsub(0.5, x);
        t = xmul(y, y);
        w[0] = xdiv(xmul(t, t), 24.0);
        break;
    case 5:
        t = xmul(y, y);
        w[2] = xadd(xmul(t, xsub(xmul(t, xsub(0.25, xdiv(y, 12.0))), 0.5)), 0.55);
        t = xmul(z, z);
        w[3] = xadd(xmul(t, xsub(xmul(t, xsub(0.25, xdiv(z, 12.0))), 0.5)), 0.55);
        y = xadd(y, 1.0);
        w[1] = xadd(xmul(y, xadd(xmul(y, xsub(xmul(y, xadd(xmul(y, xsub(xdiv(y, 24.0), 0.375)),
                                                           1.25)),
                                              1.75)),
                                 0.625)),
                    0.425);
        z = xadd(z, 1.0);
        w[4] = xadd(xmul(z, xadd(xmul(z, xsub(xmul(z, xadd(xmul(z, xsub(xdiv(z, 24.0), 0.375)),
                                                           1.25)),
                                              1.75)),
                                 0.625)),
                    0.425);
        y = xsub(1.0, x);
        t = xmul(y, y);
        w[0] = xdiv(xmul(xmul(y, t), t), 120.0);
        break;
    default:
        return;                                       // order 0: no weights (deform.c:257-258)
    }
    double last = 1.0;                                // deform.c:262-265
    for (int i = 0; i < order; ++i) last = xsub(last, w[i]);
    w[order] = last;
}

// ----------------------------------------------------------------------------
// coordinate boundary map, reference deform.c:47-128 (pre-1.6 SciPy semantics:
// wrap has period len-1, reflect maps (-1,0) onto (-1,0), constant -> -1).
// ----------------------------------------------------------------------------
EDF_COLD double edf_map_coordinate_cold(double in, int64_t len, int mode)
{
    if (in < 0) {
        switch (mode) {
        case EDF_MODE_MIRROR:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz2 = 2 * len - 2;
                in = xadd((double)(sz2 * (int64_t)(xdiv(-in, (double)sz2))), in);
                in = in <= (double)(1 - len) ? xadd(in, (double)sz2) : -in;
            }
            break;
        case EDF_MODE_REFLECT:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz2 = 2 * len;
                if (in < (double)(-sz2))
                    in = xadd((double)(sz2 * (int64_t)(xdiv(-in, (double)sz2))), in);
                in = in < (double)(-len) ? xadd(in, (double)sz2) : xsub(-in, 1.0);
            }
            break;
        case EDF_MODE_WRAP:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz = len - 1;
                in = xadd(in, (double)(sz * ((int64_t)(xdiv(-in, (double)sz)) + 1)));
            }
            break;
        case EDF_MODE_NEAREST:
            in = 0;
            break;
        case EDF_MODE_CONSTANT:
            in = -1;
            break;
        }
    } else if (in > (double)(len - 1)) {
        switch (mode) {
        case EDF_MODE_MIRROR:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz2 = 2 * len - 2;
                in = xsub(in, (double)(sz2 * (int64_t)(xdiv(in, (double)sz2))));
                if (in >= (double)len) in = xsub((double)sz2, in);
            }
            break;
        case EDF_MODE_REFLECT:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz2 = 2 * len;
                in = xsub(in, (double)(sz2 * (int64_t)(xdiv(in, (double)sz2))));
                if (in >= (double)len) in = xsub(xsub((double)sz2, in), 1.0);
            }
            break;
        case EDF_MODE_WRAP:
            if (len <= 1) {
                in = 0;
            } else {
                const int64_t sz = len - 1;
                in = xsub(in, (double)(sz * (int64_t)(xdiv(in, (double)sz))));
            }
            break;
        case EDF_MODE_NEAREST:
            in = (double)(len - 1);
            break;
        case EDF_MODE_CONSTANT:
            in = -1;
            break;
        }
    }
    return in;
}

// In-range coordinates pass through unchanged and 'constant' needs no arithmetic; everything
// else takes the out-of-line path above.  Same results as calling the full map directly.
EDF_HD double edf_map_coordinate(double in, int64_t len, int mode)
{
    if (in >= 0.0 && in <= (double)(len - 1)) return in;
    if (mode == EDF_MODE_CONSTANT) return (in < 0 || in > (double)(len - 1)) ? -1.0 : in;
    return edf_map_coordinate_cold(in, len, mode);
}

// Mirror map of a tap index that fell outside [0, len): reference deform.c:669-683
// and :796-810 (used for EVERY boundary mode, also for the control grid).
EDF_COLD int64_t edf_mirror_index_cold(int64_t idx, int64_t len)
{
    if (len <= 1) return 0;
    const int64_t s2 = 2 * len - 2;
    if (idx < 0) {
        idx = s2 * (int64_t)(int)(-idx / s2) + idx;
        idx = idx <= 1 - len ? idx + s2 : -idx;
    } else if (idx >= len) {
        idx -= s2 * (int64_t)(int)(idx / s2);
        if (idx >= len) idx = s2 - idx;
    }
    return idx;
}

EDF_HD int64_t edf_mirror_index(int64_t idx, int64_t len)
{
    if (idx >= 0 && idx < len) return idx;        // the map is the identity inside [0, len)
    return edf_mirror_index_cold(idx, len);
}

// ----------------------------------------------------------------------------
// element load / store with the reference's per-dtype conversion rules
// ----------------------------------------------------------------------------
EDF_HD double edf_load(const char* p, int dtype)    // CASE_INTERP_COEFF, deform.c:282-285
{
    switch (dtype) {
    case EDF_BOOL:
    case EDF_U8:  return (double)*(const uint8_t*)p;
    case EDF_U16: return (double)*(const uint16_t*)p;
    case EDF_U32: return (double)*(const uint32_t*)p;
    case EDF_U64: return (double)*(const unsigned long long*)p;
    case EDF_I8:  return (double)*(const int8_t*)p;
    case EDF_I16: return (double)*(const int16_t*)p;
    case EDF_I32: return (double)*(const int32_t*)p;
    case EDF_I64: return (double)*(const long long*)p;
    case EDF_F32: return (double)*(const float*)p;
    default:      return *(const double*)p;
    }
}

// deform.c:292-298: t>0 ? t+0.5 : 0, clamp to [0,max], truncate
EDF_HD double edf_round_uint(double t, double maxv)
{
    t = t > 0 ? xadd(t, 0.5) : 0.0;
    t = t > maxv ? maxv : t;
    t = t < 0 ? 0.0 : t;
    return t;
}
// deform.c:300-306: round half away from zero, clamp, truncate
EDF_HD double edf_round_int(double t, double minv, double maxv)
{
    t = t > 0 ? xadd(t, 0.5) : xsub(t, 0.5);
    t = t > maxv ? maxv : t;
    t = t < minv ? minv : t;
    return t;
}

EDF_HD void edf_store(char* p, int dtype, double t)  // deform.c:906-919
{
    switch (dtype) {
    case EDF_BOOL: *(uint8_t*)p  = (uint8_t)(int32_t)t; break;            // plain C cast
    case EDF_U8:   *(uint8_t*)p  = (uint8_t)(int32_t)edf_round_uint(t, 255.0); break;
    case EDF_U16:  *(uint16_t*)p = (uint16_t)(int32_t)edf_round_uint(t, 65535.0); break;
    case EDF_U32:  *(uint32_t*)p = (uint32_t)(long long)edf_round_uint(t, 4294967295.0); break;
    case EDF_U64:  *(unsigned long long*)p =
                       (unsigned long long)edf_round_uint(t, 18446744073709551615.0); break;
    case EDF_I8:   *(int8_t*)p   = (int8_t)(int32_t)edf_round_int(t, -128.0, 127.0); break;
    case EDF_I16:  *(int16_t*)p  = (int16_t)(int32_t)edf_round_int(t, -32768.0, 32767.0); break;
    case EDF_I32:  *(int32_t*)p  = (int32_t)edf_round_int(t, -2147483648.0, 2147483647.0); break;
    case EDF_I64:  *(long long*)p =
                       (long long)edf_round_int(t, -9223372036854775808.0, 9223372036854775807.0);
                   break;
    case EDF_F32:  *(float*)p  = (float)t; break;
    default:       *(double*)p = t; break;
    }
}

// dX[tap] += (T)coeff in the array dtype: CASE_INTERP_INCR, deform.c:309-312, :975-987.
// Integer types truncate toward zero before the add; on the device the add is an
// atomic (order-independent for integers, ~1 ulp order noise for floats).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void edf_atomic_add_sub32(char* p, uint32_t v, int bytes)
{
    // 8/16-bit element inside an aligned 32-bit word: CAS loop on the word
    uintptr_t a = (uintptr_t)p;
    uint32_t* word = (uint32_t*)(a & ~(uintptr_t)3);
    const int shift = (int)(a & 3) * 8;
    const uint32_t mask = (bytes == 1 ? 0xffu : 0xffffu) << shift;
    uint32_t old = *word, assumed;
    do {
        assumed = old;
        const uint32_t cur = (assumed & mask) >> shift;
        const uint32_t nv = ((cur + v) << shift) & mask;
        old = atomicCAS(word, assumed, (assumed & ~mask) | nv);
    } while (old != assumed);
}
#endif

EDF_HD void edf_accumulate(char* p, int dtype, double coeff)
{
#if defined(__CUDA_ARCH__)
    switch (dtype) {
    case EDF_BOOL:
    case EDF_U8:
    case EDF_I8:  edf_atomic_add_sub32(p, (uint32_t)(int32_t)coeff & 0xffu, 1); break;
    case EDF_U16:
    case EDF_I16: edf_atomic_add_sub32(p, (uint32_t)(int32_t)coeff & 0xffffu, 2); break;
    case EDF_U32:
    case EDF_I32: atomicAdd((unsigned int*)p, (unsigned int)(long long)coeff); break;
    case EDF_U64:
    case EDF_I64: atomicAdd((unsigned long long*)p, (unsigned long long)(long long)coeff); break;
    case EDF_F32: atomicAdd((float*)p, (float)coeff); break;
    default:      atomicAdd((double*)p, coeff); break;
    }
#else
    switch (dtype) {
    case EDF_BOOL:
    case EDF_U8:
    case EDF_I8:  *(uint8_t*)p  = (uint8_t)(*(uint8_t*)p + (uint8_t)(int32_t)coeff); break;
    case EDF_U16:
    case EDF_I16: *(uint16_t*)p = (uint16_t)(*(uint16_t*)p + (uint16_t)(int32_t)coeff); break;
    case EDF_U32:
    case EDF_I32: *(uint32_t*)p += (uint32_t)(long long)coeff; break;
    case EDF_U64:
    case EDF_I64: *(unsigned long long*)p += (unsigned long long)(long long)coeff; break;
    case EDF_F32: *(float*)p += (float)coeff; break;
    default:      *(double*)p += coeff; break;
    }
#endif
}

// ----------------------------------------------------------------------------
// displacement stage, exact reference order (deform.c:650-758)
// ----------------------------------------------------------------------------
// Control-grid position of output index o along axis a (deform.c:643/:655).
EDF_HD double edf_control_pos(const EdfParams& p, int a, int64_t o)
{
    return xdiv(xmul((double)(p.ncp[a] - 1), (double)(o + p.ooff[a])), (double)(p.idim[a] - 1));
}

template <int NAXIS>
EDF_HD void edf_displacement_exact(const EdfParams& p, const int64_t* o, double* displ)
{
    double  dw[NAXIS][4];
    int64_t doff[NAXIS][4];
#pragma unroll
    for (int a = 0; a < NAXIS; ++a) {
        const double cp = edf_control_pos(p, a, o[a]);
        const int64_t start = (int64_t)floor(cp) - 1;             // dorder = 3 (deform.c:375)
        edf_bspline_weights(cp, 3, dw[a]);
        const bool edge = start < 0 || start + 3 >= p.ncp[a];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            int64_t idx = start + l;
            if (edge) idx = edf_mirror_index(idx, p.ncp[a]);
            doff[a][l] = idx * p.dstr[a + 1];
        }
    }
    constexpr int NT = 1 << (2 * NAXIS);                          // 4^naxis taps
#pragma unroll
    for (int h = 0; h < NAXIS; ++h) {
        double sum = 0.0;
        for (int j = 0; j < NT; ++j) {                            // lexicographic, last axis fastest
            int64_t off = p.dstr[0] * h;
#pragma unroll
            for (int l = 0; l < NAXIS; ++l) off += doff[l][(j >> (2 * (NAXIS - 1 - l))) & 3];
            double c = edf_load(p.disp + off, p.ddtype);
#pragma unroll
            for (int l = 0; l < NAXIS; ++l) c = xmul(c, dw[l][(j >> (2 * (NAXIS - 1 - l))) & 3]);
            sum = xadd(sum, c);
        }
        displ[h] = sum;
    }
}

// Un-mapped source coordinate of output voxel o along axis h for one input
// (affine, crop offset, displacement): deform.c:771-781 before map_coordinate.
template <int NAXIS, typename I>
EDF_HD double edf_source_coordinate(const EdfParams& p, const I* o, int h, double displ_h)
{
    double cc;
    if (p.has_affine) {
        cc = 0.0;
#pragma unroll
        for (int l = 0; l < NAXIS; ++l)
            cc = xadd(cc, xmul(p.affine[h * (NAXIS + 1) + l], (double)o[l]));
        cc = xadd(cc, p.affine[h * (NAXIS + 1) + NAXIS]);
    } else {
        cc = (double)o[h];
    }
    return xadd(xadd(cc, (double)p.ooff[h]), displ_h);
}

// First tap index of the interpolation window (deform.c:784-788).
EDF_HD int64_t edf_window_start(double cc, int order)
{
    return (order & 1) ? (int64_t)floor(cc) - order / 2
                       : (int64_t)floor(xadd(cc, 0.5)) - order / 2;
}

// ----------------------------------------------------------------------------
// generic per-voxel routine: any naxis<=4, order 0..5, dtype, strides, mode.
// Forward is bit-identical to the reference (same operation order in fp64).
// ----------------------------------------------------------------------------
template <int NAXIS>
EDF_HD void edf_generic_voxel(const EdfParams& p, int64_t kk)
{
    int64_t o[NAXIS];
    {
        int64_t r = kk;
#pragma unroll
        for (int a = NAXIS - 1; a >= 0; --a) {
            o[a] = r % p.odim[a];
            r /= p.odim[a];
        }
    }
    double displ[NAXIS];
    edf_displacement_exact<NAXIS>(p, o, displ);

    for (int ii = 0; ii < p.ninputs; ++ii) {
        const EdfInputDesc& d = p.inp[ii];
        const int order = d.order;
        const int ntap = order + 1;
        bool constant = false;
        double  w[NAXIS][6];
        int64_t toff[NAXIS][6];
        int64_t obase = 0;
#pragma unroll
        for (int h = 0; h < NAXIS; ++h) obase += o[h] * d.ostr[h];
#pragma unroll
        for (int h = 0; h < NAXIS; ++h) {
            if (constant) continue;                                // deform.c:819-823 (break)
            double cc = edf_source_coordinate<NAXIS, int64_t>(p, o, h, displ[h]);
            cc = edf_map_coordinate(cc, p.idim[h], d.mode);
            if (cc > -1.0) {
                const int64_t start = edf_window_start(cc, order);
                const bool edge = start < 0 || start + order >= p.idim[h];
                for (int l = 0; l < ntap; ++l) {
                    int64_t idx = start + l;
                    if (edge) idx = edf_mirror_index(idx, p.idim[h]);
                    toff[h][l] = idx * d.istr[h];
                }
                if (order > 0) edf_bspline_weights(cc, order, w[h]);
            } else {
                constant = true;
            }
        }
        int ntaps_total = 1;
#pragma unroll
        for (int h = 0; h < NAXIS; ++h) ntaps_total *= ntap;

        for (int64_t ss = 0; ss < d.nsteps; ++ss) {
            int64_t istep = 0, ostep = 0, r = ss;                 // deform.c:829-838
            for (int q = 0; q < d.nstep_rank; ++q) {
                const int64_t c = r % d.step_dim[q];
                r /= d.step_dim[q];
                istep += d.in_step_str[q] * c;
                ostep += d.out_step_str[q] * c;
            }
            char* po = d.out + obase + ostep;
            if (!p.gradient) {
                double t;
                if (!constant) {
                    t = 0.0;
                    int tc[NAXIS];
#pragma unroll
                    for (int h = 0; h < NAXIS; ++h) tc[h] = 0;
                    for (int j = 0; j < ntaps_total; ++j) {       // deform.c:847-901
                        int64_t off = istep;
#pragma unroll
                        for (int h = 0; h < NAXIS; ++h) off += toff[h][tc[h]];
                        double c = edf_load(d.in + off, d.in_dtype);
                        if (order > 0) {
#pragma unroll
                            for (int h = 0; h < NAXIS; ++h) c = xmul(c, w[h][tc[h]]);
                        }
                        t = xadd(t, c);
#pragma unroll
                        for (int h = NAXIS - 1; h >= 0; --h) {    // odometer, last axis fastest
                            if (tc[h] < order) { tc[h]++; break; }
                            tc[h] = 0;
                        }
                    }
                } else {
                    t = d.cval;                                   // deform.c:903
                }
                edf_store(po, d.out_dtype, t);
            } else if (!constant) {                               // deform.c:926-996
                const double g = edf_load(po, d.out_dtype);
                int tc[NAXIS];
#pragma unroll
                for (int h = 0; h < NAXIS; ++h) tc[h] = 0;
                for (int j = 0; j < ntaps_total; ++j) {
                    double c = g;
                    if (order > 0) {
#pragma unroll
                        for (int h = 0; h < NAXIS; ++h) c = xmul(c, w[h][tc[h]]);
                    }
                    int64_t off = istep;
#pragma unroll
                    for (int h = 0; h < NAXIS; ++h) off += toff[h][tc[h]];
                    edf_accumulate(d.in + off, d.in_dtype, c);
#pragma unroll
                    for (int h = NAXIS - 1; h >= 0; --h) {
                        if (tc[h] < order) { tc[h]++; break; }
                        tc[h] = 0;
                    }
                }
            }
        }
    }
}
