// edf_host.h -- host-side validation and flattening shared by the C-ABI (edf_api.cu)
// and the test-only host simulator (tests/_hostsim).  No CUDA dependencies.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "edf_core.h"
#include "edf_spline_lines.h"

static thread_local char g_err[512] = "";

static int edf_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int dtype_size(int dt)
{
    switch (dt) {
    case EDF_BOOL: case EDF_U8: case EDF_I8: return 1;
    case EDF_U16: case EDF_I16: return 2;
    case EDF_U32: case EDF_I32: case EDF_F32: return 4;
    case EDF_U64: case EDF_I64: case EDF_F64: return 8;
    default: return 0;
    }
}

// ----------------------------------------------------------------------------
// problem validation + flattening (mirrors _deform_grid.c:121-255 and the
// set-up part of DeformGrid, deform.c:381-451)
// ----------------------------------------------------------------------------
static int flatten_problem(const edf_problem* pr, int gradient, EdfParams& p)
{
    if (!pr) return edf_fail(EDF_ERR_RUNTIME, "null problem");
    memset(&p, 0, sizeof(p));
    const int ninputs = pr->ninputs, naxis = pr->naxis;
    if (ninputs <= 0 || !pr->inputs || !pr->outputs)
        return edf_fail(EDF_ERR_RUNTIME, "invalid number of inputs/outputs");   // _deform_grid.c:123
    if (ninputs > EDF_MAX_INPUTS)
        return edf_fail(EDF_ERR_RUNTIME, "too many inputs (max %d)", EDF_MAX_INPUTS);
    if (naxis <= 0 || naxis > EDF_MAX_AXIS || !pr->axis)
        return edf_fail(EDF_ERR_RUNTIME, "invalid axis list");                  // _deform_grid.c:149
    if (!pr->orders || !pr->modes || !pr->cvals)
        return edf_fail(EDF_ERR_RUNTIME, "number of orders, modes, cvals must match inputs");
    p.naxis = naxis;
    p.ninputs = ninputs;
    p.gradient = gradient;

    for (int i = 0; i < ninputs; ++i) {
        const edf_array& in = pr->inputs[i];
        const edf_array& out = pr->outputs[i];
        if (in.ndim != out.ndim)
            return edf_fail(EDF_ERR_RUNTIME, "input and output dimensions should match");
        if (in.ndim < naxis || in.ndim > EDF_MAX_DIMS)
            return edf_fail(EDF_ERR_RUNTIME, "invalid axis in axis list");
        if (!dtype_size(in.dtype) || !dtype_size(out.dtype))
            return edf_fail(EDF_ERR_RUNTIME, "data type not supported");       // deform.c:889-893
        const int order = pr->orders[i];
        if (order < 0 || order > 5)
            return edf_fail(EDF_ERR_RUNTIME, "spline order not supported");
        const int mode = pr->modes[i];
        if (mode < 0 || mode > 4)
            return edf_fail(EDF_ERR_RUNTIME, "boundary mode not supported");
        if (!in.data || !out.data) {
            // empty arrays may legitimately carry null pointers
            int64_t n_in = 1, n_out = 1;
            for (int d = 0; d < in.ndim; ++d) { n_in *= in.shape[d]; n_out *= out.shape[d]; }
            if ((n_in && !in.data) || (n_out && !out.data))
                return edf_fail(EDF_ERR_VALUE, "null data pointer");
        }
        EdfInputDesc& d = p.inp[i];
        d.in = (char*)in.data;
        d.out = (char*)out.data;
        d.in_dtype = in.dtype;
        d.out_dtype = out.dtype;
        d.order = order;
        d.mode = mode;
        d.cval = pr->cvals[i];
        bool used[EDF_MAX_DIMS] = {false};
        int prev = -1;
        for (int j = 0; j < naxis; ++j) {
            const int ax = pr->axis[i * naxis + j];
            if (ax < 0 || ax >= in.ndim)
                return edf_fail(EDF_ERR_RUNTIME, "invalid axis in axis list");  // _deform_grid.c:162
            if (ax <= prev)
                return edf_fail(EDF_ERR_RUNTIME, "axis must be sorted and unique");
            prev = ax;
            used[ax] = true;
            if (in.shape[ax] != pr->inputs[0].shape[pr->axis[j]])
                return edf_fail(EDF_ERR_RUNTIME, "all inputs should have the same size");
            if (out.shape[ax] != pr->outputs[0].shape[pr->axis[j]])
                return edf_fail(EDF_ERR_RUNTIME, "all outputs should have the same size");
            d.istr[j] = in.strides[ax];
            d.ostr[j] = out.strides[ax];
        }
        d.nsteps = 1;
        int q = 0;
        for (int k = 0; k < in.ndim; ++k) {                  // deform.c:417-436
            if (used[k]) continue;
            if (in.shape[k] != out.shape[k])
                return edf_fail(EDF_ERR_RUNTIME, "input and output dimensions should match");
            d.step_dim[q] = in.shape[k];
            d.in_step_str[q] = in.strides[k];
            d.out_step_str[q] = out.strides[k];
            d.nsteps *= in.shape[k];
            ++q;
        }
        d.nstep_rank = q;
    }
    p.size = 1;
    for (int j = 0; j < naxis; ++j) {
        p.idim[j] = pr->inputs[0].shape[pr->axis[j]];
        p.odim[j] = pr->outputs[0].shape[pr->axis[j]];
        p.ooff[j] = pr->output_offset ? pr->output_offset[j] : 0;
        p.idim_m1[j] = (double)(p.idim[j] - 1);
        p.ooff_d[j] = (double)p.ooff[j];
        p.size *= p.odim[j];
    }
    const edf_array& D = pr->displacement;
    int64_t dsize = 1;
    for (int k = 0; k < D.ndim && k < EDF_MAX_DIMS; ++k) dsize *= D.shape[k];
    if (D.ndim != naxis + 1 || D.shape[0] != naxis || dsize == 0)
        return edf_fail(EDF_ERR_RUNTIME, "invalid displacement shape");         // _deform_grid.c:179
    if (D.dtype != EDF_F64 && D.dtype != EDF_F32)
        return edf_fail(EDF_ERR_RUNTIME, "displacement must be float64 or float32 coefficients");
    if (!D.data) return edf_fail(EDF_ERR_VALUE, "null displacement pointer");
    for (int j = 0; j < naxis; ++j) p.ncp[j] = D.shape[j + 1];
    for (int k = 0; k <= naxis; ++k) p.dstr[k] = D.strides[k];
    p.disp = (const char*)D.data;
    p.ddtype = D.dtype;
    p.has_affine = pr->affine ? 1 : 0;
    if (pr->affine)
        for (int k = 0; k < naxis * (naxis + 1); ++k) p.affine[k] = pr->affine[k];
    return EDF_OK;
}

static int setup_filter(EdfLineFilter& f, int order, int64_t n, int adjoint)
{
    memset(&f, 0, sizeof(f));
    f.order = order;
    if (!adjoint) {
        // SciPy >= 1.6 ni_splines.c: tabulated poles
        switch (order) {
        case 2: f.npoles = 1; f.pole[0] = -0.171572875253809902396622551580603843; break;
        case 3: f.npoles = 1; f.pole[0] = -0.267949192431122706472553658494127633; break;
        case 4: f.npoles = 2; f.pole[0] = -0.361341225900220177092212841325675255;
                              f.pole[1] = -0.013725429297339121360331226939128204; break;
        case 5: f.npoles = 2; f.pole[0] = -0.430575347099973791851434783493520110;
                              f.pole[1] = -0.043096288203264653822712376822550182; break;
        default: f.npoles = 0; break;
        }
    } else {
        // reference deform.c:1063-1084: poles from the closed forms
        switch (order) {
        case 2: f.npoles = 1; f.pole[0] = sqrt(8.0) - 3.0; break;
        case 3: f.npoles = 1; f.pole[0] = sqrt(3.0) - 2.0; break;
        case 4: f.npoles = 2;
                f.pole[0] = sqrt(664.0 - sqrt(438976.0)) + sqrt(304.0) - 19.0;
                f.pole[1] = sqrt(664.0 + sqrt(438976.0)) - sqrt(304.0) - 19.0; break;
        case 5: f.npoles = 2;
                f.pole[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
                f.pole[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5; break;
        default: f.npoles = 0; break;
        }
    }
    f.gain = 1.0;
    for (int h = 0; h < f.npoles; ++h) {
        const double z = f.pole[h];
        f.gain *= (1.0 - z) * (1.0 - 1.0 / z);
        f.pole_pow[h] = pow(z, (double)(n - 1));
        f.trunc_max[h] = (int)ceil(log(1e-15) / log(fabs(z)));
    }
    return 0;
}

