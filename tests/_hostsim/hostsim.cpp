// hostsim.cpp -- TEST-ONLY host compilation of the device code in
// elasticdeform_b200/csrc/edf_core.h / edf_spline_lines.h / edf_fast_core.h.
//
// The product library (libedf_b200.so) never runs these loops on the CPU; this
// file exists so that `pytest -m "not gpu"` can check the exact per-voxel code
// the sm_100a kernels execute against the oracle on a box without a GPU.
// Built by tests/hostsim.py with: g++ -O2 -ffp-contract=off -fPIC -shared
#include "../../elasticdeform_b200/csrc/edf_host.h"
#include "../../elasticdeform_b200/csrc/edf_fast_core.h"
#include <stdlib.h>
#include <vector>

extern "C" const char* hostsim_last_error(void) { return g_err; }

extern "C" int hostsim_deform(const edf_problem* pr, int gradient)
{
    EdfParams p;
    int rc = flatten_problem(pr, gradient, p);
    if (rc != EDF_OK) return rc;
    for (int64_t kk = 0; kk < p.size; ++kk) {
        switch (p.naxis) {
        case 1: edf_generic_voxel<1>(p, kk); break;
        case 2: edf_generic_voxel<2>(p, kk); break;
        case 3: edf_generic_voxel<3>(p, kk); break;
        case 4: edf_generic_voxel<4>(p, kk); break;
        default: return edf_fail(EDF_ERR_RUNTIME, "unsupported number of deformed axes");
        }
    }
    return EDF_OK;
}

extern "C" int hostsim_filter(const edf_array* in, const edf_array* out, int axis, int order, int adjoint)
{
    const int nd = in->ndim;
    if (axis < 0) axis += nd;
    if (axis < 0 || axis >= nd) return edf_fail(EDF_ERR_RUNTIME, "invalid axis");
    const int64_t n = in->shape[axis];
    EdfLineFilter f;
    setup_filter(f, (order < 2) ? 0 : order, n, adjoint);
    int64_t odim[EDF_MAX_DIMS], ios[EDF_MAX_DIMS], oos[EDF_MAX_DIMS], nlines = 1;
    int q = 0;
    for (int d = 0; d < nd; ++d) {
        if (d == axis) continue;
        odim[q] = in->shape[d]; ios[q] = in->strides[d]; oos[q] = out->strides[d];
        nlines *= in->shape[d]; ++q;
    }
    if (n < 1 || nlines < 1) return 0;
    std::vector<double> buf((size_t)n);
    for (int64_t line = 0; line < nlines; ++line) {
        int64_t r = line, io = 0, oo = 0;
        for (int d = q - 1; d >= 0; --d) {
            io += (r % odim[d]) * ios[d]; oo += (r % odim[d]) * oos[d]; r /= odim[d];
        }
        for (int64_t i = 0; i < n; ++i)
            buf[i] = edf_load((const char*)in->data + io + i * in->strides[axis], in->dtype);
        if (adjoint) edf_prefilter_adjoint_line(buf.data(), n, f);
        else         edf_prefilter_line(buf.data(), n, f);
        for (int64_t i = 0; i < n; ++i)
            edf_store_cast((char*)out->data + oo + i * out->strides[axis], out->dtype, buf[i]);
    }
    return 0;
}

// Fast-path coordinate pipeline (separable displacement + danger-zone fallback),
// evaluated per voxel exactly as the specialised kernels do; returns for every output
// voxel and axis the window start and the fractional offset, and the constant flag.
extern "C" int hostsim_fast_coords(const edf_problem* pr, int input_index, int64_t* starts,
                                   float* fracs, uint8_t* constant, int64_t* n_exact)
{
    EdfParams p;
    int rc = flatten_problem(pr, 0, p);
    if (rc != EDF_OK) return rc;
    return edf_fast_coords_host(p, input_index, starts, fracs, constant, n_exact);
}

// Reference-order coordinates (exact displacement evaluation for every voxel), same outputs
// as hostsim_fast_coords: what the fast pipeline has to reproduce.
template <int NAXIS>
static void exact_coords_n(const EdfParams& p, int ii, int64_t* starts, float* fracs, uint8_t* constant)
{
    const EdfInputDesc& d = p.inp[ii];
    for (int64_t kk = 0; kk < p.size; ++kk) {
        int64_t o[NAXIS], r = kk;
        for (int a = NAXIS - 1; a >= 0; --a) { o[a] = r % p.odim[a]; r /= p.odim[a]; }
        double dd[NAXIS];
        edf_displacement_exact<NAXIS>(p, o, dd);
        bool cst = false;
        for (int h = 0; h < NAXIS; ++h) {
            int st = 0; float fr = 0.f;
            const double in = edf_source_coordinate<NAXIS, int64_t>(p, o, h, dd[h]);
            if (!cst && !edf_fast_finish(p, d.mode, d.order, h, in, &st, &fr)) cst = true;
            starts[kk * NAXIS + h] = cst ? 0 : st;
            fracs[kk * NAXIS + h] = cst ? 0.f : fr;
        }
        constant[kk] = cst ? 1 : 0;
    }
}

extern "C" int hostsim_exact_coords(const edf_problem* pr, int input_index, int64_t* starts,
                                    float* fracs, uint8_t* constant)
{
    EdfParams p;
    int rc = flatten_problem(pr, 0, p);
    if (rc != EDF_OK) return rc;
    if (p.naxis == 3) exact_coords_n<3>(p, input_index, starts, fracs, constant);
    else if (p.naxis == 2) exact_coords_n<2>(p, input_index, starts, fracs, constant);
    else return edf_fail(EDF_ERR_RUNTIME, "naxis must be 2 or 3");
    return 0;
}
