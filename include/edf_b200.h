/*
 * edf_b200.h -- C-ABI of the B200-native elastic-deformation hot path.
 *
 * This is the drop-in boundary for the ONE path of gvtulder/elasticdeform that
 * this repository accelerates: the per-voxel loop behind
 *     elasticdeform._deform_grid.deform_grid        (reference _deform_grid.c:296-299, :307)
 *     elasticdeform._deform_grid.deform_grid_grad   (reference _deform_grid.c:301-304, :308)
 *     elasticdeform._deform_grid.spline_filter1d_grad (reference _deform_grid.c:61-92, :309)
 * plus the forward spline prefilter the reference borrows from SciPy
 * (scipy.ndimage.spline_filter1d, call sites deform_grid.py:160, :168, :271).
 *
 * Conventions (identical for every entry point):
 *   - plain pointers and sizes only; no Python, NumPy or torch types;
 *   - every `data` pointer is a DEVICE pointer (cudaMalloc / torch CUDA tensor);
 *     all small descriptor arrays (axis, orders, modes, cvals, affine,
 *     output_offset, shape, strides) are HOST memory, read during the call;
 *   - strides are in BYTES, like NumPy's (reference reads PyArray_STRIDE);
 *   - the callee borrows every buffer; nothing is allocated behind the caller's
 *     back except small descriptor scratch; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) and the
 *     call returns without synchronising;
 *   - return value 0 = success; a negative edf_status otherwise, with a
 *     human-readable message available from edf_last_error() (thread-local).
 *     The Python binding maps the codes to the exception types the reference
 *     raises (RuntimeError / ValueError / MemoryError, _deform_grid.c:43-46,
 *     :121-255, deform.c:742-746).
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point returns EDF_ERR_CUDA.
 */
#ifndef EDF_B200_H
#define EDF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDF_MAX_DIMS   8   /* max ndim of an input/output array               */
#define EDF_MAX_AXIS   4   /* max number of deformed axes (naxis)             */
#define EDF_MAX_INPUTS 8   /* max arrays sharing one displacement per call    */

/* dtype codes: the 11 distinct element types of the reference's switch
 * statements (deform.c:715-741, :863-888, :907-919). */
typedef enum {
    EDF_BOOL = 0, EDF_U8 = 1, EDF_U16 = 2, EDF_U32 = 3, EDF_U64 = 4,
    EDF_I8 = 5, EDF_I16 = 6, EDF_I32 = 7, EDF_I64 = 8,
    EDF_F32 = 9, EDF_F64 = 10
} edf_dtype;

/* boundary modes: reference from_scipy.h:38-47, deform_grid.py:440-454 */
typedef enum {
    EDF_MODE_NEAREST = 0, EDF_MODE_WRAP = 1, EDF_MODE_REFLECT = 2,
    EDF_MODE_MIRROR = 3, EDF_MODE_CONSTANT = 4
} edf_mode;

typedef enum {
    EDF_OK = 0,
    EDF_ERR_RUNTIME = -1,   /* reference raises RuntimeError                  */
    EDF_ERR_VALUE   = -2,   /* reference raises ValueError                    */
    EDF_ERR_MEMORY  = -3,   /* reference raises MemoryError                   */
    EDF_ERR_CUDA    = -4    /* CUDA runtime failure / no sm_100 device        */
} edf_status;

/* What DeformGrid() reads from a PyArrayObject (deform.c:381-436, :575-579). */
typedef struct {
    void*   data;                     /* device pointer to element [0,..,0]   */
    int32_t dtype;                    /* edf_dtype                            */
    int32_t ndim;                     /* <= EDF_MAX_DIMS                      */
    int64_t shape[EDF_MAX_DIMS];
    int64_t strides[EDF_MAX_DIMS];    /* bytes                                */
} edf_array;

/* One call of the reference's Py_DeformGrid_helper (_deform_grid.c:94-293):
 * `ninputs` arrays deformed by one displacement field.
 *
 * Forward  (edf_deform_grid):      inputs  = spline coefficient arrays (read),
 *                                  outputs = preallocated results (written).
 * Gradient (edf_deform_grid_grad): inputs  = dX accumulators, MUST be zeroed by
 *                                  the caller (deform_grid.py:243), updated
 *                                  with atomic adds;
 *                                  outputs = upstream gradients dY (read).
 */
typedef struct {
    int32_t ninputs;                  /* 1..EDF_MAX_INPUTS                    */
    int32_t naxis;                    /* 1..EDF_MAX_AXIS                      */
    const edf_array* inputs;          /* [ninputs] host array of descriptors  */
    const edf_array* outputs;         /* [ninputs]                            */
    edf_array displacement;           /* [naxis, P_0..P_{naxis-1}], F64 or F32
                                         B-spline COEFFICIENTS (already
                                         prefiltered, as in deform_grid.py:166) */
    const int64_t* output_offset;     /* [naxis] crop offset or NULL          */
    const int32_t* axis;              /* [ninputs*naxis] deformed axes, sorted */
    const int32_t* orders;            /* [ninputs] 0..5                       */
    const int32_t* modes;             /* [ninputs] edf_mode                   */
    const double*  cvals;             /* [ninputs]                            */
    const double*  affine;            /* [naxis*(naxis+1)] output->input map,
                                         row-major, or NULL                   */
    uint32_t flags;                   /* EDF_FLAG_*                           */
} edf_problem;

/* flags */
#define EDF_FLAG_FORCE_GENERIC 1u     /* debug: never take a specialised kernel */
#define EDF_FLAG_NO_WINDOW     2u     /* debug: direct gather / scatter kernels, no shared-memory window */
#define EDF_FLAG_STAGED_FWD    4u     /* debug: staged-window forward gather at every eligible order (2..5) */
#define EDF_FLAG_FIXED_WINDOW  8u     /* gradient through the fixed-size window kernel (better for very steep fields) */
#define EDF_FLAG_STAGED_ALL    16u    /* debug: staged-window gradient kernel at every spline order */
#define EDF_FLAG_STEEP         32u    /* hint: steep displacement field (rms gradient > ~0.25 voxel/voxel): the tap boxes
                                         of a chunk outgrow the staged window, prefer the direct / fixed-window kernels */

int edf_deform_grid(const edf_problem* problem, void* stream);
int edf_deform_grid_grad(const edf_problem* problem, void* stream);

/* `n` independent problems (a batch of volumes, each with its own displacement
 * / affine / crop) enqueued back to back on one stream. Replaces the Python
 * loop a reference user writes around deform_grid (README.md:117-133). */
int edf_deform_grid_batch(const edf_problem* problems, int32_t n, int32_t gradient,
                          void* stream);

/* A batch whose problems differ only in their data: `n` volumes of one shape / dtype / order / mode / crop, each
 * with its own input, output and displacement buffer and (optionally) its own affine map -- the data-augmentation
 * loop of README.md:117-133 with the per-volume host work reduced to three pointers.  `proto` describes the common
 * problem (ninputs == 1; its data pointers are ignored); in_ptrs / out_ptrs / disp_ptrs are `n` device addresses;
 * `affines` holds n * naxis*(naxis+1) doubles (output->input maps, row-major) or is NULL (then proto->affine, if any,
 * applies to every volume).  Same stream semantics as edf_deform_grid_batch. */
int edf_deform_grid_batch_uniform(const edf_problem* proto, int32_t n, int32_t gradient,
                                  const uint64_t* in_ptrs, const uint64_t* out_ptrs, const uint64_t* disp_ptrs,
                                  const double* affines, void* stream);

/* Mirror-boundary B-spline prefilter along one axis (orders 2..5; orders 0/1
 * copy), SciPy semantics: double line buffer, result cast to the output dtype.
 * Replaces scipy.ndimage.spline_filter1d at deform_grid.py:160/:168/:271.
 * `input` and `output` may alias exactly (in place) and must have equal shape. */
int edf_spline_filter1d(const edf_array* input, const edf_array* output,
                        int32_t axis, int32_t order, void* stream);

/* Adjoint of the above. Replaces NI_SplineFilter1DGrad (deform.c:1049-1168) behind
 * _deform_grid.spline_filter1d_grad (_deform_grid.c:61-92). Negative axis allowed. */
int edf_spline_filter1d_grad(const edf_array* input, const edf_array* output,
                             int32_t axis, int32_t order, void* stream);

/* Thread-local message for the last non-zero status returned on this thread. */
const char* edf_last_error(void);

/* Library / device introspection (no compute). */
int  edf_version(void);               /* major*10000 + minor*100 + patch      */
int  edf_device_ok(void);             /* 1 when device 0.. current is sm_100+ */

/* Debug/measurement: how many kernels this library has launched since load
 * (process-wide, all threads). bench.py reports the delta as gpu_launches.   */
uint64_t edf_launch_count(void);
/* Name of the kernel family chosen by the most recent edf_deform_grid* call on
 * this thread ("generic", "fast3d_o3", ...). */
const char* edf_last_kernel(void);
/* Debug/measurement: cycles per phase of the staged-window kernels, summed over warps, in builds compiled with
 * -DEDF_TILE_PROFILE (all zero otherwise); out16[15] = number of warps. Reads and resets. Synchronises. */
int edf_debug_tile_profile(uint64_t* out16);

#ifdef __cplusplus
}
#endif
#endif /* EDF_B200_H */
