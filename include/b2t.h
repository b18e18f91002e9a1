/*
 * b2t.h -- C ABI of libb2t.so, the B200-native (sm_100a) TEASAR hot path.
 *
 * Drop-in boundary for the per-label trace of seung-lab/kimimaro (kimimaro/trace.py) and the
 * EDT that feeds it.  Every entry point replaces one native call the reference makes on that
 * path; the reference-side binding a maintainer would add is a ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - pointers named d_* are DEVICE pointers (caller-owned, e.g. tensor.data_ptr());
 *     pointers named h_* are HOST pointers.  The library allocates nothing persistent.
 *   - volumes are Fortran-ordered: loc = x + sx*(y + sy*z)  (kimimaro/intake.py:320-322,
 *     ext/skeletontricks/skeletontricks.pyx:398).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - every function returns B2T_OK (0) or a negative status; b2t_last_error() describes it.
 *   - there is NO CPU fallback: without an sm_100 device every call fails with B2T_ERR_DEVICE.
 */
#ifndef B2T_H
#define B2T_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2T_OK 0
#define B2T_ERR_ARG (-1)       /* bad argument (shape, dtype width, null pointer) */
#define B2T_ERR_DEVICE (-2)    /* no CUDA device / wrong architecture */
#define B2T_ERR_CUDA (-3)      /* a CUDA runtime call failed, see b2t_last_error() */
#define B2T_ERR_CAPACITY (-4)  /* caller-provided buffer too small */

/* library / device ------------------------------------------------------------------------- */
int b2t_version(void);                 /* e.g. 100 = 0.1.0 */
const char* b2t_last_error(void);      /* thread-local message of the last failure */
int b2t_device_check(void);            /* B2T_OK iff current device is compute capability 10.x */

/* K1  anisotropic multi-label Euclidean distance transform ------------------------------------
 * replaces  edt.edt(labels, anisotropy, black_border)            kimimaro/intake.py:174-185
 *           edt.edt(labels, anisotropy, black_border=np.all())   kimimaro/trace.py:112-117
 *           edt.edt(cc_plane, black_border=True, anisotropy=(wx,wy))  kimimaro/intake.py:565
 * d_labels: [sx,sy,sz] unsigned ints of label_bytes in {1,2,4,8}; d_out: float32 [sx,sy,sz].
 * ndim = 2 runs the x and y passes only (sz must be 1), ndim = 3 all three passes.
 * Three launches: pass x (run scan), pass y, pass z (windowed lower envelope); sqrt fused
 * into the last one.  d_out is used in place between passes; no workspace. */
int b2t_edt(const void* d_labels, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
            float wx, float wy, float wz, int black_border, int ndim, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2T_H */
