/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * Minimal stand-in for the six FFTW3 entry points that the reference's host
 * path links against (reference: gpuParallel/fft_processor_fftw.cu:135-189,
 * gpuParallel/lagrangehalfc_impl.h:9).  FFTW3 is an external library that is
 * not vendored under /root/reference and is not installed in this image, so
 * oracle/_ref links this shim instead.  Only what the reference calls is
 * declared: one r2c plan and one c2r plan of size 2N = 2048, executed in place
 * on the buffers given at plan time.
 *
 * Semantics follow the published FFTW3 definitions:
 *   r2c:  out[k] = sum_j in[j] * exp(-2*pi*i*j*k/n),   k = 0 .. n/2
 *   c2r:  out[j] = sum_k X[k]  * exp(+2*pi*i*j*k/n),   X Hermitian-extended
 *         from in[0..n/2]; unnormalised.
 */
#ifndef ORACLE_FFTW_SHIM_H
#define ORACLE_FFTW_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct fftw_shim_plan_s *fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);

/* Shim-only extension used by the oracle restatement so that it can run the
 * very same transform arithmetic on its own (thread-private) buffers. */
void fftw_shim_execute_on(const fftw_plan p, void *in, void *out);

/* Shim-only: plain complex FFT of power-of-two length h (interleaved re,im).
 * tw = h complex values exp(sign*2*pi*i*k/h); x and y are two work arrays of
 * h complex values; returns whichever of them holds the result. */
double *fftw_shim_cfft(int h, const double *tw, int inverse, double *x, double *y);

#ifdef __cplusplus
}
#endif

#endif
