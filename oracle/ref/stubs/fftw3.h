/* Declaration-only stand-in for <fftw3.h> (FFTW 3 public API subset).
 * TEST INFRASTRUCTURE ONLY: lets the UNMODIFIED reference (tinyspec.cpp) compile in an image
 * without FFTW.  Only the five entry points the reference calls are declared
 * (rtinc/buffer.hpp:19-20, rtlib/fft.hpp:13-18); definitions live in ../fftw_shim.cpp. */
#ifndef TSREF_FFTW3_STUB_H
#define TSREF_FFTW3_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef struct tsref_fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
#ifdef __cplusplus
}
#endif
#endif
