// tsb_spectral.hpp — fused spectral synthesis (tsb_spectral.cu): device-side Hermitian spectrum -> inverse real FFT
// of twice the frame length -> triangular window -> overlap-add (BASELINE cfg5, hacks/spectral_bytebeat/*.cpp).
#pragma once
#include <cuda_runtime.h>

struct tsb_expr;

namespace tsb {
enum { SPEC_GEN_HASH = 0, SPEC_GEN_BUFFER = 1, SPEC_GEN_EXPR = 2 };
struct SpectralJob {
    int log2n;               // frame n = 2^log2n (8..18); the spectrum has 2n points
    int channels;
    long long n_frames, first_frame;
    long long hop;
    int generator;           // SPEC_GEN_*
    const double *spectrum;  // SPEC_GEN_BUFFER: device complex128 [n_frames][channels][n], bins 1..n-1 are read
    double *out;             // device float64 [channels][out_stride]: frame f is ADDED at out_off0 + f*hop
    long long out_stride, out_len, out_off0;
    double *state_edge;      // device float64 [channels][2] or nullptr: bins 0 and n carried into the next frame (in/out)
    int load_state;
    const void *expr;        // SPEC_GEN_EXPR: the caller's tsb_expr (include/tsb200.h), already checked by expr_check
};
cudaError_t op_spectral_render(int device, const SpectralJob &job, cudaStream_t st, long long *launches);
cudaError_t op_spectral_release(int device);
const char *expr_check(const struct ::tsb_expr *x);  // nullptr = well-formed
}  // namespace tsb
