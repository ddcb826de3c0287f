// tsb_granular.hpp — batched granular pitch shift with per-frame sizes (tsb_granular.cu, BASELINE cfg4).
#pragma once
#include <cuda_runtime.h>

namespace tsb {
struct GranularJob {
    const float *in;        // device float32 [lanes][in_stride]
    long long in_stride, in_len;
    float *out;             // device float32 [lanes][out_stride]: the reference's output stream
    long long out_stride, out_len;
    int lanes, channels;    // lane = stream * channels + channel; the tables are per stream
    long long frames_max;   // row length of the tables
    const long long *n_frames;  // [streams]
    const int *win, *hop, *osize, *win_id;  // [streams][frames_max]
    const long long *in_off, *out_off;
    const double *shift;
    const double *const *wtabs;  // window tables (double, the reference's libm expression), indexed by win_id
};
cudaError_t op_granular_render(const GranularJob &job, cudaStream_t st);
}  // namespace tsb
