// tsb_ops.hpp — launchers of the generic (one reference helper = one call) device operators, tsb_ops.cu.
// All pointers are device pointers, all data double precision, layouts as the reference's Buf<T>:
// planar [channel][size]; FFTBuf = interleaved (re, im); PVBuf = interleaved (amp, freq).
#pragma once
#include <cuda_runtime.h>

namespace tsb {
cudaError_t op_fft_forward(int device, const double *in, double *out, long long rows, long long n, cudaStream_t st,
                           long long *launches);
// frees the per-device FFT work buffers (four-step intermediate, Bluestein rows)
cudaError_t op_release_scratch(int device);
cudaError_t op_frft(int device, const double *in, double *out, double *scratch_g, long long nch, long long n,
                    double exponent, cudaStream_t st, long long *launches);
cudaError_t op_window(double *data, const double *wtab, long long nch, long long size, cudaStream_t st);
cudaError_t op_pv_analyze(const double *spec, double *last_phase, double *pvdata, long long nch, long long size,
                          double hop, double rate, cudaStream_t st);
cudaError_t op_pv_synthesize(double *spec, double *phase_sum, const double *pvdata, long long nch, long long size,
                             double hop, double rate, cudaStream_t st);
cudaError_t op_pv_shift(const double *tmp, double *pvdata, const double *fn_freqs, const unsigned *off,
                        const unsigned *src, long long nch, long long size, cudaStream_t st);
cudaError_t op_gather(const float *in, long long in_stride, long long in_len, long long in_off, double *frame,
                      long long lanes, long long w, cudaStream_t st);
cudaError_t op_overlap_add(const double *frame, long long osize, long long hop, float *out, long long out_stride,
                           long long out_len, long long out_off, const float *tail_in, float *tail_out, long long tail_cap,
                           long long tail_len, long long new_tail_len, long long lanes, cudaStream_t st);
cudaError_t op_timepar_scan(const double *inc, long long inc_frames, const long long *seg_tab, int n_seg, const double *base,
                            double *sum_rows, double *carry_out, long long lanes, long long m, long long lane0,
                            long long n_lanes, cudaStream_t st, long long carry_at = -1, const int *flag_in = nullptr,
                            int *flag_out = nullptr, int epoch = 0);
cudaError_t op_add_tails(void *out, int is_f64, long long out_stride, long long out_len, const long long *pos_tab,
                         int n_tails, const double *tails, long long lanes, long long n, cudaStream_t st);
cudaError_t op_fma_peak(int is_f64, void *sink, int sms, cudaStream_t st, double *tflops);
}  // namespace tsb
