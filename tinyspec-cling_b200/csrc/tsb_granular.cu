// tsb_granular.cu — batched granular time-domain pitch shift with a frame size, hop and output length that change
// from frame to frame (BASELINE cfg4; the algorithm of the reference's hacks/pitchshift-simple.cpp:1-11 on top of the
// scheduler's gather and float32 overlap-add, tinyspec.cpp:135-157,170-200):
//
//     window_hann(in)                                   rtlib/windowfn.hpp:1-7
//     out.resize(in.num_channels, in.size/shift)
//     for j: for (int k = j/shift; k < (j+1)/shift; k++) out[c][k] = in[c][j]      (nearest-neighbour resample;
//                                                       ascending j, the last writer of a k wins)
//     next_hop_ratio(n, 0.5/shift)
//
// One CTA per lane (a channel of a stream) walks the lane's frames in order; the per-frame quantities — frame size,
// input offset, output position, computed_hop, output length, shift — come from tables the caller filled by running
// the hack's control logic through the reference's clock (they depend on t only, not on the samples).  Everything the
// reference keeps between frames, the float32 overlap tail `atmp`, stays on chip (a ring indexed by output position).
// Bit-exact with the reference's float32 stream: same double products, same float32 additions in the same order.
#include <cuda_runtime.h>

#include "tsb_granular.hpp"
#include "tsb_math.cuh"

namespace tsb {

constexpr int GR_RING = 16384;  // >= the longest output frame (frame 8192 / shift 0.5)
constexpr int GR_THREADS = 256;

__global__ void __launch_bounds__(GR_THREADS) granular_kernel(GranularJob jb) {
    extern __shared__ __align__(16) float ring[];  // [GR_RING] overlap tail: slot = output position mod GR_RING, zero where nothing is pending
    for (int lane = blockIdx.x; lane < jb.lanes; lane += gridDim.x) {
        const int stream = lane / jb.channels;
        const long long F = jb.n_frames[stream];
        const long long row = (long long)stream * jb.frames_max;
        const float *src = jb.in + (long long)lane * jb.in_stride;
        float *dst = jb.out + (long long)lane * jb.out_stride;
        for (int i = threadIdx.x; i < GR_RING; i += GR_THREADS) ring[i] = 0.f;
        __syncthreads();
        for (long long f = 0; f < F; f++) {
            const int W = TSB_LDG(&jb.win[row + f]), hop = TSB_LDG(&jb.hop[row + f]), osize = TSB_LDG(&jb.osize[row + f]);
            const long long I = TSB_LDG(&jb.in_off[row + f]), P = TSB_LDG(&jb.out_off[row + f]);
            const double shift = TSB_LDG(&jb.shift[row + f]);
            const double *wtab = jb.wtabs[TSB_LDG(&jb.win_id[row + f])];
            const int span = hop > osize ? hop : osize;
            for (int k = threadIdx.x; k < span; k += GR_THREADS) {
                double v = 0.0;
                if (k < osize) {
                    // the last j (ascending) that writes out[k]: the largest j < W with (int)(j/shift) <= k
                    int j = (int)(((double)k + 1.0) * shift) + 2;
                    j = j > W - 1 ? W - 1 : (j < 0 ? 0 : j);
                    while (j > 0 && (int)__ddiv_rn((double)j, shift) > k) j--;
                    while (j + 1 < W && (int)__ddiv_rn((double)(j + 1), shift) <= k) j++;
                    const long long s = I + j;
                    const double x = (s >= 0 && s < jb.in_len) ? (double)src[s] : 0.0;
                    v = __dmul_rn(x, __ldg(&wtab[j]));  // window_hann: x[c][i] *= w
                }
                float *slot = &ring[(int)((P + k) & (GR_RING - 1))];
                const float tail = *slot;
                if (k < hop) {
                    // tinyspec.cpp:173-191: the finished sample; past the frame's end only what was pending
                    const float o = k < osize ? (float)__dadd_rn((double)tail, v) : tail;
                    if (P + k < jb.out_len) dst[P + k] = o;
                    *slot = 0.f;
                } else {
                    *slot = __fadd_rn(tail, (float)v);  // tinyspec.cpp:194-200 (a fresh slot holds 0: 0 + x = x)
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

cudaError_t op_granular_render(const GranularJob &jb, cudaStream_t st) {
    if (jb.lanes <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(granular_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_RING * (int)sizeof(float));
    if (e != cudaSuccess) return e;
    granular_kernel<<<(unsigned)jb.lanes, GR_THREADS, GR_RING * sizeof(float), st>>>(jb);
    return cudaGetLastError();
}

}  // namespace tsb
