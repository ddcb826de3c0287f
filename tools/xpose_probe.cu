// tools/xpose_probe.cu — measurement aid (not product code): what does the exchange between two radix-8 FFT passes cost
// when a warp owns a 256-point complex-double sub-transform (8 points per thread)?
//   smem : every thread stores its 8 points (16-byte, padded rows, conflict-free), __syncwarp, loads the 8 it needs
//   shfl : the same 8x8 transpose among the 8 lanes that share a column, as 3 butterfly stages of __shfl_xor_sync
//          (4 points x 4 words per stage and thread) + the selects that pick what to send and where to put what arrives
// Both variants do the same dummy arithmetic between exchanges (one DFMA per component) and run at the occupancy of
// the spectral-synthesis kernels (3 CTAs of 256 threads per SM).  Output: ns per exchange and warp, exchanges/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/xpose_probe tools/xpose_probe.cu && /tmp/xpose_probe
#include <cuda_runtime.h>

#include <cstdio>

struct __align__(16) cd {
    double x, y;
};

template <bool SHFL>
__global__ void __launch_bounds__(256, 3) xpose_kernel(double *sink, int iters, double a, double b) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cd *tile = reinterpret_cast<cd *>(raw) + warp * (8 * 36);
    const int a1 = lane >> 2, a0 = lane & 3;
    cd v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = cd{(double)(threadIdx.x * 8 + j), (double)(blockIdx.x + j)};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            v[j].x = fma(v[j].x, a, b);
            v[j].y = fma(v[j].y, a, b);
        }
        if (SHFL) {
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const int m = 1 << s;
                const bool up = (a1 >> s) & 1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (j & m) continue;
                    cd send = up ? v[j] : v[j | m];
                    cd recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 4 << s);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 4 << s);
                    if (up) v[j] = recv;
                    else v[j | m] = recv;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) tile[j * 36 + lane] = v[j];   // slot j of lane (a1, a0)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = tile[a1 * 36 + j * 4 + a0];  // slot a1 of lane (j, a0)
            __syncwarp();
        }
    }
    double acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc += v[j].x + v[j].y;
    if (acc == 12345.678) sink[0] = acc;
}

template <bool SHFL>
static double run(int iters) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *sink;
    cudaMalloc(&sink, 8);
    const size_t smem = 8 * 8 * 36 * sizeof(cd);
    cudaFuncSetAttribute(xpose_kernel<SHFL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    xpose_kernel<SHFL><<<sms * 3, 256, smem>>>(sink, 16, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    xpose_kernel<SHFL><<<sms * 3, 256, smem>>>(sink, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(sink);
    const double warps = (double)sms * 3 * 8;
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"iters\": %d, \"ns_per_exchange_per_sm\": %.3f, \"exchanges_per_s\": %.4g, "
           "\"points_per_s\": %.4g}\n",
           SHFL ? "shfl" : "smem", ms, iters, ms * 1e6 / iters / 24.0, warps * iters / (ms * 1e-3), warps * iters * 256.0 / (ms * 1e-3));
    return ms;
}

int main() {
    if (cudaSetDevice(0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int iters = 20000;
    double a = run<false>(iters), b = run<true>(iters);
    printf("{\"shfl_over_smem_time\": %.3f}\n", b / a);
    return cudaGetLastError() != cudaSuccess;
}
