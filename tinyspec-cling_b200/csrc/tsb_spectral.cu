// tsb_spectral.cu — fused spectral synthesis: device-side Hermitian spectrum -> inverse real FFT of twice the
// frame length -> triangular window -> overlap-add (BASELINE cfg5; the pattern of the reference's
// hacks/spectral_bytebeat/*.cpp, e.g. voidbubbles.cpp:4-19):
//
//     buf.resize(nch, 2n);  buf[c][i] = <formula>, buf[c][2n-i] = conj(buf[c][i])  for 0 < i < n     (the hack's loop)
//     frft(buf, buf, -1)                      rtlib/fft.hpp:3-42: y[m] = sum_k X[k] e^{-2 pi i k m / 2n} / sqrt(2n)
//     out[c][i] = buf[c][i].real() * (1.0 - abs(1.0 - i/(n/2.0)))  for i < n                          (the hack's loop)
//     overlap-add at hop                      tinyspec.cpp:170-200
//
// One quirk of the pattern is part of the result: `buf` is static and frft runs in place, so bins 0 and n — which the
// hack never writes — hold the previous frame's OUTPUT samples y[0] and y[n] when the next frame is transformed.
// Only their real parts reach the real output; they are carried from frame to frame as two doubles per channel.
//
// Real samples from a Hermitian spectrum of 2n points need one complex FFT of n points: with W = e^{-i pi / n},
//     G[k] = (X[k] + conj(X[n-k])) + i W^k (X[k] - conj(X[n-k])),   v = FFT_n(G),   y[2j] = Re v[j], y[2j+1] = Im v[j]
// and the n-point FFT is a four-step one (n = n1*n2, k = n2*k1 + k2, j = j1 + n1*j2):
//     kernel A  generator + G + n2 column FFTs of n1 points + twiddle e^{-2 pi i k2 j1 / n}  -> scratch[j1][k2]
//     kernel B  n1 row FFTs of n2 points; only j < n/2 is needed (the first n samples): unpack, 1/sqrt(2n), the
//               carried bins 0 / n added analytically, window, overlap-add into the output rows
// The scratch of one frame (channels * n * 16 bytes; 64 MiB at cfg5) is written and read back to back, i.e. from L2.
#include <cuda_runtime.h>

#include <cmath>
#include <map>
#include <vector>

#include "tsb_host.hpp"
#include "tsb_math.cuh"
#include "tsb_spectral.hpp"

namespace tsb {

using cd = cpx<double>;

// ---- generators: the spectrum value of bin i (0 < i < n) of channel c ----
// The seeded integer formula of oracle/ref/taps/bytebeat_tap.cpp / examples/bytebeat.cpp.
struct GenHash {
    uint32_t frame;
    __device__ __forceinline__ cd operator()(int c, int i) const {
        uint32_t h = frame * 2654435761u ^ ((uint32_t)c * 40503u + (uint32_t)i * 2246822519u);
        h ^= h >> 15;
        h *= 2246822519u;
        h ^= h >> 13;
        // x / 32768.0 / (double)(1 + (i >> 4)): the first division is exact, the second correctly rounded
        const double d = (double)(1 + (i >> 4));
        return mk<double>(__ddiv_rn((double)((int)(h & 0xffff) - 32768) * (1.0 / 32768.0), d),
                          __ddiv_rn((double)((int)(h >> 16) - 32768) * (1.0 / 32768.0), d));
    }
};
// Spectrum values computed by the caller: device complex128 rows [channels][n] of this frame, bins 1..n-1 are read.
struct GenBuffer {
    const cd *spec;
    long long n;
    __device__ __forceinline__ cd operator()(int c, int i) const { return spec[(long long)c * n + i]; }
};

// One CTA runs C sub-transforms of M = 2^LM points side by side, T threads each (one radix-8 butterfly per thread and
// pass): C * T = 256 threads and 64 KB of tiles, fewer when the other dimension (2^LO sub-transforms in all) is smaller.
template <int LM, int LO>
struct SubFft {
    static constexpr int M = 1 << LM;
    static constexpr int T = M / 8 < 1 ? 1 : M / 8;
    static constexpr int CMAX = (2048 / M) < 1 ? 1 : (2048 / M);
    static constexpr int C = CMAX < (1 << LO) ? CMAX : (1 << LO);
    static constexpr int NT = C * T;
};
template <int LOG2M, int T, int P, int END>
struct SpecPasses {
    static __device__ __forceinline__ cd *run(int lt, cd *src, cd *dst, const cd *tw) {
        stockham_pass_io<double, LOG2M, T, P>(
            lt, tw, [&](int a, int c) { return src[Swz<double>::f(a) ^ Swz<double>::f(c)]; },
            [&](int a, int c, cd v) { dst[Swz<double>::f(a) ^ Swz<double>::f(c)] = v; });
        __syncthreads();
        return SpecPasses<LOG2M, T, P + 1, END>::run(lt, dst, src, tw);
    }
};
template <int LOG2M, int T, int END>
struct SpecPasses<LOG2M, T, END, END> {
    static __device__ __forceinline__ cd *run(int, cd *src, cd *, const cd *) { return src; }
};

// e^{-i pi q / n} for 0 <= q < 2n, argument exact
__device__ __forceinline__ cd unit_w(long long q, double inv_n) {
    double s, c;
    sincospi(-(double)q * inv_n, &s, &c);
    return mk<double>(c, s);
}

// ---- kernel A: one CTA = C consecutive columns k2 of one channel ----
template <int L1, int L2, class Gen>
__global__ void __launch_bounds__(SubFft<L1, L2>::NT) spectral_cols_kernel(Gen gen, cd *scratch, const cd *tw1, int channels) {
    using F = SubFft<L1, L2>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2, C = F::C, NT = F::NT;
    constexpr long long n = (long long)n1 * n2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *bufA = reinterpret_cast<cd *>(smem_raw), *bufB = bufA + C * n1;
    const double inv_n = 1.0 / (double)n;
    const int groups = n2 / C;
    for (long long b = blockIdx.x; b < (long long)channels * groups; b += gridDim.x) {
        const int ch = (int)(b / groups), k2_0 = (int)(b % groups) * C;
        // G[k], k = n2*k1 + k2, for the C columns: consecutive threads take consecutive k2 (the caller's spectrum rows,
        // when the generator reads memory, are then read in contiguous pieces)
        for (int i = threadIdx.x; i < C * n1; i += NT) {
            const int c = i % C, k1 = i / C;
            const long long k = (long long)n2 * k1 + k2_0 + c;
            cd g = mk<double>(0.0, 0.0);
            if (k > 0) {  // (bins 0 and n are the carried ones: kernel B adds their contribution)
                const cd a = gen(ch, (int)k), bb = conj(gen(ch, (int)(n - k)));
                const cd s = a + bb, d = a - bb;
                const cd w = unit_w(k, inv_n);           // W^k
                const cd wd = cmul(w, d);
                g = mk<double>(s.x - wd.y, s.y + wd.x);  // s + i*W^k*d
            }
            bufA[c * n1 + Swz<double>::f(k1)] = g;
        }
        __syncthreads();
        const int gi = threadIdx.x / F::T, lt = threadIdx.x % F::T;
        cd *res = SpecPasses<L1, F::T, 0, RadixPlan<L1>::passes>::run(lt, bufA + gi * n1, bufB + gi * n1, tw1) - gi * n1;
        // A[k2][j1] * e^{-2 pi i k2 j1 / n} -> scratch[ch][j1][k2]
        cd *dst = scratch + (long long)ch * n;
        for (int i = threadIdx.x; i < C * n1; i += NT) {
            const int c = i % C, j1 = i / C;
            const long long q = (long long)(k2_0 + c) * j1;          // < n
            const cd w = unit_w(2 * q, inv_n);
            dst[(long long)j1 * n2 + k2_0 + c] = cmul(res[c * n1 + Swz<double>::f(j1)], w);
        }
        __syncthreads();
    }
}

// ---- kernel B: one CTA = C consecutive rows j1 of one channel ----
template <int L1, int L2>
__global__ void __launch_bounds__(SubFft<L2, L1>::NT)
spectral_rows_kernel(const cd *scratch, const cd *tw2, int channels, double *out, long long out_stride, long long out_len,
                     long long pos0, const double *edge_in, double *edge_out, int pairs_aligned) {
    using F = SubFft<L2, L1>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2, C = F::C, NT = F::NT;
    constexpr long long n = (long long)n1 * n2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *bufA = reinterpret_cast<cd *>(smem_raw), *bufB = bufA + C * n2;
    const double inv_sqrt = 1.0 / sqrt((double)(2 * n));
    const double half_n = (double)n / 2.0;
    const int groups = n1 / C;
    for (long long b = blockIdx.x; b < (long long)channels * groups; b += gridDim.x) {
        const int ch = (int)(b / groups), j1_0 = (int)(b % groups) * C;
        const cd *src = scratch + (long long)ch * n + (long long)j1_0 * n2;  // C rows, contiguous
        for (int i = threadIdx.x; i < C * n2; i += NT) bufA[(i / n2) * n2 + Swz<double>::f(i % n2)] = src[i];
        __syncthreads();
        const int gi = threadIdx.x / F::T, lt = threadIdx.x % F::T;
        cd *res = SpecPasses<L2, F::T, 0, RadixPlan<L2>::passes>::run(lt, bufA + gi * n2, bufB + gi * n2, tw2) - gi * n2;
        // the carried bins 0 and n of the spectrum: X0 + Xn*(-1)^m, i.e. (d + q) on even and (d - q) on odd samples
        const double d = edge_in ? edge_in[2 * ch] : 0.0, q = edge_in ? edge_in[2 * ch + 1] : 0.0;
        const double e_even = (d + q) * inv_sqrt, e_odd = (d - q) * inv_sqrt;
        double *row = out + (long long)ch * out_stride + pos0;
        // v[j], j = j1 + n1*j2, j < n/2  ->  samples m = 2j, 2j+1: consecutive threads take consecutive j1
        for (int i = threadIdx.x; i < C * (n2 / 2); i += NT) {
            const int c = i % C, j2 = i / C;
            const cd v = res[c * n2 + Swz<double>::f(j2)];
            const long long m = 2 * ((long long)(j1_0 + c) + (long long)n1 * j2);
            const double y0 = fma(v.x, inv_sqrt, e_even), y1 = fma(v.y, inv_sqrt, e_odd);
            // the hack's window, `1.0 - abs(1.0 - i/(n/2.0))`, same expression
            const double w0 = 1.0 - fabs(1.0 - (double)m / half_n), w1 = 1.0 - fabs(1.0 - (double)(m + 1) / half_n);
            if (pairs_aligned && pos0 + m + 1 < out_len) {
                double2 *p = reinterpret_cast<double2 *>(row + m);  // (pos0 and out_stride are even: 16-byte aligned)
                double2 t = *p;
                t.x += y0 * w0;
                t.y += y1 * w1;
                *p = t;
            } else {
                if (pos0 + m < out_len) row[m] += y0 * w0;
                if (pos0 + m + 1 < out_len) row[m + 1] += y1 * w1;
            }
        }
        if (edge_out && j1_0 == 0 && threadIdx.x == 0) {
            // what the in-place frft leaves in bins 0 and n for the next frame: y[0] and y[n] (j = 0 and j = n/2)
            edge_out[2 * ch] = fma(res[Swz<double>::f(0)].x, inv_sqrt, e_even);
            edge_out[2 * ch + 1] = fma(res[Swz<double>::f(n2 / 2)].x, inv_sqrt, e_even);
        }
        __syncthreads();
    }
}

// ---- host side ----
struct SpecCache {
    std::map<int, cd *> pass_tw;
    cd *scratch = nullptr;
    size_t scratch_elems = 0;
    double *edge[2] = {nullptr, nullptr};
    size_t edge_ch = 0;
};
static std::map<int, SpecCache> g_spec;

static cudaError_t spec_twiddles(SpecCache &pc, int log2m, const cd **out) {
    auto it = pc.pass_tw.find(log2m);
    if (it == pc.pass_tw.end()) {
        std::vector<double> t = pass_twiddle_table(1 << log2m);
        cd *dptr = nullptr;
        cudaError_t e = cudaMalloc(&dptr, t.size() * sizeof(double) + 16);
        if (e != cudaSuccess) return e;
        if (!t.empty()) e = cudaMemcpy(dptr, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        it = pc.pass_tw.emplace(log2m, dptr).first;
    }
    *out = it->second;
    return cudaSuccess;
}
cudaError_t op_spectral_release(int device) {
    auto it = g_spec.find(device);
    if (it == g_spec.end()) return cudaSuccess;
    SpecCache &pc = it->second;
    if (pc.scratch) cudaFree(pc.scratch);
    pc.scratch = nullptr;
    pc.scratch_elems = 0;
    for (double *&p : pc.edge) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    pc.edge_ch = 0;
    return cudaSuccess;
}

template <int L1, int L2>
static cudaError_t run_frames(SpecCache &pc, const SpectralJob &jb, cudaStream_t st, long long *launches) {
    constexpr long long n = 1LL << (L1 + L2);
    const cd *tw1, *tw2;
    cudaError_t e;
    if ((e = spec_twiddles(pc, L1, &tw1)) != cudaSuccess) return e;
    if ((e = spec_twiddles(pc, L2, &tw2)) != cudaSuccess) return e;
    const size_t need = (size_t)jb.channels * n;
    if (pc.scratch_elems < need) {
        if (pc.scratch) cudaFree(pc.scratch);
        pc.scratch = nullptr;
        pc.scratch_elems = 0;
        if ((e = cudaMalloc(&pc.scratch, need * sizeof(cd))) != cudaSuccess) return e;
        pc.scratch_elems = need;
    }
    if (pc.edge_ch < (size_t)jb.channels) {
        for (double *&p : pc.edge) {
            if (p) cudaFree(p);
            p = nullptr;
            if ((e = cudaMalloc(&p, (size_t)jb.channels * 16)) != cudaSuccess) return e;
        }
        pc.edge_ch = (size_t)jb.channels;
    }
    // the carried bins: from the caller's state, or zeros (a freshly resized buffer, rtinc/buffer.hpp:15-23)
    if (jb.state_edge && jb.load_state) e = cudaMemcpyAsync(pc.edge[0], jb.state_edge, (size_t)jb.channels * 16, cudaMemcpyDeviceToDevice, st);
    else e = cudaMemsetAsync(pc.edge[0], 0, (size_t)jb.channels * 16, st);
    if (e != cudaSuccess) return e;
    using FA = SubFft<L1, L2>;
    using FB = SubFft<L2, L1>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2;
    constexpr int CA = FA::C, CB = FB::C;
    const size_t smA = (size_t)2 * FA::C * n1 * sizeof(cd), smB = (size_t)2 * FB::C * n2 * sizeof(cd);
    auto kA_hash = spectral_cols_kernel<L1, L2, GenHash>;
    auto kA_buf = spectral_cols_kernel<L1, L2, GenBuffer>;
    auto kB = spectral_rows_kernel<L1, L2>;
    if ((e = cudaFuncSetAttribute(kA_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kA_buf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB)) != cudaSuccess) return e;
    const unsigned gridA = (unsigned)((long long)jb.channels * (n2 / CA)), gridB = (unsigned)((long long)jb.channels * (n1 / CB));
    for (long long f = 0; f < jb.n_frames; f++) {
        if (jb.generator == SPEC_GEN_HASH) {
            GenHash g{(uint32_t)(jb.first_frame + f)};
            kA_hash<<<gridA, FA::NT, smA, st>>>(g, pc.scratch, tw1, jb.channels);
        } else {
            GenBuffer g{reinterpret_cast<const cd *>(jb.spectrum) + (size_t)f * jb.channels * n, n};
            kA_buf<<<gridA, FA::NT, smA, st>>>(g, pc.scratch, tw1, jb.channels);
        }
        const long long pos0 = jb.out_off0 + f * jb.hop;
        const int aligned = ((pos0 | jb.out_stride) & 1) == 0 && (reinterpret_cast<uintptr_t>(jb.out) & 15) == 0;
        kB<<<gridB, FB::NT, smB, st>>>(pc.scratch, tw2, jb.channels, jb.out, jb.out_stride, jb.out_len, pos0, pc.edge[f & 1],
                                        pc.edge[(f + 1) & 1], aligned);
        *launches += 2;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (jb.state_edge)
        e = cudaMemcpyAsync(jb.state_edge, pc.edge[jb.n_frames & 1], (size_t)jb.channels * 16, cudaMemcpyDeviceToDevice, st);
    return e;
}

cudaError_t op_spectral_render(int device, const SpectralJob &jb, cudaStream_t st, long long *launches) {
    SpecCache &pc = g_spec[device];
    switch (jb.log2n) {
#define TSB_SPEC_CASE(L, A, B) \
    case L:                    \
        return run_frames<A, B>(pc, jb, st, launches);
        TSB_SPEC_CASE(8, 4, 4)
        TSB_SPEC_CASE(9, 5, 4)
        TSB_SPEC_CASE(10, 5, 5)
        TSB_SPEC_CASE(11, 6, 5)
        TSB_SPEC_CASE(12, 6, 6)
        TSB_SPEC_CASE(13, 7, 6)
        TSB_SPEC_CASE(14, 7, 7)
        TSB_SPEC_CASE(15, 8, 7)
        TSB_SPEC_CASE(16, 8, 8)
        TSB_SPEC_CASE(17, 9, 8)
        TSB_SPEC_CASE(18, 9, 9)
#undef TSB_SPEC_CASE
    }
    return cudaErrorInvalidValue;
}

}  // namespace tsb
