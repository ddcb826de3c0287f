// tsb_ops.cu — the reference's helper functions one by one, on the device, in double precision:
//   window_hann / window_sqrt_hann   rtlib/windowfn.hpp:1-14
//   frft(in, out, exponent)          rtlib/fft.hpp:3-42      (any size, any exponent)
//   PhaseVocoder::analyze / shift / synthesize on FFTBuf / PVBuf data   rtlib/pv.hpp:17-79 (all N bins)
// These are the "immediate mode" operators behind the source-compatible C++ facade (rtinc/): a hack that
// mixes helper calls with its own per-sample loops runs them one call at a time.  Hacks made only of helper
// calls are fused into the chain kernel instead (tsb_chain.cuh).
//
// FFT: batched complex-to-complex, forward kernel exp(-2*pi*i*j*k/n) unnormalised — the contract FFTW's
// fftw_plan_many_dft(FFTW_FORWARD) has in the reference.  Power-of-two n <= 4096: one CTA per row, Stockham
// radix-8 passes in shared memory (tsb_math.cuh).  Larger powers of two: four-step (column FFTs + twiddle,
// row FFTs) through a scratch buffer.  Any other n: Bluestein's chirp-z over the power-of-two kernels.
#include <cuda_runtime.h>

#include <cmath>
#include <complex>
#include <map>
#include <vector>

#include "tsb_host.hpp"
#include "tsb_math.cuh"
#include "tsb_ops.hpp"

namespace tsb {

using cd = cpx<double>;

// ------------------------------------------------------------------ batched pow-2 FFT, n <= 4096
struct RowAddr {            // element e of row r lives at base[(r / inner) * outer + (r % inner) * inner_stride + e * es]
    long long inner, outer, inner_stride, es;
};
TSB_HD long long row_base(const RowAddr &a, long long r) { return (r / a.inner) * a.outer + (r % a.inner) * a.inner_stride; }

template <int LOG2M, int P, int END>
struct SmemPasses {
    static constexpr int M = 1 << LOG2M;
    static constexpr int T = (M / 8) < 32 ? 32 : (M / 8);
    static __device__ __forceinline__ cd *run(cd *src, cd *dst, const cd *tw) {
        stockham_pass_io<double, LOG2M, T, P>(
            (int)threadIdx.x, tw, [&](int a, int c) { return src[Swz<double>::f(a) ^ Swz<double>::f(c)]; },
            [&](int a, int c, cd v) { dst[Swz<double>::f(a) ^ Swz<double>::f(c)] = v; });
        __syncthreads();
        return SmemPasses<LOG2M, P + 1, END>::run(dst, src, tw);
    }
};
template <int LOG2M, int END>
struct SmemPasses<LOG2M, END, END> {
    static __device__ __forceinline__ cd *run(cd *src, cd *, const cd *) { return src; }
};

// flags: bit0 = conjugate input and output (inverse transform through the forward kernel),
//        bit1 = multiply output element e of row r by exp(-2*pi*i*(r % inner_in)*e / ntw) (four-step twiddle)
template <int LOG2M>
__global__ void __launch_bounds__(((1 << LOG2M) / 8) < 32 ? 32 : ((1 << LOG2M) / 8))
fft_rows_kernel(const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw, int flags, double ntw,
                double scale) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = (M / 8) < 32 ? 32 : (M / 8);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *bufA = reinterpret_cast<cd *>(smem_raw), *bufB = bufA + M;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const cd *src = in + row_base(ia, r);
        for (int e = threadIdx.x; e < M; e += T) {
            cd v = src[e * ia.es];
            if (flags & 1) v.y = -v.y;
            bufA[Swz<double>::f(e)] = v;
        }
        __syncthreads();
        cd *res = SmemPasses<LOG2M, 0, RadixPlan<LOG2M>::passes>::run(bufA, bufB, tw);
        cd *dst = out + row_base(oa, r);
        const double rr = (double)(r % ia.inner);
        for (int e = threadIdx.x; e < M; e += T) {
            cd v = res[Swz<double>::f(e)];
            if (flags & 2) {
                double s, c;
                sincospi(-2.0 * (rr * (double)e) / ntw, &s, &c);
                v = cmul(v, mk<double>(c, s));
            }
            if (flags & 1) v.y = -v.y;
            v.x *= scale;
            v.y *= scale;
            dst[e * oa.es] = v;
        }
        __syncthreads();
    }
}

// Tiled variant for the two steps of the four-step transform, where either the elements of a row or the rows
// themselves are contiguous in memory but not both: one CTA transforms C rows that are ADJACENT in memory
// (consecutive r with inner_stride 1, or rows that are each contiguous) and moves them with the contiguous
// dimension fastest, so that every global access is a full 128-byte line instead of one 16-byte element per
// sector.  The C transforms run side by side (T threads each) with CTA-wide barriers between passes.
template <int LOG2M, int T, int P, int END>
struct TilePasses {
    static __device__ __forceinline__ cd *run(int lt, cd *src, cd *dst, const cd *tw) {
        stockham_pass_io<double, LOG2M, T, P>(
            lt, tw, [&](int a, int c) { return src[Swz<double>::f(a) ^ Swz<double>::f(c)]; },
            [&](int a, int c, cd v) { dst[Swz<double>::f(a) ^ Swz<double>::f(c)] = v; });
        __syncthreads();
        return TilePasses<LOG2M, T, P + 1, END>::run(lt, dst, src, tw);
    }
};
template <int LOG2M, int T, int END>
struct TilePasses<LOG2M, T, END, END> {
    static __device__ __forceinline__ cd *run(int, cd *src, cd *, const cd *) { return src; }
};
template <int LOG2M>
struct TileCfg {
    static constexpr int M = 1 << LOG2M;
    static constexpr int T = (M / 8) < 32 ? 32 : (M / 8);
    // C*M <= 2048 keeps the tile at 64 KB so that three CTAs share an SM and overlap each other's loads and stores
    static constexpr int C = M <= 256 ? 8 : (M <= 512 ? 4 : (M <= 1024 ? 2 : 1));
};
template <int LOG2M>
__global__ void __launch_bounds__(TileCfg<LOG2M>::C *TileCfg<LOG2M>::T)
fft_tile_kernel(const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw, int flags, double ntw,
                double scale) {
    constexpr int M = TileCfg<LOG2M>::M, T = TileCfg<LOG2M>::T, C = TileCfg<LOG2M>::C, NT = C * T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *bufA = reinterpret_cast<cd *>(smem_raw), *bufB = bufA + C * M;
    const int g = threadIdx.x / T, lt = threadIdx.x % T;
    for (long long r0 = (long long)blockIdx.x * C; r0 < rows; r0 += (long long)gridDim.x * C) {
        const bool in_rows_contig = ia.es == 1;
        for (int i = threadIdx.x; i < C * M; i += NT) {
            int c = in_rows_contig ? i / M : i % C, e = in_rows_contig ? i % M : i / C;
            cd v = in[row_base(ia, r0 + c) + e * ia.es];
            if (flags & 1) v.y = -v.y;
            bufA[c * M + Swz<double>::f(e)] = v;
        }
        __syncthreads();
        cd *res = TilePasses<LOG2M, T, 0, RadixPlan<LOG2M>::passes>::run(lt, bufA + g * M, bufB + g * M, tw);
        res -= g * M;  // base of the buffer the results are in
        const bool out_rows_contig = oa.es == 1;
        if ((flags & 2) && !out_rows_contig) {
            // four-step twiddle W_n^(j2*k1): this thread keeps one row (j2) and walks k1 in steps of NT/C, so the
            // factor advances by a constant rotation — two sincospi per thread instead of one per element
            const int c = threadIdx.x % C, e0 = threadIdx.x / C;
            const double j2 = (double)((r0 + c) % ia.inner);
            double sn, cs;
            sincospi(-2.0 * (j2 * (double)e0) / ntw, &sn, &cs);
            cd w = mk<double>(cs, sn);
            sincospi(-2.0 * (j2 * (double)(NT / C)) / ntw, &sn, &cs);
            const cd step = mk<double>(cs, sn);
            cd *dst = out + row_base(oa, r0 + c);
            for (int e = e0; e < M; e += NT / C) {
                cd v = cmul(res[c * M + Swz<double>::f(e)], w);
                if (flags & 1) v.y = -v.y;
                v.x *= scale;
                v.y *= scale;
                dst[e * oa.es] = v;
                w = cmul(w, step);
            }
        } else
        for (int i = threadIdx.x; i < C * M; i += NT) {
            int c = out_rows_contig ? i / M : i % C, e = out_rows_contig ? i % M : i / C;
            cd v = res[c * M + Swz<double>::f(e)];
            if (flags & 2) {
                double sn, cs;
                sincospi(-2.0 * ((double)((r0 + c) % ia.inner) * (double)e) / ntw, &sn, &cs);
                v = cmul(v, mk<double>(cs, sn));
            }
            if (flags & 1) v.y = -v.y;
            v.x *= scale;
            v.y *= scale;
            out[row_base(oa, r0 + c) + e * oa.es] = v;
        }
        __syncthreads();
    }
}

struct FftPlanCache {
    std::map<int, cd *> pass_tw;                                 // log2 m -> device per-pass twiddles
    std::map<long long, std::pair<cd *, cd *>> bluestein;        // n -> (chirp c[n], FFT_L of the wrapped conj chirp)
    cd *scratch = nullptr;                                       // four-step intermediate
    size_t scratch_elems = 0;
    cd *work = nullptr;                                          // Bluestein's padded rows
    size_t work_elems = 0;
};
static std::map<int, FftPlanCache> g_cache;  // per device

static cudaError_t pass_twiddles(FftPlanCache &pc, int log2m, const cd **out) {
    auto it = pc.pass_tw.find(log2m);
    if (it == pc.pass_tw.end()) {
        std::vector<double> t = pass_twiddle_table(1 << log2m);
        cd *d = nullptr;
        cudaError_t e = cudaMalloc(&d, t.size() * sizeof(double));
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        it = pc.pass_tw.emplace(log2m, d).first;
    }
    *out = it->second;
    return cudaSuccess;
}
// Both buffers grow on demand and stay (allocating gigabytes per call costs more than the transform); cudaFree
// waits for the work that still uses the old block.  op_release_scratch gives them back.
static cudaError_t ensure_work(FftPlanCache &pc, size_t elems, cd **out) {
    if (pc.work_elems < elems) {
        if (pc.work) cudaFree(pc.work);
        pc.work = nullptr;
        pc.work_elems = 0;
        cudaError_t e = cudaMalloc(&pc.work, elems * sizeof(cd));
        if (e != cudaSuccess) return e;
        pc.work_elems = elems;
    }
    *out = pc.work;
    return cudaSuccess;
}
cudaError_t op_release_scratch(int device) {
    auto it = g_cache.find(device);
    if (it == g_cache.end()) return cudaSuccess;
    FftPlanCache &pc = it->second;
    if (pc.scratch) cudaFree(pc.scratch);
    if (pc.work) cudaFree(pc.work);
    pc.scratch = pc.work = nullptr;
    pc.scratch_elems = pc.work_elems = 0;
    return cudaSuccess;
}
static cudaError_t ensure_scratch(FftPlanCache &pc, size_t elems, cd **out) {
    if (pc.scratch_elems < elems) {
        if (pc.scratch) cudaFree(pc.scratch);
        pc.scratch = nullptr;
        pc.scratch_elems = 0;
        cudaError_t e = cudaMalloc(&pc.scratch, elems * sizeof(cd));
        if (e != cudaSuccess) return e;
        pc.scratch_elems = elems;
    }
    *out = pc.scratch;
    return cudaSuccess;
}

template <int LOG2M>
static cudaError_t launch_rows(const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw, int flags,
                               double ntw, double scale, cudaStream_t st, long long *launches) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = (M / 8) < 32 ? 32 : (M / 8);
    size_t smem = 2 * M * sizeof(cd);
    auto kern = fft_rows_kernel<LOG2M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long grid = rows < 148LL * 16 ? rows : 148LL * 16;
    kern<<<(unsigned)grid, T, smem, st>>>(in, out, ia, oa, rows, tw, flags, ntw, scale);
    if (launches) ++*launches;
    return cudaGetLastError();
}
template <int LOG2M>
static cudaError_t launch_tile(const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw, int flags,
                               double ntw, double scale, cudaStream_t st, long long *launches) {
    using TC = TileCfg<LOG2M>;
    if (TC::C == 1 || rows % TC::C != 0) return launch_rows<LOG2M>(in, out, ia, oa, rows, tw, flags, ntw, scale, st, launches);
    size_t smem = 2 * (size_t)TC::C * TC::M * sizeof(cd);
    auto kern = fft_tile_kernel<LOG2M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long tiles = rows / TC::C;
    long long grid = tiles < 148LL * 12 ? tiles : 148LL * 12;
    kern<<<(unsigned)grid, TC::C * TC::T, smem, st>>>(in, out, ia, oa, rows, tw, flags, ntw, scale);
    if (launches) ++*launches;
    return cudaGetLastError();
}
static cudaError_t tile_dispatch(int log2m, const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw,
                                 int flags, double ntw, double scale, cudaStream_t st, long long *launches) {
    switch (log2m) {
#define C(L) case L: return launch_tile<L>(in, out, ia, oa, rows, tw, flags, ntw, scale, st, launches);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
    }
    return cudaErrorInvalidValue;
}
static cudaError_t rows_dispatch(int log2m, const cd *in, cd *out, RowAddr ia, RowAddr oa, long long rows, const cd *tw,
                                 int flags, double ntw, double scale, cudaStream_t st, long long *launches) {
    switch (log2m) {
#define C(L) case L: return launch_rows<L>(in, out, ia, oa, rows, tw, flags, ntw, scale, st, launches);
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
    }
    return cudaErrorInvalidValue;
}

// rows x n forward (or inverse-through-conjugation) FFT of contiguous rows, n a power of two
static cudaError_t fft_pow2(int device, const cd *in, cd *out, long long rows, long long n, bool inverse, double scale,
                            cudaStream_t st, long long *launches) {
    FftPlanCache &pc = g_cache[device];
    int lg = 0;
    while ((1LL << lg) < n) lg++;
    const int cj = inverse ? 1 : 0;
    if (lg == 0) {
        if (in != out) return cudaMemcpyAsync(out, in, rows * sizeof(cd), cudaMemcpyDeviceToDevice, st);
        return cudaSuccess;
    }
    const cd *tw;
    cudaError_t e;
    if (lg <= 12) {
        if ((e = pass_twiddles(pc, lg, &tw)) != cudaSuccess) return e;
        RowAddr a{1, n, 0, 1};
        return rows_dispatch(lg, in, out, a, a, rows, tw, cj, 1.0, scale, st, launches);
    }
    // four-step: n = n1*n2, index j = j1*n2 + j2, k = k1 + n1*k2
    int l1 = lg / 2, l2 = lg - l1;
    if (l2 > 12) return cudaErrorInvalidValue;  // n > 2^24
    long long n1 = 1LL << l1, n2 = 1LL << l2;
    cd *tmp;
    if ((e = ensure_scratch(pc, (size_t)(rows * n), &tmp)) != cudaSuccess) return e;
    // step 1: for every (row, j2): length-n1 FFT over j1 (stride n2), times W_n^(j2*k1), stored as tmp[row][k1][j2]
    if ((e = pass_twiddles(pc, l1, &tw)) != cudaSuccess) return e;
    RowAddr ia{n2, n, 1, n2}, oa{n2, n, 1, n2};
    if ((e = tile_dispatch(l1, in, tmp, ia, oa, rows * n2, tw, cj | 2, (double)n, 1.0, st, launches)) != cudaSuccess) return e;
    // step 2: for every (row, k1): length-n2 FFT over j2 (contiguous), stored at out[row][k1 + n1*k2]
    if ((e = pass_twiddles(pc, l2, &tw)) != cudaSuccess) return e;
    RowAddr ib{n1, n, n2, 1}, ob{n1, n, 1, n1};
    return tile_dispatch(l2, tmp, out, ib, ob, rows * n1, tw, cj, 1.0, scale, st, launches);
}

// ------------------------------------------------------------------ Bluestein (any n)
__global__ void bluestein_pre(const cd *in, cd *a, const cd *chirp, long long rows, long long n, long long L) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= rows * L) return;
    long long r = idx / L, m = idx % L;
    cd v = mk<double>(0, 0);
    if (m < n) v = cmul(in[r * n + m], chirp[m]);
    a[idx] = v;
}
__global__ void bluestein_mul(cd *a, const cd *bf, long long rows, long long L) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= rows * L) return;
    a[idx] = cmul(a[idx], bf[idx % L]);
}
__global__ void bluestein_post(const cd *a, cd *out, const cd *chirp, long long rows, long long n, long long L) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= rows * n) return;
    long long r = idx / n, k = idx % n;
    out[idx] = cmul(a[r * L + k], chirp[k]);
}

static cudaError_t fft_any(int device, const cd *in, cd *out, long long rows, long long n, cudaStream_t st,
                           long long *launches) {
    if (n > 0 && (n & (n - 1)) == 0) return fft_pow2(device, in, out, rows, n, false, 1.0, st, launches);
    FftPlanCache &pc = g_cache[device];
    long long L = 1;
    while (L < 2 * n - 1) L <<= 1;
    cudaError_t e;
    auto it = pc.bluestein.find(n);
    if (it == pc.bluestein.end()) {
        // chirp c[m] = exp(-i*pi*m^2/n), with m^2 reduced mod 2n in integers so the angle stays exact
        std::vector<double> c(2 * n), b(2 * L, 0.0);
        const long double pi = 3.141592653589793238462643383279502884L;
        for (long long m = 0; m < n; m++) {
            long long q = (long long)(((unsigned long long)m * (unsigned long long)m) % (unsigned long long)(2 * n));
            long double ang = -pi * (long double)q / (long double)n;
            c[2 * m] = (double)cosl(ang);
            c[2 * m + 1] = (double)sinl(ang);
        }
        for (long long m = 0; m < n; m++) {  // b[m] = conj(c[|m|]) wrapped to length L
            b[2 * m] = c[2 * m];
            b[2 * m + 1] = -c[2 * m + 1];
            if (m) {
                b[2 * (L - m)] = c[2 * m];
                b[2 * (L - m) + 1] = -c[2 * m + 1];
            }
        }
        cd *dc = nullptr, *db = nullptr;
        if ((e = cudaMalloc(&dc, n * sizeof(cd))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&db, L * sizeof(cd))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(dc, c.data(), n * sizeof(cd), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        if ((e = cudaMemcpy(db, b.data(), L * sizeof(cd), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        if ((e = fft_pow2(device, db, db, 1, L, false, 1.0, 0, nullptr)) != cudaSuccess) return e;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
        it = pc.bluestein.emplace(n, std::make_pair(dc, db)).first;
    }
    const cd *chirp = it->second.first, *bf = it->second.second;
    cd *a = nullptr;  // own work buffer (fft_pow2 may use the shared scratch for its four-step)
    if ((e = ensure_work(pc, (size_t)(rows * L), &a)) != cudaSuccess) return e;
    const int TB = 256;
    bluestein_pre<<<(unsigned)((rows * L + TB - 1) / TB), TB, 0, st>>>(in, a, chirp, rows, n, L);
    if ((e = fft_pow2(device, a, a, rows, L, false, 1.0, st, launches)) != cudaSuccess) return e;
    bluestein_mul<<<(unsigned)((rows * L + TB - 1) / TB), TB, 0, st>>>(a, bf, rows, L);
    if ((e = fft_pow2(device, a, a, rows, L, true, 1.0 / (double)L, st, launches)) != cudaSuccess) return e;
    bluestein_post<<<(unsigned)((rows * n + TB - 1) / TB), TB, 0, st>>>(a, out, chirp, rows, n, L);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

cudaError_t op_fft_forward(int device, const double *in, double *out, long long rows, long long n, cudaStream_t st,
                           long long *launches) {
    if (rows <= 0 || n <= 0) return cudaSuccess;
    return fft_any(device, reinterpret_cast<const cd *>(in), reinterpret_cast<cd *>(out), rows, n, st, launches);
}

// ------------------------------------------------------------------ frft eigen-combination, rtlib/fft.hpp:20-41
// g = unnormalised forward FFT of `in` (what fftw_execute left in `out`, copied to fft_tmp at line 19).
// in_place: the reference's loop then reads in[c][i] / in[c][j] from the buffer the FFT has just overwritten,
// and for i > n/2 from elements the loop itself has already replaced (j = n-i < i).  One thread handles the
// pair (i, n-i), low index first, which reproduces exactly that order.
struct Cplx3 {
    cd im1, im2, im3;
};
__device__ __forceinline__ cd c_add(cd a, cd b) { return mk<double>(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ cd c_sub(cd a, cd b) { return mk<double>(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
__device__ __forceinline__ cd c_mul(cd a, cd b) {  // std::complex operator*: (ac-bd, ad+bc), each op rounded
    return mk<double>(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ cd frft_combine(cd f0, cd f1, cd f2, cd f3, const Cplx3 &k) {
    const cd im = mk<double>(0.0, 1.0);
    cd b0 = c_add(c_add(c_add(f0, f1), f2), f3);
    cd b1 = c_sub(c_sub(c_add(f0, c_mul(im, f1)), f2), c_mul(im, f3));
    cd b2 = c_sub(c_add(c_sub(f0, f1), f2), f3);
    cd b3 = c_add(c_sub(c_sub(f0, c_mul(im, f1)), f2), c_mul(im, f3));
    b1 = c_mul(b1, k.im1);
    b2 = c_mul(b2, k.im2);
    b3 = c_mul(b3, k.im3);
    cd s = c_add(c_add(c_add(b0, b1), b2), b3);
    return mk<double>(s.x / 4.0, s.y / 4.0);
}
__global__ void frft_combine_kernel(const cd *in, const cd *g, cd *out, long long nch, long long n, double sqrtn, Cplx3 k,
                                    int in_place) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long half = n / 2 + 1;
    if (idx >= nch * half) return;
    long long c = idx / half, i = idx % half;
    const cd *gi = g + c * n;
    const cd *xin = in_place ? gi : in + c * n;  // in place, `in` IS the FFT output by the time the loop runs
    cd *o = out + c * n;
    long long j = i ? n - i : 0;
    cd f0 = xin[i], f2 = xin[j];
    cd f1 = mk<double>(gi[i].x / sqrtn, gi[i].y / sqrtn), f3 = mk<double>(gi[j].x / sqrtn, gi[j].y / sqrtn);
    cd oi = frft_combine(f0, f1, f2, f3, k);
    if (j != i) {
        // element j = n-i > i: its "f2" is in[c][i] — already replaced by oi when the reference gets there (in place)
        cd g0 = xin[j], g2 = in_place ? oi : xin[i];
        cd oj = frft_combine(g0, f3, g2, f1, k);
        o[j] = oj;
    }
    o[i] = oi;
}

cudaError_t op_frft(int device, const double *in, double *out, double *scratch_g, long long nch, long long n,
                    double exponent, cudaStream_t st, long long *launches) {
    if (nch * n == 0) return cudaSuccess;  // rtlib/fft.hpp:8-9
    cd *g = reinterpret_cast<cd *>(scratch_g);
    cudaError_t e = op_fft_forward(device, in, scratch_g, nch, n, st, launches);
    if (e != cudaSuccess) return e;
    std::complex<double> im(0, 1);
    std::complex<double> im1 = std::pow(im, exponent), im2 = std::pow(im, exponent * 2), im3 = std::pow(im, exponent * 3);
    Cplx3 k{mk<double>(im1.real(), im1.imag()), mk<double>(im2.real(), im2.imag()), mk<double>(im3.real(), im3.imag())};
    long long work = nch * (n / 2 + 1);
    const int TB = 128;
    frft_combine_kernel<<<(unsigned)((work + TB - 1) / TB), TB, 0, st>>>(reinterpret_cast<const cd *>(in), g,
                                                                           reinterpret_cast<cd *>(out), nch, n,
                                                                           std::sqrt((double)(size_t)n), k, in == out);
    if (launches) ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------ windows, rtlib/windowfn.hpp:1-14
__global__ void window_kernel(double *data, const double *w, long long nch, long long size) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nch * size) return;
    data[idx] = __dmul_rn(data[idx], w[idx % size]);
}
cudaError_t op_window(double *data, const double *wtab, long long nch, long long size, cudaStream_t st) {
    if (nch * size == 0) return cudaSuccess;
    const int TB = 256;
    window_kernel<<<(unsigned)((nch * size + TB - 1) / TB), TB, 0, st>>>(data, wtab, nch, size);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ PhaseVocoder on all N bins, rtlib/pv.hpp:17-48
__global__ void pv_analyze_kernel(const cd *spec, double *last_phase, double *pvdata, long long nch, long long size,
                                  double hop, double rate) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nch * size) return;
    long long j = idx % size;
    cd v = spec[idx];
    double amp = hypot(v.x, v.y);
    double phase = atan2(v.y, v.x);
    double diff = __dsub_rn(phase, last_phase[idx]);
    last_phase[idx] = phase;
    // phase_to_frequency, pv.hpp:39-45
    double expect = __dmul_rn(__dmul_rn(__dmul_rn((double)j, 2.0), 3.14159265358979323846), hop) / (double)size;
    double delta = __dsub_rn(diff, expect);
    int qpd = __double2int_rz(delta / 3.14159265358979323846);
    if (qpd >= 0) qpd += qpd & 1;
    else qpd -= qpd & 1;
    double t = __dsub_rn(delta, __dmul_rn(3.14159265358979323846, (double)qpd));
    t = __dmul_rn((double)size / hop, t);
    t = t / 2.0;
    t = t / 3.14159265358979323846;
    t = __dadd_rn((double)j, t);
    t = __dmul_rn(t, rate);
    pvdata[2 * idx] = amp;
    pvdata[2 * idx + 1] = t / (double)size;
}
__global__ void pv_synth_kernel(cd *spec, double *phase_sum, const double *pvdata, long long nch, long long size,
                                double hop, double rate) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nch * size) return;
    double freq = pvdata[2 * idx + 1], amp = pvdata[2 * idx];
    double ph = __dmul_rn(__dmul_rn(2.0 * 3.14159265358979323846, freq) / rate, hop);  // pv.hpp:46-48
    double ps = __dadd_rn(phase_sum[idx], ph);
    phase_sum[idx] = ps;
    double s, c;
    sincos(ps, &s, &c);
    spec[idx] = mk<double>(__dmul_rn(amp, c), __dmul_rn(amp, s));  // std::polar
}
// shift scatter, pv.hpp:67-78: sources of destination k listed ascending in src[first .. first+count)
__global__ void pv_shift_kernel(const double *tmp, double *pvdata, const double *fn_freqs, const unsigned *off,
                                const unsigned *src, long long nch, long long size) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nch * size) return;
    long long c = idx / size, k = idx % size;
    double amp = 0, freq = 0;
    if (k < size / 2) {
        unsigned first = off[c * (size / 2) + k], last = off[c * (size / 2) + k + 1];
        for (unsigned s = first; s < last; s++) {
            unsigned j = src[s];
            amp = __dadd_rn(amp, tmp[2 * (c * size + j)]);
            freq = fn_freqs[c * (size / 2) + j];
        }
    }
    pvdata[2 * idx] = amp;
    pvdata[2 * idx + 1] = freq;
}

cudaError_t op_pv_analyze(const double *spec, double *last_phase, double *pvdata, long long nch, long long size,
                          double hop, double rate, cudaStream_t st) {
    if (nch * size == 0) return cudaSuccess;
    const int TB = 128;
    pv_analyze_kernel<<<(unsigned)((nch * size + TB - 1) / TB), TB, 0, st>>>(reinterpret_cast<const cd *>(spec), last_phase,
                                                                              pvdata, nch, size, hop, rate);
    return cudaGetLastError();
}
cudaError_t op_pv_synthesize(double *spec, double *phase_sum, const double *pvdata, long long nch, long long size,
                             double hop, double rate, cudaStream_t st) {
    if (nch * size == 0) return cudaSuccess;
    const int TB = 128;
    pv_synth_kernel<<<(unsigned)((nch * size + TB - 1) / TB), TB, 0, st>>>(reinterpret_cast<cd *>(spec), phase_sum, pvdata,
                                                                            nch, size, hop, rate);
    return cudaGetLastError();
}
cudaError_t op_pv_shift(const double *tmp, double *pvdata, const double *fn_freqs, const unsigned *off,
                        const unsigned *src, long long nch, long long size, cudaStream_t st) {
    if (nch * size == 0) return cudaSuccess;
    const int TB = 128;
    pv_shift_kernel<<<(unsigned)((nch * size + TB - 1) / TB), TB, 0, st>>>(tmp, pvdata, fn_freqs, off, src, nch, size);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ time-parallel PV bookkeeping
// Per (lane, bin): the running phase_sum of the sequential render, from the increments the pre-pass stored
// (inc[lane][frame][bin]): sum_rows[s] receives the value at the start of segment s, carry_out the value after the
// last frame.  The additions are the reference's own — one per frame, in frame order, from `base` — so the
// time-parallel render starts every segment from exactly the state the sequential render has there (summing
// per-segment totals first regroups them and drifts by ~ulp(phase_sum) per regrouping: 1e-7 of peak at 10 min).
// A CTA owns SCAN_BINS consecutive bins of one lane: all its threads stream the increments through shared memory
// (cp.async, tiles of SCAN_TILE frames, double-buffered: the chain of additions is short, the 8 bytes per frame
// and bin are what takes time), one thread per bin then adds a tile's frames in order.
constexpr int SCAN_BINS = 16, SCAN_TILE = 128, SCAN_THREADS = 256;
__global__ void __launch_bounds__(SCAN_THREADS)
timepar_scan_kernel(const double *inc, long long inc_frames, const long long *seg_tab, int n_seg, const double *base,
                    double *sum_rows, double *carry_out, long long lanes, long long m, long long lane0, long long carry_at,
                    const int *flag_in, int *flag_out, int epoch) {
    __shared__ __align__(16) double tile[2][SCAN_TILE][SCAN_BINS];
    const long long bin_blocks = (m + SCAN_BINS - 1) / SCAN_BINS;
    // lanes lane0 .. lane0 + gridDim.x / bin_blocks - 1; base / carry_out are indexed from lane0
    const long long sub = blockIdx.x / bin_blocks, lane = lane0 + sub, bin0 = (blockIdx.x % bin_blocks) * SCAN_BINS;
    const int tid = threadIdx.x, b = tid % SCAN_BINS, r0 = tid / SCAN_BINS;
    const bool bin_ok = bin0 + b < m;
    const double *src = inc + lane * inc_frames * m + bin0 + b;
    // frames covered: the segments are consecutive, [first_0, first_last + cnt_last)
    const long long f_begin = seg_tab[0], f_end = seg_tab[SEG_COLS * (n_seg - 1)] + seg_tab[SEG_COLS * (n_seg - 1) + 1];
    auto fetch = [&](int buf, long long f0) {
        for (int r = r0; r < SCAN_TILE; r += SCAN_THREADS / SCAN_BINS) {
            const long long f = f0 + r;
            if (bin_ok && f < f_end) {
                unsigned dst = (unsigned)__cvta_generic_to_shared(&tile[buf][r][b]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src + f * m) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    const long long idx = lane * m + bin0 + b, sidx = sub * m + bin0 + b;
    int s = 0;
    long long next_start = f_begin;
    // carry_out = the running sum before frame carry_at (the next block's first lead-in frame), or after the last frame
    const bool carry_inside = carry_at >= f_begin && carry_at < f_end;
    fetch(0, f_begin);
    // Chained over GPUs (tsb_timepar_scan_chain): `base` lies in THIS GPU's memory and is written by the previous
    // rank's kernel over NVLink, one CTA's 16 bins at a time, followed by that CTA's flag (= epoch, release at system
    // scope); carry_out / flag_out point into the NEXT rank's memory.  The first tiles are already on their way.
    if (flag_in && tid < SCAN_BINS) {
        const int *fp = flag_in + blockIdx.x;
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(fp) : "memory");
        } while (v != epoch);
    }
    double c = 0.0;
    if (tid < SCAN_BINS && bin_ok && base) {
        if (flag_in) asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(c) : "l"(base + sidx) : "memory");
        else c = base[sidx];
    }
    bool sent = false;
    auto send = [&]() {  // this CTA's bins are in carry_out: tell the next rank
        if (!flag_out || sent) return;
        sent = true;
        __threadfence_system();
        __syncwarp(tid < SCAN_BINS ? 0xffffu : 0xffff0000u);
        if (tid == 0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag_out + blockIdx.x), "r"(epoch) : "memory");
    };
    int buf = 0;
    for (long long f0 = f_begin; f0 < f_end; f0 += SCAN_TILE, buf ^= 1) {
        fetch(buf ^ 1, f0 + SCAN_TILE);  // (an empty group past the end)
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();
        if (tid < SCAN_BINS && bin_ok) {
            const int cnt = (int)((f_end - f0) < SCAN_TILE ? (f_end - f0) : SCAN_TILE);
            int r = 0;
            while (r < cnt) {
                if (f0 + r == next_start) {  // a segment starts here: it starts from the sum so far
                    if (sum_rows) sum_rows[(long long)s * lanes * m + idx] = c;
                    s++;
                    next_start = s < n_seg ? seg_tab[SEG_COLS * s] : -1;
                }
                if (carry_inside && f0 + r == carry_at && carry_out) {
                    carry_out[sidx] = c;
                    send();  // (all 16 threads of the CTA reach this frame together: same loop bounds)
                }
                // up to the next segment start (or the end of the tile) nothing but the additions: the loads do
                // not depend on the sum, so the compiler batches them ahead of the dependent chain
                int stop = (next_start > f0 + r && next_start < f0 + cnt) ? (int)(next_start - f0) : cnt;
                if (carry_inside && carry_at > f0 + r && carry_at < f0 + stop) stop = (int)(carry_at - f0);
#pragma unroll 8
                for (; r < stop; r++) c = __dadd_rn(c, tile[buf][r][b]);
            }
        }
        __syncthreads();
    }
    if (tid < SCAN_BINS && bin_ok) {
        for (; s < n_seg; s++)  // (segments without frames)
            if (sum_rows) sum_rows[(long long)s * lanes * m + idx] = c;
        if (carry_out && !carry_inside) carry_out[sidx] = c;
    }
    if (tid < SCAN_BINS && flag_out) send();  // (bins past m included: the flag is per CTA)
}
cudaError_t op_timepar_scan(const double *inc, long long inc_frames, const long long *seg_tab, int n_seg, const double *base,
                            double *sum_rows, double *carry_out, long long lanes, long long m, long long lane0,
                            long long n_lanes, cudaStream_t st, long long carry_at, const int *flag_in, int *flag_out,
                            int epoch) {
    if (n_lanes * m == 0 || n_seg <= 0) return cudaSuccess;
    const long long grid = ((m + SCAN_BINS - 1) / SCAN_BINS) * n_lanes;
    timepar_scan_kernel<<<(unsigned)grid, SCAN_THREADS, 0, st>>>(inc, inc_frames, seg_tab, n_seg, base, sum_rows, carry_out, lanes, m,
                                                                 lane0, carry_at, flag_in, flag_out, epoch);
    return cudaGetLastError();
}
// out[lane][pos + i] += tail[lane][i] (i < n): a segment's overlap tail lands on the head of the next segment's
// range.  One thread per sample, fixed order: deterministic, no atomics.
template <typename real>
__global__ void add_tail_kernel(real *out, long long out_stride, long long out_len, const long long *pos_tab, int n_tails,
                                const double *tails, long long lanes, long long n) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)n_tails * lanes * n) return;
    long long i = idx % n, l = (idx / n) % lanes, t = idx / (n * lanes);
    long long p = pos_tab[t] + i;
    double v = tails[(t * lanes + l) * n + i];
    if (v != 0.0 && p < out_len) out[l * out_stride + p] += (real)v;
}
cudaError_t op_add_tails(void *out, int is_f64, long long out_stride, long long out_len, const long long *pos_tab,
                         int n_tails, const double *tails, long long lanes, long long n, cudaStream_t st) {
    long long work = (long long)n_tails * lanes * n;
    if (work == 0) return cudaSuccess;
    const int TB = 256;
    if (is_f64)
        add_tail_kernel<double><<<(unsigned)((work + TB - 1) / TB), TB, 0, st>>>((double *)out, out_stride, out_len, pos_tab,
                                                                                  n_tails, tails, lanes, n);
    else
        add_tail_kernel<float><<<(unsigned)((work + TB - 1) / TB), TB, 0, st>>>((float *)out, out_stride, out_len, pos_tab,
                                                                                 n_tails, tails, lanes, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ frame gather, tinyspec.cpp:151-157
// frame[l][i] = (double) in[l][in_off + i], zero past the end of the input (planar float rows -> planar double).
__global__ void gather_kernel(const float *in, long long in_stride, long long in_len, long long in_off, double *frame,
                              long long lanes, long long w) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= lanes * w) return;
    long long l = idx / w, i = idx % w, s = in_off + i;
    frame[idx] = (s >= 0 && s < in_len) ? (double)in[l * in_stride + s] : 0.0;
}
cudaError_t op_gather(const float *in, long long in_stride, long long in_len, long long in_off, double *frame,
                      long long lanes, long long w, cudaStream_t st) {
    if (lanes * w == 0) return cudaSuccess;
    const int TB = 256;
    gather_kernel<<<(unsigned)((lanes * w + TB - 1) / TB), TB, 0, st>>>(in, in_stride, in_len, in_off, frame, lanes, w);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ overlap-add, tinyspec.cpp:170-200
// The reference keeps the overlap in a deque<float> `atmp` and narrows to float when it emits:
//   emit  i <  min(hop, osize):  out = float(double(tail[i] or 0) + frame[i])          (:173-182)
//   emit  osize <= i < hop:      out = tail[i] or 0                                    (:184-191)
//   keep  i <  osize - hop:      tail'[i] = (tail[hop+i] exists ? tail[hop+i] + float(frame[hop+i]) : float(frame[hop+i]))
//   keep  beyond that:           tail'[i] = tail[hop+i]                                (:194-200)
// tail_in / tail_out are distinct buffers [lanes][tail_cap]; tail_len is the deque's length before this frame.
__global__ void ola_kernel(const double *frame, long long osize, long long hop, float *out, long long out_stride,
                           long long out_len, long long out_off, const float *tail_in, float *tail_out, long long tail_cap,
                           long long tail_len, long long new_tail_len, long long lanes) {
    long long span = hop > new_tail_len ? hop : new_tail_len;
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= lanes * span) return;
    long long l = idx / span, i = idx % span;
    const double *fr = frame + l * osize;
    const float *ti = tail_in + l * tail_cap;
    if (i < hop) {
        float v;
        if (i < osize) {
            float overlap = i < tail_len ? ti[i] : 0.f;
            v = (float)((double)overlap + fr[i]);
        } else {
            v = i < tail_len ? ti[i] : 0.f;
        }
        long long p = out_off + i;
        if (p < out_len) out[l * out_stride + p] = v;
    }
    if (i < new_tail_len) {
        float t;
        bool have = hop + i < tail_len;
        if (hop + i < osize) {
            float out_val = (float)fr[hop + i];
            t = have ? ti[hop + i] + out_val : out_val;
        } else {
            t = ti[hop + i];
        }
        tail_out[l * tail_cap + i] = t;
    }
}
cudaError_t op_overlap_add(const double *frame, long long osize, long long hop, float *out, long long out_stride,
                           long long out_len, long long out_off, const float *tail_in, float *tail_out, long long tail_cap,
                           long long tail_len, long long new_tail_len, long long lanes, cudaStream_t st) {
    long long span = hop > new_tail_len ? hop : new_tail_len;
    if (lanes * span == 0) return cudaSuccess;
    const int TB = 256;
    ola_kernel<<<(unsigned)((lanes * span + TB - 1) / TB), TB, 0, st>>>(frame, osize, hop, out, out_stride, out_len, out_off,
                                                                         tail_in, tail_out, tail_cap, tail_len, new_tail_len,
                                                                         lanes);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ arithmetic peak probe (measurement aid)
// Independent FMA chains, enough of them per thread to cover the pipe latency, enough CTAs to fill every SM: the
// device's non-tensor FMA rate in the given type.  bench.py reports the chain kernel's arithmetic against it.
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *sink, T a, T b, int iters) {
    T v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = (T)(threadIdx.x + q);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = v[q] * a + b;
    }
    T acc = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) acc += v[q];
    if (acc == (T)123456789) sink[0] = acc;  // never true: keeps the chains alive
}
cudaError_t op_fma_peak(int is_f64, void *sink, int sms, cudaStream_t st, double *tflops) {
    const int ctas = sms * 8, iters = is_f64 ? 4096 : 16384;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, st);
        if (is_f64) fma_peak_kernel<double><<<ctas, 256, 0, st>>>((double *)sink, 1.0000001, 1e-9, iters);
        else fma_peak_kernel<float><<<ctas, 256, 0, st>>>((float *)sink, 1.0000001f, 1e-9f, iters);
        cudaEventRecord(e1, st);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const double fmas = (double)ctas * 256 * (double)iters * 32;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    return cudaSuccess;
}

}  // namespace tsb
