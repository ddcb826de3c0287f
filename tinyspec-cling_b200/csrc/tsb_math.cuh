// tsb_math.cuh — complex arithmetic, radix-2/4/8 butterflies and the shared-memory Stockham pass
// used by every FFT in this library (sm_100a).  Everything here is written as per-thread code with
// no barriers inside, so that the very same functions can also be driven thread-by-thread from a
// host test harness (tests/emul); barriers live in the callers.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TSB_HD __host__ __device__ __forceinline__
#define TSB_D __device__ __forceinline__
#else
#define TSB_HD inline
#define TSB_D inline
#endif

#define TSB_LDG(p) tsb::ldg(p)

namespace tsb {

// columns of a time-segment row (ChainArgs::seg_tab, see SegInfo in tsb_chain.cuh)
constexpr int SEG_COLS = 8;


template <typename T>
struct alignas(2 * sizeof(T)) cpx {
    T x, y;
};

// Read-only (non-coherent) global loads for the constant tables.  Most table entries a thread reads are the
// same every frame, and the compiler would hoist them out of the frame loop and then spill them (the register
// budget is what buys residency); `asm volatile` keeps each load where it is written — an L1 hit per frame.
#if defined(__CUDA_ARCH__)
TSB_D float ldg(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
TSB_D double ldg(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
TSB_D cpx<float> ldg(const cpx<float> *p) {
    cpx<float> r;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
TSB_D cpx<double> ldg(const cpx<double> *p) {
    cpx<double> r;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
TSB_D uint32_t ldg(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
TSB_D int ldg(const int *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
TSB_D int ldg(const int16_t *p) {
    int v;
    asm volatile("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
TSB_D int ldg(const uint16_t *p) {
    int v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
TSB_D int ldg(const uint8_t *p) {
    int v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
TSB_D long long ldg(const long long *p) {
    long long v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
#else
template <typename T>
inline T ldg(const T *p) { return *p; }
inline int ldg(const int16_t *p) { return *p; }
inline int ldg(const uint16_t *p) { return *p; }
inline int ldg(const uint8_t *p) { return *p; }
#endif
template <typename T>
TSB_HD cpx<T> mk(T x, T y) {
    cpx<T> r;
    r.x = x;
    r.y = y;
    return r;
}
template <typename T>
TSB_HD cpx<T> operator+(cpx<T> a, cpx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T>
TSB_HD cpx<T> operator-(cpx<T> a, cpx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T>
TSB_HD cpx<T> cmul(cpx<T> a, cpx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename T>
TSB_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by -i (forward-transform quarter turn)
template <typename T>
TSB_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }

#ifndef TSB_F32X2
#define TSB_F32X2 1  // 0: scalar; 1: packed complex sums; 2: packed complex products too (measured slower than 1)
#endif
#if defined(__CUDA_ARCH__) && TSB_F32X2 >= 1
// sm_100a packed single precision (FADD2 / FMUL2 / FFMA2: two IEEE operations per lane and issue slot): a complex
// float lives in an aligned register pair, so complex sums are one instruction.  Each half rounds exactly like the
// scalar operation (a - b as fma(b, -1, a): the product is exact).
TSB_D float2 f2(cpx<float> a) { return make_float2(a.x, a.y); }
TSB_D cpx<float> operator+(cpx<float> a, cpx<float> b) {
    const float2 r = __fadd2_rn(f2(a), f2(b));
    return mk<float>(r.x, r.y);
}
TSB_D cpx<float> operator-(cpx<float> a, cpx<float> b) {
    const float2 r = __ffma2_rn(f2(b), make_float2(-1.f, -1.f), f2(a));
    return mk<float>(r.x, r.y);
}
#if TSB_F32X2 >= 2
TSB_D cpx<float> cmul(cpx<float> a, cpx<float> b) {  // (a.x b.x - a.y b.y, a.x b.y + a.y b.x)
    const float2 t = __fmul2_rn(make_float2(a.x, a.x), f2(b));
    const float2 r = __ffma2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x), t);
    return mk<float>(r.x, r.y);
}
#endif
#endif

constexpr int ilog2(unsigned v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// ---- forward DFTs of size 2, 4, 8 (kernel exp(-2*pi*i*j*k/R)), natural order in and out ----
template <typename T>
TSB_HD void dft2(cpx<T> &a, cpx<T> &b) {
    cpx<T> t = a;
    a = t + b;
    b = t - b;
}
template <typename T>
TSB_HD void dft4(cpx<T> &v0, cpx<T> &v1, cpx<T> &v2, cpx<T> &v3) {
    cpx<T> t0 = v0 + v2, t1 = v0 - v2, t2 = v1 + v3, t3 = mul_mi(v1 - v3);
    v0 = t0 + t2;
    v2 = t0 - t2;
    v1 = t1 + t3;
    v3 = t1 - t3;
}
template <typename T>
TSB_HD void dft8(cpx<T> *v) {
    const T h = (T)0.70710678118654752440084436210485;
    cpx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    cpx<T> o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    // w8^k * O[k], w8 = exp(-i*pi/4)
    cpx<T> p1 = mk<T>((o1.x + o1.y) * h, (o1.y - o1.x) * h);
    cpx<T> p2 = mul_mi(o2);
    cpx<T> p3 = mk<T>((o3.y - o3.x) * h, -(o3.x + o3.y) * h);
    v[0] = e0 + o0;
    v[4] = e0 - o0;
    v[1] = e1 + p1;
    v[5] = e1 - p1;
    v[2] = e2 + p2;
    v[6] = e2 - p2;
    v[3] = e3 + p3;
    v[7] = e3 - p3;
}
template <int R, typename T>
TSB_HD void dftR(cpx<T> *v) {
    if (R == 8) dft8(v);
    else if (R == 4) dft4(v[0], v[1], v[2], v[3]);
    else dft2(v[0], v[1]);
}

// Radix plan for an M-point transform: as many radix-8 passes as possible, the remainder (2 or 4)
// taken LAST so that the twiddle-free first pass is a radix-8 one.
template <int LOG2M>
struct RadixPlan {
    static constexpr int n8 = LOG2M / 3;
    static constexpr int rem = LOG2M % 3;  // 0, 1 (radix 2) or 2 (radix 4)
    static constexpr int passes = n8 + (rem ? 1 : 0);
    static constexpr int radix(int p) { return p < n8 ? 8 : (1 << rem); }
    static constexpr int ns(int p) { return p <= n8 ? (1 << (3 * p)) : (1 << (3 * n8)); }
};

// Shared-memory swizzle of the FFT buffers.  A Stockham pass reads with unit stride across lanes but
// writes with stride R*Ns, which without a swizzle lands whole half-warps in one bank.  XOR-ing low index
// bits with higher ones makes every access pattern of every pass (and the reversed M-k accesses of the
// real-FFT split) conflict-free or within 15 % of it; the functions below were chosen by exhaustive
// search over XOR / padding families (see DESIGN.md "shared-memory layout").
// f is linear over GF(2): f(a ^ c) = f(a) ^ f(c), and a + c = a ^ c when a & c == 0, so an index made of
// a per-thread part and a compile-time part with disjoint bits costs one f() per thread plus one XOR
// with a constant per access.
template <typename T>
struct Swz;
template <>
struct Swz<float> {  // 8-byte elements: 16 lanes per wavefront
    static TSB_HD constexpr int f(int i) { return i ^ ((i >> 3) & 15); }
};
template <>
struct Swz<double> {  // 16-byte elements: 8 lanes per wavefront
    static TSB_HD constexpr int f(int i) { return i ^ (((i >> 3) ^ (i >> 4)) & 7); }
};

// Layout of a buffer transformed IN PLACE (stockham_pass_io<..., INPLACE>): every pass reads and writes base + t*Ns,
// i.e. with per-thread strides 1, 8, 64, ..., and the buffer is filled in digit-reversed order (stride M/8 across
// lanes).  Folding the second and third octal digits into the first makes all of them conflict-free.
template <typename T>
struct SwzInPlace;
template <>
struct SwzInPlace<float> {
    static TSB_HD constexpr int f(int i) { return i ^ (((i >> 4) ^ (i >> 7)) & 15); }
};
template <>
struct SwzInPlace<double> {
    static TSB_HD constexpr int f(int i) { return i ^ (((i >> 3) ^ (i >> 6)) & 7); }
};

// Per-pass twiddle tables: for pass P, entry [(t-1)*Ns + jm] = exp(-2*pi*i*jm*t/(Ns*R)), 1 <= t < R, jm < Ns.
// For a fixed t consecutive threads (consecutive jm) read consecutive entries: one coalesced 8/16-byte load per
// twiddle (the [jm][t] layout costs 8x the L1 wavefronts: measured, it doubled the kernel's L1 data-pipe load).
// Passes are laid out one after the other; pass 0 (Ns = 1) needs none.
template <int LOG2M>
struct TwLayout {
    using RP = RadixPlan<LOG2M>;
    static constexpr int offset(int p) { return p <= 1 ? 0 : offset(p - 1) + RP::ns(p - 1) * (RP::radix(p - 1) - 1); }
    static constexpr int total = offset(RP::passes);
};

// One thread's share of Stockham pass `P` of an M-point forward FFT.  Butterfly j takes inputs
// in(j, t*M/R) [index = sum of the two, bit-disjoint], applies twiddle exp(-2*pi*i*(j%Ns)*t/(Ns*R)), does a
// radix-R DFT and hands out(base, t*Ns, value) with base = (j/Ns)*Ns*R + (j%Ns).  `in` / `out` are
// functors so that the first pass can read (and window) the input ring and the last pass can feed the
// overlap-add directly.  `twp` is the per-pass twiddle table (TwLayout).
//
// INPLACE: the pass of the in-place decimation-in-time transform instead — the same butterflies and twiddles, but
// the inputs are read where the outputs go, in(base, t*Ns), so a pass needs no second buffer (each thread reads its
// R points before it writes them).  The data must then start in digit-reversed order (DitOrder<LOG2M>::pos).
template <typename T, int LOG2M, int THREADS, int P, class In, class Out, int TW_DERIVE = 0, bool INPLACE = false>
TSB_HD void stockham_pass_io(int tid, const cpx<T> *twp, In in, Out out) {
    using RP = RadixPlan<LOG2M>;
    constexpr int M = 1 << LOG2M;
    constexpr int R = RP::radix(P);
    constexpr int Ns = RP::ns(P);
    constexpr int NB = M / R;  // butterflies in this pass
    constexpr int ITER = (NB + THREADS - 1) / THREADS;
#pragma unroll
    for (int it = 0; it < ITER; it++) {
        int j = tid + it * THREADS;
        if (NB % THREADS != 0 && j >= NB) break;
        cpx<T> v[R];
        if (Ns > 1) {
            // the twiddles come from global memory (L1): request them before the shared-memory inputs so that
            // their longer latency runs behind those loads
            int jm = j & (Ns - 1);
            const cpx<T> *w = twp + TwLayout<LOG2M>::offset(P) + jm;
            cpx<T> tw[R];
            if (TW_DERIVE >= 1 && R == 8) {
                // fewer table reads (the L1/shared-memory data pipe is the busiest unit of the fp64 kernel): powers
                // of the loaded twiddles by complex multiplication, at most two products deep
                tw[1] = TSB_LDG(&w[0]);
                if (TW_DERIVE >= 2) {
                    tw[2] = cmul(tw[1], tw[1]);
                    tw[4] = cmul(tw[2], tw[2]);
                } else {
                    tw[2] = TSB_LDG(&w[Ns]);
                    tw[4] = TSB_LDG(&w[3 * Ns]);
                }
                tw[3] = cmul(tw[1], tw[2]);
                tw[5] = cmul(tw[1], tw[4]);
                tw[6] = cmul(tw[2], tw[4]);
                tw[7] = cmul(tw[3], tw[4]);
            } else {
#pragma unroll
                for (int t = 1; t < R; t++) tw[t] = TSB_LDG(&w[(t - 1) * Ns]);
            }
            int base = ((j - jm) * R) + jm;
#pragma unroll
            for (int t = 0; t < R; t++) v[t] = INPLACE ? in(base, t * Ns) : in(j, t * NB);
#pragma unroll
            for (int t = 1; t < R; t++) v[t] = cmul(v[t], tw[t]);
            dftR<R>(v);
            // last pass (Ns * R == M): base == j == tid + it * THREADS, so the per-thread part of the output index is
            // tid alone and the rest a compile-time constant (what lets a caller keep per-output state in registers)
            constexpr bool LAST = Ns * R == M && (THREADS & (THREADS - 1)) == 0;
#pragma unroll
            for (int t = 0; t < R; t++) {
                if (LAST) out(tid, it * THREADS + t * Ns, v[t]);
                else out(base, t * Ns, v[t]);
            }
        } else {
#pragma unroll
            for (int t = 0; t < R; t++) v[t] = INPLACE ? in(j * R, t) : in(j, t * NB);
            dftR<R>(v);
#pragma unroll
            for (int t = 0; t < R; t++) out(j * R, t, v[t]);
        }
    }
}

// Where sample k of the natural-order input must lie for the in-place passes (INPLACE above) to end in natural order:
// pass p joins R_p sub-transforms of Ns_p points that hold the samples congruent to t modulo R_p (decimation in time),
// so k = t_{P-1} + R_{P-1} (t_{P-2} + R_{P-2} (... t_0)) sits at sum t_p Ns_p — the digits of k, reversed.
template <int LOG2M>
struct DitOrder {
    using RP = RadixPlan<LOG2M>;
    static TSB_HD constexpr int pos(int k) {
        int at = 0;
        for (int p = RP::passes - 1; p >= 0; p--) {
            at += (k & (RP::radix(p) - 1)) * RP::ns(p);
            k /= RP::radix(p);
        }
        return at;
    }
};

// ---- 256-point transforms as two radix-16 passes (the spectral-synthesis kernels, tsb_spectral.cu) ----
// Sixteen threads run one sub-transform and hold all of its points in registers between the reads and the writes of a
// pass, so a pass may write the tile it read.  Position of point i inside a sub-transform's tile row: r16_sw(i)
// (conflict-free for every access of both passes with 16-byte elements).
TSB_HD constexpr int r16_sw(int i) { return i ^ ((i >> 4) & 7); }
// 16-point forward DFT, natural order in; output X[b + 4c] is left in v[4b + c]
template <typename T>
TSB_HD void dft16_t(cpx<T> *v) {
    using cd = cpx<T>;
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, h = (T)0.70710678118654752440;
#pragma unroll
    for (int a = 0; a < 4; a++) dft4(v[a], v[a + 4], v[a + 8], v[a + 12]);
    // v[a + 4b] *= W16^(a b), W16 = e^{-i pi/8}
    auto rot = [](cd x, T c, T s) { return mk<T>(x.x * c + x.y * s, x.y * c - x.x * s); };  // x * (c - i s)
    v[5] = rot(v[5], c1, s1);                                             // W^1
    v[9] = mk<T>((v[9].x + v[9].y) * h, (v[9].y - v[9].x) * h);      // W^2
    v[13] = rot(v[13], s1, c1);                                           // W^3
    v[6] = mk<T>((v[6].x + v[6].y) * h, (v[6].y - v[6].x) * h);      // W^2
    v[10] = mul_mi(v[10]);                                                // W^4 = -i
    v[14] = mk<T>((v[14].y - v[14].x) * h, -(v[14].x + v[14].y) * h);  // W^6
    v[7] = rot(v[7], s1, c1);                                             // W^3
    v[11] = mk<T>((v[11].y - v[11].x) * h, -(v[11].x + v[11].y) * h);  // W^6
    v[15] = rot(v[15], -c1, -s1);                                         // W^9 = -W^1
#pragma unroll
    for (int b = 0; b < 4; b++) dft4(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
}
// One of the two radix-16 passes of a 256-point forward transform, thread j of 16.  PASS 0: y[16 j + t'] = DFT16_t(x[j + 16 t]);
// PASS 1: X[j + 16 t'] = DFT16_t(W256^(j t) y[j + 16 t]).  `w` = W256^j (pass 1): its powers are formed by products, at
// most four deep.  The caller separates the reads of a pass from its writes when both use the same buffer
// (rd(i) -> value, then sync(), then wr(i, value)).
// second half of a pass: twiddles (pass 1), the 16-point DFT, the stores.  HALF: only outputs X[0..127] are wanted
template <int PASS, bool HALF = false, typename T, class Wr>
TSB_HD void r16_finish(int j, cpx<T> w, cpx<T> (&v)[16], Wr wr) {
    using cd = cpx<T>;
    if (PASS == 1) {
        const cd p2 = cmul(w, w), p4 = cmul(p2, p2), p8 = cmul(p4, p4), p3 = cmul(p2, w);
        v[1] = cmul(v[1], w);
        v[2] = cmul(v[2], p2);
        v[3] = cmul(v[3], p3);
        v[4] = cmul(v[4], p4);
        v[5] = cmul(v[5], cmul(p4, w));
        v[6] = cmul(v[6], cmul(p4, p2));
        const cd p7 = cmul(p4, p3);
        v[7] = cmul(v[7], p7);
        v[8] = cmul(v[8], p8);
        v[9] = cmul(v[9], cmul(p8, w));
        v[10] = cmul(v[10], cmul(p8, p2));
        v[11] = cmul(v[11], cmul(p8, p3));
        v[12] = cmul(v[12], cmul(p8, p4));
        v[13] = cmul(v[13], cmul(p8, cmul(p4, w)));
        v[14] = cmul(v[14], cmul(p8, cmul(p4, p2)));
        v[15] = cmul(v[15], cmul(p8, p7));
    }
    dft16_t(v);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int tp = (r >> 2) + 4 * (r & 3);  // v[4b + c] holds output b + 4c
        if (HALF && (tp > 8 || (tp == 8 && j != 0))) continue;  // (PASS 1: X[j + 16 tp], tp >= 8, is the upper half; X[128] is kept)
        wr(PASS == 0 ? 16 * j + tp : j + 16 * tp, v[r]);
    }
}
template <int PASS, bool HALF = false, typename T, class Rd, class Sync, class Wr>
TSB_HD void r16_pass(int j, cpx<T> w, Rd rd, Sync sync, Wr wr) {
    cpx<T> v[16];
#pragma unroll
    for (int t = 0; t < 16; t++) v[t] = rd(j + 16 * t);
    sync();
    r16_finish<PASS, HALF>(j, w, v, wr);
}

// plain buffer-to-buffer pass (unswizzled), used by the unit tests and the standalone FFT kernels
template <typename T, int LOG2M, int THREADS, int P>
TSB_HD void stockham_pass(int tid, const cpx<T> *src, cpx<T> *dst, const cpx<T> *twp) {
    stockham_pass_io<T, LOG2M, THREADS, P>(
        tid, twp, [&](int a, int c) { return src[a + c]; }, [&](int a, int c, cpx<T> v) { dst[a + c] = v; });
}

}  // namespace tsb
