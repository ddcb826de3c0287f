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
//     kernel A  generator + G (bins k and n-k together, from the same two spectrum values) + n2 column FFTs of n1 points
//               + twiddle e^{-2 pi i k2 j1 / n}  -> scratch[j1][k2]
//     kernel B  n1 row FFTs of n2 points; only j < n/2 is needed (the first n samples): unpack, 1/sqrt(2n), the
//               carried bins 0 / n added analytically, window, overlap-add into the output rows
// The scratch of one frame (channels * n * 16 bytes; 64 MiB at cfg5) is written and read back to back, i.e. from L2.
// Both kernels are bound by the L1/shared-memory data pipe (a complex double costs a warp four wavefronts per access), so
// 256-point sub-transforms run as two radix-16 passes by 16 threads that hold the whole sub-transform in registers between
// the reads and the writes of a pass (SubFft::R16), and kernel A generates straight into its first pass (mirror elements
// of the Hermitian pairs cross the warp by shuffle).  Kernel A of frame f+1 and kernel B of frame f run on two streams.
#include <cuda_runtime.h>

#include <cstdlib>

#include <cmath>
#include <map>
#include <vector>

#include "../../include/tsb200.h"
#include "tsb_host.hpp"
#include "tsb_math.cuh"
#include "tsb_spectral.hpp"

namespace tsb {

using cd = cpx<double>;

// ---- generators: the spectrum value of bin i (0 < i < n) of channel c ----
// The seeded integer formula of oracle/ref/taps/bytebeat_tap.cpp / examples/bytebeat.cpp.
struct GenHash {
    static constexpr bool in_registers = true;  // no memory behind it: kernel A may generate straight into the first pass
    uint32_t frame;
    __device__ __forceinline__ cd operator()(int c, int i) const {
        uint32_t h = frame * 2654435761u ^ ((uint32_t)c * 40503u + (uint32_t)i * 2246822519u);
        h ^= h >> 15;
        h *= 2246822519u;
        h ^= h >> 13;
        // x / 32768.0 / (double)(1 + (i >> 4)): the first division is exact; the second as a Newton reciprocal
        // (shared by both components) and one exact-remainder correction of each quotient (Markstein).  (A table of
        // correctly rounded reciprocals saves a tenth of the kernel's instructions and costs time: the load's latency
        // sits in front of everything else the iteration does — measured, profiles/README.md)
        const double d = (double)(1 + (i >> 4));
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
        double e = fma(-d, r, 1.0);
        r = fma(r, e, r);
        e = fma(-d, r, 1.0);
        r = fma(r, e, r);
        const double a = (double)((int)(h & 0xffff) - 32768) * (1.0 / 32768.0);
        const double b = (double)((int)(h >> 16) - 32768) * (1.0 / 32768.0);
        double qa = a * r, qb = b * r;
        qa = fma(fma(-d, qa, a), r, qa);
        qb = fma(fma(-d, qb, b), r, qb);
        return mk<double>(qa, qb);
    }
};
// Spectrum values computed by the caller: device complex128 rows [channels][n] of this frame, bins 1..n-1 are read.
struct GenBuffer {
    static constexpr bool in_registers = false;  // column reads want neighbouring columns in neighbouring lanes
    const cd *spec;
    long long n;
    __device__ __forceinline__ cd operator()(int c, int i) const { return spec[(long long)c * n + i]; }
};

// A spectrum expression (TSB_GEN_EXPR, include/tsb200.h): the per-bin formula of a spectral_bytebeat hack
// (hacks/spectral_bytebeat/voidbubbles.cpp:9-13 and its siblings) as a postfix program on a stack of 64-bit slots.
// The program travels as a kernel parameter (constant bank: every thread reads the same word, a uniform load); the
// stack is a small local array.  The host has type-checked the program (expr_check): depth and types are static.
struct ExprCode {
    int n_ops;
    int ops[TSB_EXPR_MAX_OPS];
    long long ic[TSB_EXPR_MAX_CONST];
    double dc[TSB_EXPR_MAX_CONST];
};
struct GenExpr {
    static constexpr bool in_registers = true;
    ExprCode code;
    const int *itab;         // device [n_itab][n]
    const double *dtab;      // device [n_dtab][n]
    const long long *ipar;   // device, this frame's integer parameters
    const double *dpar;      // device, this frame's double parameters
    int n;
    long long frame;
    __device__ __forceinline__ static long long as_i(double v) { return __double_as_longlong(v); }
    __device__ __forceinline__ static double as_d(long long v) { return __longlong_as_double(v); }
    __device__ __noinline__ cd operator()(int c, int i) const {
        double st[TSB_EXPR_STACK];
        int sp = 0;  // next free slot
        for (int pc = 0; pc < code.n_ops; pc++) {
            const int w = code.ops[pc], arg = w >> 8;
            switch (w & 255) {
#define TSB_PUSH_I(v) st[sp++] = as_d((long long)(v))
#define TSB_PUSH_D(v) st[sp++] = (v)
#define TSB_BIN_I(expr)                                          \
    {                                                            \
        const long long b = as_i(st[sp - 1]), a = as_i(st[sp - 2]); \
        sp--;                                                    \
        st[sp - 1] = as_d((long long)(expr));                    \
    }
#define TSB_BIN_D(expr)                              \
    {                                                \
        const double b = st[sp - 1], a = st[sp - 2]; \
        sp--;                                        \
        st[sp - 1] = (expr);                         \
    }
                case TSB_X_I_BIN: TSB_PUSH_I(i); break;
                case TSB_X_I_CH: TSB_PUSH_I(c); break;
                case TSB_X_I_FRAME: TSB_PUSH_I(frame); break;
                case TSB_X_I_N: TSB_PUSH_I(n); break;
                case TSB_X_I_CONST: TSB_PUSH_I(code.ic[arg]); break;
                case TSB_X_I_TAB: TSB_PUSH_I(__ldg(&itab[(long long)arg * n + i])); break;
                case TSB_X_I_PARAM: TSB_PUSH_I(__ldg(&ipar[arg])); break;
                case TSB_X_D_CONST: TSB_PUSH_D(code.dc[arg]); break;
                case TSB_X_D_TAB: TSB_PUSH_D(__ldg(&dtab[(long long)arg * n + i])); break;
                case TSB_X_D_PARAM: TSB_PUSH_D(__ldg(&dpar[arg])); break;
                // two's-complement wrap-around, as every compiler the reference is built with does for `int`
                case TSB_X_I_ADD: TSB_BIN_I((unsigned long long)a + (unsigned long long)b) break;
                case TSB_X_I_SUB: TSB_BIN_I((unsigned long long)a - (unsigned long long)b) break;
                case TSB_X_I_MUL: TSB_BIN_I((unsigned long long)a * (unsigned long long)b) break;
                case TSB_X_I_DIV: TSB_BIN_I(b == 0 ? 0 : (b == -1 ? (long long)(0ull - (unsigned long long)a) : a / b)) break;
                case TSB_X_I_MOD: TSB_BIN_I((b == 0 || b == -1) ? 0 : a % b) break;
                case TSB_X_I_AND: TSB_BIN_I(a & b) break;
                case TSB_X_I_OR: TSB_BIN_I(a | b) break;
                case TSB_X_I_XOR: TSB_BIN_I(a ^ b) break;
                case TSB_X_I_SHL: TSB_BIN_I((unsigned long long)a << (b & 63)) break;
                case TSB_X_I_SHR: TSB_BIN_I(a >> (b & 63)) break;
                case TSB_X_I_EQ: TSB_BIN_I(a == b) break;
                case TSB_X_I_LT: TSB_BIN_I(a < b) break;
                case TSB_X_I_NEG: st[sp - 1] = as_d((long long)(0ull - (unsigned long long)as_i(st[sp - 1]))); break;
                case TSB_X_I_NOT: st[sp - 1] = as_d(~as_i(st[sp - 1])); break;
                case TSB_X_I_WRAP32: st[sp - 1] = as_d((long long)(int)(unsigned)(unsigned long long)as_i(st[sp - 1])); break;
                case TSB_X_I2D: st[sp - 1] = (double)as_i(st[sp - 1]); break;
                case TSB_X_D2I: st[sp - 1] = as_d((long long)st[sp - 1]); break;  // C truncation (saturating out of range)
                case TSB_X_D_ADD: TSB_BIN_D(__dadd_rn(a, b)) break;
                case TSB_X_D_SUB: TSB_BIN_D(__dsub_rn(a, b)) break;
                case TSB_X_D_MUL: TSB_BIN_D(__dmul_rn(a, b)) break;
                case TSB_X_D_DIV: TSB_BIN_D(__ddiv_rn(a, b)) break;
                case TSB_X_D_FMOD: TSB_BIN_D(fmod(a, b)) break;
                case TSB_X_D_POW: TSB_BIN_D(pow(a, b)) break;
                case TSB_X_D_NEG: st[sp - 1] = -st[sp - 1]; break;
                case TSB_X_D_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
                case TSB_X_D_SIN: st[sp - 1] = sin(st[sp - 1]); break;
                case TSB_X_D_COS: st[sp - 1] = cos(st[sp - 1]); break;
                case TSB_X_D_EXP: st[sp - 1] = exp(st[sp - 1]); break;
                case TSB_X_D_SQRT: st[sp - 1] = __dsqrt_rn(st[sp - 1]); break;
                case TSB_X_D_FLOOR: st[sp - 1] = floor(st[sp - 1]); break;
                case TSB_X_PICK: st[sp] = st[sp - 1 - arg]; sp++; break;
                case TSB_X_POLAR: {
                    double sn, cs;
                    sincos(st[sp - 1], &sn, &cs);
                    const double mag = st[sp - 2];
                    st[sp - 2] = __dmul_rn(mag, cs);
                    st[sp - 1] = __dmul_rn(mag, sn);
                } break;
#undef TSB_PUSH_I
#undef TSB_PUSH_D
#undef TSB_BIN_I
#undef TSB_BIN_D
                default: break;
            }
        }
        return mk<double>(st[0], st[1]);
    }
};

// One CTA runs C sub-transforms of M = 2^LM points side by side, T threads each (one radix-8 butterfly per thread and
// pass): C * T = 256 threads and 64 KB of tiles, fewer when the other dimension (2^LO sub-transforms in all) is smaller.
#ifndef TSB_SPEC_CAP_A
#define TSB_SPEC_CAP_A 2048  // points per tile of kernel A (columns x n1)
#endif
#ifndef TSB_SPEC_CAP_B
#define TSB_SPEC_CAP_B 2048  // ... of kernel B (rows x n2)
#endif
#ifndef TSB_SPEC_R16
#define TSB_SPEC_R16 1  // 256-point sub-transforms as two radix-16 passes (16 threads each) instead of 8 x 8 x 4
#endif
template <int LM, int LO, int CAP = 2048>
struct SubFft {
    static constexpr int M = 1 << LM;
    // 256 = 16 x 16: two passes, i.e. one shared-memory round trip less than 8 x 8 x 4 (both spectral kernels are bound
    // by the L1/shared-memory data pipe: 16-byte points cost four wavefronts per warp access), and the 16 threads of a
    // sub-transform hold all of its points in registers between reading and writing a pass, so one buffer is enough
    static constexpr bool R16 = TSB_SPEC_R16 && LM == 8;
    static constexpr int T = R16 ? 16 : (M / 8 < 1 ? 1 : M / 8);
    static constexpr int CMAX = (CAP / M) < 2 ? 2 : (CAP / M);
    static constexpr int C = CMAX < (1 << LO) ? CMAX : (1 << LO);
    static constexpr int NT = C * T;
    static constexpr int BUFS = R16 ? 1 : 2;  // work tiles of C * (M + 1) points
    // position of point i inside a sub-transform's tile row
    static TSB_HD constexpr int sw(int i) { return R16 ? r16_sw(i) : Swz<double>::f(i); }
};
__device__ __forceinline__ cd shfl_xor_cd(cd v, int mask) {
    return mk<double>(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
// LIN: the first pass reads its input unswizzled (a tile just copied in from global memory)
template <int LOG2M, int T, int P, int END, bool LIN = false>
struct SpecPasses {
    static __device__ __forceinline__ cd *run(int lt, const cd *src, cd *dst, cd *other, const cd *tw) {
        stockham_pass_io<double, LOG2M, T, P>(
            lt, tw,
            [&](int a, int c) { return (LIN && P == 0) ? src[a + c] : src[Swz<double>::f(a) ^ Swz<double>::f(c)]; },
            [&](int a, int c, cd v) { dst[Swz<double>::f(a) ^ Swz<double>::f(c)] = v; });
        // (radix-8 plan) a sub-transform of 256 points is run by exactly one warp (T = 32): its passes need no block barrier
        if (T == 32) __syncwarp();
        else __syncthreads();
        return SpecPasses<LOG2M, T, P + 1, END, LIN>::run(lt, dst, other, dst, tw);
    }
};
template <int LOG2M, int T, int END, bool LIN>
struct SpecPasses<LOG2M, T, END, END, LIN> {
    static __device__ __forceinline__ cd *run(int, const cd *src, cd *, cd *, const cd *) { return const_cast<cd *>(src); }
};

// e^{-i pi q / n} for 0 <= q < 2n, argument exact
__device__ __forceinline__ cd unit_w(long long q, double inv_n) {
    double s, c;
    sincospi(-(double)q * inv_n, &s, &c);
    return mk<double>(c, s);
}

#ifndef TSB_SPEC_UG
#define TSB_SPEC_UG 8  // unroll factor of kernel A's tile-path generation loop (caller spectra: 8 column reads in flight per thread, 56.1 -> 54.3 us)
#endif
#ifndef TSB_SPEC_UO
#define TSB_SPEC_UO 2  // ... of its output loop (2: two shared-memory reads in flight per thread; 54.5 us per cfg5 frame instead of 55.8)
#endif
// resident CTAs per SM the kernels are compiled for (register budget): A 3 of 256 threads (4 of the 128 threads of the
// radix-16 variant: 128 registers for the 16 points a thread holds); B 2 (3 of 128 threads with radix 16)
#ifdef TSB_SPEC_OCC_A
constexpr int spec_occ_a(int, bool) { return TSB_SPEC_OCC_A; }
#else
constexpr int spec_occ_a(int nt, bool r16) { return r16 ? (nt >= 256 ? 2 : 4) : (nt >= 256 ? 3 : (nt >= 128 ? 6 : 8)); }
#endif
#ifdef TSB_SPEC_OCC_B
constexpr int spec_occ_b(bool) { return TSB_SPEC_OCC_B; }
#else
constexpr int spec_occ_b(bool r16) { return r16 ? 3 : 2; }
#endif
// ---- kernel A: one CTA = C/2 columns k2 of one channel together with their mirror columns n2 - k2 ----
// G[k] and G[n-k] are made of the same two spectrum values: with a = X[k], b = X[n-k], sm = a + conj(b),
// wd = W^k (a - conj(b)) (and W^(n-k) = -conj(W^k)):   G[k] = sm + i wd,   G[n-k] = conj(sm) + i conj(wd).
// So a CTA takes column pairs (k2, n2 - k2) — the mirror of row k1 of column k2 is row n1-1-k1 of column n2-k2 — and
// every spectrum value is generated once instead of twice (for an interpreted expression that is half the kernel).
// Columns 0 and n2/2 mirror into themselves (rows k1 <-> n1-k1 and k1 <-> n1-1-k1): they share the pair slot that
// column 0 would otherwise leave half empty.  Tile slot 2p holds the low column of pair p, slot 2p+1 its mirror.
// wa[k1] = e^{-i pi k1 / n1} (n1 entries), wb[k2] = e^{-i pi k2 / n} (n2 entries): W^k = wa[k1] * wb[k2] for k = n2*k1 + k2
template <int L1, int L2, class Gen>
__global__ void __launch_bounds__(SubFft<L1, L2, TSB_SPEC_CAP_A>::NT, spec_occ_a(SubFft<L1, L2, TSB_SPEC_CAP_A>::NT, SubFft<L1, L2, TSB_SPEC_CAP_A>::R16))
spectral_cols_kernel(Gen gen, cd *scratch, const cd *tw1, const cd *wa, const cd *wb, int channels) {
    using F = SubFft<L1, L2, TSB_SPEC_CAP_A>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2, C = F::C, NT = F::NT, CP = C / 2;
    constexpr long long n = (long long)n1 * n2;
    static_assert(C >= 2 && C % 2 == 0 && n1 >= 2 && n2 >= 4, "column pairs need an even tile and n2 >= 4");
    // sub-transform c lives at c * SP of a tile: the stride is padded by one element so that the C columns a warp touches
    // together (consecutive threads = consecutive columns) fall into different banks
    constexpr int SP = n1 + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *bufA = reinterpret_cast<cd *>(smem_raw), *bufB = bufA + C * SP;  // (bufB: only with two work tiles, F::BUFS)
    const int groups = n2 / C;
    // generation: a thread keeps one column pair (p) and walks k1 in steps of NT / CP
    constexpr int STEPG = NT / CP;
    constexpr int UG = TSB_SPEC_UG, UO = TSB_SPEC_UO;
    const int p = threadIdx.x % CP, rg = threadIdx.x / CP;
    // output: a thread keeps one tile slot (c) and walks j1 in steps of NT / C
    constexpr int STEP = NT / C;
    const int c = threadIdx.x % C, r0 = threadIdx.x / C;
    for (long long b = blockIdx.x; b < (long long)channels * groups; b += gridDim.x) {
        const int ch = (int)(b / groups), g = (int)(b % groups);
        const int col_lo = g * CP + p;
        const bool self = col_lo == 0;  // the pair slot of the two self-mirrored columns, 0 and n2/2
        const int gi = threadIdx.x / F::T, lt = threadIdx.x % F::T;
        constexpr bool FUSE = F::R16 && Gen::in_registers;
        cd *res;
        if (FUSE && g != 0) {
            // Generation straight into the first radix-16 pass.  Lane lt of the low column of a pair (lanes 0..15 of the
            // warp) needs rows lt + 16 t of its column; the mirror of row k1 is row n1-1-k1 of the mirror column, whose
            // lanes 16..31 run the other way round: lane L and lane 31 - L need exactly each other's mirrors.  Each
            // generates the eight pairs anchored at rows lt + 16 t, t < 8, of its OWN column (G[k] = sm + i W^k d holds
            // for any anchor k), keeps that element and hands the mirror element over by shuffle: it is the
            // partner's row (15 - lt) + 16 (15 - t).  No shared-memory round trip between generator and transform.
            if constexpr (FUSE) {
                const int col_own = (gi & 1) ? n2 - (g * CP + (gi >> 1)) : g * CP + (gi >> 1);
                const cd wk = TSB_LDG(&wb[col_own]);
                cd v[16];
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const int k1 = lt + 16 * t;
                    const int k = n2 * k1 + col_own;
                    const cd a = gen(ch, k), bb = conj(gen(ch, (int)n - k));
                    const cd sm = a + bb, d = a - bb;
                    const cd wd = cmul(cmul(TSB_LDG(&wa[k1]), wk), d);
                    v[t] = mk<double>(sm.x - wd.y, sm.y + wd.x);
                    v[15 - t] = shfl_xor_cd(mk<double>(sm.x + wd.y, wd.x - sm.y), 31);
                }
                cd *col = bufA + gi * SP;
                r16_finish<0>(lt, wk, v, [&](int i, cd x) { col[F::sw(i)] = x; });
            }
        } else {
        const cd wk_lo = TSB_LDG(&wb[col_lo]);
        if (g != 0) {
            // no self-mirrored column in this group (every group but the first): bin ka of column col_lo pairs with bin
            // n - ka = (n1 - 1 - k1, n2 - col_lo) of the mirror column — straight-line code in 32-bit indices
#pragma unroll UG
            for (int it = 0; it < (n1 + STEPG - 1) / STEPG; it++) {
                const int k1 = rg + it * STEPG;
                if (n1 % STEPG != 0 && k1 >= n1) break;
                const int ka = n2 * k1 + col_lo;
                const cd a = gen(ch, ka), bb = conj(gen(ch, (int)n - ka));
                const cd sm = a + bb, d = a - bb;
                const cd wd = cmul(cmul(TSB_LDG(&wa[k1]), wk_lo), d);
                bufA[(2 * p) * SP + F::sw(k1)] = mk<double>(sm.x - wd.y, sm.y + wd.x);
                bufA[(2 * p + 1) * SP + F::sw(n1 - 1 - k1)] = mk<double>(sm.x + wd.y, wd.x - sm.y);
            }
        } else {
#pragma unroll UG
        for (int it = 0; it < (n1 + STEPG - 1) / STEPG; it++) {
            const int k1 = rg + it * STEPG;
            if (n1 % STEPG != 0 && k1 >= n1) break;
            // element A = (slot_a, k1a) of column col_a, bin ka; its mirror B = bin n - ka
            int slot_a = 2 * p, k1a = k1, col_a = col_lo;
            if (self && k1 >= n1 / 2) {
                slot_a = 2 * p + 1;
                k1a = k1 - n1 / 2;
                col_a = n2 / 2;
            }
            long long ka = (long long)n2 * k1a + col_a;
            cd wk = self ? TSB_LDG(&wb[col_a]) : wk_lo;
            const bool first = ka == 0;  // bin 0 is a carried one (kernel B adds its contribution): G[0] = 0, and this
            if (first) {                 // iteration takes the bin that mirrors into itself instead, n/2
                bufA[slot_a * SP + F::sw(0)] = mk<double>(0.0, 0.0);
                k1a = n1 / 2;
                ka = n / 2;
            }
            const long long kb = n - ka;
            const cd a = gen(ch, (int)ka), bb = conj(first ? a : gen(ch, (int)kb));
            const cd sm = a + bb, d = a - bb;
            const cd wd = cmul(cmul(TSB_LDG(&wa[k1a]), wk), d);              // W^k * d
            bufA[slot_a * SP + F::sw(k1a)] = mk<double>(sm.x - wd.y, sm.y + wd.x);  // sm + i*W^k*d
            if (!first) {
                const int slot_b = self ? slot_a : 2 * p + 1, k1b = (int)(kb >> L2);
                bufA[slot_b * SP + F::sw(k1b)] = mk<double>(sm.x + wd.y, wd.x - sm.y);  // conj(sm) + i*conj(wd)
            }
        }
        }
        __syncthreads();
        }
        if constexpr (F::R16) {
            // two radix-16 passes in the one tile: the 16 threads of a column (half a warp) hold all 256 points
            // between the reads and the writes of a pass
            cd *col = bufA + gi * SP;
            const cd w = TSB_LDG(&tw1[TwLayout<8>::offset(2) + lt]);  // W256^lt
            auto rd = [&](int i) { return col[F::sw(i)]; };
            auto wr = [&](int i, cd v) { col[F::sw(i)] = v; };
            auto sync = [] { __syncwarp(); };
            if (!(FUSE && g != 0)) r16_pass<0>(lt, w, rd, sync, wr);
            __syncwarp();
            r16_pass<1>(lt, w, rd, sync, wr);
            res = bufA;
        } else {
            res = SpecPasses<L1, F::T, 0, RadixPlan<L1>::passes>::run(lt, bufA + gi * SP, bufB + gi * SP, bufA + gi * SP, tw1) - gi * SP;
        }
        __syncthreads();
        {   // A[k2][j1] * e^{-2 pi i k2 j1 / n} -> scratch[ch][j1][k2]: the factor is base^j1 with base = wb[k2]^2; the
            // thread starts at base^r0 (square-and-multiply) and advances by base^STEP (at most 8 steps: no drift)
            const int lo = g * CP + (c >> 1);
            const int k2 = (c & 1) ? (lo == 0 ? n2 / 2 : n2 - lo) : lo;
            const cd wk2 = TSB_LDG(&wb[k2]);
            const cd base = cmul(wk2, wk2);
            cd w = mk<double>(1.0, 0.0), pw = base;
#pragma unroll
            for (int bit = 0; (1 << bit) < STEP; bit++) {
                if ((r0 >> bit) & 1) w = cmul(w, pw);
                pw = cmul(pw, pw);
            }
            const cd out_step = pw;  // base^STEP (STEP is a power of two)
            cd *dst = scratch + (long long)ch * n + k2;
#pragma unroll UO
            for (int it = 0; it < (n1 + STEP - 1) / STEP; it++) {
                const int j1 = r0 + it * STEP;
                if (n1 % STEP != 0 && j1 >= n1) break;
                dst[(long long)j1 * n2] = cmul(res[c * SP + F::sw(j1)], w);
                w = cmul(w, out_step);
            }
        }
        __syncthreads();
    }
}

// ---- kernel B: persistent CTAs, each walking tiles of C consecutive rows j1 of one channel ----
// A tile (C rows of n2 points, contiguous in the scratch) arrives by ONE bulk asynchronous copy (cp.async.bulk, the
// 1-D TMA path) that completes on an mbarrier; the copy of the next tile is issued as soon as the first pass has
// consumed the staging buffer and runs behind the remaining passes and the overlap-add of the current one.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long *bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TSB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TSB_DONE_%=;\n"
        "bra TSB_WAIT_%=;\n"
        "TSB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
template <int L1, int L2>
__global__ void __launch_bounds__(SubFft<L2, L1, TSB_SPEC_CAP_B>::NT, spec_occ_b(SubFft<L2, L1, TSB_SPEC_CAP_B>::R16))
spectral_rows_kernel(const cd *scratch, const cd *tw2, int channels, double *out, long long out_stride, long long out_len,
                     long long pos0, const double *edge_in, double *edge_out, int pairs_aligned) {
    using F = SubFft<L2, L1, TSB_SPEC_CAP_B>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2, C = F::C, NT = F::NT;
    constexpr long long n = (long long)n1 * n2;
    constexpr int SP = n2 + 1;  // padded stride of a sub-transform inside a work tile (see kernel A)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *stage = reinterpret_cast<cd *>(smem_raw);           // [C][n2], as it lies in the scratch
    cd *bufA = stage + C * n2, *bufB = bufA + C * SP;
    __shared__ __align__(8) unsigned long long bar;
    const double inv_sqrt = 1.0 / sqrt((double)(2 * n));
    const double half_n = (double)n / 2.0;
    const int groups = n1 / C;
    const long long tiles = (long long)channels * groups;
    constexpr unsigned TILE_BYTES = (unsigned)(C * n2 * sizeof(cd));
    constexpr int STEP = NT / C, NOUT = (n2 / 2 + STEP - 1) / STEP;  // output pairs per thread
    const int c = threadIdx.x % C, r0 = threadIdx.x / C;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && (long long)blockIdx.x < tiles) bulk_load(stage, scratch + (long long)blockIdx.x * C * n2, TILE_BYTES, &bar);
    unsigned phase = 0;
    cd w16 = mk<double>(1.0, 0.0);  // W256^lt of the radix-16 passes
    if constexpr (F::R16) w16 = TSB_LDG(&tw2[TwLayout<8>::offset(2) + threadIdx.x % F::T]);
    for (long long b = blockIdx.x; b < tiles; b += gridDim.x) {
        const int ch = (int)(b / groups), j1_0 = (int)(b % groups) * C;
        // the carried bins 0 and n of the spectrum (requested here, used after the transform)
        const double d = edge_in ? edge_in[2 * ch] : 0.0, q = edge_in ? edge_in[2 * ch + 1] : 0.0;
        // the output samples this thread will add to (v[j], j = j1 + n1*j2 < n/2 -> samples 2j, 2j+1): requested now,
        // they arrive while the rows are transformed
        double *row = out + (long long)ch * out_stride + pos0;
        double2 acc[NOUT];
#pragma unroll
        for (int u = 0; u < NOUT; u++) {
            const int j2 = r0 + u * STEP;
            const long long m = 2 * ((long long)(j1_0 + c) + (long long)n1 * j2);
            acc[u] = make_double2(0.0, 0.0);
            if (j2 < n2 / 2) {
                if (pairs_aligned && pos0 + m + 1 < out_len) acc[u] = *reinterpret_cast<const double2 *>(row + m);
                else {
                    if (pos0 + m < out_len) acc[u].x = row[m];
                    if (pos0 + m + 1 < out_len) acc[u].y = row[m + 1];
                }
            }
        }
        bar_wait(&bar, phase);  // the tile is in `stage`
        phase ^= 1;
        const int gi = threadIdx.x / F::T, lt = threadIdx.x % F::T;
        cd *res;
        if constexpr (F::R16) {
            cd *rowb = bufA + gi * SP;
            // first pass: stage (as it lies in the scratch) -> the work tile; then the staging buffer is free
            r16_pass<0>(lt, w16, [&](int i) { return stage[gi * n2 + i]; }, [] {}, [&](int i, cd v) { rowb[F::sw(i)] = v; });
            __syncthreads();
            if (threadIdx.x == 0 && b + gridDim.x < tiles) bulk_load(stage, scratch + (b + gridDim.x) * C * n2, TILE_BYTES, &bar);
            r16_pass<1, true>(lt, w16, [&](int i) { return rowb[F::sw(i)]; }, [] { __syncwarp(); }, [&](int i, cd v) { rowb[F::sw(i)] = v; });
            res = bufA;
        } else {
            constexpr int PASSES = RadixPlan<L2>::passes;
            // first pass: stage -> bufA; then the staging buffer is free for the next tile's copy
            stockham_pass_io<double, L2, F::T, 0>(
                lt, tw2, [&](int a, int cc) { return stage[gi * n2 + a + cc]; },
                [&](int a, int cc, cd v) { bufA[gi * SP + (Swz<double>::f(a) ^ Swz<double>::f(cc))] = v; });
            __syncthreads();
            if (threadIdx.x == 0 && b + gridDim.x < tiles) bulk_load(stage, scratch + (b + gridDim.x) * C * n2, TILE_BYTES, &bar);
            res = SpecPasses<L2, F::T, 1, PASSES>::run(lt, bufA + gi * SP, bufB + gi * SP, bufA + gi * SP, tw2) - gi * SP;
        }
        __syncthreads();
        // their contribution: X0 + Xn*(-1)^m, i.e. (d + q) on even and (d - q) on odd samples
        const double e_even = (d + q) * inv_sqrt, e_odd = (d - q) * inv_sqrt;
#pragma unroll
        for (int u = 0; u < NOUT; u++) {
            const int j2 = r0 + u * STEP;
            if (j2 >= n2 / 2) continue;
            const cd v = res[c * SP + F::sw(j2)];
            const long long m = 2 * ((long long)(j1_0 + c) + (long long)n1 * j2);
            const double y0 = fma(v.x, inv_sqrt, e_even), y1 = fma(v.y, inv_sqrt, e_odd);
            // the hack's window, `1.0 - abs(1.0 - i/(n/2.0))`, same expression
            const double w0 = 1.0 - fabs(1.0 - (double)m / half_n), w1 = 1.0 - fabs(1.0 - (double)(m + 1) / half_n);
            const double2 t = make_double2(acc[u].x + y0 * w0, acc[u].y + y1 * w1);
            if (pairs_aligned && pos0 + m + 1 < out_len) {
                *reinterpret_cast<double2 *>(row + m) = t;  // (pos0 and out_stride are even: 16-byte aligned)
            } else {
                if (pos0 + m < out_len) row[m] = t.x;
                if (pos0 + m + 1 < out_len) row[m + 1] = t.y;
            }
        }
        if (edge_out && j1_0 == 0 && threadIdx.x == 0) {
            // what the in-place frft leaves in bins 0 and n for the next frame: y[0] and y[n] (j = 0 and j = n/2)
            edge_out[2 * ch] = fma(res[F::sw(0)].x, inv_sqrt, e_even);
            edge_out[2 * ch + 1] = fma(res[F::sw(n2 / 2)].x, inv_sqrt, e_even);
        }
        __syncthreads();
    }
}

// ---- host side ----
// Static check of a spectrum expression: every opcode known, arguments in range, stack depth within TSB_EXPR_STACK,
// operand types as the opcode wants them, and exactly two doubles (re, im) left at the end.
const char *expr_check(const tsb_expr *x) {
    if (!x) return "TSB_GEN_EXPR needs args->expr";
    if (x->n_ops <= 0 || x->n_ops > TSB_EXPR_MAX_OPS || !x->ops) return "expression: n_ops out of range";
    if (x->n_iconst < 0 || x->n_iconst > TSB_EXPR_MAX_CONST || x->n_dconst < 0 || x->n_dconst > TSB_EXPR_MAX_CONST)
        return "expression: too many constants";
    if (x->n_itab < 0 || x->n_itab > TSB_EXPR_MAX_TABS || x->n_dtab < 0 || x->n_dtab > TSB_EXPR_MAX_TABS)
        return "expression: too many per-bin tables";
    if (x->n_iparam < 0 || x->n_iparam > TSB_EXPR_MAX_TABS || x->n_dparam < 0 || x->n_dparam > TSB_EXPR_MAX_TABS)
        return "expression: too many per-frame parameters";
    if ((x->n_iconst && !x->iconst) || (x->n_dconst && !x->dconst) || (x->n_itab && !x->itab) || (x->n_dtab && !x->dtab) ||
        (x->n_iparam && !x->iparam) || (x->n_dparam && !x->dparam))
        return "expression: a declared constant / table / parameter array is NULL";
    char ty[TSB_EXPR_STACK];  // 'i' or 'd'
    int sp = 0;
    auto need = [&](int k, const char *types) {
        if (sp < k) return false;
        for (int j = 0; j < k; j++)
            if (ty[sp - k + j] != types[j]) return false;
        return true;
    };
    for (int pc = 0; pc < x->n_ops; pc++) {
        const int w = x->ops[pc], op = w & 255, arg = w >> 8;
        if (w < 0) return "expression: negative program word";
        switch (op) {
            case TSB_X_I_BIN: case TSB_X_I_CH: case TSB_X_I_FRAME: case TSB_X_I_N: case TSB_X_I_CONST: case TSB_X_I_TAB:
            case TSB_X_I_PARAM: case TSB_X_D_CONST: case TSB_X_D_TAB: case TSB_X_D_PARAM: {
                const int lim = op == TSB_X_I_CONST ? x->n_iconst : op == TSB_X_I_TAB ? x->n_itab : op == TSB_X_I_PARAM ? x->n_iparam
                              : op == TSB_X_D_CONST ? x->n_dconst : op == TSB_X_D_TAB ? x->n_dtab : op == TSB_X_D_PARAM ? x->n_dparam : 1;
                if (arg >= lim) return "expression: constant / table / parameter index out of range";
                if (sp >= TSB_EXPR_STACK) return "expression: stack overflow";
                ty[sp++] = (op >= TSB_X_D_CONST) ? 'd' : 'i';
            } break;
            case TSB_X_I_ADD: case TSB_X_I_SUB: case TSB_X_I_MUL: case TSB_X_I_DIV: case TSB_X_I_MOD: case TSB_X_I_AND:
            case TSB_X_I_OR: case TSB_X_I_XOR: case TSB_X_I_SHL: case TSB_X_I_SHR: case TSB_X_I_EQ: case TSB_X_I_LT:
                if (!need(2, "ii")) return "expression: integer operator needs two integers";
                sp--;
                break;
            case TSB_X_I_NEG: case TSB_X_I_NOT: case TSB_X_I_WRAP32:
                if (!need(1, "i")) return "expression: integer operator needs an integer";
                break;
            case TSB_X_I2D:
                if (!need(1, "i")) return "expression: I2D needs an integer";
                ty[sp - 1] = 'd';
                break;
            case TSB_X_D2I:
                if (!need(1, "d")) return "expression: D2I needs a double";
                ty[sp - 1] = 'i';
                break;
            case TSB_X_D_ADD: case TSB_X_D_SUB: case TSB_X_D_MUL: case TSB_X_D_DIV: case TSB_X_D_FMOD: case TSB_X_D_POW:
                if (!need(2, "dd")) return "expression: floating-point operator needs two doubles";
                sp--;
                break;
            case TSB_X_D_NEG: case TSB_X_D_ABS: case TSB_X_D_SIN: case TSB_X_D_COS: case TSB_X_D_EXP: case TSB_X_D_SQRT:
            case TSB_X_D_FLOOR:
                if (!need(1, "d")) return "expression: floating-point function needs a double";
                break;
            case TSB_X_PICK:
                if (arg >= sp) return "expression: PICK below the stack";
                if (sp >= TSB_EXPR_STACK) return "expression: stack overflow";
                ty[sp] = ty[sp - 1 - arg];
                sp++;
                break;
            case TSB_X_POLAR:
                if (!need(2, "dd")) return "expression: POLAR needs (mag, phase)";
                break;
            default: return "expression: unknown opcode";
        }
    }
    if (sp != 2 || ty[0] != 'd' || ty[1] != 'd') return "expression: the program must leave exactly two doubles (re, im)";
    return nullptr;
}

struct SpecCache {
    std::map<int, cd *> pass_tw;
    std::map<int, std::pair<cd *, cd *>> unit_tabs;  // log2 n -> (wa[n1], wb[n2]) of kernel A
    cd *scratch = nullptr;  // two frames' worth when the frames are pipelined (scratch_slots == 2)
    size_t scratch_elems = 0;
    int scratch_slots = 0;
    // frame pipeline: kernel A of frame f+1 runs on `aux` beside kernel B of frame f (see run_frames)
    cudaStream_t aux = nullptr;
    cudaStream_t bhi = nullptr;  // TSB_SPEC_PRIO=2: kernel B on a high-priority stream of its own (joined to the caller's at the end)
    cudaEvent_t ev_end = nullptr;
    cudaEvent_t ev_start = nullptr, ev_a[4] = {nullptr, nullptr, nullptr, nullptr}, ev_b[4] = {nullptr, nullptr, nullptr, nullptr};
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
    pc.scratch_slots = 0;
    for (double *&p : pc.edge) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    pc.edge_ch = 0;
    return cudaSuccess;
}

// wa[k1] = e^{-i pi k1 / n1}, wb[k2] = e^{-i pi k2 / n}, from sincospi-quality host values
static cudaError_t spec_unit_tables(SpecCache &pc, int l1, int l2, const cd **wa, const cd **wb) {
    auto it = pc.unit_tabs.find(l1 + l2);
    if (it == pc.unit_tabs.end()) {
        const long long n1 = 1LL << l1, n2 = 1LL << l2, n = n1 * n2;
        std::vector<double> a(2 * n1), b(2 * n2);
        const long double pi = 3.14159265358979323846264338327950288L;
        for (long long k = 0; k < n1; k++) {
            a[2 * k] = (double)cosl(-pi * (long double)k / (long double)n1);
            a[2 * k + 1] = (double)sinl(-pi * (long double)k / (long double)n1);
        }
        for (long long k = 0; k < n2; k++) {
            b[2 * k] = (double)cosl(-pi * (long double)k / (long double)n);
            b[2 * k + 1] = (double)sinl(-pi * (long double)k / (long double)n);
        }
        cd *da = nullptr, *db = nullptr;
        cudaError_t e = cudaMalloc(&da, a.size() * 8);
        if (e == cudaSuccess) e = cudaMalloc(&db, b.size() * 8);
        if (e == cudaSuccess) e = cudaMemcpy(da, a.data(), a.size() * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(db, b.data(), b.size() * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        it = pc.unit_tabs.emplace(l1 + l2, std::make_pair(da, db)).first;
    }
    *wa = it->second.first;
    *wb = it->second.second;
    return cudaSuccess;
}

template <int L1, int L2>
static cudaError_t run_frames(SpecCache &pc, const SpectralJob &jb, cudaStream_t st, long long *launches) {
    constexpr long long n = 1LL << (L1 + L2);
    const cd *tw1, *tw2, *wa, *wb;
    cudaError_t e;
    if ((e = spec_twiddles(pc, L1, &tw1)) != cudaSuccess) return e;
    if ((e = spec_twiddles(pc, L2, &tw2)) != cudaSuccess) return e;
    if ((e = spec_unit_tables(pc, L1, L2, &wa, &wb)) != cudaSuccess) return e;
    // Frame pipeline (more than one frame; TSB_SPEC_OVERLAP = scratch slots, default 2, 1 = off): kernel A of frame f+1 depends on nothing kernel B of
    // frame f produces (the carried bins travel from B to B), so it runs on a second stream into the other half of
    // a two-frame scratch: A is bound by the fp64 pipe, B by the latency of its L2 round trips — side by side on an
    // SM they fill each other's gaps.  Events order A(f) -> B(f) and B(f) -> A(f+2) (same scratch half).
    static const int want_slots = [] { const char *v = getenv("TSB_SPEC_OVERLAP"); const int w = v ? atoi(v) : 2; return w < 2 ? 1 : (w > 4 ? 4 : w); }();
    const bool overlap = want_slots > 1 && jb.n_frames > 1;
    const size_t per_frame = (size_t)jb.channels * n;
    const int slots = overlap ? want_slots : 1;
    if (pc.scratch_elems < per_frame * slots) {
        if (pc.scratch) cudaFree(pc.scratch);
        pc.scratch = nullptr;
        pc.scratch_elems = 0;
        if ((e = cudaMalloc(&pc.scratch, per_frame * slots * sizeof(cd))) != cudaSuccess) return e;
        pc.scratch_elems = per_frame * slots;
    }
    if (overlap && !pc.aux) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const char *pv = getenv("TSB_SPEC_PRIO");
        // default 2: kernel B on a high-priority stream of its own.  With equal priorities the block scheduler's choice
        // between B(f) and A(f+1) differs from process to process: a frame takes 40 us in some, 49 us in others
        // (measured, profiles/r2_cfg5_priority.txt); B first is the fast order, every time.  (0: B on the caller's
        // stream, 1: A on a high-priority stream)
        const int pmode = pv ? atoi(pv) : 2;
        if ((e = cudaStreamCreateWithPriority(&pc.aux, cudaStreamNonBlocking, pmode == 1 ? prio_hi : prio_lo)) != cudaSuccess) return e;
        if (pmode == 2 && (e = cudaStreamCreateWithPriority(&pc.bhi, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return e;
        cudaEvent_t *evs[10] = {&pc.ev_start, &pc.ev_a[0], &pc.ev_a[1], &pc.ev_a[2], &pc.ev_a[3], &pc.ev_b[0], &pc.ev_b[1], &pc.ev_b[2], &pc.ev_b[3], &pc.ev_end};
        for (cudaEvent_t *ev : evs)
            if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
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
    using FA = SubFft<L1, L2, TSB_SPEC_CAP_A>;
    using FB = SubFft<L2, L1, TSB_SPEC_CAP_B>;
    constexpr int n1 = 1 << L1, n2 = 1 << L2;
    constexpr int CA = FA::C, CB = FB::C;
    const size_t smA = (size_t)FA::BUFS * FA::C * (n1 + 1) * sizeof(cd);
    const size_t smB = ((size_t)FB::C * n2 + (size_t)FB::BUFS * FB::C * (n2 + 1)) * sizeof(cd);  // staging tile + work tiles
    auto kA_hash = spectral_cols_kernel<L1, L2, GenHash>;
    auto kA_buf = spectral_cols_kernel<L1, L2, GenBuffer>;
    auto kA_expr = spectral_cols_kernel<L1, L2, GenExpr>;
    auto kB = spectral_rows_kernel<L1, L2>;
    if ((e = cudaFuncSetAttribute(kA_hash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kA_buf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kA_expr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA)) != cudaSuccess) return e;
    // TSB_GEN_EXPR: the program goes into the kernel's parameter block; tables and per-frame parameters into one
    // stream-ordered device block: [itab | ipar | dtab | dpar] (8-byte aligned pieces)
    GenExpr gx{};
    unsigned char *xblock = nullptr;
    const tsb_expr *x = reinterpret_cast<const tsb_expr *>(jb.expr);
    if (jb.generator == SPEC_GEN_EXPR) {
        gx.code.n_ops = x->n_ops;
        for (int i = 0; i < x->n_ops; i++) gx.code.ops[i] = x->ops[i];
        for (int i = 0; i < x->n_iconst; i++) gx.code.ic[i] = x->iconst[i];
        for (int i = 0; i < x->n_dconst; i++) gx.code.dc[i] = x->dconst[i];
        gx.n = (int)n;
        const size_t b_itab = (((size_t)x->n_itab * n * 4) + 7) & ~(size_t)7, b_ipar = (size_t)x->n_iparam * jb.n_frames * 8;
        const size_t b_dtab = (size_t)x->n_dtab * n * 8, b_dpar = (size_t)x->n_dparam * jb.n_frames * 8;
        if ((e = cudaMallocAsync(&xblock, b_itab + b_ipar + b_dtab + b_dpar + 8, st)) != cudaSuccess) return e;
        unsigned char *p = xblock;
        auto put = [&](const void *src, size_t raw, size_t padded) -> const void * {
            const void *at = p;
            if (raw) e = cudaMemcpyAsync(p, src, raw, cudaMemcpyHostToDevice, st);
            p += padded;
            return at;
        };
        gx.itab = (const int *)put(x->itab, (size_t)x->n_itab * n * 4, b_itab);
        if (e == cudaSuccess) gx.ipar = (const long long *)put(x->iparam, b_ipar, b_ipar);
        if (e == cudaSuccess) gx.dtab = (const double *)put(x->dtab, b_dtab, b_dtab);
        if (e == cudaSuccess) gx.dpar = (const double *)put(x->dparam, b_dpar, b_dpar);
        if (e != cudaSuccess) {
            cudaFreeAsync(xblock, st);
            return e;
        }
    }
    if ((e = cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB)) != cudaSuccess) return e;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const long long tilesB = (long long)jb.channels * (n1 / CB);
    int perB = 1;  // resident CTAs of kernel B per SM: it is launched persistent, one wave
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perB, kB, FB::NT, smB)) != cudaSuccess) return e;
    if (perB < 1) perB = 1;
    if (const char *v = getenv("TSB_SPEC_BPERSM")) {  // tuning: persistent CTAs of kernel B per SM
        const int w = atoi(v);
        if (w >= 1 && w < perB) perB = w;
    }
    const unsigned gridA = (unsigned)((long long)jb.channels * (n2 / CA));
    const unsigned gridB = (unsigned)(tilesB < (long long)sms * perB ? tilesB : (long long)sms * perB);
    if (overlap) {  // everything queued on `st` so far (tables, the caller's spectrum, earlier renders) precedes the first A
        if ((e = cudaEventRecord(pc.ev_start, st)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(pc.aux, pc.ev_start, 0)) != cudaSuccess) return e;
        if (pc.bhi && (e = cudaStreamWaitEvent(pc.bhi, pc.ev_start, 0)) != cudaSuccess) return e;
    }
    const cudaStream_t sa = overlap ? pc.aux : st;
    const cudaStream_t sb = overlap && pc.bhi ? pc.bhi : st;
    for (long long f = 0; f < jb.n_frames; f++) {
        const int slot = (int)(f % slots);
        cd *scr = pc.scratch + (size_t)slot * per_frame;
        if (overlap && f >= slots && (e = cudaStreamWaitEvent(sa, pc.ev_b[slot], 0)) != cudaSuccess) return e;
        if (jb.generator == SPEC_GEN_HASH) {
            GenHash g{(uint32_t)(jb.first_frame + f)};
            kA_hash<<<gridA, FA::NT, smA, sa>>>(g, scr, tw1, wa, wb, jb.channels);
        } else if (jb.generator == SPEC_GEN_EXPR) {
            GenExpr g = gx;
            g.frame = jb.first_frame + f;
            g.ipar = gx.ipar + (size_t)f * x->n_iparam;
            g.dpar = gx.dpar + (size_t)f * x->n_dparam;
            kA_expr<<<gridA, FA::NT, smA, sa>>>(g, scr, tw1, wa, wb, jb.channels);
        } else {
            GenBuffer g{reinterpret_cast<const cd *>(jb.spectrum) + (size_t)f * jb.channels * n, n};
            kA_buf<<<gridA, FA::NT, smA, sa>>>(g, scr, tw1, wa, wb, jb.channels);
        }
        if (overlap) {
            if ((e = cudaEventRecord(pc.ev_a[slot], sa)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(sb, pc.ev_a[slot], 0)) != cudaSuccess) return e;
        }
        const long long pos0 = jb.out_off0 + f * jb.hop;
        const int aligned = ((pos0 | jb.out_stride) & 1) == 0 && (reinterpret_cast<uintptr_t>(jb.out) & 15) == 0;
        kB<<<gridB, FB::NT, smB, sb>>>(scr, tw2, jb.channels, jb.out, jb.out_stride, jb.out_len, pos0, pc.edge[f & 1],
                                        pc.edge[(f + 1) & 1], aligned);
        if (overlap && (e = cudaEventRecord(pc.ev_b[slot], sb)) != cudaSuccess) return e;
        *launches += 2;
    }
    if (sb != st) {  // the caller's stream continues after the last kernel B
        if ((e = cudaEventRecord(pc.ev_end, sb)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(st, pc.ev_end, 0)) != cudaSuccess) return e;
    }
    if (xblock) cudaFreeAsync(xblock, st);
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
