// tsb_chain.cuh — the fused per-lane overlap-add spectral chain
//
//     gather (ring) -> analysis window -> R2C -> [phase vocoder | real spectral gain] -> C2R
//            -> synthesis window -> overlap-add -> emit hop samples
//
// for one lane (= one channel of one stream), frame after frame, with everything that the reference
// carries from frame to frame (PhaseVocoder::last_phase / phase_sum, rtinc/pv.hpp:14-15, and the
// scheduler's overlap tail `atmp`, tinyspec.cpp:43,194-200) held on chip.  It replaces, for a batch
// of lanes at once, the reference's generate_frames loop (tinyspec.cpp:122-201) together with the
// helper chain a hack runs inside it (hacks/pitchshift-pv.cpp:4-9; rtlib/pv.hpp:4-79,
// rtlib/fft.hpp:3-42, rtlib/windowfn.hpp:1-14).
//
// The code is organised as per-thread PHASES with no barrier inside a phase; the CUDA kernel
// (tsb_chain_inst.cu) calls them with __syncthreads() in between, and the CPU test harness
// (tests/emul) replays the same phases thread by thread.  Per frame (P = Stockham passes of the M-point FFT):
//     forward pass 0      ring -> analysis window -> butterflies -> buf        (gather + window fused into the load)
//     forward pass 1..P-1 buf <-> buf; pass 1 also issues the cp.async fetch of the next frame's new samples
//     PV analyse          real-FFT split + |X|, arg X, wrapped phase advance -> exchange arrays      [PV program]
//     PV synthesise       shift gather, phase accumulation, polar, C2R pre-processing -> buf          [PV program]
//     gain                split, multiply, C2R pre-processing -> buf                                  [gain program]
//     inverse pass 0..P-2 buf <-> buf
//     inverse pass P-1    butterflies -> synthesis window -> overlap-add ring -> `hop` samples to HBM (fused store)
// Shared-memory layout per lane:
//     ring[N]    float   input samples, slot = absolute sample index mod N
//     ola[N]     real    overlap-add accumulator, slot = absolute output position mod N
//     buf[2][M]  complex<real>, ping-pong FFT buffers (M = N/2), XOR-swizzled (Swz<real>); the idle one doubles
//                as the PV stage's phase-increment (and, in fp64, amplitude) exchange array
//     amp_own[M] real    PV amplitude exchange array (fp32 PV, not pipelined; pipelined it lives in ibuf[1])
//
// Real-FFT mapping: the reference runs complex FFTs of real data.  Here the N
// real samples are packed into M = N/2 complex points, transformed once, and split; with
// X = R2C(x) (kernel e^-), frft(x,+1)[k] = conj(X[k])/sqrt(N), and Re(frft(Y,-1))[n] =
// C2R(Z)[n]/sqrt(N) with Z[k] = (conj(Y[k]) + Y[N-k])/2.  1/sqrt(N) is folded into the window tables, the 1/2
// of the PV case (Y[k>=N/2] = 0 after shift) into the amplitudes handed to the synthesis.
#pragma once
#include <cmath>
#include <cstring>
#include <type_traits>

#include "tsb_math.cuh"

namespace tsb {

enum Program { PROG_PV_SHIFT = 0, PROG_GAIN = 1 };

// ---- IEEE single-rounding helpers: never contracted into FMAs, so that the fp64 phase-vocoder
// arithmetic reproduces the reference's expression-by-expression rounding (x86-64, no FMA).
#if defined(__CUDA_ARCH__)
TSB_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
TSB_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
TSB_D double sub_rn(double a, double b) { return __dsub_rn(a, b); }
TSB_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
TSB_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
TSB_D int d2i_rz(double a) { return __double2int_rz(a); }
TSB_D long long d2ll_rn(double a) { return __double2ll_rn(a); }
#else
inline double mul_rn(double a, double b) { return a * b; }  // host harness is built with -ffp-contract=off
inline double add_rn(double a, double b) { return a + b; }
inline double sub_rn(double a, double b) { return a - b; }
inline float mul_rn(float a, float b) { return a * b; }
inline float add_rn(float a, float b) { return a + b; }
inline int d2i_rz(double a) { return (int)a; }
inline long long d2ll_rn(double a) { return llrint(a); }
#endif

#define TSB_PI 3.14159265358979323846 /* M_PI */

// Compiler-level scheduling fence: keeps the unrolled per-slot bodies of the PV phases from being interleaved
// into one long stretch whose live values exceed the register budget (spills cost more than the lost overlap).
TSB_HD void sched_fence() {
#if defined(__CUDA_ARCH__)
    asm volatile("" ::: "memory");
#endif
}

// Pins a value's computation to this point of the instruction stream (keeps the optimiser from sinking it into a
// conditional block).
TSB_HD void keep_here(double &v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+d"(v));
#else
    (void)v;
#endif
}
TSB_HD void keep_here(float &v) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+f"(v));
#else
    (void)v;
#endif
}

// Tables built once per plan on the host (tsb_host.cpp) and kept in device memory.
template <typename real>
struct ChainTables {
    const cpx<real> *tw;     // per-pass twiddle tables of the M-point FFT (TwLayout<LOG2M>)
    const cpx<real> *twn;    // [M/2+1]  exp(-2*pi*i*k/N)
    const real *win_a;       // [N]      analysis window / sqrt(N)
    const real *win_s;       // [N]      synthesis window / sqrt(N) (the same table when the kinds are equal)
    // PROG_PV_SHIFT
    double inv_n;            // 1/N (exact)
    const uint32_t *map;     // [M]      (first << 16) | count into map_src: sources of destination bin k
    const uint16_t *map_src; // source bins, ascending within each destination (rtlib/pv.hpp:70-78)
    const int16_t *map1;     // [M] or nullptr: when no destination has more than one source, that source (-1: none)
    const uint8_t *src_used; // [M]      1 if some destination takes source bin j (the shift drops bins that land >= N/2)
    const uint16_t *dst_order; // [M/2] or nullptr: destination pairs (k, M-k) in the order the threads take them —
                             // pairs with a source first, so that whole slots of silent pairs are skipped
    double hop;              // PhaseVocoder::hop() (rtinc/pv.hpp:50-53), may be fractional
    double rate;             // RATE
    double ratio;            // fn(x) = ratio*x applied to the frequencies (rtlib/pv.hpp:66)
    double size_over_hop;    // size/hop()
    double inv_rate;         // 1/RATE
    double dc_pos, dc_neg;   // arg() the reference obtains for a positive / negative DC bin (see dc_phase())
    // PROG_GAIN
    const real *gain;        // [M+1]    real gain applied to bins k and (N-k)%N
    // fp32 PV only: the analysis window in double (unscaled), for the exact re-evaluation of doubtful wraps
    const double *win64;     // [N] or nullptr
};

// Launch arguments (one lane per CTA).
struct ChainArgs {
    const float *in;       // [lanes][in_stride] float32 input samples (planar per lane)
    long long in_stride;
    long long in_len;      // valid samples per lane (reads beyond are zero)
    void *out;             // [lanes][out_stride] real
    long long out_stride;
    long long out_len;     // writes beyond are dropped
    long long n_frames;
    long long in_off0;     // I_0: input offset of the first frame of this launch
    long long out_off0;    // P_0: output position of the first frame of this launch
    int hop;               // computed_hop when constant; ignored if hop_tab != nullptr
    const int *hop_tab;    // [n_frames] per-frame computed_hop (fractional hops, tinyspec.cpp:163-168)
    double *st_last_phase; // [lanes][M]   state in/out (nullptr: start from zero, do not save)
    double *st_phase_sum;  // [lanes][M]
    double *st_ola;        // [lanes][N]   overlap tail relative to the next output position
    int load_state;        // 0: start from zero (PhaseVocoder::reset, empty atmp)
    int lanes;
    // ---- time segments (tsb_timepar_*): CTA b works on segment b / lanes of lane b % lanes; state rows are
    // [segment * lanes + lane].  seg_tab[s] = {first frame (index into hop_tab), frames, I of the first frame,
    // P of the first frame, warm-up}: with warm-up the CTA first analyses the frame BEFORE the segment
    // (input offset I - prev_hop) only to obtain last_phase.
    int n_seg;                 // 0 or 1: the whole launch is one segment described by the fields above
    const long long *seg_tab;  // [n_seg][SEG_COLS]: first_frame, n_frames, in_off0, out_off0, warm-up flag, hop of the frame
                               // before, lead-in frames (see SegInfo)
    int mode;                  // 0: full chain; 1: pre-pass (analysis + shift + increment accumulation only, no output)
    double *st_warm_phase;     // pre-pass out / main in: [rows][M] last_phase after the warm-up frame (segments with warm-up)
    double *st_inc;            // pre-pass out (or nullptr): [lanes][inc_frames][M] the increment every frame adds to the
    long long inc_frames;      //   phase_sum of every destination bin, frames numbered as in hop_tab / seg_tab
    int prev_hop;              // load_state launches: hop of the frame before the first one (its samples sit at
                               // in_off0 - prev_hop of the same rows), 0 = unknown.  Only the fp32 PV reads it (WrapFix).
};

// What one CTA works on.  A segment that does not start the render begins with the frames whose output overlaps
// its first own frame (`lead` of them, N/hop - 1 for a constant hop): the pre-pass only analyses them (their
// increments belong to the previous segment), the main pass renders them without emitting anything — that
// rebuilds, with the sequential render's own arithmetic, the overlap tail the segment starts with.
struct SegInfo {
    long long first_frame, n_frames, in_off0, out_off0;  // first processed frame (a lead-in frame if lead > 0), count incl. lead
    int warm, warm_hop;
    int lead;
    long long row;  // state row = segment * lanes + lane
};

template <typename real>
struct PvMath;

#if !defined(__CUDA_ARCH__)
// host replay only (tests/emul): [0] doubtful wraps re-evaluated, [1] of them decided the other way than float32
inline long long *host_wrap_counters() {
    static long long c[2] = {0, 0};
    return c;
}
#endif

// single-instruction (MUFU) reciprocal / reciprocal square root, ~1 ulp, flush-to-zero
TSB_HD float fast_rcpf(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
TSB_HD float fast_rsqrtf(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}

// Branch-free atan2f: octant reduction + odd minimax polynomial on [0,1]; max error 3.1e-7 rad (the same
// as a correctly rounded float result near +-pi), no special-case paths.  atan2f(+-0, x<0) = +-pi as in libm.
TSB_HD float fast_atan2f(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = mn * fast_rcpf(fmaxf(mx, 1e-37f));
    float s = a * a;
    float r = 0.00282363896258175373077393f;
    r = fmaf(r, s, -0.0159569028764963150024414f);
    r = fmaf(r, s, 0.0425049886107444763183594f);
    r = fmaf(r, s, -0.0748900920152664184570312f);
    r = fmaf(r, s, 0.106347933411598205566406f);
    r = fmaf(r, s, -0.142027363181114196777344f);
    r = fmaf(r, s, 0.199926957488059997558594f);
    r = fmaf(r, s, -0.333331018686294555664062f);
    r = fmaf(r * s, a, a);
    r = ay > ax ? 1.57079632679489661923f - r : r;
    r = x < 0.f ? 3.14159265358979323846f - r : r;
    return copysignf(r, y);
}

// ---- branch-free fp64 kernels for the PV analysis (results within ~1 ulp of libm's; no special-case paths, so the
// compiler can interleave the two bins a thread analyses back to back) ----
// hardware seed reciprocals (~20 significant bits, MUFU.RCP64H / RSQ64H)
#if !defined(__CUDA_ARCH__)
inline double chop_seed(double v) {  // host harness: keep 20 mantissa bits like the hardware seed
    uint64_t u;
    memcpy(&u, &v, 8);
    u &= 0xFFFFFFFF00000000ull;
    memcpy(&v, &u, 8);
    return v;
}
#endif
TSB_HD double rcp_seed(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return chop_seed(1.0 / x);
#endif
}
TSB_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return chop_seed(1.0 / sqrt(x));
#endif
}
// n/d for d > 0 (normal range): seed, two Newton steps on the reciprocal, one correction of the quotient
TSB_HD double div_pos(double n, double d) {
    double r = rcp_seed(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double q = n * r;
    return fma(fma(-d, q, n), r, q);
}
// sqrt(x), x >= 0: coupled Newton (Goldschmidt) iteration from the seed, then one correction; exact 0 for x = 0
TSB_HD double sqrt_pos(double x) {
    double xs = x + 1e-290;  // = x for every x above 1e-274 (the addend is below half an ulp); keeps the seed off x = 0.
                             // (fmax on doubles is a compare, two selects and NaN quieting: ~8 instructions)
    double y = rsqrt_seed(xs);
    double g = xs * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, xs), h, g);
    return x > 0.0 ? g : 0.0;
}
// atan(i/64), i = 0..65, correctly rounded (128 slots: the index is masked, so that a NaN input cannot read outside)
#if defined(__CUDA_ARCH__)
static __device__ const double k_atan_tab[128] = {
#else
static const double k_atan_tab[128] = {
#endif
    0.0, 0.015623728620476831, 0.031239833430268277,
    0.046840712915969654, 0.06241880999595735, 0.0779666338315423,
    0.09347678115858947, 0.10894195698986579, 0.12435499454676144,
    0.13970887428916365, 0.15499674192394097, 0.1702119252854744,
    0.18534794999569476, 0.2003985538258785, 0.21535769969773805,
    0.23021958727684372, 0.24497866312686414, 0.2596296294082575,
    0.2741674511196588, 0.2885873618940774, 0.3028848683749714,
    0.31705575320914703, 0.3310960767041321, 0.34500217720710513,
    0.35877067027057225, 0.3723984466767542, 0.38588266939807375,
    0.39922076957525254, 0.4124104415973873, 0.42544963737004227,
    0.43833655985795783, 0.4510696559885235, 0.4636476090008061,
    0.4760693303227612, 0.48833395105640554, 0.5004408131472942,
    0.5123894603107377, 0.5241796287829132, 0.5358112379604637,
    0.5472843809874369, 0.5585993153435624, 0.5697564534829784,
    0.5807563535676704, 0.5915997103351114, 0.6022873461349642,
    0.6128202021652414, 0.6231993299340659, 0.6334258829691446,
    0.6435011087932844, 0.6534263411807619, 0.6632029927060933,
    0.6728325475937632, 0.6823165548747481, 0.6916566218531999,
    0.7008544078844502, 0.7099116184635249, 0.7188299996216245,
    0.7276113326265107, 0.7362574289814281, 0.7447701257160751,
    0.7531512809621944, 0.7614027698055784, 0.7695264804056583,
    0.7775243103733478, 0.7853981633974483, 0.7931499461096552};
// atan2(y, x): octant reduction to a/b in [0,1]; a/b = c + (a - c*b)/(b + c*a) with c = i/64 the nearest table
// point, so that atan(a/b) = atan(c) + atan(t), |t| <= 1/128 + 2^-20, where four terms of the series give
// 2^-59 relative.  One full-precision division.  Sign conventions of libm (atan2(+-0, x<0) = +-pi).
TSB_HD double fast_atan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool swap = ay > ax;                                // one compare serves max, min and the octant fix-up
    const double b = swap ? ay : ax, a = swap ? ax : ay;
    const double bs = b + 1e-290;                             // = b unless b < 1e-274 (spectra of float32 frames are 0 or
                                                              // far above that); b = 0 gives a/bs = 0
    const double q0 = a * rcp_seed(bs);                       // ~20 bits of a/b
    const double m = q0 * 64.0 + 6755399441055744.0;          // round to nearest integer in the low word
#if defined(__CUDA_ARCH__)
    const int i = __double2loint(m) & 127;
#else
    uint64_t mbits;
    memcpy(&mbits, &m, 8);
    const int i = (int)(mbits & 127);
#endif
    const double c = (m - 6755399441055744.0) * 0.015625;
    const double t = div_pos(fma(-c, bs, a), fma(c, a, bs));
    const double t2 = t * t;
    double p = fma(t2, -1.0 / 7.0, 0.2);
    p = fma(t2, p, -1.0 / 3.0);
    double v = TSB_LDG(&k_atan_tab[i]) + fma(t * t2, p, t);
    v = swap ? (1.5707963267948966 - v) + 6.123233995736766e-17 : v;
    v = x < 0.0 ? (3.141592653589793 - v) + 1.2246467991473532e-16 : v;
#if defined(__CUDA_ARCH__)
    v = (b == 0.0 && __double2hiint(x) < 0) ? 3.141592653589793 : v;  // atan2(+-0, -0) = +-pi
#else
    v = (b == 0.0 && std::signbit(x)) ? 3.141592653589793 : v;
#endif
    return copysign(v, y);
}
// trunc(fl(delta / M_PI)): the correctly rounded quotient from a reciprocal multiply and two exact-remainder
// corrections (Markstein), then the reference's (int) truncation (rtlib/pv.hpp:41).  delta = -M_PI (a DC sign
// change) gives exactly -1.
TSB_HD int wrap_quotient(double delta) {
    const double inv_pi = 0.3183098861837907;
    double q = delta * inv_pi;
    q = fma(fma(-q, TSB_PI, delta), inv_pi, q);
    q = fma(fma(-q, TSB_PI, delta), inv_pi, q);
    return d2i_rz(q);
}

// sin and cos of an unreduced phase (|x| up to ~1e14), branch-free: x*(32/pi) = k + f with the product carried in
// two doubles (exact head via fma, so the reduction is as good at 1e9 rad as at 1), sin/cos(k*pi/32) from a
// 64-entry table (the index wraps around the circle: no quadrant logic), and short series in r = f*pi/32,
// |r| <= pi/64.  Absolute error ~1e-16 (what polar() needs), independent of the argument's size.
struct SinCosEntry {
    double s, c;
};
#if defined(__CUDA_ARCH__)
static __device__ const SinCosEntry k_sincos_tab[64] = {
#else
static const SinCosEntry k_sincos_tab[64] = {
#endif
    {0.0, 1.0},
    {0.0980171403295606, 0.9951847266721969},
    {0.19509032201612828, 0.9807852804032304},
    {0.2902846772544624, 0.9569403357322088},
    {0.3826834323650898, 0.9238795325112867},
    {0.47139673682599764, 0.881921264348355},
    {0.5555702330196022, 0.8314696123025452},
    {0.6343932841636455, 0.773010453362737},
    {0.7071067811865476, 0.7071067811865476},
    {0.773010453362737, 0.6343932841636455},
    {0.8314696123025452, 0.5555702330196022},
    {0.881921264348355, 0.47139673682599764},
    {0.9238795325112867, 0.3826834323650898},
    {0.9569403357322088, 0.2902846772544624},
    {0.9807852804032304, 0.19509032201612828},
    {0.9951847266721969, 0.0980171403295606},
    {1.0, 2.4662128696421e-91},
    {0.9951847266721969, -0.0980171403295606},
    {0.9807852804032304, -0.19509032201612828},
    {0.9569403357322088, -0.2902846772544624},
    {0.9238795325112867, -0.3826834323650898},
    {0.881921264348355, -0.47139673682599764},
    {0.8314696123025452, -0.5555702330196022},
    {0.773010453362737, -0.6343932841636455},
    {0.7071067811865476, -0.7071067811865476},
    {0.6343932841636455, -0.773010453362737},
    {0.5555702330196022, -0.8314696123025452},
    {0.47139673682599764, -0.881921264348355},
    {0.3826834323650898, -0.9238795325112867},
    {0.2902846772544624, -0.9569403357322088},
    {0.19509032201612828, -0.9807852804032304},
    {0.0980171403295606, -0.9951847266721969},
    {4.9324257392842e-91, -1.0},
    {-0.0980171403295606, -0.9951847266721969},
    {-0.19509032201612828, -0.9807852804032304},
    {-0.2902846772544624, -0.9569403357322088},
    {-0.3826834323650898, -0.9238795325112867},
    {-0.47139673682599764, -0.881921264348355},
    {-0.5555702330196022, -0.8314696123025452},
    {-0.6343932841636455, -0.773010453362737},
    {-0.7071067811865476, -0.7071067811865476},
    {-0.773010453362737, -0.6343932841636455},
    {-0.8314696123025452, -0.5555702330196022},
    {-0.881921264348355, -0.47139673682599764},
    {-0.9238795325112867, -0.3826834323650898},
    {-0.9569403357322088, -0.2902846772544624},
    {-0.9807852804032304, -0.19509032201612828},
    {-0.9951847266721969, -0.0980171403295606},
    {-1.0, -1.7216825539521752e-90},
    {-0.9951847266721969, 0.0980171403295606},
    {-0.9807852804032304, 0.19509032201612828},
    {-0.9569403357322088, 0.2902846772544624},
    {-0.9238795325112867, 0.3826834323650898},
    {-0.881921264348355, 0.47139673682599764},
    {-0.8314696123025452, 0.5555702330196022},
    {-0.773010453362737, 0.6343932841636455},
    {-0.7071067811865476, 0.7071067811865476},
    {-0.6343932841636455, 0.773010453362737},
    {-0.5555702330196022, 0.8314696123025452},
    {-0.47139673682599764, 0.881921264348355},
    {-0.3826834323650898, 0.9238795325112867},
    {-0.2902846772544624, 0.9569403357322088},
    {-0.19509032201612828, 0.9807852804032304},
    {-0.0980171403295606, 0.9951847266721969}};
TSB_HD void fast_sincos(double x, double &sn, double &cs) {
    const double c1 = 10.185916357881302, c2 = -6.297176536058395e-16;  // 32/pi = c1 + c2
    const double p = x * c1;
    const double e = fma(x, c1, -p);                          // exact tail of the product
    const double m = p + 6755399441055744.0;                  // k = rint(p) in the low word
    const double kd = m - 6755399441055744.0;
#if defined(__CUDA_ARCH__)
    const int ki = __double2loint(m) & 63;
    const double2 sc = __ldg(reinterpret_cast<const double2 *>(k_sincos_tab) + ki);
    const double S = sc.x, C = sc.y;
#else
    uint64_t mb;
    memcpy(&mb, &m, 8);
    const int ki = (int)(mb & 63);
    const double S = k_sincos_tab[ki].s, C = k_sincos_tab[ki].c;
#endif
    const double f = (p - kd) + fma(x, c2, e);               // |f| <= 1/2 (+ rounding)
    const double r = f * 0.09817477042468103;                                 // pi/32
    const double r2 = r * r;
    double ps = fma(r2, -1.0 / 5040.0, 1.0 / 120.0);
    ps = fma(r2, ps, -1.0 / 6.0);
    const double sr = fma(r * r2, ps, r);                      // sin r
    double pc = fma(r2, 1.0 / 40320.0, -1.0 / 720.0);
    pc = fma(r2, pc, 1.0 / 24.0);
    pc = fma(r2, pc, -0.5);
    const double cm = r2 * pc;                                 // cos r - 1
    sn = fma(C, sr, fma(S, cm, S));
    cs = fma(-S, sr, fma(C, cm, C));
}

// Expected phase advance of bin j, `bin * 2.0 * M_PI * hop / size` evaluated left to right as rtlib/pv.hpp:40 does
// (size is a power of two: the division is an exact scaling).  Computed, not tabulated: a table of N/2 doubles
// per plan is what pushed the kernel's constant data past the L1 that is left beside the shared memory.
template <typename real>
TSB_HD double expected_advance(int j, const ChainTables<real> &tb) {
    return mul_rn(mul_rn((double)(2 * j), TSB_PI), tb.hop) * tb.inv_n;
}

// The DC bin of a real frame is real, so its phase sits exactly on the branch cut of arg() and on the
// wrap boundary of phase_to_frequency whenever its sign changes.  In the reference the value fed to
// arg() carries a ~1e-16 imaginary leak of the same sign as the bin (pow(i,2) is inexact in
// rtlib/fft.hpp:22, and frft runs in place), so arg() yields +eps or -pi+eps rather than 0 / +-pi, and
// that decides the wrap side.  Both values depend only on N; the host evaluates them with the
// reference's own expression (tsb_plan.hpp: dc_phase_consts).
template <typename real>
TSB_HD double dc_phase(double re, const ChainTables<real> &tb) {
    return re > 0 ? tb.dc_pos : (re < 0 ? tb.dc_neg : 0.0);
}

// ---- fp64: expression-faithful restatement of rtlib/pv.hpp:17-48,65-66 ----
template <>
struct PvMath<double> {
    // analysis of one bin: spectrum value X (R2C convention) -> amplitude and the phase increment the
    // synthesis stage will add for the bin this one is shifted to.
    static TSB_HD void analyze(int j, cpx<double> X, double &last_phase, const ChainTables<double> &tb, int n,
                               double &amp, double &inc, bool dc) {
        double re = X.x, im = -X.y;  // frft(.,+1) of real data = conj(R2C)/sqrt(N)
        amp = sqrt_pos(add_rn(mul_rn(re, re), mul_rn(im, im)));  // abs(), pv.hpp:20
        double pa = fast_atan2(im, re);
        keep_here(pa);  // computed for every thread (no branch around it), bin 0 then takes its own value
        double phase = dc ? dc_phase(re, tb) : pa;  // arg(), pv.hpp:21 (dc: bin 0)
        double diff = sub_rn(phase, last_phase);
        last_phase = phase;                                  // pv.hpp:23
        // phase_to_frequency, pv.hpp:39-45
        double delta = sub_rn(diff, expected_advance(j, tb));
        int qpd = wrap_quotient(delta);  // (int)(delta / M_PI)
        qpd += qpd >= 0 ? (qpd & 1) : -(qpd & 1);
        double t = sub_rn(delta, mul_rn(TSB_PI, (double)qpd));
        t = mul_rn(tb.size_over_hop, t);
        t = mul_rn(t, 0.5);  // "/2.0" is exact either way
        // The truncation above needs the true quotient (a DC sign change puts delta exactly on -pi); the two
        // divisions below only scale, so a reciprocal multiply (<= 1 ulp apart, ~1e-16 relative on the
        // increment) replaces them.
        t = mul_rn(t, 1.0 / TSB_PI);
        t = add_rn((double)j, t);
        t = mul_rn(t, tb.rate);
        double freq = mul_rn(t, 1.0 / (double)n);  // exact: n is a power of two
        freq = mul_rn(freq, tb.ratio);  // shift: freqs[j] = fn(freq), pv.hpp:66
        // frequency_to_phase, pv.hpp:46-48: 2.0*M_PI*freq/RATE*hop()
        inc = mul_rn(mul_rn(mul_rn(2.0 * TSB_PI, freq), tb.inv_rate), tb.hop);
    }
    static TSB_HD double accumulate(double phase_sum, double inc) { return add_rn(phase_sum, inc); }
    template <bool FAST = true>
    static TSB_HD cpx<double> synth(double amp, double inc, double &phase_sum) {
        phase_sum = add_rn(phase_sum, inc);  // pv.hpp:34 — never wrapped
        double s, c;
        if (FAST) {
            fast_sincos(phase_sum, s, c);
        } else {
#if defined(__CUDA_ARCH__)
            sincos(phase_sum, &s, &c);
#else
            s = sin(phase_sum);
            c = cos(phase_sum);
#endif
        }
        return mk<double>(mul_rn(amp, c), mul_rn(amp, s));  // std::polar, pv.hpp:35
    }
};

// ---- fp32: fp32 spectra / amplitudes / transcendental kernels; every phase quantity that
// accumulates (difference, wrap, increment, running sum, range reduction) stays fp64
// (SURVEY.md §7 "Twiddles").  inc = ratio*(expect_j + wrapped_delta), which is what
// pv.hpp:39-48 + the linear shift reduce to algebraically.
#ifndef TSB_WRAPFIX_INLINE
#define TSB_WRAPFIX_INLINE __forceinline__
#endif
#ifndef TSB_WRAP_K
#define TSB_WRAP_K 2e-6f   // doubt threshold, in units of sqrt(frame energy)/|X| radians (calibration: DESIGN.md §4.1)
#endif
#ifndef TSB_WRAP_C0
#define TSB_WRAP_C0 2e-6f  // ... plus this many radians (float rounding of the two phases, atan2f error)
#endif
template <>
struct PvMath<float> {
    // `doubt` (out): the wrapped phase deviation lies so close to +-pi — the boundary where (int)(delta/M_PI) changes
    // the even multiple of pi it removes (rtlib/pv.hpp:41-43) — that the fp32 transform's phase error (~1e-7 *
    // sqrt(frame energy) / |X|, plus the float rounding of the two phases) could put it on the other side than
    // the reference's double arithmetic.  A wrong side adds ratio*2*pi to phase_sum for good: harmless for an
    // integer ratio, a permanent sign flip of the bin at x1.5.  Doubtful bins are re-evaluated exactly (wrap_fix).
    // thr = WRAP_K * sqrt(frame energy), see Chain::wrap_threshold.
    static constexpr float WRAP_K = TSB_WRAP_K, WRAP_C0 = TSB_WRAP_C0;
    static TSB_HD void analyze(int j, cpx<float> X, float &last_phase, const ChainTables<float> &tb, int n,
                               float &amp, double &inc, bool dc, float thr, bool &doubt) {
        (void)n;
        float re = X.x, im = -X.y;
        float pw = fmaf(re, re, im * im);
        amp = pw * fast_rsqrtf(fmaxf(pw, 1e-37f));
        float phase = fast_atan2f(im, re);
        keep_here(phase);
        double diff = (double)phase - (double)last_phase;
        {  // bin 0: keep the reference's (double) values, stored as a sign in the float state (selects, no branch)
            double lp = last_phase > 0.f ? tb.dc_pos : (last_phase < 0.f ? tb.dc_neg : 0.0);
            double ph = dc_phase((double)re, tb);
            diff = dc ? ph - lp : diff;
            phase = dc ? (float)ph : phase;
        }
        last_phase = phase;
        double expect = expected_advance(j, tb);
        double delta = diff - expect;
        int qpd = d2i_rz(delta * (1.0 / TSB_PI));
        if (qpd >= 0) qpd += qpd & 1;
        else qpd -= qpd & 1;
        double wrapped = fma(-TSB_PI, (double)qpd, delta);
        inc = tb.ratio * (expect + wrapped);
        // |wrapped| within 2^-6 of pi?  (an integer compare on the high word of the double; only then — one bin in
        // 200 — the float test: the threshold cannot exceed 2^-6 rad, which only caps it for bins 100 dB down)
#if defined(__CUDA_ARCH__)
        const int hi = __double2hiint(wrapped) & 0x7fffffff;
#else
        uint64_t wb;
        memcpy(&wb, &wrapped, 8);
        const int hi = (int)(wb >> 32) & 0x7fffffff;
#endif
        doubt = false;
        if (hi >= 0x400901FB && !dc) {  // high word of pi - 2^-6
            const float margin = 3.14159274f - fabsf((float)wrapped);
            doubt = margin * amp < fmaf(WRAP_C0, amp, thr);
        }
    }
    // The exact side of the wrap from double-precision phases of this frame and the one before (both recomputed
    // from the input samples by Chain::wrap_fix, expressions of PvMath<double>::analyze), applied to the float32
    // increment: the value stays the one the float phases give (so that the increments of a bin keep telescoping —
    // their sum follows phase_f - phase_0 instead of random-walking), only the multiple of 2*pi*ratio is the exact one.
    static TSB_HD double exact_inc(double expect, double ratio, double inc32, double phase, double last_phase) {
        double delta = sub_rn(sub_rn(phase, last_phase), expect);
        int qpd = wrap_quotient(delta);
        qpd += qpd >= 0 ? (qpd & 1) : -(qpd & 1);
        double wrapped = fma(-TSB_PI, (double)qpd, delta);
        const double inc64 = ratio * (expect + wrapped);
        const double turn = ratio * (2.0 * TSB_PI);
        if (turn == 0.0) return inc32;
        const double k = rint((inc32 - inc64) / turn);  // 0 when the float decision was right, else +-1
#if !defined(__CUDA_ARCH__)
        host_wrap_counters()[0]++;
        host_wrap_counters()[1] += k != 0.0;
#endif
        return k == 0.0 ? inc32 : inc32 - turn * k;
    }
    static TSB_HD double accumulate(double phase_sum, double inc) { return phase_sum + inc; }
    template <bool FAST = true>
    static TSB_HD cpx<float> synth(float amp, double inc, double &phase_sum) {
        phase_sum += inc;
        // fp64 range reduction to [-pi/4, pi/4] + quadrant, then fp32 polynomials
        long long q = d2ll_rn(phase_sum * (2.0 / TSB_PI));
        double r = fma(-(double)q, 1.57079632679489655800e+00, phase_sum);
        r = fma(-(double)q, 6.12323399573676603587e-17, r);
        float x = (float)r, x2 = x * x;
        // minimax polynomials on [-pi/4, pi/4] (Cephes sinf/cosf kernels)
        float sp = -1.9515295891e-4f;
        sp = fmaf(sp, x2, 8.3321608736e-3f);
        sp = fmaf(sp, x2, -1.6666654611e-1f);
        float s = fmaf(sp * x2, x, x);
        float cp = 2.443315711809948e-5f;
        cp = fmaf(cp, x2, -1.388731625493765e-3f);
        cp = fmaf(cp, x2, 4.166664568298827e-2f);
        float c = fmaf(cp * x2, x2, fmaf(-0.5f, x2, 1.0f));
        int qi = (int)q;
        float ss = (qi & 1) ? c : s;
        float cc = (qi & 1) ? s : c;
        // quadrant signs: sin negative for q in {2,3}, cos negative for q in {1,2}; folded into the amplitude
        float as = (qi & 2) ? -amp : amp;
        float ac = ((qi + 1) & 2) ? -amp : amp;
        return mk<float>(ac * cc, as * ss);
    }
};

// OUT32 (fp64 arithmetic only): the output row and the overlap tail are float32 and are combined exactly as the
// reference's scheduler does (tinyspec.cpp:173-200): emitted sample = float(double(tail) + v), tail += float(v).
#ifndef TSB_FAST_SYNTH_MAX_S
#define TSB_FAST_SYNTH_MAX_S 2  // pair slots per thread up to which the fp64 synthesis uses fast_sincos (tuning)
#endif
template <typename real, int LOG2N, int PROG, bool OUT32 = false>
struct Chain {
    static_assert(!OUT32 || sizeof(real) == 8, "OUT32 is the float32-output variant of the fp64 chain");
    static constexpr int N = 1 << LOG2N;
    static constexpr int LOG2M = LOG2N - 1;
    static constexpr int M = N / 2;
    // Threads per lane: T0 = N/16 runs one radix-8 butterfly per thread and pass.  The PV kernels use 2*T0
    // threads: the extra warps only take part in the per-bin (PV, window, overlap-add) phases, which then own
    // 4 bins per thread instead of 8 — half the register-resident state, so more warps fit, and the per-bin
    // transcendental work (the bulk of the fp64 kernel) gets twice the threads (measured at N=1024: fp64 +45 %
    // together with the coalesced twiddle tables, fp32 +5 %).  TSB_TMUL overrides, for tuning.
    static constexpr int T0 = (N / 16) < 32 ? 32 : (N / 16);
#ifdef TSB_TMUL
    static constexpr int TMUL = N >= 256 ? TSB_TMUL : 1;
#else
    static constexpr int TMUL = (N >= 256 && N <= 4096) ? 2 : 1;  // (fp64 at 4096: one 512-thread CTA per SM; with the in-place
                                                                  // inverse transform and the register tail 27.7 -> 29.2 M frame-channels/s)
#endif
    static constexpr int T = T0 * TMUL > 1024 ? 1024 : T0 * TMUL;
    // With twice the FFT's threads, the two halves of the CTA transform two frames at once (see run_lane).
    static constexpr bool PIPE = T == 2 * T0;
    static constexpr int HALF = M / 2;                       // split pairs (k, M-k), k < M/2
    static constexpr int S = (HALF + T - 1) / T;             // pair slots per thread
    static constexpr int PASSES = RadixPlan<LOG2M>::passes;
    static constexpr bool IS_PV = PROG == PROG_PV_SHIFT;
    // fp32 PV: doubtful principal-argument wraps are re-evaluated in double precision (PvMath<float>::analyze)
    static constexpr bool WRAP_FIX = IS_PV && sizeof(real) == 4;
    // fp64 synthesis: the table-based branch-free sincos lets the compiler overlap the bins of a thread, which
    // costs registers; with 8 or more bins of state per thread (N >= 4096) that tips the FFT passes into spilling
    // (measured slower than CUDA's sincos there)
    static constexpr bool FAST_SYNTH = S <= TSB_FAST_SYNTH_MAX_S;
    // radix-8 twiddles: 0 = all seven from the table, 1 = w, w^2, w^4 from the table and the rest by products,
    // 2 = only w from the table
#ifndef TSB_TW_DERIVE64
#define TSB_TW_DERIVE64 2
#endif
#ifndef TSB_TW_DERIVE32
#define TSB_TW_DERIVE32 0
#endif
    static constexpr int TW_DERIVE = sizeof(real) == 8 ? TSB_TW_DERIVE64 : TSB_TW_DERIVE32;
    // fp64 pipelined kernels: the inverse transform runs IN PLACE (decimation in time; the synthesis writes its
    // input in digit-reversed order), so it needs one buffer instead of two — at N=1024 that is what lets a sixth
    // lane fit into an SM's shared memory (36 KB per lane).  (fp32 is bound by registers, not shared memory, and
    // parks its amplitude exchange array in the second inverse buffer.)
#ifndef TSB_INV_INPLACE
#define TSB_INV_INPLACE 1
#endif
    static constexpr bool INV_INPLACE = TSB_INV_INPLACE && PIPE && sizeof(real) == 8;
    // Register-resident overlap tail (constant hop N/4; N = 1024 in the text): the last inverse pass gives thread j the
    // sample pairs m = j + T0*u, u = 0..7, of every frame, and a hop of N/4 moves an output position from pair m of
    // one frame to pair m - 2*T0 of the next — the same thread, slot u - 2.  So the six pairs a frame leaves
    // unfinished stay in that thread's registers instead of making a read-modify-write trip through the shared-memory
    // ring every frame (160 of the kernel's ~1,480 data-pipe wavefronts per frame-channel, the busiest unit).  Same
    // additions in the same order as the ring path: bit-identical.  Launches with any other hop use the ring.
#ifndef TSB_RTAIL
#define TSB_RTAIL 1
#endif
    // frame sizes (bit log2 N) whose kernels keep the tail in registers, per precision: every size costs its kernel
    // registers, i.e. resident lanes, so each is enabled only where it was measured faster (tsb_chain_inst.cu)
    // (measured, M frame-channels/s without -> with: fp64 N=512 PV 262 -> 310, N=2048 PV 64.4 -> 69.3, gain 81.2 -> 85.8,
    // N=4096 28.7 -> 29.2 once pipelined (TMUL); fp32 N=2048 gain 128.5 -> 134.6, but PV 100.7 -> 89.2 and N=512 PV 416 -> 364: the fp32 PV
    // kernels lose more to the lanes the extra registers cost than the tail saves)
#ifndef TSB_RTAIL_SIZES64
#define TSB_RTAIL_SIZES64 ((1u << 9) | (1u << 10) | (1u << 11) | (1u << 12))
#endif
#ifndef TSB_RTAIL_SIZES32
#define TSB_RTAIL_SIZES32 (1u << 10)
#endif
#ifndef TSB_RTAIL_SIZES32_GAIN
#define TSB_RTAIL_SIZES32_GAIN ((1u << 10) | (1u << 11))
#endif
    // the last inverse pass is run by INV_NT threads; thread j gets the pairs m = j + T0*u, u < RT_PP, and a hop of
    // N/4 shifts a position by RT_SH slots
    static constexpr int INV_NT = PIPE ? T0 : T;
    static constexpr int RT_PP = M / T0, RT_SH = (N / 8) / T0;
    static constexpr bool RT_OK = TSB_RTAIL && INV_NT == T0 && RT_SH >= 1 && RT_SH * T0 * 8 == N &&
                                  (((sizeof(real) == 8 ? TSB_RTAIL_SIZES64 : (PROG == PROG_GAIN ? TSB_RTAIL_SIZES32_GAIN : TSB_RTAIL_SIZES32)) >> LOG2N) & 1u);
    static constexpr int RT_N = RT_OK ? RT_PP - RT_SH : 1;
    using C = cpx<real>;
    using Tables = ChainTables<real>;
    using SW = Swz<real>;
    using ISW = SwzInPlace<real>;  // layout of the in-place inverse buffer
    using otype = typename std::conditional<OUT32, float, real>::type;  // element of the output row and of the tail
    using OC = cpx<otype>;

    struct Smem {
        float ring[N];
        otype ola[N];
        C buf[2][M];  // ping-pong FFT buffers, swizzled (Swz<real>)
        // software pipeline (PIPE): the inverse transform of frame f runs on its own buffers while the other
        // half of the threads runs the forward transform of frame f+1 in `buf`
        C ibuf[PIPE ? (INV_INPLACE ? 1 : 2) : 1][PIPE ? M : 1];
        // PV amplitude exchange array: fp32 needs its own (the spare FFT buffer is filled by the M fp64
        // increments); in fp64 the spare buffer (16*M bytes) holds both, so nothing extra is reserved.
        real amp_own[(PROG == PROG_PV_SHIFT && sizeof(real) == 4 && !PIPE) ? M : 1];
        // fp32 PV: energy of the windowed frame (sum of squares), one partial per warp of the forward pass 0
        float energy[WRAP_FIX ? T / 32 : 1];
    };
    // Per-thread state that lives in registers for the whole launch.
    struct Regs {
        real last_phase[2 * S];
        double phase_sum[2 * S];
        // simple shift maps (at most one source per destination): the sources of this thread's bins, two int16 per
        // word, read once per launch — otherwise every synthesis phase starts by waiting for them (L1 latency
        // in front of the dependent shared-memory reads)
        uint32_t map_src[S];
        // bit 2i / 2i+1: the shift uses source bin k / M-k of slot i; bit 16+2i / 16+2i+1: destination bin k / M-k
        // of slot i has a source.  A source nobody takes only needs its phase on the last frame of a launch
        // (last_phase is state), a destination without sources stays silent and keeps its phase_sum.
        uint32_t used;
        uint32_t dpair[(S + 1) / 2];  // destination pair of slot i (16 bits each): k of the synthesis side
        uint32_t doubts;              // WRAP_FIX: doubt bits of the frame just analysed (bit 2 * slot + side)
        OC tail[RT_N];                // RT_OK: the overlap tail of this thread's sample pairs (inverse half of the CTA)
    };

    // Forward pass p writes buf[p & 1] (pass 0 reads the ring), so the packed spectrum ends up in RB.
    static constexpr int RB = (PASSES - 1) & 1;
    static constexpr int SB = RB ^ 1;
    static TSB_HD C *fwd_result(Smem &sm) { return sm.buf[RB]; }
    static TSB_HD C *fwd_spare(Smem &sm) { return sm.buf[SB]; }
    // PV exchange arrays (between the analysis and the synthesis phase)
    static TSB_HD double *xch_inc(Smem &sm) { return reinterpret_cast<double *>(fwd_spare(sm)); }
    static TSB_HD real *xch_amp(Smem &sm) {
        // fp32 pipelined: the second inverse buffer is idle between the last and the first inverse pass
        return sizeof(real) == 4 ? (PIPE ? reinterpret_cast<real *>(sm.ibuf[PIPE ? 1 : 0]) : sm.amp_own)
                                 : reinterpret_cast<real *>(xch_inc(sm) + M);
    }
    // The inverse transform starts from H: PV writes it over the consumed spectrum (RB), the gain program
    // into the spare buffer (no barrier between its split and merge).  Inverse pass p reads IB(p), writes IB(p+1).
    static constexpr int INV0 = PIPE ? 0 : (IS_PV ? RB : SB);
    static constexpr int inv_buf(int p) { return (p & 1) ? (INV0 ^ 1) : INV0; }
    static TSB_HD C *inv_buffer(Smem &sm, int p) {
        if constexpr (INV_INPLACE) return sm.ibuf[0];
        else return PIPE ? sm.ibuf[inv_buf(p)] : sm.buf[inv_buf(p)];
    }
    // where bin k of the inverse transform's input H goes
    static TSB_HD int hpos(int k) { return INV_INPLACE ? ISW::f(DitOrder<LOG2M>::pos(k)) : SW::f(k); }

    // ------------------------------------------------------------------ state in / out
    // state rows: registers <-> [row][M] arrays, thread `tid` owning bins (k, M-k) of its slots.  The register
    // arrays are passed by reference and indexed with unrolled constants so that they stay in registers.
    // kof(i): first bin of the pair held in slot i, or -1
    template <typename V, class KOf>
    static TSB_HD void regs_from(const double *arr, long long row, V (&dst)[2 * S], KOf kof) {
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = kof(i);
            if (k >= 0) {
                int kb = k == 0 ? HALF : M - k;
                dst[2 * i] = (V)arr[row * M + k];
                dst[2 * i + 1] = (V)arr[row * M + kb];
            }
        }
    }
    template <typename V, class KOf>
    static TSB_HD void regs_to(double *arr, long long row, const V (&src)[2 * S], KOf kof) {
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = kof(i);
            if (k >= 0) {
                int kb = k == 0 ? HALF : M - k;
                arr[row * M + k] = (double)src[2 * i];
                arr[row * M + kb] = (double)src[2 * i + 1];
            }
        }
    }
    // source side (analysis, last_phase): thread tid, slot i holds the pair (k, M-k), k = tid + T*i
    static TSB_HD int src_k(int tid, int i) {
        int k = tid + T * i;
        return k < HALF ? k : -1;
    }
    // destination side (synthesis, phase_sum): the plan may reorder the pairs (ChainTables::dst_order)
    static TSB_HD int dst_k(const Regs &rg, int i) {
        int k = (int)((rg.dpair[i >> 1] >> (16 * (i & 1))) & 0xffffu);
        return k < HALF ? k : -1;
    }
    template <typename V>
    static TSB_HD void regs_from(int tid, const double *arr, long long row, V (&dst)[2 * S]) {
        regs_from(arr, row, dst, [&](int i) { return src_k(tid, i); });
    }
    template <typename V>
    static TSB_HD void regs_to(int tid, double *arr, long long row, const V (&src)[2 * S]) {
        regs_to(arr, row, src, [&](int i) { return src_k(tid, i); });
    }
    static TSB_HD void load_state(int tid, Smem &sm, Regs &rg, const Tables &tb, const ChainArgs &a, const SegInfo &sg) {
        for (int i = 0; i < 2 * S; i++) {
            rg.last_phase[i] = 0;
            rg.phase_sum[i] = 0;
        }
        for (int i = 0; i < (S + 1) / 2; i++) rg.dpair[i] = 0xffffffffu;
        rg.used = 0;
#pragma unroll
        for (int i = 0; i < S; i++) {
            rg.map_src[i] = 0xffffffffu;
            const int q = tid + T * i;
            if (PROG != PROG_PV_SHIFT || q >= HALF) continue;
            // destination pair of this slot
            const int k = (HALF % T == 0 && tb.dst_order) ? (int)TSB_LDG(&tb.dst_order[q]) : q;
            rg.dpair[i >> 1] = (rg.dpair[i >> 1] & ~(0xffffu << (16 * (i & 1)))) | ((uint32_t)k << (16 * (i & 1)));
            const int kb = k == 0 ? HALF : M - k;
            uint32_t da, db;
            if (tb.map1) {
                rg.map_src[i] = ((uint32_t)TSB_LDG(&tb.map1[k]) & 0xffffu) | ((uint32_t)TSB_LDG(&tb.map1[kb]) << 16);
                da = src_lo(rg.map_src[i]) >= 0;
                db = src_hi(rg.map_src[i]) >= 0;
            } else {
                da = (TSB_LDG(&tb.map[k]) & 0xffffu) != 0;
                db = (TSB_LDG(&tb.map[kb]) & 0xffffu) != 0;
            }
            // source pair of this slot
            const int qb = is_special(i, q) ? HALF : M - q;
            uint32_t f = (uint32_t)TSB_LDG(&tb.src_used[q]) | ((uint32_t)TSB_LDG(&tb.src_used[qb]) << 1);
            rg.used |= (f | (da << 16) | (db << 17)) << (2 * i);
        }
        bool ld = a.load_state != 0;
        if (PROG == PROG_PV_SHIFT && ld) {
            if (a.mode == 1) {
                // pre-pass: increments are summed from zero; last_phase comes from the caller only for a
                // segment that starts the render (no warm-up frame)
                if (!sg.warm && a.st_last_phase) regs_from(tid, a.st_last_phase, sg.row, rg.last_phase);
            } else {
                if (sg.warm && a.st_warm_phase) regs_from(tid, a.st_warm_phase, sg.row, rg.last_phase);
                else if (a.st_last_phase) regs_from(tid, a.st_last_phase, sg.row, rg.last_phase);
                if (a.st_phase_sum) regs_from(a.st_phase_sum, sg.row, rg.phase_sum, [&](int i) { return dst_k(rg, i); });
            }
        }
        for (int i = tid; i < N; i += T) {
            otype v = 0;
            if (ld && a.st_ola && a.mode == 0) v = (otype)a.st_ola[sg.row * N + i];
            sm.ola[(int)((sg.out_off0 + i) & (N - 1))] = v;
        }
    }
    static TSB_HD void store_state(int tid, Smem &sm, Regs &rg, const ChainArgs &a, const SegInfo &sg, long long out_pos) {
        if (PROG == PROG_PV_SHIFT && a.st_last_phase && a.st_phase_sum) {
            regs_to(tid, a.st_last_phase, sg.row, rg.last_phase);
            regs_to(a.st_phase_sum, sg.row, rg.phase_sum, [&](int i) { return dst_k(rg, i); });
        }
        if (a.st_ola && a.mode == 0)
            for (int i = tid; i < N; i += T) a.st_ola[sg.row * N + i] = (double)sm.ola[(int)((out_pos + i) & (N - 1))];
    }
    // pre-pass "synthesis": only the phase accumulation of rtlib/pv.hpp:34 (same order of additions as the
    // full kernel), so that a segment's total increment per bin is known before anything is synthesised
    // inc_row: where this frame's increments go, [M] (or nullptr)
    static TSB_HD void phase_pv_accumulate(int tid, Smem &sm, Regs &rg, const Tables &tb, double *inc_row) {
        const double *inc_x = xch_inc(sm);
        const real *amp_x = xch_amp(sm);
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = dst_k(rg, i);
            if (HALF % T != 0 && k < 0) break;
            int kb = is_special(i, k) ? HALF : M - k;
            real amp;
            double inc;
            if (tb.map1) gather_bin<true>(amp_x, inc_x, tb, k, src_lo(rg.map_src[i]), amp, inc);
            else gather_bin<false>(amp_x, inc_x, tb, k, 0, amp, inc);
            rg.phase_sum[2 * i] = PvMath<real>::accumulate(rg.phase_sum[2 * i], inc);
            if (inc_row) inc_row[k] = inc;
            if (tb.map1) gather_bin<true>(amp_x, inc_x, tb, kb, src_hi(rg.map_src[i]), amp, inc);
            else gather_bin<false>(amp_x, inc_x, tb, kb, 0, amp, inc);
            rg.phase_sum[2 * i + 1] = PvMath<real>::accumulate(rg.phase_sum[2 * i + 1], inc);
            if (inc_row) inc_row[kb] = inc;
        }
    }

    // ------------------------------------------------------------------ input ring
    // Bring absolute samples [first, first+count) of this lane into the ring (zeros past in_len).
    // On the device this is an asynchronous global->shared copy (cp.async / LDGSTS): it is issued
    // right after the window phase has consumed the slots it overwrites and completes in the
    // background while the frame is transformed; ring_wait() + a barrier precede the next use.
    // NT = number of threads taking part (tid < NT)
    template <int NT = T>
    static TSB_HD void ring_fill(int tid, int lane, Smem &sm, const ChainArgs &a, long long first, int count) {
        const float *src = a.in + (long long)lane * a.in_stride;
#if defined(__CUDA_ARCH__)
        bool vec = ((first | (long long)count | a.in_stride) & 3) == 0 && first >= 0 && first + count <= a.in_len &&
                   ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
        if (vec) {
            for (int i = tid * 4; i < count; i += NT * 4) {
                long long s0 = first + i;
                unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.ring[(int)(s0 & (N - 1))]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + s0) : "memory");
            }
            return;
        }
        for (int i = tid; i < count; i += NT) {
            long long s0 = first + i;
            float *d = &sm.ring[(int)(s0 & (N - 1))];
            if (s0 >= 0 && s0 < a.in_len) {
                unsigned dst = (unsigned)__cvta_generic_to_shared(d);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src + s0) : "memory");
            } else {
                *d = 0.f;
            }
        }
#else
        for (int i = tid; i < count; i += NT) {
            long long s0 = first + i;
            sm.ring[(int)(s0 & (N - 1))] = (s0 >= 0 && s0 < a.in_len) ? src[s0] : 0.f;
        }
#endif
    }
    static TSB_HD void ring_wait() {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
    }

    // ------------------------------------------------------------------ forward passes
    // Pass 0 gathers straight from the input ring and applies the analysis window on the way
    // (rtlib/windowfn.hpp fused into the load): z[m] = x[2m]*wa[2m] + i*x[2m+1]*wa[2m+1].
    // Swizzled buffer access with index = a + c, a per-thread, c compile-time, bit-disjoint (Swz is linear).
    static TSB_HD C ld_sw(const C *buf, int a, int c) { return buf[SW::f(a) ^ SW::f(c)]; }
    static TSB_HD void st_sw(C *buf, int a, int c, C v) { buf[SW::f(a) ^ SW::f(c)] = v; }

    template <int P, int NT = T>
    static TSB_HD void fwd_pass(int tid, Smem &sm, const Tables &tb, long long in_off) {
        C *dst = sm.buf[P & 1];
        auto out = [&](int a, int c, C v) { st_sw(dst, a, c, v); };
        if (P == 0) {
            const int rb = (int)(in_off & (N - 1));
            const C *wa = reinterpret_cast<const C *>(tb.win_a);
            real e_acc = 0;  // WRAP_FIX: this thread's share of the windowed frame's energy
            if ((rb & 1) == 0) {
                const cpx<float> *ring2 = reinterpret_cast<const cpx<float> *>(sm.ring);
                const int rb2 = rb >> 1;
                auto in = [&](int a, int c) {
                    cpx<float> q = ring2[(rb2 + a + c) & (M - 1)];
                    C w = TSB_LDG(wa + a + c);
                    C z = mk<real>((real)q.x * w.x, (real)q.y * w.y);
                    if (WRAP_FIX) e_acc = fmaf(z.x, z.x, fmaf(z.y, z.y, e_acc));
                    return z;
                };
                stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE>(tid, tb.tw, in, out);
            } else {
                auto in = [&](int a, int c) {
                    int m = a + c;
                    float x0 = sm.ring[(rb + 2 * m) & (N - 1)];
                    float x1 = sm.ring[(rb + 2 * m + 1) & (N - 1)];
                    C w = TSB_LDG(wa + m);
                    C z = mk<real>((real)x0 * w.x, (real)x1 * w.y);
                    if (WRAP_FIX) e_acc = fmaf(z.x, z.x, fmaf(z.y, z.y, e_acc));
                    return z;
                };
                stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE>(tid, tb.tw, in, out);
            }
            if (WRAP_FIX) energy_store(tid, sm, (float)e_acc);
        } else {
            const C *src = sm.buf[(P - 1) & 1];
            auto in = [&](int a, int c) { return ld_sw(src, a, c); };
            stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE>(tid, tb.tw, in, out);
        }
    }
    // Warp-shuffle reduction of the per-thread energy shares; lane 0 of every warp of the pass keeps the warp's sum.
    // (Host replay: threads run one after the other, the first lane of a warp starts the sum.)
    static TSB_HD void energy_store(int tid, Smem &sm, float e) {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((tid & 31) == 0) sm.energy[tid >> 5] = e;
#else
        if ((tid & 31) == 0) sm.energy[tid >> 5] = e;
        else sm.energy[tid >> 5] += e;
#endif
    }
    // WRAP_K * sqrt(energy of the windowed frame): the doubt threshold of PvMath<float>::analyze.  NW = warps that ran
    // forward pass 0 (their partial sums; the window table carries 1/sqrt(N), so sqrt(energy) is the RMS bin magnitude).
    template <int NT>
    static TSB_HD float wrap_threshold(const Smem &sm) {
        float e = 0;
#pragma unroll
        for (int w = 0; w < (NT + 31) / 32; w++) e += sm.energy[w];
        return PvMath<float>::WRAP_K * (e * fast_rsqrtf(fmaxf(e, 1e-37f)));  // sqrt(e), one MUFU, no slow-path call
    }

    // spectrum of the real frame at bins k and M-k (k != 0) from the packed transform Zc (see file header)
    static TSB_HD void split_pair(const C *Zc, const Tables &tb, int k, C &Xa, C &Xb) {
        split_vals(Zc[SW::f(k)], conj(Zc[SW::f(M - k)]), tb, k, Xa, Xb);
    }
    static TSB_HD void split_vals(C a, C b, const Tables &tb, int k, C &Xa, C &Xb) {
        C e = mk<real>((real)0.5 * (a.x + b.x), (real)0.5 * (a.y + b.y));
        C d = a - b;
        C o = mk<real>((real)0.5 * d.y, (real)-0.5 * d.x);  // -i*d/2
        C wo = cmul(TSB_LDG(&tb.twn[k]), o);
        Xa = e + wo;           // X[k]
        Xb = conj(e - wo);     // X[M-k]
    }
    // C2R pre-processing for the pair (k, M-k), k != 0: Hermitian-half values Zk, Zp -> conj(G) entries
    static TSB_HD void merge_pair(C *H, const Tables &tb, int k, C Zk, C Zp) {
        C hk, hp;
        merge_vals(tb, k, Zk, Zp, hk, hp);
        H[hpos(k)] = hk;
        H[hpos(M - k)] = hp;
    }
    static TSB_HD void merge_vals(const Tables &tb, int k, C Zk, C Zp, C &hk, C &hp) {
        C ee = Zk + conj(Zp);
        C d = Zk - conj(Zp);
        C oo = cmul(d, conj(TSB_LDG(&tb.twn[k])));
        hk = mk<real>(ee.x - oo.y, -ee.y - oo.x);
        hp = mk<real>(ee.x + oo.y, ee.y - oo.x);
    }
    // Only the first slot of thread 0 holds the irregular pair (bins 0 and M/2): the test is compile-time
    // false for every other slot.  In the PV phases slot 0 computes the regular pair for every thread and thread 0
    // selects its values afterwards: a branch there would keep the compiler from overlapping the two bins of the
    // slot (half of a thread's transcendental work when it owns two slots).
    static TSB_HD bool is_special(int i, int k) { return i == 0 && k == 0; }

    // ------------------------------------------------------------------ fp32 PV: exact wraps for doubtful bins
    // What the frame being analysed follows (WRAP_FIX): hop of the frame before it (>= 0: its samples start that many
    // before this frame's), PH_ZERO: none — last_phase is the exact zero of a reset PhaseVocoder, PH_UNKNOWN: not known
    // (a continued render without ChainArgs::prev_hop; a warm-up frame, whose increments nobody uses): no fix.
    static constexpr int PH_ZERO = -1, PH_UNKNOWN = -2;
    // Partial sums over n = l, l + NL, ... of the frft(+1) bin j (kernel e^{+2 pi i j n / N}) of the windowed frames
    // starting at input offsets off_c and off_p, in double precision straight from the float32 input rows.
    static TSB_HD void dft_bin_partial(const float *src, long long in_len, const double *win64, long long off_c,
                                       long long off_p, int j, int l, int NL, double &rc, double &ic, double &rp, double &ip) {
        auto unit = [](int q, double &c, double &sn) {  // e^{2 pi i q / N}, q reduced mod N (table-based, no slow-path call)
            fast_sincos((double)(q & (N - 1)) * (2.0 * TSB_PI / (double)N), sn, c);
        };
        double wx, wy, sx, sy;
        unit(j * l, wx, wy);
        unit(j * NL, sx, sy);
        rc = ic = rp = ip = 0.0;
        // (plain loads, unrolled: the samples come from L2 / HBM and several iterations' requests must be in flight)
#pragma unroll 8
        for (int n = l; n < N; n += NL) {
#if defined(__CUDA_ARCH__)
            const double wv = __ldg(&win64[n]);
#else
            const double wv = win64[n];
#endif
            const long long sc = off_c + n, sp = off_p + n;
            const double xc = (sc >= 0 && sc < in_len) ? (double)src[sc] * wv : 0.0;
            const double xp = (sp >= 0 && sp < in_len) ? (double)src[sp] * wv : 0.0;
            rc = fma(xc, wx, rc);
            ic = fma(xc, wy, ic);
            rp = fma(xp, wx, rp);
            ip = fma(xp, wy, ip);
            const double nx = wx * sx - wy * sy;
            wy = fma(wx, sy, wy * sx);
            wx = nx;
        }
    }
    // bin index of doubt bit b of the thread `stid`: slot b >> 1, side b & 1 (k or M - k; thread 0 slot 0: 0 / M/2)
    static TSB_HD int doubt_bin(int stid, int b) {
        const int i = b >> 1, k = stid + T * i;
        return (b & 1) ? (is_special(i, k) ? HALF : M - k) : k;
    }
    // Every thread of the warp calls wrap_fix at the end of the analysis phase with the doubt bits of its own bins.
    // When the warp has any, it takes them one by one: all lanes sum a share of the two direct DFTs, a shuffle
    // reduction brings them together, and the owning lane replaces the bin's increment by the exact one.  The slow
    // path is a real function call (not inlined, scalars by value) so that it costs the frame loop no registers.
    struct WrapFixArgs {
        const float *src;      // this lane's input row
        long long in_len, cur_off, prev_off;
        const double *win64;
        double *inc_x;
        double hop, inv_n, ratio;
        int ph;
    };
#if defined(__CUDA_ARCH__)
    static __device__ TSB_WRAPFIX_INLINE void wrap_fix_slow(int tid, uint32_t doubts, WrapFixArgs w) {
        unsigned any = __ballot_sync(0xffffffffu, doubts != 0);
        while (any) {
            const int sl = __ffs(any) - 1;
            any &= any - 1;
            uint32_t fm = __shfl_sync(0xffffffffu, doubts, sl);
            const int stid = (tid & ~31) | sl;
            while (fm) {
                const int b = __ffs(fm) - 1;
                fm &= fm - 1;
                const int j = doubt_bin(stid, b);
                double v[4];
                dft_bin_partial(w.src, w.in_len, w.win64, w.cur_off, w.prev_off, j, tid & 31, 32, v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
                if ((tid & 31) == sl) {
                    const double expect = mul_rn(mul_rn((double)(2 * j), TSB_PI), w.hop) * w.inv_n;  // expected_advance
                    w.inc_x[j] = PvMath<float>::exact_inc(expect, w.ratio, w.inc_x[j], fast_atan2(v[1], v[0]),
                                                          w.ph >= 0 ? fast_atan2(v[3], v[2]) : 0.0);
                }
            }
        }
    }
#endif
    static TSB_HD void phase_wrap_fix(int tid, int lane, Smem &sm, Regs &rg, const Tables &tb, const ChainArgs &a,
                                      long long cur_off, int ph) {
        uint32_t doubts = rg.doubts;
#if defined(__CUDA_ARCH__)
        if (!__any_sync(0xffffffffu, doubts != 0)) return;
#else
        if (!doubts) return;
#endif
        WrapFixArgs w;
        w.src = a.in + (long long)lane * a.in_stride;
        w.in_len = a.in_len;
        w.cur_off = cur_off;
        w.prev_off = ph >= 0 ? cur_off - ph : cur_off;
        w.win64 = tb.win64;
        w.inc_x = xch_inc(sm);
        w.hop = tb.hop;
        w.inv_n = tb.inv_n;
        w.ratio = tb.ratio;
        w.ph = ph;
#if defined(__CUDA_ARCH__)
        wrap_fix_slow(tid, doubts, w);
#else
        while (doubts) {
            int b = 0;
            while (!((doubts >> b) & 1)) b++;
            doubts &= doubts - 1;
            const int j = doubt_bin(tid, b);
            double v[4] = {0, 0, 0, 0};
            for (int l = 0; l < 32; l++) {
                double q[4];
                dft_bin_partial(w.src, w.in_len, w.win64, w.cur_off, w.prev_off, j, l, 32, q[0], q[1], q[2], q[3]);
                for (int u = 0; u < 4; u++) v[u] += q[u];
            }
            w.inc_x[j] = PvMath<float>::exact_inc(expected_advance(j, tb), w.ratio, w.inc_x[j], fast_atan2(v[1], v[0]),
                                                  w.ph >= 0 ? fast_atan2(v[3], v[2]) : 0.0);
        }
#endif
    }

    // ------------------------------------------------------------------ PV phase C: split + analysis
    // (rtlib/pv.hpp:17-28 over the live bins 0..N/2-1; results go to the exchange arrays)
    // `all`: analyse every bin (last frame of a launch, warm-up frame); otherwise only the sources the shift uses.
    // Slots are taken two at a time and their k-side bins analysed together, then their (M-k)-side bins: an upward
    // shift drops the upper sources, i.e. the (M-k) side of most pairs, and this grouping lets a thread skip them
    // while the two bins it does analyse still overlap.
    // lane / a / cur_off / ph: what wrap_fix needs (fp32 only); NT0 = threads that ran forward pass 0.
    // Returns true when this thread found a doubtful wrap (fp32 only): the caller then runs phase_wrap_fix.
    static TSB_HD bool phase_pv_analyze(int tid, Smem &sm, Regs &rg, const Tables &tb, bool all, bool fix) {
        const C *Zc = fwd_result(sm);
        double *inc_x = xch_inc(sm);
        real *amp_x = xch_amp(sm);
        float thr = 0.f;
        uint32_t doubts = 0;
        if constexpr (WRAP_FIX) thr = wrap_threshold<PIPE ? T0 : T>(sm);
        auto split_slot = [&](int i, int k, C &Xa, C &Xb) {
            if (i == 0) {
                const bool sp = is_special(i, k);
                C za = Zc[SW::f(k)], zb = conj(Zc[SW::f(sp ? HALF : M - k)]);  // thread 0: Z[0], conj(Z[M/2]) = X[M/2]
                split_vals(za, zb, tb, k, Xa, Xb);
                Xa = sp ? mk<real>(za.x + za.y, 0) : Xa;  // X[0]
                Xb = sp ? zb : Xb;
            } else {
                split_pair(Zc, tb, k, Xa, Xb);
            }
        };
        // dbit: this bin's bit in `doubts` (2 * slot + side)
        auto analyze_bin = [&](int j, C X, real &last_phase, bool dc, int dbit) {
            real amp;
            double inc;
            if constexpr (sizeof(real) == 4) {
                bool doubt;
                PvMath<real>::analyze(j, X, last_phase, tb, N, amp, inc, dc, thr, doubt);
                if (WRAP_FIX) doubts |= (uint32_t)doubt << dbit;
            } else {
                (void)dbit;
                PvMath<real>::analyze(j, X, last_phase, tb, N, amp, inc, dc);
            }
            amp_x[j] = (real)0.5 * amp;  // Y[k >= N/2] = 0 after shift, so Z = conj(Y)/2: exact scaling, done here
            inc_x[j] = inc;
        };
        if constexpr (S == 1) {
            int k = tid;
            const bool sp = is_special(0, k);
            const bool in_range = !(HALF % T != 0 && k >= HALF);
            const bool need_a = in_range && (all || (rg.used & 1u)), need_b = in_range && (all || (rg.used & 2u));
            if (need_a || need_b) {
                C Xa, Xb;
                split_slot(0, k, Xa, Xb);
                if (need_a && need_b) {  // (one pair per thread: the two bins overlap only when both are wanted)
                    analyze_bin(k, Xa, rg.last_phase[0], sp, 0);
                    analyze_bin(sp ? HALF : M - k, Xb, rg.last_phase[1], false, 1);
                } else if (need_a) {
                    analyze_bin(k, Xa, rg.last_phase[0], sp, 0);
                } else {
                    analyze_bin(sp ? HALF : M - k, Xb, rg.last_phase[1], false, 1);
                }
            }
        } else {
#pragma unroll
        for (int i = 0; i < S; i += 2) {
            const int k0 = tid + T * i, k1 = tid + T * (i + 1);
            const bool two = i + 1 < S;
            const bool sp = is_special(i, k0);
            const uint32_t u = rg.used >> (2 * i);
            const bool need_a = all || (u & (two ? 5u : 1u)), need_b = all || (u & (two ? 10u : 2u));
            if (need_a || need_b) {
                C Xa0, Xb0, Xa1, Xb1;
                split_slot(i, k0, Xa0, Xb0);
                if (two) split_slot(i + 1, k1, Xa1, Xb1);
                if (need_a) {
                    analyze_bin(k0, Xa0, rg.last_phase[2 * i], sp, 2 * i);
                    if (two) analyze_bin(k1, Xa1, rg.last_phase[2 * i + 2], false, 2 * i + 2);
                }
                if (need_b) {
                    analyze_bin(sp ? HALF : M - k0, Xb0, rg.last_phase[2 * i + 1], false, 2 * i + 1);
                    if (two) analyze_bin(M - k1, Xb1, rg.last_phase[2 * i + 3], false, 2 * i + 3);
                }
            }
            sched_fence();
        }
        }
        if constexpr (WRAP_FIX) {
            if (!fix) doubts = 0;
            rg.doubts = doubts;
            return doubts != 0;
        }
        return false;
    }
    // shift gather for destination bin k (rtlib/pv.hpp:70-78: amplitudes add in ascending source
    // order, the frequency of the last source wins; bins with no source are silent with freq 0).
    // SIMPLE: at most one source per destination (any upward shift) — a table lookup and two selects.
    // `src1`: the bin's source from Regs::map_src (SIMPLE only).
    template <bool SIMPLE>
    static TSB_HD void gather_bin(const real *amp_x, const double *inc_x, const Tables &tb, int k, int src1, real &amp,
                                  double &inc) {
        if (SIMPLE) {
            int j = src1;
            int jj = j < 0 ? 0 : j;
            real a = amp_x[jj];
            double d = inc_x[jj];
            amp = j < 0 ? (real)0 : a;
            inc = j < 0 ? 0.0 : d;
            return;
        }
        amp = 0;
        inc = 0;
        uint32_t e = TSB_LDG(&tb.map[k]);
        int first = (int)(e >> 16), cnt = (int)(e & 0xffff);
#pragma unroll 1
        for (int s = 0; s < cnt; s++) {
            int j = TSB_LDG(&tb.map_src[first + s]);
            amp = amp + amp_x[j];
            inc = inc_x[j];
        }
    }
    static TSB_HD int src_lo(uint32_t w) { return (int)(short)(w & 0xffffu); }
    static TSB_HD int src_hi(uint32_t w) { return (int)w >> 16; }
    // ------------------------------------------------------------------ PV phase D: synthesis + C2R pre
    // (rtlib/pv.hpp:29-38 on the shifted bins, then Z = conj(Y) [1/2 folded into the amplitudes])
    static TSB_HD void phase_pv_synth(int tid, Smem &sm, Regs &rg, const Tables &tb) {
        if (tb.map1) phase_pv_synth_t<true>(tid, sm, rg, tb);  // uniform: decided once per plan
        else phase_pv_synth_t<false>(tid, sm, rg, tb);
    }
    template <bool SIMPLE>
    static TSB_HD void phase_pv_synth_t(int tid, Smem &sm, Regs &rg, const Tables &tb) {
        C *H = inv_buffer(sm, 0);  // not pipelined: the consumed packed spectrum's buffer; pipelined: ibuf[0]
        const double *inc_x = xch_inc(sm);
        const real *amp_x = xch_amp(sm);
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = dst_k(rg, i);
            if (HALF % T != 0 && k < 0) break;
            const bool sp = is_special(i, k);
            int kb = sp ? HALF : M - k;
            if (!((rg.used >> (16 + 2 * i)) & 3u)) {
                // neither bin has a source: silent, phase_sum unchanged (rtlib/pv.hpp:34 adds 0)
                H[hpos(k)] = mk<real>(0, 0);
                H[hpos(kb)] = mk<real>(0, 0);
                continue;
            }
            real amp;
            double inc;
            gather_bin<SIMPLE>(amp_x, inc_x, tb, k, src_lo(rg.map_src[i]), amp, inc);
            C Ya = PvMath<real>::template synth<FAST_SYNTH>(amp, inc, rg.phase_sum[2 * i]);
            gather_bin<SIMPLE>(amp_x, inc_x, tb, kb, src_hi(rg.map_src[i]), amp, inc);
            C Yb = PvMath<real>::template synth<FAST_SYNTH>(amp, inc, rg.phase_sum[2 * i + 1]);
            if (i == 0) {
                C hk, hp;
                merge_vals(tb, k, conj(Ya), conj(Yb), hk, hp);
                const real z0 = (real)2 * Ya.x;  // thread 0: Z[0] = Re Y[0] (not halved), Z[M] = 0
                H[hpos(k)] = sp ? mk<real>(z0, -z0) : hk;
                H[hpos(kb)] = sp ? mk<real>((real)2 * Yb.x, (real)-2 * Yb.y) : hp;  // 2*Z[M/2], Z = conj(Y)
            } else {
                merge_pair(H, tb, k, conj(Ya), conj(Yb));
            }
            sched_fence();
        }
    }
    // ------------------------------------------------------------------ gain program: split * g -> C2R pre
    static TSB_HD void phase_gain(int tid, Smem &sm, const Tables &tb) {
        const C *Zc = fwd_result(sm);
        C *H = inv_buffer(sm, 0);  // not pipelined: the spare forward buffer; pipelined: ibuf[0]
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = tid + T * i;
            if (HALF % T != 0 && k >= HALF) break;
            C Xa, Xb;
            if (!is_special(i, k)) {
                split_pair(Zc, tb, k, Xa, Xb);
                real ga = TSB_LDG(&tb.gain[k]), gb = TSB_LDG(&tb.gain[M - k]);
                merge_pair(H, tb, k, mk<real>(ga * Xa.x, ga * Xa.y), mk<real>(gb * Xb.x, gb * Xb.y));
            } else {
                C z0 = Zc[SW::f(0)], zh = Zc[SW::f(HALF)];
                real x0 = (z0.x + z0.y) * TSB_LDG(&tb.gain[0]);  // X[0]
                real xm = (z0.x - z0.y) * TSB_LDG(&tb.gain[M]);  // X[N/2]
                real gh = TSB_LDG(&tb.gain[HALF]);
                H[hpos(0)] = mk<real>(x0 + xm, -(x0 - xm));
                H[hpos(HALF)] = mk<real>((real)2 * gh * zh.x, (real)-2 * gh * zh.y);  // 2*Z[M/2], Z = g*conj(zh)
            }
        }
    }

    // register tail <-> the shared-memory ring (start and end of a launch; `out_pos` = position of the next frame)
    static TSB_HD void tail_load(int tid, Smem &sm, Regs &rg, long long out_pos) {
        if (PIPE && tid < T0) return;
        const int j = PIPE ? tid - T0 : tid;
#pragma unroll
        for (int u = 0; u < RT_N; u++) {
            const long long n = out_pos + 2 * (j + T0 * u);
            rg.tail[u] = mk<otype>(sm.ola[(int)(n & (N - 1))], sm.ola[(int)((n + 1) & (N - 1))]);
        }
    }
    static TSB_HD void tail_store(int tid, Smem &sm, const Regs &rg, long long out_pos) {
        if (PIPE && tid < T0) return;
        const int j = PIPE ? tid - T0 : tid;
#pragma unroll
        for (int u = 0; u < RT_PP; u++) {
            const long long n = out_pos + 2 * (j + T0 * u);
            sm.ola[(int)(n & (N - 1))] = u < RT_N ? rg.tail[u < RT_N ? u : 0].x : (otype)0;
            sm.ola[(int)((n + 1) & (N - 1))] = u < RT_N ? rg.tail[u < RT_N ? u : 0].y : (otype)0;
        }
    }
    // ------------------------------------------------------------------ inverse passes
    // The last pass hands its outputs F[m] straight to the synthesis window and the overlap-add:
    // x[2m] = Re F[m], x[2m+1] = -Im F[m]; out = x*ws; accumulate into the ring; the first `hop`
    // positions are complete (tinyspec.cpp:173-182) and leave the chip, the rest is the tail (194-200).
    template <int P, int NT = T, bool EMIT = true, bool RT = false>
    static TSB_HD void inv_pass(int tid, int lane, Smem &sm, Regs &rg, const Tables &tb, const ChainArgs &a, long long out_pos,
                                int hop) {
        const C *src = inv_buffer(sm, P);
        auto in = [&](int x, int c) { return INV_INPLACE ? src[ISW::f(x) ^ ISW::f(c)] : ld_sw(src, x, c); };
        if (P < PASSES - 1) {
            C *dst = inv_buffer(sm, P + 1);
            auto out = [&](int x, int c, C v) {
                if (INV_INPLACE) dst[ISW::f(x) ^ ISW::f(c)] = v;
                else st_sw(dst, x, c, v);
            };
            stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE, INV_INPLACE>(tid, tb.tw, in, out);
        } else {
            otype *dstg = reinterpret_cast<otype *>(a.out) + (long long)lane * a.out_stride + out_pos;
            const C *ws = reinterpret_cast<const C *>(tb.win_s);
            const int ob = (int)(out_pos & (N - 1));
            // samples of this frame that may be written: n < lim (everything when the row has room)
            const long long room = a.out_len - out_pos;
            // (a lead-in frame of a time segment writes nothing: it only rebuilds the overlap tail)
            const int lim = !EMIT ? 0 : (room >= (long long)N + hop ? 0x7fffffff : (room < 0 ? 0 : (int)room));
            // one output sample: `done` positions leave the chip, the others stay in the tail.  OUT32 follows the
            // reference's types: the finished sample is the float of (float tail + double value) (:180), the tail
            // accumulates the value narrowed first, in float (:195-197).
            // Products and sums are single IEEE operations (no FMA contraction): a sample must not depend on which
            // of the two code paths below produced it (chunked and whole renders are bit-identical).
            auto emit_val = [](otype tail, real v) -> otype { return (otype)add_rn((real)tail, v); };
            auto tail_val = [](otype tail, real v) -> otype { return OUT32 ? add_rn(tail, (otype)v) : (otype)add_rn((real)tail, v); };
            if constexpr (RT) {
                // the tail lives in registers (see RT_OK): pair u of this thread finishes for u < 2 and moves to slot
                // u - 2 otherwise; slots 6 and 7 start from nothing
                static_assert(!RT || NT == T0, "register tail: the last inverse pass is run by T0 threads");
                const bool al = (reinterpret_cast<uintptr_t>(dstg) & (2 * sizeof(otype) - 1)) == 0;
                const int emit_end = hop < lim ? hop : lim;
                OC nt[RT_N];  // the new tail (the pass does not hand its outputs over in slot order)
                auto out = [&](int x, int c, C f) {
                    const int u = c / T0;  // a constant once the pass is unrolled (last pass: x = tid, c = T0 * u)
                    const int m = x + c;
                    C w = TSB_LDG(ws + m);
                    const real vx = mul_rn(f.x, w.x), vy = -mul_rn(f.y, w.y);
                    const OC t = u < RT_N ? rg.tail[u < RT_N ? u : 0] : mk<otype>((otype)0, (otype)0);
                    if (u < RT_SH) {
                        const otype ex = emit_val(t.x, vx), ey = emit_val(t.y, vy);
                        if (al && 2 * m + 1 < emit_end) {
                            *reinterpret_cast<OC *>(dstg + 2 * m) = mk<otype>(ex, ey);
                        } else {
                            if (2 * m < emit_end) dstg[2 * m] = ex;
                            if (2 * m + 1 < emit_end) dstg[2 * m + 1] = ey;
                        }
                    } else {
                        nt[u >= RT_SH ? u - RT_SH : 0] = mk<otype>(tail_val(t.x, vx), tail_val(t.y, vy));
                    }
                };
                stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE, INV_INPLACE>(tid, tb.tw, in, out);
#pragma unroll
                for (int u = 0; u < RT_N; u++) rg.tail[u] = nt[u];
            } else if (((ob | hop | (lim == 0x7fffffff ? 0 : lim)) & 1) == 0 &&
                ((reinterpret_cast<uintptr_t>(dstg) & (2 * sizeof(otype) - 1)) == 0)) {
                // pairs (2m, 2m+1) never straddle the hop boundary or the end of the row and are aligned in the
                // ring and in the row: one 2-sample load / add / (predicated) store per pair, no branches
                OC *ola2 = reinterpret_cast<OC *>(sm.ola);
                const int ob2 = ob >> 1;
                const int emit_end = hop < lim ? hop : lim;
                auto out = [&](int x, int c, C f) {
                    const int m = x + c;
                    C w = TSB_LDG(ws + m);
                    OC *slot = &ola2[(ob2 + m) & (M - 1)];
                    const OC t = *slot;
                    const real vx = mul_rn(f.x, w.x), vy = -mul_rn(f.y, w.y);
                    const bool done = 2 * m < hop;
                    if (2 * m < emit_end) *reinterpret_cast<OC *>(dstg + 2 * m) = mk<otype>(emit_val(t.x, vx), emit_val(t.y, vy));
                    *slot = mk<otype>(done ? (otype)0 : tail_val(t.x, vx), done ? (otype)0 : tail_val(t.y, vy));
                };
                stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE, INV_INPLACE>(tid, tb.tw, in, out);
            } else {
                auto out = [&](int x, int c, C f) {
                    const int m = x + c;
                    C w = TSB_LDG(ws + m);
                    real v[2] = {mul_rn(f.x, w.x), -mul_rn(f.y, w.y)};
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        int n = 2 * m + u;
                        int slot = (ob + n) & (N - 1);
                        otype t = sm.ola[slot];
                        otype acc = tail_val(t, v[u]);
                        if (n < hop) {
                            if (n < lim) dstg[n] = emit_val(t, v[u]);
                            acc = 0;
                        }
                        sm.ola[slot] = acc;
                    }
                };
                stockham_pass_io<real, LOG2M, NT, P, decltype(in), decltype(out), TW_DERIVE, INV_INPLACE>(tid, tb.tw, in, out);
            }
            // hop larger than the frame: nothing overlaps those positions (tinyspec.cpp:184-191)
            for (int n = N + tid; n < hop; n += NT)
                if (n < lim) dstg[n] = 0;
        }
    }
};

// ---- compile-time pass loops ----
template <class CH, int P, int END>
struct FwdPasses {
    template <class Exec>
    static TSB_HD void run(Exec &ex, int lane, typename CH::Smem &sm, const typename CH::Tables &tb, const ChainArgs &a,
                           long long in_off, long long next_first, int next_cnt) {
        ex.each([&](int tid, typename CH::Regs &) {
            // after pass 0 nobody reads the ring any more: fetch the next frame's new samples behind pass 1
            if (P == 1 && next_cnt > 0) CH::ring_fill(tid, lane, sm, a, next_first, next_cnt);
            CH::template fwd_pass<P>(tid, sm, tb, in_off);
        });
        FwdPasses<CH, P + 1, END>::run(ex, lane, sm, tb, a, in_off, next_first, next_cnt);
    }
};
template <class CH, int END>
struct FwdPasses<CH, END, END> {
    template <class Exec>
    static TSB_HD void run(Exec &, int, typename CH::Smem &, const typename CH::Tables &, const ChainArgs &, long long,
                           long long, int) {}
};
template <class CH, int P, int END>
struct InvPasses {
    template <bool EMIT, bool RT, class Exec>
    static TSB_HD void run(Exec &ex, int lane, typename CH::Smem &sm, const typename CH::Tables &tb, const ChainArgs &a,
                           long long out_pos, int hop) {
        ex.each([&](int tid, typename CH::Regs &rg) {
            CH::template inv_pass<P, CH::T, EMIT, RT>(tid, lane, sm, rg, tb, a, out_pos, hop);
            if (P == END - 1) CH::ring_wait();
        });
        InvPasses<CH, P + 1, END>::template run<EMIT, RT>(ex, lane, sm, tb, a, out_pos, hop);
    }
};
template <class CH, int END>
struct InvPasses<CH, END, END> {
    template <bool EMIT, bool RT, class Exec>
    static TSB_HD void run(Exec &, int, typename CH::Smem &, const typename CH::Tables &, const ChainArgs &, long long, int) {}
};

// Software-pipelined pass pairs (CH::PIPE): in one barrier interval the first T0 threads run forward pass P of
// the NEXT frame while the other T0 threads run inverse pass P of the CURRENT frame (different buffers; the
// split is at warp granularity).  Both transforms have the same pass structure, so they stay in step.
template <class CH, int P, int END>
struct PipePasses {
    template <bool EMIT, bool RT, class Exec>
    static TSB_HD void run(Exec &ex, int lane, typename CH::Smem &sm, const typename CH::Tables &tb, const ChainArgs &a,
                           bool have_next, long long next_in_off, long long fetch_first, int fetch_cnt, bool have_cur,
                           long long out_pos, int hop) {
        auto body = [&](int tid, typename CH::Regs &rg) {
            if (tid < CH::T0) {
                if (have_next) {
                    // The samples this frame adds were requested during the previous frame's last pass (below) and
                    // had the whole PV stage to arrive (the synthesis phase ends with ring_wait + barrier).  The
                    // request for the frame after goes out in the LAST forward pass — the lightest one, opposite
                    // the inverse side's overlap-add pass — so that the two halves of the CTA reach each barrier
                    // at about the same time.
                    CH::template fwd_pass<P, CH::T0>(tid, sm, tb, next_in_off);
                    if (P == END - 1 && fetch_cnt > 0) CH::template ring_fill<CH::T0>(tid, lane, sm, a, fetch_first, fetch_cnt);
                }
            } else if (have_cur) {
                CH::template inv_pass<P, CH::T0, EMIT, RT>(tid - CH::T0, lane, sm, rg, tb, a, out_pos, hop);
            }
        };
        // The two halves work on buffers of their own (buf / ibuf) and meet again only in the PV phases: between
        // their passes each half synchronises with itself alone (a named barrier over its T0 threads), so a heavy
        // pass on one side (the windowed load, the overlap-add) no longer holds up the other side pass by pass.
        if (P < END - 1) ex.each_half(body, CH::T0);
        else ex.each(body);
        PipePasses<CH, P + 1, END>::template run<EMIT, RT>(ex, lane, sm, tb, a, have_next, next_in_off, fetch_first, fetch_cnt,
                                                           have_cur, out_pos, hop);
    }
};
template <class CH, int END>
struct PipePasses<CH, END, END> {
    template <bool EMIT, bool RT, class Exec>
    static TSB_HD void run(Exec &, int, typename CH::Smem &, const typename CH::Tables &, const ChainArgs &, bool, long long,
                           long long, int, bool, long long, int) {}
};

// The whole launch for one (segment, lane).  `Exec::each(f)` runs f(tid, regs) for every thread of the lane and
// ends with a barrier: on the device that is { f(threadIdx.x, regs); __syncthreads(); }, in the CPU
// test harness a loop over tid.  Barriers per frame: 2*PASSES + (PV ? 2 : 1).
template <class CH>
TSB_HD SegInfo seg_info(const ChainArgs &a, int lane, int seg) {
    SegInfo sg;
    if (a.seg_tab) {
        const long long *t = a.seg_tab + SEG_COLS * (long long)seg;
        sg.first_frame = TSB_LDG(&t[0]);
        sg.n_frames = TSB_LDG(&t[1]);
        sg.in_off0 = TSB_LDG(&t[2]);
        sg.out_off0 = TSB_LDG(&t[3]);
        sg.warm = (int)TSB_LDG(&t[4]);
        sg.warm_hop = (int)TSB_LDG(&t[5]);
        sg.lead = (int)TSB_LDG(&t[6]);
    } else {
        sg.first_frame = 0;
        sg.n_frames = a.n_frames;
        sg.in_off0 = a.in_off0;
        sg.out_off0 = a.out_off0;
        sg.warm = 0;
        sg.warm_hop = 0;
        sg.lead = 0;
    }
    sg.row = (long long)seg * a.lanes + lane;
    return sg;
}

// PREPASS (ChainArgs::mode == 1) is a compile-time choice: the pre-pass lives in its own kernel, so that its loop
// does not weigh on the register allocation of the render loop.
template <class CH, bool PREPASS, class Exec>
TSB_HD void run_lane(Exec &ex, int lane, int seg, typename CH::Smem &sm, const typename CH::Tables &tb, const ChainArgs &a) {
    using Regs = typename CH::Regs;
    constexpr int N = CH::N;
    static_assert(CH::PASSES >= 2, "frame sizes below 64 are not instantiated");
    // Only what the frame loop needs stays live across it (registers are what limits residency): the segment
    // description is re-read from the argument block / table at the end.
    // Frames are numbered from the segment's first OWN frame: ff < 0 are its lead-in frames (and, in the pre-pass,
    // the warm-up frame before them), so that nothing but the counter tells them apart inside the loop.
    long long in_off, out_pos;
    int own, ff_first;
    const int *hop_tab;
    constexpr bool prepass = PREPASS && CH::IS_PV;
    bool warm;
    bool store_lead;
    int warm_hop;
    int ph = 0;  // hop of the frame before the one being analysed (or a PH_* code)
    long long first_own;
    {
        SegInfo sg = seg_info<CH>(a, lane, seg);
        first_own = sg.first_frame + sg.lead;
        store_lead = sg.first_frame == 0 && sg.lead > 0;
        warm = prepass && sg.warm;
        warm_hop = sg.warm_hop;
        // a warm-up frame is the frame before the segment: it only primes last_phase
        in_off = warm ? sg.in_off0 - sg.warm_hop : sg.in_off0;
        out_pos = sg.out_off0;
        own = (int)sg.n_frames - sg.lead;
        ff_first = -sg.lead - (warm ? 1 : 0);
        hop_tab = a.hop_tab ? a.hop_tab + first_own : nullptr;  // indexed by ff below (the warm-up frame has warm_hop)
        // what the first analysed frame follows (CH::WRAP_FIX, see Chain::PH_ZERO): the warm-up frame of a segment,
        // nothing (a fresh PhaseVocoder), or the caller's previous frame
        if (warm) ph = CH::PH_UNKNOWN;  // the warm-up frame itself: its increments are not used
        else if (sg.warm) ph = sg.warm_hop;
        else if (!a.load_state) ph = CH::PH_ZERO;
        else ph = a.prev_hop > 0 ? a.prev_hop : CH::PH_UNKNOWN;
        ex.each([&](int tid, Regs &rg) {
            CH::load_state(tid, sm, rg, tb, a, sg);
            CH::ring_fill(tid, lane, sm, a, in_off, N);
            CH::ring_wait();
        });
    }
    // PV analysis of the frame at input offset `off`.  fp32: the barrier that ends the phase also tells whether any
    // thread met a doubtful wrap; if so (about one frame in 50 for noisy input) an extra phase re-evaluates those bins
    // exactly before the synthesis reads the increments.
    auto analyze_frame = [&](bool all, long long off, int prev) {
        if constexpr (CH::WRAP_FIX) {
            const bool fix = tb.win64 != nullptr && prev != CH::PH_UNKNOWN;
            if (ex.each_or([&](int tid, Regs &rg) { return CH::phase_pv_analyze(tid, sm, rg, tb, all, fix); }))
                ex.each([&](int tid, Regs &rg) { CH::phase_wrap_fix(tid, lane, sm, rg, tb, a, off, prev); });
        } else {
            ex.each([&](int tid, Regs &rg) { CH::phase_pv_analyze(tid, sm, rg, tb, all, false); });
        }
    };
    if (CH::PIPE && !prepass) {
        // frame f: [analyse f][synthesise f][inverse passes of f  ||  forward passes of f+1]; the forward passes of
        // the first frame run alone as a prologue.  `hop_of(f)` = samples emitted after frame f.
        auto hop_of = [&](int f) { return hop_tab ? TSB_LDG(&hop_tab[f]) : a.hop; };
        if (ff_first < own) {
            int hop0 = hop_of(ff_first);
            int cnt = ff_first + 1 < own ? (hop0 < N ? hop0 : N) : 0;
            PipePasses<CH, 0, CH::PASSES>::template run<false, false>(ex, lane, sm, tb, a, true, in_off, in_off + hop0 + N - cnt, cnt,
                                                                      false, out_pos, 0);
        }
        // the overlap tail in registers for the whole launch (Chain::RT_OK)
        const bool rt = CH::RT_OK && !hop_tab && a.hop == N / 4;
        if (rt) ex.each([&](int tid, Regs &rg) { CH::tail_load(tid, sm, rg, out_pos); });
        // one frame; EMIT is a type so that the lead-in frames get their own (cold) copy of the passes and the
        // steady-state loop carries nothing extra
        auto frame = [&](int ff, auto EMIT, auto RT) {
            const int hop = hop_of(ff);
            if (CH::IS_PV) {
                const bool all = ff + 1 == own;  // last_phase is state: the last frame of a launch analyses every bin
                analyze_frame(all, in_off, ph);
                if (CH::WRAP_FIX) ph = hop;
                ex.each([&](int tid, Regs &rg) {
                    CH::phase_pv_synth(tid, sm, rg, tb);
                    CH::ring_wait();  // next frame's samples (requested a frame ago) are in the ring after this barrier
                });
            } else {
                ex.each([&](int tid, Regs &) {
                    CH::phase_gain(tid, sm, tb);
                    CH::ring_wait();
                });
            }
            const bool have_next = ff + 1 < own;
            const long long next_in = in_off + hop;
            int cnt = 0;
            long long first = 0;
            if (have_next && ff + 2 < own) {  // while frame ff+1 is transformed, fetch what frame ff+2 adds
                int hop1 = hop_of(ff + 1);
                cnt = hop1 < N ? hop1 : N;
                first = next_in + hop1 + N - cnt;
            }
            PipePasses<CH, 0, CH::PASSES>::template run<decltype(EMIT)::value, decltype(RT)::value>(ex, lane, sm, tb, a, have_next, next_in,
                                                                                                   first, cnt, true, out_pos, hop);
            in_off = next_in;
            out_pos += hop;
        };
        int ff = ff_first;
        if (CH::RT_OK && rt) {
            using RTT = std::integral_constant<bool, CH::RT_OK>;
            for (; ff < 0 && ff < own; ff++) frame(ff, std::false_type{}, RTT{});
            for (; ff < own; ff++) frame(ff, std::true_type{}, RTT{});
            ex.each([&](int tid, Regs &rg) { CH::tail_store(tid, sm, rg, out_pos); });
        } else {
            for (; ff < 0 && ff < own; ff++) frame(ff, std::false_type{}, std::false_type{});
            for (; ff < own; ff++) frame(ff, std::true_type{}, std::false_type{});
        }
    } else {
    // (the overlap tail in registers, as above; never in the pre-pass, which writes no samples)
    const bool rt = CH::RT_OK && !CH::PIPE && !prepass && !hop_tab && a.hop == N / 4;
    if (CH::RT_OK && !CH::PIPE && rt) ex.each([&](int tid, Regs &rg) { CH::tail_load(tid, sm, rg, out_pos); });
    bool warm_pending = warm;
    for (int ff = ff_first; ff < own; ff++) {
        const bool is_warm = warm_pending;
        warm_pending = false;
        int hop = is_warm ? warm_hop : (hop_tab ? TSB_LDG(&hop_tab[ff]) : a.hop);
        // slots [in_off, in_off+hop) are dead once pass 0 has read the frame
        int cnt = ff + 1 < own ? (hop < N ? hop : N) : 0;
        long long first = in_off + hop + N - cnt;
        FwdPasses<CH, 0, CH::PASSES>::run(ex, lane, sm, tb, a, in_off, first, cnt);
        if (CH::IS_PV) {
            const bool all = is_warm || ff + 1 == own;
            analyze_frame(all, in_off, ph);
            if (CH::WRAP_FIX) ph = hop;
            if (prepass) {
                if (is_warm) {
                    if (a.st_warm_phase) {
                        const long long row = (long long)seg * a.lanes + lane;
                        ex.each([&](int tid, Regs &rg) { CH::regs_to(tid, a.st_warm_phase, row, rg.last_phase); });
                    }
                } else if (ff >= 0 || store_lead) {  // (lead-in frames: their increments are the previous segment's,
                                                     // unless the segment starts its block: the block's own lead-in)
                    double *inc_row = a.st_inc ? a.st_inc + ((long long)lane * a.inc_frames + first_own + ff) * CH::M : nullptr;
                    ex.each([&](int tid, Regs &rg) { CH::phase_pv_accumulate(tid, sm, rg, tb, inc_row); });
                }
                ex.each([&](int, Regs &) { CH::ring_wait(); });
                in_off += hop;
                continue;
            }
            ex.each([&](int tid, Regs &rg) { CH::phase_pv_synth(tid, sm, rg, tb); });
        } else {
            ex.each([&](int tid, Regs &) { CH::phase_gain(tid, sm, tb); });
        }
        if (CH::RT_OK && !CH::PIPE && rt) {
            constexpr bool RTT = CH::RT_OK && !CH::PIPE;
            if (ff >= 0) InvPasses<CH, 0, CH::PASSES>::template run<true, RTT>(ex, lane, sm, tb, a, out_pos, hop);
            else InvPasses<CH, 0, CH::PASSES>::template run<false, RTT>(ex, lane, sm, tb, a, out_pos, hop);
        } else {
            if (ff >= 0) InvPasses<CH, 0, CH::PASSES>::template run<true, false>(ex, lane, sm, tb, a, out_pos, hop);
            else InvPasses<CH, 0, CH::PASSES>::template run<false, false>(ex, lane, sm, tb, a, out_pos, hop);
        }
        in_off += hop;
        out_pos += hop;
    }
    if (CH::RT_OK && !CH::PIPE && rt) ex.each([&](int tid, Regs &rg) { CH::tail_store(tid, sm, rg, out_pos); });
    }
    {
        SegInfo sg = seg_info<CH>(a, lane, seg);
        ex.each([&](int tid, Regs &rg) { CH::store_state(tid, sm, rg, a, sg, out_pos); });
    }
}

}  // namespace tsb
