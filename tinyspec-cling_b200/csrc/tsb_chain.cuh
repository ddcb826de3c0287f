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
// (tsb_kernels.cu) calls them with __syncthreads() in between, and the CPU test harness
// (tests/emul) replays the same phases thread by thread.  Shared-memory layout per lane:
//     ring[N]  float   input samples, slot = absolute sample index mod N
//     ola[N]   real    overlap-add accumulator, slot = absolute output position mod N
//     bufA[M], bufB[M] complex<real>, ping-pong FFT buffers (M = N/2); the idle one doubles as the
//                      phase-increment exchange array of the PV stage
//     amp[M]   real    PV amplitude exchange array (fp32 PV only; fp64 keeps it in the idle buffer)
//
// Real-FFT mapping: the reference runs complex FFTs of real data.  Here the N
// real samples are packed into M = N/2 complex points, transformed once, and split; with
// X = R2C(x) (kernel e^-), frft(x,+1)[k] = conj(X[k])/sqrt(N), and Re(frft(Y,-1))[n] =
// C2R(Z)[n]/sqrt(N) with Z[k] = (conj(Y[k]) + Y[N-k])/2.  1/sqrt(N) (and the 1/2 of the PV case,
// where Y[k>=N/2] = 0 after shift) are folded into the window tables.
#pragma once
#include "tsb_math.cuh"

namespace tsb {

enum Program { PROG_PV_SHIFT = 0, PROG_GAIN = 1 };

// ---- IEEE single-rounding helpers: never contracted into FMAs, so that the fp64 phase-vocoder
// arithmetic reproduces the reference's expression-by-expression rounding (x86-64, no FMA).
#if defined(__CUDA_ARCH__)
TSB_D double mul_rn(double a, double b) { return __dmul_rn(a, b); }
TSB_D double add_rn(double a, double b) { return __dadd_rn(a, b); }
TSB_D double sub_rn(double a, double b) { return __dsub_rn(a, b); }
TSB_D int d2i_rz(double a) { return __double2int_rz(a); }
TSB_D long long d2ll_rn(double a) { return __double2ll_rn(a); }
#else
inline double mul_rn(double a, double b) { return a * b; }  // host harness is built with -ffp-contract=off
inline double add_rn(double a, double b) { return a + b; }
inline double sub_rn(double a, double b) { return a - b; }
inline int d2i_rz(double a) { return (int)a; }
inline long long d2ll_rn(double a) { return llrint(a); }
#endif

#define TSB_PI 3.14159265358979323846 /* M_PI */

// Tables built once per plan on the host (tsb_host.cpp) and kept in device memory.
template <typename real>
struct ChainTables {
    const cpx<real> *tw;     // [M]      exp(-2*pi*i*q/M)
    const cpx<real> *twn;    // [M/2+1]  exp(-2*pi*i*k/N)
    const real *win_a;       // [N]      analysis window / sqrt(N)
    const real *win_s;       // [N]      synthesis window * c / sqrt(N), c = 1/2 (PV) or 1 (gain)
    // PROG_PV_SHIFT
    const double *expect;    // [M]      j*2.0*M_PI*hop/N, evaluated as rtlib/pv.hpp:40 does
    const uint32_t *map;     // [M]      (first << 16) | count into map_src: sources of destination bin k
    const uint16_t *map_src; // source bins, ascending within each destination (rtlib/pv.hpp:70-78)
    double hop;              // PhaseVocoder::hop() (rtinc/pv.hpp:50-53), may be fractional
    double rate;             // RATE
    double ratio;            // fn(x) = ratio*x applied to the frequencies (rtlib/pv.hpp:66)
    double size_over_hop;    // size/hop()
    double dc_pos, dc_neg;   // arg() the reference obtains for a positive / negative DC bin (see dc_phase())
    // PROG_GAIN
    const real *gain;        // [M+1]    real gain applied to bins k and (N-k)%N
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
};

template <typename real>
struct PvMath;

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
                               double &amp, double &inc) {
        double re = X.x, im = -X.y;  // frft(.,+1) of real data = conj(R2C)/sqrt(N)
        amp = sqrt(add_rn(mul_rn(re, re), mul_rn(im, im)));  // abs(), pv.hpp:20
        double phase = j == 0 ? dc_phase(re, tb) : atan2(im, re);  // arg(), pv.hpp:21
        double diff = sub_rn(phase, last_phase);
        last_phase = phase;                                  // pv.hpp:23
        // phase_to_frequency, pv.hpp:39-45
        double delta = sub_rn(diff, TSB_LDG(&tb.expect[j]));
        int qpd = d2i_rz(delta / TSB_PI);
        if (qpd >= 0) qpd += qpd & 1;
        else qpd -= qpd & 1;
        double t = sub_rn(delta, mul_rn(TSB_PI, (double)qpd));
        t = mul_rn(tb.size_over_hop, t);
        t = mul_rn(t, 0.5);  // "/2.0" is exact either way
        t = t / TSB_PI;
        t = add_rn((double)j, t);
        t = mul_rn(t, tb.rate);
        double freq = t / (double)n;
        freq = mul_rn(freq, tb.ratio);  // shift: freqs[j] = fn(freq), pv.hpp:66
        // frequency_to_phase, pv.hpp:46-48: 2.0*M_PI*freq/RATE*hop()
        inc = mul_rn(mul_rn(2.0 * TSB_PI, freq) / tb.rate, tb.hop);
    }
    static TSB_HD cpx<double> synth(double amp, double inc, double &phase_sum) {
        phase_sum = add_rn(phase_sum, inc);  // pv.hpp:34 — never wrapped
        double s, c;
#if defined(__CUDA_ARCH__)
        sincos(phase_sum, &s, &c);
#else
        s = sin(phase_sum);
        c = cos(phase_sum);
#endif
        return mk<double>(mul_rn(amp, c), mul_rn(amp, s));  // std::polar, pv.hpp:35
    }
};

// ---- fp32: fp32 spectra / amplitudes / transcendental kernels; every phase quantity that
// accumulates (difference, wrap, increment, running sum, range reduction) stays fp64
// (SURVEY.md §7 "Twiddles").  inc = ratio*(expect_j + wrapped_delta), which is what
// pv.hpp:39-48 + the linear shift reduce to algebraically.
template <>
struct PvMath<float> {
    static TSB_HD void analyze(int j, cpx<float> X, float &last_phase, const ChainTables<float> &tb, int n,
                               float &amp, double &inc) {
        (void)n;
        float re = X.x, im = -X.y;
        amp = sqrtf(re * re + im * im);
        float phase = atan2f(im, re);
        double diff = (double)phase - (double)last_phase;
        if (j == 0) {  // DC: keep the reference's (double) values, stored as a sign in the float state
            double lp = last_phase > 0.f ? tb.dc_pos : (last_phase < 0.f ? tb.dc_neg : 0.0);
            double ph = dc_phase((double)re, tb);
            diff = ph - lp;
            phase = (float)ph;
        }
        last_phase = phase;
        double expect = TSB_LDG(&tb.expect[j]);
        double delta = diff - expect;
        int qpd = d2i_rz(delta * (1.0 / TSB_PI));
        if (qpd >= 0) qpd += qpd & 1;
        else qpd -= qpd & 1;
        double wrapped = fma(-TSB_PI, (double)qpd, delta);
        inc = tb.ratio * (expect + wrapped);
    }
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
        int qi = (int)(q & 3);
        float ss = (qi & 1) ? c : s;
        float cc = (qi & 1) ? s : c;
        if (qi == 1 || qi == 2) cc = -cc;
        if (qi >= 2) ss = -ss;
        return mk<float>(amp * cc, amp * ss);
    }
};

template <typename real, int LOG2N, int PROG>
struct Chain {
    static constexpr int N = 1 << LOG2N;
    static constexpr int LOG2M = LOG2N - 1;
    static constexpr int M = N / 2;
    static constexpr int T = (N / 16) < 32 ? 32 : (N / 16);  // threads per lane
    static constexpr int HALF = M / 2;                       // split pairs (k, M-k), k < M/2
    static constexpr int S = (HALF + T - 1) / T;             // pair slots per thread
    static constexpr int PASSES = RadixPlan<LOG2M>::passes;
    static constexpr bool IS_PV = PROG == PROG_PV_SHIFT;
    using C = cpx<real>;
    using Tables = ChainTables<real>;

    struct Smem {
        float ring[N];
        real ola[N];
        C bufA[M];
        C bufB[M];
        // PV amplitude exchange array: fp32 needs its own (the spare FFT buffer is filled by the M fp64
        // increments); in fp64 the spare buffer (16*M bytes) holds both, so nothing extra is reserved.
        real amp_own[(PROG == PROG_PV_SHIFT && sizeof(real) == 4) ? M : 1];
    };
    // Per-thread state that lives in registers for the whole launch.
    struct Regs {
        real last_phase[2 * S];
        double phase_sum[2 * S];
    };

    // where the forward transform of data that started in bufA ends up
    static TSB_HD C *fwd_result(Smem &sm) { return (PASSES & 1) ? sm.bufB : sm.bufA; }
    static TSB_HD C *fwd_spare(Smem &sm) { return (PASSES & 1) ? sm.bufA : sm.bufB; }
    // PV exchange arrays (between the analysis and the synthesis phase)
    static TSB_HD double *xch_inc(Smem &sm) { return reinterpret_cast<double *>(fwd_spare(sm)); }
    static TSB_HD real *xch_amp(Smem &sm) {
        return sizeof(real) == 4 ? sm.amp_own : reinterpret_cast<real *>(xch_inc(sm) + M);
    }

    // ------------------------------------------------------------------ state in / out
    static TSB_HD void load_state(int tid, int lane, Smem &sm, Regs &rg, const ChainArgs &a) {
        for (int i = 0; i < 2 * S; i++) {
            rg.last_phase[i] = 0;
            rg.phase_sum[i] = 0;
        }
        bool ld = a.load_state != 0;
        if (PROG == PROG_PV_SHIFT && ld && a.st_last_phase && a.st_phase_sum) {
            for (int i = 0; i < S; i++) {
                int k = tid + T * i;
                if (k >= HALF) break;
                int kb = k == 0 ? HALF : M - k;
                rg.last_phase[2 * i] = (real)a.st_last_phase[(long long)lane * M + k];
                rg.last_phase[2 * i + 1] = (real)a.st_last_phase[(long long)lane * M + kb];
                rg.phase_sum[2 * i] = a.st_phase_sum[(long long)lane * M + k];
                rg.phase_sum[2 * i + 1] = a.st_phase_sum[(long long)lane * M + kb];
            }
        }
        for (int i = tid; i < N; i += T) {
            real v = 0;
            if (ld && a.st_ola) v = (real)a.st_ola[(long long)lane * N + i];
            sm.ola[(int)((a.out_off0 + i) & (N - 1))] = v;
        }
    }
    static TSB_HD void store_state(int tid, int lane, Smem &sm, Regs &rg, const ChainArgs &a, long long out_pos) {
        if (PROG == PROG_PV_SHIFT && a.st_last_phase && a.st_phase_sum) {
            for (int i = 0; i < S; i++) {
                int k = tid + T * i;
                if (k >= HALF) break;
                int kb = k == 0 ? HALF : M - k;
                a.st_last_phase[(long long)lane * M + k] = (double)rg.last_phase[2 * i];
                a.st_last_phase[(long long)lane * M + kb] = (double)rg.last_phase[2 * i + 1];
                a.st_phase_sum[(long long)lane * M + k] = rg.phase_sum[2 * i];
                a.st_phase_sum[(long long)lane * M + kb] = rg.phase_sum[2 * i + 1];
            }
        }
        if (a.st_ola)
            for (int i = tid; i < N; i += T)
                a.st_ola[(long long)lane * N + i] = (double)sm.ola[(int)((out_pos + i) & (N - 1))];
    }

    // ------------------------------------------------------------------ input ring
    // Bring absolute samples [first, first+count) of this lane into the ring (zeros past in_len).
    // On the device this is an asynchronous global->shared copy (cp.async / LDGSTS): it is issued
    // right after the window phase has consumed the slots it overwrites and completes in the
    // background while the frame is transformed; ring_wait() + a barrier precede the next use.
    static TSB_HD void ring_fill(int tid, int lane, Smem &sm, const ChainArgs &a, long long first, int count) {
        const float *src = a.in + (long long)lane * a.in_stride;
#if defined(__CUDA_ARCH__)
        bool vec = ((first | (long long)count | a.in_stride) & 3) == 0 && first >= 0 && first + count <= a.in_len &&
                   ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
        if (vec) {
            for (int i = tid * 4; i < count; i += T * 4) {
                long long s0 = first + i;
                unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.ring[(int)(s0 & (N - 1))]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + s0) : "memory");
            }
            return;
        }
        for (int i = tid; i < count; i += T) {
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
        for (int i = tid; i < count; i += T) {
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

    // ------------------------------------------------------------------ phase A: window + pack
    // z[m] = x[2m]*wa[2m] + i*x[2m+1]*wa[2m+1]  -> bufA        (rtlib/windowfn.hpp, fused into the load)
    static TSB_HD void phase_window_pack(int tid, Smem &sm, const Tables &tb, long long in_off) {
        for (int m = tid; m < M; m += T) {
            float x0 = sm.ring[(int)((in_off + 2 * m) & (N - 1))];
            float x1 = sm.ring[(int)((in_off + 2 * m + 1) & (N - 1))];
            real w0 = TSB_LDG(&tb.win_a[2 * m]), w1 = TSB_LDG(&tb.win_a[2 * m + 1]);
            sm.bufA[m] = mk<real>((real)x0 * w0, (real)x1 * w1);
        }
    }

    // ------------------------------------------------------------------ FFT passes (ping-pong)
    template <int P>
    static TSB_HD void phase_fft_pass(int tid, const C *src, C *dst, const Tables &tb) {
        stockham_pass<real, LOG2M, T, P>(tid, src, dst, tb.tw);
    }

    // spectrum of the real frame at bins k and kb from the packed transform Zc (see file header)
    static TSB_HD void split_pair(const C *Zc, const Tables &tb, int k, C &Xa, C &Xb) {
        if (k != 0) {
            C a = Zc[k], b = conj(Zc[M - k]);
            C e = mk<real>((real)0.5 * (a.x + b.x), (real)0.5 * (a.y + b.y));
            C d = a - b;
            C o = mk<real>((real)0.5 * d.y, (real)-0.5 * d.x);  // -i*d/2
            C wo = cmul(TSB_LDG(&tb.twn[k]), o);
            Xa = e + wo;           // X[k]
            Xb = conj(e - wo);     // X[M-k]
        } else {
            C z0 = Zc[0], zh = Zc[HALF];
            Xa = mk<real>(z0.x + z0.y, 0);  // X[0]
            Xb = conj(zh);                  // X[M/2]
        }
    }
    // C2R pre-processing for the pair (k, M-k): Hermitian-half values Zk, Zp -> conj(G) entries
    static TSB_HD void merge_pair(C *H, const Tables &tb, int k, C Zk, C Zp) {
        C ee = Zk + conj(Zp);
        C d = Zk - conj(Zp);
        C oo = cmul(d, conj(TSB_LDG(&tb.twn[k])));
        H[k] = mk<real>(ee.x - oo.y, -ee.y - oo.x);
        H[M - k] = mk<real>(ee.x + oo.y, ee.y - oo.x);
    }

    // ------------------------------------------------------------------ PV phase C: split + analysis
    // (rtlib/pv.hpp:17-28 over the live bins 0..N/2-1; results go to the exchange arrays)
    static TSB_HD void phase_pv_analyze(int tid, Smem &sm, Regs &rg, const Tables &tb) {
        const C *Zc = fwd_result(sm);
        double *inc_x = xch_inc(sm);
        real *amp_x = xch_amp(sm);
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = tid + T * i;
            if (HALF % T != 0 && k >= HALF) break;
            int kb = k == 0 ? HALF : M - k;
            C Xa, Xb;
            split_pair(Zc, tb, k, Xa, Xb);
            real amp;
            double inc;
            PvMath<real>::analyze(k, Xa, rg.last_phase[2 * i], tb, N, amp, inc);
            amp_x[k] = amp;
            inc_x[k] = inc;
            PvMath<real>::analyze(kb, Xb, rg.last_phase[2 * i + 1], tb, N, amp, inc);
            amp_x[kb] = amp;
            inc_x[kb] = inc;
        }
    }
    // shift gather for destination bin k (rtlib/pv.hpp:70-78: amplitudes add in ascending source
    // order, the frequency of the last source wins; bins with no source are silent with freq 0)
    static TSB_HD void gather_bin(const real *amp_x, const double *inc_x, const Tables &tb, int k, real &amp, double &inc) {
        uint32_t e = TSB_LDG(&tb.map[k]);
        int first = (int)(e >> 16), cnt = (int)(e & 0xffff);
        amp = 0;
        inc = 0;
        for (int s = 0; s < cnt; s++) {
            int j = TSB_LDG(&tb.map_src[first + s]);
            amp = amp + amp_x[j];
            inc = inc_x[j];
        }
    }
    // ------------------------------------------------------------------ PV phase D: synthesis + C2R pre
    // (rtlib/pv.hpp:29-38 on the shifted bins, then Z = conj(Y) [1/2 folded into win_s])
    static TSB_HD void phase_pv_synth(int tid, Smem &sm, Regs &rg, const Tables &tb) {
        C *H = fwd_result(sm);  // the packed spectrum has been consumed; reuse its buffer
        const double *inc_x = xch_inc(sm);
        const real *amp_x = xch_amp(sm);
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = tid + T * i;
            if (HALF % T != 0 && k >= HALF) break;
            int kb = k == 0 ? HALF : M - k;
            real amp;
            double inc;
            gather_bin(amp_x, inc_x, tb, k, amp, inc);
            C Ya = PvMath<real>::synth(amp, inc, rg.phase_sum[2 * i]);
            gather_bin(amp_x, inc_x, tb, kb, amp, inc);
            C Yb = PvMath<real>::synth(amp, inc, rg.phase_sum[2 * i + 1]);
            if (k != 0) {
                merge_pair(H, tb, k, conj(Ya), conj(Yb));
            } else {
                real z0 = (real)2 * Ya.x;  // Z[0] = Re Y[0] (not halved), Z[M] = 0
                H[0] = mk<real>(z0, -z0);
                H[HALF] = mk<real>((real)2 * Yb.x, (real)-2 * Yb.y);  // 2*Z[M/2], Z = conj(Y)
            }
        }
    }
    // ------------------------------------------------------------------ gain program: split * g -> C2R pre
    static TSB_HD void phase_gain(int tid, Smem &sm, const Tables &tb) {
        const C *Zc = fwd_result(sm);
        C *H = fwd_spare(sm);
#pragma unroll
        for (int i = 0; i < S; i++) {
            int k = tid + T * i;
            if (HALF % T != 0 && k >= HALF) break;
            C Xa, Xb;
            if (k != 0) {
                split_pair(Zc, tb, k, Xa, Xb);
                real ga = TSB_LDG(&tb.gain[k]), gb = TSB_LDG(&tb.gain[M - k]);
                merge_pair(H, tb, k, mk<real>(ga * Xa.x, ga * Xa.y), mk<real>(gb * Xb.x, gb * Xb.y));
            } else {
                C z0 = Zc[0], zh = Zc[HALF];
                real x0 = (z0.x + z0.y) * TSB_LDG(&tb.gain[0]);  // X[0]
                real xm = (z0.x - z0.y) * TSB_LDG(&tb.gain[M]);  // X[N/2]
                real gh = TSB_LDG(&tb.gain[HALF]);
                H[0] = mk<real>(x0 + xm, -(x0 - xm));
                H[HALF] = mk<real>((real)2 * gh * zh.x, (real)-2 * gh * zh.y);  // 2*Z[M/2], Z = g*conj(zh)
            }
        }
    }
    // buffers of the inverse transform
    static TSB_HD C *inv_src(Smem &sm) { return PROG == PROG_PV_SHIFT ? fwd_result(sm) : fwd_spare(sm); }
    static TSB_HD C *inv_dst(Smem &sm) { return PROG == PROG_PV_SHIFT ? fwd_spare(sm) : fwd_result(sm); }
    static TSB_HD C *inv_result(Smem &sm) { return (PASSES & 1) ? inv_dst(sm) : inv_src(sm); }

    // ------------------------------------------------------------------ phase F: window + overlap-add + emit
    // x[2m] = Re F[m], x[2m+1] = -Im F[m]; out = x*ws; accumulate into the ring; the first `hop`
    // positions are complete (tinyspec.cpp:173-182) and leave the chip, the rest is the tail (194-200).
    static TSB_HD void phase_ola_emit(int tid, int lane, Smem &sm, const Tables &tb, const ChainArgs &a,
                                      long long out_pos, int hop) {
        const C *F = inv_result(sm);
        real *dst = reinterpret_cast<real *>(a.out) + (long long)lane * a.out_stride;
        for (int m = tid; m < M; m += T) {
            C f = F[m];
            real v[2] = {f.x * TSB_LDG(&tb.win_s[2 * m]), -f.y * TSB_LDG(&tb.win_s[2 * m + 1])};
#pragma unroll
            for (int u = 0; u < 2; u++) {
                int n = 2 * m + u;
                int slot = (int)((out_pos + n) & (N - 1));
                real acc = sm.ola[slot] + v[u];
                if (n < hop) {
                    long long p = out_pos + n;
                    if (p < a.out_len) dst[p] = acc;
                    sm.ola[slot] = 0;
                } else {
                    sm.ola[slot] = acc;
                }
            }
        }
        // hop larger than the frame: nothing overlaps those positions (tinyspec.cpp:184-191)
        for (int n = N + tid; n < hop; n += T) {
            long long p = out_pos + n;
            if (p < a.out_len) dst[p] = 0;
        }
    }
};

// ---- compile-time pass loops ----
template <class CH, int P, int END>
struct FwdPasses {
    template <class Exec>
    static TSB_HD void run(Exec &ex, typename CH::Smem &sm, const typename CH::Tables &tb, typename CH::C *src,
                           typename CH::C *dst) {
        ex.each([&](int tid, typename CH::Regs &) { CH::template phase_fft_pass<P>(tid, src, dst, tb); });
        FwdPasses<CH, P + 1, END>::run(ex, sm, tb, dst, src);
    }
};
template <class CH, int END>
struct FwdPasses<CH, END, END> {
    template <class Exec>
    static TSB_HD void run(Exec &, typename CH::Smem &, const typename CH::Tables &, typename CH::C *, typename CH::C *) {}
};

// The whole launch for one lane.  `Exec::each(f)` runs f(tid, regs) for every thread of the lane and
// ends with a barrier: on the device that is { f(threadIdx.x, regs); __syncthreads(); }, in the CPU
// test harness a loop over tid.
template <class CH, class Exec>
TSB_HD void run_lane(Exec &ex, int lane, typename CH::Smem &sm, const typename CH::Tables &tb, const ChainArgs &a) {
    using Regs = typename CH::Regs;
    constexpr int N = CH::N;
    long long in_off = a.in_off0, out_pos = a.out_off0;
    ex.each([&](int tid, Regs &rg) {
        CH::load_state(tid, lane, sm, rg, a);
        CH::ring_fill(tid, lane, sm, a, in_off, N);
        CH::ring_wait();
    });
    for (long long f = 0; f < a.n_frames; f++) {
        int hop = a.hop_tab ? TSB_LDG(&a.hop_tab[f]) : a.hop;
        ex.each([&](int tid, Regs &) { CH::phase_window_pack(tid, sm, tb, in_off); });
        // slots [in_off, in_off+hop) are dead now: start fetching the next frame's new samples
        {
            int cnt = hop < N ? hop : N;
            long long first = in_off + hop + N - cnt;
            if (f + 1 < a.n_frames)
                ex.each_nosync([&](int tid, Regs &) { CH::ring_fill(tid, lane, sm, a, first, cnt); });
        }
        FwdPasses<CH, 0, CH::PASSES>::run(ex, sm, tb, sm.bufA, sm.bufB);
        if (CH::IS_PV) {
            ex.each([&](int tid, Regs &rg) { CH::phase_pv_analyze(tid, sm, rg, tb); });
            ex.each([&](int tid, Regs &rg) { CH::phase_pv_synth(tid, sm, rg, tb); });
        } else {
            ex.each([&](int tid, Regs &) { CH::phase_gain(tid, sm, tb); });
        }
        FwdPasses<CH, 0, CH::PASSES>::run(ex, sm, tb, CH::inv_src(sm), CH::inv_dst(sm));
        ex.each([&](int tid, Regs &) {
            CH::phase_ola_emit(tid, lane, sm, tb, a, out_pos, hop);
            CH::ring_wait();
        });
        in_off += hop;
        out_pos += hop;
    }
    ex.each([&](int tid, Regs &rg) { CH::store_state(tid, lane, sm, rg, a, out_pos); });
}

}  // namespace tsb
