// tests/emul/emul.cpp — TEST INFRASTRUCTURE, CPU only.
//
// Replays the per-thread phases of the CUDA chain kernel (tinyspec-cling_b200/csrc/tsb_chain.cuh)
// thread by thread on the host, so that the kernel's arithmetic and indexing can be checked against
// the oracle in a container without a GPU.  It is compiled from the SAME headers the kernel is
// compiled from.  It is not a CPU fallback: nothing in the product loads this library.
#include <cstring>
#include <memory>
#include <vector>

#include "tsb_chain.cuh"
#include "tsb_plan.hpp"

using namespace tsb;

template <class CH>
struct HostExec {
    std::vector<typename CH::Regs> regs = std::vector<typename CH::Regs>(CH::T);
    template <class F>
    void each(F f) {
        for (int tid = 0; tid < CH::T; tid++) f(tid, regs[tid]);
    }
    template <class F>
    void each_half(F f, int) {  // (the halves are independent between their passes: any order will do)
        for (int tid = 0; tid < CH::T; tid++) f(tid, regs[tid]);
    }
    template <class F>
    bool each_or(F f) {
        bool any = false;
        for (int tid = 0; tid < CH::T; tid++) any = f(tid, regs[tid]) || any;
        return any;
    }
};

template <typename real, int LOG2N, int PROG, bool OUT32>
static int run(const PlanDesc &d, const ChainArgs &a) {
    using CH = Chain<real, LOG2N, PROG, OUT32>;
    HostTables<real> ht = build_tables<real>(d);
    ChainTables<real> tb = ht.view();
    auto sm = std::make_unique<typename CH::Smem>();
    for (int lane = 0; lane < a.lanes; lane++) {
        std::memset(sm.get(), 0, sizeof(typename CH::Smem));
        HostExec<CH> ex;
        if (a.mode == 1) run_lane<CH, true>(ex, lane, 0, *sm, tb, a);
        else run_lane<CH, false>(ex, lane, 0, *sm, tb, a);
    }
    return 0;
}

template <typename real, int PROG, bool OUT32 = false>
static int by_size(int log2n, const PlanDesc &d, const ChainArgs &a) {
    switch (log2n) {
        case 6: return run<real, 6, PROG, OUT32>(d, a);
        case 7: return run<real, 7, PROG, OUT32>(d, a);
        case 8: return run<real, 8, PROG, OUT32>(d, a);
        case 9: return run<real, 9, PROG, OUT32>(d, a);
        case 10: return run<real, 10, PROG, OUT32>(d, a);
        case 11: return run<real, 11, PROG, OUT32>(d, a);
        case 12: return run<real, 12, PROG, OUT32>(d, a);
        case 13: return run<real, 13, PROG, OUT32>(d, a);
    }
    return -1;
}

extern "C" int emul_chain_render(int precision, int log2n, int program, int win_in, int win_out, double hop_d,
                                 double rate, double ratio, const double *fn_bins, const double *gain,
                                 const float *in, long long in_stride, long long in_len, void *out,
                                 long long out_stride, long long out_len, long long n_frames, long long in_off0,
                                 long long out_off0, int hop, const int *hop_tab, double *st_last_phase,
                                 double *st_phase_sum, double *st_ola, int load_state, int lanes, int prev_hop) {
    PlanDesc d;
    d.n = 1 << log2n;
    d.program = program;
    d.win_in = win_in;
    d.win_out = win_out;
    d.hop = hop_d;
    d.rate = rate;
    d.ratio = ratio;
    d.fn_bins = fn_bins;
    d.gain = gain;
    ChainArgs a;
    a.in = in; a.in_stride = in_stride; a.in_len = in_len;
    a.out = out; a.out_stride = out_stride; a.out_len = out_len;
    a.n_frames = n_frames; a.in_off0 = in_off0; a.out_off0 = out_off0;
    a.hop = hop; a.hop_tab = hop_tab;
    a.st_last_phase = st_last_phase; a.st_phase_sum = st_phase_sum; a.st_ola = st_ola;
    a.load_state = load_state; a.lanes = lanes;
    a.n_seg = 0; a.seg_tab = nullptr; a.mode = 0; a.st_warm_phase = nullptr;
    a.st_inc = nullptr; a.inc_frames = 0;
    a.prev_hop = prev_hop;
    if (precision == 0)
        return program == PROG_PV_SHIFT ? by_size<float, PROG_PV_SHIFT>(log2n, d, a) : by_size<float, PROG_GAIN>(log2n, d, a);
    if (precision == 2)
        return program == PROG_PV_SHIFT ? by_size<double, PROG_PV_SHIFT, true>(log2n, d, a)
                                        : by_size<double, PROG_GAIN, true>(log2n, d, a);
    return program == PROG_PV_SHIFT ? by_size<double, PROG_PV_SHIFT>(log2n, d, a) : by_size<double, PROG_GAIN>(log2n, d, a);
}

// Same with the time-segment controls of ChainArgs: one segment, optional warm-up frame, pre-pass mode.
extern "C" int emul_chain_render_ex(int precision, int log2n, int program, int win_in, int win_out, double hop_d,
                                    double rate, double ratio, const double *fn_bins, const double *gain,
                                    const float *in, long long in_stride, long long in_len, void *out,
                                    long long out_stride, long long out_len, long long n_frames, long long in_off0,
                                    long long out_off0, int hop, const int *hop_tab, double *st_last_phase,
                                    double *st_phase_sum, double *st_ola, int load_state, int lanes, int mode, int warm,
                                    int warm_hop, double *st_warm_phase, double *st_inc, int lead) {
    PlanDesc d;
    d.n = 1 << log2n;
    d.program = program;
    d.win_in = win_in;
    d.win_out = win_out;
    d.hop = hop_d;
    d.rate = rate;
    d.ratio = ratio;
    d.fn_bins = fn_bins;
    d.gain = gain;
    long long seg[SEG_COLS] = {0, n_frames, in_off0, out_off0, warm, warm_hop, lead, 0};
    ChainArgs a;
    a.in = in; a.in_stride = in_stride; a.in_len = in_len;
    a.out = out; a.out_stride = out_stride; a.out_len = out_len;
    a.n_frames = n_frames; a.in_off0 = in_off0; a.out_off0 = out_off0;
    a.hop = hop; a.hop_tab = hop_tab;
    a.st_last_phase = st_last_phase; a.st_phase_sum = st_phase_sum; a.st_ola = st_ola;
    a.load_state = load_state; a.lanes = lanes;
    a.n_seg = 1; a.seg_tab = seg; a.mode = mode; a.st_warm_phase = st_warm_phase;
    a.st_inc = st_inc; a.inc_frames = n_frames;
    a.prev_hop = 0;
    if (precision == 0)
        return program == PROG_PV_SHIFT ? by_size<float, PROG_PV_SHIFT>(log2n, d, a) : by_size<float, PROG_GAIN>(log2n, d, a);
    if (precision == 2)
        return program == PROG_PV_SHIFT ? by_size<double, PROG_PV_SHIFT, true>(log2n, d, a)
                                        : by_size<double, PROG_GAIN, true>(log2n, d, a);
    return program == PROG_PV_SHIFT ? by_size<double, PROG_PV_SHIFT>(log2n, d, a) : by_size<double, PROG_GAIN>(log2n, d, a);
}

// M-point complex forward FFT through the kernel's Stockham passes (for unit tests of tsb_math.cuh)
template <typename real, int LOG2M>
static void fft_only(const double *in, double *out) {
    constexpr int M = 1 << LOG2M;
    constexpr int T = 32;
    std::vector<double> t = pass_twiddle_table(M);
    std::vector<cpx<real>> a(M), b(M), tw(t.size() / 2);
    for (size_t q = 0; q < tw.size(); q++) tw[q] = mk<real>((real)t[2 * q], (real)t[2 * q + 1]);
    for (int q = 0; q < M; q++) a[q] = mk<real>((real)in[2 * q], (real)in[2 * q + 1]);
    cpx<real> *src = a.data(), *dst = b.data();
    auto pass = [&](auto P) {
        for (int tid = 0; tid < T; tid++) stockham_pass<real, LOG2M, T, decltype(P)::value>(tid, src, dst, tw.data());
        std::swap(src, dst);
    };
    using RP = RadixPlan<LOG2M>;
    if (RP::passes > 0) pass(std::integral_constant<int, 0>{});
    if (RP::passes > 1) pass(std::integral_constant<int, 1>{});
    if (RP::passes > 2) pass(std::integral_constant<int, 2>{});
    if (RP::passes > 3) pass(std::integral_constant<int, 3>{});
    if (RP::passes > 4) pass(std::integral_constant<int, 4>{});
    for (int q = 0; q < M; q++) {
        out[2 * q] = src[q].x;
        out[2 * q + 1] = src[q].y;
    }
}
extern "C" int emul_fft(int precision, int log2m, const double *in, double *out) {
#define CASE(L)                                                     \
    case L:                                                         \
        if (precision == 0) fft_only<float, L>(in, out);           \
        else fft_only<double, L>(in, out);                          \
        return 0;
    switch (log2m) {
        CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
    }
    return -1;
}

// The two radix-16 passes of the spectral-synthesis kernels' 256-point sub-transforms (tsb_math.cuh: r16_pass), run by
// sixteen "threads" over the tile layout r16_sw.  On the device a thread reads its sixteen points, the half-warp
// synchronises, and the points are written back into the same tile; here every thread's reads precede every write
// because each pass goes from one buffer into another.  half != 0: the last pass stores X[0..128] only (kernel B).
extern "C" void emul_fft256_r16(const double *in, double *out, int half) {
    using cd = cpx<double>;
    std::vector<cd> a(256), b(256), c(256);
    for (int i = 0; i < 256; i++) {
        a[r16_sw(i)] = mk<double>(in[2 * i], in[2 * i + 1]);
        c[i] = mk<double>(0.0, 0.0);
    }
    const double tau = 6.283185307179586476925286766559;
    for (int j = 0; j < 16; j++)
        r16_pass<0>(j, mk<double>(1.0, 0.0), [&](int i) { return a[r16_sw(i)]; }, [] {}, [&](int i, cd v) { b[r16_sw(i)] = v; });
    for (int j = 0; j < 16; j++) {
        const cd w = mk<double>(cos(tau * j / 256.0), -sin(tau * j / 256.0));  // W256^j
        auto rd = [&](int i) { return b[r16_sw(i)]; };
        auto wr = [&](int i, cd v) { c[r16_sw(i)] = v; };
        if (half) r16_pass<1, true>(j, w, rd, [] {}, wr);
        else r16_pass<1>(j, w, rd, [] {}, wr);
    }
    for (int i = 0; i < 256; i++) {
        out[2 * i] = c[r16_sw(i)].x;
        out[2 * i + 1] = c[r16_sw(i)].y;
    }
}

// The branch-free fp64 kernels of the PV stages (host build of the same source), element-wise over arrays.
extern "C" void emul_pv_math(long long n, const double *y, const double *x, double *atan2_out, double *sqrt_out,
                             int *wrap_out, double *sin_out, double *cos_out) {
    for (long long i = 0; i < n; i++) {
        atan2_out[i] = fast_atan2(y[i], x[i]);
        sqrt_out[i] = sqrt_pos(fabs(x[i]));
        wrap_out[i] = wrap_quotient(x[i]);
        fast_sincos(y[i], sin_out[i], cos_out[i]);
    }
}

// fp32 PV: doubtful wraps re-evaluated in double precision since the last call / of them decided the other way
extern "C" void emul_wrap_counters(long long *fixed, long long *changed) {
    long long *c = host_wrap_counters();
    *fixed = c[0];
    *changed = c[1];
    c[0] = c[1] = 0;
}
