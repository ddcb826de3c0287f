// tsb_plan.hpp — host construction of the constant tables of one chain plan (see ChainTables in
// tsb_chain.cuh).  Header-only template over the arithmetic type; used by libtsb200 (which uploads
// the vectors) and by the CPU test harness (which points ChainTables at them directly).
#pragma once
#include <cmath>
#include <complex>
#include <vector>

#include "tsb_chain.cuh"
#include "tsb_host.hpp"

namespace tsb {

struct PlanDesc {
    int n = 0;               // frame size, power of two
    int program = PROG_PV_SHIFT;
    int win_in = WIN_SQRT_HANN, win_out = WIN_SQRT_HANN;
    double hop = 0;          // PhaseVocoder::hop() seen by the frames of this plan
    double rate = 48000;
    double ratio = 1;        // PV: frequency scale of the (linear) shift function
    const double *fn_bins = nullptr;  // PV: fn(bin centre) for j < n/2; nullptr -> centre*ratio
    const double *gain = nullptr;     // GAIN: n/2+1 real gains; nullptr -> all ones
};

// arg() of element 0 of frft(x, +1) run in place on a real frame whose DC sum has the given sign:
// rtlib/fft.hpp:20-39 evaluated for i = 0 (f0 = f2 = unnormalised FFT output, f1 = f3 = f0/sqrt(N)).
inline double dc_phase_const(int n, double sign) {
    using cd = std::complex<double>;
    cd im(0, 1);
    cd im1 = std::pow(im, 1.0), im2 = std::pow(im, 1.0 * 2), im3 = std::pow(im, 1.0 * 3);
    double sqrtn = std::sqrt((double)(size_t)n);
    cd f0(sign * sqrtn, 0.0), f1 = f0 / sqrtn, f2 = f0, f3 = f1;
    cd b0 = f0 + f1 + f2 + f3;
    cd b1 = f0 + im * f1 - f2 - im * f3;
    cd b2 = f0 - f1 + f2 - f3;
    cd b3 = f0 - im * f1 - f2 + im * f3;
    b1 *= im1;
    b2 *= im2;
    b3 *= im3;
    return std::arg((b0 + b1 + b2 + b3) / 4.0);
}

template <typename real>
struct HostTables {
    std::vector<cpx<real>> tw, twn;
    std::vector<real> win_a, win_s, gain;
    std::vector<double> win64;  // fp32 PV: the analysis window in double (ChainTables::win64)
    std::vector<uint32_t> map;
    std::vector<uint16_t> map_src;
    std::vector<int16_t> map1;  // empty unless every destination has at most one source
    std::vector<uint8_t> src_used;  // [M] 1 if some destination takes source bin j
    std::vector<uint16_t> dst_order;  // [M/2] destination pairs, those with a source first; empty = natural order
    double hop, rate, ratio, size_over_hop, dc_pos, dc_neg;
    ChainTables<real> view() const {
        ChainTables<real> t;
        t.tw = tw.data();
        t.twn = twn.data();
        t.win_a = win_a.data();
        t.win_s = win_s.empty() ? win_a.data() : win_s.data();
        t.inv_n = 1.0 / (double)win_a.size();
        t.map = map.data();
        t.map_src = map_src.data();
        t.map1 = map1.empty() ? nullptr : map1.data();
        t.src_used = src_used.data();
        t.dst_order = dst_order.empty() ? nullptr : dst_order.data();
        t.hop = hop;
        t.rate = rate;
        t.ratio = ratio;
        t.size_over_hop = size_over_hop;
        t.inv_rate = 1.0 / rate;
        t.dc_pos = dc_pos;
        t.dc_neg = dc_neg;
        t.gain = gain.data();
        t.win64 = win64.empty() ? nullptr : win64.data();
        return t;
    }
};

template <typename real>
HostTables<real> build_tables(const PlanDesc &d) {
    HostTables<real> h;
    const int n = d.n, m = n / 2;
    std::vector<double> tw = pass_twiddle_table(m), twn = twiddle_table(n, m / 2 + 1);
    h.tw.resize(tw.size() / 2);
    for (size_t q = 0; q < h.tw.size(); q++) h.tw[q] = mk<real>((real)tw[2 * q], (real)tw[2 * q + 1]);
    h.twn.resize(m / 2 + 1);
    for (int q = 0; q <= m / 2; q++) h.twn[q] = mk<real>((real)twn[2 * q], (real)twn[2 * q + 1]);
    std::vector<double> wa = window_table(d.win_in, n), ws = window_table(d.win_out, n);
    const double sqrtn = std::sqrt((double)n);
    // Same window kind on both sides (always so for the PV): one table serves both passes (half the L1 footprint
    // of the windows; the PV's 1/2 is applied to the amplitudes instead, an exact scaling).
    h.win_a.resize(n);
    for (int i = 0; i < n; i++) h.win_a[i] = (real)(wa[i] / sqrtn);
    if (d.win_in != d.win_out) {
        h.win_s.resize(n);
        for (int i = 0; i < n; i++) h.win_s[i] = (real)(ws[i] / sqrtn);
    }
    if (sizeof(real) == 4 && d.program == PROG_PV_SHIFT) h.win64 = wa;
    h.hop = d.hop;
    h.rate = d.rate;
    h.ratio = d.ratio;
    h.size_over_hop = (size_t)n / d.hop;  // size/hop(), rtlib/pv.hpp:44
    h.dc_pos = dc_phase_const(n, 1.0);
    h.dc_neg = dc_phase_const(n, -1.0);
    std::vector<double> fb(m);
    for (int j = 0; j < m; j++) {
        double centre = j / (double)(size_t)n * d.rate;  // rtlib/pv.hpp:61
        fb[j] = d.fn_bins ? d.fn_bins[j] : centre * d.ratio;
    }
    build_shift_map(n, d.rate, fb.data(), h.map, h.map_src);
    h.src_used.assign(m, 0);
    for (uint32_t e : h.map)
        for (uint32_t q = 0; q < (e & 0xffff); q++) h.src_used[h.map_src[(e >> 16) + q]] = 1;
    // Destination pairs (k, M-k) without any source are silent: when there are many (x2 leaves every odd bin
    // empty), hand the pairs out with the sourced ones first — thread t takes entries t, t+T, ... — so that
    // whole slots are silent for every thread and the synthesis skips them.  Pair 0 (bins 0 and M/2) stays first.
    {
        const int half = m / 2;
        auto sourced = [&](int k) { return (h.map[k] & 0xffff) != 0; };
        std::vector<uint16_t> with, without;
        for (int k = 1; k < half; k++) (sourced(k) || sourced(m - k) ? with : without).push_back((uint16_t)k);
        if ((int)without.size() >= half / 4) {
            h.dst_order.push_back(0);
            h.dst_order.insert(h.dst_order.end(), with.begin(), with.end());
            h.dst_order.insert(h.dst_order.end(), without.begin(), without.end());
        }
    }
    bool simple = true;
    for (uint32_t e : h.map) simple = simple && (e & 0xffff) <= 1;
    if (simple) {
        h.map1.resize(m);
        for (int k = 0; k < m; k++) h.map1[k] = (h.map[k] & 0xffff) ? (int16_t)h.map_src[h.map[k] >> 16] : (int16_t)-1;
    }
    h.gain.resize(m + 1);
    for (int k = 0; k <= m; k++) h.gain[k] = (real)(d.gain ? d.gain[k] : 1.0);
    return h;
}

}  // namespace tsb
