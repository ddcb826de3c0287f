// tsb_host.cpp — see tsb_host.hpp.
#include "tsb_host.hpp"

#include <cmath>

namespace tsb {

bool hop_fixed(double hop, uint64_t *hop_samples, uint64_t *hop_fract) {
    if (hop < 0) return false;  // "ignoring invalid hop", rtlib/control.hpp:23
    *hop_samples = (uint64_t)hop;
    *hop_fract = (uint64_t)(std::fmod(hop, 1.0) * double(uint64_t(-1)));
    return true;
}

bool SchedState::set_hop(double h) {
    uint64_t s, f;
    if (!hop_fixed(h, &s, &f)) return false;
    hop = h;
    hop_samples = s;
    hop_fract = f;
    return true;
}

double SchedState::time_secs() const {
    return (double(time_samples) + time_fract / double(uint64_t(-1))) / rate;
}

uint64_t SchedState::advance() {
    if (time_fract + hop_fract < time_fract) computed_hop = hop_samples + 1;
    else computed_hop = hop_samples;
    time_samples += computed_hop;
    time_fract += hop_fract;
    return computed_hop;
}

std::vector<FrameRec> schedule_const(SchedState &st, int64_t in_len, uint32_t n, double hop, int64_t max_frames) {
    std::vector<FrameRec> frames;
    st.new_window_size = n;
    st.set_hop(hop);
    while (max_frames < 0 || (int64_t)frames.size() < max_frames) {
        st.window_size = st.new_window_size;                                        // tinyspec.cpp:135-139
        if (in_len - st.in_pos < (int64_t)(st.computed_hop + st.window_size)) break;  // :142-143
        st.in_pos += (int64_t)st.computed_hop;                                      // :147-149
        FrameRec fr;
        fr.in_off = st.in_pos;
        fr.out_off = st.out_pos;
        fr.win = (int32_t)st.window_size;
        fr.hop_d = st.hop;
        fr.t = st.time_secs();                                                      // :162
        // the callback re-issues next_hop_*(n, hop): no change
        fr.hop = (int32_t)st.advance();                                             // :163-168
        st.out_pos += fr.hop;                                                       // :173-191
        frames.push_back(fr);
    }
    return frames;
}

std::vector<double> window_table(int kind, int n) {
    std::vector<double> w(n, 1.0);
    for (int i = 0; i < n; i++) {
        size_t ii = (size_t)i, size = (size_t)n;
        if (kind == WIN_HANN) w[i] = 0.5 * (1 - cos(2 * M_PI * ii / size));
        else if (kind == WIN_SQRT_HANN) w[i] = sqrt(0.5 * (1 - cos(2 * M_PI * ii / size)));
    }
    return w;
}

void build_shift_map(int n, double rate, const double *fn_bins, std::vector<uint32_t> &map, std::vector<uint16_t> &src) {
    int half = n / 2;
    std::vector<std::vector<uint16_t>> per(half);
    for (int j = 0; j < half; j++) {
        double q = fn_bins[j] / rate * (size_t)n;  // rtlib/pv.hpp:72
        if (!(q > -2147483648.0 && q < 2147483648.0)) continue;  // int conversion would be UB; x86 yields INT_MIN
        int k = (int)q;
        if (k >= 0 && k < half) per[k].push_back((uint16_t)j);
    }
    map.assign(half, 0);
    src.clear();
    for (int k = 0; k < half; k++) {
        map[k] = ((uint32_t)src.size() << 16) | (uint32_t)per[k].size();
        for (uint16_t j : per[k]) src.push_back(j);
    }
    if (src.empty()) src.push_back(0);
}

std::vector<double> twiddle_table(int m, int count) {
    std::vector<double> t(2 * (size_t)count);
    const long double tau = 6.283185307179586476925286766559005768L;
    for (int q = 0; q < count; q++) {
        long double a = -tau * (long double)q / (long double)m;
        t[2 * q] = (double)cosl(a);
        t[2 * q + 1] = (double)sinl(a);
    }
    return t;
}

std::vector<double> pass_twiddle_table(int m) {
    int lg = 0;
    while ((1 << lg) < m) lg++;
    int n8 = lg / 3, rem = lg % 3;
    std::vector<int> radix(n8, 8);
    if (rem) radix.push_back(1 << rem);
    std::vector<double> t;
    const long double tau = 6.283185307179586476925286766559005768L;
    int ns = 1;
    for (size_t p = 0; p < radix.size(); p++) {
        int r = radix[p];
        if (p >= 1) {
            for (int k = 1; k < r; k++)
                for (int jm = 0; jm < ns; jm++) {
                    long double a = -tau * (long double)(jm * k) / (long double)(ns * r);
                    t.push_back((double)cosl(a));
                    t.push_back((double)sinl(a));
                }
        }
        ns *= r;
    }
    if (t.empty()) { t.push_back(1.0); t.push_back(0.0); }
    return t;
}

}  // namespace tsb
