// tsb_host.hpp — host-side (CPU, no CUDA) logic of the backend: the frame schedule, the hop
// fixed-point clock, the shift map and the constant tables the kernels read.  Plain C++; shared by
// libtsb200 and by the CPU test harness under tests/emul.
#pragma once
#include <cstdint>
#include <vector>

namespace tsb {

enum WindowKind { WIN_NONE = 0, WIN_HANN = 1, WIN_SQRT_HANN = 2 };

// rtlib/control.hpp:22-29 — decomposition of a (possibly fractional) hop. false = rejected (< 0).
bool hop_fixed(double hop, uint64_t *hop_samples, uint64_t *hop_fract);

// The reference scheduler's clock state (namespace internals, tinyspec.cpp:35-47).
struct SchedState {
    uint64_t window_size = 0;
    uint64_t new_window_size = 1 << 12;  // tinyspec.cpp:38
    uint64_t hop_samples = 1 << 11;      // :39
    uint64_t hop_fract = 0;              // :40
    uint64_t computed_hop = 1 << 11;     // :41
    double hop = 2048.0;                 // :42
    uint64_t time_samples = 0, time_fract = 0;  // :45-46
    double rate = 44100;                 // :57
    // not in the reference: absolute positions, so that a batched render can address the streams
    int64_t in_pos = 0;                  // input samples consumed so far (start of the NEXT advance)
    int64_t out_pos = 0;                 // output samples emitted so far
    bool set_hop(double h);              // rtlib/control.hpp:22-29
    double time_secs() const;            // tinyspec.cpp:110-112
    // tinyspec.cpp:163-168: decide this frame's advance from the fractional clock, then move it
    uint64_t advance();
};

struct FrameRec {
    int64_t in_off;   // I_f: first input sample of the frame
    int64_t out_off;  // P_f: output position where the frame is overlap-added
    int32_t win;      // W_f
    int32_t hop;      // computed_hop_f: samples emitted after this frame
    double hop_d;     // internals::hop as seen by the frame's callback
    double t;         // t_f handed to the callback
};

// Constant-program schedule: every frame asks for the same (n, hop) via next_hop_*(n, hop), set
// before the first frame.  Restates tinyspec.cpp:135-149,161-168 for an input of
// `in_len` samples (max_frames < 0: until starved).  `st` carries on, so consecutive calls continue.
std::vector<FrameRec> schedule_const(SchedState &st, int64_t in_len, uint32_t n, double hop, int64_t max_frames);

// Window tables evaluated with the reference's expressions (rtlib/windowfn.hpp:3,10), periodic.
std::vector<double> window_table(int kind, int n);

// Shift map (rtlib/pv.hpp:59-61,70-78): fn_bins[j] = fn(j/(double)n*rate), j < n/2.
// For every destination bin k < n/2: map[k] = (first << 16) | count into `src` (ascending sources).
void build_shift_map(int n, double rate, const double *fn_bins, std::vector<uint32_t> &map, std::vector<uint16_t> &src);

// exp(-2*pi*i*q/m), q < count, evaluated in long double (re,im interleaved)
std::vector<double> twiddle_table(int m, int count);

// Per-pass twiddle tables of an m-point Stockham FFT with the radix plan of tsb_math.cuh (radix 8 passes,
// remainder 2 or 4 last): for each pass p >= 1, entries [(t-1)*Ns + jm] = exp(-2*pi*i*jm*t/(Ns*R)), t >= 1; passes
// concatenated (re,im interleaved).  Matches TwLayout<log2 m>.
std::vector<double> pass_twiddle_table(int m);


}  // namespace tsb
