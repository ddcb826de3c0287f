// rtinc/control.hpp — the control API a hack calls (reference: rtinc/control.hpp:1-7, rtlib/control.hpp:22-82,
// tinyspec.cpp:35-61,110-112) over the same process-wide state.  Like the reference's rtlib/, this header
// DEFINES that state: include it in one translation unit (the render driver).
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <string>

#include "buffer.hpp"

using func_t = std::function<void(WaveBuf &, WaveBuf &, double)>;
namespace internals {  // reference: tinyspec.cpp:35-61
uint64_t window_size = 0;
uint64_t new_window_size(1 << 12);
uint64_t hop_samples(new_window_size / 2);
uint64_t hop_fract = 0;
uint64_t computed_hop = hop_samples;
double hop = double(hop_samples);
uint64_t time_samples = 0;
uint64_t time_fract = 0;
size_t nch_in = 0, nch_out = 0, new_nch_in = 0, new_nch_out = 0;
func_t fptr(nullptr);
std::string client_name = "tinyspec_b200";
double rate = 44100;  // the offline driver sets it before the hack runs, as JACK's callback does (tinyspec.cpp:240-244)
// offline stand-in for the input queue: samples consumed so far and samples still available
int64_t in_pos = 0, in_len = 0;
bool queues_cleared = false;
}  // namespace internals
double base_time = 0;  // reference: rtlib/osc.hpp:15

#define RATE internals::rate
#define CLIENT_NAME internals::client_name

inline double current_time_secs() {  // reference: tinyspec.cpp:110-112
    return (double(internals::time_samples) + internals::time_fract / double(uint64_t(-1))) / internals::rate;
}
inline void set_hop(double new_hop) {  // reference: rtlib/control.hpp:22-29
    if (new_hop < 0) std::cerr << "ignoring invalid hop: " << new_hop << std::endl;
    else {
        uint64_t s, f;
        tsbrt::check(tsb_hop_fixed(new_hop, &s, &f), "tsb_hop_fixed");
        internals::hop = new_hop;
        internals::hop_samples = s;
        internals::hop_fract = f;
    }
}
inline void next_hop_samples(uint32_t n, double h) {
    internals::new_window_size = n;
    set_hop(h);
}
inline void next_hop_ratio(uint32_t n, double ratio = 0.25) {
    internals::new_window_size = n;
    set_hop(n * ratio);
}
inline void next_hop_hz(uint32_t n, double hz) {
    internals::new_window_size = n;
    set_hop(internals::rate / hz);
}
inline void set_process_fn(func_t fn) { internals::fptr = fn; }
inline void set_time(double secs) {  // reference: rtlib/control.hpp:45-53
    if (secs < 0)
        std::cerr << "negative time is not supported, ignoring call to set_time(" << secs << ")" << std::endl;
    else {
        base_time += current_time_secs() - secs;
        internals::time_samples = uint64_t(secs * internals::rate);
        internals::time_fract = std::fmod(secs * internals::rate, 1.0) * double(uint64_t(-1));
    }
}
inline void set_num_channels(size_t in, size_t out) {  // reference: rtlib/control.hpp:54-77 (ports are JACK's business)
    internals::queues_cleared = true;  // ainqueue / aqueue are cleared (:56-57), the overlap tail is not (:58)
    internals::new_nch_in = in;
    internals::new_nch_out = out;
}
inline void skip_to_now() {  // reference: rtlib/control.hpp:78-82: keep one frame's worth of queued input
    int64_t keep = (int64_t)(internals::computed_hop + internals::window_size);
    if (internals::in_len - internals::in_pos > keep) internals::in_len = internals::in_pos + keep;
}
