// oracle/ref/taps/granular_tap.cpp — TEST INFRASTRUCTURE: a "hack" for the UNMODIFIED reference.
// Body of the reference's hacks/pitchshift-simple.cpp:2-11 (granular time-domain pitch shift: per-frame
// varying hop, out.size != in.size, no FFT).  TSREF_VARY=0: next frame size 1800 as in the hack;
// TSREF_VARY=1: BASELINE cfg4 schedule N_next = 2^(6 + ((f+1+s) mod 8)).
#include "tap_common.inc"
const long TAP_VARY = tsref_env_long("TSREF_VARY", 0);
const long TAP_NFIRST = tsref_env_long("TSREF_NFIRST", 0);
const long TAP_STREAM = tsref_env_long("TSREF_STREAM", 0);
if (TAP_NFIRST) internals::new_window_size = TAP_NFIRST;
set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    int64_t w = in.size;
    tsref_dump("frame", &tap_frame, sizeof(int64_t));
    tsref_dump("t", &t, sizeof t);
    tsref_dump("win", &w, sizeof w);
    double shift = sin(t)/2+1;
    window_hann(in);
    out.resize(in.num_channels, in.size/shift);
    for (size_t i = 0; i < in.num_channels; i++)
        for (size_t j = 0; j < in.size; j++)
            for (int k = j/shift; k < (j+1)/shift; k++)
                out[i][k] = in[i][j];
    uint32_t n = TAP_VARY ? (1u << (6 + ((tap_frame + 1 + TAP_STREAM) % 8))) : 1800;
    next_hop_ratio(n, 0.5/shift);
    int64_t os = out.size;
    tsref_dump("osize", &os, sizeof os);
    tap_frame++;
    tsref_frame_done();
});
tsref_go();
