// oracle/ref/taps/filter_tap.cpp — TEST INFRASTRUCTURE: a "hack" for the UNMODIFIED reference.
// BASELINE cfg2 chain: window_sqrt_hann(in) -> fill_from -> frft(+1) -> real gain g[k]
// on bins k and (N-k)%N -> frft(-1) -> real part -> window_sqrt_hann(out), built only from the
// reference's own helpers.  TSREF_FILTER=0: passthrough g==1; 1: g[k] = 1/(1+k/64).
#include "tap_common.inc"
const long TAP_N = tsref_env_long("TSREF_N", 2048);
const double TAP_HOP_RATIO = tsref_env_double("TSREF_HOP_RATIO", 0.25);
const long TAP_CH = tsref_env_long("TSREF_CH", 1);
const long TAP_FILTER = tsref_env_long("TSREF_FILTER", 0);
if (TAP_CH != 2) set_num_channels(TAP_CH, TAP_CH);
next_hop_ratio(TAP_N, TAP_HOP_RATIO);
FFTBuf tap_fft;

set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    bool sel = tap_selected();
    if (sel) {
        int64_t f = tap_frame;
        tsref_dump("frame", &f, sizeof f);
        tsref_dump("t", &t, sizeof t);
        tsref_dump("in", in.data, sizeof(double)*in.size*in.num_channels);
    }
    window_sqrt_hann(in);
    tap_fft.fill_from(in);
    frft(tap_fft, tap_fft, 1);
    if (sel) tsref_dump("spec", tap_fft.data, sizeof(cplx)*tap_fft.size*tap_fft.num_channels);
    if (TAP_FILTER) {
        size_t n = tap_fft.size;
        for (size_t c = 0; c < tap_fft.num_channels; c++)
            for (size_t k = 0; k <= n/2; k++) {
                double g = 1.0/(1.0 + k/64.0);
                tap_fft[c][k] *= g;
                if (k && k != n-k) tap_fft[c][n-k] *= g;
            }
    }
    frft(tap_fft, tap_fft, -1);
    out.fill_from<cplx>(tap_fft, [](cplx x){ return x.real(); });
    window_sqrt_hann(out);
    next_hop_ratio(TAP_N, TAP_HOP_RATIO);
    if (sel) tsref_dump("out", out.data, sizeof(double)*out.size*out.num_channels);
    tap_frame++;
    tsref_frame_done();
});
tsref_go();
