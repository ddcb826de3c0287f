// oracle/ref/taps/control_tap.cpp — TEST INFRASTRUCTURE: a "hack" for the UNMODIFIED reference that exercises the
// control API from inside the frame callback (rtlib/control.hpp:22-53): a fractional hop (next_hop_samples), set_time
// (and a rejected negative one), next_hop_hz, a frame-size change through next_hop_ratio's default ratio, a rejected
// set_hop.  The same body is tinyspec-cling_b200/examples/control_times.cpp.  Dumps per frame: t, base_time, the frame
// size seen and the first input sample; the output stream pins the hop clock.
#include "tap_common.inc"
next_hop_samples(256, 64.25);
set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    double rec[4] = {t, base_time, (double)in.size, in[0][0]};
    tsref_dump("rec", rec, sizeof rec);
    for (size_t c = 0; c < out.num_channels; c++)
        for (size_t i = 0; i < out.size; i++)
            out[c][i] = 0.5*in[c][i];
    if (tap_frame == 3) set_time(12.5);
    if (tap_frame == 5) set_time(-1.0);
    if (tap_frame == 7) next_hop_hz(256, 700.0);
    if (tap_frame == 9) next_hop_ratio(512);
    if (tap_frame == 11) set_hop(-3.0);
    if (tap_frame == 13) next_hop_samples(512, 300.75);
    tap_frame++;
    tsref_frame_done();
});
tsref_go();
