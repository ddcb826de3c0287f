// oracle/ref/taps/pv_tap.cpp — TEST INFRASTRUCTURE: a "hack" for the UNMODIFIED reference.
// Same call sequence as the reference's hacks/pitchshift-pv.cpp:4-9 (analyze -> shift -> synthesize ->
// next_hop_ratio) with the BASELINE cfg1/cfg3 constants taken from the environment and set before
// the first frame, plus dumps of the per-frame double buffers and PV state.
#include "tap_common.inc"
const long TAP_N = tsref_env_long("TSREF_N", 1024);
const double TAP_HOP_RATIO = tsref_env_double("TSREF_HOP_RATIO", 0.25);
const double TAP_SHIFT = tsref_env_double("TSREF_SHIFT", 2.0);
const long TAP_CH = tsref_env_long("TSREF_CH", 2);
if (TAP_CH != 2) set_num_channels(TAP_CH, TAP_CH);
next_hop_ratio(TAP_N, TAP_HOP_RATIO);
PhaseVocoder pv;

set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    bool sel = tap_selected();
    if (sel) {
        int64_t f = tap_frame;
        tsref_dump("frame", &f, sizeof f);
        tsref_dump("t", &t, sizeof t);
        tsref_dump("in", in.data, sizeof(double)*in.size*in.num_channels);
    }
    pv.analyze(in);
    pv.shift([&](double x){ return x*TAP_SHIFT; });
    pv.synthesize(out);
    next_hop_ratio(TAP_N, TAP_HOP_RATIO);
    if (sel) {
        tsref_dump("out", out.data, sizeof(double)*out.size*out.num_channels);
        tsref_dump("last_phase", pv.last_phase.data, sizeof(double)*pv.size*pv.num_channels);
        tsref_dump("phase_sum", pv.phase_sum.data, sizeof(double)*pv.size*pv.num_channels);
        tsref_dump("pv_data", pv.data.data, sizeof(PVBin)*pv.size*pv.num_channels);
    }
    tap_frame++;
    tsref_frame_done();
});

pv.reset();
tsref_go();
