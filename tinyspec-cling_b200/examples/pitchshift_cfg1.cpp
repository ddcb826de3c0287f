// BASELINE cfg1: the reference's hacks/pitchshift-pv.cpp chain with frame 1024 / hop 256 set before the first frame.
next_hop_ratio(1024, 0.25);
PhaseVocoder pv;

set_process_fn([&](WaveBuf &in, WaveBuf &out, double) {
    pv.analyze(in);
    pv.shift([](double x){ return x*2; });
    pv.synthesize(out);
    next_hop_ratio(1024, 0.25);
});

pv.reset();
