// oracle/ref/taps/pvmore_tap.cpp — TEST INFRASTRUCTURE: a one-shot "hack" for the UNMODIFIED reference covering the rest
// of the PhaseVocoder surface (rtinc/pv.hpp:12-63, rtlib/pv.hpp:4-79): analyze -> synthesize WITHOUT shift (all N bins
// keep state), shift with a NON-LINEAR fn, the shift(function<void(WaveBuf&)>) overload, a per-object set_hop that
// differs from the global hop and use_global_hop(), reset(last, sum) one at a time, fill_from.
#include "tap_common.inc"
{
    RATE = 48000;
    uint32_t seed = 777;
    auto rnd = [&]() { seed = seed*1664525u + 1013904223u; return ((seed >> 8)/16777216.0 - 0.5); };
    const size_t n = 128;
    set_hop(32.0);   // the global hop (internals::hop)
    auto fill = [&](WaveBuf &in, int frame) {
        for (size_t i = 0; i < 2*n; i++) in.data[i] = (float)(0.3*rnd() + 0.5*sin(0.41*(i%n + frame*32) + (i/n)));
    };
    auto dump_pv = [&](const char *tag, PhaseVocoder &pv, WaveBuf &out) {
        std::string t(tag);
        tsref_dump((t + "_out").c_str(), out.data, sizeof(double)*2*n);
        tsref_dump((t + "_data").c_str(), pv.data.data, sizeof(PVBin)*2*n);
        tsref_dump((t + "_phase_sum").c_str(), pv.phase_sum.data, sizeof(double)*2*n);
        tsref_dump((t + "_last_phase").c_str(), pv.last_phase.data, sizeof(double)*2*n);
    };
    {   // 1. no shift: every one of the N bins is analysed and synthesised
        PhaseVocoder pv;
        WaveBuf in(2, n), out;
        for (int frame = 0; frame < 4; frame++) {
            fill(in, frame);
            tsref_dump("in", in.data, sizeof(double)*2*n);
            pv.analyze(in);
            pv.synthesize(out);
            dump_pv("noshift", pv, out);
        }
    }
    {   // 2. a non-linear fn (bins bunch up and spread out; several sources per destination, gaps)
        PhaseVocoder pv;
        WaveBuf in(2, n), out;
        for (int frame = 0; frame < 4; frame++) {
            fill(in, frame);
            tsref_dump("in", in.data, sizeof(double)*2*n);
            pv.analyze(in);
            pv.shift([&](double x){ return 300.0 + 0.6*x + x*x/40000.0; });
            pv.synthesize(out);
            dump_pv("nonlin", pv, out);
        }
    }
    {   // 3. the buffer overload: fn sees all bins (or all frequencies) of all channels at once
        PhaseVocoder pv;
        WaveBuf in(2, n), out;
        for (int frame = 0; frame < 3; frame++) {
            fill(in, frame);
            tsref_dump("in", in.data, sizeof(double)*2*n);
            pv.analyze(in);
            pv.shift([&](WaveBuf &b){
                for (size_t c = 0; c < b.num_channels; c++)
                    for (size_t j = 0; j < b.size; j++)
                        b[c][j] = b[c][j]*(1.25 + 0.25*c) + 3.0*(j%5);
            });
            pv.synthesize(out);
            dump_pv("bufshift", pv, out);
        }
    }
    {   // 4. per-object hop (rtinc/pv.hpp:46-53), then back to the global one; reset(last) / reset(sum) separately
        PhaseVocoder pv;
        pv.set_hop(48.5);
        WaveBuf in(2, n), out;
        for (int frame = 0; frame < 6; frame++) {
            fill(in, frame);
            tsref_dump("in", in.data, sizeof(double)*2*n);
            if (frame == 2) pv.use_global_hop();
            if (frame == 3) pv.reset(true, false);
            if (frame == 4) pv.reset(false, true);
            double h = pv.hop();
            tsref_dump("ownhop_hop", &h, sizeof h);
            pv.analyze(in);
            pv.shift([&](double x){ return x*1.5; });
            pv.synthesize(out);
            dump_pv("ownhop", pv, out);
        }
        PhaseVocoder copy;
        copy.fill_from(pv);
        fill(in, 6);
        tsref_dump("in", in.data, sizeof(double)*2*n);
        copy.analyze(in);
        copy.synthesize(out);
        dump_pv("copy", copy, out);
    }
    extern void tsref_finish();
    tsref_finish();
}
