// oracle/ref/taps/helpers_tap.cpp — TEST INFRASTRUCTURE: known-answer vectors from the reference's
// own helpers (no scheduler): windows, frft(+1/-1/fractional) on pow-2 and non-pow-2 sizes, and one
// PhaseVocoder analyze/shift/synthesize step with explicit hop.  Executes inside main() and exits.
#include "tap_common.inc"
{
    RATE = 48000;
    uint32_t lcg = 777;
    auto rnd = [&]() { lcg = lcg*1664525u + 1013904223u; return (double)(lcg >> 8)/16777216.0 - 0.5; };
    for (size_t n : {8, 64, 100, 512}) {
        WaveBuf w(2, n);
        for (size_t i = 0; i < 2*n; i++) w.data[i] = 1.0;
        window_hann(w);
        tsref_dump("hann", w.data, sizeof(double)*2*n);
        for (size_t i = 0; i < 2*n; i++) w.data[i] = 1.0;
        window_sqrt_hann(w);
        tsref_dump("sqrt_hann", w.data, sizeof(double)*2*n);
        FFTBuf x(2, n), y(2, n);
        for (size_t i = 0; i < 2*n; i++) x.data[i] = cplx(rnd(), rnd());
        tsref_dump("frft_in", x.data, sizeof(cplx)*2*n);
        for (double e : {1.0, -1.0, 2.0, 0.5}) {
            frft(x, y, e);
            tsref_dump("frft_out", y.data, sizeof(cplx)*2*n);
        }
        FFTBuf z(2, n);
        memcpy(z.data, x.data, sizeof(cplx)*2*n);
        frft(z, z, 1.0);   // in-place, as every reference call site does
        tsref_dump("frft_inplace", z.data, sizeof(cplx)*2*n);
    }
    for (size_t n : {64, 256}) {
        for (double ratio : {2.0, 1.5, 0.5}) {
            PhaseVocoder pv;
            pv.set_hop(n/4.0);
            WaveBuf in(2, n), out;
            for (int frame = 0; frame < 3; frame++) {
                for (size_t i = 0; i < 2*n; i++) in.data[i] = (float)(rnd() + 0.5*sin(0.37*(i + frame*n/4)));
                tsref_dump("pv_in", in.data, sizeof(double)*2*n);
                pv.analyze(in);
                tsref_dump("pv_analyzed", pv.data.data, sizeof(PVBin)*2*n);
                pv.shift([&](double x){ return x*ratio; });
                tsref_dump("pv_shifted", pv.data.data, sizeof(PVBin)*2*n);
                pv.synthesize(out);
                tsref_dump("pv_out", out.data, sizeof(double)*2*n);
                tsref_dump("pv_phase_sum", pv.phase_sum.data, sizeof(double)*2*n);
                tsref_dump("pv_last_phase", pv.last_phase.data, sizeof(double)*2*n);
            }
        }
    }
    // fixed-point hop clock (rtlib/control.hpp:22-29 + tinyspec.cpp:163-168) is exercised by the
    // scheduler taps; here only the set_hop decomposition.
    for (double h : {256.0, 256.5, 100.25, 0.75, 1023.999}) {
        set_hop(h);
        uint64_t v[2] = {internals::hop_samples, internals::hop_fract};
        tsref_dump("hop_fixed", v, sizeof v);
    }
    extern void tsref_finish();
    tsref_finish();
}
