// oracle/ref/taps/bytebeat_tap.cpp — TEST INFRASTRUCTURE: a "hack" for the UNMODIFIED reference.
// BASELINE cfg5 pattern (hacks/spectral_bytebeat/voidbubbles.cpp:4-19): no input, a Hermitian
// spectrum of length 2*frame filled from a seeded integer formula, frft(-1), first `frame` reals
// times a triangular window.  The spectrum formula here is integer-only (LCG hash of (frame, c, i))
// so that a device-side generator can reproduce it bit-exactly.
#include "tap_common.inc"
const long TAP_N = tsref_env_long("TSREF_N", 4096);
const long TAP_CH = tsref_env_long("TSREF_CH", 2);
next_hop_ratio(TAP_N, 0.25);
set_num_channels(0, TAP_CH);
set_process_fn([&](WaveBuf&, WaveBuf &out, double t) {
    static FFTBuf buf;
    buf.resize(out.num_channels, out.size*2);
    int n = out.size;
    for (size_t c = 0; c < out.num_channels; c++) {
        for (int i = 1; i < n; i++) {
            uint32_t h = (uint32_t)tap_frame*2654435761u ^ ((uint32_t)c*40503u + (uint32_t)i*2246822519u);
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            double re = ((int)(h & 0xffff) - 32768) / 32768.0 / (double)(1 + (i >> 4));
            double im = ((int)(h >> 16) - 32768) / 32768.0 / (double)(1 + (i >> 4));
            buf[c][i] = cplx(re, im);
            buf[c][n*2-i] = conj(buf[c][i]);
        }
    }
    frft(buf, buf, -1);
    for (size_t c = 0; c < out.num_channels; c++)
        for (int i = 0; i < n; i++)
            out[c][i] = buf[c][i].real() * (1.0-abs(1.0-i/(n/2.0)));
    if (tap_selected()) {
        tsref_dump("t", &t, sizeof t);
        tsref_dump("out", out.data, sizeof(double)*out.size*out.num_channels);
    }
    tap_frame++;
    tsref_frame_done();
});
tsref_go();
