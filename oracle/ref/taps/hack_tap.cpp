// oracle/ref/taps/hack_tap.cpp — TEST INFRASTRUCTURE: runs an UNMODIFIED hack of the reference (TSREF_HACK_FILE, e.g.
// /root/reference/hacks/spectral_bytebeat/voidbubbles.cpp, compiled where it lies — nothing is copied) and taps what
// its callback produces.  The hack installs its process function itself; the tap then wraps internals::fptr so that
// every frame's time and double `out` buffer are dumped.  The fake JACK driver starts only at tsref_go().
#include "tap_common.inc"
#include TSREF_HACK_FILE
{
    auto inner = internals::fptr;
    set_process_fn([=](WaveBuf &in, WaveBuf &out, double t) {
        inner(in, out, t);
        if (tap_selected()) {
            tsref_dump("t", &t, sizeof t);
            tsref_dump("out", out.data, sizeof(double)*out.size*out.num_channels);
        }
        tap_frame++;
        tsref_frame_done();
    });
}
tsref_go();
