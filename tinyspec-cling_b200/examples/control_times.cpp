// The control API used from inside the frame callback (the reference's rtlib/control.hpp:22-53): a fractional hop
// (next_hop_samples), set_time (and a rejected negative one), next_hop_hz, a frame-size change through
// next_hop_ratio's default ratio, a rejected set_hop.  CONTROL_DUMP: file receiving {t, base_time, frame size, first
// input sample} per frame (doubles).  The callback touches samples itself: immediate mode.
FILE *control_dump = getenv("CONTROL_DUMP") ? fopen(getenv("CONTROL_DUMP"), "wb") : nullptr;
long control_frame = 0;
next_hop_samples(256, 64.25);
set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    double rec[4] = {t, base_time, (double)in.size, in[0][0]};
    if (control_dump) { fwrite(rec, sizeof rec, 1, control_dump); fflush(control_dump); }
    for (size_t c = 0; c < out.num_channels; c++)
        for (size_t i = 0; i < out.size; i++)
            out[c][i] = 0.5*in[c][i];
    if (control_frame == 3) set_time(12.5);
    if (control_frame == 5) set_time(-1.0);
    if (control_frame == 7) next_hop_hz(256, 700.0);
    if (control_frame == 9) next_hop_ratio(512);
    if (control_frame == 11) set_hop(-3.0);
    if (control_frame == 13) next_hop_samples(512, 300.75);
    control_frame++;
});
