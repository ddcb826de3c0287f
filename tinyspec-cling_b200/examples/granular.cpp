// BASELINE cfg4: granular time-domain pitch shift with a per-frame varying frame size and hop and an output
// frame shorter/longer than the input frame (the algorithm of the reference's hacks/pitchshift-simple.cpp).
// No FFT: it exercises the frame gather, window_hann and the overlap-add with changing sizes.  The resampling
// loop is the hack's own code, so this runs in immediate mode.
// GRANULAR_VARY=1: next frame size 2^(6 + ((f+1+s) mod 8)) (64..8192) instead of the constant 1800.
long granular_frame = 0;
const long granular_vary = getenv("GRANULAR_VARY") ? atol(getenv("GRANULAR_VARY")) : 0;
const long granular_stream = getenv("GRANULAR_STREAM") ? atol(getenv("GRANULAR_STREAM")) : 0;
if (getenv("GRANULAR_NFIRST")) internals::new_window_size = atol(getenv("GRANULAR_NFIRST"));
set_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {
    double shift = sin(t)/2 + 1;
    window_hann(in);
    out.resize(in.num_channels, in.size/shift);
    for (size_t c = 0; c < in.num_channels; c++)
        for (size_t j = 0; j < in.size; j++)
            for (int k = j/shift; k < (j + 1)/shift; k++)
                out[c][k] = in[c][j];
    uint32_t n = granular_vary ? (1u << (6 + ((granular_frame + 1 + granular_stream) % 8))) : 1800;
    next_hop_ratio(n, 0.5/shift);
    granular_frame++;
});
