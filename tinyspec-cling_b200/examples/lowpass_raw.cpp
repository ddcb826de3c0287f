// A hack that mixes helpers with its own per-bin loop (cfg2 "filter" variant, g[k] = 1/(1+k/64)):
// the loop is user code on the host, so this one runs in immediate mode.
set_num_channels(1, 1);
next_hop_ratio(2048, 0.25);
FFTBuf fft;

set_process_fn([&](WaveBuf &in, WaveBuf &out, double) {
    window_sqrt_hann(in);
    fft.fill_from(in);
    frft(fft, fft, 1);
    size_t n = fft.size;
    for (size_t c = 0; c < fft.num_channels; c++)
        for (size_t k = 0; k <= n/2; k++) {
            double g = 1.0/(1.0 + k/64.0);
            fft[c][k] *= g;
            if (k && k != n-k) fft[c][n-k] *= g;
        }
    frft(fft, fft, -1);
    out.fill_from<cplx>(fft, [](cplx x){ return x.real(); });
    window_sqrt_hann(out);
    next_hop_ratio(2048, 0.25);
});
