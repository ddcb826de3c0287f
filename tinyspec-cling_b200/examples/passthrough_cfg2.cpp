// BASELINE cfg2 (pass-through variant): mono, frame 2048, hop 512, sqrt-Hann analysis and synthesis windows
// around frft(+1) / frft(-1) — a chain made only of the library's helpers.
set_num_channels(1, 1);
next_hop_ratio(2048, 0.25);
FFTBuf fft;

set_process_fn([&](WaveBuf &in, WaveBuf &out, double) {
    window_sqrt_hann(in);
    fft.fill_from(in);
    frft(fft, fft, 1);
    frft(fft, fft, -1);
    out.fill_from<cplx>(fft, [](cplx x){ return x.real(); });
    window_sqrt_hann(out);
    next_hop_ratio(2048, 0.25);
});
