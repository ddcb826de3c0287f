// BASELINE cfg5 pattern (the reference's hacks/spectral_bytebeat/*.cpp): no input; every frame fills a Hermitian
// spectrum of twice the frame length from an integer formula, inverse-transforms it with frft(-1) and keeps the
// first half under a triangular window.  The per-bin loop is the hack's own code (immediate mode); the large
// inverse FFT and the overlap-add run on the device.  BYTEBEAT_N: frame size (default 65536 -> FFT of 131072).
const long bytebeat_n = getenv("BYTEBEAT_N") ? atol(getenv("BYTEBEAT_N")) : 65536;
const long bytebeat_ch = getenv("BYTEBEAT_CH") ? atol(getenv("BYTEBEAT_CH")) : 2;
long bytebeat_frame = 0;
next_hop_ratio(bytebeat_n, 0.25);
set_num_channels(0, bytebeat_ch);
set_process_fn([&](WaveBuf &, WaveBuf &out, double) {
    static FFTBuf buf;
    buf.resize(out.num_channels, out.size*2);
    int n = out.size;
    for (size_t c = 0; c < out.num_channels; c++) {
        for (int i = 1; i < n; i++) {
            uint32_t h = (uint32_t)bytebeat_frame*2654435761u ^ ((uint32_t)c*40503u + (uint32_t)i*2246822519u);
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            double re = ((int)(h & 0xffff) - 32768)/32768.0/(double)(1 + (i >> 4));
            double im = ((int)(h >> 16) - 32768)/32768.0/(double)(1 + (i >> 4));
            buf[c][i] = cplx(re, im);
            buf[c][n*2 - i] = conj(buf[c][i]);
        }
    }
    frft(buf, buf, -1);
    for (size_t c = 0; c < out.num_channels; c++)
        for (int i = 0; i < n; i++)
            out[c][i] = buf[c][i].real()*(1.0 - abs(1.0 - i/(n/2.0)));
    bytebeat_frame++;
});
