/* oracle/oracle.c — TEST INFRASTRUCTURE.  NOT product code, never on the product path.
 *
 * A plain-C, CPU, double-precision restatement of the reference's overlap-add spectral hot path
 * (nwoeanhinnogaehr/tinyspec-cling): every function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library, and
 * only as the checker.
 *
 * Pinning: this restatement is checked (tests/test_oracle_golden.py) against fixtures in
 * tests/golden/ that were produced by running the UNMODIFIED reference compiled in the authoring
 * container (oracle/Makefile `ref` target, oracle/gen_golden.py).  The reference has no tests or
 * golden vectors of its own, and its FFT arithmetic lives in FFTW, an external and
 * version-unpinned dependency absent from /root/reference: at that one boundary parity is pinned
 * only by FFTW's documented DFT definition (forward, exp(-2*pi*i*j*k/n), unnormalised).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "signal_gen.h"

typedef double _Complex cplx;

/* ------------------------------------------------------------------ synthetic input */
void orc_gen_signal(int stream, int nch, double rate, size_t count, float *dst_interleaved) {
    ts_siggen g;
    ts_siggen_init(&g, stream, nch, rate);
    ts_siggen_fill(&g, dst_interleaved, count);
}

/* ------------------------------------------------------------------ FFT provider
 * Stands where FFTW stands in the reference (rtlib/fft.hpp:13-18): forward DFT, kernel
 * exp(-2*pi*i*j*k/n), unnormalised, double.  Power-of-two n: iterative radix-2 with long-double
 * generated twiddles; other n: direct DFT accumulated in long double. */
static void fft_forward(const cplx *in, cplx *out, size_t n) {
    const long double tau = 6.283185307179586476925286766559005768L;
    if (n == 0) return;
    if ((n & (n - 1)) == 0) {
        int lg = 0;
        while (((size_t)1 << lg) < n) lg++;
        cplx *a = (cplx *)malloc(sizeof(cplx) * n);
        cplx *w = (cplx *)malloc(sizeof(cplx) * (n / 2 + 1));
        for (size_t k = 0; k < n / 2 + 1; k++) {
            long double ang = -tau * (long double)k / (long double)n;
            w[k] = (double)cosl(ang) + I * (double)sinl(ang);
        }
        for (size_t i = 0; i < n; i++) {
            size_t r = 0;
            for (int b = 0; b < lg; b++) r |= ((i >> b) & 1u) << (lg - 1 - b);
            a[r] = in[i];
        }
        for (size_t len = 2; len <= n; len <<= 1) {
            size_t half = len >> 1, step = n / len;
            for (size_t s = 0; s < n; s += len)
                for (size_t j = 0; j < half; j++) {
                    cplx u = a[s + j], v = a[s + j + half] * w[j * step];
                    a[s + j] = u + v;
                    a[s + j + half] = u - v;
                }
        }
        memcpy(out, a, sizeof(cplx) * n);
        free(a);
        free(w);
    } else {
        long double *c = (long double *)malloc(sizeof(long double) * n);
        long double *s = (long double *)malloc(sizeof(long double) * n);
        cplx *res = (cplx *)malloc(sizeof(cplx) * n);
        for (size_t k = 0; k < n; k++) {
            long double ang = tau * (long double)k / (long double)n;
            c[k] = cosl(ang);
            s[k] = -sinl(ang);
        }
        for (size_t k = 0; k < n; k++) {
            long double re = 0, im = 0;
            for (size_t j = 0; j < n; j++) {
                size_t t = (size_t)(((unsigned long long)j * k) % n);
                long double xr = creal(in[j]), xi = cimag(in[j]);
                re += xr * c[t] - xi * s[t];
                im += xr * s[t] + xi * c[t];
            }
            res[k] = (double)re + I * (double)im;
        }
        memcpy(out, res, sizeof(cplx) * n);
        free(c);
        free(s);
        free(res);
    }
}
void orc_fft_forward(const double *in, double *out, size_t n) { fft_forward((const cplx *)in, (cplx *)out, n); }

/* ------------------------------------------------------------------ windows
 * rtlib/windowfn.hpp:1-7 (Hann) and 8-14 (sqrt-Hann): periodic, recomputed per call, i-outer. */
void orc_window_hann(double *data, size_t nch, size_t size) {
    for (size_t i = 0; i < size; i++) {
        double w = 0.5 * (1 - cos(2 * M_PI * i / size));
        for (size_t j = 0; j < nch; j++) data[j * size + i] *= w;
    }
}
void orc_window_sqrt_hann(double *data, size_t nch, size_t size) {
    for (size_t i = 0; i < size; i++) {
        double w = sqrt(0.5 * (1 - cos(2 * M_PI * i / size)));
        for (size_t j = 0; j < nch; j++) data[j * size + i] *= w;
    }
}

/* ------------------------------------------------------------------ frft
 * rtlib/fft.hpp:3-42.  `in` and `out` may alias (every reference call site is in place): the FFT
 * result lands in `out`, is copied to a temporary (line 19), and the eigen-combination loop then
 * reads in[c][i] / in[c][j] — i.e. already-overwritten data when in == out (lines 27,29). */
void orc_frft(double *in_, double *out_, size_t nch, size_t n, double exponent) {
    cplx *in = (cplx *)in_, *out = (cplx *)out_;
    if (nch * n == 0) return; /* fft.hpp:8-9 */
    cplx *tmp = (cplx *)malloc(sizeof(cplx) * nch * n);
    for (size_t c = 0; c < nch; c++) fft_forward(in + c * n, tmp + c * n, n); /* 13-18 */
    memcpy(out, tmp, sizeof(cplx) * nch * n); /* plan writes `out`; 19 copies it to fft_tmp */
    cplx im = I;
    cplx im1 = cpow(im, exponent);     /* 21 */
    cplx im2 = cpow(im, exponent * 2); /* 22 */
    cplx im3 = cpow(im, exponent * 3); /* 23 */
    double sqrtn = sqrt((double)n);    /* 24 */
    for (size_t c = 0; c < nch; c++) {
        for (size_t i = 0; i < n; i++) {
            size_t j = i ? n - i : 0; /* 27 */
            cplx f0 = in[c * n + i];
            cplx f1 = tmp[c * n + i] / sqrtn;
            cplx f2 = in[c * n + j];
            cplx f3 = tmp[c * n + j] / sqrtn;
            cplx b0 = f0 + f1 + f2 + f3;
            cplx b1 = f0 + im * f1 - f2 - im * f3;
            cplx b2 = f0 - f1 + f2 - f3;
            cplx b3 = f0 - im * f1 - f2 + im * f3;
            b1 *= im1;
            b2 *= im2;
            b3 *= im3;
            out[c * n + i] = (b0 + b1 + b2 + b3) / 4.0; /* 39 */
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------ control: hop fixed point
 * rtlib/control.hpp:22-29. Returns 0 if the hop is rejected (negative). */
int orc_set_hop(double new_hop, double *hop, uint64_t *hop_samples, uint64_t *hop_fract) {
    if (new_hop < 0) return 0;
    *hop = new_hop;
    *hop_samples = (uint64_t)new_hop;
    *hop_fract = (uint64_t)(fmod(new_hop, 1.0) * (double)(uint64_t)(-1));
    return 1;
}

/* ------------------------------------------------------------------ Buf<T> (rtinc/buffer.hpp:7-36) */
typedef struct {
    double *data; /* element = `width` doubles (1: WaveBuf, 2: FFTBuf / PVBuf) */
    size_t size, nch, width;
} obuf;
static void obuf_zero(obuf *b) { memset(b->data, 0, sizeof(double) * b->width * b->size * b->nch); }
/* resize compares PRODUCTS (buffer.hpp:16): (2,1024)->(1,2048) is a no-op keeping the old shape */
static void obuf_resize(obuf *b, size_t nch, size_t size) {
    if (b->size * b->nch != size * nch) {
        b->size = size;
        b->nch = nch;
        free(b->data);
        b->data = (double *)malloc(sizeof(double) * b->width * (nch * size + 8)); /* +slack: pitchshift-simple.cpp:8 can write one element past the end */
        obuf_zero(b);
    }
}

/* ------------------------------------------------------------------ PhaseVocoder (rtinc/pv.hpp:12-63) */
typedef struct {
    obuf data;       /* PVBin {amp,freq} */
    obuf last_phase; /* double */
    obuf phase_sum;  /* double */
    obuf fft;        /* cplx scratch */
    size_t size, nch;
    int global_hop;
    double user_hop;
    double *global_hop_ptr; /* internals::hop */
    double *rate_ptr;       /* RATE */
} orc_pv;

orc_pv *orc_pv_new(double *global_hop_ptr, double *rate_ptr) {
    orc_pv *pv = (orc_pv *)calloc(1, sizeof(orc_pv));
    pv->data.width = 2;
    pv->last_phase.width = 1;
    pv->phase_sum.width = 1;
    pv->fft.width = 2;
    pv->global_hop = 1;
    pv->global_hop_ptr = global_hop_ptr;
    pv->rate_ptr = rate_ptr;
    return pv;
}
void orc_pv_free(orc_pv *pv) {
    free(pv->data.data);
    free(pv->last_phase.data);
    free(pv->phase_sum.data);
    free(pv->fft.data);
    free(pv);
}
void orc_pv_set_hop(orc_pv *pv, double h) { pv->user_hop = h; pv->global_hop = 0; } /* pv.hpp:46-49 */
static double pv_hop(orc_pv *pv) { return pv->global_hop ? *pv->global_hop_ptr : pv->user_hop; } /* 50-53 */
static void pv_resize(orc_pv *pv, size_t nch, size_t size) { /* pv.hpp:23-31 */
    if (pv->size * pv->nch != size * nch) {
        pv->size = size;
        pv->nch = nch;
        obuf_resize(&pv->data, nch, size);
        obuf_resize(&pv->last_phase, nch, size);
        obuf_resize(&pv->phase_sum, nch, size);
    }
}
void orc_pv_reset(orc_pv *pv, int last, int sum) { /* pv.hpp:32-35 */
    if (last && pv->last_phase.data) obuf_zero(&pv->last_phase);
    if (sum && pv->phase_sum.data) obuf_zero(&pv->phase_sum);
}
double *orc_pv_data(orc_pv *pv) { return pv->data.data; }
double *orc_pv_last_phase(orc_pv *pv) { return pv->last_phase.data; }
double *orc_pv_phase_sum(orc_pv *pv) { return pv->phase_sum.data; }
double *orc_pv_fft(orc_pv *pv) { return pv->fft.data; }
size_t orc_pv_size(orc_pv *pv) { return pv->size; }
size_t orc_pv_nch(orc_pv *pv) { return pv->nch; }

/* rtlib/pv.hpp:39-45 */
static double pv_phase_to_frequency(orc_pv *pv, size_t bin, double phase_diff) {
    double hop = pv_hop(pv), rate = *pv->rate_ptr;
    size_t size = pv->size;
    double delta = phase_diff - bin * 2.0 * M_PI * hop / size;
    int qpd = (int)(delta / M_PI);
    if (qpd >= 0) qpd += qpd & 1;
    else qpd -= qpd & 1;
    return (bin + size / hop * (delta - M_PI * qpd) / 2.0 / M_PI) * rate / size;
}
/* rtlib/pv.hpp:46-48 */
static double pv_frequency_to_phase(orc_pv *pv, double freq) {
    return 2.0 * M_PI * freq / *pv->rate_ptr * pv_hop(pv);
}
/* rtlib/pv.hpp:17-28: analyze(FFTBuf&) */
void orc_pv_analyze_fft(orc_pv *pv, const double *spec, size_t nch, size_t size) {
    const cplx *x = (const cplx *)spec;
    pv_resize(pv, nch, size);
    for (size_t i = 0; i < pv->nch; i++)
        for (size_t j = 0; j < pv->size; j++) {
            cplx v = x[i * pv->size + j];
            double amp = cabs(v);
            double phase = carg(v);
            double freq = pv_phase_to_frequency(pv, j, phase - pv->last_phase.data[i * pv->size + j]);
            pv->last_phase.data[i * pv->size + j] = phase;
            pv->data.data[2 * (i * pv->size + j)] = amp;
            pv->data.data[2 * (i * pv->size + j) + 1] = freq;
        }
}
/* rtlib/pv.hpp:29-38: synthesize(FFTBuf&) */
void orc_pv_synthesize_fft(orc_pv *pv, obuf *other) {
    obuf_resize(other, pv->nch, pv->size);
    cplx *y = (cplx *)other->data;
    for (size_t i = 0; i < pv->nch; i++)
        for (size_t j = 0; j < pv->size; j++) {
            size_t e = i * pv->size + j;
            double phase = pv_frequency_to_phase(pv, pv->data.data[2 * e + 1]);
            pv->phase_sum.data[e] += phase;
            double amp = pv->data.data[2 * e], ps = pv->phase_sum.data[e];
            y[e] = amp * cos(ps) + I * (amp * sin(ps)); /* std::polar */
        }
}
/* rtlib/pv.hpp:4-9: analyze(WaveBuf&) — windows the CALLER's buffer in place */
void orc_pv_analyze(orc_pv *pv, double *wave, size_t nch, size_t size) {
    orc_window_sqrt_hann(wave, nch, size);
    obuf_resize(&pv->fft, nch, size); /* fill_from resizes (buffer.hpp:26) */
    cplx *f = (cplx *)pv->fft.data;
    for (size_t e = 0; e < nch * size; e++) f[e] = wave[e];
    orc_frft(pv->fft.data, pv->fft.data, pv->fft.nch, pv->fft.size, 1);
    orc_pv_analyze_fft(pv, pv->fft.data, pv->fft.nch, pv->fft.size);
}
/* rtlib/pv.hpp:10-16: synthesize(WaveBuf&) */
static void pv_synthesize_wave(orc_pv *pv, obuf *other) {
    obuf_resize(other, pv->nch, pv->size);
    orc_pv_synthesize_fft(pv, &pv->fft);
    orc_frft(pv->fft.data, pv->fft.data, pv->fft.nch, pv->fft.size, -1);
    obuf_resize(other, pv->fft.nch, pv->fft.size); /* fill_from (buffer.hpp:26) */
    cplx *f = (cplx *)pv->fft.data;
    for (size_t e = 0; e < other->nch * other->size; e++) other->data[e] = creal(f[e]);
    orc_window_sqrt_hann(other->data, other->nch, other->size);
}
/* flat-buffer convenience for tests: `out` must hold nch*size doubles */
void orc_pv_synthesize(orc_pv *pv, double *out) {
    obuf o = {0};
    o.width = 1;
    pv_synthesize_wave(pv, &o);
    memcpy(out, o.data, sizeof(double) * o.nch * o.size);
    free(o.data);
}
/* rtlib/pv.hpp:49-79 with fn(x) = x*ratio (hacks/pitchshift-pv.cpp:6) */
void orc_pv_shift_scale(orc_pv *pv, double ratio) {
    size_t nch = pv->nch, size = pv->size, half = size / 2;
    double rate = *pv->rate_ptr;
    double *bins = (double *)malloc(sizeof(double) * (nch * half + 1));
    double *freqs = (double *)malloc(sizeof(double) * (nch * half + 1));
    for (size_t i = 0; i < nch; i++)
        for (size_t j = 0; j < half; j++) {
            bins[i * half + j] = j / (double)size * rate;
            freqs[i * half + j] = pv->data.data[2 * (i * size + j) + 1];
        }
    for (size_t e = 0; e < nch * half; e++) bins[e] = bins[e] * ratio;   /* fn(bins)  */
    for (size_t e = 0; e < nch * half; e++) freqs[e] = freqs[e] * ratio; /* fn(freqs) */
    double *tmp = (double *)malloc(sizeof(double) * 2 * (nch * size + 1));
    memcpy(tmp, pv->data.data, sizeof(double) * 2 * nch * size);
    obuf_zero(&pv->data);
    for (size_t i = 0; i < nch; i++)
        for (size_t j = 0; j < half; j++) {
            int k = (int)(bins[i * half + j] / rate * size);
            if (k >= 0 && k < (int)size / 2) {
                pv->data.data[2 * (i * size + k)] += tmp[2 * (i * size + j)];
                pv->data.data[2 * (i * size + k) + 1] = freqs[i * half + j];
            }
        }
    free(bins);
    free(freqs);
    free(tmp);
}

/* ------------------------------------------------------------------ frame scheduler + overlap-add
 * tinyspec.cpp:113-203 restated for an offline render: all input is queued up front, frames run
 * until the starve test (142-143) fails; the output-full throttle (144-145) only paces the
 * real-time thread and does not change any value, so it is not modelled. */
typedef struct orc_sched orc_sched;
typedef void (*orc_frame_fn)(orc_sched *s, obuf *in, obuf *out, double t, void *user);
struct orc_sched {
    /* namespace internals, tinyspec.cpp:35-61 */
    uint64_t window_size, new_window_size, hop_samples, hop_fract, computed_hop;
    double hop;
    uint64_t time_samples, time_fract;
    size_t nch_in, nch_out, new_nch_in, new_nch_out;
    double rate;
    obuf audio_in, audio_out;
    /* queues */
    const float *ainqueue; size_t ain_len, ain_pos;   /* interleaved, consumed from ain_pos */
    float *atmp; size_t atmp_len, atmp_cap, atmp_pos;   /* deque<float>: valid [atmp_pos, atmp_len) */
    float *aqueue; size_t aq_len, aq_cap;
    /* taps (not in the reference): double-precision OLA of the per-frame `out`, schedule tables */
    double *ola64; size_t ola64_cap; /* [pos*nch_out + c] */
    int64_t *tab_in_off, *tab_out_off, *tab_hop, *tab_win; double *tab_t; size_t tab_cap;
    size_t frames;
};
static void sched_init(orc_sched *s, size_t nch_in, size_t nch_out, double rate) {
    memset(s, 0, sizeof *s);
    s->new_window_size = 1 << 12;             /* tinyspec.cpp:38 */
    s->hop_samples = s->new_window_size / 2;  /* 39 */
    s->computed_hop = s->hop_samples;         /* 41 */
    s->hop = (double)s->hop_samples;          /* 42 */
    s->new_nch_in = nch_in;
    s->new_nch_out = nch_out;
    s->rate = rate;
    s->audio_in.width = s->audio_out.width = 1;
}
static void sched_free(orc_sched *s) {
    free(s->audio_in.data);
    free(s->audio_out.data);
    free(s->atmp);
}
static void push_out(orc_sched *s, float v) {
    if (s->aq_len == s->aq_cap) {
        s->aq_cap = s->aq_cap ? s->aq_cap * 2 : 1 << 16;
        s->aqueue = (float *)realloc(s->aqueue, sizeof(float) * s->aq_cap);
    }
    s->aqueue[s->aq_len++] = v;
}
static void atmp_push(orc_sched *s, float v) {
    if (s->atmp_len == s->atmp_cap) {
        /* compact, then grow */
        size_t live = s->atmp_len - s->atmp_pos;
        memmove(s->atmp, s->atmp + s->atmp_pos, sizeof(float) * live);
        s->atmp_pos = 0;
        s->atmp_len = live;
        if (live * 2 + 16 > s->atmp_cap) {
            s->atmp_cap = live * 2 + (1 << 12);
            s->atmp = (float *)realloc(s->atmp, sizeof(float) * s->atmp_cap);
        }
    }
    s->atmp[s->atmp_len++] = v;
}
/* control API as seen from inside a frame callback (rtlib/control.hpp:22-41) */
void orc_sched_set_hop(orc_sched *s, double h) { orc_set_hop(h, &s->hop, &s->hop_samples, &s->hop_fract); }
void orc_sched_next_hop_samples(orc_sched *s, uint32_t n, double h) { s->new_window_size = n; orc_sched_set_hop(s, h); }
void orc_sched_next_hop_ratio(orc_sched *s, uint32_t n, double r) { s->new_window_size = n; orc_sched_set_hop(s, n * r); }
void orc_sched_next_hop_hz(orc_sched *s, uint32_t n, double hz) { s->new_window_size = n; orc_sched_set_hop(s, s->rate / hz); }
/* tinyspec.cpp:110-112 */
static double sched_time(orc_sched *s) {
    return ((double)s->time_samples + s->time_fract / (double)(uint64_t)(-1)) / s->rate;
}

static void sched_run(orc_sched *s, orc_frame_fn fn, void *user, size_t max_frames) {
    uint64_t out_pos = 0; /* P_f: sample-frame position of this frame's output start */
    while (s->frames < max_frames) {
        /* 135-141 */
        if (s->new_nch_in != s->nch_in || s->new_nch_out != s->nch_out || s->new_window_size != s->window_size) {
            s->window_size = s->new_window_size;
            s->nch_in = s->new_nch_in;
            s->nch_out = s->new_nch_out;
        }
        obuf_resize(&s->audio_in, s->nch_in, s->window_size);
        obuf_resize(&s->audio_out, s->nch_out, s->window_size);
        size_t avail = s->ain_len - s->ain_pos;
        if (avail < s->nch_in * (s->computed_hop + s->window_size)) break; /* 142-143 */
        /* 147-149 advance input by the PREVIOUS hop */
        size_t adv = s->nch_in * s->computed_hop;
        s->ain_pos += adv < avail ? adv : avail;
        /* 151-157 interleaved float -> planar double */
        for (uint64_t i = 0; i < s->window_size; i++)
            for (size_t c = 0; c < s->nch_in; c++) {
                size_t q = i * s->nch_in + c;
                s->audio_in.data[c * s->audio_in.size + i] =
                    (q < s->ain_len - s->ain_pos) ? (double)s->ainqueue[s->ain_pos + q] : 0.0;
            }
        double t = sched_time(s);
        if (s->frames < s->tab_cap) {
            s->tab_in_off[s->frames] = s->nch_in ? (int64_t)(s->ain_pos / s->nch_in) : 0;
            s->tab_out_off[s->frames] = (int64_t)out_pos;
            s->tab_win[s->frames] = (int64_t)s->window_size;
            s->tab_t[s->frames] = t;
        }
        /* 161-169 */
        fn(s, &s->audio_in, &s->audio_out, t, user);
        if (s->time_fract + s->hop_fract < s->time_fract) s->computed_hop = s->hop_samples + 1;
        else s->computed_hop = s->hop_samples;
        s->time_samples += s->computed_hop;
        s->time_fract += s->hop_fract;
        if (s->frames < s->tab_cap) s->tab_hop[s->frames] = (int64_t)s->computed_hop;
        /* tap: double OLA of this frame's `out` at P_f */
        if (s->ola64) {
            size_t osz = s->audio_out.size;
            for (size_t i = 0; i < osz; i++)
                for (size_t c = 0; c < s->nch_out; c++) {
                    size_t q = (out_pos + i) * s->nch_out + c;
                    if (q < s->ola64_cap) s->ola64[q] += s->audio_out.data[c * osz + i];
                }
        }
        /* 173-182 output region overlapping previous frame(s): float narrowing */
        size_t osz = s->audio_out.size;
        uint64_t lim = s->computed_hop < osz ? s->computed_hop : osz;
        for (uint64_t i = 0; i < lim; i++)
            for (size_t c = 0; c < s->nch_out; c++) {
                float overlap = 0;
                if (s->atmp_pos < s->atmp_len) overlap = s->atmp[s->atmp_pos++];
                push_out(s, (float)(overlap + s->audio_out.data[c * osz + i]));
            }
        /* 184-191 hop larger than the frame */
        for (int i = 0; i < (int)s->nch_out * ((int)s->computed_hop - (int)osz); i++) {
            if (s->atmp_pos < s->atmp_len) push_out(s, s->atmp[s->atmp_pos++]);
            else push_out(s, 0);
        }
        /* 194-200 save region overlapping the next frame: float accumulate */
        for (int i = 0; i < (int)osz - (int)s->computed_hop; i++)
            for (size_t c = 0; c < s->nch_out; c++) {
                float out_val = (float)s->audio_out.data[c * osz + s->computed_hop + i];
                size_t q = (size_t)i * s->nch_out + c;
                if (q < s->atmp_len - s->atmp_pos) s->atmp[s->atmp_pos + q] += out_val;
                else atmp_push(s, out_val);
            }
        out_pos += s->computed_hop;
        s->frames++;
    }
}

/* ---- frame programs (the BASELINE configs' callbacks, built from the helpers above) ---- */
typedef struct {
    orc_pv *pv;
    uint32_t n;
    double hop_ratio, shift;
    double *frame_out; size_t frame_out_cap; /* optional per-frame `out` dump, [frame][c][i] */
    double *state_dump; size_t state_every;  /* optional: last_phase,phase_sum every k frames */
    size_t state_cap;
} pv_prog;
/* hacks/pitchshift-pv.cpp:4-9 */
static void pv_frame(orc_sched *s, obuf *in, obuf *out, double t, void *user) {
    (void)t;
    pv_prog *p = (pv_prog *)user;
    orc_pv_analyze(p->pv, in->data, in->nch, in->size);
    orc_pv_shift_scale(p->pv, p->shift);
    pv_synthesize_wave(p->pv, out);
    orc_sched_next_hop_ratio(s, p->n, p->hop_ratio);
    size_t per = out->nch * out->size;
    if (p->frame_out && (s->frames + 1) * per <= p->frame_out_cap)
        memcpy(p->frame_out + s->frames * per, out->data, sizeof(double) * per);
    if (p->state_dump && p->state_every && s->frames % p->state_every == 0) {
        size_t slot = s->frames / p->state_every, sz = p->pv->nch * p->pv->size;
        if ((slot + 1) * 2 * sz <= p->state_cap) {
            memcpy(p->state_dump + slot * 2 * sz, p->pv->last_phase.data, sizeof(double) * sz);
            memcpy(p->state_dump + slot * 2 * sz + sz, p->pv->phase_sum.data, sizeof(double) * sz);
        }
    }
}

typedef struct {
    uint32_t n;
    double hop_ratio;
    const double *gain; /* n/2+1 real gains applied to bins k and (n-k)%n, or NULL = passthrough */
    int win_in, win_out; /* 0 none, 1 hann, 2 sqrt_hann */
    obuf fft;
    double *frame_out; size_t frame_out_cap;
} filter_prog;
static void apply_window(int kind, double *d, size_t nch, size_t size) {
    if (kind == 1) orc_window_hann(d, nch, size);
    else if (kind == 2) orc_window_sqrt_hann(d, nch, size);
}
/* SURVEY.md §8(d) cfg2 chain, each step one reference helper */
static void filter_frame(orc_sched *s, obuf *in, obuf *out, double t, void *user) {
    (void)t;
    filter_prog *p = (filter_prog *)user;
    apply_window(p->win_in, in->data, in->nch, in->size);
    obuf_resize(&p->fft, in->nch, in->size);
    cplx *f = (cplx *)p->fft.data;
    for (size_t e = 0; e < in->nch * in->size; e++) f[e] = in->data[e];
    orc_frft(p->fft.data, p->fft.data, p->fft.nch, p->fft.size, 1);
    if (p->gain) {
        size_t n = p->fft.size;
        for (size_t c = 0; c < p->fft.nch; c++)
            for (size_t k = 0; k <= n / 2; k++) {
                f[c * n + k] *= p->gain[k];
                if (k && k != n - k) f[c * n + n - k] *= p->gain[k];
            }
    }
    orc_frft(p->fft.data, p->fft.data, p->fft.nch, p->fft.size, -1);
    obuf_resize(out, p->fft.nch, p->fft.size);
    for (size_t e = 0; e < out->nch * out->size; e++) out->data[e] = creal(f[e]);
    apply_window(p->win_out, out->data, out->nch, out->size);
    orc_sched_next_hop_ratio(s, p->n, p->hop_ratio);
    size_t per = out->nch * out->size;
    if (p->frame_out && (s->frames + 1) * per <= p->frame_out_cap)
        memcpy(p->frame_out + s->frames * per, out->data, sizeof(double) * per);
}

typedef struct {
    double *frame_out; size_t frame_out_cap;
    uint32_t n_base; int vary; int stream;
} granular_prog;
/* hacks/pitchshift-simple.cpp:2-11 (no FFT: exercises gather/OLA with varying frame size, hop and
 * out.size != in.size).  vary=0: N_next = n_base as in the hack (1800); vary=1: SURVEY cfg4
 * N_next = 2^(6 + ((f+s) mod 8)). */
static void granular_frame(orc_sched *s, obuf *in, obuf *out, double t, void *user) {
    granular_prog *p = (granular_prog *)user;
    double shift = sin(t) / 2 + 1;
    orc_window_hann(in->data, in->nch, in->size);
    obuf_resize(out, in->nch, (size_t)(in->size / shift));
    for (size_t i = 0; i < in->nch; i++)
        for (size_t j = 0; j < in->size; j++)
            for (int k = (int)(j / shift); k < (j + 1) / shift; k++) out->data[i * out->size + k] = in->data[i * in->size + j];
    uint32_t n = p->vary ? (1u << (6 + ((s->frames + 1 + p->stream) % 8))) : p->n_base;
    orc_sched_next_hop_ratio(s, n, 0.5 / shift);
    (void)p;
}

/* ---- entry points for tests (ctypes) ---- */
typedef struct {
    size_t frames;
    size_t out_len;      /* floats written to out_stream (interleaved) */
} orc_result;

static void attach_tables(orc_sched *s, int64_t *in_off, int64_t *out_off, int64_t *hop, int64_t *win, double *t,
                          size_t cap) {
    s->tab_in_off = in_off; s->tab_out_off = out_off; s->tab_hop = hop; s->tab_win = win; s->tab_t = t;
    s->tab_cap = (in_off && out_off && hop && win && t) ? cap : 0;
}

/* Render `len` sample-frames of interleaved float input through the PV pitch-shift program.
 * out_stream (float, interleaved) receives the reference's float32 OLA stream (tinyspec.cpp:173-200),
 * ola64 (double, interleaved, zero-initialised by the caller) the double OLA of per-frame outputs. */
orc_result orc_render_pv(const float *in, size_t len, size_t nch, uint32_t n, double hop_ratio, double shift,
                         double rate, size_t max_frames, float *out_stream, size_t out_cap, double *ola64,
                         size_t ola64_cap, double *frame_out, size_t frame_out_cap, double *state_dump,
                         size_t state_every, size_t state_cap, double *final_last_phase, double *final_phase_sum,
                         int64_t *tab_in_off, int64_t *tab_out_off, int64_t *tab_hop, int64_t *tab_win,
                         double *tab_t, size_t tab_cap) {
    orc_sched s;
    sched_init(&s, nch, nch, rate);
    s.ainqueue = in; s.ain_len = len * nch;
    s.ola64 = ola64; s.ola64_cap = ola64_cap;
    attach_tables(&s, tab_in_off, tab_out_off, tab_hop, tab_win, tab_t, tab_cap);
    pv_prog p = {0};
    p.pv = orc_pv_new(&s.hop, &s.rate);
    p.n = n; p.hop_ratio = hop_ratio; p.shift = shift;
    p.frame_out = frame_out; p.frame_out_cap = frame_out_cap;
    p.state_dump = state_dump; p.state_every = state_every; p.state_cap = state_cap;
    orc_sched_next_hop_ratio(&s, n, hop_ratio); /* set before the first frame */
    orc_pv_reset(p.pv, 1, 1);                   /* hacks/pitchshift-pv.cpp:11 */
    sched_run(&s, pv_frame, &p, max_frames ? max_frames : (size_t)-1);
    orc_result r = {s.frames, s.aq_len < out_cap ? s.aq_len : out_cap};
    if (out_stream) memcpy(out_stream, s.aqueue, sizeof(float) * r.out_len);
    r.out_len = s.aq_len;
    size_t sz = p.pv->nch * p.pv->size;
    if (final_last_phase && p.pv->last_phase.data) memcpy(final_last_phase, p.pv->last_phase.data, sizeof(double) * sz);
    if (final_phase_sum && p.pv->phase_sum.data) memcpy(final_phase_sum, p.pv->phase_sum.data, sizeof(double) * sz);
    orc_pv_free(p.pv);
    free(s.aqueue);
    sched_free(&s);
    return r;
}

orc_result orc_render_filter(const float *in, size_t len, size_t nch, uint32_t n, double hop_ratio,
                             const double *gain, int win_in, int win_out, double rate, size_t max_frames,
                             float *out_stream, size_t out_cap, double *ola64, size_t ola64_cap,
                             double *frame_out, size_t frame_out_cap) {
    orc_sched s;
    sched_init(&s, nch, nch, rate);
    s.ainqueue = in; s.ain_len = len * nch;
    s.ola64 = ola64; s.ola64_cap = ola64_cap;
    filter_prog p = {0};
    p.n = n; p.hop_ratio = hop_ratio; p.gain = gain; p.win_in = win_in; p.win_out = win_out;
    p.fft.width = 2;
    p.frame_out = frame_out; p.frame_out_cap = frame_out_cap;
    orc_sched_next_hop_ratio(&s, n, hop_ratio);
    sched_run(&s, filter_frame, &p, max_frames ? max_frames : (size_t)-1);
    orc_result r = {s.frames, s.aq_len < out_cap ? s.aq_len : out_cap};
    if (out_stream) memcpy(out_stream, s.aqueue, sizeof(float) * r.out_len);
    r.out_len = s.aq_len;
    free(p.fft.data);
    free(s.aqueue);
    sched_free(&s);
    return r;
}

orc_result orc_render_granular(const float *in, size_t len, size_t nch, uint32_t n_first, uint32_t n_base, int vary,
                               int stream, double rate, size_t max_frames, float *out_stream, size_t out_cap,
                               double *ola64, size_t ola64_cap, int64_t *tab_in_off, int64_t *tab_out_off,
                               int64_t *tab_hop, int64_t *tab_win, double *tab_t, size_t tab_cap) {
    orc_sched s;
    sched_init(&s, nch, nch, rate);
    s.ainqueue = in; s.ain_len = len * nch;
    s.ola64 = ola64; s.ola64_cap = ola64_cap;
    attach_tables(&s, tab_in_off, tab_out_off, tab_hop, tab_win, tab_t, tab_cap);
    granular_prog p = {0};
    p.n_base = n_base; p.vary = vary; p.stream = stream;
    if (n_first) s.new_window_size = n_first; /* otherwise the reference defaults: 4096 / hop 2048 */
    sched_run(&s, granular_frame, &p, max_frames ? max_frames : (size_t)-1);
    orc_result r = {s.frames, s.aq_len < out_cap ? s.aq_len : out_cap};
    if (out_stream) memcpy(out_stream, s.aqueue, sizeof(float) * r.out_len);
    r.out_len = s.aq_len;
    free(s.aqueue);
    sched_free(&s);
    return r;
}

/* Hop clock only (rtlib/control.hpp:22-29 + tinyspec.cpp:110-112,163-168): run `frames` iterations
 * with a constant requested hop and report computed_hop and t for each. */
void orc_hop_clock(double hop, double rate, size_t frames, int64_t *computed_hop, double *t) {
    orc_sched s;
    sched_init(&s, 0, 0, rate);
    orc_sched_set_hop(&s, hop);
    for (size_t f = 0; f < frames; f++) {
        t[f] = sched_time(&s);
        if (s.time_fract + s.hop_fract < s.time_fract) s.computed_hop = s.hop_samples + 1;
        else s.computed_hop = s.hop_samples;
        s.time_samples += s.computed_hop;
        s.time_fract += s.hop_fract;
        computed_hop[f] = (int64_t)s.computed_hop;
    }
}
