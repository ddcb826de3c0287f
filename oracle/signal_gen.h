/* oracle/signal_gen.h — TEST INFRASTRUCTURE (not product code).
 *
 * The synthetic input of SURVEY.md §8(d): per stream s, channel c, sample n
 *     x = 0.5*sin(2*pi*f*n/rate) + 0.1*(u - 0.5),   f = 440 + 110*c + 0.37*s   (never bin-centred)
 * u from the 32-bit LCG state = state*1664525 + 1013904223, seed 12345 + s, u = (state>>8)/2^24,
 * one draw per (n, c) in n-major / c-minor order.  Values are rounded to float32 because the
 * reference's input queue is a deque<float> (tinyspec.cpp:43,214-216).
 *
 * One C definition shared by the fake JACK driver (oracle/ref/fake_jack.cpp) and the C oracle
 * (oracle/oracle.c) so both see bit-identical input (same libm sin).
 */
#ifndef TS_ORACLE_SIGNAL_GEN_H
#define TS_ORACLE_SIGNAL_GEN_H
#include <math.h>
#include <stdint.h>
#include <stddef.h>

typedef struct {
    uint32_t lcg;
    uint64_t n;      /* next sample index */
    int nch;
    int stream;
    double rate;
} ts_siggen;

static inline void ts_siggen_init(ts_siggen *g, int stream, int nch, double rate) {
    g->lcg = 12345u + (uint32_t)stream;
    g->n = 0;
    g->nch = nch;
    g->stream = stream;
    g->rate = rate;
}

/* Produce `count` sample frames, interleaved [n][c] float32. */
static inline void ts_siggen_fill(ts_siggen *g, float *dst, size_t count) {
    for (size_t i = 0; i < count; i++) {
        for (int c = 0; c < g->nch; c++) {
            g->lcg = g->lcg * 1664525u + 1013904223u;
            double u = (double)(g->lcg >> 8) / 16777216.0;
            double f = 440.0 + 110.0 * c + 0.37 * g->stream;
            double x = 0.5 * sin(2.0 * M_PI * f * (double)g->n / g->rate) + 0.1 * (u - 0.5);
            dst[i * g->nch + c] = (float)x;
        }
        g->n++;
    }
}
#endif
