/* tsb200.h — C ABI of libtsb200, the B200 (sm_100a) backend for tinyspec-cling's overlap-add
 * spectral hot path (window -> FFT -> phase vocoder / spectral gain -> inverse FFT -> overlap-add).
 *
 * The reference (nwoeanhinnogaehr/tinyspec-cling) has no binary plugin interface: a "hack" is C++
 * source compiled against rtinc/root.hpp and run frame by frame by tinyspec.cpp's scheduler
 * (SURVEY.md §8b).  This header is the boundary a binding of that interface calls into; the
 * source-compatible C++ facade (tinyspec-cling_b200/rtinc/) and the Python host mirror
 * (tinyspec-cling_b200/api.py) are both thin layers over exactly these entry points.  Each entry
 * point cites the reference code it replaces.
 *
 * Conventions: every function returns 0 on success or a negative TSB_E* code; tsb_last_error()
 * gives the message of the calling thread's last failure.  Nothing here falls back to the CPU: with
 * no usable CUDA device every device entry point fails with TSB_ECUDA.
 * A context is bound to one device and one CUDA stream and is not thread-safe (the reference runs
 * all frames on one worker thread, tinyspec.cpp:284).
 *
 * Lanes: a lane is one channel of one stream.  Channels never interact anywhere on the path
 * (rtlib/pv.hpp loops, fftw howmany, tinyspec.cpp:173-200), so S streams x C channels are S*C lanes,
 * lane = stream*C + channel, each a planar row of samples.
 */
#ifndef TSB200_H
#define TSB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TSB_ABI_VERSION 3

enum { TSB_OK = 0, TSB_EINVAL = -1, TSB_ECUDA = -2, TSB_EUNSUPPORTED = -3, TSB_ENOMEM = -4 };
/* arithmetic of the chain (phase accumulators are fp64 in all three).  TSB_F64_OUT32 computes like TSB_F64 and
 * delivers the reference's own output type: a float32 stream whose overlap tail is kept in float32 and combined
 * exactly as tinyspec.cpp:173-200 does (emitted sample = float(double(tail) + value), tail += float(value)), so it
 * is comparable bit for bit with what the reference hands to JACK and moves half the bytes of TSB_F64. */
enum { TSB_F32 = 0, TSB_F64 = 1, TSB_F64_OUT32 = 2 };
enum { TSB_PROG_PV_SHIFT = 0, TSB_PROG_GAIN = 1 };
enum { TSB_WIN_NONE = 0, TSB_WIN_HANN = 1, TSB_WIN_SQRT_HANN = 2 }; /* rtlib/windowfn.hpp:1-14 */

typedef struct tsb_ctx tsb_ctx;
typedef struct tsb_plan tsb_plan;

int tsb_abi_version(void);
const char *tsb_last_error(void);

/* ---- context ------------------------------------------------------------------------------- */
/* new_stream == 0: run on `cuda_stream`, a cudaStream_t of the caller (NULL = the default stream);
 * new_stream != 0: create and own a non-blocking stream (cuda_stream is ignored). */
int tsb_ctx_create(int device, void *cuda_stream, int new_stream, tsb_ctx **ctx);
int tsb_ctx_destroy(tsb_ctx *ctx);
int tsb_ctx_sync(tsb_ctx *ctx);
/* The context keeps device scratch between calls (staging rows of tsb_chain_render_host, the time-parallel
 * render's segment state, frft's work buffer): this waits for the stream and releases all of it. */
int tsb_ctx_trim(tsb_ctx *ctx);
/* device memory for bindings that have no CUDA runtime of their own (cgo / JNI / ctypes) */
int tsb_malloc(tsb_ctx *ctx, size_t bytes, void **dptr);
int tsb_free(tsb_ctx *ctx, void *dptr);
int tsb_memset(tsb_ctx *ctx, void *dptr, int value, size_t bytes);
int tsb_upload(tsb_ctx *ctx, void *dptr, const void *host, size_t bytes);   /* async on the ctx stream */
int tsb_download(tsb_ctx *ctx, void *host, const void *dptr, size_t bytes); /* async on the ctx stream */

/* ---- host-side control logic (no device needed) --------------------------------------------- */
/* set_hop(): rtlib/control.hpp:22-29.  Returns TSB_EINVAL for a negative hop ("ignoring invalid hop"). */
int tsb_hop_fixed(double hop, uint64_t *hop_samples, uint64_t *hop_fract);

/* Frame schedule of a render in which every frame asks for next_hop_samples(n, hop), set before the
 * first frame: restates tinyspec.cpp:135-149 (latch, starve test, advance by the PREVIOUS hop; the
 * first advance is the initial computed_hop 2048), :110-112 (t) and :163-168 (fractional-hop carry).
 * Tables may be NULL; returns the number of frames (<= cap when tables are given) in *n_frames.
 *   in_off[f]  first input sample of frame f (I_f)        out_off[f] output position of frame f (P_f)
 *   hop[f]     samples emitted after frame f              t[f]       time handed to the callback   */
int tsb_schedule_const(int64_t in_len, uint32_t n, double hop, double rate, int64_t max_frames, int64_t cap,
                       int64_t *n_frames, int64_t *in_off, int64_t *out_off, int32_t *hop_tab, double *t);

/* PhaseVocoder::shift bin remap (rtlib/pv.hpp:59-61,70-78) for fn values on the n/2 bin centres:
 * dest[j] = destination bin of source bin j, or -1 if it is dropped. */
int tsb_shift_map(uint32_t n, double rate, const double *fn_bins, int32_t *dest);

/* ---- plans ---------------------------------------------------------------------------------- */
/* A plan fixes what the reference's helpers recompute on every frame: window tables
 * (rtlib/windowfn.hpp), FFT plan/twiddles (fftw_plan_many_dft, rtlib/fft.hpp:13-16), the shift map
 * and the per-bin phase constants of PhaseVocoder (rtlib/pv.hpp:39-48). */
typedef struct tsb_plan_desc {
    uint32_t n;          /* frame size: power of two, 64..8192 */
    int32_t precision;   /* TSB_F32 | TSB_F64 | TSB_F64_OUT32 */
    int32_t program;     /* TSB_PROG_PV_SHIFT: pv.analyze(in); pv.shift(fn); pv.synthesize(out)  (hacks/pitchshift-pv.cpp:5-7)
                            TSB_PROG_GAIN:     window(in); frft(+1); spectrum *= gain; frft(-1); real; window(out) */
    int32_t win_in;      /* TSB_PROG_GAIN only (PV always uses sqrt-Hann twice, rtlib/pv.hpp:5,15) */
    int32_t win_out;
    double hop;          /* PhaseVocoder::hop() (rtinc/pv.hpp:50-53); may be fractional */
    double rate;         /* RATE */
    double ratio;        /* PV: fn(x) = ratio*x applied to the bin frequencies (rtlib/pv.hpp:66) */
    const double *fn_bins; /* PV: fn(bin centre), n/2 values, or NULL for centre*ratio (rtlib/pv.hpp:65) */
    const double *gain;  /* GAIN: n/2+1 real gains for bins k and (n-k)%n, or NULL for all ones */
} tsb_plan_desc;
int tsb_plan_create(tsb_ctx *ctx, const tsb_plan_desc *desc, tsb_plan **plan);
int tsb_plan_destroy(tsb_plan *plan);

/* ---- the fused chain ------------------------------------------------------------------------ */
/* Runs `n_frames` consecutive frames of every lane: replaces tinyspec.cpp:146-157 (gather),
 * the hack's helper chain (rtlib/pv.hpp:4-79, rtlib/fft.hpp:3-42, rtlib/windowfn.hpp) and
 * tinyspec.cpp:170-200 (overlap-add).  All pointers are DEVICE pointers; the call is asynchronous on
 * the context's stream.
 *   in        [lanes][in_stride]  float32 (the reference's queues are float, tinyspec.cpp:43)
 *   out       [lanes][out_stride] float32 (TSB_F32, TSB_F64_OUT32) or float64 (TSB_F64): position p of a lane
 *             receives the finished overlap-added sample p; n_frames frames finish sum(hop) samples
 *   hop_tab   per-frame emitted samples (device, int32) or NULL for the constant `hop`
 *   state_*   float64 [lanes][n/2], [lanes][n/2], [lanes][n]: PhaseVocoder::last_phase / phase_sum
 *             (live bins 0..n/2-1) and the overlap tail relative to the next output position.
 *             load_state = 0 starts from PhaseVocoder::reset() and an empty tail; the arrays (may be
 *             NULL) receive the state after the last frame, so a render can be continued, split into
 *             time segments, or teacher-forced.
 *   prev_hop  load_state = 1 only: computed_hop of the frame before the first one of this call (its samples
 *             start at in_off0 - prev_hop of the same rows), 0 = not known.  Only TSB_F32 reads it: a wrap of
 *             phase_to_frequency (rtlib/pv.hpp:41-43) that the float32 transform cannot decide safely is
 *             re-evaluated in double precision from the samples of the frame and of the one before it; with
 *             prev_hop = 0 the first frame of a continued render keeps the float32 decision. */
typedef struct tsb_chain_args {
    const float *in;
    int64_t in_stride, in_len;
    void *out;
    int64_t out_stride, out_len;
    int64_t n_frames;
    int64_t in_off0, out_off0;
    int32_t hop;
    const int32_t *hop_tab;
    double *state_last_phase, *state_phase_sum, *state_ola;
    int32_t load_state;
    int32_t lanes;
    int32_t prev_hop;
} tsb_chain_args;
int tsb_chain_render(tsb_ctx *ctx, tsb_plan *plan, const tsb_chain_args *args);

/* ---- fused spectral synthesis (BASELINE cfg5; the reference's hacks/spectral_bytebeat hacks, e.g. voidbubbles.cpp:4-19):
 * per frame and channel a Hermitian spectrum of 2*frame points -> frft(., -1) (rtlib/fft.hpp:3-42) -> the first
 * `frame` real samples times the hack's triangular window 1 - |1 - i/(frame/2)| -> overlap-add at `hop`
 * (tinyspec.cpp:170-200), in double precision, as two kernels per frame (four-step inverse real FFT, the
 * intermediate staying in L2).  The spectrum comes from the device:
 *   TSB_GEN_HASH    the seeded integer formula of SURVEY.md §8d / examples/bytebeat.cpp, evaluated in the kernel:
 *                   h = frame*2654435761 ^ (c*40503 + i*2246822519); h ^= h>>15; h *= 2246822519; h ^= h>>13;
 *                   X[i] = (((h & 0xffff) - 32768) + i*((h >> 16) - 32768)) / 32768.0 / (1 + (i >> 4)),  0 < i < frame
 *   TSB_GEN_BUFFER  values the caller computed: device complex128 [n_frames][channels][frame], bins 1..frame-1 read
 *   TSB_GEN_EXPR    a spectrum EXPRESSION evaluated per (frame, channel, bin) in the kernel (the
 *                   per-bin host loop of a spectral_bytebeat hack — e.g. voidbubbles.cpp:9-13,
 *                   `r = (int)(t*32)*n+i; buf[c][i] = sin(2.0*M_PI/64.*(c+((int)pow(i,1.1))&r/3|r/250|r/136|r>>16))
 *                   /pow(i+1, 0.7)*32` — as a postfix program).  The program works on a stack of 64-bit slots (int64
 *                   or double, statically typed: tsb_spectral_render checks the program before launching) and must
 *                   leave exactly two doubles: re and im of X[i].  What depends on the bin only (pow(i+1, 0.7),
 *                   (int)pow(i,1.1)) or on the frame only ((int)(t*32)*n) is evaluated by the CALLER on the host with
 *                   the reference's own libm and handed over as per-bin tables / per-frame parameters, so that only the
 *                   integer arithmetic (exact) and sin / cos / exp / fmod / pow of the coupled part run on the device.
 *                   Integers are int64 (C's 32-bit `int` wrap-around: TSB_X_I_WRAP32); x/0 and x%0 give 0.
 * In all three X[2*frame - i] = conj(X[i]).  Bins 0 and `frame`, which such a hack never writes, hold what its static buffer
 * holds there: the previous frame's output samples y[0] and y[frame] (frft runs in place) — zero before the first
 * frame; state_edge [channels][2] carries them between calls.  Frame f is ADDED to out[c][out_off0 + f*hop ...]
 * (the caller zeroes `out`; after the call positions below out_off0 + n_frames*hop are final, the rest is the tail). */
enum { TSB_GEN_HASH = 0, TSB_GEN_BUFFER = 1, TSB_GEN_EXPR = 2 };
/* one program word = opcode | (argument << 8) */
enum {
    TSB_X_I_BIN = 1,   /* push i (the bin, 0 < i < frame) */
    TSB_X_I_CH,        /* push c (the channel) */
    TSB_X_I_FRAME,     /* push first_frame + f (the frame counter) */
    TSB_X_I_N,         /* push `frame` (the hack's n = out.size) */
    TSB_X_I_CONST,     /* push iconst[arg] */
    TSB_X_I_TAB,       /* push itab[arg][i] */
    TSB_X_I_PARAM,     /* push iparam[f][arg] */
    TSB_X_D_CONST,     /* push dconst[arg] */
    TSB_X_D_TAB,       /* push dtab[arg][i] */
    TSB_X_D_PARAM,     /* push dparam[f][arg] */
    TSB_X_I_ADD, TSB_X_I_SUB, TSB_X_I_MUL, TSB_X_I_DIV, TSB_X_I_MOD, /* (a b -- a op b), C truncating / and % */
    TSB_X_I_AND, TSB_X_I_OR, TSB_X_I_XOR, TSB_X_I_SHL, TSB_X_I_SHR,  /* >> is arithmetic */
    TSB_X_I_EQ, TSB_X_I_LT,                                          /* (a b -- 0 | 1) */
    TSB_X_I_NEG, TSB_X_I_NOT, TSB_X_I_WRAP32,                        /* (a -- op a); WRAP32 = (int32_t)a */
    TSB_X_I2D,         /* (int -- double) */
    TSB_X_D2I,         /* (double -- int), C truncation */
    TSB_X_D_ADD, TSB_X_D_SUB, TSB_X_D_MUL, TSB_X_D_DIV, TSB_X_D_FMOD, TSB_X_D_POW,   /* (a b -- a op b) */
    TSB_X_D_NEG, TSB_X_D_ABS, TSB_X_D_SIN, TSB_X_D_COS, TSB_X_D_EXP, TSB_X_D_SQRT, TSB_X_D_FLOOR,
    TSB_X_PICK,        /* push a copy of the slot `arg` below the top (0 = dup) */
    TSB_X_POLAR,       /* (mag phase -- mag*cos(phase) mag*sin(phase)): std::polar */
    TSB_X__COUNT
};
#define TSB_EXPR_MAX_OPS 128
#define TSB_EXPR_MAX_CONST 24
#define TSB_EXPR_MAX_TABS 4
#define TSB_EXPR_STACK 16
typedef struct tsb_expr {        /* every pointer is a HOST pointer; copied during the call */
    int32_t n_ops;               /* <= TSB_EXPR_MAX_OPS */
    const int32_t *ops;
    int32_t n_iconst, n_dconst;  /* <= TSB_EXPR_MAX_CONST each */
    const int64_t *iconst;
    const double *dconst;
    int32_t n_itab, n_dtab;      /* <= TSB_EXPR_MAX_TABS each; per-bin tables [n_tab][frame] */
    const int32_t *itab;
    const double *dtab;
    int32_t n_iparam, n_dparam;  /* <= TSB_EXPR_MAX_TABS each; per-frame scalars [n_frames][n_param] */
    const int64_t *iparam;
    const double *dparam;
} tsb_expr;
typedef struct tsb_spectral_args {
    uint32_t frame;          /* power of two, 256 .. 262144; the inverse FFT has 2*frame points */
    int32_t channels;
    int64_t n_frames;
    int64_t first_frame;     /* frame counter of the first frame (TSB_GEN_HASH) */
    int32_t hop;             /* computed_hop, constant */
    int32_t generator;
    const double *spectrum;  /* TSB_GEN_BUFFER */
    double *out;             /* device float64 [channels][out_stride] */
    int64_t out_stride, out_len, out_off0;
    double *state_edge;      /* device float64 [channels][2] or NULL, in (load_state) / out */
    int32_t load_state;
    const tsb_expr *expr;    /* TSB_GEN_EXPR (ABI 3) */
} tsb_spectral_args;
int tsb_spectral_render(tsb_ctx *ctx, const tsb_spectral_args *args);
/* Host-side (no device needed): the static check tsb_spectral_render applies to an expression — known opcodes, indices
 * in range, stack depth <= TSB_EXPR_STACK, operand types, exactly (re, im) left.  TSB_OK or TSB_EINVAL + tsb_last_error(). */
int tsb_expr_check(const tsb_expr *expr);

/* ---- batched granular pitch shift with a frame size, hop and output length that change every frame (BASELINE cfg4;
 * the reference's hacks/pitchshift-simple.cpp:1-11 on the scheduler's gather and float32 overlap-add,
 * tinyspec.cpp:135-157,170-200): per frame window_hann(in), out.resize(in.size/shift), the hack's nearest-neighbour
 * resample `for j: for (k = j/shift; k < (j+1)/shift; k++) out[k] = in[j]`, overlap-add at computed_hop.  One launch
 * renders every lane (lane = stream*channels + channel) of every stream; the output is the reference's float32 stream
 * bit for bit.  The per-frame tables are HOST arrays [streams][frames_max], filled by running the hack's control
 * logic (it depends on t only) through the reference's clock — tinyspec-cling_b200/api.py: granular_schedule does it
 * for a Python callback: win = frame size, in_off / out_off = I_f / P_f, hop = computed_hop after
 * the callback, osize = out.size after the callback, shift = the hack's ratio.  Frame sizes up to 8192, osize up to
 * 16384. */
typedef struct tsb_granular_args {
    const float *in;          /* device float32 [lanes][in_stride] */
    int64_t in_stride, in_len;
    float *out;               /* device float32 [lanes][out_stride] */
    int64_t out_stride, out_len;
    int32_t streams, channels;
    int64_t frames_max;
    const int64_t *n_frames;  /* host [streams] */
    const int32_t *win, *hop, *osize;        /* host [streams][frames_max] */
    const int64_t *in_off, *out_off;
    const double *shift;
} tsb_granular_args;
int tsb_granular_render(tsb_ctx *ctx, const tsb_granular_args *args);

/* Measurement aid: the device's measured non-tensor FMA rate in TFLOP/s (an FMA = 2 flops), fp64 or fp32, from a
 * kernel of independent FMA chains timed with CUDA events.  bench.py reports the chain kernel's arithmetic against it. */
int tsb_probe_fma_peak(tsb_ctx *ctx, int32_t is_f64, double *tflops);

/* Same, with HOST buffers (in, out, hop_tab, state_* are host pointers): returns when the result is in `out`.
 * Large batches are cut into lane chunks whose upload, chain kernel and download overlap (separate copy engines
 * per direction, one stream per chunk); per-lane results are bit-identical to tsb_chain_render.  Pinned host
 * buffers make the copies asynchronous; pageable ones work too.  Whole rows are copied, so `out` rows come back
 * as the device computed them: zero outside the emitted range. */
int tsb_chain_render_host(tsb_ctx *ctx, tsb_plan *plan, const tsb_chain_args *args);

/* ---- time-parallel rendering ------------------------------------------------------------------ */
/* The only things a frame needs from earlier frames of its lane are last_phase (the previous frame's phases),
 * phase_sum (a running sum of per-frame increments) and the overlap tail.  A render of n_frames frames is cut
 * into time segments that run concurrently (one CTA per segment and lane):
 *   pre-pass  every segment analyses its frames (plus one warm-up frame before it, to obtain last_phase) and
 *             stores the increment each frame adds to each bin's phase_sum — no synthesis, no output;
 *   scan      per bin, the stored increments are added one frame at a time in frame order, from the carry-in:
 *             exactly the additions of the sequential render, so every segment starts from the very phase_sum the
 *             sequential render has there (summing per-segment totals first would regroup them and drift);
 *   main      the full chain per segment from that state.  A segment first replays, without emitting anything, the
 *             frames whose output overlaps its first own frame (n/hop - 1 of them for a constant hop; the pre-pass
 *             and the scan are shifted by as many frames): that rebuilds the overlap tail it starts with, with the
 *             sequential render's own arithmetic.
 * Samples, phase_sum / last_phase and the final overlap tail are bit-identical to tsb_chain_render's, in every
 * precision (the float32 tail arithmetic of TSB_F64_OUT32 included).  Cost: the lead-in frames are rendered twice.
 * This is what lets ONE long stream fill a GPU, and what a multi-GPU time split is built from: rank g renders a
 * block of frames that begins with its own `lead` lead-in frames — the frames of the block before whose output
 * reaches into its first own frame — and one warm-up frame before those (`warmup` = 1, args->in_off0 - warm_hop
 * >= 0; args->in_off0 / out_off0 / n_frames / hop_tab describe the block INCLUDING the lead-in frames, which emit
 * nothing).  It pre-passes the block, takes the running phase_sum before its first lead-in frame from rank g-1
 * (tsb_timepar_scan_at: carry in, carry out at a frame of its choice — a chain of short steps over the ranks) and
 * renders: the lead-in frames rebuild the overlap tail the block starts with, so nothing but that carry crosses a
 * block boundary and the stitched output is the one-GPU render bit for bit.  (With lead = 0 a block's first
 * samples lack the previous block's tail: tsb_add_tail adds another block's final state_ola, ~1 ulp from the
 * sequential order.)  Scratch: lanes * n_frames * n/2 doubles for the stored increments. */
typedef struct tsb_timepar tsb_timepar;
/* args is copied (device pointers must stay valid); host_hop_tab = host copy of args->hop_tab (NULL iff hop_tab is NULL) */
int tsb_timepar_create(tsb_ctx *ctx, tsb_plan *plan, const tsb_chain_args *args, const int32_t *host_hop_tab,
                       int32_t n_segments, int32_t warmup, int32_t warm_hop, int32_t lead, tsb_timepar **tp);
int tsb_timepar_segments(tsb_timepar *tp);   /* segments actually used (each must emit at least n samples) */
/* block_total (the block's increments summed from zero) / block_last_phase: device [lanes][n/2] float64, may be NULL */
int tsb_timepar_prepass(tsb_timepar *tp, double *block_total, double *block_last_phase);
/* carry_in: device [lanes][n/2] phase_sum before the block's first frame (NULL: args->state_phase_sum, or zeros);
 * carry_out (may be NULL) receives the phase_sum after its last frame.  Optional: tsb_timepar_render runs it with
 * carry_in = NULL when it was not called. */
int tsb_timepar_scan(tsb_timepar *tp, const double *carry_in, double *carry_out);
/* The same for lanes lane0 .. lane0+n_lanes-1 only (carry_in / carry_out: [n_lanes][n/2]): lanes are independent, so
 * a chain of ranks can hand the sums on lane by lane and overlap their scans.  The caller covers every lane. */
int tsb_timepar_scan_lanes(tsb_timepar *tp, int32_t lane0, int32_t n_lanes, const double *carry_in, double *carry_out);
/* The same with carry_out = the phase_sum BEFORE frame `carry_frame` of the block (0 <= carry_frame < n_frames; the
 * first lead-in frame of the next rank's block), instead of after its last frame (carry_frame < 0). */
int tsb_timepar_scan_at(tsb_timepar *tp, int32_t lane0, int32_t n_lanes, const double *carry_in, double *carry_out,
                        int64_t carry_frame);
/* The scan FUSED with the hand-over to the next GPU of a time-sharded render (one process per GPU on an NVLink box):
 * carry_in / flag_in (NULL on the first rank) are buffers in this GPU's memory — [lanes][n/2] float64 and one int32
 * per 16 bins of a lane — that the previous rank's kernel fills over NVLink; carry_out_peer / flag_out_peer (NULL
 * on the last rank) are the next rank's buffers mapped into this process (tsb_ipc_open).  Each CTA waits until its
 * own flag equals `epoch` (acquire, system scope), adds its 16 bins' increments frame by frame, writes the sums
 * before frame `carry_frame` to the peer and releases the peer's flag: no host involvement, no NCCL call, and the
 * ranks' scans overlap bin block by bin block.  Use a new epoch (and, to be safe against a rank running ahead,
 * alternate two buffer sets) for every render. */
int tsb_timepar_scan_chain(tsb_timepar *tp, const double *carry_in, const int32_t *flag_in, double *carry_out_peer,
                           int32_t *flag_out_peer, int32_t epoch, int64_t carry_frame);
/* Peer memory between the processes of one box: a 64-byte handle for an allocation made with tsb_malloc
 * (cudaIpcGetMemHandle), mapping a peer's handle into this process (peer access is enabled on demand), unmapping it. */
int tsb_ipc_export(tsb_ctx *ctx, void *dptr, void *handle64);
int tsb_ipc_open(tsb_ctx *ctx, const void *handle64, void **dptr);
int tsb_ipc_close(tsb_ctx *ctx, void *dptr);
/* initial state = args->state_* (NULL or load_state 0: reset state); they receive the final state */
int tsb_timepar_render(tsb_timepar *tp);
int tsb_timepar_destroy(tsb_timepar *tp);
/* create + prepass + render + destroy */
int tsb_chain_render_timepar(tsb_ctx *ctx, tsb_plan *plan, const tsb_chain_args *args, const int32_t *host_hop_tab,
                             int32_t n_segments);
/* out[l][pos + i] += tail[l][i], i < n (tail: device float64 [lanes][n], e.g. another block's final state_ola) */
int tsb_add_tail(tsb_ctx *ctx, tsb_plan *plan, void *out, int64_t out_stride, int64_t out_len, int64_t pos,
                 const double *tail, int64_t lanes);

/* ---- chunked streaming front end ---------------------------------------------------------------- */
/* What tinyspec.cpp's JACK callback + worker do for a live stream (tinyspec.cpp:205-239 feeding :113-203), without
 * the threads: every call appends `len` new input samples per lane (one audio period) and returns the output
 * samples that became final, for a plan rendered with next_hop_samples(n, hop) on every frame.  The first frame
 * needs 2048 + n samples, so early calls return nothing — that latency is the reference's.
 * Frames are scheduled exactly as tsb_schedule_const does for the concatenated input, and the PV state and
 * overlap tail stay on the device between calls: any chunking gives bit-identical output to one whole render. */
typedef struct tsb_stream tsb_stream;
int tsb_stream_create(tsb_ctx *ctx, tsb_plan *plan, int32_t lanes, double hop, double rate, tsb_stream **stream);
/* host_in: [lanes][in_stride] float32, `len` valid samples per row; host_out: [lanes][out_stride] float32/float64
 * (the plan's precision) receiving *out_len samples per row (never more than out_cap: size it >= len + n). */
int tsb_stream_process(tsb_stream *stream, const float *host_in, int64_t in_stride, int64_t len, void *host_out,
                       int64_t out_stride, int64_t out_cap, int64_t *out_len);
int64_t tsb_stream_frames(tsb_stream *stream);  /* frames rendered so far */
int tsb_stream_destroy(tsb_stream *stream);

/* ---- the helpers one by one ("immediate mode") ------------------------------------------------ */
/* One call = one helper of the reference's rtinc/ library, on DEVICE buffers laid out like Buf<T>
 * (planar [channel][size]; FFTBuf interleaved re,im; PVBuf interleaved amp,freq), double precision,
 * asynchronous on the context's stream.  They serve hacks that mix helper calls with their own loops;
 * chains made only of helpers are fused (tsb_chain_render). */

/* window_hann / window_sqrt_hann (rtlib/windowfn.hpp:1-14): data[c][i] *= w[i], in place. */
int tsb_window(tsb_ctx *ctx, double *data, int64_t nch, int64_t size, int32_t kind);

/* What FFTW does for the reference (fftw_plan_many_dft(FFTW_FORWARD), rtlib/fft.hpp:13-17): `rows` forward
 * DFTs of length n, kernel exp(-2*pi*i*j*k/n), unnormalised, any n >= 1, in place allowed. */
int tsb_fft_forward(tsb_ctx *ctx, const double *in, double *out, int64_t rows, int64_t n);

/* frft(in, out, exponent) (rtlib/fft.hpp:3-42), any size, any exponent; in == out reproduces the reference's
 * in-place aliasing.  nch*n == 0 returns silently (fft.hpp:8-9). */
int tsb_frft(tsb_ctx *ctx, double *in, double *out, int64_t nch, int64_t n, double exponent);

/* PhaseVocoder::analyze(FFTBuf&) (rtlib/pv.hpp:17-28,39-45): all `size` bins of every channel.
 * last_phase [nch][size] is read and updated, pvdata [nch][size][2] receives (amp, freq). */
int tsb_pv_analyze(tsb_ctx *ctx, const double *spec, double *last_phase, double *pvdata, int64_t nch, int64_t size,
                   double hop, double rate);
/* PhaseVocoder::synthesize(FFTBuf&) (rtlib/pv.hpp:29-38,46-48): phase_sum += 2*pi*freq/RATE*hop; spec = polar. */
int tsb_pv_synthesize(tsb_ctx *ctx, double *spec, double *phase_sum, const double *pvdata, int64_t nch, int64_t size,
                      double hop, double rate);
/* PhaseVocoder::shift (rtlib/pv.hpp:56-79).  The caller evaluates the user's fn on the host — it is
 * arbitrary user code — and passes fn_bins[c][j] = fn(j/size*RATE) and fn_freqs[c][j] = fn(data[c][j].freq),
 * j < size/2 (HOST arrays; tsb_pv_freqs fetches the frequencies).  The device zeroes and scatters. */
int tsb_pv_freqs(tsb_ctx *ctx, const double *pvdata, int64_t nch, int64_t size, double *host_freqs);
int tsb_pv_shift(tsb_ctx *ctx, double *pvdata, int64_t nch, int64_t size, double rate, const double *host_fn_bins,
                 const double *host_fn_freqs);

/* Frame gather (tinyspec.cpp:151-157): frame[l][i] = (double) in[l][in_off + i], i < w, zero past in_len. */
int tsb_gather_frame(tsb_ctx *ctx, const float *in, int64_t in_stride, int64_t in_len, int64_t in_off, double *frame,
                     int64_t lanes, int64_t w);
/* Overlap-add of one frame exactly as tinyspec.cpp:170-200 does it, float32 tail and float32 narrowing
 * included: emits `hop` samples at out[l][out_off ..], keeps max(tail_len, osize) - hop samples of overlap.
 * tail_in / tail_out are distinct float32 [lanes][tail_cap] buffers (swap them frame to frame); tail_len is the
 * overlap length before this frame; the new length is returned in *new_tail_len. */
int tsb_overlap_add(tsb_ctx *ctx, const double *frame, int64_t lanes, int64_t osize, int64_t hop, float *out,
                    int64_t out_stride, int64_t out_len, int64_t out_off, const float *tail_in, float *tail_out,
                    int64_t tail_cap, int64_t tail_len, int64_t *new_tail_len);

/* Number of kernel launches issued through this context so far (for benchmarks' bookkeeping). */
int64_t tsb_launch_count(tsb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
