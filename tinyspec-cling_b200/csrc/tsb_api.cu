// tsb_api.cu — the C ABI of libtsb200 (include/tsb200.h): contexts, plans, launches.
#include <cuda_runtime.h>

#include <cstdarg>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tsb200.h"
#include "tsb_launch.hpp"
#include "tsb_ops.hpp"
#include "tsb_plan.hpp"
#include "tsb_spectral.hpp"
#include "tsb_granular.hpp"
#include <map>

using namespace tsb;

#define TSB_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(TSB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));   \
    } while (0)

struct tsb_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int64_t launches;
    std::map<std::pair<int64_t, int>, double *> windows;  // (size, kind) -> device table
    double *scratch = nullptr;                            // frft's FFT output / shift's copy
    size_t scratch_bytes = 0;
    char *tp_pool = nullptr;                              // scratch of the time-parallel render, reused call to call
    size_t tp_pool_bytes = 0;
    bool tp_pool_busy = false;
    // staging of the host-buffer render, kept call to call (allocating and freeing tens of GB per call cost
    // 100-400 ms on the test boxes, more than the render itself); tsb_ctx_trim releases it
    void *host_stage[2] = {nullptr, nullptr};             // device copies of the caller's input / output rows
    size_t host_stage_bytes[2] = {0, 0};
    std::vector<cudaStream_t> host_streams;               // [0] uploads, [1] downloads, [2..] one per lane chunk
};

struct tsb_plan {
    tsb_ctx *ctx;
    tsb_plan_desc desc;
    int log2n;
    std::vector<void *> dev;  // device allocations owned by the plan
    ChainTables<float> t32;
    ChainTables<double> t64;
};

TSB_API int tsb_abi_version(void) { return TSB_ABI_VERSION; }
TSB_API const char *tsb_last_error(void) { return g_err.c_str(); }

TSB_API int tsb_ctx_create(int device, void *cuda_stream, int new_stream, tsb_ctx **out) {
    if (!out) return fail(TSB_EINVAL, "tsb_ctx_create: null out pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TSB_ECUDA, "no usable CUDA device (%s); libtsb200 has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(TSB_EINVAL, "device %d out of range (0..%d)", device, count - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TSB_EUNSUPPORTED, "device %d is sm_%d%d; libtsb200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    tsb_ctx *c = new tsb_ctx();
    c->device = device;
    c->stream = nullptr;
    c->own_stream = false;
    c->launches = 0;
    if (!new_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        cudaError_t e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e2 != cudaSuccess) {
            delete c;
            return fail(TSB_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e2));
        }
        c->own_stream = true;
    }
    *out = c;
    return TSB_OK;
}
TSB_API int tsb_ctx_destroy(tsb_ctx *c) {
    if (!c) return TSB_OK;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    for (auto &kv : c->windows) cudaFree(kv.second);
    if (c->scratch) cudaFree(c->scratch);
    if (c->tp_pool) cudaFree(c->tp_pool);
    for (void *q : c->host_stage)
        if (q) cudaFree(q);
    for (cudaStream_t st : c->host_streams) cudaStreamDestroy(st);
    delete c;
    return TSB_OK;
}
TSB_API int tsb_ctx_trim(tsb_ctx *c) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 2; i++) {
        if (c->host_stage[i]) cudaFree(c->host_stage[i]);
        c->host_stage[i] = nullptr;
        c->host_stage_bytes[i] = 0;
    }
    if (c->tp_pool && !c->tp_pool_busy) {
        cudaFree(c->tp_pool);
        c->tp_pool = nullptr;
        c->tp_pool_bytes = 0;
    }
    if (c->scratch) cudaFree(c->scratch);
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    op_release_scratch(c->device);
    op_spectral_release(c->device);
    return TSB_OK;
}
TSB_API int tsb_ctx_sync(tsb_ctx *c) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaStreamSynchronize(c->stream));
    return TSB_OK;
}
TSB_API int64_t tsb_launch_count(tsb_ctx *c) { return c ? c->launches : 0; }

TSB_API int tsb_malloc(tsb_ctx *c, size_t bytes, void **dptr) {
    if (!c || !dptr) return fail(TSB_EINVAL, "tsb_malloc: null argument");
    CU(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(TSB_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return TSB_OK;
}
TSB_API int tsb_free(tsb_ctx *c, void *dptr) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaFree(dptr));
    return TSB_OK;
}
TSB_API int tsb_memset(tsb_ctx *c, void *dptr, int value, size_t bytes) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaMemsetAsync(dptr, value, bytes, c->stream));
    return TSB_OK;
}
TSB_API int tsb_upload(tsb_ctx *c, void *dptr, const void *host, size_t bytes) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return TSB_OK;
}
TSB_API int tsb_download(tsb_ctx *c, void *host, const void *dptr, size_t bytes) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    return TSB_OK;
}

// Measured non-tensor FMA rate of the device (TFLOP/s, an FMA counted as two): fp64 if is_f64, else fp32.
TSB_API int tsb_probe_fma_peak(tsb_ctx *c, int32_t is_f64, double *tflops) {
    if (!c || !tflops) return fail(TSB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    void *sink = nullptr;
    CU(cudaMalloc(&sink, 64));
    cudaError_t e = op_fma_peak(is_f64, sink, sms, c->stream, tflops);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(TSB_ECUDA, "fma peak probe: %s", cudaGetErrorString(e));
    return TSB_OK;
}

// ---- host logic ----
TSB_API int tsb_hop_fixed(double hop, uint64_t *hop_samples, uint64_t *hop_fract) {
    if (!hop_samples || !hop_fract) return fail(TSB_EINVAL, "tsb_hop_fixed: null output");
    if (!hop_fixed(hop, hop_samples, hop_fract)) return fail(TSB_EINVAL, "ignoring invalid hop: %g", hop);
    return TSB_OK;
}
TSB_API int tsb_schedule_const(int64_t in_len, uint32_t n, double hop, double rate, int64_t max_frames, int64_t cap,
                                  int64_t *n_frames, int64_t *in_off, int64_t *out_off, int32_t *hop_tab, double *t) {
    if (!n_frames) return fail(TSB_EINVAL, "tsb_schedule_const: null n_frames");
    if (hop < 0) return fail(TSB_EINVAL, "ignoring invalid hop: %g", hop);
    if (n == 0) return fail(TSB_EINVAL, "frame size 0");
    SchedState st;
    st.rate = rate;
    std::vector<FrameRec> fr = schedule_const(st, in_len, n, hop, max_frames);
    *n_frames = (int64_t)fr.size();
    int64_t lim = (int64_t)fr.size() < cap ? (int64_t)fr.size() : cap;
    for (int64_t f = 0; f < lim; f++) {
        if (in_off) in_off[f] = fr[f].in_off;
        if (out_off) out_off[f] = fr[f].out_off;
        if (hop_tab) hop_tab[f] = fr[f].hop;
        if (t) t[f] = fr[f].t;
    }
    return TSB_OK;
}
TSB_API int tsb_shift_map(uint32_t n, double rate, const double *fn_bins, int32_t *dest) {
    if (!fn_bins || !dest || n < 2) return fail(TSB_EINVAL, "tsb_shift_map: bad argument");
    std::vector<uint32_t> map;
    std::vector<uint16_t> src;
    if (n > 65536) return fail(TSB_EUNSUPPORTED, "tsb_shift_map: n > 65536");
    build_shift_map((int)n, rate, fn_bins, map, src);
    for (uint32_t j = 0; j < n / 2; j++) dest[j] = -1;
    for (uint32_t k = 0; k < n / 2; k++) {
        uint32_t first = map[k] >> 16, cnt = map[k] & 0xffff;
        for (uint32_t s = 0; s < cnt; s++) dest[src[first + s]] = (int32_t)k;
    }
    return TSB_OK;
}

// ---- plans ----
template <typename T>
static int upload_vec(tsb_plan *p, const std::vector<T> &v, const T **dst) {
    void *d = nullptr;
    size_t bytes = v.size() * sizeof(T);
    cudaError_t e = cudaMalloc(&d, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(TSB_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    p->dev.push_back(d);
    CU(cudaMemcpy(d, v.data(), bytes, cudaMemcpyHostToDevice));
    *dst = reinterpret_cast<const T *>(d);
    return TSB_OK;
}
template <typename real>
static int upload_tables(tsb_plan *p, const PlanDesc &d, ChainTables<real> &t) {
    HostTables<real> h = build_tables<real>(d);
    t = h.view();
    int r;
    if ((r = upload_vec(p, h.tw, &t.tw))) return r;
    if ((r = upload_vec(p, h.twn, &t.twn))) return r;
    if ((r = upload_vec(p, h.win_a, &t.win_a))) return r;
    if (h.win_s.empty()) t.win_s = t.win_a;  // same window on both sides: one table
    else if ((r = upload_vec(p, h.win_s, &t.win_s))) return r;
    if ((r = upload_vec(p, h.map, &t.map))) return r;
    if ((r = upload_vec(p, h.map_src, &t.map_src))) return r;
    if (!h.map1.empty() && (r = upload_vec(p, h.map1, &t.map1))) return r;
    if ((r = upload_vec(p, h.src_used, &t.src_used))) return r;
    if (!h.dst_order.empty() && (r = upload_vec(p, h.dst_order, &t.dst_order))) return r;
    if ((r = upload_vec(p, h.gain, &t.gain))) return r;
    if (!h.win64.empty() && (r = upload_vec(p, h.win64, &t.win64))) return r;
    return TSB_OK;
}

TSB_API int tsb_plan_destroy(tsb_plan *p) {
    if (!p) return TSB_OK;
    cudaSetDevice(p->ctx->device);
    for (void *d : p->dev) cudaFree(d);
    delete p;
    return TSB_OK;
}
TSB_API int tsb_plan_create(tsb_ctx *c, const tsb_plan_desc *desc, tsb_plan **out) {
    if (!c || !desc || !out) return fail(TSB_EINVAL, "tsb_plan_create: null argument");
    uint32_t n = desc->n;
    if (n < 64 || n > 8192 || (n & (n - 1)))
        return fail(TSB_EUNSUPPORTED, "frame size %u: the fused chain supports powers of two in 64..8192", n);
    if (desc->precision != TSB_F32 && desc->precision != TSB_F64 && desc->precision != TSB_F64_OUT32) return fail(TSB_EINVAL, "bad precision %d", desc->precision);
    if (desc->program != TSB_PROG_PV_SHIFT && desc->program != TSB_PROG_GAIN) return fail(TSB_EINVAL, "bad program %d", desc->program);
    if (desc->program == TSB_PROG_PV_SHIFT && !(desc->hop > 0)) return fail(TSB_EINVAL, "phase vocoder needs hop > 0 (got %g)", desc->hop);
    if (!(desc->rate > 0)) return fail(TSB_EINVAL, "bad rate %g", desc->rate);
    CU(cudaSetDevice(c->device));
    tsb_plan *p = new tsb_plan();
    p->ctx = c;
    p->desc = *desc;
    p->log2n = ilog2(n);
    PlanDesc d;
    d.n = (int)n;
    d.program = desc->program;
    bool pv = desc->program == TSB_PROG_PV_SHIFT;
    d.win_in = pv ? (int)WIN_SQRT_HANN : desc->win_in;
    d.win_out = pv ? (int)WIN_SQRT_HANN : desc->win_out;
    d.hop = desc->hop > 0 ? desc->hop : 1.0;
    d.rate = desc->rate;
    d.ratio = desc->ratio;
    d.fn_bins = desc->fn_bins;
    d.gain = desc->gain;
    int r = desc->precision == TSB_F32 ? upload_tables<float>(p, d, p->t32) : upload_tables<double>(p, d, p->t64);
    if (r) {
        tsb_plan_destroy(p);
        return r;
    }
    p->desc.fn_bins = nullptr;
    p->desc.gain = nullptr;
    *out = p;
    return TSB_OK;
}

// ---- the fused chain ----
static chain_launch_fn launcher(int log2n) {
    switch (log2n) {
        case 6: return launch_chain_6;
        case 7: return launch_chain_7;
        case 8: return launch_chain_8;
        case 9: return launch_chain_9;
        case 10: return launch_chain_10;
        case 11: return launch_chain_11;
        case 12: return launch_chain_12;
        case 13: return launch_chain_13;
    }
    return nullptr;
}
static int check_args(const tsb_chain_args *a) {
    if (!a) return fail(TSB_EINVAL, "null args");
    if (a->lanes < 0 || a->n_frames < 0) return fail(TSB_EINVAL, "negative lanes / n_frames");
    if (a->lanes && a->n_frames && (!a->in || !a->out)) return fail(TSB_EINVAL, "null in/out");
    if (!a->hop_tab && a->hop < 0) return fail(TSB_EINVAL, "negative hop");
    if (a->in_off0 < 0 || a->out_off0 < 0) return fail(TSB_EINVAL, "negative offset");
    if (a->in_len < 0 || a->out_len < 0) return fail(TSB_EINVAL, "negative row length");
    if (a->lanes > 1 && (a->in_stride < a->in_len || a->out_stride < a->out_len))
        return fail(TSB_EINVAL, "row stride shorter than the row (in %lld < %lld or out %lld < %lld)", (long long)a->in_stride,
                    (long long)a->in_len, (long long)a->out_stride, (long long)a->out_len);
    return TSB_OK;
}
static ChainArgs to_args(const tsb_chain_args *a) {
    ChainArgs k;
    k.in = a->in; k.in_stride = a->in_stride; k.in_len = a->in_len;
    k.out = a->out; k.out_stride = a->out_stride; k.out_len = a->out_len;
    k.n_frames = a->n_frames; k.in_off0 = a->in_off0; k.out_off0 = a->out_off0;
    k.hop = a->hop; k.hop_tab = a->hop_tab;
    k.st_last_phase = a->state_last_phase; k.st_phase_sum = a->state_phase_sum; k.st_ola = a->state_ola;
    k.load_state = a->load_state; k.lanes = a->lanes;
    k.n_seg = 0; k.seg_tab = nullptr; k.mode = 0; k.st_warm_phase = nullptr;
    k.st_inc = nullptr; k.inc_frames = 0;
    k.prev_hop = a->prev_hop;
    return k;
}
TSB_API int tsb_chain_render(tsb_ctx *c, tsb_plan *p, const tsb_chain_args *a) {
    if (!c || !p) return fail(TSB_EINVAL, "null context / plan");
    int r = check_args(a);
    if (r) return r;
    if (a->lanes == 0 || a->n_frames == 0) return TSB_OK;
    CU(cudaSetDevice(c->device));
    ChainArgs k = to_args(a);
    chain_launch_fn fn = launcher(p->log2n);
    if (!fn) return fail(TSB_EUNSUPPORTED, "no kernel for frame size 2^%d", p->log2n);
    const void *tb = p->desc.precision == TSB_F32 ? (const void *)&p->t32 : (const void *)&p->t64;
    cudaError_t e = fn(p->desc.precision, p->desc.program, tb, k, c->stream);
    if (e != cudaSuccess) return fail(TSB_ECUDA, "chain kernel launch: %s", cudaGetErrorString(e));
    c->launches++;
    return TSB_OK;
}

// Host-buffer render.  Lanes are independent and a lane's rows are contiguous in the caller's arrays, so the
// batch is cut into lane chunks and pipelined: chunk k's input upload (one contiguous copy), its chain kernel on
// its own stream, and its output download overlap the other chunks' (uploads and downloads use separate copy
// engines; the chunk kernels share the SMs).  Row-strided 2-D copies of time blocks were measured 10-20x slower
// than contiguous ones on this platform (tools/pcie_probe.py), hence lanes, not time, are the pipeline axis.
// Per-lane results are bit-identical to tsb_chain_render.
TSB_API int tsb_chain_render_host(tsb_ctx *c, tsb_plan *p, const tsb_chain_args *a_in) {
    if (!c || !p) return fail(TSB_EINVAL, "null context / plan");
    int r = check_args(a_in);
    if (r) return r;
    if (a_in->lanes == 0 || a_in->n_frames == 0) return TSB_OK;
    tsb_chain_args a_norm = *a_in;
    if (a_norm.lanes == 1) {  // a single row: its stride means nothing (numpy reports 1 for it), the row is what is copied
        a_norm.in_stride = a_norm.in_len;
        a_norm.out_stride = a_norm.out_len;
    }
    const tsb_chain_args *a = &a_norm;
    CU(cudaSetDevice(c->device));
    const int64_t n = p->desc.n, m = n / 2, lanes = a->lanes, F = a->n_frames;
    const size_t osz = p->desc.precision == TSB_F64 ? 8 : 4;
    const size_t in_bytes = (size_t)lanes * a->in_stride * 4, out_bytes = (size_t)lanes * a->out_stride * osz;
    // chunks: enough lanes each to keep copies large (>= 256 MB) and at most 32 chunks
    int64_t nc = (int64_t)((in_bytes + out_bytes) / (512ull << 20));
    nc = nc < 1 ? 1 : (nc > 32 ? 32 : nc);
    if (const char *ev = getenv("TSB_HOST_CHUNKS")) nc = atoi(ev) > 0 ? atoi(ev) : nc;  // tuning
    if (nc > lanes) nc = lanes;
    const bool trace = getenv("TSB_HOST_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_alloc = 0, t_enq = 0, t_sync = 0;
    struct Res {
        std::vector<void *> ptrs;
        std::vector<cudaEvent_t> evs;
        ~Res() {
            for (cudaEvent_t e : evs) cudaEventDestroy(e);
            for (void *q : ptrs) cudaFree(q);
        }
        void *get(size_t bytes) {
            void *q = nullptr;
            if (cudaMalloc(&q, bytes ? bytes : 16) != cudaSuccess) return nullptr;
            ptrs.push_back(q);
            return q;
        }
        cudaEvent_t ev() {
            cudaEvent_t e = nullptr;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            evs.push_back(e);
            return e;
        }
    } res;
    // staging rows and streams live in the context (grow-only; tsb_ctx_trim frees them)
    const size_t want[2] = {in_bytes, out_bytes};
    for (int i = 0; i < 2; i++) {
        if (c->host_stage_bytes[i] >= want[i]) continue;
        if (c->host_stage[i]) cudaFree(c->host_stage[i]);
        c->host_stage[i] = nullptr;
        c->host_stage_bytes[i] = 0;
        if (cudaMalloc(&c->host_stage[i], want[i] ? want[i] : 16) != cudaSuccess) {
            c->host_stage[i] = nullptr;
            (void)cudaGetLastError();
            return fail(TSB_ENOMEM, "device allocation of %zu bytes for the host-buffer render failed", want[i]);
        }
        c->host_stage_bytes[i] = want[i];
    }
    while ((int64_t)c->host_streams.size() < nc + 2) {
        cudaStream_t st = nullptr;
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        c->host_streams.push_back(st);
    }
    float *din = (float *)c->host_stage[0];
    char *dout = (char *)c->host_stage[1];
    int32_t *dhops = nullptr;
    if (a->hop_tab) {
        dhops = (int32_t *)res.get((size_t)F * 4);
        if (!dhops) return fail(TSB_ENOMEM, "device allocation failed");
        CU(cudaMemcpyAsync(dhops, a->hop_tab, (size_t)F * 4, cudaMemcpyHostToDevice, c->stream));
    }
    double *host_state[3] = {a->state_last_phase, a->state_phase_sum, a->state_ola};
    double *dev_state[3] = {nullptr, nullptr, nullptr};
    size_t st_w[3] = {(size_t)m, (size_t)m, (size_t)n};
    for (int i = 0; i < 3; i++) {
        if (!host_state[i]) continue;
        dev_state[i] = (double *)res.get((size_t)lanes * st_w[i] * 8);
        if (!dev_state[i]) return fail(TSB_ENOMEM, "device allocation failed");
        if (a->load_state) CU(cudaMemcpyAsync(dev_state[i], host_state[i], (size_t)lanes * st_w[i] * 8, cudaMemcpyHostToDevice, c->stream));
    }
    cudaEvent_t setup = res.ev();
    CU(cudaEventRecord(setup, c->stream));
    t_alloc = now();
    cudaStream_t s_in = c->host_streams[0], s_out = c->host_streams[1];
    CU(cudaStreamWaitEvent(s_in, setup, 0));
    cudaStream_t saved = c->stream;
    int rc = TSB_OK;
    // Chunk sizes ramp up at the start and down at the end (weights 1,2,3,4,...,4,3,2,1): the download cannot start
    // before the first chunk is uploaded and rendered, and nothing overlaps the last chunk's render and download,
    // so those two chunks are kept small; a lane takes the same time to render however few lanes share the launch.
    std::vector<int64_t> chunk(nc, 0);
    {
        std::vector<int64_t> wgt(nc);
        int64_t wsum = 0;
        for (int64_t k = 0; k < nc; k++) {
            int64_t edge = k < nc - 1 - k ? k : nc - 1 - k;
            wgt[k] = edge < 3 ? edge + 1 : 4;
            wsum += wgt[k];
        }
        int64_t given = 0;
        for (int64_t k = 0; k < nc; k++) {
            chunk[k] = lanes * wgt[k] / wsum;
            given += chunk[k];
        }
        for (int64_t k = nc / 2; given < lanes; k = (k + 1) % nc) {  // remainder to the middle chunks
            chunk[k]++;
            given++;
        }
    }
    int64_t l0 = 0;
    for (int64_t k = 0; k < nc && rc == TSB_OK; k++) {
        int64_t cnt = chunk[k];
        if (cnt == 0) continue;
        cudaEvent_t in_done = res.ev(), k_done = res.ev();
        cudaStream_t s_k = c->host_streams[2 + k];
        cudaError_t e = cudaMemcpyAsync(din + (size_t)l0 * a->in_stride, a->in + (size_t)l0 * a->in_stride,
                                        (size_t)cnt * a->in_stride * 4, cudaMemcpyHostToDevice, s_in);
        if (e == cudaSuccess) e = cudaEventRecord(in_done, s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, in_done, 0);
        if (e != cudaSuccess) { rc = fail(TSB_ECUDA, "upload: %s", cudaGetErrorString(e)); break; }
        tsb_chain_args d = *a;
        d.in = din + (size_t)l0 * a->in_stride;
        d.out = dout + (size_t)l0 * a->out_stride * osz;
        d.hop_tab = dhops;
        d.state_last_phase = dev_state[0] ? dev_state[0] + (size_t)l0 * m : nullptr;
        d.state_phase_sum = dev_state[1] ? dev_state[1] + (size_t)l0 * m : nullptr;
        d.state_ola = dev_state[2] ? dev_state[2] + (size_t)l0 * n : nullptr;
        d.lanes = (int32_t)cnt;
        e = cudaMemsetAsync((void *)d.out, 0, (size_t)cnt * a->out_stride * osz, s_k);
        if (e != cudaSuccess) { rc = fail(TSB_ECUDA, "memset: %s", cudaGetErrorString(e)); break; }
        c->stream = s_k;  // launch this chunk on its own stream
        rc = tsb_chain_render(c, p, &d);
        c->stream = saved;
        if (rc) break;
        e = cudaEventRecord(k_done, s_k);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_out, k_done, 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char *)a->out + (size_t)l0 * a->out_stride * osz, dout + (size_t)l0 * a->out_stride * osz,
                                (size_t)cnt * a->out_stride * osz, cudaMemcpyDeviceToHost, s_out);
        for (int i = 0; i < 3 && e == cudaSuccess; i++)
            if (host_state[i])
                e = cudaMemcpyAsync(host_state[i] + (size_t)l0 * st_w[i], dev_state[i] + (size_t)l0 * st_w[i],
                                    (size_t)cnt * st_w[i] * 8, cudaMemcpyDeviceToHost, s_out);
        if (e != cudaSuccess) { rc = fail(TSB_ECUDA, "download: %s", cudaGetErrorString(e)); break; }
        l0 += cnt;
    }
    t_enq = now();
    for (int64_t k = 0; k < nc + 2; k++) cudaStreamSynchronize(c->host_streams[k]);
    CU(cudaStreamSynchronize(c->stream));
    t_sync = now();
    if (trace)
        fprintf(stderr, "[tsb host render] chunks %lld  alloc %.1f ms  enqueue %.1f ms  wait %.1f ms  (in %.2f GB, out %.2f GB)\n",
                (long long)nc, t_alloc - t_begin, t_enq - t_alloc, t_sync - t_enq, in_bytes / 1e9, out_bytes / 1e9);
    return rc;
}

// ---- the helpers one by one ----
static int ctx_scratch(tsb_ctx *c, size_t bytes, double **out) {
    if (c->scratch_bytes < bytes) {
        // the stream may still be using the old block
        CU(cudaStreamSynchronize(c->stream));
        if (c->scratch) cudaFree(c->scratch);
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        cudaError_t e = cudaMalloc(&c->scratch, bytes);
        if (e != cudaSuccess) return fail(TSB_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        c->scratch_bytes = bytes;
    }
    *out = c->scratch;
    return TSB_OK;
}
TSB_API int tsb_window(tsb_ctx *c, double *data, int64_t nch, int64_t size, int32_t kind) {
    if (!c) return fail(TSB_EINVAL, "null context");
    if (nch < 0 || size < 0) return fail(TSB_EINVAL, "negative shape");
    if (kind != TSB_WIN_HANN && kind != TSB_WIN_SQRT_HANN && kind != TSB_WIN_NONE) return fail(TSB_EINVAL, "bad window kind %d", kind);
    if (nch * size == 0 || kind == TSB_WIN_NONE) return TSB_OK;
    CU(cudaSetDevice(c->device));
    auto key = std::make_pair(size, (int)kind);
    auto it = c->windows.find(key);
    if (it == c->windows.end()) {
        std::vector<double> w = window_table(kind, (int)size);
        double *d = nullptr;
        cudaError_t e = cudaMalloc(&d, w.size() * sizeof(double));
        if (e != cudaSuccess) return fail(TSB_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
        CU(cudaMemcpyAsync(d, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        it = c->windows.emplace(key, d).first;
    }
    CU(op_window(data, it->second, nch, size, c->stream));
    c->launches++;
    return TSB_OK;
}
TSB_API int tsb_fft_forward(tsb_ctx *c, const double *in, double *out, int64_t rows, int64_t n) {
    if (!c) return fail(TSB_EINVAL, "null context");
    if (rows < 0 || n < 0) return fail(TSB_EINVAL, "negative shape");
    if (n > (1LL << 24)) return fail(TSB_EUNSUPPORTED, "FFT length %lld > 2^24", (long long)n);
    CU(cudaSetDevice(c->device));
    long long l = 0;
    CU(op_fft_forward(c->device, in, out, rows, n, c->stream, &l));
    c->launches += l;
    return TSB_OK;
}
TSB_API int tsb_frft(tsb_ctx *c, double *in, double *out, int64_t nch, int64_t n, double exponent) {
    if (!c) return fail(TSB_EINVAL, "null context");
    if (nch < 0 || n < 0) return fail(TSB_EINVAL, "negative shape");
    if (nch * n == 0) return TSB_OK;  // rtlib/fft.hpp:8-9
    if (n > (1LL << 23)) return fail(TSB_EUNSUPPORTED, "frft length %lld > 2^23", (long long)n);
    CU(cudaSetDevice(c->device));
    double *g;
    int r = ctx_scratch(c, (size_t)(nch * n) * 16, &g);
    if (r) return r;
    long long l = 0;
    CU(op_frft(c->device, in, out, g, nch, n, exponent, c->stream, &l));
    c->launches += l;
    return TSB_OK;
}
TSB_API int tsb_pv_analyze(tsb_ctx *c, const double *spec, double *last_phase, double *pvdata, int64_t nch, int64_t size,
                           double hop, double rate) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(op_pv_analyze(spec, last_phase, pvdata, nch, size, hop, rate, c->stream));
    c->launches++;
    return TSB_OK;
}
TSB_API int tsb_pv_synthesize(tsb_ctx *c, double *spec, double *phase_sum, const double *pvdata, int64_t nch, int64_t size,
                              double hop, double rate) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(op_pv_synthesize(spec, phase_sum, pvdata, nch, size, hop, rate, c->stream));
    c->launches++;
    return TSB_OK;
}
TSB_API int tsb_pv_freqs(tsb_ctx *c, const double *pvdata, int64_t nch, int64_t size, double *host_freqs) {
    if (!c || !host_freqs) return fail(TSB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    int64_t half = size / 2;
    for (int64_t ch = 0; ch < nch; ch++)  // freq is the second double of each PVBin
        CU(cudaMemcpy2DAsync(host_freqs + ch * half, 8, pvdata + 2 * ch * size + 1, 16, 8, (size_t)half,
                             cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TSB_OK;
}
TSB_API int tsb_pv_shift(tsb_ctx *c, double *pvdata, int64_t nch, int64_t size, double rate, const double *host_fn_bins,
                         const double *host_fn_freqs) {
    if (!c || !host_fn_bins || !host_fn_freqs) return fail(TSB_EINVAL, "null argument");
    if (nch * size == 0) return TSB_OK;
    CU(cudaSetDevice(c->device));
    const int64_t half = size / 2;
    std::vector<unsigned> off(nch * half + 1, 0), src;
    for (int64_t ch = 0; ch < nch; ch++) {
        std::vector<std::vector<unsigned>> per(half);
        for (int64_t j = 0; j < half; j++) {
            double q = host_fn_bins[ch * half + j] / rate * (size_t)size;  // rtlib/pv.hpp:72
            if (!(q > -2147483648.0 && q < 2147483648.0)) continue;
            int k = (int)q;
            if (k >= 0 && k < (int)half) per[k].push_back((unsigned)j);
        }
        for (int64_t k = 0; k < half; k++) {
            off[ch * half + k] = (unsigned)src.size();
            for (unsigned j : per[k]) src.push_back(j);
        }
    }
    off[nch * half] = (unsigned)src.size();
    if (src.empty()) src.push_back(0);
    // scratch: [copy of pvdata | fn_freqs | off | src]
    size_t b_tmp = (size_t)(nch * size) * 16, b_fr = (size_t)(nch * half) * 8, b_off = off.size() * 4, b_src = src.size() * 4;
    double *scr;
    int r = ctx_scratch(c, b_tmp + b_fr + b_off + b_src + 64, &scr);
    if (r) return r;
    char *p = reinterpret_cast<char *>(scr);
    double *d_tmp = reinterpret_cast<double *>(p);
    double *d_fr = reinterpret_cast<double *>(p + b_tmp);
    unsigned *d_off = reinterpret_cast<unsigned *>(p + b_tmp + b_fr);
    unsigned *d_src = reinterpret_cast<unsigned *>(p + b_tmp + b_fr + b_off);
    CU(cudaMemcpyAsync(d_tmp, pvdata, b_tmp, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(d_fr, host_fn_freqs, b_fr, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_off, off.data(), b_off, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_src, src.data(), b_src, cudaMemcpyHostToDevice, c->stream));
    CU(op_pv_shift(d_tmp, pvdata, d_fr, d_off, d_src, nch, size, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // off/src/host arrays go out of scope
    c->launches++;
    return TSB_OK;
}

TSB_API int tsb_gather_frame(tsb_ctx *c, const float *in, int64_t in_stride, int64_t in_len, int64_t in_off, double *frame,
                             int64_t lanes, int64_t w) {
    if (!c) return fail(TSB_EINVAL, "null context");
    if (lanes < 0 || w < 0) return fail(TSB_EINVAL, "negative shape");
    CU(cudaSetDevice(c->device));
    CU(op_gather(in, in_stride, in_len, in_off, frame, lanes, w, c->stream));
    c->launches++;
    return TSB_OK;
}
TSB_API int tsb_overlap_add(tsb_ctx *c, const double *frame, int64_t lanes, int64_t osize, int64_t hop, float *out,
                            int64_t out_stride, int64_t out_len, int64_t out_off, const float *tail_in, float *tail_out,
                            int64_t tail_cap, int64_t tail_len, int64_t *new_tail_len) {
    if (!c || !new_tail_len) return fail(TSB_EINVAL, "null argument");
    if (lanes < 0 || osize < 0 || hop < 0 || tail_len < 0) return fail(TSB_EINVAL, "negative shape");
    int64_t keep = (tail_len > osize ? tail_len : osize) - hop;
    if (keep < 0) keep = 0;
    if (keep > tail_cap || tail_len > tail_cap) return fail(TSB_EINVAL, "overlap of %lld samples exceeds tail_cap %lld", (long long)keep, (long long)tail_cap);
    if (tail_in == tail_out && keep > 0) return fail(TSB_EINVAL, "tail_in and tail_out must be distinct buffers");
    CU(cudaSetDevice(c->device));
    CU(op_overlap_add(frame, osize, hop, out, out_stride, out_len, out_off, tail_in, tail_out, tail_cap, tail_len, keep, lanes,
                      c->stream));
    *new_tail_len = keep;
    c->launches++;
    return TSB_OK;
}

// ---- fused spectral synthesis (cfg5) ----
TSB_API int tsb_spectral_render(tsb_ctx *c, const tsb_spectral_args *a) {
    if (!c || !a) return fail(TSB_EINVAL, "null argument");
    const uint32_t n = a->frame;
    if (n < 256 || n > (1u << 18) || (n & (n - 1))) return fail(TSB_EUNSUPPORTED, "frame size %u: fused spectral synthesis supports powers of two in 256..262144", n);
    if (a->channels < 0 || a->n_frames < 0 || a->hop < 0) return fail(TSB_EINVAL, "negative channels / n_frames / hop");
    if (a->generator != TSB_GEN_HASH && a->generator != TSB_GEN_BUFFER && a->generator != TSB_GEN_EXPR) return fail(TSB_EINVAL, "bad generator %d", a->generator);
    if (a->generator == TSB_GEN_EXPR) {
        if (const char *why = expr_check(a->expr)) return fail(TSB_EINVAL, "%s", why);
    }
    if (a->generator == TSB_GEN_BUFFER && !a->spectrum && a->n_frames * a->channels) return fail(TSB_EINVAL, "TSB_GEN_BUFFER needs a spectrum");
    if (a->channels == 0 || a->n_frames == 0) return TSB_OK;
    if (!a->out || a->out_off0 < 0 || a->out_len < 0 || (a->channels > 1 && a->out_stride < a->out_len)) return fail(TSB_EINVAL, "bad output rows");
    CU(cudaSetDevice(c->device));
    SpectralJob jb;
    jb.log2n = ilog2(n);
    jb.channels = a->channels;
    jb.n_frames = a->n_frames;
    jb.first_frame = a->first_frame;
    jb.hop = a->hop;
    jb.generator = a->generator == TSB_GEN_HASH ? (int)SPEC_GEN_HASH : a->generator == TSB_GEN_EXPR ? (int)SPEC_GEN_EXPR : (int)SPEC_GEN_BUFFER;
    jb.expr = a->generator == TSB_GEN_EXPR ? a->expr : nullptr;
    jb.spectrum = a->spectrum;
    jb.out = a->out;
    jb.out_stride = a->out_stride;
    jb.out_len = a->out_len;
    jb.out_off0 = a->out_off0;
    jb.state_edge = a->state_edge;
    jb.load_state = a->load_state;
    long long l = 0;
    cudaError_t e = op_spectral_render(c->device, jb, c->stream, &l);
    c->launches += l;
    if (e != cudaSuccess) return fail(TSB_ECUDA, "spectral synthesis: %s", cudaGetErrorString(e));
    return TSB_OK;
}

TSB_API int tsb_expr_check(const tsb_expr *x) {
    if (const char *why = expr_check(x)) return fail(TSB_EINVAL, "%s", why);
    return TSB_OK;
}

// ---- batched granular render (cfg4) ----
static int window_table_dev(tsb_ctx *c, int64_t size, int kind, double **out) {
    auto key = std::make_pair(size, kind);
    auto it = c->windows.find(key);
    if (it == c->windows.end()) {
        std::vector<double> w = window_table(kind, (int)size);
        double *d = nullptr;
        cudaError_t e = cudaMalloc(&d, (w.size() ? w.size() : 1) * sizeof(double));
        if (e != cudaSuccess) return fail(TSB_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
        CU(cudaMemcpy(d, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
        it = c->windows.emplace(key, d).first;
    }
    *out = it->second;
    return TSB_OK;
}
TSB_API int tsb_granular_render(tsb_ctx *c, const tsb_granular_args *a) {
    if (!c || !a) return fail(TSB_EINVAL, "null argument");
    if (a->streams < 0 || a->channels <= 0 || a->frames_max < 0) return fail(TSB_EINVAL, "bad streams / channels / frames_max");
    if (a->streams == 0 || a->frames_max == 0) return TSB_OK;
    if (!a->in || !a->out || !a->n_frames || !a->win || !a->hop || !a->osize || !a->in_off || !a->out_off || !a->shift)
        return fail(TSB_EINVAL, "null table or buffer");
    CU(cudaSetDevice(c->device));
    const size_t S = (size_t)a->streams, FM = (size_t)a->frames_max, cells = S * FM;
    // validate, and number the distinct frame sizes (their Hann tables are cached per context)
    std::map<int32_t, int32_t> ids;
    std::vector<int32_t> win_id(cells, 0);
    std::vector<const double *> wt;
    for (size_t s = 0; s < S; s++) {
        if (a->n_frames[s] < 0 || a->n_frames[s] > a->frames_max) return fail(TSB_EINVAL, "n_frames[%zu] out of range", s);
        for (int64_t f = 0; f < a->n_frames[s]; f++) {
            const size_t q = s * FM + (size_t)f;
            const int32_t W = a->win[q];
            if (W < 1 || W > 8192) return fail(TSB_EUNSUPPORTED, "frame size %d (stream %zu, frame %lld): 1..8192 supported", W, s, (long long)f);
            if (a->osize[q] < 0 || a->osize[q] > 16384) return fail(TSB_EUNSUPPORTED, "output frame of %d samples: 0..16384 supported", a->osize[q]);
            if (a->hop[q] < 0 || a->hop[q] > 16384) return fail(TSB_EUNSUPPORTED, "hop %d: 0..16384 supported", a->hop[q]);
            if (!(a->shift[q] > 0)) return fail(TSB_EINVAL, "shift must be positive");
            auto it = ids.find(W);
            if (it == ids.end()) {
                double *d = nullptr;
                int r = window_table_dev(c, W, TSB_WIN_HANN, &d);
                if (r) return r;
                it = ids.emplace(W, (int32_t)wt.size()).first;
                wt.push_back(d);
            }
            win_id[q] = it->second;
        }
    }
    // one device block for the tables
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_nf = al(S * 8), b_i32 = al(cells * 4), b_i64 = al(cells * 8), b_wt = al(wt.size() * 8);
    const size_t total = b_nf + 4 * b_i32 + 2 * b_i64 + b_i64 + b_wt;
    char *blk = nullptr;
    cudaError_t e = cudaMalloc(&blk, total);
    if (e != cudaSuccess) return fail(TSB_ENOMEM, "granular tables (%zu bytes): %s", total, cudaGetErrorString(e));
    char *q = blk;
    auto put = [&](const void *h, size_t bytes, size_t slot) {
        char *d = q;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream);
        q += slot;
        return d;
    };
    GranularJob jb;
    jb.n_frames = (const long long *)put(a->n_frames, S * 8, b_nf);
    jb.win = (const int *)put(a->win, cells * 4, b_i32);
    jb.hop = (const int *)put(a->hop, cells * 4, b_i32);
    jb.osize = (const int *)put(a->osize, cells * 4, b_i32);
    jb.win_id = (const int *)put(win_id.data(), cells * 4, b_i32);
    jb.in_off = (const long long *)put(a->in_off, cells * 8, b_i64);
    jb.out_off = (const long long *)put(a->out_off, cells * 8, b_i64);
    jb.shift = (const double *)put(a->shift, cells * 8, b_i64);
    jb.wtabs = (const double *const *)put(wt.data(), wt.size() * 8, b_wt);
    jb.in = a->in; jb.in_stride = a->in_stride; jb.in_len = a->in_len;
    jb.out = a->out; jb.out_stride = a->out_stride; jb.out_len = a->out_len;
    jb.lanes = a->streams * a->channels;
    jb.channels = a->channels;
    jb.frames_max = a->frames_max;
    if (e == cudaSuccess) e = op_granular_render(jb, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // the host tables / win_id / wt are about to go away
    cudaFree(blk);
    if (e != cudaSuccess) return fail(TSB_ECUDA, "granular render: %s", cudaGetErrorString(e));
    c->launches++;
    return TSB_OK;
}

// ---- time-parallel rendering ----
struct tsb_timepar {
    tsb_ctx *ctx;
    tsb_plan *plan;
    tsb_chain_args args;
    int n_seg, warmup, warm_hop;
    size_t lanes, n, m;
    std::vector<long long> seg;  // [n_seg][SEG_COLS]
    long long *d_seg = nullptr, *d_pos = nullptr;
    double *d_last = nullptr, *d_sum = nullptr, *d_warm = nullptr, *d_ola = nullptr;
    double *d_inc = nullptr;  // [lanes][n_frames][m] per-frame increments of every destination bin (pre-pass)
    char *block = nullptr;   // one allocation holding all of the above
    bool from_pool = false;
    bool prepassed = false, scanned = false;
};
TSB_API int tsb_timepar_destroy(tsb_timepar *tp) {
    if (!tp) return TSB_OK;
    cudaSetDevice(tp->ctx->device);
    if (tp->from_pool) {
        tp->ctx->tp_pool_busy = false;  // work queued on the stream keeps using it until it completes: stream order
    } else if (tp->block) {
        cudaStreamSynchronize(tp->ctx->stream);
        cudaFree(tp->block);
    }
    delete tp;
    return TSB_OK;
}
TSB_API int tsb_timepar_segments(tsb_timepar *tp) { return tp ? tp->n_seg : 0; }
TSB_API int tsb_timepar_create(tsb_ctx *c, tsb_plan *p, const tsb_chain_args *a, const int32_t *host_hop_tab,
                               int32_t n_segments, int32_t warmup, int32_t warm_hop, int32_t block_lead, tsb_timepar **out) {
    if (!c || !p || !out) return fail(TSB_EINVAL, "null argument");
    int r = check_args(a);
    if (r) return r;
    if ((a->hop_tab == nullptr) != (host_hop_tab == nullptr)) return fail(TSB_EINVAL, "host_hop_tab must mirror args->hop_tab");
    if (p->desc.program != TSB_PROG_PV_SHIFT && warmup) return fail(TSB_EINVAL, "warm-up frames only exist for the phase vocoder");
    if (warmup && (warm_hop < 0 || a->in_off0 - warm_hop < 0)) return fail(TSB_EINVAL, "warm-up frame would start before the input");
    if (block_lead < 0 || (block_lead > 0 && block_lead >= a->n_frames)) return fail(TSB_EINVAL, "lead-in frames (%d) must leave the block a frame of its own", block_lead);
    if (block_lead > 0 && !warmup) return fail(TSB_EINVAL, "a block with lead-in frames needs the warm-up frame before them");
    CU(cudaSetDevice(c->device));
    tsb_timepar *tp = new tsb_timepar();
    tp->ctx = c;
    tp->plan = p;
    tp->args = *a;
    tp->warmup = warmup;
    tp->warm_hop = warm_hop;
    tp->lanes = (size_t)a->lanes;
    tp->n = p->desc.n;
    tp->m = tp->n / 2;
    const int64_t F = a->n_frames;
    // every segment must emit at least n samples, so that a tail never reaches past the next segment
    int64_t hop_min = a->hop;
    if (host_hop_tab) {
        hop_min = host_hop_tab[0];
        for (int64_t f = 0; f < F; f++) hop_min = host_hop_tab[f] < hop_min ? host_hop_tab[f] : hop_min;
    }
    int64_t min_frames = hop_min > 0 ? ((int64_t)tp->n + hop_min - 1) / hop_min : F;
    if (min_frames < 1) min_frames = 1;
    int64_t nseg = n_segments < 1 ? 1 : n_segments;
    const int64_t F_own = F - block_lead;  // the block's first `block_lead` frames are its lead-in (segment 0's)
    if (F > 0 && nseg > F_own / min_frames) nseg = F_own / min_frames;
    if (nseg < 1) nseg = 1;
    tp->n_seg = (int)nseg;
    tp->seg.assign((size_t)SEG_COLS * nseg, 0);
    std::vector<long long> pos(nseg, 0);
    auto hop_of = [&](int64_t fr) { return (int64_t)(host_hop_tab ? host_hop_tab[fr] : a->hop); };
    // offsets of every frame of the block (and one past the end)
    std::vector<int64_t> in_at((size_t)F + 1), out_at((size_t)F + 1);
    in_at[0] = a->in_off0;
    out_at[0] = a->out_off0;
    for (int64_t fr = 0; fr < F; fr++) {
        in_at[fr + 1] = in_at[fr] + hop_of(fr);
        out_at[fr + 1] = out_at[fr] + hop_of(fr);
    }
    int64_t f = block_lead;
    for (int64_t s = 0; s < nseg; s++) {
        int64_t cnt = F_own / nseg + (s < F_own % nseg ? 1 : 0);
        // lead-in: the earlier frames whose output reaches into this segment's first frame (N/hop - 1 of them for a
        // constant hop).  The segment replays them silently, so it starts with the overlap tail the sequential
        // render has there; segment 0 of the block has none (it starts the render, or takes its tail from the
        // block before: tsb_add_tail)
        int64_t lead = 0;
        while (s > 0 && lead < f && out_at[f - lead - 1] + (int64_t)tp->n > out_at[f]) lead++;
        if (lead > f - 1) lead = f > 0 ? f - 1 : 0;  // keep one frame before it for the warm-up analysis (segments hold
                                                    // at least n/hop_min frames, so this never bites)
        if (s == 0) lead = block_lead;              // the block's own lead-in; its warm-up frame lies before the block
        const int64_t f0 = f - lead;
        long long *t = &tp->seg[(size_t)SEG_COLS * s];
        t[0] = f0;
        t[1] = cnt + lead;
        t[2] = in_at[f0];
        t[3] = out_at[f0];
        t[4] = f0 == 0 ? (warmup ? 1 : 0) : 1;
        t[5] = f0 == 0 ? warm_hop : hop_of(f0 - 1);
        t[6] = lead;
        pos[s] = out_at[f];
        f += cnt;
    }
    size_t rows = (size_t)nseg * tp->lanes;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_seg = al(tp->seg.size() * 8), b_pos = al((size_t)nseg * 8), b_m = al(rows * tp->m * 8), b_n = al(rows * tp->n * 8);
    const bool pv = p->desc.program == TSB_PROG_PV_SHIFT;
    const size_t b_inc = pv ? al(tp->lanes * (size_t)(F > 0 ? F : 1) * tp->m * 8) : 0;
    const size_t total = b_seg + b_pos + 3 * b_m + b_n + b_inc;
    // one block for everything; the context keeps it for the next call (allocation costs as much as the render of
    // a single stream).  Reuse is safe in stream order; a second live tsb_timepar gets its own block.
    if (!c->tp_pool_busy) {
        if (c->tp_pool_bytes < total) {
            cudaError_t es = cudaStreamSynchronize(c->stream);
            if (es != cudaSuccess) {
                delete tp;
                return fail(TSB_ECUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(es));
            }
            if (c->tp_pool) cudaFree(c->tp_pool);
            c->tp_pool = nullptr;
            c->tp_pool_bytes = 0;
            if (cudaMalloc(&c->tp_pool, total) == cudaSuccess) c->tp_pool_bytes = total;
        }
        if (c->tp_pool_bytes >= total) {
            tp->block = c->tp_pool;
            tp->from_pool = true;
            c->tp_pool_busy = true;
        }
    }
    if (!tp->block && cudaMalloc(&tp->block, total) != cudaSuccess) {
        tsb_timepar_destroy(tp);
        return fail(TSB_ENOMEM, "time-parallel scratch (%zu bytes)", total);
    }
    char *q = tp->block;
    tp->d_seg = (long long *)q; q += b_seg;
    tp->d_pos = (long long *)q; q += b_pos;
    tp->d_last = (double *)q; q += b_m;
    tp->d_sum = (double *)q; q += b_m;
    tp->d_warm = (double *)q; q += b_m;
    tp->d_ola = (double *)q; q += b_n;
    tp->d_inc = pv ? (double *)q : nullptr;
    cudaError_t e = cudaMemcpyAsync(tp->d_seg, tp->seg.data(), tp->seg.size() * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tp->d_pos, pos.data(), (size_t)nseg * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // seg / pos are host temporaries
    if (e != cudaSuccess) {
        tsb_timepar_destroy(tp);  // gives the pool block back / frees the private one
        return fail(TSB_ECUDA, "time-parallel tables: %s", cudaGetErrorString(e));
    }
    *out = tp;
    return TSB_OK;
}
static int timepar_launch(tsb_timepar *tp, int mode) {
    tsb_ctx *c = tp->ctx;
    tsb_plan *p = tp->plan;
    ChainArgs k = to_args(&tp->args);
    k.n_seg = tp->n_seg;
    k.seg_tab = tp->d_seg;
    k.mode = mode;
    k.st_last_phase = tp->d_last;
    k.st_phase_sum = tp->d_sum;
    k.st_ola = tp->d_ola;
    k.st_warm_phase = tp->d_warm;
    k.st_inc = mode == 1 ? tp->d_inc : nullptr;
    k.inc_frames = tp->args.n_frames;
    k.load_state = 1;
    chain_launch_fn fn = launcher(p->log2n);
    if (!fn) return fail(TSB_EUNSUPPORTED, "no kernel for frame size 2^%d", p->log2n);
    const void *tb = p->desc.precision == TSB_F32 ? (const void *)&p->t32 : (const void *)&p->t64;
    cudaError_t e = fn(p->desc.precision, p->desc.program, tb, k, c->stream);
    if (e != cudaSuccess) return fail(TSB_ECUDA, "chain kernel launch: %s", cudaGetErrorString(e));
    c->launches++;
    return TSB_OK;
}
// row block 0 of a scratch array <- the caller's state (or zeros)
static int seed_rows(tsb_timepar *tp, double *dst, const double *user, size_t width) {
    size_t bytes = tp->lanes * width * 8;
    if (user && tp->args.load_state) CU(cudaMemcpyAsync(dst, user, bytes, cudaMemcpyDeviceToDevice, tp->ctx->stream));
    else CU(cudaMemsetAsync(dst, 0, bytes, tp->ctx->stream));
    return TSB_OK;
}
TSB_API int tsb_timepar_prepass(tsb_timepar *tp, double *block_total, double *block_last_phase) {
    if (!tp) return fail(TSB_EINVAL, "null argument");
    tsb_ctx *c = tp->ctx;
    CU(cudaSetDevice(c->device));
    const size_t rows = (size_t)tp->n_seg * tp->lanes;
    if (tp->plan->desc.program != TSB_PROG_PV_SHIFT || tp->args.n_frames == 0 || tp->lanes == 0) {
        // nothing crosses frames but the overlap tail: no pre-pass needed
        if (block_total) CU(cudaMemsetAsync(block_total, 0, tp->lanes * tp->m * 8, c->stream));
        if (block_last_phase) CU(cudaMemsetAsync(block_last_phase, 0, tp->lanes * tp->m * 8, c->stream));
        CU(cudaMemsetAsync(tp->d_sum, 0, rows * tp->m * 8, c->stream));
        tp->prepassed = true;
        tp->scanned = true;  // no phase state: every segment starts from zeros
        return TSB_OK;
    }
    CU(cudaMemsetAsync(tp->d_last, 0, rows * tp->m * 8, c->stream));
    int r = seed_rows(tp, tp->d_last, tp->args.state_last_phase, tp->m);
    if (r) return r;
    if ((r = timepar_launch(tp, 1))) return r;
    if (block_total)
        CU(op_timepar_scan(tp->d_inc, tp->args.n_frames, tp->d_seg, tp->n_seg, nullptr, nullptr, block_total, (long long)tp->lanes,
                           (long long)tp->m, 0, (long long)tp->lanes, c->stream));
    if (block_last_phase)
        CU(cudaMemcpyAsync(block_last_phase, tp->d_last + (rows - tp->lanes) * tp->m, tp->lanes * tp->m * 8,
                           cudaMemcpyDeviceToDevice, c->stream));
    tp->prepassed = true;
    tp->scanned = false;
    if (block_total) c->launches++;
    return TSB_OK;
}
TSB_API int tsb_timepar_scan_lanes(tsb_timepar *tp, int32_t lane0, int32_t n_lanes, const double *carry_in, double *carry_out) {
    return tsb_timepar_scan_at(tp, lane0, n_lanes, carry_in, carry_out, -1);
}
TSB_API int tsb_timepar_scan_at(tsb_timepar *tp, int32_t lane0, int32_t n_lanes, const double *carry_in, double *carry_out,
                                int64_t carry_frame) {
    if (!tp) return fail(TSB_EINVAL, "null argument");
    if (carry_frame >= tp->args.n_frames) return fail(TSB_EINVAL, "carry_frame %lld beyond the block's %lld frames", (long long)carry_frame, (long long)tp->args.n_frames);
    if (!tp->prepassed) return fail(TSB_EINVAL, "tsb_timepar_scan before tsb_timepar_prepass");
    if (lane0 < 0 || n_lanes < 0 || (size_t)lane0 + (size_t)n_lanes > tp->lanes)
        return fail(TSB_EINVAL, "lanes %d..%d out of range (0..%zu)", lane0, lane0 + n_lanes, tp->lanes);
    tsb_ctx *c = tp->ctx;
    CU(cudaSetDevice(c->device));
    const tsb_chain_args &a = tp->args;
    const size_t bytes = (size_t)n_lanes * tp->m * 8;
    if (tp->plan->desc.program != TSB_PROG_PV_SHIFT || a.n_frames == 0 || n_lanes == 0) {
        if (carry_out && bytes) {
            if (carry_in) CU(cudaMemcpyAsync(carry_out, carry_in, bytes, cudaMemcpyDeviceToDevice, c->stream));
            else CU(cudaMemsetAsync(carry_out, 0, bytes, c->stream));
        }
        tp->scanned = true;
        return TSB_OK;
    }
    const double *base = carry_in;
    if (!base && a.state_phase_sum && a.load_state) base = a.state_phase_sum + (size_t)lane0 * tp->m;
    CU(op_timepar_scan(tp->d_inc, a.n_frames, tp->d_seg, tp->n_seg, base, tp->d_sum, carry_out, (long long)tp->lanes,
                       (long long)tp->m, lane0, n_lanes, c->stream, carry_frame));
    c->launches++;
    tp->scanned = true;
    return TSB_OK;
}
// The same scan as ONE kernel fused with the transfer to the next GPU (a chain of ranks, one process per GPU):
// carry_in / flag_in are buffers in THIS GPU's memory that the previous rank's kernel writes over NVLink (peer-mapped
// through tsb_ipc_*), carry_out_peer / flag_out_peer point into the next rank's; every CTA (16 bins of a lane) waits for
// its own flag, adds its frames and hands its bins on, so the ranks' scans overlap bin block by bin block.
TSB_API int tsb_timepar_scan_chain(tsb_timepar *tp, const double *carry_in, const int32_t *flag_in, double *carry_out_peer,
                                   int32_t *flag_out_peer, int32_t epoch, int64_t carry_frame) {
    if (!tp) return fail(TSB_EINVAL, "null argument");
    if (!tp->prepassed) return fail(TSB_EINVAL, "tsb_timepar_scan_chain before tsb_timepar_prepass");
    if ((carry_in == nullptr) != (flag_in == nullptr) || (carry_out_peer == nullptr) != (flag_out_peer == nullptr))
        return fail(TSB_EINVAL, "carry and flag pointers come in pairs");
    if (carry_frame >= tp->args.n_frames) return fail(TSB_EINVAL, "carry_frame %lld beyond the block's %lld frames", (long long)carry_frame, (long long)tp->args.n_frames);
    tsb_ctx *c = tp->ctx;
    CU(cudaSetDevice(c->device));
    const tsb_chain_args &a = tp->args;
    if (tp->plan->desc.program != TSB_PROG_PV_SHIFT || a.n_frames == 0 || tp->lanes == 0)
        return fail(TSB_EUNSUPPORTED, "the chained scan is for phase-vocoder blocks with frames");
    const double *base = carry_in;
    if (!base && a.state_phase_sum && a.load_state) base = a.state_phase_sum;
    CU(op_timepar_scan(tp->d_inc, a.n_frames, tp->d_seg, tp->n_seg, base, tp->d_sum, carry_out_peer, (long long)tp->lanes,
                       (long long)tp->m, 0, (long long)tp->lanes, c->stream, carry_frame, flag_in, flag_out_peer, epoch));
    c->launches++;
    tp->scanned = true;
    return TSB_OK;
}
// Peer memory between the processes of one box (cudaIpc*): export an allocation made with tsb_malloc, map a peer's.
TSB_API int tsb_ipc_export(tsb_ctx *c, void *dptr, void *handle64) {
    if (!c || !dptr || !handle64) return fail(TSB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return TSB_OK;
}
TSB_API int tsb_ipc_open(tsb_ctx *c, const void *handle64, void **dptr) {
    if (!c || !dptr || !handle64) return fail(TSB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TSB_OK;
}
TSB_API int tsb_ipc_close(tsb_ctx *c, void *dptr) {
    if (!c) return fail(TSB_EINVAL, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaIpcCloseMemHandle(dptr));
    return TSB_OK;
}
TSB_API int tsb_timepar_scan(tsb_timepar *tp, const double *carry_in, double *carry_out) {
    if (!tp) return fail(TSB_EINVAL, "null argument");
    return tsb_timepar_scan_lanes(tp, 0, (int32_t)tp->lanes, carry_in, carry_out);
}
TSB_API int tsb_timepar_render(tsb_timepar *tp) {
    if (!tp) return fail(TSB_EINVAL, "null argument");
    if (!tp->prepassed) return fail(TSB_EINVAL, "tsb_timepar_render before tsb_timepar_prepass");
    tsb_ctx *c = tp->ctx;
    CU(cudaSetDevice(c->device));
    if (tp->args.n_frames == 0 || tp->lanes == 0) return TSB_OK;
    const size_t rows = (size_t)tp->n_seg * tp->lanes;
    const tsb_chain_args &a = tp->args;
    int r;
    if (!tp->scanned && (r = tsb_timepar_scan(tp, nullptr, nullptr))) return r;
    r = seed_rows(tp, tp->d_last, a.state_last_phase, tp->m);  // segment 0 without warm-up starts from the caller's phases
    if (r) return r;
    CU(cudaMemsetAsync(tp->d_ola, 0, rows * tp->n * 8, c->stream));
    if ((r = seed_rows(tp, tp->d_ola, a.state_ola, tp->n))) return r;
    if ((r = timepar_launch(tp, 0))) return r;
    // (no tail fix-up between the segments: each one rebuilt the tail it starts with from its lead-in frames)
    const size_t lastrow = rows - tp->lanes;
    if (a.state_last_phase) CU(cudaMemcpyAsync(a.state_last_phase, tp->d_last + lastrow * tp->m, tp->lanes * tp->m * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (a.state_phase_sum) CU(cudaMemcpyAsync(a.state_phase_sum, tp->d_sum + lastrow * tp->m, tp->lanes * tp->m * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (a.state_ola) CU(cudaMemcpyAsync(a.state_ola, tp->d_ola + lastrow * tp->n, tp->lanes * tp->n * 8, cudaMemcpyDeviceToDevice, c->stream));
    tp->prepassed = false;
    return TSB_OK;
}
TSB_API int tsb_chain_render_timepar(tsb_ctx *c, tsb_plan *p, const tsb_chain_args *a, const int32_t *host_hop_tab,
                                     int32_t n_segments) {
    tsb_timepar *tp = nullptr;
    int r = tsb_timepar_create(c, p, a, host_hop_tab, n_segments, 0, 0, 0, &tp);
    if (r) return r;
    r = tsb_timepar_prepass(tp, nullptr, nullptr);
    if (!r) r = tsb_timepar_render(tp);
    tsb_timepar_destroy(tp);
    return r;
}
TSB_API int tsb_add_tail(tsb_ctx *c, tsb_plan *p, void *out, int64_t out_stride, int64_t out_len, int64_t pos,
                         const double *tail, int64_t lanes) {
    if (!c || !p || !out || !tail) return fail(TSB_EINVAL, "null argument");
    CU(cudaSetDevice(c->device));
    long long *d_pos = nullptr;
    CU(cudaMallocAsync((void **)&d_pos, 8, c->stream));
    long long hp = pos;
    CU(cudaMemcpyAsync(d_pos, &hp, 8, cudaMemcpyHostToDevice, c->stream));
    CU(op_add_tails(out, p->desc.precision == TSB_F64, out_stride, out_len, d_pos, 1, tail, lanes, (long long)p->desc.n, c->stream));
    CU(cudaFreeAsync(d_pos, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // &hp
    c->launches++;
    return TSB_OK;
}

// ---- chunked streaming front end ----
struct tsb_stream {
    tsb_ctx *ctx;
    tsb_plan *plan;
    int lanes;
    double hop;
    SchedState st;
    int64_t pushed = 0;   // input samples received per lane
    int64_t base = 0;     // absolute index of the first sample kept in d_in
    int64_t cap = 0;      // samples per row of d_in
    int64_t frames = 0;
    float *d_in = nullptr;
    void *d_out = nullptr;
    int64_t out_cap = 0;
    int32_t *d_hops = nullptr;
    int64_t hops_cap = 0;
    double *d_last = nullptr, *d_sum = nullptr, *d_ola = nullptr;
    bool started = false;
    int32_t last_hop = 0;  // computed_hop of the last frame rendered (tsb_chain_args::prev_hop of the next call)
};
TSB_API int tsb_stream_destroy(tsb_stream *s) {
    if (!s) return TSB_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->d_in);
    cudaFree(s->d_out);
    cudaFree(s->d_hops);
    cudaFree(s->d_last);
    cudaFree(s->d_sum);
    cudaFree(s->d_ola);
    delete s;
    return TSB_OK;
}
TSB_API int64_t tsb_stream_frames(tsb_stream *s) { return s ? s->frames : 0; }
TSB_API int tsb_stream_create(tsb_ctx *c, tsb_plan *p, int32_t lanes, double hop, double rate, tsb_stream **out) {
    if (!c || !p || !out) return fail(TSB_EINVAL, "null argument");
    if (lanes <= 0) return fail(TSB_EINVAL, "lanes must be positive");
    if (hop < 0) return fail(TSB_EINVAL, "ignoring invalid hop: %g", hop);
    CU(cudaSetDevice(c->device));
    tsb_stream *s = new tsb_stream();
    s->ctx = c;
    s->plan = p;
    s->lanes = lanes;
    s->hop = hop;
    s->st.rate = rate;
    const size_t n = p->desc.n, m = n / 2;
    cudaError_t e = cudaMalloc(&s->d_last, (size_t)lanes * m * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_sum, (size_t)lanes * m * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_ola, (size_t)lanes * n * 8);
    if (e != cudaSuccess) {
        tsb_stream_destroy(s);
        return fail(TSB_ENOMEM, "stream state: %s", cudaGetErrorString(e));
    }
    *out = s;
    return TSB_OK;
}
TSB_API int tsb_stream_process(tsb_stream *s, const float *host_in, int64_t in_stride, int64_t len, void *host_out,
                               int64_t out_stride, int64_t out_cap, int64_t *out_len) {
    if (!s || !out_len) return fail(TSB_EINVAL, "null argument");
    if (len < 0 || (len > 0 && !host_in)) return fail(TSB_EINVAL, "bad input chunk");
    tsb_ctx *c = s->ctx;
    CU(cudaSetDevice(c->device));
    const int64_t n = s->plan->desc.n, lanes = s->lanes;
    const size_t osz = s->plan->desc.precision == TSB_F64 ? 8 : 4;
    *out_len = 0;
    // append the chunk to the pending input (rows hold absolute samples [base, pushed))
    int64_t have = s->pushed - s->base, need = have + len;
    if (need > s->cap) {
        int64_t cap = need + (need >> 1) + 2 * n + 4096;
        cap = (cap + 3) & ~(int64_t)3;
        float *nd = nullptr;
        cudaError_t e = cudaMalloc(&nd, (size_t)lanes * cap * 4);
        if (e != cudaSuccess) return fail(TSB_ENOMEM, "stream input buffer: %s", cudaGetErrorString(e));
        if (have > 0)
            CU(cudaMemcpy2DAsync(nd, (size_t)cap * 4, s->d_in, (size_t)s->cap * 4, (size_t)have * 4, (size_t)lanes,
                                 cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(s->d_in);
        s->d_in = nd;
        s->cap = cap;
    }
    if (len > 0)
        CU(cudaMemcpy2DAsync(s->d_in + have, (size_t)s->cap * 4, host_in, (size_t)in_stride * 4, (size_t)len * 4, (size_t)lanes,
                             cudaMemcpyHostToDevice, c->stream));
    s->pushed += len;
    // frames that can run now (tinyspec.cpp:142-149 on the concatenated input)
    // (on a copy of the scheduler: it is committed only once the frames are on their way, so that an error below
    // leaves the stream where it was and the frames are scheduled again by the next call)
    int64_t out0 = s->st.out_pos;
    SchedState st_next = s->st;
    std::vector<FrameRec> fr = schedule_const(st_next, s->pushed, (uint32_t)n, s->hop, -1);
    if (!fr.empty()) {
        int64_t total = st_next.out_pos - out0;
        if (total > out_cap) return fail(TSB_EINVAL, "out_cap %lld too small for %lld finished samples", (long long)out_cap, (long long)total);
        if (total > s->out_cap) {
            cudaFree(s->d_out);
            s->d_out = nullptr;
            int64_t cap = total + (total >> 1) + 1024;
            cudaError_t e = cudaMalloc(&s->d_out, (size_t)lanes * cap * osz);
            if (e != cudaSuccess) return fail(TSB_ENOMEM, "stream output buffer: %s", cudaGetErrorString(e));
            s->out_cap = cap;
        }
        if ((int64_t)fr.size() > s->hops_cap) {
            cudaFree(s->d_hops);
            s->d_hops = nullptr;
            s->hops_cap = (int64_t)fr.size() * 2 + 64;
            cudaError_t e = cudaMalloc(&s->d_hops, (size_t)s->hops_cap * 4);
            if (e != cudaSuccess) return fail(TSB_ENOMEM, "stream hop table: %s", cudaGetErrorString(e));
        }
        std::vector<int32_t> hops(fr.size());
        for (size_t i = 0; i < fr.size(); i++) hops[i] = fr[i].hop;
        CU(cudaMemcpyAsync(s->d_hops, hops.data(), hops.size() * 4, cudaMemcpyHostToDevice, c->stream));
        tsb_chain_args a;
        memset(&a, 0, sizeof a);
        a.in = s->d_in;
        a.in_stride = s->cap;
        a.in_len = s->pushed - s->base;
        a.out = s->d_out;
        a.out_stride = s->out_cap;
        a.out_len = total;
        a.n_frames = (int64_t)fr.size();
        a.in_off0 = fr[0].in_off - s->base;
        a.out_off0 = 0;
        a.hop = hops[0];
        a.hop_tab = s->d_hops;
        a.state_last_phase = s->d_last;
        a.state_phase_sum = s->d_sum;
        a.state_ola = s->d_ola;
        a.load_state = s->started ? 1 : 0;
        a.lanes = (int32_t)lanes;
        a.prev_hop = s->started ? s->last_hop : 0;  // the kept input starts with the previous frame (see below)
        // positions no frame reaches (hop > frame: the reference emits zeros there, tinyspec.cpp:184-191)
        if (total > 0) CU(cudaMemset2DAsync(s->d_out, (size_t)s->out_cap * osz, 0, (size_t)total * osz, (size_t)lanes, c->stream));
        int r = tsb_chain_render(c, s->plan, &a);
        if (r) return r;
        s->st = st_next;
        s->started = true;
        s->last_hop = hops.back();
        s->frames += (int64_t)fr.size();
        if (total > 0)
            CU(cudaMemcpy2DAsync(host_out, (size_t)out_stride * osz, s->d_out, (size_t)s->out_cap * osz, (size_t)total * osz,
                                 (size_t)lanes, cudaMemcpyDeviceToHost, c->stream));
        *out_len = total;
        // drop what no later frame can read: the next frame starts at in_pos + computed_hop >= in_pos
        int64_t keep_from = s->st.in_pos;
        if (keep_from > s->base) {
            int64_t left = s->pushed - keep_from;
            CU(cudaStreamSynchronize(c->stream));
            if (left > 0) {
                // rows overlap themselves when shifting left: go through the output buffer's neighbour-free scratch
                float *tmp = nullptr;
                CU(cudaMalloc(&tmp, (size_t)lanes * left * 4));
                CU(cudaMemcpy2DAsync(tmp, (size_t)left * 4, s->d_in + (keep_from - s->base), (size_t)s->cap * 4, (size_t)left * 4,
                                     (size_t)lanes, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpy2DAsync(s->d_in, (size_t)s->cap * 4, tmp, (size_t)left * 4, (size_t)left * 4, (size_t)lanes,
                                     cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                cudaFree(tmp);
            }
            s->base = keep_from;
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    return TSB_OK;
}
