// rtinc/driver.hpp — the offline frame scheduler: what tinyspec.cpp's generate_frames + JACK callback do
// (tinyspec.cpp:113-239), for a whole render at once and for many independent streams.
//
//   fused mode      per frame the hack's callback runs on symbolic buffers and records its helper calls;
//                   runs of frames with the same record become one tsb_chain_render launch over
//                   frames x lanes (lane = channel of a stream).  PV state and the overlap tail stay on the device.
//   immediate mode  the reference's loop verbatim: gather (device) -> callback with real buffers whose helpers
//                   are device operators -> overlap-add (device, float32 tail exactly as tinyspec.cpp:170-200).
#pragma once
#include <chrono>
#include <memory>

#include "control.hpp"
#include "pv.hpp"

namespace tsbrt {

struct Options {
    double seconds = 5.0, rate = 48000.0;
    int streams = 1, precision = TSB_F64, device = 0;
    std::string mode = "fused", in_path, out_path;
    long long max_frames = -1;
    bool timing = false, out_f64 = false, timepar = true;
};
inline Options parse_args(int argc, char **argv) {
    Options o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) fatal("missing value after " + a);
            return argv[++i];
        };
        if (a == "--seconds") o.seconds = atof(next().c_str());
        else if (a == "--rate") o.rate = atof(next().c_str());
        else if (a == "--streams") o.streams = atoi(next().c_str());
        else if (a == "--precision") o.precision = next() == "f32" ? TSB_F32 : TSB_F64;
        else if (a == "--mode") o.mode = next();
        else if (a == "--in") o.in_path = next();
        else if (a == "--out") o.out_path = next();
        else if (a == "--out-f64") o.out_f64 = true;
        else if (a == "--no-timepar") o.timepar = false;
        else if (a == "--max-frames") o.max_frames = atoll(next().c_str());
        else if (a == "--device") o.device = atoi(next().c_str());
        else if (a == "--timing") o.timing = true;
        else fatal("unknown argument " + a + "\nusage: ts_render [--seconds S] [--rate R] [--streams N] [--precision f32|f64] "
                   "[--mode fused|immediate] [--in planar.f32] [--out planar.f32] [--out-f64] [--no-timepar] [--max-frames K] [--timing]");
    }
    if (o.mode != "fused" && o.mode != "immediate") fatal("--mode must be fused or immediate");
    return o;
}

// SURVEY.md §8d synthetic input: 0.5*sin(2*pi*f*n/rate) + 0.1*(u-0.5), f = 440 + 110*c + 0.37*s, 32-bit LCG noise
inline void synth_stream(int s, int nch, double rate, size_t len, float *planar, size_t stride) {
    uint32_t lcg = 12345u + (uint32_t)s;
    for (size_t n = 0; n < len; n++)
        for (int c = 0; c < nch; c++) {
            lcg = lcg * 1664525u + 1013904223u;
            double u = (double)(lcg >> 8) / 16777216.0;
            double f = 440.0 + 110.0 * c + 0.37 * s;
            planar[(size_t)c * stride + n] = (float)(0.5 * sin(2.0 * M_PI * f * (double)n / rate) + 0.1 * (u - 0.5));
        }
}

struct DevBuf {
    void *d = nullptr;
    size_t bytes = 0;
    void ensure(size_t n, bool zero = false) {
        Runtime &rt = Runtime::get();
        if (bytes < n) {
            if (d) tsb_free(rt.context(), d);
            check(tsb_malloc(rt.context(), n ? n : 16, &d), "tsb_malloc");
            bytes = n;
            zero = true;
        }
        if (zero && bytes) check(tsb_memset(rt.context(), d, 0, bytes), "tsb_memset");
    }
    ~DevBuf() {
        if (d && Runtime::get().ctx) tsb_free(Runtime::get().ctx, d);
    }
};

struct SegmentKey {
    int program = -1;
    uint32_t n = 0;
    double hop_d = 0, ratio = 1;
    int win_in = 0, win_out = 0;
    std::vector<double> fn_bins;
    const void *pv = nullptr;
    int pv_epoch = 0;
    bool operator==(const SegmentKey &o) const {
        return program == o.program && n == o.n && hop_d == o.hop_d && ratio == o.ratio && win_in == o.win_in &&
               win_out == o.win_out && fn_bins == o.fn_bins && pv == o.pv && pv_epoch == o.pv_epoch;
    }
};

struct Renderer {
    Options opt;
    WaveBuf audio_in, audio_out;  // reference: tinyspec.cpp:44
    size_t lanes_in = 0, lanes_out = 0;
    int64_t in_len = 0, out_cap = 0, out_pos = 0;
    std::vector<float> host_in;    // planar [lanes_in][in_len]
    std::vector<float> host_out;   // planar [lanes_out][out_cap]
    std::vector<double> host_out64;
    long long frames = 0, launches0 = 0, timepar_launches = 0;
    double wall = 0;

    explicit Renderer(const Options &o) : opt(o) {}

    // what tinyspec.cpp's init_audio does before the hack body runs (:245-258)
    void init() {
        Runtime::get().device = opt.device;
        internals::rate = opt.rate;
        set_num_channels(2, 2);
    }
    bool latch_and_check() {  // reference: tinyspec.cpp:135-145
        using namespace internals;
        if (new_nch_in != nch_in || new_nch_out != nch_out || new_window_size != window_size) {
            window_size = new_window_size;
            nch_in = new_nch_in;
            nch_out = new_nch_out;
        }
        audio_in.resize(nch_in, window_size);
        audio_out.resize(nch_out, window_size);
        if (nch_in > 0 && internals::in_len - internals::in_pos < (int64_t)(computed_hop + window_size)) return false;
        if (nch_in == 0 && out_pos >= (int64_t)(opt.seconds * opt.rate)) return false;
        if (opt.max_frames >= 0 && frames >= opt.max_frames) return false;
        return true;
    }
    void advance_clock() {  // reference: tinyspec.cpp:163-168
        using namespace internals;
        if (time_fract + hop_fract < time_fract) computed_hop = hop_samples + 1;
        else computed_hop = hop_samples;
        time_samples += computed_hop;
        time_fract += hop_fract;
    }
    void load_input() {
        size_t nin = internals::new_nch_in, nout = internals::new_nch_out;
        lanes_in = nin * opt.streams;
        lanes_out = nout * opt.streams;
        in_len = (int64_t)(opt.seconds * opt.rate);
        host_in.assign(lanes_in * (size_t)in_len, 0.f);
        if (!opt.in_path.empty()) {
            FILE *f = fopen(opt.in_path.c_str(), "rb");
            if (!f) fatal("cannot open " + opt.in_path);
            fseek(f, 0, SEEK_END);
            long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            if (lanes_in == 0 || sz % (long)(4 * lanes_in)) fatal("input file is not planar float32 [lanes][len]");
            in_len = sz / (long)(4 * lanes_in);
            host_in.resize(lanes_in * (size_t)in_len);
            if (fread(host_in.data(), 4, host_in.size(), f) != host_in.size()) fatal("short read on " + opt.in_path);
            fclose(f);
        } else {
            for (int s = 0; s < opt.streams; s++)
                if (nin) synth_stream(s, (int)nin, opt.rate, (size_t)in_len, host_in.data() + (size_t)s * nin * in_len, (size_t)in_len);
        }
        internals::in_len = in_len;
        internals::in_pos = 0;
        out_cap = (nin ? in_len : (int64_t)(opt.seconds * opt.rate)) + (1 << 18);
    }

    // ------------------------------------------------------------------ fused mode
    struct Pending {
        SegmentKey key;
        int64_t in_off0 = 0, out_off0 = 0;
        std::vector<int32_t> hops;
    };
    DevBuf d_in, d_out, d_hops, d_last, d_sum, d_ola;
    std::map<std::string, tsb_plan *> plans;
    bool state_valid = false;
    int32_t last_hop = 0;  // computed_hop of the last frame rendered so far
    const void *state_pv = nullptr;
    int state_epoch = -1;
    uint32_t state_n = 0;

    static bool classify(const std::vector<Op> &tr, const WaveBuf *in, const WaveBuf *out, SegmentKey &k, std::string &why) {
        for (const Op &op : tr)
            if (op.kind == OP_OTHER) {
                why = op.why;
                return false;
            }
        if (tr.size() == 3 && tr[0].kind == OP_PV_ANALYZE_WAVE && tr[1].kind == OP_PV_SHIFT && tr[2].kind == OP_PV_SYNTH_WAVE &&
            tr[0].a == tr[1].a && tr[0].a == tr[2].a && tr[0].b == in && tr[2].b == out) {
            k.program = TSB_PROG_PV_SHIFT;
            k.pv = tr[0].a;
            k.hop_d = tr[0].p;
            k.ratio = tr[1].p;
            k.fn_bins = tr[1].fn_bins;
            k.win_in = k.win_out = TSB_WIN_SQRT_HANN;
            if (tr[2].p != tr[0].p) {
                why = "PhaseVocoder::hop() changed between analyze and synthesize";
                return false;
            }
            return true;
        }
        // window(in); fft.fill_from(in); frft(fft,fft,+1); frft(fft,fft,-1); out.fill_from(fft, real); window(out)
        size_t i = 0;
        int w_in = TSB_WIN_NONE, w_out = TSB_WIN_NONE;
        if (i < tr.size() && tr[i].kind == OP_WINDOW && tr[i].a == in) w_in = (int)tr[i++].p;
        if (i + 3 < tr.size() && tr[i].kind == OP_WIDEN && tr[i].a == in && tr[i + 1].kind == OP_FRFT && tr[i + 1].p == 1.0 &&
            tr[i + 1].a == tr[i].b && tr[i + 1].b == tr[i].b && tr[i + 2].kind == OP_FRFT && tr[i + 2].p == -1.0 &&
            tr[i + 2].a == tr[i].b && tr[i + 2].b == tr[i].b && tr[i + 3].kind == OP_REAL && tr[i + 3].a == tr[i].b &&
            tr[i + 3].b == out) {
            i += 4;
            if (i < tr.size() && tr[i].kind == OP_WINDOW && tr[i].a == out) w_out = (int)tr[i++].p;
            if (i == tr.size()) {
                k.program = TSB_PROG_GAIN;
                k.win_in = w_in;
                k.win_out = w_out;
                k.hop_d = internals::hop;
                return true;
            }
        }
        why = "the callback's helper calls are not a chain the library fuses (PhaseVocoder analyze -> shift -> synthesize, "
              "or window -> frft(+1) -> frft(-1) -> real part -> window)";
        return false;
    }

    // fp64 arithmetic delivers the reference's float32 stream (float32 overlap tail and narrowing done on the
    // device exactly as tinyspec.cpp:173-200) unless the unrounded float64 accumulation was asked for (--out-f64)
    int chain_precision() const { return (opt.precision == TSB_F64 && !opt.out_f64) ? (int)TSB_F64_OUT32 : opt.precision; }
    size_t out_elem_size() const { return chain_precision() == TSB_F64 ? 8 : 4; }

    void flush(Pending &p) {
        if (p.hops.empty()) return;
        Runtime &rt = Runtime::get();
        const SegmentKey &k = p.key;
        const size_t lanes = lanes_in, n = k.n, esz = out_elem_size();
        // plan (cached per distinct chain)
        char buf[256];
        snprintf(buf, sizeof buf, "%d|%u|%.17g|%.17g|%d|%d|%zu|", k.program, k.n, k.hop_d, k.ratio, k.win_in, k.win_out, k.fn_bins.size());
        std::string id = buf;
        for (double v : k.fn_bins) id.append((const char *)&v, 8);
        tsb_plan *&plan = plans[id];
        if (!plan) {
            tsb_plan_desc d;
            memset(&d, 0, sizeof d);
            d.n = k.n;
            d.precision = chain_precision();
            d.program = k.program;
            d.win_in = k.win_in;
            d.win_out = k.win_out;
            d.hop = k.hop_d > 0 ? k.hop_d : 1.0;
            d.rate = internals::rate;
            d.ratio = k.ratio;
            d.fn_bins = k.fn_bins.empty() ? nullptr : k.fn_bins.data();
            check(tsb_plan_create(rt.context(), &d, &plan), "tsb_plan_create");
        }
        // state: PhaseVocoder::resize / reset zero it (rtinc/pv.hpp:23-35); a size change also drops it
        // (reset(last, sum) zeroes only what it names, rtinc/pv.hpp:32-35; a resize zeroes both)
        const bool same_pv = state_valid && state_pv == k.pv && state_n == n;
        bool zero_last = !same_pv, zero_sum = !same_pv;
        PhaseVocoder *pvp = (PhaseVocoder *)k.pv;
        if (same_pv && state_epoch != k.pv_epoch) {
            zero_last = pvp->_reset_last;
            zero_sum = pvp->_reset_sum;
        }
        if (pvp) pvp->_reset_last = pvp->_reset_sum = false;
        bool keep_ola = state_valid && state_n == n;
        if (state_valid && state_n != n) fatal("frame size changed from " + std::to_string(state_n) + " to " + std::to_string(n) +
                                               " mid-render: fused mode needs a constant frame size, use --mode immediate");
        d_last.ensure(lanes * (n / 2) * 8, zero_last);
        d_sum.ensure(lanes * (n / 2) * 8, zero_sum);
        d_ola.ensure(lanes * n * 8, !keep_ola);
        tsb_chain_args a;
        memset(&a, 0, sizeof a);
        a.in = (const float *)d_in.d;
        a.in_stride = in_len;
        a.in_len = in_len;
        a.out = d_out.d;
        a.out_stride = out_cap;
        a.out_len = out_cap;
        a.n_frames = (int64_t)p.hops.size();
        a.in_off0 = p.in_off0;
        a.out_off0 = p.out_off0;
        bool constant = true;
        for (int32_t h : p.hops) constant = constant && h == p.hops[0];
        a.hop = p.hops[0];
        if (!constant) {
            d_hops.ensure(p.hops.size() * 4);
            check(tsb_upload(rt.context(), d_hops.d, p.hops.data(), p.hops.size() * 4), "tsb_upload");
            a.hop_tab = (const int32_t *)d_hops.d;
        }
        a.state_last_phase = (double *)d_last.d;
        a.state_phase_sum = (double *)d_sum.d;
        a.state_ola = (double *)d_ola.d;
        a.load_state = 1;  // arrays were zeroed above when the state does not carry over
        a.lanes = (int32_t)lanes;
        // the frame before this run (its samples are in the same rows): lets the float32 chain decide doubtful wraps exactly
        a.prev_hop = (same_pv && !zero_last) ? last_hop : 0;
        last_hop = p.hops.back();
        // A few lanes cannot fill the GPU frame after frame (a lane is one CTA walking its frames in order): cut the
        // run into concurrent time segments instead — samples and PV state bit-identical to the sequential launch.
        const int64_t frames_run = (int64_t)p.hops.size();
        const size_t inc_bytes = lanes * (size_t)frames_run * (n / 2) * 8;
        int32_t nseg = (int32_t)((148 * 4 + lanes - 1) / lanes);
        if (nseg > frames_run / 16) nseg = (int32_t)(frames_run / 16);
        if (opt.timepar && nseg >= 2 && inc_bytes <= ((size_t)4 << 30)) {
            check(tsb_chain_render_timepar(rt.context(), plan, &a, constant ? nullptr : p.hops.data(), nseg),
                  "tsb_chain_render_timepar");
            timepar_launches++;
        } else {
            check(tsb_chain_render(rt.context(), plan, &a), "tsb_chain_render");
        }
        rt.sync();  // p.hops is about to go away
        state_valid = true;
        state_pv = k.pv;
        state_epoch = k.pv_epoch;
        state_n = (uint32_t)n;
        (void)esz;
        p.hops.clear();
    }

    void render_fused() {
        Runtime &rt = Runtime::get();
        if (internals::new_nch_in == 0) fatal("fused mode needs an input (a hack without inputs fills its spectrum itself): use --mode immediate");
        const size_t esz = out_elem_size();
        d_in.ensure(host_in.size() * 4);
        check(tsb_upload(rt.context(), d_in.d, host_in.data(), host_in.size() * 4), "tsb_upload");
        d_out.ensure(lanes_out * (size_t)out_cap * esz, true);
        Pending pend;
        while (latch_and_check()) {
            using namespace internals;
            in_pos += (int64_t)computed_hop;  // reference: tinyspec.cpp:147-149
            int64_t frame_in = in_pos, frame_out = out_pos;
            rt.tracing = true;
            rt.raw_access = false;
            rt.trace.clear();
            if (!fptr) fatal("the hack never called set_process_fn");
            fptr(audio_in, audio_out, current_time_secs());
            rt.tracing = false;
            advance_clock();
            SegmentKey k;
            std::string why;
            if (rt.raw_access) fatal("the frame callback reads or writes samples itself (operator[]), which cannot be batched: use --mode immediate");
            if (!classify(rt.trace, &audio_in, &audio_out, k, why)) fatal(why + ": use --mode immediate");
            k.n = (uint32_t)window_size;
            if (k.pv) k.pv_epoch = ((const PhaseVocoder *)k.pv)->_state_epoch;
            if (queues_cleared && frames > 0)
                fatal("set_num_channels mid-render drops the queued input (rtlib/control.hpp:56-57): how much that is depends on "
                      "the audio period, which an offline render does not have; call it before the first frame");
            queues_cleared = false;
            if (nch_in != nch_out) fatal("fused chains need as many output as input channels");
            if (audio_out.size != window_size || audio_out.num_channels != nch_out) fatal("fused chains need out to keep the frame's shape");
            if (!(k == pend.key)) {
                flush(pend);
                pend.key = k;
                pend.in_off0 = frame_in;
                pend.out_off0 = frame_out;
            }
            pend.hops.push_back((int32_t)computed_hop);
            out_pos += (int64_t)computed_hop;
            frames++;
            if (out_pos + (int64_t)window_size > out_cap) fatal("output capacity exceeded");
        }
        flush(pend);
        rt.sync();
        size_t total = lanes_out * (size_t)out_cap;
        if (esz == 4) {
            host_out.resize(total);
            check(tsb_download(rt.context(), host_out.data(), d_out.d, total * 4), "tsb_download");
            rt.sync();
        } else {
            host_out64.resize(total);
            check(tsb_download(rt.context(), host_out64.data(), d_out.d, total * 8), "tsb_download");
            rt.sync();
            host_out.resize(total);
            for (size_t i = 0; i < total; i++) host_out[i] = (float)host_out64[i];
        }
    }

    // ------------------------------------------------------------------ immediate mode
    void render_immediate() {
        Runtime &rt = Runtime::get();
        if (opt.streams != 1) fatal("immediate mode renders one stream (a hack's state is per process in the reference too)");
        const int64_t tail_cap = 1 << 19;
        DevBuf tail[2], dout;
        d_in.ensure(host_in.size() * 4 + 16);
        if (!host_in.empty()) check(tsb_upload(rt.context(), d_in.d, host_in.data(), host_in.size() * 4), "tsb_upload");
        size_t nout_max = internals::new_nch_out ? internals::new_nch_out : 1;
        dout.ensure(nout_max * (size_t)out_cap * 4, true);
        tail[0].ensure(nout_max * tail_cap * 4, true);
        tail[1].ensure(nout_max * tail_cap * 4, true);
        int cur = 0;
        int64_t tail_len = 0;
        while (latch_and_check()) {
            using namespace internals;
            if (nch_out > nout_max) fatal("set_num_channels raised the output channel count mid-render");
            in_pos += (int64_t)computed_hop;
            if (nch_in) {
                check(tsb_gather_frame(rt.context(), (const float *)d_in.d, in_len, internals::in_len, in_pos,
                                       (double *)audio_in._to_device(), (int64_t)nch_in, (int64_t)window_size),
                      "tsb_gather_frame");
                audio_in._dev_written();
            }
            if (!fptr) fatal("the hack never called set_process_fn");
            fptr(audio_in, audio_out, current_time_secs());
            advance_clock();
            int64_t osize = (int64_t)audio_out.size, hopn = (int64_t)computed_hop, new_len = 0;
            if (out_pos + hopn > out_cap) fatal("output capacity exceeded");
            if (nch_out) {
                check(tsb_overlap_add(rt.context(), (const double *)audio_out._to_device(), (int64_t)audio_out.num_channels, osize,
                                      hopn, (float *)dout.d, out_cap, out_cap, out_pos, (const float *)tail[cur].d,
                                      (float *)tail[cur ^ 1].d, tail_cap, tail_len, &new_len),
                      "tsb_overlap_add");
                cur ^= 1;
                tail_len = new_len;
            }
            out_pos += hopn;
            frames++;
        }
        rt.sync();
        lanes_out = internals::nch_out;
        host_out.resize(lanes_out * (size_t)out_cap);
        if (!host_out.empty()) check(tsb_download(rt.context(), host_out.data(), dout.d, host_out.size() * 4), "tsb_download");
        rt.sync();
    }

    int run() {
        load_input();
        Runtime &rt = Runtime::get();
        rt.context();
        auto t0 = std::chrono::steady_clock::now();
        launches0 = tsb_launch_count(rt.ctx);
        if (opt.mode == "fused") render_fused();
        else render_immediate();
        wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!opt.out_path.empty()) {
            FILE *f = fopen(opt.out_path.c_str(), "wb");
            if (!f) fatal("cannot open " + opt.out_path);
            bool f64 = opt.out_f64 && !host_out64.empty();
            for (size_t l = 0; l < lanes_out; l++) {  // planar [lanes][out_pos]
                if (f64) fwrite(host_out64.data() + l * (size_t)out_cap, 8, (size_t)out_pos, f);
                else fwrite(host_out.data() + l * (size_t)out_cap, 4, (size_t)out_pos, f);
            }
            fclose(f);
        }
        if (opt.timing)
            printf("{\"frames\": %lld, \"wall_s\": %.6f, \"streams\": %d, \"lanes\": %zu, \"out_samples_per_lane\": %lld, "
                   "\"mode\": \"%s\", \"precision\": \"%s\", \"gpu_launches\": %lld, \"time_parallel_renders\": %lld}\n",
                   frames, wall, opt.streams, lanes_out, (long long)out_pos, opt.mode.c_str(),
                   opt.precision == TSB_F32 ? "f32" : "f64", (long long)(tsb_launch_count(rt.ctx) - launches0),
                   timepar_launches);
        for (auto &kv : plans) tsb_plan_destroy(kv.second);
        return 0;
    }
};

}  // namespace tsbrt
