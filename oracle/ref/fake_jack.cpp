// oracle/ref/fake_jack.cpp — TEST INFRASTRUCTURE (not product code).
//
// An offline, deterministic stand-in for the JACK server, so that the UNMODIFIED reference
// (/root/reference/tinyspec.cpp, compiled where it lies with -DHACK=<tap>) can be rendered without
// audio hardware.  It defines the 13 JACK entry points the reference calls (declared in
// stubs/jack/jack.h) and drives the registered process callback (tinyspec.cpp:205-239) from a
// thread with an exact idle handshake: the reference worker holds internals::exec_mtx for one
// whole frame iteration (tinyspec.cpp:132), so when we hold exec_mtx + data_mtx and the
// starve/full predicate of tinyspec.cpp:142-145 is true, the worker can make no further progress
// until the next callback.  Frame contents never depend on timing; with the handshake the
// output stream does not either.
//
// Environment (all optional):
//   TSREF_RATE       sample rate reported to the reference            (48000)
//   TSREF_SAMPLES    input samples per channel to feed                (48000)
//   TSREF_LEN        JACK period, samples per callback                (1024)
//   TSREF_STREAM     stream index s of the synthetic signal           (0)
//   TSREF_INPUT      raw interleaved float32 file to feed instead of the generator
//   TSREF_OUT        file receiving the valid (non-underrun) output stream, interleaved float32
//   TSREF_DUMP_DIR   directory for tap dumps written through tsref_dump()
//   TSREF_MAX_FRAMES stop after the tap reported this many frames (for input-less hacks)
//   TSREF_TIMING     1 = print wall time first->last process callback + frame count as JSON
#include <jack/jack.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../signal_gen.h"
#include "rtinc/buffer.hpp"  // the reference's own Buf<T> (needed to name internals::fptr's type)

namespace internals {
extern uint64_t system_frame_size, window_size, new_window_size, computed_hop;
extern std::deque<float> aqueue, atmp, ainqueue;
extern size_t nch_in, nch_out, new_nch_in, new_nch_out;
extern std::mutex exec_mtx, data_mtx;
extern std::function<void(Buf<double> &, Buf<double> &, double)> fptr;
}  // namespace internals

struct _jack_port {
    std::string name;
    unsigned long flags;
    std::vector<float> buf;
};
struct _jack_client {
    std::string name;
};

namespace {
_jack_client g_client;
std::vector<_jack_port *> g_in_ports, g_out_ports;
JackProcessCallback g_process = nullptr;
JackSampleRateCallback g_srate = nullptr;
std::atomic<bool> g_go{false};
std::atomic<long> g_frames{0};
std::map<std::string, FILE *> g_dumps;
std::thread g_driver;

long env_long(const char *k, long d) {
    const char *v = getenv(k);
    return v && *v ? atol(v) : d;
}
double env_dbl(const char *k, double d) {
    const char *v = getenv(k);
    return v && *v ? atof(v) : d;
}

// true when the reference worker cannot run another frame (tinyspec.cpp:142-145), evaluated with
// the frame size the worker would latch next (tinyspec.cpp:135-139).
bool worker_idle(uint64_t len) {
    std::lock_guard<std::mutex> e(internals::exec_mtx);
    std::lock_guard<std::mutex> d(internals::data_mtx);
    uint64_t ws = internals::new_window_size;
    size_t nin = internals::new_nch_in, nout = internals::new_nch_out;
    bool starved = internals::ainqueue.size() < nin * (internals::computed_hop + ws);
    uint64_t sys = internals::system_frame_size ? internals::system_frame_size : len;
    bool full = internals::aqueue.size() >= std::max(sys, ws) * nout;
    return starved || full;
}
void wait_idle(uint64_t len) {
    // The worker is woken asynchronously (condvar); give it the chance to consume the wake-up,
    // then require the predicate to hold on several consecutive looks.
    int stable = 0;
    while (stable < 3) {
        if (worker_idle(len)) {
            stable++;
            std::this_thread::sleep_for(std::chrono::microseconds(stable == 1 ? 50 : 20));
        } else {
            stable = 0;
            std::this_thread::yield();
        }
    }
}

void driver_main() {
    const double rate = env_dbl("TSREF_RATE", 48000.0);
    const long total = env_long("TSREF_SAMPLES", 48000);
    const long len = env_long("TSREF_LEN", 1024);
    const int stream = (int)env_long("TSREF_STREAM", 0);
    const long max_frames = env_long("TSREF_MAX_FRAMES", -1);
    const bool timing = env_long("TSREF_TIMING", 0) != 0;
    const char *in_path = getenv("TSREF_INPUT");
    const char *out_path = getenv("TSREF_OUT");

    if (g_srate) g_srate((jack_nframes_t)rate, nullptr);
    while (!g_go.load()) std::this_thread::sleep_for(std::chrono::milliseconds(1));

    const size_t nin = g_in_ports.size(), nout = g_out_ports.size();
    FILE *fin = in_path ? fopen(in_path, "rb") : nullptr;
    FILE *fout = out_path ? fopen(out_path, "wb") : nullptr;
    if (in_path && !fin) { fprintf(stderr, "fake_jack: cannot open %s\n", in_path); _Exit(2); }
    ts_siggen gen;
    ts_siggen_init(&gen, stream, (int)nin, rate);
    std::vector<float> inter(nin * len), outi(nout * len);

    auto t0 = std::chrono::steady_clock::now();
    long fed = 0;
    uint64_t out_written = 0;
    auto drain_ports = [&](long n, size_t valid) {
        // `valid` interleaved samples at the head of this period are real output, the rest is
        // underrun zero-fill (tinyspec.cpp:222-232).
        if (!fout) return;
        size_t k = 0;
        for (long i = 0; i < n && k < valid; i++)
            for (size_t c = 0; c < nout && k < valid; c++, k++) outi[k] = g_out_ports[c]->buf[i];
        fwrite(outi.data(), sizeof(float), k, fout);
        out_written += k;
    };
    while (true) {
        bool more_input = nin > 0 && fed < total;
        if (nin > 0 && !more_input) break;
        if (max_frames >= 0 && g_frames.load() >= max_frames) break;
        long n = nin > 0 ? std::min(len, total - fed) : len;
        if (nin > 0) {
            if (fin) {
                size_t got = fread(inter.data(), sizeof(float), nin * n, fin);
                for (size_t i = got; i < nin * (size_t)n; i++) inter[i] = 0.f;
            } else {
                ts_siggen_fill(&gen, inter.data(), n);
            }
            for (size_t c = 0; c < nin; c++) {
                g_in_ports[c]->buf.resize(n);
                for (long i = 0; i < n; i++) g_in_ports[c]->buf[i] = inter[i * nin + c];
            }
        }
        for (size_t c = 0; c < nout; c++) g_out_ports[c]->buf.assign(n, 0.f);
        size_t avail;
        {
            std::lock_guard<std::mutex> d(internals::data_mtx);
            avail = internals::aqueue.size();
        }
        g_process((jack_nframes_t)n, nullptr);
        drain_ports(n, std::min(avail, nout * (size_t)n));
        fed += n;
        wait_idle(len);
    }
    // Input exhausted: keep waking the worker (a zero-length period is a pure wake-up) and pull
    // what it produces until it is starved with an empty output queue.
    for (int rounds = 0; rounds < 1000000; rounds++) {
        size_t pulled = 0;
        {
            std::lock_guard<std::mutex> d(internals::data_mtx);
            pulled = internals::aqueue.size();
            if (fout) {
                std::vector<float> tmp(internals::aqueue.begin(), internals::aqueue.end());
                fwrite(tmp.data(), sizeof(float), tmp.size(), fout);
                out_written += tmp.size();
            }
            internals::aqueue.clear();
        }
        if (max_frames >= 0) break;
        long before = g_frames.load();
        for (size_t c = 0; c < nout; c++) g_out_ports[c]->buf.clear();
        for (size_t c = 0; c < nin; c++) g_in_ports[c]->buf.clear();
        g_process(0, nullptr);
        wait_idle(len);
        if (pulled == 0 && g_frames.load() == before) {
            std::lock_guard<std::mutex> d(internals::data_mtx);
            if (internals::aqueue.empty()) break;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (timing) {
        double sec = std::chrono::duration<double>(t1 - t0).count();
        printf("{\"frames\": %ld, \"wall_s\": %.6f, \"in_samples\": %ld, \"out_samples\": %llu, "
               "\"nch_in\": %zu, \"nch_out\": %zu}\n",
               g_frames.load(), sec, fed, (unsigned long long)out_written, nin, nout);
    }
    if (fout) fclose(fout);
    for (auto &kv : g_dumps) fclose(kv.second);
    fflush(nullptr);
    _Exit(0);  // main() of the reference sleeps forever (tinyspec.cpp:290)
}
}  // namespace

// ---- hooks used by the tap hacks (declared at block scope inside the hack body) ----
void tsref_go() { g_go.store(true); }
void tsref_frame_done() { g_frames.fetch_add(1); }
long tsref_env_long(const char *k, long d) { return env_long(k, d); }
double tsref_env_double(const char *k, double d) { return env_dbl(k, d); }
void tsref_finish() {
    for (auto &kv : g_dumps) fclose(kv.second);
    fflush(nullptr);
    _Exit(0);
}
void tsref_dump(const char *name, const void *p, size_t bytes) {
    const char *dir = getenv("TSREF_DUMP_DIR");
    if (!dir) return;
    FILE *&f = g_dumps[name];
    if (!f) {
        std::string path = std::string(dir) + "/" + name + ".bin";
        f = fopen(path.c_str(), "wb");
        if (!f) { fprintf(stderr, "fake_jack: cannot open %s\n", path.c_str()); _Exit(2); }
    }
    fwrite(p, 1, bytes, f);
}

// ---- JACK API ----
extern "C" {
jack_client_t *jack_client_open(const char *name, jack_options_t, jack_status_t *status, ...) {
    g_client.name = name ? name : "tsref";
    if (status) *status = 0;
    return &g_client;
}
char *jack_get_client_name(jack_client_t *c) { return const_cast<char *>(c->name.c_str()); }
int jack_set_process_callback(jack_client_t *, JackProcessCallback cb, void *) { g_process = cb; return 0; }
int jack_set_sample_rate_callback(jack_client_t *, JackSampleRateCallback cb, void *) { g_srate = cb; return 0; }
int jack_activate(jack_client_t *) {
    g_driver = std::thread(driver_main);
    g_driver.detach();
    return 0;
}
int jack_deactivate(jack_client_t *) { return 0; }
void *jack_port_get_buffer(jack_port_t *port, jack_nframes_t n) {
    if (port->buf.size() < n) port->buf.resize(n, 0.f);
    return port->buf.data();
}
jack_port_t *jack_port_register(jack_client_t *, const char *port_name, const char *, unsigned long flags,
                                unsigned long) {
    _jack_port *p = new _jack_port{port_name, flags, {}};
    (flags & JackPortIsInput ? g_in_ports : g_out_ports).push_back(p);
    return p;
}
int jack_port_unregister(jack_client_t *, jack_port_t *port) {
    auto &v = (port->flags & JackPortIsInput) ? g_in_ports : g_out_ports;
    for (size_t i = 0; i < v.size(); i++)
        if (v[i] == port) { v.erase(v.begin() + i); break; }
    delete port;
    return 0;
}
const char **jack_get_ports(jack_client_t *, const char *, const char *, unsigned long) { return nullptr; }
int jack_connect(jack_client_t *, const char *, const char *) { return 0; }
int jack_disconnect(jack_client_t *, const char *, const char *) { return 0; }
void jack_free(void *) {}
}
