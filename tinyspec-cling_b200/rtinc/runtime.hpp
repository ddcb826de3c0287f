// rtinc/runtime.hpp — the device runtime behind the source-compatible helper headers.
//
// The reference (nwoeanhinnogaehr/tinyspec-cling) compiles a "hack" as statements inside main() against
// rtinc/root.hpp and runs its frame callback on one worker thread (tinyspec.cpp:113-203,288-289).  This
// header set keeps those names and semantics but executes on a B200 through libtsb200 (include/tsb200.h):
//
//   * immediate mode: every helper call (window_*, frft, PhaseVocoder::*) is one device operator on the
//     buffer's device copy; Buf<T> keeps a host and a device copy and moves data only when the other side
//     is touched (operator[] = host access).  Works for any hack; one launch per helper per frame.
//   * fused mode: the callback is run per frame on symbolic buffers and only RECORDS its helper calls; when
//     the record is a chain the library fuses (PhaseVocoder analyze -> shift(linear fn) -> synthesize, or
//     window -> frft(+1) -> frft(-1) -> real -> window), consecutive frames with the same record become one
//     tsb_chain_render launch over all frames x all lanes.  Per-frame control logic (next_hop_*, t) still runs.
//
// Nothing here computes audio on the CPU: the only host arithmetic is the user's own code.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../../include/tsb200.h"

namespace tsbrt {

[[noreturn]] inline void fatal(const std::string &msg) {
    std::cerr << "FATAL: " << msg << std::endl;
    std::exit(1);
}
inline void check(int rc, const char *what) {
    if (rc != TSB_OK) fatal(std::string(what) + ": " + tsb_last_error());
}

enum OpKind { OP_WINDOW, OP_WIDEN, OP_REAL, OP_FRFT, OP_PV_ANALYZE_WAVE, OP_PV_SHIFT, OP_PV_SYNTH_WAVE, OP_OTHER };
struct Op {
    OpKind kind;
    const void *a = nullptr, *b = nullptr;  // buffers / objects involved
    double p = 0;                           // window kind, frft exponent or shift ratio
    std::vector<double> fn_bins;            // OP_PV_SHIFT: fn(bin centre), size/2 values
    std::string why;                        // OP_OTHER: what could not be recorded
};

struct Runtime {
    tsb_ctx *ctx = nullptr;
    bool tracing = false;      // fused mode: helpers record instead of executing
    bool raw_access = false;   // a traced callback touched samples itself
    std::vector<Op> trace;
    static Runtime &get() {
        static Runtime r;
        return r;
    }
    tsb_ctx *context() {
        if (!ctx) check(tsb_ctx_create(device, nullptr, 1, &ctx), "tsb_ctx_create");
        return ctx;
    }
    int device = 0;
    void sync() { check(tsb_ctx_sync(context()), "tsb_ctx_sync"); }
};

// Device copy of one Buf<T>: which side holds the current data.
struct Mirror {
    void *d = nullptr;
    size_t bytes = 0;
    bool host_valid = true, dev_valid = false;
    void release() {
        if (d && Runtime::get().ctx) tsb_free(Runtime::get().ctx, d);
        d = nullptr;
        bytes = 0;
        dev_valid = false;
    }
    void *to_device(const void *host, size_t n) {
        Runtime &rt = Runtime::get();
        if (bytes < n) {
            release();
            check(tsb_malloc(rt.context(), n ? n : 16, &d), "tsb_malloc");
            bytes = n;
        }
        if (!dev_valid) {
            if (n) check(tsb_upload(rt.context(), d, host, n), "tsb_upload");
            rt.sync();
            dev_valid = true;
        }
        return d;
    }
    void to_host(void *host, size_t n) {
        if (host_valid) return;
        Runtime &rt = Runtime::get();
        if (n) check(tsb_download(rt.context(), host, d, n), "tsb_download");
        rt.sync();
        host_valid = true;
    }
};

}  // namespace tsbrt
