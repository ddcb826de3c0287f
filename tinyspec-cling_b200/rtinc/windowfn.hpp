// rtinc/windowfn.hpp — window_hann / window_sqrt_hann (reference: rtlib/windowfn.hpp:1-14), in place, on the device.
#pragma once
#include "buffer.hpp"

inline void _tsb_window(WaveBuf &data, int kind) {
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_WINDOW;
        op.a = &data;
        op.p = kind;
        rt.trace.push_back(op);
        return;
    }
    if (data.size * data.num_channels == 0) return;
    tsbrt::check(tsb_window(rt.context(), (double *)data._to_device(), (int64_t)data.num_channels, (int64_t)data.size, kind),
                 "tsb_window");
    data._dev_written();
}
inline void window_hann(WaveBuf &data) { _tsb_window(data, TSB_WIN_HANN); }
inline void window_sqrt_hann(WaveBuf &data) { _tsb_window(data, TSB_WIN_SQRT_HANN); }
