// rtinc/fft.hpp — cplx, FFTBuf and frft (reference: rtinc/fft.hpp:1-4, rtlib/fft.hpp:3-42), on the device:
// any size, any exponent, in == out reproduces the reference's in-place aliasing (tsb_frft).
#pragma once
#include <cassert>
#include <complex>

#include "buffer.hpp"
using cplx = std::complex<double>;
using FFTBuf = Buf<cplx>;

inline void frft(FFTBuf &in, FFTBuf &out, double exponent) {
    assert(in.size == out.size);
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_FRFT;
        op.a = &in;
        op.b = &out;
        op.p = exponent;
        rt.trace.push_back(op);
        return;
    }
    if (in.num_channels * in.size == 0) return;  // rtlib/fft.hpp:8-9
    double *din = (double *)in._to_device();
    double *dout = &in == &out ? din : (double *)out._to_device();
    tsbrt::check(tsb_frft(rt.context(), din, dout, (int64_t)in.num_channels, (int64_t)in.size, exponent), "tsb_frft");
    out._dev_written();
}

namespace tsbrt {
template <typename T, typename U>
inline OpKind probe_mix(std::function<T(U)> &) { return OP_OTHER; }
inline OpKind probe_mix(std::function<cplx(double)> &f) {
    return (f(2.5) == cplx(2.5, 0) && f(-1.0) == cplx(-1.0, 0) && f(0.0) == cplx(0, 0)) ? OP_WIDEN : OP_OTHER;
}
inline OpKind probe_mix(std::function<double(cplx)> &f) {
    return (f(cplx(2.5, -7)) == 2.5 && f(cplx(-1, 3)) == -1.0 && f(cplx(0, 1)) == 0.0) ? OP_REAL : OP_OTHER;
}
}  // namespace tsbrt

// fill_from (reference: rtinc/buffer.hpp:24-30).  The functor is arbitrary user code, so the conversion runs on
// the host; while a callback is being traced, the two conversions the reference's own helpers use (widen a
// real buffer to complex, take the real part) are recognised by probing the functor.
template <typename T>
template <typename U>
void Buf<T>::fill_from(Buf<U> &other, std::function<T(U)> mix) {
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    resize(other.num_channels, other.size);
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_OTHER;
        op.why = "fill_from with a functor that is neither a widening nor a real-part conversion";
        op.a = &other;
        op.b = this;
        op.kind = tsbrt::probe_mix(mix);
        rt.trace.push_back(op);
        return;
    }
    other._to_host();
    for (size_t i = 0; i < num_channels; i++)
        for (size_t j = 0; j < size; j++) data[i * size + j] = mix(other.data[i * other.size + j]);
    _host_written();
}
