// rtinc/pv.hpp — PVBin, PVBuf and PhaseVocoder with the reference's public surface (rtinc/pv.hpp:1-63) and
// semantics (rtlib/pv.hpp:4-79), executing on the device.
#pragma once
#include "fft.hpp"
#include "windowfn.hpp"

struct PVBin {
    double amp, freq;
    PVBin() : amp(0), freq(0) {}
    PVBin(double amp, double freq) : amp(amp), freq(freq) {}
};
using PVBuf = Buf<PVBin>;

namespace internals {
extern double hop;
extern double rate;
}

struct PhaseVocoder {
    PVBuf data;
    Buf<double> last_phase;
    Buf<double> phase_sum;
    size_t size = 0;
    size_t num_channels = 0;
    FFTBuf fft;
    bool global_hop;
    double user_hop;
    PhaseVocoder() { use_global_hop(); }
    PhaseVocoder(size_t num_channels, size_t size) : PhaseVocoder() { resize(num_channels, size); }
    void resize(size_t num_channels, size_t size) {
        if (this->size * this->num_channels != size * num_channels) {  // reference: rtinc/pv.hpp:24
            this->size = size;
            this->num_channels = num_channels;
            data.resize(num_channels, size);
            last_phase.resize(num_channels, size);
            phase_sum.resize(num_channels, size);
            _state_epoch++;
            _reset_last = _reset_sum = true;
        }
    }
    void reset(bool last = true, bool sum = true) {
        if (last && last_phase.data) last_phase.zero();
        if (sum && phase_sum.data) phase_sum.zero();
        if (last || sum) _state_epoch++;
        _reset_last |= last;
        _reset_sum |= sum;
    }
    void fill_from(PhaseVocoder &other) {
        num_channels = other.num_channels;
        size = other.size;
        global_hop = other.global_hop;
        user_hop = other.user_hop;
        data.fill_from(other.data);
        last_phase.fill_from(other.last_phase);
        phase_sum.fill_from(other.phase_sum);
    }
    void use_global_hop() { global_hop = true; }
    void set_hop(double new_hop) {
        user_hop = new_hop;
        global_hop = false;
    }
    double hop() {
        if (global_hop) return internals::hop;
        else return user_hop;
    }
    PVBin *operator[](int index) { return data[index]; }
    void analyze(WaveBuf &other);
    void synthesize(WaveBuf &other);
    void analyze(FFTBuf &other);
    void synthesize(FFTBuf &other);
    double phase_to_frequency(size_t bin, double phase_diff);
    double frequency_to_phase(double freq);
    void shift(std::function<double(double)> fn);
    void shift(std::function<void(WaveBuf &)> fn);
    // runtime bookkeeping (fused mode keeps this object's state in chain layout on the device)
    int _state_epoch = 0;
    bool _reset_last = false, _reset_sum = false;
};

inline void PhaseVocoder::analyze(WaveBuf &other) {  // reference: rtlib/pv.hpp:4-9
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    if (rt.tracing) {
        resize(other.num_channels, other.size);
        tsbrt::Op op;
        op.kind = tsbrt::OP_PV_ANALYZE_WAVE;
        op.a = this;
        op.b = &other;
        op.p = hop();
        rt.trace.push_back(op);
        return;
    }
    window_sqrt_hann(other);
    fft.fill_from(other);
    frft(fft, fft, 1);
    analyze(fft);
}
inline void PhaseVocoder::synthesize(WaveBuf &other) {  // reference: rtlib/pv.hpp:10-16
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    other.resize(num_channels, size);
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_PV_SYNTH_WAVE;
        op.a = this;
        op.b = &other;
        op.p = hop();
        rt.trace.push_back(op);
        return;
    }
    synthesize(fft);
    frft(fft, fft, -1);
    other.fill_from<cplx>(fft, [](cplx x) { return x.real(); });
    window_sqrt_hann(other);
}
inline void PhaseVocoder::analyze(FFTBuf &other) {  // reference: rtlib/pv.hpp:17-28
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    resize(other.num_channels, other.size);
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_OTHER;
        op.why = "PhaseVocoder::analyze(FFTBuf&) outside analyze(WaveBuf&)";
        rt.trace.push_back(op);
        return;
    }
    if (num_channels * size == 0) return;
    tsbrt::check(tsb_pv_analyze(rt.context(), (const double *)other._to_device(), (double *)last_phase._to_device(),
                                (double *)data._to_device(), (int64_t)num_channels, (int64_t)size, hop(), internals::rate),
                 "tsb_pv_analyze");
    last_phase._dev_written();
    data._dev_written();
}
inline void PhaseVocoder::synthesize(FFTBuf &other) {  // reference: rtlib/pv.hpp:29-38
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    other.resize(num_channels, size);
    if (rt.tracing) {
        tsbrt::Op op;
        op.kind = tsbrt::OP_OTHER;
        op.why = "PhaseVocoder::synthesize(FFTBuf&) outside synthesize(WaveBuf&)";
        rt.trace.push_back(op);
        return;
    }
    if (num_channels * size == 0) return;
    tsbrt::check(tsb_pv_synthesize(rt.context(), (double *)other._to_device(), (double *)phase_sum._to_device(),
                                   (const double *)data._to_device(), (int64_t)num_channels, (int64_t)size, hop(),
                                   internals::rate),
                 "tsb_pv_synthesize");
    phase_sum._dev_written();
    other._dev_written();
}
inline double PhaseVocoder::phase_to_frequency(size_t bin, double phase_diff) {  // reference: rtlib/pv.hpp:39-45
    double delta = phase_diff - bin * 2.0 * M_PI * hop() / size;
    int qpd = (int)(delta / M_PI);
    if (qpd >= 0) qpd += qpd & 1;
    else qpd -= qpd & 1;
    return (bin + size / hop() * (delta - M_PI * qpd) / 2.0 / M_PI) * internals::rate / size;
}
inline double PhaseVocoder::frequency_to_phase(double freq) {  // reference: rtlib/pv.hpp:46-48
    return 2.0 * M_PI * freq / internals::rate * hop();
}
inline void PhaseVocoder::shift(std::function<double(double)> fn) {  // reference: rtlib/pv.hpp:49-55
    shift([&](WaveBuf &buf) {
        for (size_t i = 0; i < buf.num_channels; i++)
            for (size_t j = 0; j < buf.size; j++) buf.data[i * buf.size + j] = fn(buf.data[i * buf.size + j]);
    });
}
inline void PhaseVocoder::shift(std::function<void(WaveBuf &)> fn) {  // reference: rtlib/pv.hpp:56-79
    tsbrt::Runtime &rt = tsbrt::Runtime::get();
    const size_t half = size / 2;
    bool was_tracing = rt.tracing, was_raw = rt.raw_access;
    rt.tracing = false;  // the scratch buffers below are ordinary host buffers, not the callback's
    WaveBuf bins(num_channels, half);
    for (size_t i = 0; i < num_channels; i++)
        for (size_t j = 0; j < half; j++) bins.data[i * half + j] = j / (double)size * internals::rate;
    if (was_tracing) {
        // fn acts on values, not on signal data: evaluate it on the bin centres (as the reference does) and
        // probe it for linearity, which is what the fused kernel can apply to the per-bin frequencies.
        fn(bins);
        WaveBuf probe(num_channels, half);
        const double pts[4] = {1.0, 1000.0, -3.7, 12345.678};
        for (size_t e = 0; e < num_channels * half; e++) probe.data[e] = pts[e % 4];
        fn(probe);
        tsbrt::Op op;
        op.kind = tsbrt::OP_PV_SHIFT;
        op.a = this;
        double ratio = (num_channels * half != 0) ? probe.data[0] : 1.0;
        bool linear = true, same = true;
        for (size_t e = 0; e < num_channels * half; e++) linear = linear && probe.data[e] == ratio * pts[e % 4];
        for (size_t i = 1; i < num_channels; i++)
            for (size_t j = 0; j < half; j++) same = same && bins.data[i * half + j] == bins.data[j];
        if (!linear || !same) {
            op.kind = tsbrt::OP_OTHER;
            op.why = !linear ? "shift function is not of the form x -> r*x" : "shift function differs between channels";
        }
        op.p = ratio;
        op.fn_bins.assign(bins.data, bins.data + half);
        rt.tracing = true;
        rt.raw_access = was_raw;
        rt.trace.push_back(op);
        return;
    }
    WaveBuf freqs(num_channels, half);
    if (num_channels * half != 0) {
        tsbrt::check(tsb_pv_freqs(rt.context(), (const double *)data._to_device(), (int64_t)num_channels, (int64_t)size,
                                  freqs.data),
                     "tsb_pv_freqs");
    }
    fn(bins);
    fn(freqs);
    bins._to_host();
    freqs._to_host();
    if (num_channels * size != 0)
        tsbrt::check(tsb_pv_shift(rt.context(), (double *)data._to_device(), (int64_t)num_channels, (int64_t)size,
                                  internals::rate, bins.data, freqs.data),
                     "tsb_pv_shift");
    data._dev_written();
}
