// rtinc/buffer.hpp — Buf<T>: planar [channel][size] buffer, the data type of the operator API.
// Same public surface and semantics as the reference's rtinc/buffer.hpp:7-37 (data/size/num_channels, resize
// keyed on the element COUNT and zero-filling, fill_from with a converting functor, zero, operator[] -> row
// pointer), plus a device copy managed by the runtime (see runtime.hpp).  Touching samples through
// operator[] is host access: the device copy is fetched first and considered stale afterwards.
#pragma once
#include "runtime.hpp"

template <typename T>
struct Buf {
    T *data = nullptr;
    size_t size = 0;
    size_t num_channels = 0;
    tsbrt::Mirror _mirror;
    Buf() {}
    Buf(size_t num_channels, size_t size) { resize(num_channels, size); }
    Buf(const Buf &) = delete;
    Buf &operator=(const Buf &) = delete;
    ~Buf() {
        _mirror.release();
        if (data) std::free(data);
    }
    size_t _bytes() const { return sizeof(T) * size * num_channels; }
    void resize(size_t num_channels, size_t size) {
        if (this->size * this->num_channels != size * num_channels) {  // reference: buffer.hpp:16
            this->size = size;
            this->num_channels = num_channels;
            if (data) std::free(data);
            void *p = nullptr;
            if (posix_memalign(&p, 64, sizeof(T) * num_channels * size + 64)) tsbrt::fatal("out of memory");
            data = (T *)p;
            zero();
        }
    }
    template <typename U>
    void fill_from(Buf<U> &other, std::function<T(U)> mix = [](U x) { return T(x); });
    void zero() {
        for (size_t i = 0; i < size * num_channels; i++) data[i] = {};
        _host_written();
    }
    T *operator[](int index) {
        tsbrt::Runtime &rt = tsbrt::Runtime::get();
        if (rt.tracing) rt.raw_access = true;
        _to_host();
        _host_written();
        return data + index * size;
    }
    // ---- residency (used by the helper implementations) ----
    void _to_host() { _mirror.to_host(data, _bytes()); }
    void _host_written() {
        _mirror.host_valid = true;
        _mirror.dev_valid = false;
    }
    void *_to_device() { return _mirror.to_device(data, _bytes()); }
    void _dev_written() {
        _mirror.dev_valid = true;
        _mirror.host_valid = false;
    }
};
using WaveBuf = Buf<double>;
