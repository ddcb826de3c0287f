"""oracle/orc.py — TEST INFRASTRUCTURE: ctypes binding of the plain-C oracle (oracle/oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")


def build(force=False):
    """Compile oracle.c -> _build/liboracle.so (gcc only; seconds)."""
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "signal_gen.h"))):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class _Result(C.Structure):
    _fields_ = [("frames", C.c_size_t), ("out_len", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P = C.c_void_p
        L.orc_gen_signal.argtypes = [C.c_int, C.c_int, C.c_double, C.c_size_t, P]
        L.orc_fft_forward.argtypes = [P, P, C.c_size_t]
        L.orc_window_hann.argtypes = [P, C.c_size_t, C.c_size_t]
        L.orc_window_sqrt_hann.argtypes = [P, C.c_size_t, C.c_size_t]
        L.orc_frft.argtypes = [P, P, C.c_size_t, C.c_size_t, C.c_double]
        L.orc_set_hop.argtypes = [C.c_double, P, P, P]
        L.orc_set_hop.restype = C.c_int
        L.orc_pv_new.argtypes = [P, P]
        L.orc_pv_new.restype = P
        L.orc_pv_free.argtypes = [P]
        L.orc_pv_set_hop.argtypes = [P, C.c_double]
        L.orc_pv_reset.argtypes = [P, C.c_int, C.c_int]
        for name in ("orc_pv_data", "orc_pv_last_phase", "orc_pv_phase_sum", "orc_pv_fft"):
            getattr(L, name).argtypes = [P]
            getattr(L, name).restype = P
        L.orc_pv_size.argtypes = [P]
        L.orc_pv_size.restype = C.c_size_t
        L.orc_pv_nch.argtypes = [P]
        L.orc_pv_nch.restype = C.c_size_t
        L.orc_pv_analyze.argtypes = [P, P, C.c_size_t, C.c_size_t]
        L.orc_pv_analyze_fft.argtypes = [P, P, C.c_size_t, C.c_size_t]
        L.orc_pv_shift_scale.argtypes = [P, C.c_double]
        L.orc_pv_synthesize.argtypes = [P, P]
        L.orc_render_pv.restype = _Result
        L.orc_render_pv.argtypes = [P, C.c_size_t, C.c_size_t, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                    C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, C.c_size_t,
                                    C.c_size_t, P, P, P, P, P, P, P, C.c_size_t]
        L.orc_render_filter.restype = _Result
        L.orc_render_filter.argtypes = [P, C.c_size_t, C.c_size_t, C.c_uint32, C.c_double, P, C.c_int, C.c_int,
                                        C.c_double, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, C.c_size_t]
        L.orc_render_granular.restype = _Result
        L.orc_render_granular.argtypes = [P, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                          C.c_double, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, P, P, P, P,
                                          C.c_size_t]
        L.orc_hop_clock.argtypes = [C.c_double, C.c_double, C.c_size_t, P, P]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def gen_signal(stream, nch, count, rate=48000.0):
    """Interleaved float32 [count, nch] (SURVEY §8d generator)."""
    out = np.empty((count, nch), np.float32)
    lib().orc_gen_signal(stream, nch, rate, count, _p(out))
    return out


def fft_forward(x):
    x = np.ascontiguousarray(x, np.complex128)
    y = np.empty_like(x)
    lib().orc_fft_forward(_p(x), _p(y), x.shape[-1])
    return y


def window(data, kind):
    d = np.array(data, np.float64, order="C")
    nch, size = d.shape
    (lib().orc_window_hann if kind == "hann" else lib().orc_window_sqrt_hann)(_p(d), nch, size)
    return d


def frft(x, exponent, inplace=True):
    x = np.array(x, np.complex128, order="C")
    nch, n = x.shape
    if inplace:
        lib().orc_frft(_p(x), _p(x), nch, n, exponent)
        return x
    y = np.zeros_like(x)
    lib().orc_frft(_p(x), _p(y), nch, n, exponent)
    return y


def set_hop(h):
    hop = C.c_double()
    hs = C.c_uint64()
    hf = C.c_uint64()
    ok = lib().orc_set_hop(h, C.byref(hop), C.byref(hs), C.byref(hf))
    return ok, hs.value, hf.value


def hop_clock(hop, rate, frames):
    ch = np.empty(frames, np.int64)
    t = np.empty(frames, np.float64)
    lib().orc_hop_clock(hop, rate, frames, _p(ch), _p(t))
    return ch, t


class PV:
    """Step-wise PhaseVocoder oracle with explicit hop (rtinc/pv.hpp:12-63)."""

    def __init__(self, hop, rate=48000.0):
        self._hop = C.c_double(hop)
        self._rate = C.c_double(rate)
        self.h = lib().orc_pv_new(C.addressof(self._hop), C.addressof(self._rate))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_pv_free(self.h)
            self.h = None

    def _view(self, fn, width):
        n = lib().orc_pv_nch(self.h) * lib().orc_pv_size(self.h) * width
        ptr = getattr(lib(), fn)(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,)).copy()

    def analyze(self, wave):
        w = np.array(wave, np.float64, order="C")
        lib().orc_pv_analyze(self.h, _p(w), w.shape[0], w.shape[1])
        return w  # windowed in place, like the reference

    def shift(self, ratio):
        lib().orc_pv_shift_scale(self.h, ratio)

    def shift_map(self, fn_bins, ratio):
        """PhaseVocoder::shift (rtlib/pv.hpp:56-79) with fn given by its values on the bin centres (`fn_bins`,
        size/2 of them, what decides the destination bin, :72) and fn(freq) = ratio*freq on the frequencies (:66) —
        the two things a fused plan is built from.  Pure-Python scatter in the reference's order: ascending j,
        amplitudes add, the frequency of the last writer wins, bins >= size/2 end up zero."""
        nch, size = lib().orc_pv_nch(self.h), lib().orc_pv_size(self.h)
        ptr = lib().orc_pv_data(self.h)
        d = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(nch * size * 2,)).reshape(nch, size, 2)
        rate, half = self._rate.value, size // 2
        tmp = d.copy()
        d[:] = 0.0
        for i in range(nch):
            for j in range(half):
                q = fn_bins[j] / rate * size
                if not (-2147483648.0 < q < 2147483648.0):
                    continue
                k = int(q)  # truncation toward zero, :72
                if 0 <= k < half:
                    d[i, k, 0] += tmp[i, j, 0]
                    d[i, k, 1] = tmp[i, j, 1] * ratio

    def synthesize(self):
        nch, size = lib().orc_pv_nch(self.h), lib().orc_pv_size(self.h)
        out = np.empty((nch, size), np.float64)
        lib().orc_pv_synthesize(self.h, _p(out))
        return out

    @property
    def data(self):
        nch, size = lib().orc_pv_nch(self.h), lib().orc_pv_size(self.h)
        return self._view("orc_pv_data", 2).reshape(nch, size, 2)

    @property
    def last_phase(self):
        nch, size = lib().orc_pv_nch(self.h), lib().orc_pv_size(self.h)
        return self._view("orc_pv_last_phase", 1).reshape(nch, size)

    @property
    def phase_sum(self):
        nch, size = lib().orc_pv_nch(self.h), lib().orc_pv_size(self.h)
        return self._view("orc_pv_phase_sum", 1).reshape(nch, size)


def n_frames(length, n, hop):
    """Frame count for a length-`length` input at constant integer hop (tinyspec.cpp:142-149)."""
    if length < 2048 + n:
        return 0
    return (length - 2048 - n) // hop + 1


def render_pv(x, n, hop_ratio, shift, rate=48000.0, max_frames=0, want_frames=False, state_every=0,
              want_tables=False):
    """x: interleaved float32 [L, nch].  Returns dict(stream=float32 [P, nch], ola64=float64 [P+n, nch],
    frames=F, frame_out=[F, nch, n] (optional), state=[slots, 2, nch, n] (optional), last_phase, phase_sum)."""
    x = np.ascontiguousarray(x, np.float32)
    L, nch = x.shape
    hop = n * hop_ratio
    est_frames = max(1, int((L - 2048 - n) / max(hop, 1e-9)) + 4)
    if max_frames:
        est_frames = min(est_frames, max_frames)
    out_cap = (L + 2 * n + 4096) * nch
    stream = np.zeros(out_cap, np.float32)
    ola64 = np.zeros(out_cap, np.float64)
    frame_out = np.zeros((est_frames, nch, n), np.float64) if want_frames else None
    slots = (est_frames + state_every - 1) // state_every + 1 if state_every else 0
    state = np.zeros((slots, 2, nch, n), np.float64) if state_every else None
    lp = np.zeros((nch, n), np.float64)
    ps = np.zeros((nch, n), np.float64)
    tabs = [np.zeros(est_frames, np.int64) for _ in range(4)] + [np.zeros(est_frames, np.float64)] \
        if want_tables else [None] * 5
    r = lib().orc_render_pv(_p(x), L, nch, n, hop_ratio, shift, rate, max_frames, _p(stream), out_cap, _p(ola64),
                            out_cap, _p(frame_out), frame_out.size if want_frames else 0, _p(state), state_every,
                            state.size if state_every else 0, _p(lp), _p(ps), *[_p(t) for t in tabs],
                            est_frames if want_tables else 0)
    F = r.frames
    res = dict(frames=F, stream=stream[:r.out_len].reshape(-1, nch), ola64=ola64.reshape(-1, nch),
               last_phase=lp, phase_sum=ps)
    if want_frames:
        res["frame_out"] = frame_out[:F]
    if state_every:
        res["state"] = state[:(F + state_every - 1) // state_every]
    if want_tables:
        res["tables"] = dict(in_off=tabs[0][:F], out_off=tabs[1][:F], hop=tabs[2][:F], win=tabs[3][:F], t=tabs[4][:F])
    return res


WIN = {"none": 0, None: 0, "hann": 1, "sqrt_hann": 2}


def render_filter(x, n, hop_ratio, gain=None, win_in="sqrt_hann", win_out="sqrt_hann", rate=48000.0, max_frames=0,
                  want_frames=False):
    x = np.ascontiguousarray(x, np.float32)
    L, nch = x.shape
    hop = n * hop_ratio
    est_frames = max(1, int((L - 2048 - n) / max(hop, 1e-9)) + 4)
    if max_frames:
        est_frames = min(est_frames, max_frames)
    out_cap = (L + 2 * n + 4096) * nch
    stream = np.zeros(out_cap, np.float32)
    ola64 = np.zeros(out_cap, np.float64)
    frame_out = np.zeros((est_frames, nch, n), np.float64) if want_frames else None
    g = None if gain is None else np.ascontiguousarray(gain, np.float64)
    assert g is None or g.shape == (n // 2 + 1,)
    r = lib().orc_render_filter(_p(x), L, nch, n, hop_ratio, _p(g), WIN[win_in], WIN[win_out], rate, max_frames,
                                _p(stream), out_cap, _p(ola64), out_cap, _p(frame_out),
                                frame_out.size if want_frames else 0)
    res = dict(frames=r.frames, stream=stream[:r.out_len].reshape(-1, nch), ola64=ola64.reshape(-1, nch))
    if want_frames:
        res["frame_out"] = frame_out[:r.frames]
    return res


def render_granular(x, n_first=0, n_base=1800, vary=False, stream_idx=0, rate=48000.0, max_frames=0):
    x = np.ascontiguousarray(x, np.float32)
    L, nch = x.shape
    out_cap = (4 * L + 65536) * nch
    cap = max_frames or (L // 16 + 16)
    stream = np.zeros(out_cap, np.float32)
    ola64 = np.zeros(out_cap, np.float64)
    tabs = [np.zeros(cap, np.int64) for _ in range(4)] + [np.zeros(cap, np.float64)]
    r = lib().orc_render_granular(_p(x), L, nch, n_first, n_base, int(vary), stream_idx, rate, max_frames,
                                  _p(stream), out_cap, _p(ola64), out_cap, *[_p(t) for t in tabs], cap)
    F = r.frames
    return dict(frames=F, stream=stream[:r.out_len].reshape(-1, nch), ola64=ola64.reshape(-1, nch),
                tables=dict(in_off=tabs[0][:F], out_off=tabs[1][:F], hop=tabs[2][:F], win=tabs[3][:F], t=tabs[4][:F]))


# ---- the spectral-bytebeat pattern (BASELINE cfg5), numpy restatement ----
def bytebeat_spectrum(frame_idx, ch, n):
    """Bins 1..n-1 of the seeded integer generator of oracle/ref/taps/bytebeat_tap.cpp:17-24 (the hack's per-bin
    loop, hacks/spectral_bytebeat/voidbubbles.cpp:9-15 with an integer-only formula): complex128 [ch, n], bin 0 = 0."""
    i = np.arange(n, dtype=np.uint32)[None, :]
    c = np.arange(ch, dtype=np.uint32)[:, None]
    with np.errstate(over="ignore"):
        h = (np.uint32(frame_idx) * np.uint32(2654435761)) ^ (c * np.uint32(40503) + i * np.uint32(2246822519))
        h ^= h >> np.uint32(15)
        h *= np.uint32(2246822519)
        h ^= h >> np.uint32(13)
    d = (1 + (np.arange(n) >> 4)).astype(np.float64)[None, :]
    re = ((h & np.uint32(0xffff)).astype(np.int64) - 32768) / 32768.0 / d
    im = ((h >> np.uint32(16)).astype(np.int64) - 32768) / 32768.0 / d
    x = re + 1j * im
    x[:, 0] = 0
    return x


def render_bytebeat(n, ch, frames, hop=None, spectrum=None):
    """The pattern of hacks/spectral_bytebeat/*.cpp (voidbubbles.cpp:4-19) frame by frame: a STATIC buffer of 2n points,
    bins i and 2n-i (0 < i < n) rewritten every frame, frft(buf, buf, -1) IN PLACE (rtlib/fft.hpp:3-42: forward DFT
    over sqrt(2n)) — so bins 0 and n hold the previous frame's output samples — out = Re(buf[:n]) * (1 - |1 - i/(n/2)|),
    overlap-added at `hop` in double (the scheduler's own stream is float32, tinyspec.cpp:170-200).
    spectrum(f) -> complex [ch, n] overrides the seeded generator.  Returns dict(out [frames, ch, n], ola64 [ch, L])."""
    hop = n // 4 if hop is None else hop
    buf = np.zeros((ch, 2 * n), np.complex128)
    i = np.arange(n)
    tri = 1.0 - np.abs(1.0 - i / (n / 2.0))
    outs = np.zeros((frames, ch, n))
    ola = np.zeros((ch, (frames - 1) * hop + n if frames else 0))
    for f in range(frames):
        x = bytebeat_spectrum(f, ch, n) if spectrum is None else spectrum(f)
        buf[:, 1:n] = x[:, 1:n]
        buf[:, 2 * n - 1:n:-1] = np.conj(x[:, 1:n])
        buf = np.fft.fft(buf, axis=1) / np.sqrt(2.0 * n)
        outs[f] = buf[:, :n].real * tri
        ola[:, f * hop:f * hop + n] += outs[f]
    return dict(out=outs, ola64=ola)


# ---- the reference's stateless spectral_bytebeat hacks, restated from their sources (numpy int64 arithmetic with C's
# truncating division; pow / sin / fmod through numpy or, where a truncation follows, Python's math = the C library) ----
def _cdiv(a, b):
    """C's truncating integer division on int64 arrays"""
    a, b = np.asarray(a, np.int64), np.asarray(b, np.int64)
    q = np.abs(a) // np.abs(b)
    return np.where((a < 0) == (b < 0), q, -q)


def _cmod(a, b):
    """C's % (sign of the dividend)"""
    return np.asarray(a, np.int64) - _cdiv(a, b) * np.asarray(b, np.int64)


def _wrap32(v):
    v = int(v) & 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def hack_spectrum(hack, t, ch, n):
    """Bins 1..n-1 that hacks/spectral_bytebeat/<hack>.cpp writes for callback time t: complex128 [ch, n], bin 0 = 0.
    voidbubbles.cpp:9-13, acidbattery.cpp:9-17, andshuffle.cpp:9-14, brokenconvo.cpp:9-11, ghostacid.cpp:9-12."""
    import math
    i = np.arange(n, dtype=np.int64)[None, :]
    c = np.arange(ch, dtype=np.int64)[:, None]
    fi = np.arange(n, dtype=np.float64)[None, :]
    powtab = lambda e, off=0: np.array([math.pow(k + off, e) if k + off > 0 else 1.0 for k in range(n)])[None, :]
    with np.errstate(all="ignore"):
        if hack == "voidbubbles":
            r = _wrap32(int(t * 32) * n) + i
            ipow = np.array([int(math.pow(k, 1.1)) for k in range(n)], np.int64)[None, :]
            k = ((c + ipow) & _cdiv(r, 3)) | _cdiv(r, 250) | _cdiv(r, 136) | (r >> 16)
            re = np.sin(2.0 * math.pi / 64. * k.astype(np.float64)) / powtab(0.7, 1) * 32
            x = re + 0j
        elif hack == "acidbattery":
            r = _wrap32(int(t * 32) * n) + i + 0 * c
            k = _cmod((r >> 14) ^ (r >> 17), 64) + 2
            mag = np.sin(powtab(0.1) * k * 16 + _cmod(r & (r >> 13), k).astype(np.float64)
                         / (1 + _cmod(_cdiv(r, 3) & _cdiv(r, 1023), k).astype(np.float64))) / powtab(1.0, 1) * 32.0
            x = mag + 1j * mag
        elif hack == "andshuffle":
            r = int(_cdiv(_wrap32(int(t * 8) * n), 2)) + n - i
            j = _cdiv(r, 128) + c
            m = _cmod(((_cdiv(j, 5) & _cdiv(j, 7) & (_cdiv(j, 64) ^ _cdiv(j, 127))) ^ (j >> 16)), 128)
            ii = ((i * i + (1 << 31)) % (1 << 32)) - (1 << 31)   # i*i in C's 32-bit int
            mag = np.sin(1 + np.fmod(ii.astype(np.float64), 1 + m.astype(np.float64))) / powtab(0.75) * 32
            x = mag + 1j * mag
        elif hack == "brokenconvo":
            r = int(_cdiv(_wrap32(int(t * math.pi) * n), 2)) + i
            m = _cdiv(r, 77) & _cdiv(r, 38) & (_cdiv(r, 24) + c)
            re = np.sin(2.0 * math.pi * (fi / (1 + np.fmod(1 + m.astype(np.float64), fi + 1)))) / powtab(0.75, 1) * 64
            x = re + 0j
        elif hack == "ghostacid":
            r = _wrap32(int(t * 4) * n) + i
            den = 1 + (_cdiv(r, 97) ^ (r & _cdiv(r, 2))).astype(np.float64)
            mag = np.sin(math.pi * np.fmod((_cdiv(i, 24) + c).astype(np.float64), 1 + r / den / 7)) / powtab(1.0, 1) * 128.0
            x = mag + 1j * mag
        else:
            raise ValueError(hack)
    x = np.array(np.broadcast_to(x, (ch, n)), np.complex128)
    x[:, 0] = 0
    return x


HACK_FRAMES = {"voidbubbles": 1 << 12, "acidbattery": 1 << 11, "andshuffle": 1 << 17, "brokenconvo": 1 << 15,
               "ghostacid": 1 << 14}


def render_hack(hack, frames, ch=2, rate=48000.0):
    """render_bytebeat with the spectrum of an unmodified reference hack, at the frame size and hop (n/2) it sets;
    callback time of frame f = f*hop/rate (tinyspec.cpp:110-112,163-168 with an integral hop)."""
    n = HACK_FRAMES[hack]
    hop = n // 2
    return render_bytebeat(n, ch, frames, hop=hop, spectrum=lambda f: hack_spectrum(hack, f * hop / rate, ch, n))
