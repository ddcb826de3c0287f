"""ctypes binding of libtsb200 (include/tsb200.h).  Device buffers are torch CUDA tensors: torch is
used for device memory and streams only."""
import ctypes as C
import os

import numpy as np

F32, F64, F64_OUT32 = 0, 1, 2
PROG_PV_SHIFT, PROG_GAIN = 0, 1
WIN_NONE, WIN_HANN, WIN_SQRT_HANN = 0, 1, 2
_WIN = {None: 0, "none": 0, "hann": 1, "sqrt_hann": 2, 0: 0, 1: 1, 2: 2}
_PREC = {"f32": 0, "f64": 1, "f64o32": 2, 0: 0, 1: 1, 2: 2}
_PROG = {"pv": 0, "pv_shift": 0, "gain": 1, 0: 0, 1: 1}

_HERE = os.path.dirname(os.path.abspath(__file__))


class TsbError(RuntimeError):
    pass


def lib_path():
    # TSB200_LIB: load another build of the library (kernel tuning experiments)
    return os.environ.get("TSB200_LIB") or os.path.join(_HERE, "libtsb200.so")


class _PlanDesc(C.Structure):
    _fields_ = [("n", C.c_uint32), ("precision", C.c_int32), ("program", C.c_int32), ("win_in", C.c_int32),
                ("win_out", C.c_int32), ("hop", C.c_double), ("rate", C.c_double), ("ratio", C.c_double),
                ("fn_bins", C.c_void_p), ("gain", C.c_void_p)]


class _SpectralArgs(C.Structure):
    _fields_ = [("frame", C.c_uint32), ("channels", C.c_int32), ("n_frames", C.c_int64), ("first_frame", C.c_int64),
                ("hop", C.c_int32), ("generator", C.c_int32), ("spectrum", C.c_void_p), ("out", C.c_void_p),
                ("out_stride", C.c_int64), ("out_len", C.c_int64), ("out_off0", C.c_int64), ("state_edge", C.c_void_p),
                ("load_state", C.c_int32), ("expr", C.c_void_p)]


TSB_EXPR_MAX_CONST = 24


class _Expr(C.Structure):  # tsb_expr (include/tsb200.h)
    _fields_ = [("n_ops", C.c_int32), ("ops", C.c_void_p), ("n_iconst", C.c_int32), ("n_dconst", C.c_int32),
                ("iconst", C.c_void_p), ("dconst", C.c_void_p), ("n_itab", C.c_int32), ("n_dtab", C.c_int32),
                ("itab", C.c_void_p), ("dtab", C.c_void_p), ("n_iparam", C.c_int32), ("n_dparam", C.c_int32),
                ("iparam", C.c_void_p), ("dparam", C.c_void_p)]


class _GranularArgs(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("in_stride", C.c_int64), ("in_len", C.c_int64), ("out", C.c_void_p),
                ("out_stride", C.c_int64), ("out_len", C.c_int64), ("streams", C.c_int32), ("channels", C.c_int32),
                ("frames_max", C.c_int64), ("n_frames", C.c_void_p), ("win", C.c_void_p), ("hop", C.c_void_p),
                ("osize", C.c_void_p), ("in_off", C.c_void_p), ("out_off", C.c_void_p), ("shift", C.c_void_p)]


class _ChainArgs(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("in_stride", C.c_int64), ("in_len", C.c_int64), ("out", C.c_void_p),
                ("out_stride", C.c_int64), ("out_len", C.c_int64), ("n_frames", C.c_int64), ("in_off0", C.c_int64),
                ("out_off0", C.c_int64), ("hop", C.c_int32), ("hop_tab", C.c_void_p),
                ("state_last_phase", C.c_void_p), ("state_phase_sum", C.c_void_p), ("state_ola", C.c_void_p),
                ("load_state", C.c_int32), ("lanes", C.c_int32), ("prev_hop", C.c_int32)]


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise TsbError(f"{path} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')} -j` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    P = C.c_void_p
    L.tsb_last_error.restype = C.c_char_p
    L.tsb_ctx_create.argtypes = [C.c_int, P, C.c_int, C.POINTER(P)]
    L.tsb_ctx_destroy.argtypes = [P]
    L.tsb_ctx_sync.argtypes = [P]
    L.tsb_ctx_trim.argtypes = [P]
    L.tsb_launch_count.argtypes = [P]
    L.tsb_launch_count.restype = C.c_int64
    L.tsb_probe_fma_peak.argtypes = [P, C.c_int32, C.POINTER(C.c_double)]
    L.tsb_hop_fixed.argtypes = [C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tsb_schedule_const.argtypes = [C.c_int64, C.c_uint32, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                     C.POINTER(C.c_int64), P, P, P, P]
    L.tsb_shift_map.argtypes = [C.c_uint32, C.c_double, P, P]
    L.tsb_plan_create.argtypes = [P, C.POINTER(_PlanDesc), C.POINTER(P)]
    L.tsb_plan_destroy.argtypes = [P]
    L.tsb_chain_render.argtypes = [P, P, C.POINTER(_ChainArgs)]
    L.tsb_chain_render_host.argtypes = [P, P, C.POINTER(_ChainArgs)]
    L.tsb_spectral_render.argtypes = [P, C.POINTER(_SpectralArgs)]
    L.tsb_granular_render.argtypes = [P, C.POINTER(_GranularArgs)]
    L.tsb_expr_check.argtypes = [C.POINTER(_Expr)]
    I64, D = C.c_int64, C.c_double
    L.tsb_window.argtypes = [P, P, I64, I64, C.c_int32]
    L.tsb_fft_forward.argtypes = [P, P, P, I64, I64]
    L.tsb_frft.argtypes = [P, P, P, I64, I64, D]
    L.tsb_pv_analyze.argtypes = [P, P, P, P, I64, I64, D, D]
    L.tsb_pv_synthesize.argtypes = [P, P, P, P, I64, I64, D, D]
    L.tsb_pv_freqs.argtypes = [P, P, I64, I64, P]
    L.tsb_pv_shift.argtypes = [P, P, I64, I64, D, P, P]
    I32 = C.c_int32
    L.tsb_timepar_create.argtypes = [P, P, C.POINTER(_ChainArgs), P, I32, I32, I32, I32, C.POINTER(P)]
    L.tsb_timepar_segments.argtypes = [P]
    L.tsb_timepar_prepass.argtypes = [P, P, P]
    L.tsb_timepar_scan.argtypes = [P, P, P]
    L.tsb_timepar_scan_lanes.argtypes = [P, I32, I32, P, P]
    L.tsb_timepar_scan_at.argtypes = [P, I32, I32, P, P, C.c_int64]
    L.tsb_timepar_scan_chain.argtypes = [P, P, P, P, P, I32, C.c_int64]
    L.tsb_ipc_export.argtypes = [P, P, P]
    L.tsb_ipc_open.argtypes = [P, P, C.POINTER(P)]
    L.tsb_ipc_close.argtypes = [P, P]
    L.tsb_malloc.argtypes = [P, C.c_size_t, C.POINTER(P)]
    L.tsb_free.argtypes = [P, P]
    L.tsb_memset.argtypes = [P, P, C.c_int, C.c_size_t]
    L.tsb_timepar_render.argtypes = [P]
    L.tsb_timepar_destroy.argtypes = [P]
    L.tsb_chain_render_timepar.argtypes = [P, P, C.POINTER(_ChainArgs), P, I32]
    L.tsb_add_tail.argtypes = [P, P, P, I64, I64, I64, P, I64]
    L.tsb_stream_create.argtypes = [P, P, I32, D, D, C.POINTER(P)]
    L.tsb_stream_process.argtypes = [P, P, I64, I64, P, I64, I64, C.POINTER(I64)]
    L.tsb_stream_frames.argtypes = [P]
    L.tsb_stream_frames.restype = I64
    L.tsb_stream_destroy.argtypes = [P]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise TsbError(f"libtsb200 error {rc}: {_load().tsb_last_error().decode()}")


def expr_check(ops, n_iconst=0, n_dconst=0, n_itab=0, n_dtab=0, n_iparam=0, n_dparam=0):
    """tsb_expr_check on a raw program (host-side): raises TsbError with the reason when it is malformed"""
    ops = np.ascontiguousarray(ops, np.int32)
    dummy = np.zeros(max(1, TSB_EXPR_MAX_CONST), np.int64)
    x = _Expr()
    x.n_ops, x.ops = len(ops), ops.ctypes.data
    x.n_iconst, x.n_dconst, x.n_itab, x.n_dtab, x.n_iparam, x.n_dparam = n_iconst, n_dconst, n_itab, n_dtab, n_iparam, n_dparam
    x.iconst = x.dconst = x.itab = x.dtab = x.iparam = x.dparam = dummy.ctypes.data
    _check(_load().tsb_expr_check(C.byref(x)))


def abi_version():
    return _load().tsb_abi_version()


def hop_fixed(hop):
    """set_hop() decomposition (rtlib/control.hpp:22-29) -> (hop_samples, hop_fract)."""
    s, f = C.c_uint64(), C.c_uint64()
    _check(_load().tsb_hop_fixed(float(hop), C.byref(s), C.byref(f)))
    return s.value, f.value


def schedule_const(in_len, n, hop, rate=48000.0, max_frames=-1):
    """Frame tables of a constant-(n, hop) render (tinyspec.cpp:135-168)."""
    L = _load()
    nf = C.c_int64()
    _check(L.tsb_schedule_const(in_len, n, float(hop), rate, max_frames, 0, C.byref(nf), None, None, None, None))
    F = nf.value
    tabs = dict(in_off=np.zeros(F, np.int64), out_off=np.zeros(F, np.int64), hop=np.zeros(F, np.int32),
                t=np.zeros(F, np.float64))
    if F:
        _check(L.tsb_schedule_const(in_len, n, float(hop), rate, max_frames, F, C.byref(nf),
                                    tabs["in_off"].ctypes.data, tabs["out_off"].ctypes.data,
                                    tabs["hop"].ctypes.data, tabs["t"].ctypes.data))
    tabs["n_frames"] = F
    return tabs


def granular_schedule(in_len, callback, rate=48000.0, n_first=4096, max_frames=None):
    """The reference's frame scheduler and clock (tinyspec.cpp:135-168) run with a hack's CONTROL logic only, for hacks
    whose frame size / hop / output length change every frame (BASELINE cfg4): callback(f, t, W) -> (osize, shift,
    n_next, hop_next) is what the hack decides in frame f at time t with frame size W — the size of `out` after the
    callback, the resampling ratio, and the arguments of its next_hop_* call (the next frame size and the new hop).
    Returns the tables tsb_granular_render takes: win, in_off, out_off, hop (computed_hop after the callback), osize,
    shift, t."""
    rows = dict(win=[], in_off=[], out_off=[], hop=[], osize=[], shift=[], t=[])
    new_window, computed_hop = int(n_first), 2048          # tinyspec.cpp:38-41
    hop_s, hop_f = hop_fixed(2048.0)
    in_pos = out_pos = ts = tf = 0
    f = 0
    while max_frames is None or f < max_frames:
        W = new_window                                      # latched for this frame, :135-139
        if in_len - in_pos < computed_hop + W:              # starved, :142-143
            break
        in_pos += computed_hop                              # :147-149
        t = (float(ts) + float(tf) / 18446744073709551616.0) / rate   # current_time_secs, :110-112
        osize, shift, n_next, hop_next = callback(f, t, W)
        new_window = int(n_next)
        if hop_next >= 0:                                   # set_hop ignores a negative hop, rtlib/control.hpp:23
            hop_s, hop_f = hop_fixed(float(hop_next))
        computed_hop = hop_s + (1 if tf + hop_f >= 1 << 64 else 0)   # :163-168
        ts += computed_hop
        tf = (tf + hop_f) & ((1 << 64) - 1)
        for k, v in (("win", W), ("in_off", in_pos), ("out_off", out_pos), ("hop", computed_hop), ("osize", int(osize)),
                     ("shift", float(shift)), ("t", t)):
            rows[k].append(v)
        out_pos += computed_hop
        f += 1
    out = {k: np.asarray(v, np.float64 if k in ("shift", "t") else np.int64) for k, v in rows.items()}
    out["n_frames"] = f
    return out


def shift_map(n, rate, fn_bins):
    fb = np.ascontiguousarray(fn_bins, np.float64)
    assert fb.shape == (n // 2,)
    dest = np.zeros(n // 2, np.int32)
    _check(_load().tsb_shift_map(n, rate, fb.ctypes.data, dest.ctypes.data))
    return dest


def bind_host_to_gpu(device=0):
    """Restrict the calling process to the CPUs next to GPU `device` (NVML's CPU affinity of the device) so that
    pinned host buffers allocated afterwards sit on that NUMA node: the host-buffer render moves its data at PCIe
    rate, and with several ranks on one box a remote node's memory and the socket interconnect become the limit.
    Returns the CPU set used, or None when NVML / the platform gives nothing (nothing is changed then)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device)
        bus_id = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        use = cpus & os.sched_getaffinity(0)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None


class Context:
    """One device + one CUDA stream: stream="torch" (default) runs on torch's current stream of that device,
    stream="new" on a stream the context creates and owns, an int is taken as a cudaStream_t."""

    def __init__(self, device=0, stream="torch"):
        L = _load()
        sp, new = None, 0
        if stream == "torch":
            import torch
            sp = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else None
        elif stream == "new":
            new = 1
        elif stream is not None:
            sp = int(stream)
        h = C.c_void_p()
        _check(L.tsb_ctx_create(device, sp, new, C.byref(h)))
        self.h = h
        self.device = device

    def sync(self):
        _check(_load().tsb_ctx_sync(self.h))

    def trim(self):
        """release the device scratch the context keeps between calls (host-render staging, time-parallel state)"""
        _check(_load().tsb_ctx_trim(self.h))

    @property
    def launches(self):
        return _load().tsb_launch_count(self.h)

    def spectral_render(self, frame, channels, n_frames, hop, out=None, spectrum=None, first_frame=0, out_off0=0,
                        state_edge=None, load_state=False, expr=None, times=None):
        """tsb_spectral_render: the spectral-bytebeat pattern (BASELINE cfg5) fused on the device.  spectrum=None: the
        seeded integer generator; else CUDA complex128 [n_frames, channels, frame]; expr = an expr.Program (a hack's
        per-bin formula, TSB_GEN_EXPR) evaluated in the kernel, with `times` = the callback time of every frame
        (default: (first_frame + f)*hop/48000, the reference's clock for an integral hop, tinyspec.cpp:110-112).
        Frames are ADDED into `out` (CUDA float64 [channels, >= out_off0 + (n_frames-1)*hop + frame], zeros by default);
        returns it."""
        import torch
        if out is None:
            out = torch.zeros((channels, out_off0 + max(n_frames - 1, 0) * hop + frame), dtype=torch.float64,
                              device=torch.device("cuda", self.device))
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.shape[0] == channels
        a = _SpectralArgs()
        a.frame, a.channels, a.n_frames, a.first_frame, a.hop = frame, channels, n_frames, first_frame, int(hop)
        a.generator = 2 if expr is not None else (0 if spectrum is None else 1)
        keep = None
        if expr is not None:
            assert spectrum is None, "give a spectrum buffer or an expression, not both"
            if times is None:
                times = [(first_frame + f) * int(hop) / 48000.0 for f in range(n_frames)]
            assert len(times) == n_frames
            itab, dtab = expr.tables(frame)
            ipar, dpar = expr.params(times, frame)
            ops, ic, dc = (np.ascontiguousarray(expr.ops, np.int32), np.ascontiguousarray(expr.iconst, np.int64),
                           np.ascontiguousarray(expr.dconst, np.float64))
            x = _Expr()
            x.n_ops, x.ops = len(ops), ops.ctypes.data
            x.n_iconst, x.n_dconst, x.iconst, x.dconst = len(ic), len(dc), ic.ctypes.data, dc.ctypes.data
            x.n_itab, x.n_dtab, x.itab, x.dtab = itab.shape[0], dtab.shape[0], itab.ctypes.data, dtab.ctypes.data
            x.n_iparam, x.n_dparam, x.iparam, x.dparam = ipar.shape[1], dpar.shape[1], ipar.ctypes.data, dpar.ctypes.data
            keep = (x, ops, ic, dc, itab, dtab, ipar, dpar)
            a.expr = C.addressof(x)
        if spectrum is not None:
            assert spectrum.is_cuda and spectrum.dtype == torch.complex128 and spectrum.is_contiguous()
            assert tuple(spectrum.shape) == (n_frames, channels, frame)
            a.spectrum = spectrum.data_ptr()
        a.out, a.out_stride, a.out_len, a.out_off0 = out.data_ptr(), out.stride(0) if channels > 1 else out.shape[1], out.shape[1], out_off0
        if state_edge is not None:
            assert state_edge.is_cuda and state_edge.dtype == torch.float64 and tuple(state_edge.shape) == (channels, 2)
            a.state_edge = state_edge.data_ptr()
        a.load_state = int(load_state)
        _check(_load().tsb_spectral_render(self.h, C.byref(a)))
        del keep  # (the expression's host arrays are copied during the call)
        self._keep_spec = (out, spectrum, state_edge)
        return out

    def granular_render(self, x, channels, schedules, out=None):
        """tsb_granular_render: x CUDA float32 [streams*channels, L]; schedules = one granular_schedule() result per
        stream.  Returns the float32 output streams [streams*channels, max total hop] (the reference's stream per lane;
        a shorter stream is zero past its end)."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
        streams = len(schedules)
        assert x.shape[0] == streams * channels
        fm = max(1, max(int(s["n_frames"]) for s in schedules))
        tabs = {k: np.zeros((streams, fm), np.float64 if k == "shift" else (np.int64 if k in ("in_off", "out_off") else np.int32))
                for k in ("win", "hop", "osize", "in_off", "out_off", "shift")}
        nf = np.zeros(streams, np.int64)
        total = 0
        for i, s in enumerate(schedules):
            F = int(s["n_frames"])
            nf[i] = F
            for k in tabs:
                tabs[k][i, :F] = s[k][:F]
            if F:
                total = max(total, int(s["out_off"][F - 1] + s["hop"][F - 1]))
        if out is None:
            out = torch.zeros((x.shape[0], max(total, 1)), dtype=torch.float32, device=x.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == x.shape[0]
        a = _GranularArgs()
        a.in_, a.in_stride, a.in_len = x.data_ptr(), x.stride(0) if x.shape[0] > 1 else x.shape[1], x.shape[1]
        a.out, a.out_stride, a.out_len = out.data_ptr(), out.stride(0) if out.shape[0] > 1 else out.shape[1], out.shape[1]
        a.streams, a.channels, a.frames_max = streams, channels, fm
        a.n_frames = nf.ctypes.data
        for k in tabs:
            setattr(a, k, tabs[k].ctypes.data)
        _check(_load().tsb_granular_render(self.h, C.byref(a)))
        return out

    # ---- raw device memory and peer mapping (multi-GPU chains, shard.GpuBackend.link) ----
    def malloc(self, nbytes):
        p = C.c_void_p()
        _check(_load().tsb_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        _check(_load().tsb_free(self.h, ptr))

    def memset(self, ptr, value, nbytes):
        _check(_load().tsb_memset(self.h, ptr, value, nbytes))

    def ipc_export(self, ptr):
        buf = C.create_string_buffer(64)
        _check(_load().tsb_ipc_export(self.h, ptr, buf))
        return bytes(buf.raw)

    def ipc_open(self, handle):
        p = C.c_void_p()
        _check(_load().tsb_ipc_open(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        _check(_load().tsb_ipc_close(self.h, ptr))

    def fma_peak(self, precision="f64"):
        """measured non-tensor FMA rate of the device in TFLOP/s (tsb_probe_fma_peak)"""
        v = C.c_double()
        _check(_load().tsb_probe_fma_peak(self.h, 1 if precision == "f64" else 0, C.byref(v)))
        return v.value

    # ---- the helpers one by one (tsb200.h "immediate mode"); tensors are CUDA float64 / complex128 ----
    def window(self, data, kind):
        """window_hann / window_sqrt_hann in place on a float64 [nch, size] tensor."""
        assert data.is_cuda and data.is_contiguous() and data.dim() == 2
        _check(_load().tsb_window(self.h, data.data_ptr(), data.shape[0], data.shape[1], _WIN[kind]))
        return data

    def fft_forward(self, x, out=None):
        """rows of unnormalised forward DFTs (the FFTW contract), complex128 [rows, n], any n."""
        import torch
        assert x.is_cuda and x.dtype == torch.complex128 and x.is_contiguous() and x.dim() == 2
        out = torch.empty_like(x) if out is None else out
        _check(_load().tsb_fft_forward(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1]))
        return out

    def frft(self, x, exponent, out=None):
        """frft(in, out, exponent); out=None or out is x -> in place, like every reference call site."""
        import torch
        assert x.is_cuda and x.dtype == torch.complex128 and x.is_contiguous() and x.dim() == 2
        out = x if out is None else out
        _check(_load().tsb_frft(self.h, x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], float(exponent)))
        return out

    def pv_analyze(self, spec, last_phase, pvdata, hop, rate=48000.0):
        nch, size = spec.shape
        _check(_load().tsb_pv_analyze(self.h, spec.data_ptr(), last_phase.data_ptr(), pvdata.data_ptr(), nch, size,
                                      float(hop), float(rate)))

    def pv_synthesize(self, spec, phase_sum, pvdata, hop, rate=48000.0):
        nch, size = spec.shape
        _check(_load().tsb_pv_synthesize(self.h, spec.data_ptr(), phase_sum.data_ptr(), pvdata.data_ptr(), nch, size,
                                         float(hop), float(rate)))

    def pv_shift(self, pvdata, fn, rate=48000.0):
        """PhaseVocoder::shift with a Python fn applied elementwise on the host (it is user code)."""
        nch, size, _ = pvdata.shape
        half = size // 2
        freqs = np.zeros((nch, half))
        _check(_load().tsb_pv_freqs(self.h, pvdata.data_ptr(), nch, size, freqs.ctypes.data))
        bins = np.tile(np.arange(half) / float(size) * rate, (nch, 1))
        fb = np.ascontiguousarray(fn(bins), np.float64)
        ff = np.ascontiguousarray(fn(freqs), np.float64)
        _check(_load().tsb_pv_shift(self.h, pvdata.data_ptr(), nch, size, float(rate), fb.ctypes.data, ff.ctypes.data))

    def close(self):
        if getattr(self, "h", None):
            _load().tsb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    def __init__(self, ctx, n, precision="f64", program="pv", hop=None, rate=48000.0, ratio=1.0, fn_bins=None,
                 gain=None, win_in="sqrt_hann", win_out="sqrt_hann"):
        self.ctx, self.n = ctx, n
        self.precision, self.program = _PREC[precision], _PROG[program]
        d = _PlanDesc()
        d.n, d.precision, d.program = n, self.precision, self.program
        d.win_in, d.win_out = _WIN[win_in], _WIN[win_out]
        d.hop = float(n / 4 if hop is None else hop)
        d.rate, d.ratio = rate, ratio
        self._fb = None if fn_bins is None else np.ascontiguousarray(fn_bins, np.float64)
        self._g = None if gain is None else np.ascontiguousarray(gain, np.float64)
        if self._fb is not None:
            assert self._fb.shape == (n // 2,)
            d.fn_bins = self._fb.ctypes.data
        if self._g is not None:
            assert self._g.shape == (n // 2 + 1,)
            d.gain = self._g.ctypes.data
        h = C.c_void_p()
        _check(_load().tsb_plan_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    @property
    def out_dtype(self):
        import torch
        return torch.float64 if self.precision == F64 else torch.float32

    def _args(self, lanes, in_ptr, in_stride, in_len, out_ptr, out_stride, out_len, n_frames, in_off0, out_off0, hop,
              hop_tab_ptr, st, load_state):
        a = _ChainArgs()
        a.in_, a.in_stride, a.in_len = in_ptr, in_stride, in_len
        a.out, a.out_stride, a.out_len = out_ptr, out_stride, out_len
        a.n_frames, a.in_off0, a.out_off0 = n_frames, in_off0, out_off0
        a.hop, a.hop_tab = int(hop), hop_tab_ptr
        a.state_last_phase, a.state_phase_sum, a.state_ola = st
        a.load_state, a.lanes = int(load_state), lanes
        return a

    def render(self, x, n_frames, hop, out=None, in_off0=2048, out_off0=0, hop_tab=None, state=None,
               load_state=False, out_len=None):
        """Device path.  x: CUDA float32 [lanes, L]; returns out [lanes, out_len] (plan dtype), asynchronous on
        the context stream.  state: dict(last_phase [lanes,n/2], phase_sum [lanes,n/2], ola [lanes,n]) of CUDA
        float64 tensors, updated in place."""
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
        lanes, L = x.shape
        ht = None
        if hop_tab is not None:
            ht = torch.as_tensor(np.asarray(hop_tab, np.int32)).to(x.device) if not torch.is_tensor(hop_tab) else hop_tab
            assert ht.dtype == torch.int32 and ht.numel() == n_frames
            total = int(ht.sum().item())
        else:
            total = n_frames * int(hop)
        if out is None:
            out = torch.zeros((lanes, out_len if out_len is not None else out_off0 + total), dtype=self.out_dtype,
                              device=x.device)
        assert out.is_cuda and out.dtype == self.out_dtype and out.is_contiguous() and out.shape[0] == lanes
        st = (None, None, None)
        if state is not None:
            for k, w in (("last_phase", self.n // 2), ("phase_sum", self.n // 2), ("ola", self.n)):
                t = state[k]
                assert t.is_cuda and t.dtype == torch.float64 and t.shape == (lanes, w) and t.is_contiguous()
            st = (state["last_phase"].data_ptr(), state["phase_sum"].data_ptr(), state["ola"].data_ptr())
        a = self._args(lanes, x.data_ptr(), x.stride(0) if lanes > 1 else L, L, out.data_ptr(),
                       out.stride(0) if lanes > 1 else out.shape[1], out.shape[1], n_frames,
                       in_off0, out_off0, hop, None if ht is None else ht.data_ptr(), st, load_state)
        _check(_load().tsb_chain_render(self.ctx.h, self.h, C.byref(a)))
        self._keep = (x, out, ht, state)
        return out

    def _device_args(self, x, n_frames, hop, out, in_off0, out_off0, hop_tab, state, load_state, out_len):
        import torch
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
        lanes, L = x.shape
        ht_host = None
        ht = None
        if hop_tab is not None:
            ht_host = np.ascontiguousarray(hop_tab.cpu().numpy() if torch.is_tensor(hop_tab) else hop_tab, np.int32)
            ht = torch.from_numpy(ht_host).to(x.device)
            assert ht_host.size == n_frames
            total = int(ht_host.sum())
        else:
            total = n_frames * int(hop)
        if out is None:
            out = torch.zeros((lanes, out_len if out_len is not None else out_off0 + total), dtype=self.out_dtype,
                              device=x.device)
        assert out.is_cuda and out.dtype == self.out_dtype and out.is_contiguous() and out.shape[0] == lanes
        st = (None, None, None)
        if state is not None:
            for k, w in (("last_phase", self.n // 2), ("phase_sum", self.n // 2), ("ola", self.n)):
                t = state[k]
                assert t.is_cuda and t.dtype == torch.float64 and t.shape == (lanes, w) and t.is_contiguous()
            st = (state["last_phase"].data_ptr(), state["phase_sum"].data_ptr(), state["ola"].data_ptr())
        a = self._args(lanes, x.data_ptr(), x.stride(0) if lanes > 1 else L, L, out.data_ptr(),
                       out.stride(0) if lanes > 1 else out.shape[1], out.shape[1], n_frames,
                       in_off0, out_off0, hop, None if ht is None else ht.data_ptr(), st, load_state)
        return a, out, ht, ht_host

    def render_timepar(self, x, n_frames, hop, n_segments, out=None, in_off0=2048, out_off0=0, hop_tab=None,
                       state=None, load_state=False, out_len=None):
        """render() bit for bit (samples and PV state), with the frames of every lane cut into `n_segments`
        concurrent time segments (tsb_chain_render_timepar)."""
        a, out, ht, ht_host = self._device_args(x, n_frames, hop, out, in_off0, out_off0, hop_tab, state, load_state,
                                                out_len)
        _check(_load().tsb_chain_render_timepar(self.ctx.h, self.h, C.byref(a),
                                                None if ht_host is None else ht_host.ctypes.data, int(n_segments)))
        self._keep = (x, out, ht, state)
        return out

    def timepar(self, x, n_frames, hop, n_segments, out=None, in_off0=2048, out_off0=0, hop_tab=None, state=None,
                load_state=False, out_len=None, warmup=False, warm_hop=0, lead=0):
        """Two-step form (pre-pass, then render) for renders split across GPUs: returns a TimePar.  lead: the block's
        first `lead` frames are lead-in frames (they emit nothing and rebuild the overlap tail the block starts with)."""
        a, out, ht, ht_host = self._device_args(x, n_frames, hop, out, in_off0, out_off0, hop_tab, state, load_state,
                                                out_len)
        return TimePar(self, a, out, ht_host, n_segments, warmup, warm_hop, keep=(x, ht, state), lead=lead)

    def add_tail(self, out, pos, tail):
        """out[l][pos + i] += tail[l][i]: another block's final overlap tail onto the head of this block's output."""
        import torch
        assert tail.is_cuda and tail.dtype == torch.float64 and tail.is_contiguous() and tail.shape == (out.shape[0], self.n)
        _check(_load().tsb_add_tail(self.ctx.h, self.h, out.data_ptr(), out.stride(0), out.shape[1], int(pos),
                                    tail.data_ptr(), out.shape[0]))

    def render_host(self, x, n_frames, hop, out=None, in_off0=2048, out_off0=0, hop_tab=None, state=None,
                    load_state=False, out_len=None):
        """Host path (tsb_chain_render_host): x / out are numpy arrays or CPU torch tensors (pinned or not);
        returns when `out` holds the result."""
        import torch
        xt = torch.as_tensor(x)
        assert not xt.is_cuda and xt.dtype == torch.float32 and xt.dim() == 2 and xt.is_contiguous()
        lanes, L = xt.shape
        ht = None if hop_tab is None else np.ascontiguousarray(hop_tab, np.int32)
        total = int(ht.sum()) if ht is not None else n_frames * int(hop)
        if out is None:
            out = torch.zeros((lanes, out_len if out_len is not None else out_off0 + total), dtype=self.out_dtype)
        ot = torch.as_tensor(out)
        assert ot.dtype == self.out_dtype and ot.is_contiguous() and ot.shape[0] == lanes
        st = (None, None, None)
        if state is not None:
            for k, w in (("last_phase", self.n // 2), ("phase_sum", self.n // 2), ("ola", self.n)):
                assert state[k].dtype == np.float64 and state[k].shape == (lanes, w)
            st = tuple(state[k].ctypes.data for k in ("last_phase", "phase_sum", "ola"))
        # (a size-1 dimension reports an arbitrary stride)
        a = self._args(lanes, xt.data_ptr(), xt.stride(0) if lanes > 1 else L, L, ot.data_ptr(),
                       ot.stride(0) if lanes > 1 else ot.shape[1], ot.shape[1], n_frames,
                       in_off0, out_off0, hop, None if ht is None else ht.ctypes.data, st, load_state)
        _check(_load().tsb_chain_render_host(self.ctx.h, self.h, C.byref(a)))
        return out

    def close(self):
        if getattr(self, "h", None):
            _load().tsb_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TimePar:
    """tsb_timepar_*: pre-pass (block totals) and render of one block of frames cut into time segments."""

    def __init__(self, plan, args, out, ht_host, n_segments, warmup, warm_hop, keep, lead=0):
        self.plan, self.args, self.out, self._keep, self._ht = plan, args, out, keep, ht_host
        h = C.c_void_p()
        _check(_load().tsb_timepar_create(plan.ctx.h, plan.h, C.byref(args), None if ht_host is None else ht_host.ctypes.data,
                                          int(n_segments), int(bool(warmup)), int(warm_hop), int(lead), C.byref(h)))
        self.h = h

    @property
    def segments(self):
        return _load().tsb_timepar_segments(self.h)

    def prepass(self, want_totals=False):
        """Analyse the block and store every frame's increments.  want_totals: also return (block_total,
        block_last_phase), CUDA float64 [lanes, n/2] — an extra pass over the stored increments."""
        import torch
        if not want_totals:
            _check(_load().tsb_timepar_prepass(self.h, None, None))
            return None
        lanes, m = self.out.shape[0], self.plan.n // 2
        tot = torch.zeros((lanes, m), dtype=torch.float64, device=self.out.device)
        last = torch.zeros((lanes, m), dtype=torch.float64, device=self.out.device)
        _check(_load().tsb_timepar_prepass(self.h, tot.data_ptr(), last.data_ptr()))
        return tot, last

    def scan(self, carry_in=None, lane0=0, n_lanes=None, carry_frame=-1):
        """phase_sum after the block's last frame — or before its frame `carry_frame` — from the phase_sum before its
        first (CUDA float64 [n_lanes, n/2]; None: the state given at creation, or zeros): the block's stored
        increments added frame by frame.  lane0 / n_lanes restrict it to some lanes (a chain of ranks hands the sums
        on lane by lane)."""
        import torch
        lanes, m = self.out.shape[0], self.plan.n // 2
        n_lanes = lanes - lane0 if n_lanes is None else n_lanes
        if carry_in is not None:
            assert carry_in.is_cuda and carry_in.dtype == torch.float64 and carry_in.shape == (n_lanes, m) and carry_in.is_contiguous()
        out = torch.empty((n_lanes, m), dtype=torch.float64, device=self.out.device)
        _check(_load().tsb_timepar_scan_at(self.h, int(lane0), int(n_lanes), None if carry_in is None else carry_in.data_ptr(),
                                           out.data_ptr(), int(carry_frame)))
        self._keep_carry = carry_in
        return out

    def scan_chain(self, carry_in, flag_in, carry_out_peer, flag_out_peer, epoch, carry_frame=-1):
        """tsb_timepar_scan_chain: raw device pointers (ints; None = first / last rank)"""
        _check(_load().tsb_timepar_scan_chain(self.h, carry_in, flag_in, carry_out_peer, flag_out_peer, int(epoch),
                                              int(carry_frame)))

    def render(self):
        _check(_load().tsb_timepar_render(self.h))
        return self.out

    def close(self):
        if getattr(self, "h", None):
            _load().tsb_timepar_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stream:
    """tsb_stream_*: chunked (audio-period) front end; process(chunk) -> the samples that became final."""

    def __init__(self, plan, lanes, hop, rate=48000.0):
        self.plan, self.lanes = plan, lanes
        h = C.c_void_p()
        _check(_load().tsb_stream_create(plan.ctx.h, plan.h, lanes, float(hop), float(rate), C.byref(h)))
        self.h = h

    def process(self, chunk):
        """chunk: float32 [lanes, len] (numpy).  Returns [lanes, k] in the plan's precision, k >= 0."""
        x = np.ascontiguousarray(chunk, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.lanes
        cap = x.shape[1] + 2 * self.plan.n + 4096
        out = np.zeros((self.lanes, cap), np.float64 if self.plan.precision == F64 else np.float32)
        n = C.c_int64()
        _check(_load().tsb_stream_process(self.h, x.ctypes.data, x.shape[1], x.shape[1], out.ctypes.data, cap, cap,
                                          C.byref(n)))
        return out[:, :n.value]

    @property
    def frames(self):
        return _load().tsb_stream_frames(self.h)

    def close(self):
        if getattr(self, "h", None):
            _load().tsb_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
