"""tests/emul/emul.py — TEST INFRASTRUCTURE: ctypes binding of the CPU replay of the chain kernel."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("TSB_EMUL_LIB") or os.path.join(HERE, "_build", "libtsb_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB)
        P = C.c_void_p
        _lib.emul_chain_render.argtypes = [C.c_int] * 5 + [C.c_double] * 3 + [P, P, P, C.c_longlong, C.c_longlong, P,
                                                                             C.c_longlong, C.c_longlong,
                                                                             C.c_longlong, C.c_longlong,
                                                                             C.c_longlong, C.c_int, P, P, P, P,
                                                                             C.c_int, C.c_int, C.c_int]
        _lib.emul_fft.argtypes = [C.c_int, C.c_int, P, P]
        _lib.emul_pv_math.argtypes = [C.c_longlong, P, P, P, P, P, P, P]
        _lib.emul_pv_math.restype = None
        _lib.emul_chain_render_ex.argtypes = _lib.emul_chain_render.argtypes[:-1] + [C.c_int, C.c_int, C.c_int, P, P, C.c_int]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def wrap_counters():
    """(doubtful wraps the fp32 PV re-evaluated in double precision, of them decided the other way) since the last call"""
    a, b = C.c_longlong(), C.c_longlong()
    lib().emul_wrap_counters(C.byref(a), C.byref(b))
    return a.value, b.value


def pv_math(y, x):
    """(fast_atan2(y, x), sqrt_pos(|x|), wrap_quotient(x), fast_sincos(y)) element-wise: the fp64 kernels of the PV stages"""
    y = np.ascontiguousarray(y, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    n = y.size
    at, sq, sn, cs = (np.empty(n) for _ in range(4))
    wr = np.empty(n, np.int32)
    lib().emul_pv_math(n, _p(y), _p(x), _p(at), _p(sq), _p(wr), _p(sn), _p(cs))
    return at, sq, wr, sn, cs


WIN = {"none": 0, None: 0, "hann": 1, "sqrt_hann": 2}


def chain_render(x, n, hop, n_frames, precision="f64", program="pv", ratio=2.0, gain=None, fn_bins=None,
                 win_in="sqrt_hann", win_out="sqrt_hann", rate=48000.0, hop_d=None, hop_tab=None, in_off0=2048,
                 out_off0=0, state=None, out_len=None, prev_hop=0):
    """x: planar float32 [lanes, L].  Returns (out [lanes, out_len], state dict)."""
    x = np.ascontiguousarray(x, np.float32)
    lanes, L = x.shape
    m = n // 2
    if out_len is None:
        out_len = out_off0 + (int(np.sum(hop_tab)) if hop_tab is not None else n_frames * hop)
    dt = np.float64 if precision == "f64" else np.float32
    out = np.zeros((lanes, out_len), dt)
    prec = {"f32": 0, "f64": 1, "f64o32": 2}[precision]
    load = state is not None
    if state is None:
        state = dict(last_phase=np.zeros((lanes, m)), phase_sum=np.zeros((lanes, m)), ola=np.zeros((lanes, n)))
    ht = None if hop_tab is None else np.ascontiguousarray(hop_tab, np.int32)
    g = None if gain is None else np.ascontiguousarray(gain, np.float64)
    fb = None if fn_bins is None else np.ascontiguousarray(fn_bins, np.float64)
    r = lib().emul_chain_render(prec, int(np.log2(n)), 0 if program == "pv" else 1,
                                WIN[win_in], WIN[win_out], float(hop if hop_d is None else hop_d), rate, ratio,
                                _p(fb), _p(g), _p(x), L, L, _p(out), out_len, out_len, n_frames, in_off0, out_off0,
                                int(hop), _p(ht), _p(state["last_phase"]), _p(state["phase_sum"]), _p(state["ola"]),
                                int(load), lanes, int(prev_hop))
    assert r == 0
    return out, state


def segment_render(x, n, hop, ratio, in_off0, n_frames, lead, warm, mode, state, warm_phase, inc=None, precision="f64",
                   out_len=None):
    """One time segment as the kernel runs it (ChainArgs::seg_tab row): `n_frames` frames from input offset `in_off0`,
    the first `lead` of them lead-in frames, optionally preceded by a warm-up frame; mode 1 = pre-pass."""
    x = np.ascontiguousarray(x, np.float32)
    lanes, L = x.shape
    dt = np.float64 if precision == "f64" else np.float32
    out_len = n_frames * hop if out_len is None else out_len
    out = np.zeros((lanes, max(out_len, 1)), dt)
    r = lib().emul_chain_render_ex({"f32": 0, "f64": 1, "f64o32": 2}[precision], int(np.log2(n)), 0, 2, 2, float(hop), 48000.0,
                                   ratio, None, None, _p(x), L, L, _p(out), out.shape[1], out_len, n_frames, in_off0, 0,
                                   int(hop), None, _p(state["last_phase"]), _p(state["phase_sum"]), _p(state["ola"]), 1,
                                   lanes, mode, int(warm), int(hop), _p(warm_phase), _p(inc), int(lead))
    assert r == 0
    return out


class ReplayBackend:
    """CPU stand-in for the device side of tinyspec-cling_b200/shard.py (kernel phases replayed thread by thread),
    so the multi-rank orchestration can be tested with gloo in a container without a GPU."""

    def __init__(self, n, hop_d, ratio, precision="f64", rate=48000.0):
        self.n, self.hop_d, self.ratio, self.precision, self.rate = n, hop_d, ratio, precision, rate

    def _call(self, out_len, mode, state, inc=None):
        b, x = self._b, self._x
        lanes, L = x.shape
        ht = b["hop_tab"]
        dt = np.float64 if self.precision == "f64" else np.float32
        out = np.zeros((lanes, max(out_len, 1)), dt)
        r = lib().emul_chain_render_ex(1 if self.precision == "f64" else 0, int(np.log2(self.n)), 0, 2, 2, float(self.hop_d),
                                       self.rate, self.ratio, None, None, _p(x), L, L, _p(out), out.shape[1], out_len,
                                       int(ht.size), b["in_off0"], 0, int(ht[0]), _p(ht), _p(state["last_phase"]),
                                       _p(state["phase_sum"]), _p(state["ola"]), 1, lanes, mode, int(b["warm"]),
                                       int(b["warm_hop"]), _p(self._wp), _p(inc), int(b["lead"]))
        assert r == 0
        return out

    def prepass(self, x, block):
        lanes = x.shape[0]
        m = self.n // 2
        self._x, self._b = np.ascontiguousarray(x, np.float32), block
        self._wp = np.zeros((lanes, m))
        st = dict(last_phase=np.zeros((lanes, m)), phase_sum=np.zeros((lanes, m)), ola=np.zeros((lanes, self.n)))
        self._inc = np.zeros((lanes, max(int(block["hop_tab"].size), 1), m))
        self._call(0, 1, st, self._inc)

    def zero_carry(self):
        return np.zeros((self._x.shape[0], self.n // 2))

    def scan(self, carry, lane0=0, n_lanes=None, carry_frame=-1):
        """the running sum of lanes lane0.. before frame `carry_frame` of the block (after its last frame when < 0),
        from the sum before its first: the block's increments added frame by frame, exactly the additions of the
        sequential render (tsb_timepar_scan_at)"""
        lanes = self._x.shape[0]
        n_lanes = lanes - lane0 if n_lanes is None else n_lanes
        ps = np.ascontiguousarray(carry, np.float64).copy()
        if not hasattr(self, "_carry"):
            self._carry = np.zeros((lanes, self.n // 2))
        self._carry[lane0:lane0 + n_lanes] = ps
        noted = None
        for f in range(int(self._b["hop_tab"].size)):
            if f == carry_frame:
                noted = ps.copy()
            ps = ps + self._inc[lane0:lane0 + n_lanes, f, :]
        return ps if noted is None else noted

    def render(self):
        lanes = self._x.shape[0]
        m = self.n // 2
        st = dict(last_phase=np.zeros((lanes, m)), phase_sum=np.ascontiguousarray(self._carry, np.float64).copy(),
                  ola=np.zeros((lanes, self.n)))
        out = self._call(int(self._b["hop_tab"].sum()), 0, st)
        return out[:, self._b["out_pad"]:]
