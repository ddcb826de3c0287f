"""tests/emul/emul.py — TEST INFRASTRUCTURE: ctypes binding of the CPU replay of the chain kernel."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libtsb_emul.so")
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
                                                                             C.c_int, C.c_int]
        _lib.emul_fft.argtypes = [C.c_int, C.c_int, P, P]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


WIN = {"none": 0, None: 0, "hann": 1, "sqrt_hann": 2}


def chain_render(x, n, hop, n_frames, precision="f64", program="pv", ratio=2.0, gain=None, fn_bins=None,
                 win_in="sqrt_hann", win_out="sqrt_hann", rate=48000.0, hop_d=None, hop_tab=None, in_off0=2048,
                 out_off0=0, state=None, out_len=None):
    """x: planar float32 [lanes, L].  Returns (out [lanes, out_len], state dict)."""
    x = np.ascontiguousarray(x, np.float32)
    lanes, L = x.shape
    m = n // 2
    if out_len is None:
        out_len = out_off0 + (int(np.sum(hop_tab)) if hop_tab is not None else n_frames * hop)
    dt = np.float64 if precision == "f64" else np.float32
    out = np.zeros((lanes, out_len), dt)
    load = state is not None
    if state is None:
        state = dict(last_phase=np.zeros((lanes, m)), phase_sum=np.zeros((lanes, m)), ola=np.zeros((lanes, n)))
    ht = None if hop_tab is None else np.ascontiguousarray(hop_tab, np.int32)
    g = None if gain is None else np.ascontiguousarray(gain, np.float64)
    fb = None if fn_bins is None else np.ascontiguousarray(fn_bins, np.float64)
    r = lib().emul_chain_render(1 if precision == "f64" else 0, int(np.log2(n)), 0 if program == "pv" else 1,
                                WIN[win_in], WIN[win_out], float(hop if hop_d is None else hop_d), rate, ratio,
                                _p(fb), _p(g), _p(x), L, L, _p(out), out_len, out_len, n_frames, in_off0, out_off0,
                                int(hop), _p(ht), _p(state["last_phase"]), _p(state["phase_sum"]), _p(state["ola"]),
                                int(load), lanes)
    assert r == 0
    return out, state
