"""CPU-only tests of the product's host side: the C-ABI library loads and exports what include/tsb200.h
declares, the schedule / hop clock / shift map restate the reference (checked against the oracle and
the golden fixtures), and device entry points fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import tsb200
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "tsb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(tsb200.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tsb200.h but not exported"
    assert tsb200.abi_version() == 3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tsb200.TsbError, match="no usable CUDA device"):
        tsb200.Context(0, stream=None)


def test_hop_fixed(golden):
    g = golden("helpers")
    for h, (hs, hf) in zip(g["hops"], g["hop_fixed"].reshape(-1, 2)):
        assert tsb200.hop_fixed(float(h)) == (int(hs), int(hf))
    with pytest.raises(tsb200.TsbError, match="invalid hop"):
        tsb200.hop_fixed(-0.5)


@pytest.mark.parametrize("n,hop,length", [(1024, 256.0, 24000), (512, 512 * 0.2003, 16000), (64, 16.0, 6000),
                                          (4096, 1024.0, 18432), (256, 0.75, 3000), (128, 300.5, 9000),
                                          (1024, 256.0, 3000), (1024, 256.0, 3072)])
def test_schedule_matches_oracle(n, hop, length):
    """tsb_schedule_const against the oracle's restatement of tinyspec.cpp:135-168 (itself pinned to the
    compiled reference), including fractional hops, hop > frame and inputs too short for any frame."""
    x = np.zeros((length, 1), np.float32)
    ref = orc.render_pv(x, n, hop / n, 1.0, want_tables=True)["tables"]
    got = tsb200.schedule_const(length, n, hop)
    assert got["n_frames"] == len(ref["in_off"])
    assert np.array_equal(got["in_off"], ref["in_off"])
    assert np.array_equal(got["out_off"], ref["out_off"])
    assert np.array_equal(got["hop"], ref["hop"])
    assert np.array_equal(got["t"], ref["t"])
    if got["n_frames"]:
        assert got["in_off"][0] == 2048  # first advance is the initial computed_hop


def test_schedule_cfg1_frame_count():
    assert tsb200.schedule_const(2880000, 1024, 256.0)["n_frames"] == 11239
    assert tsb200.schedule_const(2880000, 1024, 256.0, max_frames=100)["n_frames"] == 100


@pytest.mark.parametrize("n", [64, 256, 1024])
@pytest.mark.parametrize("ratio", [2.0, 1.5, 0.5, 0.8, 1.0])
def test_shift_map_matches_oracle(n, ratio):
    """Bin remap of PhaseVocoder::shift: put a distinct amplitude in every bin and let the oracle's shift move
    them; every destination must hold the sum of the sources tsb_shift_map assigns to it."""
    rate = 48000.0
    pv = orc.PV(hop=n / 4.0, rate=rate)
    pv.analyze(np.random.default_rng(0).standard_normal((1, n)))
    before = pv.data[0].copy()
    pv.shift(ratio)
    after = pv.data[0]
    centres = np.arange(n // 2) / float(n) * rate
    dest = tsb200.shift_map(n, rate, centres * ratio)
    amp = np.zeros(n)
    freq = np.zeros(n)
    for j in range(n // 2):
        if dest[j] >= 0:
            amp[dest[j]] += before[j, 0]
            freq[dest[j]] = before[j, 1] * ratio
    assert np.array_equal(amp, after[:, 0])
    assert np.array_equal(freq, after[:, 1])


def _granular_callback(stream, vary):
    """the control logic of examples/granular.cpp (hacks/pitchshift-simple.cpp:3,5,10 with BASELINE cfg4's frame sizes)"""
    import math

    def cb(f, t, W):
        shift = math.sin(t) / 2 + 1
        n = (1 << (6 + ((f + 1 + stream) % 8))) if vary else 1800
        return int(W / shift), shift, n, n * (0.5 / shift)
    return cb


@pytest.mark.parametrize("vary,n_first,stream", [(False, 4096, 0), (True, 64, 2), (True, 64, 7)])
def test_granular_schedule_matches_the_oracle(vary, n_first, stream):
    """tsb200.granular_schedule (the reference's scheduler and clock driven by a hack's control logic only) against the
    frame tables of the C oracle's full render: frame sizes, input offsets, output positions, computed hops and t."""
    from oracle import orc
    samples = 60000
    x = orc.gen_signal(stream, 2, samples)
    r = orc.render_granular(x, n_first=0 if n_first == 4096 else n_first, vary=vary, stream_idx=stream)
    s = tsb200.granular_schedule(samples, _granular_callback(stream, vary), n_first=n_first)
    assert s["n_frames"] == r["frames"] > 0
    for k in ("win", "in_off", "out_off", "hop", "t"):
        assert np.array_equal(s[k], r["tables"][k]), k
