"""The source-compatible C++ header set (tinyspec-cling_b200/rtinc) and the offline render driver.

CPU: the reference's own hacks compile UNCHANGED against our headers (where /root/reference exists) and the
example hacks build.  GPU: rendered output against the oracle, in fused and in immediate mode."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "tinyspec-cling_b200", "driver")
EX = os.path.join(ROOT, "tinyspec-cling_b200", "examples")
REF_HACKS = "/root/reference/hacks"


def build(hack, out=None):
    args = ["make", "-C", DRV, "HACK=" + hack]
    if out:
        args.append("OUT=" + out)
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out or os.path.join(DRV, "ts_render_" + os.path.splitext(os.path.basename(hack))[0])


@pytest.mark.skipif(not os.path.isdir(REF_HACKS), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("hack", ["pitchshift-pv", "readme", "pitchshift-simple", "karplus-strong", "sinosc", "bytebeat",
                                  "osc", "tidal", "pwm", "pwm_bb", "connect", "spectral_bytebeat/voidbubbles",
                                  "spectral_bytebeat/andshuffle", "spectral_bytebeat/knead"])
def test_reference_hacks_compile_unchanged(hack, tmp_path):
    build(os.path.join(REF_HACKS, hack + ".cpp"), out=str(tmp_path / "r"))


def test_examples_build():
    for h in ("pitchshift_cfg1", "passthrough_cfg2", "lowpass_raw", "granular", "bytebeat", "control_times"):
        assert os.path.exists(build(os.path.join(EX, h + ".cpp")))


def render(binary, *args, lanes=1, f64=False, env=None):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bin")
        cmd = [binary, "--out", out, "--timing"] + [str(a) for a in args] + (["--out-f64"] if f64 else [])
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        if r.returncode != 0:
            return r, None, None
        info = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        y = np.fromfile(out, np.float64 if f64 else np.float32).reshape(lanes, -1)
        return r, info, y


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_fused_pitchshift(precision):
    from oracle import orc
    b = build(os.path.join(EX, "pitchshift_cfg1.cpp"))
    streams, secs = 3, 0.5
    r, info, y = render(b, "--seconds", secs, "--streams", streams, "--precision", precision, lanes=2 * streams,
                        f64=precision == "f64")
    assert r.returncode == 0, r.stderr
    # all frames x all lanes in one chain launch — here, with a handful of lanes, as one time-parallel render
    # (pre-pass, scan, chain)
    assert info["mode"] == "fused" and info["time_parallel_renders"] == 1 and info["gpu_launches"] == 3
    for s in range(streams):
        x = orc.gen_signal(s, 2, int(secs * 48000))
        ref = orc.render_pv(x, 1024, 0.25, 2.0)
        assert info["frames"] == ref["frames"]
        want = ref["ola64"][:y.shape[1]].T
        got = y[2 * s:2 * s + 2]
        if precision == "f64":
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
        else:
            assert np.sqrt(np.mean((got - want) ** 2) / np.mean(want ** 2)) <= 1e-5


@pytest.mark.gpu
def test_fused_pitchshift_default_output_is_the_reference_float_stream():
    """fp64 mode without --out-f64: the fused launch itself keeps the float32 overlap tail and narrows as
    tinyspec.cpp:173-200 does, so the file is the reference's float32 stream."""
    from oracle import orc
    b = build(os.path.join(EX, "pitchshift_cfg1.cpp"))
    r, info, y = render(b, "--seconds", 1.0, "--streams", 2, lanes=4)
    assert r.returncode == 0, r.stderr
    assert info["mode"] == "fused" and info["time_parallel_renders"] == 1 and info["gpu_launches"] == 3
    for s in range(2):
        ref = orc.render_pv(orc.gen_signal(s, 2, 48000), 1024, 0.25, 2.0)
        want = ref["stream"].T
        got = y[2 * s:2 * s + 2, :want.shape[1]]
        assert (got == want).mean() > 0.999
        assert np.abs(got.astype(np.float64) - want).max() <= 2.0 ** -23 * np.abs(want).max()


@pytest.mark.gpu
def test_immediate_pitchshift_matches_reference_stream():
    """Immediate mode runs the reference's loop with device operators, float32 overlap tail included: the output
    is the reference's float32 stream up to the last bit of float rounding."""
    from oracle import orc
    b = build(os.path.join(EX, "pitchshift_cfg1.cpp"))
    r, info, y = render(b, "--seconds", 0.25, "--mode", "immediate", lanes=2)
    assert r.returncode == 0, r.stderr
    x = orc.gen_signal(0, 2, 12000)
    ref = orc.render_pv(x, 1024, 0.25, 2.0)
    assert info["frames"] == ref["frames"] and y.shape[1] == ref["stream"].shape[0]
    d = np.abs(y.T - ref["stream"])
    assert d.max() <= 1.2e-7 * np.abs(ref["stream"]).max()
    assert (d == 0).mean() > 0.99


@pytest.mark.gpu
def test_fused_passthrough_and_immediate_lowpass():
    from oracle import orc
    x = orc.gen_signal(0, 1, 24000)
    b = build(os.path.join(EX, "passthrough_cfg2.cpp"))
    r, info, y = render(b, "--seconds", 0.5, f64=True)
    assert r.returncode == 0, r.stderr
    ref = orc.render_filter(x, 2048, 0.25, None)
    assert info["frames"] == ref["frames"] and info["gpu_launches"] <= 3   # gain chain: no pre-pass / scan kernels
    assert np.abs(y[0] - ref["ola64"][:y.shape[1], 0]).max() < 1e-12
    b = build(os.path.join(EX, "lowpass_raw.cpp"))
    r, info, y = render(b, "--seconds", 0.5, "--mode", "immediate")
    assert r.returncode == 0, r.stderr
    gain = 1.0 / (1.0 + np.arange(1025) / 64.0)
    ref = orc.render_filter(x, 2048, 0.25, gain)
    assert info["frames"] == ref["frames"]
    d = np.abs(y[0] - ref["stream"][:, 0])
    assert d.max() <= 1.2e-7 * np.abs(ref["stream"]).max() and (d == 0).mean() > 0.99


@pytest.mark.gpu
def test_fused_mode_refuses_raw_access_loudly():
    b = build(os.path.join(EX, "lowpass_raw.cpp"))
    r, _, _ = render(b, "--seconds", 0.2)
    assert r.returncode != 0 and "immediate" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,env", [("granular_1800", {}),
                                      ("granular_vary", {"GRANULAR_VARY": "1", "GRANULAR_NFIRST": "64", "GRANULAR_STREAM": "2"})])
def test_immediate_granular_cfg4_bit_exact(golden, name, env):
    """cfg4: per-frame varying frame size (64..8192 or 1800) and hop, out.size != in.size.  Gather, window_hann
    and the float32 overlap-add run on the device; the result is the compiled reference's stream bit for bit."""
    from oracle import orc
    g = golden(name)
    samples, _, _, stream = (int(v) for v in g["meta"])
    b = build(os.path.join(EX, "granular.cpp"))
    x = orc.gen_signal(stream, 2, samples)
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "in.f32")
        np.ascontiguousarray(x.T).tofile(inp)
        r, info, y = render(b, "--mode", "immediate", "--in", inp, lanes=2, env=env)
    assert r.returncode == 0, r.stderr
    assert info["frames"] == len(g["t"])
    assert y.shape[1] == g["stream"].shape[0]
    assert np.array_equal(y.T, g["stream"])


@pytest.mark.gpu
def test_immediate_bytebeat_cfg5_pattern(golden):
    """cfg5 pattern at a small size: Hermitian spectrum of 2*frame points -> frft(-1) on the device -> triangular
    window (hack code) -> overlap-add; against the compiled reference's stream."""
    g = golden("bytebeat_n2048")
    n, ch, frames = (int(v) for v in g["meta"])
    b = build(os.path.join(EX, "bytebeat.cpp"))
    r, info, y = render(b, "--mode", "immediate", "--max-frames", frames, "--seconds", 100, lanes=ch,
                        env={"BYTEBEAT_N": str(n), "BYTEBEAT_CH": str(ch)})
    assert r.returncode == 0, r.stderr
    assert info["frames"] == frames and y.shape[1] == g["stream"].shape[0]
    d = np.abs(y.T - g["stream"])
    assert d.max() <= 1.2e-7 * np.abs(g["stream"]).max() and (d == 0).mean() > 0.99


@pytest.mark.gpu
def test_immediate_bytebeat_full_size(golden):
    """The real cfg5 size through the facade: frame 65536 -> frft of 131072 points on the device (four-step), the
    hack's own per-bin loops on the host; against the compiled reference's stream (every 32nd sample kept)."""
    g = golden("bytebeat_n65536")
    n, ch, frames = (int(v) for v in g["meta"])
    every = int(g["every"][0])
    b = build(os.path.join(EX, "bytebeat.cpp"))
    r, info, y = render(b, "--mode", "immediate", "--max-frames", frames, "--seconds", 100, lanes=ch,
                        env={"BYTEBEAT_N": str(n), "BYTEBEAT_CH": str(ch)})
    assert r.returncode == 0, r.stderr
    assert info["frames"] == frames and y.shape[1] == int(g["stream_len"][0])
    d = np.abs(y.T[::every] - g["stream"])
    assert d.max() <= 1.2e-7 * np.abs(g["stream"]).max() and (d == 0).mean() > 0.99


@pytest.mark.gpu
def test_control_api_inside_the_callback(golden):
    """set_time (and a rejected negative one), next_hop_samples with a fractional hop, next_hop_hz, a frame-size change
    through next_hop_ratio's default ratio, a rejected set_hop — called from inside the frame callback
    (rtlib/control.hpp:22-53).  Per frame t, base_time, the frame size and the first input sample, and the whole output
    stream, equal what the unmodified reference produced (oracle/ref/taps/control_tap.cpp)."""
    from oracle import orc
    g = golden("control_times")
    samples, stream = (int(v) for v in g["meta"])
    b = build(os.path.join(EX, "control_times.cpp"))
    x = orc.gen_signal(stream, 2, samples)
    with tempfile.TemporaryDirectory() as d:
        inp, dump = os.path.join(d, "in.f32"), os.path.join(d, "rec.bin")
        np.ascontiguousarray(x.T).tofile(inp)
        r, info, y = render(b, "--mode", "immediate", "--in", inp, lanes=2, env={"CONTROL_DUMP": dump})
        assert r.returncode == 0, r.stderr
        rec = np.fromfile(dump, np.float64).reshape(-1, 4)
    assert "negative time is not supported" in r.stderr and "ignoring invalid hop" in r.stderr
    assert info["frames"] == g["rec"].shape[0]
    assert np.array_equal(rec, g["rec"])
    want = g["stream"].reshape(-1, 2)
    assert y.shape[1] == want.shape[0] and np.array_equal(y.T, want)


@pytest.mark.gpu
def test_skip_to_now_and_mid_render_channel_change():
    """skip_to_now keeps one frame's worth of the queued input (rtlib/control.hpp:78-82): offline, where the whole input
    is queued up front, the render ends after the frame that is already scheduled.  set_num_channels mid-render drops
    the queued input in the reference (:56-57) — how much depends on the audio period — so the fused driver refuses it
    loudly instead of guessing."""
    with tempfile.TemporaryDirectory() as d:
        h1 = os.path.join(d, "skip.cpp")
        open(h1, "w").write("long fr = 0;\nnext_hop_ratio(256, 0.25);\nset_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {\n"
                            "  for (size_t c = 0; c < out.num_channels; c++) for (size_t i = 0; i < out.size; i++) out[c][i] = in[c][i];\n"
                            "  if (++fr == 5) skip_to_now();\n});\n")
        b1 = build(h1, out=os.path.join(d, "skip"))
        r, info, y = render(b1, "--mode", "immediate", "--seconds", 0.5, lanes=2)
        assert r.returncode == 0, r.stderr
        assert info["frames"] == 6          # the fifth frame called it; one more frame was still queued
        h2 = os.path.join(d, "chan.cpp")
        open(h2, "w").write("long fr = 0;\nPhaseVocoder pv;\nnext_hop_ratio(256, 0.25);\nset_process_fn([&](WaveBuf &in, WaveBuf &out, double t) {\n"
                            "  pv.analyze(in); pv.shift([](double x){ return x*2; }); pv.synthesize(out);\n"
                            "  if (++fr == 3) set_num_channels(2, 2);\n});\n")
        b2 = build(h2, out=os.path.join(d, "chan"))
        r, info, y = render(b2, "--seconds", 0.5, lanes=2)
        assert r.returncode != 0 and "set_num_channels mid-render" in r.stderr
