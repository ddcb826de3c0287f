"""The C oracle (oracle/oracle.c) pinned against fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import orc


def _slices(sizes, per):
    off = 0
    for n in sizes:
        yield n, off, off + per(n)
        off += per(n)


def test_windows(golden):
    g = golden("helpers")
    for n, a, b in _slices(g["frft_sizes"], lambda n: 2 * n):
        ones = np.ones((2, n))
        assert np.array_equal(orc.window(ones, "hann").ravel(), g["hann"][a:b])
        assert np.array_equal(orc.window(ones, "sqrt_hann").ravel(), g["sqrt_hann"][a:b])


def test_frft(golden):
    g = golden("helpers")
    exps = g["frft_exps"]
    off_out = 0
    for n, a, b in _slices(g["frft_sizes"], lambda n: 2 * n):
        x = g["frft_in"][a:b].reshape(2, n)
        for e in exps:
            want = g["frft_out"][off_out:off_out + 2 * n].reshape(2, n)
            off_out += 2 * n
            got = orc.frft(x, float(e), inplace=False)
            assert np.abs(got - want).max() <= 1e-15 * max(1.0, np.abs(want).max()), (n, e)
        got = orc.frft(x, 1.0, inplace=True)
        want = g["frft_inplace"][a:b].reshape(2, n)
        assert np.abs(got - want).max() <= 1e-15 * np.abs(want).max()
        # documented identities: e=-1 -> G[k]/sqrt(N), e=+1 -> G[(N-k)%N]/sqrt(N)
        G = np.fft.fft(x, axis=-1) / np.sqrt(n)
        assert np.abs(orc.frft(x, -1.0, inplace=False) - G).max() < 1e-13
        assert np.abs(orc.frft(x, 1.0, inplace=False) - np.roll(G[:, ::-1], 1, axis=-1)).max() < 1e-13


def test_set_hop_fixed_point(golden):
    g = golden("helpers")
    fixed = g["hop_fixed"].reshape(-1, 2)
    for h, (hs, hf) in zip(g["hops"], fixed):
        ok, s, f = orc.set_hop(float(h))
        assert ok and s == hs and f == hf
    assert orc.set_hop(-1.0)[0] == 0  # rejected, rtlib/control.hpp:23


def test_pv_steps(golden):
    g = golden("helpers")
    off = 0
    for n in g["pv_sizes"]:
        for ratio in g["pv_ratios"]:
            pv = orc.PV(hop=n / 4.0)
            for _ in range(3):
                sl = slice(off, off + 2 * n)
                sl2 = slice(2 * off, 2 * off + 4 * n)
                pv.analyze(g["pv_in"][sl].reshape(2, n))
                assert np.array_equal(pv.data.ravel(), g["pv_analyzed"][sl2])
                pv.shift(float(ratio))
                assert np.array_equal(pv.data.ravel(), g["pv_shifted"][sl2])
                out = pv.synthesize()
                assert np.array_equal(out.ravel(), g["pv_out"][sl])
                assert np.array_equal(pv.phase_sum.ravel(), g["pv_phase_sum"][sl])
                assert np.array_equal(pv.last_phase.ravel(), g["pv_last_phase"][sl])
                off += 2 * n


@pytest.mark.parametrize("name", ["pv_cfg1_short", "pv_x1p5_n256", "pv_down_n512_fracthop"])
def test_pv_render(golden, name):
    g = golden(name)
    n, ch, samples, stream = (int(v) for v in g["meta"])
    hop_ratio, shift = (float(v) for v in g["params"])
    x = orc.gen_signal(stream, ch, samples)
    r = orc.render_pv(x, n, hop_ratio, shift, want_frames=True, state_every=1, want_tables=True)
    assert np.array_equal(r["stream"], g["stream"])          # float32 OLA stream, bit-exact
    f = g["frame"]
    assert np.array_equal(r["frame_out"][f], g["out"])       # per-frame doubles
    assert np.array_equal(r["state"][f][:, 0], g["last_phase"])
    assert np.array_equal(r["state"][f][:, 1], g["phase_sum"])
    assert np.array_equal(r["tables"]["t"][f], g["t"])
    # frame f reads input [I_f, I_f+N): check the gather against the dumped `in`
    for i, fr in enumerate(f):
        o = r["tables"]["in_off"][fr]
        assert np.array_equal(x[o:o + n].T.astype(np.float64), g["in"][i])
    if hop_ratio * n == int(hop_ratio * n):
        assert r["frames"] == orc.n_frames(samples, n, int(hop_ratio * n))
        assert r["tables"]["in_off"][0] == 2048


def test_pv_cfg3_shape(golden):
    g = golden("pv_cfg3_shape")
    n, ch, samples, stream = (int(v) for v in g["meta"])
    x = orc.gen_signal(stream, ch, samples)
    r = orc.render_pv(x, n, 0.25, 1.5, state_every=1)
    assert np.array_equal(r["stream"][-8192:], g["stream"])
    assert np.array_equal(r["state"][g["frame"]][:, 1], g["phase_sum"])


def test_pv_cfg1_60s(golden):
    """Full-length cfg1 (60 s, 11,239 frames): final PV state and the stream tail."""
    g = golden("pv_cfg1_60s_tail")
    x = orc.gen_signal(0, 2, 2880000)
    r = orc.render_pv(x, 1024, 0.25, 2.0)
    assert r["frames"] == 11239
    assert r["stream"].shape[0] == int(g["stream_len"][0]) // 2
    assert np.array_equal(r["stream"][-8192:], g["stream"])
    assert np.array_equal(r["phase_sum"], g["phase_sum"][-1])
    assert np.array_equal(r["last_phase"], g["last_phase"][-1])


@pytest.mark.parametrize("name", ["filter_pass_n2048", "filter_lp_n2048", "filter_lp_n512_stereo"])
def test_filter_render(golden, name):
    g = golden(name)
    n, ch, samples, filt = (int(v) for v in g["meta"])
    hop_ratio = float(g["params"][0])
    x = orc.gen_signal(0, ch, samples)
    gain = 1.0 / (1.0 + np.arange(n // 2 + 1) / 64.0) if filt else None
    r = orc.render_filter(x, n, hop_ratio, gain, want_frames=True)
    assert np.array_equal(r["stream"], g["stream"])
    assert np.array_equal(r["frame_out"][g["frame"]], g["out"])


def test_passthrough_cola_known_answer(golden):
    """sqrt-Hann analysis+synthesis at hop N/4 gives y[p] = 2*x[p+2048] in the steady region."""
    n, hop = 2048, 512
    x = orc.gen_signal(0, 1, 24000)
    r = orc.render_filter(x, n, 0.25, None)
    F = r["frames"]
    p = np.arange(n - hop, (F - 1) * hop)
    assert np.abs(r["ola64"][p, 0] - 2.0 * x[p + 2048, 0].astype(np.float64)).max() < 1e-13


@pytest.mark.parametrize("name,vary", [("granular_1800", False), ("granular_vary", True)])
def test_granular(golden, name, vary):
    g = golden(name)
    samples, _, n_first, stream = (int(v) for v in g["meta"])
    x = orc.gen_signal(stream, 2, samples)
    r = orc.render_granular(x, n_first=n_first, vary=vary, stream_idx=stream)
    assert r["frames"] == len(g["t"])
    assert np.array_equal(r["tables"]["win"], g["win"])
    assert np.array_equal(r["tables"]["t"], g["t"])
    assert np.array_equal(r["stream"], g["stream"])


@pytest.mark.parametrize("name", ["bytebeat_n2048", "bytebeat_n65536"])
def test_bytebeat_restatement(golden, name):
    """The numpy restatement of the spectral-bytebeat pattern (BASELINE cfg5: static 2n-point buffer, frft(-1) in
    place, triangular window) against what the compiled reference produced — at the small size and at cfg5's own
    frame 65536 (every 32nd sample of each frame; any sample depends on every bin)."""
    g = golden(name)
    n, ch, frames = (int(v) for v in g["meta"])
    every = int(g["every"][0])
    r = orc.render_bytebeat(n, ch, frames)
    peak = np.abs(g["out"]).max()
    assert np.abs(r["out"][:, :, ::every] - g["out"]).max() <= 1e-12 * peak
    assert np.abs(r["out"].sum(axis=2) - g["frame_sum"]).max() <= 1e-9 * np.abs(g["frame_sum"]).max()
    assert np.abs((r["out"] ** 2).sum(axis=2) - g["frame_sumsq"]).max() <= 1e-12 * g["frame_sumsq"].max()
    # the reference's own stream is the float32 narrowing of the overlap-add (tinyspec.cpp:173-200)
    L = int(g["stream_len"][0])
    s = r["ola64"][:, :L].T[::every]
    assert np.abs(s - g["stream"]).max() <= 2e-7 * np.abs(g["stream"]).max()
    # the in-place quirk matters: with bins 0 and n cleared every frame the result is measurably different
    if name == "bytebeat_n2048":
        buf_clean = np.zeros((ch, 2 * n), np.complex128)
        x = orc.bytebeat_spectrum(1, ch, n)
        buf_clean[:, 1:n] = x[:, 1:n]
        buf_clean[:, 2 * n - 1:n:-1] = np.conj(x[:, 1:n])
        clean = (np.fft.fft(buf_clean, axis=1) / np.sqrt(2.0 * n))[:, :n].real * (1.0 - np.abs(1.0 - np.arange(n) / (n / 2.0)))
        assert np.abs(clean - g["out"][1]).max() > 1e-6 * peak
