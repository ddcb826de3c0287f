"""GPU parity of the one-helper-per-call device operators (tsb200.h "immediate mode") against the C oracle
and the golden fixtures of the compiled reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import tsb200
    from oracle import orc
    return dict(torch=torch, tsb=tsb200, orc=orc, ctx=tsb200.Context(0))


def dev(env, a):
    return env["torch"].from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [1, 2, 8, 64, 100, 512, 1000, 1800, 4096, 8192, 65536, 262144])
def test_fft_forward_any_n(env, n):
    rng = np.random.default_rng(n)
    rows = 3 if n <= 8192 else 2
    x = rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))
    got = env["ctx"].fft_forward(dev(env, x)).cpu().numpy()
    want = np.fft.fft(x, axis=-1)
    assert np.abs(got - want).max() <= 2e-14 * np.abs(want).max() * max(1.0, np.log2(max(n, 2)))


def test_fft_in_place_and_large_batch(env):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((700, 2048)) + 1j * rng.standard_normal((700, 2048))
    xd = dev(env, x)
    env["ctx"].fft_forward(xd, out=xd)
    assert np.abs(xd.cpu().numpy() - np.fft.fft(x, axis=-1)).max() < 1e-11


@pytest.mark.parametrize("kind", ["hann", "sqrt_hann"])
@pytest.mark.parametrize("n", [8, 100, 1024])
def test_window_bit_exact(env, kind, n):
    x = np.random.default_rng(0).standard_normal((2, n))
    got = env["ctx"].window(dev(env, x), kind).cpu().numpy()
    assert np.array_equal(got, env["orc"].window(x, kind))


def test_frft_vs_golden(env, golden):
    """frft for exponents 1, -1, 2, 0.5 on sizes 8, 64, 100, 512: out of place and in place."""
    g = golden("helpers")
    off_in = off_out = 0
    for n in g["frft_sizes"]:
        x = g["frft_in"][off_in:off_in + 2 * n].reshape(2, n)
        for e in g["frft_exps"]:
            want = g["frft_out"][off_out:off_out + 2 * n].reshape(2, n)
            off_out += 2 * n
            xd = dev(env, x)
            out = env["torch"].zeros_like(xd)
            env["ctx"].frft(xd, float(e), out=out)
            assert np.abs(out.cpu().numpy() - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (n, e)
        want = g["frft_inplace"][off_in:off_in + 2 * n].reshape(2, n)
        xd = dev(env, x)
        env["ctx"].frft(xd, 1.0)
        assert np.abs(xd.cpu().numpy() - want).max() <= 1e-13 * np.abs(want).max()
        off_in += 2 * n


@pytest.mark.parametrize("n,e", [(1800, 1.0), (1800, -1.0), (4096, 0.3), (131072, -1.0)])
def test_frft_vs_oracle(env, n, e):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    want = env["orc"].frft(x, e, inplace=True)
    xd = dev(env, x)
    env["ctx"].frft(xd, e)
    assert np.abs(xd.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()


def test_frft_round_trip_and_empty(env):
    x = np.random.default_rng(2).standard_normal((4, 2048)) + 0j
    xd = dev(env, x)
    env["ctx"].frft(xd, 1.0)
    env["ctx"].frft(xd, -1.0)
    assert np.abs(xd.cpu().numpy() - x).max() < 1e-12
    empty = env["torch"].zeros((0, 16), dtype=env["torch"].complex128, device="cuda")
    env["ctx"].frft(empty, 1.0)  # N*C == 0 returns silently (rtlib/fft.hpp:8-9)


def test_pv_steps_vs_golden(env, golden):
    """PhaseVocoder::analyze(WaveBuf) -> shift -> synthesize(WaveBuf) rebuilt from the device helpers, three
    frames at sizes 64/256 and ratios 2, 1.5, 0.5, against the compiled reference's dumps (all N bins)."""
    torch, ctx = env["torch"], env["ctx"]
    g = golden("helpers")
    off = 0
    for n in g["pv_sizes"]:
        n = int(n)
        for ratio in g["pv_ratios"]:
            hop = n / 4.0
            last = torch.zeros((2, n), dtype=torch.float64, device="cuda")
            psum = torch.zeros((2, n), dtype=torch.float64, device="cuda")
            pvd = torch.zeros((2, n, 2), dtype=torch.float64, device="cuda")
            for _ in range(3):
                sl, sl2 = slice(off, off + 2 * n), slice(2 * off, 2 * off + 4 * n)
                wave = dev(env, g["pv_in"][sl].reshape(2, n))
                ctx.window(wave, "sqrt_hann")                       # rtlib/pv.hpp:5
                spec = torch.complex(wave, torch.zeros_like(wave))  # fill_from, :6
                ctx.frft(spec, 1.0)                                 # :7
                ctx.pv_analyze(spec, last, pvd, hop)                # :8
                want = g["pv_analyzed"][sl2].reshape(2, n, 2)
                got = pvd.cpu().numpy()
                assert np.abs(got[..., 0] - want[..., 0]).max() < 1e-13
                assert np.abs(got[..., 1] - want[..., 1]).max() < 1e-6   # Hz; 1e-9 rad of phase ~ 3e-8 Hz
                ctx.pv_shift(pvd, lambda v: v * float(ratio))
                want = g["pv_shifted"][sl2].reshape(2, n, 2)
                got = pvd.cpu().numpy()
                assert np.abs(got[..., 0] - want[..., 0]).max() < 1e-13
                assert np.abs(got[..., 1] - want[..., 1]).max() < 1e-6
                ctx.pv_synthesize(spec, psum, pvd, hop)             # :12
                ctx.frft(spec, -1.0)                                # :13
                out = spec.real.contiguous()                        # :14
                ctx.window(out, "sqrt_hann")                        # :15
                want = g["pv_out"][sl].reshape(2, n)
                assert np.abs(out.cpu().numpy() - want).max() <= 1e-9 * np.abs(want).max()
                assert np.abs(psum.cpu().numpy() - g["pv_phase_sum"][sl].reshape(2, n)).max() < 1e-8
                assert np.abs(last.cpu().numpy() - g["pv_last_phase"][sl].reshape(2, n)).max() < 1e-12
                off += 2 * n


# ---- fused spectral synthesis (BASELINE cfg5): tsb_spectral_render ----
@pytest.mark.parametrize("name", ["bytebeat_n2048", "bytebeat_n65536"])
def test_spectral_render_matches_the_compiled_reference(env, golden, name):
    """Generator on the device -> inverse real FFT of 2*frame points (four-step) -> triangular window -> overlap-add,
    against the per-frame double `out` buffers and the stream of the UNMODIFIED reference running
    oracle/ref/taps/bytebeat_tap.cpp — at cfg5's real frame size 65536 (FFT of 131072 points) too."""
    torch = env["torch"]
    g = golden(name)
    n, ch, frames = (int(v) for v in g["meta"])
    every, hop = int(g["every"][0]), n // 4
    peak = np.abs(g["out"]).max()
    # frame by frame (each into its own zeroed row) = the hack's `out` buffer of that frame; the carried bins 0 / n
    # travel through state_edge
    edge = torch.zeros((ch, 2), dtype=torch.float64, device="cuda")
    for f in range(frames):
        o = env["ctx"].spectral_render(n, ch, 1, hop, first_frame=f, state_edge=edge, load_state=f > 0)
        got = o.cpu().numpy()[:, :n]
        assert np.abs(got[:, ::every] - g["out"][f]).max() <= 1e-9 * peak, f
    # all frames in one call: the overlap-added stream (the reference's is its float32 narrowing)
    out = env["ctx"].spectral_render(n, ch, frames, hop)
    L = int(g["stream_len"][0])
    s = out.cpu().numpy()[:, :L].T[::every]
    assert np.abs(s - g["stream"]).max() <= 2e-7 * np.abs(g["stream"]).max()


@pytest.mark.parametrize("n,ch,frames", [(256, 3, 5), (512, 1, 4), (4096, 2, 5), (8192, 2, 3), (16384, 5, 3), (32768, 1, 3),
                                         (65536, 64, 2), (131072, 2, 3), (262144, 1, 2)])
def test_spectral_render_vs_oracle_all_sizes(env, n, ch, frames):
    """every supported frame size (both splits of the four-step transform, odd and even log2) against the numpy
    restatement, cfg5's 64 channels included; 1e-9 of peak (BASELINE.json fp64 bound), measured ~1e-13"""
    orc = env["orc"]
    hop = n // 4
    ref = orc.render_bytebeat(n, ch, frames, hop)
    got = env["ctx"].spectral_render(n, ch, frames, hop).cpu().numpy()
    assert got.shape == ref["ola64"].shape
    assert np.abs(got - ref["ola64"]).max() <= 1e-9 * np.abs(ref["ola64"]).max()
    assert np.abs(got - ref["ola64"]).max() <= 1e-11 * np.abs(ref["ola64"]).max()   # (what it actually achieves)


def test_spectral_render_caller_spectrum_and_continuation(env):
    """TSB_GEN_BUFFER with the generator's own values is the TSB_GEN_HASH render (to the last bits: the kernel divides
    by a corrected Newton reciprocal); a render continued
    through state_edge and the tail in `out` is the whole render bit for bit; an arbitrary (non-generator) spectrum
    matches the oracle fed the same values; odd hop / odd offset take the unaligned store path."""
    torch, orc = env["torch"], env["orc"]
    n, ch, frames, hop = 4096, 2, 6, 1024
    whole = env["ctx"].spectral_render(n, ch, frames, hop)
    spec = np.stack([orc.bytebeat_spectrum(f, ch, n) for f in range(frames)])
    via_buf = env["ctx"].spectral_render(n, ch, frames, hop, spectrum=dev(env, spec))
    assert float((whole - via_buf).abs().max()) <= 1e-14 * float(whole.abs().max())
    out = torch.zeros_like(whole)
    edge = torch.zeros((ch, 2), dtype=torch.float64, device="cuda")
    env["ctx"].spectral_render(n, ch, 2, hop, out=out, state_edge=edge)
    env["ctx"].spectral_render(n, ch, frames - 2, hop, out=out, first_frame=2, out_off0=2 * hop, state_edge=edge, load_state=True)
    assert torch.equal(out, whole)
    rng = np.random.default_rng(5)
    arb = rng.standard_normal((frames, ch, n)) + 1j * rng.standard_normal((frames, ch, n))
    hop2 = 1001
    ref = orc.render_bytebeat(n, ch, frames, hop2, spectrum=lambda f: arb[f])
    got = env["ctx"].spectral_render(n, ch, frames, hop2, spectrum=dev(env, arb)).cpu().numpy()
    assert np.abs(got - ref["ola64"]).max() <= 1e-11 * np.abs(ref["ola64"]).max()


def test_spectral_render_rejects_bad_arguments(env):
    tsb = env["tsb"]
    with pytest.raises(tsb.TsbError, match="powers of two"):
        env["ctx"].spectral_render(3000, 1, 1, 750)
    with pytest.raises(tsb.TsbError, match="powers of two"):
        env["ctx"].spectral_render(128, 1, 1, 32)


# ---- the rest of the PhaseVocoder surface against the compiled reference (oracle/ref/taps/pvmore_tap.cpp) ----
def _pv_frame(env, wave_np, last, psum, pvd, hop, shift=None):
    """PhaseVocoder::analyze(WaveBuf) [-> shift] -> synthesize(WaveBuf) from the device helpers (rtlib/pv.hpp:4-16)"""
    torch, ctx = env["torch"], env["ctx"]
    wave = dev(env, wave_np)
    ctx.window(wave, "sqrt_hann")
    spec = torch.complex(wave, torch.zeros_like(wave))
    ctx.frft(spec, 1.0)
    ctx.pv_analyze(spec, last, pvd, hop)
    if shift is not None:
        ctx.pv_shift(pvd, shift)
    ctx.pv_synthesize(spec, psum, pvd, hop)
    ctx.frft(spec, -1.0)
    out = spec.real.contiguous()
    ctx.window(out, "sqrt_hann")
    return out.cpu().numpy()


def _pv_check(g, tag, f, n, out, last, psum, pvd):
    want = g[tag + "_out"].reshape(-1, 2, n)[f]
    assert np.abs(out - want).max() <= 1e-9 * max(np.abs(want).max(), 1e-300), (tag, f)
    d = g[tag + "_data"].reshape(-1, 2, n, 2)[f]
    got = pvd.cpu().numpy()
    assert np.abs(got[..., 0] - d[..., 0]).max() < 1e-12, (tag, f)
    assert np.abs(got[..., 1] - d[..., 1]).max() < 1e-5, (tag, f)          # Hz
    assert np.abs(psum.cpu().numpy() - g[tag + "_phase_sum"].reshape(-1, 2, n)[f]).max() < 1e-8, (tag, f)
    assert np.abs(last.cpu().numpy() - g[tag + "_last_phase"].reshape(-1, 2, n)[f]).max() < 1e-11, (tag, f)


def test_pv_without_shift_nonlinear_fn_buffer_fn_own_hop(env, golden):
    """PV without shift (all N bins carry state), a non-linear fn (several sources per destination, gaps), the
    shift(function<void(WaveBuf&)>) overload, a per-object hop with use_global_hop(), reset(last) and reset(sum) one at a
    time, and fill_from — each frame against the unmodified reference's dumps of out, data, phase_sum, last_phase."""
    torch = env["torch"]
    g = golden("pvmore")
    n = int(g["n"][0])
    frames_in = g["in"].reshape(-1, 2, n)
    zeros = lambda *s: torch.zeros(s, dtype=torch.float64, device="cuda")
    idx = 0
    # 1. no shift
    last, psum, pvd = zeros(2, n), zeros(2, n), zeros(2, n, 2)
    for f in range(4):
        out = _pv_frame(env, frames_in[idx], last, psum, pvd, 32.0)
        _pv_check(g, "noshift", f, n, out, last, psum, pvd)
        idx += 1
    # 2. non-linear fn
    last, psum, pvd = zeros(2, n), zeros(2, n), zeros(2, n, 2)
    for f in range(4):
        out = _pv_frame(env, frames_in[idx], last, psum, pvd, 32.0, lambda x: 300.0 + 0.6 * x + x * x / 40000.0)
        _pv_check(g, "nonlin", f, n, out, last, psum, pvd)
        idx += 1
    # 3. fn on the whole buffer of bins / frequencies
    last, psum, pvd = zeros(2, n), zeros(2, n), zeros(2, n, 2)
    def buf_fn(b):
        c = np.arange(b.shape[0])[:, None]
        j = np.arange(b.shape[1])[None, :]
        return b * (1.25 + 0.25 * c) + 3.0 * (j % 5)
    for f in range(3):
        out = _pv_frame(env, frames_in[idx], last, psum, pvd, 32.0, buf_fn)
        _pv_check(g, "bufshift", f, n, out, last, psum, pvd)
        idx += 1
    # 4. per-object hop 48.5, back to the global 32 at frame 2, reset(last) at 3, reset(sum) at 4
    last, psum, pvd = zeros(2, n), zeros(2, n), zeros(2, n, 2)
    hops = g["ownhop_hop"]
    assert list(hops) == [48.5, 48.5, 32.0, 32.0, 32.0, 32.0]
    for f in range(6):
        if f == 3:
            last.zero_()
        if f == 4:
            psum.zero_()
        out = _pv_frame(env, frames_in[idx], last, psum, pvd, float(hops[f]), lambda x: x * 1.5)
        _pv_check(g, "ownhop", f, n, out, last, psum, pvd)
        idx += 1
    # 5. fill_from: a copy continues from the same state (global hop), no shift
    out = _pv_frame(env, frames_in[idx], last.clone(), psum.clone(), pvd.clone(), 32.0)
    want = g["copy_out"].reshape(2, n)
    assert np.abs(out - want).max() <= 1e-9 * np.abs(want).max()


# ---- batched granular render with per-frame sizes (BASELINE cfg4): tsb_granular_render ----
def _granular_cb(stream, vary):
    import math

    def cb(f, t, W):
        shift = math.sin(t) / 2 + 1
        n = (1 << (6 + ((f + 1 + stream) % 8))) if vary else 1800
        return int(W / shift), shift, n, n * (0.5 / shift)
    return cb


@pytest.mark.parametrize("name,vary", [("granular_1800", False), ("granular_vary", True)])
def test_granular_render_is_the_reference_stream(env, golden, name, vary):
    """one launch: gather, window_hann, the hack's nearest-neighbour resample and the float32 overlap-add with a frame
    size (64..8192 or 1800), hop and output length that change every frame — the compiled reference's stream bit for bit"""
    tsb, orc = env["tsb"], env["orc"]
    g = golden(name)
    samples, _, n_first, stream = (int(v) for v in g["meta"])
    x = orc.gen_signal(stream, 2, samples)
    s = tsb.granular_schedule(samples, _granular_cb(stream, vary), n_first=n_first or 4096)
    assert s["n_frames"] == len(g["t"]) and np.array_equal(s["osize"], g["osize"])
    y = env["ctx"].granular_render(dev(env, x.T), 2, [s]).cpu().numpy()
    assert y.shape[1] == g["stream"].shape[0]
    assert np.array_equal(y.T, g["stream"])


def test_granular_render_256_streams(env):
    """BASELINE cfg4 as specified: 256 stereo streams, each with its own frame-size schedule (64..8192, offset by the
    stream index), in ONE launch; every stream against the oracle's render of it, bit for bit."""
    tsb, orc = env["tsb"], env["orc"]
    streams, samples = 256, 40000
    xs = [orc.gen_signal(st, 2, samples) for st in range(streams)]
    scheds = [tsb.granular_schedule(samples, _granular_cb(st, True), n_first=64) for st in range(streams)]
    x = np.concatenate([v.T for v in xs], axis=0)
    ctx = env["ctx"]
    l0 = ctx.launches
    y = ctx.granular_render(dev(env, x), 2, scheds).cpu().numpy()
    assert ctx.launches - l0 == 1
    for st in range(0, streams, 5):
        r = orc.render_granular(xs[st], n_first=64, vary=True, stream_idx=st)
        L = r["stream"].shape[0]
        assert L == int(scheds[st]["out_off"][-1] + scheds[st]["hop"][-1])
        assert np.array_equal(y[2 * st:2 * st + 2, :L].T, r["stream"]), st
        assert not y[2 * st:2 * st + 2, L:].any()
