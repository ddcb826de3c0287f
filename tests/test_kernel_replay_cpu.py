"""CPU replay of the CUDA chain kernel's per-thread phases (tests/emul, compiled from the same headers as
the kernel) against the oracle: checks the kernel's arithmetic and indexing without a GPU.  The replay is
test infrastructure, not a fallback — nothing in the product loads it."""
import os
import sys

import numpy as np
import pytest

from oracle import orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul  # noqa: E402


def maxrel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def rmsrel(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2) / np.mean(b ** 2)))


@pytest.mark.parametrize("lg", range(3, 13))
def test_stockham_fft(lg):
    import ctypes as C
    m = 1 << lg
    rng = np.random.default_rng(lg)
    x = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    want = np.fft.fft(x)
    for prec, tol in ((1, 1e-15), (0, 4e-7)):
        y = np.empty(m, np.complex128)
        assert emul.lib().emul_fft(prec, lg, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)) == 0
        assert maxrel(y, want) < tol


@pytest.mark.parametrize("n,hr,shift,ch,samples,stream", [
    (1024, 0.25, 2.0, 2, 14000, 0), (256, 0.25, 1.5, 2, 8000, 0), (64, 0.25, 0.5, 1, 5000, 5),
    (4096, 0.25, 1.5, 1, 24000, 0), (8192, 0.25, 1.5, 1, 36000, 2), (512, 0.5, 1.2, 1, 9000, 0),
    # used-bin bookkeeping: halving (two sources per destination, silent upper destinations, reordered pairs), x3
    (1024, 0.25, 0.5, 1, 14000, 2), (1024, 0.25, 3.0, 1, 14000, 3), (512, 0.25, 4.0, 1, 9000, 4)])
def test_pv_chain(n, hr, shift, ch, samples, stream):
    x = orc.gen_signal(stream, ch, samples)
    ref = orc.render_pv(x, n, hr, shift)
    F, hop = ref["frames"], int(n * hr)
    want = ref["ola64"][:F * hop].T
    out, st = emul.chain_render(x.T, n, hop, F, precision="f64", ratio=shift)
    assert maxrel(out, want) < 1e-11
    dps = np.abs(st["phase_sum"] - ref["phase_sum"][:, :n // 2])
    assert dps.max() < 1e-8          # same wrap side in every bin of every frame
    out32, st32 = emul.chain_render(x.T, n, hop, F, precision="f32", ratio=shift)
    assert rmsrel(out32, want) < 1e-5
    assert (np.abs(st32["phase_sum"] - ref["phase_sum"][:, :n // 2]) > 1).sum() == 0


def test_fp32_doubtful_wraps_are_decided_in_double():
    """fp32 mode, non-integer ratio: a wrap of phase_to_frequency (rtlib/pv.hpp:41-43) that the float32 transform puts
    within its own phase error of the +-pi boundary is re-evaluated in double precision from the input samples
    (Chain::phase_wrap_fix).  Over 60 s of cfg1 at x1.5 the float decision is wrong a few times — each one used to be
    a permanent sign flip of that bin — and every one is caught: the phase sums follow the fp64 render."""
    n, hop, ratio = 1024, 256, 1.5
    x = orc.gen_signal(0, 2, 60 * 48000).T
    F = (x.shape[1] - 2048 - n) // hop + 1
    out64, st64 = emul.chain_render(x, n, hop, F, precision="f64", ratio=ratio)
    emul.wrap_counters()
    out32, st32 = emul.chain_render(x, n, hop, F, precision="f32", ratio=ratio)
    fixed, changed = emul.wrap_counters()
    assert fixed > 50 and changed >= 1                      # the slow path ran, and it mattered
    assert fixed < 0.05 * F * 2                             # ... rarely: a few per hundred frames
    dps = np.abs(st32["phase_sum"] - st64["phase_sum"])
    assert (dps > 1).sum() == 0 and dps.max() < 1e-3         # no flips: the sums differ by float phase noise only
    assert rmsrel(out32, out64) < 1e-5
    # a continued render decides the same way when it is told where the previous frame lies (ChainArgs::prev_hop)
    a = F // 3
    o1, st = emul.chain_render(x, n, hop, a, precision="f32", ratio=ratio)
    o2, st = emul.chain_render(x, n, hop, F - a, precision="f32", ratio=ratio, in_off0=2048 + a * hop, state=st,
                               prev_hop=hop)
    assert np.array_equal(np.concatenate([o1, o2], axis=1), out32)
    assert np.array_equal(st["phase_sum"], st32["phase_sum"])


def test_pv_chain_fractional_hop_and_segments():
    x = orc.gen_signal(3, 1, 12000)
    n, hop_d, shift = 512, 512 * 0.2003, 0.75
    ref = orc.render_pv(x, n, 0.2003, shift, want_tables=True)
    tabs, F = ref["tables"], ref["frames"]
    out, _ = emul.chain_render(x.T, n, 0, F, precision="f64", ratio=shift, hop_d=hop_d, hop_tab=tabs["hop"])
    assert maxrel(out, ref["ola64"][:out.shape[1]].T) < 1e-11
    # the same render as two launches handing the state over: bit-identical
    a = F // 3
    o2 = np.zeros_like(out)
    o1, st = emul.chain_render(x.T, n, 0, a, precision="f64", ratio=shift, hop_d=hop_d, hop_tab=tabs["hop"][:a],
                               out_len=out.shape[1])
    o2, st = emul.chain_render(x.T, n, 0, F - a, precision="f64", ratio=shift, hop_d=hop_d, hop_tab=tabs["hop"][a:],
                               in_off0=int(tabs["in_off"][a]), out_off0=int(tabs["out_off"][a]), state=st,
                               out_len=out.shape[1])
    p = int(tabs["out_off"][a])
    assert np.array_equal(np.concatenate([o1[:, :p], o2[:, p:]], axis=1), out)


@pytest.mark.parametrize("n,filt", [(2048, True), (64, True), (512, False)])
def test_gain_chain(n, filt):
    x = orc.gen_signal(0, 2, 2048 + 10 * n)
    gain = 1.0 / (1.0 + np.arange(n // 2 + 1) / 64.0) if filt else None
    ref = orc.render_filter(x, n, 0.25, gain)
    F = ref["frames"]
    out, _ = emul.chain_render(x.T, n, n // 4, F, precision="f64", program="gain", gain=gain)
    assert maxrel(out, ref["ola64"][:out.shape[1]].T) < 1e-13
    out, _ = emul.chain_render(x.T, n, n // 4, F, precision="f32", program="gain", gain=gain)
    assert rmsrel(out, ref["ola64"][:out.shape[1]].T) < 1e-5


@pytest.mark.parametrize("n,hr,shift,samples", [(64, 1.6, 1.5, 9000), (128, 1.0, 2.0, 9000), (256, 0.3, 1.5, 9000),
                                                (64, 0.01, 1.2, 2600)])
def test_pv_chain_hop_edge_cases(n, hr, shift, samples):
    """hop larger than the frame (tinyspec.cpp:184-191), hop equal to the frame, fractional hops, hops of 0/1."""
    x = orc.gen_signal(1, 2, samples)
    ref = orc.render_pv(x, n, hr, shift, want_tables=True)
    tabs, F = ref["tables"], ref["frames"]
    const = bool(np.all(tabs["hop"] == tabs["hop"][0]))
    out, _ = emul.chain_render(x.T, n, int(tabs["hop"][0]), F, precision="f64", ratio=shift, hop_d=n * hr,
                               hop_tab=None if const else tabs["hop"])
    assert maxrel(out, ref["ola64"][:out.shape[1]].T) < 1e-11


@pytest.mark.parametrize("n,hr,shift,ch,samples", [(1024, 0.25, 2.0, 2, 30000), (256, 0.3, 1.5, 2, 9000),
                                                   (64, 1.6, 1.5, 1, 9000), (512, 0.2003, 0.75, 1, 12000)])
def test_pv_chain_float32_output_is_the_reference_stream(n, hr, shift, ch, samples):
    """TSB_F64_OUT32: fp64 arithmetic, float32 tail and output combined as tinyspec.cpp:173-200 does.  The result is
    the reference's own float32 stream (what it hands to JACK): identical except where the 1e-10 arithmetic
    difference straddles a float32 rounding boundary, and then off by one float32 step of the running tail."""
    x = orc.gen_signal(2, ch, samples)
    ref = orc.render_pv(x, n, hr, shift, want_tables=True)
    tabs, F = ref["tables"], ref["frames"]
    const = bool(np.all(tabs["hop"] == tabs["hop"][0]))
    out, st = emul.chain_render(x.T, n, int(tabs["hop"][0]), F, precision="f64o32", ratio=shift, hop_d=n * hr,
                                hop_tab=None if const else tabs["hop"])
    want = ref["stream"][:out.shape[1]].T
    assert out.dtype == np.float32 and want.dtype == np.float32
    differ = np.count_nonzero(out != want)
    assert differ <= 1e-3 * out.size
    assert np.abs(out.astype(np.float64) - want).max() <= 2.0 ** -23 * np.abs(want).max()
    # the tail handed on is float32-valued
    assert np.array_equal(st["ola"], st["ola"].astype(np.float32).astype(np.float64))


def test_gain_chain_float32_output_is_the_reference_stream():
    x = orc.gen_signal(0, 2, 2048 + 12 * 2048)
    gain = 1.0 / (1.0 + np.arange(1025) / 64.0)
    ref = orc.render_filter(x, 2048, 0.25, gain)
    out, _ = emul.chain_render(x.T, 2048, 512, ref["frames"], precision="f64o32", program="gain", gain=gain)
    want = ref["stream"][:out.shape[1]].T
    assert np.count_nonzero(out != want) <= 1e-3 * out.size
    assert np.abs(out.astype(np.float64) - want).max() <= 2.0 ** -23 * np.abs(want).max()


def test_fp64_pv_math_kernels():
    """The branch-free replacements of atan2 / sqrt / (int)(x/M_PI) / sincos in the fp64 chain, against libm."""
    rng = np.random.default_rng(5)
    n = 400000
    mag = 10.0 ** rng.integers(-6, 6, n)
    x = (rng.random(n) - 0.5) * mag
    y = (rng.random(n) - 0.5) * 10.0 ** rng.integers(-6, 6, n)
    y[::1000] = x[::1000] * (np.arange(n)[::1000] // 1000 % 65) / 64.0      # on the table points
    at, sq, wr, sn, cs = emul.pv_math(y, x)
    want = np.arctan2(y, x)
    ulp = np.spacing(np.abs(want))
    assert np.max(np.abs(at - want) / ulp) <= 2.0
    assert np.array_equal(sq, np.sqrt(np.abs(x)))                          # correctly rounded
    # the wrap needs the reference's exact truncated quotient, also on and next to multiples of pi
    d = np.concatenate([x, np.pi * np.arange(-300, 300), np.nextafter(np.pi * np.arange(-300, 300), 1e9),
                        np.nextafter(np.pi * np.arange(-300, 300), -1e9)])
    _, _, wr, _, _ = emul.pv_math(d, d)
    assert np.array_equal(wr, np.trunc(d / np.pi).astype(np.int32))
    # sincos of unreduced phases up to 1e9 rad (a one-hour render reaches 8e8): absolute error ~1 ulp of 1
    ph = (rng.random(n) - 0.5) * 2.0 * 10.0 ** rng.integers(-3, 10, n)
    _, _, _, sn, cs = emul.pv_math(ph, ph)
    ld = ph.astype(np.longdouble)
    assert np.max(np.abs(sn - np.sin(ld).astype(np.float64))) < 2.3e-16
    assert np.max(np.abs(cs - np.cos(ld).astype(np.float64))) < 2.3e-16
    # signs and zeros as libm has them
    yy = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, -0.0])
    xx = np.array([-0.0, -0.0, 0.0, 0.0, -0.0, 0.0, -2.0, -2.0])
    at, sq, _, _, _ = emul.pv_math(yy, xx)
    assert np.array_equal(at, np.arctan2(yy, xx)) and np.array_equal(np.signbit(at), np.signbit(np.arctan2(yy, xx)))
    assert sq[0] == 0.0


def _oracle_render_with_map(x, n, hop, fn_bins, ratio, rate=48000.0):
    """Frame by frame with the oracle's PhaseVocoder steps and an arbitrary bin map (what `fn_bins` + `ratio` of a
    fused plan mean): tinyspec.cpp:135-200 schedule (constant integer hop), float64 overlap-add."""
    L, ch = x.shape
    F = orc.n_frames(L, n, hop)
    pv = orc.PV(float(hop), rate)
    out = np.zeros((ch, F * hop + n))
    for f in range(F):
        o = 2048 + f * hop
        pv.analyze(np.ascontiguousarray(x[o:o + n].T, np.float64))
        pv.shift_map(fn_bins, ratio)
        out[:, f * hop:f * hop + n] += pv.synthesize()
    return out[:, :F * hop], F, pv


@pytest.mark.parametrize("kind", ["reverse", "random", "sparse"])
def test_pv_chain_arbitrary_bin_map(kind):
    """Bin maps a linear shift never produces — a permutation, random collisions and drops, a few survivors — through
    the plan's used-source flags, destination order and multi-source gather."""
    n, hop, ch = 256, 64, 1
    half, rate = n // 2, 48000.0
    centre = np.arange(half) / n * rate
    rng = np.random.default_rng(11)
    if kind == "reverse":
        fb, ratio = centre[::-1].copy() + 1.0, 1.0
    elif kind == "random":
        fb, ratio = rng.integers(-20, half + 40, half) / n * rate + 1.0, 1.3
    else:
        fb = np.full(half, 1e9)
        keep = rng.choice(half, 9, replace=False)
        fb[keep] = centre[rng.choice(half, 9, replace=False)] + 1.0
        ratio = 0.8
    x = orc.gen_signal(7, ch, 2048 + n + 40 * hop)
    want, F, pv = _oracle_render_with_map(x, n, hop, fb, ratio)
    out, st = emul.chain_render(x.T, n, hop, F, precision="f64", ratio=ratio, fn_bins=fb)
    assert maxrel(out, want) < 1e-11
    assert np.abs(st["phase_sum"] - pv.phase_sum[:, :half]).max() < 1e-8
    assert np.abs(st["last_phase"] - pv.last_phase[:, :half]).max() < 1e-11
    out32, _ = emul.chain_render(x.T, n, hop, F, precision="f32", ratio=ratio, fn_bins=fb)
    assert rmsrel(out32, want) < 1e-5


@pytest.mark.parametrize("precision", ["f64", "f64o32", "f32"])
def test_time_segment_with_lead_in_frames_is_the_sequential_render(precision):
    """A time segment that replays, silently, the frames whose output overlaps its first frame starts with the very
    overlap tail (and, through the pre-pass and the frame-by-frame scan, the very phase state) the sequential
    render has there: its samples are bit-identical, float32 tail arithmetic included."""
    n, hop, ratio, ch = 256, 64, 1.5, 2
    x = orc.gen_signal(3, ch, 2048 + n + 47 * hop).T
    F = orc.n_frames(x.shape[1], n, hop)
    seq, st_seq = emul.chain_render(x, n, hop, F, precision=precision, ratio=ratio)
    first, lead = 20, n // hop - 1              # segment 1 emits frames [20, F); frames 17..19 are its lead-in
    m = n // 2
    zeros = lambda: dict(last_phase=np.zeros((ch, m)), phase_sum=np.zeros((ch, m)), ola=np.zeros((ch, n)))
    f0 = first - lead
    # pre-pass of both segments: per-frame increments (segment 1: warm-up frame 16, lead-in 17..19 analysed only)
    wp0, wp1 = np.zeros((ch, m)), np.zeros((ch, m))
    inc0 = np.zeros((ch, first, m))              # segment 0's own frames 0..19 (contiguous [lanes][frames][bins] blocks)
    emul.segment_render(x, n, hop, ratio, 2048, first, 0, False, 1, zeros(), wp0, inc0, precision)
    inc1 = np.zeros((ch, F - f0, m))
    emul.segment_render(x, n, hop, ratio, 2048 + f0 * hop, F - f0, lead, True, 1, zeros(), wp1, inc1, precision)
    # a segment that starts its block (first_frame 0) stores its lead-in frames' increments too — the block before
    # holds the same values for the same frames, bit for bit (what a time-sharded render relies on)
    assert np.array_equal(inc1[:, :lead], inc0[:, f0:first])
    inc = np.concatenate([inc0, inc1[:, lead:]], axis=1)
    # scan: the sequential additions, noted where segment 1 starts processing
    ps = np.zeros((ch, m))
    for f in range(f0):
        ps = ps + inc[:, f]
    # main pass of segment 1 from (last_phase after frame 16, phase_sum before frame 17, empty tail)
    st1 = zeros()
    st1["phase_sum"][:] = ps
    out1 = emul.segment_render(x, n, hop, ratio, 2048 + f0 * hop, F - f0, lead, True, 0, st1, wp1, None, precision,
                               out_len=(F - f0) * hop)
    assert np.all(out1[:, :lead * hop] == 0)     # nothing is emitted for the lead-in frames
    assert np.array_equal(out1[:, lead * hop:], seq[:, first * hop:])
    for k in ("last_phase", "phase_sum", "ola"):
        assert np.array_equal(st1[k], st_seq[k])


@pytest.mark.parametrize("precision", ["f64", "f64o32", "f32"])
@pytest.mark.parametrize("n,program", [(1024, "pv"), (512, "pv"), (2048, "pv"), (2048, "gain"), (4096, "pv")])
def test_register_overlap_tail_is_the_ring_path(precision, n, program):
    """At a constant hop of N/4 some kernels keep the overlap tail in registers (Chain::RT_OK, tsb_chain.cuh: N=1024
    everywhere, 512 and 2048 in fp64, 2048 gain in fp32): same samples, same state as the shared-memory ring path
    (forced here by a hop table of the same hops), across a state hand-over and with a row that ends mid-frame."""
    rng = np.random.default_rng(3)
    hop, F = n // 4, 21
    x = (rng.random((2, 2048 + F * hop + n), dtype=np.float32) - 0.5)
    a, sa = emul.chain_render(x, n, hop, F, precision=precision, program=program, ratio=1.5)
    b, sb = emul.chain_render(x, n, hop, F, precision=precision, program=program, ratio=1.5, hop_tab=np.full(F, hop, np.int32))
    assert np.abs(a).max() > 0.1
    assert np.array_equal(a, b)
    for k in ("ola", "phase_sum", "last_phase"):
        assert np.array_equal(sa[k], sb[k])
    a1, s1 = emul.chain_render(x, n, hop, 9, precision=precision, program=program, ratio=1.5)
    a2, _ = emul.chain_render(x, n, hop, F - 9, precision=precision, program=program, ratio=1.5, in_off0=2048 + 9 * hop,
                              state=s1, prev_hop=hop)
    assert np.array_equal(np.concatenate([a1, a2], axis=1), a)
    c, _ = emul.chain_render(x, n, hop, F, precision=precision, program=program, ratio=1.5, out_len=F * hop - 301)
    assert np.array_equal(c, a[:, :F * hop - 301])


def test_pv_chain_with_its_own_hop():
    """PhaseVocoder::set_hop (rtinc/pv.hpp:46-53): the hop inside phase_to_frequency / frequency_to_phase is the
    object's own, not the scheduler's.  The fused plan takes it as `hop_d`; frames still advance by the scheduler's."""
    n, hs, ch, pv_hop, ratio = 1024, 256, 2, 200.0, 1.5
    x = orc.gen_signal(0, ch, 2048 + 30 * hs + n)
    F = orc.n_frames(x.shape[0], n, hs)
    pv = orc.PV(pv_hop)
    acc = np.zeros((ch, F * hs + n))
    for f in range(F):
        fr = np.ascontiguousarray(x[2048 + f * hs:2048 + f * hs + n].T.astype(np.float64))
        pv.analyze(fr)
        pv.shift(ratio)
        acc[:, f * hs:f * hs + n] += pv.synthesize()
    out, _ = emul.chain_render(np.ascontiguousarray(x.T), n, hs, F, precision="f64", ratio=ratio, hop_d=pv_hop)
    assert np.abs(out[:, :F * hs] - acc[:, :F * hs]).max() <= 1e-9 * np.abs(acc).max()
    wrong, _ = emul.chain_render(np.ascontiguousarray(x.T), n, hs, F, precision="f64", ratio=ratio)
    assert np.abs(wrong[:, :F * hs] - acc[:, :F * hs]).max() > 0.1 * np.abs(acc).max()  # the scheduler's hop is another PV


def test_radix16_passes_of_the_spectral_kernels():
    """tsb_math.cuh r16_pass: 256 points = two radix-16 passes by sixteen threads over the r16_sw tile layout (what
    kernels A and B of tsb_spectral.cu run for cfg5's 256 x 256 four-step transform), against numpy's FFT; the
    half-output variant of kernel B must deliver X[0..128] and leave the rest alone; and the layout must be free of
    shared-memory bank conflicts for every access of both passes (16-byte elements: 8 lanes per wavefront)."""
    import ctypes as C
    rng = np.random.default_rng(16)
    x = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    want = np.fft.fft(x)
    y = np.empty(256, np.complex128)
    fn = emul.lib().emul_fft256_r16
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    fn.restype = None
    fn(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 0)
    assert maxrel(y, want) < 2e-15
    fn(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), 1)
    assert maxrel(y[:129], want[:129]) < 2e-15
    assert np.all(y[129:] == 0)
    sw = lambda i: i ^ ((i >> 4) & 7)
    for t in range(16):  # lanes j = 8 q .. 8 q + 7 form one wavefront
        for q in range(2):
            lanes = range(8 * q, 8 * q + 8)
            assert len({sw(j + 16 * t) % 8 for j in lanes}) == 8   # reads of both passes, writes of pass 1
            assert len({sw(16 * j + t) % 8 for j in lanes}) == 8   # writes of pass 0
