"""GPU parity: the CUDA chain (through the C ABI, tinyspec-cling_b200/api.py) against the C oracle on
the same seeded inputs, and against the golden fixtures of the compiled reference.

Tolerances (BASELINE.json north_star): fp64 mode <= 1e-9 max error relative to the signal peak;
fp32 mode <= 1e-5 RMS-relative (>= 100 dB SNR)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F64_MAXREL = 1e-9
F32_RMSREL = 1e-5


@pytest.fixture(scope="module")
def env():
    import torch
    import tsb200
    from oracle import orc
    ctx = tsb200.Context(0)
    return dict(torch=torch, tsb=tsb200, orc=orc, ctx=ctx)


def maxrel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def rmsrel(a, b):
    return float(np.sqrt(np.mean((a - b) ** 2) / np.mean(b ** 2)))


def run_pv(env, x, n, hop_ratio, shift, precision, host=False, **kw):
    """x: interleaved float32 [L, ch] (the oracle's layout). Returns (out [ch, P], frames, state)."""
    torch, tsb = env["torch"], env["tsb"]
    L, ch = x.shape
    sch = tsb.schedule_const(L, n, n * hop_ratio)
    F = sch["n_frames"]
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=n * hop_ratio, ratio=shift)
    xp = np.ascontiguousarray(x.T)
    const = bool(np.all(sch["hop"] == sch["hop"][0]))
    hop_tab = None if const else sch["hop"]
    if host:
        st = dict(last_phase=np.zeros((ch, n // 2)), phase_sum=np.zeros((ch, n // 2)), ola=np.zeros((ch, n)))
        out = plan.render_host(xp, F, int(sch["hop"][0]), in_off0=int(sch["in_off"][0]), hop_tab=hop_tab, state=st)
        return out.numpy(), F, st
    xd = torch.from_numpy(xp).cuda()
    st = dict(last_phase=torch.zeros((ch, n // 2), dtype=torch.float64, device="cuda"),
              phase_sum=torch.zeros((ch, n // 2), dtype=torch.float64, device="cuda"),
              ola=torch.zeros((ch, n), dtype=torch.float64, device="cuda"))
    out = plan.render(xd, F, int(sch["hop"][0]), in_off0=int(sch["in_off"][0]), hop_tab=hop_tab, state=st, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), F, {k: v.cpu().numpy() for k, v in st.items()}


CASES = [  # n, hop_ratio, shift, channels, samples, stream
    (1024, 0.25, 2.0, 2, 24000, 0),     # cfg1 shape (hacks/pitchshift-pv.cpp x2)
    (1024, 0.25, 1.5, 2, 48000, 1),
    (4096, 0.25, 1.5, 8, 30000, 0),     # cfg3 shape
    (256, 0.25, 1.5, 2, 12000, 0),
    (64, 0.25, 0.5, 1, 6000, 5),        # smallest frame, downward shift (several sources per bin)
    (128, 0.5, 1.2, 3, 6000, 2),
    (512, 0.2003, 0.75, 1, 16000, 3),   # fractional hop: computed_hop alternates 102/103
    (2048, 0.125, 2.0, 2, 30000, 0),
    (8192, 0.25, 1.5, 1, 60000, 2),     # largest fused frame
    (64, 1.6, 1.5, 2, 9000, 1),         # hop larger than the frame (tinyspec.cpp:184-191), fractional
    (128, 1.0, 2.0, 2, 9000, 1),        # hop equal to the frame: no overlap at all
    (256, 0.3, 1.5, 2, 9000, 1),        # odd hops 76/77
    # shift maps that exercise the used-bin bookkeeping: halving (two sources per destination, the upper half of
    # the destinations silent and reordered), x3 and x4 (most sources dropped, most destinations silent)
    (1024, 0.25, 0.5, 2, 30000, 2),
    (1024, 0.25, 3.0, 2, 30000, 3),
    (512, 0.25, 4.0, 1, 20000, 4),
    (2048, 0.25, 2.0, 2, 40000, 0),
    (4096, 0.25, 0.5, 1, 40000, 1),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_r{c[1]}_x{c[2]}_c{c[3]}")
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_pv_vs_oracle(env, case, precision):
    n, hr, shift, ch, samples, stream = case
    orc = env["orc"]
    x = orc.gen_signal(stream, ch, samples)
    ref = orc.render_pv(x, n, hr, shift)
    out, F, st = run_pv(env, x, n, hr, shift, precision)
    assert F == ref["frames"]
    want = ref["ola64"][:out.shape[1]].T
    if precision == "f64":
        assert maxrel(out, want) <= F64_MAXREL
        # PV state: same wrap side everywhere (no 2*pi*ratio jumps), tiny numeric difference
        dps = np.abs(st["phase_sum"] - ref["phase_sum"][:, :n // 2])
        assert dps.max() < 1e-6, f"{int((dps > 1).sum())} wrap flips"
        assert np.abs(st["last_phase"] - ref["last_phase"][:, :n // 2]).max() < 1e-9
        # narrowed to float it reproduces the reference's float32 output stream to float rounding
        P = ref["stream"].shape[0]
        assert np.abs(out[:, :P].T.astype(np.float32) - ref["stream"]).max() <= 2.4e-7 * np.abs(want).max() + 1e-7
    else:
        assert rmsrel(out, want) <= F32_RMSREL
        dps = np.abs(st["phase_sum"] - ref["phase_sum"][:, :n // 2])
        assert (dps > 1).sum() == 0, f"{int((dps > 1).sum())} wrap flips"


@pytest.mark.parametrize("name,ratio", [("pv_cfg1_60s_tail", 2.0), ("pv_cfg3_shape", 1.5)])
def test_full_size_render_vs_reference_binary(env, golden, name, ratio):
    """BASELINE configs[0] at full length (60 s, 11,239 frames) and cfg3's 8 x 4096 shape, against the tail of the
    float32 stream and the final PhaseVocoder state the compiled, unmodified reference produced — with the mode
    that delivers the reference's own stream type (fp64 arithmetic, float32 tail and narrowing)."""
    g = golden(name)
    n, ch, samples, stream = (int(v) for v in g["meta"])
    x = env["orc"].gen_signal(stream, ch, samples)
    out, F, st = run_pv(env, x, n, 0.25, ratio, "f64o32")
    P = int(g["stream_len"][0]) // ch
    assert F * (n // 4) == P
    tail, want = out[:, P - 8192:P].T, g["stream"]
    assert (tail == want).mean() > 0.999
    assert np.abs(tail.astype(np.float64) - want).max() <= 2.0 ** -23 * np.abs(want).max()
    if int(g["frame"][-1]) == F - 1:
        m = n // 2
        assert np.abs(st["phase_sum"] - g["phase_sum"][-1][:, :m]).max() < 1e-6   # of up to 1e7 rad
        assert np.abs(st["last_phase"] - g["last_phase"][-1][:, :m]).max() < 1e-9


@pytest.mark.parametrize("name", ["pv_cfg1_short", "pv_x1p5_n256", "pv_down_n512_fracthop"])
def test_pv_vs_golden(env, golden, name):
    """Directly against what the compiled, unmodified reference produced."""
    g = golden(name)
    n, ch, samples, stream = (int(v) for v in g["meta"])
    hr, shift = (float(v) for v in g["params"])
    x = env["orc"].gen_signal(stream, ch, samples)
    out, F, st = run_pv(env, x, n, hr, shift, "f64")
    P = g["stream"].shape[0]
    peak = np.abs(g["stream"]).max()
    assert np.abs(out[:, :P].T.astype(np.float32) - g["stream"]).max() <= 2.4e-7 * peak + 1e-7
    last = int(g["frame"][-1])
    if last == F - 1:
        assert np.abs(st["phase_sum"] - g["phase_sum"][-1][:, :n // 2]).max() < 1e-6


def test_pv_segments_bit_identical(env):
    """Splitting a render into launches that hand the state over gives bit-identical output (what
    time-segment sharding and teacher-forced checks rely on)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop, ch = 1024, 256, 2
    x = orc.gen_signal(0, ch, 40000)
    sch = tsb.schedule_const(x.shape[0], n, hop)
    F = sch["n_frames"]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    # x2 and x0.5: bins the shift does not use are analysed only on the last frame of each launch, and the state
    # handed over must still be the reference's
    for precision, ratio in (("f64", 1.5), ("f32", 1.5), ("f64", 2.0), ("f32", 2.0), ("f64", 0.5)):
        plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=hop, ratio=ratio)
        whole = plan.render(xd, F, hop).clone()
        st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
              (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
        out = torch.zeros_like(whole)
        f0 = 0
        for seg in (1, 7, 40, F - 48):
            plan.render(xd, seg, hop, out=out, in_off0=int(sch["in_off"][f0]), out_off0=int(sch["out_off"][f0]),
                        state=st, load_state=f0 > 0)
            f0 += seg
        torch.cuda.synchronize()
        assert f0 == F
        assert torch.equal(out, whole)
        again = plan.render(xd, F, hop)
        torch.cuda.synchronize()
        assert torch.equal(again, whole)  # deterministic: no atomics anywhere


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f64", "f64o32", "f32"])
@pytest.mark.parametrize("n,program", [(1024, "pv"), (1024, "gain"), (512, "pv"), (2048, "pv"), (2048, "gain"), (4096, "pv")])
def test_register_overlap_tail_is_the_ring_path(env, precision, n, program):
    """At a constant hop of N/4 some kernels keep the overlap tail in registers (Chain::RT_OK: N=1024 everywhere, 512 and
    2048 in fp64, 2048 gain in fp32): bit-identical to the shared-memory ring path, which a hop table of the same hops
    selects; also with a row that ends mid-frame."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    hop, ch = n // 4, 3
    x = orc.gen_signal(0, ch, 30000)
    sch = tsb.schedule_const(x.shape[0], n, hop)
    F = sch["n_frames"]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    plan = tsb.Plan(env["ctx"], n, precision=precision, program=program, hop=hop, ratio=1.5)
    a = plan.render(xd, F, hop).clone()
    b = plan.render(xd, F, hop, hop_tab=np.full(F, hop, np.int32)).clone()
    torch.cuda.synchronize()
    assert float(a.abs().max()) > 0.1
    assert torch.equal(a, b)
    short = torch.zeros((ch, F * hop - 301), dtype=a.dtype, device="cuda")
    plan.render(xd, F, hop, out=short)
    torch.cuda.synchronize()
    assert torch.equal(short, a[:, :F * hop - 301])


@pytest.mark.parametrize("n,ch,filt", [(2048, 1, False), (2048, 1, True), (512, 2, True), (64, 2, True), (8192, 1, True)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gain_vs_oracle(env, n, ch, filt, precision):
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    x = orc.gen_signal(0, ch, 2048 + 12 * n)
    gain = 1.0 / (1.0 + np.arange(n // 2 + 1) / 64.0) if filt else None
    ref = orc.render_filter(x, n, 0.25, gain)
    F = ref["frames"]
    assert F >= 8
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="gain", gain=gain)
    out = plan.render(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), F, n // 4)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    want = ref["ola64"][:out.shape[1]].T
    if precision == "f64":
        assert maxrel(out, want) <= 1e-12
        if not filt:  # COLA known answer: y[p] = 2*x[p+2048] in the steady region
            p = np.arange(n - n // 4, (F - 1) * (n // 4))
            assert np.abs(out[:, p] - 2.0 * x[p + 2048].T).max() < 1e-12
    else:
        assert rmsrel(out, want) <= F32_RMSREL


@pytest.mark.parametrize("name", ["filter_pass_n2048", "filter_lp_n2048", "filter_lp_n512_stereo"])
def test_gain_vs_golden(env, golden, name):
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    g = golden(name)
    n, ch, samples, filt = (int(v) for v in g["meta"])
    hr = float(g["params"][0])
    x = orc.gen_signal(0, ch, samples)
    gain = 1.0 / (1.0 + np.arange(n // 2 + 1) / 64.0) if filt else None
    sch = tsb.schedule_const(samples, n, n * hr)
    plan = tsb.Plan(env["ctx"], n, precision="f64", program="gain", gain=gain)
    out = plan.render(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), sch["n_frames"], int(n * hr))
    torch.cuda.synchronize()
    got = out.cpu().numpy().T.astype(np.float32)
    P = g["stream"].shape[0]
    assert np.abs(got[:P] - g["stream"]).max() <= 2.4e-7 * np.abs(g["stream"]).max() + 1e-7


def test_host_path_matches_device_path(env):
    orc = env["orc"]
    x = orc.gen_signal(0, 2, 30000)
    a, F, sa = run_pv(env, x, 1024, 0.25, 2.0, "f64", host=False)
    b, F2, sb = run_pv(env, x, 1024, 0.25, 2.0, "f64", host=True)
    assert F == F2 and np.array_equal(a, b)
    assert np.array_equal(sa["phase_sum"], sb["phase_sum"])


@pytest.mark.parametrize("n,hr,samples,ch", [(256, 0.25, 2200000, 40), (512, 0.2003, 900000, 150)])
def test_host_path_pipelined_chunks_bit_identical(env, n, hr, samples, ch):
    """Large enough (> 1 GB of rows) for several pipelined lane chunks, constant and fractional hop: same bits
    as one launch over device-resident buffers."""
    x = np.random.default_rng(3).uniform(-0.5, 0.5, (samples, ch)).astype(np.float32)
    a, F, sa = run_pv(env, x, n, hr, 1.5, "f32", host=False)
    b, F2, sb = run_pv(env, x, n, hr, 1.5, "f32", host=True)
    assert F == F2 and F > 1500
    assert np.array_equal(a, b)
    for k in ("last_phase", "phase_sum", "ola"):
        assert np.array_equal(sa[k], sb[k])


def test_many_lanes_match_single_lane(env):
    """Lanes are independent: a batch of 300 streams gives, lane by lane, what each gives alone."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop = 1024, 256
    xs = [orc.gen_signal(s, 1, 8000)[:, 0] for s in range(6)]
    X = np.stack([xs[i % 6] for i in range(300)])
    sch = tsb.schedule_const(8000, n, hop)
    plan = tsb.Plan(env["ctx"], n, precision="f32", program="pv", hop=hop, ratio=2.0)
    out = plan.render(torch.from_numpy(X).cuda(), sch["n_frames"], hop)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for i in range(6, 300):
        assert np.array_equal(out[i], out[i % 6])
    ref = orc.render_pv(xs[3][:, None], n, 0.25, 2.0)
    assert rmsrel(out[3], ref["ola64"][:out.shape[1], 0]) <= F32_RMSREL


def test_errors_are_loud(env):
    tsb = env["tsb"]
    with pytest.raises(tsb.TsbError):
        tsb.Plan(env["ctx"], 1000, program="pv", hop=250)      # not a power of two
    with pytest.raises(tsb.TsbError):
        tsb.Plan(env["ctx"], 16384, program="pv", hop=4096)    # beyond the fused range
    with pytest.raises(tsb.TsbError):
        tsb.hop_fixed(-1.0)                                    # rtlib/control.hpp:23


# ---------------------------------------------------------------- time-parallel rendering (tsb_timepar_*)
@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("n,hr,shift,ch,samples,nseg", [(1024, 0.25, 2.0, 2, 96000, 37), (1024, 0.25, 1.5, 1, 480000, 444),
                                                        (4096, 0.25, 1.5, 8, 120000, 5), (512, 0.2003, 0.75, 1, 40000, 16)])
def test_timepar_matches_sequential_and_oracle(env, precision, n, hr, shift, ch, samples, nseg):
    """Segments of one stream rendered concurrently give the sequential render bit for bit — samples, PV state and
    overlap tail, both precisions: the scan adds the stored increments frame by frame like the sequential render
    does, and every segment replays silently the frames whose output overlaps its first one, which rebuilds the
    tail it starts with."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    x = orc.gen_signal(1, ch, samples)
    sch = tsb.schedule_const(samples, n, n * hr)
    F = sch["n_frames"]
    const = bool(np.all(sch["hop"] == sch["hop"][0]))
    hop_tab = None if const else sch["hop"]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=n * hr, ratio=shift)
    mk = lambda: {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
                  (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
    s_seq, s_par = mk(), mk()
    seq = plan.render(xd, F, int(sch["hop"][0]), hop_tab=hop_tab, state=s_seq).clone()
    par = plan.render_timepar(xd, F, int(sch["hop"][0]), nseg, hop_tab=hop_tab, state=s_par)
    torch.cuda.synchronize()
    seq, par = seq.cpu().numpy(), par.cpu().numpy()
    peak = np.abs(seq).max()
    assert torch.equal(s_par["phase_sum"], s_seq["phase_sum"])
    assert torch.equal(s_par["last_phase"], s_seq["last_phase"])
    assert np.array_equal(par, seq)
    assert torch.equal(s_par["ola"], s_seq["ola"])
    if precision == "f64" and samples <= 120000:
        ref = orc.render_pv(x, n, hr, shift)
        assert maxrel(par, ref["ola64"][:par.shape[1]].T) <= F64_MAXREL
    again = plan.render_timepar(xd, F, int(sch["hop"][0]), nseg, hop_tab=hop_tab)
    torch.cuda.synchronize()
    assert np.array_equal(again.cpu().numpy(), par)  # deterministic


def test_timepar_two_blocks_like_two_gpus(env):
    """The multi-GPU time split, run as two blocks on one device: block 1 starts mid-render with a warm-up frame
    from its input halo, the running phase_sum handed on by block 0's scan, and receives block 0's overlap tail."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop, ch, samples = 1024, 256, 2, 200000
    x = orc.gen_signal(0, ch, samples)
    sch = tsb.schedule_const(samples, n, hop)
    F = sch["n_frames"]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    plan = tsb.Plan(env["ctx"], n, precision="f64", program="pv", hop=hop, ratio=1.5)
    whole = plan.render(xd, F, hop).clone()
    zeros = lambda w: torch.zeros((ch, w), dtype=torch.float64, device="cuda")
    F0 = F // 2
    out = torch.zeros_like(whole)
    # block 0: frames [0, F0)
    st0 = dict(last_phase=zeros(n // 2), phase_sum=zeros(n // 2), ola=zeros(n))
    tp0 = plan.timepar(xd, F0, hop, 24, out=out, state=st0)
    tot0, _ = tp0.prepass(want_totals=True)
    # block 1: frames [F0, F): needs input from I_F0 - hop on (the halo)
    st1 = dict(last_phase=zeros(n // 2), phase_sum=zeros(n // 2), ola=zeros(n))
    tp1 = plan.timepar(xd, F - F0, hop, 24, out=out, in_off0=int(sch["in_off"][F0]), out_off0=int(sch["out_off"][F0]),
                       state=st1, load_state=True, warmup=True, warm_hop=hop)
    tp1.prepass()
    carry = tp0.scan()                       # running sum after block 0 ...
    assert torch.equal(carry, tot0)          # ... which from a zero start is also its total
    # ... is where block 1 starts; handed on lane by lane, as a chain of ranks does
    end = torch.cat([tp1.scan(carry[c:c + 1].contiguous(), c, 1) for c in range(ch)])
    tp1.render()
    tp0.render()
    plan.add_tail(out, int(sch["out_off"][F0]), st0["ola"])
    torch.cuda.synchronize()
    st = {k: zeros(w) for k, w in (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
    whole = plan.render(xd, F, hop, state=st).clone()
    assert torch.equal(end, st["phase_sum"]) and torch.equal(st1["phase_sum"], st["phase_sum"])
    d = np.abs(out.cpu().numpy() - whole.cpu().numpy())
    assert d.max() <= 4 * np.finfo(np.float64).eps * float(whole.abs().max())
    assert (d == 0).mean() > 1 - 1.01 * n / whole.shape[1]   # only block 0's tail is added afterwards


@pytest.mark.parametrize("n,hop_d,precision,world", [(1024, 256.0, "f64", 2), (256, 51.3, "f64", 3), (1024, 256.0, "f32", 3),
                                                     (4096, 1024.0, "f64o32", 2)])
def test_time_blocks_with_lead_in_are_the_whole_render(env, n, hop_d, precision, world):
    """The multi-GPU time split as shard.render_time_sharded runs it, with the ranks' blocks rendered one after the
    other on this device: a block starts with its lead-in frames (one warm-up frame before them), takes the running
    phase_sum the block before noted for it, and needs nothing else — the stitched output is the whole render bit for
    bit, in every precision and with a fractional hop (every block slices its own hop table)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    from tinyspec_cling_b200 import shard
    ch, samples = 2, 150000
    x = orc.gen_signal(1, ch, samples)
    sch = tsb.schedule_const(samples, n, hop_d)
    F = sch["n_frames"]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=hop_d, ratio=1.5)
    whole = plan.render(xd, F, int(sch["hop"][0]), hop_tab=sch["hop"]).clone()
    blocks = shard.plan_time_blocks(sch, n, world)
    backs = []
    for b in blocks:
        be = shard.GpuBackend(plan, 6)
        be.prepass(xd[:, b["in_lo"]:b["in_hi"]].contiguous(), b)
        backs.append(be)
    carry = backs[0].zero_carry()
    for b, be in zip(blocks, backs):       # the carry chain, lane by lane as the ranks do it
        carry = torch.cat([be.scan(carry[c:c + 1].contiguous(), c, 1, b["carry_frame"]) for c in range(ch)])
    got = torch.cat([be.render() for be in backs], dim=1)
    torch.cuda.synchronize()
    assert got.shape == whole.shape
    assert torch.equal(got, whole)
    st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
          (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
    plan.render(xd, F, int(sch["hop"][0]), hop_tab=sch["hop"], state=st)
    assert torch.equal(carry, st["phase_sum"])   # the last block's scan ends on the render's final phase_sum


@pytest.mark.parametrize("n,hr,precision", [(1024, 0.25, "f64"), (512, 0.2003, "f32"), (64, 1.6, "f64")])
def test_streaming_front_end_bit_identical(env, n, hr, precision):
    """Audio periods of arbitrary, varying length pushed one by one (tsb_stream_process) give exactly the samples
    of one whole render, and nothing comes out before the reference's start-up latency (2048 + frame samples)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    ch, samples = 2, 40000
    x = orc.gen_signal(4, ch, samples)
    whole, F, _ = run_pv(env, x, n, hr, 1.5, precision)
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=n * hr, ratio=1.5)
    st = tsb.Stream(plan, ch, n * hr)
    rng = np.random.default_rng(0)
    pos, parts, first_output_at = 0, [], None
    while pos < samples:
        k = int(min(samples - pos, rng.choice([1, 64, 256, 1024, 4096, 777])))
        y = st.process(np.ascontiguousarray(x[pos:pos + k].T))
        pos += k
        if y.shape[1] and first_output_at is None:
            first_output_at = pos
        parts.append(y)
    got = np.concatenate(parts, axis=1)
    assert st.frames == F and got.shape == whole.shape
    assert np.array_equal(got, whole)
    assert first_output_at >= 2048 + n


def test_timepar_gain_program(env):
    """Without a phase vocoder only the overlap tail crosses segments, and the lead-in frames rebuild it: the
    time-parallel render of the spectral-gain chain is the sequential one bit for bit."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop = 2048, 512
    x = orc.gen_signal(0, 2, 150000)
    sch = tsb.schedule_const(x.shape[0], n, hop)
    gain = 1.0 / (1.0 + np.arange(n // 2 + 1) / 64.0)
    plan = tsb.Plan(env["ctx"], n, precision="f64", program="gain", gain=gain)
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    seq = plan.render(xd, sch["n_frames"], hop).clone()
    par = plan.render_timepar(xd, sch["n_frames"], hop, 40)
    torch.cuda.synchronize()
    assert torch.equal(par, seq)
    ref = orc.render_filter(x, n, 0.25, gain)
    assert maxrel(par.cpu().numpy(), ref["ola64"][:par.shape[1]].T) <= 1e-12


def test_timepar_rejects_bad_arguments(env):
    torch, tsb = env["torch"], env["tsb"]
    plan = tsb.Plan(env["ctx"], 1024, precision="f32", program="pv", hop=256, ratio=2.0)
    x = torch.zeros((2, 20000), dtype=torch.float32, device="cuda")
    with pytest.raises(tsb.TsbError, match="warm-up frame would start before the input"):
        plan.timepar(x, 10, 256, 2, in_off0=100, warmup=True, warm_hop=256)
    tp = plan.timepar(x, 60, 256, 1000)       # asks for far more segments than frames allow
    assert 1 <= tp.segments <= 60 // 4        # every segment emits at least one frame length
    with pytest.raises(tsb.TsbError, match="before tsb_timepar_prepass"):
        tp.render()


def test_two_minute_render_does_not_diverge(env):
    """BASELINE.json: the PV phase state must not diverge over long renders.  120 s of cfg1 (22,489 frames,
    phase_sum up to 1.8e7 rad): fp64 stays within 1e-9 of the oracle with every wrap on the oracle's side; fp32 at
    the hack's own integer ratio (x2, hacks/pitchshift-pv.cpp:6) stays within 1e-5 RMS (a wrap taken on the other
    side shifts a bin by 4*pi there, i.e. not at all)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop, ch, samples = 1024, 256, 2, 120 * 48000
    x = orc.gen_signal(0, ch, samples)
    ref = orc.render_pv(x, n, 0.25, 2.0)
    F = ref["frames"]
    want = ref["ola64"][:F * hop].T
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    for precision in ("f64", "f32"):
        plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=hop, ratio=2.0)
        st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
              (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
        out = plan.render(xd, F, hop, state=st)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        dps = np.abs(st["phase_sum"].cpu().numpy() - ref["phase_sum"][:, :n // 2])
        tail = slice(-5 * 48000, None)
        if precision == "f64":
            assert maxrel(got, want) <= F64_MAXREL
            assert dps.max() < 1e-6                      # no wrap flips; ulp(phase_sum) is 4e-9 here
        else:
            assert rmsrel(got, want) <= F32_RMSREL
            assert rmsrel(got[:, tail], want[:, tail]) <= F32_RMSREL   # no growth towards the end
            flips = dps > 1
            assert np.all(np.abs(dps[flips] / (4 * np.pi) - np.round(dps[flips] / (4 * np.pi))) < 1e-3)  # multiples of 4*pi


@pytest.mark.gpu
def test_cfg3_ten_minutes_teacher_forced(env):
    """BASELINE configs[2] (8 channels, frame 4096, hop 1024, x1.5) over 10 minutes of audio, the way SURVEY.md §7
    (hard part 2) prescribes for renders too long for a free-running 1e-9: every 2048 frames the oracle's PhaseVocoder
    state (last_phase, phase_sum) is loaded into the GPU render, which must then reproduce the oracle's next 64 frames to
    1e-9 of peak, however late in the render; the free-running render is held to the drift two correct double
    implementations show against each other (6e-10 at 10 minutes between the reference's own FFT backends)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import cfg3_teacher_forced
    r = cfg3_teacher_forced.run(600.0, every=2048, window=64, procs=8)
    assert r["frames"] == 28120 and r["windows"] >= 13
    assert r["teacher_forced_max_rel_worst"] <= F64_MAXREL
    assert r["free_running_max_rel_worst"] <= 5e-9
    assert max(x["max_phase_sum_rad"] for x in r["rows"]) > 5e7   # the regime the check is about


@pytest.mark.gpu
def test_fp32_two_minutes_at_x1p5_never_flips(env):
    """BASELINE.json: "PV phase state must never diverge" — fp32 mode at a NON-integer ratio (cfg3's x1.5), where a wrap
    decided on the other side than the reference's double arithmetic is a permanent sign flip of the bin
    (rtlib/pv.hpp:39-45).  120 s of cfg1 at x1.5 flipped several bins before doubtful wraps were re-evaluated in
    double precision (Chain::phase_wrap_fix); now: zero flips against the oracle, 1e-5 RMS held to the end, and the
    time-parallel fp32 render is the sequential one bit for bit."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop, ch, samples, ratio = 1024, 256, 2, 120 * 48000, 1.5
    x = orc.gen_signal(0, ch, samples)
    ref = orc.render_pv(x, n, 0.25, ratio)
    F = ref["frames"]
    want = ref["ola64"][:F * hop].T
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    plan = tsb.Plan(env["ctx"], n, precision="f32", program="pv", hop=hop, ratio=ratio)
    outs = []
    for nseg in (0, 148):
        st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
              (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
        out = plan.render_timepar(xd, F, hop, nseg, state=st) if nseg else plan.render(xd, F, hop, state=st)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        dps = np.abs(st["phase_sum"].cpu().numpy() - ref["phase_sum"][:, :n // 2])
        assert (dps > 1).sum() == 0, f"{int((dps > 1).sum())} wrap flips"
        assert dps.max() < 1e-3
        tail = slice(-5 * 48000, None)
        assert rmsrel(got, want) <= F32_RMSREL
        assert rmsrel(got[:, tail], want[:, tail]) <= F32_RMSREL
        outs.append((got, st["phase_sum"].cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("n,hr,shift,ch,samples", [(1024, 0.25, 2.0, 2, 96000), (4096, 0.25, 1.5, 8, 60000),
                                                   (256, 0.3, 1.5, 2, 9000), (64, 1.6, 1.5, 1, 9000),
                                                   (512, 0.2003, 0.75, 1, 16000), (8192, 0.25, 1.5, 1, 60000)])
def test_f64_out32_is_the_reference_float_stream(env, n, hr, shift, ch, samples):
    """TSB_F64_OUT32 (fp64 arithmetic, float32 tail/output combined as tinyspec.cpp:173-200): the output is the
    reference's own float32 stream — identical samples except where the ~1e-10 arithmetic difference straddles a
    float32 rounding boundary (then one float32 step of the tail) — and the PV state is that of the fp64 mode."""
    orc = env["orc"]
    x = orc.gen_signal(2, ch, samples)
    ref = orc.render_pv(x, n, hr, shift)
    out, F, st = run_pv(env, x, n, hr, shift, "f64o32")
    assert out.dtype == np.float32
    want = ref["stream"][:out.shape[1]].T
    assert np.count_nonzero(out != want) <= 1e-3 * out.size
    assert np.abs(out.astype(np.float64) - want).max() <= 2.0 ** -23 * np.abs(want).max()
    _, _, st64 = run_pv(env, x, n, hr, shift, "f64")
    assert np.array_equal(st["phase_sum"], st64["phase_sum"]) and np.array_equal(st["last_phase"], st64["last_phase"])
    # host-buffer path: same bits, also from the staging buffers the context kept, also after releasing them
    for rep in range(3):
        outh, _, _ = run_pv(env, x, n, hr, shift, "f64o32", host=True)
        assert np.array_equal(outh, out)
        if rep == 1:
            env["ctx"].trim()


@pytest.mark.gpu
def test_f64_out32_stream_and_timepar(env):
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hr, ch, samples = 1024, 0.25, 2, 120000
    x = orc.gen_signal(4, ch, samples)
    whole, F, _ = run_pv(env, x, n, hr, 1.5, "f64o32")
    plan = tsb.Plan(env["ctx"], n, precision="f64o32", program="pv", hop=n * hr, ratio=1.5)
    st = tsb.Stream(plan, ch, n * hr)
    parts = [st.process(np.ascontiguousarray(x[p:p + 5000].T)) for p in range(0, samples, 5000)]
    got = np.concatenate(parts, axis=1)
    assert got.dtype == np.float32 and np.array_equal(got, whole)
    # time-parallel: every segment rebuilds its float32 overlap tail from its lead-in frames: the same samples
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    sch = tsb.schedule_const(samples, n, n * hr)
    tp = plan.render_timepar(xd, F, int(sch["hop"][0]), in_off0=int(sch["in_off"][0]), n_segments=16)
    torch.cuda.synchronize()
    tp = tp.cpu().numpy()
    assert tp.dtype == np.float32
    assert np.array_equal(tp, whole)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["reverse", "random", "sparse"])
def test_pv_arbitrary_bin_map(env, kind):
    """Bin maps a linear shift never produces (a permutation, random collisions and drops, a few survivors) through
    the plan's used-source flags, destination order and multi-source gather, against a frame-by-frame composition of
    the oracle's PhaseVocoder steps (oracle PV.shift_map = rtlib/pv.hpp:56-79 with fn given on the bin centres)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hop, ch = 1024, 256, 2
    half, rate = n // 2, 48000.0
    centre = np.arange(half) / n * rate
    rng = np.random.default_rng(11)
    if kind == "reverse":
        fb, ratio = centre[::-1].copy() + 1.0, 1.0
    elif kind == "random":
        fb, ratio = rng.integers(-20, half + 40, half) / n * rate + 1.0, 1.3
    else:
        fb = np.full(half, 1e9)
        keep = rng.choice(half, 30, replace=False)
        fb[keep] = centre[rng.choice(half, 30, replace=False)] + 1.0
        ratio = 0.8
    x = orc.gen_signal(7, ch, 2048 + n + 30 * hop)
    F = orc.n_frames(x.shape[0], n, hop)
    pv = orc.PV(float(hop), rate)
    want = np.zeros((ch, F * hop + n))
    for f in range(F):
        o = 2048 + f * hop
        pv.analyze(np.ascontiguousarray(x[o:o + n].T, np.float64))
        pv.shift_map(fb, ratio)
        want[:, f * hop:f * hop + n] += pv.synthesize()
    want = want[:, :F * hop]
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    for precision in ("f64", "f32"):
        plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=hop, ratio=ratio, fn_bins=fb)
        st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
              (("last_phase", half), ("phase_sum", half), ("ola", n))}
        out = plan.render(xd, F, hop, state=st).cpu().numpy()
        if precision == "f64":
            assert maxrel(out, want) <= F64_MAXREL
            assert np.abs(st["phase_sum"].cpu().numpy() - pv.phase_sum[:, :half]).max() < 1e-8
            assert np.abs(st["last_phase"].cpu().numpy() - pv.last_phase[:, :half]).max() < 1e-10
        else:
            assert rmsrel(out, want) <= F32_RMSREL


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_fused_pv_with_its_own_hop(env, precision):
    """PhaseVocoder::set_hop (rtinc/pv.hpp:46-53): the hop inside phase_to_frequency / frequency_to_phase is the
    object's own (the plan's `hop`), while frames advance by the scheduler's hop (render's `hop`)."""
    torch, tsb, orc = env["torch"], env["tsb"], env["orc"]
    n, hs, ch, pv_hop, ratio = 1024, 256, 2, 200.0, 1.5
    x = orc.gen_signal(0, ch, 2048 + 60 * hs + n)
    F = orc.n_frames(x.shape[0], n, hs)
    pv = orc.PV(pv_hop)
    acc = np.zeros((ch, F * hs + n))
    for f in range(F):
        fr = np.ascontiguousarray(x[2048 + f * hs:2048 + f * hs + n].T.astype(np.float64))
        pv.analyze(fr)
        pv.shift(ratio)
        acc[:, f * hs:f * hs + n] += pv.synthesize()
    plan = tsb.Plan(env["ctx"], n, precision=precision, program="pv", hop=pv_hop, ratio=ratio)
    out = plan.render(torch.from_numpy(np.ascontiguousarray(x.T)).cuda(), F, hs).cpu().numpy().astype(np.float64)
    ref = acc[:, :F * hs]
    if precision == "f64":
        assert np.abs(out - ref).max() <= 1e-9 * np.abs(ref).max()
    else:
        assert np.sqrt(np.mean((out - ref) ** 2)) <= 1e-5 * np.sqrt(np.mean(ref ** 2))
