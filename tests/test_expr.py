"""Spectrum expressions (TSB_GEN_EXPR): the expression compiler, the transcriptions of the reference's
stateless spectral_bytebeat hacks and the static program check, on the CPU; the device interpreter inside the fused
spectral synthesis against the UNMODIFIED hacks compiled from /root/reference (tests/golden/hack_*.npz) on the GPU."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tsb200  # noqa: E402
from oracle import orc  # noqa: E402
import expr_eval  # noqa: E402

X = tsb200.expr
HACKS = ["voidbubbles", "acidbattery", "andshuffle", "brokenconvo", "ghostacid"]


def test_opcodes_match_the_header():
    """expr.OPS is the TSB_X_* enumeration of include/tsb200.h, in order"""
    src = open(os.path.join(ROOT, "include", "tsb200.h")).read()
    body = src[src.index("TSB_X_I_BIN = 1"):src.index("TSB_X__COUNT")]
    names = re.findall(r"TSB_X_([A-Z0-9_]+)", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == X.OPS[1:]
    for key, val in (("MAX_OPS", X.MAX_OPS), ("MAX_CONST", X.MAX_CONST), ("MAX_TABS", X.MAX_TABS), ("STACK", X.STACK)):
        assert int(re.search(r"#define TSB_EXPR_%s (\d+)" % key, src).group(1)) == val


@pytest.mark.parametrize("hack", HACKS)
def test_hack_programs_equal_the_oracle_restatements(hack):
    """the compiled program of a hack, run by the test-side numpy interpreter, against the oracle's direct restatement
    of the same C source (itself pinned to the compiled unmodified hack, test_oracle_hack_fixtures) — several callback
    times, a small frame so the CPU suite stays fast; both go through numpy's sin, so the match is to the last bits"""
    h = X.HACKS[hack]()
    prog = h["program"]
    tsb200.expr_check(prog.ops, len(prog.iconst), len(prog.dconst), len(prog.itabs), len(prog.dtabs), len(prog.iparams),
                      len(prog.dparams))
    n, ch = 1024, 3
    for t in (0.0, 0.0427, 1.7, 33.3, 471.25):
        want = orc.hack_spectrum(hack, t, ch, n)
        got = expr_eval.evaluate(prog, X.OPS, t, 0, ch, n)
        d = np.abs(got[:, 1:] - want[:, 1:]).max()
        assert d <= 1e-13 * max(1.0, np.abs(want).max()), (hack, t, d)


@pytest.mark.parametrize("hack", HACKS)
def test_oracle_hack_fixtures(hack):
    """oracle.render_hack (numpy restatement of the hack's formula + the bytebeat pattern) against the per-frame double
    buffers of the UNMODIFIED hack compiled from /root/reference (oracle/gen_golden.py hacks)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hack_%s.npz" % hack))
    n, ch, frames = (int(v) for v in g["meta"])
    assert n == orc.HACK_FRAMES[hack] == X.HACKS[hack]()["frame"]
    frames = min(frames, 6 if n <= 1 << 14 else 3)
    every = int(g["every"][0])
    assert np.array_equal(g["t"][:frames], np.arange(frames) * (n // 2) / 48000.0)
    r = orc.render_hack(hack, frames, ch)
    peak = np.abs(g["out"]).max()
    assert np.abs(r["out"][:, :, ::every] - g["out"][:frames]).max() <= 1e-12 * peak


def test_expression_builder_semantics():
    """C's conversions and operators: int/int truncates, a double operand promotes, % follows the dividend's sign,
    wrap to 32 bits on request, x/0 = x%0 = 0, constants are pooled, cplx(mag, mag) evaluates mag once"""
    i = X.bin()
    p = X.compile(X.f64((i - 5) / 2), X.f64((i - 5) % 3))
    v = expr_eval.evaluate(p, X.OPS, 0.0, 0, 1, 11)[0]
    assert list(v.real) == [-2, -2, -1, -1, 0, 0, 0, 1, 1, 2, 2] and list(v.imag) == [-2, -1, 0, -2, -1, 0, 1, 2, 0, 1, 2]
    p = X.compile((i / 2.0), X.f64((i * 70000 * 70000).i32()))
    v = expr_eval.evaluate(p, X.OPS, 0.0, 0, 1, 4)[0]
    assert list(v.real) == [0, 0.5, 1.0, 1.5]
    assert list(v.imag) == [0, float(np.int32(np.uint32(4900000000 & 0xffffffff))), float(np.int64(9800000000).astype(np.int32)),
                            float(np.int64(14700000000).astype(np.int32))]
    p = X.compile(X.f64(7 // (i - 1)), X.f64(7 % (i - 1)))
    v = expr_eval.evaluate(p, X.OPS, 0.0, 0, 1, 3)[0]
    assert list(v.real) == [-7, 0, 7] and list(v.imag) == [0, 0, 0]
    m = X.sin(X.f64(i) * 0.5) * 3.0
    p = X.compile(m, m)
    assert X.OPS[p.ops[-1] & 255] == "PICK" and len(p.dconst) == 2
    re_, im_ = X.polar(2.0, X.f64(i) * 0.25), None
    p = X.compile(X.polar(2.0, X.f64(i) * 0.25))
    v = expr_eval.evaluate(p, X.OPS, 0.0, 0, 1, 8)[0]
    assert np.allclose(v, 2.0 * np.exp(1j * 0.25 * np.arange(8)), rtol=0, atol=1e-15)
    with pytest.raises(TypeError):
        _ = X.sin(X.f64(i)) & 3
    t = X.dparam(lambda t, n: t * n) + X.f64(X.iparam(lambda t, n: int(t * 8)) + X.frame_counter() + X.size())
    v = expr_eval.evaluate(X.compile(t), X.OPS, 2.5, 7, 2, 16)
    assert np.all(v == 2.5 * 16 + 20 + 7 + 16)


def test_expr_check_rejects_malformed_programs():
    C = X.CODE
    ok = [C["I_BIN"], C["I2D"], C["PICK"]]
    tsb200.expr_check(ok)
    bad = {
        "unknown opcode": [200],
        "needs two integers": [C["I_BIN"], C["I2D"], C["I_BIN"], C["I_ADD"]],
        "needs a double": [C["I_BIN"], C["D_SIN"]],
        "exactly two doubles": [C["I_BIN"], C["I2D"]],
        "index out of range": [C["D_CONST"] | (3 << 8), C["PICK"]],
        "PICK below": [C["I_BIN"], C["PICK"] | (1 << 8)],
        "stack overflow": [C["I_BIN"]] * 17,
        "n_ops out of range": [C["I_BIN"]] * 200,
    }
    for why, ops in bad.items():
        with pytest.raises(tsb200.TsbError, match=why):
            tsb200.expr_check(ops, n_dconst=2)
    with pytest.raises(ValueError):
        big = X.bin()
        for k in range(70):
            big = big + k
        X.compile(X.f64(big))


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return tsb200.Context(0)


@pytest.mark.gpu
@pytest.mark.parametrize("hack", HACKS)
def test_spectral_expr_matches_the_unmodified_hacks(ctx, hack):
    """The hack's formula as an expression evaluated inside kernel A of the fused spectral synthesis, at the hack's own
    frame size and hop, against the UNMODIFIED hack compiled from the reference: every frame's double `out` buffer
    (<= 1e-9 of peak, BASELINE.json's fp64 bound) and the overlap-added stream (the reference's is float32)."""
    import torch
    g = np.load(os.path.join(ROOT, "tests", "golden", "hack_%s.npz" % hack))
    n, ch, frames = (int(v) for v in g["meta"])
    every, hop = int(g["every"][0]), n // 2
    prog = X.HACKS[hack]()["program"]
    peak = np.abs(g["out"]).max()
    edge = torch.zeros((ch, 2), dtype=torch.float64, device="cuda")
    worst = 0.0
    for f in range(frames):
        o = ctx.spectral_render(n, ch, 1, hop, first_frame=f, state_edge=edge, load_state=f > 0, expr=prog, times=[float(g["t"][f])])
        got = o.cpu().numpy()[:, :n]
        worst = max(worst, np.abs(got[:, ::every] - g["out"][f]).max() / peak)
    print(hack, "worst frame error / peak", worst)
    assert worst <= 1e-9
    l0 = ctx.launches
    out = ctx.spectral_render(n, ch, frames, hop, expr=prog, times=list(g["t"]))
    assert ctx.launches - l0 == 2 * frames
    L = int(g["stream_len"][0])
    s = out.cpu().numpy()[:, :L].T[::every]
    assert np.abs(s - g["stream"]).max() <= 2e-7 * np.abs(g["stream"]).max()


@pytest.mark.gpu
def test_spectral_expr_vs_oracle_64_channels(ctx):
    """voidbubbles' formula at BASELINE cfg5's shape (64 channels, frame 65536, hop 16384; default clock) against the
    numpy restatement; integer ops with negative operands and fmod / pow / exp / polar through a synthetic program"""
    n, ch, frames, hop = 65536, 64, 2, 16384
    prog = X.voidbubbles()["program"]
    want = orc.render_bytebeat(n, ch, frames, hop=hop, spectrum=lambda f: orc.hack_spectrum("voidbubbles", f * hop / 48000.0, ch, n))
    got = ctx.spectral_render(n, ch, frames, hop, expr=prog).cpu().numpy()
    assert np.abs(got - want["ola64"]).max() <= 1e-9 * np.abs(want["ola64"]).max()
    i, c = X.bin(), X.channel()
    k = ((i - 300) // 7) % 5 ^ (c - 1) * 3 | (-i >> 2)
    ph = X.fmod(X.f64(k) * 0.37, 3.0) + X.exp(-X.f64(i) / 512.0) + X.pow(X.f64(i + 1), 0.25) + X.floor(X.f64(i) / 3.0) * 0.125
    prog2 = X.compile(X.polar(X.sqrt(X.f64(1 + (i & 15))) / X.f64(i + 1), ph))
    n, ch, frames, hop = 1024, 3, 3, 256
    spec = lambda f: expr_eval.evaluate(prog2, X.OPS, 0.0, f, ch, n)
    want = orc.render_bytebeat(n, ch, frames, hop=hop, spectrum=spec)
    got = ctx.spectral_render(n, ch, frames, hop, expr=prog2).cpu().numpy()
    assert np.abs(got - want["ola64"]).max() <= 1e-9 * np.abs(want["ola64"]).max()
