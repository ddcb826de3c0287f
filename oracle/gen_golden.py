"""oracle/gen_golden.py — TEST INFRASTRUCTURE: regenerate tests/golden/*.npz.

Runs the UNMODIFIED reference (oracle/_ref/ref_* binaries = /root/reference/tinyspec.cpp compiled where
it lies + stub headers + offline fake JACK + FFTW-ABI shim; `make -C oracle ref`) on the synthetic
inputs of SURVEY.md §8(d) and stores what it produced.  The fixtures travel to the GPU box; the
reference tree does not.  Usage:  python oracle/gen_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(HERE, "..", "tests", "golden")


def run_ref(binary, env, dumps=(), want_out=True):
    with tempfile.TemporaryDirectory() as d:
        e = dict(os.environ, TSREF_DUMP_DIR=d, **{k: str(v) for k, v in env.items()})
        if want_out:
            e["TSREF_OUT"] = os.path.join(d, "out.f32")
        subprocess.run([os.path.join(REF, binary)], env=e, cwd=d, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        res = {}
        for name, dt in dumps:
            p = os.path.join(d, name + ".bin")
            res[name] = np.fromfile(p, dtype=dt) if os.path.exists(p) else np.zeros(0, dt)
        if want_out:
            res["stream"] = np.fromfile(e["TSREF_OUT"], dtype=np.float32)
        return res


PV_DUMPS = [("frame", np.int64), ("t", np.float64), ("in", np.float64), ("out", np.float64),
            ("last_phase", np.float64), ("phase_sum", np.float64), ("pv_data", np.float64)]


def pv_case(name, n, hop_ratio, shift, ch, samples, first, every, stream=0, tail_only=False, binary="ref_pv_tap"):
    r = run_ref(binary, dict(TSREF_N=n, TSREF_HOP_RATIO=hop_ratio, TSREF_SHIFT=shift, TSREF_CH=ch,
                             TSREF_SAMPLES=samples, TSREF_STREAM=stream, TSREF_DUMP_FIRST=first,
                             TSREF_DUMP_EVERY=every), PV_DUMPS)
    k = len(r["frame"])
    out = dict(meta=np.array([n, ch, samples, stream], np.int64), params=np.array([hop_ratio, shift], np.float64),
               frame=r["frame"], t=r["t"], out=r["out"].reshape(k, ch, n),
               last_phase=r["last_phase"].reshape(k, ch, n), phase_sum=r["phase_sum"].reshape(k, ch, n),
               pv_data=r["pv_data"].reshape(k, ch, n, 2), stream_len=np.array([len(r["stream"])], np.int64))
    if not tail_only:
        out["in"] = r["in"].reshape(k, ch, n)
    s = r["stream"].reshape(-1, ch)
    out["stream"] = s[-8192:] if tail_only else s
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "frames dumped", k, "stream", s.shape)


def filter_case(name, n, hop_ratio, ch, samples, filt, first, every):
    r = run_ref("ref_filter_tap", dict(TSREF_N=n, TSREF_HOP_RATIO=hop_ratio, TSREF_CH=ch, TSREF_SAMPLES=samples,
                                       TSREF_FILTER=filt, TSREF_DUMP_FIRST=first, TSREF_DUMP_EVERY=every),
                [("frame", np.int64), ("t", np.float64), ("in", np.float64), ("out", np.float64),
                 ("spec", np.complex128)])
    k = len(r["frame"])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), meta=np.array([n, ch, samples, filt], np.int64),
                        params=np.array([hop_ratio], np.float64), frame=r["frame"], t=r["t"],
                        out=r["out"].reshape(k, ch, n), spec=r["spec"].reshape(k, ch, n),
                        stream=r["stream"].reshape(-1, ch))
    print(name, "frames dumped", k, "stream", r["stream"].shape)


def helpers_case():
    r = run_ref("ref_helpers_tap", {}, [("hann", np.float64), ("sqrt_hann", np.float64),
                                        ("frft_in", np.complex128), ("frft_out", np.complex128),
                                        ("frft_inplace", np.complex128), ("pv_in", np.float64),
                                        ("pv_analyzed", np.float64), ("pv_shifted", np.float64),
                                        ("pv_out", np.float64), ("pv_phase_sum", np.float64),
                                        ("pv_last_phase", np.float64), ("hop_fixed", np.uint64)], want_out=False)
    np.savez_compressed(os.path.join(GOLD, "helpers.npz"), frft_sizes=np.array([8, 64, 100, 512]),
                        frft_exps=np.array([1.0, -1.0, 2.0, 0.5]), pv_sizes=np.array([64, 256]),
                        pv_ratios=np.array([2.0, 1.5, 0.5]), hops=np.array([256.0, 256.5, 100.25, 0.75, 1023.999]),
                        **r)
    print("helpers", {k: v.shape for k, v in r.items()})


def granular_case(name, samples, vary, n_first=0, stream=0):
    r = run_ref("ref_granular_tap", dict(TSREF_SAMPLES=samples, TSREF_VARY=vary, TSREF_NFIRST=n_first,
                                         TSREF_STREAM=stream, TSREF_DUMP_FIRST=1 << 30),
                [("frame", np.int64), ("t", np.float64), ("win", np.int64), ("osize", np.int64)])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"),
                        meta=np.array([samples, vary, n_first, stream], np.int64), t=r["t"], win=r["win"],
                        osize=r["osize"], stream=r["stream"].reshape(-1, 2))
    print(name, "frames", len(r["t"]), "stream", r["stream"].shape)


def bytebeat_case(name, n, ch, frames, every=1):
    """every > 1 (the full cfg5 size): keep every `every`-th sample of each frame's double `out` and of the stream —
    each sample depends on every bin of the spectrum, so a subset pins the transform as well as all of them"""
    r = run_ref("ref_bytebeat_tap", dict(TSREF_N=n, TSREF_CH=ch, TSREF_MAX_FRAMES=frames, TSREF_DUMP_FIRST=frames),
                [("t", np.float64), ("out", np.float64)])
    k = len(r["t"])
    out = r["out"].reshape(k, ch, n)
    stream = r["stream"].reshape(-1, ch)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), meta=np.array([n, ch, frames], np.int64), t=r["t"],
                        out=out[:, :, ::every], stream=stream[::every], every=np.array([every], np.int64),
                        stream_len=np.array([stream.shape[0]], np.int64),
                        frame_sum=out.sum(axis=2), frame_sumsq=(out ** 2).sum(axis=2))
    print(name, "frames dumped", k, "stream", stream.shape)


def hack_case(hack, n, frames, every=1):
    """An UNMODIFIED spectral_bytebeat hack of the reference (ref/taps/hack_tap.cpp includes the file where it lies):
    its callback times and per-frame double `out` buffers (two channels, frame n, hop n/2 as the hack sets them)."""
    r = run_ref("ref_hack_%s_mkl" % hack, dict(TSREF_MAX_FRAMES=frames, TSREF_DUMP_FIRST=frames),
                [("t", np.float64), ("out", np.float64)])
    k = len(r["t"])
    out = r["out"].reshape(k, 2, n)
    stream = r["stream"].reshape(-1, 2)
    np.savez_compressed(os.path.join(GOLD, "hack_%s.npz" % hack), meta=np.array([n, 2, frames], np.int64), t=r["t"],
                        out=out[:, :, ::every], stream=stream[::every], every=np.array([every], np.int64),
                        stream_len=np.array([stream.shape[0]], np.int64),
                        frame_sum=out.sum(axis=2), frame_sumsq=(out ** 2).sum(axis=2))
    print("hack", hack, "frames dumped", k, "stream", stream.shape, "t", r["t"][:4])


def hack_cases():
    hack_case("voidbubbles", 1 << 12, 24)
    hack_case("acidbattery", 1 << 11, 24)
    hack_case("brokenconvo", 1 << 15, 8, every=8)
    hack_case("ghostacid", 1 << 14, 8, every=4)
    hack_case("andshuffle", 1 << 17, 4, every=32)


def control_case(name, samples):
    r = run_ref("ref_control_tap", dict(TSREF_SAMPLES=samples, TSREF_STREAM=1), [("rec", np.float64)])
    rec = r["rec"].reshape(-1, 4)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), meta=np.array([samples, 1], np.int64), rec=rec,
                        stream=r["stream"].reshape(-1, 2))
    print(name, "frames", rec.shape[0], "stream", r["stream"].shape)


def pvmore_case():
    tags = ["noshift", "nonlin", "bufshift", "ownhop", "copy"]
    dumps = [("in", np.float64), ("ownhop_hop", np.float64)]
    for t in tags:
        dumps += [(t + "_out", np.float64), (t + "_data", np.float64), (t + "_phase_sum", np.float64),
                  (t + "_last_phase", np.float64)]
    r = run_ref("ref_pvmore_tap", {}, dumps, want_out=False)
    np.savez_compressed(os.path.join(GOLD, "pvmore.npz"), n=np.array([128]), frames=np.array([4, 4, 3, 6, 1]), **r)
    print("pvmore", {k: v.shape for k, v in r.items()})


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    if sys.argv[1:] == ["hacks"]:  # only the fixtures of the unmodified spectral_bytebeat hacks
        hack_cases()
        sys.exit(0)
    helpers_case()
    # cfg1 shape, short: every frame's state would be 1.3 MB; keep the first 6 + every 16th
    pv_case("pv_cfg1_short", 1024, 0.25, 2.0, 2, 24000, 6, 16)
    pv_case("pv_x1p5_n256", 256, 0.25, 1.5, 2, 12000, 4, 32)
    pv_case("pv_down_n512_fracthop", 512, 0.2003, 0.75, 1, 16000, 4, 32, stream=3)
    pv_case("pv_cfg3_shape", 4096, 0.25, 1.5, 8, 16384 + 2048, 0, 8, tail_only=True)
    # cfg1 full length (60 s): final state + last 8192 output samples pin long-run phase accumulation
    pv_case("pv_cfg1_60s_tail", 1024, 0.25, 2.0, 2, 2880000, 0, 11238, tail_only=True)
    filter_case("filter_pass_n2048", 2048, 0.25, 1, 24000, 0, 3, 16)
    filter_case("filter_lp_n2048", 2048, 0.25, 1, 24000, 1, 3, 16)
    filter_case("filter_lp_n512_stereo", 512, 0.5, 2, 12000, 1, 3, 16)
    granular_case("granular_1800", 60000, 0)
    granular_case("granular_vary", 60000, 1, n_first=64, stream=2)
    bytebeat_case("bytebeat_n2048", 2048, 2, 12)
    # BASELINE cfg5 at its real size: frame 65536 -> frft of 131072 points (two channels, six frames, every 32nd sample)
    bytebeat_case("bytebeat_n65536", 65536, 2, 6, every=32)
    control_case("control_times", 12000)
    pvmore_case()
    hack_cases()
