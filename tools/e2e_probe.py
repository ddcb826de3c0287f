"""tools/e2e_probe.py — developer probe: where the time of the host-buffer render (tsb_chain_render_host, bench.py's
e2e leg) goes: raw pinned copy rates in each direction and both at once, then the call itself with its phase trace
(TSB_HOST_TRACE) for several chunk counts (TSB_HOST_CHUNKS)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tsb200  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f64o32"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 370
N, HOP, CH, L = 1024, 256, 2, 48000 * 60
lanes = S * CH
ctx = tsb200.Context(0)
F = tsb200.schedule_const(L, N, HOP, 48000.0)["n_frames"]
plan = tsb200.Plan(ctx, N, precision=prec, program="pv", hop=HOP, rate=48000.0, ratio=2.0)
xh = torch.empty((lanes, L), dtype=torch.float32, pin_memory=True)
xh.normal_()
oh = torch.empty((lanes, F * HOP), dtype=plan.out_dtype, pin_memory=True)
xd = torch.empty((lanes, L), dtype=torch.float32, device="cuda")
od = torch.empty((lanes, F * HOP), dtype=plan.out_dtype, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, nbytes, label):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:50s} {nbytes / dt / 1e9:7.1f} GB/s  ({dt * 1e3:.1f} ms)", flush=True)


for _ in range(2):
    timeit(lambda: xd.copy_(xh, non_blocking=True), xh.numel() * 4, "H2D whole input")
    timeit(lambda: oh.copy_(od, non_blocking=True), oh.numel() * oh.element_size(), "D2H whole output")


def both():
    with torch.cuda.stream(s1):
        xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        oh.copy_(od, non_blocking=True)


timeit(both, xh.numel() * 4 + oh.numel() * oh.element_size(), "H2D input + D2H output concurrently")
del xd, od
torch.cuda.empty_cache()
os.environ["TSB_HOST_TRACE"] = "1"
for chunks in ("", "4", "8", "16", "32", "64"):
    if chunks:
        os.environ["TSB_HOST_CHUNKS"] = chunks
    for rep in range(2):
        t0 = time.perf_counter()
        plan.render_host(xh, F, HOP, out=oh)
        dt = time.perf_counter() - t0
        print(f"chunks={chunks or 'default'} rep {rep}: {dt * 1e3:.1f} ms total -> {S * F / dt / 1e6:.2f} M frames/s", flush=True)
