"""tools/long_render_check.py — BASELINE.json: "PV phase state must never diverge over a 10-minute render".
Renders cfg1 (stereo 48 kHz, N=1024, hop 256) for --seconds with the oracle (CPU, minutes) and on the GPU
(fp64 sequential, fp64 time-parallel, fp32), and reports output error, phase_sum error and wrap flips."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tsb200
from oracle import orc

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=600.0)
ap.add_argument("--ratio", type=float, default=1.5)
ap.add_argument("--cache", default=None, help="npz path: reuse the oracle's result between runs (comparing library builds)")
args = ap.parse_args()
n, hop, ch = 1024, 256, 2
L = int(args.seconds * 48000)
x = orc.gen_signal(0, ch, L)
t0 = time.time()
if args.cache and os.path.exists(args.cache):
    z = np.load(args.cache)
    ref = {"frames": int(z["frames"]), "ola64": z["ola64"], "phase_sum": z["phase_sum"]}
else:
    ref = orc.render_pv(x, n, 0.25, args.ratio)
    if args.cache:
        np.savez(args.cache, frames=ref["frames"], ola64=ref["ola64"], phase_sum=ref["phase_sum"])
t_cpu = time.time() - t0
F = ref["frames"]
ctx = tsb200.Context(0)
xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
want = ref["ola64"][:F * hop].T
peak = np.abs(want).max()
res = {"seconds": args.seconds, "ratio": args.ratio, "frames": F, "oracle_cpu_s": t_cpu,
       "max_phase_sum_rad": float(np.abs(ref["phase_sum"]).max())}
for name, precision, nseg in (("f64", "f64", 0), ("f64_timepar_592seg", "f64", 592), ("f32", "f32", 0),
                              ("f32_timepar_592seg", "f32", 592)):
    plan = tsb200.Plan(ctx, n, precision=precision, program="pv", hop=hop, ratio=args.ratio)
    st = {k: torch.zeros((ch, w), dtype=torch.float64, device="cuda") for k, w in
          (("last_phase", n // 2), ("phase_sum", n // 2), ("ola", n))}
    out = plan.render_timepar(xd, F, hop, nseg, state=st) if nseg else plan.render(xd, F, hop, state=st)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    dps = np.abs(st["phase_sum"].cpu().numpy() - ref["phase_sum"][:, :n // 2])
    # error over time: last 5 seconds vs first 5 seconds
    tail = slice(-5 * 48000, None)
    res[name] = {"max_rel": float(np.abs(got - want).max() / peak),
                 "rms_rel": float(np.sqrt(np.mean((got - want) ** 2) / np.mean(want ** 2))),
                 "rms_rel_last_5s": float(np.sqrt(np.mean((got[:, tail] - want[:, tail]) ** 2) / np.mean(want[:, tail] ** 2))),
                 "phase_sum_max_abs_diff_rad": float(dps.max()), "wrap_flips": int((dps > 1).sum())}
print(json.dumps(res))
