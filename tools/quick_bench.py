"""tools/quick_bench.py — developer probe (not the contract bench): time the fused chain at a few
shapes with CUDA events.  usage: python tools/quick_bench.py [n] [lanes,lanes,...] [seconds] [precisions]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tsb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lanes_list = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [148, 592, 1184, 2368]
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
precs = sys.argv[4].split(",") if len(sys.argv) > 4 else ["f32", "f64"]
prog = sys.argv[5] if len(sys.argv) > 5 else "pv"
hop = n // 4
L = int(seconds * 48000)
ctx = tsb200.Context(0)
sch = tsb200.schedule_const(L, n, hop)
F = sch["n_frames"]
for precision in precs:
    plan = tsb200.Plan(ctx, n, precision=precision, program=prog, hop=hop, ratio=2.0)
    for lanes in lanes_list:
        x = (torch.rand((lanes, L), device="cuda") - 0.5)
        out = torch.zeros((lanes, F * hop), dtype=plan.out_dtype, device="cuda")
        for _ in range(2):
            plan.render(x, F, hop, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            plan.render(x, F, hop, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fc = F * lanes
        osz = 4 if precision == "f32" else 8
        balg = hop * (4 + osz)
        print(f"{prog} n={n} {precision} lanes={lanes:5d} frames={F} {ms:9.3f} ms  {fc / ms / 1e3:9.2f} M frame-ch/s  "
              f"{fc * hop / ms / 1e6:8.3f} G out-samples/s  alg {fc * balg / ms / 1e6:8.1f} GB/s "
              f"({fc * balg / ms / 1e6 / 6545 * 100:5.2f}% of 6545)", flush=True)
        del x, out
