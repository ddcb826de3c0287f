"""tools/cfg3_timeshard.py — BASELINE cfg3 on N GPUs of one box: phase-vocoder pitch shift x1.5, frame 4096,
hop 1024, 8 channels, ONE long render split by time segment across ranks (shard.py: input halo, the running
phase_sum handed down the ranks, overlap-tail send/recv over NCCL), every rank's block itself cut into concurrent
time segments.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/cfg3_timeshard.py --seconds 3600 [--precision f64] [--check-seconds 30]
Rank 0 prints one JSON line.  --check-seconds T first renders T seconds both sharded and on rank 0 alone and
compares them (<= 1e-9 of peak in f64)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import tsb200
from tinyspec_cling_b200 import shard

N, HOP, CH, RATE, RATIO = 4096, 1024, 8, 48000.0, 1.5


def synth(lo, hi, device):
    """Samples [lo, hi) of the 8 channels: sine + hash noise, a pure function of the sample index so that every
    rank can generate exactly its own range."""
    n = torch.arange(lo, hi, device=device, dtype=torch.float64)
    c = torch.arange(CH, device=device, dtype=torch.float64)[:, None]
    f = 440.0 + 110.0 * c
    noise = torch.frac(torch.sin(n[None, :] * 12.9898 + c * 78.233) * 43758.5453) - 0.5
    return (0.5 * torch.sin(2 * np.pi * f * n[None, :] / RATE) + 0.1 * noise).float().contiguous()


def sharded(ctx, plan, sch, rank, world, dev, nseg):
    b = shard.plan_time_blocks(sch, N, world)[rank]
    x = synth(b["in_lo"], b["in_hi"], dev)
    be = shard.GpuBackend(plan, HOP, nseg)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = shard.render_time_sharded(be, x, b, dist)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, b, float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--check-seconds", type=float, default=30.0)
    ap.add_argument("--segments", type=int, default=74, help="time segments per rank (x 8 lanes = CTAs)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = tsb200.Context(local)
    plan = tsb200.Plan(ctx, N, precision=args.precision, program="pv", hop=HOP, rate=RATE, ratio=RATIO)
    res = {"workload": "cfg3", "frame": N, "hop": HOP, "channels": CH, "ratio": RATIO, "n_gpus": world,
           "precision": args.precision, "segments_per_rank": args.segments}
    if args.check_seconds > 0:
        L = int(args.check_seconds * RATE)
        sch = tsb200.schedule_const(L, N, HOP, RATE)
        out, b, _ = sharded(ctx, plan, sch, rank, world, dev, args.segments)
        sizes = [0] * world
        gathered = [None] * world
        dist.all_gather_object(gathered, (b["out_lo"], out.cpu().numpy()))
        if rank == 0:
            got = np.concatenate([g[1] for g in sorted(gathered, key=lambda g: g[0])], axis=1)
            x = synth(0, L, dev)
            whole = plan.render(x, sch["n_frames"], HOP)
            torch.cuda.synchronize()
            whole = whole.cpu().numpy()
            err = float(np.abs(got - whole[:, :got.shape[1]]).max() / np.abs(whole).max())
            res["check"] = {"seconds": args.check_seconds, "frames": int(sch["n_frames"]), "max_rel_err_vs_single_gpu": err}
            assert got.shape[1] == sch["n_frames"] * HOP
            assert err <= (1e-9 if args.precision == "f64" else 1e-5), err
    L = int(args.seconds * RATE)
    sch = tsb200.schedule_const(L, N, HOP, RATE)
    for _ in range(2):
        out, b, ms = sharded(ctx, plan, sch, rank, world, dev, args.segments)
    F = int(sch["n_frames"])
    res.update({"seconds": args.seconds, "frames": F, "ms": ms, "frames_per_s": F / ms * 1e3,
                "frame_channels_per_s": F * CH / ms * 1e3, "out_samples_per_s": F * CH * HOP / ms * 1e3,
                "x_realtime": args.seconds / (ms * 1e-3)})
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
