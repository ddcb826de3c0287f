"""tools/cfg3_timeshard.py — BASELINE cfg3 on N GPUs of one box: phase-vocoder pitch shift x1.5, frame 4096,
hop 1024, 8 channels, ONE long render split by time segment across ranks (shard.py: input halo of lead-in and
warm-up frames, the running phase_sum handed down the ranks over NCCL), every rank's block itself cut into
concurrent time segments.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/cfg3_timeshard.py --seconds 3600 [--precision f64]
Rank 0 prints one JSON line: the sharded render's time, the same render on rank 0 alone, the strong-scaling
efficiency, and how many samples of the stitched output differ from the one-GPU render (none).  bench.py runs
`run()` with a shorter render as its `time_sharded` figure."""
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


def _timed(fn, dist, dev, warmup, reps):
    """mean device time of fn() over `reps` runs after `warmup`, max over ranks (CUDA events on the current stream)"""
    ms = []
    out = None
    for i in range(warmup + reps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= warmup:
            ms.append(float(t.item()))
    return out, sum(ms) / len(ms)


def run(dist, rank, world, dev, ctx, seconds=600.0, precision="f64", segments=74, warmup=2, reps=3, check=True, peer=True):
    """cfg3 (8 channels, frame 4096, hop 1024, x1.5) for `seconds` of audio as ONE render split by time over the
    `world` ranks, and the same render on rank 0 alone (time-parallel over `segments` segments): times, the
    strong-scaling efficiency and whether the stitched output is the one-GPU render bit for bit.  dist: the
    initialised torch.distributed module, or None for a single process.  Every rank returns the same dict."""
    plan = tsb200.Plan(ctx, N, precision=precision, program="pv", hop=HOP, rate=RATE, ratio=RATIO)
    L = int(seconds * RATE)
    sch = tsb200.schedule_const(L, N, HOP, RATE)
    F = int(sch["n_frames"])
    res = {"workload": "cfg3", "what": "ONE 8-channel render split by time segment across the GPUs (halo = lead-in + warm-up "
           "frames; the exact running phase_sum is handed down the ranks)", "frame": N, "hop": HOP, "channels": CH,
           "ratio": RATIO, "seconds": seconds, "frames": F, "n_gpus": world, "precision": precision,
           "segments_per_gpu": segments}
    # the same render on one GPU (rank 0; the others wait) — first, before the ranks map each other's memory: the
    # baseline is the plain one-GPU render a single process would do
    whole, ms1 = None, 0.0
    if rank == 0:
        xw = synth(0, L, dev)
        # (with the segment count of a one-GPU run: the render is bit-identical whatever the segmentation, and more
        # segments only add lead-in frames — the baseline should be the fastest one-GPU render, not a handicapped one)
        whole, ms1 = _timed(lambda: plan.render_timepar(xw, F, HOP, segments), None, dev, 1, reps)
        del xw
    t = torch.tensor([ms1], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms1 = float(t.item())
    out = None
    if world > 1:
        b = shard.plan_time_blocks(sch, N, world)[rank]
        x = synth(b["in_lo"], b["in_hi"], dev)
        be = shard.GpuBackend(plan, segments)
        if peer:
            be.link(dist, CH)
        out, ms = _timed(lambda: shard.render_time_sharded(be, x, b, dist), dist, dev, warmup, reps)
        res["carry_transport"] = ("scan kernels chained over NVLink peer memory (tsb_timepar_scan_chain)" if peer else
                                  "NCCL send/recv per lane")
        res["halo_samples"] = int(b["in_off0"]) if rank else None
    if world > 1 and os.environ.get("TSB_SHARD_TRACE"):
        # where the time goes: the three steps one by one, each synchronised (device time, per rank)
        import time as _t
        def step(fn):
            torch.cuda.synchronize(); dist.barrier(); t0 = _t.perf_counter(); fn(); torch.cuda.synchronize()
            return (_t.perf_counter() - t0) * 1e3
        for _ in range(2):
            tp = step(lambda: be.prepass(x, b))
            tc = step(lambda: be.scan_chain(b["carry_frame"])) if be.linked else -1.0
            tr = step(lambda: be.render())
        allv = [None] * world
        dist.all_gather_object(allv, (rank, round(tp, 3), round(tc, 3), round(tr, 3)))
        res["trace_prepass_chain_render_ms"] = allv
    if world == 1:
        ms = ms1
    res.update({"ms": ms, "ms_1gpu": ms1, "frames_per_s": F / ms * 1e3, "frame_channels_per_s": F * CH / ms * 1e3,
                "out_samples_per_s": F * CH * HOP / ms * 1e3, "x_realtime": seconds / (ms * 1e-3),
                "efficiency_vs_1gpu": ms1 / (world * ms)})
    if check and world > 1:
        # the ranks hand their blocks to rank 0, which compares them with its own render of the whole thing
        blocks = shard.plan_time_blocks(sch, N, world)
        bad = torch.zeros(2, dtype=torch.float64, device=dev)
        if rank == 0:
            peak = float(whole.abs().max().item())
            for r in range(world):
                br = blocks[r]
                part = out if r == 0 else torch.empty((CH, br["out_hi"] - br["out_lo"]), dtype=whole.dtype, device=dev)
                if r > 0:
                    dist.recv(part, r)
                ref = whole[:, br["out_lo"]:br["out_hi"]]
                bad[0] = max(float(bad[0].item()), float((part - ref).abs().max().item()) / peak)
                bad[1] += float((part != ref).sum().item())
        else:
            dist.send(out.contiguous(), 0)
        dist.broadcast(bad, 0)
        res["max_rel_err_vs_1gpu"] = float(bad[0].item())
        res["samples_differing_from_1gpu"] = int(bad[1].item())
    elif world == 1:
        res["max_rel_err_vs_1gpu"] = 0.0
        res["samples_differing_from_1gpu"] = 0
    del whole, out
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--segments", type=int, default=74, help="time segments per rank (x 8 lanes = CTAs)")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--nccl-carry", action="store_true", help="hand the phase carry on with NCCL send/recv instead of peer memory")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    d = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        d = dist
    ctx = tsb200.Context(local)
    res = run(d, rank, world, dev, ctx, args.seconds, args.precision, args.segments, check=not args.no_check,
              peer=not args.nccl_carry)
    if rank == 0:
        print(json.dumps(res))
    if d is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
