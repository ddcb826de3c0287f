#!/usr/bin/env python
"""bench.py — the contract benchmark (see DESIGN.md "Measurement").

One step = one pass of the fused hot path (gather -> sqrt-Hann -> R2C -> phase-vocoder analyse / shift x2 /
synthesise -> C2R -> sqrt-Hann -> overlap-add) over one batch of synthetic streams.

Workload (config.workload = "cfg1x"): BASELINE.json configs[0] — hacks/pitchshift-pv.cpp, stereo 48 kHz,
frame 1024, hop 256, 60 s per stream (11,239 frames) — replicated over S independent streams per GPU so
that one launch fills the device (a single stereo stream is 2 lanes; SURVEY.md §8d asks for both figures,
the single-stream one is reported as `single_stream`).  N GPUs: every rank renders its own S streams
(stream sharding, no data-path collective) -> "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision f64|f32] [--streams S] [--seconds T]
  python bench.py --impl reference ...   # the reference's own CPU implementation on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAME, HOP, CH, RATE, SHIFT = 1024, 256, 2, 48000.0, 2.0
# One lane = one CTA of 128 threads; resident CTAs per SM: 4 (fp64 kernel: 128 regs — the overlap tail lives in
# registers — 37 KB smem) / 5 (fp32: 96 regs, 18 KB).  148 SMs x that many CTAs = one full wave of lanes; 2 lanes per
# stereo stream; both precisions run two waves.
STREAMS = {"f64": 148 * 4, "f32": 148 * 5}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def b_alg(precision):
    """Algorithmic bytes per frame-channel: hop*(s_in + s_out); the 32*N_live/F_launch state
    term is counted from the state arrays actually read/written once per launch."""
    return HOP * (4 + (8 if precision == "f64" else 4))


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.mark_at = index, [], None, 0

    def mark(self):
        """rows from here on belong to the load being measured (nvidia-smi is already looping by then)"""
        self.mark_at = len(self.rows)

    def count(self):
        return len(self.rows) - self.mark_at

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows[self.mark_at:]:
            f = [v.strip() for v in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_input(torch, lanes, length, device, seed):
    """0.5*sin(2*pi*f*n/48000) + 0.1*(u-0.5), f = 440 + 110*c + 0.37*s (SURVEY.md §8d shape; u from torch's
    generator instead of the LCG — values need not match the oracle here, parity is tests/' job)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = torch.arange(length, device=device, dtype=torch.float64)
    lane = torch.arange(lanes, device=device)
    f = 440.0 + 110.0 * (lane % CH).double() + 0.37 * (lane // CH).double()
    x = torch.empty((lanes, length), dtype=torch.float32, device=device)
    step = 64
    for a in range(0, lanes, step):
        b = min(lanes, a + step)
        ph = (2.0 * 3.141592653589793 / RATE) * f[a:b, None] * n[None, :]
        x[a:b] = (0.5 * torch.sin(ph) + 0.1 * (torch.rand((b - a, length), generator=g, device=device,
                                                          dtype=torch.float64) - 0.5)).float()
    return x


def ref_binary():
    for name in ("ref_pv_tap_mkl", "ref_pv_tap"):
        p = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(p):
            try:
                env = dict(os.environ, TSREF_SAMPLES="8192", TSREF_N=str(N_FRAME))
                subprocess.run([p], env=env, cwd="/tmp", check=True, timeout=60, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
                return p, ("MKL DFTI standing in for FFTW" if name.endswith("_mkl") else "radix-2 FFT shim standing in for FFTW")
            except Exception:
                continue
    return None, None


def host_cpus():
    """the CPUs this process may run on (cgroup/affinity aware), in order"""
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


# One MKL/OpenMP thread per reference process: the reference is single-threaded (one worker, tinyspec.cpp:284);
# without this every copy's MKL pool spawns a thread per core and P copies oversubscribe the host (round-1 flaw).
ONE_THREAD = {"OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "MKL_DYNAMIC": "FALSE"}


def run_reference(procs, seconds):
    """`procs` concurrent copies of the compiled, unmodified reference (oracle/_ref), one stereo stream each
    (its own README's way to scale, README.md:85-86), each pinned to its own host core with one FFT thread
    (BASELINE.md §3.3).  Time = the binary's own first->last process-callback wall (oracle/ref/fake_jack.cpp,
    TSREF_TIMING), the slowest copy counting for all.  Returns (frames per process, seconds, kind, what, procs)."""
    binary, fft = ref_binary()
    samples = int(seconds * RATE)
    if binary is None:
        # the oracle port as a stand-in (kind "port")
        from oracle import orc
        x = orc.gen_signal(0, CH, samples)
        t0 = time.perf_counter()
        r = orc.render_pv(x, N_FRAME, HOP / N_FRAME, SHIFT)
        return r["frames"], time.perf_counter() - t0, "port", "C oracle (oracle/oracle.c), radix-2 FFT", 1
    import tempfile
    env = dict(os.environ, TSREF_SAMPLES=str(samples), TSREF_N=str(N_FRAME), TSREF_HOP_RATIO=str(HOP / N_FRAME),
               TSREF_SHIFT=str(SHIFT), TSREF_CH=str(CH), TSREF_TIMING="1", TSREF_DUMP_FIRST="0", **ONE_THREAD)
    cpus = host_cpus()
    dirs = [tempfile.mkdtemp(prefix="tsref") for _ in range(procs)]

    def pin(cpu):
        return lambda: os.sched_setaffinity(0, {cpu})

    ps = [subprocess.Popen([binary], env=dict(env, TSREF_STREAM=str(i)), cwd=d, stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, text=True, preexec_fn=pin(cpus[i % len(cpus)]))
          for i, d in enumerate(dirs)]
    outs = []
    for p in ps:
        try:
            outs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append("")
    frames, walls = 0, []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("{"):
                j = json.loads(line)
                frames = j["frames"]
                walls.append(j["wall_s"])
    for d in dirs:
        try:
            os.rmdir(d)
        except OSError:
            pass
    if len(walls) != procs:
        raise SystemExit(f"reference arm: {procs - len(walls)} of {procs} reference processes did not finish")
    return frames, max(walls), "reference", f"unmodified reference (oracle/_ref), {fft}", procs


def frames_per_stream(length):
    """frames the scheduler runs over `length` input samples (tinyspec.cpp:142-149: the first advance is 2048)"""
    return (length - 2048 - N_FRAME) // HOP + 1


def workload_config(seconds, streams, precision):
    """`config` of the JSON line — the same dict for both arms (the reference arm runs a bounded sample of it)"""
    L = int(seconds * RATE)
    return {"workload": "cfg1x", "what": "hacks/pitchshift-pv.cpp chain (BASELINE configs[0]) x S streams",
            "frame": N_FRAME, "hop": HOP, "channels": CH, "rate": RATE, "shift": SHIFT,
            "seconds_per_stream": seconds, "frames_per_stream": frames_per_stream(L), "streams_per_gpu": streams,
            "lanes_per_gpu": streams * CH, "sharding": "streams across GPUs, no collective",
            "l2": f"inputs larger than L2 ({streams * CH * L * 4 / 1e9:.1f} GB in per GPU)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--streams", type=int, default=0,
                    help="independent stereo streams per GPU (default: one full wave of resident CTAs, see STREAMS)")
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-out", default="f32", choices=["f32", "f64"],
                    help="sample type the end-to-end leg returns to the host: the reference's float32 stream (default) "
                         "or, in fp64 mode, the unrounded float64 accumulation")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary figures (other precision, single stream, "
                                                             "time-sharded cfg3, cfg5)")
    ap.add_argument("--ts-seconds", type=float, default=600.0,
                    help="audio seconds of the time-sharded cfg3 render (BASELINE configs[2] is 3600 s; 0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L = int(args.seconds * RATE)
    peak, peak_src = peaks()

    if args.impl == "reference":
        if rank != 0:
            return
        cores = len(host_cpus())
        S = args.streams or STREAMS[args.precision]
        times, frames = [], 0
        for i in range(args.warmup + args.steps):
            frames, wall, kind, what, used = run_reference(cores, args.seconds)
            if i >= args.warmup:
                times.append(wall)
        tot = sum(times)
        fps = used * frames * len(times) / tot
        line = {"impl": "reference", "metric": "stft_pv_istft_frames_per_s", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "frame_channels_per_s": fps * CH, "out_samples_per_s": fps * CH * HOP,
                "frames_per_s_per_core": fps / used,
                "config": workload_config(args.seconds, S, args.precision),
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": kind,
                                 "sample": f"bounded sample of the workload: {used} of its streams per step, one "
                                           f"reference process per host core (pinned, 1 FFT thread each), 1 stereo "
                                           f"stream x {args.seconds:g} s = {frames} frames each; time = the slowest "
                                           f"process's own first->last callback wall; {what}"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import tsb200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libtsb200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = tsb200.Context(local)
    sch = tsb200.schedule_const(L, N_FRAME, HOP, RATE)
    F = sch["n_frames"]
    S = args.streams or STREAMS[args.precision]
    lanes = S * CH
    other = "f32" if args.precision == "f64" else "f64"
    S_other = args.streams or STREAMS[other]
    lanes_alloc = lanes if args.no_extras else max(lanes, S_other * CH)
    x_all = synth_input(torch, lanes_alloc, L, dev, 1234 + rank)
    x = x_all[:lanes]

    def timed(plan, xin, out, steps, warmup, sampler=None):
        if sampler:
            sampler.start()  # nvidia-smi takes a few hundred ms to deliver its first row: start it before the warm-up
        for _ in range(warmup):
            plan.render(xin, F, HOP, out=out)
        barrier()
        if sampler:
            sampler.mark()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            plan.render(xin, F, HOP, out=out)
        e1.record()
        barrier()
        ms, launches = e0.elapsed_time(e1), ctx.launches - l0
        clocks = None
        if sampler:
            # a timed region shorter than a few sampling periods may have caught no row: keep the same launch
            # running (untimed) until there are three, so the clocks are always those under this load
            extra, t_end = 0, time.time() + 3.0
            while sampler.count() < 3 and time.time() < t_end:
                plan.render(xin, F, HOP, out=out)
                torch.cuda.synchronize()
                extra += 1
            clocks = sampler.stop()
            clocks["window"] = "timed region" if not extra else f"timed region + {extra} identical untimed launches"
        return max_over_ranks(ms), launches, clocks

    def device_leg(precision, xin, steps, warmup, sampler=None):
        plan = tsb200.Plan(ctx, N_FRAME, precision=precision, program="pv", hop=HOP, rate=RATE, ratio=SHIFT)
        out = torch.zeros((xin.shape[0], F * HOP), dtype=plan.out_dtype, device=dev)
        ms, launches, clocks = timed(plan, xin, out, steps, warmup, sampler)
        del out
        return ms / steps, launches, clocks

    ms_step, launches, clocks = device_leg(args.precision, x, args.steps, args.warmup, ClockSampler(local))
    frames_all = world * S * F
    fps = frames_all / (ms_step * 1e-3)
    fcs = fps * CH
    balg = b_alg(args.precision)
    state_bytes = lanes * (N_FRAME // 2) * 8 * 2 * 2 * 0  # state arrays are not passed in the bench (reset start)
    alg_bytes_launch = S * F * CH * balg + state_bytes
    achieved = alg_bytes_launch / (ms_step * 1e-3) / 1e9
    # per-launch DRAM traffic and executed arithmetic of this kernel from the committed `ncu --set full` capture of the
    # same launch shape (profiles/ncu_traffic.json: what was measured, how, from which report)
    traffic, flops_fc, ncu_src = None, None, None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(args.precision)
        if tj:
            traffic = tj["dram_bytes_per_frame_channel"] * S * F * CH
            flops_fc = tj.get("flops_per_frame_channel")
            ncu_src = tj.get("source")
    # compute-side roofline: the kernel is issue/arithmetic-bound, so its executed flops (ncu's per-opcode counts: FMA = 2)
    # are reported against the device's measured FMA rate in the same type
    fma_peak = ctx.fma_peak(args.precision)

    line = {"metric": "stft_pv_istft_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "frame_channels_per_s": fcs, "out_samples_per_s": fcs * HOP,
            "config": workload_config(args.seconds, S, args.precision),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_frame_channel": balg,
                         "kernel": "tsb::chain_kernel<" + ("double" if args.precision == "f64" else "float") + ",10,PV>",
                         "compute_peak_tflops": fma_peak,
                         "compute_peak_source": "measured here: tsb_probe_fma_peak (independent " + args.precision + " FMA chains, CUDA events)",
                         "compute_achieved_tflops": None if flops_fc is None else flops_fc * fcs / world / 1e12,
                         "compute_frac": None if flops_fc is None else flops_fc * fcs / world / 1e12 / fma_peak,
                         "flops_per_frame_channel": flops_fc, "ncu_source": ncu_src,
                         "note": "the fused chain is issue/latency-bound, not HBM-bound (SURVEY.md §7 hard part 1): "
                                 "compute_frac (executed flops / measured FMA peak) is the yardstick that applies; "
                                 "see profiles/ for the ncu SOL"},
            "clocks": clocks, "gpu_launches": launches}

    # ---- end to end through the C ABI with HOST buffers (tsb_chain_render_host) ----
    if not args.no_e2e:
        # the reference hands float32 to JACK (tinyspec.cpp:43,173-200): the fp64 chain delivers that stream type
        # (TSB_F64_OUT32: same arithmetic and PV state as the device leg, float32 tail/output as the reference's)
        eprec = "f64o32" if (args.precision == "f64" and args.e2e_out == "f32") else args.precision
        # pinned buffers on the NUMA node next to this rank's GPU (several ranks share the host)
        bound = None if os.environ.get("TSB_NO_NUMA_BIND") else tsb200.bind_host_to_gpu(local)
        plan = tsb200.Plan(ctx, N_FRAME, precision=eprec, program="pv", hop=HOP, rate=RATE, ratio=SHIFT)
        # the end-to-end leg is bound by the host copies whatever the batch: it keeps to 370 streams per GPU (8.5 GB each
        # way per rank; eight ranks pin 136 GB of the one host between them)
        S_e = min(S, 370)
        lanes = S_e * CH
        xh = torch.empty((lanes, L), dtype=torch.float32, pin_memory=True)
        xh.copy_(x[:lanes])
        oh = torch.empty((lanes, F * HOP), dtype=plan.out_dtype, pin_memory=True)
        torch.cuda.empty_cache()
        # the floor: the bare copies of the same buffers, both directions at once, nothing else (every rank at the
        # same time: with several ranks the host's memory and PCIe complex are shared)
        xd2 = torch.empty((lanes, L), dtype=torch.float32, device=dev)
        od2 = torch.empty((lanes, F * HOP), dtype=plan.out_dtype, device=dev)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        floor = []
        for i in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                xd2.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_dn):
                oh.copy_(od2, non_blocking=True)
            torch.cuda.synchronize()
            floor.append((time.perf_counter() - t0) * 1e3)
        floor_ms = max_over_ranks(min(floor[1:]))
        del xd2, od2
        torch.cuda.empty_cache()
        for _ in range(2):
            plan.render_host(xh, F, HOP, out=oh)  # warm-up
        barrier()
        t0 = time.perf_counter()
        esteps = max(1, args.steps)
        for _ in range(esteps):
            plan.render_host(xh, F, HOP, out=oh)
        barrier()
        ems = max_over_ranks((time.perf_counter() - t0) * 1e3) / esteps
        line["e2e"] = {"value": world * S_e * F / (ems * 1e-3), "unit": "frames/s", "streams_per_gpu": S_e,
                       "h2d_bytes_per_step": lanes * L * 4,
                       "d2h_bytes_per_step": lanes * F * HOP * oh.element_size(), "ms_per_step": ems, "steps": esteps,
                       "out_dtype": str(oh.dtype).replace("torch.", ""), "arithmetic": args.precision,
                       "host_cpus_bound": len(bound) if bound else None,
                       "copy_floor_ms": floor_ms,
                       "copy_floor_gbs_per_gpu": (lanes * L * 4 + lanes * F * HOP * oh.element_size()) / floor_ms / 1e6,
                       "copy_floor_note": "the same pinned buffers copied up and down concurrently with nothing else, all "
                                          "ranks at once (max over ranks): what the host/PCIe side allows",
                       "api": "tsb_chain_render_host (pinned host buffers; upload + kernel + download inside the timed region)"}
        del oh, xh
        ctx.trim()  # the context keeps the staging rows between calls: release them for the legs below
    # ---- secondary figures (rank 0 describes them; all ranks run them to stay in step) ----
    if not args.no_extras:
        oms, _, _ = device_leg(other, x_all[:S_other * CH], max(1, args.steps), 2)
        ofps = world * S_other * F / (oms * 1e-3)
        line["other_precision"] = {"dtype": other, "value": ofps, "unit": "frames/s", "ms_per_step": oms,
                                   "streams_per_gpu": S_other,
                                   "roofline_frac": ofps / world * CH * b_alg(other) / 1e9 / peak}
        if world == 1:
            plan1 = tsb200.Plan(ctx, N_FRAME, precision=args.precision, program="pv", hop=HOP, rate=RATE, ratio=SHIFT)
            x1 = x_all[:CH].contiguous()
            o1 = torch.zeros((CH, F * HOP), dtype=plan1.out_dtype, device=dev)
            nseg = 296
            for _ in range(2):
                plan1.render_timepar(x1, F, HOP, nseg, out=o1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                plan1.render_timepar(x1, F, HOP, nseg, out=o1)
            e1.record()
            torch.cuda.synchronize()
            line["single_stream"] = {"value": F / (e0.elapsed_time(e1) / 3 * 1e-3), "unit": "frames/s",
                                     "note": f"BASELINE configs[0] as is: ONE stereo 60 s stream (2 lanes), rendered as {nseg} "
                                             "concurrent time segments (tsb_chain_render_timepar: pre-pass + scan + main; bit-identical to the sequential render)"}
    # ---- BASELINE configs[2] in small: ONE long 8-channel render split by time segment across the ranks ----
    if not args.no_extras and args.ts_seconds > 0:
        del x, x_all
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cfg3_timeshard
        line["time_sharded"] = cfg3_timeshard.run(dist, rank, world, dev, ctx, seconds=args.ts_seconds, precision="f64")
    # ---- BASELINE configs[4]: spectral bytebeat, frame 65536 (inverse FFT of 131072 points), 64 channels, fp64 ----
    if not args.no_extras:
        n5, ch5, f5 = 65536, 64, 24
        hop5 = n5 // 4
        out5 = torch.zeros((ch5, (f5 - 1) * hop5 + n5), dtype=torch.float64, device=dev)
        l0 = ctx.launches
        for _ in range(2):
            ctx.spectral_render(n5, ch5, f5, hop5, out=out5)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps5 = 3
        for _ in range(reps5):
            ctx.spectral_render(n5, ch5, f5, hop5, out=out5)
        e1.record()
        barrier()
        ms5 = max_over_ranks(e0.elapsed_time(e1)) / reps5
        fc5 = world * ch5 * f5 / (ms5 * 1e-3)
        balg5 = (n5 + 1) * 16 + hop5 * 8   # the Hermitian half spectrum + hop output samples
        line["cfg5"] = {"what": "spectral bytebeat pattern (hacks/spectral_bytebeat/*.cpp): seeded integer spectrum generated on the "
                                "device -> inverse real FFT of 131072 points -> triangular window -> overlap-add, fp64; every rank "
                                "renders its own 64 channels", "frame": n5, "fft_points": 2 * n5, "hop": hop5,
                        "channels_per_gpu": ch5, "frames": f5, "ms_per_render": ms5, "frames_per_s": fc5 / ch5,
                        "frame_channels_per_s": fc5, "out_samples_per_s": fc5 * hop5, "gpu_launches_per_render": (ctx.launches - l0) // (2 + reps5),
                        "roofline": {"bound": "hbm", "alg_bytes_per_frame_channel": balg5, "achieved": fc5 / world * balg5 / 1e9,
                                     "peak": peak, "unit": "GB/s", "frac": fc5 / world * balg5 / 1e9 / peak}}
        if world == 1:
            # the reference's own voidbubbles.cpp formula (libm sin / pow, integer bytebeat) as a spectrum expression
            # evaluated in the kernel (TSB_GEN_EXPR), same shape
            import tsb200
            prog = tsb200.expr.voidbubbles()["program"]
            for _ in range(2):
                ctx.spectral_render(n5, ch5, f5, hop5, out=out5, expr=prog)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps5):
                ctx.spectral_render(n5, ch5, f5, hop5, out=out5, expr=prog)
            e1.record()
            torch.cuda.synchronize()
            msx = e0.elapsed_time(e1) / reps5
            line["cfg5"]["voidbubbles_expr"] = {"what": "hacks/spectral_bytebeat/voidbubbles.cpp:9-13 as a TSB_GEN_EXPR program "
                                                        "(36 words, interpreted per bin in kernel A), same shape",
                                                "ms_per_render": msx, "frame_channels_per_s": ch5 * f5 / (msx * 1e-3)}
            # the same with the spectrum READ from HBM (TSB_GEN_BUFFER): the case SURVEY.md §8d's algorithmic bytes
            # describe literally (1.5 GB of spectra for the 24 frames: larger than L2)
            g5 = torch.Generator(device=dev)
            g5.manual_seed(5)
            spec5 = torch.view_as_complex(torch.rand((f5, ch5, n5, 2), generator=g5, device=dev, dtype=torch.float64) - 0.5)
            for _ in range(2):
                ctx.spectral_render(n5, ch5, f5, hop5, out=out5, spectrum=spec5)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps5):
                ctx.spectral_render(n5, ch5, f5, hop5, out=out5, spectrum=spec5)
            e1.record()
            torch.cuda.synchronize()
            msb = e0.elapsed_time(e1) / reps5
            fcb = ch5 * f5 / (msb * 1e-3)
            line["cfg5"]["spectrum_from_hbm"] = {"what": "caller's complex128 spectra resident in HBM instead of the in-kernel generator",
                                                 "ms_per_render": msb, "frame_channels_per_s": fcb,
                                                 "roofline_frac": fcb * balg5 / 1e9 / peak}
            del spec5
            bb = os.path.join(ROOT, "oracle", "_ref", "ref_bytebeat_tap_mkl")
            if os.path.exists(bb):
                try:
                    import tempfile
                    d = tempfile.mkdtemp(prefix="tsref")
                    env = dict(os.environ, TSREF_N=str(n5), TSREF_CH="2", TSREF_MAX_FRAMES="12", TSREF_DUMP_FIRST="0",
                               TSREF_TIMING="1", **ONE_THREAD)
                    o = subprocess.run([bb], env=env, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                       timeout=120, preexec_fn=lambda: os.sched_setaffinity(0, {host_cpus()[0]})).stdout
                    j = [json.loads(ln) for ln in o.splitlines() if ln.startswith("{")][-1]
                    line["cfg5"]["cpu_reference"] = {"frame_channels_per_s": 2 * j["frames"] / j["wall_s"], "cores": 1,
                                                     "sample": f"unmodified reference (oracle/_ref/ref_bytebeat_tap_mkl: the same "
                                                               f"seeded formula in the hack's host loop, MKL DFTI for FFTW), frame {n5}, "
                                                               f"2 channels, {j['frames']} frames, first->last callback wall"}
                except Exception as ex:  # the baseline is informative; the GPU figures stand without it
                    line["cfg5"]["cpu_reference"] = {"unavailable": str(ex)[:200]}
        del out5
    # ---- CPU baseline: rank 0, N=1 only, bounded sample ----
    if rank == 0 and world == 1:
        frames, wall, kind, what, used = run_reference(1, min(args.seconds, 60.0))
        line["cpu_baseline"] = {"value": frames / wall, "unit": "frames/s", "cores": used, "kind": kind,
                                "sample": f"1 process pinned to 1 core (1 FFT thread), 1 stereo stream x "
                                          f"{min(args.seconds, 60.0):g} s = {frames} frames, its own first->last "
                                          f"callback wall; {what}"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
