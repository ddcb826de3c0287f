"""Sharding one box's GPUs over the path.

* by stream: lanes never interact, so rank r simply renders streams split_streams(S, world, r): no collective.
* by time segment, for one long render (BASELINE cfg3): rank r owns a contiguous block of frames.  Three things
  cross a block boundary: the input halo (the block's first frame needs one warm-up frame before
  it, and its last frame reaches frame-hop samples past the block's own hops), the per-bin phase carry and the
  overlap tail (frame-hop samples that land on the next block's head).  The carry is the running phase_sum itself:
  every rank pre-passes its block concurrently (storing each frame's increments), then the sum travels down the
  ranks — rank r receives it, adds its block's increments one frame at a time (tsb_timepar_scan, ~0.1 ms) and
  passes it on — so that every block starts from exactly the state the sequential render has there; summing block
  totals first would regroup the additions and drift by ~ulp(phase_sum) per block and segment.  Carry and tail
  are [lanes, n/2] / [lanes, n] float64 neighbour messages, latency-bound, over NCCL (NVLink P2P).

The device work sits behind a small backend interface (prepass / scan / render / add_tail) so that the orchestration
can be exercised with gloo on CPU ranks against a replay of the kernel (tests/test_shard_gloo.py).
"""
import numpy as np


def split_streams(n_streams, world, rank):
    """Contiguous block of streams for `rank`: (first, count); counts differ by at most one."""
    base, rem = divmod(n_streams, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def plan_time_blocks(sch, n, world):
    """Cut a schedule (tsb200.schedule_const) into `world` contiguous frame blocks.  Per rank: frame range, the
    input range it must hold (halo included), where its first frame starts inside that range, its output range."""
    F = int(sch["n_frames"])
    blocks = []
    for r in range(world):
        f0, cnt = split_streams(F, world, r)
        if cnt == 0:
            blocks.append(dict(rank=r, first_frame=f0, n_frames=0, in_lo=0, in_hi=0, in_off0=0, out_lo=0, out_hi=0,
                               warm=False, warm_hop=0))
            continue
        f1 = f0 + cnt
        warm = f0 > 0
        warm_hop = int(sch["hop"][f0 - 1]) if warm else 0
        in_lo = int(sch["in_off"][f0]) - warm_hop
        in_hi = int(sch["in_off"][f1 - 1]) + n
        out_lo = int(sch["out_off"][f0])
        out_hi = int(sch["out_off"][f1 - 1]) + int(sch["hop"][f1 - 1])
        blocks.append(dict(rank=r, first_frame=f0, n_frames=cnt, in_lo=in_lo, in_hi=in_hi,
                           in_off0=int(sch["in_off"][f0]) - in_lo, out_lo=out_lo, out_hi=out_hi, warm=warm,
                           warm_hop=warm_hop))
    return blocks


def render_time_sharded(backend, x_block, block, dist, device_tensor=None):
    """One rank's part of a time-sharded render.  `x_block` holds input samples [block.in_lo, block.in_hi) of every
    lane; returns this rank's output samples [block.out_lo, block.out_hi).

    backend.prepass(x, n_frames, in_off0, warm)   analyses the block, keeps every frame's increments
    backend.scan(carry_in [k, n/2], lane0, k) -> carry_out: the running phase_sum of those lanes after the block
    backend.render() -> (out [lanes, out_hi-out_lo], tail [lanes, n])
    backend.add_tail(out, tail) -> out with the previous block's tail added at its head
    `dist` is torch.distributed (nccl on GPUs, gloo in the CPU test); device_tensor(t) moves a numpy array or
    tensor to the tensor type the process group communicates (identity by default)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    to_t = device_tensor or (lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)))
    backend.prepass(x_block, block["n_frames"], block["in_off0"], block["warm"])
    # the running phase_sum travels down the ranks: receive, add this block's increments in frame order, pass on
    # ... lane by lane (lanes are independent): while this rank adds lane c, the next one already works on lane c-1
    zero = backend.zero_carry()
    is_t = torch.is_tensor(zero)
    sends = []
    for c in range(zero.shape[0]):
        carry = to_t(zero[c:c + 1]).clone()
        if rank > 0:
            dist.recv(carry, rank - 1)
        carry_out = backend.scan(carry if is_t else carry.numpy(), c, 1)
        if rank + 1 < world:
            sends.append(dist.isend(to_t(carry_out), rank + 1))
    for r in sends:
        r.wait()
    out, tail = backend.render()
    tail_t = to_t(tail)
    reqs = []
    if rank + 1 < world:
        reqs.append(dist.isend(tail_t, rank + 1))
    if rank > 0:
        prev = torch.empty_like(tail_t)
        dist.recv(prev, rank - 1)
        out = backend.add_tail(out, prev if torch.is_tensor(tail) else prev.numpy())
    for r in reqs:
        r.wait()
    return out


class GpuBackend:
    """The device side of render_time_sharded on libtsb200 (tsb_timepar_*): every block is itself cut into
    `n_segments` concurrent time segments so that a handful of lanes fills the GPU."""

    def __init__(self, plan, hop, n_segments, hop_tab=None):
        self.plan, self.hop, self.n_segments, self.hop_tab = plan, hop, n_segments, hop_tab

    def prepass(self, x, n_frames, in_off0, warm):
        import torch
        lanes, n = x.shape[0], self.plan.n
        self.state = dict(last_phase=torch.zeros((lanes, n // 2), dtype=torch.float64, device=x.device),
                          phase_sum=torch.zeros((lanes, n // 2), dtype=torch.float64, device=x.device),
                          ola=torch.zeros((lanes, n), dtype=torch.float64, device=x.device))
        self.tp = self.plan.timepar(x, n_frames, self.hop, self.n_segments, in_off0=in_off0, out_off0=0,
                                    hop_tab=self.hop_tab, state=self.state, load_state=True, warmup=warm,
                                    warm_hop=self.hop)
        self.tp.prepass()

    def zero_carry(self):
        return self.state["phase_sum"].clone().zero_()

    def scan(self, carry, lane0=0, n_lanes=None):
        return self.tp.scan(carry.contiguous(), lane0, n_lanes)

    def render(self):
        out = self.tp.render()
        return out, self.state["ola"]

    def add_tail(self, out, tail):
        self.plan.add_tail(out, 0, tail.contiguous())
        return out
