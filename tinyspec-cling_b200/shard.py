"""Sharding one box's GPUs over the path.

* by stream: lanes never interact, so rank r simply renders streams split_streams(S, world, r): no collective.
* by time segment, for one long render (BASELINE cfg3): rank r owns a contiguous block of frames.  What a block needs
  from before its first frame: the input halo, the per-bin phase carry and the overlap tail.  A rank
  r > 0 starts `lead` frames early — the frames of the block before whose output reaches into its first own frame,
  frame/hop - 1 of them for a constant hop — and one warm-up frame before those: the warm-up frame gives
  last_phase, the lead-in frames are rendered without emitting anything and so rebuild the overlap tail with the
  sequential render's own arithmetic (no tail message, no fix-up).  Its input range grows by those lead + 1 hops
  (the halo).  The carry is the running phase_sum itself: every rank pre-passes its block concurrently (storing each
  frame's increments), then the sum travels down the ranks — rank r receives the sum before its first lead-in frame,
  adds its block's increments one frame at a time (tsb_timepar_scan_at), noting the value before rank r+1's first
  lead-in frame, and passes that on — so that every block starts from exactly the state the one-GPU render has
  there.  The stitched output is the one-GPU render bit for bit.  The carry messages are [lanes, n/2] float64,
  latency-bound, over NCCL P2P (NVLink), lane by lane so that consecutive ranks overlap.

The device work sits behind a small backend interface (prepass / scan / render) so that the orchestration can be
exercised with gloo on CPU ranks against a replay of the kernel (tests/test_shard_gloo.py).
"""
import numpy as np


def split_streams(n_streams, world, rank):
    """Contiguous block of streams for `rank`: (first, count); counts differ by at most one."""
    base, rem = divmod(n_streams, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def plan_time_blocks(sch, n, world):
    """Cut a schedule (tsb200.schedule_const) into `world` contiguous frame blocks.  Per rank: its own frame range,
    the frames it starts early (lead-in) and the warm-up frame before them, the input range it must hold (halo
    included), where its first processed frame starts inside that range, its output range, and the frame of the
    PREVIOUS rank's block (block-relative, lead-in included) before which that rank notes the carry for it."""
    F = int(sch["n_frames"])
    in_off, out_off, hop = sch["in_off"], sch["out_off"], sch["hop"]
    blocks = []
    for r in range(world):
        f0, cnt = split_streams(F, world, r)
        if cnt == 0:
            blocks.append(dict(rank=r, first_frame=f0, n_frames=0, lead=0, first_proc=f0, in_lo=0, in_hi=0, in_off0=0,
                               out_lo=0, out_hi=0, out_pad=0, warm=False, warm_hop=0))
            continue
        f1 = f0 + cnt
        lead = 0
        while f0 - lead - 1 >= 1 and int(out_off[f0 - lead - 1]) + n > int(out_off[f0]):
            lead += 1  # (stops at frame 1: frame 0 stays the warm-up frame)
        if r == 0 or f0 == 0:
            lead = 0
        first = f0 - lead
        warm = first > 0
        warm_hop = int(hop[first - 1]) if warm else 0
        in_lo = int(in_off[first]) - warm_hop
        in_hi = int(in_off[f1 - 1]) + n
        blocks.append(dict(rank=r, first_frame=f0, n_frames=cnt, lead=lead, first_proc=first, in_lo=in_lo, in_hi=in_hi,
                           in_off0=int(in_off[first]) - in_lo, out_lo=int(out_off[f0]),
                           out_hi=int(out_off[f1 - 1]) + int(hop[f1 - 1]), out_pad=int(out_off[f0]) - int(out_off[first]),
                           warm=warm, warm_hop=warm_hop, hop_tab=np.ascontiguousarray(hop[first:f1], np.int32)))
    for r, b in enumerate(blocks):
        # where, inside this block (lead-in frames included), the next non-empty block's first processed frame lies
        nxt = next((c for c in blocks[r + 1:] if c["n_frames"]), None)
        b["carry_frame"] = -1 if (nxt is None or not b["n_frames"]) else nxt["first_proc"] - b["first_proc"]
    return blocks


def render_time_sharded(backend, x_block, block, dist, device_tensor=None):
    """One rank's part of a time-sharded render.  `x_block` holds input samples [block.in_lo, block.in_hi) of every
    lane; returns this rank's output samples [block.out_lo, block.out_hi).

    backend.prepass(x, block)                     analyses the block (lead-in included), keeps every frame's increments
    backend.scan(carry_in [k, n/2], lane0, k, carry_frame) -> the running phase_sum of those lanes before frame
                                                  carry_frame of the block (after its last frame when < 0)
    backend.render() -> out [lanes, out_hi-out_lo]
    `dist` is torch.distributed (nccl on GPUs, gloo in the CPU test); device_tensor(t) moves a numpy array or
    tensor to the tensor type the process group communicates (identity by default)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    to_t = device_tensor or (lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)))
    backend.prepass(x_block, block)
    if getattr(backend, "linked", False):
        # peer memory: every rank launches ONE scan kernel; the carries flow from GPU to GPU, CTA to CTA
        backend.scan_chain(block["carry_frame"])
        return backend.render()
    # the running phase_sum travels down the ranks: receive, add this block's increments in frame order, pass on
    # ... lane by lane (lanes are independent): while this rank adds lane c, the next one already works on lane c-1
    zero = backend.zero_carry()
    is_t = torch.is_tensor(zero)
    sends = []
    for c in range(zero.shape[0]):
        carry = to_t(zero[c:c + 1]).clone()
        if rank > 0:
            dist.recv(carry, rank - 1)
        carry_out = backend.scan(carry if is_t else carry.numpy(), c, 1, block["carry_frame"])
        if rank + 1 < world:
            sends.append(dist.isend(to_t(carry_out), rank + 1))
    for r in sends:
        r.wait()
    return backend.render()


class GpuBackend:
    """The device side of render_time_sharded on libtsb200 (tsb_timepar_*): every block is itself cut into
    `n_segments` concurrent time segments so that a handful of lanes fills the GPU."""

    def __init__(self, plan, n_segments):
        self.plan, self.n_segments = plan, n_segments
        self.linked = False

    def link(self, dist, lanes):
        """Map the next rank's carry buffers into this process (cudaIpc over NVLink) so that the phase carry travels
        from scan kernel to scan kernel without the host or NCCL (tsb_timepar_scan_chain).  Two buffer sets, used
        alternately, each [lanes][n/2] float64 + one int32 flag per 16 bins of a lane."""
        ctx, m = self.plan.ctx, self.plan.n // 2
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._carry_bytes = ((lanes * m * 8 + 255) // 256) * 256
        self._flag_bytes = ((lanes * ((m + 15) // 16) * 4 + 255) // 256) * 256
        self._set_bytes = self._carry_bytes + self._flag_bytes
        self._buf = ctx.malloc(2 * self._set_bytes)
        ctx.memset(self._buf, 0, 2 * self._set_bytes)
        ctx.sync()
        handles = [None] * self.world
        dist.all_gather_object(handles, ctx.ipc_export(self._buf))
        self._next = ctx.ipc_open(handles[self.rank + 1]) if self.rank + 1 < self.world else None
        self._epoch = 0
        dist.barrier()
        self.linked = True

    def scan_chain(self, carry_frame):
        self._epoch += 1
        off = (self._epoch & 1) * self._set_bytes
        mine = (self._buf + off, self._buf + off + self._carry_bytes) if self.rank > 0 else (None, None)
        peer = (self._next + off, self._next + off + self._carry_bytes) if self._next is not None else (None, None)
        self.tp.scan_chain(mine[0], mine[1], peer[0], peer[1], self._epoch, carry_frame)

    def prepass(self, x, block):
        import torch
        lanes, n = x.shape[0], self.plan.n
        self.block = block
        if getattr(self, "tp", None) is not None:
            self.tp.close()  # give the context's scratch block back BEFORE the next one is asked for (else the new
            self.tp = None   # block is a private multi-GB cudaMalloc every other render)
        self.state = dict(last_phase=torch.zeros((lanes, n // 2), dtype=torch.float64, device=x.device),
                          phase_sum=torch.zeros((lanes, n // 2), dtype=torch.float64, device=x.device),
                          ola=torch.zeros((lanes, n), dtype=torch.float64, device=x.device))
        ht = block["hop_tab"]
        # (the output row starts at the first lead-in frame's position: out_pad samples nobody writes, then the block's)
        self.tp = self.plan.timepar(x, int(ht.size), int(ht[0]), self.n_segments, in_off0=block["in_off0"], out_off0=0,
                                    hop_tab=ht, state=self.state, load_state=True, warmup=block["warm"],
                                    warm_hop=block["warm_hop"], lead=block["lead"])
        self.tp.prepass()

    def zero_carry(self):
        return self.state["phase_sum"].clone().zero_()

    def scan(self, carry, lane0=0, n_lanes=None, carry_frame=-1):
        return self.tp.scan(carry.contiguous(), lane0, n_lanes, carry_frame)

    def render(self):
        return self.tp.render()[:, self.block["out_pad"]:]
