"""The N>1 path on CPU: two and three gloo ranks run the time-sharded render orchestration (shard.py: halo, the
running phase_sum handed down the ranks, overlap-tail send/recv) over a CPU replay of the kernel, and the stitched
result is checked against the oracle and the one-process render.  Also the stream split and block planning."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tsb200  # noqa: F401  (registers tinyspec_cling_b200)
from tinyspec_cling_b200 import shard
from oracle import orc

N, HOP, RATIO, SAMPLES, CH = 256, 64, 1.5, 9000, 2


def test_split_streams_covers_everything():
    for S in (0, 1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            got = [shard.split_streams(S, world, r) for r in range(world)]
            assert sum(c for _, c in got) == S
            assert all(got[r][0] + got[r][1] == got[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_plan_time_blocks_ranges():
    sch = tsb200.schedule_const(SAMPLES, N, HOP)
    blocks = shard.plan_time_blocks(sch, N, 3)
    assert sum(b["n_frames"] for b in blocks) == sch["n_frames"]
    assert blocks[0]["in_lo"] == 2048 and not blocks[0]["warm"] and blocks[0]["lead"] == 0
    for a, b in zip(blocks, blocks[1:]):
        assert a["out_hi"] == b["out_lo"] and b["warm"]
        assert b["lead"] == N // HOP - 1                                   # the frames whose output reaches into the block
        first = b["first_frame"] - b["lead"]
        assert b["in_lo"] == sch["in_off"][first] - HOP                     # ... and one warm-up frame before them: the halo
        assert b["out_pad"] == b["lead"] * HOP and b["hop_tab"].size == b["lead"] + b["n_frames"]
        assert a["carry_frame"] == a["lead"] + a["n_frames"] - b["lead"]    # where the previous rank notes the carry
    assert blocks[-1]["carry_frame"] == -1


def _worker(rank, world, port, q, case):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    n, hop_d, ratio, samples = case
    x = orc.gen_signal(0, CH, samples).T
    sch = tsb200.schedule_const(samples, n, hop_d)
    b = shard.plan_time_blocks(sch, n, world)[rank]
    backend = emul.ReplayBackend(n, hop_d, ratio)
    out = shard.render_time_sharded(backend, np.ascontiguousarray(x[:, b["in_lo"]:b["in_hi"]]), b, dist)
    q.put((rank, b["out_lo"], np.asarray(out)))
    dist.barrier()
    dist.destroy_process_group()


# (frame, hop, ratio, samples): cfg3's shape in small, and a fractional hop (computed_hop alternates 51 / 52:
# every block's warm-up hop, lead-in frames and hop table are its own)
CASES = {"const": (N, float(HOP), RATIO, SAMPLES), "fractional": (256, 51.3, 1.5, 9000)}


@pytest.mark.parametrize("world,case", [(2, "const"), (3, "const"), (3, "fractional")])
def test_time_sharded_render_gloo(world, case):
    n, hop_d, ratio, samples = CASES[case]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, CASES[case])) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([p[2] for p in parts], axis=1)
    x = orc.gen_signal(0, CH, samples)
    ref = orc.render_pv(x, n, hop_d / n, ratio, want_tables=True)
    want = ref["ola64"][:got.shape[1]].T
    assert got.shape[1] == int(ref["tables"]["hop"].sum())
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    # Every block starts from the state the one-process render has there (the exact running phase_sum handed down
    # the ranks, last_phase from its warm-up frame, the overlap tail rebuilt by its lead-in frames): the stitched
    # output IS the one-process render.
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    seq, _ = emul.chain_render(x.T, n, int(ref["tables"]["hop"][0]), ref["frames"], precision="f64", ratio=ratio, hop_d=hop_d,
                               hop_tab=ref["tables"]["hop"])
    assert np.array_equal(got, seq)
