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
    assert blocks[0]["in_lo"] == 2048 and not blocks[0]["warm"]
    for a, b in zip(blocks, blocks[1:]):
        assert a["out_hi"] == b["out_lo"] and b["warm"]
        assert b["in_lo"] == sch["in_off"][b["first_frame"]] - HOP          # one warm-up frame of halo
        assert a["in_hi"] - sch["in_off"][b["first_frame"]] == N - HOP      # frame - hop samples of overlap


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    x = orc.gen_signal(0, CH, SAMPLES).T
    sch = tsb200.schedule_const(SAMPLES, N, HOP)
    b = shard.plan_time_blocks(sch, N, world)[rank]
    backend = emul.ReplayBackend(N, HOP, RATIO)
    out = shard.render_time_sharded(backend, np.ascontiguousarray(x[:, b["in_lo"]:b["in_hi"]]), b, dist)
    q.put((rank, b["out_lo"], np.asarray(out)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_time_sharded_render_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([p[2] for p in parts], axis=1)
    ref = orc.render_pv(orc.gen_signal(0, CH, SAMPLES), N, HOP / N, RATIO)
    want = ref["ola64"][:got.shape[1]].T
    assert got.shape[1] == ref["frames"] * HOP
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    # The phase carry travels down the ranks as the exact running sum, so every block starts from the state the
    # one-process render has there: the samples are identical except where a block's overlap tail was added
    # afterwards (the same addends in another order).
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
    import emul
    seq, _ = emul.chain_render(orc.gen_signal(0, CH, SAMPLES).T, N, HOP, ref["frames"], precision="f64", ratio=RATIO)
    same = got == seq
    assert same.mean() > 1 - 1.01 * (world - 1) * N / got.shape[1]
    assert np.abs(got - seq).max() <= 4 * np.finfo(np.float64).eps * np.abs(seq).max()
