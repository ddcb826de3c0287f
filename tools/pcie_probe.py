"""tools/pcie_probe.py — developer probe: pinned host<->device copy bandwidth, 1-D vs row-strided 2-D, one
direction vs both at once (what bounds bench.py's e2e leg)."""
import time
import torch

dev = torch.device("cuda", 0)
lanes, L = 1036, 2880000
rows_w = 180000  # one time block of a row
h_in = torch.empty((lanes, L), dtype=torch.float32, pin_memory=True)
h_out = torch.empty((lanes, L), dtype=torch.float32, pin_memory=True)
d_a = torch.empty((lanes, L), dtype=torch.float32, device=dev)
d_blk = torch.empty((lanes, rows_w), dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, nbytes, label):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:55s} {nbytes / dt / 1e9:7.1f} GB/s  ({dt * 1e3:.1f} ms)", flush=True)


for _ in range(2):
    timeit(lambda: d_a.copy_(h_in, non_blocking=True), h_in.numel() * 4, "H2D 1-D whole (11.9 GB)")
    timeit(lambda: h_out.copy_(d_a, non_blocking=True), h_in.numel() * 4, "D2H 1-D whole")


def both():
    with torch.cuda.stream(s1):
        d_a[:lanes // 2].copy_(h_in[:lanes // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        h_out[lanes // 2:].copy_(d_a[lanes // 2:], non_blocking=True)


timeit(both, h_in.numel() * 4, "H2D half + D2H half concurrently (1-D)")
timeit(lambda: d_blk.copy_(h_in[:, :rows_w], non_blocking=True), lanes * rows_w * 4, "H2D 2-D block (1036 rows x 720 KB)")
timeit(lambda: h_out[:, :rows_w].copy_(d_blk, non_blocking=True), lanes * rows_w * 4, "D2H 2-D block")


def both2d():
    with torch.cuda.stream(s1):
        d_blk.copy_(h_in[:, :rows_w], non_blocking=True)
    with torch.cuda.stream(s2):
        h_out[:, rows_w:2 * rows_w].copy_(d_a[:, rows_w:2 * rows_w], non_blocking=True)


timeit(both2d, 2 * lanes * rows_w * 4, "H2D 2-D + D2H 2-D concurrently")
