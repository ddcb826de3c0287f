"""tools/fft_bench.py — developer probe: batched forward FFT (tsb_fft_forward) throughput vs the HBM roofline.
Algorithmic bytes = read + write each element once = 32 B per complex128 point."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tsb200
ctx = tsb200.Context(0)
for n, rows in ((1024, 65536), (4096, 16384), (8192, 8192), (65536, 1024), (131072, 512), (262144, 256), (1800, 32768)):
    x = torch.randn((rows, n), dtype=torch.float64, device="cuda") + 0j
    y = torch.empty_like(x)
    for _ in range(2):
        ctx.fft_forward(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.fft_forward(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = rows * n * 32 / 1e9
    print(f"n={n:7d} rows={rows:6d}  {ms:8.3f} ms  {rows / ms * 1e3:12.0f} transforms/s  alg {gb / ms * 1e3:7.1f} GB/s "
          f"({gb / ms * 1e3 / 6545 * 100:5.1f}% of 6545)", flush=True)
