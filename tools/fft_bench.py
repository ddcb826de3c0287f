"""tools/fft_bench.py — developer probe: batched forward FFT (tsb_fft_forward) throughput vs the HBM roofline, with
cuFFT timed beside it AS A COMPARISON ONLY (cuFFT is not on the product path — it is reached here
through torch.fft.fft, which is cuFFT's Z2Z plan).  Algorithmic bytes = read + write each element once = 32 B per
complex128 point.  Prints one line per shape and a JSON summary (kept under profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsb200

ctx = tsb200.Context(0)
PEAK = 6545.0
rows_out = []


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n, rows in ((1024, 65536), (4096, 16384), (8192, 8192), (65536, 1024), (131072, 512), (262144, 256), (1800, 32768)):
    x = torch.randn((rows, n), dtype=torch.float64, device="cuda") + 0j
    y = torch.empty_like(x)
    ms = timed(lambda: ctx.fft_forward(x, out=y))
    ms_cufft = timed(lambda: torch.fft.fft(x, dim=-1, out=y))
    ref = torch.fft.fft(x, dim=-1)
    ctx.fft_forward(x, out=y)
    torch.cuda.synchronize()
    err = float((y - ref).abs().max() / ref.abs().max())
    gb = rows * n * 32 / 1e9
    rows_out.append({"n": n, "rows": rows, "tsb_ms": ms, "tsb_gbs": gb / ms * 1e3, "tsb_frac_hbm": gb / ms * 1e3 / PEAK,
                     "cufft_ms": ms_cufft, "cufft_gbs": gb / ms_cufft * 1e3, "cufft_frac_hbm": gb / ms_cufft * 1e3 / PEAK,
                     "tsb_over_cufft_time": ms / ms_cufft, "max_rel_diff": err})
    print(f"n={n:7d} rows={rows:6d}  tsb {ms:8.3f} ms {gb / ms * 1e3:7.1f} GB/s ({gb / ms * 1e3 / PEAK * 100:5.1f}%)   "
          f"cuFFT {ms_cufft:8.3f} ms {gb / ms_cufft * 1e3:7.1f} GB/s ({gb / ms_cufft * 1e3 / PEAK * 100:5.1f}%)   "
          f"tsb/cuFFT time {ms / ms_cufft:5.2f}   max rel diff {err:.1e}", flush=True)
    del x, y, ref
print(json.dumps({"what": "batched complex128 forward FFT: tsb_fft_forward vs cuFFT (torch.fft.fft), comparison only",
                  "hbm_peak_gbs": PEAK, "shapes": rows_out}))
