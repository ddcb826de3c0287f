"""tools/cfg5_bench.py — developer probe: time the fused spectral synthesis (tsb_spectral_render, BASELINE cfg5)
at a few shapes with CUDA events.  usage: python tools/cfg5_bench.py [frame] [channels] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tsb200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 24
hop = n // 4
ctx = tsb200.Context(0)
out = torch.zeros((ch, (frames - 1) * hop + n), dtype=torch.float64, device="cuda")
for _ in range(2):
    ctx.spectral_render(n, ch, frames, hop, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    ctx.spectral_render(n, ch, frames, hop, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fc = ch * frames / (ms * 1e-3)
balg = (n + 1) * 16 + hop * 8
def report(tag, ms):
    fc = ch * frames / (ms * 1e-3)
    print(f"cfg5 {tag} frame={n} ch={ch} frames={frames}: {ms:.3f} ms  {ms / frames * 1e3:.1f} us/frame  {fc / 1e6:.3f} M frame-ch/s  "
          f"alg {fc * balg / 1e9:.1f} GB/s ({fc * balg / 1e9 / 6545 * 100:.1f}% of 6545)")


# the same with the spectrum read from HBM (TSB_GEN_BUFFER): what SURVEY.md §8d's algorithmic bytes describe
if len(sys.argv) <= 4 or sys.argv[4] != "nobuf":
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    spec = torch.view_as_complex(torch.rand((frames, ch, n, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5)
    for _ in range(2):
        ctx.spectral_render(n, ch, frames, hop, out=out, spectrum=spec)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ctx.spectral_render(n, ch, frames, hop, out=out, spectrum=spec)
    e1.record()
    torch.cuda.synchronize()
    report("spectrum-from-HBM", e0.elapsed_time(e1) / reps)
print(f"cfg5 generator frame={n} ch={ch} frames={frames}: {ms:.3f} ms  {ms / frames * 1e3:.1f} us/frame  {fc / 1e6:.3f} M frame-ch/s  "
      f"alg {fc * balg / 1e9:.1f} GB/s ({fc * balg / 1e9 / 6545 * 100:.1f}% of 6545)  "
      f"{fc * 5 * n * 16 / 1e12:.2f} TFLOP/s (5 n log2 n of the n-point complex FFT)")
