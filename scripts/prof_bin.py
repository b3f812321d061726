"""uv-binning pass alone, for `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
(run under gpurun): 1e7 device-resident visibilities, bin width 1e3 lambda, three calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frank_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
gen = torch.Generator(device='cuda').manual_seed(7)
uv = 1.9e7 * torch.sqrt(torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
V = torch.complex(torch.randn(n, device='cuda', dtype=torch.float64, generator=gen),
                  torch.randn(n, device='cuda', dtype=torch.float64, generator=gen))
w = 1e4 * (0.5 + 1.5 * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
width = 1e3
uv_max = float(uv.max().item())
nbins = int(np.ceil(uv_max / width)); nbins += int(nbins * width < uv_max)
ctx = _lib.get_context(0)
for _ in range(3):
    idx, counts, sums, err = ctx.uv_bin_dev(uv, V, w, width, nbins)
ctx.timer_start()
for _ in range(5):
    ctx.uv_bin_dev(uv, V, w, width, nbins)
ms = ctx.timer_stop() / 5
print(f'n={n} nbins={nbins} counts.sum={int(counts.sum())} {ms:.3f} ms per call, {72 * n / ms / 1e6:.1f} GB/s algorithmic (72 B/vis)')
