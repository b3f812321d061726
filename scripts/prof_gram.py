"""Run the mapping once on synthetic visibilities already resident on the GPU (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.geometry import FixedGeometry
from frank_b200.statistical_models import VisibilityMapping
from frank_b200.constants import rad_to_arcsec
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
debris = len(sys.argv) > 4 and sys.argv[4] == 'debris'
g = FixedGeometry(30., 40., 1e-3, -2e-3)
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
gen = torch.Generator(device='cuda').manual_seed(1)
q = 0.98 * dht.q[-1] * torch.sqrt(torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
th = 2 * np.pi * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen)
ud, vd = q * torch.cos(th), q * torch.sin(th)
u, v = g.reproject(ud, vd) if False else (ud / np.cos(np.deg2rad(30.)) * np.cos(np.deg2rad(40.)) + vd * np.sin(np.deg2rad(40.)),
                                          -ud / np.cos(np.deg2rad(30.)) * np.sin(np.deg2rad(40.)) + vd * np.cos(np.deg2rad(40.)))
V = torch.complex(torch.exp(-(q / 1e6) ** 2), torch.zeros_like(q))
w = 1e4 * (0.5 + 1.5 * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
vm = VisibilityMapping(dht, g, verbose=False, check_qbounds=False, **(dict(vis_model='debris', scale_height=lambda R: 0.05 * R) if debris else {}))
for _ in range(reps):
    m = vm.map_visibilities(u, v, V, w)
    print(vm.last_timing, n * N / vm.last_timing['gram_ms'] / 1e6, 'Gvis.mode/s')
