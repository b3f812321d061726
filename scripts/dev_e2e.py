"""End-to-end timing of the host entry point's copy / compute pipeline for a few chunking settings (dev tool).
    python scripts/dev_e2e.py [n_vis] [N]"""
import sys
import time
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.geometry import FixedGeometry
from frank_b200.hankel import DiscreteHankelTransform
import bench

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
geom = FixedGeometry(*bench.GEOM)
dht = DiscreteHankelTransform(bench.RMAX / rad_to_arcsec, N)
ctx = _lib.get_context(0)
ctx.dht_setup(dht)
u, v, V, w = bench.synthetic_visibilities_device(n, dht, seed=1)
Vr = torch.view_as_real(V).contiguous()
gdev = geom.device_scalars()
out = torch.zeros(N * N + N + 1, dtype=torch.float64, device='cuda')
q_last, scale = float(dht.q[-1]), float(np.cos(np.deg2rad(bench.GEOM[0])))
for _ in range(3):
    ctx.map_visibilities(n, u, v, Vr, w, 1, gdev, 0, scale, None, True, q_last, out[:N * N], out[N * N:N * N + N], out[N * N + N:], host=False)
ctx.timer_start()
for _ in range(5):
    ctx.map_visibilities(n, u, v, Vr, w, 1, gdev, 0, scale, None, True, q_last, out[:N * N], out[N * N:N * N + N], out[N * N + N:], host=False)
print(f"device-resident step: {ctx.timer_stop() / 5:.2f} ms  {ctx.last_map_timing()}")
pin = [x.cpu().pin_memory() for x in (u, v, Vr, w)]
page = [np.array(x.numpy()) for x in pin]
hout = np.zeros(N * N + N + 1)
for kind, arrs in (('pinned', pin), ('pageable', page)):
    for growth, kmax, chunk, thr in [(2.0, 4, 250000, 8), (3.0, 4, 250000, 8), (1.5, 6, 250000, 8), (1.0, 8, 250000, 8), (1.0, 4, 250000, 8), (1.0, 2, 250000, 8), (1.0, 1, 250000, 8)]:
        ctx.set_option('map_growth', growth); ctx.set_option('map_kmax', kmax); ctx.set_option('map_chunk', chunk)
        ctx.set_option('stage_threads', thr)
        for _ in range(2):
            ctx.map_visibilities(n, *arrs, 1, gdev, 0, scale, None, True, q_last, hout[:N * N], hout[N * N:N * N + N], hout[N * N + N:], host=True)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            ctx.map_visibilities(n, *arrs, 1, gdev, 0, scale, None, True, q_last, hout[:N * N], hout[N * N:N * N + N], hout[N * N + N:], host=True)
        dt = (time.perf_counter() - t0) / reps * 1e3
        tm = ctx.last_map_timing()
        print(f"{kind:9s} growth={growth} kmax={kmax} threads={thr}: {dt:.2f} ms e2e   gram {tm['gram_ms']:.2f} prep {tm['prep_ms']:.2f} fin {tm['finalize_ms']:.2f} copy {tm['copy_ms']:.2f}", flush=True)
