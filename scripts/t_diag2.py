import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import frank_oracle as fo
from frank_b200.geometry import FixedGeometry
from frank_b200.radial_fitters import FrankFitter
u, v, V, w, odht = fo.synthetic_disc(200000, 300, analytic=True)
def run(tag):
    FF = FrankFitter(1.6, 300, FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, store_iteration_diagnostics=True)
    pre = FF.preprocess_visibilities(u, v, V, w)
    t = time.time(); sol = FF.fit_preprocessed(pre); dt = time.time() - t
    print(tag, f'{dt*1e3:.1f} ms', FF.iteration_diagnostics['num_iterations'], flush=True)
run('plain')
import torch
torch.cuda.set_device(0); x = torch.zeros(10, device='cuda'); torch.cuda.synchronize()
run('after torch cuda init')
big = torch.zeros(50_000_000, dtype=torch.float64, device='cuda'); torch.cuda.synchronize()
run('after 400MB torch alloc')
hp = torch.zeros(50_000_000, dtype=torch.float64).pin_memory()
run('after 400MB pinned alloc')
import threading, subprocess
def smi():
    for _ in range(20):
        subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader'], capture_output=True); time.sleep(0.2)
th = threading.Thread(target=smi); th.start()
run('with nvidia-smi polling')
th.join()
run('after polling')
