import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import frank_oracle as fo
from frank_b200.geometry import FixedGeometry
from frank_b200.radial_fitters import FrankFitter
u, v, V, w, odht = fo.synthetic_disc(200000, 300, analytic=True)
for diag in [False, True, False, True]:
    FF = FrankFitter(1.6, 300, FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, store_iteration_diagnostics=diag)
    pre = FF.preprocess_visibilities(u, v, V, w)
    t = time.time(); sol = FF.fit_preprocessed(pre); dt = time.time() - t
    print('diag', diag, f'{dt*1e3:.1f} ms')
