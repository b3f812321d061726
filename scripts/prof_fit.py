import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import frank_oracle as fo
from frank_b200.geometry import FixedGeometry
from frank_b200.radial_fitters import FrankFitter
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
u, v, V, w, odht = fo.synthetic_disc(100000, N, analytic=True)
FF = FrankFitter(1.6, N, FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 2000, convergence_failure='ignore')
sol = FF.fit(u, v, V, w)
print('done', sol.MAP[:3])
