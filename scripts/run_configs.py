"""Run the BASELINE.json configurations that fit one GPU and print one JSON line each (results table of BASELINE.md)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.geometry import FixedGeometry
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.radial_fitters import FrankFitter
from frank_b200.debris_fitters import FrankDebrisFitter
from frank_b200.statistical_models import VisibilityMapping
from frank_b200.filter import CriticalFilter
from frank_b200.utilities import UVDataBinner
import bench

which = sys.argv[1:] or ['1', '2', '3', '4', '5']
geom = FixedGeometry(*bench.GEOM)


def data(n, N, seed=1):
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    return dht, bench.synthetic_visibilities_device(n, dht, seed)


def timed_fit(FF, u, v, V, w, reps=2):
    out = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pre = FF.preprocess_visibilities(u, v, V, w)
        t1 = time.perf_counter()
        sol = FF.fit_preprocessed(pre)
        t2 = time.perf_counter()
        out = {'map_s': t1 - t0, 'solver_s': t2 - t1, 'fit_s': t2 - t0, 'iterations': int(FF.iteration_diagnostics['num_iterations']),
               'gram_ms': FF._vis_map.last_timing['gram_ms']}
    return out, sol


if '1' in which or '2' in which:
    for tag, n in [('1', 1_000_000), ('2', 10_000_000)]:
        if tag not in which:
            continue
        dht, (u, v, V, w) = data(n, 300)
        FF = FrankFitter(1.6, 300, geom, alpha=1.05, weights_smooth=1e-4, verbose=False, store_iteration_diagnostics=True)
        r, sol = timed_fit(FF, u, v, V, w)
        r.update(config=tag, n_vis=n, N=300, method='Normal', gvis_mode_per_s=n * 300 / r['gram_ms'] / 1e6)
        print(json.dumps(r), flush=True)

if '3' in which:
    n, N = 10_000_000, 500
    dht, (u, v, V, w) = data(n, N)
    vm = VisibilityMapping(dht, geom, verbose=False)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); m = vm.map_visibilities(u, v, V, w); t1 = time.perf_counter()
    r = {'config': '3 mapping (N=500, 1e7 visibilities)', 'n_vis': n, 'N': N,
         'map_s': t1 - t0, 'gram_ms': vm.last_timing['gram_ms'], 'gvis_mode_per_s': n * N / vm.last_timing['gram_ms'] / 1e6}
    print(json.dumps(r), flush=True)
    # LogNormal fit (device-resident loop) on a size where the reference's log-normal path is numerically well defined
    dht2, (u2, v2, V2, w2) = data(1_000_000, 100)
    FL = FrankFitter(1.6, 100, geom, alpha=1.3, weights_smooth=1e-2, method='LogNormal', verbose=False, store_iteration_diagnostics=True)
    t0 = time.perf_counter(); sol = FL.fit(u2, v2, V2, w2); t1 = time.perf_counter()
    print(json.dumps({'config': '3b LogNormal N=100, 1e6 vis', 'fit_s': t1 - t0, 'iterations': int(FL.iteration_diagnostics['num_iterations']),
                      'newton': sol._fit._status}), flush=True)

if '3full' in which:
    # BASELINE.json configs[2] end to end: method='LogNormal', N=500, 1e7 visibilities (alpha=1.3, wsmooth=1e-2)
    n, N = 10_000_000, 500
    dht, (u, v, V, w) = data(n, N)
    FL = FrankFitter(1.6, N, geom, alpha=1.3, weights_smooth=1e-2, method='LogNormal', verbose=False, store_iteration_diagnostics=True,
                     convergence_failure='warn')
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pre = FL.preprocess_visibilities(u, v, V, w)
    t1 = time.perf_counter()
    r = {'config': '3 (LogNormal MAP fit, N=500, 1e7 visibilities)', 'n_vis': n, 'N': N, 'map_s': t1 - t0,
         'gram_ms': FL._vis_map.last_timing['gram_ms']}
    try:
        sol = FL.fit_preprocessed(pre)
        t2 = time.perf_counter()
        r.update(solver_s=t2 - t1, fit_s=t2 - t0, iterations=int(FL.iteration_diagnostics['num_iterations']), newton=sol._fit._status)
    except Exception as e:      # at Rmax = 1.6" the reference itself aborts at this N (tests/golden/make_golden.py gen_config3)
        r.update(solver_s=time.perf_counter() - t1, outcome=type(e).__name__ + ': ' + str(e)[:80])
    print(json.dumps(r), flush=True)
    # the same shape at Rmax = 1.0" (where the reference completes, fixture config3_lognormal_N500): solver time on the fixture's M, j
    g3 = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'config3_lognormal_N500.npz'))
    Mref = np.zeros((N, N)); Mref[np.triu_indices(N)] = g3['M_upper']; Mref = Mref + np.triu(Mref, 1).T
    FL = FrankFitter(1.0, N, geom, alpha=1.3, weights_smooth=1e-2, method='LogNormal', verbose=False, store_iteration_diagnostics=True)
    pre = {'M': Mref, 'j': g3['j'], 'null_likelihood': float(g3['H0']), 'hash': [False, FL._DHT, geom, 'opt_thick', None]}
    for rep in range(2):
        t1 = time.perf_counter()
        try:
            sol = FL.fit_preprocessed(pre)
            r = {'config': '3 solver on the reference fixture (LogNormal, N=500, Rmax=1.0)', 'solver_s': time.perf_counter() - t1,
                 'iterations': int(FL.iteration_diagnostics['num_iterations']), 'reference_iterations': int(g3['num_iterations']),
                 'newton': sol._fit._status}
        except Exception as e:
            r = {'config': '3 solver on the reference fixture (LogNormal, N=500, Rmax=1.0)', 'solver_s': time.perf_counter() - t1,
                 'outcome': type(e).__name__ + ': ' + str(e)[:80]}
    print(json.dumps(r), flush=True)

if '4' in which:
    n, N = 1_000_000, 300
    dht, (u, v, V, w) = data(n, N)
    FF = FrankFitter(1.6, N, geom, verbose=False, convergence_failure='ignore')
    pre = FF.preprocess_visibilities(u, v, V, w)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sols = FF.fit_sweep_preprocessed(pre, alphas=np.linspace(1.01, 1.5, 8), weights_smooths=np.logspace(-4, -1, 8), on_cholesky_failure='flag')
        t1 = time.perf_counter()
    it = np.array(FF.sweep_diagnostics['num_iterations'])
    print(json.dumps({'config': '4 (FrankFitter.fit_sweep, one rank)', 'grid_points': len(sols), 'N': N, 'sweep_s': t1 - t0,
                      'iterations_min_max_sum': [int(it.min()), int(it.max()), int(it.sum())],
                      'us_per_point_iteration': (t1 - t0) / it.sum() * 1e6, 'converged': int(np.sum(FF.sweep_diagnostics['converged'])),
                      'cholesky_failed_points': FF.sweep_diagnostics['cholesky_failed_points_this_rank']}), flush=True)

if '5' in which:
    n, N = int(os.environ.get('CFG5_NVIS', 100_000_000)), 2000
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    u, v, V, w = bench.synthetic_visibilities_device(n, dht, 5)
    vm = VisibilityMapping(dht, geom, vis_model='debris', scale_height=lambda r: 0.05 * r, verbose=False)
    freqs = torch.randint(0, 4, (n,), device='cuda').to(torch.float64)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = vm.map_visibilities(u, v, V, w, frequencies=freqs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    r = {'config': '5 (one device call, channel in the sort key)', 'n_vis': n, 'N': N, 'channels': 4, 'vis_model': 'debris', 'map_s': t1 - t0,
         'gvis_mode_per_s': n * N / (t1 - t0) / 1e9, 'timing_ms': {k: float(x) for k, x in vm.last_timing.items()}}
    print(json.dumps(r), flush=True)
    # on-GPU deprojection + uv binning of all visibilities, device resident (apply_correction -> q -> UVDataBinner)
    del m
    ctx = _lib.get_context()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        up, vp, wp, Vp, gq = ctx.apply_correction_dev(u, v, V, geom.device_scalars(), want_q=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        b = UVDataBinner(gq, Vp, w, 1e3)
        t2 = time.perf_counter()
        del up, vp, wp
    print(json.dumps({'config': '5 deprojection + binning (device resident)', 'n_vis': n, 'bin_width': 1e3, 'nbins': len(b),
                      'apply_correction_s': t1 - t0, 'apply_correction_gbs': 80 * n / (t1 - t0) / 1e9, 'bin_s': t2 - t1,
                      'bin_gbs_algorithmic': 72 * n / (t2 - t1) / 1e9}), flush=True)
