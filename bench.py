#!/usr/bin/env python
"""Benchmark of the hot path: visibilities -> GP normal equations (H^T W H, H^T W V).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload config2|config5]

A "step" is one pass of `VisibilityMapping.map_visibilities` (geometry pre-pass + baseline sort + fused J0/Gram kernel +
split-K reduction, and with N > 1 ranks the library's in-call NCCL all-reduce) over one batch of synthetic visibilities.
`value` is Gvis.mode/s = n_vis * N_modes / t / 1e9 with the inputs already resident in HBM; `e2e` is the same metric
through the C ABI with HOST buffers (pinned; a pageable run is reported next to it), host->device and device->host copies
inside the timed region.

  --workload config2 (default)  BASELINE.json configs[1]: Normal-fit mapping, 1e7 unbinned visibilities per GPU, N = 300
                                (weak scaling: every rank maps its own 1e7 visibilities).
  --workload config5            BASELINE.json configs[4]: N = 2000, 1e8 visibilities IN TOTAL in 4 frequency channels, debris
                                scale-height factor (strong scaling: 1e8 / N_gpus visibilities per rank).

--impl reference times the CPU restatement of the reference's NumPy/SciPy path (oracle/frank_oracle.py -- the reference
itself is pure Python and is not present on the GPU box) on a bounded sample of the same workload, on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RMAX = 1.6
GEOM = (30., 40., 1e-3, -2e-3)
WORKLOADS = {
    'config2': dict(n_total=None, n_per_gpu=10_000_000, N=300, nchan=1, model='opt_thick', scaling='weak',
                    name='Normal fit mapping, {n:.0e} unbinned visibilities per GPU, N=300 (BASELINE.json configs[1])'),
    'config5': dict(n_total=100_000_000, n_per_gpu=None, N=2000, nchan=4, model='debris', scaling='strong',
                    name='large-scale mapping, 1e8 visibilities in total over the GPUs, N=2000, 4 channels, debris scale height '
                         '(BASELINE.json configs[4])'),
}
# measured on this pool's B200 with probes/fp64_probe.cu (profiles/r01_fp64_probe.txt): raw mma.sync m8n8k4 f64
# issue rate; MEASURED_PEAKS.json carries no FP64 figure (cuBLAS DGEMM on the same box: 35.46 TFLOP/s)
FP64_DMMA_PEAK_TFLOPS = 37.1
FP64_DGEMM_TFLOPS = 35.46


def gram_traffic(n, N):
    """DRAM bytes of one k_gram launch from the committed ncu --set full capture of this workload, else None."""
    for name in ('r02_gram_traffic.json', 'r01_gram_traffic.json'):
        try:
            with open(os.path.join(ROOT, 'profiles', name)) as fh:
                t = json.load(fh)
            if t['n_vis'] == n and t['N'] == N:
                return t['dram_bytes_read'] + t['dram_bytes_write'], name
        except Exception:
            pass
    return None, None


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def synthetic_visibilities_device(n, dht, seed):
    """Synthetic Gaussian-ring-like visibilities of the BASELINE.md shape, generated on the GPU with torch
    (closed-form Gaussian Hankel pair: the cost of the path does not depend on the values)."""
    import torch
    gen = torch.Generator(device='cuda').manual_seed(seed)
    inc, PA = np.deg2rad(GEOM[0]), np.deg2rad(GEOM[1])
    q = 0.98 * dht.q[-1] * torch.sqrt(torch.rand(n, device='cuda', dtype=torch.float64, generator=gen) * (1 - 1e-5) + 1e-5)
    th = 2 * np.pi * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen)
    ud, vd = q * torch.cos(th) / np.cos(inc), q * torch.sin(th)       # re-project (frank/geometry.py:115-127)
    u = ud * np.cos(PA) + vd * np.sin(PA)
    v = -ud * np.sin(PA) + vd * np.cos(PA)
    del ud, vd, th
    s = 0.3 / (3600 * 180 / np.pi)
    Vd = np.cos(inc) * 2 * np.pi * s * s * 3e9 * torch.exp(-2 * np.pi ** 2 * s * s * q * q)
    del q
    w = 1e4 * (0.5 + 1.5 * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
    noise = torch.randn(n, 2, device='cuda', dtype=torch.float64, generator=gen) / torch.sqrt(w)[:, None]
    V = torch.complex(Vd + noise[:, 0], noise[:, 1])
    return u.contiguous(), v.contiguous(), V.contiguous(), w.contiguous()


def cpu_sample(n_sample, N, reps=1, threads=None, want_result=False):
    """Time the oracle's map_visibilities (NumPy/SciPy restatement of the reference path) on a bounded sample."""
    from oracle import frank_oracle as fo
    u, v, V, w, dht = fo.synthetic_disc(n_sample, N, RMAX, analytic=True)
    best, res = None, None
    ctxm = None
    if threads is not None:
        try:
            from threadpoolctl import threadpool_limits
            ctxm = threadpool_limits(limits=threads)
        except Exception:
            ctxm = None
    try:
        for _ in range(reps):
            t = time.perf_counter()
            res = fo.map_visibilities(dht, u, v, V, w, *GEOM)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
    finally:
        if ctxm is not None:
            ctxm.unregister() if hasattr(ctxm, 'unregister') else ctxm.restore_original_limits()
    if want_result:
        return n_sample * N / best / 1e9, best, (u, v, V, w, res)
    return n_sample * N / best / 1e9, best


def run_reference(args, rank):
    """The reference's CPU implementation of the path (oracle port: same scipy.special.j0 + numpy.dot chunk loop) with
    all the host threads BLAS will use, on a bounded sample of the workload."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    N = wl['N']
    n = wl['n_per_gpu'] or wl['n_total']
    n_sample = 200_000 if N <= 500 else 6_000
    vals = []
    for _ in range(args.warmup):
        cpu_sample(max(2000, n_sample // 10), N)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample(n_sample, N)[0])
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = float(np.mean(vals))
    cores = os.cpu_count()
    line = {
        'impl': 'reference', 'metric': 'Gvis.mode/s for H^T W H (+ H^T W V), map_visibilities', 'value': value,
        'unit': 'Gvis.mode/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': wl['scaling'], 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': wl['name'].format(n=float(n)), 'n_vis': n, 'N': N, 'sample_n_vis': n_sample},
        'cpu_baseline': {'value': value, 'unit': 'Gvis.mode/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n_sample} of {n} visibilities per step (linear in n_vis); oracle/frank_oracle.map_visibilities '
                                   f'= the reference chunk loop (scipy.special.j0 + numpy.dot), BLAS threads unrestricted '
                                   f'(the reference recommends ONE thread, frank/fit.py:29-37: this favours the reference)'},
        'e2e': {'value': value, 'unit': 'Gvis.mode/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def parity_block(ctx, vm, N, model_scale, q_last, gdev):
    """Strict parity figures on a 300 000-visibility sample of the workload: the GPU's M against the oracle's (= the
    reference's, bit for bit, tests/test_oracle_golden.py) with NO floor, and the reference's own self-noise on the same
    data (visibilities permuted)."""
    from oracle import frank_oracle as fo
    n_s = 300_000
    cpu_val, cpu_dt, (u, v, V, w, ref) = cpu_sample(n_s, N, want_result=True)
    got = vm.map_visibilities(u, v, V, w)
    perm = np.random.default_rng(1).permutation(n_s)
    dht = fo.DHTTables(RMAX / fo.RAD_TO_ARCSEC, N)
    ref2 = fo.map_visibilities(dht, u[perm], v[perm], V[perm], w[perm], *GEOM)
    d = np.sqrt(np.diag(ref['M']))

    def errs(A, B):
        return {'max_per_entry_rel': float(np.max(np.abs(A - B) / np.abs(B))),
                'max_rel_to_sqrt_MkkMll': float(np.max(np.abs(A - B) / np.outer(d, d))),
                'max_norm_rel': float(np.max(np.abs(A - B)) / np.max(np.abs(B)))}
    return cpu_val, cpu_dt, {
        'sample_n_vis': n_s, 'N': N,
        'M_gpu_vs_reference': errs(got['M'], ref['M']),
        'M_reference_vs_itself_permuted': errs(ref2['M'], ref['M']),
        'j_max_norm_rel': float(np.max(np.abs(got['j'] - ref['j'])) / np.max(np.abs(ref['j']))),
        'H0_rel': float(abs(got['null_likelihood'] - ref['null_likelihood']) / abs(ref['null_likelihood'])),
        'note': 'strict figures, no floor; north-star bar 1e-10 per entry -- the reference misses it against itself when its '
                'visibilities are permuted (entries that cancel to 1e-7..1e-9 of sqrt(Mkk Mll)); tests hold '
                '|dM| <= 1e-10 |M| + 16 eps sqrt(Mkk Mll) and max(1e-10, 4 x self-noise) at BASELINE size'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='frank_b200', choices=['frank_b200', 'reference'])
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS))
    ap.add_argument('--n-vis', type=int, default=None, help='override the visibilities per GPU (config2) / in total (config5)')
    ap.add_argument('--no-fit', action='store_true', help='skip the end-to-end FrankFitter.fit timing')
    ap.add_argument('--no-extras', action='store_true', help='skip the binning pass, the parity block and the fit')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]

    import torch
    import torch.distributed as dist
    from frank_b200 import _lib, distributed
    from frank_b200.constants import rad_to_arcsec
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.radial_fitters import FrankFitter
    from frank_b200.statistical_models import VisibilityMapping

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    N, nchan = wl['N'], wl['nchan']
    if wl['scaling'] == 'weak':
        n = args.n_vis or wl['n_per_gpu']
        n_total = n * world
    else:
        n_total = args.n_vis or wl['n_total']
        lo, hi = distributed.shard_bounds(n_total, rank, world)
        n = hi - lo
    geom = FixedGeometry(*GEOM)
    dht = DiscreteHankelTransform(RMAX / rad_to_arcsec, N)
    debris = wl['model'] == 'debris'
    vm = VisibilityMapping(dht, geom, vis_model=wl['model'], scale_height=(lambda r: 0.05 * r) if debris else None, verbose=False,
                           device=local_rank)
    ctx = _lib.get_context(local_rank)
    ctx.dht_setup(dht)
    if world > 1:
        distributed.init_library_comm(ctx)          # the library all-reduces (M, j, H0) in-call, on its own stream
    u, v, V, w = synthetic_visibilities_device(n, dht, seed=12345 + rank)
    chan = None
    chan_sharding = None
    if nchan > 1:
        # channel-major sharding (frank_b200.distributed.channel_major_order): a rank holds as few channels as possible, so
        # its tiles of 64 baseline-sorted visibilities stay narrow in baseline (sparse regime)
        gen = torch.Generator(device='cuda').manual_seed(777 + rank)
        if world % nchan == 0:
            chan = torch.full((n,), rank // (world // nchan), device='cuda', dtype=torch.int32)
            chan_sharding = 'channel-major: one channel per rank'
        elif nchan % world == 0:
            k = nchan // world
            chan = rank * k + torch.randint(0, k, (n,), device='cuda', dtype=torch.int32, generator=gen)
            chan_sharding = f'channel-major: {k} channels per rank' if world > 1 else None
        else:
            chan = torch.randint(0, nchan, (n,), device='cuda', dtype=torch.int32, generator=gen)
            chan_sharding = 'every rank holds a share of every channel'
    gdev = geom.device_scalars()
    nM, nj = nchan * N * N, nchan * N
    out = torch.zeros(nM + nj + 1, dtype=torch.float64, device='cuda')
    Md, jd, H0d = out[:nM], out[nM:nM + nj], out[nM + nj:]
    Vr = torch.view_as_real(V).contiguous()
    q_last = float(dht.q[-1])
    model_code = _lib.MODEL_CODE[wl['model']]
    model_scale = vm._model_scale()
    H2 = vm._H2
    torch.cuda.synchronize()

    def step_device():
        rc, _, _ = ctx.map_visibilities(n, u, v, Vr, w, 1, gdev, model_code, model_scale, H2, True, q_last, Md, jd, H0d, host=False,
                                        chan=chan, nchan=nchan)
        return rc

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    gram_ms, prep_ms, fin_ms = [], [], []
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(args.steps):
        step_device()
        tm = ctx.last_map_timing()
        gram_ms.append(tm['gram_ms']); prep_ms.append(tm['prep_ms']); fin_ms.append(tm['finalize_ms'])
    dev_ms = ctx.timer_stop()            # CUDA events on the library's stream (kernels and the collective are launched on it)
    wall_ms = (time.perf_counter() - t0) * 1e3
    sync_all()
    t = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_total * N / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers: pinned (headline) and pageable ------------------------------
    hu, hv, hV, hw = [x.cpu().pin_memory() for x in (u, v, Vr, w)]
    hchan = None if chan is None else chan.cpu().pin_memory()
    hout = torch.zeros(nM + nj + 1, dtype=torch.float64).pin_memory()
    hM, hj, hH0 = hout[:nM], hout[nM:nM + nj], hout[nM + nj:]

    def e2e_run(arrs, ch, steps):
        def step_host():
            ctx.map_visibilities(n, arrs[0], arrs[1], arrs[2], arrs[3], 1, gdev, model_code, model_scale, H2, True, q_last, hM, hj, hH0,
                                 host=True, chan=ch, nchan=nchan)
        step_host()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        tt = torch.tensor([ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(2, min(args.steps, 5)) if args.workload == 'config2' else 1      # config5 steps take tens of seconds
    e2e_ms = e2e_run((hu, hv, hV, hw), hchan, e2e_steps)
    e2e_value = n_total * N / (e2e_ms * 1e-3) / 1e9
    pu, pv, pV, pw = [np.array(x.numpy()) for x in (hu, hv, hV, hw)]           # fresh pageable copies (what a frank user passes)
    pchan = None if hchan is None else np.array(hchan.numpy())
    e2e_page_ms = e2e_run((pu, pv, pV, pw), pchan, e2e_steps)
    sampler.stop_flag = True             # clocks sampled across the device-resident and the end-to-end timed regions
    sampler.join(timeout=2)

    extras = not args.no_extras and args.workload == 'config2'
    # ---- the HBM-bound uv-binning pass (UVDataBinner) timed alone on rank 0, device-resident arrays ----------
    binning = None
    if rank == 0 and extras:
        quv = torch.hypot(u, v).contiguous()
        bin_width = 1e3                                       # lambda (BASELINE.json configs[4])
        uv_max = float(quv.max().item())
        nbins = int(np.ceil(uv_max / bin_width))
        nbins += int(nbins * bin_width < uv_max)
        for _ in range(2):
            ctx.uv_bin_dev(quv, V, w, bin_width, nbins)
        reps = 5
        ctx.timer_start()
        for _ in range(reps):
            ctx.uv_bin_dev(quv, V, w, bin_width, nbins)
        bin_ms = ctx.timer_stop() / reps
        bin_bytes = 72 * n                                    # (32 B read + 4 B index written) x 2 passes
        binning = {'n_vis': n, 'nbins': nbins, 'bin_width_lambda': bin_width, 'ms': bin_ms, 'algorithmic_bytes': bin_bytes,
                   'achieved_gbs': bin_bytes / (bin_ms * 1e-3) / 1e9, 'Gvis_per_s': n / (bin_ms * 1e-3) / 1e9}
        del quv

    # ---- whole fit (map + power-spectrum loop) on rank 0, pageable NumPy inputs as a frank user passes them -------------
    fit = None
    if not args.no_fit and rank == 0 and extras:
        FF = FrankFitter(RMAX, N, geom, alpha=1.05, weights_smooth=1e-4, verbose=False, device=local_rank,
                         store_iteration_diagnostics=True)
        pVc = pV.view(np.complex128).reshape(-1)
        ctx.comm_destroy() if world > 1 else None            # the fit below is a single-rank job
        t0 = time.perf_counter()
        pre = FF.preprocess_visibilities(pu, pv, pVc, pw)
        t_map = time.perf_counter() - t0
        t0 = time.perf_counter()
        FF.fit_preprocessed(pre)             # warm-up: the first solve of a process pays one-off CUDA start-up costs
        t_first = time.perf_counter() - t0
        loops, loop_clocks = [], []
        for _ in range(3):
            smp = ClockSampler(local_rank)                    # the loop is a chain of small kernels: record the SM clock it ran at
            smp.start()
            t0 = time.perf_counter()
            FF.fit_preprocessed(pre)
            loops.append(time.perf_counter() - t0)
            smp.stop_flag = True
            smp.join(timeout=2)
            loop_clocks.append(smp.summary()['sm_mhz'])
        t_loop = float(np.median(loops))
        fit = {'fit_wall_s': t_map + t_loop, 'map_s': t_map, 'solver_loop_s': t_loop, 'solver_loop_runs_s': loops, 'solver_loop_sm_mhz': loop_clocks,
               'solver_loop_first_call_s': t_first,
               'iterations': int(FF.iteration_diagnostics['num_iterations']), 'method': 'Normal', 'alpha': 1.05, 'wsmooth': 1e-4,
               'inputs': 'pageable host NumPy arrays (fresh np.array copies)'}

    if rank == 0:
        g_ms = float(np.mean(gram_ms))
        NT = (N + 1 + 7) // 8
        useful = (N * (N + 1) + 2 * N) * n * 1.0               # executed upper triangle + j
        executed = NT * (NT + 1) // 2 * 128 * n * 1.0          # DMMA flops issued: 8x8 tiles x 2 flops x (n / 4 k-steps x 4)
        achieved = useful / (g_ms * 1e-3) / 1e12
        traffic, traffic_src = gram_traffic(n, N)
        line = {
            'metric': 'Gvis.mode/s for H^T W H (+ H^T W V), map_visibilities', 'value': value, 'unit': 'Gvis.mode/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': wl['scaling'], 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': wl['name'].format(n=float(n)), 'n_vis_per_gpu': n, 'n_vis_total': n_total, 'N': N, 'channels': nchan,
                       'channel_sharding': chan_sharding,
                       'vis_model': wl['model'], 'Rmax_arcsec': RMAX, 'geometry': GEOM,
                       'l2': f'inputs ({40 * n / 1e6:.0f} MB per step) exceed the 126 MB L2',
                       'multi_gpu': 'visibility shards; the library all-reduces (M, j, H0) over its NCCL communicator inside the call, on its own stream'},
            'e2e': {'value': e2e_value, 'unit': 'Gvis.mode/s', 'h2d_bytes_per_step': (40 + (4 if nchan > 1 else 0)) * n,
                    'd2h_bytes_per_step': 8 * (nM + nj + 1), 'ms_per_step': e2e_ms, 'steps': e2e_steps, 'host_buffers': 'pinned',
                    'pageable': {'value': n_total * N / (e2e_page_ms * 1e-3) / 1e9, 'ms_per_step': e2e_page_ms,
                                 'note': 'same call with pageable NumPy arrays: gathered into the pinned staging ring by host threads'},
                    'pipeline': 'chunks of growing size over two lanes (copy of chunk k+1 under the kernels of chunk k)'},
            'gpu_launches': (13 + 3 * (2 if N > 310 else 1) + (nchan - 1) + (1 if nchan > 1 else 0)) * args.steps,
            'kernels_per_step': ['k_prep', 'k_prep_reduce', 'k_items_from_rec', 'k_sort_hist/scan/scatter x passes', 'k_chan_starts (multi-channel)',
                                 'k_seg_finish', 'k_sort_gather', 'k_gram x channels', 'k_gram_accumulate', 'k_map_result', 'k_gram_scale'],
            'roofline': {'bound': 'tensor', 'kernel': 'k_gram (fused J0 + FP64 DMMA Gram)', 'achieved': achieved,
                         'peak': FP64_DMMA_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': achieved / FP64_DMMA_PEAK_TFLOPS,
                         'frac_of_cublas_dgemm': achieved / FP64_DGEMM_TFLOPS,
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'algorithmic_bytes': 40 * n, 'kernel_ms': g_ms,
                         'flop_count': 'useful FP64 flops N(N+1)+2N per visibility (upper triangle of H^T W H plus H^T W V; SURVEY 8d)',
                         'executed_dmma_tflops': executed / (g_ms * 1e-3) / 1e12,
                         'full_matrix_equivalent_tflops': (2 * N * N + 2 * N) * n / (g_ms * 1e-3) / 1e12,
                         'peak_source': 'builder-measured FP64 mma.sync m8n8k4 issue rate on this pool (profiles/r01_fp64_probe.txt; '
                                        'cuBLAS DGEMM 35.46 TFLOP/s on the same box); MEASURED_PEAKS.json has no FP64 entry'},
            'stage_ms': {'prepass_and_sort': float(np.mean(prep_ms)), 'gram': g_ms, 'finalize': float(np.mean(fin_ms)),
                         'wall_per_step': wall_ms / args.steps},
            'clocks': sampler.summary(),
        }
        if extras:
            cpu_val, cpu_dt, par = parity_block(ctx, vm, N, model_scale, q_last, gdev)
            line['parity'] = par
        else:
            n_s = 300_000 if N <= 500 else 6_000
            cpu_val, cpu_dt = cpu_sample(n_s, N)
        line['cpu_baseline'] = {'value': cpu_val, 'unit': 'Gvis.mode/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': f'{300_000 if extras or N <= 500 else 6_000} of {n} visibilities ({cpu_dt:.1f} s); '
                                          'oracle/frank_oracle.map_visibilities = the reference chunk loop (scipy.special.j0 + numpy.dot), '
                                          'BLAS threads unrestricted (the reference recommends one thread, which is slower)'}
        if fit:
            line['fit'] = fit
        peaks = load_peaks()
        if peaks:
            line['measured_peaks'] = {k: peaks.get(k) for k in ('hbm_gbs', 'bf16_tflops')}
        hbm_peak = float(peaks.get('hbm_gbs') or 6551.7)     # MEASURED_PEAKS.json copy bandwidth of this pool's B200
        prep_bytes = 72 * n                                   # 40 B read + 32 B record written
        line['hbm_passes'] = {
            'peak_gbs': hbm_peak,
            'prepass_and_sort': {'ms': float(np.mean(prep_ms)), 'algorithmic_bytes': prep_bytes,
                                 'achieved_gbs': prep_bytes / (float(np.mean(prep_ms)) * 1e-3) / 1e9,
                                 'frac': prep_bytes / (float(np.mean(prep_ms)) * 1e-3) / 1e9 / hbm_peak,
                                 'note': 'geometry pre-pass + stable radix sort by (channel, baseline) + SoA gather'}}
        if binning:
            binning['frac'] = binning['achieved_gbs'] / hbm_peak
            line['hbm_passes']['uv_binning'] = binning
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
