#!/usr/bin/env python
"""Benchmark of the hot path: visibilities -> GP normal equations (H^T W H, H^T W V), BASELINE.json config 2.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of `VisibilityMapping.map_visibilities` (geometry pre-pass + baseline sort + fused
J0/Gram kernel + split-K reduction) over one batch of n_vis synthetic visibilities at N_modes collocation
points.  `value` is Gvis.mode/s = n_vis * N_modes / t / 1e9 with the inputs already resident in HBM; `e2e` is
the same metric through the C ABI with HOST (pinned) buffers, host->device and device->host copies inside the
timed region.  With --gpus N > 1 (launched by torch.distributed.run) every rank maps its own n_vis
visibilities and the partial (M, j, H0) are summed with one NCCL all-reduce per step (weak scaling).

--impl reference times the CPU restatement of the reference's NumPy/SciPy path (oracle/frank_oracle.py --
the reference itself is pure Python and is not present on the GPU box) on a bounded sample of the same
workload, on the host cores.
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

N_VIS = 10_000_000
N_MODES = 300
RMAX = 1.6
GEOM = (30., 40., 1e-3, -2e-3)
# measured on this pool's B200 with probes/fp64_probe.cu (profiles/r01_fp64_probe.txt): raw mma.sync m8n8k4 f64
# issue rate; MEASURED_PEAKS.json carries no FP64 figure
FP64_DMMA_PEAK_TFLOPS = 37.1


def gram_traffic(n, N):
    """DRAM bytes of one k_gram launch from the committed ncu --set full capture of this workload, else None."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r01_gram_traffic.json')) as fh:
            t = json.load(fh)
        if t['n_vis'] == n and t['N'] == N:
            return t['dram_bytes_read'] + t['dram_bytes_write']
    except Exception:
        pass
    return None


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
    s = 0.3 / (3600 * 180 / np.pi)
    Vd = np.cos(inc) * 2 * np.pi * s * s * 3e9 * torch.exp(-2 * np.pi ** 2 * s * s * q * q)
    w = 1e4 * (0.5 + 1.5 * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
    noise = torch.randn(n, 2, device='cuda', dtype=torch.float64, generator=gen) / torch.sqrt(w)[:, None]
    V = torch.complex(Vd + noise[:, 0], noise[:, 1])
    return u.contiguous(), v.contiguous(), V.contiguous(), w.contiguous()


def cpu_sample(n_sample, reps=1):
    """Time the oracle's map_visibilities (NumPy/SciPy restatement of the reference path) on a bounded sample."""
    from oracle import frank_oracle as fo
    u, v, V, w, dht = fo.synthetic_disc(n_sample, N_MODES, RMAX, analytic=True)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        fo.map_visibilities(dht, u, v, V, w, *GEOM)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n_sample * N_MODES / best / 1e9, best


def run_reference(args, rank):
    if rank != 0:
        return
    n_sample = 200_000
    vals = []
    for _ in range(args.warmup):
        cpu_sample(20_000)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_sample(n_sample)[0])
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = float(np.mean(vals))
    cores = os.cpu_count()
    line = {
        'impl': 'reference', 'metric': 'Gvis.mode/s for H^T W H (+ H^T W V), map_visibilities', 'value': value,
        'unit': 'Gvis.mode/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'Normal fit mapping, {N_VIS:.0e} unbinned visibilities, N={N_MODES} (BASELINE.json configs[1])',
                   'n_vis': N_VIS, 'N': N_MODES, 'sample_n_vis': n_sample},
        'cpu_baseline': {'value': value, 'unit': 'Gvis.mode/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{n_sample} of {N_VIS} visibilities per step (linear in n_vis); oracle/frank_oracle.map_visibilities '
                                   f'= the reference chunk loop (scipy.special.j0 + numpy.dot), BLAS threads unrestricted'},
        'e2e': {'value': value, 'unit': 'Gvis.mode/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='frank_b200', choices=['frank_b200', 'reference'])
    ap.add_argument('--n-vis', type=int, default=N_VIS)
    ap.add_argument('--no-fit', action='store_true', help='skip the end-to-end FrankFitter.fit timing')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from frank_b200 import _lib
    from frank_b200.constants import rad_to_arcsec
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.radial_fitters import FrankFitter
    from frank_b200.statistical_models import VisibilityMapping

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    n = args.n_vis
    N = N_MODES
    geom = FixedGeometry(*GEOM)
    dht = DiscreteHankelTransform(RMAX / rad_to_arcsec, N)
    vm = VisibilityMapping(dht, geom, verbose=False, device=local_rank)
    ctx = _lib.get_context(local_rank)
    ctx.dht_setup(dht)
    u, v, V, w = synthetic_visibilities_device(n, dht, seed=12345 + rank)
    gdev = geom.device_scalars()
    out = torch.zeros(N * N + N + 1, dtype=torch.float64, device='cuda')
    Md, jd, H0d = out[:N * N], out[N * N:N * N + N], out[N * N + N:]
    Vr = torch.view_as_real(V).contiguous()
    q_last = float(dht.q[-1])
    model_scale = float(np.cos(np.deg2rad(GEOM[0])))

    def step_device():
        rc, qmin, qmax = ctx.map_visibilities(n, u, v, Vr, w, 1, gdev, 0, model_scale, None, True, q_last, Md, jd, H0d, host=False)
        if world > 1:
            dist.all_reduce(out)          # partial M, j, H0 of independent visibilities
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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ctx.timer_start()
    ev0.record()
    for _ in range(args.steps):
        step_device()
        tm = ctx.last_map_timing()
        gram_ms.append(tm['gram_ms']); prep_ms.append(tm['prep_ms']); fin_ms.append(tm['finalize_ms'])
    ev1.record()
    lib_ms = ctx.timer_stop()            # CUDA events on the library's stream
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = max(lib_ms, ev0.elapsed_time(ev1)) if world > 1 else lib_ms
    sync_all()
    t = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n * N / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host (pinned) buffers -----------------------------------------
    hu, hv, hV, hw = [x.cpu().pin_memory() for x in (u, v, Vr, w)]
    hout = torch.zeros(N * N + N + 1, dtype=torch.float64).pin_memory()
    hM, hj, hH0 = hout[:N * N], hout[N * N:N * N + N], hout[N * N + N:]

    def step_host():
        ctx.map_visibilities(n, hu, hv, hV, hw, 1, gdev, 0, model_scale, None, True, q_last, hM, hj, hH0, host=True)
        if world > 1:
            out.copy_(hout, non_blocking=True)
            dist.all_reduce(out)
            hout.copy_(out)

    e2e_steps = max(2, min(args.steps, 5))
    step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * N / (float(t.item()) * 1e-3) / 1e9
    sampler.stop_flag = True             # clocks sampled across the device-resident and the end-to-end timed regions
    sampler.join(timeout=2)

    # ---- the HBM-bound uv-binning pass (UVDataBinner) timed alone on rank 0, device-resident arrays ----------
    binning = None
    if rank == 0:
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
        binning = {'kernels': ['k_bin_index', 'k_sort_hist x2', 'k_sort_scan x2', 'k_sort_scatter x2', 'k_bin_starts', 'k_bin_reduce'],
                   'n_vis': n, 'nbins': nbins, 'bin_width_lambda': bin_width, 'ms': bin_ms, 'algorithmic_bytes': bin_bytes,
                   'achieved_gbs': bin_bytes / (bin_ms * 1e-3) / 1e9, 'Gvis_per_s': n / (bin_ms * 1e-3) / 1e9}
        del quv

    # ---- whole fit (map + power-spectrum loop) on rank 0 ------------------------------------------------
    fit = None
    if not args.no_fit and rank == 0:
        FF = FrankFitter(RMAX, N, geom, alpha=1.05, weights_smooth=1e-4, verbose=False, device=local_rank,
                         store_iteration_diagnostics=True)
        hVc = torch.view_as_complex(hV).numpy()
        t0 = time.perf_counter()
        pre = FF.preprocess_visibilities(hu.numpy(), hv.numpy(), hVc, hw.numpy())
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
               'inputs': 'host numpy arrays (pageable)'}

    if rank == 0:
        g_ms = float(np.mean(gram_ms))
        useful = (N * (N + 1) + 2 * N) * n                     # executed upper triangle + j
        nt = (N + 1 + 7) // 8
        executed = nt * (nt + 1) // 2 * 128 * n                # DMMA flops issued: 8x8 tiles x 2 flops x (n / 4 k-steps x 4)
        achieved = useful / (g_ms * 1e-3) / 1e12
        cpu_val, cpu_dt = cpu_sample(300_000)
        line = {
            'metric': 'Gvis.mode/s for H^T W H (+ H^T W V), map_visibilities', 'value': value, 'unit': 'Gvis.mode/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'Normal fit mapping, {n:.0e} unbinned visibilities per GPU, N={N} (BASELINE.json configs[1])',
                       'n_vis_per_gpu': n, 'N': N, 'Rmax_arcsec': RMAX, 'geometry': GEOM,
                       'l2': 'inputs (400 MB per step) exceed the 126 MB L2', 'multi_gpu': 'visibility shards + NCCL all-reduce of (M, j, H0)'},
            'e2e': {'value': e2e_value, 'unit': 'Gvis.mode/s', 'h2d_bytes_per_step': 40 * n, 'd2h_bytes_per_step': 8 * (N * N + N + 1),
                    'ms_per_step': float(t.item()), 'steps': e2e_steps, 'host_buffers': 'pinned'},
            'gpu_launches': 12 * args.steps,
            'kernels_per_step': ['k_prep', 'k_prep_reduce', 'k_items_from_rec', 'k_sort_hist x2', 'k_sort_scan x2', 'k_sort_scatter x2', 'k_sort_gather',
                                 'k_gram', 'k_gram_finalize'],
            'roofline': {'bound': 'tensor', 'kernel': 'k_gram (fused J0 + FP64 DMMA Gram)', 'achieved': achieved,
                         'peak': FP64_DMMA_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': achieved / FP64_DMMA_PEAK_TFLOPS,
                         'traffic': gram_traffic(n, N), 'traffic_unit': 'bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_gram_traffic.json)',
                         'algorithmic_bytes': 4 * 24 * n, 'kernel_ms': g_ms,
                         'flop_count': 'useful FP64 flops N(N+1)+2N per visibility (upper triangle of H^T W H plus H^T W V; SURVEY 8d)',
                         'executed_dmma_tflops': executed / (g_ms * 1e-3) / 1e12,
                         'full_matrix_equivalent_tflops': (2 * N * N + 2 * N) * n / (g_ms * 1e-3) / 1e12,
                         'j0_evaluations_per_s': 760 * n / (g_ms * 1e-3) if N == 300 else None,
                         'peak_source': 'FP64 mma.sync m8n8k4 issue rate measured on this pool (profiles/r01_fp64_probe.txt); '
                                        'MEASURED_PEAKS.json has no FP64 entry'},
            'stage_ms': {'prepass_and_sort': float(np.mean(prep_ms)), 'gram': g_ms, 'finalize': float(np.mean(fin_ms)),
                         'wall_per_step': wall_ms / args.steps},
            'cpu_baseline': {'value': cpu_val, 'unit': 'Gvis.mode/s', 'cores': os.cpu_count(), 'kind': 'port',
                             'sample': f'300000 of {n} visibilities ({cpu_dt:.1f} s); oracle/frank_oracle.map_visibilities = the reference '
                                       'chunk loop (scipy.special.j0 + numpy.dot), BLAS threads unrestricted'},
            'clocks': sampler.summary(),
        }
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
                                 'note': 'geometry pre-pass + 2-pass stable radix sort + SoA gather + work-table upload'}}
        if binning:
            binning['frac'] = binning['achieved_gbs'] / hbm_peak
            line['hbm_passes']['uv_binning'] = binning
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
