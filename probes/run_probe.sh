#!/bin/bash
# First GPU call: FP64 pipe denominators (raw DFMA/DMMA issue rate + cuBLAS DGEMM via torch).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/probe_smi.txt 2>&1
timeout 300 ./build/fp64_probe > gpurun_out/fp64_probe.txt 2>&1
timeout 300 python - > gpurun_out/dgemm_probe.txt 2>&1 <<'PY'
import torch, time
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"cuBLAS DGEMM {n}^3 best of 5: {best:.2f} ms = {2*n**3/best/1e9:.2f} TFLOP/s")
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    c = a @ b; k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS DGEMM sustained 4 s: {2*n**3*k/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s over {k} calls")
a = torch.randn(n, 512, dtype=torch.float64, device='cuda')
for _ in range(2): c = a.T @ a
torch.cuda.synchronize()
e0.record(); c = a.T @ a; e1.record(); torch.cuda.synchronize()
print(f"cuBLAS DGEMM (512x8192)x(8192x512): {e0.elapsed_time(e1):.3f} ms = {2*512*512*n/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s")
PY
cat gpurun_out/fp64_probe.txt gpurun_out/dgemm_probe.txt
