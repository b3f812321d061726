set -x
python -m pytest tests/test_gpu_binner.py tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -15
python scripts/prof_bin.py 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_s2b.json 2> gpurun_out/r01_bench_s2b.err; tail -c 1500 gpurun_out/r01_bench_s2b.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r01_bench_s2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['stage_ms'], d['hbm_passes'], d['fit'])
P
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r01_bin_launches2.csv python scripts/prof_bin.py > /dev/null 2>&1
