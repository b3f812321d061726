set -x
python -m pytest tests/test_gpu_binner.py -m gpu -x -q 2>&1 | tail -8
CFG5_NVIS=100000000 timeout 600 python scripts/run_configs.py 5 2>&1 | grep -v WARNING | tail -4
