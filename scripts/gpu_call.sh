set -x
python -m pytest tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -15
python scripts/run_configs.py 3full 2>&1 | grep -v WARNING | tail -5
