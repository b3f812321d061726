set -x
python scripts/dev_loop_var.py 2>&1 | grep -v WARNING | tail -2
python -m pytest tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4
