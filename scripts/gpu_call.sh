set -x
FB_SOLVE_MU=smem python scripts/dev_loop_var.py 2>&1 | grep -v WARNING | tail -2
python scripts/dev_loop_var.py 2>&1 | grep -v WARNING | tail -2
python -m pytest tests/test_gpu_fit.py -m gpu -x -q -k "gaussian_model or frank_fitter_normal or solver_loop or sweep" 2>&1 | tail -3
