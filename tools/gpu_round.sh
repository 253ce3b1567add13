cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gradient.py -m gpu -q --tb=short 2>&1 | tail -5 | cut -c1-400
timeout 600 python tools/_grad_time.py 2>&1 | tail -4
