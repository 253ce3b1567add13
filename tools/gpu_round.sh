cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "resume" 2>&1 | tail -30 | cut -c1-500
