cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multigeometry.py -m gpu -q --tb=short -x 2>&1 | tail -8 | cut -c1-500
