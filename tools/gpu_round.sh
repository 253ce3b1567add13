# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_gradient.py -m gpu -q --tb=short 2>&1 | tail -15 | cut -c1-500
