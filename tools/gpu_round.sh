cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mcmc_graph.py tests/test_gpu_xla_shim.py tests/test_gpu_sharding.py -m gpu -q --tb=short -k "mcmc or proposal or graph or xla or shard or chain" 2>&1 | tail -30 | cut -c1-500
