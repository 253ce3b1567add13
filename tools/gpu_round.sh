# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q --tb=short -x 2>&1 | tail -4 | cut -c1-300
for v in "DPE_X=1" "DPE_DET_FWD_128=1"; do echo "== $v"; env $v timeout 600 python tools/mcmc_timing.py Benzene 1024 2>&1 | grep "n_inter=20 graph=True" | tail -1; done
