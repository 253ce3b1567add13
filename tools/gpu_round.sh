cd $GRAFT_REPO_ROOT
timeout 600 python tools/mcmc_timing.py N2 4096 2>&1 | tail -10
timeout 600 python tools/mcmc_timing.py HChain10 512 2>&1 | tail -10
