cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gradient.py -m gpu -q --tb=short 2>&1 | tail -30 | cut -c1-600
timeout 600 python tools/_grad_time.py 2>&1 | tail
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/grad_launches.csv python tools/_grad_time.py > /dev/null 2>&1
