# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "large or every_kernel" 2>&1 | tail -4 | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --no-weight-sharing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['secondary']
print('N2', round(d['ms_per_step'],3), 'benzene', s['value'], s['ms_per_step'], s['roofline']['eloc_stages_ms'])"
