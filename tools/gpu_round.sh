cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_multigeometry.py -m gpu -q --tb=short -x 2>&1 | tail -6 | cut -c1-400
for v in "DPE_X=1" "DPE_DET_FACTOR_SINGLE=1"; do
env $v timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --no-weight-sharing --secondary '' 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v N2 ms/step', round(d['ms_per_step'],3), 'det_factor', r['eloc_stages_ms']['det_factor'], 'det_trace', r['eloc_stages_ms']['det_trace'])"
done
