# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_multigeometry.py -m gpu -q --tb=short -x 2>&1 | tail -4 | cut -c1-300
for i in 1 2; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --no-weight-sharing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['secondary']
print('N2 ms/step', round(d['ms_per_step'],3), 'conv', r['eloc_stages_ms']['schnet_conv'], 'main', r['eloc_stages_ms']['main_layer'], '| benzene', round(s['ms_per_step'],1), 'conv', s['roofline']['eloc_stages_ms']['schnet_conv'])"
done
