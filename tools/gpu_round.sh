cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_multigeometry.py -m gpu -q --tb=line 2>&1 | grep -E "AssertionError|Error|passed|failed" | cut -c1-600
for v in "" "DPE_DET_EPI_GROUPS=1"; do
  env $v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['secondary']; r2=s['roofline']
print('VARIANT [$v]', 'ms/step', round(d['ms_per_step'],3), 'eloc', round(r['eloc_pass_ms'],3), 'frac', round(r['frac'],3), round(r['class_frac'],3), 'main', r['eloc_stages_ms']['main_layer'], 'orb', r['eloc_stages_ms']['orbitals'])
print('   benzene ms/step', round(s['ms_per_step'],2), 'evals/s', round(s['value'],1), 'stages', r2['eloc_stages_ms'])"
done
