# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_multigeometry.py tests/test_gpu_gradient.py -m gpu -q --tb=short -x 2>&1 | tail -4 | cut -c1-300
timeout 600 python tools/mcmc_timing.py N2 4096 2>&1 | grep "n_inter=20 graph=True" | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --no-weight-sharing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; s=d['secondary']
print('N2 ms/step', round(d['ms_per_step'],3), {k:r['eloc_stages_ms'][k] for k in ('el_ion_stream','pair_stream','main_layer')}, 'fwd eion', r['forward_stages_ms']['el_ion_stream'], '| benzene', round(s['ms_per_step'],1), s['roofline']['eloc_stages_ms']['el_ion_stream'])"
