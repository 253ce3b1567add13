cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | cut -c1-400
timeout 600 python tools/mcmc_timing.py N2 4096 2>&1 | grep "n_inter=20 graph=True" | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --secondary '' --no-weight-sharing 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step', round(d['ms_per_step'],3), 'stages', r['eloc_stages_ms'], 'fwd', r['forward_stages_ms'])"
