cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo skip tests
for v in "DPE_MCMC_GRAPH=0" "DPE_MCMC_GRAPH=1"; do
env $v timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --secondary '' > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -3 gpurun_out/bench_now.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_now.json').read().strip().splitlines()[-1])
print('$v ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'])
c=d['cadence']; print('  cadence', c['value'], c['ms_per_epoch'], c['metropolis_ms_per_step'], 'opt epoch', c['optimisation_epoch']['ms_per_epoch'])
PY
done
