cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'step_ms', d.get('step_ms'), 'clocks', d['clocks'])
PY
}
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --secondary '' --no-weight-sharing > gpurun_out/s1.json 2>/dev/null; show gpurun_out/s1.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --secondary '' --no-weight-sharing > gpurun_out/s2.json 2>/dev/null; show gpurun_out/s2.json
CUDA_VISIBLE_DEVICES=1 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-cadence --secondary '' --no-weight-sharing > gpurun_out/s1b.json 2>/dev/null; show gpurun_out/s1b.json
