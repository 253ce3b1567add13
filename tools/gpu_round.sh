set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "AssertionError|Error|passed|failed" | cut -c1-900 > gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_pytest.log
tail -2 gpurun_out/${TAG}_smoke.log
tail -c 400 gpurun_out/${TAG}_bench.err
