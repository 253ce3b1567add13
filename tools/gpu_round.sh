set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "AssertionError|Error|passed|failed" | cut -c1-900 > gpurun_out/r02d_pytest.log
timeout 900 python tools/parity_table.py gpurun_out/r02d_parity_table.md > gpurun_out/r02d_parity_stdout.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d_smoke.log 2>&1
cat gpurun_out/r02d_pytest.log
tail -3 gpurun_out/r02d_smoke.log
