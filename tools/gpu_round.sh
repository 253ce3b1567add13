cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_paths.py > gpurun_out/memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|done|E_mean" gpurun_out/memcheck.log | head -20
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_paths.py > gpurun_out/racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|hazard|done" gpurun_out/racecheck.log | head -20
