# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_rows -s 5 -c 2 -o $O/rows -f python tools/profile_step.py N2 4096 eloc > /dev/null 2>&1
ncu -i $O/rows.ncu-rep --page raw --csv > $O/rows.raw.csv 2>/dev/null
ncu -i $O/rows.ncu-rep --page source --csv --print-source sass --launch-skip 1 --launch-count 1 > $O/rows.sass.csv 2>/dev/null
rm -f $O/rows.ncu-rep
