# scratch runner for gpurun calls during development: edit, then  gpurun -- 'bash tools/gpu_round.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_n2.csv python tools/profile_step.py N2 4096 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_benzene1024.csv python tools/profile_step.py Benzene 1024 mcmc,eloc > /dev/null 2>&1
cap() { # name regex skip count what
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/$1 -f python tools/profile_step.py N2 4096 $5 > /dev/null 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
cap conv 'k_conv_fused' 4 4 eloc
cap fwd_kernels 'k_envelope_fwd|k_det_fwd_half|k_conv_fwd|k_pair_stream_tc|k_eion_stream|k_act_mean_fwd' 6 10 mcmc
cap det_factor 'k_det_factor_half' 1 1 eloc
