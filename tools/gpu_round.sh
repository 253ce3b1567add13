cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
O=gpurun_out/r02
rm -f gpurun_out/tl.txt
DPE_GEMM_TIMELINE=gpurun_out/tl.txt timeout 300 python tools/profile_step.py N2 4096 eloc > /dev/null 2>&1
cap() { # name regex skip count what
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $O/$1 -f python tools/profile_step.py N2 4096 $5 > /dev/null 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1.raw.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
cap grad_atb 'k_atb' 212 30 eloc,grad
cap grad_pair 'k_bw_pair|k_bw_eion|k_det_inverse|k_bw_orbitals|k_gemm_nt' 22 8 eloc,grad
ls -la $O | head -30
