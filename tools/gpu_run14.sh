set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export WG_PAIR=1
python tools/profile_step.py > gpurun_out/plain_k2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_wn_pair_kernel -s 110 -c 3 -o gpurun_out/r02_wn_pair python tools/profile_step.py > gpurun_out/ncu_k2_full.log 2>&1
unset WG_PAIR
python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:tf32_(gate|res)_kernel" -s 30 -c 4 -o gpurun_out/r02_tf32 python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/ncu_k1_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/plain_k2.log | tail -2
