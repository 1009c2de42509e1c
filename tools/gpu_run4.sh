set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/plain_k2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/r02_launches_k2.csv python tools/profile_step.py > gpurun_out/ncu_k2_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_wn_layer_kernel -s 110 -c 3 -o gpurun_out/r02_wn_layer python tools/profile_step.py > gpurun_out/ncu_k2_full.log 2>&1
python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/plain_k1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_k1_tf32x3.csv python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/ncu_k1_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32_ -s 30 -c 4 -o gpurun_out/r02_tf32 python tools/profile_step.py 1 200 256 tf32x3 > gpurun_out/ncu_k1_full.log 2>&1
ls -la gpurun_out/ | tail -12
tail -3 gpurun_out/plain_k2.log gpurun_out/plain_k1.log
