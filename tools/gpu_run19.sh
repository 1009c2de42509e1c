set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_5_ragged.py -x -q -k "hundreds" 2>&1 | tail -3
timeout 1200 python tools/soak.py 150 20 > gpurun_out/soak.log 2>&1; echo "rc=$?" >> gpurun_out/soak.log
cat gpurun_out/soak.log | tail -16
