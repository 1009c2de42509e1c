set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_6_tf32x3.py -x -q -s 2>&1 | grep -E "err|passed|failed|Error" > gpurun_out/pytest_tf32_bk16.log; cat gpurun_out/pytest_tf32_bk16.log
for bk in 32 16 32 16; do WG_TF32_BK=$bk python tools/latency_probe.py --modes tf32x3 --frames 200,860 2>/dev/null | sed "s/^/BK=$bk /"; done | tee gpurun_out/latency_tf32_bk.log
timeout 600 python -m pytest tests/test_gpu_5_ragged.py -x -q -k "no_write_outside" 2>&1 | tail -3
