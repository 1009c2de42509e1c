set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_6_tf32x3.py -x -q > gpurun_out/pytest_tf32.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tf32.log
tail -30 gpurun_out/pytest_tf32.log
timeout 600 python -m pytest tests/test_gpu_2_bf16.py -x -q -k "concurrent or waveglow512_full" > gpurun_out/pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sel.log
tail -15 gpurun_out/pytest_sel.log
timeout 600 python tools/latency_probe.py > gpurun_out/latency.jsonl 2> gpurun_out/latency.err; echo "rc=$?" >> gpurun_out/latency.err
cat gpurun_out/latency.jsonl; tail -5 gpurun_out/latency.err
