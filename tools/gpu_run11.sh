set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_5_ragged.py tests/test_gpu_4_tts.py -x -q > gpurun_out/pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sel.log
tail -5 gpurun_out/pytest_sel.log
for p in -1 -1; do
WG_PAIR=$p python bench.py --workload k5 --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['detail']
print('WG_PAIR=$p', k['s_per_sweep'], k['device_span_s_per_sweep_max'], k['clocks']['sm_mhz'], k['rank0_host_s_last_sweep'])"
done
