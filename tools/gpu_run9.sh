set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for p in 0 1 0 1; do
WG_PAIR=$p python bench.py --workload k5 --steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); k=d['detail']
print('WG_PAIR=$p', k['s_per_sweep'], k['device_span_s_per_sweep_max'], k['clocks']['sm_mhz'], k['rank0_host_s_last_sweep'])"
done
