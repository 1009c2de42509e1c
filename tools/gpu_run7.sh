set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
WG_PAIR=1 timeout 1200 python -m pytest tests/test_gpu_2_bf16.py tests/test_gpu_5_ragged.py -x -q > gpurun_out/pytest_pair.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_pair.log
tail -6 gpurun_out/pytest_pair.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1]); ex=d.pop('extras',{})
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
for k,v in ex.items(): print(k, json.dumps(v)[:1800])
PY
