set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?" >> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1]); ex=d.pop('extras',{})
print({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks','cpu_baseline')}, d['e2e'])
for k,v in ex.items(): print(k, json.dumps(v)[:1200])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "rc=$?"
head -c 400 gpurun_out/bench_2gpu_ref.json
