set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "rc=$?" >> gpurun_out/bench_${n}gpu.err
tail -2 gpurun_out/bench_${n}gpu.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_${n}gpu.json') if x.startswith('{')]
d=json.loads(l[-1]); ex=d.pop('extras',{})
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'])
k=ex['k5_strong_scaling']; print({a:k[a] for a in ('value','s_per_sweep','load_imbalance','device_span_s_per_sweep_max','rank0_host_s_last_sweep')})
PY
done
