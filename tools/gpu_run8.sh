set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/layer_timing.py > gpurun_out/layer_timing.jsonl 2> gpurun_out/layer_timing.err; echo "rc=$?" >> gpurun_out/layer_timing.err
cat gpurun_out/layer_timing.jsonl; tail -3 gpurun_out/layer_timing.err
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
tail -6 gpurun_out/bench_default.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_default.json') if x.startswith('{')]
d=json.loads(l[-1]); ex=d.pop('extras',{})
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','clocks')})[:900])
print(json.dumps(d['roofline'])[:500])
for k,v in ex.items(): print(k, json.dumps(v)[:900])
PY
