# The round-end check on a B200 box: `gpurun -- bash tools/gpu_check.sh` -> GPU tests, parity prints, default bench line, reference arm (outputs in gpurun_out/).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -7 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_6_tf32x3.py tests/test_gpu_5_ragged.py -q -s 2>&1 | grep -E "err|max-abs|passed|failed" > gpurun_out/pytest_prints.log
cat gpurun_out/pytest_prints.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> gpurun_out/bench_default.err
tail -6 gpurun_out/bench_default.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_default.json') if x.startswith('{')]
d=json.loads(l[-1])
ex=d.pop('extras',{})
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','clocks','cpu_baseline')},indent=0)[:1500])
print(json.dumps(d['roofline'])[:600])
for k,v in ex.items(): print(k, json.dumps(v)[:1400])
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json | head -c 900; tail -4 gpurun_out/bench_reference.err
