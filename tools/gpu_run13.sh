set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/pair_ab.py 5 > gpurun_out/pair_ab.jsonl 2> gpurun_out/pair_ab.err; echo "rc=$?" >> gpurun_out/pair_ab.err
cat gpurun_out/pair_ab.jsonl; tail -8 gpurun_out/pair_ab.err
