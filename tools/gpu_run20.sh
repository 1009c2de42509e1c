set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/pair_ab512.py 3 > gpurun_out/pair_ab512.jsonl 2> gpurun_out/pair_ab512.err; echo "rc=$?" >> gpurun_out/pair_ab512.err
cat gpurun_out/pair_ab512.jsonl; tail -6 gpurun_out/pair_ab512.err
