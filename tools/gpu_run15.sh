set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/sanitize_step.py > gpurun_out/sanitize_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_step.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -4 gpurun_out/sanitize_plain.log; tail -12 gpurun_out/sanitize_memcheck.log
