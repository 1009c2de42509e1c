set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export WG_LIB_PATH=$PWD/text_to_speech_b200/libwg_b200_probes.so WG_PAIR=1 WG_PAIR_EPI=16
python tools/profile_step.py > gpurun_out/plain_p16.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/r02_launches_k2_pair16.csv python tools/profile_step.py > gpurun_out/ncu_p16.log 2>&1
tail -2 gpurun_out/plain_p16.log
