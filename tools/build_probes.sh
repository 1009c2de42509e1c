#!/bin/sh
# A -DWG_PROBES build of the library next to the shipped one (result-breaking A/B probes and the 16-epilogue-warp pair kernel):
#   sh tools/build_probes.sh && WG_LIB_PATH=$PWD/text_to_speech_b200/libwg_b200_probes.so WG_PAIR=1 WG_PAIR_EPI=16 python tools/pair_ab.py
cd "$(dirname "$0")/.." && python - <<'PY'
import subprocess
from text_to_speech_b200 import _lib
out = _lib.LIB_PATH.replace("libwg_b200.so", "libwg_b200_probes.so")
subprocess.run(_lib.nvcc_command(out, extra=("-DWG_PROBES",)), check=True)
print("built", out)
PY
