"""Regenerates profiles/<prefix>_sass_inventory.md and the listings of the shipped hot kernels from libwg_b200.so
(cuobjdump -sass; runs on the build container, no GPU needed). Usage: python tools/sass_inventory.py [prefix]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "text_to_speech_b200", "libwg_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "SYNCS", "ELECT", "HMMA", "FFMA2", "FADD2", "FMUL2",
        "LDGSTS", "PREEXIT", "ACQBULK", "MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "FFMA", "LDS", "STS", "LDG", "STG", "BAR.SYNC", "HGMMA"]
LISTINGS = {   # demangled-name fragment -> file suffix
    "wg::tc_wn_layer_kernel<false, false>": "tc_wn_layer_kernel_0",
    "wg::tc_wn_layer_kernel<false, true>": "tc_wn_layer_kernel_first",
    "wg::tc_wn_pair_kernel<false, false, 8>": "tc_wn_pair_kernel_0",
    "wg::tf32_gate_kernel<false, 32, true, 8>": "tf32_gate_kernel_0",
    "wg::tf32_flow_kernel": "tf32_flow_kernel",
    "wg::flow_boundary_kernel": "flow_boundary_kernel",
    "mel_frames_kernel": "mel_frames_kernel",
    "lstm_mma_kernel": "lstm_mma_kernel",
}


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = [m.group(1), []]
            kernels.append(cur)
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            cur[1].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    names = subprocess.run(["c++filt"], input="\n".join(k[0] for k in kernels), capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS inventory of libwg_b200.so (cuobjdump -sass, sm_100a) -- regenerate with `python tools/sass_inventory.py`", "",
           "Counts of the mnemonics that identify the paths: UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st,",
           "UTMALDG/UTMASTG/UTMAPF = TMA load/store/prefetch, SYNCS = mbarrier ops, ELECT = elect.sync, HMMA = mma.sync (only in the",
           "Tacotron2 decoder LSTM, whose M = 16 batch tile is too small for tcgen05), FFMA2/FADD2/FMUL2 = packed fp32x2 math,",
           "LDGSTS = cp.async, PREEXIT / ACQBULK = griddepcontrol.launch_dependents / .wait (programmatic dependent launch). No HGMMA",
           "(sm_90 wgmma) anywhere; every WaveGlow tensor-core instruction is tcgen05.", "",
           "| kernel | instructions | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for (mangled, body), name in zip(kernels, names):
        short = re.sub(r"\(.*", "", name)
        ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?", "", l) for l in body]
        cnt = []
        for c in COLS:
            if c in ("FFMA", "HMMA"):
                n = sum(1 for o in ops if re.match(rf"{c}[ .]", o) and not o.startswith(c + "2"))
            else:
                n = sum(1 for o in ops if o.startswith(c))
            cnt.append(str(n) if n else "")
        out.append(f"| `{short}` | {len(ops)} | " + " | ".join(cnt) + " |")
        for frag, suffix in LISTINGS.items():
            if frag in name:
                with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_{suffix}.txt"), "w") as f:
                    f.write(f"// {name}\n" + "\n".join(body) + "\n")
    with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_inventory.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"{len(kernels)} kernels")


if __name__ == "__main__":
    main()
