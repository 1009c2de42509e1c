"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/wg_b200.h declares,
validates arguments, and fails LOUDLY (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from text_to_speech_b200 import _lib
from text_to_speech_b200.weights import WaveGlowHParams, generate_weights


def _header_functions():
    src = open(os.path.join(ROOT, "include", "wg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib_built):
    for name in _header_functions():
        assert hasattr(lib_built, name), f"{name} not exported"
    assert lib_built.wg_abi_version() == _lib.ABI_VERSION


def test_sass_is_blackwell_native(lib_built):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):     # tcgen05.mma, tcgen05.ld, TMA
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HGMMA" not in sass
    # per kernel: the BF16 layer kernels (single CTA and CTA pair), the WaveGlow-512 pair and the fp32-grade tf32x3 pair
    # all issue tcgen05.mma (UTCHMMA) fed by TMA (UTMALDG) and read their accumulators with tcgen05.ld (LDTM)
    bodies = {}
    for chunk in sass.split("Function : ")[1:]:
        bodies[chunk.split("\n", 1)[0].strip()] = chunk
    for frag in ("tc_wn_layer_kernel", "tc_wn_pair_kernel", "tc512_gate_kernel", "tc512_res_kernel", "tf32_gate_kernel", "tf32_res_kernel",
                 "tf32_flow_kernel"):
        hits = [b for name, b in bodies.items() if frag in name]
        assert hits, f"no kernel named *{frag}* in the library"
        for b in hits:
            assert "UTCHMMA" in b and "UTMALDG" in b and "LDTM" in b, f"{frag}: not a tcgen05/TMA kernel"
    # the one-launch-per-flow kernel writes its outputs with TMA stores (UTMASTG) and runs on CTA pairs (.2CTA MMAs)
    flow = next(b for name, b in bodies.items() if "tf32_flow_kernel" in name)
    assert "UTMASTG" in flow and "2CTA" in flow


def test_null_arguments_are_rejected(lib_built):
    h = ctypes.c_void_p()
    assert lib_built.wg_create(None, None, 0, 0, ctypes.byref(h)) == -1
    assert b"NULL" in lib_built.wg_last_error(None)
    n = ctypes.c_size_t()
    assert lib_built.wg_workspace_bytes(None, 1, 1, ctypes.byref(n)) == -1
    assert lib_built.wg_infer(None, None, None, 1.0, 0, 1, 1, None, None, 0, None) == -1
    assert lib_built.wg_last_launch_count(None) == 0
    lib_built.wg_destroy(None)      # must be a no-op


def test_bad_hparams_are_rejected_before_touching_cuda(lib_built):
    from text_to_speech_b200.engine import WaveGlowEngine, WaveGlowError
    hp = WaveGlowHParams(n_flows=2, n_early_every=2, n_layers=1, n_channels=16, kernel_size=5)
    w = generate_weights(hp, 1)
    with pytest.raises(WaveGlowError, match="kernel_size"):
        WaveGlowEngine(hp, w, mode="fp32")
    hp = WaveGlowHParams(n_flows=2, n_early_every=2, n_layers=1, n_channels=24)
    with pytest.raises(WaveGlowError, match="n_channels"):
        WaveGlowEngine(hp, generate_weights(hp, 1), mode="fp32")
    with pytest.raises(ValueError, match="mode"):
        WaveGlowEngine(hp, generate_weights(hp, 1), mode="int8")


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_gpu_means_loud_failure_not_fallback(lib_built):
    from text_to_speech_b200.engine import WaveGlowEngine, WaveGlowError
    hp = WaveGlowHParams(n_flows=2, n_early_every=2, n_layers=1, n_channels=16)
    with pytest.raises(WaveGlowError, match="no CUDA device|no CPU fallback"):
        WaveGlowEngine(hp, generate_weights(hp, 1), mode="fp32")


def test_weight_validation(lib_built):
    from text_to_speech_b200.weights import check_weights
    hp = WaveGlowHParams(n_flows=2, n_early_every=2, n_layers=1, n_channels=16)
    w = generate_weights(hp, 1)
    bad = dict(w)
    bad.pop("block-0/start_conv/bias")
    with pytest.raises(ValueError, match="missing"):
        check_weights(hp, bad)
    bad = dict(w)
    bad["upsample/kernel"] = np.zeros((1024, 80, 79), np.float32)
    with pytest.raises(ValueError, match="shape"):
        check_weights(hp, bad)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "text_to_speech_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle/", text, flags=re.M), \
                    f"{fn} reaches into oracle/"
