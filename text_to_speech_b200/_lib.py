"""ctypes binding of libwg_b200.so (C ABI: include/wg_b200.h) + the in-tree nvcc build recipe.

The product path has NO fallback: if the shared library is missing or a CUDA device is absent the
loader / wg_create raise. PyTorch is used by the host layer only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# WG_LIB_PATH: load another build of the library (development: a -DWG_PROBES build made with tools/build_probes.sh)
LIB_PATH = os.environ.get("WG_LIB_PATH") or os.path.join(_HERE, "libwg_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "engine.cu"), os.path.join(_HERE, "csrc", "mel.cu"),
           os.path.join(_HERE, "csrc", "taco.cu")]
HEADERS = [os.path.join(_HERE, "csrc", f) for f in ("common.cuh", "simt_kernels.cuh", "tc_kernels.cuh", "tc_pair_kernels.cuh", "tc_c512_kernels.cuh", "tc_tf32_kernels.cuh", "tc_tf32_flow_kernel.cuh")] + \
          [os.path.join(os.path.dirname(_HERE), "include", f) for f in ("wg_b200.h", "wg_mel_b200.h", "wg_taco_b200.h")]

WG_OK = 0
WG_MODE_FP32, WG_MODE_BF16, WG_MODE_TF32X3 = 0, 1, 2
MODES = {"fp32": WG_MODE_FP32, "bf16": WG_MODE_BF16, "tf32x3": WG_MODE_TF32X3}
ABI_VERSION = 2

# every symbol include/wg_b200.h declares (tests check the library exports all of them)
EXPORTS = ["wg_abi_version", "wg_create", "wg_destroy", "wg_last_error", "wg_workspace_bytes", "wg_infer",
           "wg_infer_host", "wg_workspace_bytes_ragged", "wg_infer_ragged", "wg_infer_host_ragged", "wg_last_launch_count", "wg_profile_enable", "wg_profile_read", "wg_debug_read_timing", "wg_debug_pair_info", "wg_debug_infer_prefix", "wg_debug_get_spect",
           "wg_debug_gemm_bf16"]
# ... and include/wg_mel_b200.h
MEL_EXPORTS = ["wg_mel_create", "wg_mel_destroy", "wg_mel_last_error", "wg_mel_frames", "wg_mel_spectrogram",
               "wg_mel_spectrogram_host"]


class WgConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n_mel_channels", "n_flows", "n_group", "n_early_every", "n_early_size", "n_layers",
                 "n_channels", "kernel_size", "mode")]


class WgMelConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("sampling_rate", "n_mel_channels", "filter_length", "hop_length", "win_length")] + \
               [("clip_val", ctypes.c_float)]


TACO_EXPORTS = ["wg_taco_create", "wg_taco_destroy", "wg_taco_last_error", "wg_taco_decode", "wg_taco_set_graph_chunk"]


class WgTacoConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n_mel_channels", "prenet_dim", "embedding_dim", "attention_rnn_dim", "decoder_rnn_dim",
                 "attention_dim", "attention_filters", "attention_kernel_size")] + [("prenet_drop_rate", ctypes.c_float),
                                                                                    ("lstm_weight_dtype", ctypes.c_int32)]


class WgTensor(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.POINTER(ctypes.c_float)),
                ("ndim", ctypes.c_int32), ("shape", ctypes.c_int64 * 4)]


def nvcc_command(out=LIB_PATH, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", *extra, "-o", out] + SOURCES


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    """Compiles csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def load_library():
    """dlopens libwg_b200.so and declares the prototypes. Raises if the library is absent --
    there is deliberately no Python/CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the B200 WaveGlow runtime has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, f32p = c.c_void_p, c.c_int32, c.POINTER(c.c_float)
    lib.wg_abi_version.restype = c.c_int
    lib.wg_abi_version.argtypes = []
    lib.wg_create.restype = c.c_int
    lib.wg_create.argtypes = [c.POINTER(WgConfig), c.POINTER(WgTensor), i32, i32, c.POINTER(vp)]
    lib.wg_destroy.restype = None
    lib.wg_destroy.argtypes = [vp]
    lib.wg_last_error.restype = c.c_char_p
    lib.wg_last_error.argtypes = [vp]
    lib.wg_workspace_bytes.restype = c.c_int
    lib.wg_workspace_bytes.argtypes = [vp, i32, i32, c.POINTER(c.c_size_t)]
    lib.wg_infer.restype = c.c_int
    lib.wg_infer.argtypes = [vp, vp, vp, c.c_float, i32, i32, i32, vp, vp, c.c_size_t, vp]
    lib.wg_infer_host.restype = c.c_int
    lib.wg_infer_host.argtypes = [vp, f32p, f32p, c.c_float, i32, i32, i32, f32p]
    i32p = c.POINTER(i32)
    lib.wg_workspace_bytes_ragged.restype = c.c_int
    lib.wg_workspace_bytes_ragged.argtypes = [vp, i32, i32, i32p, c.POINTER(c.c_size_t)]
    lib.wg_infer_ragged.restype = c.c_int
    lib.wg_infer_ragged.argtypes = [vp, vp, vp, c.c_float, i32, i32, i32, i32p, vp, vp, c.c_size_t, vp]
    lib.wg_infer_host_ragged.restype = c.c_int
    lib.wg_infer_host_ragged.argtypes = [vp, f32p, f32p, c.c_float, i32, i32, i32, i32p, f32p]
    lib.wg_last_launch_count.restype = c.c_int
    lib.wg_last_launch_count.argtypes = [vp]
    lib.wg_profile_enable.restype = c.c_int
    lib.wg_profile_enable.argtypes = [vp, i32]
    lib.wg_profile_read.restype = c.c_int
    lib.wg_profile_read.argtypes = [vp, c.POINTER(c.c_double), c.POINTER(i32)]
    lib.wg_debug_read_timing.restype = c.c_int
    lib.wg_debug_read_timing.argtypes = [vp, c.POINTER(c.c_uint64)]
    lib.wg_debug_pair_info.restype = c.c_int
    lib.wg_debug_pair_info.argtypes = [vp, c.POINTER(i32), c.POINTER(i32)]
    lib.wg_debug_infer_prefix.restype = c.c_int
    lib.wg_debug_infer_prefix.argtypes = [vp, vp, vp, c.c_float, i32, i32, i32, vp, c.c_size_t, vp, i32, i32, vp, vp]
    lib.wg_debug_get_spect.restype = c.c_int
    lib.wg_debug_get_spect.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.wg_debug_gemm_bf16.restype = c.c_int
    lib.wg_debug_gemm_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.wg_mel_create.restype = c.c_int
    lib.wg_mel_create.argtypes = [c.POINTER(WgMelConfig), f32p, f32p, i32, c.POINTER(vp)]
    lib.wg_mel_destroy.restype = None
    lib.wg_mel_destroy.argtypes = [vp]
    lib.wg_mel_last_error.restype = c.c_char_p
    lib.wg_mel_last_error.argtypes = [vp]
    lib.wg_mel_frames.restype = c.c_int
    lib.wg_mel_frames.argtypes = [vp, c.c_int64, c.POINTER(c.c_int64)]
    lib.wg_mel_spectrogram.restype = c.c_int
    lib.wg_mel_spectrogram.argtypes = [vp, vp, i32, c.c_int64, vp, vp]
    lib.wg_mel_spectrogram_host.restype = c.c_int
    lib.wg_mel_spectrogram_host.argtypes = [vp, f32p, i32, c.c_int64, f32p]
    lib.wg_taco_create.restype = c.c_int
    lib.wg_taco_create.argtypes = [c.POINTER(WgTacoConfig), c.POINTER(WgTensor), i32, i32, c.POINTER(vp)]
    lib.wg_taco_destroy.restype = None
    lib.wg_taco_destroy.argtypes = [vp]
    lib.wg_taco_last_error.restype = c.c_char_p
    lib.wg_taco_last_error.argtypes = [vp]
    lib.wg_taco_decode.restype = c.c_int
    lib.wg_taco_decode.argtypes = [vp, vp, c.POINTER(i32), i32, i32, i32, i32, i32, c.c_uint64, vp, vp, vp, vp,
                                   c.POINTER(i32), vp]
    lib.wg_taco_set_graph_chunk.restype = c.c_int
    lib.wg_taco_set_graph_chunk.argtypes = [vp, i32]
    if lib.wg_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libwg_b200.so ABI {lib.wg_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib
