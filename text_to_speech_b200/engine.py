"""WaveGlowEngine: Python handle on one wg_engine (one GPU). Host logic only -- every FLOP of the
path runs in libwg_b200.so. PyTorch provides device tensors, the caller's stream and pinned memory."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .weights import WaveGlowHParams, load_weights, check_weights, HOP


class WaveGlowError(RuntimeError):
    """Non-zero wg_status (mirrors the RuntimeError the reference's TensorRT runtime raises,
    utils/keras/runtimes/tensorrt_runtime.py:169,198-199)."""


class WaveGlowEngine:
    def __init__(self, hparams: WaveGlowHParams, weights: dict, *, mode: str = "bf16", device: int = 0):
        if mode not in _lib.MODES:
            raise ValueError(f"mode must be one of {sorted(_lib.MODES)}, got {mode!r}")
        check_weights(hparams, weights)
        self._lib = _lib.load_library()
        self.hp, self.mode, self.device = hparams, mode, int(device)
        cfg = _lib.WgConfig(hparams.n_mel_channels, hparams.n_flows, hparams.n_group, hparams.n_early_every,
                            hparams.n_early_size, hparams.n_layers, hparams.n_channels, hparams.kernel_size,
                            _lib.MODES[mode])
        names = [k for k in weights if not k.startswith("__")]
        keep = []           # keep the float32 host arrays alive during wg_create
        arr = (_lib.WgTensor * len(names))()
        for i, name in enumerate(names):
            a = np.ascontiguousarray(weights[name], dtype=np.float32)
            keep.append(a)
            arr[i].name = name.encode()
            arr[i].data = a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
        handle = ctypes.c_void_p()
        rc = self._lib.wg_create(ctypes.byref(cfg), arr, len(names), self.device, ctypes.byref(handle))
        if rc != _lib.WG_OK:
            raise WaveGlowError(f"wg_create failed ({rc}): {self._lib.wg_last_error(None).decode()}")
        self._h = handle
        self._ws = None          # torch uint8 workspace, grown on demand

    @classmethod
    def from_file(cls, path, **kw):
        hp, w = load_weights(path)
        return cls(hp, w, **kw)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wg_destroy(self._h)
            self._h = None

    __del__ = close

    # -- helpers ----------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != _lib.WG_OK:
            raise WaveGlowError(f"{what} failed ({rc}): {self._lib.wg_last_error(self._h).decode()}")

    def workspace_bytes(self, B, T):
        n = ctypes.c_size_t()
        self._check(self._lib.wg_workspace_bytes(self._h, B, T, ctypes.byref(n)), "wg_workspace_bytes")
        return n.value

    def workspace_bytes_ragged(self, B, T, lengths):
        n = ctypes.c_size_t()
        self._check(self._lib.wg_workspace_bytes_ragged(self._h, B, T, lengths, ctypes.byref(n)), "wg_workspace_bytes_ragged")
        return n.value

    def _lengths(self, lengths, B, T):
        """Per-utterance frame counts -> a C int32 array (validated here so a bad list never reaches the C side)."""
        ls = [int(x) for x in lengths]
        if len(ls) != B or any(l <= 0 or l > T for l in ls):
            raise ValueError(f"lengths must hold {B} frame counts in [1, {T}], got {ls[:8]}{'...' if len(ls) > 8 else ''}")
        return (ctypes.c_int32 * B)(*ls)

    def _workspace(self, B, T, lengths=None):
        import torch
        need = self.workspace_bytes(B, T) if lengths is None else self.workspace_bytes_ragged(B, T, lengths)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=f"cuda:{self.device}")
        off = (-self._ws.data_ptr()) % 1024
        return self._ws.data_ptr() + off, self._ws.numel() - off

    @property
    def last_launch_count(self):
        return self._lib.wg_last_launch_count(self._h)

    def profile_enable(self, on=True):
        self._check(self._lib.wg_profile_enable(self._h, int(bool(on))), "wg_profile_enable")

    def profile_read(self):
        """(sum of WN-layer kernel durations in ms, number of launches) since the last read."""
        ms, n = ctypes.c_double(), ctypes.c_int32()
        self._check(self._lib.wg_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n)), "wg_profile_read")
        return ms.value, n.value

    def pair_info(self):
        """(CTA pairs resident at once on this device, whether the last infer used the CTA-pair layer kernel)."""
        a, b = ctypes.c_int32(), ctypes.c_int32()
        self._check(self._lib.wg_debug_pair_info(self._h, ctypes.byref(a), ctypes.byref(b)), "wg_debug_pair_info")
        return a.value, bool(b.value)

    def read_layer_timing(self):
        buf = (ctypes.c_uint64 * 128)()
        self._check(self._lib.wg_debug_read_timing(self._h, buf), "wg_debug_read_timing")
        return list(buf)

    # -- device-resident call (inputs already in HBM) ------------------------------------------------
    def infer_device(self, mel, z=None, sigma=1.0, deterministic=False, out=None, lengths=None):
        """mel [B,T,n_mel] / z [B,32T,8] / out [B,256T]: float32 CUDA tensors on this engine's device.
        `lengths`: optional per-utterance frame counts (<= T) of a padded batch -- wg_infer_ragged: utterance b is
        computed exactly as if it were passed alone with its own length; the waveform tail beyond 256*lengths[b] is 0.
        Asynchronous on torch's current stream."""
        import torch
        dev = torch.device("cuda", self.device)
        if mel.device != dev or mel.dtype != torch.float32 or mel.dim() != 3 or mel.shape[2] != self.hp.n_mel_channels:
            raise ValueError(f"mel must be a float32 [B,T,{self.hp.n_mel_channels}] tensor on {dev}")
        mel = mel.contiguous()
        B, T = int(mel.shape[0]), int(mel.shape[1])
        Lg = T * HOP // self.hp.n_group
        if not deterministic:
            if z is None:
                raise ValueError("z is required unless deterministic (the Runtime layer draws it when omitted)")
            if z.device != dev or z.dtype != torch.float32 or tuple(z.shape) != (B, Lg, self.hp.n_group):
                raise ValueError(f"z must be a float32 [{B},{Lg},{self.hp.n_group}] tensor on {dev}, got {tuple(z.shape)}")
            z = z.contiguous()
        if out is None:
            out = torch.empty(B, T * HOP, dtype=torch.float32, device=dev)
        elif out.device != dev or out.dtype != torch.float32 or tuple(out.shape) != (B, T * HOP) or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 [{B},{T * HOP}] tensor on {dev}")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if lengths is not None:
            lens = self._lengths(lengths, B, T)
            ws_ptr, ws_bytes = self._workspace(B, T, lens)
            rc = self._lib.wg_infer_ragged(self._h, mel.data_ptr(), 0 if deterministic else z.data_ptr(), float(sigma),
                                           int(bool(deterministic)), B, T, lens, out.data_ptr(), ws_ptr, ws_bytes, stream)
            self._check(rc, "wg_infer_ragged")
            return out
        ws_ptr, ws_bytes = self._workspace(B, T)
        rc = self._lib.wg_infer(self._h, mel.data_ptr(), 0 if deterministic else z.data_ptr(), float(sigma),
                                int(bool(deterministic)), B, T, out.data_ptr(), ws_ptr, ws_bytes, stream)
        self._check(rc, "wg_infer")
        return out

    # -- host call through the C ABI's own staging (no torch on the data path) --------------------------
    def infer_host(self, mel, z=None, sigma=1.0, deterministic=False, lengths=None):
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim != 3 or mel.shape[2] != self.hp.n_mel_channels or mel.shape[0] == 0 or mel.shape[1] == 0:
            raise ValueError(f"mel must be a non-empty float32 [B,T,{self.hp.n_mel_channels}] array, got {mel.shape}")
        B, T = mel.shape[0], mel.shape[1]
        out = np.empty((B, T * HOP), dtype=np.float32)
        f32p = ctypes.POINTER(ctypes.c_float)
        zp = None
        if not deterministic:
            if z is None:
                raise ValueError("z is required unless deterministic (the Runtime layer draws it when omitted)")
            z = np.ascontiguousarray(z, dtype=np.float32)
            Lg = T * HOP // self.hp.n_group
            if z.shape != (B, Lg, self.hp.n_group):
                raise ValueError(f"z must be a float32 [{B},{Lg},{self.hp.n_group}] array, got {z.shape}")
            zp = z.ctypes.data_as(f32p)
        if lengths is not None:
            rc = self._lib.wg_infer_host_ragged(self._h, mel.ctypes.data_as(f32p), zp, float(sigma), int(bool(deterministic)),
                                                B, T, self._lengths(lengths, B, T), out.ctypes.data_as(f32p))
            self._check(rc, "wg_infer_host_ragged")
            return out
        rc = self._lib.wg_infer_host(self._h, mel.ctypes.data_as(f32p), zp, float(sigma), int(bool(deterministic)),
                                     B, T, out.ctypes.data_as(f32p))
        self._check(rc, "wg_infer_host")
        return out

    # -- test hooks ---------------------------------------------------------------------------------
    def debug_prefix(self, mel, z, sigma, stop_flow, stop_layer):
        import torch
        dev = torch.device("cuda", self.device)
        B, T = int(mel.shape[0]), int(mel.shape[1])
        M = B * T * HOP // self.hp.n_group
        h = torch.zeros(M, self.hp.n_channels, dtype=torch.float32, device=dev)
        acc = torch.zeros(M, 8, dtype=torch.float32, device=dev)
        ws_ptr, ws_bytes = self._workspace(B, T)
        rc = self._lib.wg_debug_infer_prefix(self._h, mel.data_ptr(), z.data_ptr(), float(sigma), 0, B, T, ws_ptr,
                                             ws_bytes, torch.cuda.current_stream(dev).cuda_stream, stop_flow,
                                             stop_layer, h.data_ptr(), acc.data_ptr())
        self._check(rc, "wg_debug_infer_prefix")
        return h, acc

    def debug_spect(self, B, T):
        import torch
        dev = torch.device("cuda", self.device)
        M = B * T * HOP // self.hp.n_group
        S = self.hp.n_mel_channels * self.hp.n_group
        out = torch.empty(M, S, dtype=torch.float32, device=dev)
        ws_ptr, _ = self._workspace(B, T)
        rc = self._lib.wg_debug_get_spect(self._h, B, T, ws_ptr, out.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
        self._check(rc, "wg_debug_get_spect")
        return out


def debug_gemm_bf16(A, W, bias=None):
    """D = A @ W^T + bias through the stand-alone tcgen05 GEMM (A [M,K], W [N,K] bf16 CUDA tensors)."""
    import torch
    lib = _lib.load_library()
    M, K = A.shape
    N = W.shape[0]
    D = torch.empty(M, N, dtype=torch.float32, device=A.device)
    rc = lib.wg_debug_gemm_bf16(A.data_ptr(), W.data_ptr(), 0 if bias is None else bias.data_ptr(), D.data_ptr(),
                                M, N, K, torch.cuda.current_stream(A.device).cuda_stream)
    if rc != _lib.WG_OK:
        raise WaveGlowError(f"wg_debug_gemm_bf16 failed ({rc}): {lib.wg_last_error(None).decode()}")
    return D
