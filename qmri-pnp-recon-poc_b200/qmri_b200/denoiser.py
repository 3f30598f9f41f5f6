"""Denoiser plug-in: the reference's DRUNet (`UNetRes`) and its MATLAB wrapper.

Mirrors:

* ``UNetRes(in_nc, out_nc=10, nc=[64,128,256,512], nb=4, act_mode='R', 'strideconv',
  'convtranspose')`` - ``PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:68-117``
  (configuration of ``PyTorch_Denoiser/main_train.py:247``), weights handed over as the
  ``state_dict()`` (64 bias-free tensors, PyTorch layouts).
* ``denoiseImage_PnP_ADMM(A, net, onnx_dagnetwork, residual_noise)`` -
  ``main_files/utils/denoiseImage_PnP_ADMM.m:1-117``.
* ``build_noise_map(noise_std, rows, cols)`` - ``main_files/utils/build_noise_map.m:16-34``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import Context, as_f, check, dtype_code, ptr

NC = (64, 128, 256, 512)
NB = 4


def state_dict_keys(in_nc=10):
    """Key order of ``UNetRes(...).state_dict()`` with the expected shapes."""
    keys = [("m_head.weight", (NC[0], in_nc, 3, 3))]
    for lvl in range(3):
        for b in range(NB):
            for j in (0, 2):
                keys.append((f"m_down{lvl + 1}.{b}.res.{j}.weight", (NC[lvl], NC[lvl], 3, 3)))
        keys.append((f"m_down{lvl + 1}.{NB}.weight", (NC[lvl + 1], NC[lvl], 2, 2)))
    for b in range(NB):
        for j in (0, 2):
            keys.append((f"m_body.{b}.res.{j}.weight", (NC[3], NC[3], 3, 3)))
    for lvl in (2, 1, 0):
        keys.append((f"m_up{lvl + 1}.0.weight", (NC[lvl + 1], NC[lvl], 2, 2)))  # ConvTranspose2d: [Cin, Cout, kh, kw]
        for b in range(1, NB + 1):
            for j in (0, 2):
                keys.append((f"m_up{lvl + 1}.{b}.res.{j}.weight", (NC[lvl], NC[lvl], 3, 3)))
    keys.append(("m_tail.weight", (10, NC[0], 3, 3)))
    return keys


def load_checkpoint_state_dict(path_or_obj):
    """Weights of a trained ``UNetRes`` as the reference saves them: ``torch.save({'model_state_dict': net.state_dict(),
    'epoch': ..., 'loss': ...}, path)`` (``PyTorch_Denoiser/main_train.py``; read back at ``main_test.py:260-262``).
    Accepts the path of such a file, the loaded dict, or a bare ``state_dict``; strips ``module.`` prefixes left by
    ``DataParallel``.  Returns ``(state_dict, in_nc)`` with ``in_nc`` read off ``m_head.weight``."""
    obj = path_or_obj
    if isinstance(obj, (str, bytes)) or hasattr(obj, "__fspath__"):
        import torch
        obj = torch.load(obj, map_location="cpu", weights_only=True)
    if not isinstance(obj, dict):
        raise TypeError("checkpoint must be a dict (state_dict, or a dict holding 'model_state_dict')")
    for k in ("model_state_dict", "state_dict", "model"):
        if k in obj and isinstance(obj[k], dict):
            obj = obj[k]
            break
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in obj.items()}
    if "m_head.weight" not in sd:
        raise KeyError("checkpoint holds no UNetRes weights (m_head.weight missing)")
    in_nc = int(tuple(sd["m_head.weight"].shape)[1])
    want = dict(state_dict_keys(in_nc))
    extra = [k for k in sd if k not in want]
    if extra:
        raise KeyError(f"unexpected tensors in the checkpoint (this build runs the bias-free UNetRes of main_train.py:247): {extra[:4]}")
    return sd, in_nc


class UNetRes:
    """The built-in on-device denoiser.  ``state_dict`` maps the reference's keys to arrays
    (numpy or torch tensors)."""

    @classmethod
    def from_checkpoint(cls, path_or_obj, ctx=None):
        """Trained-weight import (``main_test.py:260-262``): see ``load_checkpoint_state_dict``."""
        sd, in_nc = load_checkpoint_state_dict(path_or_obj)
        return cls(sd, in_nc=in_nc, ctx=ctx)

    @classmethod
    def from_onnx(cls, path_or_bytes, ctx=None):
        """Import of the reference's ONNX export (``utils.py:444-485``; ``importONNXNetwork`` at
        ``main_recon_tsmis_FFT.m:138-152``): the graph's initializers are the weights - see ``onnx_import``."""
        from .onnx_import import load_onnx_state_dict
        sd, in_nc = load_onnx_state_dict(path_or_bytes)
        return cls(sd, in_nc=in_nc, ctx=ctx)

    def __init__(self, state_dict, in_nc=10, ctx=None):
        self.ctx = ctx or Context.default()
        self.in_nc = int(in_nc)
        arrays = []
        for key, shape in state_dict_keys(self.in_nc):
            if key not in state_dict:
                raise KeyError(f"state_dict is missing {key}")
            w = state_dict[key]
            if hasattr(w, "detach"):
                w = w.detach().cpu().numpy()
            w = np.ascontiguousarray(np.asarray(w, dtype=np.float32))
            if tuple(w.shape) != shape:
                raise ValueError(f"{key}: expected shape {shape}, got {tuple(w.shape)}")
            arrays.append(w)
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        h = C.c_void_p()
        check(self.ctx.lib.qmri_unetres_load(self.ctx.handle, self.in_nc, C.cast(ptrs, C.c_void_p), len(arrays), C.byref(h)))
        self.handle = h

    def set_precision(self, mode):
        """'fp32' (CUDA cores, exact mode) or 'tc' (tcgen05 split-bf16 tensor path)."""
        code = {"fp32": 0, "tc": 1}.get(mode, mode)
        check(self.ctx.lib.qmri_unetres_set_precision(self.handle, int(code)))

    def forward(self, x):
        """PyTorch layout: ``S x in_nc x H x W`` float32 -> ``S x 10 x H x W``."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        squeeze = x.ndim == 3
        if squeeze:
            x = x[None]
        S, Cin, H, W = x.shape
        if Cin != self.in_nc:
            raise ValueError(f"network expects {self.in_nc} input channels, got {Cin}")
        out = np.zeros((S, 10, H, W), np.float32)
        check(self.ctx.lib.qmri_unetres_forward(self.handle, ptr(x), ptr(out), S, H, W))
        return out[0] if squeeze else out

    __call__ = forward

    def denoise(self, A):
        """MATLAB layout: ``H x W x in_nc [x S]`` -> ``H x W x 10 [x S]``, same class as A."""
        A = as_f(A)
        if np.iscomplexobj(A):
            raise TypeError("denoiser input must be real")  # validateattributes(... 'real' ...), denoiseImage_PnP_ADMM.m:121
        batched = A.ndim == 4
        S = A.shape[3] if batched else 1
        H, W, Cin = A.shape[:3]
        if Cin != self.in_nc:
            raise ValueError(f"network expects {self.in_nc} input channels, got {Cin}")
        shape = (H, W, 10, S) if batched else (H, W, 10)
        out = np.zeros(shape, A.dtype, order="F")
        check(self.ctx.lib.qmri_unetres_denoise(self.handle, ptr(A), dtype_code(A), ptr(out), dtype_code(out), S, H, W))
        return out

    def flops(self, S, H, W):
        return float(self.ctx.lib.qmri_unetres_flops(self.handle, S, H, W))

    def close(self):
        if self.handle:
            self.ctx.lib.qmri_unetres_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def denoiseImage_PnP_ADMM(A, net, onnx_dagnetwork=True, residual_noise=False):
    """``I = denoiseImage_PnP_ADMM(A, net, onnx_dagnetwork, residual_noise)``.

    ``onnx_dagnetwork`` only selects how MATLAB addresses the last conv layer
    (``denoiseImage_PnP_ADMM.m:83-96``) and has no numerical effect; anything but a
    boolean is the reference's "print and return" soft failure (``:92-95``).
    ``residual_noise=True`` returns ``A - CNN(A)`` (``:100-102``)."""
    A = np.asarray(A)
    if A.size == 0 or A.ndim > 4 or not np.all(np.isfinite(A)):
        raise ValueError("A must be a nonempty finite real array with at most 4 dimensions")  # :119-127
    if not isinstance(onnx_dagnetwork, (bool, np.bool_)):
        print("Error: onnx_dagnetwork not set for denoiseImage_PnP_ADMM()")
        return None
    if not isinstance(residual_noise, (bool, np.bool_)):
        print("Error: residual_noise not set for denoiseImage_PnP_ADMM()")
        return None
    res = net.denoise(A)
    if residual_noise:
        if A.shape[2] != res.shape[2]:
            raise ValueError("residual_noise needs equal input and output channel counts")
        return (A.astype(np.float32) - res.astype(np.float32)).astype(A.dtype)
    return res


def build_noise_map(noise_std, rows, cols):
    """``repmat(noise_std, rows, cols)`` (``build_noise_map.m:34``)."""
    return np.tile(np.atleast_2d(np.asarray(noise_std, dtype=np.float64)), (int(rows), int(cols)))
