"""Oracle: the DRUNet denoiser (``UNetRes``) restated with torch functional ops.

Test infrastructure (see ``oracle/__init__.py``).  Follows
``PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:68-117`` with the
configuration of ``PyTorch_Denoiser/main_train.py:247``
(``nc=[64,128,256,512], nb=4, act_mode='R', strideconv / convtranspose``, all
``bias=False``) and the blocks of ``basicblock.py:61-98`` (conv),
``:211-223`` (ResBlock = x + conv(relu(conv(x)))), ``:413-419`` (ConvTranspose2d
2x2 stride 2) and ``:437-443`` (Conv2d 2x2 stride 2).

It exists because ``/root/reference`` does not travel to the GPU box: the
restatement does.  ``tests/test_oracle_denoiser.py`` checks it bit-for-bit
against the imported reference module (same ``state_dict``) whenever
``/root/reference`` is present, and against golden vectors generated from that
import otherwise.

``make_state_dict`` constructs the layers in the reference's construction order
so that ``torch.manual_seed(s)`` + default PyTorch init (``main_train.py:264-266``
does nothing else) yields the same random weights as
``torch.manual_seed(s); UNetRes(...)``.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as Fn

NC = (64, 128, 256, 512)
NB = 4


def layer_names(in_nc=10):
    """Ordered (key, kind, cin, cout) list of the 64 bias-free convolutions."""
    L = [("m_head.weight", "c3", in_nc, NC[0])]
    for lvl in range(3):
        for b in range(NB):
            for j in (0, 2):
                L.append((f"m_down{lvl + 1}.{b}.res.{j}.weight", "c3", NC[lvl], NC[lvl]))
        L.append((f"m_down{lvl + 1}.{NB}.weight", "down", NC[lvl], NC[lvl + 1]))
    for b in range(NB):
        for j in (0, 2):
            L.append((f"m_body.{b}.res.{j}.weight", "c3", NC[3], NC[3]))
    for lvl in (2, 1, 0):
        L.append((f"m_up{lvl + 1}.0.weight", "up", NC[lvl + 1], NC[lvl]))
        for b in range(1, NB + 1):
            for j in (0, 2):
                L.append((f"m_up{lvl + 1}.{b}.res.{j}.weight", "c3", NC[lvl], NC[lvl]))
    L.append(("m_tail.weight", "c3", NC[0], 10))
    return L


def make_state_dict(in_nc=10, seed=0, dtype=torch.float32):
    """Seeded default-init weights, identical to ``torch.manual_seed(seed); UNetRes(in_nc, 10, ...)``."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    for key, kind, cin, cout in layer_names(in_nc):
        if kind == "c3":
            m = nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
        elif kind == "down":
            m = nn.Conv2d(cin, cout, 2, 2, 0, bias=False)
        else:
            m = nn.ConvTranspose2d(cin, cout, 2, 2, 0, bias=False)
        sd[key] = m.weight.detach().to(dtype)
    return sd


def _res(x, sd, prefix):
    y = Fn.conv2d(x, sd[prefix + ".res.0.weight"], padding=1)
    y = Fn.relu(y)
    y = Fn.conv2d(y, sd[prefix + ".res.2.weight"], padding=1)
    return x + y


def unetres_forward(sd, x0):
    """``UNetRes.forward`` (``network_unet.py:106-117``); x0 is NCHW."""
    x1 = Fn.conv2d(x0, sd["m_head.weight"], padding=1)
    skips = [x1]
    x = x1
    for lvl in (1, 2, 3):
        for b in range(NB):
            x = _res(x, sd, f"m_down{lvl}.{b}")
        x = Fn.conv2d(x, sd[f"m_down{lvl}.{NB}.weight"], stride=2)
        skips.append(x)
    x4 = x
    for b in range(NB):
        x = _res(x, sd, f"m_body.{b}")
    for lvl in (3, 2, 1):
        x = x + skips[lvl]
        x = Fn.conv_transpose2d(x, sd[f"m_up{lvl}.0.weight"], stride=2)
        for b in range(1, NB + 1):
            x = _res(x, sd, f"m_up{lvl}.{b}")
    return Fn.conv2d(x + skips[0], sd["m_tail.weight"], padding=1)


def denoise_matlab_layout(sd, A):
    """``denoiseImage_PnP_ADMM(A, net, true, false)`` (``denoiseImage_PnP_ADMM.m:73-114``).

    A: numpy H x W x Cin (MATLAB layout, values in [0,1]) -> H x W x 10, computed
    in single precision and cast back to the input class.  Layout convention of
    ``PyTorch_Denoiser/utils.py:349-366``: ``t[c,h,w] = A[h,w,c]``.
    """
    import numpy as np
    a = np.asarray(A)
    t = torch.from_numpy(np.ascontiguousarray(np.transpose(a, (2, 0, 1))[None]).astype(np.float32))
    with torch.no_grad():
        o = unetres_forward(sd, t)[0]
    return np.transpose(o.numpy(), (1, 2, 0)).astype(a.dtype)
