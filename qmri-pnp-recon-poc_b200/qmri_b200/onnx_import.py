"""Trained-weight import from the reference's ONNX export, without the ``onnx`` package.

The reference exports its trained ``UNetRes`` with ``torch.onnx.export(..., export_params=True, opset_version=9,
do_constant_folding=True)`` (``PyTorch_Denoiser/utils.py:444-485``) and MATLAB reads it back with ``importONNXNetwork``
(``main_recon_tsmis_FFT.m:138-152``).  The weights of such a file are the ``initializer`` tensors of its graph; this module
reads exactly those from the protobuf wire format (ONNX ``ModelProto.graph = 7``, ``GraphProto.node = 1`` /
``initializer = 5``, ``TensorProto.dims = 1 / data_type = 2 / float_data = 4 / name = 8 / raw_data = 9``,
``NodeProto.input = 1 / op_type = 4``) and hands them to :class:`qmri_b200.UNetRes` as a ``state_dict``.

Tensor names: the TorchScript exporter keeps the parameter names (``m_head.weight``, ``m_down1.0.res.0.weight`` ...).  If a
file carries anonymous names instead, the weights are taken in the order the graph's ``Conv`` / ``ConvTranspose`` nodes consume
them, which is the network's forward order = the ``state_dict`` order (head, down1-3, body, up3-1, tail).
"""
from __future__ import annotations

import struct

import numpy as np

from .denoiser import state_dict_keys


def _varint(buf, pos):
    r, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not (b & 0x80):
            return r, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Yield (field number, wire type, value) of one protobuf message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _tensor(buf):
    dims, dtype, name, raw, floats = [], None, "", None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(v)
            else:  # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(d)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            if wt == 2:
                floats.append(np.frombuffer(bytes(v), dtype="<f4"))
            else:
                floats.append(np.array(struct.unpack("<f", v), np.float32))
        elif fno == 8:
            name = bytes(v).decode("utf-8", errors="replace")
        elif fno == 9:
            raw = bytes(v)
    if dtype != 1:
        return name, None   # not FLOAT: not a conv weight
    if raw is not None:
        arr = np.frombuffer(raw, dtype="<f4")
    elif floats:
        arr = np.concatenate(floats)
    else:
        arr = np.zeros(0, np.float32)
    if int(np.prod(dims)) != arr.size:
        raise ValueError(f"initializer {name!r}: {arr.size} values for dims {dims}")
    return name, arr.reshape(dims).astype(np.float32)


def read_initializers(path_or_bytes):
    """-> (dict name -> float32 array, list of weight-initializer names in Conv / ConvTranspose node order)."""
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = memoryview(bytes(path_or_bytes))
    else:
        with open(path_or_bytes, "rb") as f:
            data = memoryview(f.read())
    graph = None
    for fno, wt, v in _fields(data):
        if fno == 7 and wt == 2:
            graph = v
    if graph is None:
        raise ValueError("not an ONNX ModelProto: no graph (field 7)")
    inits, conv_order = {}, []
    for fno, wt, v in _fields(graph):
        if fno == 5 and wt == 2:
            name, arr = _tensor(v)
            if arr is not None:
                inits[name] = arr
        elif fno == 1 and wt == 2:
            inputs, op = [], ""
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 2:
                    inputs.append(bytes(v2).decode("utf-8", errors="replace"))
                elif f2 == 4 and w2 == 2:
                    op = bytes(v2).decode("utf-8", errors="replace")
            if op in ("Conv", "ConvTranspose") and len(inputs) >= 2:
                conv_order.append(inputs[1])
    return inits, conv_order


def load_onnx_state_dict(path_or_bytes):
    """Weights of an exported ``UNetRes`` as ``(state_dict, in_nc)``; raises if the file does not hold that network."""
    inits, conv_order = read_initializers(path_or_bytes)
    if "m_head.weight" in inits:
        in_nc = int(inits["m_head.weight"].shape[1])
        keys = state_dict_keys(in_nc)
        missing = [k for k, _ in keys if k not in inits]
        if missing:
            raise KeyError(f"ONNX file is missing UNetRes weights: {missing[:4]}")
        sd = {k: inits[k] for k, _ in keys}
    else:
        order = [n for n in conv_order if n in inits]
        if len(order) != 64:
            raise KeyError(f"ONNX graph has {len(order)} Conv / ConvTranspose weights, UNetRes has 64")
        in_nc = int(inits[order[0]].shape[1])
        sd = {k: inits[n] for (k, _), n in zip(state_dict_keys(in_nc), order)}
    for k, shape in state_dict_keys(in_nc):
        if tuple(sd[k].shape) != tuple(shape):
            raise ValueError(f"{k}: ONNX initializer has shape {tuple(sd[k].shape)}, UNetRes expects {tuple(shape)}")
    return sd, in_nc
