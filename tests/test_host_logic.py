"""Host-side logic of the Python mirror that needs no GPU: argument validation, key order, helpers."""
import numpy as np
import pytest

import qmri_b200 as q
from oracle import unetres


def test_state_dict_key_order_matches_reference_order():
    for in_nc in (10, 11):
        mine = q.state_dict_keys(in_nc)
        ref = unetres.layer_names(in_nc)
        assert [k for k, _ in mine] == [k for k, *_ in ref]
        sd = unetres.make_state_dict(in_nc, seed=0)
        for k, shape in mine:
            assert tuple(sd[k].shape) == shape


def test_build_noise_map():
    nm = q.build_noise_map(0.01, 224, 224)
    assert nm.shape == (224, 224) and nm.dtype == np.float64 and np.all(nm == 0.01)


def test_param_validation_without_gpu():
    from qmri_b200.admm import _make_params
    with pytest.raises(KeyError, match="param.iter"):
        _make_params({"gamma": 0.05}, [])
    with pytest.raises(TypeError, match="fft_operator"):
        _make_params({"iter": 1, "gamma": 0.05, "F": object(), "X0": 0, "net": lambda v: v}, [])


def test_denoise_wrapper_validates_input():
    class Net:
        def denoise(self, A):
            return A[:, :, :10]
    with pytest.raises(ValueError):
        q.denoiseImage_PnP_ADMM(np.full((8, 8, 10), np.nan), Net())
    with pytest.raises(ValueError):
        q.denoiseImage_PnP_ADMM(np.zeros((0, 8, 10)), Net())
    A = np.random.default_rng(0).random((8, 8, 10))
    assert np.array_equal(q.denoiseImage_PnP_ADMM(A, Net(), True, False), A)
    assert np.allclose(q.denoiseImage_PnP_ADMM(A, Net(), True, True), 0)


def test_checkpoint_import_reads_the_reference_format(tmp_path):
    """main_test.py:260-262 loads {'model_state_dict', 'epoch', 'loss'}; DataParallel prefixes and bare state_dicts are accepted too."""
    import torch
    import qmri_b200 as q
    from oracle import unetres
    sd = unetres.make_state_dict(11, seed=3)
    f = tmp_path / "ckpt.pt"
    torch.save({"model_state_dict": sd, "epoch": 7, "loss": 0.1}, f)
    got, in_nc = q.load_checkpoint_state_dict(str(f))
    assert in_nc == 11 and set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    got2, in_nc2 = q.load_checkpoint_state_dict({"module." + k: v for k, v in sd.items()})
    assert in_nc2 == 11 and set(got2) == set(sd)
    assert [k for k, _ in q.state_dict_keys(11)] == list(sd.keys())        # same key order as UNetRes(...).state_dict()
    import pytest
    with pytest.raises(KeyError):
        q.load_checkpoint_state_dict({"m_head.weight": sd["m_head.weight"], "m_head.bias": torch.zeros(64)})
    with pytest.raises(KeyError):
        q.load_checkpoint_state_dict({"something": torch.zeros(1)})


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port, no GPU): one JSON line with the contract's keys."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-iters", "1", "--atoms", "5000", "--ref-match-px", "256"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["metric"].startswith("ADMM slice-iterations/s") and line["unit"] == "slice-iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["scaling"] == "strong" and line["config"]["slices"] == 120 and line["config"]["iters_per_step"] == 100


# ---- ONNX initializer import (utils.py:444-485 writes the file, main_recon_tsmis_FFT.m:138-152 reads it) ----------------------
def _pb_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb_field(fno, wt, payload):
    tag = _pb_varint((fno << 3) | wt)
    return tag + (_pb_varint(len(payload)) + payload if wt == 2 else payload)


def _onnx_tensor(name, arr, raw=True, packed_dims=True):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    if packed_dims:
        body = _pb_field(1, 2, b"".join(_pb_varint(d) for d in arr.shape))
    else:
        body = b"".join(_pb_field(1, 0, _pb_varint(d)) for d in arr.shape)
    body += _pb_field(2, 0, _pb_varint(1))                    # data_type = FLOAT
    body += _pb_field(8, 2, name.encode())
    body += _pb_field(9, 2, arr.tobytes()) if raw else _pb_field(4, 2, arr.tobytes())
    return body


def _onnx_model(tensors, nodes, extra_int_tensor=True):
    """Minimal ModelProto the way torch.onnx.export lays it out: ir_version, producer, graph{node*, name, initializer*}, opset."""
    g = b""
    for op, inputs in nodes:
        n = b"".join(_pb_field(1, 2, i.encode()) for i in inputs) + _pb_field(2, 2, b"out") + _pb_field(4, 2, op.encode())
        g += _pb_field(1, 2, n)
    g += _pb_field(2, 2, b"torch-jit-export")
    for name, arr, raw, packed in tensors:
        g += _pb_field(5, 2, _onnx_tensor(name, arr, raw, packed))
    if extra_int_tensor:  # an int64 shape constant, as constant folding leaves behind: must be ignored
        t = _pb_field(1, 0, _pb_varint(2)) + _pb_field(2, 0, _pb_varint(7)) + _pb_field(8, 2, b"shape_const") + _pb_field(9, 2, np.array([1, 2], "<i8").tobytes())
        g += _pb_field(5, 2, t)
    return _pb_field(1, 0, _pb_varint(4)) + _pb_field(2, 2, b"pytorch") + _pb_field(7, 2, g) + _pb_field(8, 2, _pb_field(2, 0, _pb_varint(9)))


@pytest.mark.parametrize("named", [True, False])
def test_onnx_initializer_import(named, tmp_path):
    import qmri_b200 as q
    rng = np.random.default_rng(5)
    keys = q.state_dict_keys(11)
    sd = {k: (rng.standard_normal(shape) * 0.01).astype(np.float32) for k, shape in keys}
    names = {k: (k if named else f"onnx::Conv_{100 + 3 * i}") for i, (k, _) in enumerate(keys)}
    # initializers in shuffled order (the exporter does not promise state_dict order), mixed raw_data / float_data, packed / unpacked dims
    order = list(rng.permutation(len(keys)))
    tensors = [(names[keys[i][0]], sd[keys[i][0]], bool(i % 3), bool(i % 2)) for i in order]
    nodes = []
    for k, shape in keys:
        nodes.append(("ConvTranspose" if k.startswith("m_up") and k.endswith(".0.weight") else "Conv", ["act", names[k]]))
        nodes.append(("Relu", ["act"]))
    blob = _onnx_model(tensors, nodes)
    f = tmp_path / "net.onnx"
    f.write_bytes(blob)
    got, in_nc = q.load_onnx_state_dict(str(f))
    assert in_nc == 11 and [k for k in got] == [k for k, _ in keys]
    assert all(np.array_equal(got[k], sd[k]) for k in sd)
    got2, _ = q.load_onnx_state_dict(blob)                    # bytes are accepted too
    assert all(np.array_equal(got2[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        q.load_onnx_state_dict(b"\x08\x04")                    # no graph
    with pytest.raises(KeyError):
        q.load_onnx_state_dict(_onnx_model(tensors[:10], nodes[:20]))


def test_oracle_metrics_sanity():
    """oracle/metrics.py against closed forms: identical images, constant offset, and the window weights."""
    from oracle import metrics
    rng = np.random.default_rng(0)
    a = rng.random((40, 30))
    assert metrics.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    assert metrics.psnr(a, a) == np.inf
    assert metrics.psnr(a, a + 0.1) == pytest.approx(20.0, abs=1e-9)           # mse 0.01 -> 10 log10(1/0.01)
    # constant images: sigma terms vanish, ssim = (2 mu_x mu_y + C1) / (mu_x^2 + mu_y^2 + C1)
    x, y = np.full((24, 24), 0.3), np.full((24, 24), 0.5)
    assert metrics.ssim(x, y) == pytest.approx((2 * 0.15 + 1e-4) / (0.09 + 0.25 + 1e-4), rel=1e-9)
    # hole filling: a ring keeps its interior, a ring broken on a diagonal still does (8-connected background needs a full gap)
    pd = np.zeros((20, 20))
    pd[5:15, 5:15] = 1.0
    pd[7:13, 7:13] = 0.0
    m = metrics.getmask_fromPD(pd, 0.15)
    assert m[9, 9] == 1 and m[0, 0] == 0 and m.sum() == 100
    pd[5, 5] = 0.0                                                               # corner pixel removed: the hole touches it only diagonally
    m = metrics.getmask_fromPD(pd, 0.15)
    assert m[9, 9] == 1                                                          # (6,6) is still ring, so the background cannot enter
    pd[6, 6] = 0.0                                                               # now a diagonal 8-connected path of zeros reaches the hole
    m = metrics.getmask_fromPD(pd, 0.15)
    assert m[9, 9] == 0
    # Philox known answers (Random123 kat_vectors)
    assert [hex(v) for v in metrics.philox4x32_10(np.zeros(4, np.uint64), (0, 0))] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in metrics.philox4x32_10(np.full(4, 0xFFFFFFFF, np.uint64), (0xFFFFFFFF, 0xFFFFFFFF))] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
