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
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-iters", "1"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["metric"].startswith("ADMM slice-iterations/s") and line["unit"] == "slice-iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
