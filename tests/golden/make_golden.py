"""Generates the golden fixtures in this directory.  Run in the build container only:

    python tests/golden/make_golden.py

* unetres_{10,11}ch.npz - produced by IMPORTING THE REFERENCE's own network
  (/root/reference/PyTorch_Denoiser/zhang_dpir_testing_code/network_unet.py:68-117, UNetRes) with
  torch.manual_seed(0) default-initialised weights (main_train.py:247,264-266) on a seeded input;
  stores the input, the reference output and per-layer weight checksums.  These pin oracle/unetres.py
  (and through it the CUDA denoiser) to the reference implementation.
* appendix_a.json - the known-answer values of SURVEY.md Appendix A (an independent float64
  restatement made during the survey); the MATLAB parts of the reference cannot run here, so these
  are the only external pins for masks / operator / x-update / matching ("parity unpinned").
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/PyTorch_Denoiser"


def unetres_fixture(in_nc):
    sys.path.insert(0, REF)
    from zhang_dpir_testing_code.network_unet import UNetRes
    torch.manual_seed(0)
    net = UNetRes(in_nc=in_nc, out_nc=10, nc=[64, 128, 256, 512], nb=4, act_mode="R",
                  downsample_mode="strideconv", upsample_mode="convtranspose").eval()
    sd = net.state_dict()
    torch.manual_seed(100 + in_nc)
    x = torch.rand(1, in_nc, 32, 40)
    if in_nc == 11:
        x[:, 10] = 0.01
    with torch.no_grad():
        y = net(x)
    keys = list(sd.keys())
    sums = np.array([float(sd[k].double().sum()) for k in keys])
    sq = np.array([float((sd[k].double() ** 2).sum()) for k in keys])
    np.savez_compressed(os.path.join(HERE, f"unetres_{in_nc}ch.npz"), x=x.numpy(), y=y.numpy(), keys=np.array(keys),
                        weight_sum=sums, weight_sumsq=sq, nparams=np.array(sum(v.numel() for v in sd.values())))


def appendix_a():
    a = {
        "spiral_counts": [620, 615, 618, 619, 621, 613, 621, 621, 621, 615],
        "spiral_idx_sums_1based": [15734107, 15349027, 15459027, 15477829, 15504623, 15191686, 15517756, 15511852, 15456821, 15203314],
        "spiral_first5_frame1": [1, 5, 10, 17, 26], "spiral_last3_frame1": [50165, 50171, 50175],
        "spiral_first5_frame10": [1, 4, 9, 15, 24], "spiral_last3_frame10": [50166, 50172, 50176],
        "epi_count": 672, "epi_idx_sums_1based": {"1": 16828896, "2": 16829568, "10": 16834944},
        "epi_first5_frame1": [2, 67, 132, 226, 291],
        "X0_norm2": 1.260914984248e+05,
        "spiral": {"nmeas": 6184, "y_norm2": 3.124627321012e+04, "y1": [2.189889727512e+00, 0.0],
                   "y2": [1.893362639834e-01, 7.895610818791e-02], "yend": [-3.238625689794e-01, -9.212102677081e-01],
                   "x0_111": [9.820602753177e-02, -3.741220677461e-01], "x0_end": [-3.762419983709e-02, -2.627527254764e-01],
                   "x_norm2": 1.089099617982e+05, "x_111": [1.138065603836e-01, -6.286806154603e-02],
                   "x_100_50_5": [6.803444208467e-01, 2.255400393693e-02]},
        "epi": {"nmeas": 6720, "y_norm2": 6.178596834661e+03, "y1": [-5.431534163719e-02, -5.551850572933e-03],
                "y2": [-7.178162927376e-04, 6.883457740922e-04], "yend": [1.512647044792e-02, -4.165966072218e-02],
                "x0_111": [9.270777601293e-02, 8.492868021432e-03], "x0_end": [1.865734490314e-02, 2.342772577123e-02],
                "x_norm2": 1.041382994863e+05, "x_111": [1.522659404176e-01, -3.248275374206e-02],
                "x_100_50_5": [6.691885559181e-01, 6.113009617138e-02]},
        "match": {"K": 1000, "dm_sum": 17205590, "dm_first5": [319, 319, 319, 319, 319], "dm_25000": 319,
                  "T1_sum": 7.199068769e+04, "T2_sum": 8.303992052e+03, "abs_pd_sum": 2.397099755e+04,
                  "pd1": [-5.178696521e-01, -5.458844639e-02], "min_gap": 5.9e-08, "n_gap_lt_1e-6": 22, "n_gap_lt_1e-4": 2402},
    }
    with open(os.path.join(HERE, "appendix_a.json"), "w") as f:
        json.dump(a, f, indent=1)


if __name__ == "__main__":
    unetres_fixture(10)
    unetres_fixture(11)
    appendix_a()
    print("golden fixtures written to", HERE)
