#!/usr/bin/env python
"""bench.py - PnP-ADMM MRF reconstruction hot path on B200 (contract: see the task's bench section).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--slices S] [--iters I]

Metric (BASELINE.json): ADMM iterations/s on a 224 x 224 x 10 slice batch (slice-iterations per second,
whole job over all N GPUs).  Workload at N = 1: BASELINE configs[1] - PnP-ADMM, single slice per GPU, spiral
mask (771-sample curve, 6184 measurements), rho = 0.05, 100 iterations, random-init DRUNet (UNetRes 10->10)
denoiser.  One "step" = one x = PnP_ADMM(y, param) reconstruction of the per-GPU slice batch (`iters`
iterations: fused x-update kernel + 64-conv denoiser each).  N > 1: every rank reconstructs its own slices
(slice sharding, no collective on the path) -> weak scaling.

  value : slice-iterations/s with y, X0 already resident in HBM (qmri_admm_run only), device-timed
  e2e   : the same through the reference-facing call PnP_ADMM(y, param) with HOST buffers: H2D of y and X0 and
          D2H of x inside the timed region
  roofline       : dominant kernel = the denoiser's conv kernels (tensor-pipe bound), timed live at the workload's slice count;
                   `traffic` = DRAM bytes per forward from the committed ncu capture (profiles/traffic.json)
  roofline_fwd_S15 : the same forward at 15 slices per GPU (machine filled), timed live
  roofline_k1/k2 : the x-update kernels (HBM bound, 120 slices) and dictionary matching (fp32 pipe), timed live
  clocks         : NVML samples taken during a second, identical timed pass (its time is reported too)
  cpu_baseline   : the oracle loop (NumPy x-update + PyTorch-CPU UNetRes) on the host cores, bounded sample

`--impl reference` times that CPU implementation as the reference arm (no MATLAB/Octave exists here; the
reference is MATLAB + PyTorch, so its CPU path is restated by oracle/ - kind "port").
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "qmri-pnp-recon-poc_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_IMG, C_CH = 224, 10
SPIRAL_S = 771
RHO = 0.05
FLOPS_PER_SLICE_FWD = 213.25e9


def load_traffic():
    """DRAM bytes per unit of work from the committed ncu captures (profiles/traffic.json; how each was taken is in the file)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def synthetic_slices(S, seed):
    """Seeded smooth brain-like 10-channel real TSMIs (N x M x C x S).  Throughput does not depend on content."""
    rng = np.random.default_rng(seed)
    n = np.arange(N_IMG)[:, None, None, None] / N_IMG - 0.5
    m = np.arange(N_IMG)[None, :, None, None] / N_IMG - 0.5
    c = np.arange(C_CH)[None, None, :, None]
    s = rng.uniform(0.8, 1.2, size=(1, 1, 1, S))
    brain = ((n / 0.42) ** 2 + (m / 0.36) ** 2 <= 1.0) * 1.0
    X = brain * (np.cos(0.35 * c) * np.exp(-4 * (n ** 2 + m ** 2)) * s + 0.15 * np.sin(9 * n * s + c) * np.cos(7 * m))
    X = X + 0.01 * rng.standard_normal(X.shape) * brain
    X *= np.sign(X[:, :, 0:1, :] + 1e-30)  # channel 1 non-negative, like main_synthesize_tsmis.m:97-98
    return np.asfortranarray(X)


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region through NVML.  Samples are taken by the timing loop itself once
    all timed steps have been enqueued (the GPU is then busy executing them, the host has nothing to launch): a background
    `nvidia-smi -lms` child or NVML thread was measured to stall this process's kernel launches for tens of ms per query
    (driver lock), inflating the single-slice loop by up to 4x.  Falls back to one nvidia-smi query per sample."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index, period_s=0.05):
        self.idx = gpu_index
        self.period = period_s
        self.sm, self.reasons, self.mx = [], set(), None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.idx
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[self.idx])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.sample()              # first query outside the timed region (lazy NVML initialisation takes ~0.5 s)
        self.sm, self.reasons = [], set()

    def sample(self):
        try:
            self._sample()
        except Exception:
            pass

    def _sample(self):
        if self.nvml is not None:
            self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            for nm, bit in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(nm)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.sm.append(float(f[0]))
        self.mx = float(f[1])
        for nm, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
            if val.lower().startswith("active"):
                self.reasons.add(nm)

    def stop(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_reference_loop(n_iters, threads, seed=0):
    """The reference's CPU path restated (oracle/): NumPy exact x-update + PyTorch CPU UNetRes, one slice."""
    import torch
    from oracle import sampling, unetres
    from oracle.admm import pnp_admm
    torch.set_num_threads(threads)
    P = sampling.setup_subsampling_spiralgrided(N_IMG, N_IMG, SPIRAL_S, np.eye(C_CH))
    F = sampling.FOperator(P)
    X = synthetic_slices(1, seed)[..., 0]
    Y = F.forward(X)
    X0 = F.adjoint(Y)
    sd = unetres.make_state_dict(10, seed=0)
    param = {"iter": n_iters, "gamma": RHO, "F": F, "X0": X0, "net": lambda v: unetres.denoise_matlab_layout(sd, v)}
    # warm-up of the conv kernels (oneDNN primitive creation)
    unetres.denoise_matlab_layout(sd, np.zeros((N_IMG, N_IMG, 10)))
    t0 = time.perf_counter()
    pnp_admm(Y, param, solver="closed")
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_iters = max(1, args.ref_iters)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_loop(1, threads)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference_loop(sample_iters, threads)
    value = args.steps * sample_iters / t
    line = {
        "metric": "ADMM slice-iterations/s (224x224x10 slice batch)", "value": value, "unit": "slice-iterations/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 x-update / f32 denoiser", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: PnP-ADMM single slice, spiral 771, rho 0.05, random-init UNetRes 10->10",
                   "slices_per_gpu": 1, "iters_per_step": sample_iters},
        "cpu_baseline": {"value": value, "unit": "slice-iterations/s", "cores": threads, "kind": "port",
                         "sample": f"{sample_iters} ADMM iterations of one slice per step (oracle: NumPy closed-form x-update + PyTorch-CPU UNetRes); "
                                   "no MATLAB/Octave in the image, the reference's MATLAB+PyTorch CPU path is restated by oracle/"},
        "e2e": {"value": value, "unit": "slice-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import torch
    import qmri_b200 as q
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = q.Context(local_rank)
    # a dedicated torch stream shared with the library: CUDA events recorded on it see the library's kernels
    # (the legacy default stream has handle 0, which qmri_ctx_set_stream reads as "use your own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    peaks = load_peaks()
    S, iters = args.slices, args.iters

    # ---- problem (seeded synthetic data; weights random-init like the reference without its ONNX blobs) ----
    V = np.eye(C_CH)
    P = q.setup_subsampling_spiralgrided(N_IMG, N_IMG, SPIRAL_S, V, ctx=ctx)
    F = q.fft_operator(P)
    X = synthetic_slices(S, seed=1000 + rank)
    Y = F.forward(X)
    rng = np.random.default_rng(rank)
    Y = Y + (rng.standard_normal(Y.shape) + 1j * rng.standard_normal(Y.shape)) * np.sqrt(np.mean(np.abs(Y) ** 2) / 10 ** 3.0 / 2)
    X0 = F.adjoint(Y)
    sd = make_weights()
    net = q.UNetRes(sd, in_nc=10, ctx=ctx)
    if args.precision:
        net.set_precision(args.precision)
    param = {"iter": iters, "gamma": RHO, "cg_tol": 1e-4, "F": F, "X0": X0, "net": net, "denoiser_type": "single_level"}
    sess = q.AdmmSession(param, S)

    # pinned host buffers for the end-to-end leg
    def pinned(a):
        t = torch.empty(a.size * (2 if np.iscomplexobj(a) else 1), dtype=torch.float64).pin_memory()
        v = t.numpy().view(a.dtype).reshape(a.shape, order="F")
        v[...] = a
        return t, v
    _ty, Yp = pinned(np.asfortranarray(Y.reshape(P.nmeas, S)))
    _tx, X0p = pinned(np.asfortranarray(X0.reshape(N_IMG, N_IMG, C_CH, S)))
    _to, Xout = pinned(np.zeros((N_IMG, N_IMG, C_CH, S), np.complex128, order="F"))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup, sampler=None, collective=True, label=None):
        """collective=False: a leg only rank 0 runs (no barrier / all-reduce, which the other ranks would never join)."""
        sync = barrier if collective else torch.cuda.synchronize
        for _ in range(warmup):
            flush.zero_()          # warm-up mirrors the timed loop (the first fill kernel launch loads its module lazily)
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        l0 = ctx.launch_count
        t_host = time.time()
        e0.record(stream)
        for i in range(steps):
            flush.zero_()          # evict L2 between steps (256 MiB > 126 MB L2)
            fn()
            marks[i].record(stream)
        e1.record(stream)
        t_host = time.time() - t_host
        if sampler is not None:    # every step is enqueued: sample clocks while the GPU works through them (host idle)
            sampler.sample()
            while not e1.query():
                time.sleep(0.1)    # sparse on purpose: every NVML query measurably perturbs the running kernels
                sampler.sample()
        sync()
        ms = e0.elapsed_time(e1)
        if rank == 0 and label:
            per = [round(([e0] + marks)[i].elapsed_time(marks[i]), 2) for i in range(steps)]
            sys.stderr.write(f"[bench] {label}: per-step ms {per}, host enqueue {t_host * 1e3:.1f} ms\n")
        return (max_over_ranks(ms) if collective else ms), ctx.launch_count - l0

    # ---- value: loop only, inputs resident in HBM -----------------------------------------------------
    sess.upload_raw(Yp, X0p)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # Two identical timed passes back to back.  Pass 1 carries no NVML query at all and gives `value`; pass 2 is timed the
    # same way while rank 0 samples clocks / throttle reasons during it.  Both times are reported (clocks.sampled_pass_ms_per_step):
    # on some boxes a single NVML query stalls the running kernels for tens of ms, which would otherwise be booked on the kernels.
    ms_dev, launches = timed(lambda: sess.run(iters), args.steps, args.warmup, label="value pass")
    ms_sampled, _ = timed(lambda: sess.run(iters), args.steps, 0, sampler if rank == 0 else None, label="clock-sampled pass")
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled_pass_ms_per_step"] = ms_sampled / args.steps
        clocks["how"] = "NVML queries during a second, identical timed pass run right after the one `value` is taken from"
    value = world * S * iters * args.steps / (ms_dev * 1e-3)

    # ---- e2e: the public call with host buffers ----------------------------------------------------------
    def e2e_step():
        sess.upload_raw(Yp, X0p)
        sess.run(iters)
        sess.download(Xout)        # synchronises
    ms_e2e, _ = timed(e2e_step, args.steps, 1, label="e2e pass")
    e2e_value = world * S * iters * args.steps / (ms_e2e * 1e-3)

    # ---- live roofline legs -------------------------------------------------------------------------------
    import ctypes as C
    hw = N_IMG * N_IMG
    vin = torch.rand(S * 10 * hw, device="cuda")
    vout = torch.empty_like(vin)
    def fwd():
        q._capi.check(ctx.lib.qmri_unetres_forward_dev(net.handle, C.c_void_p(vin.data_ptr()), C.c_void_p(vout.data_ptr()), None, None, S, N_IMG, N_IMG))
    ms_fwd, n_fwd_launch = timed(fwd, 5, 3)
    fwd_tflops = net.flops(S, N_IMG, N_IMG) * 5 / (ms_fwd * 1e-3) / 1e12
    tensor_peak = peaks["bf16_tflops_sustained"]
    traffic = load_traffic()
    t_fwd = traffic.get(f"unetres_forward_S{S}")
    roofline = {"bound": "tensor", "kernel": "UNetRes forward (64 conv launches; 3x3 convs = 97% of flops)", "achieved": fwd_tflops,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": fwd_tflops / tensor_peak,
                "traffic": (t_fwd["dram_read_bytes"] + t_fwd["dram_write_bytes"]) if t_fwd else None,
                "traffic_note": (t_fwd or {}).get("how", "no committed ncu capture for this slice count"),
                "peak_source": f"{peaks['src']} dense bf16 (sustained); the fp32 exact mode runs on CUDA cores, see DESIGN.md",
                "ms_per_forward": ms_fwd / 5, "precision_mode": args.precision}
    # K1 at a batch larger than L2 (algorithmic bytes: 20 B per pixel-channel per iteration)
    roof_k1 = None
    if rank == 0 and world == 1 and not args.skip_extra:  # single-GPU run only: the scaling runs stay short
        S1 = args.k1_slices
        Xb = synthetic_slices(S1, seed=5)
        Yb = F.forward(Xb)
        sess1 = q.AdmmSession(dict(param, X0=np.zeros((N_IMG, N_IMG, C_CH, S1)), net=net), S1)
        sess1.upload(Yb, F.adjoint(Yb))
        sess1.run(1)
        reps = 20
        ms_k1, _ = timed(lambda: sess1.xupdate_only(reps), 3, 2, collective=False)
        t_launch = ms_k1 * 1e-3 / (3 * reps)
        gbs = 20.0 * hw * C_CH * S1 / t_launch / 1e9
        t_k1 = traffic.get(f"k1_xupdate_S{S1}")
        roof_k1 = {"bound": "hbm", "kernel": "x-update = stream_fwd_kernel + stream_solve_kernel + stream_adj_kernel (3 launches; the single cluster "
                                             "kernel xupdate_kernel serves one-slice batches)",
                   "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": gbs / peaks["hbm_gbs"], "traffic": (t_k1["dram_read_bytes"] + t_k1["dram_write_bytes"]) if t_k1 else None,
                   "traffic_note": (t_k1 or {}).get("how", "no committed ncu capture for this slice count"),
                   "slices": S1, "us_per_slice_iteration": 1e6 * t_launch / S1,
                   "algorithmic_bytes_per_launch": 20 * hw * C_CH * S1}
        sess1.close()
        # K2: one slice against a 100k-atom dictionary, complex data (40 flop per px-atom)
        K = args.match_atoms
        rngd = np.random.default_rng(3)
        D = rngd.standard_normal((K, C_CH)).astype(np.float32)
        D /= np.linalg.norm(D, axis=1, keepdims=True)
        d = q.Dictionary({"D": D, "normD": np.ones(K, np.float32), "lut": rngd.random((K, 2)).astype(np.float32)}, ctx=ctx)
        npix = hw * max(1, min(S, 4))
        xr = torch.randn(C_CH * npix, device="cuda")
        xi = torch.randn(C_CH * npix, device="cuda")
        qm = torch.empty(2 * npix, device="cuda")
        pdv = torch.empty(2 * npix, device="cuda")
        def match():
            q._capi.check(ctx.lib.qmri_match_dev(d.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(xi.data_ptr()), npix,
                                                 C.c_void_p(qm.data_ptr()), C.c_void_p(pdv.data_ptr()), None, None))
        ms_m, _ = timed(match, 3, 2, collective=False)
        pxa = npix * K * 3 / (ms_m * 1e-3)
        fp32_peak = 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
        roof_k2 = {"bound": "fp32", "kernel": "match_kernel", "achieved": pxa * 40 / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                   "frac": pxa * 40 / 1e12 / fp32_peak, "px_atoms_per_s": pxa, "atoms": K, "pixels": npix}
        d.close()
    else:
        roof_k2 = None

    # (runs after the x-update / matching legs: fifteen slices of tensor work pull the GPU into its power cap for a while)
    # the same forward with the machine filled (15 slices = BASELINE configs[2]'s per-GPU batch): what the conv kernels reach when
    # a layer has enough tiles for 148 SMs - the single-slice figure above is bound by per-layer latency (64 dependent layers)
    roof_fwd15 = None
    if rank == 0 and world == 1 and not args.skip_extra and S != 15:
        S15 = 15
        vin15 = torch.rand(S15 * 10 * hw, device="cuda")
        vout15 = torch.empty_like(vin15)
        def fwd15():
            q._capi.check(ctx.lib.qmri_unetres_forward_dev(net.handle, C.c_void_p(vin15.data_ptr()), C.c_void_p(vout15.data_ptr()), None, None, S15, N_IMG, N_IMG))
        ms15, _ = timed(fwd15, 5, 3, collective=False)
        tf15 = net.flops(S15, N_IMG, N_IMG) * 5 / (ms15 * 1e-3) / 1e12
        t15 = traffic.get("unetres_forward_S15")
        roof_fwd15 = {"bound": "tensor", "kernel": "UNetRes forward, 15 slices", "achieved": tf15, "peak": tensor_peak, "unit": "TFLOP/s",
                      "frac": tf15 / tensor_peak, "frac_of_3_product_ceiling": 3 * tf15 / tensor_peak,
                      "traffic": (t15["dram_read_bytes"] + t15["dram_write_bytes"]) if t15 else None, "ms_per_forward": ms15 / 5,
                      "note": "split-bf16 operands need 3 bf16 products per fp32 product (1e-4 parity bar): ceiling = peak / 3"}
        del vin15, vout15
    if rank == 0:
        threads = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.skip_cpu:
            t = cpu_reference_loop(args.cpu_iters, threads)
            cpu = {"value": args.cpu_iters / t, "unit": "slice-iterations/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_iters} ADMM iterations of one slice (oracle: NumPy closed-form x-update + PyTorch-CPU UNetRes, "
                             f"{threads} threads); MATLAB/Octave absent - reference CPU path restated by oracle/"}
        line = {
            "metric": "ADMM slice-iterations/s (224x224x10 slice batch)", "value": value, "unit": "slice-iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (x-update fp32 FFT; denoiser " + args.precision + ")", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: PnP-ADMM, spiral 771 (6184 meas/slice), rho 0.05, random-init UNetRes 10->10",
                       "slices_per_gpu": S, "iters_per_step": iters, "l2": "256 MiB buffer written between steps",
                       "parallelism": f"slice-sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": "slice-iterations/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(Yp.nbytes + X0p.nbytes), "d2h_bytes_per_step": int(Xout.nbytes)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_fwd_S15": roof_fwd15,
            "roofline_k1": roof_k1,
            "roofline_k2": roof_k2,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    sess.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def make_weights():
    """torch.manual_seed(0) default-init UNetRes weights in the reference's construction order (main_train.py:247,264-266)."""
    import torch
    import torch.nn as nn
    import qmri_b200 as q
    torch.manual_seed(0)
    sd = {}
    for key, shape in q.state_dict_keys(10):
        if re.fullmatch(r"m_up\d\.0\.weight", key):
            m = nn.ConvTranspose2d(shape[0], shape[1], 2, 2, 0, bias=False)
        elif shape[2] == 2:
            m = nn.Conv2d(shape[1], shape[0], 2, 2, 0, bias=False)
        else:
            m = nn.Conv2d(shape[1], shape[0], 3, 1, 1, bias=False)
        sd[key] = m.weight.detach()
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slices", type=int, default=1, help="slices per GPU (configs[1] = 1; configs[2] = 15)")
    ap.add_argument("--iters", type=int, default=100, help="ADMM iterations per reconstruction (param.iter)")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"], help="denoiser precision mode: tc (tcgen05 split-bf16, default) / fp32 (CUDA cores)")
    ap.add_argument("--k1-slices", type=int, default=120)
    ap.add_argument("--match-atoms", type=int, default=100000)
    ap.add_argument("--cpu-iters", type=int, default=12)
    ap.add_argument("--ref-iters", type=int, default=4)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-extra", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand without torchrun: start one rank per GPU the way the driver does
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
