#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the numbagg hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of one hot-path function over one batch of synthetic input (SURVEY.md
8(d) generators).  Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  the oracle port (oracle/nbg_oracle.c, OpenMP over rows like numba's parallel
                target) timed on this box's host cores on a bounded sample of the same workload
  e2e           same metric through the public numpy API: H2D of pinned inputs + kernels + D2H
`--impl reference` times the CPU implementation only (no GPU work), same metric/config.

Multi-GPU (torchrun): every rank runs the same per-GPU batch on its own device (independent
row shards, no data-path collective => weak scaling); time = max over ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed `ncu --set full` capture of this exact workload (profiles/r01_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return float(json.load(f)[workload]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ workloads
# name -> (family, func, dtype, rows, n, params).  Shapes: BASELINE.json configs / SURVEY 8(d).
WORKLOADS = {
    # configs[1]: low-cardinality grouped reductions, labels shared by all rows
    "cfg2_group_nansum": ("group", "group_nansum", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
    "cfg2_group_nanmean": ("group", "group_nanmean", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
    "cfg2_group_nanstd": ("group", "group_nanstd", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
    # configs[0]: README 2-D move_mean (tiny: launch-latency bound) and its >= 1 GB scaling
    "cfg1_move_mean": ("move", "move_mean", "f64", 100, 10_000, dict(window=20, min_count=1)),
    "cfg1s_move_mean": ("move", "move_mean", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    "cfg1s_move_sum": ("move", "move_sum", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    "cfg1s_move_std": ("move", "move_std", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    "cfg1s_move_var": ("move", "move_var", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    "cfg1s_move_cov": ("move", "move_cov", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    "cfg1s_move_corr": ("move", "move_corr", "f64", 2000, 100_000, dict(window=20, min_count=1)),
    # configs[2]: one long core axis, 30 % NaN
    "cfg3_move_exp_nanmean": ("exp", "move_exp_nanmean", "f64", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_ffill": ("fill", "ffill", "f64", 1, 1_000_000_000, dict()),
    "cfg3_bfill": ("fill", "bfill", "f64", 1, 1_000_000_000, dict()),
    # configs[3]: wide windows on float32, 10 % NaN, min_count=500
    "cfg4_move_std": ("move", "move_std", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_cov": ("move", "move_cov", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_corr": ("move", "move_corr", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    # configs[4]: high-cardinality 1-D grouped reductions (per-element labels)
    "cfg5_group_nansum1d": ("group1d", "group_nansum", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
    "cfg5_group_nanargmax": ("group1d", "group_nanargmax", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
    "cfg5_group_nanfirst": ("group1d", "group_nanfirst", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
    "cfg5_group_nanvar": ("group1d", "group_nanvar", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
}
# SURVEY 8(f) rank 1 (plain NaN reductions) on the config-2 / config-4 shapes
WORKLOADS.update({
    "red_nansum_f32": ("reduce", "nansum", "f32", 10_000, 1_000_000, dict(axis=-1)),
    "red_nanmean_f32": ("reduce", "nanmean", "f32", 10_000, 1_000_000, dict(axis=-1)),
    "red_nanvar_f32": ("reduce", "nanvar", "f32", 10_000, 1_000_000, dict(axis=-1)),
    "red_nanmax_f32": ("reduce", "nanmax", "f32", 10_000, 1_000_000, dict(axis=-1)),
    "red_nanargmax_f32": ("reduce", "nanargmax", "f32", 10_000, 1_000_000, dict(axis=-1)),
    "red_nansum_f64": ("reduce", "nansum", "f64", 2000, 1_000_000, dict(axis=-1)),
    "red_nanstd_f64": ("reduce", "nanstd", "f64", 2000, 1_000_000, dict(axis=-1)),
    "red_nansum_f32_axis0": ("reduce", "nansum", "f32", 1_000_000, 1000, dict(axis=0)),
    "red_nanvar_f64_axis0": ("reduce", "nanvar", "f64", 1_000_000, 1000, dict(axis=0)),
    "red_nanmean_f32_short": ("reduce", "nanmean", "f32", 10_000_000, 100, dict(axis=-1)),
    "red_nansum_f64_all": ("reduce", "nansum", "f64", 1, 1_000_000_000, dict(axis=None)),
    # selection (multi-pass by construction; roofline counts ONE read)
    "quant_median_long": ("quantile", "nanquantile", "f64", 2000, 1_000_000, dict(quantiles=0.5, axis=-1)),
    "quant_quartiles_short": ("quantile", "nanquantile", "f64", 1_000_000, 1000, dict(quantiles=[0.25, 0.5, 0.75], axis=-1)),
    # matrix functions, (obs, vars) -> (obs, vars, vars); write-bound by construction
    "mat_move_cov": ("matrix", "move_covmatrix", "f64", 200_000, 32, dict(window=100, min_count=10)),
})
DEFAULT_WORKLOAD = "cfg2_group_nansum"
TWO_INPUT = {"move_cov", "move_corr", "move_exp_nancov", "move_exp_nancorr"}
NP_DT = {"f32": np.float32, "f64": np.float64}


def nan_frac(family):
    return 0.3 if family in ("exp", "fill") else 0.1


def alg_bytes(family, func, dt, rows, n, params):
    """Compulsory traffic: inputs read once, outputs written once."""
    s = 4 if dt == "f32" else 8
    if family in ("move", "exp", "fill"):
        return rows * n * s * (3 if func in TWO_INPUT else 2)
    if family == "reduce":
        outs = {None: 1, -1: rows, 0: n}[params["axis"]]
        return rows * n * s + outs * 8
    if family == "quantile":
        return rows * n * s + rows * 8 * np.size(params["quantiles"])
    if family == "matrix":
        return rows * n * s + rows * n * n * s
    K = params["num_labels"]
    if family == "group":
        return rows * n * s + n * 8 + rows * K * s
    return rows * n * (s + 8) + K * s  # group1d: int64 label per element


def host_batch(family, func, dt, rows, n, params, seed=0):
    """Synthetic input on the host (bounded rows) -- SURVEY 8(d) generators."""
    rs = np.random.RandomState(seed)
    a = rs.rand(rows, n)
    a = np.where(a > nan_frac(family), a, np.nan).astype(NP_DT[dt])
    args = [a if family != "group1d" else a.reshape(-1)]
    if func in TWO_INPUT:
        args.append((a.astype(np.float64) ** 2 + 1).astype(NP_DT[dt]))
    kwargs = dict(params)
    if family == "reduce" and rows == 1:
        args = [a.reshape(-1)]
    if family == "group":
        args.append(np.random.RandomState(0).randint(0, params["num_labels"], size=n).astype(np.int64))
        kwargs["axis"] = -1
    elif family == "group1d":
        args.append(np.random.RandomState(0).randint(0, params["num_labels"], size=rows * n).astype(np.int64))
    if family == "exp" and dt == "f32":
        kwargs["alpha"] = np.float32(kwargs["alpha"])
    return args, kwargs


# ------------------------------------------------------------------------------ CPU baseline
def cpu_baseline(family, func, dt, rows, n, params, budget_s=12.0, reps=3):
    """Oracle port on the host cores, bounded sample (rows or a prefix of a 1-D input)."""
    from oracle import oracle

    cores = os.cpu_count() or 1
    if family in ("group1d",) or rows == 1:
        # 1-D inputs use ONE core by construction in the reference (gufunc parallelism spans
        # outer dims only); sample a prefix
        sn = min(n, 50_000_000 if family != "group1d" else 20_000_000)
        srows, used = 1, 1
    else:
        per_row = n
        # ~2e8 elements per call (the quantile port sorts row by row in NumPy: 2e7)
        srows = max(1, min(rows, int((2.0e7 if family == "quantile" else 2.0e8) // per_row)))
        # the quantile port is a NumPy loop; the matrix port parallelises over batch items (one here)
        sn, used = n, (1 if family in ("quantile", "matrix") else min(cores, srows))
    p = dict(params)
    if family == "group1d":
        p = dict(params)
    args, kwargs = host_batch(family, func, dt, srows, sn, p)
    f = getattr(oracle, func)
    f(*args, **kwargs)  # warm
    ts = []
    t_end = time.perf_counter() + budget_s
    for _ in range(reps):
        t0 = time.perf_counter()
        f(*args, **kwargs)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() > t_end:
            break
    best = min(ts)
    elems = srows * sn
    return dict(
        value=elems / best, unit="elements/s", cores=used, kind="port",
        sample=f"{func} on {srows}x{sn} {dt} (oracle/nbg_oracle.c, OpenMP over rows, best of {len(ts)}); host has {cores} cores",
    ), (srows, sn)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 3 + i and s[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.samples))


# ------------------------------------------------------------------------------ main arms
def run_reference(args, wl):
    family, func, dt, rows, n, params = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, (srows, sn) = cpu_baseline(family, func, dt, rows, n, params, budget_s=60.0, reps=max(1, args.steps))
    line = dict(
        impl="reference", metric="elements/s", value=base["value"], unit="elements/s", n_gpus=args.gpus,
        steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * srows * sn / base["value"],
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype=dt, data="synthetic",
        config=dict(workload=args.workload, func=func, shape=[rows, n], sample_shape=[srows, sn], **_jsonable(params)),
        cpu_baseline=base,
        e2e=dict(value=base["value"], unit="elements/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line))


def _jsonable(d):
    return {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in d.items()}


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import numbagg_b200 as nb
    from numbagg_b200 import decorators as D

    family, func, dt, rows, n, params = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    tdt = torch.float32 if dt == "f32" else torch.float64

    # ---- device-resident synthetic input (per-GPU batch; generated on device: 40 GB does not
    # fit the dev container's host RAM)
    g = torch.Generator(device=device).manual_seed(rank)
    shape = (rows, n) if family != "group1d" else (rows * n,)
    a = torch.empty(shape, dtype=tdt, device=device)
    chunk = 1 << 28
    flat = a.view(-1)
    for s in range(0, flat.numel(), chunk):
        seg = flat[s:s + chunk]
        seg.uniform_(0, 1, generator=g)
        seg[seg <= nan_frac(family)] = float("nan")
    tensors = [a]
    if func in TWO_INPUT:
        tensors.append(a * a + 1)
    labels = None
    if family == "group":
        labels = torch.from_numpy(np.random.RandomState(0).randint(0, params["num_labels"], size=n)).to(device)
    elif family == "group1d":
        labels = torch.randint(0, params["num_labels"], (rows * n,), generator=g, device=device, dtype=torch.int64)

    qdev = torch.tensor(np.atleast_1d(params["quantiles"]), dtype=torch.float64, device=device) if family == "quantile" else None

    def step_device():
        if family == "move":
            return D.run_move(func, tensors, params["window"], params["min_count"], -1)
        if family == "exp":
            al = float(np.float32(params["alpha"])) if dt == "f32" else params["alpha"]
            return D.run_move_exp(func, tensors, al, 0.0, -1)[0]
        if family == "fill":
            return D.run_fill(func, a, n, -1)[0]
        if family == "reduce":
            axes = (0, 1) if params["axis"] is None else (params["axis"] % 2,)
            return D.run_reduce(func, a, axes)
        if family == "quantile":
            return D.run_quantile(a, qdev, (params["axis"] % 2,))
        if family == "matrix":
            return D.run_matrix(func, a, window=params["window"], min_count=params["min_count"])
        v2 = a if family == "group" else a.view(1, -1)
        return D.run_group(func, v2, labels, params["num_labels"], 1)

    elements = rows * n
    abytes = alg_bytes(family, func, dt, rows, n, params)

    # ---- timed region: device-resident
    for _ in range(max(3, args.warmup)):
        out = step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    # Launch-bound workloads (a step is a few microseconds of GPU work, less than the host
    # needs to issue it): the K steps are captured once into a CUDA graph and replayed, so
    # the timed region holds exactly the K steps' kernels and no Python.
    use_graph = abytes < (256 << 20) and not args.no_graph
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(args.steps):
                out = step_device()
        launches = nb.launch_count() - launches0
        graph.replay()  # warm replay
        torch.cuda.synchronize()
        ev0.record()
        graph.replay()
        ev1.record()
    else:
        ev0.record()
        for _ in range(args.steps):
            out = step_device()
        ev1.record()
        launches = None
    torch.cuda.synchronize()
    if launches is None:
        launches = nb.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    # keep the clocks sampled for at least ~1.5 s of the same load so the record means something
    t_extra = time.perf_counter()
    while rank == 0 and time.perf_counter() - t_extra < 1.5:
        step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        dist.barrier()
    ms_step = ms_total / args.steps
    value = world * elements / (ms_step * 1e-3)

    # ---- e2e through the public numpy API on a bounded host batch: every step copies the
    # pinned host inputs to the device, runs the kernels and copies the result back to the host
    budget = 4 << 30  # bytes of host input per step
    item = a.element_size()
    if rows > 1:
        erows = max(1, min(rows, budget // (n * item * len(tensors))))
        en = n
    else:
        erows = 1
        en = min(n, budget // (item * len(tensors) + (8 if family == "group1d" else 0)))
        en -= en % 4
    e_args, e_kwargs = host_batch(family, func, dt, erows, en, params, seed=rank)
    pinned = []
    for x in e_args:
        px = nb.empty_pinned(x.shape, x.dtype)
        px[...] = x
        pinned.append(px)
    del e_args
    f_public = getattr(nb, func)
    res = f_public(*pinned, **e_kwargs)
    h2d = sum(x.nbytes for x in pinned)
    d2h = res.nbytes

    for _ in range(2):
        f_public(*pinned, **e_kwargs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        res = f_public(*pinned, **e_kwargs)  # H2D + kernels + D2H (returns a host array)
    torch.cuda.synchronize()
    e_s = (time.perf_counter() - t0) / e_steps
    if world > 1:
        t = torch.tensor([e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_s = float(t.item())
    e2e_value = world * erows * en / e_s

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = abytes / (ms_step * 1e-3) / 1e9
        base, _ = cpu_baseline(family, func, dt, rows, n, params)
        line = dict(
            metric="elements/s", value=value, unit="elements/s", n_gpus=world, steps=args.steps,
            warmup=max(3, args.warmup), ms_per_step=ms_step, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype=dt, data="synthetic",
            config=dict(workload=args.workload, func=func, shape=[rows, n], per_gpu_batch=[rows, n],
                        nan_fraction=nan_frac(family), l2="inputs larger than L2 (no flush needed)" if abytes > (256 << 20) else "input smaller than L2: launch-latency bound config, reported as is",
                        e2e_batch=[erows, en], launch="cuda graph of the K steps" if use_graph else "stream", **_jsonable(params)),
            clocks=clocks,
            e2e=dict(value=e2e_value, unit="elements/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=e_s * 1e3, note="public numpy API; pinned host input; H2D + kernels + D2H timed"),
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=measured_traffic(args.workload), peak_source=peak_src,
                          note="algorithmic bytes per step / CUDA-event step time (all kernels of the step)"),
            cpu_baseline=base,
        )
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true",
                    help="launch-bound workloads: issue the steps from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
