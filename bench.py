#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the numbagg hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
                    [--no-parity] [--no-sharded] [--no-e2e] [--no-graph]

One "step" = one pass of one hot-path function over the WHOLE BASELINE configuration (SURVEY.md
8(d) generators).  Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM copy peak
  cpu_baseline  (N=1) the reference's own Numba path (oracle/_ref, `kind: "reference"`) -- or, when
                that copy is missing, the C port (oracle/nbg_oracle.c, `kind: "port"`) -- timed on
                this box's host cores on a bounded sample of the same workload
  e2e           same metric through the public numpy API: H2D of pinned inputs + kernels + D2H
  parity        full-size spot checks of THIS run's outputs against the oracle (observed error
                next to the tolerance)
  sharded       (N>1) the core-axis / element-sharded forms of configs 3-5 over NCCL: elements/s,
                bytes exchanged, and parity against the unsharded single-GPU result
`--impl reference` times the CPU implementation only (no GPU work), same metric/config.

Multi-GPU (torchrun): STRONG scaling -- the fixed configuration is split across the ranks: rows
with no data-path collective where the config has rows (cfg1/2/4), the core axis / the elements
with one exchange step where it is one long slice (cfg3: carry all-gather, cfg5: per-label
partial all-reduce).  time = max over ranks, value = all elements / that time.
"""

from __future__ import annotations

import os
import sys

_RANK = int(os.environ.get("RANK", "0"))
if _RANK == 0:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms (rank 0 only) must see the whole
    # host.  Set BEFORE numpy / torch / numba load an OpenMP runtime.
    _n = str(os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = _n
    os.environ["NUMBA_NUM_THREADS"] = _n

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
GEN_BLOCK = 25_000_000     # elements per deterministic generation block (shards align to it)


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed `ncu --set full` capture of this exact workload (profiles/r02_traffic.json, falling
    back to round 1's file)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)[workload]["dram_bytes_per_launch"])
        except Exception:
            continue
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
    "cfg2_group_nancount": ("group", "group_nancount", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
    "cfg2_group_nanmax": ("group", "group_nanmax", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
    "cfg2_group_nanargmax": ("group", "group_nanargmax", "f32", 10_000, 1_000_000, dict(num_labels=1000)),
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
    "cfg3_move_exp_nansum": ("exp", "move_exp_nansum", "f64", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nancount": ("exp", "move_exp_nancount", "f64", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nanvar": ("exp", "move_exp_nanvar", "f64", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nanstd": ("exp", "move_exp_nanstd", "f64", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nancov": ("exp", "move_exp_nancov", "f64", 1, 500_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nancorr": ("exp", "move_exp_nancorr", "f64", 1, 500_000_000, dict(alpha=0.1)),
    "cfg3_move_exp_nanmean_f32": ("exp", "move_exp_nanmean", "f32", 1, 1_000_000_000, dict(alpha=0.1)),
    "cfg3_ffill": ("fill", "ffill", "f64", 1, 1_000_000_000, dict()),
    "cfg3_bfill": ("fill", "bfill", "f64", 1, 1_000_000_000, dict()),
    # configs[3]: wide windows on float32, 10 % NaN, min_count=500
    "cfg4_move_std": ("move", "move_std", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_var": ("move", "move_var", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_mean": ("move", "move_mean", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_cov": ("move", "move_cov", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    "cfg4_move_corr": ("move", "move_corr", "f32", 1000, 1_000_000, dict(window=1000, min_count=500)),
    # configs[4]: high-cardinality 1-D grouped reductions (per-element labels)
    "cfg5_group_nansum1d": ("group1d", "group_nansum", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
    "cfg5_group_nanmean": ("group1d", "group_nanmean", "f64", 1, 2_000_000_000, dict(num_labels=10_000_000)),
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
    "mat_move_corr": ("matrix", "move_corrmatrix", "f64", 200_000, 32, dict(window=100, min_count=10)),
})
DEFAULT_WORKLOAD = "cfg2_group_nansum"
TWO_INPUT = {"move_cov", "move_corr", "move_exp_nancov", "move_exp_nancorr"}
NP_DT = {"f32": np.float32, "f64": np.float64}
RTOL = {"f32": 1e-5, "f64": 1e-12}
EXACT = {"ffill", "bfill", "group_nancount", "group_nanargmax", "group_nanargmin", "group_nanfirst", "group_nanlast",
         "group_nanmax", "group_nanmin", "group_nanany", "group_nanall", "nancount", "nanargmax", "nanargmin",
         "nanmax", "nanmin", "allnan", "anynan"}


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


def shared_labels(n, K):
    return np.random.RandomState(0).randint(0, K, size=n).astype(np.int64)


def host_batch(family, func, dt, rows, n, params, seed=0):
    """Synthetic input on the host (bounded size) -- SURVEY 8(d) generators."""
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
        args.append(shared_labels(n, params["num_labels"]))
        kwargs["axis"] = -1
    elif family == "group1d":
        args.append(np.random.RandomState(1).randint(0, params["num_labels"], size=rows * n).astype(np.int64))
    if family == "exp" and dt == "f32":
        kwargs["alpha"] = np.float32(kwargs["alpha"])
    return args, kwargs


# ------------------------------------------------------------------------------ CPU arms
def _omp_team():
    """Threads an OpenMP parallel region gets in this process (what the C port will use)."""
    import ctypes

    for lib in ("libgomp.so.1", "libomp.so", "libiomp5.so"):
        try:
            return int(ctypes.CDLL(lib).omp_get_max_threads())
        except Exception:
            continue
    return None


def _cpu_sample_shape(family, rows, n):
    """Bounded sample of the workload (about 2e8 elements; whole rows where there are rows)."""
    if family == "group1d" or rows == 1:
        return 1, min(n, 100_000_000 if family != "group1d" else 20_000_000)
    budget = 2.0e7 if family == "quantile" else (4.0e7 if family == "matrix" else 2.0e8)
    return max(1, min(rows, int(budget // n))), n


def cpu_arm(family, func, dt, rows, n, params, budget_s=20.0, reps=3, allow_port=True):
    """Time the reference's own Numba path (oracle/_ref/numbagg) on the host cores; when that copy
    is not there, the C port.  Returns (dict for `cpu_baseline`, sample shape)."""
    cores = os.cpu_count() or 1
    srows, sn = _cpu_sample_shape(family, rows, n)
    one_slice = family == "group1d" or rows == 1  # gufunc parallelism spans outer dims only
    args, kwargs = host_batch(family, func, dt, srows, sn, params)
    kind, info, f = None, {}, None
    try:
        from oracle import ref_install

        numbagg, info = ref_install.import_reference(cores)
        f = getattr(numbagg, func)
        kind = "reference"
        used = 1 if one_slice else min(int(info["num_threads"]), srows)
        team = int(info["num_threads"])
        how = (f"numbagg {func} via numba {info['numba']} target={getattr(f, 'target', '?')}, "
               f"numba.get_num_threads()={team}")
    except Exception as ex:  # noqa: BLE001
        if not allow_port:
            raise
        from oracle import oracle

        f = getattr(oracle, func)
        kind = "port"
        team = _omp_team() or 1
        used = 1 if (one_slice or family in ("quantile", "matrix")) else min(team, srows)
        how = f"oracle/nbg_oracle.c (OpenMP over rows, omp_get_max_threads()={team}); reference import failed: {str(ex)[:80]}"
    if not one_slice and srows >= 2 and cores > 1 and team <= 1:
        raise RuntimeError(f"CPU arm would run on ONE thread of {cores} cores (OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}): "
                           "refusing to report a baseline that is not the parallel path")
    f(*args, **kwargs)  # JIT / warm
    ts = []
    t_end = time.perf_counter() + budget_s
    for _ in range(max(1, reps)):
        t0 = time.perf_counter()
        f(*args, **kwargs)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() > t_end:
            break
    best = min(ts)
    elems = srows * sn
    return dict(
        value=elems / best, unit="elements/s", cores=used, kind=kind,
        sample=f"{func} on {srows}x{sn} {dt}: {how}; best of {len(ts)}; host has {cores} cores"
               + ("; 1-D input: one core by construction (gufunc parallelism spans outer dims only)" if one_slice else ""),
    ), (srows, sn), ts


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


# ------------------------------------------------------------------------------ reference arm
def _jsonable(d):
    return {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in d.items()}


def run_reference(args, wl):
    family, func, dt, rows, n, params = wl
    if _RANK != 0:
        return  # one CPU arm per box: the other ranks exit 0 without work
    base, (srows, sn), ts = cpu_arm(family, func, dt, rows, n, params, budget_s=120.0,
                                    reps=max(1, args.steps) + max(0, args.warmup))
    line = dict(
        impl="reference", metric="elements/s", value=base["value"], unit="elements/s", n_gpus=args.gpus,
        steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * min(ts),
        higher_is_better=True, scaling="strong", vs_baseline=None, dtype=dt, data="synthetic",
        config=dict(workload=args.workload, func=func, shape=[rows, n], sample_shape=[srows, sn], **_jsonable(params)),
        cpu_baseline=base,
        e2e=dict(value=base["value"], unit="elements/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        gpu_launches=0,
    )
    print(json.dumps(line))


# ------------------------------------------------------------------------------ device data
def gen_flat(torch, device, dtype, lo, hi, nan_fraction, seed):
    """Elements [lo, hi) of the workload's flattened input, generated on the device in blocks of
    GEN_BLOCK elements each seeded by its block index -- any rank / any world size sees the same
    global array (lo and hi are multiples of GEN_BLOCK or the array end)."""
    out = torch.empty(hi - lo, dtype=dtype, device=device)
    g = torch.Generator(device=device)
    pos = lo
    while pos < hi:
        blk = pos // GEN_BLOCK
        end = min(hi, (blk + 1) * GEN_BLOCK)
        full = torch.empty(min(GEN_BLOCK, end - blk * GEN_BLOCK), dtype=dtype, device=device)
        g.manual_seed(seed * 1_000_003 + blk)
        full.uniform_(0, 1, generator=g)
        seg = full[pos - blk * GEN_BLOCK: end - blk * GEN_BLOCK]
        seg[seg <= nan_fraction] = float("nan")
        out[pos - lo: end - lo] = seg
        pos = end
    return out


def gen_labels_flat(torch, device, lo, hi, K, seed=7):
    out = torch.empty(hi - lo, dtype=torch.int64, device=device)
    g = torch.Generator(device=device)
    pos = lo
    while pos < hi:
        blk = pos // GEN_BLOCK
        end = min(hi, (blk + 1) * GEN_BLOCK)
        g.manual_seed(seed * 1_000_003 + blk)
        full = torch.randint(0, K, (min(GEN_BLOCK, end - blk * GEN_BLOCK),), generator=g, device=device, dtype=torch.int64)
        out[pos - lo: end - lo] = full[pos - blk * GEN_BLOCK: end - blk * GEN_BLOCK]
        pos = end
    return out


def shard_bounds(total, rank, world, unit):
    """[lo, hi) of `total` items for this rank, cut at multiples of `unit` items."""
    units = (total + unit - 1) // unit
    base, rem = divmod(units, world)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(total, lo_u * unit), min(total, hi_u * unit)


# ------------------------------------------------------------------------------ parity helpers
def compare(func, dt, got, exp, scale=0.0):
    """Observed error of `got` against the oracle's `exp` and the verdict under the north_star
    tolerance: bit-exact classes must be equal; the others |got-exp| <= rtol*|exp| + rtol*scale
    (scale = magnitude of the sums an output is a difference of; 0 where outputs are plain sums)."""
    got = np.asarray(got)
    exp = np.asarray(exp)
    if got.shape != exp.shape:
        return dict(ok=False, reason=f"shape {got.shape} vs {exp.shape}")
    nan_ok = bool(np.array_equal(np.isnan(got), np.isnan(exp))) if exp.dtype.kind == "f" else True
    if func in EXACT or exp.dtype.kind != "f":
        if exp.dtype.kind == "f":
            eq = (got == exp) | (np.isnan(got) & np.isnan(exp))
        else:
            eq = got == exp
        return dict(ok=bool(eq.all()) and nan_ok, compared=int(exp.size), tolerance="bit-exact", mismatches=int((~eq).sum()))
    rtol = RTOL[dt]
    g, e = got.astype(np.float64), exp.astype(np.float64)
    fin = np.isfinite(g) & np.isfinite(e)
    err = np.abs(g[fin] - e[fin])
    rel = err / np.maximum(np.abs(e[fin]), 1e-300)
    bound = rtol * np.abs(e[fin]) + rtol * scale
    return dict(ok=bool(nan_ok and (err <= bound).all()), compared=int(exp.size), nan_masks_equal=nan_ok,
                max_rel_err=float(rel.max()) if rel.size else 0.0,
                max_err_over_bound=float((err / np.maximum(bound, 1e-300)).max()) if err.size else 0.0,
                rtol=rtol, abs_floor=rtol * scale)


def parity_checks(torch, nb, D, oracle, wl, tensors, labels, out, step_on, lo_elem, ends_row=True):
    """Spot checks of this rank's full-size run against the oracle: bounded CPU work."""
    family, func, dt, rows, n, params = wl
    res = {}
    rs = np.random.RandomState(123)
    a = tensors[0]
    if family in ("group", "move", "reduce") and a.dim() == 2 and a.shape[0] > 1:
        # rows are independent: first rows + random rows of this rank's shard
        nrows = a.shape[0]
        per = max(1, int(1.3e8 // (a.shape[1] * (2 if func in TWO_INPUT else 1))))
        first = list(range(min(nrows, max(1, per // 2))))
        rand = sorted(set(rs.randint(0, nrows, size=max(1, per // 2)).tolist()))
        for tag, idx in (("first_rows", first), ("random_rows", rand)):
            ti = torch.tensor(idx, device=a.device)
            host = [t.index_select(0, ti).cpu().numpy() for t in tensors]
            got = out.index_select(0, ti).cpu().numpy()
            f = getattr(oracle, func)
            if family == "group":
                exp = f(host[0], labels.cpu().numpy(), num_labels=params["num_labels"], axis=-1)
                m = float(np.nanmax(np.abs(host[0])))
                per_group = host[0].shape[1] / params["num_labels"]
                scale = {"group_nansum": m * per_group, "group_nanmean": m, "group_nanstd": m * m, "group_nanvar": m * m}.get(func, 0.0)
                if func == "group_nanstd":  # compare variances: the floor belongs to the variance
                    res[tag] = compare("group_nanvar", dt, got.astype(np.float64) ** 2, exp.astype(np.float64) ** 2, scale)
                    res[tag]["compared_as"] = "variance"
                    continue
            elif family == "move":
                exp = f(*host, window=params["window"], min_count=params["min_count"])
                m = float(np.nanmax(np.abs(host[0])))
                scale = 0.0 if "corr" in func else (m * m if any(t in func for t in ("var", "std", "cov")) else m)
                if func == "move_std":
                    res[tag] = compare("move_var", dt, got.astype(np.float64) ** 2, exp.astype(np.float64) ** 2, scale)
                    res[tag]["compared_as"] = "variance"
                    continue
                if func == "move_corr":
                    # correlations divide by differences of running sums the reference never re-syncs: absolute
                    # floor 1e-9 (float64) / 1e-4 (float32) on a value in [-1, 1], the same as smoke() and the
                    # sharded block use (observed on config 1s, window 20: 9e-12)
                    scale = 1e3 if dt == "f64" else 10.0
            else:
                exp = f(host[0], axis=-1) if params["axis"] == -1 else None
                if exp is None:
                    continue
                scale = float(np.nanmax(np.abs(host[0]))) * (host[0].shape[1] if func == "nansum" and dt == "f32" else 0.0) * 6e-8 / RTOL[dt]
            res[tag] = compare(func, dt, got, exp, scale)
            res[tag]["rows"] = len(idx)
    elif family in ("exp", "fill"):
        nloc = a.shape[-1]
        flat_in = [t.reshape(-1) for t in tensors]
        flat_out = out.reshape(-1)
        f = getattr(oracle, func)
        kw = dict(alpha=params["alpha"]) if family == "exp" else {}
        big, small, warm = min(nloc, 100_000_000), min(nloc, 1_000_000), min(nloc, 20_000)
        if func != "bfill":
            if lo_elem == 0:  # this shard starts the row: its first 1e8 outputs, exactly as the reference scans them
                host = [t[:big].cpu().numpy() for t in flat_in]
                res["first_1e8"] = compare(func, dt, flat_out[:big].cpu().numpy(), f(*host, **kw))
            # the far end, with a warm-up long enough (2e4 >> 7070 steps of decay / a valid value) for the
            # state to be independent of what precedes it
            s0 = max(0, nloc - small - warm)
            host = [t[s0:].cpu().numpy() for t in flat_in]
            res["last_1e6"] = compare(func, dt, flat_out[nloc - small:].cpu().numpy(), f(*host, **kw)[-small:])
        else:
            if ends_row:  # bfill scans from the row end: the last 1e8 outputs exactly
                host = [t[nloc - big:].cpu().numpy() for t in flat_in]
                res["last_1e8"] = compare(func, dt, flat_out[nloc - big:].cpu().numpy(), f(*host, **kw))
            host = [t[: small + warm].cpu().numpy() for t in flat_in]
            res["first_1e6"] = compare(func, dt, flat_out[:small].cpu().numpy(), f(*host, **kw)[:small])
    elif family == "group1d":
        # per-label results of a 2e8-element prefix run through both
        m = min(a.numel(), 200_000_000)
        va, la = a.reshape(-1)[:m], labels[:m]
        got = step_on(va.view(1, -1), la).reshape(-1).cpu().numpy()
        exp = getattr(oracle, func)(va.cpu().numpy(), la.cpu().numpy(), num_labels=params["num_labels"])
        mval = float(np.nanmax(np.abs(va[:1_000_000].cpu().numpy())))
        per_group = m / params["num_labels"]
        scale = {"group_nansum": 0.0, "group_nanmean": 0.0, "group_nanvar": mval * mval}.get(func, 0.0)
        res["prefix_2e8"] = compare(func, dt, got, exp, scale)
        res["prefix_2e8"]["labels"] = params["num_labels"]
        res["prefix_2e8"]["elements_per_label"] = per_group
    return res


# ------------------------------------------------------------------------------ sharded block
def sharded_block(torch, dist, nd, D, device, rank, world, steps=10):
    """Core-axis / element-sharded forms of BASELINE configs 3-5 over NCCL, at
    full size: every rank regenerates the whole input (same seeds), computes the UNSHARDED result on
    its own GPU, runs the sharded form on its contiguous shard and compares its whole shard."""
    results = {}

    def timed(fn):
        for _ in range(5):  # NCCL sets collectives / p2p channels up lazily: first calls are not representative
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t.item())

    def agree(ok):
        t = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def same(x, y, rtol):
        if rtol == 0:
            return bool(torch.equal(torch.nan_to_num(x, nan=-7.0), torch.nan_to_num(y, nan=-7.0)))
        if not bool(torch.equal(torch.isnan(x), torch.isnan(y))):
            return False
        return bool(torch.allclose(x, y, rtol=rtol, atol=0.0, equal_nan=True))

    def max_rel(x, y):
        fin = torch.isfinite(x) & torch.isfinite(y)
        if not bool(fin.any()):
            return 0.0
        return float(((x[fin] - y[fin]).abs() / y[fin].abs().clamp_min(1e-300)).max().item())

    # ---- config 3: one row of 1e9 float64, 30 % NaN
    n = 1_000_000_000
    lo, hi = shard_bounds(n, rank, world, GEN_BLOCK)
    full = gen_flat(torch, device, torch.float64, 0, n, 0.3, seed=3).view(1, n)
    shard = full[:, lo:hi].contiguous()
    lens = [shard_bounds(n, r, world, GEN_BLOCK)[1] - shard_bounds(n, r, world, GEN_BLOCK)[0] for r in range(world)]
    ref = D.run_move_exp("move_exp_nanmean", [full], 0.1, 0.0, -1)[0]
    got, ms = timed(lambda: nd.move_exp_sharded("move_exp_nanmean", shard, alpha=0.1))
    results["cfg3_move_exp_nanmean"] = dict(
        elements_per_s=n / (ms * 1e-3), ms_per_step=ms, sharding="core axis, contiguous", collective="all_gather of per-shard aggregates (NCCL)",
        bytes_exchanged_per_rank=world * 11 * 8, passes_over_shard=1,
        head_recomputed=nd.exp_forget_length(0.1), parity_ok=agree(same(got, ref[:, lo:hi], 1e-12)),
        max_rel_err=max_rel(got, ref[:, lo:hi]), tolerance=1e-12, compared="whole shard vs unsharded single-GPU result")
    del ref, got
    for f in ("ffill", "bfill"):
        ref = D.run_fill(f, full, n, -1)[0]
        got, ms = timed(lambda: nd.fill_sharded(f, shard, shard_lens=lens))
        results[f"cfg3_{f}"] = dict(
            elements_per_s=n / (ms * 1e-3), ms_per_step=ms, sharding="core axis, contiguous",
            collective="all_gather of per-shard aggregates (NCCL) + in-place patch of the sentinel run",
            bytes_exchanged_per_rank=world * 3 * 8, passes_over_shard=1,
            parity_ok=agree(same(got, ref[:, lo:hi], 0)), tolerance="bit-exact", compared="whole shard vs unsharded single-GPU result")
        del ref, got
    del full, shard
    torch.cuda.empty_cache()

    # ---- config 4: 1000 x 1e6 float32, window 1000, core axis sharded (halo exchange)
    rows, n4, w = 1000, 1_000_000, 1000
    a = gen_flat(torch, device, torch.float32, 0, rows * n4, 0.1, seed=4).view(rows, n4)
    b = a * a + 1
    clo, chi = shard_bounds(n4, rank, world, 1000)
    lens4 = [shard_bounds(n4, r, world, 1000)[1] - shard_bounds(n4, r, world, 1000)[0] for r in range(world)]
    sa, sb = a[:, clo:chi].contiguous(), b[:, clo:chi].contiguous()
    for f, ins, full_ins in (("move_std", [sa], [a]), ("move_corr", [sa, sb], [a, b])):
        ref = D.run_move(f, full_ins, w, 500, -1)
        got, ms = timed(lambda: nd.move_sharded(f, *ins, window=w, min_count=500, shard_lens=lens4))
        refs = ref[:, clo:chi]
        # the head of a shard is recomputed from (halo, shard[:window]): its window sums are formed in a
        # different order than in the unsharded tiles (float32 outputs: rtol 1e-5)
        results[f"cfg4_{f}"] = dict(
            elements_per_s=rows * n4 / (ms * 1e-3), ms_per_step=ms, sharding="core axis, contiguous column blocks",
            collective="send/recv of the predecessor's last `window` columns (NCCL p2p), overlapped with the interior",
            bytes_exchanged_per_rank=len(ins) * rows * w * 4, passes_over_shard=1,
            parity_ok=agree(same(got, refs, 1e-5 if f == "move_std" else 1e-4)), max_rel_err=max_rel(got, refs),
            tolerance=1e-5 if f == "move_std" else 1e-4, compared="whole shard vs unsharded single-GPU result")
        del ref, got
    del a, b, sa, sb
    torch.cuda.empty_cache()

    # ---- config 5: 2e9 float64 elements, 1e7 labels, element shards (partial-state all-reduce)
    n5, K = 2_000_000_000, 10_000_000
    elo, ehi = shard_bounds(n5, rank, world, GEN_BLOCK)
    v = gen_flat(torch, device, torch.float64, 0, n5, 0.1, seed=5)
    lab = gen_labels_flat(torch, device, 0, n5, K)
    sv, sl = v[elo:ehi].view(1, -1), lab[elo:ehi]
    for f, tol, planes in (("group_nansum", 1e-12, 1), ("group_nanvar", 1e-10, 3), ("group_nanargmax", 0, 2)):
        ref = D.run_group(f, v.view(1, -1), lab, K, 1)
        got, ms = timed(lambda: nd.group_sharded(f, sv, sl, num_labels=K, index_offset=elo))
        results[f"cfg5_{f}"] = dict(
            elements_per_s=n5 / (ms * 1e-3), ms_per_step=ms, sharding="elements, contiguous",
            collective="all_reduce on the partial state's channel planes (NCCL): " +
                       {"group_nansum": "SUM", "group_nanvar": "SUM x3", "group_nanargmax": "MAX(key) then MIN(index | key == max)"}[f],
            bytes_exchanged_per_rank=planes * K * 8, parity_ok=agree(same(got, ref, tol)),
            max_rel_err=max_rel(got, ref) if tol else 0.0, tolerance=tol if tol else "bit-exact",
            compared="all 1e7 labels vs unsharded single-GPU result")
        del ref, got
    return results


# ------------------------------------------------------------------------------ our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import numbagg_b200 as nb
    from numbagg_b200 import decorators as D
    from numbagg_b200 import distributed as nd

    family, func, dt, rows, n, params = wl
    rank = _RANK
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    tdt = torch.float32 if dt == "f32" else torch.float64
    item = 4 if dt == "f32" else 8
    K = params.get("num_labels")

    # ---- this rank's share of the fixed configuration (strong scaling)
    core_sharded = world > 1 and rows == 1 and family in ("exp", "fill", "group1d")
    replica_mode = world > 1 and (family in ("quantile", "matrix") or (family == "reduce" and params["axis"] != -1)
                                  or (rows == 1 and not core_sharded))
    if rows > 1 and not replica_mode:
        unit = max(1, GEN_BLOCK // n) if (rows * n) % GEN_BLOCK == 0 and GEN_BLOCK % n == 0 else 1
        r_lo, r_hi = shard_bounds(rows, rank, world, unit)
        my_rows, my_n = r_hi - r_lo, n
        e_lo, e_hi = r_lo * n, r_hi * n
        sharding = "rows (independent slices, no data-path collective)" if world > 1 else "none"
    elif replica_mode:
        my_rows, my_n, e_lo, e_hi = rows, n, 0, rows * n
        sharding = "replicas only (this workload does not shard over rows)"
    else:
        e_lo, e_hi = shard_bounds(n, rank, world, GEN_BLOCK) if core_sharded else (0, n)
        my_rows, my_n = 1, e_hi - e_lo
        sharding = {"exp": "core axis (one carry exchange: all_gather of shard aggregates)",
                    "fill": "core axis (one carry exchange: all_gather of shard aggregates)",
                    "group1d": "elements (one combine: all_reduce of per-label partials)"}.get(family, "replicas") if world > 1 else "none"
    if (rows * n) % GEN_BLOCK == 0 or world == 1:
        flat = gen_flat(torch, device, tdt, e_lo, e_hi, nan_frac(family), seed=2)
    else:
        flat = gen_flat(torch, device, tdt, 0, rows * n, nan_frac(family), seed=2)[e_lo:e_hi].clone()
    a = flat.view(my_rows, my_n) if family != "group1d" else flat
    tensors = [a]
    if func in TWO_INPUT:
        tensors.append(a * a + 1)
    labels = None
    if family == "group":
        labels = torch.from_numpy(shared_labels(n, K)).to(device)
    elif family == "group1d":
        labels = gen_labels_flat(torch, device, e_lo, e_hi, K)
    lens = [shard_bounds(n, r, world, GEN_BLOCK)[1] - shard_bounds(n, r, world, GEN_BLOCK)[0] for r in range(world)] if core_sharded else None
    qdev = torch.tensor(np.atleast_1d(params["quantiles"]), dtype=torch.float64, device=device) if family == "quantile" else None
    al = None
    if family == "exp":
        al = float(np.float32(params["alpha"])) if dt == "f32" else params["alpha"]

    def group1d_on(v2, lab2):
        return D.run_group(func, v2, lab2, K, 1)

    def step_device():
        if family == "move":
            return D.run_move(func, tensors, params["window"], params["min_count"], -1)
        if family == "exp":
            if core_sharded:
                return nd.move_exp_sharded(func, *tensors, alpha=al)
            return D.run_move_exp(func, tensors, al, 0.0, -1)[0]
        if family == "fill":
            if core_sharded:
                return nd.fill_sharded(func, a, shard_lens=lens)
            return D.run_fill(func, a, n, -1)[0]
        if family == "reduce":
            axes = (0, 1) if params["axis"] is None else (params["axis"] % 2,)
            return D.run_reduce(func, a, axes)
        if family == "quantile":
            return D.run_quantile(a, qdev, (params["axis"] % 2,))
        if family == "matrix":
            return D.run_matrix(func, a, window=params["window"], min_count=params["min_count"])
        if family == "group":
            return D.run_group(func, a, labels, K, 1)
        if core_sharded:
            return nd.group_sharded(func, a.view(1, -1), labels, num_labels=K, index_offset=e_lo)
        return group1d_on(a.view(1, -1), labels)

    elements = rows * n  # the whole configuration: every rank holds 1/world of it
    my_elements = my_rows * my_n
    abytes = alg_bytes(family, func, dt, my_rows, my_n, params)

    # ---- timed region: device-resident
    warm = max(3, args.warmup)
    for _ in range(warm):
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
    use_graph = abytes < (256 << 20) and not args.no_graph and not core_sharded
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
    while rank == 0 and world == 1 and time.perf_counter() - t_extra < 1.5:
        step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        dist.barrier()
    ms_step = ms_total / args.steps
    replicas = world if replica_mode else 1
    value = replicas * elements / (ms_step * 1e-3)

    # ---- full-size parity spot checks of this run's outputs (rank 0's shard)
    parity = None
    if rank == 0 and not args.no_parity and family in ("group", "move", "exp", "fill", "group1d", "reduce"):
        try:
            from oracle import oracle

            p_out = out
            if core_sharded and family == "group1d":
                p_out = None
            parity = parity_checks(torch, nb, D, oracle, wl, tensors, labels, p_out,
                                   group1d_on if family == "group1d" else None, e_lo, ends_row=(e_hi == rows * n))
        except Exception as ex:  # noqa: BLE001
            parity = dict(error=f"{type(ex).__name__}: {str(ex)[:200]}")
    if world > 1:
        dist.barrier()

    # ---- e2e through the public numpy API: every step copies the pinned host inputs to the device, runs
    # the kernels and copies the result back to the host.  Host batch = this rank's share of the SAME
    # configuration when host RAM allows, else the largest prefix of it that fits.
    e2e = None
    if not args.no_e2e:
        try:
            import psutil

            avail = int(psutil.virtual_memory().available)
        except Exception:
            avail = 32 << 30
        per_elem = item * len(tensors) + (8 if family == "group1d" else 0)
        out_per_elem = item if family in ("move", "exp", "fill") else 0
        budget = int(avail * 0.35 / max(1, min(world, 8)))
        e_elems = min(my_elements, max(1, budget // (per_elem + out_per_elem)))
        if my_rows > 1:
            erows, en = max(1, min(my_rows, e_elems // my_n)), my_n
        else:
            erows, en = 1, max(4, (e_elems // 4) * 4)
        pinned = []
        for tsr in tensors:
            src = tsr[:erows] if my_rows > 1 else tsr.reshape(-1)[:en]
            h = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            h.copy_(src)
            pinned.append(h.numpy())
        if family == "group":
            pinned.append(shared_labels(n, K))
        elif family == "group1d":
            hl = torch.empty(en, dtype=torch.int64, pin_memory=True)
            hl.copy_(labels[:en])
            pinned.append(hl.numpy())
        # free the device-resident inputs: the public call brings its own copy
        del tensors, a, flat, out
        labels = None
        torch.cuda.empty_cache()
        e_kwargs = dict(params)
        if family == "group":
            e_kwargs["axis"] = -1
        if family == "exp" and dt == "f32":
            e_kwargs["alpha"] = np.float32(e_kwargs["alpha"])
        f_public = getattr(nb, func)
        if my_rows == 1 and family in ("exp", "fill"):
            pinned = [p.reshape(1, -1) for p in pinned]
        res = f_public(*pinned, **e_kwargs)
        h2d = sum(x.nbytes for x in pinned)
        d2h = res.nbytes
        f_public(*pinned, **e_kwargs)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e_steps = max(2, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            res = f_public(*pinned, **e_kwargs)  # H2D + kernels + D2H (returns a host array)
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t0) / e_steps
        e_elems_all = torch.tensor([float(erows * en), e_s], dtype=torch.float64, device=device)
        if world > 1:
            t_el = e_elems_all[:1].clone()
            t_s = e_elems_all[1:].clone()
            dist.all_reduce(t_el, op=dist.ReduceOp.SUM)
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
            tot_el, e_s = float(t_el.item()), float(t_s.item())
        else:
            tot_el = float(erows * en)
        e2e = dict(value=tot_el / e_s, unit="elements/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                   ms_per_step=e_s * 1e3, batch_per_rank=[erows, en], same_config=bool(erows * en == my_elements),
                   note="public numpy API; pinned host input; H2D + kernels + D2H timed; PCIe-bound by construction")
        del pinned, res
        torch.cuda.empty_cache()

    # ---- sharded forms of configs 3-5 over NCCL (N > 1)
    sharded = None
    if world > 1 and not args.no_sharded:
        try:
            torch.cuda.empty_cache()
            sharded = sharded_block(torch, dist, nd, D, device, rank, world)
        except Exception as ex:  # noqa: BLE001
            sharded = dict(error=f"{type(ex).__name__}: {str(ex)[:300]}")

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = abytes / (ms_step * 1e-3) / 1e9
        base = None
        if world == 1 and not args.no_cpu:
            try:
                base = cpu_arm(family, func, dt, rows, n, params)[0]
            except Exception as ex:  # noqa: BLE001
                base = dict(error=f"{type(ex).__name__}: {str(ex)[:200]}")
        line = dict(
            metric="elements/s", value=value, unit="elements/s", n_gpus=world, steps=args.steps,
            warmup=warm, ms_per_step=ms_step, higher_is_better=True, scaling="strong",
            vs_baseline=None, dtype=dt, data="synthetic",
            config=dict(workload=args.workload, func=func, shape=[rows, n], per_gpu_share=[my_rows, my_n], sharding=sharding,
                        nan_fraction=nan_frac(family),
                        l2="inputs larger than L2 (no flush needed)" if abytes > (256 << 20) else "input smaller than L2: launch-latency bound config, reported as is",
                        launch="cuda graph of the K steps" if use_graph else "stream", **_jsonable(params)),
            clocks=clocks,
            e2e=e2e,
            gpu_launches=int(launches),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=measured_traffic(args.workload), peak_source=peak_src,
                          note="algorithmic bytes of one GPU's share per step / CUDA-event step time (all kernels of the step); "
                               "traffic = ncu dram bytes of the dominant kernel from the committed capture of this workload (profiles/)"),
            cpu_baseline=base,
            parity=parity,
        )
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size parity spot checks")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded block (configs 3-5 over NCCL)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-in / host-out leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
