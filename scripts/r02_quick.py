"""Round-2 A/B timing harness (GPU box): builds each BASELINE-config input ONCE and times the device
entry points under several environment settings (the C library reads its tuning variables per call).
CUDA events, inputs larger than L2.  Prints one JSON line per (workload, setting).

    python scripts/r02_quick.py cfg2 cfg3 cfg4 cfg5 cfg1s [--steps 10]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numbagg_b200 import decorators as D  # noqa: E402

PEAK = 6447.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def gen(shape, dt, nan_frac, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.empty(shape, dtype=dt, device=dev)
    flat = a.view(-1)
    for s in range(0, flat.numel(), 1 << 28):
        seg = flat[s:s + (1 << 28)]
        seg.uniform_(0, 1, generator=g)
        seg[seg <= nan_frac] = float("nan")
    return a


def timeit(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(name, fn, abytes, settings, steps):
    for env in settings:
        saved = {k: os.environ.get(k) for k in env}
        for k, v in env.items():
            os.environ[k] = str(v)
        try:
            ms = timeit(fn, steps)
            gbs = abytes / ms / 1e6
            print(json.dumps(dict(workload=name, env=env, ms=round(ms, 4), gbs=round(gbs, 1), frac=round(gbs / PEAK, 4))), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(json.dumps(dict(workload=name, env=env, error=str(ex)[:300])), flush=True)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    steps = 10
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    if "persist" in args:
        # trigger the persisting-L2 set-aside (per-element-label path) BEFORE timing unrelated kernels
        v = gen((1, 50_000_000), torch.float64, 0.1)
        lab = torch.randint(0, 1_000_000, (50_000_000,), device=dev, dtype=torch.int64)
        D.run_group("group_nansum", v, lab, 1_000_000, 1)
        torch.cuda.synchronize()
        del v, lab
    if "cfg2" in args:
        rows, n, K = 10_000, 1_000_000, 1000
        a = gen((rows, n), torch.float32, 0.1)
        labels = torch.from_numpy(np.random.RandomState(0).randint(0, K, size=n)).to(dev)
        ab = rows * n * 4 + n * 8 + rows * K * 4
        rb2 = [dict(NBG_RB2_OFF=1), dict(NBG_RB2_S=3)]
        if "sweep" in args:
            for NW in (16, 31):
                for S in (2, 3):
                    rb2.append(dict(NBG_RB2_NW=NW, NBG_RB2_S=S))
            rb2.append(dict(NBG_RB2_NW=31, NBG_RB2_S=2, NBG_RB2_NSEG=2))
            rb2.append(dict(NBG_RB2_NW=31, NBG_RB2_S=2, NBG_RB2_PD=4))
            rb2.append(dict(NBG_RB2_NW=31, NBG_RB2_S=4))
        for f, sets in (("group_nansum", rb2), ("group_nancount", rb2[:2]), ("group_nansum_of_squares", rb2[:2]),
                        ("group_nanmean", [{}]), ("group_nanstd", [{}]), ("group_nanmax", [{}]), ("group_nanargmax", [{}])):
            run("cfg2_" + f, lambda: D.run_group(f, a, labels, K, 1), ab, sets, steps)
        # 1/8 of the rows: the per-GPU share of the strong-scaled config at N=8
        a8 = a[:1250]
        ab8 = 1250 * n * 4 + n * 8 + 1250 * K * 4
        run("cfg2/8_group_nansum", lambda: D.run_group("group_nansum", a8, labels, K, 1), ab8, [dict(NBG_RB2_S=3)], steps)
        run("cfg2/8_group_nansum", lambda: D.run_group("group_nansum", a8, labels, K, 1), ab8, [dict(NBG_RB2_OFF=1)], steps)
        del a, a8
        torch.cuda.empty_cache()
    pf = [dict(NBG_PREFETCH_TILES=0), {}, dict(NBG_PREFETCH_TILES=2000)]
    if "cfg3" in args:
        n = 1_000_000_000
        a = gen((1, n), torch.float64, 0.3)
        ab = n * 16
        run("cfg3_move_exp_nanmean", lambda: D.run_move_exp("move_exp_nanmean", [a], 0.1, 0.0, -1)[0], ab, pf, steps)
        run("cfg3_ffill", lambda: D.run_fill("ffill", a, n, -1)[0], ab, pf, steps)
        run("cfg3_bfill", lambda: D.run_fill("bfill", a, n, -1)[0], ab, pf, steps)
        for f in ("move_exp_nansum", "move_exp_nancount", "move_exp_nanvar", "move_exp_nanstd"):
            run("cfg3_" + f, lambda: D.run_move_exp(f, [a], 0.1, 0.0, -1)[0], ab, [{}], steps)
        del a
        torch.cuda.empty_cache()
    if "pfsweep" in args:
        n = 1_000_000_000
        a = gen((1, n), torch.float64, 0.3)
        sw = [dict(NBG_PREFETCH_TILES=v) for v in (0, 150, 300, 600, 888, 1200)]
        run("cfg3_ffill", lambda: D.run_fill("ffill", a, n, -1)[0], n * 16, sw, steps)
        run("cfg3_move_exp_nanmean", lambda: D.run_move_exp("move_exp_nanmean", [a], 0.1, 0.0, -1)[0], n * 16, sw, steps)
        del a
        torch.cuda.empty_cache()
        b = gen((2000, 100_000), torch.float64, 0.1)
        run("cfg1s_move_mean", lambda: D.run_move("move_mean", [b], 20, 1, -1), 2000 * 100_000 * 16, [dict(NBG_PREFETCH_TILES=v) for v in (0, 100, 200, 296, 600)], steps)
        del b
    if "cfg4" in args:
        rows, n = 1000, 1_000_000
        a = gen((rows, n), torch.float32, 0.1)
        b = a * a + 1
        for f, nin in (("move_std", 1), ("move_var", 1), ("move_mean", 1), ("move_cov", 2), ("move_corr", 2)):
            ts = [a, b][:nin]
            run("cfg4_" + f, lambda: D.run_move(f, ts, 1000, 500, -1), rows * n * 4 * (nin + 1), pf[:2], steps)
        del a, b
        torch.cuda.empty_cache()
    if "cfg4g" in args:
        rows, n = 1000, 1_000_000
        a = gen((rows, n), torch.float32, 0.1)
        for f in ("move_std", "move_var", "move_mean"):
            run("cfg4_" + f, lambda: D.run_move(f, [a], 1000, 500, -1), rows * n * 8,
                [dict(NBG_PFX="all", NBG_PFX_GEOM=0), dict(NBG_PFX="all", NBG_PFX_GEOM=1), dict(NBG_PFX="off", NBG_PFX_GEOM=0)], steps)
        del a
        torch.cuda.empty_cache()
    if "mat" in args:
        no, nv = 200_000, 32
        for dt in (torch.float64, torch.float32):
            a = gen((no, nv), dt, 0.1)
            ob = no * nv * nv * a.element_size()
            tag = "f64" if dt == torch.float64 else "f32"
            al = torch.full((no,), 0.05, dtype=dt, device=dev)
            sw = [dict(NBG_MAT_NOSEG=1), dict()] if dt == torch.float64 else [dict()]
            run(f"mat_move_cov_{tag}", lambda: D.run_matrix("move_covmatrix", a, window=100, min_count=10), ob, sw, 3)
            run(f"mat_move_corr_{tag}", lambda: D.run_matrix("move_corrmatrix", a, window=100, min_count=10), ob, [dict()], 3)
            run(f"mat_exp_cov_{tag}", lambda: D.run_matrix("move_exp_nancovmatrix", a, alpha=al), ob, sw, 3)
            run(f"mat_exp_corr_{tag}", lambda: D.run_matrix("move_exp_nancorrmatrix", a, alpha=al), ob, [dict()], 3)
            at = a.T.contiguous()
            run(f"mat_static_cov_{tag}", lambda: D.run_matrix("nancovmatrix", at), no * nv * a.element_size(), sw, 3)
            run(f"mat_static_corr_{tag}", lambda: D.run_matrix("nancorrmatrix", at), no * nv * a.element_size(), [dict()], 3)
            del at
            del a
            torch.cuda.empty_cache()
    if "cfg3x" in args:
        n = 1_000_000_000
        a = gen((1, n), torch.float64, 0.3)
        for f in ("move_exp_nanmean", "move_exp_nanvar", "move_exp_nanstd"):
            run("cfg3_" + f, lambda: D.run_move_exp(f, [a], 0.1, 0.0, -1)[0], n * 16, [{}], steps)
        run("cfg3_move_exp_nanvar_gate", lambda: D.run_move_exp("move_exp_nanvar", [a], 0.1, 0.5, -1)[0], n * 16, [{}], steps)
        del a
        torch.cuda.empty_cache()
        n = 500_000_000
        a = gen((1, n), torch.float64, 0.3)
        b = a * a + 1
        for f in ("move_exp_nancov", "move_exp_nancorr"):
            run("cfg3h_" + f, lambda: D.run_move_exp(f, [a, b], 0.1, 0.0, -1)[0], n * 24, [{}], steps)
        del a, b
        torch.cuda.empty_cache()
    if "cfg5q" in args:
        n, K = 2_000_000_000, 10_000_000
        a = gen((n,), torch.float64, 0.1)
        g = torch.Generator(device=dev).manual_seed(1)
        labels = torch.randint(0, K, (n,), generator=g, device=dev, dtype=torch.int64)
        ab = n * 16 + K * 8
        v2 = a.view(1, -1)
        for f in ("group_nansum", "group_nanmean", "group_nanvar", "group_nanargmax", "group_nanfirst", "group_nanmax"):
            run("cfg5_" + f, lambda: D.run_group(f, v2, labels, K, 1), ab, [{}], max(3, steps // 2))
        del a, labels, v2
        torch.cuda.empty_cache()
    if "cfg2nc" in args:
        rows, n, K = 10_000, 1_000_000, 1000
        a = gen((rows, n), torch.float32, 0.1)
        labels = torch.from_numpy(np.random.RandomState(0).randint(0, K, size=n)).to(dev)
        ab = rows * n * 4 + n * 8 + rows * K * 4
        run("cfg2_group_nansum", lambda: D.run_group("group_nansum", a, labels, K, 1), ab, [dict(NBG_RB2_NC=1), dict(NBG_RB2_NC=2), dict(NBG_RB2_NC=4)], steps)
        run("cfg2_group_nanmean", lambda: D.run_group("group_nanmean", a, labels, K, 1), ab, [dict(NBG_RB2_OFF=1), dict(NBG_RB2_NC=2), dict(NBG_RB2_NC=3), dict(NBG_RB2_NC=4)], steps)
        run("cfg2_group_nanstd", lambda: D.run_group("group_nanstd", a, labels, K, 1), ab, [dict(NBG_RB2_OFF=1), dict(NBG_RB2_NC=3), dict(NBG_RB2_NC=4), dict(NBG_RB2_NC=6)], steps)
        del a
        torch.cuda.empty_cache()
    if "cfg2nw" in args:
        rows, n, K = 10_000, 1_000_000, 1000
        a = gen((rows, n), torch.float32, 0.1)
        labels = torch.from_numpy(np.random.RandomState(0).randint(0, K, size=n)).to(dev)
        ab = rows * n * 4 + n * 8 + rows * K * 4
        sets = [dict(NBG_RB2_NW=w) for w in (16, 8, 12, 24)] + [dict(NBG_RB2_NW=12, NBG_RB2_PD=2), dict(NBG_RB2_NW=16, NBG_RB2_PD=2), dict(NBG_RB2_NW=8, NBG_RB2_S=3), dict(NBG_RB2_NW=12, NBG_RB2_S=3)]
        run("cfg2_group_nansum", lambda: D.run_group("group_nansum", a, labels, K, 1), ab, sets, steps)
        del a
        torch.cuda.empty_cache()
    if "cfg5p" in args:
        n, K = 2_000_000_000, 10_000_000
        a = gen((n,), torch.float64, 0.1)
        g = torch.Generator(device=dev).manual_seed(1)
        labels = torch.randint(0, K, (n,), generator=g, device=dev, dtype=torch.int64)
        ab = n * 16 + K * 8
        v2 = a.view(1, -1)
        for f in ("group_nanvar", "group_nanmean", "group_nanstd"):
            run("cfg5_" + f, lambda: D.run_group(f, v2, labels, K, 1), ab, [dict(NBG_GROUP_PARTITION=0), dict(NBG_GROUP_PARTITION=1)], 3)
        del a, labels, v2
        torch.cuda.empty_cache()
    if "cfg1s" in args:
        rows, n = 2000, 100_000
        a = gen((rows, n), torch.float64, 0.1)
        b = a * a + 1
        for f, nin in (("move_mean", 1), ("move_std", 1), ("move_corr", 2)):
            ts = [a, b][:nin]
            run("cfg1s_" + f, lambda: D.run_move(f, ts, 20, 1, -1), rows * n * 8 * (nin + 1), pf[:2], steps)
        del a, b
        torch.cuda.empty_cache()
    if "cfg5" in args:
        n, K = 2_000_000_000, 10_000_000
        a = gen((n,), torch.float64, 0.1)
        g = torch.Generator(device=dev).manual_seed(1)
        labels = torch.randint(0, K, (n,), generator=g, device=dev, dtype=torch.int64)
        ab = n * 16 + K * 8
        v2 = a.view(1, -1)
        for f in ("group_nansum", "group_nanvar", "group_nanargmax", "group_nanfirst"):
            run("cfg5_" + f, lambda: D.run_group(f, v2, labels, K, 1), ab, [dict(NBG_L2_PERSIST=0, NBG_L2_HINT=0), dict(NBG_L2_PERSIST=0, NBG_L2_HINT=1), dict(NBG_L2_PERSIST=1, NBG_L2_HINT=0)], max(3, steps // 3))
        run("cfg5_group_nanmean", lambda: D.run_group("group_nanmean", v2, labels, K, 1), ab, [dict(NBG_L2_HINT=1)], 3)
        lab32 = labels.to(torch.int32)
        run("cfg5_group_nansum_i32labels", lambda: D.run_group("group_nansum", v2, lab32, K, 1), n * 12 + K * 8, [dict(NBG_L2_HINT=1)], 3)
        del lab32
        torch.cuda.empty_cache()
        slab = labels.sort().values
        run("cfg5_group_nansum_sorted", lambda: D.run_group("group_nansum", v2, slab, K, 1), ab, [dict(NBG_L2_HINT=1)], 3)


if __name__ == "__main__":
    main()
