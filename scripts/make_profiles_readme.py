"""profiles/README.md from profiles/r02_bench_all.jsonl (scripts/r02_profiles.sh) and round 1's table.
usage: python scripts/make_profiles_readme.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    rows = {}
    try:
        for line in open(os.path.join(P, name)):
            try:
                d = json.loads(line)
            except Exception:
                continue
            if "config" in d:
                rows[d["config"]["workload"]] = d
    except FileNotFoundError:
        pass
    return rows


def fmt(x):
    return f"{x:.1f}" if x >= 10 else f"{x:.3g}"


r1, r2 = load("r01_bench_all.jsonl"), load("r02_bench_all.jsonl")
try:
    traffic = json.load(open(os.path.join(P, "r02_traffic.json")))
except Exception:
    traffic = {}
out = ["# profiles — round 2", "",
       "All numbers from one B200 (`gpurun`), device-resident inputs, CUDA events on the launching stream, >= 3",
       "warm-up steps, inputs larger than L2 (except cfg1).  `roofline` = algorithmic bytes / step time / measured copy",
       "bandwidth (6447.8 GB/s, MEASURED_PEAKS.json); 0.87 of it = north_star's 70 % of 8 TB/s.  `r01` = the same",
       "column one round ago.  `traffic` = ncu DRAM bytes of the dominant kernel / algorithmic bytes (this round's",
       "`--set full` captures, `r02_traffic.json`).  `e2e` = public numpy API with pinned host buffers (H2D + kernels +",
       "D2H inside the timed region; PCIe-bound by construction).  `cpu` = numbagg's own Numba path (`kind:",
       "reference`, oracle/_ref) on the box's host cores, 1-D inputs on one core like the reference's gufunc.  `parity`",
       "= the run's own output spot-checked against the oracle at full size.  The `red_*` rows only READ, so they can",
       "exceed the COPY bandwidth used as `peak`.  cfg1 is launch-bound (K steps replayed from one CUDA graph).", "",
       "| workload | shape | Gel/s | ms/step | roofline | r01 | traffic | e2e Gel/s | cpu Gel/s (kind, cores) | parity |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---|"]
for wl, d in r2.items():
    c = d["config"]
    e = (d.get("e2e") or {}).get("value")
    cb = d.get("cpu_baseline") or {}
    par = d.get("parity") or {}
    ok = all(v.get("ok", False) for v in par.values() if isinstance(v, dict)) if par else None
    tr = traffic.get(wl, {}).get("dram_bytes_per_launch")
    alg = d["roofline"]["achieved"] * 1e9 * d["ms_per_step"] / 1e3
    old = r1.get(wl, {}).get("roofline", {}).get("frac")
    out.append(f"| {wl} ({c['func']}, {d['dtype']}) | {c['shape'][0]}x{c['shape'][1]}{' axis=' + str(c['axis']) if 'axis' in c else ''} | "
               f"{fmt(d['value'] / 1e9)} | {d['ms_per_step']:.3f} | {d['roofline']['frac']:.3g} | {'' if old is None else format(old, '.3g')} | "
               f"{'' if not tr else format(tr / alg, '.2f')} | {'' if not e else fmt(e / 1e9)} | "
               f"{'' if not cb else fmt(cb['value'] / 1e9) + ' (' + str(cb.get('kind')) + ', ' + str(cb.get('cores')) + ')'} | "
               f"{'' if ok is None else ('ok' if ok else 'FAIL')} |")
out += ["", "Multi-GPU (strong scaling of the fixed configs + sharded forms over NCCL with in-run parity): DESIGN.md §5,",
        "`r02_bench_n2.json`, `r02_bench_n4.json`, `r02_bench_n8*.json`.", "", "Files:", "",
        "* `r02_bench_all.jsonl` — the raw bench.py JSON lines behind the table (`scripts/r02_profiles.sh`).",
        "* `r02_launches_default_bench*.csv` — ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of the default bench command: the row-bins kernel is 98 % of a step, plan / histogram / init / finalize the rest.",
        "* `r02_ncu_*.txt` — per-kernel summaries of `ncu --set full` captures (scripts/ncu_summary.py): duration, DRAM bytes, pipe utilisation, stall reasons, top stalled SASS instructions; `r02_traffic.json` is derived from them (scripts/collect_r02_profiles.py).",
        "* `r02_sass_summary.txt` — `cuobjdump -sass` census per kernel family (UBLKCP / UBLKPF = TMA bulk copy / L2 prefetch, SYNCS = mbarrier, RED / ATOMG / ATOMS, F2F, MUFU, DADD / DMUL / DFMA, ...; scripts/sass_summary.py).  No UTMALDG / UTCHMMA: 1-D row tiles, no contraction anywhere on this path.",
        "* `r02_parity_observed.json` — worst observed error per (function, dtype) over everything the GPU test session compared in the tolerance class, next to the bound.",
        "* `r02_prefetch_sweep.jsonl` — L2 prefetch distance sweep of the one-tile-per-CTA kernels.",
        "* `r02_launches_quant_median_long.csv` — launch list of one long-row `nanquantile` call (two histogram passes, compaction, candidate sort; the skipped tail passes take 4 µs each).",
        "* `r02_pytest_gpu.log` — the `-m gpu` suite on the box.",
        "* `r01_*` — round 1's evidence, kept for comparison.",
        "* `experiments/` — raw A/B timings behind statements in DESIGN.md: config-4 prefix geometry (`r02_exp18_cfg4g`), matrix segments (`r02_exp20_mat`, `r02_final2_mat`), exp read-outs (`r02_exp22_cfg3`), rb2 label split and warp sweep (`r02_exp24_cfg2`, `r02_exp29_cfg2`, `r02_exp7_cfg2`), config-5 L2 policies and the partition experiment (`r02_exp3_cfg5`, `r02_exp32_cfg5`, `r02_exp33_part_raw.csv`), sharded call breakdown (`r02_exp16_breakdown.txt`).", ""]
open(os.path.join(P, "README.md"), "w").write("\n".join(out))
print("\n".join(out[:60]))
