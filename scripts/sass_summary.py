"""profiles/rNN_sass_summary.txt: what the built libnbg_b200.so contains (cuobjdump -sass): per kernel
family, the counts of the SASS mnemonics that prove the sm_100a-native paths -- UBLKCP (1-D TMA bulk
copies, both directions), UBLKPF (bulk L2 prefetch), SYNCS.* (mbarrier), RED/ATOM (global reductions),
ATOMS (shared atomics), BAR, LDS/STS, F2F / MUFU (conversion / special-function pipe), DADD/DMUL/DFMA.
usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "numbagg_b200", "libnbg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fam = collections.OrderedDict()
cur = None
archs = collections.Counter()
WANT = ["UBLKCP", "UBLKPF", "SYNCS", "UTMALDG", "UTCHMMA", "RED", "ATOMG", "ATOM", "ATOMS", "BAR", "LDS", "STS", "LDG", "STG",
        "F2F", "MUFU", "DADD", "DMUL", "DFMA", "SHFL", "CCTL", "NANOSLEEP"]
nfunc = 0
for line in out.splitlines():
    m = re.match(r"\s*arch = (sm_\w+)", line)
    if m:
        archs[m.group(1)] += 1
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        nfunc += 1
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        base = re.sub(r"<.*", "", name.replace("void ", "")).replace("nbg::", "")
        cur = fam.setdefault(base, dict(variants=0, insts=0, c=collections.Counter()))
        cur["variants"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["insts"] += 1
        for w in WANT:
            if op == w or op.startswith(w + ".") or (w == "RED" and op == "REDG") or (w == "ATOM" and op == "ATOM"):
                cur["c"][w] += 1
print(f"libnbg_b200.so: {nfunc} kernels (template instances), architectures: {dict(archs)}")
print("counts are summed over all template instances of a kernel family\n")
hdr = ["kernel family", "variants", "SASS insts"] + WANT
print(" | ".join(hdr))
for k, v in fam.items():
    print(" | ".join([k, str(v["variants"]), str(v["insts"])] + [str(v["c"].get(w, 0)) for w in WANT]))
tot = collections.Counter()
for v in fam.values():
    tot.update(v["c"])
print("\nTOTAL " + ", ".join(f"{w}={tot.get(w, 0)}" for w in WANT))
print("No UTMALDG / UTCHMMA by design: every tile is a 1-D row span (bulk copies need no tensor map) and no kernel contracts (no tensor cores on this path).")
