"""TEST / BENCH INFRASTRUCTURE -- recipe that makes the UNMODIFIED reference importable on the GPU box.

    python -m oracle.ref_install          (also run by __graft_entry__.build())

numbagg is pure Python + Numba (no build step), so "installing" it is a copy of its package
directory from the read-only mount (/root/reference/numbagg, minus its test suite) into
`oracle/_ref/numbagg`.  `oracle/_ref/` is git-ignored (no reference source ever enters the
history) but NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot, where
`bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time it (numba's parallel target
on the box's host cores).  Nothing in `numbagg_b200/` may import it (tests/test_abi_cpu.py greps).

On a machine without /root/reference this is a no-op and whatever copy is already present is used;
if there is none, bench.py falls back to the C port (oracle/nbg_oracle.c) and says `kind: "port"`.
"""

from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/numbagg"
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "numbagg")


def install(verbose: bool = False) -> str | None:
    """Copy the reference package when its mount is present.  Returns the directory that should
    be put on sys.path to import it (or None when no copy exists)."""
    if os.path.isdir(SRC):
        os.makedirs(DST_ROOT, exist_ok=True)
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("test", "__pycache__", "*.pyc"))
        if verbose:
            print(f"reference copied: {SRC} -> {DST}")
    return DST_ROOT if os.path.isfile(os.path.join(DST, "__init__.py")) else None


def available() -> str | None:
    return DST_ROOT if os.path.isfile(os.path.join(DST, "__init__.py")) else None


def import_reference(num_threads: int | None = None):
    """Import the reference from oracle/_ref with numba's thread pool sized BEFORE numba loads.
    torchrun exports OMP_NUM_THREADS=1 to every rank; numba's omp layer and NUMBA_NUM_THREADS would
    silently inherit it, so both are set explicitly here.  Returns (module, info dict)."""
    root = available()
    if root is None:
        raise ImportError("oracle/_ref/numbagg is missing (run `python -m oracle.ref_install` where /root/reference exists)")
    n = int(num_threads or os.cpu_count() or 1)
    already = "numba" in sys.modules
    if not already:
        os.environ["NUMBA_NUM_THREADS"] = str(n)
        os.environ["OMP_NUM_THREADS"] = str(n)
    if root not in sys.path:
        sys.path.insert(0, root)
    import numba
    import numbagg

    if already and numba.config.NUMBA_NUM_THREADS >= n:
        numba.set_num_threads(n)
    info = dict(
        numbagg_file=os.path.relpath(numbagg.__file__, os.path.dirname(HERE)),
        numba=numba.__version__,
        num_threads=int(numba.get_num_threads()),
        cpu_count=os.cpu_count(),
    )
    return numbagg, info


if __name__ == "__main__":
    print(install(verbose=True))
