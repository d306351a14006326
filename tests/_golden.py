"""Loader for tests/golden/*.npz (written by oracle/gen_golden.py from the reference)."""

import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star parity classes
EXACT_FUNCS = {
    "ffill", "bfill", "group_nancount", "group_nanargmax", "group_nanargmin", "group_nanfirst",
    "group_nanlast", "group_nanmin", "group_nanmax", "group_nanany", "group_nanall",
}


class Case:
    def __init__(self, suite, idx, entry, data):
        self.suite, self.idx, self.entry = suite, idx, entry
        self.func = entry["func"]
        self.args = [data[k] for k in entry["args"]]
        if "layout" in entry:  # re-create the strides the reference saw (npz stores C order)
            perm = entry["layout"]
            self.args[0] = np.ascontiguousarray(self.args[0].transpose(perm)).transpose(np.argsort(perm))
        self.kwargs = {}
        for k, v in entry["kwargs"].items():
            if isinstance(v, dict) and "np_scalar" in v:
                v = np.dtype(v["np_scalar"]).type(v["value"])
            elif isinstance(v, dict) and "tuple" in v:
                v = tuple(v["tuple"])
            self.kwargs[k] = v
        for k, key in entry["array_kwargs"].items():
            self.kwargs[k] = data[key]
        self.expected = data[entry["out"]]

    @property
    def id(self):
        kw = ",".join(
            f"{k}={'arr' if isinstance(v, np.ndarray) else v}" for k, v in self.kwargs.items()
        )
        a0 = self.args[0]
        return f"{self.idx}-{self.func}-{a0.dtype}{list(a0.shape)}-{kw}"


def load_suite(name):
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    with np.load(path) as z:
        data = {k: z[k] for k in z.files}
    manifest = json.loads(bytes(data.pop("manifest")).decode())
    return [Case(name, i, e, data) for i, e in enumerate(manifest)]


def all_cases(*names):
    out = []
    for n in names or ("moving", "moving_exp", "fill", "grouped"):
        out.extend(load_suite(n))
    return out
