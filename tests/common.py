"""Shared helpers for the test-suite: golden fixture loading."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    with open(os.path.join(GOLDEN_DIR, "INDEX.txt")) as f:
        return [l.strip() for l in f if l.strip()]


def load_golden(name):
    """-> dict(X, D, Y, dY, dX, kwargs, is_list) exactly as the reference was called."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    kw = dict(meta["kwargs"])
    if kw.get("crop") is not None:
        kw["crop"] = tuple(slice(a, b) for a, b in kw["crop"])
    if kw.get("affine") is not None:
        kw["affine"] = np.array(kw["affine"])
    if "axis" in kw:
        if isinstance(kw["axis"], dict):
            kw["axis"] = tuple(kw["axis"]["tuple"])
        elif isinstance(kw["axis"], list):
            kw["axis"] = [tuple(a) for a in kw["axis"]]
    n = meta["n"]
    g = {"kwargs": kw, "is_list": meta["is_list"], "grad": meta["grad"],
         "D": z["displacement"],
         "X": [z["x%d" % i] for i in range(n)], "Y": [z["y%d" % i] for i in range(n)]}
    if meta["grad"]:
        g["dY"] = [z["dy%d" % i] for i in range(n)]
        g["dX"] = [z["dx%d" % i] for i in range(n)]
    return g


def call_args(g, what="X"):
    """Positional input in the form the reference was called with (list or single array)."""
    v = g[what]
    return list(v) if g["is_list"] else v[0]


def as_list(v):
    return v if isinstance(v, (list, tuple)) else [v]


def grad_kwargs(g):
    kw = dict(g["kwargs"])
    kw["X_shape"] = [x.shape for x in g["X"]] if g["is_list"] else g["X"][0].shape
    return kw
