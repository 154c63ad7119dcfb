"""Load tests/golden/golden.json (float.hex() strings -> float64)."""
import json
import os

import numpy as np

_G = None


def unhex(v):
    if v is None:
        return None
    if isinstance(v, str):
        return float.fromhex(v)
    return np.array([unhex(x) for x in v], dtype=np.float64)


def spec_from_hex(s):
    out = dict(s)
    for k in ("mparams", "time", "x0"):
        out[k] = [float(v) for v in unhex(s[k])]
    out["Xb"] = [[float(v) for v in row] for row in unhex(s["Xb"])]
    out["xtol"] = float.fromhex(s["xtol"])
    return out


def golden():
    global _G
    if _G is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")) as f:
            _G = json.load(f)
    return _G


def by_name(kind, name):
    for e in golden()[kind]:
        if e["spec"]["name"] == name:
            return e
    raise KeyError(name)
