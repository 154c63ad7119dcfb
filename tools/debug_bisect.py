"""Bisect a solver-path difference: CPU MINPACK (oracle hybrd) driven by GPU residuals."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from golden_util import golden, by_name, spec_from_hex, unhex
from gpu_util import gpu_residual, gpu_solve, gpu_fdjac
from backends import OracleBackend
from oracle import pyminpack as pm
import scenarios as S
ora = OracleBackend()
e = [c for c in golden()["cont_param"] if c["spec"]["name"] == "di_cont_muT"][0]
spec = spec_from_hex(e["spec"]); spec["mparams"][2] = 0.02
o = ora.solve(spec); print("oracle", o["info"], o["nfev"])
g = gpu_solve(spec); print("gpu   ", g["info"][0], g["nfev"][0], np.linalg.norm(g["x"][0]-o["x"]))
h = pm.hybrd(lambda x: gpu_residual(spec, x), spec["x0"], xtol=spec["xtol"]); print("cpu-minpack on gpu residuals", h["info"], h["nfev"], np.linalg.norm(h["x"]-o["x"]))
J1 = gpu_fdjac(spec); J2 = ora.fdjac(spec)
np.set_printoptions(linewidth=250, precision=3)
print("max |J1-J2|", np.max(np.abs(J1-J2)))
print((J1 != 0).astype(int)); print((J2 != 0).astype(int))
d = np.abs(J1-J2); print(d)
