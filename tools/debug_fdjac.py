import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from golden_util import by_name, spec_from_hex
from gpu_util import gpu_fdjac, gpu_residual
from backends import OracleBackend
ora = OracleBackend()
names = sys.argv[1:] or ["goddard_stage4_singular"]
for name in names:
    spec = spec_from_hex(by_name("residual", name)["spec"])
    want = ora.fdjac(spec); got = gpu_fdjac(spec)
    dev = np.zeros_like(want)
    for seed in (1, 2, 3):
        dev = np.maximum(dev, np.abs(ora.fdjac(spec, noise_ulps=8.0, seed=seed) - want))
    colmax = np.max(np.abs(want), axis=0)
    x0 = np.array(spec["x0"]); f0 = ora.residual(spec)
    h = np.sqrt(1e-15) * np.abs(x0); h[h == 0] = np.sqrt(1e-15)
    scale = max(np.max(np.abs(x0)), np.max(np.abs(f0)), 1.0)
    quantum = 2.220446049250313e-16 * scale / h
    tol = 4 * dev + 8 * quantum[None, :] + 1e-9 * colmax[None, :]
    ratio = np.abs(got - want) / tol
    idx = np.dstack(np.unravel_index(np.argsort(-ratio.ravel())[:25], ratio.shape))[0]
    print(name, "scale", scale)
    gr = gpu_residual(spec)
    print("  residual max abs diff", np.max(np.abs(gr - f0)), "at", np.argmax(np.abs(gr - f0)))
    for i, j in idx:
        print("  row %3d col %3d got % .6e want % .6e dev %.3e colmax %.3e x_j % .3e h %.2e f_i % .3e quantum %.2e ratio %.2e" % (i, j, got[i, j], want[i, j], dev[i, j], colmax[j], x0[j], h[j], f0[i], quantum[j], ratio[i, j]))
