import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from golden_util import by_name, spec_from_hex
from gpu_util import gpu_fdjac, gpu_residual
from backends import OracleBackend
ora = OracleBackend()
for name in ["vtol_wp1", "goddard_stage1", "interceptor_init"]:
    spec = spec_from_hex(by_name("residual", name)["spec"])
    want = ora.fdjac(spec); got = gpu_fdjac(spec)
    bad = np.argwhere((want == 0) & (got != 0))
    print(name, "P", want.shape[0], "bad entries", len(bad))
    x0 = np.array(spec["x0"])
    for i, j in bad[:12]:
        print("   row", i, "col", j, "got", got[i, j], "x_j", x0[j], "h", 3.1622776601683795e-08 * abs(x0[j]))
    f0 = gpu_residual(spec); fo = ora.residual(spec)
    print("   residual max diff", np.max(np.abs(f0 - fo)))
