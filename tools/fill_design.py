"""Fill the [PLACEHOLDER] numbers of DESIGN.md from the final bench line (profiles/<tag>_bench.json) and, when given,
the strong-scaling lines.  usage: python tools/fill_design.py r2_final [strong_1.json strong_2.json ...]"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
b = json.load(open(os.path.join(ROOT, "profiles", tag + "_bench.json")))
k = b["roofline"]["kernels"]
def kk(prefix):
    for name, v in k.items():
        if name.startswith(prefix):
            return v
    return None
qp, ch, jac, asm, integ = kk("hybrd_qpass"), kk("hybrd_chain"), kk("hybrd_jac"), kk("assemble"), kk("integrate")
s2 = b.get("stage2") or {}
warm, kd = s2.get("warm_start", {}), s2.get("continuation_KD", {})
pct = lambda x: "%.1f %%" % (100 * x)
kilo = lambda x: "%.1f k" % (x / 1e3)
rep = {
    "VALUE": kilo(b["value"]), "E2E": kilo(b["e2e"]["value"]), "STEP_MS": "%.0f" % b["roofline"]["profiled_step_ms"],
    "QPASS_MS": "%.0f" % qp["ms"], "QPASS_GBS": "%.0f" % qp["achieved"], "QPASS_FRAC": pct(qp["frac"]),
    "CHAIN_MS": "%.0f" % ch["ms"], "CHAIN_FRAC": pct(ch["frac"]),
    "JAC_MS": "%.0f" % jac["ms"], "JAC_FRAC": pct(jac["frac"]), "JAC_US": "%.2f" % (jac["ms"] * 1e3 / jac["units"]),
    "ASM_MS": "%.0f" % asm["ms"], "ASM_FRAC": pct(asm["frac"]), "INT_MS": "%.0f" % integ["ms"],
    "WHOLE_FRAC": pct(b["whole_solve"]["fp64_frac"]), "RK4ONLY_FRAC": pct(b["whole_solve"]["rk4_only_frac"]),
    "CPU": "%.0f" % b["cpu_baseline"]["value"],
    "RK4_FRAC": pct(b["rk4_kernel"]["frac"]), "RK4_GSTEPS": "%.1f" % (b["rk4_kernel"]["rk4_steps_per_s"] / 1e9),
    "P_INFO": pct(b["parity"]["identical_info"]), "P_BOTH": pct(b["parity"]["identical_info_nfev"]),
    "P_Z": "%.2f" % b["parity"]["converged_rate_z"],
    "WARM_VALUE": kilo(warm.get("value", 0)), "WARM_E2E": kilo(warm.get("e2e", {}).get("value", 0)),
    "WARM_BOTH": pct(warm.get("parity", {}).get("identical_info_nfev", 0)), "WARM_X": pct(warm.get("parity", {}).get("x_within_xtol", 0) or 0),
    "KD_VALUE": kilo(kd.get("value", 0)), "KD_E2E": kilo(kd.get("e2e", {}).get("value", 0)),
    "KD_INFO": pct(kd.get("parity", {}).get("identical_info", 0)),
}
if len(sys.argv) > 2:
    lines = [json.load(open(f)) for f in sys.argv[2:]]
    lines.sort(key=lambda l: l["n_gpus"])
    base = lines[0]["value"] / lines[0]["n_gpus"]
    rep["STRONG"] = ", ".join("N=%d %s solves/s (%.3f)" % (l["n_gpus"], kilo(l["value"]), l["value"] / (base * l["n_gpus"])) for l in lines)
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
for key, val in rep.items():
    s = s.replace("[" + key + "]", val)
open(p, "w").write(s)
left = sorted(set(re.findall(r"\[[A-Z0-9_]+\]", s)))
print("filled; placeholders left:", left)
