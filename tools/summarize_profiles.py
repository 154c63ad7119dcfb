"""Copy a profiling pass (tools/profile_final.sh) from gpurun_out/<tag>/ into profiles/ and write the
launch-list summary next to it.  usage: python tools/summarize_profiles.py r1_final"""
import collections, csv, gzip, io, json, os, re, shutil, sys
tag = sys.argv[1]
O, P = os.path.join("gpurun_out", tag), "profiles"
d = json.load(open(O + "/bench.json")); r = json.load(open(O + "/bench_reference.json"))
lines = open(O + "/launches.csv").read().splitlines()
s = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
agg = collections.defaultdict(lambda: [0, 0.0])
for x in csv.DictReader(io.StringIO("\n".join(lines[s:]))):
    if x["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    k = re.sub(r"\(.*", "", x["Kernel Name"])[:60]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(a[1] for a in agg.values())
out = ["# %s: ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 python bench.py --batch 20000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1" % tag,
       "# (cold-cache, serialised launch times: compare SHARES, not absolutes; launches 200..599 = rounds ~40..120 of the warm-up solve)"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-62s launches=%4d total=%10.1f us share=%.3f avg=%8.1f us" % (k, a[0], a[1], a[1] / tot, a[1] / a[0]))
kk = d["roofline"]["kernels"]; tk = sum(v["ms"] for v in kk.values())
out.append("# live CUDA-event shares over the timed steps of bench.py (batch 1e5, no profiler): " +
           ", ".join("%s %.3f" % (k.split("<")[0], v["ms"] / tk) for k, v in kk.items()))
# DRAM traffic of the captured hybrd_res launch
rows = list(csv.reader(open(O + "/ncu_res_raw.csv"))); hdr, units, vals = rows[0], rows[1], rows[2]
g = lambda n: float(vals[hdr.index(n)].replace(",", ""))
src = gzip.open(O + "/ncu_res_source.csv.gz", "rt").read().splitlines()
s2 = next(i for i, l in enumerate(src) if l.startswith('"Address"'))
its = max(int(x["Instructions Executed"]) for x in csv.DictReader(io.StringIO("\n".join(src[s2:]))) if "REDG.E.ADD.64" in x["Source"])
out.append("# hybrd_res_kernel capture (--set full): duration %.1f %s, dram read %.1f + write %.1f MB, %d Broyden iterations in the launch"
           " -> %.1f KB of DRAM traffic per iteration (algorithmic 182.0 KB)" % (g("gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")],
           g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), its, (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e3 / its))
out.append("# bench: %.0f solves/s (e2e %.0f), reference arm %.1f solves/s on %d host cores, RK4 kernel %.1f%% of the measured FP64 peak, dominant kernel %s at %.1f%% of %s" % (
    d["value"], d["e2e"]["value"], r["value"], r["cpu_baseline"]["cores"], 100 * d["rk4_kernel"]["frac"], d["roofline"]["kernel"],
    100 * d["roofline"]["frac"], d["roofline"]["bound"]))
open(os.path.join(P, tag + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
for f in ("bench.json", "bench_reference.json", "launches.csv", "probe_traj.log"):
    shutil.copy(os.path.join(O, f), os.path.join(P, tag + "_" + f))
for k in ("res", "jac", "traj"):
    for suf in ("_raw.csv", "_details.txt", "_source.csv.gz"):
        shutil.copy(os.path.join(O, "ncu_" + k + suf), os.path.join(P, tag + "_ncu_" + k + suf))
