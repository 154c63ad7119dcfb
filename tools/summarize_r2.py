"""Copy a profiling pass (tools/profile_r2.sh) from gpurun_out/<tag>/ into profiles/, write the launch-list summary and
profiles/r2_traffic.json: DRAM bytes per unit of work of the solver kernels, from the `ncu --set full` captures, tied
to the sources they were captured from (bench.py reads it for roofline.traffic).
usage: python tools/summarize_r2.py r2_final"""
import collections, csv, gzip, io, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
tag = sys.argv[1]
O, P = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
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
out = ["# %s: ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 python bench.py --batch 20000 --steps 1 --warmup 1 ..." % tag,
       "# (cold-cache, serialised launch times: compare SHARES, not absolutes; 480 launches = 80 rounds of the warm-up solve)"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-62s launches=%4d total=%10.1f us share=%.3f avg=%8.1f us" % (k, a[0], a[1], a[1] / tot, a[1] / a[0]))
kk = d["roofline"]["kernels"]; tk = sum(v["ms"] for v in kk.values())
out.append("# live CUDA-event shares over the profiled pass of bench.py (batch 1e5, no profiler): " +
           ", ".join("%s %.3f" % (k.split("<")[0], v["ms"] / tk) for k, v in kk.items()))


def capture(name):
    rows = list(csv.reader(open(O + "/ncu_%s_raw.csv" % name))); hdr, units, vals = rows[0], rows[1], rows[2]
    def g(n):
        v, u = float(vals[hdr.index(n)].replace(",", "")), units[hdr.index(n)]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    src = gzip.open(O + "/ncu_%s_source.csv.gz" % name, "rt").read().splitlines()
    s2 = next(i for i, l in enumerate(src) if l.startswith('"Address"'))
    # units of work in the captured launch: executions of the kernel's 64-bit counter atomic
    its = [int(x["Instructions Executed"]) for x in csv.DictReader(io.StringIO("\n".join(src[s2:]))) if "RED" in x["Source"] and ".64" in x["Source"]]
    return dict(kernel=vals[hdr.index("Kernel Name")], dram_bytes=g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
                duration_us=float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
                {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0),
                units=max(its) if its else None,
                registers=int(float(vals[hdr.index("launch__registers_per_thread")])),
                issue_active_pct=float(vals[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                warps_active_pct=float(vals[hdr.index("sm__warps_active.avg.pct_of_peak_sustained_active")]))


traffic = {"source_digest": bench.source_digest(), "tag": tag, "workload": "goddard", "P": 85, "kernels": {}}      # profile_r2.sh captures the default workload
names = {"chain": "hybrd_chain_kernel", "qpass": "hybrd_qpass_kernel", "jac": "hybrd_jac_kernel"}
for short, full in names.items():
    try:
        c = capture(short)
    except Exception as e:           # a capture that did not happen is simply absent
        out.append("# %s: no capture (%s)" % (full, e))
        continue
    if c["units"]:
        c["dram_bytes_per_unit"] = c["dram_bytes"] / c["units"]
        traffic["kernels"][full] = c
    out.append("# %s capture (--set full): %.1f us, DRAM %.1f MB, %s units of work in the launch -> %s KB per unit; %d registers, "
               "issue slots busy %.1f %%, warps active %.1f %%" % (full, c["duration_us"], c["dram_bytes"] / 1e6, c["units"],
               ("%.1f" % (c["dram_bytes"] / c["units"] / 1e3)) if c["units"] else "?", c["registers"], c["issue_active_pct"], c["warps_active_pct"]))
json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
out.append("# bench: %.0f solves/s (e2e %.0f), reference arm %.1f solves/s on %d host cores, RK4 kernel %.1f%% of the measured FP64 peak, dominant kernel %s at %.1f%% of %s" % (
    d["value"], d["e2e"]["value"], r["value"], r["cpu_baseline"]["cores"], 100 * d["rk4_kernel"]["frac"], d["roofline"]["kernel"],
    100 * d["roofline"]["frac"], d["roofline"]["bound"]))
open(os.path.join(P, tag + "_launches_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
for f in ("bench.json", "bench_reference.json", "launches.csv", "probe_traj.log"):
    shutil.copy(os.path.join(O, f), os.path.join(P, tag + "_" + f))
for k in ("chain", "qpass", "jac", "int", "asm", "traj"):
    for suf in ("_raw.csv", "_details.txt", "_source.csv.gz"):
        if os.path.exists(os.path.join(O, "ncu_" + k + suf)):
            shutil.copy(os.path.join(O, "ncu_" + k + suf), os.path.join(P, tag + "_ncu_" + k + suf))
