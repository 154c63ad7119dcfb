"""Summarise an `ncu --page source --csv` export: stall reasons overall and the hottest SASS regions."""
import csv, gzip, sys, io
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
lines = f.read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stalls = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
tot = {k: 0 for k in stalls}
total = 0
for r in rows:
    s = int(r["# Samples"] or 0)
    total += s
    for k in stalls:
        tot[k] += int(r[k] or 0)
print("kernel:", lines[0][:150])
print("total samples", total, " instructions", len(rows))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-28s %8d  %.1f%%" % (k, v, 100.0 * v / max(total, 1)))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
print("hottest instructions (index: samples, main stall, SASS):")
for i in sorted(order):
    r = rows[i]
    main = max(stalls, key=lambda k: int(r[k] or 0))
    print("  %6d: %6s %-16s exec=%-8s %s" % (i, r["# Samples"], main.replace("stall_", ""), r["Instructions Executed"], r["Source"].strip()[:90]))
