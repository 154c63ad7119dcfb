"""Map the stall samples of an `ncu --page source --csv` export to source lines, using the line info of
the in-tree libsocp_b200.so (nvdisasm -g).  usage: ncu_lines.py <source.csv.gz> <mangled-kernel-prefix> [top]"""
import re, gzip, csv, io, sys, subprocess, os, tempfile, glob
from collections import defaultdict
src, prefix = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "socp_b200", "libsocp_b200.so")], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = next(i for i in range(start + 1, len(txt)) if txt[i].startswith("//---------------------"))
cur, insts = None, []
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
lines = gzip.open(src, "rt").read().splitlines()
s = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[s:]))))
print("sass instructions: cubin %d, report %d" % (len(insts), len(rows)))
agg = defaultdict(lambda: [0, 0, 0])
for (loc, ins), r in zip(insts, rows):
    a = agg[loc]
    a[0] += int(r["# Samples"]); a[1] += int(r["stall_barrier"]); a[2] += int(r["Instructions Executed"])
tot = sum(a[0] for a in agg.values())
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s line %4d  samples %6d (%4.1f%%) barrier %6d  exec %d" % (loc[0], loc[1], a[0], 100 * a[0] / tot, a[1], a[2]))
