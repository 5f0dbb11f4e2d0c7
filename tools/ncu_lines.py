"""Attribute the per-SASS-instruction counters of an ncu report (--page source --csv) to CUDA source lines, using
nvdisasm -g line markers of the same kernel (the report itself carries no CUDA view when the sources are not resolvable).
usage: python tools/ncu_lines.py <sass.csv> <nvdisasm.txt> <mangled kernel name> [top N]"""
import csv, re, sys
sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. instruction -> line from nvdisasm
lines = open(dis, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." + kname in l)
cur = None; seq = []
for l in lines[start + 1:]:
    if l.strip().startswith(".section") or l.strip().startswith("//-----"):
        if seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2)))
# 2. counters from the report (same instruction order)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]; ci = h.index("Instructions Executed"); si = h.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) > ci]
assert abs(len(data) - len(seq)) < 8, (len(data), len(seq))
agg = {}
for (loc, txt), r in zip(seq, data):
    a = agg.setdefault(loc, [0, 0]); a[0] += int(r[ci] or 0); a[1] += int(r[si] or 0)
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
src = {}
print(f"total warp instructions {tot}, samples {ts}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if loc:
        try:
            if loc[0] not in src: src[loc[0]] = open("zeekstd_b200/csrc/" + loc[0]).read().split("\n")
            text = src[loc[0]][loc[1] - 1].strip()[:110]
        except Exception: pass
    print(f"{a[0] / tot * 100:5.1f}% inst {a[1] / max(ts,1) * 100:5.1f}% smp  {loc}  {text}")
