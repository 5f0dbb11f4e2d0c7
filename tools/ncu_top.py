"""Summarise an ncu report's per-CUDA-source-line warp-stall samples:
   python tools/ncu_top.py rep.ncu-rep kernel_regex [N]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
agg = {}
for hi in his:
    hdr = rows[hi]
    ni = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    fname = rows[hi - 2][1] if hi >= 2 else ""
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[0] == "Line No": break
        if r[2] != "-": continue            # only the per-line aggregate rows (Address == '-')
        key = (fname.split("/")[-1], int(r[0]))
        n = int(r[ni] or 0)
        a = agg.setdefault(key, [0, 0, r[1], {}])
        a[0] += n; a[1] += int(r[ie] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: a[3][hdr[i][6:]] = a[3].get(hdr[i][6:], 0) + v
tot = sum(a[0] for a in agg.values()) or 1
print(f"kernel {kre}: {tot} samples over {len(agg)} source lines")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*a[0]/tot:5.1f}% {f}:{ln:<4d} inst={a[1]:<9d} {a[2].strip()[:100]}  | " + ", ".join(f"{k}:{v}" for k, v in st))
