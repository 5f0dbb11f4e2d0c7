import numpy as np, sys
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
ok = (t[:, 7] > 0)
t = t[ok]
t0 = t[:, 0].min()
names = ["begin", "meta_loaded", "started", "lits_done", "deps_resolved", "first_ready", "matches_done", "marked_done"]
d = np.diff(t, axis=1)
print("chunks traced:", len(t))
print("median cycles per stage:", {f"{names[i]}->{names[i+1]}": int(np.median(d[:, i])) for i in range(7)})
print("p90    cycles per stage:", {f"{names[i]}->{names[i+1]}": int(np.percentile(d[:, i], 90)) for i in range(7)})
done = np.sort(t[:, 7])
print("median gap between consecutive chunk completions:", int(np.median(np.diff(done))), "mean", int(np.mean(np.diff(done))))
print("total span for", len(t), "chunks:", int(done[-1] - t0))
for c in range(40, 56):
    print(c, (t[c] - t0).tolist())
