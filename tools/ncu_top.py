"""top stalled SASS lines of one kernel in an ncu report: python tools/ncu_top.py report.ncu-rep <kernel substring> [n]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page holds one block per kernel: a "Kernel Name" row, a header row, then instruction rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[]); blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    h = b["rows"][0]
    ci, cs, cx = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    items, tot, agg = [], 0.0, {}
    for k, r in enumerate(b["rows"][1:]):
        try:
            v = float(r[cs])
        except Exception:
            continue
        tot += v
        top = max(stall_cols, key=lambda i: float(r[i] or 0))
        items.append((v, k, r[ci][:80], h[top], r[cx]))
        for i in stall_cols:
            agg[h[i]] = agg.get(h[i], 0) + float(r[i] or 0)
    print(b["name"][:100], "samples", tot)
    print("  ", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for v, k, sl, top, x in sorted(items, reverse=True)[:n]:
        print(f"{100 * v / tot:5.1f}% #{k:5d} {top:>18} exec={x:>9}  {sl}")
