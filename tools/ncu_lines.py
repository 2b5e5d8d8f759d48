#!/usr/bin/env python
"""Join an ncu SASS-level source page (per-instruction warp-stall samples) with nvdisasm line info -> samples per CUDA
source line.  usage: ncu_lines.py <report.ncu-rep> <object.o> <kernel substring> [top N]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def main():
    rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # offset -> (file, line) for the kernel's text section
    off2line, cur, in_k = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            in_k = kname in ln
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    base = None
    per_line = defaultdict(lambda: defaultdict(float))
    total = 0
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        n = float(r[col["# Samples"]] or 0)
        ex = float(r[col["Instructions Executed"]] or 0)
        key = off2line.get(off, ((None, 0), ""))[0]
        d = per_line[key]
        d["samples"] += n
        d["inst"] += ex
        for sc in stall_cols:
            d[sc] += float(r[col[sc]] or 0)
        total += n
    print(f"total samples {total:.0f}")
    items = sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]
    for key, d in items:
        stalls = sorted(((v, k) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:3]
        ss = " ".join(f"{k[6:]}={v:.0f}" for v, k in stalls if v > 0)
        print(f"{str(key[0]):>18}:{key[1]:<5} samples {d['samples']:7.0f} ({100*d['samples']/max(total,1):5.1f}%) inst {d['inst']:9.0f}  {ss}")


if __name__ == "__main__":
    main()
