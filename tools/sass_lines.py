#!/usr/bin/env python
"""Code footprint per source line range of one kernel: tools/sass_lines.py <object.o> <kernel substring> [bucket]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
obj, kname = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt, files, cur, in_k = Counter(), Counter(), None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        in_k = kname in ln
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln) and cur:
        files[cur[0]] += 1
        if cur[0].startswith("learner_fast") or cur[0].startswith("learner_small"):
            cnt[cur[1] // bucket * bucket] += 1
        else:
            cnt[cur[0]] += 1
print("instructions by file:", dict(files))
for k, v in sorted(cnt.items(), key=lambda kv: (isinstance(kv[0], str), kv[0])):
    print(f"{k!s:>28}: {v:5d} instr {v*16/1024:6.1f} KB")
