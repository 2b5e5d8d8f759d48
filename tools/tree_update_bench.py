"""Times srlx_tree_update / srlx_tree_sample / srlx_tree_add (the IPriorityMemory seam kernels) for a few batch sizes with
CUDA events, on whichever library SRLX_LIB points at.  usage: [SRLX_LIB=...] python tools/tree_update_bench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
cap = 1_000_000
tree = torch.zeros(2 * cap - 1, dtype=torch.float64, device=dev)
meta = torch.zeros(C.sizeof(_lib.SrlxState), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
_lib.check(lib.srlx_tree_clear(tree.data_ptr(), cap, meta.data_ptr(), s))
g = torch.Generator(device=dev).manual_seed(0)
warm = torch.rand(100_000, dtype=torch.float64, device=dev, generator=g)
_lib.check(lib.srlx_tree_add(tree.data_ptr(), cap, meta.data_ptr(), warm.data_ptr(), 100_000, 0.8, 1e-4, 0, s))
torch.cuda.synchronize()
for n in (1, 8, 32, 64, 256):
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    w = torch.empty(n, dtype=torch.float32, device=dev)
    pr = torch.rand(n, dtype=torch.float32, device=dev, generator=g)
    reps = 200
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for phase in range(2):  # 0 = warm-up
        ev[0].record()
        for i in range(reps):
            _lib.check(lib.srlx_tree_sample(tree.data_ptr(), cap, meta.data_ptr(), n, 100 + i, 0.4, 1000.0, 1, i + 1, None, 9999,
                                            idx.data_ptr(), w.data_ptr(), None, s))
        ev[1].record()
        for i in range(reps):
            _lib.check(lib.srlx_tree_update(tree.data_ptr(), cap, meta.data_ptr(), idx.data_ptr(), pr.data_ptr(), n, 0.8, 1e-4, s))
        ev[2].record()
        for i in range(reps):
            _lib.check(lib.srlx_tree_add(tree.data_ptr(), cap, meta.data_ptr(), warm.data_ptr(), n, 0.8, 1e-4, 0, s))
        ev[3].record()
        torch.cuda.synchronize()
    print(f"n={n:4d}  sample {ev[0].elapsed_time(ev[1]) * 1e3 / reps:8.1f} us   update {ev[1].elapsed_time(ev[2]) * 1e3 / reps:8.1f} us   "
          f"add {ev[2].elapsed_time(ev[3]) * 1e3 / reps:8.1f} us   ({os.path.basename(_lib.LIB_PATH)})", flush=True)
