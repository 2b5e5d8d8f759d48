"""latency of srlx_tree_seam by phase: empty launch (launch + mapped flag), ops only, sample only, both; wall clock per call incl. poll"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simple_distributed_rl_b200.memory import DeviceProportionalMemory

m = DeviceProportionalMemory(1_000_000, 0.8, 0.4, 1000, has_duplicate=True)
step = 0
for _ in range(100_000):
    m.add((step,) * 4, random.random()); step += 1
m._flush()
def run(n_ops, batch, reps=2000):
    m._ops_idx[:64] = np.random.randint(999_999, 1_999_998, 64)
    m._ops_val[:64] = np.random.rand(64)
    m._ops_idx[64] = -1; m._ops_val[64] = 0.5
    t0 = time.perf_counter()
    for i in range(reps):
        m._n_ops = n_ops
        m._launch(batch, i, None, 9999)
    return (time.perf_counter() - t0) / reps * 1e6
for n_ops, batch in ((0, 0), (1, 0), (64, 0), (65, 0), (0, 64), (65, 64), (0, 32), (0, 1)):
    print(f"n_ops={n_ops:3d} batch={batch:3d}: {run(n_ops, batch):7.2f} us per call (launch + kernel + poll)")
