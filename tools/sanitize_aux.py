"""Small workloads of the auxiliary kernels for compute-sanitizer (memcheck / racecheck / synccheck): the IPriorityMemory seam
(hashed exact-order update, sampler), the PPO returns scan (scalar and 16-byte paths, GAE and MC) and the R2D2 sequence targets."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200.memory import DeviceProportionalMemory  # noqa: E402
from simple_distributed_rl_b200.returns import returns_scan, sequence_targets  # noqa: E402

rng = np.random.default_rng(0)
m = DeviceProportionalMemory(1000, 0.8, 0.4, 1000, has_duplicate=True)
for i in range(300):
    m.add(i, float(rng.random()))
for step in range(5):
    b, w, idx = m.sample(64, step)
    m.update(idx, rng.random(64))
m2 = DeviceProportionalMemory(37, 0.6, 0.4, 1000, has_duplicate=False)
for i in range(37):
    m2.add(i, float(rng.random()))
b, w, idx = m2.sample(8, 1)
m2.update(idx, rng.random(8))
print("tree ok", m.length(), m2.length())
dev = "cuda:0"
for T, E in ((130, 64), (57, 33)):
    r = torch.randn(T, E, device=dev)
    d = (torch.rand(T, E, device=dev) < 0.05).to(torch.uint8)
    v, nv = torch.randn(T, E, device=dev), torch.randn(T, E, device=dev)
    out, valid = returns_scan(r, d, v, nv)
    out2, _ = returns_scan(r.double(), d, method="MC")
    out3, _ = returns_scan(r, d, method="MC", tail_is_episode_end=True)
print("returns ok", float(out.sum()), float(out2.sum()))
B, T, A = 70, 80, 4
t, tm, k = sequence_targets(torch.randn(B, T + 1, A, device=dev), torch.randn(B, T + 1, A, device=dev), torch.randint(0, A, (B, T), device=dev),
                            torch.rand(B, T, device=dev).double() * 0.9 + 0.05, torch.randn(B, T, device=dev).double(),
                            torch.rand(B, T, device=dev) < 0.05, enable_rescale=True)
torch.cuda.synchronize()
print("sequence targets ok", float(t.sum()))
