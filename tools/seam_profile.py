"""where an epoch of the reference's SumTree speed-test goes on the IPriorityMemory seam: host time of add / sample / update, time inside
the launch + poll, kernel duration (CUDA events)"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_distributed_rl_b200.memory import DeviceProportionalMemory

m = DeviceProportionalMemory(1_000_000, 0.8, 0.4, 1000, has_duplicate=True)
step = 0
for _ in range(100_000):
    m.add((step,) * 4, random.random()); step += 1
t = dict(add=0.0, sample=0.0, update=0.0, rand=0.0, launch=0.0)
orig = m._launch
def timed_launch(*a, **k):
    t0 = time.perf_counter(); orig(*a, **k); t["launch"] += time.perf_counter() - t0
m._launch = timed_launch
N = 5000
for _ in range(N):
    t0 = time.perf_counter(); m.add((step,) * 4, random.random()); step += 1
    t1 = time.perf_counter(); b, w, ua = m.sample(64, step)
    t2 = time.perf_counter(); pr = [random.random() for _ in range(64)]
    t3 = time.perf_counter(); m.update(ua, pr)
    t4 = time.perf_counter()
    t["add"] += t1 - t0; t["sample"] += t2 - t1; t["rand"] += t3 - t2; t["update"] += t4 - t3
print({k: round(v / N * 1e6, 2) for k, v in t.items()}, "us per epoch")
# kernel duration alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot = 0.0
for _ in range(200):
    m.add((step,) * 4, random.random()); step += 1
    m.update(ua, pr)
    e0.record(); m.sample(64, step); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print("kernel (events around launch + poll):", round(tot / 200 * 1e3, 2), "us")
