"""SumTree / ProportionalMemory micro-benchmark under the reference's own protocol.

Protocol (tests/quick/rl/memories/speedtest.py:30-58 of the reference): capacity 1 M, alpha 0.8, beta 0.4 over 1000 steps,
has_duplicate=True; 100 000 warm-up adds with random priorities, then 5 000 x {add 1, sample 64, update 64 random priorities};
the figure is the wall time of the whole run.  Arms:

  python   the reference's ProportionalMemory (srl/rl/memories/priority_memories/proportional_memory.py), only where
           /root/reference is importable (the build container); elsewhere the oracle's restatement (oracle/sumtree.py) stands in
  cpp      the reference's pybind11 module compiled from its own sources (oracle/_ref, built by oracle/Makefile)
  seam     DeviceProportionalMemory driven item by item through the IPriorityMemory methods, host lists in and out: add / update
           append to an op list in mapped pinned host memory, every sample() is ONE launch (srlx_tree_seam) that applies the list,
           draws the batch and writes the results + a sequence word back to mapped host memory, which the host polls
  device   the same three C-ABI calls (srlx_tree_add / _sample / _update) with every operand resident in HBM and no host
           synchronisation inside the loop: what a device-side consumer sees (3 launches per epoch)
  fused    for scale: the fused engine does add (8192 leaves) / sample / update inside rollout + learner kernels; its SumTree
           share per update comes from tools/phase_clocks.py, not from this script

Run on the GPU box:  python tools/sumtree_speedtest.py --out gpurun_out/sumtree_speedtest.json
This is measurement tooling: it may import oracle/ (CPU arms); the product path does not.
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CAPACITY, WARMUP, BATCH, EPOCHS = 1_000_000, 100_000, 64, 5_000
ALPHA, BETA0, BETA_STEPS = 0.8, 0.4, 1000


def protocol(memory, warmup=WARMUP, epochs=EPOCHS, batch_size=BATCH):
    """The reference's _speed_test loop, statement for statement, returning (total seconds, seconds of the epoch part)."""
    t0 = time.perf_counter()
    step = 0
    for _ in range(warmup):
        memory.add((step, step, step, step), random.random())
        step += 1
    t1 = time.perf_counter()
    for _ in range(epochs):
        memory.add((step, step, step, step), random.random())
        step += 1
        batches, weights, update_args = memory.sample(batch_size, step)
        assert len(batches) == batch_size and len(weights) == batch_size
        memory.update(update_args, [random.random() for _ in range(batch_size)])
    t2 = time.perf_counter()
    return t2 - t0, t2 - t1


def arm_python():
    ref = "/root/reference"
    if os.path.isdir(os.path.join(ref, "srl")):
        sys.path.insert(0, ref)
        try:
            from srl.rl.memories.priority_memories.proportional_memory import ProportionalMemory
        finally:
            sys.path.remove(ref)
        return "reference ProportionalMemory (python)", ProportionalMemory(CAPACITY, ALPHA, BETA0, BETA_STEPS, has_duplicate=True)
    from oracle.sumtree import ProportionalMemory

    class Port:  # the port has no payload list and takes its uniforms from a callback: adapt it to the protocol's calls
        def __init__(self):
            self.m = ProportionalMemory(CAPACITY, ALPHA, BETA0, BETA_STEPS, has_duplicate=True)

        def add(self, batch, priority=None):
            self.m.add(priority)

        def sample(self, batch_size, step):
            idx, w, _, _ = self.m.sample(batch_size, step, lambda i, k: random.random())
            return idx, w, idx

        def update(self, indices, priorities):
            self.m.update(indices, priorities)

    return "oracle restatement (python, oracle/sumtree.py)", Port()


def arm_cpp():
    import glob
    import importlib.util

    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "proportional_memory_cpp*.so"))
    if not so:
        return None, None
    spec = importlib.util.spec_from_file_location("proportional_memory_cpp", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return "reference C++ module (oracle/_ref)", mod.ProportionalMemory(CAPACITY, ALPHA, BETA0, BETA_STEPS, True)


def arm_seam():
    from simple_distributed_rl_b200.memory import DeviceProportionalMemory

    return "DeviceProportionalMemory (IPriorityMemory seam: one launch per sample, mapped host memory)", DeviceProportionalMemory(CAPACITY, ALPHA, BETA0, BETA_STEPS, has_duplicate=True)


def run_device_resident(epochs=EPOCHS):
    """The three C-ABI calls with operands in HBM; timed with CUDA events around the epoch loop (no host sync inside)."""
    import ctypes as C

    import torch

    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    tree = torch.zeros(2 * CAPACITY - 1, dtype=torch.float64, device=dev)
    meta = torch.zeros(C.sizeof(_lib.SrlxState), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(0)
    eps = 0.0001
    _lib.check(lib.srlx_tree_clear(tree.data_ptr(), CAPACITY, meta.data_ptr(), s))
    warm = torch.rand(WARMUP, dtype=torch.float64, device=dev, generator=g)
    add_p = torch.rand(epochs, dtype=torch.float64, device=dev, generator=g)
    upd_p = torch.rand(epochs, BATCH, dtype=torch.float32, device=dev, generator=g)
    idx = torch.empty(BATCH, dtype=torch.int64, device=dev)
    w = torch.empty(BATCH, dtype=torch.float32, device=dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    _lib.check(lib.srlx_tree_add(tree.data_ptr(), CAPACITY, meta.data_ptr(), warm.data_ptr(), WARMUP, ALPHA, eps, 0, s))  # one launch
    e1.record()
    a_ptr, u_ptr = add_p.data_ptr(), upd_p.data_ptr()
    for i in range(epochs):
        _lib.check(lib.srlx_tree_add(tree.data_ptr(), CAPACITY, meta.data_ptr(), a_ptr + 8 * i, 1, ALPHA, eps, 0, s))
        _lib.check(lib.srlx_tree_sample(tree.data_ptr(), CAPACITY, meta.data_ptr(), BATCH, WARMUP + i + 1, BETA0, float(BETA_STEPS), 1,
                                        i + 1, None, 9999, idx.data_ptr(), w.data_ptr(), None, s))
        _lib.check(lib.srlx_tree_update(tree.data_ptr(), CAPACITY, meta.data_ptr(), idx.data_ptr(), u_ptr + 4 * BATCH * i, BATCH,
                                        ALPHA, eps, s))
    e2.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # invariant after the run: every internal node is the sum of its children (size-independent parity property)
    t = tree.cpu().numpy()
    n_int = CAPACITY - 1
    import numpy as np

    err = float(np.max(np.abs(t[:n_int] - (t[1: 2 * n_int: 2] + t[2: 2 * n_int + 1: 2]))))
    return {"arm": "device-resident C-ABI calls (3 launches per epoch, no host sync)", "total_s": wall,
            "warmup_ms_device": e0.elapsed_time(e1), "epochs_ms_device": e1.elapsed_time(e2),
            "us_per_epoch_device": e1.elapsed_time(e2) * 1e3 / epochs, "us_per_epoch_wall": None,
            "tree_sum_invariant_max_abs_err": err, "root": float(t[0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip", default="", help="comma list of arms to skip: python,cpp,seam,device")
    a = ap.parse_args()
    skip = set(x for x in a.skip.split(",") if x)
    rows = []
    random.seed(0)
    for key, make in (("python", arm_python), ("cpp", arm_cpp), ("seam", arm_seam)):
        if key in skip:
            continue
        name, mem = make()
        if mem is None:
            rows.append({"arm": key, "unavailable": "oracle/_ref not built"})
            continue
        total, ep = protocol(mem)
        rows.append({"arm": name, "total_s": total, "epochs_s": ep, "us_per_epoch_wall": ep * 1e6 / EPOCHS,
                     "us_per_warmup_add": (total - ep) * 1e6 / WARMUP})
        print(json.dumps(rows[-1]), flush=True)
    if "device" not in skip:
        rows.append(run_device_resident())
        print(json.dumps(rows[-1]), flush=True)
    out = {"protocol": {"capacity": CAPACITY, "warmup_adds": WARMUP, "epochs": EPOCHS, "batch": BATCH, "alpha": ALPHA,
                        "beta_initial": BETA0, "beta_steps": BETA_STEPS, "source": "reference tests/quick/rl/memories/speedtest.py:30-58"},
           "host_cores": os.cpu_count(), "rows": rows}
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
