"""Throughput of the tensor-core inference mode (csrc/qnet_tc.cu) against the fp32 seam and against the measured bf16 peak.

    python tools/qnet_tc_bench.py [--out profiles/r2_x_qnet_tc_bench.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import _lib  # noqa: E402
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("bf16_tflops", 0) or 0)
    lib = _lib.load()
    out = {"peak_bf16_tflops_measured": peak, "dense": [], "qnet": []}
    for M, N, K in [(8192, 512, 512), (65536, 512, 512), (262144, 512, 512), (262144, 1024, 1024), (65536, 4096, 4096)]:
        x = torch.randn((M, K), device="cuda").to(torch.bfloat16)
        w = torch.randn((N, K), device="cuda").to(torch.bfloat16)
        b = torch.randn(N, device="cuda")
        y = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        ms = timeit(lambda: _lib.check(lib.srlx_dense_bf16_tc(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, 0, M, N, K, 1, s)))
        ms_t = timeit(lambda: torch.relu(torch.addmm(b.to(torch.bfloat16), x, w.T)))
        fl = 2.0 * M * N * K
        rec = {"M": M, "N": N, "K": K, "ms": ms, "tflops": fl / ms / 1e9, "frac_of_measured_peak": (fl / ms / 1e9 / peak) if peak else None,
               "torch_cublas_ms": ms_t, "torch_cublas_tflops": fl / ms_t / 1e9, "bytes_gb_s": (M * K + N * K + M * N) * 2 / ms / 1e6}
        out["dense"].append(rec)
        print(rec, flush=True)
    for name, kw, n in [("dqn_mlp512x512", dict(env="CartPole-v1", algo="dqn", hidden=(512, 512), mem_kind=0), 262144),
                        ("rainbow_default", dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3), 262144),
                        ("dqn_mlp64x64", dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), mem_kind=0), 262144)]:
        eng = DeviceEngine(EngineConfig(n_envs=8, ring_rows=4, batch_size=4, warmup_size=4, **kw))
        x = torch.randn((n, eng.D), device="cuda")
        ms_tc = timeit(lambda: eng.pred_q_tc(x), iters=10)
        xh = x.cpu().numpy()
        q = torch.empty((n, eng.A), dtype=torch.float32, device="cuda")
        import ctypes as C
        try:
            ms_f32 = timeit(lambda: _lib.check(lib.srlx_qnet_forward(C.byref(eng.c), 0, x.data_ptr(), n, 0, q.data_ptr(), eng._stream())), iters=5)
        except _lib.SrlxError as e:  # the fp32 seam keeps the whole network in shared memory: wide layers do not fit
            ms_f32 = None
        rec = {"net": name, "states": n, "tc_ms": ms_tc, "fp32_seam_ms": ms_f32, "speedup": (ms_f32 / ms_tc) if ms_f32 else None,
               "states_per_s_tc": n / ms_tc * 1e3, "fp32_seam": "ok" if ms_f32 else "network too large for the shared-memory-resident fp32 kernel"}
        out["qnet"].append(rec)
        print(rec, flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
