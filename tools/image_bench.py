"""Image path on one B200: the observation pipeline's throughput against the HBM roofline, and the conv Q-network's update at the
reference's Atari setting (DQN block 32 / 64 / 64 filters, 84 x 84 x 4 stacks, 512 hidden units, batch 32; and batch 256), with the CPU
restatement (oracle/image.py, oracle/imageq.py: torch CPU = the reference's own arithmetic) timed beside it.

    python tools/image_bench.py [--out gpurun_out/image_bench.json] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import _lib, image  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters  # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--frames", type=int, default=8192)
    args = ap.parse_args()
    lib = _lib.load()
    out = {"config": "image pipeline: 210 x 160 x 3 uint8 frames -> 84 x 84 gray '0to1' (InputImageBlockConfig DQN default); "
                     "conv Q-net: DQN block, 84 x 84 x 4, hidden 512, 6 actions"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # ---- pipeline: n frames (> L2) per launch; algorithmic bytes = frame in (uint8) + frame out (float32)
    n = args.frames
    pipe = image.DeviceImagePipeline((210, 160, 3), "RGB", "GRAY_HW1", (84, 84), "0to1")
    frames = torch.randint(0, 256, (n, 210, 160, 3), dtype=torch.uint8, device="cuda")
    dst = torch.empty((n, 84, 84, 1), dtype=torch.float32, device="cuda")
    ms = timed(lambda: pipe(frames, out=dst), 10)
    bytes_ = n * (210 * 160 * 3 + 84 * 84 * 4)
    out["pipeline"] = {"frames_per_launch": n, "ms_per_launch": ms, "frames_per_s": n / (ms * 1e-3), "algorithmic_bytes_per_frame": 210 * 160 * 3 + 84 * 84 * 4,
                       "achieved_GBps": bytes_ / (ms * 1e-3) / 1e9, "input_mb": frames.numel() / 1e6}
    pipe8 = image.DeviceImagePipeline((210, 160, 3), "RGB", "GRAY_HW", (84, 84), "")
    dst8 = torch.empty((n, 84, 84), dtype=torch.uint8, device="cuda")
    ms8 = timed(lambda: pipe8(frames, out=dst8), 10)
    out["pipeline_uint8_out"] = {"ms_per_launch": ms8, "frames_per_s": n / (ms8 * 1e-3), "achieved_GBps": n * (210 * 160 * 3 + 84 * 84) / (ms8 * 1e-3) / 1e9}
    del frames, dst, dst8
    # ---- conv Q-net
    rng = np.random.default_rng(0)
    for B in (32, 256):
        spec = image.ImageNetSpec((84, 84, 4), "IMAGE_MAP", 6)
        res = {}
        for u8 in (False, True):
            net = image.ImageQNet(spec, batch_size=B, uint8_states=u8)
            fr = torch.randint(0, 256, (2, B, 84, 84, 4), dtype=torch.uint8, device="cuda")
            x = fr if u8 else fr.float() / 255.0
            a = torch.randint(0, 6, (B,), dtype=torch.int32, device="cuda")
            r, ud, w = torch.randn(B, device="cuda"), torch.ones(B, device="cuda"), torch.rand(B, device="cuda") * 0.7 + 0.3
            l0 = lib.srlx_launch_count()
            net.train(x[0], x[1], a, r, ud, w)
            launches = lib.srlx_launch_count() - l0
            ms_u = timed(lambda: net.train(x[0], x[1], a, r, ud, w), 50, warm=5)
            ms_f = timed(lambda: net.pred_q(x[0]), 50, warm=5)
            # host batches (the plug-in trainer's path): pinned host arrays, upload inside the timed region, loss + priorities read back
            hx = [t.cpu().pin_memory() for t in (x[0], x[1])]
            hv = [t.cpu().pin_memory() for t in (a, r, ud, w)]

            def e2e():
                loss, pri, _ = net.train(hx[0].cuda(non_blocking=True), hx[1].cuda(non_blocking=True), *[t.cuda(non_blocking=True) for t in hv])
                torch.cat([loss, pri]).cpu()

            ms_e = timed(e2e, 30, warm=3)
            res["uint8_states" if u8 else "float32_states"] = {"ms_per_update": ms_u, "updates_per_s": 1e3 / ms_u, "samples_per_s": B * 1e3 / ms_u,
                                                               "launches_per_update": int(launches), "ms_per_forward": ms_f,
                                                               "e2e_ms_per_update_host_batches": ms_e, "h2d_bytes_per_update": int(sum(t.numel() * t.element_size() for t in hx + hv))}
        # flops of one update: 3 forward passes + backward (2x forward, no input gradient for conv 1)
        fl = 0
        for (c, h, w_, k, st, p, oh, ow, f, cf, off) in spec.conv_geo:
            fl += 2 * oh * ow * f * (c * k * k + 1)
        for (o, kk, off) in spec.dense:
            fl += 2 * o * (kk + 1)
        res["flops_per_sample_forward"] = fl
        res["tflops_update_float32_states"] = 5 * fl * B / (res["float32_states"]["ms_per_update"] * 1e-3) / 1e12
        if not args.no_cpu:
            from oracle import imageq as oq

            torch.set_num_threads(max(1, os.cpu_count() or 1))
            sd = {k: v.numpy() for k, v in spec.init_state_dict(0).items()}
            ora = oq.ImageQ(sd, (84, 84, 4), "IMAGE_MAP")
            st = rng.random((2, B, 84, 84, 4), dtype=np.float32)
            aa, rr, uu, ww = rng.integers(0, 6, B), rng.normal(0, 1, B).astype(np.float32), np.ones(B, np.int64), rng.uniform(0.3, 1, B).astype(np.float32)
            ora.train(st[0], st[1], aa, rr, uu, ww)
            t0 = time.perf_counter()
            n_it = 5 if B == 32 else 2
            for _ in range(n_it):
                ora.train(st[0], st[1], aa, rr, uu, ww)
            res["cpu_port_ms_per_update"] = (time.perf_counter() - t0) / n_it * 1e3
            res["cpu_threads"] = torch.get_num_threads()
        out[f"imageq_batch{B}"] = res
    if not args.no_cpu:
        from oracle import image as oimg

        f = rng.integers(0, 256, (210, 160, 3), dtype=np.uint8)
        t0 = time.perf_counter()
        for _ in range(20):
            oimg.process(f, "RGB", "GRAY_HW1", (84, 84), "0to1")
        out["pipeline"]["cpu_port_frames_per_s"] = 20 / (time.perf_counter() - t0)
        try:
            import cv2

            t0 = time.perf_counter()
            for _ in range(2000):
                g = cv2.resize(cv2.cvtColor(f, cv2.COLOR_RGB2GRAY), (84, 84)).astype(np.float32)
                g /= np.uint8(255)
            out["pipeline"]["cpu_cv2_frames_per_s_1core"] = 2000 / (time.perf_counter() - t0)
        except ImportError:
            pass
    hbm = peaks.get("hbm_gbs")
    if hbm:
        out["pipeline"]["hbm_peak_GBps"] = hbm
        out["pipeline"]["roofline_frac"] = out["pipeline"]["achieved_GBps"] / hbm
    print("IMAGEBENCH " + json.dumps(out), flush=True)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
