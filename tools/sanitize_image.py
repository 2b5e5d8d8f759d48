"""Small workloads of the image path for compute-sanitizer (memcheck / racecheck / synccheck): both image kernels (staged and
one-thread-per-output), the conv Q-network's forward / update on the default tiles, or -- with SRLX_IMAGE_TC3=1 in the environment -- on
the tcgen05 tiles, and the tcgen05 GEMM tap with split-K."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import _lib, image  # noqa: E402

rng = np.random.default_rng(0)
for shape, st, it, rs, nm in [((64, 48, 3), "RGB", "GRAY_HW1", (28, 36), "0to1"), ((50, 70, 3), "RGB", "GRAY_HW", (33, 21), ""),
                              ((32, 32), "GRAY_HW", "RGB", (20, 24), "-1to1"), ((40, 40, 3), "RGB", "RGB", None, "0to1")]:
    pipe = image.DeviceImagePipeline(shape, st, it, rs, nm, trimming=(2, 3, 38, 30) if rs is None else None)
    out = pipe(rng.integers(0, 256, size=(5,) + shape, dtype=np.uint8))
    torch.cuda.synchronize()
    print("pipe ok", tuple(out.shape), out.dtype)
for obs, stype, u8, hidden, duel in [((28, 36, 4), "IMAGE_MAP", True, (32,), None), ((3, 30, 26), "GRAY_HW", False, (), None),
                                     ((20, 24, 3), "RGB", False, (24, 16), None), ((28, 36, 2), "IMAGE_MAP", True, (24, 16), "max")]:
    spec = image.ImageNetSpec(obs, stype, 5, filters=8, hidden=hidden, dueling=duel)
    net = image.ImageQNet(spec, batch_size=6, uint8_states=u8, target_model_update_interval=2)
    fr = rng.integers(0, 256, size=(2, 6) + obs, dtype=np.uint8)
    x = fr if u8 else (fr / 255.0).astype(np.float32)
    for _ in range(3):
        loss, pri, tq = net.train(x[0], x[1], rng.integers(0, 5, 6), rng.normal(0, 1, 6).astype(np.float32), np.ones(6, np.float32),
                                  rng.uniform(0.3, 1, 6).astype(np.float32))
    q = net.pred_q(x[0][:4])
    torch.cuda.synchronize()
    print("imageq ok", obs, float(loss), net.train_count, net.sync_count, tuple(q.shape))
lib = _lib.load()
for (M, N, K) in [(130, 70, 100), (40, 300, 2000)]:
    A, B, C = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda"), torch.zeros(M, N, device="cuda")
    ws = torch.empty(1 << 20, device="cuda")
    _lib.check(lib.srlx_sgemm_tc3(A.data_ptr(), K, 1, B.data_ptr(), N, 1, C.data_ptr(), N, M, N, K, 0, 0, ws.data_ptr(), ws.numel(),
                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print("tc3 ok", float((C - A @ B).abs().max()))
