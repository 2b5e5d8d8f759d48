"""A few updates of the conv Q-network at the Atari setting, for `ncu --metrics gpu__time_duration.sum` launch lists (tools/launch_summary.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_distributed_rl_b200 import image  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spec = image.ImageNetSpec((84, 84, 4), "IMAGE_MAP", 6)
net = image.ImageQNet(spec, batch_size=B, uint8_states=True)
fr = torch.randint(0, 256, (2, B, 84, 84, 4), dtype=torch.uint8, device="cuda")
a = torch.randint(0, 6, (B,), dtype=torch.int32, device="cuda")
r, ud, w = torch.randn(B, device="cuda"), torch.ones(B, device="cuda"), torch.rand(B, device="cuda") * 0.7 + 0.3
for _ in range(3):
    net.train(fr[0], fr[1], a, r, ud, w)
torch.cuda.synchronize()
