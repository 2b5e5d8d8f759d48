"""Debug helper: the atari-shaped lockstep case of tests/test_image_gpu.py, gradient and parameter differences per key and update."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from simple_distributed_rl_b200 import image as im
import test_image_gpu as T
import torch.nn.functional as F

obs_shape, A, B = (84, 84, 4), 6, 32
spec, net, ora = T._oracle_pair(im, obs_shape, "IMAGE_MAP", A, 32, (512,), B, True, False, seed=5, sync=3)
rng = np.random.default_rng(11)
for u in range(5):
    fr, st, a, r, ud, w = T._batch(rng, B, obs_shape, A)
    net.train(st[:, 0], st[:, 1], a, r, ud, w, phases=1)
    got = spec.to_state_dict(net.grads.cpu().numpy())
    tq = torch.tensor(ora.calc_target_q(st[:, 1], r, ud))
    q = ora.forward(ora.p, torch.tensor(st[:, 0]))
    q = torch.sum(q * F.one_hot(torch.tensor(a), A).float(), dim=1)
    for p_ in ora.p.values():
        p_.grad = None
    F.huber_loss(tq * torch.tensor(w), q * torch.tensor(w)).backward()
    print(f"--- update {u}")
    for k in spec.keys():
        want = ora.p[k].grad.numpy(); g = got[k].numpy()
        d = np.abs(g - want)
        print(f"grad {k:50s} max|g| {np.abs(want).max():.3e} med|g| {np.median(np.abs(want)):.3e} maxdiff {d.max():.3e} rel-to-max {d.max()/np.abs(want).max():.2e}  frac(|d|>1e-3|g|) {(d > 1e-3*np.abs(want)+1e-12).mean():.4f}")
    net.train(st[:, 0], st[:, 1], a, r, ud, w, phases=2)
    ora.train(st[:, 0], st[:, 1], a, r, ud, w)
    sd, osd = net.state_dict(), ora.state_dict()
    for k in spec.keys():
        d = np.abs(sd[k].numpy() - osd[k])
        print(f"par  {k:50s} maxdiff {d.max():.3e} frac>2e-5 {(d > 2e-5).mean():.5f} frac>2e-6 {(d > 2e-6).mean():.5f}")
