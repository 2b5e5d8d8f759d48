"""CPU tests of the PPO host logic and of the oracle restatement (no GPU): parameter layout of the two-stack network view, the keras
Adam / staircase decay restatement, and that the loss of oracle/ppo.py behaves as the reference's clipped objective must."""
import math

import numpy as np
import torch

from oracle import ppo as oppo
from simple_distributed_rl_b200.ppo import PPOConfig, PPONetSpec


def test_two_stack_view_shares_the_trunk_and_covers_every_parameter_once():
    spec = PPONetSpec(3, (64, 64), (32,), (16, 8), 2)
    nv, npol = spec.nets()
    assert nv.n_layers == 4 and npol.n_layers == 5 and nv.n_actions == 1 and npol.n_actions == 2
    assert [nv.w_off[i] for i in range(2)] == [npol.w_off[i] for i in range(2)]          # trunk: same offsets in both stacks
    assert [nv.b_off[i] for i in range(2)] == [npol.b_off[i] for i in range(2)]
    covered = np.zeros(spec.n_params, dtype=np.int32)
    for _, out, k, w, b in spec.layers:
        covered[w:w + out * k] += 1
        covered[b:b + out] += 1
    assert (covered == 1).all()
    p = spec.init_params(0, True)
    name = {l[0]: l for l in spec.layers}
    for n in ("trunk0", "value0", "policy1"):
        _, out, k, w, b = name[n]
        assert abs(p[w:w + out * k].std() - math.sqrt(2.0 / k)) < 0.25 * math.sqrt(2.0 / k) and (p[b:b + out] == 0).all()   # he_normal, zero bias
    _, out, k, w, b = name["value_out"]
    assert abs(np.linalg.norm(p[w:w + k]) - 1.0) < 1e-5                                                                     # orthogonal row


def test_keras_adam_restatement_and_staircase_decay():
    adam = oppo.KerasAdam(3, lr=0.1, decay_steps=2, decay_rate=0.5)
    p = torch.zeros(3)
    g = torch.tensor([1.0, -2.0, 0.5])
    lrs = []
    for step in range(5):
        q = adam.apply(p, g)
        m_hat = 1.0  # constant gradient: m / sqrt(v) -> sign(g) after bias correction
        lrs.append(float((p - q)[0]))
        p = q
    # |delta| = lr_t * m_hat / (sqrt(v_hat) + eps) ~= lr * 0.5^floor(step / 2) for a constant gradient
    np.testing.assert_allclose(lrs, [0.1, 0.1, 0.05, 0.05, 0.025], rtol=1e-4)


def test_clipped_objective_stops_the_gradient_outside_the_clip_range():
    spec = PPONetSpec(3, (8,), (), (), 2)
    cfg = PPOConfig(hidden_block=(8,), value_block=(), policy_block=(), baseline_type="", entropy_weight=0.0, value_loss_weight=0.0,
                    global_gradient_clip_norm=0.0, enable_value_clip=False)
    p = spec.init_params(1, True)
    rng = np.random.default_rng(0)
    x, a = rng.normal(size=(32, 3)).astype(np.float32), rng.normal(size=32).astype(np.float32)
    with torch.no_grad():
        v, po = oppo.forward(spec.layers, spec.stack_v, spec.stack_p, torch.as_tensor(p), torch.as_tensor(x))
        lp = oppo.normal_logprob(torch.as_tensor(a), po[:, 0], torch.clamp(po[:, 1], math.log(1e-10), math.log(10))).numpy()
    ret = np.ones(32, dtype=np.float32)  # positive advantage everywhere
    # old log-probs far BELOW the new ones: ratio >> 1 + clip with adv > 0 -> the clipped branch is the minimum and has no gradient
    _, info = oppo.train_update(spec, p, oppo.KerasAdam(spec.n_params, 1e-3), cfg, True, x, a, np.zeros(32), lp - 5.0, ret)
    assert np.abs(info["grad"]).max() == 0.0
    # ratio == 1: the plain policy gradient
    _, info = oppo.train_update(spec, p, oppo.KerasAdam(spec.n_params, 1e-3), cfg, True, x, a, np.zeros(32), lp, ret)
    assert np.abs(info["grad"]).max() > 0.0 and abs(info["policy_loss"] + 1.0) < 1e-5
