"""CPU (torch fp32) restatement of the reference's PPO (TEST INFRASTRUCTURE) -- **parity unpinned by execution**: the reference's PPO
is TensorFlow-only (srl/algorithms/ppo/ppo.py:6, config.py:125) and TensorFlow is not available where this was written, so this file
restates the TF code line by line in torch instead of being checked against a run of it.  What CAN be pinned is pinned elsewhere: the
worker's GAE / Monte-Carlo accumulation (oracle/gae.py against goldens produced by the reference's own Worker.on_step).

  forward            ActorCriticNetwork.call (ppo.py:88-101): trunk MLP (relu) -> value MLP -> Dense(1); -> policy MLP -> NormalDistBlock
                     (loc Dense, log_scale Dense clipped to log(stable_gradients_scale_range), normal_dist_block.py:134-155) or
                     CategoricalDistBlock (logits Dense, categorical_dist_block.py:149-153)
  policy             Worker.policy (:307-356): sample / mean, log_prob floored at log(1e-6), env action = clip(rescale_from(a))
  train_update       Trainer._train (:208-291) + compute_train_loss (:103-169), tf.clip_by_global_norm, keras Adam
                     (alpha = lr_t sqrt(1 - b2^t) / (1 - b1^t); p -= alpha m / (sqrt(v) + 1e-7)), ExponentialDecay(staircase=True)
"""
import math

import numpy as np
import torch

from . import philox

LOG_1E6 = math.log(1e-6)


def forward(layers, stack_v, stack_p, params: torch.Tensor, x: torch.Tensor):
    """layers: PPONetSpec.layers entries (name, out, k, w_off, b_off); returns (v [n], policy outputs [n, n_out])."""

    def run(stack):
        h = x
        for i, (_, out, k, w, b) in enumerate(stack):
            h = torch.nn.functional.linear(h, params[w:w + out * k].view(out, k), params[b:b + out])
            if i < len(stack) - 1:
                h = torch.relu(h)
        return h

    return run(stack_v)[:, 0], run(stack_p)


def normal_logprob(x, loc, ls):
    return -0.5 * math.log(2 * math.pi) - ls - 0.5 * (((x - loc) / torch.exp(ls)) ** 2)


def policy_noise(seed, e, g):
    """The N(0,1) draw of env e at vector step g (csrc/ppo.cu: Box-Muller cos branch over Philox(seed, STREAM_POLICY, (e, g)))."""
    w = philox.words(seed, philox.STREAM_POLICY, e, g & 0xFFFFFFFF, g >> 32)
    u1 = ((np.uint32(w[0]) >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = (np.uint32(w[1]) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    return np.float32(r * np.float32(np.cos((np.float32(2.0) * u2).astype(np.float64) * np.pi)))


class KerasAdam:
    def __init__(self, n, lr, decay_steps=0, decay_rate=1.0, b1=0.9, b2=0.999, eps=1e-7):
        self.m, self.v = torch.zeros(n), torch.zeros(n)
        self.lr, self.decay_steps, self.decay_rate, self.b1, self.b2, self.eps, self.iterations = lr, decay_steps, decay_rate, b1, b2, eps, 0

    def apply(self, params: torch.Tensor, g: torch.Tensor):
        step = self.iterations
        lr = self.lr * (self.decay_rate ** (step // self.decay_steps)) if self.decay_steps else self.lr
        t = step + 1
        alpha = np.float32(lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))
        self.m = self.m + (g - self.m) * np.float32(1 - self.b1)
        self.v = self.v + (g * g - self.v) * np.float32(1 - self.b2)
        self.iterations += 1
        return params - alpha * self.m / (torch.sqrt(self.v) + np.float32(self.eps))


def train_update(spec, params: np.ndarray, adam: KerasAdam, cfg, continuous, states, actions, old_v, old_logp, ret):
    """One Trainer._train on a given minibatch.  cfg: the PPOConfig fields.  Returns (new params, info)."""
    p = torch.tensor(np.asarray(params, dtype=np.float32), requires_grad=True)
    x = torch.as_tensor(np.asarray(states, dtype=np.float32))
    if cfg.enable_state_normalized:
        x = (x - x.mean(dim=0, keepdim=True)) / (x.std(dim=0, unbiased=False, keepdim=True) + 1e-8)
    v_target = torch.as_tensor(np.asarray(ret, dtype=np.float32))
    adv = v_target.clone()
    if cfg.baseline_type == "ave":
        adv = adv - adv.mean()
    elif cfg.baseline_type == "std":
        adv = adv / (adv.std(unbiased=False) + 1e-8)
    elif cfg.baseline_type == "normal":
        adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    v, po = forward(spec.layers, spec.stack_v, spec.stack_p, p, x)
    act = torch.as_tensor(np.asarray(actions, dtype=np.float32))
    if continuous:
        lo, hi = math.log(cfg.stable_gradients_scale_range[0]), math.log(cfg.stable_gradients_scale_range[1])
        new_logpi = normal_logprob(act, po[:, 0], torch.clamp(po[:, 1], lo, hi))
    else:
        new_logpi = torch.log_softmax(po, dim=-1).gather(1, act.long()[:, None])[:, 0]
    if cfg.baseline_type in ("advantage", "v"):
        adv = adv - v.detach()
    ratio = torch.exp(new_logpi - torch.as_tensor(np.asarray(old_logp, dtype=np.float32)))
    if cfg.surrogate_type == "clip":
        rc = torch.clamp(ratio, 1 - cfg.policy_clip_range, 1 + cfg.policy_clip_range)
        policy_loss = torch.minimum(ratio * adv, rc * adv)
    else:
        policy_loss = ratio * adv
    policy_loss = -policy_loss.mean()
    ov = torch.as_tensor(np.asarray(old_v, dtype=np.float32))
    if cfg.enable_value_clip:
        vc = torch.maximum(torch.minimum(v, ov + cfg.value_clip_range), ov - cfg.value_clip_range)
        value_loss = torch.maximum((v - v_target) ** 2, (vc - v_target) ** 2)
    else:
        value_loss = (v - v_target) ** 2
    value_loss = cfg.value_loss_weight * value_loss.mean()
    entropy_loss = cfg.entropy_weight * -(-torch.exp(new_logpi) * new_logpi).mean()
    loss = policy_loss + value_loss + entropy_loss
    loss.backward()
    g = p.grad.detach()
    norm = float(torch.sqrt((g.double() ** 2).sum()))
    if cfg.global_gradient_clip_norm != 0:
        g = g * np.float32(cfg.global_gradient_clip_norm / max(norm, cfg.global_gradient_clip_norm))
    new_p = adam.apply(p.detach(), g)
    return new_p.numpy(), dict(policy_loss=float(policy_loss.detach()), value_loss=float(value_loss.detach()), entropy_loss=float(entropy_loss.detach()), grad=g.numpy(),
                               grad_norm=norm)
