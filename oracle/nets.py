"""CPU (torch fp32) restatement of the reference Q-networks and of one trainer update (TEST INFRASTRUCTURE).

  NetSpec / flat layout   the srlx_net layout of include/srlx.h, with converters to/from the reference state_dict keys
                          (dqn QNetwork srl/algorithms/dqn/model_torch.py:17-29; rainbow QNetwork
                          srl/algorithms/rainbow/model_torch.py:15-29; MLPBlock srl/rl/torch_/blocks/mlp_block.py:9-47;
                          DuelingNetworkBlock srl/rl/torch_/blocks/dueling_network.py:8-59;
                          NoisyLinear srl/rl/torch_/modules/noisy_linear.py:8-52)
  forward                 y = relu(x W^T + b) ..., dueling combine, noisy W = mu + sigma * eps with eps INJECTED
  train_update            one Trainer.train(): dqn/model_torch.py:90-132, rainbow/model_torch.py:85-122 given the sampled
                          batch, IS weights and the three noise draws; uses torch autograd + torch.optim.Adam (the same
                          third-party code the reference calls).

Pinned by tests/golden/trainer_*.npz = the reference Trainer.train() executed here on frozen batches.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import targets

DUEL_NONE, DUEL_AVERAGE, DUEL_MAX, DUEL_NAIVE = 0, 1, 2, 3
_DUEL = {None: DUEL_NONE, "none": DUEL_NONE, "average": DUEL_AVERAGE, "max": DUEL_MAX, "": DUEL_NAIVE, "naive": DUEL_NAIVE}

NOISE_KIND_ROLLOUT = 0
NOISE_KIND_TRAIN = 1  # call_id = update*3 + pass (0: online(s), 1: online(s'), 2: target(s'))
NOISE_KIND_PRED = 3


@dataclass
class NetSpec:
    in_dim: int
    hidden: Tuple[int, ...]
    n_actions: int
    dueling: Optional[str] = None  # None | "average" | "max" | "" (naive)
    noisy: bool = False

    def __post_init__(self):
        self.duel = _DUEL[self.dueling]
        hid = tuple(int(h) for h in self.hidden)
        outs, ks = [], []
        k = self.in_dim
        if self.duel == DUEL_NONE:
            for h in hid:
                outs.append(h)
                ks.append(k)
                k = h
            outs.append(self.n_actions)
            ks.append(k)
        else:
            for h in hid[:-1]:
                outs.append(h)
                ks.append(k)
                k = h
            H = hid[-1]
            outs.append(2 * H)
            ks.append(k)
            outs.append(1 + self.n_actions)
            ks.append(H)
        self.out_dim, self.k_dim = outs, ks
        self.n_layers = len(outs)
        self.w_off, self.b_off = [], []
        off = 0
        for o, kk in zip(outs, ks):
            self.w_off.append(off)
            off += o * kk
            self.b_off.append(off)
            off += o
        self.n_params = off

    # ---- reference state_dict <-> flat ------------------------------------------------------------------
    def _keys(self, algo):
        """[(flat_off, shape, key_mu, key_sigma)] in flat order; algo in {"dqn", "rainbow"}."""
        wk, bk = ("w_mu", "b_mu") if self.noisy else ("weight", "bias")
        ws, bs = "w_sigma", "b_sigma"
        ent = []
        n_trunk = self.n_layers - 1 if self.duel == DUEL_NONE else self.n_layers - 2
        for l in range(n_trunk):
            base = f"hidden_block.hidden_layers.{2 * l}."
            ent.append((self.w_off[l], (self.out_dim[l], self.k_dim[l]), base + wk, base + ws))
            ent.append((self.b_off[l], (self.out_dim[l],), base + bk, base + bs))
        if self.duel == DUEL_NONE:
            l = self.n_layers - 1
            if algo == "dqn":
                base = "out_layer."
                ent.append((self.w_off[l], (self.out_dim[l], self.k_dim[l]), base + "weight", None))
                ent.append((self.b_off[l], (self.out_dim[l],), base + "bias", None))
            else:  # rainbow MLP: block.add_layer(nn.Linear(...)) is a plain Linear even when noisy (dueling_network.py:125-127)
                base = f"hidden_block.hidden_layers.{2 * n_trunk}."
                ent.append((self.w_off[l], (self.out_dim[l], self.k_dim[l]), base + "weight", None))
                ent.append((self.b_off[l], (self.out_dim[l],), base + "bias", None))
        else:
            base = f"hidden_block.hidden_layers.{2 * n_trunk}."
            lh, lo = self.n_layers - 2, self.n_layers - 1
            H, K, A = self.k_dim[lo], self.k_dim[lh], self.n_actions
            ent.append((self.w_off[lh], (H, K), base + "v_layers.0." + wk, base + "v_layers.0." + ws))
            ent.append((self.w_off[lh] + H * K, (H, K), base + "adv_layers.0." + wk, base + "adv_layers.0." + ws))
            ent.append((self.b_off[lh], (H,), base + "v_layers.0." + bk, base + "v_layers.0." + bs))
            ent.append((self.b_off[lh] + H, (H,), base + "adv_layers.0." + bk, base + "adv_layers.0." + bs))
            ent.append((self.w_off[lo], (1, H), base + "v_layers.2." + wk, base + "v_layers.2." + ws))
            ent.append((self.w_off[lo] + H, (A, H), base + "adv_layers.2." + wk, base + "adv_layers.2." + ws))
            ent.append((self.b_off[lo], (1,), base + "v_layers.2." + bk, base + "v_layers.2." + bs))
            ent.append((self.b_off[lo] + 1, (A,), base + "adv_layers.2." + bk, base + "adv_layers.2." + bs))
        return ent

    def from_state_dict(self, sd, algo):
        mu = np.zeros(self.n_params, dtype=np.float32)
        sigma = np.zeros(self.n_params, dtype=np.float32)
        for off, shape, kmu, ksig in self._keys(algo):
            n = int(np.prod(shape))
            mu[off : off + n] = sd[kmu].detach().cpu().numpy().reshape(-1)
            if self.noisy and ksig is not None and ksig in sd:
                sigma[off : off + n] = sd[ksig].detach().cpu().numpy().reshape(-1)
        return mu, (sigma if self.noisy else None)

    def to_state_dict(self, mu, sigma, algo):
        sd = {}
        for off, shape, kmu, ksig in self._keys(algo):
            n = int(np.prod(shape))
            sd[kmu] = torch.tensor(np.asarray(mu[off : off + n]).reshape(shape).copy())
            if self.noisy and ksig is not None:
                sd[ksig] = torch.tensor(np.asarray(sigma[off : off + n]).reshape(shape).copy())
        return sd

    def sigma_mask(self, algo="rainbow"):
        """1 where a sigma parameter exists (the plain out Linear of a noisy rainbow MLP has none)."""
        m = np.zeros(self.n_params, dtype=np.float32)
        if not self.noisy:
            return m
        for off, shape, kmu, ksig in self._keys(algo):
            if ksig is not None:
                m[off : off + int(np.prod(shape))] = 1
        return m


def forward(spec: NetSpec, mu: torch.Tensor, sigma: Optional[torch.Tensor], noise: Optional[torch.Tensor], x: torch.Tensor):
    """Q(x) from flat parameters.  noise: flat N(0,1) draw (same layout) for this forward call, or None."""
    h = x
    for l in range(spec.n_layers):
        o, k = spec.out_dim[l], spec.k_dim[l]
        W = mu[spec.w_off[l] : spec.w_off[l] + o * k].view(o, k)
        b = mu[spec.b_off[l] : spec.b_off[l] + o]
        if spec.noisy and sigma is not None and noise is not None:
            W = W + sigma[spec.w_off[l] : spec.w_off[l] + o * k].view(o, k) * noise[spec.w_off[l] : spec.w_off[l] + o * k].view(o, k)
            b = b + sigma[spec.b_off[l] : spec.b_off[l] + o] * noise[spec.b_off[l] : spec.b_off[l] + o]
        last = l == spec.n_layers - 1
        if last and spec.duel != DUEL_NONE:
            H = k
            v = F.linear(h[:, :H], W[:1], b[:1])
            adv = F.linear(h[:, H:], W[1:], b[1:])
            if spec.duel == DUEL_AVERAGE:
                h = v + adv - torch.mean(adv, dim=-1, keepdim=True)
            elif spec.duel == DUEL_MAX:
                h = v + adv - torch.max(adv, dim=-1, keepdim=True)[0]
            else:
                h = v + adv
        else:
            h = F.linear(h, W, b)
            if not last:
                h = torch.relu(h)
    return h


def np_forward(spec, mu, sigma, noise, x):
    with torch.no_grad():
        t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32))
        return forward(spec, t(mu), t(sigma), t(noise), t(x)).numpy()


class AdamState:
    """torch.optim.Adam over the flat buffers (lr, betas=(0.9,0.999), eps=1e-8: model_torch.py:78)."""

    def __init__(self, spec: NetSpec, mu, sigma, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.spec = spec
        self.mu = torch.tensor(np.asarray(mu, dtype=np.float32), requires_grad=True)
        ps = [self.mu]
        self.sigma = None
        if spec.noisy:
            self.sigma = torch.tensor(np.asarray(sigma, dtype=np.float32), requires_grad=True)
            ps.append(self.sigma)
        self.opt = torch.optim.Adam(ps, lr=lr, betas=betas, eps=eps)


def train_update(
    spec: NetSpec,
    st: AdamState,
    tgt_mu,
    tgt_sigma,
    *,
    algo,
    states,
    actions,
    rewards,
    dones,
    weights,
    discount,
    multisteps=1,
    retrace_h=1.0,
    enable_double_dqn=True,
    enable_rescale=False,
    noise=(None, None, None),
    sigma_mask=None,
    huber_delta=1.0,
    next_invalid=None,
):
    """One Trainer.train() given an already-sampled batch.  next_invalid: optional bool [B, M, A], invalid actions of the M next states.

    states [B, M+1, D]; actions/rewards/dones [B, M] where dones = "terminated" (so undone = 1 - dones for the 1-step
    algorithms, dqn.py:243).  noise = (online(s) draw, online(s') draw, target(s') draw), flat, or None when not noisy.
    Returns dict(loss, target_q, q, priorities, grad_mu, grad_sigma).
    """
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32))
    n0, n1, n2 = [t(n) for n in noise]
    tmu, tsig = t(tgt_mu), t(tgt_sigma)

    def pred_q(x):
        with torch.no_grad():
            return forward(spec, st.mu, st.sigma, n1, t(x)).numpy()

    def pred_target_q(x):
        with torch.no_grad():
            return forward(spec, tmu, tsig, n2, t(x)).numpy()

    B, M = actions.shape
    if algo == "rainbow" and multisteps > 1:
        target_q, state, act = targets.rainbow_target(
            pred_q, pred_target_q, states, actions, rewards, dones, discount, multisteps, retrace_h,
            enable_double_dqn, enable_rescale, n_actions=spec.n_actions, next_invalid=next_invalid)
    else:
        target_q = targets.dqn_target(
            pred_q, pred_target_q, states[:, 1, :], rewards[:, 0].astype(np.float32), (1 - dones[:, 0]).astype(np.int64),
            discount, enable_double_dqn, enable_rescale, next_invalid=None if next_invalid is None else next_invalid[:, 0, :])
        state, act = states[:, 0, :], actions[:, 0]

    onehot = torch.as_tensor(np.eye(spec.n_actions, dtype=np.float32)[act])
    w = t(weights)
    tq = t(target_q)
    q_all = forward(spec, st.mu, st.sigma, n0, t(state))
    q = torch.sum(q_all * onehot, dim=1)
    loss = torch.nn.HuberLoss(delta=huber_delta)(tq * w, q * w)
    st.opt.zero_grad()
    loss.backward()
    if spec.noisy and sigma_mask is not None:
        st.sigma.grad *= torch.as_tensor(sigma_mask)
    g_mu = st.mu.grad.detach().numpy().copy()
    g_sigma = st.sigma.grad.detach().numpy().copy() if spec.noisy else None
    st.opt.step()
    pri = np.abs((tq - q).detach().numpy())
    return dict(loss=float(loss.item()), target_q=np.asarray(target_q), q=q.detach().numpy().copy(), priorities=pri,
                grad_mu=g_mu, grad_sigma=g_sigma)
