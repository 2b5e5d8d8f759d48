"""Q-network description and flat parameter layout (include/srlx.h `srlx_net`) <-> reference state_dict keys.

Mirrors the modules the reference builds:
  dqn QNetwork      srl/algorithms/dqn/model_torch.py:17-29      (hidden_block.hidden_layers.{2i}.*, out_layer.*)
  rainbow QNetwork  srl/algorithms/rainbow/model_torch.py:15-29  (hidden_block.hidden_layers.{2i}.*, dueling block keys
                    v_layers.{0,2}.*, adv_layers.{0,2}.*; NoisyLinear keys w_mu/w_sigma/b_mu/b_sigma)
so `call_backup()` / `call_restore()` interchange checkpoints with reference-trained parameters
(srl/rl/torch_/helper.py:60-93).  Initialisers follow srl/rl/torch_/blocks/mlp_block.py:26-33 (he_normal + zero bias for
hidden Linear), torch defaults for the output / dueling Linear, noisy_linear.py:24-33 for NoisyLinear.
"""
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

_DUEL = {None: _lib.DUEL_NONE, "none": _lib.DUEL_NONE, "average": _lib.DUEL_AVERAGE, "max": _lib.DUEL_MAX,
         "": _lib.DUEL_NAIVE, "naive": _lib.DUEL_NAIVE}


@dataclass
class NetSpec:
    in_dim: int
    hidden: Tuple[int, ...]
    n_actions: int
    dueling: Optional[str] = None  # None | "average" | "max" | "" (naive)
    noisy: bool = False
    algo: str = "dqn"  # "dqn" | "rainbow": state_dict naming + which layers are NoisyLinear

    def __post_init__(self):
        self.duel = _DUEL[self.dueling]
        hid = tuple(int(h) for h in self.hidden)
        if self.duel != _lib.DUEL_NONE and len(hid) < 1:
            raise ValueError("a dueling head needs at least one layer size")
        outs, ks = [], []
        k = int(self.in_dim)
        if self.duel == _lib.DUEL_NONE:
            for h in hid:
                outs.append(h); ks.append(k); k = h
            outs.append(self.n_actions); ks.append(k)
        else:
            for h in hid[:-1]:
                outs.append(h); ks.append(k); k = h
            H = hid[-1]
            outs.append(2 * H); ks.append(k)
            outs.append(1 + self.n_actions); ks.append(H)
        if len(outs) > _lib.SRLX_MAX_LAYERS:
            raise ValueError(f"too many layers ({len(outs)} > {_lib.SRLX_MAX_LAYERS})")
        self.out_dim, self.k_dim = outs, ks
        self.n_layers = len(outs)
        self.w_off, self.b_off = [], []
        off = 0
        for o, kk in zip(outs, ks):
            self.w_off.append(off); off += o * kk
            self.b_off.append(off); off += o
        self.n_params = off
        # which layers draw noise: every NoisyLinear; the plain out Linear appended to a noisy rainbow MLP does not
        self.layer_noisy = [1 if self.noisy else 0] * self.n_layers
        if self.noisy and self.duel == _lib.DUEL_NONE:
            self.layer_noisy[-1] = 0

    # ---- C struct -------------------------------------------------------------------------------------
    def to_c(self) -> "_lib.SrlxNet":
        n = _lib.SrlxNet()
        n.n_layers, n.in_dim, n.n_params, n.n_actions = self.n_layers, self.in_dim, self.n_params, self.n_actions
        n.dueling, n.noisy = self.duel, int(self.noisy)
        for l in range(self.n_layers):
            n.out_dim[l], n.k_dim[l], n.w_off[l], n.b_off[l] = self.out_dim[l], self.k_dim[l], self.w_off[l], self.b_off[l]
            n.layer_noisy[l] = self.layer_noisy[l]
        return n

    # ---- state_dict keys -------------------------------------------------------------------------------
    def entries(self):
        """[(flat_off, shape, mu_key, sigma_key or None)] in flat order."""
        wk, bk = ("w_mu", "b_mu") if self.noisy else ("weight", "bias")
        ent = []
        n_trunk = self.n_layers - 1 if self.duel == _lib.DUEL_NONE else self.n_layers - 2
        for l in range(n_trunk):
            base = f"hidden_block.hidden_layers.{2 * l}."
            ent.append((self.w_off[l], (self.out_dim[l], self.k_dim[l]), base + wk, base + "w_sigma" if self.noisy else None))
            ent.append((self.b_off[l], (self.out_dim[l],), base + bk, base + "b_sigma" if self.noisy else None))
        if self.duel == _lib.DUEL_NONE:
            l = self.n_layers - 1
            base = "out_layer." if self.algo == "dqn" else f"hidden_block.hidden_layers.{2 * n_trunk}."
            ent.append((self.w_off[l], (self.out_dim[l], self.k_dim[l]), base + "weight", None))
            ent.append((self.b_off[l], (self.out_dim[l],), base + "bias", None))
        else:
            base = f"hidden_block.hidden_layers.{2 * n_trunk}."
            lh, lo = self.n_layers - 2, self.n_layers - 1
            H, K, A = self.k_dim[lo], self.k_dim[lh], self.n_actions
            sg = (lambda s: s) if self.noisy else (lambda s: None)
            ent += [
                (self.w_off[lh], (H, K), base + "v_layers.0." + wk, sg(base + "v_layers.0.w_sigma")),
                (self.w_off[lh] + H * K, (H, K), base + "adv_layers.0." + wk, sg(base + "adv_layers.0.w_sigma")),
                (self.b_off[lh], (H,), base + "v_layers.0." + bk, sg(base + "v_layers.0.b_sigma")),
                (self.b_off[lh] + H, (H,), base + "adv_layers.0." + bk, sg(base + "adv_layers.0.b_sigma")),
                (self.w_off[lo], (1, H), base + "v_layers.2." + wk, sg(base + "v_layers.2.w_sigma")),
                (self.w_off[lo] + H, (A, H), base + "adv_layers.2." + wk, sg(base + "adv_layers.2.w_sigma")),
                (self.b_off[lo], (1,), base + "v_layers.2." + bk, sg(base + "v_layers.2.b_sigma")),
                (self.b_off[lo] + 1, (A,), base + "adv_layers.2." + bk, sg(base + "adv_layers.2.b_sigma")),
            ]
        return ent

    def from_state_dict(self, sd):
        mu = np.zeros(self.n_params, dtype=np.float32)
        sigma = np.zeros(self.n_params, dtype=np.float32) if self.noisy else None
        for off, shape, kmu, ksig in self.entries():
            n = int(np.prod(shape))
            t = sd[kmu]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"state_dict[{kmu!r}] has shape {tuple(t.shape)}, expected {tuple(shape)}")
            mu[off:off + n] = t.detach().cpu().numpy().reshape(-1)
            if ksig is not None:
                sigma[off:off + n] = sd[ksig].detach().cpu().numpy().reshape(-1)
        return mu, sigma

    def to_state_dict(self, mu, sigma=None):
        sd = {}
        for off, shape, kmu, ksig in self.entries():
            n = int(np.prod(shape))
            sd[kmu] = torch.from_numpy(np.ascontiguousarray(mu[off:off + n]).reshape(shape).copy())
            if ksig is not None:
                sd[ksig] = torch.from_numpy(np.ascontiguousarray(sigma[off:off + n]).reshape(shape).copy())
        return sd

    # ---- initialisation ---------------------------------------------------------------------------------
    def init_params(self, seed: int = 0):
        """Draw initial parameters with the reference's initialisers (torch CPU generator)."""
        gen = torch.Generator().manual_seed(int(seed))
        mu = np.zeros(self.n_params, dtype=np.float32)
        sigma = np.zeros(self.n_params, dtype=np.float32) if self.noisy else None
        n_trunk = self.n_layers - 1 if self.duel == _lib.DUEL_NONE else self.n_layers - 2

        def default_linear(out_f, in_f):
            bound = 1.0 / math.sqrt(in_f)  # nn.Linear.reset_parameters: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(in))
            w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
            b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
            return w, b

        def he_linear(out_f, in_f):
            w = torch.randn(out_f, in_f, generator=gen) * math.sqrt(2.0 / in_f)  # kaiming_normal_ (he_normal)
            return w, torch.zeros(out_f)

        def noisy_linear(out_f, in_f):
            stdv = 1.0 / math.sqrt(in_f)
            w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * stdv
            b = (torch.rand(out_f, generator=gen) * 2 - 1) * stdv
            return w, b, 0.5 * stdv

        for off, shape, kmu, ksig in self.entries():
            pass
        for l in range(self.n_layers):
            o, k = self.out_dim[l], self.k_dim[l]
            is_trunk = l < n_trunk
            if self.duel != _lib.DUEL_NONE and not is_trunk:
                # dueling hidden (2H x K) / out ((1+A) x H): two independent Linear each
                parts = [(o // 2, k), (o // 2, k)] if l == self.n_layers - 2 else [(1, k), (self.n_actions, k)]
                ws, bs, sgs = [], [], []
                for (oo, kk) in parts:
                    if self.noisy:
                        w, b, s = noisy_linear(oo, kk); sgs.append(s)
                    else:
                        w, b = default_linear(oo, kk)
                    ws.append(w); bs.append(b)
                W, Bv = torch.cat(ws, 0), torch.cat(bs, 0)
                mu[self.w_off[l]:self.w_off[l] + o * k] = W.numpy().reshape(-1)
                mu[self.b_off[l]:self.b_off[l] + o] = Bv.numpy()
                if self.noisy:
                    r0 = 0
                    for (oo, kk), s in zip(parts, sgs):
                        sigma[self.w_off[l] + r0 * k:self.w_off[l] + (r0 + oo) * k] = s
                        sigma[self.b_off[l] + r0:self.b_off[l] + r0 + oo] = s
                        r0 += oo
                continue
            if self.layer_noisy[l]:
                w, b, s = noisy_linear(o, k)
                sigma[self.w_off[l]:self.w_off[l] + o * k] = s
                sigma[self.b_off[l]:self.b_off[l] + o] = s
            elif is_trunk:
                w, b = he_linear(o, k)
            else:
                w, b = default_linear(o, k)
            mu[self.w_off[l]:self.w_off[l] + o * k] = w.numpy().reshape(-1)
            mu[self.b_off[l]:self.b_off[l] + o] = b.numpy()
        return mu, sigma
