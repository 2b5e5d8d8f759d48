"""CPU restatement of the reference's image Q-network and its DQN trainer (TEST INFRASTRUCTURE; only tests/, smoke() and bench.py's CPU
legs may import this).

Follows
  srl/rl/torch_/blocks/input_image_reshape_block.py:23-70   state batch -> (batch, ch, h, w)
  srl/rl/torch_/blocks/dqn_image_block.py:10-62             Conv2d(k8 s4 p3) -> Conv2d(k4 s2 p2) -> Conv2d(k3 s1 p1), replicate padding, ReLU
  srl/rl/torch_/blocks/mlp_block.py, srl/algorithms/dqn/model_torch.py:17-29   Flatten -> Linear + ReLU ... -> Linear(A)
  srl/algorithms/dqn/dqn.py:143-173                         calc_target_q (double DQN, rescaling; no invalid actions)
  srl/algorithms/dqn/model_torch.py:75-131                  Trainer.train: HuberLoss(target * w, q * w), Adam, priorities, target sync
  srl/rl/torch_/blocks/dueling_network.py:8-59, srl/algorithms/rainbow/model_torch.py:15-29,85-122, rainbow_nomultisteps.py:10-43
                                                            rainbow (multisteps = 1) over the same image block: dueling head, float32 targets
written over a plain dict of arrays under the reference's state_dict keys (torch.nn.functional calls, no reference module).
Pinned by tests/golden/imageq_*.npz -- the reference's own Trainer.train on frozen batches (tests/golden/make_golden_image.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

CONV = [(8, 4, 3), (4, 2, 2), (3, 1, 1)]  # (kernel, stride, padding) of the DQN image block


def to_nchw(x: torch.Tensor, obs_shape, stype: str) -> torch.Tensor:
    """InputImageReshapeBlock.forward for a batch"""
    n = len(obs_shape)
    if stype == "GRAY_HW":
        return x.unsqueeze(1) if n == 2 else x
    if stype == "GRAY_HW1":
        return x.permute(0, 3, 1, 2) if n == 3 else x.reshape(x.shape[:4])
    assert n == 3, (stype, obs_shape)
    return x.permute(0, 3, 1, 2)  # RGB / IMAGE_MAP / FEATURE_MAP


def rescaling(x, eps=0.001):
    return np.sign(x) * (np.sqrt(np.abs(x) + 1.0) - 1.0) + eps * x


def inverse_rescaling(x, eps=0.001):
    n = np.sqrt(1.0 + 4.0 * eps * (np.abs(x) + 1.0 + eps)) - 1.0
    n = n / (2.0 * eps)
    return np.sign(x) * ((n**2) - 1.0)


class ImageQ:
    def __init__(self, sd, obs_shape, stype, double=True, rescale=False, discount=0.99, lr=0.001, sync_interval=1000, target_sd=None, dueling=None):
        self.obs_shape, self.stype = tuple(obs_shape), stype
        self.p = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True) for k, v in sd.items()}
        tsd = sd if target_sd is None else target_sd
        self.t = {k: torch.tensor(np.asarray(v), dtype=torch.float32) for k, v in tsd.items()}
        self.double, self.rescale, self.discount, self.sync_interval = double, rescale, discount, sync_interval
        self.opt = torch.optim.Adam(list(self.p.values()), lr=lr)
        self.train_count = 0
        self.sync_count = 0
        self.n_hidden = sum(1 for k in sd if k.startswith("hidden_block.hidden_layers.") and k.endswith(".weight") and "_layers." not in k[27:])
        self.dueling = dueling  # None | "average" | "max" | "": the block sits at hidden_block.hidden_layers.{2 * n_hidden}
        self.target_f32 = dueling is not None

    def forward(self, params, x: torch.Tensor) -> torch.Tensor:
        x = to_nchw(x, self.obs_shape, self.stype)
        for i, (k, s, pad) in enumerate(CONV):
            x = F.pad(x, (pad, pad, pad, pad), mode="replicate")
            x = F.relu(F.conv2d(x, params[f"in_block.image_block.image_layers.{2 * i}.weight"], params[f"in_block.image_block.image_layers.{2 * i}.bias"], stride=s))
        x = x.flatten(1)
        for i in range(self.n_hidden):
            x = F.relu(F.linear(x, params[f"hidden_block.hidden_layers.{2 * i}.weight"], params[f"hidden_block.hidden_layers.{2 * i}.bias"]))
        if self.dueling is None:
            return F.linear(x, params["out_layer.weight"], params["out_layer.bias"])
        base = f"hidden_block.hidden_layers.{2 * self.n_hidden}."
        v = F.linear(F.relu(F.linear(x, params[base + "v_layers.0.weight"], params[base + "v_layers.0.bias"])), params[base + "v_layers.2.weight"],
                     params[base + "v_layers.2.bias"])
        adv = F.linear(F.relu(F.linear(x, params[base + "adv_layers.0.weight"], params[base + "adv_layers.0.bias"])), params[base + "adv_layers.2.weight"],
                       params[base + "adv_layers.2.bias"])
        if self.dueling == "average":
            return v + adv - torch.mean(adv, dim=-1, keepdim=True)
        if self.dueling == "max":
            return v + adv - torch.max(adv, dim=-1, keepdim=True)[0]
        return v + adv

    def pred_q(self, state, target=False):
        with torch.no_grad():
            return self.forward(self.t if target else self.p, torch.tensor(np.asarray(state, np.float32))).numpy()

    def calc_target_q(self, n_state, reward, undone):
        n_q_target = self.pred_q(n_state, target=True)
        if self.double:
            n_q = self.pred_q(n_state)
            maxq = n_q_target[np.arange(len(reward)), np.argmax(n_q, axis=1)]
        else:
            maxq = np.max(n_q_target, axis=1)
        if self.rescale:
            maxq = inverse_rescaling(maxq)
        if self.target_f32:  # rainbow_nomultisteps.py:13-17,36: every array float32
            target_q = np.asarray(reward, np.float32) + np.asarray(undone, np.float32) * self.discount * maxq
        else:  # dqn.py: undone stays an int array -> float64
            target_q = np.asarray(reward, np.float32) + np.asarray(undone, np.int64) * self.discount * maxq
        if self.rescale:
            target_q = rescaling(target_q)
        return target_q.astype(np.float32)

    def train(self, state, n_state, action, reward, undone, weights):
        """one Trainer.train(); returns (loss, priorities, target_q)"""
        target_q = torch.tensor(self.calc_target_q(n_state, reward, undone))
        w = torch.tensor(np.asarray(weights, np.float32))
        q = self.forward(self.p, torch.tensor(np.asarray(state, np.float32)))
        q = torch.sum(q * F.one_hot(torch.tensor(np.asarray(action, np.int64)), q.shape[1]).float(), dim=1)
        loss = F.huber_loss(target_q * w, q * w)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        pri = np.abs((target_q - q).detach().numpy())
        if self.train_count % self.sync_interval == 0:
            self.t = {k: v.detach().clone() for k, v in self.p.items()}
            self.sync_count += 1
        self.train_count += 1
        return float(loss.item()), pri, target_q.numpy()

    def state_dict(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}
