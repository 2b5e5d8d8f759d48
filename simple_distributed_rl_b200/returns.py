"""Worker-side returns of the reference's PPO on a device rollout buffer (the worker half).

`ppo.Worker.on_step` (srl/algorithms/ppo/ppo.py:357-406) turns a finished episode into per-step "discounted_reward" values --
GAE advantages (`experience_collection_method="GAE"`, the default; `discount`, `gae_discount` of ppo/config.py:67-70) or
Monte-Carlo returns ("MC") -- after clipping the reward (`reward_clip`).  A vectorised on-policy rollout holds the same
quantities time-major, [T, E] with E env copies: `returns_scan` runs the accumulation for every column in one launch of
csrc/returns.cu (srlx_returns_scan) and returns what the worker would have handed to `memory.add()`, in place.

The PPO trainer (clipped surrogate, `ppo.py:103-291`) is not built yet; this is the piece of the path that is a pure function of
the rollout buffer.  There is no CPU fallback."""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

METHODS = {"GAE": _lib.RETURNS_GAE, "MC": _lib.RETURNS_MC}


def returns_scan(reward: torch.Tensor, done: torch.Tensor, value: Optional[torch.Tensor] = None, next_value: Optional[torch.Tensor] = None,
                 discount: float = 0.9, gae_discount: float = 0.9, method: str = "GAE", reward_clip: Optional[Tuple[float, float]] = None,
                 tail_is_episode_end: bool = False, with_valid: bool = True):
    """reward [T, E] float32 (float64 allowed for MC: the reference accumulates python floats), done [T, E] uint8 / bool,
    value / next_value [T, E] float32 (GAE) -> (returns float32 [T, E], valid uint8 [T, E] or None).  valid = 0 marks the
    steps of episodes that have not ended inside the buffer (the reference emits an episode only when it is done)."""
    if method not in METHODS:
        raise ValueError(f"experience_collection_method {method!r} (GAE or MC)")
    if not reward.is_cuda:
        raise _lib.SrlxError("returns_scan needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    T, E = reward.shape
    dev = reward.device
    done8 = done.to(torch.uint8).contiguous()
    r32 = r64 = None
    if reward.dtype == torch.float64:
        if method != "MC":
            raise ValueError("float64 rewards are only meaningful for MC (GAE is float32 arithmetic in the reference)")
        r64 = reward.contiguous()
    else:
        r32 = reward.to(torch.float32).contiguous()
    if method == "GAE":
        if value is None or next_value is None:
            raise ValueError("GAE needs value and next_value")
        value, next_value = value.to(torch.float32).contiguous(), next_value.to(torch.float32).contiguous()
        if value.shape != reward.shape or next_value.shape != reward.shape:
            raise ValueError("value / next_value must have the shape of reward")
    if done8.shape != reward.shape:
        raise ValueError("done must have the shape of reward")
    out = torch.empty((T, E), dtype=torch.float32, device=dev)
    valid = torch.empty((T, E), dtype=torch.uint8, device=dev) if with_valid else None
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    lo, hi = (reward_clip if reward_clip is not None else (0.0, 0.0))
    with torch.cuda.device(dev):
        _lib.check(lib.srlx_returns_scan(ptr(r32), ptr(r64), ptr(value), ptr(next_value), ptr(done8), ptr(out), ptr(valid), int(T), int(E),
                                         float(discount), float(gae_discount), METHODS[method], int(tail_is_episode_end),
                                         int(reward_clip is not None), float(lo), float(hi), torch.cuda.current_stream(dev).cuda_stream))
    return out, valid


def sequence_targets(q_online: torch.Tensor, q_target: torch.Tensor, actions: torch.Tensor, mu: torch.Tensor, rewards: torch.Tensor,
                     dones: torch.Tensor, discount: float = 0.997, retrace_h: float = 1.0, enable_double_dqn: bool = True,
                     enable_rescale: bool = False, enable_retrace: bool = True):
    """The per-sequence target loop of the reference's R2D2 trainer (srl/algorithms/r2d2/r2d2.py:150-203; SURVEY.md 8a R14, the
    target / Retrace / priority half) on the device: q_online / q_target [B, T+1, A] float32, actions [B, T], mu (stored behaviour
    probabilities) / rewards [B, T] (float64: the reference keeps them as python floats), dones [B, T] ->
    (target float64 [B, T], td_mean float64 [B], td_is_float64 bool [B]).  `td_mean` is the sequence's priority input (:204);
    where the reference holds it as float32 its value here is that float32 widened.  Not implemented: the invalid-action mask of the
    reference loop (next_invalid_actions -> -inf before the argmax, calc_epsilon_greedy_probs with invalid actions, r2d2.py:160-176) --
    the envs on the device path have none.  The LSTM Q-network, burn-in and the sequence replay around this loop are not built.
    There is no CPU fallback."""
    if not q_online.is_cuda:
        raise _lib.SrlxError("sequence_targets needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    B, T1, A = q_online.shape
    T = T1 - 1
    dev = q_online.device
    qo, qt = q_online.to(torch.float32).contiguous(), q_target.to(torch.float32).contiguous()
    if qt.shape != qo.shape or tuple(actions.shape) != (B, T):
        raise ValueError("q_target must have the shape of q_online [B, T+1, A] and actions the shape [B, T]")
    for name, x in (("mu", mu), ("rewards", rewards), ("dones", dones)):
        if tuple(x.shape) != (B, T):  # the kernel indexes [b][t] with T steps per row: any other shape reads out of bounds
            raise ValueError(f"{name} must have the shape [B, T] = {(B, T)}, got {tuple(x.shape)}")
    if T < 1 or T > 128:
        raise ValueError(f"sequence length T = {T} outside [1, 128]")
    act = actions.to(torch.int32).contiguous()
    mu64, r64 = mu.to(torch.float64).contiguous(), rewards.to(torch.float64).contiguous()
    dn = dones.to(torch.uint8).contiguous()
    target = torch.empty((B, T), dtype=torch.float64, device=dev)
    td_mean = torch.empty(B, dtype=torch.float64, device=dev)
    td_kind = torch.empty(B, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.srlx_sequence_targets(qo.data_ptr(), qt.data_ptr(), act.data_ptr(), mu64.data_ptr(), r64.data_ptr(), dn.data_ptr(),
                                             target.data_ptr(), td_mean.data_ptr(), td_kind.data_ptr(), int(B), int(T), int(A),
                                             float(discount), float(retrace_h), int(enable_double_dqn), int(enable_rescale),
                                             int(enable_retrace), torch.cuda.current_stream(dev).cuda_stream))
    return target, td_mean, td_kind.bool()
