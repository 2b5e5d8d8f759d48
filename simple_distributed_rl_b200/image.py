"""The reference's image configs on device: the observation pipeline and the conv Q-network with its DQN trainer.

    reference                                                              here                           libsrlx entry (csrc/imageq.cu)
    ImageProcessor.remap_observation (rl/processors/image_processor.py)    DeviceImagePipeline            srlx_image_process
    InputImageBlock + DQNImageBlock + hidden block + out layer             ImageNetSpec (layout, keys)    -
      (rl/torch_/blocks/*.py, algorithms/dqn/model_torch.py:17-29)
    Parameter.pred_q / pred_target_q (model_torch.py:60-72)                ImageQNet.pred_q               srlx_imageq_forward
    Trainer.train (model_torch.py:75-131) + calc_target_q (dqn.py:143-173) ImageQNet.train                srlx_imageq_train
    PriorityReplayBuffer demo memory (priority_replay_buffer.py:177-240)   DemoMixMemory                  (host logic over the device memories)

Everything numeric runs in libsrlx.so; there is no CPU fallback (a missing library or a CPU tensor raises)."""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

GRAY_HW, GRAY_HW1, RGB = "GRAY_HW", "GRAY_HW1", "RGB"  # srl.base.define.SpaceTypes names
_NORM = {"": 0, "0to1": 1, "-1to1": 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# =====================================================================================================================
class DeviceImagePipeline:
    """ImageProcessor (image_processor.py:17-154) for batches of uint8 frames resident in HBM.

    src_shape / src_type describe the env's observation space ((H, W), (H, W, 1) or (H, W, 3); "GRAY_HW" / "GRAY_HW1" / "RGB"),
    image_type / resize (w, h) / normalize_type / trimming (top, left, bottom, right) are the reference class's fields.  One kernel
    launch per batch: colour conversion, trimming, resize and normalisation fused, frames read once."""

    def __init__(self, src_shape: Sequence[int], src_type: str, image_type: str = GRAY_HW, resize: Optional[Tuple[int, int]] = None,
                 normalize_type: str = "", trimming: Optional[Tuple[int, int, int, int]] = None, max_val: float = 255.0,
                 device: str = "cuda:0"):
        if image_type not in (GRAY_HW, GRAY_HW1, RGB) or src_type not in (GRAY_HW, GRAY_HW1, RGB):
            raise ValueError(f"image types are GRAY_HW / GRAY_HW1 / RGB (got {src_type!r} -> {image_type!r})")  # image_processor.py:39-50
        if normalize_type not in _NORM:
            raise ValueError(f"normalize_type {normalize_type!r}")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SrlxError("DeviceImagePipeline needs a CUDA device: there is no CPU fallback")
        self.src_shape = tuple(int(v) for v in src_shape)
        H, W = self.src_shape[:2]
        src_c = 3 if src_type == RGB else 1
        if self.src_shape not in ((H, W, src_c), (H, W)) or (src_type == GRAY_HW1 and len(self.src_shape) != 3):
            raise ValueError(f"source shape {self.src_shape} does not fit {src_type}")
        out_c = 3 if image_type == RGB else 1
        top, left, bottom, right = (0, 0, H, W)
        if trimming is not None:  # image_processor.py:55-69
            top, left, bottom, right = trimming
            if not (top < bottom and left < right):
                raise ValueError("trimming: top < bottom and left < right")
            top, left, bottom, right = max(top, 0), max(left, 0), min(bottom, H), min(right, W)
        th, tw = bottom - top, right - left
        oh, ow = (th, tw) if resize is None else (int(resize[1]), int(resize[0]))
        self.out_shape = (oh, ow) if image_type == GRAY_HW else (oh, ow, out_c)
        if image_type == GRAY_HW and src_type == GRAY_HW1 and resize is None:
            self.out_shape = (oh, ow, 1)  # nothing drops the trailing axis in the reference either (cv2.resize would have)
        self.out_dtype = torch.uint8 if normalize_type == "" else torch.float32
        p = _lib.SrlxImageProc()
        p.src_h, p.src_w, p.src_c = H, W, src_c
        p.top, p.left, p.trim_h, p.trim_w = top, left, th, tw
        p.out_h, p.out_w, p.out_c = oh, ow, out_c
        p.resize, p.normalize, p.max_val = int(resize is not None), _NORM[normalize_type], float(max_val)
        self._tables = []
        if resize is not None:
            for dst, src, border in ((ow, tw, 1), (oh, th, 0)):
                idx, coef = np.zeros(dst, np.int32), np.zeros((dst, 2), np.int32)
                _lib.check(self.lib.srlx_image_linear_table(dst, src, border, idx.ctypes.data, coef.ctypes.data))
                self._tables += [torch.from_numpy(idx).to(self.device), torch.from_numpy(coef).to(self.device)]
            p.x_idx, p.x_coef, p.y_idx, p.y_coef = (t.data_ptr() for t in self._tables)
        self.c = p

    def __call__(self, frames, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8 [n, *src_shape] (or one frame) on the device, or a numpy array (copied over) -> [n, *out_shape]."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames)).to(self.device, non_blocking=True)
        single = frames.dim() == len(self.src_shape)
        if single:
            frames = frames.unsqueeze(0)
        if frames.dtype != torch.uint8 or tuple(frames.shape[1:]) != self.src_shape or not frames.is_cuda:
            raise ValueError(f"frames must be a CUDA uint8 tensor [n, {self.src_shape}], got {frames.dtype} {tuple(frames.shape)} on {frames.device}")
        frames = frames.contiguous()
        n = frames.shape[0]
        if n == 0:
            return torch.empty((0,) + self.out_shape, dtype=self.out_dtype, device=self.device)
        if out is None:
            out = torch.empty((n,) + self.out_shape, dtype=self.out_dtype, device=self.device)
        elif out.dtype != self.out_dtype or tuple(out.shape) != (n,) + self.out_shape or not out.is_contiguous():
            raise ValueError("out has the wrong dtype / shape")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_image_process(C.byref(self.c), frames.data_ptr(), n, out.data_ptr(), int(np.prod(self.out_shape)), _stream()))
        return out[0] if single else out


# =====================================================================================================================
DQN_CONV = ((8, 4, 3), (4, 2, 2), (3, 1, 1))  # DQNImageBlock: (kernel, stride, replicate padding), filters f, 2f, 2f


@dataclass
class ImageNetSpec:
    """Layout of the conv Q-network (include/srlx.h `srlx_imageq`) <-> the reference's state_dict.

    obs_shape / obs_type: the RL observation space the reference hands the network (after processors and window stacking):
    (H, W) GRAY_HW, (H, W, 1) GRAY_HW1, (H, W, C) RGB / IMAGE_MAP / FEATURE_MAP -- channel-fastest frames -- or (len, H, W)
    GRAY_HW / (len, H, W, 1) GRAY_HW1 -- a stack of planes (input_image_reshape_block.py:23-70)."""
    obs_shape: Tuple[int, ...]
    obs_type: str
    n_actions: int
    filters: int = 32
    hidden: Tuple[int, ...] = (512,)
    conv: Tuple[Tuple[int, int, int], ...] = DQN_CONV
    conv_filters: Optional[Tuple[int, ...]] = None
    dueling: Optional[str] = None  # None: Linear(A) (dqn); "average" | "max" | "" (naive): rainbow's DuelingNetworkBlock, hidden[-1] = its units

    def __post_init__(self):
        s, t = tuple(int(v) for v in self.obs_shape), self.obs_type
        self.obs_shape = s
        if t == GRAY_HW and len(s) == 2:
            C_, H, W, planes = 1, s[0], s[1], False
        elif t == GRAY_HW and len(s) == 3:
            C_, H, W, planes = s[0], s[1], s[2], True
        elif t == GRAY_HW1 and len(s) == 3 and s[-1] == 1:
            C_, H, W, planes = 1, s[0], s[1], False
        elif t == GRAY_HW1 and len(s) == 4 and s[-1] == 1:
            C_, H, W, planes = s[0], s[1], s[2], True
        elif t in (RGB, "IMAGE_MAP", "FEATURE_MAP", "COLOR") and len(s) == 3:
            C_, H, W, planes = s[2], s[0], s[1], False
        else:
            raise ValueError(f"unknown space_type: {t} {s}")  # the reference's message
        self.in_c, self.in_h, self.in_w, self.planes = C_, H, W, planes
        # element strides (b, c, h, w) of a contiguous state batch
        self.in_strides = (C_ * H * W, H * W, W, 1) if planes else (C_ * H * W, 1, W * C_, C_)
        if self.conv_filters is None:
            self.conv_filters = tuple([self.filters] + [self.filters * 2] * (len(self.conv) - 1))
        if len(self.conv) > _lib.SRLX_MAX_CONV or len(self.hidden) + 1 > _lib.SRLX_MAX_LAYERS:
            raise ValueError("too many layers")
        self.conv_geo = []  # (C, H, W, k, s, p, OH, OW, F, c_fast, off)
        off, c, h, w = 0, C_, H, W
        for l, ((k, st, p), f) in enumerate(zip(self.conv, self.conv_filters)):
            oh, ow = (h + 2 * p - k) // st + 1, (w + 2 * p - k) // st + 1
            if h + 2 * p < k or w + 2 * p < k:
                raise ValueError(f"conv layer {l}: empty output for a {h} x {w} input")
            c_fast = (l > 0) or (self.in_strides[1] == 1)
            self.conv_geo.append((c, h, w, k, st, p, oh, ow, f, c_fast, off))
            off += f * (c * k * k + 1)
            c, h, w = f, oh, ow
        self.flat = c * h * w
        self.last_chw = (c, h, w)
        self.dense = []  # (out, k, off)
        kk = self.flat
        hid = tuple(int(v) for v in self.hidden)
        self.duel = {None: _lib.DUEL_NONE, "none": _lib.DUEL_NONE, "average": _lib.DUEL_AVERAGE, "max": _lib.DUEL_MAX, "": _lib.DUEL_NAIVE,
                     "naive": _lib.DUEL_NAIVE}[self.dueling]
        if self.duel != _lib.DUEL_NONE:
            if len(hid) < 1:
                raise ValueError("a dueling head needs at least one layer size (its hidden units)")
            self.duel_hidden = hid[-1]
            outs = hid[:-1] + (2 * hid[-1], 1 + int(self.n_actions))  # [value-hidden ; advantage-hidden], then [V ; Adv]
        else:
            self.duel_hidden = 0
            outs = hid + (int(self.n_actions),)
        if len(outs) > _lib.SRLX_MAX_LAYERS:
            raise ValueError("too many layers")
        for out in outs:
            self.dense.append((out, kk, off))
            off += out * (kk + 1)
            kk = out
        self.n_params = off

    # ---- state_dict <-> flat -----------------------------------------------------------------------------------------
    def _conv_key(self, l):
        return f"in_block.image_block.image_layers.{2 * l}."

    def _dense_key(self, l):
        return "out_layer." if l == len(self.dense) - 1 else f"hidden_block.hidden_layers.{2 * l}."

    def _dense_parts(self):
        """[(layer l, row0, rows, col0, cols, key prefix)]: which Linear of the reference fills which sub-block of dense layer l."""
        n, parts = len(self.dense), []
        if self.duel == _lib.DUEL_NONE:
            for l, (out, kk, off) in enumerate(self.dense):
                parts.append((l, 0, out, 0, kk, self._dense_key(l)))
            return parts
        for l in range(n - 2):
            out, kk, off = self.dense[l]
            parts.append((l, 0, out, 0, kk, f"hidden_block.hidden_layers.{2 * l}."))
        base, H, A = f"hidden_block.hidden_layers.{2 * (n - 2)}.", self.duel_hidden, int(self.n_actions)
        kk = self.dense[n - 2][1]
        parts += [(n - 2, 0, H, 0, kk, base + "v_layers.0."), (n - 1, 0, 1, 0, H, base + "v_layers.2."),
                  (n - 2, H, H, 0, kk, base + "adv_layers.0."), (n - 1, 1, A, H, H, base + "adv_layers.2.")]
        return parts

    def keys(self) -> List[str]:
        ks = []
        for l in range(len(self.conv_geo)):
            ks += [self._conv_key(l) + "weight", self._conv_key(l) + "bias"]
        for (_, _, _, _, _, key) in self._dense_parts():
            ks += [key + "weight", key + "bias"]
        return ks

    def from_state_dict(self, sd) -> np.ndarray:
        flat = np.zeros(self.n_params, np.float32)
        arr = lambda v: v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)  # noqa: E731
        for l, (c, h, w, k, st, p, oh, ow, f, c_fast, off) in enumerate(self.conv_geo):
            W_, b = arr(sd[self._conv_key(l) + "weight"]), arr(sd[self._conv_key(l) + "bias"])
            if W_.shape != (f, c, k, k):
                raise ValueError(f"state_dict[{self._conv_key(l)}weight] has shape {W_.shape}, expected {(f, c, k, k)}")
            cols = W_.transpose(0, 2, 3, 1).reshape(f, -1) if c_fast else W_.reshape(f, -1)
            flat[off:off + f * (c * k * k + 1)] = np.concatenate([cols, b.reshape(f, 1)], axis=1).reshape(-1)
        for (l, r0, rows, c0, cols, key) in self._dense_parts():
            out, kk, off = self.dense[l]
            W_, b = arr(sd[key + "weight"]), arr(sd[key + "bias"])
            if W_.shape != (rows, cols):
                raise ValueError(f"state_dict[{key}weight] has shape {W_.shape}, expected {(rows, cols)}")
            if l == 0:  # torch flattens (c, h, w); the device's conv output is (h, w, c)
                c, h, w = self.last_chw
                W_ = W_.reshape(rows, c, h, w).transpose(0, 2, 3, 1).reshape(rows, cols)
            blk = flat[off:off + out * (kk + 1)].reshape(out, kk + 1)  # a view: off-branch blocks of a dueling head stay zero
            blk[r0:r0 + rows, c0:c0 + cols] = W_
            blk[r0:r0 + rows, kk] = b
        return flat

    def to_state_dict(self, flat: np.ndarray):
        sd = {}
        for l, (c, h, w, k, st, p, oh, ow, f, c_fast, off) in enumerate(self.conv_geo):
            blk = np.asarray(flat[off:off + f * (c * k * k + 1)]).reshape(f, c * k * k + 1)
            cols = blk[:, :-1]
            W_ = cols.reshape(f, k, k, c).transpose(0, 3, 1, 2) if c_fast else cols.reshape(f, c, k, k)
            sd[self._conv_key(l) + "weight"] = torch.from_numpy(np.ascontiguousarray(W_).copy())
            sd[self._conv_key(l) + "bias"] = torch.from_numpy(blk[:, -1].copy())
        for (l, r0, rows, c0, cols, key) in self._dense_parts():
            out, kk, off = self.dense[l]
            blk = np.asarray(flat[off:off + out * (kk + 1)]).reshape(out, kk + 1)
            W_ = blk[r0:r0 + rows, c0:c0 + cols]
            if l == 0:
                c, h, w = self.last_chw
                W_ = W_.reshape(rows, h, w, c).transpose(0, 3, 1, 2).reshape(rows, cols)
            sd[key + "weight"] = torch.from_numpy(np.ascontiguousarray(W_).copy())
            sd[key + "bias"] = torch.from_numpy(blk[r0:r0 + rows, kk].copy())
        return sd

    def init_state_dict(self, seed: int = 0):
        """The reference's initialisers: torch defaults for Conv2d and the output Linear (U(+-1/sqrt(fan_in))), he_normal + zero bias
        for the hidden Linear layers (srl/rl/torch_/blocks/mlp_block.py:26-33)."""
        gen = torch.Generator().manual_seed(int(seed))
        sd = {}
        for l, (c, h, w, k, st, p, oh, ow, f, c_fast, off) in enumerate(self.conv_geo):
            bound = 1.0 / math.sqrt(c * k * k)
            sd[self._conv_key(l) + "weight"] = (torch.rand(f, c, k, k, generator=gen) * 2 - 1) * bound
            sd[self._conv_key(l) + "bias"] = (torch.rand(f, generator=gen) * 2 - 1) * bound
        for (l, r0, rows, c0, cols, key) in self._dense_parts():
            if "hidden_block.hidden_layers" in key and "_layers." not in key.split("hidden_layers.")[1]:  # MLPBlock Linear: he_normal, zero bias
                sd[key + "weight"] = torch.randn(rows, cols, generator=gen) * math.sqrt(2.0 / cols)
                sd[key + "bias"] = torch.zeros(rows)
            else:  # out_layer and the dueling block's Linear layers: torch defaults
                bound = 1.0 / math.sqrt(cols)
                sd[key + "weight"] = (torch.rand(rows, cols, generator=gen) * 2 - 1) * bound
                sd[key + "bias"] = (torch.rand(rows, generator=gen) * 2 - 1) * bound
        return sd


# =====================================================================================================================
class ImageQNet:
    """The conv Q-network and its DQN trainer in HBM.  `pred_q` / `pred_target_q` / `train` take state batches as CUDA tensors shaped
    [n, *obs_shape] -- float32 as the reference's memory holds them, or uint8 frames with `uint8_states=True` (normalised "0to1" on the
    fly inside the first im2col; 4x fewer bytes to store and move) -- or numpy arrays (copied over)."""

    def __init__(self, spec: ImageNetSpec, batch_size: int = 32, enable_double_dqn: bool = True, enable_rescale: bool = False,
                 discount: float = 0.99, lr: float = 0.001, target_model_update_interval: int = 1000, adam_beta1: float = 0.9,
                 adam_beta2: float = 0.999, adam_eps: float = 1e-8, uint8_states: bool = False, max_val: float = 255.0, seed: int = 0,
                 device: str = "cuda:0", batch_cap: Optional[int] = None, target_f32: Optional[bool] = None):
        """target_f32: the float32 target arithmetic of rainbow_nomultisteps.py (default: on with a dueling head, i.e. for rainbow specs)"""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.SrlxError("ImageQNet needs a CUDA device: there is no CPU fallback")
        self.spec, self.batch_size, self.uint8_states = spec, int(batch_size), bool(uint8_states)
        cap = int(batch_cap or batch_size)
        q = _lib.SrlxImageQ()
        q.in_c, q.in_h, q.in_w = spec.in_c, spec.in_h, spec.in_w
        q.in_sb, q.in_sc, q.in_sh, q.in_sw = spec.in_strides
        q.in_u8, q.in_max_val = int(self.uint8_states), float(max_val)
        q.n_conv = len(spec.conv_geo)
        for l, (c, h, w, k, st, p, oh, ow, f, c_fast, off) in enumerate(spec.conv_geo):
            q.conv_f[l], q.conv_k[l], q.conv_s[l], q.conv_p[l], q.conv_oh[l], q.conv_ow[l], q.conv_off[l] = f, k, st, p, oh, ow, off
        q.n_dense = len(spec.dense)
        for l, (out, kk, off) in enumerate(spec.dense):
            q.dense_out[l], q.dense_k[l], q.dense_off[l] = out, kk, off
        q.n_actions, q.n_params, q.batch_cap = spec.n_actions, spec.n_params, cap
        q.enable_double_dqn, q.enable_rescale, q.target_update_interval = int(enable_double_dqn), int(enable_rescale), int(target_model_update_interval)
        q.dueling, q.duel_hidden = spec.duel, spec.duel_hidden
        q.target_f32 = int(spec.duel != _lib.DUEL_NONE if target_f32 is None else target_f32)
        q.discount, q.lr, q.adam_beta1, q.adam_beta2, q.adam_eps = discount, lr, adam_beta1, adam_beta2, adam_eps
        f32 = dict(dtype=torch.float32, device=self.device)
        self.params, self.target = torch.zeros(spec.n_params, **f32), torch.zeros(spec.n_params, **f32)
        self.adam_m, self.adam_v, self.grads = torch.zeros(spec.n_params, **f32), torch.zeros(spec.n_params, **f32), torch.zeros(spec.n_params, **f32)
        self.counters = torch.zeros(4, dtype=torch.int64, device=self.device)
        q.params, q.target, q.adam_m, q.adam_v, q.grads = (t.data_ptr() for t in (self.params, self.target, self.adam_m, self.adam_v, self.grads))
        q.counters = self.counters.data_ptr()
        n_ws = int(self.lib.srlx_imageq_ws_floats(C.byref(q)))
        if n_ws == 0:
            raise _lib.SrlxError(f"srlx_imageq_ws_floats: {self.lib.srlx_last_error().decode()}")
        self.ws = torch.zeros(n_ws, **f32)
        q.ws, q.ws_floats = self.ws.data_ptr(), n_ws
        self.c = q
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_imageq_init(C.byref(q), _stream()))
        self._pri = torch.zeros(cap, **f32)
        self._tq = torch.zeros(cap, **f32)
        self._loss = torch.zeros(1, **f32)
        self.load_state_dict(spec.init_state_dict(seed))

    # ---- parameters ---------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd, target_sd=None):
        flat = torch.from_numpy(self.spec.from_state_dict(sd))
        self.params.copy_(flat)
        self.target.copy_(flat if target_sd is None else torch.from_numpy(self.spec.from_state_dict(target_sd)))

    def state_dict(self, target: bool = False):
        return self.spec.to_state_dict((self.target if target else self.params).cpu().numpy())

    @property
    def train_count(self) -> int:
        return int(self.counters[0].item())

    @property
    def sync_count(self) -> int:
        return int(self.counters[2].item())

    # ---- batches ------------------------------------------------------------------------------------------------------
    def _states(self, x) -> torch.Tensor:
        """uint8 frames (only on a network built with uint8_states=True: its workspace holds the normalised copy) or float32 states;
        sets the struct's in_u8 for the call that follows."""
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x if x.dtype == np.uint8 else x.astype(np.float32, copy=False))
            x = torch.from_numpy(x).to(self.device, non_blocking=True)
        u8 = x.dtype == torch.uint8
        if u8 and not self.uint8_states:
            raise ValueError("uint8 states need a network built with uint8_states=True")
        if not x.is_cuda or x.dtype not in (torch.uint8, torch.float32) or tuple(x.shape[1:]) != self.spec.obs_shape:
            raise ValueError(f"states must be a CUDA uint8 / float32 tensor [n, {self.spec.obs_shape}], got {x.dtype} {tuple(x.shape)} on {x.device}")
        self.c.in_u8 = int(u8)
        return x.contiguous()

    def _vec(self, x, dtype) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype={torch.float32: np.float32, torch.int32: np.int32}[dtype])).to(self.device, non_blocking=True)

    def _forward(self, state, use_target: int) -> torch.Tensor:
        x = self._states(state)
        n, cap = x.shape[0], self.c.batch_cap
        out = torch.empty((n, self.spec.n_actions), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for i in range(0, n, cap):
                m = min(cap, n - i)
                _lib.check(self.lib.srlx_imageq_forward(C.byref(self.c), use_target, x[i:i + m].data_ptr(), m, out[i:i + m].data_ptr(), _stream()))
        return out

    def pred_q(self, state) -> torch.Tensor:
        return self._forward(state, 0)

    def pred_target_q(self, state) -> torch.Tensor:
        return self._forward(state, 1)

    def train(self, state, n_state, action, reward, undone, weights, phases: int = 3):
        """One Trainer.train() on the batch.  Returns (loss [1], priorities [B], target_q [B]) as device tensors (views of buffers that
        the next call overwrites)."""
        s, ns = self._states(state), self._states(n_state)
        if s.dtype != ns.dtype:
            raise ValueError("state and n_state must have the same dtype")
        B = s.shape[0]
        a, r = self._vec(action, torch.int32), self._vec(reward, torch.float32)
        u, w = self._vec(undone, torch.float32), self._vec(weights, torch.float32)
        if not (ns.shape[0] == a.numel() == r.numel() == u.numel() == w.numel() == B):
            raise ValueError("batch arrays disagree in length")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_imageq_train(C.byref(self.c), s.data_ptr(), ns.data_ptr(), a.data_ptr(), r.data_ptr(), u.data_ptr(), w.data_ptr(),
                                                  B, self._pri.data_ptr(), self._loss.data_ptr(), self._tq.data_ptr(), int(phases), _stream()))
        return self._loss, self._pri[:B], self._tq[:B]


    def apply_gradients(self):
        """The optimiser half of an update (phases = 2): Adam on whatever sits in `self.grads`, target sync, counters."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_imageq_train(C.byref(self.c), None, None, None, None, None, None, 1, None, None, None, 2, _stream()))

    def train_data_parallel(self, state, n_state, action, reward, undone, weights, group=None):
        """One Trainer.train() of ONE trainer over the ranks of a torch.distributed group (the reference's distributed mode has a single
        trainer, srl/base/run/play_mp.py:352-462): every rank runs forward / backward on ITS shard of the batch (equal shard sizes), the
        flat gradient is averaged with one NCCL all-reduce over NVLink -- the Huber loss is a mean over the global batch -- and every rank
        applies the same Adam step, so parameters, moments and target network stay identical on all ranks.  Returns this rank's
        (loss of its shard, priorities, target_q)."""
        import torch.distributed as dist

        out = self.train(state, n_state, action, reward, undone, weights, phases=1)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=group)
            self.grads.div_(dist.get_world_size(group))
        self.apply_gradients()
        return out


# =====================================================================================================================
class DemoMixMemory:
    """PriorityReplayBuffer's demo memory (srl/rl/memories/priority_replay_buffer.py:177-240) over any IPriorityMemory (the device
    memories of memory.py): a second, uniform buffer of demonstration batches; every sample() draws batch_size - demo_batch_size items
    from the priority memory and demo_batch_size = max(1, int(batch_size * demo_ratio)) from the demo buffer.

    As in the reference, sample() appends ONE weight of 1.0 however many demo items it drew (np.append(weights, 1.0), :237-239), and
    update() only reaches the priority memory (update_args cover its items only)."""

    def __init__(self, memory, batch_size: int, demo_ratio: float = 1.0 / 256.0, capacity: int = 100_000, warmup_size: int = 1000, seed: Optional[int] = None):
        import random

        self.memory = memory
        self.demo_batch_size = max(1, int(batch_size * demo_ratio))
        self.batch_size = batch_size - self.demo_batch_size
        if not (warmup_size <= capacity):
            raise ValueError(f"assert {warmup_size} <= {capacity}")
        if not (self.batch_size > 0):
            raise ValueError(f"assert {self.batch_size} > 0")
        if not (self.batch_size <= warmup_size):
            raise ValueError(f"assert {self.batch_size} <= {warmup_size}")
        self.capacity, self.warmup_size = capacity, warmup_size
        self.demo: List = []
        self.demo_idx = 0
        self.select_memory = "main"
        self.step = 0
        self._rng = random.Random(seed) if seed is not None else random

    def length(self) -> int:
        return self.memory.length() + len(self.demo)

    def add(self, batch, priority: Optional[float] = None) -> None:
        if self.select_memory == "demo":  # ReplayBuffer.add (replay_buffer.py): ring of `capacity` items
            if len(self.demo) < self.capacity:
                self.demo.append(batch)
            else:
                self.demo[self.demo_idx] = batch
            self.demo_idx = (self.demo_idx + 1) % self.capacity
        else:
            self.memory.add(batch, priority)

    def is_warmup_needed(self) -> bool:
        return self.memory.length() < self.warmup_size

    def sample(self, step: int = -1, batch_size: int = -1):
        if self.memory.length() < self.warmup_size:
            return None
        batch_size = batch_size if batch_size > -1 else self.batch_size
        step = step if step > -1 else self.step
        batches, weights, update_args = self.memory.sample(batch_size, step)
        weights = np.asarray(weights, dtype=np.float32)
        if len(self.demo) < self.demo_batch_size:  # ReplayBuffer.sample below its warmup_size returns None -> batches.extend(None) raises
            raise TypeError("'NoneType' object is not iterable (the demo memory holds fewer than demo_batch_size items)")
        batches = list(batches) + self._rng.sample(self.demo, self.demo_batch_size)
        weights = np.append(weights, 1.0).astype(weights.dtype)
        return batches, weights, update_args

    def update(self, update_args, priorities, step: int = -1) -> None:
        priorities = priorities[: self.batch_size]  # :244-246: the demo items' TD errors are dropped
        self.memory.update(update_args, priorities)
        self.step = step

    def call_backup(self, **kwargs):
        return [self.memory.backup(), [self.demo[:], self.demo_idx, False]]  # :249-250, ReplayBuffer.call_backup

    def call_restore(self, data, **kwargs) -> None:
        self.memory.restore(data[0])
        self.demo, self.demo_idx = list(data[1][0])[-self.capacity:], int(data[1][1]) % self.capacity
