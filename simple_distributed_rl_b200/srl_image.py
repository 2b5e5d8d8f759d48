"""The image configs of the reference through its own plug-in API (needs `srl` importable): the UNMODIFIED `srl.Runner(env,
dqn.Config(...)).train()` with an image observation space runs its processors, network and trainer in libsrlx.so.

    import srl
    from srl.algorithms import dqn
    from simple_distributed_rl_b200 import srl_image
    srl_image.register()                                            # "DQN:torch": the reference's Memory and Worker, the device Parameter / Trainer
    cfg = dqn.Config()                                              # input_block.image = the DQN block (input_block.py:108-121)
    cfg.input_block.image.processors = [srl_image.DeviceImageProcessor(SpaceTypes.GRAY_HW1, (84, 84), normalize_type="0to1")]   # optional
    cfg.window_length = 4
    srl.Runner("ALE/Pong-v5", cfg).train(max_train_count=...)
    srl_image.unregister()

What plugs in where:
  srl/base/rl/processor.py / srl/base/env/processor.py   DeviceImageProcessor: the reference's ImageProcessor dataclass with
                                                          remap_observation (image_processor.py:104-154) on the device; the space
                                                          logic (remap_observation_space, :31-102) is inherited unchanged
  srl/algorithms/dqn/dqn.py:134-173                       ImageParameter(CommonInterfaceParameter): online / target conv Q-network in
                                                          HBM; call_backup / call_restore = the reference state_dict
  srl/algorithms/dqn/model_torch.py:75-131                ImageTrainer.train(): memory.sample() (the reference's own replay, host
                                                          lists) -> one batch upload -> srlx_imageq_train -> memory.update()
Memory and Worker stay the reference's classes (dqn.Memory, dqn.Worker): the worker's policy calls ImageParameter.pred_q.
Not covered (raises): activation other than relu, image blocks other than "DQN" (R2D3 / AlphaZero / MuZero blocks), invalid actions.
"""
from dataclasses import dataclass
from typing import Any, cast

import numpy as np
import torch

from srl.algorithms.dqn.dqn import CommonInterfaceParameter
from srl.base.rl import registration as rl_registration
from srl.base.rl.trainer import RLTrainer
from srl.base.spaces.space import SpaceBase
from srl.rl.processors.image_processor import ImageProcessor

from . import _lib
from .image import DeviceImagePipeline, ImageNetSpec, ImageQNet

_MOD = __name__


def _device_of(config) -> str:
    dev = str(getattr(config, "used_device_torch", "cuda") or "cuda")
    if not torch.cuda.is_available() or dev.startswith("cpu"):
        raise _lib.SrlxError(f"the device classes need a CUDA device (used_device_torch = {dev!r}, cuda available = "
                             f"{torch.cuda.is_available()}): there is no CPU fallback; srl_image.unregister() restores the torch classes")
    return "cuda:0" if dev == "cuda" else dev


@dataclass
class DeviceImageProcessor(ImageProcessor):
    """Drop-in for srl.rl.processors.image_processor.ImageProcessor: same fields, same spaces, pixels on the device."""

    def remap_observation(self, state, prev_space: SpaceBase, new_space: SpaceBase, **kwargs):
        state = np.asarray(state)
        if "float" in str(state.dtype):  # the reference neither converts nor resizes float frames (:126-137): nothing for the device to do
            return super().remap_observation(state, prev_space, new_space, **kwargs)
        pipe = getattr(self, "_pipe", None)
        key = (tuple(state.shape), prev_space.stype)
        if pipe is None or self._pipe_key != key:
            pipe = DeviceImagePipeline(state.shape, prev_space.stype.name, self.image_type.name, self.resize, self.normalize_type, self.trimming,
                                       max_val=float(self.max_val))
            self._pipe, self._pipe_key = pipe, key
        return pipe(state.astype(np.uint8)).cpu().numpy()


def spec_from_config(config) -> ImageNetSpec:
    """dqn.Config (after setup) -> ImageNetSpec; raises for what the device network does not build."""
    obs, act = config.observation_space, config.action_space
    if not obs.is_image_like():
        raise _lib.SrlxError(f"srl_image handles image observation spaces (got {obs}); srl_classes.register() covers value observations")
    img = config.input_block.image
    if img.name != "DQN":
        raise NotImplementedError(f"image block {img.name!r}: only the DQN block is built on the device")
    if str(img.kwargs.get("activation", "relu")).lower() != "relu":
        raise NotImplementedError("image block activation other than relu")
    hk = dict(getattr(config.hidden_block, "kwargs", {}) or {})
    if config.hidden_block.name != "MLP" or str(hk.get("activation", "relu")).lower() != "relu":
        raise NotImplementedError(f"hidden block {config.hidden_block.name!r} / activation {hk.get('activation')!r}")
    return ImageNetSpec(tuple(obs.shape), obs.stype.name, int(act.n), filters=int(img.kwargs.get("filters", 32)), hidden=tuple(hk["layer_sizes"]))


class ImageParameter(CommonInterfaceParameter):
    def setup(self) -> None:
        super().setup()
        cfg = self.config
        self.spec = spec_from_config(cfg)
        self.net = ImageQNet(self.spec, batch_size=cfg.batch_size, enable_double_dqn=cfg.enable_double_dqn, enable_rescale=cfg.enable_rescale,
                             discount=cfg.discount, lr=float(cfg.lr),
                             target_model_update_interval=cfg.target_model_update_interval, device=_device_of(cfg),
                             seed=int(torch.initial_seed() % (2**31)))
        self.np_dtype = cfg.get_dtype("np")

    def call_restore(self, data: Any, from_serialized: bool = False, **kwargs) -> None:
        if from_serialized:
            import pickle

            data = pickle.loads(data)
        self.net.load_state_dict(data)  # model_torch.py:49-51: online and target both take the restored weights

    def call_backup(self, serialized: bool = False, **kwargs) -> Any:
        sd = self.net.state_dict()
        if serialized:
            import pickle

            return pickle.dumps(sd)
        return sd

    def summary(self, **kwargs):
        print(f"ImageQNet on {self.net.device}: {self.spec}")

    def pred_q(self, state) -> np.ndarray:
        return self.net.pred_q(np.asarray(state, dtype=np.float32)).cpu().numpy().astype(self.np_dtype, copy=False)

    def pred_target_q(self, state) -> np.ndarray:
        return self.net.pred_target_q(np.asarray(state, dtype=np.float32)).cpu().numpy().astype(self.np_dtype, copy=False)


class ImageTrainer(RLTrainer):
    def on_setup(self) -> None:
        self.net = cast(ImageParameter, self.parameter).net
        self.sync_count = 0

    def train(self) -> None:
        batches = self.memory.sample()
        if batches is None:
            return
        batches, weights, update_args = batches
        state, n_state, onehot_action, reward, undone, next_invalid_actions = zip(*batches)
        if any(len(v) > 0 for v in next_invalid_actions):
            raise NotImplementedError("invalid actions with an image observation space")
        action = np.argmax(np.asarray(onehot_action, dtype=np.float32), axis=1).astype(np.int32)
        loss, pri, _ = self.net.train(np.asarray(state, dtype=np.float32), np.asarray(n_state, dtype=np.float32), action,
                                      np.asarray(reward, dtype=np.float32), np.asarray(undone, dtype=np.float32),
                                      np.asarray(weights, dtype=np.float32))
        host = torch.cat([loss, pri]).cpu().numpy()  # one read: the reference reads loss.item() and the priorities here too
        self.info["loss"] = float(host[0])
        self.memory.update(update_args, host[1:], self.train_count)
        if self.train_count % self.config.target_model_update_interval == 0:
            self.sync_count += 1  # the device synced inside the update (model_torch.py:124-127)
        self.info["sync"] = self.sync_count
        self.train_count += 1


# ---------------------------------------------------------------------------------------------------------------------
_saved = {}


def register() -> None:
    """Take over "DQN:torch" with the reference's Memory / Worker and the device Parameter / Trainer."""
    from srl.algorithms import dqn

    reg = rl_registration._registry
    key = rl_registration._create_registry_key(dqn.Config().set_torch())
    if key not in _saved:
        _saved[key] = list(reg[key])
    mem_ep, _, _, worker_ep = _saved[key]
    reg[key] = [mem_ep, f"{_MOD}:ImageParameter", f"{_MOD}:ImageTrainer", worker_ep]


def unregister() -> None:
    reg = rl_registration._registry
    for k, v in _saved.items():
        reg[k] = list(v)
    _saved.clear()
