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
Rainbow with multisteps = 1 ("Rainbow_no_multisteps:torch": the dueling head of srl/rl/torch_/blocks/dueling_network.py, average / max / naive,
float32 targets of rainbow_nomultisteps.py) runs through the same classes.
Not covered (raises): activation other than relu, image blocks other than "DQN" (R2D3 / AlphaZero / MuZero blocks), invalid actions,
NoisyNet and n-step Retrace targets over image states.
"""
from dataclasses import dataclass
from typing import Any, cast

import numpy as np
import torch

from srl.algorithms.dqn.dqn import CommonInterfaceParameter
from srl.base.rl import registration as rl_registration
from srl.base.rl.memory import RLMemory
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

    def __getstate__(self):
        # the reference deep-copies / pickles configs (RLConfig.copy, train_mp): the cached device pipeline (CUDA tensors, a ctypes handle)
        # stays behind and is rebuilt on first use
        d = dict(self.__dict__)
        d.pop("_pipe", None)
        d.pop("_pipe_key", None)
        return d

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
    rainbow = hasattr(config, "multisteps")
    if rainbow and (int(config.multisteps) != 1 or bool(getattr(config, "enable_noisy_dense", False))):
        raise NotImplementedError("rainbow over image states: multisteps = 1 without NoisyNet is built on the device (Rainbow_no_multisteps)")
    filters = int(img.kwargs.get("filters", 32))
    if config.hidden_block.name == "MLP" and not rainbow:
        if str(hk.get("activation", "relu")).lower() != "relu":
            raise NotImplementedError(f"hidden block activation {hk.get('activation')!r}")
        return ImageNetSpec(tuple(obs.shape), obs.stype.name, int(act.n), filters=filters, hidden=tuple(hk["layer_sizes"]))
    if config.hidden_block.name == "DuelingNetwork" and rainbow:
        acts = (hk.get("mlp_kwargs", {}).get("activation", "relu"), hk.get("dueling_kwargs", {}).get("activation", "relu"))
        if any(str(a).lower() != "relu" for a in acts):
            raise NotImplementedError(f"dueling network activations {acts!r}")
        return ImageNetSpec(tuple(obs.shape), obs.stype.name, int(act.n), filters=filters, hidden=tuple(hk["layer_sizes"]),
                            dueling=hk.get("dueling_kwargs", {}).get("dueling_type", "average"))
    raise NotImplementedError(f"hidden block {config.hidden_block.name!r} for {'rainbow' if rainbow else 'dqn'} over image states "
                              "(dqn: MLP; rainbow: DuelingNetwork)")


class _ImageParameterMixin:
    """setup / backup / restore / pred_q over ImageQNet, shared by the dqn and the rainbow parameter classes"""

    def setup(self) -> None:
        super().setup()
        cfg = self.config
        self.spec = spec_from_config(cfg)
        self.net = ImageQNet(self.spec, batch_size=cfg.batch_size, enable_double_dqn=cfg.enable_double_dqn, enable_rescale=cfg.enable_rescale,
                             discount=cfg.discount, lr=float(cfg.lr),
                             target_model_update_interval=cfg.target_model_update_interval, device=_device_of(cfg), uint8_states=True,
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


class ImageParameter(_ImageParameterMixin, CommonInterfaceParameter):
    """srl/algorithms/dqn/model_torch.py:34-72"""


def _rainbow_parameter_class():
    from srl.algorithms.rainbow.rainbow import CommonInterfaceParameter as RainbowParameterBase

    class ImageRainbowParameter(_ImageParameterMixin, RainbowParameterBase):
        """srl/algorithms/rainbow/model_torch.py:32-72 (multisteps = 1: rainbow_nomultisteps.py)"""

    return ImageRainbowParameter


ImageRainbowParameter = _rainbow_parameter_class()


class DeviceImageMemory(RLMemory):
    """RLPriorityReplayBuffer (srl/rl/memories/priority_replay_buffer.py:177-274) for image states, resident in HBM as BYTES.

    The reference keeps every item as two float32 stacks in host memory (2 x 113 KB at 84 x 84 x 4; optionally zlib-compressed) and the
    trainer re-uploads 7 MB per update.  Here an item is what the worker hands to `add` -- [state, n_state, onehot action, reward, undone,
    next_invalid_actions] (dqn.py:229-246) -- with the two states stored as the uint8 frames they came from (a "0to1" state is k / 255
    exactly; `add` checks it): 56 KB per item, 100 k items (the reference's default capacity) = 5.6 GB, 2 M items = 113 GB of the 180.
    `sample()` gathers the batch on the device and hands device tensors to ImageTrainer; nothing but the 56 KB of a new item crosses
    PCIe.  Memory kinds: "ReplayBuffer" (random.sample over the filled slots, replay_buffer.py:92-101) and "Proportional"
    (memory.DeviceProportionalMemory: the device SumTree; the payload is the slot)."""

    STAGE = 32

    def setup(self) -> None:
        import random

        from .memory import DeviceProportionalMemory

        cfg, mem = self.config, self.config.memory
        self.device = torch.device(_device_of(cfg))
        self.capacity, self.warmup_size, self.batch_size = int(mem.capacity), int(mem.warmup_size), int(cfg.batch_size)
        if not (self.warmup_size <= self.capacity and 0 < self.batch_size <= self.warmup_size):
            raise ValueError(f"assert 0 < batch_size ({self.batch_size}) <= warmup_size ({self.warmup_size}) <= capacity ({self.capacity})")
        if getattr(mem, "enable_demo_memory", False):
            raise NotImplementedError("demo memory over DeviceImageMemory: wrap the device memory in image.DemoMixMemory")
        obs = tuple(cfg.observation_space.shape)
        self.obs_shape = obs
        name = str(getattr(mem, "name", "ReplayBuffer"))
        mk = dict(getattr(mem, "kwargs", {}) or {})
        if name == "ReplayBuffer":
            self.per = None
        elif name in ("Proportional", "Proportional_cpp"):
            self.per = DeviceProportionalMemory(self.capacity, alpha=mk.get("alpha", 0.6), beta_initial=mk.get("beta_initial", 0.4),
                                                beta_steps=mk.get("beta_steps", 1_000_000), has_duplicate=mk.get("has_duplicate", True),
                                                epsilon=mk.get("epsilon", 0.0001), device=str(self.device))
        else:
            raise NotImplementedError(f"memory {name!r} over image states (ReplayBuffer, Proportional)")
        u8 = dict(dtype=torch.uint8, device=self.device)
        self.S, self.NS = torch.zeros((self.capacity,) + obs, **u8), torch.zeros((self.capacity,) + obs, **u8)
        self.act = torch.zeros(self.capacity, dtype=torch.int32, device=self.device)
        self.rew = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        self.undone = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        n = int(np.prod(obs))
        self._st_s = torch.zeros((self.STAGE, 2, n), dtype=torch.uint8).pin_memory()
        self._st_v = torch.zeros((self.STAGE, 3), dtype=torch.float32).pin_memory()
        self._n_stage, self._count, self.step = 0, 0, 0
        self._rng = random
        self.register_worker_func_custom(self.add, self.serialize)
        self.register_trainer_recv_func(self.sample)
        self.register_trainer_send_func(self.update)

    @staticmethod
    def _to_u8(x) -> np.ndarray:
        x = np.asarray(x, dtype=np.float32)
        b = np.rint(x * np.float32(255)).astype(np.uint8)
        back = b.astype(np.float32)
        back /= np.uint8(255)
        if not np.array_equal(back, x):
            raise NotImplementedError("DeviceImageMemory stores '0to1'-normalised uint8 frames; this state is not k / 255 "
                                      "(use the reference's Memory: srl_image.register(device_memory=False))")
        return b.reshape(-1)

    def add(self, batch: Any, priority=None, serialized: bool = False) -> None:
        if serialized:
            import pickle

            batch = pickle.loads(batch)
        state, n_state, onehot_action, reward, undone, next_invalid_actions = batch
        if len(next_invalid_actions) > 0:
            raise NotImplementedError("invalid actions with an image observation space")
        i = self._n_stage
        self._st_s[i, 0] = torch.from_numpy(self._to_u8(state))
        self._st_s[i, 1] = torch.from_numpy(self._to_u8(n_state))
        self._st_v[i, 0], self._st_v[i, 1], self._st_v[i, 2] = float(np.argmax(onehot_action)), float(reward), float(undone)
        self._n_stage += 1
        if self.per is not None:
            self.per.add(self._count % self.capacity, priority)  # payload = slot; the SumTree's write cursor runs in step with ours
        self._count += 1
        if self._n_stage == self.STAGE:
            self.flush()

    def serialize(self, batch: Any, priority=None) -> Any:
        import pickle

        return (pickle.dumps(batch), priority)

    def flush(self) -> None:
        n = self._n_stage
        if n == 0:
            return
        first = self._count - n
        slots = torch.arange(first, first + n, dtype=torch.int64) % self.capacity
        s = self._st_s[:n].to(self.device, non_blocking=True)
        v = self._st_v[:n].to(self.device, non_blocking=True)
        slots = slots.to(self.device, non_blocking=True)
        self.S.index_copy_(0, slots, s[:, 0].reshape((n,) + self.obs_shape))
        self.NS.index_copy_(0, slots, s[:, 1].reshape((n,) + self.obs_shape))
        self.act.index_copy_(0, slots, v[:, 0].to(torch.int32))
        self.rew.index_copy_(0, slots, v[:, 1])
        self.undone.index_copy_(0, slots, v[:, 2])
        torch.cuda.current_stream(self.device).synchronize()  # the pinned staging rows are reused by the next add
        self._n_stage = 0

    def length(self) -> int:
        return min(self._count, self.capacity)

    def is_warmup_needed(self) -> bool:
        return self.length() < self.warmup_size

    def sample(self, step: int = -1, batch_size: int = -1):
        if self.length() < self.warmup_size:
            return None
        self.flush()
        batch_size = batch_size if batch_size > -1 else self.batch_size
        step = step if step > -1 else self.step
        if self.per is None:
            slots = self._rng.sample(range(self.length()), batch_size)
            weights, update_args = np.ones(batch_size, np.float32), []
        else:
            slots, weights, update_args = self.per.sample(batch_size, step)
            weights = np.asarray(weights, dtype=np.float32)
        idx = torch.as_tensor(np.asarray(slots, dtype=np.int64)).to(self.device, non_blocking=True)
        batch = dict(state=self.S.index_select(0, idx), n_state=self.NS.index_select(0, idx), action=self.act.index_select(0, idx),
                     reward=self.rew.index_select(0, idx), undone=self.undone.index_select(0, idx))
        return batch, weights, update_args

    def update(self, update_args, priorities, step: int) -> None:
        if self.per is not None:
            self.per.update(update_args, np.asarray(priorities))
        self.step = step

    # the reference's formats (priority_replay_buffer.py:252-258; replay_buffer.py:103-127): items as the worker's float32 lists
    def _item(self, slot: int):
        A = int(self.config.action_space.n)
        f = lambda t: (t[slot].cpu().numpy().astype(np.float32) / np.uint8(255)).astype(np.float32)  # noqa: E731
        return [f(self.S), f(self.NS), np.eye(A, dtype=np.float32)[int(self.act[slot])].tolist(), float(self.rew[slot]), int(self.undone[slot]), []]

    def call_backup(self, **kwargs) -> Any:
        self.flush()
        n = self.length()
        items = [self._item(i) for i in range(n)]
        if self.per is None:  # ReplayBuffer.call_backup: [buffer, idx, compress]
            return [items, self._count % self.capacity if n == self.capacity else n, False]
        data = self.per.backup()  # [capacity, max_priority, size, write, tree, payloads]: the payloads become the worker's items
        data[5] = items + [None] * (self.capacity - n)
        return [data, None]

    @staticmethod
    def _decode(item):
        """an item as the reference stores it with memory.compress = True (zlib over pickle, priority_replay_buffer.py:205-214)"""
        if isinstance(item, (bytes, bytearray)):
            import pickle
            import zlib

            return pickle.loads(zlib.decompress(item))
        return item

    def call_restore(self, data: Any, **kwargs) -> None:
        if self.per is None:
            items, write = [self._decode(it) for it in list(data[0])[-self.capacity:]], int(data[1])
        else:
            pd = list(data[0])
            if int(pd[0]) != self.capacity:
                raise NotImplementedError(f"restoring a proportional image memory of capacity {pd[0]} into one of {self.capacity}")
            items, write = [self._decode(it) for it in pd[5][:int(pd[2])]], int(pd[3])
        self._n_stage, self._count = 0, 0
        saved, self.per = self.per, None  # refill the rings in slot order without touching the tree
        try:
            for it in items:
                self.add(it)
            self.flush()
        finally:
            self.per = saved
        n = len(items)
        self._count = n if n < self.capacity else self.capacity + write % self.capacity
        if self.per is not None:
            pd[5] = list(range(self.capacity))  # payload = slot
            self.per.restore(pd)


class ImageTrainer(RLTrainer):
    def on_setup(self) -> None:
        self.net = cast(ImageParameter, self.parameter).net
        self.sync_count = 0

    def _finish(self, loss, pri, update_args) -> None:
        host = torch.cat([loss, pri]).cpu().numpy()  # one read: the reference reads loss.item() and the priorities here too
        self.info["loss"] = float(host[0])
        self.memory.update(update_args, host[1:], self.train_count)
        if self.train_count % self.config.target_model_update_interval == 0:
            self.sync_count += 1  # the device synced inside the update (model_torch.py:124-127)
        self.info["sync"] = self.sync_count
        self.train_count += 1

    def train(self) -> None:
        batches = self.memory.sample()
        if batches is None:
            return
        batches, weights, update_args = batches
        if isinstance(batches, dict):  # DeviceImageMemory: uint8 states and the scalars are already in HBM
            loss, pri, _ = self.net.train(batches["state"], batches["n_state"], batches["action"], batches["reward"], batches["undone"],
                                          np.asarray(weights, dtype=np.float32))
            return self._finish(loss, pri, update_args)
        state, n_state, onehot_action, reward, undone, next_invalid_actions = zip(*batches)
        if any(len(v) > 0 for v in next_invalid_actions):
            raise NotImplementedError("invalid actions with an image observation space")
        action = np.argmax(np.asarray(onehot_action, dtype=np.float32), axis=1).astype(np.int32)
        loss, pri, _ = self.net.train(np.asarray(state, dtype=np.float32), np.asarray(n_state, dtype=np.float32), action,
                                      np.asarray(reward, dtype=np.float32), np.asarray(undone, dtype=np.float32),
                                      np.asarray(weights, dtype=np.float32))
        self._finish(loss, pri, update_args)


# ---------------------------------------------------------------------------------------------------------------------
_saved = {}


def register(device_memory: bool = False) -> None:
    """Take over "DQN:torch" and "Rainbow_no_multisteps:torch": the reference's Worker, the device Parameter / Trainer, and either the reference's Memory (host lists,
    the batch uploaded per update) or DeviceImageMemory (uint8 frames resident in HBM, device_memory=True)."""
    from srl.algorithms import dqn

    from srl.algorithms import rainbow

    reg = rl_registration._registry
    for cfg, par in ((dqn.Config().set_torch(), "ImageParameter"), (rainbow.Config(multisteps=1).set_torch(), "ImageRainbowParameter")):
        key = rl_registration._create_registry_key(cfg)
        if key not in _saved:
            _saved[key] = list(reg[key])
        mem_ep, _, _, worker_ep = _saved[key]
        reg[key] = [f"{_MOD}:DeviceImageMemory" if device_memory else mem_ep, f"{_MOD}:{par}", f"{_MOD}:ImageTrainer", worker_ep]


def unregister() -> None:
    reg = rl_registration._registry
    for k, v in _saved.items():
        reg[k] = list(v)
    _saved.clear()
