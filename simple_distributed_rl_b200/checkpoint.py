"""Checkpoint / wire compatibility of the device state with reference-trained artefacts.

File format        `save_file` / `load_file` restate srl/utils/common.py:117-152: a pickle, LZMA-compressed (xz container, magic
                   fd 37 7a 58 5a 00) when compress=True; load sniffs the magic.
Parameter          `RLParameter.save` = save_file(path, call_backup()) (srl/base/rl/parameter.py:38-51); for the torch DQN / Rainbow
                   parameters call_backup() is `q_online.state_dict()` (srl/algorithms/dqn/model_torch.py:47-52,
                   srl/rl/torch_/helper.py:76-93): an OrderedDict of torch tensors under the reference's module keys
                   (netspec.NetSpec.to_state_dict / from_state_dict hold the key map).
Memory             `RLMemory.save` = save_file(path, call_backup()) (srl/base/rl/memory.py:119-150) with
                   call_backup() = [memory.backup(), demo_memory backup or None] (srl/rl/memories/priority_replay_buffer.py:252-258)
                     ReplayBuffer.backup()        = [memory list, idx]                         (priority_memories/replay_buffer.py:40-55)
                     ProportionalMemory.backup()  = [capacity, max_priority, size, write, tree list, data list]
                                                                                              (proportional_memory.py:179-205)
                   Each item is the record the worker handed to memory.add(), zlib(pickle(.)) when memory.compress
                   (priority_replay_buffer.py:205-217):
                     DQN / Rainbow multisteps=1   [state, next_state, onehot_action, reward, int(not terminated), next_invalid_actions]
                                                  (srl/algorithms/dqn/dqn.py:226-246, rainbow/rainbow_nomultisteps.py:91-108)
                     Rainbow multisteps=M>1       M+1 entries [state_k, onehot(a_{k-1}), r_{k-1}, terminated_{k-1}, next_invalid]
                                                  (rainbow.py:345-375; the trainer reads entry 0's state only, :191-199)

The device ring stores ONE record per env step, time-major (slot = row * E + env), and rebuilds n-step windows by index.  A
reference memory is a flat list of self-contained items.  The mapping used here:

  export  items are emitted env-major -- env 0's windows oldest to newest, then env 1's, ... -- so the item stream is E
          concatenated trajectories.  Padded window tails are rebuilt exactly as the sampler rebuilds them (repeated last
          state, reward 0, terminated 1, random action).  For the proportional memory a fresh tree is built over the emitted
          items' leaf priorities (leaf i = item i; every inner node = left + right).
  import  the item stream (oldest first) is cut into E equal chunks; chunk c becomes column c of the ring: item j's FIRST
          transition is ring row j, and the last item's window supplies the M-1 rows after it, so export -> import is the
          identity on everything the sampler can reach.  A step is an episode end when it terminated or when the next window
          entry is padding; the last row of every column is marked done (a window never runs across the seam into steps
          taken after the import).  Leaf priorities follow their items.

Everything here is host-side numpy on plain arrays (no torch, no CUDA), so the CPU tests drive it against the imported reference.
"""
import binascii
import lzma
import pickle
import zlib
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

_XZ_MAGIC = b"fd377a585a00"


# ---- files (srl/utils/common.py:117-152) -------------------------------------------------------------------------------
def save_file(path: str, dat: Any, compress: bool = True) -> None:
    import os

    try:
        if compress:
            with lzma.open(path, "w") as f:
                f.write(pickle.dumps(dat))
        else:
            with open(path, "wb") as f:
                pickle.dump(dat, f)
    except Exception:
        if os.path.isfile(path):
            os.remove(path)
        raise


def load_file(path: str) -> Any:
    with open(path, "rb") as f:
        is_xz = binascii.hexlify(f.read(6)) == _XZ_MAGIC
    if is_xz:
        with lzma.open(path) as f:
            return pickle.loads(f.read())
    with open(path, "rb") as f:
        return pickle.load(f)


# ---- parameters -------------------------------------------------------------------------------------------------------
def parameter_backup(spec, mu: np.ndarray, sigma: Optional[np.ndarray]) -> "OrderedDict":
    """RLParameter.call_backup() of the torch parameter: state_dict of the online network (torch tensors, module key order)."""
    return OrderedDict(spec.to_state_dict(mu, sigma))


def parameter_restore(spec, data: Dict[str, Any]) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    import torch

    return spec.from_state_dict({k: (v if hasattr(v, "detach") else torch.as_tensor(np.asarray(v))) for k, v in data.items()})


# ---- ring arrays -------------------------------------------------------------------------------------------------------
class RingView:
    """The host copy of what the converters need: plain numpy arrays + the few scalars that define validity."""

    def __init__(self, n_envs: int, ring_rows: int, multisteps: int, n_actions: int, obs_dim: int, vec_steps: int = 0):
        self.E, self.R, self.M, self.A, self.D = int(n_envs), int(ring_rows), int(multisteps), int(n_actions), int(obs_dim)
        cap = self.E * self.R
        self.vec_steps = int(vec_steps)
        self.obs = np.zeros((cap, self.D), dtype=np.float32)
        self.next_obs = np.zeros((cap, self.D), dtype=np.float32)
        self.action = np.zeros(cap, dtype=np.int32)
        self.reward = np.zeros(cap, dtype=np.float32)
        self.term = np.zeros(cap, dtype=np.uint8)
        self.done = np.zeros(cap, dtype=np.uint8)
        self.invalid: Optional[np.ndarray] = None  # [cap] uint32 bit masks: invalid actions of the row's NEXT state (None: no masks)
        self.leaf_priority: Optional[np.ndarray] = None  # [cap] float64 (proportional memory), by slot
        self.max_priority = 1.0

    @property
    def capacity(self) -> int:
        return self.E * self.R

    def valid_rows(self) -> Tuple[int, int]:
        """(first vector step, count) of the rows whose M-step window is complete and still in the ring
        (oracle/engine.py::_valid_range, csrc learner: the uniform sampler's range)."""
        g_lo = max(0, self.vec_steps - self.R)
        return g_lo, max(0, self.vec_steps - self.M + 1 - g_lo)


def philox_pad_action(seed: int, env: int, vec_step: int, n_actions: int) -> int:
    """The action the device's sampler puts on a padded window step: Philox4x32-10(counter (env, step_lo, step_hi,
    STREAM_PAD_ACTION = 6), key = seed), word 0 mapped to [0, n_actions) by multiply-shift (csrc/philox.cuh, learner kernels)."""
    m32 = 0xFFFFFFFF
    c = [env & m32, vec_step & m32, (vec_step >> 32) & m32, 6]
    k0, k1 = seed & m32, (seed >> 32) & m32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & m32, p1 & m32, ((p0 >> 32) ^ c[3] ^ k1) & m32, p0 & m32]
        k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
    return (c[0] * n_actions) >> 32


def _onehot(a: int, n: int) -> List[int]:
    v = [0] * n  # DiscreteSpace.get_onehot (srl/base/spaces/discrete.py:116-121): a python list of ints
    v[int(a)] = 1
    return v


def _pack(item: Any, compress: bool, level: int = -1) -> Any:
    return zlib.compress(pickle.dumps(item), level=level) if compress else item


def _unpack(item: Any) -> Any:
    return pickle.loads(zlib.decompress(item)) if isinstance(item, (bytes, bytearray)) else item


def export_items(ring: RingView, pad_action=None, compress: bool = False) -> Tuple[List[Any], Optional[np.ndarray]]:
    """Ring -> reference items, env-major; returns (items, leaf priorities of the items or None).
    pad_action(env, vector_step) -> int supplies the random action of a padded tail step (the device draws it from its own
    Philox stream, learner kernels / oracle.engine.window); None -> action 0 (the trainer multiplies it with Q of a
    terminated step's successor, whose gain is masked by terminated=1, rainbow.py:257)."""
    E, R, M, A = ring.E, ring.R, ring.M, ring.A
    g_lo, n_g = ring.valid_rows()
    inv = (lambda sl: [a for a in range(A) if (int(ring.invalid[sl]) >> a) & 1]) if ring.invalid is not None else (lambda sl: [])
    items: List[Any] = []
    pri = [] if ring.leaf_priority is not None else None
    for e in range(E):
        for g in range(g_lo, g_lo + n_g):
            slot = (g % R) * E + e
            if M == 1:
                item = [ring.obs[slot].copy(), ring.next_obs[slot].copy(), _onehot(ring.action[slot], A), float(ring.reward[slot]),
                        int(not ring.term[slot]), inv(slot)]
            else:
                item = [[ring.obs[slot].copy(), None, None, None, None]]
                ended, last_state = False, None
                for k in range(M):
                    if not ended:
                        sk = ((g + k) % R) * E + e
                        last_state = ring.next_obs[sk].copy()
                        item.append([last_state, _onehot(ring.action[sk], A), float(ring.reward[sk]), int(ring.term[sk]), inv(sk)])
                        ended = bool(ring.done[sk])
                    else:
                        a = int(pad_action(e, g + k)) if pad_action is not None else 0
                        item.append([last_state.copy(), _onehot(a, A), 0, 1, []])
            items.append(_pack(item, compress))
            if pri is not None:
                pri.append(float(ring.leaf_priority[slot]))
    return items, (np.asarray(pri, dtype=np.float64) if pri is not None else None)


def build_sum_tree(leaves: np.ndarray, capacity: int) -> np.ndarray:
    """Flat SumTree (2*capacity-1 doubles, leaf j at j+capacity-1, children 2i+1 / 2i+2; proportional_memory.py:13-47) whose
    inner nodes are exactly left + right."""
    tree = np.zeros(2 * capacity - 1, dtype=np.float64)
    tree[capacity - 1: capacity - 1 + len(leaves)] = leaves
    n_inner = capacity - 1  # inner nodes are [0, capacity-1); heap depth d holds nodes [2^d - 1, 2^(d+1) - 1)
    d = max(n_inner, 1).bit_length()
    while d >= 0:  # deepest level first: a node's children (one level down) are final when it is summed
        p = np.arange((1 << d) - 1, min((1 << (d + 1)) - 1, n_inner))
        if len(p):
            tree[p] = tree[2 * p + 1] + tree[2 * p + 2]
        d -= 1
    return tree


def memory_backup(ring: RingView, proportional: bool, pad_action=None, compress: bool = False) -> list:
    """RLPriorityReplayBuffer.call_backup() for the ring: [inner memory backup, None (no demo memory)]."""
    items, pri = export_items(ring, pad_action, compress)
    cap, n = ring.capacity, len(items)
    if not proportional:
        return [[items, n if n < cap else 0], None]
    leaves = np.zeros(cap, dtype=np.float64)
    leaves[:n] = pri
    data: List[Any] = list(items) + [None] * (cap - n)
    return [[cap, float(ring.max_priority), n, n % cap, build_sum_tree(leaves, cap).tolist(), data], None]


def _first_transition(item: Any, M: int):
    """(state, next_state, action, reward, terminated, done, tail) of the first step of a reference item; tail = the later
    (state', action, reward, terminated, done) steps a multistep window carries."""
    it = _unpack(item)
    bits = lambda lst: sum(1 << int(a) for a in (lst or []))  # noqa: E731
    if M == 1:
        s, ns, oh, r, undone = it[0], it[1], it[2], it[3], it[4]
        term = 0 if undone else 1
        return (np.asarray(s, np.float32), np.asarray(ns, np.float32), int(np.argmax(oh)), float(r), term, term, [],
                bits(it[5] if len(it) > 5 else None))
    steps = []
    for k in range(1, M + 1):
        ns, oh, r, t = it[k][0], it[k][1], it[k][2], it[k][3]
        steps.append([np.asarray(ns, np.float32), int(np.argmax(oh)), float(r), int(t), int(t), bits(it[k][4] if len(it[k]) > 4 else None)])
    for k in range(M - 1):  # step k ended its episode when step k+1 is padding: same state, reward 0, terminated 1
        nxt = steps[k + 1]
        if nxt[3] == 1 and nxt[2] == 0.0 and np.array_equal(nxt[0], steps[k][0]):
            steps[k][4] = 1
    real, ended = [], False
    for st in steps:  # the steps after an episode end are padding, not data
        if ended:
            break
        real.append(st)
        ended = bool(st[4])
    f = real[0]
    return np.asarray(it[0][0], np.float32), f[0], f[1], f[2], f[3], f[4], real[1:], f[5]


def _ordered_items(inner: list, proportional: bool) -> Tuple[List[Any], Optional[np.ndarray], float]:
    """Items of a reference memory backup oldest first (+ their leaf priorities, max_priority)."""
    if not proportional:
        mem, idx = list(inner[0]), int(inner[1])
        return (mem[idx:] + mem[:idx]) if idx < len(mem) else mem, None, 1.0
    cap, maxp, size, write, tree, data = int(inner[0]), float(inner[1]), int(inner[2]), int(inner[3]), inner[4], inner[5]
    order = list(range(write, cap)) + list(range(write)) if size >= cap else list(range(size))
    order = [i for i in order if data[i] is not None]
    pri = np.asarray([tree[i + cap - 1] for i in order], dtype=np.float64)
    return [data[i] for i in order], pri, maxp


class DiscontinuousMemoryError(ValueError):
    """A multistep memory whose item stream is not one contiguous trajectory per ring column."""


def memory_restore(data: list, n_envs: int, ring_rows: int, multisteps: int, n_actions: int, obs_dim: int, proportional: bool,
                   check_continuity: bool = True) -> RingView:
    """RLPriorityReplayBuffer.call_restore() into ring arrays (see the module docstring for the layout).

    With multisteps > 1 the ring keeps only the FIRST transition of every item and rebuilds the M-step window from the rows that
    follow, so within a column item j+1 must be the next step of item j's trajectory (or item j's first step ended its episode).
    A memory filled by several actors (play_mp interleaves their items) or any other non-contiguous stream does not have that
    property; importing it would stitch windows from unrelated trajectories, so it raises DiscontinuousMemoryError instead
    (check_continuity=False skips the check for a caller that knows better)."""
    inner = data[0] if (len(data) == 2 and isinstance(data[0], list) and (data[1] is None or isinstance(data[1], list))) else data
    src_prop = len(inner) == 6 and not isinstance(inner[0], list)
    items, pri, maxp = _ordered_items(inner, src_prop)
    E, R, M = int(n_envs), int(ring_rows), int(multisteps)
    n_g = min(len(items) // E, R - (M - 1))
    ring = RingView(E, R, M, n_actions, obs_dim)
    ring.max_priority = maxp
    if proportional:
        ring.leaf_priority = np.zeros(ring.capacity, dtype=np.float64)
    if n_g <= 0:
        return ring
    items = items[len(items) - n_g * E:]  # the newest n_g * E items
    if pri is not None:
        pri = pri[len(pri) - n_g * E:]
    for e in range(E):
        tail = []
        prev_ns, prev_done = None, True
        for j in range(n_g):
            s, ns, a, r, term, done, tail, mask = _first_transition(items[e * n_g + j], M)
            if mask:
                if ring.invalid is None:
                    ring.invalid = np.zeros(ring.capacity, dtype=np.uint32)
                ring.invalid[j * E + e] = mask
            if check_continuity and M > 1 and not prev_done and not np.array_equal(s, prev_ns):
                raise DiscontinuousMemoryError(
                    f"memory_restore: item {e * n_g + j} does not continue the trajectory of the item before it (its state is not the "
                    f"previous item's next state and that step did not end an episode). With multisteps = {M} the ring rebuilds "
                    "windows from consecutive rows, so the item stream must be one contiguous trajectory per column -- a memory "
                    "filled by interleaved actors (play_mp) cannot be imported this way")
            prev_ns, prev_done = ns, bool(done)
            slot = j * E + e
            ring.obs[slot], ring.next_obs[slot] = s, ns
            ring.action[slot], ring.reward[slot], ring.term[slot], ring.done[slot] = a, r, term, done
            if proportional:
                ring.leaf_priority[slot] = pri[e * n_g + j] if pri is not None else maxp
        prev_next = ring.next_obs[(n_g - 1) * E + e]
        for k in range(M - 1):  # rows after the last window start: real steps carried by the last item, else a cut
            slot = (n_g + k) * E + e
            if k < len(tail):
                ns, a, r, term, done, mask = tail[k]
                if mask:
                    if ring.invalid is None:
                        ring.invalid = np.zeros(ring.capacity, dtype=np.uint32)
                    ring.invalid[slot] = mask
                ring.obs[slot], ring.next_obs[slot] = prev_next, ns
                ring.action[slot], ring.reward[slot], ring.term[slot], ring.done[slot] = a, r, term, done
                prev_next = ns
            else:  # the episode ended inside the last window: the row is never reached (the window is cut before it)
                ring.obs[slot], ring.next_obs[slot] = prev_next, prev_next
                ring.action[slot], ring.reward[slot], ring.term[slot], ring.done[slot] = 0, 0.0, 1, 1
        ring.done[(n_g + M - 2) * E + e] = 1  # seam: nothing recorded after the import continues this trajectory
    ring.vec_steps = n_g + M - 1
    return ring
