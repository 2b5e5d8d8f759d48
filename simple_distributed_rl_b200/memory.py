"""DeviceProportionalMemory: the reference's IPriorityMemory interface over the device SumTree.

Drop-in for srl/rl/memories/priority_memories/proportional_memory.py::ProportionalMemory and its pybind11 twin
(cpp_module/src/proportional_memory.cpp:87-275): same constructor arguments, same methods
(clear / length / add / sample / update / backup / restore), `sample` returns (batches, weights, tree indices) and the
indices are handed back verbatim to `update`; `backup()` returns the same 6-element list
[capacity, max_priority, size, write, tree[:], data[:]] so memories interchange with the reference.
Use it through `PriorityReplayBufferConfig.set_custom("simple_distributed_rl_b200.memory:DeviceProportionalMemory", {...})`
(srl/rl/memories/priority_replay_buffer.py:111-117,149-152).

The python payloads stay in a host list (the reference's C++ module also keeps py::object payloads); only the priority
arithmetic lives on the GPU.  This seam exists for parity (the reference's own KATs run against it); the fast path is
the fused engine (engine.py), where payloads never leave HBM.
"""
import ctypes as C
from typing import Any, List, Optional

import numpy as np
import torch

from . import _lib


class DeviceProportionalMemory:
    """One kernel launch per `sample()`: `add` / `update` only append to an op list in mapped pinned host memory; the next `sample`
    (or anything that reads the tree) launches `srlx_tree_seam`, which applies the list in program order with the reference's
    sequential association, draws the batch, and writes indices, weights and a sequence word straight back into mapped host memory,
    where the host polls the word -- no stream synchronisation, no staging copies, no per-call allocations."""

    MAX_OPS = 4096
    MAX_BATCH = 1024

    def __init__(self, capacity: int, alpha: float = 0.6, beta_initial: float = 0.4, beta_steps: int = 1_000_000,
                 has_duplicate: bool = True, epsilon: float = 0.0001, device="cuda:0", seed: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("DeviceProportionalMemory needs a CUDA device (no CPU fallback)")
        self.capacity = int(capacity)
        self.alpha, self.beta_initial, self.beta_steps = float(alpha), float(beta_initial), float(beta_steps)
        self.has_duplicate, self.epsilon = bool(has_duplicate), float(epsilon)
        self.device = torch.device(device)
        self.seed = int(seed)
        self._tree = torch.zeros(2 * self.capacity - 1, dtype=torch.float64, device=self.device)
        self._meta = torch.zeros(C.sizeof(_lib.SrlxState), dtype=torch.uint8, device=self.device)
        self._draws = 0
        # mapped pinned host memory: 2 x [ops_idx int64 x MAX_OPS | ops_val float64 x MAX_OPS] | out_idx int64 x MAX_BATCH | out_w float32 x
        # MAX_BATCH | flag uint64.  Two op lists, so that a launch can be left running (`_launch(wait=False)`) while the host fills
        # the other one.  (Launching the apply kernel eagerly from `update` was measured: the second launch per epoch costs more host time
        # than the overlap saves, 63 vs 57 us per epoch of the reference's speed-test; `update` only appends.)
        nb = 2 * self.MAX_OPS * 16 + self.MAX_BATCH * 12 + 8
        hp, dp = C.c_void_p(), C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_host_alloc(nb, C.byref(hp), C.byref(dp)))
        self._host_ptr, base = hp.value, dp.value
        view = lambda ct, n, off: np.ctypeslib.as_array((ct * n).from_address(hp.value + off))  # noqa: E731
        ob = self.MAX_OPS * 16  # bytes of one op list
        o2, o3, o4 = 2 * ob, 2 * ob + self.MAX_BATCH * 8, 2 * ob + self.MAX_BATCH * 12
        self._bufs = [(view(C.c_int64, self.MAX_OPS, k * ob), view(C.c_double, self.MAX_OPS, k * ob + self.MAX_OPS * 8)) for k in (0, 1)]
        self._out_idx, self._out_w = view(C.c_int64, self.MAX_BATCH, o2), view(C.c_float, self.MAX_BATCH, o3)
        self._flag = view(C.c_uint64, 1, o4)
        self._n_ops, self._seq, self._cur, self._buf_seq = 0, 0, 0, [0, 0]
        self._descs = []
        for k in (0, 1):
            d = _lib.SrlxSeam()
            d.tree, d.meta, d.ops_idx, d.ops_val = self._tree.data_ptr(), self._meta.data_ptr(), base + k * ob, base + k * ob + self.MAX_OPS * 8
            d.out_tree_idx, d.out_weights, d.flag, d.capacity = base + o2, base + o3, base + o4, self.capacity
            d.alpha, d.epsilon, d.beta_initial, d.beta_steps = self.alpha, self.epsilon, self.beta_initial, self.beta_steps
            d.has_duplicate = int(self.has_duplicate)
            self._descs.append((d, C.byref(d)))
        self._ops_idx, self._ops_val = self._bufs[0]
        self._call, self._dev_index = self.lib.srlx_tree_seam_desc, self.device.index or 0
        self.clear()

    def __del__(self):
        try:
            if getattr(self, "_host_ptr", None):
                torch.cuda.synchronize(self.device)
                self.lib.srlx_host_free(self._host_ptr)
                self._host_ptr = None
        except Exception:
            pass

    # -- helpers
    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _launch(self, batch: int, step: int, u_ptr, max_tries: int, wait: bool = True):
        """apply the pending ops of the current list, then draw `batch` items (0: apply only).  wait: return when the kernel's sequence
        word has arrived; else return at once and continue in the other op list."""
        self._seq += 1
        self._draws += 1 if batch else 0
        seed = (self.seed * 0x9E3779B97F4A7C15 + self._draws) & 0xFFFFFFFFFFFFFFFF
        ref = self._descs[self._cur][1]
        if torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self.device):
                rc = self._call(ref, self._n_ops, batch, max(int(step), 0), seed, u_ptr, max_tries, self._seq, self._s())
        else:
            rc = self._call(ref, self._n_ops, batch, max(int(step), 0), seed, u_ptr, max_tries, self._seq, self._s())
        if rc:
            _lib.check(rc)
        self._n_ops = 0
        if wait:
            self._wait(self._seq)
        else:
            self._buf_seq[self._cur] = self._seq
            self._cur ^= 1
            self._ops_idx, self._ops_val = self._bufs[self._cur]
            self._wait(self._buf_seq[self._cur])  # the launch that read the list we are about to fill (long done)

    def _wait(self, seq: int):
        flag, spins = self._flag, 0
        while flag[0] < seq:
            spins += 1
            if spins > 20_000_000:  # ~10 s: surface a dead kernel instead of spinning forever
                torch.cuda.synchronize(self.device)
                if flag[0] < seq:
                    raise _lib.SrlxError("srlx_tree_seam did not complete")

    def _flush(self):
        if self._n_ops:
            self._launch(0, 0, None, 1)
        self._wait(self._seq)

    def _read_meta(self) -> "_lib.SrlxState":
        self._flush()
        return _lib.SrlxState.from_buffer_copy(self._meta.cpu().numpy().tobytes())

    def _write_meta(self, st):
        self._meta.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))

    # -- IPriorityMemory (imemory.py:7-34)
    def clear(self) -> None:
        self._wait(self._seq)
        self._n_ops = 0
        _lib.check(self.lib.srlx_tree_clear(self._tree.data_ptr(), self.capacity, self._meta.data_ptr(), self._s()))
        self.data: List[Any] = [None] * self.capacity
        self._write = 0
        self._size = 0

    def length(self) -> int:
        return self._size

    def add(self, batch: Any, priority: Optional[float] = None, _restore_skip: bool = False) -> None:
        n = self._n_ops
        if n == self.MAX_OPS:
            self._flush()
            n = 0
        if priority is None:
            self._ops_idx[n] = -2
        else:
            self._ops_idx[n] = -3 if _restore_skip else -1
            self._ops_val[n] = priority
        self._n_ops = n + 1
        w = self._write
        self.data[w] = batch
        self._write = w + 1 if w + 1 < self.capacity else 0
        if self._size < self.capacity:
            self._size += 1

    def sample(self, batch_size: int, step: int, uniforms: Optional[np.ndarray] = None):
        if not 1 <= batch_size <= self.MAX_BATCH:
            raise ValueError(f"batch_size {batch_size} out of range [1, {self.MAX_BATCH}]")
        u_ptr, max_tries = None, 9999
        if uniforms is not None:
            u = torch.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(self.device)
            u_ptr, max_tries = u.data_ptr(), int(u.shape[1])
        # the Philox stream is keyed by (seed, draw counter) so consecutive sample() calls are independent draws
        self._launch(int(batch_size), step, u_ptr, max_tries)
        indices = self._out_idx[:batch_size].tolist()
        data, off = self.data, self.capacity - 1
        return [data[i - off] for i in indices], self._out_w[:batch_size].copy(), indices

    def update(self, indices: List[Any], priorities: np.ndarray) -> None:
        n = len(indices)
        if n == 0:
            return
        if n > self.MAX_OPS:
            raise ValueError(f"update of {n} items at once (max {self.MAX_OPS})")
        if self._n_ops + n > self.MAX_OPS:
            self._flush()
        a = self._n_ops
        self._ops_idx[a:a + n] = indices
        self._ops_val[a:a + n] = np.asarray(priorities, dtype=np.float32)  # the trainer hands float32 TD errors
        self._n_ops = a + n

    def backup(self):
        st = self._read_meta()
        return [self.capacity, float(st.max_priority), int(st.mem_size), int(st.vec_steps), self._tree.cpu().numpy().tolist(), self.data[:]]

    def restore(self, data) -> None:
        if self.capacity == data[0]:
            st = self._read_meta()
            st.max_priority, st.mem_size, st.vec_steps = float(data[1]), int(data[2]), int(data[3])
            self._write_meta(st)
            self._tree.copy_(torch.as_tensor(np.asarray(data[4], dtype=np.float64)))
            self.data = list(data[5][:])
            self._size, self._write = int(data[2]), int(data[3])
        else:  # different capacity: re-add item by item with the stored priorities (proportional_memory.py:196-205)
            self.clear()
            capacity, size, tree, tree_data = data[0], data[2], data[4], data[5]
            for i in range(size):
                self.add(tree_data[i], tree[i + capacity - 1], _restore_skip=True)

    # -- extras used by the parity tests
    @property
    def max_priority(self) -> float:
        return float(self._read_meta().max_priority)

    def tree_array(self) -> np.ndarray:
        self._flush()
        return self._tree.cpu().numpy()


class DeviceRankBasedMemory:
    """The reference's RankBasedMemory (srl/rl/memories/priority_memories/rankbased_memory.py:15-77) over device kernels
    (csrc/rankbased.cu): same constructor arguments and IPriorityMemory methods (imemory.py:7-34), `sample` returns (batches,
    float64 weights, numpy array of item indices) and the indices go back verbatim to `update`; `backup()` is the reference's
    4-element list [capacity, buffer, priorities, pos].  Use through
    `PriorityReplayBufferConfig.set_custom("simple_distributed_rl_b200.memory:DeviceRankBasedMemory", {...})`.

    Where the reference runs np.argsort over all N priorities on the host at every sample, this sorts them with a radix sort on the
    GPU, keeps the rank cdf (a function of N and alpha only) between calls, and replays numpy's choice(..., replace=False)
    algorithm on a uniform stream -- Philox by default, or `uniforms=` to inject the stream a seeded np.random would produce
    (parity tests).  Ties between equal priorities are ranked by ascending item index (numpy leaves their order unspecified).
    Payloads stay in a host list, as in the reference."""

    def __init__(self, capacity: int = 100_000, alpha: float = 0.6, beta_initial: float = 0.4, beta_steps: int = 1_000_000,
                 device="cuda:0", seed: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.SrlxError("DeviceRankBasedMemory needs a CUDA device (no CPU fallback)")
        self.capacity = int(capacity)
        self.alpha, self.beta_initial, self.beta_steps = float(alpha), float(beta_initial), float(beta_steps)
        self.device, self.seed = torch.device(device), int(seed)
        nbytes = int(self.lib.srlx_rank_scratch_bytes(self.capacity))
        if nbytes == 0:
            raise ValueError("capacity outside [1, 2^27]")
        self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._draws = 0
        self.clear()

    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def clear(self) -> None:
        self.buffer: List[Any] = []
        self._pri = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        self.pos = 0
        self._cdf_n = -1

    def length(self) -> int:
        return len(self.buffer)

    def add(self, batch: Any, priority: Optional[float] = None) -> None:
        if len(self.buffer) < self.capacity:
            self.buffer.append(batch)
        else:
            self.buffer[self.pos] = batch
        self._pri[self.pos] = float("nan") if priority is None else float(priority)  # numpy stores None as nan (:40)
        self.pos = (self.pos + 1) % self.capacity

    def sample(self, batch_size: int, step: int, uniforms: Optional[np.ndarray] = None):
        beta = self.beta_initial + (1 - self.beta_initial) * step / self.beta_steps
        beta = 1.0 if beta > 1 else float(beta)
        n = len(self.buffer)
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        w = torch.empty(batch_size, dtype=torch.float64, device=self.device)
        u_ptr, n_u = None, 0
        if uniforms is not None:
            u = torch.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(self.device)
            u_ptr, n_u = u.data_ptr(), int(u.numel())
        self._draws += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_rank_sample(self._pri.data_ptr(), self.capacity, n, self.alpha, beta, int(batch_size), u_ptr, n_u,
                                                 self.seed & 0xFFFFFFFFFFFFFFFF, self._draws, int(self._cdf_n != n), self._scratch.data_ptr(),
                                                 idx.data_ptr(), w.data_ptr(), None, None, self._s()))
        self._cdf_n = n
        sampled = idx.cpu().numpy()
        return [self.buffer[i] for i in sampled], w.cpu().numpy(), sampled

    def update(self, indices: List[Any], priorities: np.ndarray) -> None:
        n = len(indices)
        if n == 0:
            return
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(self.device)
        val = torch.as_tensor(np.asarray(priorities, dtype=np.float32)).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_rank_update(self._pri.data_ptr(), idx.data_ptr(), val.data_ptr(), n, self._s()))

    def backup(self):
        return [self.capacity, self.buffer[:], self._pri.cpu().numpy(), self.pos]

    def restore(self, data) -> None:
        self.buffer = list(data[1][:])
        pri = np.zeros(self.capacity, dtype=np.float32)
        src = np.asarray(data[2], dtype=np.float32)
        pri[: min(len(src), self.capacity)] = src[: self.capacity]
        self._pri.copy_(torch.as_tensor(pri))
        self.pos = int(data[3])
        self._cdf_n = -1

    def argsort(self) -> np.ndarray:
        """Item indices by descending priority (tests / tools)."""
        n = len(self.buffer)
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.srlx_rank_argsort(self._pri.data_ptr(), self.capacity, n, self._scratch.data_ptr(), out.data_ptr(), self._s()))
        return out.cpu().numpy().astype(np.int64)
