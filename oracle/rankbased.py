"""CPU restatement of the reference's rank-based replay memory (TEST INFRASTRUCTURE).

Follows srl/rl/memories/priority_memories/rankbased_memory.py:15-77 (RankBasedMemory: priorities float32 [capacity], add / sample /
update) with the one random source -- np.random.choice(sorted_indices, size=B, p=probs, replace=False) (:55) -- restated from
numpy's legacy RandomState.choice (numpy/random/mtrand.pyx, the `replace=False, p given` branch) so that the uniform stream is an
ARGUMENT: the tests feed the same stream (np.random.RandomState(seed).random_sample) to the reference, to this restatement and to
the CUDA kernels (csrc/rankbased.cu).

Pinned by tests/golden/rankbased.npz (the reference class run with np.random seeded, tests/golden/make_golden.py::gen_rankbased).
"""
import numpy as np


def choice_without_replacement(p, size, uniforms):
    """numpy RandomState.choice(len(p), size, replace=False, p=p) on an explicit uniform stream; returns (indices, uniforms used)."""
    p = np.array(p, dtype=np.float64, copy=True)
    found = np.zeros(size, dtype=np.int64)
    n_uniq, used = 0, 0
    while n_uniq < size:
        need = size - n_uniq
        x = np.asarray(uniforms[used:used + need], dtype=np.float64)
        assert len(x) == need, "uniform stream exhausted"
        used += need
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, unique_indices = np.unique(new, return_index=True)
        unique_indices.sort()
        new = new.take(unique_indices)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found, used


class RankBasedMemory:
    """rankbased_memory.py:15-77 without the python payload list (payload = the item index)."""

    def __init__(self, capacity=100_000, alpha=0.6, beta_initial=0.4, beta_steps=1_000_000):
        self.capacity, self.alpha, self.beta_initial, self.beta_steps = int(capacity), alpha, beta_initial, beta_steps
        self.clear()

    def clear(self):
        self.size = 0
        self.priorities = np.zeros(self.capacity, dtype=np.float32)
        self.pos = 0

    def length(self):
        return self.size

    def add(self, priority=None):
        self.size = min(self.size + 1, self.capacity)
        self.priorities[self.pos] = np.nan if priority is None else priority  # numpy stores None as nan (:40)
        self.pos = (self.pos + 1) % self.capacity

    def sorted_indices(self):
        return np.argsort(-self.priorities[: self.size], kind="stable")  # ties: ascending item index (numpy's default leaves it open)

    def sample(self, batch_size, step, uniforms):
        beta = self.beta_initial + (1 - self.beta_initial) * step / self.beta_steps
        if beta > 1:
            beta = 1
        N = self.size
        sorted_indices = self.sorted_indices()
        ranks = np.arange(1, N + 1)
        probs = (1 / ranks) ** self.alpha
        probs /= probs.sum()
        found, used = choice_without_replacement(probs, batch_size, uniforms)
        sampled = sorted_indices[found]
        weights = (N * probs[found]) ** (-beta)
        weights = weights / weights.max()
        return sampled, weights, found, used

    def update(self, indices, priorities):
        for idx, td in zip(indices, priorities):
            self.priorities[idx] = td
