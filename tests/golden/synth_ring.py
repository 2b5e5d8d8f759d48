"""Synthetic replay contents for the learner goldens (tests/golden/make_learner_golden.py generates the reference's outputs on them;
the tests rebuild the same ring from the seed when a fixture is too large to store).  numpy only -- no reference, no CUDA."""
import numpy as np

from simple_distributed_rl_b200 import checkpoint as ck


def ring_seed_of(case: str) -> int:
    return sum(ord(c) * (i + 1) for i, c in enumerate(case))


def synth_ring(kw, seed):
    """A ring of plausible trajectories: next_obs of a step is obs of the next one unless the episode ended."""
    D, A = (4, 2) if kw["env"] == "CartPole-v1" else (2, 4)
    E, R, M, G = kw["n_envs"], kw["ring_rows"], kw["multisteps"], kw["vec_steps"]
    rng = np.random.default_rng(seed)
    v = ck.RingView(E, R, M, A, D, vec_steps=G)
    cur = rng.normal(0, 0.5, size=(E, D)).astype(np.float32)
    for g in range(G):
        row = g % R
        sl = slice(row * E, (row + 1) * E)
        nxt = (cur + rng.normal(0, 0.2, size=(E, D))).astype(np.float32)
        done = rng.random(E) < 0.12
        term = done & (rng.random(E) < 0.7)
        v.obs[sl], v.next_obs[sl] = cur, nxt
        v.action[sl] = rng.integers(0, A, size=E)
        v.reward[sl] = rng.normal(0, 1, size=E).astype(np.float32)
        v.term[sl], v.done[sl] = term, done
        fresh = rng.normal(0, 0.5, size=(E, D)).astype(np.float32)
        cur = np.where(done[:, None], fresh, nxt)
    if kw["mem_kind"]:
        g_lo, n_g = v.valid_rows()
        leaves = np.zeros(E * R)
        for g in range(g_lo, g_lo + n_g):
            row = g % R
            p = rng.random(E) ** 2 + 0.01
            p[rng.random(E) < kw.get("zero_leaves", 0.0)] = 0.0  # the reference re-draws on a zero-priority leaf
            leaves[row * E:(row + 1) * E] = p
        v.leaf_priority = leaves
        v.max_priority = float(max(1.0, leaves.max()))
    return v
