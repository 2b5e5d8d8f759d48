"""oracle/sumtree.py vs the REFERENCE's own native SumTree (srl/rl/memories/priority_memories/cpp_module/src/
proportional_memory.cpp, compiled by oracle/Makefile into oracle/_ref/).  Skipped when the module was never built."""
import glob
import importlib.util
import math
import os

import numpy as np
import pytest

from oracle import sumtree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_ref():
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", "proportional_memory_cpp*.so"))
    if not hits:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("proportional_memory_cpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("cap", [4, 10, 37, 1000])
def test_tree_state_matches_reference_cpp(cap):
    ref = _load_ref().ProportionalMemory(cap, 0.6, 0.4, 1000, True, 0.0001)
    orc = sumtree.ProportionalMemory(cap, 0.6, 0.4, 1000, True, 0.0001)
    rng = np.random.default_rng(cap)
    for i in range(3 * cap):
        p = None if rng.random() < 0.3 else float(rng.normal())
        ref.add(("item", i), p)
        orc.add(p)
    b = ref.backup()  # [capacity, max_priority, size, write, tree[:], data[:]]
    assert b[0] == cap and b[2] == orc.size and b[3] == orc.tree.write
    # -ffast-math may contract (abs(p)+eps)^alpha differently: compare at 1e-12, not bit for bit
    np.testing.assert_allclose(np.array(b[4]), orc.tree.tree, rtol=1e-12, atol=1e-15)
    assert math.isclose(b[1], orc.max_priority, rel_tol=1e-12)
    # update through both, indices as the reference returns them (tree indices)
    for _ in range(50):
        idx = rng.integers(0, cap, size=5) + cap - 1
        td = rng.normal(size=5).astype(np.float32).astype(np.float64)  # the C++ update() takes vector<float> (:208)
        ref.update([int(i) for i in idx], [float(t) for t in td])
        orc.update(idx, td)
    b = ref.backup()
    np.testing.assert_allclose(np.array(b[4]), orc.tree.tree, rtol=1e-12, atol=1e-15)
    assert math.isclose(b[1], orc.max_priority, rel_tol=1e-12)


@pytest.mark.parametrize("alpha", [0, 0.2, 0.5, 0.8, 1.0])
def test_IS_weights_match_reference_cpp(alpha):
    """tests/quick/rl/memories/test_priority_memories.py:97-117 run against the reference C++ module and the oracle."""
    eps = 0.0001
    ref = _load_ref().ProportionalMemory(10, alpha, 1, 1000000, False, eps)
    orc = sumtree.ProportionalMemory(10, alpha, 1, 1000000, False, eps)
    pri = [1, 2, 4, 3]
    for i, p in enumerate(pri):
        ref.add((i, i, i, i), p)
        orc.add(p)
    batches, w_ref, idx_ref = ref.sample(4, 1)
    rng = np.random.default_rng(0)
    idx, w, _, _ = orc.sample(4, 1, lambda i, k: float(rng.random()))
    w_by_item_ref = {b[0]: float(x) for b, x in zip(batches, w_ref)}
    w_by_item = {int(j - 9): float(x) for j, x in zip(idx, w)}
    assert sorted(w_by_item) == [0, 1, 2, 3] == sorted(w_by_item_ref)
    for k in range(4):
        assert math.isclose(w_by_item[k], w_by_item_ref[k], rel_tol=1e-6)  # the C++ module returns float32 weights
