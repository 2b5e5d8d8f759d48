"""CPU oracle for the srl hot path (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` is a CPU restatement of the reference algorithms
(pocokhc/simple_distributed_rl v1.4.5, paths relative to the reference root) that the
CUDA path in ``simple_distributed_rl_b200`` is checked against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` leg may
import it.  The product package never imports ``oracle``.

Pinning status (see DESIGN.md "Oracle"):
  * sumtree / replay / targets / nets / grid : pinned against the reference itself, executed in
    the build container, frozen as ``tests/golden/*.npz`` by ``tests/golden/make_golden.py``,
    plus the reference's own KATs (tests/quick/rl/memories/test_priority_memories.py).
  * cartpole : restates third-party gymnasium==1.2.0 ``classic_control/cartpole.py`` (not vendored,
    not installed) -> **parity unpinned** against gymnasium; pinned only oracle<->device.
"""
