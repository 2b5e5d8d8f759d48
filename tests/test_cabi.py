"""The drop-in boundary on a machine without a GPU: libsrlx.so loads, exports every function include/srlx.h declares, its struct
sizes match the ctypes mirror, argument errors come back as error codes with a message (no compute calls here), the product
package never imports the oracle, and every product entry point refuses to run without the CUDA path (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simple_distributed_rl_b200")


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "srlx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # comments mention symbol names too
    return sorted(set(re.findall(r"\b(srlx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from simple_distributed_rl_b200 import _lib

    names = _declared_functions()
    assert len(names) >= 20 and "srlx_learn" in names and "srlx_vec_step" in names and "srlx_returns_scan" in names
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/srlx.h but not exported: {missing}"
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert set(names) <= bound | {"srlx_version", "srlx_last_error"}, sorted(set(names) - bound)  # the ctypes mirror binds them all


def test_struct_sizes_and_version():
    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()  # raises on an ABI mismatch between the header structs and the ctypes mirror
    assert lib.srlx_sizeof_engine() == C.sizeof(_lib.SrlxEngine)
    assert lib.srlx_sizeof_state() == C.sizeof(_lib.SrlxState) == 128
    assert lib.srlx_sizeof_net() == C.sizeof(_lib.SrlxNet)
    assert lib.srlx_version() >= 1


def test_argument_errors_are_codes_with_a_message():
    """Validation happens before any CUDA call, so it can be exercised without a device: NULL engine, NULL buffers, bad method."""
    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()
    assert lib.srlx_learn(None, 1, 0) != 0 and b"NULL" in lib.srlx_last_error()
    rc = lib.srlx_returns_scan(None, None, None, None, None, None, None, 4, 4, 0.9, 0.9, 7, 0, 0, 0.0, 0.0, 0)
    assert rc != 0 and b"method" in lib.srlx_last_error()
    rc = lib.srlx_returns_scan(None, None, None, None, None, None, None, 4, 4, 0.9, 0.9, _lib.RETURNS_GAE, 0, 0, 0.0, 0.0, 0)
    assert rc != 0 and b"NULL" in lib.srlx_last_error()
    with pytest.raises(_lib.SrlxError, match="NULL"):
        _lib.check(lib.srlx_learn(None, 1, 0))


def test_missing_library_fails_loudly(monkeypatch):
    from simple_distributed_rl_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "no_such_dir", "libsrlx.so"))
    with pytest.raises(_lib.SrlxError, match="no CPU fallback"):
        _lib.load()


def test_no_cpu_fallback_anywhere():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusals below are for machines without one")
    from simple_distributed_rl_b200 import _lib
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig
    from simple_distributed_rl_b200.memory import DeviceProportionalMemory
    from simple_distributed_rl_b200.returns import returns_scan

    with pytest.raises((_lib.SrlxError, RuntimeError, AssertionError)):
        DeviceEngine(EngineConfig(env="Grid", algo="dqn", hidden=(8,), n_envs=4, ring_rows=4, batch_size=2, warmup_size=4))
    with pytest.raises(_lib.SrlxError):
        DeviceProportionalMemory(16)
    with pytest.raises(_lib.SrlxError):
        returns_scan(torch.zeros(3, 4), torch.zeros(3, 4, dtype=torch.uint8), torch.zeros(3, 4), torch.zeros(3, 4))


def test_product_package_never_imports_the_oracle():
    offenders = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    for f in os.listdir(os.path.join(PKG, "csrc")):
        if f.endswith((".cu", ".cuh")):
            assert "#include \"../../oracle" not in open(os.path.join(PKG, "csrc", f)).read()


def test_env_tables_and_engine_config_host_logic():
    """Host-side tables the kernels consume (envspec.py) and the config checks that run before any device call."""
    from simple_distributed_rl_b200 import _lib
    from simple_distributed_rl_b200.envspec import division_table, make_env_spec

    eng = _lib.SrlxEngine()
    make_env_spec("Grid").fill(eng)
    assert (eng.env_id, eng.obs_dim, eng.n_actions, eng.trunc_limit, eng.grid_w, eng.grid_h, eng.grid_n_starts) == (_lib.ENV_GRID, 2, 4, 51, 6, 5, 1)
    cdf = np.array(eng.grid_slip_cdf[:16]).reshape(4, 4)
    assert np.allclose(cdf[:, -1], 1.0) and np.all(np.diff(cdf, axis=1) >= 0)
    make_env_spec("Pendulum-v1", action_division_num=5).fill(eng)
    assert eng.n_actions == 5 and [eng.act_tbl[i] for i in range(5)] == [-2.0, -1.0, 0.0, 1.0, 2.0]
    assert [float(x) for x in division_table(-2.0, 2.0, 3)] == [-2.0, 0.0, 2.0]
    with pytest.raises(ValueError):
        make_env_spec("LunarLander-v2")
    with pytest.raises(ValueError):
        make_env_spec("Pendulum-v1", action_division_num=1)
