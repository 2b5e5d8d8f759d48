"""CPU: host logic of the image path -- layout <-> reference state_dict, the demo memory's mixing rules, loud failure without CUDA."""
import numpy as np
import pytest
import torch

from simple_distributed_rl_b200 import _lib, image


def test_spec_layout_round_trip_every_input_kind():
    for shape, t in [((28, 36, 4), "IMAGE_MAP"), ((4, 28, 36), "GRAY_HW"), ((28, 36), "GRAY_HW"), ((28, 36, 1), "GRAY_HW1"),
                     ((3, 28, 36, 1), "GRAY_HW1"), ((20, 24, 3), "RGB")]:
        sp = image.ImageNetSpec(shape, t, 5, filters=8, hidden=(16, 8))
        sd = sp.init_state_dict(1)
        assert list(sd.keys()) == sp.keys()
        flat = sp.from_state_dict(sd)
        assert flat.shape == (sp.n_params,) and sp.n_params == sum(v.numel() for v in sd.values())
        back = sp.to_state_dict(flat)
        assert all(torch.equal(sd[k], back[k]) for k in sd)
    # the DQN block's shapes for an Atari stack (dqn_image_block.py:25-27: 84 -> 21 -> 11 -> 11)
    sp = image.ImageNetSpec((84, 84, 4), "IMAGE_MAP", 6)
    assert [(g[6], g[7], g[8]) for g in sp.conv_geo] == [(21, 21, 32), (11, 11, 64), (11, 11, 64)] and sp.flat == 64 * 11 * 11


class _Mem:
    def __init__(self):
        self.items, self.updates = [], []

    def length(self):
        return len(self.items)

    def add(self, batch, priority=None):
        self.items.append(batch)

    def sample(self, batch_size, step):
        return self.items[:batch_size], [0.5] * batch_size, list(range(batch_size))

    def update(self, update_args, priorities):
        self.updates.append((list(update_args), np.array(priorities)))

    def backup(self):
        return list(self.items)

    def restore(self, data):
        self.items = list(data)


def test_demo_memory_mixing_rules():
    """priority_replay_buffer.py:177-246: demo_batch_size = max(1, int(B * ratio)); main batch = B - demo; ONE weight of 1.0 appended;
    update() trims the priorities to the main batch."""
    m = image.DemoMixMemory(_Mem(), batch_size=32, demo_ratio=1 / 16, capacity=100, warmup_size=40, seed=0)
    assert m.demo_batch_size == 2 and m.batch_size == 30
    m.select_memory = "demo"
    for i in range(5):
        m.add(("demo", i))
    m.select_memory = "main"
    for i in range(39):
        m.add(("main", i))
    assert m.sample() is None and m.is_warmup_needed() and m.length() == 44
    m.add(("main", 39))
    batches, weights, args = m.sample()
    assert len(batches) == 32 and [b[0] for b in batches].count("demo") == 2 and all(b[0] == "main" for b in batches[:30])
    assert weights.dtype == np.float32 and len(weights) == 31 and weights[-1] == 1.0  # the reference's np.append(weights, 1.0)
    m.update(args, np.arange(32, dtype=np.float32), 7)
    assert len(m.memory.updates[-1][1]) == 30 and m.step == 7
    data = m.call_backup()
    m2 = image.DemoMixMemory(_Mem(), batch_size=32, demo_ratio=1 / 16, capacity=100, warmup_size=40)
    m2.call_restore(data)
    assert m2.length() == m.length() and m2.demo == m.demo
    # demo ring wraps at capacity
    m3 = image.DemoMixMemory(_Mem(), batch_size=4, demo_ratio=0.25, capacity=3, warmup_size=3)
    m3.select_memory = "demo"
    for i in range(5):
        m3.add(i)
    assert m3.demo == [3, 4, 2]
    with pytest.raises(ValueError):
        image.DemoMixMemory(_Mem(), batch_size=1, demo_ratio=1.0, capacity=10, warmup_size=5)  # nothing left for the main batch


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_lib.SrlxError):
        image.DeviceImagePipeline((10, 10, 3), "RGB", device="cpu")
    with pytest.raises(_lib.SrlxError):
        image.ImageQNet(image.ImageNetSpec((28, 36, 1), "GRAY_HW1", 3, filters=8, hidden=(16,)))


def test_linear_table_is_a_host_function_and_matches_the_oracle():
    from oracle import image as oimg

    lib = _lib.load()
    for dst, src, border in [(84, 210, 1), (84, 160, 0), (96, 64, 1), (72, 48, 0), (5, 5, 1), (1, 7, 0), (77, 3, 1)]:
        idx, coef = np.zeros(dst, np.int32), np.zeros((dst, 2), np.int32)
        _lib.check(lib.srlx_image_linear_table(dst, src, border, idx.ctypes.data, coef.ctypes.data))
        oi, oc = oimg.linear_table(dst, src, bool(border))
        assert np.array_equal(idx, oi) and np.array_equal(coef, oc), (dst, src, border)


def test_bench_image_reference_arm_prints_the_contract_line():
    """bench.py --workload image --impl reference: the reference's ImageProcessor (or the numpy port without cv2 / baseline/_ref) on the host
    cores, one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "image", "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_device_image_processor_copies_and_pickles_without_its_device_cache():
    """the reference deep-copies configs (RLConfig.copy) and pickles them for train_mp: the processor must survive both with a cached
    pipeline attached"""
    import copy
    import pickle

    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = next((p for p in ("/root/reference", os.path.join(root, "baseline", "_ref")) if os.path.isfile(os.path.join(p, "srl", "__init__.py"))), None)
    if ref is None:
        pytest.skip("reference not present")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from srl.base.define import SpaceTypes

    from simple_distributed_rl_b200 import srl_image

    p = srl_image.DeviceImageProcessor(SpaceTypes.GRAY_HW1, (84, 84), normalize_type="0to1")
    p._pipe, p._pipe_key = _lib.load(), ("not", "copyable")  # a ctypes handle, as the cached pipeline holds
    for q in (copy.deepcopy(p), pickle.loads(pickle.dumps(p))):
        assert q == p and not hasattr(q, "_pipe") and q.resize == (84, 84) and q.normalize_type == "0to1"
    # decode of compressed items (memory.compress = True files of the reference)
    import zlib

    item = [np.zeros((2, 2), np.float32), np.ones((2, 2), np.float32), [1.0, 0.0], 0.5, 1, []]
    got = srl_image.DeviceImageMemory._decode(zlib.compress(pickle.dumps(item)))
    assert np.array_equal(got[1], item[1]) and got[3] == 0.5 and srl_image.DeviceImageMemory._decode(item) is item
