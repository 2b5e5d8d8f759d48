"""GPU: the image pipeline and the conv Q-network (csrc/imageq.cu) through the C ABI against
  - the reference's own outputs: tests/golden/image_processor.npz (ImageProcessor, bit for bit), imageq_*.npz (dqn Trainer.train with
    the DQN image block on frozen batches: Q, target_q, loss, priorities, parameters after every update);
  - the oracle (oracle/image.py, oracle/imageq.py) at the shapes the reference's defaults give (210 x 160 x 3 Atari frames, 84 x 84 x 4
    stacks, 32 / 64 / 64 filters, 512 hidden units, batch 32);
  - the reference Runner over the registered classes (srl_image.register()) on an image env."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from image_cases import PROC_CASES  # noqa: E402


@pytest.fixture(scope="module")
def im():
    from simple_distributed_rl_b200 import image

    return image


# ---- processor --------------------------------------------------------------------------------------------------------
def test_processor_matches_the_reference_class_bit_for_bit(im):
    g = np.load(os.path.join(GOLD, "image_processor.npz"))
    for i, (src_t, shape, img_t, resize, norm, trim) in enumerate(PROC_CASES):
        pipe = im.DeviceImagePipeline(shape, src_t, img_t, resize, norm, trim, max_val=float(g[f"max_val_{i}"]))
        got = pipe(g[f"frames_{i}"]).cpu().numpy()
        want = g[f"out_{i}"]
        assert got.dtype == want.dtype and got.shape == want.shape, (i, got.dtype, got.shape, want.shape)
        assert np.array_equal(got, want), (i, np.abs(got.astype(np.float64) - want).max())
        one = pipe(g[f"frames_{i}"][0]).cpu().numpy()  # a single frame, as ImageProcessor.remap_observation takes it
        assert np.array_equal(one, want[0])


def test_processor_random_shapes_against_the_oracle_and_strided_output(im):
    from oracle import image as oimg

    rng = np.random.default_rng(3)
    for it in range(40):
        H, W, h, w = (int(v) for v in rng.integers(2, 200, size=4))
        if it % 2:
            W = 16 * int(rng.integers(1, 13))  # frame size a multiple of 16 bytes: the shared-memory staged kernel; else one thread per output
        src_t = ("RGB", "GRAY_HW")[int(rng.integers(2))]
        img_t = ("RGB", "GRAY_HW", "GRAY_HW1")[int(rng.integers(3))]
        norm = ("", "0to1", "-1to1")[int(rng.integers(3))]
        frames = rng.integers(0, 256, size=(4, H, W, 3) if src_t == "RGB" else (4, H, W), dtype=np.uint8)
        pipe = im.DeviceImagePipeline(frames.shape[1:], src_t, img_t, (w, h), norm)
        got = pipe(frames).cpu().numpy()
        for f, gt in zip(frames, got):
            assert np.array_equal(gt, oimg.process(f, src_t, img_t, (w, h), norm)), (H, W, h, w, src_t, img_t, norm)
    # empty batch
    pipe = im.DeviceImagePipeline((10, 12, 3), "RGB", "GRAY_HW", (5, 6), "0to1")
    assert pipe(np.zeros((0, 10, 12, 3), np.uint8)).shape == (0, 6, 5)


def test_processor_at_atari_batch_size_property(im):
    """4096 Atari frames (413 MB) -> 84 x 84 gray "0to1": batch == frame by frame, constant frames stay constant, output in [0, 1]."""
    pipe = im.DeviceImagePipeline((210, 160, 3), "RGB", "GRAY_HW1", (84, 84), "0to1")
    gen = torch.Generator(device="cuda").manual_seed(0)
    frames = torch.randint(0, 256, (4096, 210, 160, 3), dtype=torch.uint8, device="cuda", generator=gen)
    frames[7] = 93
    out = pipe(frames)
    assert out.shape == (4096, 84, 84, 1) and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    gray93 = np.float32((93 * 9798 + 93 * 19235 + 93 * 3735 + 16384) >> 15) / np.float32(255)
    assert torch.all(out[7] == float(gray93))
    for i in (0, 1234, 4095):
        assert torch.equal(pipe(frames[i]), out[i])


# ---- conv Q-network vs the reference's trainer ------------------------------------------------------------------------------
def _net_from_golden(im, g, uint8=False, cap=None):
    duel = str(g["dueling"]) if "dueling" in g.files else "none"  # rainbow (multisteps = 1) files carry the dueling type
    spec = im.ImageNetSpec(tuple(g["obs_shape"]), str(g["obs_stype"]), int(g["n_actions"]), filters=int(g["filters"]), hidden=tuple(g["hidden"]),
                           dueling=None if duel == "none" else duel)
    net = im.ImageQNet(spec, batch_size=g["frames"].shape[1], enable_double_dqn=bool(g["double"]), enable_rescale=bool(g["rescale"]),
                       discount=float(g["discount"]), lr=float(g["lr"]), target_model_update_interval=1000, uint8_states=uint8, batch_cap=cap)
    keys = [str(k) for k in g["keys"]]
    assert keys == spec.keys()
    net.load_state_dict({k: g["p0/" + k] for k in keys}, {k: g["t0/" + k] for k in keys})
    return net, keys


@pytest.mark.parametrize("uint8", [False, True], ids=["f32", "u8"])
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "imageq_*.npz"))), ids=lambda p: os.path.basename(p)[7:-4])
def test_imageq_matches_the_reference_trainer(im, path, uint8):
    g = np.load(path)
    net, keys = _net_from_golden(im, g, uint8)
    frames = g["frames"]
    states = frames.astype(np.float32)
    states /= np.uint8(255)
    x = frames if uint8 else states
    np.testing.assert_allclose(net.pred_q(x[0, :, 0]).cpu().numpy(), g["q_before"], rtol=1e-4, atol=1e-5)
    for u in range(len(g["losses"])):
        loss, pri, tq = net.train(x[u, :, 0], x[u, :, 1], g["actions"][u], g["rewards"][u], g["undone"][u], g["weights"][u])
        np.testing.assert_allclose(tq.cpu().numpy(), g["target_q"][u], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(float(loss), g["losses"][u], rtol=1e-4)
        np.testing.assert_allclose(pri.cpu().numpy(), g["priorities"][u], rtol=1e-3, atol=1e-5)
        sd = net.state_dict()
        for k in keys:  # post-Adam parameters: 1e-4 of the parameter scale (Adam's first steps are +-lr whatever the gradient's size)
            np.testing.assert_allclose(sd[k].numpy(), g[f"p{u + 1}/" + k], rtol=1e-4, atol=2e-6, err_msg=f"update {u} {k}")
    assert net.train_count == len(g["losses"]) and net.sync_count == 1


def _oracle_pair(im, obs_shape, stype, A, filters, hidden, B, double, rescale, seed, sync=1000, uint8=False, cap=None):
    from oracle import imageq as oq

    spec = im.ImageNetSpec(obs_shape, stype, A, filters=filters, hidden=hidden)
    net = im.ImageQNet(spec, batch_size=B, enable_double_dqn=double, enable_rescale=rescale, target_model_update_interval=sync, seed=seed,
                       uint8_states=uint8, batch_cap=cap)
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    tsd = {k: (v + 0.01 * torch.randn(v.shape, generator=torch.Generator().manual_seed(seed + 1))).numpy() for k, v in net.state_dict().items()}
    net.load_state_dict(sd, tsd)
    ora = oq.ImageQ(sd, obs_shape, stype, double, rescale, 0.99, 0.001, sync, tsd)
    return spec, net, ora


def _batch(rng, B, obs_shape, A):
    fr = rng.integers(0, 256, size=(B, 2) + tuple(obs_shape), dtype=np.uint8)
    st = fr.astype(np.float32)
    st /= np.uint8(255)
    return fr, st, rng.integers(0, A, B), rng.normal(0, 1, B).astype(np.float32), (rng.random(B) > 0.2).astype(np.int64), rng.uniform(0.3, 1, B).astype(np.float32)


@pytest.mark.parametrize("obs_shape,stype,filters,hidden,B,double,rescale", [
    ((84, 84, 4), "IMAGE_MAP", 32, (512,), 32, True, False),   # the reference's Atari setting: DQN block, window_length 4, batch 32
    ((4, 40, 52), "GRAY_HW", 16, (64, 32), 16, False, True),    # a stack of planes (NCHW input), two hidden layers
    ((30, 26), "GRAY_HW", 8, (), 5, True, True),                # one gray frame, no hidden layer, odd batch
    ((72, 96, 3), "RGB", 16, (128,), 8, True, False),           # the R2D3 processor's frame size through the DQN block
], ids=["atari84x4", "planes", "gray_nohidden", "rgb96x72"])
def test_imageq_lockstep_with_the_oracle_unresynced(im, obs_shape, stype, filters, hidden, B, double, rescale):
    """5 consecutive updates, no resynchronisation: Q, target, loss, priorities every update; parameters and the target net at the end
    (sync interval 3: two syncs inside the run)."""
    A = 6
    spec, net, ora = _oracle_pair(im, obs_shape, stype, A, filters, hidden, B, double, rescale, seed=5, sync=3)
    rng = np.random.default_rng(11)
    for u in range(5):
        fr, st, a, r, ud, w = _batch(rng, B, obs_shape, A)
        np.testing.assert_allclose(net.pred_q(st[:, 0]).cpu().numpy(), ora.pred_q(st[:, 0]), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(net.pred_target_q(st[:, 1]).cpu().numpy(), ora.pred_q(st[:, 1], target=True), rtol=2e-4, atol=2e-5)
        loss, pri, tq = net.train(st[:, 0], st[:, 1], a, r, ud, w)
        oloss, opri, otq = ora.train(st[:, 0], st[:, 1], a, r, ud, w)
        np.testing.assert_allclose(tq.cpu().numpy(), otq, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(float(loss), oloss, rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(pri.cpu().numpy(), opri, rtol=2e-3, atol=2e-5)
    # 5 Adam steps of lr = 1e-3 move a weight by at most 5e-3.  The functions above agree to 2e-4 at every update; the PARAMETERS are held to
    # a statistical bar, because Adam divides by sqrt(v) + 1e-8: an entry whose gradient is of the order of its own rounding error (1e-8:
    # half of the 4 M entries of the first dense layer see gradients below 1e-5) takes a step of up to +-lr whose size the rounding
    # decides, and after a large step (update 2 of this stream: gradients 50 x larger) ReLU / Huber-region flips of single samples change
    # gradients by 1e-3 relative.  Measured (tools/image_debug.py): <= 3 % of a block's entries off by more than 2e-5, none by more than
    # 4.3e-4.  Bar: <= 5 % beyond 2 % of one step, nothing beyond half a step.  (Single updates from the reference's own parameters are held
    # to rtol 1e-4 / atol 2e-6 in test_imageq_matches_the_reference_trainer.)
    def close(a, b, what):
        d = np.abs(a - b)
        bad = d > 2e-5 + 1e-3 * np.abs(b)
        assert bad.mean() <= 0.05 and d.max() <= 5e-4, (what, float(bad.mean()), float(d.max()))

    sd, osd, tsd = net.state_dict(), ora.state_dict(), net.state_dict(target=True)
    for k in spec.keys():
        close(sd[k].numpy(), osd[k], k)
        close(tsd[k].numpy(), ora.t[k].numpy(), "target " + k)
    assert net.train_count == 5 and net.sync_count == ora.sync_count == 2


def test_imageq_gradient_against_autograd(im):
    """phases = 1 leaves the gradient of one batch in net.grads: every block against torch autograd on the oracle (rel 1e-4 of the block's
    largest entry)."""
    obs_shape, A, B = (44, 36, 3), 5, 12
    spec, net, ora = _oracle_pair(im, obs_shape, "RGB", A, 8, (48,), B, True, False, seed=9)
    fr, st, a, r, ud, w = _batch(np.random.default_rng(2), B, obs_shape, A)
    p_before = net.params.clone()
    net.train(st[:, 0], st[:, 1], a, r, ud, w, phases=1)
    assert torch.equal(net.params, p_before) and net.train_count == 0
    import torch.nn.functional as F

    tq = torch.tensor(ora.calc_target_q(st[:, 1], r, ud))
    q = ora.forward(ora.p, torch.tensor(st[:, 0]))
    q = torch.sum(q * F.one_hot(torch.tensor(a), A).float(), dim=1)
    F.huber_loss(tq * torch.tensor(w), q * torch.tensor(w)).backward()
    got = spec.to_state_dict(net.grads.cpu().numpy())
    for k in spec.keys():
        want = ora.p[k].grad.numpy()
        np.testing.assert_allclose(got[k].numpy(), want, rtol=1e-3, atol=1e-4 * np.abs(want).max(), err_msg=k)
    net.train(st[:, 0], st[:, 1], a, r, ud, w, phases=2)  # Adam on the gradient left there
    assert net.train_count == 1 and not torch.equal(net.params, p_before)


def test_imageq_uint8_states_equal_float_states_bitwise_and_chunked_forward(im):
    obs_shape, A, B = (84, 84, 4), 4, 8
    _, net_f, _ = _oracle_pair(im, obs_shape, "IMAGE_MAP", A, 32, (512,), B, True, False, seed=1)
    _, net_u, _ = _oracle_pair(im, obs_shape, "IMAGE_MAP", A, 32, (512,), B, True, False, seed=1, uint8=True, cap=3)
    fr, st, a, r, ud, w = _batch(np.random.default_rng(4), B, obs_shape, A)
    qf, qu = net_f.pred_q(st[:, 0]), net_u.pred_q(fr[:, 0])  # cap = 3: 8 states in chunks of 3, 3, 2
    assert torch.equal(qf, qu)
    assert torch.equal(net_u.pred_q(st[:, 0]), qu)  # a uint8-capable network takes float32 states too (the worker's policy does)
    with pytest.raises(ValueError):
        net_f.pred_q(fr[:, 0])  # uint8 frames need a network built with uint8_states=True
    with pytest.raises(Exception):
        net_u.train(fr[:, 0], fr[:, 1], a, r, ud, w)  # batch 8 > batch_cap 3


def test_imageq_state_dict_round_trip_and_bad_shapes(im):
    spec = im.ImageNetSpec((28, 36, 1), "GRAY_HW1", 3, filters=8, hidden=(16,))
    net = im.ImageQNet(spec, batch_size=4, seed=3)
    sd = net.state_dict()
    net2 = im.ImageQNet(spec, batch_size=4, seed=4)
    net2.load_state_dict(sd)
    x = np.random.default_rng(0).random((4, 28, 36, 1), dtype=np.float32)
    assert torch.equal(net.pred_q(x), net2.pred_q(x)) and torch.equal(net2.pred_q(x), net2.pred_target_q(x))
    bad = dict(sd)
    bad["out_layer.weight"] = torch.zeros(3, 17)
    with pytest.raises(ValueError):
        net2.load_state_dict(bad)
    with pytest.raises(ValueError):
        im.ImageNetSpec((28, 36, 3, 2), "RGB", 3)  # the reference's reshape block raises for this too
    with pytest.raises(ValueError):
        im.ImageNetSpec((1, 1, 1), "GRAY_HW1", 3)  # padded frame (7) smaller than the first kernel (8): torch raises as well


# ---- the reference Runner over the registered classes ------------------------------------------------------------------------
def test_reference_runner_trains_image_dqn_on_device(srl_mod, tmp_path):
    """srl.Runner(image env, dqn.Config()).train(): the reference's loop, memory and worker; processor, network and trainer on the device.
    The parameter file it writes loads into the reference's own torch Parameter and gives the same Q."""
    import srl
    from srl.base.define import SpaceTypes
    from srl.rl.processors.image_processor import ImageProcessor

    sys.path.insert(0, os.path.dirname(__file__))
    import image_env
    from simple_distributed_rl_b200 import srl_image

    dqn, _ = srl_mod
    image_env.register()

    def make_cfg(proc_cls):
        cfg = dqn.Config(batch_size=16, lr=1e-3, epsilon=0.3, target_model_update_interval=25)
        cfg.input_block.image.set_dqn_block(filters=8)
        cfg.input_block.image.processors = [proc_cls(SpaceTypes.GRAY_HW1, (36, 28), normalize_type="0to1")]
        cfg.hidden_block.set((32,))
        cfg.window_length = 2
        cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 500, 32, False
        return cfg

    # (a) device memory: uint8 frames resident in HBM, uniform and proportional; file round trip in the reference's item format
    for mem_kind in ("uniform", "per"):
        srl_image.register(device_memory=True)
        try:
            cfg = make_cfg(srl_image.DeviceImageProcessor)
            if mem_kind == "per":
                cfg.memory.set_proportional()
            runner = srl.Runner("PixelGrid-b200", cfg)
            state = runner.train(max_train_count=40)
            mem = state.memory
            assert type(mem).__name__ == "DeviceImageMemory" and mem.S.dtype == torch.uint8 and mem.S.is_cuda
            assert state.trainer.get_train_count() == 40 and np.isfinite(state.trainer.info["loss"]) and mem.length() >= 32 + 40 - 2
            # what sits in HBM is what the reference's processor + window stacking produced, as bytes: a frame's background is blue = 40,
            # gray = (40 * 3735 + 16384) >> 15 = 5 after cv2's conversion
            mem.flush()
            assert int(mem.S[:mem.length()].min()) >= 0 and 5 in mem.S[:mem.length()].unique().tolist()
            if mem_kind == "per":
                assert mem.per.length() == mem.length() and mem.per.max_priority > 0
            path_m = str(tmp_path / f"m_{mem_kind}.dat")
            runner.save_memory(path_m)
            n_before, s_before = mem.length(), mem.S[:mem.length()].clone()
            runner2 = srl.Runner("PixelGrid-b200", cfg)
            runner2.load_memory(path_m)
            mem2 = runner2.make_memory()
            mem2.flush()
            assert mem2.length() == n_before and torch.equal(mem2.S[:n_before], s_before)
            item = mem2._item(0)
            assert len(item) == 6 and item[0].dtype == np.float32 and item[0].shape == (28, 36, 2)  # the worker's record (dqn.py:229-246)
        finally:
            srl_image.unregister()
    # (b) the reference's own Memory (host lists), device processor / network / trainer
    srl_image.register()
    try:
        runner = srl.Runner("PixelGrid-b200", make_cfg(srl_image.DeviceImageProcessor))
        state = runner.train(max_train_count=60)
        assert type(state.trainer).__name__ == "ImageTrainer" and type(state.parameter).__name__ == "ImageParameter"
        assert type(state.memory).__module__.startswith("srl.") and type(state.worker.worker).__module__.startswith("srl.")
        assert state.trainer.get_train_count() == 60 and state.parameter.net.train_count == 60 and state.parameter.net.sync_count == 3
        assert np.isfinite(state.trainer.info["loss"]) and state.trainer.info["sync"] == 3
        assert tuple(runner.rl_config.observation_space.shape) == (28, 36, 2)
        assert len(runner.evaluate(max_episodes=2)) == 2
        path = str(tmp_path / "p.dat")
        runner.save_parameter(path)
        frame = image_env.PixelGrid().reset()
        q_dev = state.parameter.pred_q(np.zeros((1, 28, 36, 2), np.float32) + 0.25)
        # the device processor == the reference's on the env's own frames
        ref_p, dev_p = ImageProcessor(SpaceTypes.GRAY_HW1, (36, 28), normalize_type="0to1"), srl_image.DeviceImageProcessor(SpaceTypes.GRAY_HW1, (36, 28), normalize_type="0to1")
        sp = image_env.PixelGrid().observation_space
        ns = ref_p.remap_observation_space(sp)
        assert ns == dev_p.remap_observation_space(sp)
        assert np.array_equal(ref_p.remap_observation(frame, sp, ns), dev_p.remap_observation(frame, sp, ns))
    finally:
        srl_image.unregister()
    ref_runner = srl.Runner("PixelGrid-b200", make_cfg(ImageProcessor))
    ref_runner.set_device("CPU")
    ref_runner.load_parameter(path)
    par = ref_runner.make_parameter()
    assert type(par).__module__.startswith("srl.")
    np.testing.assert_allclose(par.pred_q(np.zeros((1, 28, 36, 2), np.float32) + 0.25), q_dev, rtol=1e-4, atol=5e-5)  # torch CPU conv vs device


def test_imageq_on_the_tcgen05_tiles_in_a_subprocess():
    """SRLX_IMAGE_TC3=1 (read when the library loads): every map of the network on the tcgen05 3 x TF32 tiles (csrc/gemm_tc3.cuh) -- the
    reference-trainer goldens and the gradient check again, to the same tolerances."""
    import subprocess

    env = dict(os.environ, SRLX_IMAGE_TC3="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k",
                        "matches_the_reference_trainer or gradient_against_autograd or planes or gray_nohidden"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_image_dqn_learns_through_the_reference_runner(srl_mod):
    """Learning gate of the image path: srl.Runner on PixelGrid (optimal return 0.94) over the device processor / conv Q-network / trainer
    / uint8 replay reaches >= 0.85 after 3000 updates (tools/image_learning_check.py: 0.94 on seeds 1, 2, 3, ~3 s of training each)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    sys.path.insert(0, os.path.dirname(__file__))
    import image_env
    import image_learning_check as lc
    from simple_distributed_rl_b200 import srl_image

    image_env.register()
    srl_image.register(device_memory=True)
    try:
        reward, _ = lc.run(seed=1, n_train=3000)
    finally:
        srl_image.unregister()
    assert reward >= 0.85, reward


def test_reference_runner_trains_image_rainbow_dueling_on_device(srl_mod, tmp_path):
    """Rainbow with multisteps = 1 ("Rainbow_no_multisteps:torch") and the dueling block over the DQN image block: the reference's loop,
    Memory (proportional) and Worker; network / trainer on the device; the parameter file loads into the reference's torch Parameter."""
    import srl
    from srl.base.define import SpaceTypes
    from srl.rl.processors.image_processor import ImageProcessor

    sys.path.insert(0, os.path.dirname(__file__))
    import image_env
    from simple_distributed_rl_b200 import srl_image

    _, rainbow = srl_mod
    image_env.register()

    def make_cfg():
        cfg = rainbow.Config(batch_size=16, lr=1e-3, epsilon=0.3, target_model_update_interval=25, multisteps=1, enable_noisy_dense=False)
        cfg.input_block.image.set_dqn_block(filters=8)
        cfg.input_block.image.processors = [ImageProcessor(SpaceTypes.GRAY_HW1, (36, 28), normalize_type="0to1")]
        cfg.hidden_block.set_dueling_network((24, 16), dueling_type="average")
        cfg.window_length = 2
        cfg.memory.set_proportional()
        cfg.memory.capacity, cfg.memory.warmup_size, cfg.memory.compress = 500, 32, False
        return cfg

    for device_memory in (False, True):
        srl_image.register(device_memory=device_memory)
        try:
            runner = srl.Runner("PixelGrid-b200", make_cfg())
            state = runner.train(max_train_count=50)
            assert type(state.trainer).__name__ == "ImageTrainer" and type(state.parameter).__name__ == "ImageRainbowParameter"
            assert type(state.worker.worker).__module__ == "srl.algorithms.rainbow.rainbow_nomultisteps"
            assert type(state.memory).__name__ == ("DeviceImageMemory" if device_memory else "Memory")
            net = state.parameter.net
            assert net.train_count == 50 and net.sync_count == 2 and np.isfinite(state.trainer.info["loss"]) and net.c.target_f32 == 1
            # the off-branch blocks of the dueling output layer never move
            out, kk, off = net.spec.dense[-1]
            blk = net.params[off:off + out * (kk + 1)].reshape(out, kk + 1).cpu().numpy()
            H = net.spec.duel_hidden
            assert np.all(blk[0, H:2 * H] == 0) and np.all(blk[1:, :H] == 0) and np.any(blk[0, :H] != 0) and np.any(blk[1:, H:2 * H] != 0)
            path = str(tmp_path / f"p_{int(device_memory)}.dat")
            runner.save_parameter(path)
            q_dev = state.parameter.pred_q(np.zeros((1, 28, 36, 2), np.float32) + 0.25)
        finally:
            srl_image.unregister()
        ref_runner = srl.Runner("PixelGrid-b200", make_cfg())
        ref_runner.set_device("CPU")
        ref_runner.load_parameter(path)
        par = ref_runner.make_parameter()
        assert type(par).__module__.startswith("srl.")
        np.testing.assert_allclose(par.pred_q(np.zeros((1, 28, 36, 2), np.float32) + 0.25), q_dev, rtol=1e-4, atol=5e-5)
