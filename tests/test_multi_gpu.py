"""Single-learner data parallelism over the GPUs of one node: the in-kernel gradient all-reduce and the global PER
normalisation of csrc/learner_fast.cu against the oracle.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a one-GPU box.

One process drives one engine per device (parallel.link_engines).  The torchrun / CUDA-IPC path (parallel.link_engine_distributed)
is exercised by tools/dp_check.py and bench.py --gpus N."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import engine as oeng  # noqa: E402
from oracle import nets as onets  # noqa: E402


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


KW = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=64,
          ring_rows=16, batch_size=32, warmup_size=128, target_update_interval=3, per_beta_steps=50.0)
LEARNER_SEED = 77


def _make(world, kw=KW, debug=True):
    from simple_distributed_rl_b200 import parallel
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    engs = [DeviceEngine(EngineConfig(**kw, seed=10 + r), device=f"cuda:{r}", debug=debug) for r in range(world)]
    for e in engs:
        e.run(kw["ring_rows"] + 3, 0)  # different seeds: every rank fills its own shard with its own trajectories
    parallel.link_engines(engs, learner_seed=LEARNER_SEED)
    return engs


def _noise(eng, call_id):
    from simple_distributed_rl_b200 import _lib

    out = torch.empty(eng.spec.n_params, dtype=torch.float32, device=eng.device)
    with torch.cuda.device(eng.device):
        _lib.check(eng.lib.srlx_noise_fill(LEARNER_SEED, 1, int(call_id), out.data_ptr(), eng.spec.n_params, eng._stream()))
    return out.cpu().numpy()


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_update_equals_one_trainer_on_the_global_batch(world):
    """Every update: (1) all ranks hold bit-identical parameters, moments and target network afterwards; (2) they equal ONE
    Trainer.train() of the oracle (rainbow/model_torch.py:85-122) on the union of the ranks' batches -- world x 32 items, Huber
    mean over all of them -- to 1e-4; (3) the IS weights are those of one memory holding all shards: (N p / total)^-beta with the
    global N and total, divided by the maximum over the whole global batch (proportional_memory.py:159-167); (4) every rank's
    SumTree receives the priorities of its own items."""
    _need_gpus(world)
    engs = _make(world)
    e0 = engs[0]
    cfg = oeng.EngineConfig(**KW, seed=10)
    spec = onets.NetSpec(e0.D, tuple(KW["hidden"]), e0.A, KW["dueling"], KW["noisy"])
    mu, sigma = e0.get_params()
    adam = onets.AdamState(spec, mu, sigma, lr=cfg.lr)
    tm, ts = e0.get_target()
    mask = spec.sigma_mask("rainbow")
    B, M, D = KW["batch_size"], KW["multisteps"], e0.D
    for u in range(7):
        trees = [e.t["tree"].cpu().numpy() for e in engs]
        sizes = [e.read_state().mem_size for e in engs]
        for e in engs:
            e.learn(1)
        for e in engs:
            torch.cuda.synchronize(e.device)
            e.check_dp_alive()
        # (1) identical replicas
        for k in ("params", "params_sigma", "target", "target_sigma", "adam_m", "adam_v"):
            ref = e0.t[k].cpu()
            for e in engs[1:]:
                assert torch.equal(ref, e.t[k].cpu()), (u, k)
        # the global batch as the ranks sampled it
        states, acts, rews, terms, weights, idxs = [], [], [], [], [], []
        for e in engs:
            win = e.t["dbg_windows"].cpu().numpy()
            ns = B * (M + 1) * D
            states.append(win[:ns].reshape(B, M + 1, D))
            acts.append(win[ns:ns + B * M].reshape(B, M).astype(np.int64))
            rews.append(win[ns + B * M:ns + 2 * B * M].reshape(B, M))
            terms.append(win[ns + 2 * B * M:].reshape(B, M))
            weights.append(e.t["dbg_weights"].cpu().numpy())
            idxs.append(e.t["dbg_sample_idx"].cpu().numpy())
        # (3) IS weights of one memory over all shards
        beta_step = max(u - 1, 0)
        beta = min(1.0, cfg.per_beta_initial + (1 - cfg.per_beta_initial) * beta_step / cfg.per_beta_steps)
        n_glob, total_glob = float(sum(sizes)), float(sum(t[0] for t in trees))
        raw = [(n_glob * (t[i] / total_glob)) ** (-beta) for t, i in zip(trees, idxs)]
        wmax = max(r.max() for r in raw)
        for r, w in zip(raw, weights):
            np.testing.assert_allclose(w, r / wmax, rtol=1e-6)
        assert max(w.max() for w in weights) == 1.0
        # (2) one trainer on the union
        noise = tuple(_noise(e0, u * 3 + p) for p in range(3))
        res = onets.train_update(spec, adam, tm, ts, algo="rainbow", states=np.concatenate(states), actions=np.concatenate(acts),
                                 rewards=np.concatenate(rews), dones=np.concatenate(terms), weights=np.concatenate(weights),
                                 discount=cfg.discount, multisteps=M, retrace_h=cfg.retrace_h, enable_double_dqn=True, enable_rescale=False,
                                 noise=noise, sigma_mask=mask)
        mu_d, sg_d = e0.get_params()
        np.testing.assert_allclose(mu_d, adam.mu.detach().numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(sg_d, adam.sigma.detach().numpy(), rtol=1e-4, atol=2e-5)
        g = e0.t["dbg_grads"].cpu().numpy()
        P = spec.n_params
        np.testing.assert_allclose(g[:P], res["grad_mu"], rtol=1e-3, atol=2e-6)
        np.testing.assert_allclose(g[P:], res["grad_sigma"], rtol=1e-3, atol=2e-6)
        tq = np.concatenate([e.t["dbg_target_q"].cpu().numpy() for e in engs])
        np.testing.assert_allclose(tq, res["target_q"], rtol=1e-4, atol=1e-5)
        if u % KW["target_update_interval"] == 0:
            tm, ts = adam.mu.detach().numpy().copy(), adam.sigma.detach().numpy().copy()
        tm_d, ts_d = e0.get_target()
        np.testing.assert_allclose(tm_d, tm, rtol=1e-4, atol=2e-5)
        # (4) every shard's tree got the new priorities of its own items (the last item touching a leaf wins)
        off = 0
        for e, t_old, idx in zip(engs, trees, idxs):
            t_new = e.t["tree"].cpu().numpy()
            pr = (np.abs(res["priorities"][off:off + B].astype(np.float64)) + cfg.per_epsilon) ** cfg.per_alpha
            last = {int(i): p for i, p in zip(idx, pr)}
            for i, p in last.items():
                assert math.isclose(t_new[i], p, rel_tol=1e-3, abs_tol=1e-5), (u, i)
            cap = e.cap
            np.testing.assert_allclose(t_new[0], t_new[cap - 1:].sum(), rtol=1e-9)
            off += B
        # re-synchronise the oracle (as the lockstep tests do): fp32 ulps are amplified by Adam's g / sqrt(v)
        adam.mu.data.copy_(torch.as_tensor(mu_d))
        adam.sigma.data.copy_(torch.as_tensor(sg_d))
        tm, ts = e0.get_target()
        st = e0.read_state()
        assert st.train_count == u + 1 and all(e.read_state().train_count == u + 1 for e in engs)


def test_data_parallel_many_updates_per_launch_equal_one_by_one():
    """n updates in one launch per rank (the exchange runs inside the persistent kernels, flags carry the update number) ==
    n launches of one update: bit-identical state on every rank."""
    _need_gpus(2)
    a, b = _make(2, debug=False), _make(2, debug=False)
    for n in (3, 8):
        for e in a:
            e.learn(n)
        for _ in range(n):
            for e in b:
                e.learn(1)
        for e in a + b:
            torch.cuda.synchronize(e.device)
            e.check_dp_alive()
        for ea, eb in zip(a, b):
            for k in ("params", "params_sigma", "target", "adam_m", "adam_v", "tree"):
                assert torch.equal(ea.t[k], eb.t[k]), (n, k)
    assert torch.equal(a[0].t["params"].cpu(), a[1].t["params"].cpu())


def test_data_parallel_learning_gate_cartpole():
    """The single learner over two GPUs has to LEARN: the bench configuration (Rainbow default shape) with 512 env copies per GPU,
    global batch 64; greedy evaluation past 150 (a random policy scores ~22) -- the N > 1 learning gate."""
    _need_gpus(2)
    from simple_distributed_rl_b200 import parallel
    from simple_distributed_rl_b200.engine import EngineConfig
    from simple_distributed_rl_b200.runner import VecRunner

    kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=512,
              ring_rows=256, batch_size=32, warmup_size=1000, lr=1e-3, target_update_interval=1000)
    runners = [VecRunner(EngineConfig(**kw, seed=1 + r), device=f"cuda:{r}") for r in range(2)]
    engs = [r.engine for r in runners]
    parallel.link_engines(engs, learner_seed=5)
    for step in range(400):
        for e in engs:
            e.vec_step()
        for e in engs:
            e.learn(128)  # one update of the global batch (64 items) per 8 env steps
    for e in engs:
        torch.cuda.synchronize(e.device)
        e.check_dp_alive()
    assert torch.equal(engs[0].t["params"].cpu(), engs[1].t["params"].cpu())
    st = engs[0].read_state()
    assert st.train_count > 40_000
    for r in runners:
        assert float(np.mean(r.evaluate(max_episodes=50, test_epsilon=0.0))) >= 150.0


def test_r2d2_actor_shards_one_learner_under_torchrun():
    """BASELINE configs[3]'s layout in small: two ranks (torchrun, NCCL), each with its own env copies and replay shard, ONE trainer step
    per update on the global batch (gradient all-reduce between backward and Adam, simple_distributed_rl_b200/r2d2.py::
    link_data_parallel): parameters, target network and Adam state stay bit-identical on both ranks while they keep acting and learning."""
    import json
    import os
    import subprocess
    import sys

    _need_gpus(2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, R2D2_SMALL="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(root, "tools", "r2d2_dp_check.py")], capture_output=True, text=True, env=env,
                       timeout=600)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("R2D2DP ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(line[-1][len("R2D2DP "):])
    assert out["world"] == 2 and out["replicas_bit_identical"] is True and out["train_count"] >= 20 and out["global_batch"] == 64


def test_image_trainer_data_parallel_equals_one_trainer_on_the_global_batch():
    """The conv Q-network's trainer over 2 GPUs (image.ImageQNet.train(phases=1) -> gradient average -> apply_gradients(), what
    train_data_parallel does with an NCCL all-reduce under torchrun): two ranks with half the batch each == one network on the whole
    batch, update after update, and the ranks stay bit-identical."""
    _need_gpus(2)
    from simple_distributed_rl_b200 import image as im

    obs, A, B = (84, 84, 4), 6, 16
    spec = im.ImageNetSpec(obs, "IMAGE_MAP", A)
    one = im.ImageQNet(spec, batch_size=2 * B, seed=3, target_model_update_interval=2, device="cuda:0")
    ranks = [im.ImageQNet(spec, batch_size=B, seed=3, target_model_update_interval=2, device=f"cuda:{r}") for r in range(2)]
    rng = np.random.default_rng(0)
    for u in range(4):
        fr = rng.integers(0, 256, size=(2, 2 * B) + obs, dtype=np.uint8)
        st = (fr / np.float32(255)).astype(np.float32)
        a, r_ = rng.integers(0, A, 2 * B), rng.normal(0, 1, 2 * B).astype(np.float32)
        ud, w = (rng.random(2 * B) > 0.2).astype(np.float32), rng.uniform(0.3, 1, 2 * B).astype(np.float32)
        loss1, pri1, tq1 = one.train(st[0], st[1], a, r_, ud, w)
        outs = []
        for k, net in enumerate(ranks):
            sl = slice(k * B, (k + 1) * B)
            outs.append(net.train(st[0][sl], st[1][sl], a[sl], r_[sl], ud[sl], w[sl], phases=1))
        g = (ranks[0].grads + ranks[1].grads.to("cuda:0")) / 2  # = dist.all_reduce(SUM) / world
        ranks[0].grads.copy_(g)
        ranks[1].grads.copy_(g.to("cuda:1"))
        for net in ranks:
            net.apply_gradients()
        assert torch.equal(ranks[0].params.cpu(), ranks[1].params.cpu()) and torch.equal(ranks[0].target.cpu(), ranks[1].target.cpu())
        np.testing.assert_allclose(torch.cat([o[2].cpu() for o in outs]).numpy(), tq1.cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(torch.cat([o[1].cpu() for o in outs]).numpy(), pri1.cpu().numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(float(outs[0][0] + outs[1][0].cpu().to(outs[0][0].device)) / 2, float(loss1), rtol=1e-5)
        d = (ranks[0].params - one.params).abs()
        # Adam's +-lr steps on entries whose gradient is rounding noise (see test_image_gpu.py): statistical bar on the parameters
        assert float((d > 2e-5).float().mean()) <= 0.02 and float(d.max()) <= 5e-4, (u, float((d > 2e-5).float().mean()), float(d.max()))
    assert ranks[0].train_count == ranks[1].train_count == one.train_count == 4 and ranks[0].sync_count == one.sync_count == 2


def test_image_network_on_the_tcgen05_tiles_on_the_second_device():
    """Function attributes (the tcgen05 tiles' 97 KB of dynamic shared memory) are per device: a network on cuda:1 of a process that already
    ran one on cuda:0 must still launch -- forward maps of 256 states go to the tcgen05 engine by the default rule."""
    _need_gpus(2)
    from simple_distributed_rl_b200 import image as im

    spec = im.ImageNetSpec((84, 84, 4), "IMAGE_MAP", 6)
    x = np.random.default_rng(0).integers(0, 256, size=(256, 84, 84, 4), dtype=np.uint8)
    qs = []
    for d in (0, 1):
        net = im.ImageQNet(spec, batch_size=256, uint8_states=True, seed=3, device=f"cuda:{d}")
        qs.append(net.pred_q(x).cpu())
    assert torch.equal(qs[0], qs[1])
