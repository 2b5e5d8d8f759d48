"""GPU tests of the tensor-core inference mode (csrc/qnet_tc.cu: tcgen05.mma + TMEM + TMA): the dense bf16 kernel against torch, the
whole-network forward against the fp32 seam.  An explicit non-parity mode: the bar is bf16 resolution, stated here."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dense(x, w, b, relu, out_f32):
    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()
    M, K = x.shape
    N = w.shape[0]
    y = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    _lib.check(lib.srlx_dense_bf16_tc(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), None if b is None else b.data_ptr(), y.data_ptr(),
                                      y.stride(0), int(out_f32), M, N, K, int(relu), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("M,N,K,relu,out_f32,bias", [
    (128, 128, 64, False, True, False),      # one tile, one k-block: the descriptors alone
    (128, 128, 256, False, True, True),      # four k-blocks: one trip round the stage ring
    (256, 256, 1024, True, False, True),     # sixteen k-blocks: stages reused four times (empty-barrier parity), 4 CTAs
    (300, 200, 72, True, False, True),       # ragged M, N, K (TMA zero fill, masked stores)
    (1000, 512, 512, True, False, True),
    (4096, 64, 8, True, False, True),        # the observation layer: K padded to 8
    (129, 3, 1024, False, True, True),       # the output layer: three columns, fp32 out
    (70000, 512, 512, True, False, True),    # 2 188 CTAs: more tiles than SMs
])
def test_dense_bf16_tc_equals_torch(M, N, K, relu, out_f32, bias):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    x = (torch.randn((M, K), device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn((N, K), device="cuda", generator=g) * (1.0 / K ** 0.5)).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    y = _dense(x, w, b, relu, out_f32)
    ref = x.float() @ w.float().T
    if b is not None:
        ref = ref + b
    if relu:
        ref = ref.relu()
    assert torch.isfinite(y.float()).all()
    # fp32 accumulation of exact bf16 products: only the summation order (and the bf16 rounding of the output) differs
    tol = dict(rtol=2e-4, atol=2e-4) if out_f32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.float(), ref, **tol)


@pytest.mark.parametrize("kw", [
    dict(env="CartPole-v1", algo="dqn", hidden=(512, 512), mem_kind=0, n_envs=8, ring_rows=4, batch_size=4, warmup_size=4),
    dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=8, ring_rows=4,
         batch_size=4, warmup_size=4),
    dict(env="Pendulum-v1", algo="rainbow", hidden=(256, 128), dueling="max", noisy=False, mem_kind=0, n_envs=8, ring_rows=4, batch_size=4,
         warmup_size=4),
    dict(env="Grid", algo="dqn", hidden=(64, 64), mem_kind=0, n_envs=8, ring_rows=4, batch_size=4, warmup_size=4),
], ids=["dqn_512x512", "rainbow_default_noisy_dueling", "pendulum_duelmax_256x128", "grid_64x64"])
def test_pred_q_tc_agrees_with_the_fp32_seam(kw):
    """RLParameter.pred_q on the tensor cores vs the oracle's torch fp32 forward (oracle/nets.py, pinned by the trainer goldens) on
    20 000 states: bf16 operands carry 8 bits of mantissa, so the bar is 2e-2 of the Q scale (the fp32 seam srlx_qnet_forward is held
    to 1e-4 against the same oracle; a 512 x 512 layer does not fit its shared-memory-resident weights at all)."""
    from oracle import nets as onets
    from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig

    dev = DeviceEngine(EngineConfig(**kw, seed=4))
    x = np.random.default_rng(0).normal(size=(20_000, dev.D)).astype(np.float32)
    spec = onets.NetSpec(dev.D, tuple(dev.cfg.hidden), dev.A, dev.cfg.dueling, dev.cfg.noisy)
    noise = dev.noise(3, 9) if dev.cfg.noisy else None

    def ref(target, xs):
        mu, sg = dev.get_target() if target else dev.get_params()
        return onets.np_forward(spec, mu, sg, noise, xs)

    dev.pred_q = lambda xs, target=False, noise_call_id=0: ref(target, xs)
    want = dev.pred_q(x, noise_call_id=9)
    got = dev.pred_q_tc(x, noise_call_id=9).cpu().numpy()
    scale = np.abs(want).max()
    assert np.isfinite(got).all() and scale > 0
    assert np.abs(got - want).max() <= 2e-2 * scale, (np.abs(got - want).max(), scale)
    want_t = dev.pred_q(x[:1000], target=True, noise_call_id=9)
    got_t = dev.pred_q_tc(x[:1000], target=True, noise_call_id=9).cpu().numpy()
    assert np.abs(got_t - want_t).max() <= 2e-2 * np.abs(want_t).max()
    # the greedy action agrees except where the two best Q values are within the bf16 error
    top2 = np.sort(want, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4e-2 * scale
    assert (np.argmax(got, axis=1)[clear] == np.argmax(want, axis=1)[clear]).all()
