"""CPU tests of the R2D2 host side and its oracle (no GPU): the keras-LSTM restatement of oracle/r2d2.py against torch.nn.LSTM (an
independent implementation with the same gate order i, f, g, o), the flat parameter layout against keras weight lists, the worker's
recent_* lists, keras Huber, keras Adam against torch.optim.Adam (which keras' formula equals up to where epsilon sits)."""
import numpy as np
import pytest
import torch

from oracle import r2d2 as orc
from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2NetSpec


def _weights(D, u, hidden, dueling, A, seed=0):
    spec = R2D2NetSpec(D, u, hidden, dueling, A)
    w = spec.init_keras(seed)
    rng = np.random.default_rng(seed)
    return spec, [x + rng.normal(0, 0.05, x.shape).astype(np.float32) for x in w]  # biases non-zero


def test_oracle_lstm_equals_torch_nn_lstm():
    D, u, B, T = 3, 8, 5, 7
    spec, w = _weights(D, u, (6,), None, 4)
    net = orc.QNet(w, 1, None)
    ref = torch.nn.LSTM(D, u, batch_first=True)
    with torch.no_grad():  # torch: weight_ih [4u][D] rows gate-major (i, f, g, o), bias_ih + bias_hh
        ref.weight_ih_l0.copy_(torch.as_tensor(w[0]).T)
        ref.weight_hh_l0.copy_(torch.as_tensor(w[1]).T)
        ref.bias_ih_l0.copy_(torch.as_tensor(w[2]))
        ref.bias_hh_l0.zero_()
    g = torch.Generator().manual_seed(1)
    x, h0, c0 = torch.randn(B, T, D, generator=g), torch.randn(B, u, generator=g), torch.randn(B, u, generator=g)
    with torch.no_grad():
        out_ref, (hn, cn) = ref(x, (h0[None], c0[None]))
        h, c, outs = h0, c0, []
        for t in range(T):
            h, c = net.step(x[:, t], h, c)
            outs.append(h)
    np.testing.assert_allclose(torch.stack(outs, 1).numpy(), out_ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c.numpy(), cn[0].numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("hidden,dueling", [((16, 16), None), ((12,), "average"), ((8, 12), "max"), ((12,), "")])
def test_flat_layout_roundtrip_and_forward(hidden, dueling):
    """from_keras / to_keras are inverse; the flat layout ([x | h | 1] . W^T with rows unit * 4 + gate, bias as last column, dueling
    block as one wide layer with structural zeros) computes what the keras-order oracle computes."""
    D, u, A, B = 3, 8, 5, 6
    spec, w = _weights(D, u, hidden, dueling, A, seed=3)
    p = spec.from_keras(w)
    for a, b in zip(spec.to_keras(p), w):
        np.testing.assert_array_equal(a, b)
    net = orc.QNet(w, len(hidden) if dueling is None else len(hidden) - 1, dueling)
    rng = np.random.default_rng(0)
    x, h, c = (rng.normal(size=s).astype(np.float32) for s in ((B, D), (B, u), (B, u)))
    with torch.no_grad():
        h2, c2 = net.step(torch.as_tensor(x), torch.as_tensor(h), torch.as_tensor(c))
        q = net.head(h2).numpy()
    # numpy evaluation straight from the flat buffer
    K = spec.K
    Wl = p[:4 * u * K].reshape(u, 4, K)
    z = np.einsum("bk,ugk->bug", np.concatenate([x, h, np.ones((B, 1), np.float32)], 1), Wl)
    sig = lambda v: 1 / (1 + np.exp(-v))  # noqa: E731
    cf = sig(z[:, :, 1]) * c + sig(z[:, :, 0]) * np.tanh(z[:, :, 2])
    hf = sig(z[:, :, 3]) * np.tanh(cf)
    np.testing.assert_allclose(hf, h2.numpy(), rtol=1e-5, atol=1e-6)
    a = hf
    for l, (out, k, off) in enumerate(spec.head):
        Wd = p[off:off + out * (k + 1)].reshape(out, k + 1)
        a = np.concatenate([a, np.ones((B, 1), np.float32)], 1) @ Wd.T
        if l < len(spec.head) - 1:
            a = np.maximum(a, 0)
    if dueling is None:
        qf = a
    else:
        v, adv = a[:, :1], a[:, 1:]
        qf = v + adv - (adv.mean(1, keepdims=True) if dueling == "average" else adv.max(1, keepdims=True) if dueling == "max" else 0)
        out, k, off = spec.head[-1]
        assert np.all(p[off:off + out * (k + 1)].reshape(out, k + 1)[spec.zero_mask] == 0)
    np.testing.assert_allclose(qf, q, rtol=1e-4, atol=1e-5)


def test_worker_lists_follow_the_reference_shapes_and_padding():
    """r2d2.py:221-303: every step adds one item of burnin + S + 1 states; an episode's end adds S - 1 padded items whose newest
    entries are (dummy state, random action, 1/A, reward 0, done True); the hidden state of an item belongs to its first state."""
    bi, S, D, u, A = 2, 3, 2, 4, 3
    w = orc.WorkerLists(bi, S, D, u, A, rand_action=lambda kind, j: {"reset": 1, "tail": 2}[kind])
    w.on_reset(np.array([1.0, 1.0], np.float32))
    for t in range(4):
        hid = (np.full(u, t + 1, np.float32), np.full(u, -(t + 1), np.float32))
        w.on_step(t % A, 0.5, float(t), terminated=(t == 3), done=(t == 3), next_state=np.full(D, t + 2, np.float32), hidden_after=hid)
    assert len(w.items) == 4 + (S - 1)
    it = w.items[0]
    assert it["states"].shape == (bi + S + 1, D) and len(it["actions"]) == S
    assert np.all(it["states"][:-2] == 0) and np.all(it["states"][-2] == 1) and np.all(it["states"][-1] == 2)
    assert it["actions"] == [1, 1, 0] and it["probs"][:2] == [1 / 3, 1 / 3] and it["dones"] == [False] * 3
    assert np.all(it["hidden_states"][0] == 0)
    last = w.items[-1]
    assert last["actions"] == [0, 2, 2] and last["rewards"] == [3.0, 0.0, 0.0] and last["dones"] == [True, True, True]
    assert np.all(last["states"][-1] == 0) and np.all(last["states"][-2] == 0) and np.all(last["states"][-3] == 5)
    # states[0] of the last item is the state of step 1 (value 2); its hidden state is the one after step 0
    assert np.all(last["states"][0] == 2) and np.all(last["hidden_states"][0] == 1) and np.all(last["hidden_states"][1] == -1)


def test_keras_huber_and_adam_restatements():
    y, p = torch.tensor([[0.0, 2.0], [1.0, -3.0]]), torch.tensor([[0.5, 0.0], [1.0, 0.0]])
    ref = torch.nn.functional.huber_loss(p, y, delta=1.0)
    assert abs(float(orc.huber_keras(y, p)) - float(ref)) < 1e-7
    # keras Adam == torch Adam when epsilon is negligible against sqrt(v)
    spec, w = _weights(2, 3, (4,), None, 2)
    tr = orc.Trainer(w, 1, None, 0, 1, 0.99, 1e-3, 10, True, False, False, 1.0)
    ref_w = [torch.tensor(x.copy(), requires_grad=True) for x in w]
    opt = torch.optim.Adam(ref_w, lr=1e-3, eps=1e-7)
    rng = np.random.default_rng(0)
    for _ in range(3):
        grads = [rng.normal(size=x.shape).astype(np.float32) for x in w]
        for t, g in zip(ref_w, grads):
            t.grad = torch.as_tensor(g)
        opt.step()
        tr.apply(grads)
    for a, b in zip(tr.weights(), ref_w):
        np.testing.assert_allclose(a, b.detach().numpy(), rtol=1e-5, atol=1e-7)
    assert tr.sync_count == 1 and tr.train_count == 3


def test_config_defaults_are_the_reference_defaults(srl_mod):
    """R2D2Config restates srl/algorithms/r2d2/config.py (importing the config module needs no TensorFlow)."""
    from srl.algorithms.r2d2.config import Config

    ref, own = Config(), R2D2Config()
    for k in ("test_epsilon", "epsilon", "batch_size", "lstm_units", "burnin", "sequence_length", "discount", "lr",
              "target_model_update_interval", "enable_double_dqn", "enable_rescale", "enable_retrace", "retrace_h"):
        assert getattr(ref, k) == getattr(own, k), k
    assert ref.memory.capacity == own.capacity and ref.memory.warmup_size == own.warmup_size and ref.memory.name == own.memory
    assert ref.hidden_block.name == "DuelingNetwork" and tuple(ref.hidden_block.kwargs["layer_sizes"]) == own.hidden_layers
    assert ref.hidden_block.kwargs["dueling_kwargs"]["dueling_type"] == own.dueling_type
    ref.set_atari_config()
    own.set_atari_config()
    for k in ("burnin", "sequence_length", "discount", "lr", "batch_size", "target_model_update_interval", "enable_rescale", "enable_retrace"):
        assert getattr(ref, k) == getattr(own, k), k
    assert ref.memory.name == own.memory and ref.memory.kwargs["alpha"] == own.per_alpha and ref.memory.kwargs["beta_initial"] == own.per_beta_initial
