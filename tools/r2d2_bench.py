"""R2D2 on device at the shape BASELINE configs[3] names (2048 env copies, LSTM 512, burn-in 40, sequence 80, batch 64, prioritized
sequence replay; CartPole-v1 / Pendulum-v1 in place of the Box2D LunarLander): time of one vector step and of one trainer update
(eager launches and under a CUDA graph), CUDA events on the launching stream after warm-up."""
import argparse
import json
import sys, os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Engine


def timed(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def cpu_port(eng, n):
    """oracle/r2d2.py (the torch restatement of Trainer._train_on_batches + keras Adam) on the host cores, same shapes, random batch"""
    import time

    import numpy as np

    from oracle import r2d2 as orc

    c = eng.cfg
    tr = orc.Trainer(eng.get_weights(), len(c.hidden_layers) - (0 if c.dueling_type is None else 1), c.dueling_type, c.burnin,
                     c.sequence_length, c.discount, c.lr, c.target_model_update_interval, c.enable_double_dqn, c.enable_rescale,
                     c.enable_retrace, c.retrace_h)
    rng = np.random.default_rng(0)
    B, S, W1 = eng.B, eng.S, eng.W + 1
    states = rng.normal(size=(B, W1, eng.D)).astype(np.float32)
    actions = rng.integers(0, eng.A, size=(B, S)).tolist()
    probs = np.full((B, S), 0.5).tolist()
    rewards = rng.normal(size=(B, S)).tolist()
    dones = np.zeros((B, S), bool).tolist()
    h0 = np.zeros((B, eng.u), np.float32)
    times = []
    for i in range(n + 1):
        t0 = time.perf_counter()
        out = tr.train_on_batches(states, actions, probs, rewards, dones, h0, h0, np.ones(B, np.float32))
        tr.apply(out["grads"])
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times[1:]))
    return dict(kind="port", update_ms=ms, updates_per_s=1e3 / ms, cores=torch.get_num_threads(),
                sample=f"{n} updates of oracle/r2d2.py (torch fp32 CPU) on a random batch of the same shape after 1 warm-up")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="CartPole-v1")
    ap.add_argument("--n-envs", type=int, default=2048)
    ap.add_argument("--units", type=int, default=512)
    ap.add_argument("--burnin", type=int, default=40)
    ap.add_argument("--seq", type=int, default=80)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--rows", type=int, default=512)
    ap.add_argument("--memory", default="Proportional")
    ap.add_argument("--out", default="")
    ap.add_argument("--cpu-baseline", type=int, default=0, help="time N updates of the CPU port (oracle/r2d2.py, torch fp32, all host threads) "
                    "on a batch of the same shape: the reference's own R2D2 is TensorFlow-only and cannot run here")
    a = ap.parse_args()
    cfg = R2D2Config(env=a.env, n_envs=a.n_envs, lstm_units=a.units, hidden_layers=(512,), dueling_type="average", burnin=a.burnin,
                     sequence_length=a.seq, batch_size=a.batch, capacity=a.n_envs * a.rows, warmup_size=a.n_envs * 4, memory=a.memory,
                     enable_rescale=True, enable_retrace=False, lr=1e-4, target_model_update_interval=2500)
    eng = R2D2Engine(cfg)
    for _ in range(a.burnin + a.seq + 8):
        eng.vec_step(True)
    eng.learn(2)
    torch.cuda.synchronize()
    l0 = eng.lib.srlx_launch_count()
    eng.learn(1)
    launches = eng.lib.srlx_launch_count() - l0
    step_ms = timed(lambda: eng.vec_step(True), 50)
    learn_ms = timed(lambda: eng.learn(1), 20)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.learn(1)
    graph_ms = timed(g.replay, 20)
    st = eng.read_state()
    P = eng.spec.n_params
    W, B, u, K, S = eng.W, eng.B, eng.u, eng.K, eng.S
    # fp32 FLOPs of one update: LSTM forward (W + 1 steps x 2 networks), BPTT (dh + dW over S steps), head forward / backward
    lstm_f = 2.0 * B * K * 4 * u
    flops = 2 * (W + 1) * lstm_f + S * (2.0 * B * 4 * u * u) + 2.0 * S * B * 4 * u * K
    out = dict(config=dict(env=a.env, n_envs=a.n_envs, lstm_units=a.units, burnin=a.burnin, sequence_length=a.seq, batch_size=a.batch,
                           ring_rows=eng.R, memory=a.memory, n_params=P),
               vec_step_ms=step_ms, env_steps_per_s=a.n_envs / step_ms * 1e3, update_ms_eager=learn_ms, update_ms_graph=graph_ms,
               updates_per_s_graph=1e3 / graph_ms, launches_per_update=int(launches), lstm_tflops_graph=flops / graph_ms / 1e9,
               train_count=int(st.train_count), loss=st.last_loss, mem_size=int(st.mem_size))
    if a.cpu_baseline > 0:
        out["cpu_baseline"] = cpu_port(eng, a.cpu_baseline)
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
