#!/usr/bin/env python
"""bench.py -- env-steps/s + trainer-updates/s of the actor/replay/learner hot path (BASELINE.json metric).

Workload (configs[2], the configuration the metric is quoted on; fits one GPU):
  Rainbow as the reference implements it (DoubleDQN + dueling(512,) + NoisyNet + 3-step Retrace + proportional PER;
  srl/algorithms/rainbow/rainbow.py:57-108) on CartPole-v1, 8192 vectorised envs per GPU, SumTree replay of 2M
  transitions per GPU (ring 256 rows x 8192 envs), batch 32, lr 1e-3, target sync every 1000 updates.
One bench "step" = S = --vec-steps-per-step (default 16) consecutive passes of the hot path, each pass ONE vector step of all
E envs (E env steps: policy forward, env.step, ring write, replay add) followed by E/train_interval trainer updates (RunContext.train_interval, srl/base/context.py:60; default here 10, i.e.
one Trainer.train() per 10 env steps -- the ratio the two north-star targets, >= 1M env-steps/s and >= 100k updates/s,
imply; the reference's own default of 1 makes env-steps/s == updates/s) -- all enqueued on one CUDA stream with no host round trip in between.

  python bench.py [--gpus N] [--steps K] [--warmup W]            own arm (CUDA, libsrlx.so)
  python bench.py --impl reference [...]                          the CPU path: the UNMODIFIED reference (srl.Runner.train from
                                                                  baseline/_ref, one process per host core) when that install is
                                                                  present, else the oracle port of the loop (oracle/engine.py)
N > 1: launched under torchrun, one rank per GPU; each rank owns its env shard + replay shard (weak scaling);
see DESIGN.md "Multi-GPU" for what is exchanged.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def rainbow_kwargs(n_envs, ring_rows, warmup_size, seed):
    return dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1,
                multisteps=3, n_envs=n_envs, ring_rows=ring_rows, batch_size=32, warmup_size=warmup_size, seed=seed,
                enable_double_dqn=True, target_update_interval=1000, lr=1e-3, discount=0.99)


def dqn_kwargs(n_envs, ring_rows, warmup_size, seed):
    """BASELINE configs[1]: DQN (double, no dueling / noisy / n-step), MLP[64,64], uniform replay."""
    return dict(env="CartPole-v1", algo="dqn", hidden=(64, 64), dueling=None, noisy=False, mem_kind=0, multisteps=1,
                n_envs=n_envs, ring_rows=ring_rows, batch_size=32, warmup_size=warmup_size, seed=seed, epsilon=0.1,
                enable_double_dqn=True, target_update_interval=1000, lr=1e-3, discount=0.99)


def dqn_default_kwargs(n_envs, ring_rows, warmup_size, seed):
    """The reference's default DQN: one hidden layer of 512 (MLPBlockConfig default), double DQN, uniform ReplayBuffer."""
    kw = dqn_kwargs(n_envs, ring_rows, warmup_size, seed)
    kw["hidden"] = (512,)
    return kw


WORKLOADS = {"rainbow": rainbow_kwargs, "dqn": dqn_kwargs, "dqn_default": dqn_default_kwargs}


def workload_kwargs(args, **kw):
    return WORKLOADS[args.workload](**kw)


def workload_config(args, world):
    name = {"dqn": "DQN(double) MLP[64,64] uniform replay CartPole-v1 (BASELINE configs[1]; not the headline config)",
            "dqn_default": "DQN(double) MLP[512] uniform replay CartPole-v1 (the reference's default DQN config; not the headline config)",
            "rainbow": "Rainbow(double+dueling512+noisy+3step-retrace+PER) CartPole-v1 (BASELINE configs[2])"}[args.workload]
    return {"workload": name,
            "n_envs_per_gpu": args.envs, "replay_capacity_per_gpu": args.envs * args.ring_rows, "batch_size": 32,
            "multisteps": 3 if args.workload == "rainbow" else 1, "train_interval": args.train_interval,
            "vec_steps_per_step": args.vec_steps_per_step,
            "updates_per_step_per_gpu": (args.envs // args.train_interval) * args.vec_steps_per_step,
            "env_steps_per_step": args.envs * world * args.vec_steps_per_step,
            "parallelism": (f"shard{world}: env/replay/SumTree shards + learner replica per GPU, parameters averaged by one "
                            f"NCCL all-reduce per step (the single-learner mode over the same shards is measured in the same run: key "
                            f"single_learner)") if world > 1 else "single",
            "l2": "flushed between timed steps (256 MiB write), device-timed and e2e alike"}


def algorithmic_bytes_per_update(n_params_total, batch=32, multisteps=3, depth=21):
    """gather + weights fwd/bwd (5 passes) + Adam (7 x 4 B per parameter) + SumTree sample/update
    (depth 0 = uniform replay: no tree)."""
    gather = batch * multisteps * 44
    weights = 5 * 4 * n_params_total
    adam = 7 * 4 * n_params_total
    tree = batch * depth * 8 + batch * depth * 16
    return gather + weights + adam + tree + 4 * batch * 4


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid, self.proc, self.path = uuid, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
            time.sleep(0.35)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU path -- used by cpu_baseline and by --impl reference
#   kind "reference": the UNMODIFIED reference, srl.Runner(...).train() = core_play.play (srl/base/run/core_play.py:115-214), from
#                     the offline install under baseline/_ref (python -m pip install --no-index --no-deps --target baseline/_ref
#                     <copy of /root/reference>; __graft_entry__.build() makes it when /root/reference is present), on the CPU
#                     restatement of CartPole-v1 registered with the reference's env registry (oracle/ref_envs.py; gymnasium is absent)
#   kind "port":      the oracle's sequential port of the vectorised loop (oracle/engine.py) when that install is missing
# ---------------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "srl", "__init__.py"))


def cpu_reference_run(train_interval, env_steps, warmup_env_steps, threads, workload="rainbow", capacity=2_000_000):
    """One process of the unmodified reference: srl.Runner("CartPole-v1", <workload config>).train(max_steps=env_steps,
    train_interval=train_interval) on one env (the reference's loop is one env per process), memory.capacity as the workload
    says, after a warm-up train() that fills the memory past warmup_size.  Rates as print_progress.py:224-237 defines them."""
    import torch

    torch.set_num_threads(max(1, threads))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import srl
    from srl.algorithms import dqn, rainbow

    from oracle.ref_envs import register_restated_envs

    try:
        register_restated_envs()
    except AssertionError:
        pass  # already registered in this process
    if workload == "rainbow":
        c = rainbow.Config(multisteps=3, enable_noisy_dense=True, enable_double_dqn=True, batch_size=32, lr=1e-3, discount=0.99,
                           target_model_update_interval=1000)  # hidden block default: dueling (512,)
        c.memory.set_proportional()
    else:
        c = dqn.Config(batch_size=32, lr=1e-3, discount=0.99, target_model_update_interval=1000, enable_double_dqn=True, epsilon=0.1)
        c.hidden_block.set((64, 64) if workload == "dqn" else (512,))
        c.memory.set_replay_buffer()
    c.framework = "torch"
    c.memory.capacity = int(capacity)
    c.memory.warmup_size = 1000
    c.memory.compress = False
    runner = srl.Runner("CartPole-v1", c)
    runner.set_device("CPU")
    runner.set_seed(1 + os.getpid() % 1000)
    runner.train(max_steps=max(1200, warmup_env_steps), train_interval=train_interval, enable_progress=False)
    t0 = time.perf_counter()
    st = runner.train(max_steps=env_steps, train_interval=train_interval, enable_progress=False)
    dt = time.perf_counter() - t0
    return dict(env_steps_per_s=st.total_step / dt, updates_per_s=st.train_count / dt, env_steps=int(st.total_step), seconds=dt,
                sample=f"srl.Runner('CartPole-v1', {workload} config).train(max_steps={env_steps}, train_interval={train_interval}): "
                       f"1 env per process (the reference's loop), memory.capacity {capacity}, batch 32, device CPU, torch threads="
                       f"{threads}, CartPole-v1 = CPU restatement registered with the reference (gymnasium absent)")


def cpu_port_run(n_envs, train_interval, steps, warmup, budget_s, threads, workload="rainbow"):
    """Times `steps` steps of the sequential CPU port (oracle/engine.py) on a bounded sample of the workload:
    n_envs env copies instead of 8192, same network / algorithm / train_interval."""
    import torch

    from oracle import engine as oeng
    from simple_distributed_rl_b200.netspec import NetSpec

    torch.set_num_threads(max(1, threads))
    U = max(1, n_envs // train_interval)
    kw = WORKLOADS[workload](n_envs, ring_rows=64, warmup_size=n_envs, seed=1)
    spec = NetSpec(4, kw["hidden"], 2, kw["dueling"], kw["noisy"], kw["algo"])
    mu, sigma = spec.init_params(0)
    orc = oeng.OracleEngine(oeng.EngineConfig(**kw), mu, sigma)
    for _ in range(3):  # prefill so the first update has M-step windows to sample
        orc.vec_step()
    for _ in range(warmup):
        orc.vec_step()
        orc.learn(U)
    t0 = time.perf_counter()
    done, tc0 = 0, orc.train_count
    for _ in range(steps):
        orc.vec_step()
        orc.learn(U)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(env_steps_per_s=done * n_envs / dt, updates_per_s=(orc.train_count - tc0) / dt, env_steps=done * n_envs, seconds=dt,
                sample=f"{done} passes x ({n_envs} env copies + {U} updates), replay capacity {64 * n_envs}, same net/algorithm/"
                       f"train_interval, sequential CPU port (oracle/engine.py), torch threads={threads}")


def _cpu_worker(a):
    kind, a = a[0], a[1:]
    return cpu_reference_run(*a) if kind == "reference" else cpu_port_run(*a)


def reference_arm(args):
    """All host cores: one sequential replica of the loop per core (the same sharding the GPU arm uses across ranks; the
    reference's own multi-core mode, train_mp, likewise runs one sequential actor loop per process,
    srl/base/run/play_mp.py:539-552).  One bench step = --ref-env-steps-per-step env steps in every process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = min(os.cpu_count() or 1, 64)
    K = max(1, args.steps)
    if reference_available() and not args.force_port:
        kind = "reference"
        a = (kind, args.train_interval, K * args.ref_env_steps_per_step, max(1, args.warmup) * args.ref_env_steps_per_step, 1,
             args.workload, args.envs * args.ring_rows)
    else:
        kind = "port"
        a = (kind, args.cpu_envs, args.train_interval, K * max(1, args.ref_env_steps_per_step // args.cpu_envs), 1, 150.0, 1, args.workload)
    with mp.get_context("spawn").Pool(cores) as pool:
        rs = pool.map(_cpu_worker, [a] * cores)
    secs = max(x["seconds"] for x in rs)
    env_steps = sum(x["env_steps"] for x in rs)
    value = env_steps / secs
    upd = sum(x["updates_per_s"] * x["seconds"] for x in rs) / secs
    sample = f"{cores} processes x [" + rs[0]["sample"] + "]"
    cfg = workload_config(args, 1)
    # the workload keys are the own arm's; "cpu_arm" says what the CPU processes actually ran (a bounded sample of that workload)
    cfg["cpu_arm"] = {"kind": kind, "processes": cores, "n_envs_per_process": 1 if kind == "reference" else args.cpu_envs,
                      "replay_capacity_per_process": args.envs * args.ring_rows if kind == "reference" else 64 * args.cpu_envs,
                      "env_steps_per_step_all_processes": env_steps // K, "sample": sample}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": max(1, args.warmup), "ms_per_step": 1e3 * secs / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "trainer_updates_per_sec": upd, "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------------
def own_arm(args):
    import torch
    import torch.distributed as dist

    from simple_distributed_rl_b200 import _lib, parallel
    from simple_distributed_rl_b200.engine import EngineConfig
    from simple_distributed_rl_b200.runner import VecRunner

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    E, R, TI, S = args.envs, args.ring_rows, args.train_interval, args.vec_steps_per_step
    U = E // TI
    kw = workload_kwargs(args, n_envs=E, ring_rows=R, warmup_size=1000, seed=1 + rank)
    runner = VecRunner(EngineConfig(**kw), device=dev)
    eng = runner.engine
    lib = eng.lib
    P_total = eng.spec.n_params * (2 if kw["noisy"] else 1)  # mu (+ sigma)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    sync_tensors = [eng.t["params"]] + ([eng.t["params_sigma"]] if "params_sigma" in eng.t else [])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(events=None):
        """S passes of (vector step + U updates); the replicas' online parameters are averaged once per step (N > 1)."""
        t_roll = []
        for _ in range(S):
            if events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            eng.vec_step()
            if events is not None:
                e1.record()
                t_roll.append((e0, e1))
            eng.learn(U)
        if world > 1:  # replicas' online parameters averaged over NVLink once per step (parallel.py)
            parallel.average_parameters(sync_tensors)
        return t_roll

    # fill the whole ring so sampling spans the full 2M-slot replay (steady state), then warm up
    eng.run(R, 0)
    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()

    # ---- timed region: K steps, device-timed, L2 flushed between steps ----------------------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    rolls = []
    st0 = eng.read_state()
    launches0 = lib.srlx_launch_count()
    clocks = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", ""))
    if rank == 0:
        clocks.start()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        rolls.extend(one_step(events=True))
        ev[k][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    launches = lib.srlx_launch_count() - launches0
    st1 = eng.read_state()
    t_dev = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(K))  # ms
    t_roll = sum(a.elapsed_time(b) for a, b in rolls)
    t_learn = t_dev - t_roll  # learner launches (+ the once-per-step parameter average when N > 1)
    tt = torch.tensor([t_dev, t_roll, t_learn], dtype=torch.float64, device=dev)
    cnt = torch.tensor([st1.total_step - st0.total_step, st1.train_count - st0.train_count], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_dev, t_roll, t_learn = [float(x) for x in tt.tolist()]
    env_steps, updates = [float(x) for x in cnt.tolist()]
    value = env_steps / (t_dev * 1e-3)
    upd_rate = updates / (t_dev * 1e-3)

    # ---- e2e: the same K steps through the public API (VecRunner.train), host in the loop, L2 flushed once per bench step -----
    class _Flush:
        def __init__(self):
            self.n = 0

        def on_step_end(self, context, state):
            self.n += 1
            if self.n % S == 0:
                flush.zero_()
            return False

    barrier()
    t0 = time.perf_counter()
    rs = runner.train(max_steps=K * S * E, train_interval=TI, callbacks=[_Flush()])
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    e2e_n = torch.tensor([float(rs.total_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_n, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_n.item()) / float(e2e_t.item())
    kname, cluster, smem = eng.learner_info()
    chunk = 256 if kname == "learner_fast_kernel" else U  # srlx_learn issues the fast kernel in launches of <= 256 updates
    upl = [min(chunk, U - i) for i in range(0, U, chunk)]  # updates per learner launch within one pass
    n_launch_per_pass = 2 + len(upl) * (2 if kw["noisy"] and kname == "learner_fast_kernel" else 1)
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": S * n_launch_per_pass * C.sizeof(_lib.SrlxEngine),  # the engine block rides in as kernel parameters
           "d2h_bytes_per_step": S * C.sizeof(_lib.SrlxState),
           "trainer_updates_per_sec": float(rs.train_count) * world / float(e2e_t.item()),
           "note": f"VecRunner.train(max_steps=K*S*E): per pass {n_launch_per_pass} launches + one pinned 128 B counter read + host stop "
                   "checks and callbacks; L2 flushed once per bench step; envs are generated on device by design, so there is no "
                   "bulk host input on this path"}

    # ---- the opt-in PRE-SAMPLED mode (EngineConfig.presample): batch t+1 is drawn before update t's priorities reach the tree, the
    #      staleness the reference's own memory process has in distributed mode; the SumTree chain leaves the critical path ----------
    presampled = None
    if world == 1 and kw["mem_kind"] == 1 and eng.learner_info()[0] == "learner_fast_kernel" and not args.no_presample:
        eng.c.presample = 1
        for _ in range(2):
            for _ in range(S):
                eng.vec_step()
                eng.learn(U)
        barrier()
        p0 = eng.read_state()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Kp = max(2, K // 2)
        flush.zero_()
        a.record()
        for _ in range(Kp):
            for _ in range(S):
                eng.vec_step()
                eng.learn(U)
        b.record()
        barrier()
        p1 = eng.read_state()
        pms = a.elapsed_time(b)
        pn = float(p1.train_count - p0.train_count)
        presampled = {"env_steps_per_sec": Kp * S * E / (pms * 1e-3), "trainer_updates_per_sec": pn / (pms * 1e-3),
                      "us_per_update_whole_pass": 1e3 * pms / pn, "steps": Kp,
                      "what": "EngineConfig.presample = True: inside a launch batch t+1 is sampled before update t is applied (one update "
                              "of staleness, srl/base/run/play_mp_memory.py:253-350); NOT the headline: `value` is the sequential order"}
        eng.c.presample = 0

    # ---- N > 1: the SINGLE-LEARNER mode measured next to the replicas: every update is one Trainer.train() on the global
    #      batch (world x 32 items), gradients summed over NVLink inside the learner kernel, identical Adam on every rank ----------
    single = None
    if world > 1:
        parallel.link_engine_distributed(eng, learner_seed=12345)
        for _ in range(2):
            for _ in range(S):
                eng.vec_step()
                eng.learn(U)
        barrier()
        sl0 = eng.read_state()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ks = max(2, K // 2)
        a.record()
        for _ in range(Ks):
            flush.zero_()
            for _ in range(S):
                eng.vec_step()
                eng.learn(U)
        b.record()
        barrier()
        eng.check_dp_alive()
        sl1 = eng.read_state()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        flat = torch.cat([eng.t[k].reshape(-1) for k in ("params", "target", "adam_m", "adam_v")])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([1.0 if torch.equal(ref, flat) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        n_upd = float(sl1.train_count - sl0.train_count)
        sl_ms = float(ms.item())
        single = {"env_steps_per_sec": world * Ks * S * E / (sl_ms * 1e-3), "trainer_updates_per_sec": n_upd / (sl_ms * 1e-3),
                  "global_batch": 32 * world, "items_per_sec": 32 * world * n_upd / (sl_ms * 1e-3), "us_per_update": 1e3 * sl_ms / n_upd,
                  "replica_us_per_update": 1e3 * t_learn / (updates / world),
                  "exchange_us_per_update": 1e3 * sl_ms / n_upd - 1e3 * t_dev / (updates / world),
                  "steps": Ks, "replicas_bit_identical": bool(same.item() == 1.0),
                  "what": "one trainer over all shards: per-update gradient all-reduce + global PER scalars as NVLink peer stores "
                          "inside learner_fast_kernel (words carry their own update tag), identical Adam on every rank; "
                          "exchange_us_per_update = whole-pass time per update minus the replica mode's"}
        eng.set_data_parallel(1, 0, [], 0, 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (learner: one launch = up to 256 dependent updates) ---------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (float(peaks["hbm_gbs"]), "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    bytes_per_update = algorithmic_bytes_per_update(P_total, multisteps=kw["multisteps"],
                                                    depth=(E * R - 1).bit_length() if kw["mem_kind"] else 0)
    n_launch = len(upl) * S * K                      # learner launches in the timed region
    upd_per_rank = updates / world                   # updates one rank's learner did in the timed region
    learn_ms_per_launch = t_learn / n_launch         # average launch duration (CUDA events on the launching stream)
    # algorithmic bytes per (average) launch / average launch duration == bytes_per_update * updates / learner time
    achieved = bytes_per_update * (upd_per_rank / n_launch) / (learn_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full captures
        for fn, note in (("r1_j_learner_ncu_summary.json", "256 updates per launch"), ("r1_o_learner_small_ncu_summary.json", "1024 updates per launch"),
                         ("r2_learner_ncu_summary.json", "256 updates per launch"), ("r2_learner_small_ncu_summary.json", "per launch")):
            fp = os.path.join(ROOT, "profiles", fn)
            if not os.path.exists(fp):
                continue
            for k in json.load(open(fp)):
                if kname in k["kernel"] and int(k.get("cluster", 0)) == cluster:
                    traffic, traffic_src = k["dram_bytes"], f"profiles/{fn} ({note})"
    except Exception:
        pass
    roofline = {"kernel": kname, "cluster_ctas": cluster, "smem_bytes_per_cta": smem, "bound": "hbm", "achieved": achieved,
                "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_update": bytes_per_update,
                "updates_per_launch": upl, "updates_per_launch_mean": upd_per_rank / n_launch, "launches_per_step": len(upl) * S,
                "launch_ms": learn_ms_per_launch, "learner_ms_per_step": t_learn / K, "updates_per_step": upd_per_rank / K,
                "share_of_step": t_learn / t_dev, "us_per_update": 1e3 * t_learn / upd_per_rank,
                "formula": "achieved = algorithmic_bytes_per_update * updates_per_step / learner_ms_per_step",
                "note": "consecutive updates are data-dependent (weights_t -> weights_t+1, priorities_t -> sample_t+1): the "
                        "limiter is the dependent-step latency of one 16-SM cluster, not HBM; the kernel keeps weights, Adam "
                        "state and the top of the SumTree in shared memory, so its DRAM traffic is below the algorithmic "
                        "figure (which counts 5 weight passes + Adam per update); learner time = step time minus the rollout "
                        "launches and includes the ~1% noise_precompute / tree_blk_build launches; see DESIGN.md"}
    rollout_bytes = 76 * E
    roll_ms = t_roll / (K * S)
    roofline_rollout = {"kernel": "rollout_kernel+post_step_kernel", "bound": "hbm", "achieved": rollout_bytes / (roll_ms * 1e-3) / 1e9,
                        "peak": peak_gbs, "unit": "GB/s", "frac": rollout_bytes / (roll_ms * 1e-3) / 1e9 / peak_gbs,
                        "algorithmic_bytes_per_env_step": 76, "launch_ms": roll_ms, "share_of_step": t_roll / t_dev}

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        if reference_available() and not args.force_port:
            n = max(2000, int(args.cpu_seconds * 1500))
            r = cpu_reference_run(TI, n, 1200, 1, args.workload, E * R)
            cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": 1, "kind": "reference", "sample": r["sample"],
                   "trainer_updates_per_sec": r["updates_per_s"]}
        else:
            r = cpu_port_run(args.cpu_envs, TI, steps=10_000, warmup=1, budget_s=args.cpu_seconds, threads=1, workload=args.workload)
            cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["sample"],
                   "trainer_updates_per_sec": r["updates_per_s"],
                   "note": "baseline/_ref (the reference install) is absent; BASELINE.md section 2b: the unmodified reference measured "
                           "216 updates/s per core where this port measures 100"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "trainer_updates_per_sec": upd_rate, "wall_ms_per_step": 1e3 * t_wall / K,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
            "roofline_rollout": roofline_rollout, "cpu_baseline": cpu, "single_learner": single, "presampled": presampled,
            "final_loss": float(st1.last_loss), "episodes": int(st1.episode_count),
            "mean_episode_len": float(st1.episode_len_sum) / max(1, st1.episode_count)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# --workload image: the observation pipeline of the image configs, an HBM-bound byte kernel.  A step = one pass of
# ImageProcessor.remap_observation (srl/rl/processors/image_processor.py:104-154: RGB -> gray, cv2 linear resize 210 x 160 -> 84 x 84,
# "0to1" normalisation: the InputImageBlockConfig "DQN" default, input_block.py:205) over a batch of --frames synthetic Atari frames
# (826 MB at the default 8192: larger than L2).  Metric frames/s; every rank processes its own frames (weak scaling, no collective).
# ---------------------------------------------------------------------------------------------------------------------
IMG_SRC, IMG_DST = (210, 160, 3), (84, 84)
IMG_BYTES = IMG_SRC[0] * IMG_SRC[1] * IMG_SRC[2] + IMG_DST[0] * IMG_DST[1] * 4  # algorithmic bytes per frame: uint8 in, float32 out


def image_config(args):
    return {"workload": "ImageProcessor RGB 210x160x3 uint8 -> GRAY_HW1 84x84 float32 '0to1' (InputImageBlockConfig DQN default; SURVEY 8f rank 4)",
            "frames_per_launch": args.frames, "launches_per_step": args.launches_per_step, "frames_per_step": args.frames * args.launches_per_step,
            "l2": "inputs larger than L2 (frames_per_launch x 100.8 KB in, x 28.2 KB out)"}


def _image_cpu_worker(a):
    kind, n_frames, seed = a
    import numpy as np
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(64,) + IMG_SRC, dtype=np.uint8)
    if kind == "reference":  # the reference's own class (cv2 underneath), from baseline/_ref
        sys.path.insert(0, REF_DIR)
        from srl.base.define import SpaceTypes
        from srl.base.spaces.box import BoxSpace
        from srl.rl.processors.image_processor import ImageProcessor
        sp = BoxSpace(IMG_SRC, 0, 255, np.uint8, SpaceTypes.RGB)
        proc = ImageProcessor(SpaceTypes.GRAY_HW1, (IMG_DST[1], IMG_DST[0]), normalize_type="0to1")
        ns = proc.remap_observation_space(sp)
        fn = lambda f: proc.remap_observation(f, sp, ns)  # noqa: E731
    else:
        from oracle import image as oimg
        fn = lambda f: oimg.process(f, "RGB", "GRAY_HW1", (IMG_DST[1], IMG_DST[0]), "0to1")  # noqa: E731
    fn(frames[0])
    t0 = time.perf_counter()
    for i in range(n_frames):
        fn(frames[i & 63])
    return n_frames, time.perf_counter() - t0


def image_cpu_run(n_procs, n_frames):
    kind = "reference"
    try:
        import cv2  # noqa: F401
        assert os.path.isfile(os.path.join(REF_DIR, "srl", "__init__.py"))
    except Exception:
        kind = "port"
        n_frames = max(50, n_frames // 40)  # the numpy port is ~40 x slower than cv2
    if n_procs == 1:
        res = [_image_cpu_worker((kind, n_frames, 1))]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(n_procs) as pool:
            res = pool.map(_image_cpu_worker, [(kind, n_frames, 1 + i) for i in range(n_procs)])
    total, t = sum(r[0] for r in res), max(r[1] for r in res)
    what = ("srl ImageProcessor(GRAY_HW1, (84, 84), '0to1').remap_observation (cv2 cvtColor + resize) from baseline/_ref" if kind == "reference"
            else "oracle/image.py (numpy restatement of the cv2 fixed-point algorithms; cv2 or baseline/_ref absent)")
    return {"value": total / t, "kind": kind, "cores": n_procs, "sample": f"{what}: {n_frames} synthetic 210x160x3 frames per process, {n_procs} process(es)"}


def image_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    K, W = args.steps, max(3, args.warmup)
    cores = os.cpu_count() or 1
    image_cpu_run(cores, 200 * W)
    t0 = time.perf_counter()
    r = image_cpu_run(cores, 2000 * K)
    line = {"impl": "reference", "metric": "frames_per_sec", "value": r["value"], "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": K, "warmup": W, "ms_per_step": 1e3 * (time.perf_counter() - t0) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": image_config(args),
            "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def image_arm(args):
    import torch
    import torch.distributed as dist

    from simple_distributed_rl_b200 import _lib, image

    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n, K, W, R = args.frames, args.steps, max(3, args.warmup), args.launches_per_step
    pipe = image.DeviceImagePipeline(IMG_SRC, "RGB", "GRAY_HW1", (IMG_DST[1], IMG_DST[0]), "0to1", device=str(dev))
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    frames = torch.randint(0, 256, (n,) + IMG_SRC, dtype=torch.uint8, device=dev, generator=gen)
    out = torch.empty((n,) + IMG_DST + (1,), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(W * R):
        pipe(frames, out=out)
    clk = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", ""))
    barrier()
    clk.start()
    l0 = lib.srlx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K * R):
        pipe(frames, out=out)  # inputs (826 MB) and outputs (231 MB) are larger than L2: every launch streams from HBM
    e1.record()
    barrier()
    launches = lib.srlx_launch_count() - l0
    clocks = clk.stop()
    t_dev = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev = float(t_dev[0])
    value = n * K * R * world / (t_dev * 1e-3)
    # end to end through the public API with HOST buffers: pinned uint8 frames -> device, the kernel, float32 states -> pinned host memory
    n_e = min(n, 2048)
    h_in = torch.empty((n_e,) + IMG_SRC, dtype=torch.uint8).pin_memory()
    h_in.copy_(frames[:n_e].cpu())
    h_out = torch.empty((n_e,) + IMG_DST + (1,), dtype=torch.float32).pin_memory()
    for _ in range(2):
        h_out.copy_(pipe(h_in.to(dev, non_blocking=True)), non_blocking=True)
    barrier()
    e0.record()
    for _ in range(K):
        h_out.copy_(pipe(h_in.to(dev, non_blocking=True)), non_blocking=True)
    e1.record()
    barrier()
    t_e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e = {"value": n_e * K * world / (float(t_e[0]) * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h_in.numel()), "d2h_bytes_per_step": int(h_out.numel() * 4),
           "frames_per_step": n_e, "note": "DeviceImagePipeline(frames) on pinned host uint8 frames, float32 states copied back to pinned host memory: PCIe-bound"}
    if rank == 0:
        peak, peak_src = 6543.1, "fallback"
        try:
            peak, peak_src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic, traffic_src = None, None
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", "r3u_image_process_ncu_summary.json")))[0]
            traffic, traffic_src = d["dram_bytes"] * n / 8192.0, "profiles/r3u_image_process_ncu_summary.json (8192 frames per launch)"
        except Exception:
            pass
        achieved = IMG_BYTES * n * K * R / (t_dev * 1e-3) / 1e9
        roofline = {"kernel": "image_process_staged_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_frame": IMG_BYTES,
                    "frames_per_launch": n, "launch_ms": t_dev / (K * R), "formula": "achieved = algorithmic_bytes_per_frame * frames_per_launch / launch duration"}
        cpu = None
        if not args.no_cpu_baseline:
            r = image_cpu_run(1, 20000)
            cpu = {"value": r["value"], "unit": "frames/s", "cores": 1, "kind": r["kind"], "sample": r["sample"]}
        line = {"metric": "frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_dev / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": image_config(args),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="rainbow", choices=sorted(WORKLOADS) + ["image"],
                    help="rainbow = BASELINE configs[2] (the headline, default); dqn = configs[1] (use --envs 4096) and dqn_default = "
                         "the reference's default DQN config: side measurements; image = the observation pipeline of the image configs "
                         "(SURVEY 8f rank 4; metric frames/s)")
    ap.add_argument("--frames", type=int, default=8192, help="--workload image: frames per launch")
    ap.add_argument("--launches-per-step", type=int, default=100, help="--workload image: launches in one bench step (100 x 0.27 ms: a 0.5 s timed region)")
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--ring-rows", type=int, default=256)
    ap.add_argument("--train-interval", type=int, default=10,
                    help="env steps per trainer update (RunContext.train_interval); 10 = the ratio of the two north-star targets "
                         "(>= 1M env-steps/s with >= 100k updates/s)")
    ap.add_argument("--vec-steps-per-step", type=int, default=16,
                    help="passes of the hot path (vector step + its updates) in one bench step: 16 makes 20 steps a ~2.4 s timed region")
    ap.add_argument("--ref-env-steps-per-step", type=int, default=1000,
                    help="--impl reference: env steps every CPU process does in one bench step (a bounded sample of the workload)")
    ap.add_argument("--cpu-envs", type=int, default=64, help="env copies in the bounded sample of the oracle port (fallback CPU arm)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-presample", action="store_true", help="skip the extra measurement of the opt-in pre-sampled mode")
    ap.add_argument("--force-port", action="store_true", help="CPU arms: use the oracle port even when baseline/_ref is present")
    args = ap.parse_args()
    if args.workload == "image":
        return image_reference_arm(args) if args.impl == "reference" else image_arm(args)
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
