#!/usr/bin/env python
"""bench.py -- env-steps/s + trainer-updates/s of the actor/replay/learner hot path (BASELINE.json metric).

Workload (configs[2], the configuration the metric is quoted on; fits one GPU):
  Rainbow as the reference implements it (DoubleDQN + dueling(512,) + NoisyNet + 3-step Retrace + proportional PER;
  srl/algorithms/rainbow/rainbow.py:57-108) on CartPole-v1, 8192 vectorised envs per GPU, SumTree replay of 2M
  transitions per GPU (ring 256 rows x 8192 envs), batch 32, lr 1e-3, target sync every 1000 updates.
One bench "step" = ONE vector step of all E envs (E env steps: policy forward, env.step, ring write, replay add)
followed by E/train_interval trainer updates (RunContext.train_interval, srl/base/context.py:60; default here 10, i.e.
one Trainer.train() per 10 env steps -- the ratio the two north-star targets, >= 1M env-steps/s and >= 100k updates/s,
imply; the reference's own default of 1 makes env-steps/s == updates/s) -- all enqueued on one CUDA stream with no host round trip in between.

  python bench.py [--gpus N] [--steps K] [--warmup W]            own arm (CUDA, libsrlx.so)
  python bench.py --impl reference [...]                          the CPU path (oracle port of the reference loop)
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
            "updates_per_step_per_gpu": args.envs // args.train_interval,
            "env_steps_per_step": args.envs * world,
            "parallelism": (f"shard{world}: env/replay/SumTree shards + learner replica per GPU, parameters averaged by one "
                            f"NCCL all-reduce per step") if world > 1 else "single",
            "l2": "flushed between timed steps (256 MiB write)"}


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
# CPU path (oracle port of the reference loop) -- used by cpu_baseline and by --impl reference
# ---------------------------------------------------------------------------------------------------------------------
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
    return dict(env_steps_per_s=done * n_envs / dt, updates_per_s=(orc.train_count - tc0) / dt, steps=done, seconds=dt,
                sample=f"{done} steps x ({n_envs} env copies + {U} updates), same net/algorithm/train_interval, "
                       f"sequential CPU port (oracle/engine.py), torch threads={threads}")


def _cpu_worker(a):
    return cpu_port_run(*a)


def reference_arm(args):
    """All host cores: one sequential replica of the loop per core, each on its own shard of env copies (the same
    sharding the GPU arm uses across ranks; the reference's own multi-core mode, train_mp, likewise runs one
    sequential actor loop per process, srl/base/run/play_mp.py:539-552)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = min(os.cpu_count() or 1, 64)
    a = (args.cpu_envs, args.train_interval, args.steps, min(args.warmup, 1), 150.0, 1, args.workload)
    with mp.get_context("spawn").Pool(cores) as pool:
        rs = pool.map(_cpu_worker, [a] * cores)
    r = dict(env_steps_per_s=sum(x["env_steps_per_s"] for x in rs), updates_per_s=sum(x["updates_per_s"] for x in rs),
             steps=min(x["steps"] for x in rs), seconds=max(x["seconds"] for x in rs),
             sample=f"{cores} processes x [" + rs[0]["sample"] + "]")
    line = {"impl": "reference", "metric": METRIC, "value": r["env_steps_per_s"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * r["seconds"] / max(1, r["steps"]),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "trainer_updates_per_sec": r["updates_per_s"],
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": r["env_steps_per_s"], "unit": UNIT, "cores": cores, "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["env_steps_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    E, R, TI = args.envs, args.ring_rows, args.train_interval
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

    # fill the whole ring so sampling spans the full 2M-slot replay (steady state), then warm up
    eng.run(R, 0)
    for _ in range(max(3, args.warmup)):
        eng.vec_step()
        eng.learn(U)
        if world > 1:
            parallel.average_parameters(sync_tensors)
    barrier()

    # ---- timed region: K steps, device-timed per phase, L2 flushed between steps ------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
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
        eng.vec_step()
        ev[k][1].record()
        eng.learn(U)
        if world > 1:  # replicas' online parameters averaged over NVLink once per step (parallel.py)
            parallel.average_parameters(sync_tensors)
        ev[k][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if rank == 0 else None
    launches = lib.srlx_launch_count() - launches0
    st1 = eng.read_state()
    t_roll = sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(K))  # ms
    t_learn = sum(ev[k][1].elapsed_time(ev[k][2]) for k in range(K))
    t_dev = t_roll + t_learn
    tt = torch.tensor([t_dev, t_roll, t_learn], dtype=torch.float64, device=dev)
    cnt = torch.tensor([st1.total_step - st0.total_step, st1.train_count - st0.train_count], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_dev, t_roll, t_learn = [float(x) for x in tt.tolist()]
    env_steps, updates = [float(x) for x in cnt.tolist()]
    value = env_steps / (t_dev * 1e-3)
    upd_rate = updates / (t_dev * 1e-3)

    # ---- e2e: the same K steps through the public API (VecRunner.train), host in the loop ----------------------
    barrier()
    t0 = time.perf_counter()
    rs = runner.train(max_steps=K * E, train_interval=TI)
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    e2e_n = torch.tensor([float(rs.total_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_n, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_n.item()) / float(e2e_t.item())
    n_launch_per_step = 3
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": n_launch_per_step * C.sizeof(_lib.SrlxEngine),  # the engine block rides in as kernel parameters
           "d2h_bytes_per_step": C.sizeof(_lib.SrlxState),
           "trainer_updates_per_sec": float(rs.train_count) * world / float(e2e_t.item()),
           "note": "VecRunner.train(max_steps=K*E): per step 3 launches + one pinned 128 B counter read + host stop checks; "
                   "envs are generated on device by design, so there is no bulk host input on this path"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (learner_kernel: one launch = U dependent updates) ---------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (float(peaks["hbm_gbs"]), "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    bytes_per_update = algorithmic_bytes_per_update(P_total, multisteps=kw["multisteps"],
                                                    depth=(E * R - 1).bit_length() if kw["mem_kind"] else 0)
    kname, cluster, smem = eng.learner_info()
    chunk = 256 if kname == "learner_fast_kernel" else U  # srlx_learn issues the fast kernel in launches of <= 256 updates
    n_launch = (U + chunk - 1) // chunk
    learn_ms_per_launch = t_learn / (K * n_launch)
    achieved = bytes_per_update * min(U, chunk) / (learn_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full captures
        for fn, note in (("r1_j_learner_ncu_summary.json", "256 updates per launch"), ("r1_o_learner_small_ncu_summary.json", "1024 updates per launch")):
            for k in json.load(open(os.path.join(ROOT, "profiles", fn))):
                if kname in k["kernel"] and int(k.get("cluster", 0)) == cluster:
                    traffic, traffic_src = k["dram_bytes"], f"profiles/{fn} ({note})"
    except Exception:
        pass
    roofline = {"kernel": kname, "cluster_ctas": cluster, "smem_bytes_per_cta": smem, "bound": "hbm", "achieved": achieved,
                "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_update": bytes_per_update, "updates_per_launch": min(U, chunk),
                "launches_per_step": n_launch, "launch_ms": learn_ms_per_launch, "share_of_step": t_learn / t_dev,
                "us_per_update": 1e3 * t_learn / (K * U),
                "note": "consecutive updates are data-dependent (weights_t -> weights_t+1, priorities_t -> sample_t+1): the "
                        "limiter is the dependent-step latency of one 16-SM cluster, not HBM; the kernel keeps weights, Adam "
                        "state and the top of the SumTree in shared memory, so its DRAM traffic is below the algorithmic "
                        "figure (which counts 5 weight passes + Adam per update); launch_ms includes the ~1% noise_precompute "
                        "launches; see DESIGN.md"}
    rollout_bytes = 76 * E
    roll_ms = t_roll / K
    roofline_rollout = {"kernel": "rollout_kernel+post_step_kernel", "bound": "hbm", "achieved": rollout_bytes / (roll_ms * 1e-3) / 1e9,
                        "peak": peak_gbs, "unit": "GB/s", "frac": rollout_bytes / (roll_ms * 1e-3) / 1e9 / peak_gbs,
                        "algorithmic_bytes_per_env_step": 76, "launch_ms": roll_ms, "share_of_step": t_roll / t_dev}

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(args.cpu_envs, TI, steps=10_000, warmup=1, budget_s=args.cpu_seconds, threads=1, workload=args.workload)
        cpu = {"value": r["env_steps_per_s"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["sample"],
               "trainer_updates_per_sec": r["updates_per_s"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "trainer_updates_per_sec": upd_rate, "wall_ms_per_step": 1e3 * t_wall / K,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
            "roofline_rollout": roofline_rollout, "cpu_baseline": cpu,
            "final_loss": float(st1.last_loss), "episodes": int(st1.episode_count),
            "mean_episode_len": float(st1.episode_len_sum) / max(1, st1.episode_count)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="rainbow", choices=sorted(WORKLOADS),
                    help="rainbow = BASELINE configs[2] (the headline, default); dqn = configs[1] (use --envs 4096) and dqn_default = "
                         "the reference's default DQN config: side measurements")
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--ring-rows", type=int, default=256)
    ap.add_argument("--train-interval", type=int, default=10,
                    help="env steps per trainer update (RunContext.train_interval); 10 = the ratio of the two north-star targets "
                         "(>= 1M env-steps/s with >= 100k updates/s)")
    ap.add_argument("--cpu-envs", type=int, default=64, help="env copies in the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
