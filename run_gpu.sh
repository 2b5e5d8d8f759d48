timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_b.json')); print(d['value'], d['trainer_updates_per_sec'], d['roofline']['us_per_update'], d['roofline_rollout']['launch_ms'], d['e2e']['value'], d['final_loss'], d['mean_episode_len'])"
tail -3 gpurun_out/bench_r1_b.err
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig
import torch
for kw in [dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=48, ring_rows=12, batch_size=32, warmup_size=96),
           dict(env="CartPole-v1", algo="rainbow", hidden=(64, 64), dueling="average", noisy=False, mem_kind=1, multisteps=3, n_envs=32, ring_rows=9, batch_size=16, warmup_size=64, has_duplicate=False, epsilon=0.25)]:
    d = DeviceEngine(EngineConfig(**kw))
    d.run(6, 3)
    torch.cuda.synchronize()
    print(d.read_state().train_count)
PY
timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py 2>&1 | tail -8
timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py 2>&1 | tail -8
