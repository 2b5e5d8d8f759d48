cat > /tmp/prof.py <<'PY'
import sys; sys.path.insert(0,'.')
from simple_distributed_rl_b200.engine import DeviceEngine, EngineConfig
import torch
kw = dict(env="CartPole-v1", algo="rainbow", hidden=(512,), dueling="average", noisy=True, mem_kind=1, multisteps=3, n_envs=8192, ring_rows=256, batch_size=32, warmup_size=1000, seed=1)
d = DeviceEngine(EngineConfig(**kw))
d.run(256, 0)
for rep in range(3):
    d.learn(128)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:learner_kernel -s 2 -c 1 -o gpurun_out/prof_learner_r1_b -f python /tmp/prof.py > gpurun_out/ncu_c.log 2>&1; tail -3 gpurun_out/ncu_c.log
