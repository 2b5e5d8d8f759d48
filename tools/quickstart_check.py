"""runs the README quick start with small numbers (GPU box; the reference from baseline/_ref or /root/reference)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
    if os.path.isfile(os.path.join(p, "srl", "__init__.py")):
        sys.path.insert(0, p)
        break
import srl
from srl.algorithms import rainbow
from simple_distributed_rl_b200 import srl_classes
srl_classes.register()
cfg = rainbow.Config(multisteps=3, enable_noisy_dense=True); cfg.memory.set_proportional()
cfg.memory.warmup_size = 64
runner = srl.Runner("CartPole-v1", cfg)
runner.train(max_train_count=200); print("1a", runner.evaluate(max_episodes=2))
state = srl_classes.train_vectorized(runner, num_envs=1024, max_steps=1024 * 300, train_interval=10)
print("1b", state.total_step, state.train_count, runner.evaluate(max_episodes=2))
from simple_distributed_rl_b200.srl_plugin import DeviceRunner
dev = DeviceRunner("CartPole-v1", cfg, num_envs=1024)
dev.train(max_steps=1024 * 300, train_interval=10); print("2", sum(dev.evaluate(max_episodes=20)) / 20)
srl_classes.unregister()
from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Runner
r = R2D2Runner(R2D2Config(env="Pendulum-v1", n_envs=64, lstm_units=64, hidden_layers=(64,), memory="Proportional"))
r.train(max_train_count=500); print("3a", r.evaluate(max_episodes=3))
from simple_distributed_rl_b200.ppo import PPOConfig, PPORunner
p = PPORunner(PPOConfig(env="Pendulum-v1", n_envs=1024, horizon=200, gae_discount=0.95))
p.train(max_rollouts=2); print("3b", p.evaluate(max_episodes=3))
