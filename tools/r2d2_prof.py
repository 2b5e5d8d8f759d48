"""one R2D2 update + one vector step at the configs[3] shape, for an ncu launch list"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simple_distributed_rl_b200.r2d2 import R2D2Config, R2D2Engine

cfg = R2D2Config(env="CartPole-v1", n_envs=2048, lstm_units=512, hidden_layers=(512,), dueling_type="average", burnin=40, sequence_length=80,
                 batch_size=64, capacity=2048 * 256, warmup_size=2048 * 4, memory="Proportional", enable_rescale=True, enable_retrace=False)
eng = R2D2Engine(cfg)
for _ in range(130):
    eng.vec_step(True)
eng.learn(2)
torch.cuda.synchronize()
