"""simple_distributed_rl_b200 -- B200-native rollout / replay / PER / TD-update engine behind the SRL plugin API.

Only the hot path of pocokhc/simple_distributed_rl lives here:
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/srlx.h) -> libsrlx.so
  _lib.py      ctypes binding of the C ABI (fails loudly when the library is missing)
  netspec.py   Q-network description / flat parameter layout <-> reference state_dict keys
  envspec.py   closed-form environment tables (Grid, CartPole)
  engine.py    DeviceEngine: owns the HBM buffers (torch CUDA tensors) and drives the kernels
  memory.py    DeviceProportionalMemory: IPriorityMemory-compatible seam over the device SumTree
  runner.py    VecRunner: srl.Runner-like facade (train / evaluate / RunState counters / callbacks)
  srl_plugin.py registration of the device classes with an installed `srl`
"""
__version__ = "0.1.0"
