"""Loader of tests/golden/learner_*.npz (the reference's Trainer.train + memory sample/update on a frozen replay memory,
tests/golden/make_learner_golden.py) shared by the CPU oracle test and the GPU device test."""
import ast
import glob
import os
import sys

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)

from synth_ring import synth_ring  # noqa: E402

from simple_distributed_rl_b200 import checkpoint as ck  # noqa: E402

PATHS = sorted(glob.glob(os.path.join(GOLDEN, "learner_*.npz")))
IDS = [os.path.basename(p)[len("learner_"):-4] for p in PATHS]


def load_case(path):
    """-> (engine kwargs, RingView, npz).  The ring is stored in the fixture, or rebuilt from its seed and checked against the
    stored checksums when the fixture would be too large."""
    g = np.load(path)
    kw = ast.literal_eval(str(g["kw"]))
    alpha, beta0, beta_steps, per_eps = [float(x) for x in g["per"]]
    kw.update(seed=int(g["engine_seed"]), target_update_interval=int(g["target_interval"]), warmup_size=kw["batch_size"],
              per_alpha=alpha, per_beta_initial=beta0, per_beta_steps=beta_steps, per_epsilon=per_eps, discount=float(g["discount"]),
              lr=float(g["lr"]))
    if "ring_obs" in g.files:
        D, A = g["ring_obs"].shape[1], (2 if kw["env"] == "CartPole-v1" else 4)
        v = ck.RingView(kw["n_envs"], kw["ring_rows"], kw["multisteps"], A, D, vec_steps=int(g["vec_steps"]))
        v.obs[:], v.next_obs[:], v.action[:], v.reward[:] = g["ring_obs"], g["ring_next_obs"], g["ring_action"], g["ring_reward"]
        v.term[:], v.done[:] = g["ring_term"], g["ring_done"]
        if kw["mem_kind"]:
            v.leaf_priority = g["leaf_priority"].astype(np.float64)
    else:
        v = synth_ring(dict(kw, vec_steps=int(g["vec_steps"])), int(g["ring_seed"]))
    v.max_priority = float(g["max_priority0"])
    chk = [float(v.obs.astype(np.float64).sum()), float(v.reward.astype(np.float64).sum()), float(v.action.sum()),
           float(0.0 if v.leaf_priority is None else v.leaf_priority.sum())]
    np.testing.assert_allclose(chk, g["ring_checksum"], rtol=1e-12)  # the rebuilt ring is the one the reference trained on
    return kw, v, g
