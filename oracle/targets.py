"""CPU restatement of the TD-target computations on the hot path (TEST INFRASTRUCTURE).

  dqn_target      <- CommonInterfaceParameter.calc_target_q, srl/algorithms/dqn/dqn.py:144-176 and the identical
                     1-step Rainbow variant srl/algorithms/rainbow/rainbow_nomultisteps.py:10-43
  rainbow_target  <- CommonInterfaceParameter.calc_target_q (n-step + Retrace), srl/algorithms/rainbow/rainbow.py:185-287
  rescaling / inverse_rescaling <- srl/rl/functions.py:10-17
  huber_loss_and_grad <- torch.nn.HuberLoss(delta=1, reduction="mean") applied to (target*w, q*w) as in
                     srl/algorithms/dqn/model_torch.py:115 / rainbow/model_torch.py:105

Inputs are dense arrays instead of the reference's python lists of records; the arithmetic (dtype promotions included)
follows the reference line by line.  Pinned by tests/golden/targets.npz = the reference functions executed here on frozen
batches (tests/golden/make_golden.py).
"""
import numpy as np


def rescaling(x, eps=0.001):
    return np.sign(x) * (np.sqrt(np.abs(x) + 1.0) - 1.0) + eps * x


def inverse_rescaling(x, eps=0.001):
    n = np.sqrt(1.0 + 4.0 * eps * (np.abs(x) + 1.0 + eps)) - 1.0
    n = n / (2.0 * eps)
    return np.sign(x) * ((n**2) - 1.0)


def dqn_target(pred_q, pred_target_q, n_state, reward, undone, discount, enable_double_dqn=True, enable_rescale=False,
               next_invalid=None, np_dtype=np.float32):
    """dqn.py:144-176.  n_state [B,D] f32, reward [B] f32, undone [B] int, next_invalid optional bool [B,A]."""
    batch_size = len(n_state)
    n_q_target = np.array(pred_target_q(n_state))
    if enable_double_dqn:
        n_q = np.array(pred_q(n_state))
        if next_invalid is not None and next_invalid.any():
            n_q[next_invalid] = np.min(n_q)  # global batch min, dqn.py:160
        n_act_idx = np.argmax(n_q, axis=1)
        maxq = n_q_target[np.arange(batch_size), n_act_idx]
    else:
        if next_invalid is not None and next_invalid.any():
            n_q_target[next_invalid] = np.min(n_q_target)
        maxq = np.max(n_q_target, axis=1)
    if enable_rescale:
        maxq = inverse_rescaling(maxq)
    target_q = reward + undone * discount * maxq
    if enable_rescale:
        target_q = rescaling(target_q)
    return target_q.astype(np_dtype)


def rainbow_target(pred_q, pred_target_q, states, actions, rewards, dones, discount, multisteps, retrace_h=1.0,
                   enable_double_dqn=True, enable_rescale=False, n_actions=None, np_dtype=np.float32, next_invalid=None):
    """rainbow.py:185-287.

    states  [B, M+1, D]   tracking "state" of the M+1 window entries
    actions [B, M] int    action index of entries 1..M (one-hot in the reference)
    rewards [B, M], dones [B, M]  entries 1..M ("terminated")
    next_invalid optional bool [B, M, A]: invalid actions of entries 1..M (rainbow.py:236-249: -inf before the argmax)
    Returns (target_q [B], state [B,D], action_idx [B]).
    """
    B, M = actions.shape
    assert M == multisteps
    A = n_actions
    multi_discounts = np.tile(np.array([discount**n for n in range(M)], dtype=np_dtype), (B, 1))
    onehot = np.eye(A, dtype=np_dtype)[actions]  # [B,M,A]
    reward = rewards.astype(np_dtype)
    done = dones.astype(np_dtype)

    state = states[:, 0, :]
    n_state = states[:, 1:, :]
    action = onehot[:, 0, :]
    n_action = onehot[:, 1:, :]

    if enable_double_dqn:
        online_state = n_state
    else:
        online_state = n_state[:, :-1, :]
    target_state = n_state
    online_shape1 = online_state.shape[1]
    online_flat = np.reshape(online_state, (B * online_shape1,) + online_state.shape[2:])
    target_flat = np.reshape(target_state, (B * M,) + target_state.shape[2:])

    if online_shape1 > 0:
        q_online = np.array(pred_q(online_flat))
        q_online = np.reshape(q_online, (B, online_shape1) + q_online.shape[1:])
    else:
        q_online = np.zeros((B, 0, A), dtype=np_dtype)
    q_target = np.array(pred_target_q(target_flat))
    q_target = np.reshape(q_target, (B, M) + q_target.shape[1:])

    q = np.sum(q_online[:, : M - 1, :] * n_action, axis=2)
    q = np.insert(q, 0, 0, axis=1)

    if enable_double_dqn:
        if next_invalid is not None and next_invalid.any():
            q_online[next_invalid] = -np.inf
        n_act_idx = np.argmax(q_online, axis=2)
    else:
        if next_invalid is not None and next_invalid.any():
            q_target[next_invalid] = -np.inf
        n_act_idx = np.argmax(q_target, axis=2)
    maxq = np.take_along_axis(q_target, np.expand_dims(n_act_idx, axis=2), axis=2)
    maxq = np.squeeze(maxq, axis=2)
    if enable_rescale:
        maxq = inverse_rescaling(maxq)
    gains = reward + (1 - done) * discount * maxq
    if enable_rescale:
        gains = rescaling(gains)
    td_errors = gains - q

    # Retrace with the reference's index shift (rainbow.py:267): action taken at s_{j+1} vs greedy action at s_{j+2}
    pi_probs = np.argmax(n_action, axis=2) == n_act_idx[:, 1:]
    pi_probs = np.transpose(pi_probs, (1, 0))
    retrace_list = [np.ones((B,))]
    retrace = np.ones((B,))
    for n in range(M - 1):
        retrace *= retrace_h * pi_probs[n]
        retrace_list.append(retrace.copy())
    retrace_list = np.asarray(retrace_list).transpose((1, 0))
    target_q = np.sum(td_errors * multi_discounts * retrace_list, axis=1, dtype=np_dtype)
    return target_q, state, actions[:, 0]


def huber_loss_and_grad(target_q, q, weights, delta=1.0):
    """loss = mean(huber(q*w - target*w)); returns (loss, dloss/dq [B], priorities=|target-q| [B]) in float32."""
    t = (target_q * weights).astype(np.float32)
    p = (q * weights).astype(np.float32)
    d = p - t
    ad = np.abs(d)
    per = np.where(ad <= delta, 0.5 * d * d, delta * (ad - 0.5 * delta)).astype(np.float32)
    loss = np.float32(per.mean())
    dq = (np.clip(d, -delta, delta) * weights / len(q)).astype(np.float32)
    pri = np.abs(target_q - q).astype(np.float32)
    return loss, dq, pri
