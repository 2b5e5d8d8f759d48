/*
 * srlx.h -- C ABI of libsrlx.so: the B200 (sm_100a) rollout -> replay -> PER sample -> TD/Adam update path
 * behind the SRL (pocokhc/simple_distributed_rl v1.4.5) plugin API.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No torch / Python types cross this boundary.
 *   - Every buffer is CALLER-OWNED device memory (the Python host side allocates torch CUDA tensors and passes
 *     data_ptr()).  The library keeps no state between calls except the thread-local error string.
 *   - Every call takes the CUDA stream as a uintptr_t-sized handle and is asynchronous on it unless noted "sync".
 *   - Return value: 0 = ok, <0 = error; srlx_last_error() returns a thread-local message.
 *   - One set of buffers is driven from one host thread at a time (documented, not locked), like the reference's
 *     GIL-serialised pybind11 module (srl/rl/memories/priority_memories/cpp_module/src/proportional_memory.cpp).
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   srlx_tree_*      <- ProportionalMemory / SumTree: srl/rl/memories/priority_memories/proportional_memory.py:13-205
 *                       and its pybind11 twin cpp_module/src/proportional_memory.cpp:14-275
 *   srlx_vec_step    <- core_play.play loop body srl/base/run/core_play.py:115-214 (policy -> env.step -> on_step),
 *                       EnvRun.step srl/base/env/env_run.py:254-366, WorkerRun srl/base/rl/worker_run.py:310-401,
 *                       Grid srl/envs/grid.py:173-208,340-378, dqn.Worker.policy/on_step srl/algorithms/dqn/dqn.py:192-246,
 *                       rainbow.Worker srl/algorithms/rainbow/rainbow.py:301-400
 *   srlx_learn       <- dqn Trainer.train srl/algorithms/dqn/model_torch.py:90-132, rainbow Trainer.train
 *                       srl/algorithms/rainbow/model_torch.py:85-122, calc_target_q dqn.py:144-176,
 *                       rainbow.py:185-287, rainbow_nomultisteps.py:10-43, PriorityReplayBuffer.sample/update
 *                       srl/rl/memories/priority_replay_buffer.py:228-250, ReplayBuffer.sample
 *                       srl/rl/memories/priority_memories/replay_buffer.py:34-36
 *   srlx_qnet_forward<- RLParameter.pred_q / pred_target_q srl/algorithms/dqn/model_torch.py:58-70
 */
#ifndef SRLX_H_
#define SRLX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRLX_VERSION 100 /* 0.1.0 */

#define SRLX_MAX_LAYERS 6   /* dense layers incl. the output layer */
#define SRLX_MAX_OBS 16     /* observation floats per env step */
#define SRLX_MAX_ACTIONS 16 /* discrete actions */
#define SRLX_MAX_MULTISTEPS 8
#define SRLX_MAX_BATCH 256

/* ---- environments (closed-form, stepped on device) ------------------------------------------------------ */
/* SRLX_ENV_EXTERNAL: the env is stepped by a host loop (the reference's core_play.play through the plug-in classes); transitions
 * enter through srlx_ext_step and srlx_vec_step is refused */
enum { SRLX_ENV_GRID = 0, SRLX_ENV_CARTPOLE = 1, SRLX_ENV_PENDULUM = 2, SRLX_ENV_EXTERNAL = 3 };

/* dueling combine: srl/rl/torch_/blocks/dueling_network.py:51-58 */
enum { SRLX_DUEL_NONE = 0, SRLX_DUEL_AVERAGE = 1, SRLX_DUEL_MAX = 2, SRLX_DUEL_NAIVE = 3 };

/* replay kind: srl/rl/memories/priority_replay_buffer.py:119-152 */
enum { SRLX_MEM_UNIFORM = 0, SRLX_MEM_PROPORTIONAL = 1 };

/* Q-network: a stack of dense layers (ReLU between, last one linear).  Flat parameter layout, fp32:
 *   for l in 0..n_layers-1:  W_l [out_l][in_l] row-major (torch nn.Linear layout) at w_off[l], b_l [out_l] at b_off[l].
 * Dueling head (srl/rl/torch_/blocks/dueling_network.py:8-59): the LAST HIDDEN layer is the concatenation
 * [value-hidden(H) ; advantage-hidden(H)] (width 2H) and the output layer has 1+A rows of H inputs each:
 * row 0 (V) reads hidden[0:H], rows 1..A (advantages) read hidden[H:2H].
 * Noisy layers (srl/rl/torch_/modules/noisy_linear.py:8-52): a second flat buffer "sigma" with the same layout;
 * effective weight = mu + sigma * N(0,1), noise re-drawn per forward call. */
typedef struct srlx_net {
  int32_t n_layers;
  int32_t in_dim;
  int32_t out_dim[SRLX_MAX_LAYERS]; /* rows of W_l; last = n_actions (plain) or 1+n_actions (dueling) */
  int32_t k_dim[SRLX_MAX_LAYERS];   /* columns of W_l (in_l; for the dueling output layer this is H) */
  int32_t w_off[SRLX_MAX_LAYERS];
  int32_t b_off[SRLX_MAX_LAYERS];
  int32_t n_params; /* floats in one flat buffer */
  int32_t n_actions;
  int32_t dueling; /* SRLX_DUEL_* */
  int32_t noisy;   /* 0/1: any layer is a NoisyLinear */
  int32_t layer_noisy[SRLX_MAX_LAYERS]; /* per layer: the plain out Linear of a noisy rainbow MLP is NOT noisy
                                           (srl/rl/models/config/dueling_network.py:125-127) */
} srlx_net;

/* Device-resident counters / scalars shared by the kernels (one per engine, 256-byte aligned). */
typedef struct srlx_state {
  uint64_t vec_steps;     /* g: vector steps executed so far */
  uint64_t total_step;    /* RunState.total_step  (srl/base/context.py:297-343): env steps = vec_steps * n_envs */
  uint64_t train_count;   /* RLTrainer.train_count (srl/base/rl/trainer.py:14-62) */
  uint64_t episode_count; /* RunState.episode_count */
  uint64_t sync_count;    /* trainer.info["sync"] */
  uint64_t adam_step;     /* torch.optim.Adam state["step"] */
  uint64_t mem_size;      /* items currently sampleable (ProportionalMemory.size / len(ReplayBuffer.memory)) */
  uint64_t sample_retries;/* PER rejections (priority==0 or duplicate), diagnostics */
  double max_priority;    /* ProportionalMemory.max_priority (proportional_memory.py:116) */
  double episode_reward_sum;   /* sum of finished-episode rewards (for mean reward reporting) */
  double last_loss;       /* trainer.info["loss"] of the last update */
  double loss_sum;        /* sum of losses since last host reset */
  uint64_t episode_len_sum;
  uint64_t reserved[3];
} srlx_state;

/* Engine configuration: dimensions, hyper-parameters and the device pointers of every caller-owned buffer. */
typedef struct srlx_engine {
  /* ---- sizes ---- */
  int32_t env_id;          /* SRLX_ENV_* */
  int32_t n_envs;          /* E */
  int32_t obs_dim;         /* D */
  int32_t n_actions;       /* A */
  int32_t ring_rows;       /* R: ring capacity = R*E slots; slot(g,e) = (g % R)*E + e */
  int32_t multisteps;      /* M >= 1 (rainbow.Config.multisteps) */
  int32_t batch_size;      /* B */
  int32_t mem_kind;        /* SRLX_MEM_* */
  /* ---- algorithm switches (dqn.py:50-103, rainbow.py:57-108) ---- */
  int32_t enable_double_dqn;
  int32_t enable_rescale;
  int32_t enable_reward_clip;
  int32_t has_duplicate;   /* ProportionalMemory.has_duplicate */
  int32_t target_update_interval;
  int32_t trunc_limit;     /* episode is truncated when step_num reaches this (CartPole 500; Grid 51 = max_episode_steps+1) */
  int32_t trunc_overrides_term; /* gymnasium TimeLimit semantics: truncated wins over terminated */
  int32_t presample;       /* 0: every batch is sampled after the previous update's priorities are in the tree (Trainer.train in sequence);
                              1: batch t+1 is drawn BEFORE update t is applied (one update of staleness, as the reference's own memory
                              process does in distributed mode, srl/base/run/play_mp_memory.py:253-350): the SumTree chain leaves the
                              critical path (csrc/learner_fast.cu only) */
  uint64_t seed;
  uint64_t warmup_size;
  /* ---- hyper-parameters ---- */
  double epsilon;          /* epsilon-greedy (ignored when net.noisy, rainbow.py:305-309) */
  double discount;
  double lr;
  double adam_beta1, adam_beta2, adam_eps;
  double retrace_h;
  double per_alpha, per_beta_initial, per_beta_steps, per_epsilon;
  double reward_shift, reward_scale; /* WorkerRun: (r + shift) * scale, worker_run.py:348 */
  double huber_delta;
  /* ---- Grid env tables (srl/envs/grid.py) ---- */
  int32_t grid_w, grid_h;
  int8_t grid_field[64];       /* row-major [h][w]; 9 wall, 1 goal, -1 hole, 2 start, 0 road */
  int32_t grid_n_starts;
  int32_t grid_starts[16];     /* x | y<<8 */
  double grid_slip_cdf[16];    /* [chosen action][4] normalised cdf in the order np.random.choice sees (UP,DOWN,RIGHT,LEFT) */
  int32_t grid_slip_action[4]; /* executed action for each cdf slot: {3,1,2,0} */
  double grid_move_reward, grid_goal_reward, grid_hole_reward;
  /* ---- discretised continuous action (Pendulum): value of action index a, the reference's BoxSpace division table
   *      (srl/base/spaces/box.py:317-366, RLConfig.action_division_num srl/base/rl/config.py:55), float32 values ---- */
  double act_tbl[16];
  /* ---- network ---- */
  srlx_net net;
  /* ---- device buffers (caller-owned) ---- */
  srlx_state* state;        /* 1 */
  double* env_state;        /* [E][4] f64 (CartPole x,x_dot,theta,theta_dot; Grid x,y in [0],[1]) */
  int32_t* env_step_num;    /* [E] EnvRun._step_num */
  uint32_t* env_episode;    /* [E] episodes started by this env (RNG counter) */
  double* env_ep_reward;    /* [E] EnvRun._episode_rewards */
  uint8_t* env_needs_reset; /* [E] */
  double* env_first_ep_reward; /* [E] optional (NULL ok): reward of the FIRST episode env e finished (Runner.evaluate) */
  int32_t* env_last_ep_len;    /* [E] optional: length of the last finished episode, 0 = none finished yet */
  float* ring_obs;          /* [R*E][D] state            */
  float* ring_next_obs;     /* [R*E][D] next_state       */
  int32_t* ring_action;     /* [R*E] action index (one-hot in the reference record) */
  float* ring_reward;       /* [R*E] */
  uint8_t* ring_term;       /* [R*E] worker.terminated (NOT truncated) */
  uint8_t* ring_done;       /* [R*E] episode ended at this step (terminated or truncated) */
  double* tree;             /* [2*R*E-1] SumTree nodes, leaf j at j + R*E - 1 (proportional_memory.py:13-47); NULL for uniform */
  double* tree_scratch;     /* [2*(E+2)] scratch for the per-step bulk add; NULL for uniform */
  float* params;            /* [n_params] online mu */
  float* params_sigma;      /* [n_params] online sigma (noisy) or NULL */
  float* target;            /* [n_params] target mu */
  float* target_sigma;      /* [n_params] or NULL */
  float* adam_m;            /* [n_params * (1+noisy)] exp_avg   (mu block then sigma block) */
  float* adam_v;            /* [n_params * (1+noisy)] exp_avg_sq */
  /* ---- optional debug taps (NULL in production) ---- */
  float* dbg_q;             /* [E][A] Q values the policy saw in the last vector step */
  int32_t* dbg_action;      /* [E] */
  int64_t* dbg_sample_idx;  /* [B] tree indices (PER) or slots (uniform) of the last update */
  float* dbg_weights;       /* [B] IS weights of the last update */
  float* dbg_target_q;      /* [B] */
  float* dbg_q_sa;          /* [B] */
  float* dbg_grads;         /* [n_params*(1+noisy)] gradient of the last update */
  float* dbg_windows;       /* [B][M+1][D] states, then [B][M] (action, reward, term) as floats */
  long long* dbg_clock;     /* [64] clock64() stamps of the learner phases of the second-to-last update of a launch (CTA 0):
                               [0..7] compute warps, [16..24] aux / memory warps; see learner.cu, learner_fast.cu */
  /* ---- scratch (caller-owned) ---- */
  float* noise_scratch;     /* NoisyNet draws of a chunk of updates, precomputed by all SMs and streamed by the learner
                               (learner_fast.cu); NULL -> the generic learner draws in-kernel */
  uint64_t noise_scratch_bytes;
  double* tree_blk;         /* optional: blocked copy of the SumTree levels below the shared-memory-cached top (each 5-level
                               subtree = 62 contiguous doubles in a 512-byte block), rebuilt by srlx_learn at every call and
                               kept in step by the learner, so the sampler fetches a subtree with one coalesced load per
                               lane; srlx_tree_blk_bytes(capacity) bytes; NULL -> the sampler reads the flat tree */
  uint64_t tree_blk_bytes;
  /* ---- linear epsilon schedule (srl/rl/schedulers/schedulers/linear.py:16-21, read per step at dqn.py:196 / rainbow.py:312 with
   *      the worker's step_in_training, here the vector step count g): eps_phase_steps == 0 -> constant `epsilon`;
   *      else training rollouts use g >= phase ? eps_end : epsilon - ((epsilon - eps_end) / phase) * g ---- */
  double eps_end;
  uint64_t eps_phase_steps;
  /* ---- data-parallel single learner over the GPUs of one node (csrc/learner_fast.cu): every rank samples
   *      batch_size items from its own replay shard, the gradients of the global batch (dp_world x batch_size items) are summed
   *      over NVLink inside the learner kernel every update (peer stores into dp_peer[r] + flags), the IS weights use the global
   *      N / total / max (srl/rl/memories/priority_memories/proportional_memory.py:138-167), and every rank applies the identical
   *      Adam step -- the reference's ONE trainer (srl/base/run/play_mp.py:352-462) with its memory sharded over the GPUs ---- */
  uint64_t learner_seed;  /* NoisyNet draws of the trainer (0 -> seed): the same on every rank, so every replica builds the same weights */
  int32_t dp_world;       /* ranks exchanging gradients per update; 0 / 1 = off */
  int32_t dp_rank;
  void* dp_peer[8];       /* dp_peer[r]: the exchange buffer of rank r as addressable from this rank (own: srlx_dp_alloc or any
                             zeroed device buffer of srlx_dp_bytes(); peers: srlx_dp_open of their IPC handle, or a peer-enabled pointer) */
  uint64_t dp_bytes;      /* size of every exchange buffer */
  /* ---- invalid-action masks (dqn.py:156-165 with its batch-minimum fill, rainbow_nomultisteps.py:19-31, rainbow.py:236-252 with
   *      -inf per window step): bit a of ring_invalid[slot] = action a is invalid in that row's NEXT state.  NULL: the env has no
   *      invalid actions (the device envs).  Rows enter through srlx_ext_step_masked; the generic learner (csrc/learner.cu) reads
   *      them, and srlx_learn dispatches to it whenever the buffer is present ---- */
  uint32_t* ring_invalid; /* [R*E] or NULL */
  /* ---- any epsilon schedule (SchedulerConfig with several phases / cosine / polynomial ..., srl/rl/schedulers/scheduler.py:232-345) as
   *      a table: training rollouts of vector step g use eps_table[min(g, eps_table_len - 1)] (the host fills it by stepping the
   *      reference's own scheduler object 0, 1, 2, ... until it is constant).  NULL: `epsilon` / the linear phase above ---- */
  const double* eps_table;
  uint64_t eps_table_len;
} srlx_engine;

/* ---- PPO (R15; BASELINE configs[4]) -- csrc/ppo.cu restates srl/algorithms/ppo/ppo.py for E vectorised env copies ------------------- */
typedef struct srlx_ppo_state {
  uint64_t train_count;  /* RLTrainer.train_count: minibatch updates done */
  uint64_t adam_step;    /* optimizer.iterations */
  double policy_loss, value_loss, entropy_loss; /* trainer.info of the last update (ppo.py:271-273) */
  double grad_norm;      /* global gradient norm before clipping */
  uint64_t reserved[2];
} srlx_ppo_state;

typedef struct srlx_ppo {
  srlx_engine env;       /* env tables, env buffers, counters (`state`), seed, reward_shift / reward_scale; ring / tree / Q-net fields unused */
  /* ActorCriticNetwork (ppo.py:55-101) as two dense stacks over ONE flat parameter buffer: value stack = trunk + value block + 1
   * output, policy stack = trunk + policy block + outputs (continuous: loc, log_scale; discrete: logits); the trunk layers carry
   * the same w_off / b_off in both.  ReLU between layers, last layer linear. */
  srlx_net net_v, net_p;
  int32_t n_params;
  int32_t continuous;          /* 1: Normal policy on a 1-D Box action (Pendulum-v1); 0: Categorical over env.n_actions */
  int32_t horizon;             /* T: rows of the rollout buffer */
  int32_t batch_size;          /* <= 32 */
  int32_t baseline_type;       /* 0 none, 1 "ave", 2 "std", 3 "normal", 4 "advantage" (ppo.py:220-232, :125-126) */
  int32_t surrogate_clip;      /* 1 "clip", 0 "" */
  int32_t enable_value_clip, state_normalized;
  int32_t method;              /* SRLX_RETURNS_GAE / SRLX_RETURNS_MC */
  int32_t reward_clip_enable;
  uint64_t lr_decay_steps;     /* 0: constant lr; else keras ExponentialDecay(staircase=True): lr * rate^floor(step / steps) */
  double discount, gae_discount, policy_clip_range, value_clip_range, lr, lr_decay_rate, value_loss_weight, entropy_weight, grad_clip_norm;
  double adam_beta1, adam_beta2, adam_eps;  /* keras Adam: 0.9, 0.999, 1e-7 */
  double log_scale_lo, log_scale_hi;        /* log of stable_gradients_scale_range (normal_dist_block.py:96-100,148-153) */
  double action_low, action_high;           /* the env's Box: rescale_from [-1, 1] then clip (np_array.py:64-67,93-95) */
  double reward_clip_lo, reward_clip_hi;
  /* device buffers (caller-owned) */
  float* params; float* adam_m; float* adam_v;  /* [n_params] */
  float* buf_obs;          /* [T][E][D] worker batch["state"] */
  float* buf_action;       /* [T][E] the policy's action (before rescaling) or the action index */
  float* buf_v;            /* [T][E] batch["v"]: V(s) at acting time (old_v of the value clip) */
  float* buf_logp;         /* [T][E] batch["log_prob"] */
  float* buf_reward;       /* [T][E] */
  unsigned char* buf_done; /* [T][E] */
  float* buf_vnew;         /* [T+1][E] V(s) with the parameters at the end of the rollout (GAE) */
  float* buf_ret;          /* [T][E] batch["discounted_reward"] */
  unsigned char* buf_valid;/* [T][E] 1 where the reference would have emitted the step (its episode ended inside the buffer) */
  srlx_ppo_state* pstate;
  int32_t* dbg_idx;        /* optional [batch_size]: flat buffer indices of the last minibatch */
  float* dbg_grads;        /* optional [n_params]: clipped gradient of the last update */
  float* grad_scratch;     /* [n_params]: gradient accumulator of the update kernel when it does not fit shared memory (wide blocks) */
} srlx_ppo;

size_t srlx_sizeof_ppo(void);
size_t srlx_sizeof_ppo_state(void);
/* one vector step of all E env copies under the current policy (Worker.policy + env.step); training != 0 stores row vec_steps % T */
int srlx_ppo_vec_step(const srlx_ppo* ppo, int training, uintptr_t cuda_stream);
/* V(s) of n states [n][D] with the current parameters */
int srlx_ppo_values(const srlx_ppo* ppo, const float* obs_dev, uint64_t n, float* out_dev, uintptr_t cuda_stream);
/* Worker.on_step at episode end for the whole buffer (ppo.py:375-404): values with the current parameters, GAE / MC -> buf_ret, buf_valid */
int srlx_ppo_finish_rollout(const srlx_ppo* ppo, uintptr_t cuda_stream);
/* n_updates x Trainer._train (ppo.py:208-291) on the finished buffer */
int srlx_ppo_learn(const srlx_ppo* ppo, uint32_t n_updates, uintptr_t cuda_stream);

/* ---- R2D2 (R14; BASELINE configs[3]) -- csrc/r2d2.cu restates srl/algorithms/r2d2/r2d2.py for E vectorised env copies --------------
 * Q-network (r2d2.py:27-63): Flatten -> keras LSTM(lstm_units) -> hidden block (MLP + Dense(A), or MLP + dueling block).
 * Parameter layout (one flat fp32 buffer; every dense map carries its bias as a LAST COLUMN, activations carry a trailing 1):
 *   lstm   W[4u][in + u + 1]   row j = unit * 4 + gate (gate order i, f, c~, o as keras), columns [x (in) | h (u) | bias]
 *   head l W[out_l][k_l + 1]   ReLU after every layer but the last; dueling: the last hidden layer is [value-hidden(H) ; advantage-
 *          hidden(H)] (width 2H) and the output layer has 1 + A rows over 2H inputs whose off-branch blocks are structural zeros
 *          (row 0 reads [0,H), rows 1..A read [H,2H)); duel_hidden = H tells the optimiser which entries never move.
 * Replay: one ring COLUMN per env copy with its own cursor (slot(e, c) = (c % R) * E + e).  A row stores one step of the worker's
 * recent_* lists (r2d2.py:232-283): state, next state, action, behaviour probability, reward, terminated, the LSTM state BEFORE the step,
 * and the step's index in its episode; the seq_len-1 padded steps the reference appends at an episode's end (:285-303) are written as
 * rows too, the padding in front of an episode's first step (:232-247) is rebuilt by index.  Item = anchor row = the newest
 * transition of a (burnin + seq_len + 1)-state window, exactly one item per memory.add() of the reference. */
typedef struct srlx_r2d2 {
  srlx_engine env;  /* env tables + env buffers + `state` counters + seed, epsilon, discount, lr, adam_*, per_*, batch_size, mem_kind,
                       has_duplicate, warmup_size, target_update_interval, enable_double_dqn, enable_rescale, retrace_h, ring_rows (R),
                       tree ([2*R*E-1], proportional only); the Q-net / ring / learner fields of srlx_engine are unused */
  int32_t lstm_units, burnin, seq_len, enable_retrace;
  int32_t n_head, dueling;                 /* head layers (incl. the output layer); SRLX_DUEL_* */
  int32_t head_out[SRLX_MAX_LAYERS], head_k[SRLX_MAX_LAYERS], head_off[SRLX_MAX_LAYERS]; /* rows, inputs (without the bias column), offset */
  int32_t lstm_off, n_params;
  int32_t duel_hidden;                     /* H of the dueling block (0: no dueling block) */
  int32_t no_persistent;                   /* != 0: one launch per time step instead of the persistent unroll kernels (diagnostic) */
  double test_epsilon;
  /* parameters */
  float* params; float* target; float* adam_m; float* adam_v; float* grads; /* [n_params] each */
  /* replay ring, [R*E] rows */
  uint32_t* cursor;        /* [E] rows written so far by env e */
  float* ring_obs; float* ring_next_obs;   /* [R*E][D] */
  int32_t* ring_action; double* ring_prob; double* ring_reward; unsigned char* ring_done; int32_t* ring_tstep; /* [R*E] */
  float* ring_h; float* ring_c;            /* [R*E][u] */
  /* rollout workspace */
  float* roll_xh;          /* [E][in+u+1] the step's LSTM input [x | h_{t-1} | 1] (the library writes the trailing 1) */
  float* roll_h;           /* [E][u+1] h_t carried to the next step, last column 1 (head input) */
  float* roll_c;           /* [E][u] */
  float* roll_gates;       /* optional [E][4u]: gate pre-activations of the step (E > 64: the step runs as tiled GEMM + cell update) */
  float* roll_act[SRLX_MAX_LAYERS]; /* [E][head_out[l] + 1] (last column 1); the last one [E][head_out] */
  unsigned char* roll_reset; /* [E] */
  uint32_t* new_c0; uint32_t* new_n;  /* [E] first new row / rows written by the last vector step */
  int64_t* add_idx; double* add_pri;  /* [E*(2*seq_len+burnin+seq_len)] replay-add list (proportional) */
  /* learner workspace; z = 0 online, 1 target; W = burnin + seq_len; rows t-major */
  float* xh;               /* [2][W+2][B][in+u+1], last column 1 */
  float* cbuf;             /* [2][W+2][B][u] */
  float* gates;            /* [W+1][B][4u] online gate activations */
  float* dgates;           /* [W+1][B][4u] */
  float* dc;               /* [B][u] */
  float* gemm_ws;          /* optional split-K workspace of the narrow weight-gradient maps (32 * max_l head_out[l] * (head_k[l] + 1) floats is enough) */
  uint64_t gemm_ws_floats;
  uint32_t* bar;           /* [8] barrier words of the persistent unroll kernels ([0..3]; NULL: one launch per time step) and of the
                              cooperative replay add ([4], [5]; required with proportional replay) */
  float* act[SRLX_MAX_LAYERS];  /* [2][(seq_len+1)*B][head_out[l] + 1] (last column 1); last layer [2][rows][head_out] */
  float* dact[SRLX_MAX_LAYERS]; /* [(seq_len+1)*B][head_out[l]] gradient wrt the layer's output */
  float* dh;               /* [(seq_len+1)*B][u] gradient wrt the LSTM outputs */
  float* q;                /* [2][B][seq_len+1][A] */
  int64_t* sel;            /* [B] tree indices (proportional) or slots (uniform) of the batch */
  float* weights;          /* [B] IS weights */
  int32_t* b_actions; double* b_mu; double* b_rewards; unsigned char* b_dones; /* [B][seq_len] */
  double* b_target; double* b_tdmean; unsigned char* b_tdkind;                 /* [B][seq_len], [B], [B] */
} srlx_r2d2;

size_t srlx_sizeof_r2d2(void);
/* one vector step (Worker.policy + env.step + Worker.on_step, r2d2.py:249-318); training != 0 stores rows and adds items */
int srlx_r2d2_vec_step(const srlx_r2d2* r, int training, uintptr_t cuda_stream);
/* n_updates x Trainer.train (r2d2.py:90-215); no-op while mem_size < warmup_size */
int srlx_r2d2_learn(const srlx_r2d2* r, uint32_t n_updates, uintptr_t cuda_stream);
/* the same update in two halves, for a data-parallel trainer over several GPUs: phases = 1 runs sample .. backward and leaves the
 * gradient of this rank's batch in r->grads (the caller all-reduces it: NCCL over NVLink), phases = 2 runs Adam, the priority update,
 * the target sync and the counters; phases = 3 = srlx_r2d2_learn.  One update per call when split. */
int srlx_r2d2_learn_phase(const srlx_r2d2* r, uint32_t n_updates, int phases, uintptr_t cuda_stream);
/* q_out[n][A] = Q of ONE step from the given LSTM state: obs [n][D], h / c [n][u] in, h_out / c_out [n][u] out (may alias h / c);
 * test tap and evaluation helper: runs in the learner workspace (r->xh, r->cbuf, r->act), so n <= batch_size and never between the
 * kernels of an update */
int srlx_r2d2_forward(const srlx_r2d2* r, int use_target, const float* obs_dev, const float* h_dev, const float* c_dev, uint32_t n,
                      float* q_out_dev, float* h_out_dev, float* c_out_dev, uintptr_t cuda_stream);
/* C[M][N] = A . B with arbitrary strides (the fp32 GEMM every R2D2 layer runs on; test tap): A(m,k) = a[m*sa_m + k*sa_k],
 * B(k,n) = b[k*sb_k + n*sb_n], C row stride ldc; relu != 0 clamps at 0; accumulate != 0 adds to C */
int srlx_sgemm(const float* a_dev, long long sa_m, long long sa_k, const float* b_dev, long long sb_k, long long sb_n, float* c_dev,
               long long ldc, int M, int N, int K, int relu, int accumulate, uintptr_t cuda_stream);

/* ---- library ------------------------------------------------------------------------------------------ */
int srlx_version(void);
const char* srlx_last_error(void);
size_t srlx_sizeof_engine(void);
size_t srlx_sizeof_state(void);
size_t srlx_sizeof_net(void);
/* Number of kernels launched by this library in this process since load (bench.py "gpu_launches"). */
uint64_t srlx_launch_count(void);

/* ---- RNG taps (parity of the device Philox4x32-10 streams with oracle/philox.py) ------------------------- */
/* out[i*4..i*4+3] = philox4x32-10(ctr = (a0+i, b, c, stream), key = seed) for i < n. */
int srlx_philox_words(uint64_t seed, uint32_t stream, uint32_t a0, uint32_t b, uint32_t c, uint32_t* out_dev,
                      size_t n, uintptr_t cuda_stream);
/* out[i] = x[i]^a as the fused learner evaluates the priority (|td| + eps)^alpha and the IS weights (a short fp64
 * log/exp, csrc/learner_fast.cu::pow_chain): test tap for its parity with libm pow (max rel. difference 4e-15). */
int srlx_dbg_pow(const double* x_dev, double a, double* out_dev, size_t n, uintptr_t cuda_stream);
/* The N(0,1) noise tensor a NoisyLinear forward call `call_id` of kind `kind` uses (flat param layout). */
int srlx_noise_fill(uint64_t seed, uint32_t kind, uint64_t call_id, float* out_dev, size_t n_params,
                    uintptr_t cuda_stream);

/* ---- SumTree / ProportionalMemory (R7) -- narrow seam, mirrors the pybind11 class method by method -------- */
/* tree: [2*capacity-1] doubles; meta: srlx_state (uses mem_size, max_priority, vec_steps as the ring `write`). */
int srlx_tree_clear(double* tree, uint64_t capacity, srlx_state* meta, uintptr_t cuda_stream);
/* add n items at ring positions write, write+1, ... (mod capacity).  priorities_dev == NULL -> max_priority
 * (proportional_memory.py:120-129); restore_skip != 0 -> priorities are used as-is. */
int srlx_tree_add(double* tree, uint64_t capacity, srlx_state* meta, const double* priorities_dev, uint64_t n,
                  double alpha, double epsilon, int restore_skip, uintptr_t cuda_stream);
/* sample `batch` leaves.  If u01_dev != NULL it holds pre-drawn uniforms [batch][max_tries] (row i, attempt k) so the
 * oracle can replay the identical draw; otherwise Philox(seed, STREAM_SAMPLE, (i | k<<16, step_lo, step_hi)).
 * out_tree_idx: tree indices (leaf + capacity - 1), the `indices` the reference returns; out_weights fp32 IS weights
 * normalised by their max; out_priority (optional) the leaf priorities. */
int srlx_tree_sample(const double* tree, uint64_t capacity, srlx_state* meta, uint32_t batch, uint64_t step,
                     double beta_initial, double beta_steps, int has_duplicate, uint64_t seed,
                     const double* u01_dev, uint32_t max_tries, int64_t* out_tree_idx, float* out_weights,
                     double* out_priority, uintptr_t cuda_stream);
/* update(indices, priorities): leaf <- (|p|+eps)^alpha, propagate the change to the root in index order, track
 * max_priority (proportional_memory.py:171-177). */
int srlx_tree_update(double* tree, uint64_t capacity, srlx_state* meta, const int64_t* tree_idx_dev,
                     const float* priorities_dev, uint32_t n, double alpha, double epsilon, uintptr_t cuda_stream);
/* The whole seam in one launch per sample(): the add / update calls made since the last sample (op list in program order; idx >= 0:
 * update of that tree index with raw priority val; -1: add with raw priority; -2: add with priority None; -3: add with the stored
 * priority, ProportionalMemory.restore) are applied with the reference's sequential association, then `batch` items are drawn
 * (batch == 0: apply only).  ops_* / out_* / flag may be mapped pinned HOST memory (srlx_host_alloc): the kernel reads and writes it
 * directly and stores `seq` into *flag last, so the host polls one word instead of synchronising the stream. */
int srlx_tree_seam(double* tree_dev, uint64_t capacity, srlx_state* meta_dev, const int64_t* ops_idx, const double* ops_val, uint32_t n_ops,
                   double alpha, double epsilon, uint32_t batch, uint64_t step, double beta_initial, double beta_steps, int has_duplicate,
                   uint64_t seed, const double* u01_dev, uint32_t max_tries, int64_t* out_tree_idx, float* out_weights,
                   unsigned long long* flag, unsigned long long seq, uintptr_t cuda_stream);
/* the same call with its per-memory constants in a caller-owned HOST struct (fewer arguments to marshal per call) */
typedef struct srlx_seam {
  double* tree; srlx_state* meta; const int64_t* ops_idx; const double* ops_val; int64_t* out_tree_idx; float* out_weights;
  unsigned long long* flag;
  uint64_t capacity;
  double alpha, epsilon, beta_initial, beta_steps;
  int32_t has_duplicate, reserved;
} srlx_seam;
int srlx_tree_seam_desc(const srlx_seam* desc, uint32_t n_ops, uint32_t batch, uint64_t step, uint64_t seed, const double* u01_dev,
                        uint32_t max_tries, unsigned long long seq, uintptr_t cuda_stream);
/* mapped pinned host memory (zeroed): *host_ptr for the CPU, *dev_ptr for kernels */
int srlx_host_alloc(size_t bytes, void** host_ptr, void** dev_ptr);
int srlx_host_free(void* host_ptr);
/* descend only: out_tree_idx[i] = SumTree._retrieve(tree, 0, vals[i]) (proportional_memory.py:57-66). */
int srlx_tree_retrieve(const double* tree, uint64_t capacity, const double* vals_dev, uint32_t n,
                       int64_t* out_tree_idx, uintptr_t cuda_stream);

/* ---- large-batch Q-net inference on the tensor cores (csrc/qnet_tc.cu: tcgen05.mma + TMEM accumulators + TMA-staged operands) ------
 * An explicit NON-parity mode of RLParameter.pred_q / pred_target_q (srl/algorithms/dqn/model_torch.py:58-70): bf16 operands, fp32
 * accumulation; the fp32 seam (srlx_qnet_forward) stays the parity path.
 * srlx_dense_bf16_tc: Y[M][N] = act(X[M][K] . W[N][K]^T + b); X / W bf16 with K contiguous (row strides ldx / ldw, multiples of 8),
 * bias fp32 [N] or NULL, Y bf16 (out_f32 == 0) or fp32, relu != 0 applies max(., 0).
 * srlx_qnet_forward_tc: the whole network (NoisyLinear draws and the dueling head included) for n states; workspace from
 * srlx_qnet_tc_workspace_bytes. */
int srlx_dense_bf16_tc(const void* x_dev, int ldx, const void* w_dev, int ldw, const float* bias_dev, void* y_dev, int ldy, int out_f32,
                       int M, int N, int K, int relu, uintptr_t cuda_stream);
size_t srlx_qnet_tc_workspace_bytes(const srlx_engine* eng, uint32_t n);
int srlx_qnet_forward_tc(const srlx_engine* eng, int use_target, const float* obs_dev, uint32_t n, uint64_t noise_call_id, float* q_out_dev,
                         void* workspace_dev, size_t workspace_bytes, uintptr_t cuda_stream);

/* ---- rank-based prioritized replay (csrc/rankbased.cu) <- RankBasedMemory
 *      srl/rl/memories/priority_memories/rankbased_memory.py:15-77 behind the IPriorityMemory seam (imemory.py:7-34) --------------
 * priorities: float32 [capacity], item i's priority as the reference stores it (|td| as handed to update(); None -> NaN, sorted last).
 * srlx_rank_sample = sample(batch_size, step) for the first n items: argsort(-priorities[:n]) (radix sort on device), rank
 * probabilities (1/rank)^alpha, np.random.choice(..., replace=False) replayed on the uniform stream u01_dev (NULL: Philox(seed,
 * draw_id)), IS weights (n * prob)^-beta / max as float64.  rebuild_cdf != 0 when n or alpha changed since the last call on this
 * scratch.  out_idx: item indices (what the reference returns as `sampled_indices`); out_ranks (optional): their 0-based ranks;
 * out_used (optional): uniforms consumed.  srlx_rank_update = update(indices, priorities): priorities[idx[i]] = values[i] in order. */
size_t srlx_rank_scratch_bytes(uint64_t capacity);
int srlx_rank_sample(const float* priorities_dev, uint64_t capacity, uint32_t n, double alpha, double beta, uint32_t batch,
                     const double* u01_dev, uint32_t n_u, uint64_t seed, uint64_t draw_id, int rebuild_cdf, void* scratch_dev,
                     int64_t* out_idx_dev, double* out_weights_dev, uint32_t* out_ranks_dev, uint32_t* out_used_dev, uintptr_t cuda_stream);
int srlx_rank_update(float* priorities_dev, const int64_t* idx_dev, const float* values_dev, uint32_t n, uintptr_t cuda_stream);
int srlx_rank_argsort(const float* priorities_dev, uint64_t capacity, uint32_t n, void* scratch_dev, uint32_t* out_sorted_idx_dev,
                      uintptr_t cuda_stream);

/* ---- PPO worker-side returns (R15, the worker half) ------------------------------------------------------ */
#define SRLX_RETURNS_GAE 0
#define SRLX_RETURNS_MC 1
/* ppo.Worker.on_step at episode end (srl/algorithms/ppo/ppo.py:357-406) over a time-major rollout buffer [n_steps][n_envs]:
 * method GAE: out = delta + discount*gae_discount*out_next with delta = r - v at the last step of an episode, else
 * r + discount*next_v - v (float32, the reference's rounding order); method MC: out = r + discount*out_next in float64
 * (reward_f64_dev if given, else reward_dev), stored as float32.  done[t][e] != 0 ends an episode; steps after a column's
 * last episode end are not emitted by the reference yet: out = 0, valid = 0 (valid_dev may be NULL) unless
 * tail_is_episode_end.  clip_enable: reward clip to [clip_lo, clip_hi] first (ppo.py:362-367). */
int srlx_returns_scan(const float* reward_dev, const double* reward_f64_dev, const float* value_dev, const float* next_value_dev,
                      const unsigned char* done_dev, float* out_dev, unsigned char* valid_dev, uint32_t n_steps, uint32_t n_envs,
                      double discount, double gae_discount, int method, int tail_is_episode_end, int clip_enable, double clip_lo,
                      double clip_hi, uintptr_t cuda_stream);

/* ---- R2D2 trainer, per-sequence targets (R14, the target / Retrace / priority half) ----------------------------- */
/* r2d2.Trainer._train_on_batches, the loop between the network forwards and the loss (srl/algorithms/r2d2/r2d2.py:150-203),
 * for n_seq stored sequences of seq_len (<= 128) steps: q_online / q_target [n_seq][seq_len+1][n_actions] float32 (the online
 * and target Q of every step after burn-in), actions int32 / mu (behaviour probability) / rewards float64 / dones uint8
 * [n_seq][seq_len] -> target_out float64 [n_seq][seq_len] (the regression target of Q_online(s_t)[a_t]), td_mean_out [n_seq]
 * (mean TD error of the sequence: the priority input, :204) and td_kind_out (0: the reference holds it as float32, 1: float64).
 * Bit-exact with the reference's numpy/python arithmetic (tests/golden/r2d2_targets.npz). */
int srlx_sequence_targets(const float* q_online_dev, const float* q_target_dev, const int32_t* actions_dev, const double* mu_dev,
                          const double* rewards_dev, const unsigned char* dones_dev, double* target_out_dev, double* td_mean_out_dev,
                          unsigned char* td_kind_out_dev, uint32_t n_seq, uint32_t seq_len, uint32_t n_actions, double discount,
                          double retrace_h, int enable_double_dqn, int enable_rescale, int enable_retrace, uintptr_t cuda_stream);

/* ---- engine (R1-R6, R8-R12) ------------------------------------------------------------------------------ */
/* Zero the counters, mark every env for reset, clear ring flags and the tree. */
int srlx_engine_reset(const srlx_engine* eng, uintptr_t cuda_stream);
/* n_steps x { one vector step of all E envs (policy -> env.step -> ring write -> replay add) followed by
 * updates_per_step trainer updates }.  training==0: evaluation rollouts (test_epsilon in eng->epsilon, nothing stored). */
int srlx_engine_run(const srlx_engine* eng, uint32_t n_steps, uint32_t updates_per_step, int training,
                    uintptr_t cuda_stream);
/* Which kernel srlx_learn runs for this engine: 1 = the short-critical-path kernel for single-hidden-layer networks
 * (csrc/learner_fast.cu; *cluster_size CTAs, *smem_bytes of shared memory each), 0 = the generic kernel (csrc/learner.cu). */
int srlx_learner_info(const srlx_engine* eng, int* cluster_size, size_t* smem_bytes);
/* Bytes the optional srlx_engine.tree_blk buffer needs for a SumTree of `capacity` leaves. */
size_t srlx_tree_blk_bytes(uint64_t capacity);
/* The two halves separately (parity tests drive them one at a time). */
int srlx_vec_step(const srlx_engine* eng, int training, uintptr_t cuda_stream);
int srlx_learn(const srlx_engine* eng, uint32_t n_updates, uintptr_t cuda_stream);
/* ---- plug-in seams for the reference's own host loop (srl.Runner / core_play.play, srl/base/run/core_play.py:115-214) ------------
 * srlx_ext_step <- RLWorker.on_step -> RLMemory.add (srl/algorithms/dqn/dqn.py:213-246, rainbow.py:333-400,
 *                  srl/rl/memories/priority_replay_buffer.py:205-217): one row of E externally produced records (s, s', a, r,
 *                  terminated, episode-ended) is written at the ring cursor and added to the replay memory exactly as a device
 *                  vector step would (n-step windows are rebuilt by index, so records must arrive in trajectory order per column).
 * srlx_env_reset_obs / srlx_env_step_actions <- EnvBase.reset / EnvBase.step (srl/base/env/base.py:60-137) for the closed-form
 *                  envs: reset-if-needed (force != 0: always) + observation; one step per env copy with CALLER-supplied actions
 *                  -> (next observation [E][D] float32, raw reward f64, terminated, truncated at trunc_limit).  No policy, no ring. */
int srlx_ext_step(const srlx_engine* eng, const float* obs_dev, const float* next_obs_dev, const int32_t* action_dev,
                  const float* reward_dev, const unsigned char* term_dev, const unsigned char* done_dev, uintptr_t cuda_stream);
/* srlx_ext_step with the next state's invalid-action mask of every record (bit a = action a invalid; NULL = none) */
int srlx_ext_step_masked(const srlx_engine* eng, const float* obs_dev, const float* next_obs_dev, const int32_t* action_dev,
                         const float* reward_dev, const unsigned char* term_dev, const unsigned char* done_dev,
                         const uint32_t* next_invalid_dev, uintptr_t cuda_stream);
int srlx_env_reset_obs(const srlx_engine* eng, int force, float* out_obs_dev, uintptr_t cuda_stream);
int srlx_env_step_actions(const srlx_engine* eng, const int32_t* actions_dev, float* out_obs_dev, double* out_reward_dev,
                          unsigned char* out_term_dev, unsigned char* out_trunc_dev, uintptr_t cuda_stream);
/* ---- exchange buffers of the data-parallel learner (one per rank; peers write gradients and flags into them over NVLink) ------
 * srlx_dp_bytes: bytes one buffer needs for this engine.  srlx_dp_alloc: cudaMalloc + zero + IPC handle (64 bytes) for the other
 * processes of the node; srlx_dp_open / srlx_dp_close: map / unmap a peer's buffer from its handle; srlx_dp_free: release an own
 * buffer; srlx_dp_enable_peer: cudaDeviceEnablePeerAccess both ways between two devices driven by ONE process. */
size_t srlx_dp_bytes(const srlx_engine* eng);
int srlx_dp_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]);
int srlx_dp_open(const unsigned char handle[64], void** ptr_out);
int srlx_dp_close(void* ptr);
int srlx_dp_free(void* ptr);
int srlx_dp_enable_peer(int device_a, int device_b);
/* Inference seam: q_out[n][A] = Q(obs[n][D]) with `params` (pred_q) or `target` (pred_target_q); noise_call_id is the
 * NoisyLinear draw to use (ignored when !noisy; kind = 3). */
int srlx_qnet_forward(const srlx_engine* eng, int use_target, const float* obs_dev, uint32_t n, uint64_t noise_call_id,
                      float* q_out_dev, uintptr_t cuda_stream);

/* ---- Image observation pipeline + conv Q-network -- csrc/imageq.cu -------------------------------------------
 * srlx_image_process <- ImageProcessor.remap_observation (srl/rl/processors/image_processor.py:104-154) for a BATCH of uint8 frames:
 *   colour conversion (cv2.cvtColor COLOR_RGB2GRAY, or gray -> 3 equal channels), trimming, cv2.resize (INTER_LINEAR, the 11-bit
 *   fixed-point form OpenCV uses for uint8) and the normalisation, in ONE pass over the frames (read uint8 once, write once).
 *   Bit-exact with cv2 4.13 / the reference class (tests/golden/image_processor.npz).
 * srlx_image_linear_table: the per-axis index / coefficient table of cv2.resize (host; border_reset = 1 for x, 0 for y). */
typedef struct srlx_image_proc {
  int32_t src_h, src_w, src_c;        /* source frames uint8 [n][src_h][src_w][src_c], src_c = 1 or 3 */
  int32_t top, left, trim_h, trim_w;  /* trimming window (0, 0, src_h, src_w: none) */
  int32_t out_h, out_w, out_c;        /* out_c = 1 or 3; out_c != src_c converts */
  int32_t resize;                     /* != 0: through the tables; 0: out_h == trim_h and out_w == trim_w */
  int32_t normalize;                  /* 0: uint8 out; 1: float32 v / max_val ("0to1"); 2: float32 v * 2 / max_val - 1 ("-1to1") */
  float max_val;
  const int32_t* x_idx; const int32_t* x_coef; const int32_t* y_idx; const int32_t* y_coef; /* device [out_w], [out_w][2], [out_h], [out_h][2] */
} srlx_image_proc;
int srlx_image_linear_table(int32_t dst, int32_t src, int border_reset, int32_t* idx_out_host, int32_t* coef_out_host);
/* out_dev: uint8 or float32 [n][out_h][out_w][out_c], frame f at out_dev + f * out_frame_stride elements */
int srlx_image_process(const srlx_image_proc* p, const unsigned char* src_dev, uint32_t n, void* out_dev, uint64_t out_frame_stride,
                       uintptr_t cuda_stream);

/* Conv Q-network of the image configs: InputImageBlock (reshape + DQNImageBlock: Conv2d layers with replicate padding and ReLU,
 * srl/rl/torch_/blocks/dqn_image_block.py:10-62, input_image_block.py:42-75) -> Flatten -> MLP hidden block -> Linear(A)
 * (srl/algorithms/dqn/model_torch.py:17-29), and its trainer (model_torch.py:75-131 with dqn.py:143-173 calc_target_q): double DQN,
 * rescaling, Huber loss with the IS weight inside, torch Adam, priorities |target - q|, target sync at train_count % interval == 0.
 * Every map is im2col (replicate padding = index clamp) + the strided GEMM family of csrc/gemm.cuh (3 x TF32 tensor-core tiles at
 * fp32 accuracy).  Parameter layout (one flat fp32 buffer, bias = LAST COLUMN of every block):
 *   conv l   W[F_l][k*k*C + 1]   column order (kh, kw, c) when the layer's input is channel-fastest (NHWC: every layer after the
 *                                first, and a first layer fed NHWC frames), (c, kh, kw) when it is NCHW (a stack of gray frames)
 *   dense l  W[out_l][k_l + 1]   dense 0 reads the last conv output flattened as (h, w, c)
 * (netspec.ImageNetSpec converts from / to the reference's state_dict.)  Activations are NHWC. */
#define SRLX_MAX_CONV 4
typedef struct srlx_imageq {
  int32_t in_c, in_h, in_w;
  int64_t in_sb, in_sc, in_sh, in_sw;   /* element strides of a state batch */
  int32_t in_u8; float in_max_val;      /* != 0: states are uint8 frames, value = v / in_max_val formed on the fly (the "0to1" normalisation) */
  int32_t n_conv;
  int32_t conv_f[SRLX_MAX_CONV], conv_k[SRLX_MAX_CONV], conv_s[SRLX_MAX_CONV], conv_p[SRLX_MAX_CONV];
  int32_t conv_oh[SRLX_MAX_CONV], conv_ow[SRLX_MAX_CONV], conv_off[SRLX_MAX_CONV];
  int32_t n_dense;                      /* hidden layers + the output layer */
  int32_t dense_out[SRLX_MAX_LAYERS], dense_k[SRLX_MAX_LAYERS], dense_off[SRLX_MAX_LAYERS];
  int32_t n_actions, n_params, batch_cap;
  int32_t enable_double_dqn, enable_rescale;
  uint32_t target_update_interval;
  int32_t dueling, duel_hidden;         /* SRLX_DUEL_*; != NONE: the rainbow head (srl/rl/torch_/blocks/dueling_network.py:8-59) as the last two
                                           dense layers -- [value-hidden (H) ; advantage-hidden (H)] of width 2H, then 1 + A rows over 2H inputs whose
                                           off-branch blocks are structural zeros (row 0 reads [0, H), rows 1..A read [H, 2H); Adam never moves them)
                                           -- followed by Q = V + Adv - mean / max / nothing */
  int32_t target_f32;                   /* != 0: the target arithmetic of rainbow_nomultisteps.py:10-43 (float32 throughout); 0: dqn.py:143-173
                                           (numpy promotes reward + undone * discount * maxq to float64) */
  double discount, lr, adam_beta1, adam_beta2, adam_eps;
  float* params; float* target; float* adam_m; float* adam_v; float* grads;  /* [n_params] each */
  uint64_t* counters;                   /* device [4]: train_count, adam_step, sync_count, reserved */
  float* ws; uint64_t ws_floats;        /* workspace, srlx_imageq_ws_floats(q) floats */
} srlx_imageq;
size_t srlx_sizeof_imageq(void);
uint64_t srlx_imageq_ws_floats(const srlx_imageq* q);
/* write the constant columns of the workspace (run once after allocation, and after any foreign write to ws) */
int srlx_imageq_init(const srlx_imageq* q, uintptr_t cuda_stream);
/* q_out[n][A] = Q(state[n]) with `params` (pred_q) or `target` (pred_target_q); n <= batch_cap */
int srlx_imageq_forward(const srlx_imageq* q, int use_target, const void* state_dev, uint32_t n, float* q_out_dev, uintptr_t cuda_stream);
/* one Trainer.train() on a batch in device memory: action int32 [B], reward / undone / weights float32 [B] ->
 * priorities_out [B] (|target_q - q|), loss_out [1], target_q_out [B] (may be NULL).  phases: 1 = forward + backward (gradient left
 * in q->grads), 2 = Adam + target sync + counters, 3 = both. */
/* C[M][N] = A . B with arbitrary strides on the tcgen05 tiles at fp32 accuracy (3 x TF32 through tcgen05.mma.kind::tf32, accumulator in
 * tensor memory; csrc/gemm_tc3.cuh) -- the GEMM every map of the conv Q-network runs on; test tap with srlx_sgemm's arguments plus an
 * optional split-K workspace (ws_dev may be NULL: no split). */
int srlx_sgemm_tc3(const float* a_dev, long long sa_m, long long sa_k, const float* b_dev, long long sb_k, long long sb_n, float* c_dev,
                   long long ldc, int M, int N, int K, int relu, int accumulate, float* ws_dev, uint64_t ws_floats, uintptr_t cuda_stream);
int srlx_imageq_train(const srlx_imageq* q, const void* state_dev, const void* n_state_dev, const int32_t* action_dev,
                      const float* reward_dev, const float* undone_dev, const float* weights_dev, uint32_t batch,
                      float* priorities_out_dev, float* loss_out_dev, float* target_q_out_dev, int phases, uintptr_t cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* SRLX_H_ */
