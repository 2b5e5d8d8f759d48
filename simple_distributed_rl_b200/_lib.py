"""ctypes binding of libsrlx.so (C ABI: include/srlx.h).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRLX_LIB") or os.path.join(_HERE, "libsrlx.so")  # SRLX_LIB: diagnostic builds (phase clocks)

SRLX_MAX_LAYERS = 6
ENV_GRID, ENV_CARTPOLE, ENV_PENDULUM, ENV_EXTERNAL = 0, 1, 2, 3
RETURNS_GAE, RETURNS_MC = 0, 1  # srlx_returns_scan methods
DUEL_NONE, DUEL_AVERAGE, DUEL_MAX, DUEL_NAIVE = 0, 1, 2, 3
MEM_UNIFORM, MEM_PROPORTIONAL = 0, 1
NOISE_KIND_ROLLOUT, NOISE_KIND_TRAIN, NOISE_KIND_PRED = 0, 1, 3
STREAM_ENV_RESET, STREAM_ENV_STEP, STREAM_POLICY, STREAM_NOISE, STREAM_SAMPLE, STREAM_PAD_ACTION, STREAM_UNIFORM_SAMPLE = 1, 2, 3, 4, 5, 6, 7


class SrlxNet(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32),
        ("in_dim", C.c_int32),
        ("out_dim", C.c_int32 * SRLX_MAX_LAYERS),
        ("k_dim", C.c_int32 * SRLX_MAX_LAYERS),
        ("w_off", C.c_int32 * SRLX_MAX_LAYERS),
        ("b_off", C.c_int32 * SRLX_MAX_LAYERS),
        ("n_params", C.c_int32),
        ("n_actions", C.c_int32),
        ("dueling", C.c_int32),
        ("noisy", C.c_int32),
        ("layer_noisy", C.c_int32 * SRLX_MAX_LAYERS),
    ]


class SrlxState(C.Structure):
    _fields_ = [
        ("vec_steps", C.c_uint64),
        ("total_step", C.c_uint64),
        ("train_count", C.c_uint64),
        ("episode_count", C.c_uint64),
        ("sync_count", C.c_uint64),
        ("adam_step", C.c_uint64),
        ("mem_size", C.c_uint64),
        ("sample_retries", C.c_uint64),
        ("max_priority", C.c_double),
        ("episode_reward_sum", C.c_double),
        ("last_loss", C.c_double),
        ("loss_sum", C.c_double),
        ("episode_len_sum", C.c_uint64),
        ("reserved", C.c_uint64 * 3),
    ]


_P = C.c_void_p


class SrlxEngine(C.Structure):
    _fields_ = [
        ("env_id", C.c_int32), ("n_envs", C.c_int32), ("obs_dim", C.c_int32), ("n_actions", C.c_int32),
        ("ring_rows", C.c_int32), ("multisteps", C.c_int32), ("batch_size", C.c_int32), ("mem_kind", C.c_int32),
        ("enable_double_dqn", C.c_int32), ("enable_rescale", C.c_int32), ("enable_reward_clip", C.c_int32),
        ("has_duplicate", C.c_int32), ("target_update_interval", C.c_int32), ("trunc_limit", C.c_int32),
        ("trunc_overrides_term", C.c_int32), ("presample", C.c_int32),
        ("seed", C.c_uint64), ("warmup_size", C.c_uint64),
        ("epsilon", C.c_double), ("discount", C.c_double), ("lr", C.c_double),
        ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
        ("retrace_h", C.c_double),
        ("per_alpha", C.c_double), ("per_beta_initial", C.c_double), ("per_beta_steps", C.c_double), ("per_epsilon", C.c_double),
        ("reward_shift", C.c_double), ("reward_scale", C.c_double), ("huber_delta", C.c_double),
        ("grid_w", C.c_int32), ("grid_h", C.c_int32),
        ("grid_field", C.c_int8 * 64),
        ("grid_n_starts", C.c_int32),
        ("grid_starts", C.c_int32 * 16),
        ("grid_slip_cdf", C.c_double * 16),
        ("grid_slip_action", C.c_int32 * 4),
        ("grid_move_reward", C.c_double), ("grid_goal_reward", C.c_double), ("grid_hole_reward", C.c_double),
        ("act_tbl", C.c_double * 16),
        ("net", SrlxNet),
        ("state", _P), ("env_state", _P), ("env_step_num", _P), ("env_episode", _P), ("env_ep_reward", _P),
        ("env_needs_reset", _P), ("env_first_ep_reward", _P), ("env_last_ep_len", _P),
        ("ring_obs", _P), ("ring_next_obs", _P), ("ring_action", _P), ("ring_reward", _P), ("ring_term", _P),
        ("ring_done", _P),
        ("tree", _P), ("tree_scratch", _P),
        ("params", _P), ("params_sigma", _P), ("target", _P), ("target_sigma", _P), ("adam_m", _P), ("adam_v", _P),
        ("dbg_q", _P), ("dbg_action", _P), ("dbg_sample_idx", _P), ("dbg_weights", _P), ("dbg_target_q", _P),
        ("dbg_q_sa", _P), ("dbg_grads", _P), ("dbg_windows", _P), ("dbg_clock", _P),
        ("noise_scratch", _P), ("noise_scratch_bytes", C.c_uint64),
        ("tree_blk", _P), ("tree_blk_bytes", C.c_uint64),
        ("eps_end", C.c_double), ("eps_phase_steps", C.c_uint64),
        ("learner_seed", C.c_uint64), ("dp_world", C.c_int32), ("dp_rank", C.c_int32), ("dp_peer", _P * 8), ("dp_bytes", C.c_uint64),
        ("ring_invalid", _P), ("eps_table", _P), ("eps_table_len", C.c_uint64),
    ]


class SrlxPpoState(C.Structure):
    _fields_ = [("train_count", C.c_uint64), ("adam_step", C.c_uint64), ("policy_loss", C.c_double), ("value_loss", C.c_double),
                ("entropy_loss", C.c_double), ("grad_norm", C.c_double), ("reserved", C.c_uint64 * 2)]


class SrlxPpo(C.Structure):
    _fields_ = [
        ("env", SrlxEngine), ("net_v", SrlxNet), ("net_p", SrlxNet),
        ("n_params", C.c_int32), ("continuous", C.c_int32), ("horizon", C.c_int32), ("batch_size", C.c_int32),
        ("baseline_type", C.c_int32), ("surrogate_clip", C.c_int32), ("enable_value_clip", C.c_int32), ("state_normalized", C.c_int32),
        ("method", C.c_int32), ("reward_clip_enable", C.c_int32),
        ("lr_decay_steps", C.c_uint64),
        ("discount", C.c_double), ("gae_discount", C.c_double), ("policy_clip_range", C.c_double), ("value_clip_range", C.c_double),
        ("lr", C.c_double), ("lr_decay_rate", C.c_double), ("value_loss_weight", C.c_double), ("entropy_weight", C.c_double),
        ("grad_clip_norm", C.c_double),
        ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
        ("log_scale_lo", C.c_double), ("log_scale_hi", C.c_double), ("action_low", C.c_double), ("action_high", C.c_double),
        ("reward_clip_lo", C.c_double), ("reward_clip_hi", C.c_double),
        ("params", _P), ("adam_m", _P), ("adam_v", _P),
        ("buf_obs", _P), ("buf_action", _P), ("buf_v", _P), ("buf_logp", _P), ("buf_reward", _P), ("buf_done", _P), ("buf_vnew", _P),
        ("buf_ret", _P), ("buf_valid", _P), ("pstate", _P), ("dbg_idx", _P), ("dbg_grads", _P), ("grad_scratch", _P),
    ]


class SrlxR2d2(C.Structure):
    _fields_ = [
        ("env", SrlxEngine),
        ("lstm_units", C.c_int32), ("burnin", C.c_int32), ("seq_len", C.c_int32), ("enable_retrace", C.c_int32),
        ("n_head", C.c_int32), ("dueling", C.c_int32),
        ("head_out", C.c_int32 * SRLX_MAX_LAYERS), ("head_k", C.c_int32 * SRLX_MAX_LAYERS), ("head_off", C.c_int32 * SRLX_MAX_LAYERS),
        ("lstm_off", C.c_int32), ("n_params", C.c_int32), ("duel_hidden", C.c_int32), ("no_persistent", C.c_int32),
        ("test_epsilon", C.c_double),
        ("params", _P), ("target", _P), ("adam_m", _P), ("adam_v", _P), ("grads", _P),
        ("cursor", _P), ("ring_obs", _P), ("ring_next_obs", _P), ("ring_action", _P), ("ring_prob", _P), ("ring_reward", _P),
        ("ring_done", _P), ("ring_tstep", _P), ("ring_h", _P), ("ring_c", _P),
        ("roll_xh", _P), ("roll_h", _P), ("roll_c", _P), ("roll_gates", _P), ("roll_act", _P * SRLX_MAX_LAYERS), ("roll_reset", _P),
        ("new_c0", _P), ("new_n", _P), ("add_idx", _P), ("add_pri", _P),
        ("xh", _P), ("cbuf", _P), ("gates", _P), ("dgates", _P), ("dc", _P), ("gemm_ws", _P), ("gemm_ws_floats", C.c_uint64), ("bar", _P),
        ("act", _P * SRLX_MAX_LAYERS), ("dact", _P * SRLX_MAX_LAYERS), ("dh", _P), ("q", _P), ("sel", _P), ("weights", _P),
        ("b_actions", _P), ("b_mu", _P), ("b_rewards", _P), ("b_dones", _P), ("b_target", _P), ("b_tdmean", _P), ("b_tdkind", _P),
    ]


SRLX_MAX_CONV = 4


class SrlxImageProc(C.Structure):
    _fields_ = [("src_h", C.c_int32), ("src_w", C.c_int32), ("src_c", C.c_int32),
                ("top", C.c_int32), ("left", C.c_int32), ("trim_h", C.c_int32), ("trim_w", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32),
                ("resize", C.c_int32), ("normalize", C.c_int32), ("max_val", C.c_float),
                ("x_idx", _P), ("x_coef", _P), ("y_idx", _P), ("y_coef", _P)]


class SrlxImageQ(C.Structure):
    _fields_ = [
        ("in_c", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("in_sb", C.c_int64), ("in_sc", C.c_int64), ("in_sh", C.c_int64), ("in_sw", C.c_int64),
        ("in_u8", C.c_int32), ("in_max_val", C.c_float),
        ("n_conv", C.c_int32),
        ("conv_f", C.c_int32 * SRLX_MAX_CONV), ("conv_k", C.c_int32 * SRLX_MAX_CONV), ("conv_s", C.c_int32 * SRLX_MAX_CONV),
        ("conv_p", C.c_int32 * SRLX_MAX_CONV), ("conv_oh", C.c_int32 * SRLX_MAX_CONV), ("conv_ow", C.c_int32 * SRLX_MAX_CONV),
        ("conv_off", C.c_int32 * SRLX_MAX_CONV),
        ("n_dense", C.c_int32),
        ("dense_out", C.c_int32 * SRLX_MAX_LAYERS), ("dense_k", C.c_int32 * SRLX_MAX_LAYERS), ("dense_off", C.c_int32 * SRLX_MAX_LAYERS),
        ("n_actions", C.c_int32), ("n_params", C.c_int32), ("batch_cap", C.c_int32),
        ("enable_double_dqn", C.c_int32), ("enable_rescale", C.c_int32), ("target_update_interval", C.c_uint32),
        ("dueling", C.c_int32), ("duel_hidden", C.c_int32), ("target_f32", C.c_int32),
        ("discount", C.c_double), ("lr", C.c_double), ("adam_beta1", C.c_double), ("adam_beta2", C.c_double), ("adam_eps", C.c_double),
        ("params", _P), ("target", _P), ("adam_m", _P), ("adam_v", _P), ("grads", _P),
        ("counters", _P), ("ws", _P), ("ws_floats", C.c_uint64),
    ]


class SrlxSeam(C.Structure):
    _fields_ = [("tree", _P), ("meta", _P), ("ops_idx", _P), ("ops_val", _P), ("out_tree_idx", _P), ("out_weights", _P), ("flag", _P),
                ("capacity", C.c_uint64), ("alpha", C.c_double), ("epsilon", C.c_double), ("beta_initial", C.c_double),
                ("beta_steps", C.c_double), ("has_duplicate", C.c_int32), ("reserved", C.c_int32)]


class SrlxError(RuntimeError):
    pass


# every symbol include/srlx.h declares: (name, restype, argtypes)
_u64, _u32, _i32, _dbl, _sz, _uptr = C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_size_t, C.c_size_t
SYMBOLS = [
    ("srlx_version", C.c_int, []),
    ("srlx_last_error", C.c_char_p, []),
    ("srlx_sizeof_engine", _sz, []),
    ("srlx_sizeof_state", _sz, []),
    ("srlx_sizeof_net", _sz, []),
    ("srlx_launch_count", _u64, []),
    ("srlx_philox_words", C.c_int, [_u64, _u32, _u32, _u32, _u32, _P, _sz, _uptr]),
    ("srlx_noise_fill", C.c_int, [_u64, _u32, _u64, _P, _sz, _uptr]),
    ("srlx_dbg_pow", C.c_int, [_P, _dbl, _P, _sz, _uptr]),
    ("srlx_tree_clear", C.c_int, [_P, _u64, _P, _uptr]),
    ("srlx_tree_add", C.c_int, [_P, _u64, _P, _P, _u64, _dbl, _dbl, _i32, _uptr]),
    ("srlx_tree_sample", C.c_int, [_P, _u64, _P, _u32, _u64, _dbl, _dbl, _i32, _u64, _P, _u32, _P, _P, _P, _uptr]),
    ("srlx_tree_update", C.c_int, [_P, _u64, _P, _P, _P, _u32, _dbl, _dbl, _uptr]),
    ("srlx_tree_seam", C.c_int, [_P, _u64, _P, _P, _P, _u32, _dbl, _dbl, _u32, _u64, _dbl, _dbl, _i32, _u64, _P, _u32, _P, _P, _P, _u64, _uptr]),
    ("srlx_tree_seam_desc", C.c_int, [C.POINTER(SrlxSeam), _u32, _u32, _u64, _u64, _P, _u32, _u64, _uptr]),
    ("srlx_host_alloc", C.c_int, [_sz, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("srlx_host_free", C.c_int, [_P]),
    ("srlx_tree_retrieve", C.c_int, [_P, _u64, _P, _u32, _P, _uptr]),
    ("srlx_engine_reset", C.c_int, [C.POINTER(SrlxEngine), _uptr]),
    ("srlx_engine_run", C.c_int, [C.POINTER(SrlxEngine), _u32, _u32, _i32, _uptr]),
    ("srlx_vec_step", C.c_int, [C.POINTER(SrlxEngine), _i32, _uptr]),
    ("srlx_learn", C.c_int, [C.POINTER(SrlxEngine), _u32, _uptr]),
    ("srlx_learner_info", C.c_int, [C.POINTER(SrlxEngine), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    ("srlx_tree_blk_bytes", _sz, [_u64]),
    ("srlx_qnet_forward", C.c_int, [C.POINTER(SrlxEngine), _i32, _P, _u32, _u64, _P, _uptr]),
    ("srlx_sizeof_ppo", _sz, []),
    ("srlx_sizeof_ppo_state", _sz, []),
    ("srlx_ppo_vec_step", C.c_int, [C.POINTER(SrlxPpo), _i32, _uptr]),
    ("srlx_ppo_values", C.c_int, [C.POINTER(SrlxPpo), _P, _u64, _P, _uptr]),
    ("srlx_ppo_finish_rollout", C.c_int, [C.POINTER(SrlxPpo), _uptr]),
    ("srlx_ppo_learn", C.c_int, [C.POINTER(SrlxPpo), _u32, _uptr]),
    ("srlx_sizeof_r2d2", _sz, []),
    ("srlx_r2d2_vec_step", C.c_int, [C.POINTER(SrlxR2d2), _i32, _uptr]),
    ("srlx_r2d2_learn", C.c_int, [C.POINTER(SrlxR2d2), _u32, _uptr]),
    ("srlx_r2d2_learn_phase", C.c_int, [C.POINTER(SrlxR2d2), _u32, _i32, _uptr]),
    ("srlx_r2d2_forward", C.c_int, [C.POINTER(SrlxR2d2), _i32, _P, _P, _P, _u32, _P, _P, _P, _uptr]),
    ("srlx_sgemm", C.c_int, [_P, C.c_longlong, C.c_longlong, _P, C.c_longlong, C.c_longlong, _P, C.c_longlong, _i32, _i32, _i32, _i32, _i32, _uptr]),
    ("srlx_dense_bf16_tc", C.c_int, [_P, _i32, _P, _i32, _P, _P, _i32, _i32, _i32, _i32, _i32, _i32, _uptr]),
    ("srlx_qnet_tc_workspace_bytes", _sz, [C.POINTER(SrlxEngine), _u32]),
    ("srlx_qnet_forward_tc", C.c_int, [C.POINTER(SrlxEngine), _i32, _P, _u32, _u64, _P, _P, _sz, _uptr]),
    ("srlx_rank_scratch_bytes", _sz, [_u64]),
    ("srlx_rank_sample", C.c_int, [_P, _u64, _u32, _dbl, _dbl, _u32, _P, _u32, _u64, _u64, _i32, _P, _P, _P, _P, _P, _uptr]),
    ("srlx_rank_update", C.c_int, [_P, _P, _P, _u32, _uptr]),
    ("srlx_rank_argsort", C.c_int, [_P, _u64, _u32, _P, _P, _uptr]),
    ("srlx_dp_bytes", _sz, [C.POINTER(SrlxEngine)]),
    ("srlx_dp_alloc", C.c_int, [_sz, C.POINTER(C.c_void_p), C.c_char_p]),
    ("srlx_dp_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    ("srlx_dp_close", C.c_int, [_P]),
    ("srlx_dp_free", C.c_int, [_P]),
    ("srlx_dp_enable_peer", C.c_int, [_i32, _i32]),
    ("srlx_ext_step", C.c_int, [C.POINTER(SrlxEngine), _P, _P, _P, _P, _P, _P, _uptr]),
    ("srlx_ext_step_masked", C.c_int, [C.POINTER(SrlxEngine), _P, _P, _P, _P, _P, _P, _P, _uptr]),
    ("srlx_env_reset_obs", C.c_int, [C.POINTER(SrlxEngine), _i32, _P, _uptr]),
    ("srlx_env_step_actions", C.c_int, [C.POINTER(SrlxEngine), _P, _P, _P, _P, _P, _uptr]),
    ("srlx_sequence_targets", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _u32, _u32, _u32, _dbl, _dbl, _i32, _i32, _i32, _uptr]),
    ("srlx_sgemm_tc3", C.c_int, [_P, C.c_longlong, C.c_longlong, _P, C.c_longlong, C.c_longlong, _P, C.c_longlong, _i32, _i32, _i32, _i32, _i32, _P, _u64, _uptr]),
    ("srlx_image_linear_table", C.c_int, [_i32, _i32, _i32, _P, _P]),
    ("srlx_image_process", C.c_int, [C.POINTER(SrlxImageProc), _P, _u32, _P, _u64, _uptr]),
    ("srlx_sizeof_imageq", _sz, []),
    ("srlx_imageq_ws_floats", _u64, [C.POINTER(SrlxImageQ)]),
    ("srlx_imageq_init", C.c_int, [C.POINTER(SrlxImageQ), _uptr]),
    ("srlx_imageq_forward", C.c_int, [C.POINTER(SrlxImageQ), _i32, _P, _u32, _P, _uptr]),
    ("srlx_imageq_train", C.c_int, [C.POINTER(SrlxImageQ), _P, _P, _P, _P, _P, _P, _u32, _P, _P, _P, _i32, _uptr]),
    ("srlx_returns_scan", C.c_int, [_P, _P, _P, _P, _P, _P, _P, _u32, _u32, _dbl, _dbl, _i32, _i32, _i32, _dbl, _dbl, _uptr]),
]

_lib = None


def load():
    """Load libsrlx.so (built by `make -C simple_distributed_rl_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrlxError(
            f"{LIB_PATH} not found: the CUDA extension is required (there is no CPU fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C simple_distributed_rl_b200/csrc`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.srlx_sizeof_ppo() != C.sizeof(SrlxPpo) or lib.srlx_sizeof_ppo_state() != C.sizeof(SrlxPpoState):
        raise SrlxError(f"ABI mismatch: C sizes ppo/ppo_state = {lib.srlx_sizeof_ppo()}/{lib.srlx_sizeof_ppo_state()}, "
                        f"ctypes = {C.sizeof(SrlxPpo)}/{C.sizeof(SrlxPpoState)}")
    if lib.srlx_sizeof_imageq() != C.sizeof(SrlxImageQ):
        raise SrlxError(f"ABI mismatch: C size imageq = {lib.srlx_sizeof_imageq()}, ctypes = {C.sizeof(SrlxImageQ)}")
    if lib.srlx_sizeof_r2d2() != C.sizeof(SrlxR2d2):
        raise SrlxError(f"ABI mismatch: C size r2d2 = {lib.srlx_sizeof_r2d2()}, ctypes = {C.sizeof(SrlxR2d2)}")
    if lib.srlx_sizeof_engine() != C.sizeof(SrlxEngine) or lib.srlx_sizeof_state() != C.sizeof(SrlxState) or lib.srlx_sizeof_net() != C.sizeof(SrlxNet):
        raise SrlxError(
            f"ABI mismatch: C sizes engine/state/net = {lib.srlx_sizeof_engine()}/{lib.srlx_sizeof_state()}/{lib.srlx_sizeof_net()}, "
            f"ctypes = {C.sizeof(SrlxEngine)}/{C.sizeof(SrlxState)}/{C.sizeof(SrlxNet)}"
        )
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SrlxError(f"libsrlx call failed ({rc}): {load().srlx_last_error().decode()}")
