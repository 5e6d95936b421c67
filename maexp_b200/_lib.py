"""ctypes binding of libmaexp_b200.so (the C-ABI declared in include/maexp_b200.h).

There is no CPU fallback: importing this module without the built CUDA library raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAEXP_LIB") or os.path.join(HERE, "libmaexp_b200.so")  # MAEXP_LIB: experiment builds (tools/)

MAX_AGENTS, MAX_TYPES, MAX_RINGS = 16, 8, 4


class MaexpError(RuntimeError):
    pass


class EnvCfg(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("num_agents", C.c_int32), ("type_objects", C.c_int32),
        ("size_x", C.c_int32), ("size_y", C.c_int32), ("visible_range", C.c_int32), ("past_step", C.c_int32),
        ("agent_view", C.c_int32), ("object_view", C.c_int32), ("area_mask", C.c_int32), ("noise", C.c_int32),
        ("reward_mode", C.c_int32), ("max_episode_length", C.c_int32), ("num_objects", C.c_int32),
        ("keep_objects_num", C.c_int32), ("shuffle_init_xys", C.c_int32), ("shared_max_is_mean", C.c_int32),
        ("auto_reset", C.c_int32), ("num_destinations", C.c_int32), ("random_destinations", C.c_int32),
        ("seed", C.c_uint32), ("first_global_env_id", C.c_uint32),
        ("init_xys", (C.c_int32 * 2) * MAX_AGENTS), ("task_mask", C.c_int32 * MAX_AGENTS),
        ("task_order", (C.c_int32 * MAX_TYPES) * MAX_AGENTS),
        ("flip_noise_range", C.c_float * MAX_RINGS), ("std_distribution", C.c_float * MAX_RINGS),
        ("noise_threshold", C.c_float * MAX_RINGS), ("env_kind", C.c_int32),
        ("dynamic_obstacles", C.c_int32), ("num_init_xys", C.c_int32),
    ]


class Tables(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wall", "objects_area", "noise_area", "destination", "los")]


class ObsOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates",
                                           "positions", "objects", "destination", "actions_taken", "rewards_i8", "done_env",
                                           "heatmaps", "objects_bits")]


class ReplayView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "actions", "rewards", "dones", "was_reset", "positions")] + \
               [(n, C.c_int32) for n in ("num_slots", "num_envs", "num_agents", "slot_floats", "head", "count")] + \
               [(n, C.c_void_p) for n in ("objects_bits", "destination", "dest_table")] + \
               [(n, C.c_int32) for n in ("type_objects", "size_x")]


class Batch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "next_obs", "actions", "rewards", "dones", "picked", "pos", "next_pos", "error_flag",
                                          "obj_rows", "next_obj_rows", "agents", "next_agents", "dest_rows", "next_dest_rows")]

REPLAY_ERR_NO_VALID, REPLAY_ERR_INDEX_RANGE, REPLAY_ERR_INDEX_INVALID, REPLAY_ERR_NOT_UNIQUE = 1, 2, 4, 8


class LearnerCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("kind", "num_agents", "batch_size", "channels", "visible_range", "num_actions",
                                          "hidden", "conv1_channels", "conv2_channels", "num_atoms", "obs_stride")] + \
               [(n, C.c_float) for n in ("gamma", "v_min", "v_max", "delta_z", "learning_rate", "alpha",
                                          "one_minus_alpha", "eps", "max_norm")] + \
               [(n, C.c_int32) for n in ("num_quantiles_dash", "num_cosines", "embed_dim", "num_heads", "mlp_hidden",
                                          "block_loop", "area_mask", "size_x", "size_y", "patch_size", "local_patch_size",
                                          "relative_patch_size", "iqn_trunk", "optimizer")] + \
               [(n, C.c_float) for n in ("beta1", "beta2")] + [("attention_hard", C.c_int32), ("af_lambd", C.c_float)] + \
               [(n, C.c_int32) for n in ("objects_channel", "agents_channel", "destination_channel", "type_objects",
                                          "team_agents")]


class DpPeers(C.Structure):
    _fields_ = [("world_size", C.c_int32), ("rank", C.c_int32), ("grads", C.c_void_p * 8), ("partials", C.c_void_p * 8),
                ("signals", C.c_void_p * 8), ("multicast_grads", C.c_void_p)]


LEARNER_KINDS = {"dqn_conv_mlp": 0, "c51": 1, "qr": 2, "dqn_mlp": 3, "iqn": 4, "da3": 5, "da6": 6}
BUF_ONLINE, BUF_TARGET, BUF_GRADS, BUF_SQUARE_AVG, BUF_SUPPORT, BUF_EXP_AVG, BUF_ADAM_STEPS = range(7)

# every symbol include/maexp_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "maexp_env_create": (C.c_int, [C.POINTER(EnvCfg), C.POINTER(Tables), C.POINTER(C.c_void_p)]),
    "maexp_env_destroy": (None, [C.c_void_p]),
    "maexp_env_channels": (C.c_int, [C.c_void_p]),
    "maexp_env_slot_floats": (C.c_int, [C.c_void_p]),
    "maexp_env_agent_parallel": (C.c_int, [C.c_void_p]),
    "maexp_env_reset": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(ObsOut), C.c_void_p]),
    "maexp_env_step": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ObsOut),
                                 C.c_void_p]),
    "maexp_env_step_host": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "maexp_env_heatmap_ints": (C.c_int64, [C.c_void_p]),
    "maexp_env_step_host_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "maexp_env_host_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "maexp_env_step_host_compact_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "maexp_env_get_state": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_void_p]),
    "maexp_env_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_env_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int32),
                                  C.c_void_p]),
    "maexp_replay_sample": (C.c_int, [C.POINTER(ReplayView), C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.POINTER(Batch), C.c_void_p]),
    "maexp_learner_create": (C.c_int, [C.POINTER(LearnerCfg), C.POINTER(C.c_void_p)]),
    "maexp_learner_destroy": (None, [C.c_void_p]),
    "maexp_learner_num_params": (C.c_int64, [C.c_void_p]),
    "maexp_learner_num_tensors": (C.c_int, [C.c_void_p]),
    "maexp_learner_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]),
    "maexp_learner_buffer": (C.c_void_p, [C.c_void_p, C.c_int]),
    "maexp_tc_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "maexp_tc_gemm_tn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.c_void_p]),
    "maexp_learner_attention_tokens": (C.c_int, [C.c_void_p, C.c_int]),
    "maexp_learner_attention": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "maexp_learner_bind_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_learner_bind_canvas_inputs": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6),
    "maexp_learner_set_los": (C.c_int, [C.c_void_p, C.c_void_p]),
    "maexp_learner_sync_target": (C.c_int, [C.c_void_p, C.c_void_p]),
    "maexp_learner_q_values": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "maexp_learner_grads": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    "maexp_learner_q_values_iqn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_learner_grads_iqn": (C.c_int, [C.c_void_p] + [C.c_void_p] * 9 + [C.c_void_p]),
    "maexp_learner_apply": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_learner_bind_grads": (C.c_int, [C.c_void_p, C.c_void_p]),
    "maexp_learner_apply_data_parallel": (C.c_int, [C.c_void_p, C.POINTER(DpPeers), C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_learner_update": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p]),
    "maexp_heatmap_observation_add": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "maexp_decode_relative": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p]),
    "maexp_select_actions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "maexp_loss_dqn": (C.c_int, [C.c_void_p] * 5 + [C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "maexp_loss_c51": (C.c_int, [C.c_void_p] * 6 + [C.c_float] * 4 + [C.c_int32] * 4 + [C.c_void_p] * 3),
    "maexp_loss_qr": (C.c_int, [C.c_void_p] * 6 + [C.c_float] + [C.c_int32] * 4 + [C.c_void_p] * 3),
    "maexp_loss_iqn": (C.c_int, [C.c_void_p] * 6 + [C.c_float] + [C.c_int32] * 5 + [C.c_void_p] * 3),
    "maexp_clip_rmsprop_step": (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32] + [C.c_float] * 5 + [C.c_void_p] * 3),
    "maexp_fill_uniform": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_void_p]),
    "maexp_last_error": (C.c_char_p, []),
    "maexp_version": (C.c_int, []),
}

_LIB = None


def load():
    """Load the CUDA library; raises MaexpError when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MaexpError(f"{LIB_PATH} is missing: run `python -m maexp_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _LIB = lib
    return _LIB


def check(rc):
    if rc != 0:
        raise MaexpError(load().maexp_last_error().decode() or f"maexp error {rc}")
