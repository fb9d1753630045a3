"""ctypes binding of the C ABI declared in include/dcc_b200.h.

Fails loudly: if libdcc_b200.so is missing or a call returns a non-zero status a DccError is raised —
there is no CPU or PyTorch fallback for any entry point.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdcc_b200.so")


class DccError(RuntimeError):
    pass


class EnvCfg(C.Structure):
    """struct dcc_env_cfg (include/dcc_b200.h)."""
    _fields_ = [
        ("n_envs", C.c_int32), ("n_agents", C.c_int32), ("n_pois", C.c_int32), ("max_ep_len", C.c_int32),
        ("reference_compat", C.c_int32), ("reserved0", C.c_int32),
        ("r_cover", C.c_double), ("r_comm", C.c_double), ("comm_r_scale", C.c_double),
        ("comm_force_scale", C.c_double), ("dt", C.c_double), ("damping", C.c_double), ("max_speed", C.c_double),
        ("sensitivity", C.c_double), ("m_energy", C.c_double), ("rew_cover", C.c_double), ("rew_done", C.c_double),
        ("rew_out", C.c_double), ("contact_margin", C.c_double),
    ]


class MappoCfg(C.Structure):
    """struct dcc_mappo_cfg (include/dcc_b200.h)."""
    _fields_ = [
        ("n_agents", C.c_int32), ("obs_dim", C.c_int32), ("hidden", C.c_int32), ("act_dim", C.c_int32),
        ("chunk_rows", C.c_int32), ("gemm_backend", C.c_int32),
        ("clip_param", C.c_float), ("entropy_coef", C.c_float), ("value_loss_coef", C.c_float),
        ("huber_delta", C.c_float), ("max_grad_norm", C.c_float), ("gamma", C.c_float), ("gae_lambda", C.c_float),
        ("opti_eps", C.c_float), ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("vn_beta", C.c_double),
        ("use_huber_loss", C.c_int32), ("use_clipped_value_loss", C.c_int32), ("use_max_grad_norm", C.c_int32),
        ("use_valuenorm", C.c_int32), ("use_gae", C.c_int32), ("use_feature_normalization", C.c_int32),
        ("weight_decay", C.c_float), ("use_relu", C.c_int32), ("layer_N", C.c_int32), ("recurrent_N", C.c_int32),
    ]


_VP = C.c_void_p
# name -> (restype, argtypes); every symbol include/dcc_b200.h declares
SIGNATURES = {
    "dcc_env_cfg_default": (C.c_int, [C.POINTER(EnvCfg)]),
    "dcc_env_obs_dim": (C.c_int, [C.c_int32, C.c_int32]),
    "dcc_env_create": (C.c_int, [C.POINTER(EnvCfg), _VP, C.c_int, C.POINTER(_VP)]),
    "dcc_env_destroy": (C.c_int, [_VP]),
    "dcc_env_set_poi_layouts": (C.c_int, [_VP, _VP, _VP]),
    "dcc_env_reset": (C.c_int, [_VP, _VP, _VP]),
    "dcc_env_step": (C.c_int, [_VP] * 10),
    "dcc_env_step_host": (C.c_int, [_VP] * 7),
    "dcc_env_reset_host": (C.c_int, [_VP, _VP, _VP]),
    "dcc_env_get_state": (C.c_int, [_VP, _VP, _VP, _VP]),
    "dcc_env_set_state": (C.c_int, [_VP, _VP, _VP, _VP]),
    "dcc_env_snapshot_state": (C.c_int, [_VP, _VP, _VP, _VP]),
    "dcc_env_state_ptrs": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    "dcc_env_set_launch": (C.c_int, [_VP, C.c_int, C.c_int]),
    "dcc_env_use_specialized": (C.c_int, [_VP, C.c_int]),
    "dcc_env_launch_count": (C.c_int64, [_VP]),
    "dcc_mappo_cfg_default": (C.c_int, [C.POINTER(MappoCfg)]),
    "dcc_mappo_create": (C.c_int, [C.POINTER(MappoCfg), C.c_int, C.POINTER(_VP)]),
    "dcc_mappo_destroy": (C.c_int, [_VP]),
    "dcc_mappo_param_count": (C.c_int64, [_VP, C.c_int]),
    "dcc_mappo_chunk_rows": (C.c_int, [_VP]),
    "dcc_mappo_gemm_backend": (C.c_int, [_VP]),
    "dcc_mappo_launch_count": (C.c_int64, [_VP]),
    "dcc_mappo_act": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_uint64, C.c_uint64, C.c_int, _VP, _VP, _VP, _VP]),
    "dcc_mappo_evaluate": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int, _VP, _VP, _VP, _VP]),
    "dcc_mappo_set_env_layout": (C.c_int, [_VP, C.c_int, _VP, C.c_double]),
    "dcc_mappo_act_state": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int, C.c_uint64, C.c_uint64, C.c_int, _VP, _VP, _VP, _VP]),
    "dcc_mappo_evaluate_state": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int, _VP, _VP, _VP, _VP]),
    "dcc_mappo_epoch_grads_state": (C.c_int, [_VP] * 13 + [C.c_double, C.c_int, C.c_int, _VP, _VP]),
    "dcc_mappo_minibatch_grads_state": (C.c_int, [_VP] * 13 + [C.c_double, _VP, C.c_int64, _VP, C.c_double, _VP, _VP]),
    "dcc_obs_from_state": (C.c_int, [_VP, _VP, _VP, C.c_int, _VP, _VP]),
    "dcc_rollout_insert": (C.c_int, [_VP, _VP, C.c_int, C.c_int, _VP, _VP, _VP]),
    "dcc_mappo_gae": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int, C.c_int, _VP, _VP]),
    "dcc_mappo_train_begin": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, _VP, _VP]),
    "dcc_mappo_epoch_grads": (C.c_int, [_VP] * 12 + [C.c_double, C.c_int, C.c_int, _VP, _VP]),
    "dcc_mappo_minibatch_stats": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _VP]),
    "dcc_mappo_minibatch_grads": (C.c_int, [_VP] * 12 + [C.c_double, _VP, C.c_int64, _VP, C.c_double, _VP, _VP]),
    "dcc_mappo_rnn_pass_seqs": (C.c_int, [_VP, C.c_int]),
    "dcc_mappo_act_rnn": (C.c_int, [_VP] * 4 + [C.c_int, _VP, _VP, _VP, C.c_int, C.c_uint64, C.c_uint64, C.c_int] + [_VP] * 6),
    "dcc_mappo_seq_grads": (C.c_int, [_VP] * 15 + [C.c_double, _VP, C.c_int64, C.c_int, _VP, C.c_double, _VP, _VP]),
    "dcc_mappo_apply": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP, C.c_float, C.c_int64, _VP, _VP]),
    "dcc_op_gemm": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _VP, C.c_int, _VP, C.c_int,
                              _VP, C.c_int, C.c_int, _VP]),
    "dcc_host_alloc": (C.c_int, [C.POINTER(_VP), C.c_size_t]),
    "dcc_host_free": (C.c_int, [_VP]),
    "dcc_status_string": (C.c_char_p, [C.c_int]),
    "dcc_last_cuda_error": (C.c_char_p, []),
    "dcc_abi_version": (C.c_int, []),
}

_lib = None


def load():
    """dlopen libdcc_b200.so and bind every declared symbol.  Raises DccError if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build when the library is missing or older than its sources (content hash, build.py); a fresh clone works
    # without a separate build step as long as nvcc is there.  DCC_NO_AUTOBUILD=1 skips the check.
    if not os.environ.get("DCC_NO_AUTOBUILD"):
        from . import build as _build
        try:
            if _build._stale():
                _build.build()
        except RuntimeError as e:
            if not os.path.exists(LIB_PATH):
                raise DccError("%s is missing and could not be built (%s). There is no CPU fallback." % (LIB_PATH, e))
            import warnings
            warnings.warn("libdcc_b200.so is stale and could not be rebuilt (%s); using the existing library" % e)
    if not os.path.exists(LIB_PATH):
        raise DccError("%s is missing: build it with `python -m dcc_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        lib = load()
        msg = lib.dcc_status_string(int(status)).decode()
        if status == -2:
            msg += ": " + lib.dcc_last_cuda_error().decode()
        raise DccError("%s failed: %s (status %d)" % (what or "dcc call", msg, status))
