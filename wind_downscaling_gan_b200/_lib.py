"""ctypes binding of libwdg.so (C ABI in include/wdg.h).

There is no CPU fallback: importing this module without the built library, or
calling into it without an sm_100 device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwdg.so")

_lib = None


class WdgError(RuntimeError):
    pass


def _declare(lib):
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    fp = C.POINTER(C.c_float)
    i64p = C.POINTER(C.c_int64)
    lib.wdg_last_error.restype = C.c_char_p
    lib.wdg_last_error.argtypes = []
    sigs = {
        "wdg_device_info": [i, C.POINTER(i), C.POINTER(i), C.POINTER(i)],
        "wdg_generator_create": [C.POINTER(vp), i, i, i, i, i, i],
        "wdg_generator_num_weights": [vp],
        "wdg_generator_weight_info": [vp, i, C.POINTER(C.c_char_p), i64p, C.POINTER(i)],
        "wdg_generator_set_weight": [vp, C.c_char_p, vp, i64p, i],
        "wdg_generator_get_weight": [vp, C.c_char_p, vp, C.c_int64],
        "wdg_generator_finalize": [vp],
        "wdg_generator_set_precision": [vp, i],
        "wdg_generator_get_precision": [vp],
        "wdg_generator_workspace_bytes": [vp, i, i, C.POINTER(sz)],
        "wdg_generator_bind": [vp, i, i, vp, sz, vp],
        "wdg_generator_forward": [vp, vp, vp, vp, vp],
        "wdg_generator_forward_gen_noise": [vp, vp, C.c_float, C.c_uint64, C.c_uint64, vp, vp, vp],
        "wdg_generator_io_bytes": [vp, i, i, C.POINTER(sz)],
        "wdg_generator_predict_host": [vp, vp, vp, vp, vp, vp],
        "wdg_generator_predict_host_gen_noise": [vp, vp, C.c_float, C.c_uint64, C.c_uint64, vp, vp, vp],
        "wdg_generator_launches_per_forward": [vp],
        "wdg_generator_pipeline_schedule": [i, i, C.POINTER(i), i],
        "wdg_generator_debug_read": [vp, i, vp, C.c_int64],
        "wdg_generator_profile": [vp, i],
        "wdg_generator_stage_ms": [vp, vp, i],
        "wdg_patch_scratch_bytes": [i, i, i, i, C.POINTER(sz)],
        "wdg_gather_normalise": [vp, vp, vp, i, i, i, vp, i, vp, i, i, i, vp, vp, vp, vp, vp],
        "wdg_gather_normalise_regrid": [vp, vp, i, i, i, vp, vp, vp, i, i, vp, vp, C.c_float, i, i, vp, i, vp, i, i, i, vp, vp, vp, vp, vp],
        "wdg_stitch": [vp, vp, i, vp, i, i, i, i, i, i, vp, i, vp, i, vp, vp],
        "wdg_stitch_accum": [vp, vp, i, vp, i, i, i, i, i, i, vp, i, vp, i, vp, i, vp],
    }
    ll, f, d = C.c_longlong, C.c_float, C.c_double
    ip = C.POINTER(i)
    sigs.update({
        "wdg_train_set_precision": [i],
        "wdg_train_get_precision": [],
        "wdg_conv2d_fwd": [vp, vp, vp, vp, ip, i, vp],
        "wdg_conv2d_fwd_act": [vp, vp, vp, vp, ip, i, f, vp],
        "wdg_conv2d_bwd_data": [vp, vp, vp, ip, i, vp],
        "wdg_conv2d_bwd_weight_scratch": [ip, C.POINTER(sz), ip],
        "wdg_conv2d_bwd_weight": [vp, vp, vp, ip, vp, i, vp],
        "wdg_colsum": [i, vp, i, i, vp, i, i, ll, i, vp, vp, i, vp],
        "wdg_leaky_relu_fwd": [vp, ll, f, vp],
        "wdg_leaky_relu_bwd": [vp, vp, ll, f, vp],
        "wdg_axpby": [vp, i, i, vp, i, i, f, vp, i, i, f, ll, i, i, vp],
        "wdg_lerp_batch": [vp, vp, vp, vp, ll, ll, vp],
        "wdg_bias_act": [vp, i, i, vp, ll, i, f, vp],
        "wdg_transpose01": [vp, vp, i, i, ll, vp],
        "wdg_col2im": [vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "wdg_bn_train_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, ll, i, f, f, vp, vp],
        "wdg_bn_infer": [vp, vp, vp, vp, vp, vp, ll, i, f, vp, vp],
        "wdg_bn_finalize_apply": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ll, ll, i, f, f, vp],
        "wdg_bn_bwd_sums": [vp, vp, vp, vp, vp, vp, ll, i, vp, vp],
        "wdg_bn_bwd_dx": [vp, vp, vp, vp, vp, vp, vp, vp, ll, ll, i, f, vp],
        "wdg_bn_train_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, ll, i, f, vp, vp],
        "wdg_ln_fwd": [vp, vp, i, i, vp, vp, vp, vp, ll, i, f, vp],
        "wdg_ln_bwd": [vp, i, i, vp, vp, vp, vp, vp, vp, vp, ll, i, f, vp, vp],
        "wdg_lstm_gates_fwd": [vp, vp, vp, vp, ll, i, vp],
        "wdg_lstm_gates_bwd": [vp, vp, vp, vp, vp, vp, ll, i, vp],
        "wdg_lstm_small_fwd": [vp, vp, vp, vp, vp, vp, i, i, i, i, vp],
        "wdg_lstm_small_bwd_data": [vp, vp, vp, i, i, i, i, vp],
        "wdg_lstm16_pack": [vp, vp, vp],
        "wdg_round_tf32": [vp, ll, vp],
        "wdg_lstm16_fwd_step": [vp, vp, vp, vp, vp, vp, i, i, i, vp],
        "wdg_lstm16_bwd_step": [vp, vp, vp, vp, vp, vp, vp, i, i, i, vp],
        "wdg_lstm128_pack": [vp, vp, vp],
        "wdg_lstm128_fwd_step": [vp, vp, vp, vp, vp, vp, i, i, i, vp],
        "wdg_upconv5x5_workspace_bytes": [ll, i, C.POINTER(sz)],
        "wdg_upconv5x5_fwd": [vp, vp, vp, vp, vp, ll, i, vp, sz, vp],
        "wdg_upsample2x_fwd": [vp, vp, ll, i, i, i, vp],
        "wdg_upsample2x_bwd": [vp, vp, ll, i, i, i, vp],
        "wdg_dense_mean_fwd": [vp, vp, vp, vp, i, i, i, vp],
        "wdg_dense_mean_bwd": [vp, vp, vp, vp, vp, vp, i, i, i, vp],
        "wdg_reduce": [i, vp, vp, ll, d, vp, vp, vp],
        "wdg_gp_norm": [vp, vp, i, ll, i, vp],
        "wdg_adam": [vp, vp, vp, vp, ll, f, f, f, f, vp],
        "wdg_sn_update": [vp, vp, i, i, vp, vp],
        "wdg_adam_lr": [vp, vp, f, f, f, vp],
        "wdg_adam_dev": [vp, vp, vp, vp, ll, vp, f, f, f, vp],
        "wdg_noise_normal_state": [vp, ll, f, vp, vp],
        "wdg_rng_advance": [vp, C.c_uint64, vp],
        "wdg_uniform_state": [vp, ll, vp, vp],
        "wdg_noise_normal": [vp, ll, f, C.c_uint64, C.c_uint64, vp],
        "wdg_critic_create": [C.POINTER(vp), i, i, i, i, i, i, i],
        "wdg_critic_num_weights": [vp],
        "wdg_critic_weight_info": [vp, i, C.POINTER(C.c_char_p), i64p, ip, i64p, ip],
        "wdg_critic_context_bytes": [vp, i, i, i, C.POINTER(sz)],
        "wdg_critic_scratch_bytes": [vp, i, i, C.POINTER(sz)],
        "wdg_critic_sn_update": [vp, vp, vp, sz, vp],
        "wdg_critic_forward": [vp, vp, vp, vp, vp, i, i, i, vp, sz, vp, sz, vp],
        "wdg_critic_backward": [vp, vp, vp, i, i, i, vp, vp, vp, vp, sz, vp],
        "wdg_critic_backward_input": [vp, vp, vp, i, i, i, vp, vp, vp, sz, vp],
        "wdg_critic_backward_weights": [vp, vp, vp, i, i, i, vp, vp, vp, sz, vp],
        "wdg_metrics_pointwise_scratch": [i, C.POINTER(sz)],
        "wdg_metrics_pointwise": [vp, vp, i, ll, i, vp, vp, vp],
        "wdg_metric_lsd": [vp, vp, i, i, i, i, i, vp, vp, vp],
        "wdg_metric_spatial_ks": [vp, vp, i, i, i, i, i, i, vp, vp],
    })
    for name, args in sigs.items():
        if not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i
    lib.wdg_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.wdg_philox4x32_10.restype = None
    lib.wdg_generator_destroy.argtypes = [vp]
    lib.wdg_generator_destroy.restype = None
    lib.wdg_crc32c.argtypes = [C.c_uint32, vp, sz]
    lib.wdg_crc32c.restype = C.c_uint32
    lib.wdg_critic_destroy.argtypes = [vp]
    lib.wdg_critic_destroy.restype = None
    for name in ("wdg_critic_num_floats", "wdg_critic_num_trainable_floats"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = C.c_int64
    del fp


def lib():
    """Load libwdg.so (once).  Raises WdgError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WdgError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "This package has no CPU / PyTorch fallback.")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


calls = 0   # library entry points called so far (every one launches at least one kernel on the training path)


def check(rc):
    global calls
    calls += 1
    if rc != 0:
        raise WdgError(lib().wdg_last_error().decode())
