"""ctypes binding of libladiff_b200.so (the C-ABI declared in include/ladiff_b200.h).

There is no CPU fallback: if the shared library is missing the first call raises, telling the
user to build it (``python -m ladiffcodec_b200.build`` / ``__graft_entry__.build()``).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libladiff_b200.so")
ABI_VERSION = 2
MAX_RATIOS = 8

ERR_NAMES = {-1: "LADIFF_ERR_ARG", -2: "LADIFF_ERR_STATE", -3: "LADIFF_ERR_KEY", -4: "LADIFF_ERR_CUDA",
             -5: "LADIFF_ERR_UNSUPPORTED", -6: "LADIFF_ERR_WORKSPACE"}

c_i32, c_i64, c_u64, c_vp, c_cp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_char_p


class LadiffConfig(ctypes.Structure):
    _fields_ = [
        ("rep_dims", c_i32), ("diff_dims", c_i32), ("n_filters", c_i32), ("lstm_layers", c_i32),
        ("n_enc_ratios", c_i32), ("enc_ratios", c_i32 * MAX_RATIOS),
        ("quantization", c_i32), ("n_q", c_i32), ("n_q_used", c_i32), ("run_diff", c_i32),
        ("cond_channels", c_i32), ("n_upsampling_ratios", c_i32), ("upsampling_ratios", c_i32 * MAX_RATIOS),
        ("unet_scale_cond", c_i32), ("sample_rate", c_i32), ("reserved", c_i32 * 8),
    ]


# name -> (restype, argtypes); every symbol include/ladiff_b200.h declares
SIGNATURES = {
    "ladiff_last_error": (c_cp, []),
    "ladiff_abi_version": (c_i32, []),
    "ladiff_act_dtype": (c_cp, []),
    "ladiff_create": (c_i32, [ctypes.POINTER(LadiffConfig), ctypes.POINTER(c_vp)]),
    "ladiff_destroy": (c_i32, [c_vp]),
    "ladiff_load_weight": (c_i32, [c_vp, c_cp, c_vp, ctypes.POINTER(c_i64), c_i32]),
    "ladiff_finalize": (c_i32, [c_vp]),
    "ladiff_expected_keys": (c_i32, [c_vp]),
    "ladiff_expected_key_at": (c_i32, [c_vp, c_i32, ctypes.POINTER(c_cp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i32)]),
    "ladiff_workspace_bytes": (c_i64, [c_vp, c_i32, c_i32]),
    "ladiff_synthesize_workspace_bytes": (c_i64, [c_vp, c_vp, c_i32, c_i32]),
    "ladiff_get_cond": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_encode": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_rvq_decode": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp]),
    "ladiff_rvq_encode": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ladiff_upsample_layer": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "ladiff_unet_forward": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_ddpm_steps": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_u64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "ladiff_set_clip_offset": (c_i32, [c_vp, c_u64]),
    "ladiff_ddim_steps": (c_i32, [c_vp, c_vp, c_vp, ctypes.POINTER(c_i32), c_i32, ctypes.c_double, c_vp, c_i64, c_u64, c_i32, c_i32, c_i32,
                                  c_vp, c_i64, c_vp]),
    "ladiff_ddim_times": (c_i32, [c_i32, c_i32, ctypes.POINTER(c_i32)]),
    "ladiff_randn": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_u64, c_i32, c_vp]),
    "ladiff_q_sample": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp, c_i64, c_vp]),
    "ladiff_axpby": (c_i32, [c_vp, ctypes.c_double, c_vp, ctypes.c_double, c_i64, c_vp]),
    "ladiff_p_losses": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_sdsdr": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i64, ctypes.c_double, c_vp]),
    "ladiff_synthesize_ddim": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, ctypes.c_double, c_vp, c_vp, c_i64, c_u64, c_vp, c_vp, c_vp,
                                       c_i64, c_vp]),
    "ladiff_decode": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_synthesize": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_i64, c_u64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_synthesize_codes": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_u64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "ladiff_normalize_clips": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp]),
    "ladiff_op_conv1d_cl": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp]),
    "ladiff_op_conv1d_bench": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                       ctypes.POINTER(ctypes.c_float), ctypes.c_char_p, c_i32]),
    "ladiff_op_fullattn": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32]),
    "ladiff_set_conv_impl": (c_i32, [c_vp, c_i32]),
    "ladiff_take_launch_count": (c_i64, [c_vp]),
    "ladiff_set_skip_ops": (c_i32, [c_vp, c_i32]),
    "ladiff_set_profiling": (c_i32, [c_vp, c_i32]),
    "ladiff_profile_report": (c_i32, [c_vp, ctypes.POINTER(ctypes.c_double)]),
    "ladiff_profile_dump": (c_i32, [c_vp, ctypes.c_char_p, c_i64]),
}

_lib = None


class LadiffError(RuntimeError):
    pass


def get_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LadiffError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
                "`python -m ladiffcodec_b200.build` (needs nvcc). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.ladiff_abi_version() != ABI_VERSION:
            raise LadiffError(f"ABI mismatch: library {lib.ladiff_abi_version()} != binding {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def act_dtype():
    """torch dtype of the UNet's 16-bit operands / stored activations in the loaded build."""
    import torch
    return {"f16": torch.float16, "bf16": torch.bfloat16}[get_lib().ladiff_act_dtype().decode()]


def check(rc, what=""):
    if rc != 0:
        msg = get_lib().ladiff_last_error().decode("utf-8", "replace")
        raise LadiffError(f"{what or 'ladiff call'} failed with {ERR_NAMES.get(rc, rc)}: {msg}")


def make_config(*, rep_dims, diff_dims, n_filters, lstm, enc_ratios, quantization, n_q, n_q_used, run_diff,
                cond_channels, upsampling_ratios, unet_scale_cond, sample_rate):
    c = LadiffConfig()
    c.rep_dims, c.diff_dims, c.n_filters, c.lstm_layers = rep_dims, diff_dims, n_filters, lstm
    if len(enc_ratios) > MAX_RATIOS or len(upsampling_ratios or []) > MAX_RATIOS:
        raise ValueError("too many ratios")
    c.n_enc_ratios = len(enc_ratios)
    for i, r in enumerate(enc_ratios):
        c.enc_ratios[i] = int(r)
    c.quantization, c.n_q, c.n_q_used, c.run_diff = int(quantization), int(n_q), int(n_q_used), int(run_diff)
    c.cond_channels = cond_channels
    ups = list(upsampling_ratios) if (upsampling_ratios is not None and run_diff) else []
    c.n_upsampling_ratios = len(ups)
    for i, r in enumerate(ups):
        c.upsampling_ratios[i] = int(r)
    c.unet_scale_cond, c.sample_rate = int(bool(unet_scale_cond)), int(sample_rate)
    return c
