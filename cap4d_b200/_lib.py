"""ctypes binding of libcap4d_b200.so (the C ABI declared in include/cap4d_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry point fails
loudly.  The library is built in-tree by `python -m cap4d_b200.build` / `__graft_entry__.build()`.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcap4d_b200.so")

MAX_LEVELS = 8
N_CLASSES = 6
CLASS_NAMES = ("conv3x3", "linear", "attention", "groupnorm", "layernorm", "other")


class UnetConfig(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int),
        ("out_channels", c_int),
        ("model_channels", c_int),
        ("condition_channels", c_int),
        ("num_res_blocks", c_int),
        ("n_levels", c_int),
        ("channel_mult", c_int * MAX_LEVELS),
        ("n_attention_resolutions", c_int),
        ("attention_resolutions", c_int * MAX_LEVELS),
        ("num_head_channels", c_int),
        ("time_steps", c_int),
    ]


MAX_GROUPS_PER_CALL = 16


class SamplerStores(ctypes.Structure):
    _fields_ = [(name, c_void_p) for name in (
        "ref_z", "ref_mask", "ref_pos", "gen_z", "gen_mask", "gen_pos",
        "ref_z_u", "ref_mask_u", "ref_pos_u", "gen_z_u", "gen_mask_u", "gen_pos_u", "latents")]


class SamplerCall(ctypes.Structure):
    _fields_ = [("timestep", c_int64), ("x_coef", c_float), ("e_coef", c_float), ("n_groups", ctypes.c_int32),
                ("pad_", ctypes.c_int32), ("groups", ctypes.c_int32 * MAX_GROUPS_PER_CALL)]


class VaeConfig(ctypes.Structure):
    _fields_ = [
        ("ch", c_int),
        ("n_levels", c_int),
        ("ch_mult", c_int * MAX_LEVELS),
        ("num_res_blocks", c_int),
        ("z_channels", c_int),
        ("embed_dim", c_int),
        ("out_ch", c_int),
    ]


# name -> (restype, argtypes); must list every symbol include/cap4d_b200.h declares
SIGNATURES = {
    "cap4d_b200_unet_create": (c_int, [POINTER(UnetConfig), POINTER(c_void_p)]),
    "cap4d_b200_unet_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "cap4d_b200_unet_num_params": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_unet_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_int64), POINTER(c_int)]),
    "cap4d_b200_unet_finalize": (c_int, [c_void_p]),
    "cap4d_b200_unet_set_precision": (c_int, [c_void_p, c_int]),
    "cap4d_b200_unet_set_ref_views": (c_int, [c_void_p, c_int]),
    "cap4d_b200_unet_ref_view_violations": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_vae_create": (c_int, [POINTER(VaeConfig), POINTER(c_void_p)]),
    "cap4d_b200_vae_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "cap4d_b200_vae_num_params": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_vae_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_int64), POINTER(c_int)]),
    "cap4d_b200_vae_finalize": (c_int, [c_void_p]),
    "cap4d_b200_vae_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_size_t)]),
    "cap4d_b200_vae_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_float, c_void_p,
                                      c_size_t, c_void_p]),
    "cap4d_b200_vae_decode_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_float, c_void_p,
                                         c_size_t, c_void_p]),
    "cap4d_b200_vae_num_launches": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_vae_has_encoder": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_vae_encode_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_size_t)]),
    "cap4d_b200_vae_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cap4d_b200_vae_destroy": (c_int, [c_void_p]),
    "cap4d_b200_unet_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "cap4d_b200_unet_forward": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
         c_size_t, c_void_p],
    ),
    "cap4d_b200_unet_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "cap4d_b200_unet_num_launches": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_unet_class_stats": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_int)]),
    "cap4d_b200_unet_class_exec_flops": (c_int, [c_void_p, POINTER(c_double)]),
    "cap4d_b200_unet_forward_timed": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
         c_size_t, c_void_p, POINTER(c_float)],
    ),
    "cap4d_b200_unet_enable_taps": (c_int, [c_void_p, c_int]),
    "cap4d_b200_unet_num_taps": (c_int, [c_void_p, POINTER(c_int)]),
    "cap4d_b200_unet_tap_info": (c_int, [c_void_p, c_int, c_char_p, c_int, POINTER(c_void_p), POINTER(c_int64),
                                         POINTER(c_int), POINTER(c_int)]),
    "cap4d_b200_unet_collect_timings": (c_int, [c_void_p, POINTER(c_float), POINTER(c_int)]),
    "cap4d_b200_unet_destroy": (c_int, [c_void_p]),
    "cap4d_b200_cfg_ddim_update": (
        c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    "cap4d_b200_sampler_gather": (
        c_int, [POINTER(SamplerStores), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cap4d_b200_sampler_update": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "cap4d_b200_sampler_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cap4d_b200_sampler_unpack": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cap4d_b200_gemm_bf16": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                POINTER(c_float), c_int]),
    "cap4d_b200_gemm_mixed": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                POINTER(c_float), c_int]),
    "cap4d_b200_conv3x3_bf16": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                c_void_p, POINTER(c_float), c_int]),
    "cap4d_b200_upsample_conv3x3_bf16": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_float),
                c_int]),
    "cap4d_b200_attention_bf16": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, POINTER(c_float), c_int]),
    "cap4d_b200_attention_trace": (
        c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "cap4d_b200_groupnorm_bf16": (
        c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p,
                c_void_p, c_void_p, POINTER(c_float), c_int]),
    "cap4d_b200_layernorm_bf16": (
        c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, POINTER(c_float), c_int]),
    "cap4d_b200_cond_pos_enc": (
        c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "cap4d_b200_cond_workspace_bytes": (c_int, [c_int, c_int, POINTER(c_size_t)]),
    "cap4d_b200_cond_ray_map": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cap4d_b200_last_error": (c_char_p, []),
    "cap4d_b200_version": (c_char_p, []),
}

_lib = None


def load():
    """Load the shared library (once) and bind every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"cap4d_b200: {LIB_PATH} is missing. The CUDA extension is the product; there is no fallback. "
            "Build it with `python -m cap4d_b200.build` (needs nvcc with sm_100a support)."
        )
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().cap4d_b200_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"cap4d_b200: {what} failed (status {status}): {last_error()}")
