"""ctypes binding of include/cvaegan_b200.h (the C-ABI drop-in boundary).

There is NO fallback: if `libcvaegan_b200.so` is missing or a call fails, this module raises.
Build the library in-tree with `python -c "import __graft_entry__ as g; g.build()"`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcvaegan_b200.so")

NET_ENCODER, NET_GENERATOR, NET_DISCRIMINATOR, NET_CLASSIFIER = 0, 1, 2, 3
NET_NAMES = ("encoder", "generator", "discriminator", "classifier")
STEP_NO_UPDATE, STEP_LOCAL_BN, VISIT_LAMBDA_ZERO, STEP_PRIOR_ONLY, STEP_CVAE = 1, 2, 4, 8, 16
GRAD_TAIL = 16
ABI_VERSION = 2      # CVG_ABI_VERSION of include/cvaegan_b200.h this binding was written against


class CvgError(RuntimeError):
    pass


class CvgConfig(C.Structure):
    _fields_ = [
        ("feature_num", C.c_int32), ("label_num", C.c_int32), ("z_size", C.c_int32), ("max_batch", C.c_int32),
        ("world_size", C.c_int32), ("rank", C.c_int32),
        ("lambda_recon", C.c_float), ("lambda_kl", C.c_float), ("lambda_adv", C.c_float),
        ("g_lr", C.c_float), ("d_lr", C.c_float), ("c_lr", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("bn_momentum", C.c_float), ("bn_eps", C.c_float), ("ln_eps", C.c_float), ("sn_eps", C.c_float),
        ("lrelu_slope", C.c_float), ("dropout_p", C.c_float),
        ("hidden", C.c_int32 * 3),     # (0, 0, 0): the reference's widths; else the widened model's three hidden widths
        ("unconditional", C.c_int32),  # 1: E / G / D without the one-hot label columns (sibling trainer VAE-GAN)
    ]


class CvgTensorDesc(C.Structure):
    _fields_ = [("key", C.c_char * 96), ("kind", C.c_int32), ("ndim", C.c_int32), ("shape", C.c_int64 * 2),
                ("offset", C.c_int64)]


class CvgNoise(C.Structure):
    _fields_ = [("z", C.c_void_p), ("eps", C.c_void_p), ("d_mask1", C.c_void_p), ("d_mask2", C.c_void_p),
                ("c_mask1", C.c_void_p), ("c_mask2", C.c_void_p)]


_P, _I, _I64, _U64, _F = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); every symbol include/cvaegan_b200.h declares
SIGNATURES = {
    "cvg_last_error": (C.c_char_p, []),
    "cvg_abi_version": (_I, []),
    "cvg_config_bytes": (_I, []),
    "cvg_create": (_I, [C.POINTER(CvgConfig), C.POINTER(_P)]),
    "cvg_destroy": (None, [_P]),
    "cvg_net_sizes": (_I, [_P, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    "cvg_tensor_table": (_I, [_P, _I, C.POINTER(CvgTensorDesc), C.c_int32, C.POINTER(C.c_int32)]),
    "cvg_workspace_bytes": (_I64, [_P]),
    "cvg_bind_net": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "cvg_bind_workspace": (_I, [_P, _P, _I64, _P]),
    "cvg_set_adam_step": (_I, [_P, _I, _I64]),
    "cvg_get_adam_step": (_I64, [_P, _I]),
    "cvg_comm_unique_id": (_I, [_P]),
    "cvg_comm_init": (_I, [_P, _P, _I, _I]),
    "cvg_nvl_local_handle": (_I, [_P, _P]),
    "cvg_nvl_attach": (_I, [_P, _P]),
    "cvg_nvl_disable": (_I, [_P]),
    "cvg_step_d": (_I, [_P, _P, _I, _I, C.POINTER(CvgNoise), _U64, _U64, _I, _P, _P]),
    "cvg_step_c": (_I, [_P, _P, _I, _I, C.POINTER(CvgNoise), _U64, _U64, _I, _P, _P]),
    "cvg_step_g": (_I, [_P, _P, _I, _I, C.POINTER(CvgNoise), _U64, _U64, _F, _I, _P, _P]),
    "cvg_step_classifier": (_I, [_P, _P, _P, _I, C.POINTER(CvgNoise), _U64, _U64, _F, _F, _F, _F, _I, _P, _P]),
    "cvg_visit": (_I, [_P, _I, _I, _I64, _P, _I64, _P, _I, _I, _I, _I, _P, _P]),
    "cvg_ctl_set": (_I, [_P, _U64, _U64, _I, _F, _I, _P]),
    "cvg_adam": (_I, [_P, _I, _P]),
    "cvg_sample_rows": (_I, [_P, _P, _I64, _I64, _I64, _I, _U64, _U64, _P, _P, _P]),
    "cvg_generate": (_I, [_P, _I, _I64, _P, _U64, _U64, _I, _P, _P]),
    "cvg_generate_filter": (_I, [_P, _I, _I64, _F, _P, _U64, _U64, _P, _P, _I64, _P, _P, _P, _P]),
    "cvg_filter_logits": (_I, [_P, _I64, _I, _I, _F, _P, _P]),
    "cvg_filter_compact": (_I, [_P, _P, _I64, _I, _I, _I, _F, _U64, _P, _P, _I64, _P, _P]),
    "cvg_classifier_forward": (_I, [_P, _P, _I64, _P, _P]),
    "cvg_encoder_forward": (_I, [_P, _P, _I, _I64, _P, _P, _P]),
    "cvg_patience_scan": (_I, [_P, _I64, _I64, _I, _I, C.POINTER(_I64), C.POINTER(_I64)]),
    "cvg_debug_read": (_I, [_P, C.c_char_p, _I, _I, _P, C.POINTER(C.c_int), _P]),
    "cvg_debug_tc_counters": (_I, [_P, _P]),
    "cvg_profile_enable": (_I, [_P, _I]),
    "cvg_profile_read": (_I, [_P, _I, C.POINTER(_I64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cvg_launch_count": (_I64, [_P]),
    "cvg_debug_set": (_I, [_P, C.c_char_p, _I]),
    "cvg_debug_get": (_I, [_P, C.c_char_p, C.POINTER(_I)]),
    "cvg_debug_mk_cycles": (_I, [_P, _P, _I, C.POINTER(_I)]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvgError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no fallback path. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.cvg_abi_version() != ABI_VERSION:
        raise CvgError("libcvaegan_b200.so ABI version mismatch")
    if lib.cvg_config_bytes() != C.sizeof(CvgConfig):
        raise CvgError(f"CvgConfig layout mismatch: library {lib.cvg_config_bytes()} bytes, binding {C.sizeof(CvgConfig)}")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().cvg_last_error()
        raise CvgError(msg.decode("utf-8", "replace") if msg else f"cvaegan_b200 call failed (rc={rc})")
