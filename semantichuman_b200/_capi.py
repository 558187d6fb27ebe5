"""ctypes binding of libshb200.so (the C ABI declared in include/shb200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the import of the
product fails loudly (north_star: "no CPU fallback").  Build it with ``python semantichuman_b200/_build.py``
or ``__graft_entry__.build()``.
"""
import ctypes
import os

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libshb200.so")

c_int, c_i64, c_size, c_vp, c_float = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float

# name -> (restype, argtypes); one entry per declaration in include/shb200.h
SIGNATURES = {
    "shb_abi_version": (c_int, []),
    "shb_set_persistent_sms": (c_int, [c_int]),
    "shb_error_string": (ctypes.c_char_p, [c_int]),
    "shb_build_inverse_spiral_csr": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    "shb_dense_to_csr": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "shb_csr_transpose": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "shb_l1_loss_workspace": (c_size, [c_i64]),
    "shb_l1_loss_fwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_size, c_vp, c_int, c_vp]),
    "shb_l1_loss_bwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_vp]),
    "shb_colsum": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    "shb_partnorm_loss_fwd_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    "shb_pair_loss_workspace": (c_size, [c_int] * 3),
    "shb_pair_loss_grad_acc_bytes": (c_size, [c_int] * 3),
    "shb_pair_loss_fwd": (c_int, [c_vp] * 9 + [c_float, c_int, c_vp, c_vp, c_vp, c_size] + [c_int] * 5 + [c_vp]),
    "shb_pair_loss_bwd": (c_int, [c_vp] * 6 + [c_size] + [c_int] * 4 + [c_vp]),
    "shb_group_linear_gather_fwd": (c_int, [c_vp] * 8 + [c_int] * 5 + [c_vp]),
    "shb_group_linear_gather_bwd": (c_int, [c_vp] * 10 + [c_int] * 6 + [c_vp]),
    "shb_group_linear_scatter_fwd": (c_int, [c_vp] * 8 + [c_int] * 5 + [c_vp]),
    "shb_group_linear_scatter_bwd": (c_int, [c_vp] * 10 + [c_int] * 6 + [c_vp]),
    "shb_adam_tick": (c_int, [c_vp, c_vp]),
    "shb_adam_step": (c_int, [c_int] + [c_vp] * 7 + [ctypes.c_double] * 3 + [c_float] * 2 + [c_vp]),
    "shb_adam_step_mixed": (c_int, [c_int] + [c_vp] * 8 + [ctypes.c_double] * 3 + [c_float] * 2 + [c_vp]),
    "shb_cast_bf16": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "shb_slab_tensor_bytes": (c_size, [c_int] * 4),
    "shb_slab_from_rows": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp] + [c_int] * 7 + [c_vp]),
    "shb_slab_to_rows": (c_int, [c_vp, c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    "shb_slab_l1_workspace": (c_size, []),
    "shb_slab_l1_fwd": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_size, c_vp] + [c_int] * 4 + [c_vp]),
    "shb_slab_l1_bwd": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    "shb_slab_pool": (c_int, [c_vp] * 6 + [c_int] * 6 + [c_vp]),
    "shb_slab_weight_image_bytes": (c_size, [c_int] * 4),
    "shb_slab_weight_images": (c_int, [c_vp, c_vp, c_vp] + [c_int] * 6 + [c_vp]),
    "shb_slab_weight_images_batch": (c_int, [c_int] + [c_vp] * 8 + [c_int, c_vp]),
    "shb_slab_conv_supported": (c_int, [c_int] * 4),
    "shb_slab_conv": (c_int, [c_vp] * 7 + [c_int] * 10 + [c_vp]),
    "shb_build_conv_groups": (c_int, [c_vp, c_vp] + [c_int] * 4 + [c_vp] * 6),
    "shb_slab_gconv_plan": (c_int, [c_int] * 4 + [c_vp, c_vp]),
    "shb_slab_gconv": (c_int, [c_vp] * 5 + [c_int] * 3 + [c_vp] * 4 + [c_int] * 10 + [c_vp]),
    "shb_slab_wgrad_supported": (c_int, [c_int] * 4),
    "shb_slab_wgrad_workspace": (c_size, [c_int] * 4),
    "shb_slab_wgrad": (c_int, [c_vp] * 6 + [c_size] + [c_int] * 10 + [c_vp]),
}

ACT_ENUM = {"identity": 0, "relu": 1, "elu": 2, "leaky_relu": 3, "sigmoid": 4, "tanh": 5}
F32, BF16 = 0, 1


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: the CUDA extension has not been built "
            "(run `python semantichuman_b200/_build.py`); semantichuman_b200 has no CPU fallback")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.shb_abi_version() != 2:
        raise ImportError("libshb200.so ABI version mismatch")
    return lib


lib = _load()
LIB_PATH = _LIB_PATH


def check(code, what):
    if code != 0:
        msg = lib.shb_error_string(int(code)).decode()
        raise RuntimeError(f"{what} failed with code {code}: {msg}")
