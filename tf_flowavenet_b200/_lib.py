"""ctypes binding of libflowavenet_b200.so (the C ABI declared in include/flowavenet_b200.h).

There is deliberately NO fallback: if the shared library is missing or fails to load, importing the
product API raises.  Build it with ``python -m tf_flowavenet_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflowavenet_b200.so")

FWN_FP32, FWN_MIXED_BF16, FWN_MIXED_FP16 = 0, 1, 2


class FwnConfig(C.Structure):
    _fields_ = [("n_block", C.c_int32), ("n_flow", C.c_int32), ("n_layer", C.c_int32), ("num_mels", C.c_int32),
                ("filter_size", C.c_int32), ("affine", C.c_int32), ("causal", C.c_int32), ("n_upsample", C.c_int32),
                ("upsample_scales", C.c_int32 * 4), ("gin_channels", C.c_int32), ("n_speakers", C.c_int32),
                ("precision", C.c_int32)]


_p, _i, _l, _fp = C.c_void_p, C.c_int, C.c_int64, C.c_void_p  # device pointers travel as void*

# name -> (restype, argtypes); every symbol include/flowavenet_b200.h declares
SIGNATURES = {
    "fwn_last_error": (C.c_char_p, []),
    "fwn_abi_version": (_i, []),
    "fwn_create": (_i, [C.POINTER(FwnConfig), C.POINTER(_p)]),
    "fwn_destroy": (_i, [_p]),
    "fwn_num_params": (_i, [_p]),
    "fwn_param_info": (_i, [_p, _i, C.POINTER(C.c_char_p), C.POINTER(_l * 4), C.POINTER(_i)]),
    "fwn_set_param": (_i, [_p, C.c_char_p, _fp, _l, _p]),
    "fwn_get_param": (_i, [_p, C.c_char_p, _fp, _l, _p]),
    "fwn_prepack": (_i, [_p, _p]),
    "fwn_workspace_bytes": (_l, [_p, _i, _i]),
    "fwn_forward": (_i, [_p, _fp, _fp, _fp, _i, _i, _fp, _fp, _fp, _i, _p, _l, _p]),
    "fwn_reverse": (_i, [_p, _fp, _fp, _fp, _i, _i, _fp, _p, _l, _p]),
    "fwn_forward_host": (_i, [_p, _fp, _fp, _fp, _i, _i, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fwn_reverse_host": (_i, [_p, _fp, _fp, _fp, _i, _i, _fp]),
    "fwn_reverse_chunk": (_i, [_p, _fp, _fp, _i, _i, _i, _i, _fp, _p, _l, _p]),
    "fwn_train_enable": (_i, [_p, _p]),
    "fwn_train_workspace_bytes": (_l, [_p, _i, _i]),
    "fwn_param_floats": (_l, [_p]),
    "fwn_param_offset": (_l, [_p, _i]),
    "fwn_grad_floats": (_l, [_p]),
    "fwn_params_ptr": (_i, [_p, C.POINTER(_p)]),
    "fwn_loss_and_grads": (_i, [_p, _fp, _fp, _fp, _i, _i, _fp, _fp, _fp, _l, _p, _l, _p]),
    "fwn_grad_bucket_count": (_i, [_p]),
    "fwn_grad_bucket_range": (_i, [_p, _i, C.POINTER(_l), C.POINTER(_l)]),
    "fwn_grad_bucket_wait": (_i, [_p, _i, _p]),
    "fwn_grad_global_norm": (_i, [_p, _fp, _fp, _p]),
    "fwn_apply_gradients": (_i, [_p, _fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _l, _p]),
    "fwn_set_split_terms": (_i, [_p, _i, _i]),
    "fwn_set_train_compute": (_i, [_p, _i]),
    "fwn_set_layer_fusion": (_i, [_p, _i]),
    "fwn_wgrad_bf16": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _p]),
    "fwn_set_train_exact_forward": (_i, [_p, _i]),
    "fwn_get_train_state": (_i, [_p, _i, _fp, _l, _p]),
    "fwn_set_train_state": (_i, [_p, _i, _fp, _l, _p]),
    "fwn_repack": (_i, [_p, _p]),
    "fwn_last_launches": (_l, [_p]),
    "fwn_profile_enable": (_i, [_p, _i]),
    "fwn_profile_read": (_i, [_p, C.POINTER(C.c_double * 8), C.POINTER(_l * 8), C.POINTER(C.c_double * 8)]),
    "fwn_receptive_halo": (_i, [_p]),
    "fwn_squeeze": (_i, [_fp, _fp, _i, _i, _i, _p]),
    "fwn_unsqueeze": (_i, [_fp, _fp, _i, _i, _i, _p]),
    "fwn_change_order": (_i, [_fp, _fp, _l, _i, _p]),
    "fwn_actnorm_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _l, _i, _p]),
    "fwn_actnorm_rev": (_i, [_fp, _fp, _fp, _fp, _l, _i, _p]),
    "fwn_actnorm_ddi": (_i, [_fp, _fp, _fp, _l, _i, _p]),
    "fwn_affine_fwd": (_i, [_fp, _fp, _fp, _fp, _l, _i, _i, _p]),
    "fwn_affine_rev": (_i, [_fp, _fp, _fp, _l, _i, _i, _p]),
    "fwn_upsample_stage": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _p]),
    "fwn_conv1d": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fwn_conv1d_bf16": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "fwn_zero_conv1d": (_i, [_fp, _fp, _fp, _fp, _fp, _l, _i, _i, _p]),
    "fwn_gated_activation": (_i, [_fp, _fp, _fp, _l, _p]),
    "fwn_residual_scale": (_i, [_fp, _fp, _fp, _l, _p]),
    "fwn_add": (_i, [_fp, _fp, _fp, _l, _i, _p]),
    "fwn_log_p": (_i, [_fp, _fp, _l, _p]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it is missing -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `python -m tf_flowavenet_b200.build` "
                              "(libflowavenet_b200 has no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here == header/library drift
            fn.restype, fn.argtypes = res, args
        if l.fwn_abi_version() != 1:
            raise ImportError("libflowavenet_b200 ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().fwn_last_error().decode("utf-8", "replace")
        if msg == "g is None":  # model.py:320-321 raises ValueError('g is None')
            raise ValueError(msg)
        raise RuntimeError("libflowavenet_b200: " + msg)


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor, or NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
