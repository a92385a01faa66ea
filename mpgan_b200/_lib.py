"""ctypes binding of libmpgan_b200.so (the C ABI in include/mpgan_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_VARIANT = os.environ.get("MPG_LIB_VARIANT", "")   # experiment / trace builds (mpgan_b200/build.py)
LIB_PATH = os.path.join(_HERE, "lib", f"libmpgan_b200{'_' + _VARIANT if _VARIANT else ''}.so")

_f = C.c_void_p  # device pointers travel as integers
_i, _fl, _u64, _u32, _sz = C.c_int, C.c_float, C.c_uint64, C.c_uint32, C.c_size_t

_SIGS = {
    "mpg_version": (C.c_int, []),
    "mpg_last_error": (C.c_char_p, []),
    "mpg_features": (C.c_int, []),
    "mpg_launch_count": (C.c_ulonglong, []),
    "mpg_probe": (None, [_i, _f, _f]),
    "mpg_linear_fwd": (C.c_int, [_f, _i, _f, _f, _f, _i, _i, _i, _i, _fl, _fl, _u64, _f, _u32, _i, _f]),
    "mpg_linear_bwd": (C.c_int, [_f, _f, _f, _i, _f, _f, _f, _i, _i, _f, _f, _i, _i, _i, _i, _fl, _fl, _u64, _f,
                                 _u32, _i, _f]),
    "mpg_fn_supported": (C.c_int, [_i, _i, _i, _i, _i, _fl]),
    "mpg_fn_workspace_bytes": (C.c_size_t, [_i] * 5),
    "mpg_fn_fwd": (C.c_int, [_f, _i, _i, _f, _i, _i, _i, _f, _f, _f, _f, _f, _f, _i, _i, _i, _fl, _fl, _u64, _f, _f,
                             _sz, _f, _f, _f, _f]),
    "mpg_fn_bwd": (C.c_int, [_f, _f, _f, _f, _i, _i, _f, _i, _i, _i, _f, _f, _f, _i, _i, _i, _fl, _fl, _u64, _f, _f,
                             _sz, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_edge_workspace_bytes": (C.c_size_t, [_i] * 6),
    "mpg_edge_fwd": (C.c_int, [_f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _fl, _fl,
                               _u64, _f, _i, _f, _sz, _f, _f]),
    "mpg_edge_bwd": (C.c_int, [_f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _fl, _fl,
                               _u64, _f, _i, _f, _sz, _f, _f, _i, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_edge_fwd_workspace_bytes": (C.c_size_t, [_i] * 6),
    "mpg_edge_bwd_saved": (C.c_int, [_f, _sz, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                     _fl, _fl, _u64, _f, _i, _f, _sz, _f, _f, _i, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_rank_mask": (C.c_int, [_f, _i, _f, _i, _i, _i, _f, _f]),
    "mpg_particle_order": (C.c_int, [_f, _i, _i, _f, _f, _f]),
    "mpg_batch_order": (C.c_int, [_f, _i, _i, _f, _f]),
    "mpg_permute_rows": (C.c_int, [_f, _i, _f, _i, _f, _i, _i, _i, _i, _f]),
    "mpg_ls_loss_fwd": (C.c_int, [_f, _i, _i, _fl, _fl, _f, _f]),
    "mpg_ls_loss_bwd": (C.c_int, [_f, _f, _i, _i, _fl, _fl, _f, _f]),
    "mpg_split_mask": (C.c_int, [_f, _i, _i, _f, _f]),
    "mpg_gen_tail_fwd": (C.c_int, [_f, _f, _f, _i, _i, _i, _f]),
    "mpg_gen_tail_bwd": (C.c_int, [_f, _f, _f, _i, _i, _i, _i, _f]),
    "mpg_pool_fwd": (C.c_int, [_f, _f, _f, _i, _i, _i, _i, _f]),
    "mpg_pool_bwd": (C.c_int, [_f, _f, _f, _i, _i, _i, _i, _f]),
    "mpg_unary_fwd": (C.c_int, [_f, _f, _sz, _i, _f]),
    "mpg_unary_bwd": (C.c_int, [_f, _f, _f, _sz, _i, _f]),
    "mpg_sn_fwd": (C.c_int, [_f, _f, _f, _f, _f, _i, _i, _f]),
    "mpg_sn_bwd": (C.c_int, [_f, _f, _f, _f, _f, _f, _i, _i, _f]),
    "mpg_rmsprop": (C.c_int, [_f, _f, _f, _sz, _fl, _fl, _fl, _fl, _f]),
    "mpg_attn_fwd": (C.c_int, [_f, _i, _f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _f, _f, _f]),
    "mpg_attn_bwd": (C.c_int, [_f, _i, _f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f]),
    "mpg_residual_dropout_fwd": (C.c_int, [_f, _f, _f, _sz, _i, _fl, _u64, _f, _u32, _f]),
    "mpg_residual_dropout_bwd": (C.c_int, [_f, _f, _sz, _i, _fl, _u64, _f, _u32, _f]),
    "mpg_knn_select": (C.c_int, [_f, _i, _f, _i, _i, _i, _i, _i, _f, _f]),
    "mpg_edge_nbr_fwd": (C.c_int, [_f, _i, _i, _f, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                   _fl, _fl, _u64, _f, _f, _sz, _f, _f]),
    "mpg_edge_nbr_bwd": (C.c_int, [_f, _i, _i, _f, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                   _fl, _fl, _u64, _f, _f, _sz, _f, _f, _i, _f, _f, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_cond_columns": (C.c_int, [_f, _i, _f, _i, _f, _sz, _i, _i, _f]),
    "mpg_edge_bwd2_workspace_bytes": (C.c_size_t, [_i] * 6),
    "mpg_edge_bwd2": (C.c_int, [_f, _i, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _fl, _fl, _u64,
                                _f, _f, _sz, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_compact_map_ints": (C.c_size_t, [_i, _i]),
    "mpg_compact_map": (C.c_int, [_f, _i, _i, _f, _f, _f]),
    "mpg_edge_set_compaction": (C.c_int, [_f]),
    "mpg_split_mask_bwd": (C.c_int, [_f, _f, _i, _sz, _f]),
    "mpg_pool_dmask": (C.c_int, [_f, _f, _f, _i, _i, _i, _fl, _f]),
    "mpg_gen_postprocess": (C.c_int, [_f, _i, _f, _i, _sz, _i, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), _i, _f]),
    "mpg_layernorm_fwd": (C.c_int, [_f, _f, _f, _f, _f, _f, _sz, _i, _fl, _f]),
    "mpg_layernorm_bwd": (C.c_int, [_f, _f, _f, _f, _f, _f, _f, _f, _sz, _i, _f]),
    "mpg_mab_supported": (C.c_int, [_i, _i, _i, _i]),
    "mpg_mab_workspace_bytes": (C.c_size_t, [_i]),
    "mpg_mab_fwd": (C.c_int, [_f, _i, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _fl, _fl, _fl, _u64, _f, _i,
                              _f, _sz, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_mab_bwd": (C.c_int, [_f, _i, _f, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _fl, _fl, _fl, _u64, _f, _i,
                              _f, _sz, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f]),
    "mpg_peer_flag_words": (C.c_size_t, [_i, _i]),
    "mpg_allreduce_rmsprop": (C.c_int, [_f, _f, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _f, _sz, _i, _i, _i, _fl,
                                        _fl, _fl, _f]),
}

_lib = None


def lib():
    """Loads the shared library on first use; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m mpgan_b200.build` "
                "(there is no CPU / PyTorch fallback for the mpgan_b200 ops)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().mpg_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Raises for CPU tensors: no fallback path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("mpgan_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype not in (torch.float32, torch.int64, torch.int32, torch.uint8):
        raise RuntimeError(f"mpgan_b200 ops are fp32 at the boundary, got {t.dtype}")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
