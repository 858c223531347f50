"""ctypes binding of ``libfreegaussian_b200.so`` (the C ABI declared in include/fg_api.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libfreegaussian_b200.so"

ABI_VERSION = 12

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pi = C.POINTER(C.c_int)



class AdamSegment(C.Structure):  # fg_adam_segment
    _fields_ = [("param", _vp), ("grad", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("n", _i64), ("first", _i64),
                ("row_len", C.c_int32), ("split", C.c_int32), ("lr", _f32), ("lr_rest", _f32)]


class RefineConfig(C.Structure):  # fg_refine_config
    _fields_ = [("densify_grad_thresh", _f32), ("densify_size_thresh", _f32), ("split_screen_size", _f32),
                ("cull_alpha_thresh", _f32), ("cull_scale_thresh", _f32), ("cull_screen_size", _f32),
                ("max_dim", _f32), ("n_split_samples", C.c_int32), ("use_screen", C.c_int32),
                ("cull_big", C.c_int32), ("densify", C.c_int32)]


class RefineArray(C.Structure):  # fg_refine_array
    _fields_ = [("in_", _vp), ("out", _vp), ("row_floats", C.c_int32), ("zero_new", C.c_int32)]


class MlpPackSegment(C.Structure):  # fg_mlp_pack_segment
    _fields_ = [("src", _vp), ("dst_hi", _vp), ("dst_lo", _vp), ("src_ld", C.c_int32), ("src_col0", C.c_int32),
                ("rows", C.c_int32), ("cols", C.c_int32), ("dst_ld", C.c_int32), ("dst_col0", C.c_int32),
                ("transpose", C.c_int32)]


class ProjectBwdPub(C.Structure):  # fg_project_bwd_pub
    _fields_ = [("campos", _vp), ("mask", _vp), ("prefix", _vp), ("rgb", _vp), ("offsets", _vp), ("nnz", _vp),
                ("words", C.c_int32), ("phase", C.c_int32)]


XCHG_MAX_RANKS = 16
XCHG_FLAG_BYTES = 65536


class XchgPeers(C.Structure):  # fg_xchg_peers
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("buf", _vp * XCHG_MAX_RANKS), ("mc", _vp),
                ("flags", _vp * XCHG_MAX_RANKS)]


ADAM_MAX_SEGMENTS = 8
MLP_PACK_MAX_SEGMENTS = 32
MLP_EMBED_LD = 96
MLP_HEAD_LD = 32
MLP_RELU, MLP_LINEAR, MLP_DGRAD = 0, 1, 2
REFINE_MAX_ARRAYS = 24

# name -> (restype, argtypes); mirrors include/fg_api.h one to one
SIGNATURES = {
    "fg_adam_step": (_i32, [_i32, C.POINTER(AdamSegment), _i32, C.c_double, C.c_double, C.c_double, _vp]),
    "fg_refine_workspace_bytes": (_i64, [_i64]),
    "fg_refine_plan": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp, C.POINTER(RefineConfig), _vp, _vp, _vp, _i64, _vp]),
    "fg_refine_map": (_i32, [_i64, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "fg_refine_gather": (_i32, [_i64, _i64, _vp, _i32, C.POINTER(RefineArray), _vp]),
    "fg_refine_children": (_i32, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_last_error": (C.c_char_p, []),
    "fg_abi_version": (_i32, []),
    "fg_launch_count": (C.c_longlong, []),
    "fg_set_option": (_i32, [C.c_char_p, _i32]),
    "fg_measure_fp32_tflops": (_i32, [C.POINTER(C.c_double), _vp]),
    "fg_project_fwd": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _f32, _i32,
                              _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                              _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_project_bwd": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _f32,
                              _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(ProjectBwdPub), _vp]),
    "fg_xchg_pub_bytes": (_i64, [_i32, _i32]),
    "fg_xchg_pub_layout": (_i32, [_i32, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                  C.POINTER(C.c_int32)]),
    "fg_xchg_barrier": (_i32, [C.POINTER(XchgPeers), C.c_uint32, _vp]),
    "fg_xchg_allreduce_f32": (_i32, [C.POINTER(XchgPeers), _i64, _i64, C.c_uint32, _i32, _vp]),
    "fg_xchg_sh_bwd_views": (_i32, [C.POINTER(XchgPeers), _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "fg_scan_workspace_bytes": (_i64, [_i64]),
    "fg_exclusive_scan_i32": (_i32, [_i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fg_isect_emit": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_radix_sort_workspace_bytes": (_i64, [_i64]),
    "fg_radix_sort_pairs_u64_u32": (_i32, [_i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _pi, _vp]),
    "fg_radix_sort_pairs_u32_u32": (_i32, [_i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _pi, _vp]),
    "fg_isect_offsets": (_i32, [_i64, _vp, _i32, _i32, _i32, _vp, _vp]),
    "fg_isect_depth_keys": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "fg_depth_sort_workspace_bytes": (_i64, [_i64]),
    "fg_depth_sort_visible": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fg_gather_i32": (_i32, [_i64, _vp, _vp, _vp, _vp]),
    "fg_isect_emit_tiles": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_isect_offsets_tiles": (_i32, [_i64, _vp, _i32, _i32, _i32, _vp, _vp]),
    "fg_isect_ids_from_tiles": (_i32, [_i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "fg_bin_coarse_dims": (_i32, [_i32, _i32, _pi, _pi]),
    "fg_bin_count": (_i32, [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_bin_tile_scan_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "fg_bin_tile_scan": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fg_bin_coarse_emit": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_bin_ranked_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "fg_bin_count_cells": (_i32, [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp]),
    "fg_bin_cell_scan": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp]),
    "fg_bin_ranked_emit": (_i32, [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "fg_bin_fine": (_i32, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "fg_rasterize_fwd": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "fg_rasterize_bwd": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_densify_stats": (_i32, [_i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp]),
    "fg_assign_masks": (_i32, [_i64, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "fg_l1_ssim_workspace_floats": (_i64, [_i32, _i32]),
    "fg_l1_ssim_fwd": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp]),
    "fg_l1_ssim_bwd": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp]),
    "fg_depth_fixup_fwd": (_i32, [_i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "fg_depth_fixup_bwd": (_i32, [_i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "fg_render_front_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "fg_render_front": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _f32, _f32, _i32,
                               _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                               _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64), _vp, _i64, _vp,
                               _i64, _vp, _i64, _vp]),
    "fg_render_back_workspace_bytes": (_i64, [_i32, _i32, _i32, _i64]),
    "fg_render_back": (_i32, [_i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _i64,
                              _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fg_mlp_linear": (_i32, [_i32, _i64, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_mlp_wgrad": (_i32, [_i64, _vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "fg_mlp_pack": (_i32, [_i32, C.POINTER(MlpPackSegment), _vp]),
    "fg_deform_embed": (_i32, [_i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "fg_deform_embed_bwd": (_i32, [_i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "fg_deform_apply_fwd": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_deform_apply_bwd": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_pack_workspace_bytes": (_i64, [_i64]),
    "fg_pack_plan": (_i32, [_i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fg_pack_gather": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp, _vp]),
    "fg_pack_remap": (_i32, [_i64, _vp, _vp, _vp]),
    "fg_rows_workspace_bytes": (_i64, [_i64]),
    "fg_rows_active": (_i32, [_i64, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "fg_rows_gather": (_i32, [_i64, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fg_time_branch_fwd": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_time_branch_bwd": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fg_knn_workspace_bytes": (_i64, [_i64]),
    "fg_knn_f32": (_i32, [_i64, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
}


class FgError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load the library once.  Raises (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FgError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the render path)"
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if L.fg_abi_version() != ABI_VERSION:
            raise FgError(f"ABI mismatch: library {L.fg_abi_version()} vs binding {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().fg_last_error().decode()
        if code == 1:
            raise AssertionError(msg)  # gsplat convention: bad arguments are assertion failures
        raise FgError(f"fg error {code}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensor must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), "internal: non-contiguous tensor handed to the C ABI"
    return t.data_ptr()


def on_device_of(arg):
    """Decorator: run the call with the CUDA device of tensor argument ``arg`` (position or keyword name) current, so
    the kernels go to that device's current stream even when the caller's current device is another one.  Non-CUDA
    arguments pass through untouched (the entry points then raise their own "no CPU path" errors)."""
    import functools
    import inspect

    def deco(fn):
        names = list(inspect.signature(fn).parameters)
        idx = arg if isinstance(arg, int) else names.index(arg)
        name = names[idx]

        @functools.wraps(fn)
        def wrapped(*a, **kw):
            import torch
            t = a[idx] if idx < len(a) else kw.get(name)
            if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
                with torch.cuda.device(t.device):
                    return fn(*a, **kw)
            return fn(*a, **kw)

        return wrapped

    return deco


def launch_count() -> int:
    return int(lib().fg_launch_count())
