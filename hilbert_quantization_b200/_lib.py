"""ctypes binding of libhq_b200.so (include/hq_b200.h).  No CPU fallback: importing this
module without a loadable CUDA library raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_i64, _i32, _p, _dbl = C.c_int64, C.c_int, C.c_void_p, C.c_double


class IndexLayout(C.Structure):
    """struct hq_index_layout"""
    _fields_ = [("L", C.c_int32), ("Lsum", C.c_int32), ("lvl_off", C.c_int32 * 8),
                ("lvl_w", C.c_int32 * 8), ("lvl_keff", C.c_int32 * 8)]


# name -> (restype, argtypes); mirrors include/hq_b200.h declaration by declaration
SIGNATURES = {
    "hq_version": (_i32, []),
    "hq_last_error": (C.c_char_p, []),
    "hq_sm_count": (_i32, []),
    "hq_launch_count": (_i64, [_i32]),
    "hq_d2xy_batch": (_i32, [_i32, _i64, _i64, _p, _p, _p]),
    "hq_xy2d_batch": (_i32, [_i32, _p, _p, _i64, _p, _p]),
    "hq_map_to_2d": (_i32, [_p, _i64, _i64, _i64, _i32, _i32, _p, _i64, _p]),
    "hq_map_from_2d": (_i32, [_p, _i64, _i32, _i64, _i64, _i32, _p, _i64, _p]),
    "hq_fused_scratch_bytes": (_i64, [_i64, _i32, _i32]),
    "hq_fused_scratch_bytes_min_level": (_i64, [_i64, _i32, _i32, _i32]),
    "hq_map_index_fused": (_i32, [_p, _i32, _i64, _i64, _i64, _i32, _p, _i64, _p, _i64, _p, _i32, _i32, _p, _i64, _p, _i64, _p]),
    "hq_map_index_fused_ml": (_i32, [_p, _i32, _i64, _i64, _i64, _i32, _p, _i64, _p, _i64, _p, _i32, _i32, _i32, _p, _i64, _p, _i64, _p]),
    "hq_map_index_stream": (_i32, [_p, _i64, _i32, _p, _p, _i32, _i32, _i32, _p, _i64, _p, _i64, _p]),
    "hq_block_means": (_i32, [_p, _i64, _i32, _i32, _i64, _i32, _i32, _p, _p, _i32, _p, _i64, _p]),
    "hq_map_index_quant": (_i32, [_p, _i64, _i64, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p, _i64, _p, _p, _i64, _p]),
    "hq_quantize_u8": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "hq_dequantize_u8": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "hq_index_row_lengths": (_i32, [_p, _i64, C.POINTER(IndexLayout), _p, _p]),
    "hq_filter_level": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _i32, _p, _p, _i32, _p, _i64, _dbl, _p, _i64, _p, _p, _p, _p]),
    "hq_filter_select": (_i32, [_p, _i64, _i64, _i32, _p, _p, _dbl, _p, _i64, _p, _p]),
    "hq_filter_select_prev": (_i32, [_p, _i64, _i64, _i32, _p, _p, _i64, _p, _p, _dbl, _p, _i64, _p, _p]),
    "hq_filter_fast_supported": (_i32, [C.POINTER(IndexLayout)]),
    "hq_filter_level_norms": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _p, _p, _p]),
    "hq_filter_fast_scratch_bytes": (_i64, [_i64, _i32, C.POINTER(IndexLayout)]),
    "hq_filter_fast_mode": (_i32, [_i64, _i32, C.POINTER(IndexLayout)]),
    "hq_filter_fast_fallback_offset": (_i64, [_i64, _i32, C.POINTER(IndexLayout)]),
    "hq_filter_fast_window_layout": (_i32, [_i64, _i32, C.POINTER(IndexLayout), _p]),
    "hq_filter_rows_max_queries": (_i32, []),
    "hq_filter_rows_cols": (_i32, [C.POINTER(IndexLayout)]),
    "hq_filter_rows_pack": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _p, _p]),
    "hq_filter_fast_rows": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _p, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _i32, _p,
                                   _p, _i64, _p, _p, _p, _i64, _p]),
    "hq_filter_fast": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _p, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _i32, _p,
                              _p, _i64, _p, _p, _p, _i64, _p]),
    "hq_filter_tc_supported": (_i32, [C.POINTER(IndexLayout)]),
    "hq_filter_tc_plan": (_i32, [_i64, _i32, _p, _p]),
    "hq_filter_tc_packed_cols": (_i32, [C.POINTER(IndexLayout)]),
    "hq_filter_tc_valid_pitch": (_i64, [_i64]),
    "hq_filter_tc_pack": (_i32, [_p, _p, _i64, C.POINTER(IndexLayout), _i32, _p, _p]),
    "hq_filter_tc_valid": (_i32, [_p, _i64, C.POINTER(IndexLayout), _p, _i64, _p]),
    "hq_row_norms": (_i32, [_p, _i64, _i64, _i64, _p, _p]),
    "hq_rerank_scores_f32": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _i64, _p, _i64, _p, _i64, _p]),
    "hq_paired_cosine01": (_i32, [_p, _p, _p, _p, _i64, _i64, _i64, _p, _p]),
    "hq_rerank_sparse_topk_supported": (_i32, [_i64, _i64, _i64, _i32]),
    "hq_rerank_sparse_topk_scratch_bytes": (_i64, [_i32, _i32]),
    "hq_rerank_sparse_topk": (_i32, [_p, _p, _i64, _i64, _i64, _p, _i64, C.c_float, _p, _p, _i32, _i64, _p, _i64, _i32, _i64, _p, _p, _p, _i64, _p]),
    "hq_rerank_scores_sparse_f32": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _i64, _p, _i64, _p, _i64, _p]),
    "hq_topk_from_scores": (_i32, [_p, _i64, _i64, _i32, _i32, _i64, _p, _p, _p]),
    "hq_topk_chunked_scratch_bytes": (_i64, [_i64, _i32, _i32]),
    "hq_topk_from_scores_chunked": (_i32, [_p, _i64, _i64, _i32, _i32, _i64, _p, _p, _p, _i64, _p]),
    "hq_rerank_scratch_bytes": (_i64, [_i64, _i32]),
    "hq_rerank_topk_f32": (_i32, [_p, _p, _i64, _i64, _i64, _p, _p, _i32, _i64, _p, _i64, _i32, _i64, _p, _p, _p, _i64, _p]),
    "hq_to_bf16": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _p]),
    "hq_rerank_bf16_scratch_bytes": (_i64, [_i64, _i32, _i32]),
    "hq_to_bf16_unit": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "hq_shard_ingest_supported": (_i32, [_i64]),
    "hq_shard_ingest": (_i32, [_p, _i64, _i64, _i64, _p, _i32, _p, _i64, _p, _p, _i64, _p]),
    "hq_rerank_topk_unit_bf16": (_i32, [_p, _i64, _p, _i64, _p, _p, _i32, _i64, _i64, _p, _i64, _p, _i64, _p, _i32, _p, _i64, _i32, _i64,
                                        C.c_float, _p, _p, _p, _p, _i64, _p]),
    "hq_bf16_unit_error_max": (_i32, [_p, _i64, _i64, _i64, _p, _p, _i64, _p, _p]),
    "hq_comprehensive_scores": (_i32, [_p, _i64, _i32, _i32, _i64, _p, _i32, _i64, _p, _p, _i64, _p, _p]),
    "hq_offset_square_means": (_i32, [_p, _i64, _i32, _i64, _p, _i64, _p]),
    "hq_pearson01_matrix": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i32, _p, _i64, _p]),
    "hq_gcut_hist": (_i32, [_p, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _p]),
    "hq_gcut_scan": (_i32, [_p, _i32, _p, _p, _p, _p, _p]),
    "hq_gcut_ties": (_i32, [_p, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _p]),
    "hq_gcut_apply": (_i32, [_p, _i64, _i64, _i32, _p, _i64, _p, _p, _p, _p]),
    "hq_kernel_timing": (_i32, [_i32]),
    "hq_kernel_timing_read": (_i32, [_p, _p, _i32]),
    "hq_topk_merge": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "hq_topk_merge_strided": (_i32, [_p, _p, _i32, _i32, _i32, _i64, _i64, _p, _p, _p]),
    "hq_core_level_sims": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _i32, _p, _p]),
}

HQ_OK, HQ_EINVAL, HQ_ECUDA, HQ_EUNSUPPORTED = 0, -1, -2, -3


class HQLibraryError(RuntimeError):
    """The CUDA library is missing, unloadable or returned HQ_ECUDA."""


def _load():
    path = _build.LIB_PATH
    if not os.path.exists(path) or _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine and no prebuilt library
            if not os.path.exists(path):
                raise HQLibraryError(
                    f"libhq_b200.so is not built ({path}) and could not be compiled here: {e}. "
                    "There is no CPU fallback; run `python -m hilbert_quantization_b200.build`.") from e
            import warnings
            warnings.warn(f"libhq_b200.so is older than its sources and the rebuild failed ({e}); loading the STALE library "
                          f"{path}", RuntimeWarning)
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise HQLibraryError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise HQLibraryError(f"{path} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
LIB_PATH = _build.LIB_PATH


def last_error() -> str:
    return lib.hq_last_error().decode("utf-8", "replace")


def check(rc: int, exc=None):
    """Map a status code to an exception (HQ_EINVAL -> `exc` or ValueError)."""
    if rc == HQ_OK:
        return
    msg = last_error()
    if rc == HQ_EINVAL:
        raise (exc or ValueError)(msg)
    if rc == HQ_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise HQLibraryError(msg)
