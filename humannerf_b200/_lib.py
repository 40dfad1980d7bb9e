"""ctypes binding of libhumannerf_b200.so (the C ABI in include/humannerf_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a
`RuntimeError` is raised.  Only raw device pointers, sizes and the current CUDA
stream cross the boundary; PyTorch merely owns the memory.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

HN_ABI_VERSION = 6
HN_MAX_LAYERS = 16
PRECISION_FP32 = 0
PRECISION_BF16 = 1
BWD_DATA = 1          # phases of hn_field_bwd_bf16_phases
BWD_WEIGHTS = 2

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libhumannerf_b200.so')
_lib: Optional[C.CDLL] = None

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int


class FieldParams(C.Structure):
    _fields_ = [('nr_layers', C.c_int32), ('nr_width', C.c_int32), ('nr_skip', C.c_int32),
                ('nr_freqs', C.c_int32), ('cond_dim', C.c_int32), ('cn_layers', C.c_int32),
                ('cn_width', C.c_int32), ('cn_skip', C.c_int32), ('cn_freqs', C.c_int32),
                ('reserved', C.c_int32),
                ('nr_w', _p * HN_MAX_LAYERS), ('nr_b', _p * HN_MAX_LAYERS),
                ('cn_w', _p * HN_MAX_LAYERS), ('cn_b', _p * HN_MAX_LAYERS)]


class WgradJob(C.Structure):
    _fields_ = [('dz', _p), ('x', _p), ('dw', _p), ('db', _p), ('f_dz', C.c_int32), ('f_x', C.c_int32),
                ('m', C.c_int32), ('x_f0', C.c_int32), ('n', C.c_int32), ('n_valid', C.c_int32),
                ('dw_ld', C.c_int32), ('dw_col0', C.c_int32), ('bias_col', C.c_int32), ('reserved', C.c_int32)]


class AdamTensor(C.Structure):
    _fields_ = [('param', _p), ('grad', _p), ('exp_avg', _p), ('exp_avg_sq', _p), ('numel', C.c_int64)]


class FieldGrads(C.Structure):
    _fields_ = [('nr_w', _p * HN_MAX_LAYERS), ('nr_b', _p * HN_MAX_LAYERS),
                ('cn_w', _p * HN_MAX_LAYERS), ('cn_b', _p * HN_MAX_LAYERS)]


# name -> (restype, argtypes); every symbol declared in include/humannerf_b200.h
SIGNATURES = {
    'hn_version': (_int, []),
    'hn_last_error': (C.c_char_p, []),
    'hn_launch_count': (_i64, []),
    'hn_has_tcgen05': (_int, []),
    'hn_gen_rays': (_int, [_int, _int, _p, _p, _p, _p, _p]),
    'hn_rays_bbox': (_int, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    'hn_compact_rays_workspace_bytes': (_i64, [_i64]),
    'hn_compact_rays': (_int, [_i64] + [_p] * 13),
    'hn_sample_z': (_int, [_i64, _int, _p, _p, _p, _p, _p, _p]),
    'hn_motion_field_workspace_bytes': (_i64, [_int, _int]),
    'hn_motion_field_fwd': (_int, [_i64, _int, _int, _int] + [_p] * 13 + [_i64, _p]),
    'hn_motion_field_bwd': (_int, [_i64, _int, _int, _int] + [_p] * 19 + [_i64, _p]),
    'hn_field_workspace_bytes': (_i64, [C.POINTER(FieldParams), _i64, _int, _int]),
    'hn_field_fwd': (_int, [C.POINTER(FieldParams), _i64, _p, _p, _p, _int, _p, _p, _p, _p, _i64, _int, _int, _p, _p, _p]),
    'hn_field_bwd_scratch_bytes': (_i64, [C.POINTER(FieldParams), _i64]),
    'hn_field_bwd': (_int, [C.POINTER(FieldParams), C.POINTER(FieldGrads), _i64, _p, _p, _p, _int, _p, _p, _p, _p,
                            _p, _i64, _p, _i64, _p]),
    'hn_field_bwd_bf16_scratch_bytes': (_i64, [C.POINTER(FieldParams), _i64]),
    'hn_field_bwd_bf16': (_int, [C.POINTER(FieldParams), C.POINTER(FieldGrads), _i64, _p, _p, _p, _p, _int, _p, _p, _p,
                                 _p, _p, _i64, _p, _p, _i64, _p, _p]),
    'hn_field_bwd_bf16_phases': (_int, [C.POINTER(FieldParams), C.POINTER(FieldGrads), _i64, _p, _p, _p, _p, _int, _p, _p, _p,
                                 _p, _p, _i64, _p, _p, _i64, _p, _int, _p]),
    'hn_field_packed_bytes': (_i64, [C.POINTER(FieldParams)]),
    'hn_field_pack_bf16': (_int, [C.POINTER(FieldParams), _p, _p, _p]),
    'hn_field_repack_bf16': (_int, [C.POINTER(FieldParams), _p, _p, _p]),
    'hn_pose_fk_fwd': (_int, [_int, _p, _p, _p, _p, _p, _p, _p]),
    'hn_pose_fk_bwd': (_int, [_int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'hn_adam_plan_bytes': (_i64, [C.POINTER(AdamTensor), _int]),
    'hn_adam_n_chunks': (_i64, [C.POINTER(AdamTensor), _int]),
    'hn_adam_plan': (_int, [C.POINTER(AdamTensor), _int, _p, _i64]),
    'hn_adam_step': (_int, [_p, _i64, _p, _p, _p, _p, _p, C.c_float, C.c_float, C.c_float, _p]),
    'hn_composite_fwd': (_int, [_i64, _int] + [_p] * 15),
    'hn_composite_bwd': (_int, [_i64, _int] + [_p] * 12),
    'hn_tile_image_bytes': (_i64, [_i64, _int]),
    'hn_tile_pack': (_int, [_i64, _int, _int, _p, _i64, _p, _p]),
    'hn_tile_unpack': (_int, [_i64, _int, _int, _p, _p, _i64, _p]),
    'hn_tile_wgrad': (_int, [C.POINTER(WgradJob), _int, _i64, _p, _p]),
    'hn_scatter_rows': (_int, [_i64, _int, _p, _p, _p, _p]),
    'hn_gather_rows': (_int, [_i64, _int, _p, _p, _p, _p]),
    'hn_vol_decoder_workspace_bytes': (_i64, [_int, _int, _int, _p]),
    'hn_vol_decoder_scratch_bytes': (_i64, [_int, _int, _int, _p]),
    'hn_vol_decoder_fwd': (_int, [_int, _int, _int] + [_p] * 9 + [_i64, _p, _i64, _int, _p]),
    'hn_vol_decoder_bwd': (_int, [_int, _int, _int] + [_p] * 7 + [_i64, _p, _i64] + [_p] * 5 + [_int, _p]),
    'hn_patch_prepare': (_int, [_int, _int] + [_p] * 7),
    'hn_select_patches': (_int, [_int, _int, _int, _int, C.c_double] + [_p] * 22),
    'hn_lpips_head_fwd': (_int, [_int, _int, _int, _p, _i64, _i64, _i64, _p, _p, _p]),
    'hn_lpips_head_bwd': (_int, [_int, _int, _int, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
}


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library once; fail loudly if it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise RuntimeError(
            f'humannerf_b200: {_LIB_PATH} is missing - run `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or ./build.sh).  There is no CPU or PyTorch fallback for the rendering hot path.')
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.hn_version() != HN_ABI_VERSION:
        raise RuntimeError(f'humannerf_b200: ABI version {lib.hn_version()} != {HN_ABI_VERSION}')
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hn_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'humannerf_b200.{what} failed (code {rc}): {msg}')


def launch_count() -> int:
    return int(load().hn_launch_count())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('humannerf_b200: expected a CUDA tensor (there is no CPU path)')
    if not t.is_contiguous():
        raise RuntimeError('humannerf_b200: expected a contiguous tensor')
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
