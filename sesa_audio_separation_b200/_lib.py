"""ctypes binding of the C-ABI in include/sesa_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is NO
fallback: a missing library or a missing CUDA device raises immediately.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsesa_b200.so')

ACT_NONE, ACT_GELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3


class SesaError(RuntimeError):
    pass


class GemmEpilogue(ctypes.Structure):
    _fields_ = [('rownorm', c_int32), ('act', c_int32), ('residual', c_int32), ('glu', c_int32),
                ('rot_cols', c_int32), ('rot_dim', c_int32), ('pos_div', c_int32), ('pos_mod', c_int32),
                ('rot', c_void_p)]


# numpy mirror of struct sesa_gemm_group (64 bytes)
GEMM_GROUP_DTYPE = np.dtype([('A', '<u8'), ('W', '<u8'), ('bias', '<u8'), ('C', '<u8'),
                             ('M', '<i4'), ('N', '<i4'), ('K', '<i4'), ('_pad', '<i4'),
                             ('lda', '<i8'), ('ldw', '<i8'), ('ldc', '<i8')])
assert GEMM_GROUP_DTYPE.itemsize == 72

# numpy mirror of struct sesa_tc_problem (272 bytes)
TC_PROBLEM_DTYPE = np.dtype([('A', '<u8'), ('W', '<u8'), ('bias', '<u8'), ('rowscale', '<u8'), ('C', '<u8'),
                             ('P', '<u8'), ('lda', '<i8'), ('a_plane', '<i8'), ('ldw', '<i8'), ('w_plane', '<i8'),
                             ('ldc', '<i8'), ('ldp', '<i8'), ('p_plane', '<i8'),
                             ('M', '<i4'), ('N', '<i4'), ('K', '<i4'), ('_pad', '<i4'),
                             ('conv_taps', '<i4'), ('conv_cin', '<i4'), ('conv_B', '<i4'), ('conv_T', '<i4'),
                             ('conv_F', '<i4'), ('conv_inT', '<i4'), ('conv_inF', '<i4'), ('conv_stride', '<i4'),
                             ('conv_dt', '<i4', (9,)), ('conv_df', '<i4', (9,)),
                             ('row_map', '<i4'), ('rm_F', '<i4'), ('rm_dt', '<i4'), ('rm_df', '<i4'),
                             ('rowss', '<u8'), ('ss_out', '<u8'), ('ss_slots', '<i4'), ('p_cols', '<i4'),
                             ('c_col0', '<i4'), ('ss_ld', '<i4')])
assert TC_PROBLEM_DTYPE.itemsize == 272

_SIGS = {
    'sesa_abi_version': (c_int, []),
    'sesa_last_error': (c_char_p, []),
    'sesa_device_info': (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    'sesa_pad_reflect': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p]),
    'sesa_frame_chunks': (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64,
                                  c_void_p, c_void_p]),
    'sesa_stft': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_int, c_int,
                          c_int, c_void_p]),
    'sesa_mask_istft': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    'sesa_gemm_simt': (c_int, [c_void_p, c_int, c_int, c_int, POINTER(GemmEpilogue), c_void_p]),
    'sesa_attention_simt': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int64, c_int64, c_int64, c_void_p]),
    'sesa_gemm_tc_table_bytes': (c_int64, [c_int]),
    'sesa_gemm_tc_build': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, POINTER(c_int)]),
    'sesa_gemm_tc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(GemmEpilogue), c_void_p]),
    'sesa_attention_tc': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int,
                                  c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    'sesa_prep_rows': (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int64, c_int64, c_int,
                               c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    'sesa_band_prep': (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                               c_void_p]),
    'sesa_split_weight': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    'sesa_rmsnorm': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'sesa_add_inplace': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'sesa_gather_rows': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    'sesa_overlap_add': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int,
                                 c_void_p, c_int, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p]),
    'sesa_instnorm_stats': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p,
                                    ctypes.c_float, c_void_p]),
    'sesa_norm_act_split': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_int, c_void_p, c_int64, c_int64, c_void_p]),
    'sesa_transpose_add': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p]),
    'sesa_transpose_add_stats': (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p,
                                         ctypes.c_float, c_void_p]),
    'sesa_mdx_pack': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'sesa_mdx_final_concat': (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64, c_void_p,
                                      c_int64, c_int64, c_void_p]),
    'sesa_mdx_unpack': (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'sesa_rmsnorm_planes': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]),
    'sesa_overlap_accumulate': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int,
                                        c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_int64, c_int64, c_int64,
                                        c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    'sesa_pad_reflect_slice': (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                       c_void_p]),
    'sesa_tta_variants': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    'sesa_tta_combine': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p]),
    'sesa_ensemble_wave': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def load():
    """Load libsesa_b200.so and bind every symbol of include/sesa_b200.h (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SesaError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                        'at the repo root (nvcc, sm_100a). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sesa_abi_version() != 3:
        raise SesaError('libsesa_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().sesa_last_error()
        raise SesaError(f'sesa_b200 error {status}: {msg.decode() if msg else "?"}')


# ---- launch accounting / per-kernel-class timing (bench.py, profiling); off by default
LAUNCHES = 0
_profile = None   # dict: class -> [ (start_event, end_event), ... ] when enabled
_CLASS = {'sesa_gemm_simt': 'gemm_simt', 'sesa_gemm_tc': 'gemm_tc', 'sesa_prep_rows': 'prep_rows', 'sesa_rmsnorm_planes': 'prep_rows', 'sesa_rmsnorm': 'prep_rows', 'sesa_add_inplace': 'prep_rows', 'sesa_band_prep': 'prep_rows', 'sesa_instnorm_stats': 'norm', 'sesa_norm_act_split': 'norm', 'sesa_transpose_add': 'norm', 'sesa_transpose_add_stats': 'norm', 'sesa_attention_simt': 'attention',
          'sesa_attention_tc': 'attention', 'sesa_stft': 'stft', 'sesa_mask_istft': 'mask_istft',
          'sesa_overlap_add': 'overlap_add', 'sesa_overlap_accumulate': 'overlap_add', 'sesa_pad_reflect_slice': 'framing',
          'sesa_tta_variants': 'tta', 'sesa_tta_combine': 'tta', 'sesa_ensemble_wave': 'ensemble', 'sesa_frame_chunks': 'framing', 'sesa_pad_reflect': 'framing'}


def profile_start():
    global _profile
    _profile = {}


def profile_stop():
    """Returns {class: (n_launches, total_ms)} measured with CUDA events on the launch stream."""
    global _profile
    import torch
    torch.cuda.synchronize()
    out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in _profile.items()}
    _profile = None
    return out


def call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if _profile is None:
        check(getattr(load(), name)(*args))
        return
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(getattr(load(), name)(*args))
    b.record()
    _profile.setdefault(_CLASS.get(name, 'other'), []).append((a, b))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SesaError('sesa_audio_separation_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists')
    load()
