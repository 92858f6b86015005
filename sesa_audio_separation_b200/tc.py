"""Host helpers of the tensor-core (tcgen05) kernels: bf16 hi/lo plane buffers and grouped-GEMM tables.

A "plane pair" is a bf16 tensor of shape [2, rows, ld]: plane 0 = bf16(x), plane 1 = bf16(x - plane0).
The tables hold the TMA tensor maps of sesa_gemm_tc (encoded on the host by the C library) and are
built once per workspace; nothing here does arithmetic.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import TC_PROBLEM_DTYPE, call


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def round8(n):
    return (int(n) + 7) // 8 * 8


def alloc_planes(rows, cols, device, planes=2):
    """bf16 [planes, rows, round8(cols)] zero-initialised (padding columns must stay zero)."""
    return torch.zeros(planes, int(rows), round8(cols), device=device, dtype=torch.bfloat16)


def split_weight(w):
    """fp32 weight [N, K] on the device -> bf16 planes [2, N, round8(K)] (through the C library)."""
    w = w.contiguous().float()
    n, k = w.shape
    out = torch.empty(2, n, round8(k), device=w.device, dtype=torch.bfloat16)
    call('sesa_split_weight', _ptr(w), n, k, _ptr(out), round8(k), _stream())
    return out


def prep_rows(x, rows, dim, ldx, planes, normalize, gate_w=None, gate_b=None, gates=None, ldg=0, rowinv=None,
              out_planes=2, ss_slots=1):
    """x fp32 [rows, ldx] -> bf16 planes (+ gate logits, + inverse row norms).  normalize: False/0 raw, True/1 unit-norm
    rows, 2 = raw planes and rowinv[rows, ss_slots] = (sum of squares, 0, ...) for a GEMM's fused row scale."""
    n_gates = 0 if gate_w is None else int(gate_w.shape[0])
    call('sesa_prep_rows', _ptr(x), ldx, rows, dim, int(normalize),
         _ptr(planes) if planes is not None else None,
         planes.shape[-1] if planes is not None else 0,
         planes.stride(0) if planes is not None else 0, out_planes,
         _ptr(gate_w) if gate_w is not None else None, _ptr(gate_b) if gate_b is not None else None, n_gates,
         _ptr(gates) if gates is not None else None, ldg, _ptr(rowinv) if rowinv is not None else None, ss_slots,
         _stream())


# CTA pairs (tcgen05 cta_group::2, 256 x 256 pair tiles) are the default for the 256-wide tile: they halve the W-tile
# traffic and reach 1.2-1.55 PFLOP/s of issued MMA where single-CTA tiles stall on L2 (profiles/r1_microbench.md)
DEFAULT_CTA_GROUP = 2
SS_SLOTS_PER_BLOCK = 2   # SESA_TC_SS_SLOTS_PER_BLOCK of include/sesa_b200.h


class TcGemmTable:
    """Device table of one (grouped) sesa_gemm_tc launch.

    problems: list of dicts with keys A (bf16 planes tensor or (ptr, ld, plane_stride)), W (same), M, N, K and
    optional bias, rowscale, C (ptr, ldc), P (ptr, ldp, plane_stride).
    """

    def __init__(self, problems, device, block_n=256, cta_group=None):
        if cta_group is None:
            cta_group = DEFAULT_CTA_GROUP if block_n == 256 else 1
        n = len(problems)
        arr = np.zeros(n, dtype=TC_PROBLEM_DTYPE)
        for i, p in enumerate(problems):
            a_ptr, lda, a_pl = p['A']
            w_ptr, ldw, w_pl = p['W']
            c_ptr, ldc = p.get('C') or (0, 0)
            p_ptr, ldp, p_pl = p.get('P') or (0, 0, 0)
            rec = arr[i]
            for name, val in (('A', a_ptr), ('W', w_ptr), ('bias', p.get('bias') or 0), ('rowscale', p.get('rowscale') or 0),
                              ('C', c_ptr), ('P', p_ptr), ('lda', lda), ('a_plane', a_pl), ('ldw', ldw), ('w_plane', w_pl),
                              ('ldc', ldc), ('ldp', ldp), ('p_plane', p_pl), ('M', p['M']), ('N', p['N']), ('K', p['K'])):
                rec[name] = val
            conv = p.get('conv')
            if conv is not None:      # implicit-GEMM convolution: dict(cin, B, T, F, inT, inF, stride, taps=[(dt, df), ...])
                taps = conv['taps']
                rec['conv_taps'] = len(taps)
                for name in ('cin', 'B', 'T', 'F', 'inT', 'inF', 'stride'):
                    rec['conv_' + name] = conv[name]
                rec['conv_dt'][:len(taps)] = [t[0] for t in taps]
                rec['conv_df'][:len(taps)] = [t[1] for t in taps]
            for name in ('rowss', 'ss_out', 'ss_slots', 'p_cols', 'c_col0', 'ss_ld'):
                if p.get(name):
                    rec[name] = p[name]
            rm = p.get('row_map')
            if rm is not None:        # (F_in, dt, df): 2x up-sampling scatter
                rec['row_map'], rec['rm_F'], rec['rm_dt'], rec['rm_df'] = 1, rm[0], rm[1], rm[2]
        lib = _lib.load()
        nbytes = int(lib.sesa_gemm_tc_table_bytes(n))
        host = np.zeros(nbytes, dtype=np.uint8)
        tiles = ctypes.c_int(0)
        _lib.check(lib.sesa_gemm_tc_build(arr.ctypes.data_as(ctypes.c_void_p), n, block_n, cta_group,
                                          host.ctypes.data_as(ctypes.c_void_p), ctypes.byref(tiles)))
        self.n, self.block_n, self.cta_group, self.tiles = n, block_n, cta_group, int(tiles.value)
        self.dev = torch.from_numpy(host).to(device)
        self.flops = sum(2 * int(p['M']) * int(p['N']) * int(p['K']) for p in problems)

    def run(self, ep, nsplit=3, out_planes=2):
        call('sesa_gemm_tc', _ptr(self.dev), self.n, self.tiles, self.block_n, self.cta_group, nsplit, out_planes, ctypes.byref(ep),
             _stream())


def planes_arg(t, col_offset=0):
    """(ptr, ld, plane_stride) of a [planes, rows, ld] bf16 tensor, optionally starting at a column."""
    return (t.data_ptr() + 2 * int(col_offset), t.shape[-1], t.stride(0))
