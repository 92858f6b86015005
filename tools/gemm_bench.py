#!/usr/bin/env python
"""Micro-benchmark of the tensor-core GEMM on the four transformer-layer shapes of BS-RoFormer C2
(M = tokens of `--chunks` chunks).  CUDA-event timing, L2 flushed between repetitions by the working set
(A/C planes are >> 126 MB).  Used for ncu captures of a single kernel (profiles/)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sesa_audio_separation_b200 import _lib, tc  # noqa: E402
from sesa_audio_separation_b200._lib import GemmEpilogue  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=2)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--nsplit', type=int, default=3)
    ap.add_argument('--block-n', type=int, default=256)
    ap.add_argument('--only', default='')
    ap.add_argument('--cta-group', type=int, default=2)
    args = ap.parse_args()
    _lib.require_cuda()
    dev = 'cuda'
    M, D = 49662 * args.chunks, 512
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(M, D, device=dev, generator=g)
    xp = tc.alloc_planes(M, D, dev)
    tc.prep_rows(x, M, D, D, xp, True)
    hp = tc.alloc_planes(M, 4 * D, dev)
    qkv = torch.zeros(M, 1544, device=dev)
    w = {k: tc.split_weight(torch.randn(n, kk, device=dev, generator=g) / kk ** 0.5)
         for k, (n, kk) in dict(qkv=(1536, 512), out=(512, 512), ff1=(2048, 512), ff2=(512, 2048)).items()}
    w['qkvg'] = tc.split_weight(torch.randn(1544, 512, device=dev, generator=g) / 512 ** 0.5)
    gates = torch.zeros(M, 8, device=dev)
    qkvp = tc.alloc_planes(M, 1536, dev)
    xp2 = tc.alloc_planes(M, D, dev)
    ss = torch.ones(M, 4, device=dev)
    bg = torch.zeros(1544, device=dev)
    b1 = torch.randn(2048, device=dev, generator=g)
    b2 = torch.randn(512, device=dev, generator=g)
    ang = torch.einsum('i,j->ij', torch.arange(801, dtype=torch.float32), 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64)))
    rot = torch.stack([ang.cos(), ang.sin()], -1).reshape(801, -1, 4).permute(1, 0, 2).contiguous().to(dev)   # quad-major
    rot_band = torch.stack([ang[:62].cos(), ang[:62].sin()], -1).reshape(62, -1, 4).permute(1, 0, 2).contiguous().to(dev)

    def table(A, W, N, K, **kw):
        return tc.TcGemmTable([dict(A=tc.planes_arg(A), W=tc.planes_arg(W), M=M, N=N, K=K, **kw)], dev, block_n=args.block_n, cta_group=args.cta_group)
    cases = {
        'qkv': (table(xp, w['qkv'], 1536, 512, C=(qkv.data_ptr(), 1544)),
                GemmEpilogue(0, 0, 0, 0, 1024, 64, 62, 801, rot.data_ptr()), 2 * M * 1536 * 512),
        'out': (table(xp, w['out'], 512, 512, C=(x.data_ptr(), 512)), GemmEpilogue(0, 0, 1, 0, 0, 0, 1, 1, None), 2 * M * 512 * 512),
        'ff1': (table(xp, w['ff1'], 2048, 512, bias=b1.data_ptr(), P=tc.planes_arg(hp)),
                GemmEpilogue(0, 1, 0, 0, 0, 0, 1, 1, None), 2 * M * 2048 * 512),
        'ff2': (table(hp, w['ff2'], 512, 2048, bias=b2.data_ptr(), C=(x.data_ptr(), 512)),
                GemmEpilogue(0, 0, 1, 0, 0, 0, 1, 1, None), 2 * M * 512 * 2048),
        'qkvg': (table(xp, w['qkvg'], 1544, 512, C=(gates.data_ptr(), 8), P=tc.planes_arg(qkvp), bias=bg.data_ptr(),
                       rowss=ss.data_ptr(), ss_slots=4, p_cols=1536, c_col0=1536),
                 GemmEpilogue(0, 0, 0, 0, 1024, 64, 62, 801, rot.data_ptr()), 2 * M * 1544 * 512),
        'qkvg_band': (table(xp, w['qkvg'], 1544, 512, C=(gates.data_ptr(), 8), P=tc.planes_arg(qkvp), bias=bg.data_ptr(),
                            rowss=ss.data_ptr(), ss_slots=4, p_cols=1536, c_col0=1536),
                      GemmEpilogue(0, 0, 0, 0, 1024, 64, 1, 62, rot_band.data_ptr()), 2 * M * 1544 * 512),
        'qkvp': (table(xp, w['qkv'], 1536, 512, P=tc.planes_arg(qkvp), rowss=ss.data_ptr(), ss_slots=4),
                 GemmEpilogue(0, 0, 0, 0, 1024, 64, 62, 801, rot.data_ptr()), 2 * M * 1536 * 512),
        'outp': (table(xp, w['out'], 512, 512, C=(x.data_ptr(), 512), P=tc.planes_arg(xp2), ss_out=ss.data_ptr()),
                 GemmEpilogue(0, 0, 1, 0, 0, 0, 1, 1, None), 2 * M * 512 * 512),
        'ff2p': (table(hp, w['ff2'], 512, 2048, bias=b2.data_ptr(), C=(x.data_ptr(), 512), P=tc.planes_arg(xp2), ss_out=ss.data_ptr()),
                 GemmEpilogue(0, 0, 1, 0, 0, 0, 1, 1, None), 2 * M * 512 * 2048),
        'plain': (table(xp, w['ff1'], 2048, 512, C=(hp.data_ptr(), 2048)), GemmEpilogue(0, 0, 0, 0, 0, 0, 1, 1, None), 2 * M * 2048 * 512),
    }
    for name, (tab, ep, flops) in cases.items():
        if args.only and name not in args.only.split(','):
            continue
        for _ in range(2):
            tab.run(ep, nsplit=args.nsplit)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            tab.run(ep, nsplit=args.nsplit)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(f'{name:9s} M={M} nsplit={args.nsplit} BN={args.block_n} CG={args.cta_group}: {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s algorithmic '
              f'({flops * (3 if args.nsplit == 3 else 1) / ms / 1e9:8.1f} MMA TFLOP/s)')


if __name__ == '__main__':
    main()
