#!/usr/bin/env python
"""Micro-benchmark of the tensor-core attention on the two BS-RoFormer C2 geometries (time: 62*B sequences of
801 frames; band: 801*B sequences of 62 bands; 8 heads x 64).  CUDA-event timing."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sesa_audio_separation_b200 import _lib, tc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=2)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--nsplit', type=int, default=3)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    _lib.require_cuda()
    dev = 'cuda'
    B, T, F, H, dh = args.chunks, 801, 62, 8, 64
    inner = H * dh
    M = B * T * F
    g = torch.Generator(device=dev).manual_seed(0)
    planes = (torch.randn(2, M, 3 * inner, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    planes[1] *= 2.0 ** -9
    gates = torch.randn(M, 8, device=dev, generator=g)
    out = tc.alloc_planes(M, inner, dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    geo = {'time': ((B * F, T, F, T * F, 1, F), 4 * T * T * dh * H * B * F), 'band': ((B * T, F, 1, F, 0, 1), 4 * F * F * dh * H * B * T)}
    for name, (a, flops) in geo.items():
        if args.only and name not in args.only.split(','):
            continue
        def run():
            _lib.call('sesa_attention_tc', P(planes), planes.shape[-1], planes.stride(0), P(gates), 8, P(out),
                      out.shape[-1], out.stride(0), H, dh, *a, T if name == 'band' else 0, args.nsplit, 2, st)
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(f'{name:5s} B={B} nsplit={args.nsplit}: {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s algorithmic')


if __name__ == '__main__':
    main()
