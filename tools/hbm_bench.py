#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound kernels of the hot path at the BASELINE C2 shapes (BS-RoFormer, n_fft 2048,
hop 441, chunk 352 800, stereo): STFT, fused mask+iSTFT, chunk framing and the streamed demix overlap-add.  CUDA-event timing on
the launching stream; L2 is flushed between repetitions by writing a 512 MB buffer.  Bytes are the ALGORITHMIC bytes of
SURVEY 8d (each tensor read or written once).  Used for the ncu captures under profiles/."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sesa_audio_separation_b200 import _lib  # noqa: E402
from sesa_audio_separation_b200.plan import make_plan, windowing_array  # noqa: E402
from sesa_audio_separation_b200.roformer import _istft_envelope, _twiddle  # noqa: E402

HBM_PEAK = 6552.0


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def timed(fn, reps, flush):
    st = torch.cuda.current_stream()
    ms = []
    for _ in range(reps + 2):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        b.synchronize()
        ms.append(a.elapsed_time(b))
    ms = sorted(ms[2:])
    return ms[len(ms) // 2]


def measure(chunks=4, reps=7, stems=1, seconds=180.0, only='', verbose=False):
    """-> {kernel: {'us', 'bytes', 'gbs', 'frac'}} for stft / mask_istft / overlap_add / framing."""
    _lib.require_cuda()
    _lib.load()
    dev = 'cuda'
    S = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    B, C, L, N, hop, NS = chunks, 2, 352800, 2048, 441, stems
    T, F = 1 + L // hop, N // 2 + 1
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(128 * 1024 * 1024, device=dev)
    win = torch.hann_window(N).to(dev)
    tw = _twiddle(N).to(dev)
    env = _istft_envelope(torch.hann_window(N), N, hop, T, L).to(dev)
    audio = torch.randn(B, C, L, device=dev, generator=g)
    spec = torch.empty(B * T, F, C, 2, device=dev)
    mask = torch.randn(NS, B * T, F * C * 2, device=dev, generator=g)
    out = torch.empty(B, NS, C, L, device=dev)
    res = {}
    sel = set(only.split(',')) if only else None

    def report(name, ms, nbytes):
        gbs = nbytes / 1e9 / (ms / 1e3)
        res[name] = {'us': ms * 1e3, 'bytes': nbytes, 'gbs': gbs, 'frac': gbs / HBM_PEAK}
        if verbose:
            print(f'{name:12s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / HBM_PEAK * 100:5.1f} % of {HBM_PEAK:.0f}', flush=True)

    if sel is None or 'stft' in sel:
        ms = timed(lambda: _lib.call('sesa_stft', P(audio), P(spec), P(win), P(tw), B, C, L, N, hop, 0, F, S), reps, flush)
        report('stft', ms, B * (4 * C * L + 8 * C * F * T))
    else:
        _lib.call('sesa_stft', P(audio), P(spec), P(win), P(tw), B, C, L, N, hop, 0, F, S)
    if sel is None or 'istft' in sel:
        ms = timed(lambda: _lib.call('sesa_mask_istft', P(spec), P(mask), None, None, P(out), P(win), P(env), P(tw), B, NS, C, N,
                                     hop, T, L, 0, 0, S), reps, flush)
        report('mask_istft', ms, B * (8 * C * F * T + NS * (8 * C * F * T + 4 * C * L)))
    # demix overlap-add over a whole track
    length = int(seconds * 44100)
    plan = make_plan(length, L, 4, 1)
    starts = torch.tensor(plan.starts, dtype=torch.int64).to(dev)
    lens = torch.tensor(plan.lens, dtype=torch.int64).to(dev)
    modes = torch.tensor(plan.modes, dtype=torch.int32).to(dev)
    kinds = torch.tensor(plan.kinds, dtype=torch.int32).to(dev)
    if sel is None or 'ola' in sel:
        # the streamed overlap-add of the product path: one engine batch (B chunks, mid-track) folded into the running sums
        y = torch.randn(B, NS, C, L, device=dev, generator=g)
        result = torch.empty(NS * C, (length + 3) // 4 * 4, device=dev)
        partial = torch.zeros(NS * C, (plan.padded + 3) // 4 * 4 + 4, device=dev)
        window = windowing_array(L, plan.fade).to(dev)
        crop = plan.border if plan.pad else 0
        span = -(-L // plan.step)
        k0 = plan.n_chunks // 2
        ms = timed(lambda: _lib.call('sesa_overlap_accumulate', P(y), k0, B, P(starts), P(lens), P(kinds), plan.n_chunks, plan.step,
                                     L, plan.fade, P(window), NS, C, plan.padded, k0, k0 + B + span - 1, P(partial),
                                     partial.shape[1], 0, crop, length, P(result), result.shape[1], 0, length, S), reps, flush)
        # chunk outputs read once, finished regions written once, the (span-1) open regions read and written back
        report('overlap_add', ms, 4 * NS * C * (B * L + B * plan.step + 2 * (span - 1) * plan.step))
        del y, result, partial
    if sel is None or 'frame' in sel:
        padded = torch.randn(C, plan.padded, device=dev, generator=g)
        chunks_t = torch.empty(B, C, L, device=dev)
        ms = timed(lambda: _lib.call('sesa_frame_chunks', P(padded), plan.padded, C, P(starts), P(lens), P(modes), B, L,
                                     P(chunks_t), S), reps, flush)
        report('framing', ms, 2 * 4 * B * C * L)
    torch.cuda.synchronize()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=4)
    ap.add_argument('--reps', type=int, default=7)
    ap.add_argument('--stems', type=int, default=1)
    ap.add_argument('--seconds', type=float, default=180.0)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    measure(args.chunks, args.reps, args.stems, args.seconds, args.only, verbose=True)


if __name__ == '__main__':
    main()
