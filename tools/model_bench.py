#!/usr/bin/env python
"""Per-model forward benchmark on the BASELINE configs (configs/*.yaml): ms per chunk batch, x realtime at the
config's overlap, per-kernel-class breakdown.  Random-init weights, synthetic input, CUDA-event timing."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sesa_audio_separation_b200 as sesa  # noqa: E402
from sesa_audio_separation_b200 import _lib  # noqa: E402

MODELS = {'bs': ('bs_roformer', 'config_bs_roformer_vocals.yaml'), 'mel4': ('mel_band_roformer', 'config_mel_band_roformer_4stem.yaml'),
          'mdx': ('mdx23c', 'config_vocals_mdx23c.yaml')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='mdx', choices=list(MODELS))
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--precision', default='fp32')
    args = ap.parse_args()
    mt, fn = MODELS[args.model]
    model, cfg = sesa.get_model_from_config(mt, os.path.join(ROOT, 'configs', fn))
    model.eval().to('cuda').set_precision(args.precision)
    L = int(cfg.audio.chunk_size)
    x = torch.randn(args.batch, 2, L, device='cuda') * 0.1
    for _ in range(2):
        y = model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        y = model(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    _lib.profile_start()
    model(x)
    prof = _lib.profile_stop()
    ov = int(cfg.inference.num_overlap)
    audio_s = args.batch * (L / ov) / 44100.0
    print(json.dumps({'model': args.model, 'batch': args.batch, 'precision': args.precision, 'ms_per_batch': ms,
                      'ms_per_chunk': ms / args.batch, 'x_realtime_at_overlap': audio_s / (ms / 1e3), 'overlap': ov,
                      'breakdown_ms': {k: [n, round(t, 3)] for k, (n, t) in sorted(prof.items())},
                      'mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == '__main__':
    main()
