#!/usr/bin/env python
"""torchrun --nproc-per-node W tools/shard_check.py : chunk-range sharding of ONE track over W GPUs (NCCL send/recv
halo exchange, distributed.py) against the unsharded demix on rank 0 — must be bit-identical.  Also times the
sharded run (BASELINE config 3 style: Mel-Band-RoFormer, 4 stems)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sesa_audio_separation_b200 as sesa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seconds', type=float, default=60.0)
    ap.add_argument('--model', default='mel4')
    ap.add_argument('--engine-batch', type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    mt, fn = {'mel4': ('mel_band_roformer', 'config_mel_band_roformer_4stem.yaml'),
              'bs': ('bs_roformer', 'config_bs_roformer_vocals.yaml')}[args.model]
    model, cfg = sesa.get_model_from_config(mt, os.path.join(ROOT, 'configs', fn))   # same seed on every rank
    model.eval().to(dev)
    g = torch.Generator().manual_seed(99)
    mix = (0.1 * torch.randn(2, int(args.seconds * 44100), generator=g)).to(dev)
    eng = sesa.DemixEngine(cfg, model, dev, engine_batch=args.engine_batch, world=world, rank=rank)
    eng.run(mix, to_host=False)                       # warm-up
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    res = eng.run(mix, to_host=False)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        single = sesa.DemixEngine(cfg, model, dev, engine_batch=args.engine_batch)
        single.run(mix, to_host=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = single.run(mix, to_host=False)
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t0
        same = bool(torch.equal(res, ref))
        print(json.dumps({'world': world, 'model': args.model, 'seconds': args.seconds, 'n_chunks': eng.plan.n_chunks,
                          'bit_identical_to_unsharded': same, 'sharded_s': dt, 'unsharded_s': dt1,
                          'x_realtime_sharded': args.seconds / dt, 'x_realtime_unsharded': args.seconds / dt1,
                          'max_abs_diff': float((res - ref).abs().max())}))
        assert same
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
