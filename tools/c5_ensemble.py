#!/usr/bin/env python
"""BASELINE config 5: avg_wave ensemble of BS-RoFormer + Mel-Band-RoFormer + MDX23C vocals over synthetic 4-min tracks,
track-sharded over the GPUs of one box (independent tracks: no collective), the three estimates averaged ON the device.

  python tools/c5_ensemble.py [--tracks 64] [--seconds 240]          (one GPU)
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_ensemble.py --tracks 64

Prints one JSON line: whole-job x realtime (audio seconds of all tracks / wall time of the slowest rank), H2D + D2H
inside the timed region (host mixes in, ensembled vocals out)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sesa_audio_separation_b200 as sesa  # noqa: E402
from sesa_audio_separation_b200.ensemble import ensemble_tracks  # noqa: E402
from bench import synth_track  # noqa: E402

MEMBERS = [('bs_roformer', 'config_bs_roformer_vocals.yaml'), ('mel_band_roformer', 'config_mel_band_roformer_vocals.yaml'),
           ('mdx23c', 'config_vocals_mdx23c.yaml')]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tracks', type=int, default=8)
    ap.add_argument('--seconds', type=float, default=240.0)
    ap.add_argument('--out-dir', default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    members = []
    for mt, fn in MEMBERS:
        model, cfg = sesa.get_model_from_config(mt, os.path.join(ROOT, 'configs', fn))
        members.append((cfg, model.eval().to(dev)))
    # two distinct synthetic host mixes per rank stand for its share of the tracks (every track is separated in full; only
    # the synthesis is shared); entries of other ranks are skipped by ensemble_tracks' round-robin sharding
    pool = [synth_track(args.seconds, 5000 + 2 * rank + j).numpy() for j in range(2)]
    tracks = [(f'track{i:03d}', pool[(i // world) % 2]) for i in range(args.tracks)]
    ensemble_tracks(members, tracks[rank:rank + 1], dev)                                # warm-up (workspaces, TMA tables)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    done = ensemble_tracks(members, tracks, dev, out_dir=args.out_dir, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'workload': f'avg_wave ensemble of BS-RoFormer + Mel-Band-RoFormer + MDX23C vocals over {args.tracks} synthetic '
                                      f'{args.seconds:.0f}-s tracks, track-sharded over {world} GPU(s), averaged on the device',
                          'value': args.tracks * args.seconds / float(dt.item()), 'unit': 'x realtime (whole job, end to end)',
                          'n_gpus': world, 'seconds_wall': float(dt.item()), 'tracks_on_rank0': len(done)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
