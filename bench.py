#!/usr/bin/env python
"""bench.py — x-realtime of the chunked separation hot path (BASELINE.json metric) on N B200s.

Workload (BASELINE.json configs[1]): BS-RoFormer vocals (dim 512, depth 12, 62 bands, 8x64 heads,
hop 441), chunk 352800, overlap 4, on a 3-min synthetic 44.1 kHz stereo mix (96 chunks per track),
random-init weights.  One "step" = one full demix of one track per GPU (tracks are sharded across
ranks, no data-path collective => weak scaling).  `value` = seconds of audio separated per second with
the mix resident in HBM; `e2e` = the same through the public demix() call with HOST buffers (H2D of the
mix and D2H of the stems inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
MODEL_CFG = dict(dim=512, depth=12, stereo=True, num_stems=1, time_transformer_depth=1,
                 freq_transformer_depth=1, dim_head=64, heads=8, stft_n_fft=2048, stft_hop_length=441,
                 stft_win_length=2048, mask_estimator_depth=2, mlp_expansion_factor=4)
CHUNK, OVERLAP, CFG_BATCH = 352800, 4, 1


def synth_track(seconds, seed):
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * SR)
    x = 0.1 * torch.randn(2, n, generator=g)
    t = torch.arange(n, dtype=torch.float64) / SR
    for c in range(2):
        for _ in range(3):
            f = 50.0 + 4000.0 * torch.rand(1, generator=g).item()
            x[c] += (0.4 * torch.sin(2 * np.pi * f * t)).float()
    return (x * (0.9 / x.abs().max())).contiguous()


def flops_per_chunk(cfg, L):
    """Algorithmic FLOPs of one chunk forward: GEMMs + attention."""
    D, H, dh = cfg['dim'], cfg['heads'], cfg['dim_head']
    inner = H * dh
    T = 1 + L // cfg['stft_hop_length']
    fpb = (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)
    dins = [4 * f for f in fpb]
    nb = len(dins)
    M = T * nb
    per_layer = 2 * M * (D * (3 * inner + H) + inner * D + 2 * D * 4 * D)
    n_layers = cfg['depth'] * (cfg['time_transformer_depth'] + cfg['freq_transformer_depth'])
    hid = D * cfg['mlp_expansion_factor']
    band = 2 * T * sum(dins) * D
    mask = cfg['num_stems'] * 2 * T * sum(D * hid + hid * 2 * d for d in dins)
    gemm = n_layers * per_layer + band + mask
    att = cfg['depth'] * (cfg['time_transformer_depth'] * nb * 4 * T * T * dh * H +
                          cfg['freq_transformer_depth'] * T * 4 * nb * nb * dh * H)
    return gemm, att


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def secondary_rooflines(prof, n_chunks, att_f, tf_peak, hbm_peak, seconds, mma_mult):
    """Achieved rate of every other kernel class against the roofline that bounds it (SURVEY 8d byte/FLOP figures)."""
    T, F, C, N, L = 1 + CHUNK // 441, 1025, 2, 1, CHUNK
    per_chunk_bytes = {
        'stft': 4 * C * L + 8 * C * F * T,                               # audio read + spectrogram write
        'mask_istft': 8 * C * F * T + 8 * N * C * F * T + 4 * N * C * L,  # spec + mask read, chunk output write
        'framing': 2 * 4 * C * L,                                         # chunk gather (read + write)
    }
    out = {}
    for k, b in per_chunk_bytes.items():
        n, ms = prof.get(k, (0, 0.0))
        if ms > 0:
            gbs = b * n_chunks / 1e9 / (ms / 1e3)
            out[k] = {'bound': 'hbm', 'ms_per_step': round(ms, 3), 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                      'frac': gbs / hbm_peak}
    n, ms = prof.get('overlap_add', (0, 0.0))
    if ms > 0:
        b = 4 * N * C * L * n_chunks + 4 * N * C * int(seconds * SR)
        out['overlap_add'] = {'bound': 'hbm', 'ms_per_step': round(ms, 3), 'achieved': b / 1e9 / (ms / 1e3), 'peak': hbm_peak,
                              'unit': 'GB/s', 'frac': b / 1e9 / (ms / 1e3) / hbm_peak}
    n, ms = prof.get('attention', (0, 0.0))
    if ms > 0:
        tf = att_f * n_chunks / 1e12 / (ms / 1e3)
        out['attention'] = {'bound': 'tensor', 'ms_per_step': round(ms, 3), 'achieved': tf, 'peak': tf_peak, 'unit': 'TFLOP/s',
                            'frac': tf / tf_peak, 'mma_frac_of_peak': tf * mma_mult / tf_peak}
    return out


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1397.8), d.get('hbm_gbs', 6552.0), 'measured (MEASURED_PEAKS.json, sustained bf16)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


def cpu_baseline(state_dict, steps, warmup):
    """The CPU oracle (restatement of the reference's demix+forward, pinned to it by tests/golden) timed on
    this host's cores on a BOUNDED sample: one chunk forward per step (of the 96 the track needs)."""
    from oracle import roformer as orof
    torch.set_num_threads(os.cpu_count())
    sd = {k: v.detach().cpu().float() for k, v in state_dict.items()}
    x = synth_track(CHUNK / SR, 99)[None, :, :CHUNK]
    with torch.inference_mode():
        for _ in range(warmup):
            orof.bs_roformer_forward(sd, MODEL_CFG, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            orof.bs_roformer_forward(sd, MODEL_CFG, x)
        dt = (time.perf_counter() - t0) / steps
    audio_per_chunk = (CHUNK // OVERLAP) / SR
    return dict(value=audio_per_chunk / dt, unit='x realtime', cores=os.cpu_count(), kind='port',
                sample=f'{steps} chunk forward(s) of 96 (352800 samples, batch 1, fp32, {os.cpu_count()} threads): '
                       f'{dt:.2f} s per chunk; x realtime = audio advanced per chunk (2.0 s at overlap 4) / chunk time',
                s_per_chunk=dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'bf16', 'fp32_simt'])
    ap.add_argument('--seconds', type=float, default=180.0)
    ap.add_argument('--engine-batch', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-rates', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    base_cfg = {'workload': 'BS-RoFormer vocals dim512 depth12 62 bands, 3-min 44.1 kHz stereo track, chunk 352800 '
                            'overlap 4 (96 chunks/track), one track per GPU', 'track_seconds': args.seconds,
                'chunk_size': CHUNK, 'num_overlap': OVERLAP, 'config_batch_size': CFG_BATCH,
                'parallelism': f'track-sharded x{world}', 'l2': 'working set >> L2 (>=1 GB of activations per chunk batch)'}

    if args.impl == 'reference':
        if rank != 0:
            return
        import sesa_audio_separation_b200 as sesa
        model = sesa.BSRoformer(**MODEL_CFG, seed=0)
        cb = cpu_baseline(model.state_dict(), max(1, args.steps), 1 if args.warmup > 0 else 0)
        line = {'impl': 'reference', 'metric': 'seconds of audio separated per second (x realtime), BS-RoFormer vocals',
                'value': cb['value'], 'unit': 'x realtime', 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': cb['s_per_chunk'] * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': dict(base_cfg, note='reference CPU path (oracle port of utils.demix + BSRoformer.forward; the '
                               'reference package itself cannot be installed/imported on the GPU box: absent third-party deps)'),
                'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                'e2e': {'value': cb['value'], 'unit': 'x realtime', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200 import _lib
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    model = sesa.BSRoformer(**MODEL_CFG, seed=0).eval().to(dev)
    if hasattr(model, 'set_precision'):
        model.set_precision(args.precision)
    config = sesa.ConfigDict(dict(audio=dict(chunk_size=CHUNK, sample_rate=SR),
                                  inference=dict(num_overlap=OVERLAP, batch_size=CFG_BATCH),
                                  training=dict(instruments=['vocals', 'other'], target_instrument='vocals')))
    mix_host = synth_track(args.seconds, 1234 + rank).pin_memory()
    mix_dev = mix_host.to(dev)
    eng = sesa.DemixEngine(config, model, dev, engine_batch=args.engine_batch)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        eng.run(mix_dev, to_host=False)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.run(mix_dev, to_host=False)
    e1.record()
    barrier()
    launches = _lib.LAUNCHES - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- end to end through the public API with host buffers
    res = sesa.demix(config, model, mix_host.numpy(), dev, 'bs_roformer', engine_batch=args.engine_batch)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = sesa.demix(config, model, mix_host.numpy(), dev, 'bs_roformer', engine_batch=args.engine_batch)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    out_bytes = sum(v.nbytes for v in res.values())

    # ---- per-kernel-class device time of one more step (CUDA events around every launch)
    _lib.profile_start()
    eng.run(mix_dev, to_host=False)
    prof = _lib.profile_stop()

    if rank == 0:
        audio_s = args.seconds * world * args.steps
        n_chunks = eng.plan.n_chunks
        gemm_f, att_f = flops_per_chunk(MODEL_CFG, CHUNK)
        tf_peak, hbm_peak, peak_src = peaks()
        tc_mode = args.precision != 'fp32_simt'
        gemm_cls = 'gemm_tc' if tc_mode else 'gemm_simt'
        gemm_n, gemm_ms = prof.get(gemm_cls, (0, 0.0))
        breakdown = {k: {'launches': n, 'ms': round(t, 3)} for k, (n, t) in sorted(prof.items())}
        # the tensor-core class runs every GEMM of the forward, the band-split Linears included
        cls_f = gemm_f
        achieved = (cls_f * n_chunks / 1e12) / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
        mma_mult = 3 if args.precision == 'fp32' else 1
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r1_gemm_tc_traffic.json')
        if tc_mode and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get('dram_bytes_per_launch_avg')
        line = {
            'metric': 'seconds of audio separated per second (x realtime), BS-RoFormer vocals',
            'value': audio_s / (ms_total / 1e3), 'unit': 'x realtime', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'bf16x3 (split-bf16 tensor-core products, fp32 accumulate: fp32-parity mode)',
                      'bf16': 'bf16 (fp32 accumulate)', 'fp32_simt': 'f32'}[args.precision], 'data': 'synthetic',
            'config': dict(base_cfg, engine_batch=args.engine_batch, precision=args.precision),
            'e2e': {'value': audio_s / float(t_e2e.item()), 'unit': 'x realtime',
                    'h2d_bytes_per_step': int(mix_host.numel() * 4), 'd2h_bytes_per_step': int(out_bytes)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'kernel': ('gemm_tc_kernel<256,NSPLIT,flavour> (tcgen05 grouped GEMM: to_qkv / to_out / FeedForward / '
                                    'MaskEstimator launches)') if tc_mode else 'gemm_simt_kernel', 'bound': 'tensor',
                         'achieved': achieved, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf_peak,
                         'traffic': traffic, 'peak_source': peak_src, 'launches_per_step': gemm_n,
                         'algorithmic_tflop_per_chunk': cls_f / 1e12,
                         'mma_tflops_issued': achieved * mma_mult,
                         'mma_frac_of_peak': achieved * mma_mult / tf_peak,
                         'note': ('fp32-parity mode issues 3 bf16 MMAs per algorithmic product (hi.hi + hi.lo + lo.hi), so '
                                  'frac (algorithmic) is bounded by 1/3; mma_frac_of_peak is the tensor-pipe view')
                                 if mma_mult == 3 else '',
                         'share_of_step': gemm_ms / max(1e-9, sum(t for _, t in prof.values()))},
            'breakdown_ms_per_step': breakdown, 'attention_tflop_per_chunk': att_f / 1e12,
            'kernels': secondary_rooflines(prof, n_chunks, att_f, tf_peak, hbm_peak, args.seconds, mma_mult),
        }
        if world == 1 and tc_mode and not args.no_kernel_rates:
            # the HBM-bound kernels again, each alone at the 4-chunk launch shape with L2 flushed between repetitions: the
            # per-launch event timing of the profiling pass above carries ~10 us of bracketing per launch, which matters
            # for 10-100 us kernels
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import hbm_bench
            iso = hbm_bench.measure(chunks=args.engine_batch, reps=5, seconds=args.seconds)
            for k, v in iso.items():
                line['kernels'].setdefault(k, {'bound': 'hbm', 'peak': hbm_peak, 'unit': 'GB/s'})
                line['kernels'][k].update({'isolated_us_per_launch': round(v['us'], 1), 'isolated_achieved': v['gbs'],
                                           'isolated_frac': v['gbs'] / hbm_peak})
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(model.state_dict(), 1, 0)
            line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
