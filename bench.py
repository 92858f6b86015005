#!/usr/bin/env python
"""bench.py — x-realtime of the chunked separation hot path (BASELINE.json metric) on N B200s.

Headline workload (BASELINE.json configs[1], "C2"): BS-RoFormer vocals (dim 512, depth 12, 62 bands, 8x64 heads,
hop 441), chunk 352800, overlap 4, on a 3-min synthetic 44.1 kHz stereo mix (96 chunks per track), random-init weights.
One "step" = one full demix of one track per GPU (tracks are sharded across ranks, no data-path collective => weak
scaling).  `value` = seconds of audio separated per second with the mix resident in HBM; `e2e` = the same through the
public demix() call with HOST buffers (H2D of the mix and D2H of the stems inside the timed region).

The same JSON line also carries
  * `configs`: sub-records for the other single-GPU BASELINE configurations the repo claims — C1 (MDX23C vocals, 30-s
    track, 27 chunks) and C3's model on one GPU (Mel-Band-RoFormer 4-stem, 10-min track) — each with x realtime,
    ms per chunk, the dominant kernel's roofline fraction and the bf16-mode number (N = 1 only);
  * `strong` (N > 1 only): BASELINE configs[2] — ONE 10-min track through Mel-Band-RoFormer 4-stem, chunk-range sharded
    over the N ranks with the NCCL halo exchange (the path's only collective), against the same track on one GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]

`--impl reference` times the reference's own CPU implementation of the path (the oracle port of utils.demix +
BSRoformer.forward with the reference's stock CPU attention, F.scaled_dot_product_attention) on this box's host cores;
a step there is one chunk of the real 96-chunk loop (framing, forward, windowed accumulation).
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
BS_DIM_INPUTS = [4 * f for f in (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)]
MDX23C_TFLOP_PER_CHUNK = 2.434        # SURVEY 8d probe (FlopCounterMode on the reference TFC_TDF_net, conv 2.255 + linear 0.179)


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


def roformer_flops(dim, depth, heads, dim_head, dim_inputs, L, hop, n_mask_linears, stems, expansion=4, sub_depth=(1, 1)):
    """Algorithmic FLOPs of one chunk forward of a band-split RoFormer: (GEMMs, attention)."""
    D, inner = dim, heads * dim_head
    T = 1 + L // hop
    nb = len(dim_inputs)
    M = T * nb
    per_layer = 2 * M * (D * (3 * inner + heads) + inner * D + 2 * D * 4 * D)
    n_layers = depth * (sub_depth[0] + sub_depth[1])
    hid = D * expansion
    band = 2 * T * sum(dim_inputs) * D
    mask = 0
    for d in dim_inputs:
        dims = [D] + [hid] * (n_mask_linears - 1) + [2 * d]
        mask += 2 * T * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    gemm = n_layers * per_layer + band + stems * mask
    att = depth * (sub_depth[0] * nb * 4 * T * T * dim_head * heads + sub_depth[1] * T * 4 * nb * nb * dim_head * heads)
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


def secondary_rooflines(prof, n_chunks, att_f, tf_peak, hbm_peak, seconds, mma_mult, engine_batch):
    """Achieved rate of every other kernel class against the roofline that bounds it (SURVEY 8d byte/FLOP figures)."""
    T, F, C, N, L = 1 + CHUNK // 441, 1025, 2, 1, CHUNK
    step = CHUNK // OVERLAP
    per_chunk_bytes = {
        'stft': 4 * C * L + 8 * C * F * T,                               # audio read + spectrogram write
        'mask_istft': 8 * C * F * T + 8 * N * C * F * T + 4 * N * C * L,  # spec + mask read, chunk output write
        'framing': 2 * 4 * C * L,                                         # chunk gather (read + write)
        # streamed overlap-add: chunk output read, finished region written, open regions read + written per engine batch
        'overlap_add': 4 * N * C * (L + step) + 2 * 4 * N * C * (OVERLAP - 1) * step / engine_batch,
    }
    out = {}
    for k, b in per_chunk_bytes.items():
        n, ms = prof.get(k, (0, 0.0))
        if ms > 0:
            gbs = b * n_chunks / 1e9 / (ms / 1e3)
            out[k] = {'bound': 'hbm', 'ms_per_step': round(ms, 3), 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                      'frac': gbs / hbm_peak}
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


def cpu_reference(state_dict, seconds, timed_chunks, warm_chunks):
    """The reference's CPU path — the oracle port of utils.demix driving BSRoformer.forward (stock CPU attention =
    F.scaled_dot_product_attention, attend.py:56,89-93), pinned to the reference by tests/golden — timed on this host's
    cores on a BOUNDED sample of the same job: the first ``warm_chunks + timed_chunks`` chunks of the real 96-chunk loop
    over the 3-min track (framing, forward, window multiply and accumulation included).  x realtime extrapolates by chunk
    count: track seconds / (96 x mean seconds per chunk), so the border-padding overhead (96 chunks for 90 chunks' worth
    of audio) is inside the number."""
    from oracle import demix as odemix
    from oracle import roformer as orof
    from oracle.demix import demix_schedule
    torch.set_num_threads(os.cpu_count())
    sd = {k: v.detach().cpu().float() for k, v in state_dict.items()}
    mix = synth_track(seconds, 1234).numpy()
    n_total = len(demix_schedule(mix.shape[1], CHUNK, OVERLAP, CFG_BATCH)['chunks'])
    stamps = []

    def model_fn(a):
        stamps.append(time.perf_counter())
        return orof.bs_roformer_forward(sd, MODEL_CFG, a)
    with torch.inference_mode():
        odemix.demix(mix, model_fn, CHUNK, OVERLAP, CFG_BATCH, 1, max_chunks=warm_chunks + timed_chunks)
    stamps.append(time.perf_counter())
    per = np.diff(stamps)[warm_chunks:]
    dt = float(per.mean())
    return dict(value=seconds / (n_total * dt), unit='x realtime', cores=os.cpu_count(), kind='port',
                sample=f'chunks {warm_chunks}..{warm_chunks + len(per) - 1} of the {n_total}-chunk demix loop of the {seconds:.0f}-s track '
                       f'(oracle.demix + BS-RoFormer forward with SDPA, fp32, batch 1, {os.cpu_count()} threads): '
                       f'{dt:.2f} s per chunk; x realtime = {seconds:.0f} s / ({n_total} chunks x s per chunk)',
                s_per_chunk=dt)


# ----------------------------------------------------------------------------------------------------------------------
def time_runs(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def sub_config(name, dev, tf_peak, hbm_peak):
    """One more BASELINE configuration on this GPU: x realtime (device-resident and end to end), ms per chunk, the
    dominant kernel class against its roofline, and the bf16-mode number."""
    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200 import _lib
    spec = {'c1_mdx23c': ('mdx23c', 'config_vocals_mdx23c.yaml', 30.0,
                          'MDX23C TFC-TDF-v3 vocals (112 M parameters), 30-s 44.1 kHz stereo track, chunk 261120 overlap 4 (27 chunks)'),
            'c3_mel4': ('mel_band_roformer', 'config_mel_band_roformer_4stem.yaml', 600.0,
                        'Mel-Band-RoFormer 4-stem (dim 384, depth 6, 60 mel bands, 832.6 M parameters), 10-min 44.1 kHz stereo track, '
                        'chunk 352800 overlap 2, one GPU')}[name]
    mt, fn, seconds, workload = spec
    model, cfg = sesa.get_model_from_config(mt, os.path.join(ROOT, 'configs', fn))
    model.eval().to(dev)
    mix_host = synth_track(seconds, 4321).pin_memory()
    mix_dev = mix_host.to(dev)
    eng = sesa.DemixEngine(cfg, model, dev, engine_batch=4)

    def barrier():
        torch.cuda.synchronize()
    rec = {'workload': workload, 'unit': 'x realtime'}
    ms = time_runs(lambda: eng.run(mix_dev, to_host=False), 2, 1, barrier)
    n_chunks = eng.plan.n_chunks
    rec.update(value=seconds / (ms / 1e3), ms_per_track=ms, n_chunks=n_chunks, ms_per_chunk=ms / n_chunks, precision='fp32')
    res = sesa.demix(cfg, model, mix_host.numpy(), dev, mt)      # warm-up: page-locked result block enters the host cache
    del res
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sesa.demix(cfg, model, mix_host.numpy(), dev, mt)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    rec['e2e'] = {'value': seconds / t_e2e, 'unit': 'x realtime', 'h2d_bytes_per_step': int(mix_host.numel() * 4),
                  'd2h_bytes_per_step': int(sum(v.nbytes for v in res.values()))}
    del res
    _lib.profile_start()
    eng.run(mix_dev, to_host=False)
    prof = _lib.profile_stop()
    rec['breakdown_ms_per_track'] = {k: {'launches': n, 'ms': round(t, 3)} for k, (n, t) in sorted(prof.items())}
    gemm_ms = prof.get('gemm_tc', (0, 0.0))[1]
    L = int(cfg.audio.chunk_size)
    if mt == 'mdx23c':
        gemm_f, att_f = MDX23C_TFLOP_PER_CHUNK * 1e12, 0.0
    else:
        m = model
        gemm_f, att_f = roformer_flops(m.dim, m.depth, m.heads, m.dim_head, list(m.dim_inputs), L, m.hop, m.n_mask_linears,
                                       m.num_stems, m.mlp_expansion_factor, (m.t_depth, m.f_depth))
    if gemm_ms > 0:
        tf = gemm_f * n_chunks / 1e12 / (gemm_ms / 1e3)
        rec['roofline'] = {'kernel': 'gemm_tc_kernel (tcgen05: ' + ('implicit-GEMM convs + TDF Linears' if mt == 'mdx23c' else
                                                                     'transformer / band-split / mask-estimator GEMMs') + ')',
                           'bound': 'tensor', 'achieved': tf, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': tf / tf_peak,
                           'mma_frac_of_peak': 3 * tf / tf_peak, 'algorithmic_tflop_per_chunk': gemm_f / 1e12,
                           'share_of_step': gemm_ms / max(1e-9, sum(t for _, t in prof.values()))}
    att_ms = prof.get('attention', (0, 0.0))[1]
    if att_ms > 0:
        tf = att_f * n_chunks / 1e12 / (att_ms / 1e3)
        rec['attention'] = {'achieved': tf, 'unit': 'TFLOP/s', 'frac': tf / tf_peak, 'mma_frac_of_peak': 3 * tf / tf_peak,
                            'ms': round(att_ms, 3)}
    model.set_precision('bf16')
    ms16 = time_runs(lambda: eng.run(mix_dev, to_host=False), 1, 1, barrier)
    rec['bf16_mode'] = {'value': seconds / (ms16 / 1e3), 'unit': 'x realtime', 'ms_per_chunk': ms16 / n_chunks}
    del eng, model, mix_dev
    torch.cuda.empty_cache()
    return rec


def strong_scaling(dev, rank, world, steps, warmup):
    """BASELINE configs[2]: ONE 10-min track, Mel-Band-RoFormer 4-stem, chunk-range sharded over the ranks with the NCCL
    halo exchange (distributed.py) — against the same track on one GPU, and bit-compared with it."""
    import torch.distributed as dist
    import sesa_audio_separation_b200 as sesa
    seconds = 600.0
    model, cfg = sesa.get_model_from_config('mel_band_roformer', os.path.join(ROOT, 'configs', 'config_mel_band_roformer_4stem.yaml'))
    model.eval().to(dev)                                      # same seed => the same weights on every rank
    mix_host = synth_track(seconds, 777)                      # every rank reads "the same file"
    mix_dev = mix_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    one = sesa.DemixEngine(cfg, model, dev, engine_batch=4)
    ms1 = time_runs(lambda: one.run(mix_dev, to_host=False), 1, 1, barrier)      # all ranks at once: same power envelope
    ref = one.run(mix_dev, to_host=False) if rank == 0 else None
    t_1gpu = max_over_ranks(ms1)
    eng = sesa.DemixEngine(cfg, model, dev, engine_batch=4, world=world, rank=rank)
    msN = max_over_ranks(time_runs(lambda: eng.run(mix_dev, to_host=False), steps, warmup, barrier))
    res = eng.run(mix_dev, to_host=False)
    torch.cuda.synchronize()
    tm = eng.timings()
    halo_ms = max_over_ranks(tm.get('halo_ms', 0.0))
    gather_ms = max_over_ranks(tm.get('gather_ms', 0.0))
    halo_bytes = int(max_over_ranks(tm.get('halo_bytes', 0)))
    same = bool(torch.equal(res, ref)) if rank == 0 else True
    # end to end: every rank uploads only the slice of the (host) mix its chunks read, rank 0 downloads the stems
    eng.run(mix_host.numpy(), to_host=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.run(mix_host.numpy(), to_host=True)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) / steps)
    h2d = int(max_over_ranks(eng.stats.get('h2d_bytes', 0)))
    rec = None
    if rank == 0:
        rec = {'workload': 'Mel-Band-RoFormer 4-stem (dim 384, depth 6, 60 mel bands), ONE 10-min 44.1 kHz stereo track, chunk 352800 '
                           f'overlap 2 ({eng.plan.n_chunks} chunks), contiguous chunk ranges over {world} GPUs, NCCL halo send/recv',
               'scaling': 'strong', 'value': seconds / (msN / 1e3), 'unit': 'x realtime', 'ms_per_track': msN,
               't_1gpu': t_1gpu, 'value_1gpu': seconds / (t_1gpu / 1e3), 'speedup': t_1gpu / msN,
               'efficiency': t_1gpu / (world * msN), 'halo_bytes': halo_bytes, 'halo_ms': halo_ms, 'gather_ms': gather_ms,
               'bit_identical': same, 'steps': steps,
               'e2e': {'value': seconds / t_e2e, 'unit': 'x realtime', 'h2d_bytes_per_step_per_rank_max': h2d,
                       'd2h_bytes_per_step': int(res.numel() * 4)},
               'note': 'halo_ms = device time the compute stream waited for the incoming halo (max over ranks); gather_ms = the '
                       "grouped point-to-point gather of the owned ranges on rank 0; t_1gpu = the same track on each GPU alone "
                       '(all GPUs busy at once, max over ranks)'}
    del eng, one, model
    torch.cuda.empty_cache()
    return rec


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
    ap.add_argument('--no-configs', action='store_true', help='skip the C1 / C3 sub-records (N = 1)')
    ap.add_argument('--no-strong', action='store_true', help='skip the chunk-range-sharded strong-scaling record (N > 1)')
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
        cb = cpu_reference(model.state_dict(), args.seconds, max(3, args.steps), min(args.warmup, 2))
        line = {'impl': 'reference', 'metric': 'seconds of audio separated per second (x realtime), BS-RoFormer vocals',
                'value': cb['value'], 'unit': 'x realtime', 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': cb['s_per_chunk'] * 1e3, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': dict(base_cfg),      # the same keys and values as the b200 arm's `config`
                'note': 'reference CPU path: the oracle port of utils.demix + BSRoformer.forward with the '
                        "reference's stock CPU attention (SDPA); the reference package itself cannot be installed/imported "
                        'on the GPU box (script collection, absent third-party deps); a step = one chunk of the 96-chunk loop',
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
    n_chunks = eng.plan.n_chunks
    state_dict = model.state_dict()
    del eng, model, mix_dev
    torch.cuda.empty_cache()

    strong = None
    if world > 1 and not args.no_strong:
        try:
            strong = strong_scaling(dev, rank, world, steps=max(2, min(args.steps, 5)), warmup=1)
        except Exception as e:          # (a failure every rank hits alike: the weak-scaling line above still goes out)
            strong = {'error': f'{type(e).__name__}: {e}'[:300]}

    if rank == 0:
        audio_s = args.seconds * world * args.steps
        gemm_f, att_f = roformer_flops(MODEL_CFG['dim'], MODEL_CFG['depth'], MODEL_CFG['heads'], MODEL_CFG['dim_head'], BS_DIM_INPUTS,
                                       CHUNK, 441, 2, 1)
        tf_peak, hbm_peak, peak_src = peaks()
        tc_mode = args.precision != 'fp32_simt'
        gemm_cls = 'gemm_tc' if tc_mode else 'gemm_simt'
        gemm_n, gemm_ms = prof.get(gemm_cls, (0, 0.0))
        breakdown = {k: {'launches': n, 'ms': round(t, 3)} for k, (n, t) in sorted(prof.items())}
        # the tensor-core class runs every GEMM of the forward, the band-split Linears included
        cls_f = gemm_f
        achieved = (cls_f * n_chunks / 1e12) / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
        mma_mult = 3 if args.precision == 'fp32' else 1
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'r2_gemm_tc_traffic.json')
        if tc_mode and os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get('dram_bytes_per_launch_avg'), tj.get('source')
        line = {
            'metric': 'seconds of audio separated per second (x realtime), BS-RoFormer vocals',
            'value': audio_s / (ms_total / 1e3), 'unit': 'x realtime', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'bf16x3 (split-bf16 tensor-core products, fp32 accumulate: fp32-parity mode)',
                      'bf16': 'bf16 (fp32 accumulate)', 'fp32_simt': 'f32'}[args.precision], 'data': 'synthetic',
            'config': dict(base_cfg),
            'engine': {'engine_batch': args.engine_batch, 'precision': args.precision},
            'e2e': {'value': audio_s / float(t_e2e.item()), 'unit': 'x realtime',
                    'h2d_bytes_per_step': int(mix_host.numel() * 4), 'd2h_bytes_per_step': int(out_bytes)},
            'gpu_launches': int(launches), 'clocks': clocks,
            'roofline': {'kernel': ('gemm_tc_kernel<256,NSPLIT,flavour> (tcgen05 grouped GEMM: to_qkv / to_out / FeedForward / '
                                    'MaskEstimator launches)') if tc_mode else 'gemm_simt_kernel', 'bound': 'tensor',
                         'achieved': achieved, 'peak': tf_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf_peak,
                         'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src, 'launches_per_step': gemm_n,
                         'algorithmic_tflop_per_chunk': cls_f / 1e12,
                         'mma_tflops_issued': achieved * mma_mult,
                         'mma_frac_of_peak': achieved * mma_mult / tf_peak,
                         'note': ('fp32-parity mode issues 3 bf16 MMAs per algorithmic product (hi.hi + hi.lo + lo.hi), so '
                                  'frac (algorithmic) is bounded by 1/3; mma_frac_of_peak is the tensor-pipe view')
                                 if mma_mult == 3 else '',
                         'share_of_step': gemm_ms / max(1e-9, sum(t for _, t in prof.values()))},
            'breakdown_ms_per_step': breakdown, 'attention_tflop_per_chunk': att_f / 1e12,
            'kernels': secondary_rooflines(prof, n_chunks, att_f, tf_peak, hbm_peak, args.seconds, mma_mult, args.engine_batch),
        }
        if strong is not None:
            line['strong'] = strong
        if world == 1 and tc_mode and not args.no_kernel_rates:
            # the HBM-bound kernels again, each alone at the 4-chunk launch shape with L2 flushed between repetitions: the
            # per-launch event timing of the profiling pass above carries ~10 us of bracketing per launch, which matters
            # for 10-100 us kernels
            sys.path.insert(0, os.path.join(ROOT, 'tools'))
            import hbm_bench
            try:
                iso = hbm_bench.measure(chunks=args.engine_batch, reps=5, seconds=args.seconds)
            except Exception as e:
                iso = {}
                line['kernels']['isolated_error'] = f'{type(e).__name__}: {e}'[:300]
            for k, v in iso.items():
                line['kernels'].setdefault(k, {'bound': 'hbm', 'peak': hbm_peak, 'unit': 'GB/s'})
                line['kernels'][k].update({'isolated_us_per_launch': round(v['us'], 1), 'isolated_achieved': v['gbs'],
                                           'isolated_frac': v['gbs'] / hbm_peak})
            torch.cuda.empty_cache()
            # and at a launch size that saturates the memory system (16 chunks per launch = 4x the product's engine batch):
            # separates what the KERNEL reaches from what a 10-70 us launch can reach at all
            try:
                big = hbm_bench.measure(chunks=16, reps=5, seconds=args.seconds)
            except Exception as e:
                big = {}
                line['kernels']['saturated_error'] = f'{type(e).__name__}: {e}'[:300]
            for k, v in big.items():
                line['kernels'].setdefault(k, {'bound': 'hbm', 'peak': hbm_peak, 'unit': 'GB/s'})
                line['kernels'][k].update({'saturated_us_per_launch': round(v['us'], 1), 'saturated_achieved': v['gbs'],
                                           'saturated_frac': v['gbs'] / hbm_peak, 'saturated_chunks_per_launch': 16})
            torch.cuda.empty_cache()
        if world == 1 and tc_mode and args.precision == 'fp32' and not args.no_configs:
            line['configs'] = {}
            for name in ('c1_mdx23c', 'c3_mel4'):
                try:
                    line['configs'][name] = sub_config(name, dev, tf_peak, hbm_peak)
                except Exception as e:      # a sub-record must never cost the headline line
                    line['configs'][name] = {'error': f'{type(e).__name__}: {e}'[:300]}
                    torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference(state_dict, args.seconds, 2, 1)
                line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
            except Exception as e:
                line['cpu_baseline'] = {'error': f'{type(e).__name__}: {e}'[:300]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
