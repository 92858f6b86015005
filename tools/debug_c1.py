#!/usr/bin/env python
"""Two diagnostics behind the full-config parity tests (tests/test_gpu_parity.py):
 (1) torch.istft evaluated on CUDA vs on the CPU for the oracle's call shapes (where, and for which shapes, they part);
 (2) BASELINE C1 (MDX23C, 30-s track, 27 chunks): error of the engine per step-long region of the track against the
     oracle evaluated on CUDA, and for the worst chunk engine vs oracle-on-CUDA vs oracle-on-CPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sesa_audio_separation_b200 as sesa  # noqa: E402
from oracle import demix as odemix  # noqa: E402
from oracle import mdx23c as omdx  # noqa: E402
from oracle.weights import fill_state_dict, synth_mix  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def istft_probe():
    g = torch.Generator().manual_seed(0)
    for n_fft, hop, sig, T, use_len in [(2048, 441, 8, 401, False), (2048, 441, 8, 401, True), (2048, 441, 2, 401, False),
                                        (2048, 441, 8, 801, False), (2048, 441, 1, 50, False), (8192, 1024, 4, 256, False),
                                        (2048, 512, 8, 401, False)]:
        z = torch.randn(sig, n_fft // 2 + 1, T, dtype=torch.complex64, generator=g)
        w = torch.hann_window(n_fft)
        length = hop * (T - 1) if use_len else None
        a = torch.istft(z, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=w, length=length)
        b = torch.istft(z.cuda(), n_fft=n_fft, hop_length=hop, win_length=n_fft, window=w.cuda(), length=length).cpu()
        d = (a - b).abs()
        i = int(d.argmax())
        print(f'istft n_fft {n_fft} hop {hop} signals {sig} frames {T} length={length}: max|d| {float(d.max()):.3e} / max|ref| '
              f'{float(a.abs().max()):.3e} at signal {i // a.shape[1]} sample {i % a.shape[1]} of {a.shape[1]}; '
              f'frac of samples with |d| > 1e-5*max: {float((d > 1e-5 * a.abs().max()).float().mean()):.4f}', flush=True)


def c1_probe():
    model, cfg = sesa.get_model_from_config('mdx23c', os.path.join(ROOT, 'configs', 'config_vocals_mdx23c.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=21)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    L, ov, bs = int(cfg.audio.chunk_size), int(cfg.inference.num_overlap), int(cfg.inference.batch_size)
    mix = synth_mix(30 * 44100, 2, seed=93)
    eng = sesa.DemixEngine(cfg, model, 'cuda', engine_batch=4)
    est = eng.run(mix)
    plan = eng.plan
    ocfg = dict(audio=dict(cfg.audio), model=dict(cfg.model), num_target_instruments=2)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = odemix.demix(mix, lambda a: omdx.mdx23c_forward(sd_gpu, ocfg, a.cuda()).cpu(), L, ov, bs, 2)
    peak = np.abs(ref).max()
    step, border = plan.step, plan.border
    print('C1 track: global max_rel', np.abs(ref - est).max() / peak)
    for r in range(0, -(-plan.padded // step)):
        a, b = max(r * step - border, 0), min((r + 1) * step - border, mix.shape[1])
        if b <= a:
            continue
        d = np.abs(ref[..., a:b] - est[..., a:b])
        print(f'  region {r:2d} samples [{a}, {b}): max|d|/peak {d.max() / peak:.3e}  region peak/peak {np.abs(ref[..., a:b]).max() / peak:.3f}')
    # chunk by chunk: engine vs oracle on CUDA vs oracle on CPU
    padded = np.pad(mix, ((0, 0), (border, border)), mode='reflect')
    torch.set_num_threads(os.cpu_count())
    for k in (0, 5, plan.n_chunks - 3, plan.n_chunks - 2, plan.n_chunks - 1):
        s, n = plan.starts[k], plan.lens[k]
        part = torch.from_numpy(padded[:, s:s + n].copy())
        if n < L:
            mode = 'reflect' if n > L // 2 else 'constant'
            part = torch.nn.functional.pad(part[None], (0, L - n), mode=mode)[0]
        x = part[None]
        with torch.inference_mode():
            y = model(x.cuda()).cpu().numpy()
            og = omdx.mdx23c_forward(sd_gpu, ocfg, x.cuda()).cpu().numpy()
            oc = omdx.mdx23c_forward(sd, ocfg, x).numpy()
        pk = np.abs(oc).max()
        print(f'  chunk {k} (len {n}): engine vs oracle-CPU {np.abs(y - oc).max() / pk:.3e}; engine vs oracle-CUDA '
              f'{np.abs(y - og).max() / pk:.3e}; oracle-CUDA vs oracle-CPU {np.abs(og - oc).max() / pk:.3e}', flush=True)


if __name__ == '__main__':
    istft_probe()
    c1_probe()
