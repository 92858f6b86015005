#!/usr/bin/env python
"""Where does the oracle's Mel-Band forward evaluated with CUDA tensors part from the same oracle on the CPU?
Runs oracle/roformer.py's stages one by one on both devices (fp32, TF32 off) and prints the max relative difference of
every intermediate, feeding each CUDA stage with the CPU stage's input so the first diverging OP is isolated."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sesa_audio_separation_b200 as sesa  # noqa: E402
from oracle import roformer as orof  # noqa: E402
from oracle.weights import fill_state_dict, synth_mix  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = dict(dim=384, depth=1, stereo=True, num_stems=4, time_transformer_depth=1, freq_transformer_depth=1, num_bands=60,
           dim_head=64, heads=8, stft_n_fft=2048, stft_hop_length=441, stft_win_length=2048, mask_estimator_depth=2,
           sample_rate=44100)
x = torch.from_numpy(synth_mix(441 * 400, 2, seed=42))[None]
model = sesa.MelBandRoformer(**cfg)
sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=41)
sdg = {k: v.cuda() for k, v in sd.items()}


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    return float((a.double() - b.double()).abs().max() / a.double().abs().max())


with torch.inference_mode():
    s, heads, dh, n_fft, hop, num_stems = 2, 8, 64, 2048, 441, 4
    fi, nfpb, nbpf, _ = orof.mel_band_index_maps(cfg)
    dim_inputs = tuple(2 * f * s for f in nfpb)
    st_c, win_c = orof._stft(x, n_fft, hop, n_fft)
    st_g, win_g = orof._stft(x.cuda(), n_fft, hop, n_fft)
    print('stft            ', rel(st_c, st_g))
    b, fs, t, _ = st_c.shape
    xc = st_c[:, fi].permute(0, 2, 1, 3).reshape(b, t, -1)
    xg = st_c.cuda()[:, fi.cuda()].permute(0, 2, 1, 3).reshape(b, t, -1)
    print('gather          ', rel(xc, xg))
    bc = orof.band_split(xc, sd, dim_inputs)
    bg = orof.band_split(xc.cuda(), sdg, dim_inputs)
    print('band_split      ', rel(bc, bg))
    # one transformer at a time
    tc_ = bc.permute(0, 2, 1, 3).reshape(b * bc.shape[2], t, -1)
    a_c = orof.attention(tc_, sd, 'layers.0.0.layers.0.0.', heads, dh)
    a_g = orof.attention(tc_.cuda(), sdg, 'layers.0.0.layers.0.0.', heads, dh)
    print('time attention  ', rel(a_c, a_g))
    f_c = orof.feedforward(tc_, sd, 'layers.0.0.layers.0.1.')
    f_g = orof.feedforward(tc_.cuda(), sdg, 'layers.0.0.layers.0.1.')
    print('feedforward     ', rel(f_c, f_g))
    lc = orof.axial_layers(bc, sd, 1, 1, 1, heads, dh, True)
    lg = orof.axial_layers(bc.cuda(), sdg, 1, 1, 1, heads, dh, True)
    print('axial layers    ', rel(lc, lg))
    mc = torch.stack([orof.mask_estimator(lc, sd, n, dim_inputs, 3) for n in range(num_stems)], dim=1)
    mg = torch.stack([orof.mask_estimator(lc.cuda(), sdg, n, dim_inputs, 3) for n in range(num_stems)], dim=1)
    print('mask estimator  ', rel(mc, mg))
    for n in range(num_stems):
        print('   stem', n, rel(mc[:, n], mg[:, n]))
    masks = torch.view_as_complex(mc.reshape(b, num_stems, t, -1, 2).permute(0, 1, 3, 2, 4).contiguous())
    spec = torch.view_as_complex(st_c.contiguous())[:, None]
    idx = fi[None, None, :, None].expand(b, num_stems, -1, t)
    sum_c = torch.zeros(b, num_stems, fs, t, dtype=spec.dtype).scatter_add_(2, idx, masks)
    sum_g = torch.zeros(b, num_stems, fs, t, dtype=spec.dtype, device='cuda').scatter_add_(2, idx.cuda(), masks.cuda())
    print('complex scatter ', rel(sum_c, sum_g))
    sum_r = torch.view_as_complex(torch.zeros(b, num_stems, fs, t, 2, device='cuda').index_add_(2, fi.cuda(), torch.view_as_real(masks.cuda())))
    print('real index_add  ', rel(sum_c, sum_r))
    denom = nbpf.repeat_interleave(s)[:, None]
    sp_c = spec * (sum_c / denom.clamp(min=1e-8))
    sp_g = spec.cuda() * (sum_c.cuda() / denom.cuda().clamp(min=1e-8))
    print('mask multiply   ', rel(sp_c, sp_g))
    y_c = orof._istft(sp_c, s, n_fft, hop, n_fft, win_c, None)
    y_g = orof._istft(sp_c.cuda(), s, n_fft, hop, n_fft, win_g, None)
    print('istft           ', rel(y_c, y_g))
