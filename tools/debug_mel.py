#!/usr/bin/env python
"""Evidence for the note in tests/test_gpu_parity.py::test_full_size_mel_band_roformer_chunk_vs_oracle_on_gpu: the oracle's
Mel-Band forward evaluated with CUDA tensors disagrees with the same oracle on the CPU once stems x frames is large
(complex scatter_add_ on CUDA), while the engine matches the CPU oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sesa_audio_separation_b200 as sesa  # noqa: E402
from conftest import max_rel, snr_db  # noqa: E402
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
model.load_state_dict(sd)
model.eval().to('cuda')
y = model(x.cuda()).cpu().numpy()
with torch.inference_mode():
    ref_gpu = orof.mel_band_roformer_forward({k: v.cuda() for k, v in sd.items()}, cfg, x.cuda()).cpu().numpy()
    ref_cpu = orof.mel_band_roformer_forward(sd, cfg, x).numpy()
print('oracle on CUDA vs oracle on CPU: max_rel %.3e snr %.1f dB' % (max_rel(ref_cpu, ref_gpu), snr_db(ref_cpu, ref_gpu)))
print('engine vs oracle on CPU:        max_rel %.3e snr %.1f dB' % (max_rel(ref_cpu, y), snr_db(ref_cpu, y)))
