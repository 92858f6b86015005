"""Restatements of the two third-party pieces of arithmetic on the hot path.

Neither package is vendored or version-pinned by the reference (requirements.txt:16 ``librosa``,
requirements.txt:37 ``rotary_embedding_torch``) and neither is installed here, so these follow
the published algorithms.  Status: parity unpinned (no upstream vectors available offline).

* ``RotaryEmbedding`` — lucidrains/rotary-embedding-torch, ``lang`` frequencies, theta=10000,
  interleaved pairs.  Call sites: models/bs_roformer/bs_roformer.py:112-113,384-385 and
  models/bs_roformer/mel_band_roformer.py:121-122,381-382.
* ``mel`` — librosa.filters.mel (Slaney scale, Slaney norm).  Call site:
  models/bs_roformer/mel_band_roformer.py:410; only ``weights > 0`` is consumed (:425).
"""
import numpy as np
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    """freqs = 1 / theta^(arange(0, dim, 2) / dim), kept as a parameter named ``freqs`` so that the
    reference state_dict layout (``layers.i.j.layers.k.0.rotary_embed.freqs``) is reproduced."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        return apply_rotary(t, self.freqs)


def rotary_angles(n, freqs):
    """angles (n, 2*len(freqs)) with every frequency repeated for the two members of a pair."""
    pos = torch.arange(n, dtype=freqs.dtype, device=freqs.device)
    ang = torch.einsum('i,j->ij', pos, freqs)
    return torch.repeat_interleave(ang, 2, dim=-1)


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rotary(t, freqs):
    """t: (..., n, d).  Positions 0..n-1 along dim -2, computed in fp32, cast back to t.dtype."""
    ang = rotary_angles(t.shape[-2], freqs.float())
    tf = t.float()
    out = tf * ang.cos() + rotate_half(tf) * ang.sin()
    return out.to(t.dtype)


# ----------------------------------------------------------------------------- librosa.filters.mel

def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    out = np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)
    return out


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, dtype=np.float32):
    """Slaney-scale triangular filter bank with Slaney (area) normalisation, (n_mels, 1+n_fft//2)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=dtype)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def istft_exact(z, **kw):
    """torch.istft evaluated where it is accurate.  Measured on B200 / torch 2.11 (tools/debug_c1.py, kept under
    profiles/r2_torch_istft_cuda.md): on CUDA tensors torch.istft is exact for 1-2 signals but off by up to 1.6e-2 of the
    peak for a batch of 8 signals and 7e-3 for 4, for any n_fft / hop / length, while torch.stft on CUDA and everything
    else in the forward agree with the CPU to ~1e-6.  The CPU evaluation is the reference's own CPU path, so a CUDA
    spectrogram is inverted on the CPU and the waveform handed back on its device."""
    if z.device.type == 'cpu':
        return torch.istft(z, **kw)
    kw = dict(kw)
    if kw.get('window') is not None:
        kw['window'] = kw['window'].cpu()
    return torch.istft(z.cpu(), **kw).to(z.device)
