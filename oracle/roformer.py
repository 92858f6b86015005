"""CPU oracle: functional fp32 restatement of the BS-RoFormer / Mel-Band-RoFormer inference forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works on a reference ``state_dict`` (same key
layout as the reference modules) so no reference class is needed at run time.  Every function cites
the reference lines it follows; paths are relative to /root/reference.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from .third_party import apply_rotary, istft_exact, mel as _mel

DEFAULT_FREQS_PER_BANDS = (  # models/bs_roformer/bs_roformer.py:315-324
    (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)
)


def rmsnorm(x, gamma):
    """models/bs_roformer/bs_roformer.py:43-50 — F.normalize (L2, eps 1e-12) * sqrt(dim) * gamma."""
    return F.normalize(x, dim=-1) * (x.shape[-1] ** 0.5) * gamma


def attention(x, sd, p, heads, dim_head):
    """models/bs_roformer/bs_roformer.py:106-121.  The softmax core follows Attend's dispatch
    (models/bs_roformer/attend.py:76-126): with ``flash_attn=True`` (every shipped config) a CPU tensor goes to
    ``F.scaled_dot_product_attention`` with all back ends allowed (``cpu_config``, :56,89-93) — this is the reference's
    stock CPU path and the one the timed CPU baseline must run; a CUDA fp32 tensor takes the explicit einsum softmax
    (:113-126), which is what the reference needs for fp32 on a GPU (flash-only SDPA rejects fp32, SURVEY a-10)."""
    b, n, _ = x.shape
    xn = rmsnorm(x, sd[p + 'norm.gamma'])
    qkv = F.linear(xn, sd[p + 'to_qkv.weight'])
    qkv = qkv.reshape(b, n, 3, heads, dim_head).permute(2, 0, 3, 1, 4)  # qkv b h n d
    q, k, v = qkv[0], qkv[1], qkv[2]
    freqs = sd[p + 'rotary_embed.freqs']
    q = apply_rotary(q, freqs)
    k = apply_rotary(k, freqs)
    if x.device.type == 'cpu':
        out = F.scaled_dot_product_attention(q, k, v, dropout_p=0.)
    else:
        sim = torch.einsum('bhid,bhjd->bhij', q, k) * (dim_head ** -0.5)
        attn = sim.softmax(dim=-1)
        out = torch.einsum('bhij,bhjd->bhid', attn, v)
    gates = F.linear(xn, sd[p + 'to_gates.weight'], sd[p + 'to_gates.bias'])  # b n h
    out = out * gates.permute(0, 2, 1).unsqueeze(-1).sigmoid()
    out = out.permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    return F.linear(out, sd[p + 'to_out.0.weight'])


def feedforward(x, sd, p):
    """models/bs_roformer/bs_roformer.py:55-74 (dropout = identity in eval)."""
    h = rmsnorm(x, sd[p + 'net.0.gamma'])
    h = F.linear(h, sd[p + 'net.1.weight'], sd[p + 'net.1.bias'])
    h = F.gelu(h)
    return F.linear(h, sd[p + 'net.4.weight'], sd[p + 'net.4.bias'])


def transformer(x, sd, p, depth, heads, dim_head, norm_output):
    """models/bs_roformer/bs_roformer.py:211-217 (BS: norm_output=False :376;
    Mel: output RMSNorm, mel_band_roformer.py:218,226)."""
    for j in range(depth):
        x = attention(x, sd, f'{p}layers.{j}.0.', heads, dim_head) + x
        x = feedforward(x, sd, f'{p}layers.{j}.1.') + x
    if norm_output:
        x = rmsnorm(x, sd[p + 'norm.gamma'])
    return x


def band_split(x, sd, dim_inputs):
    """models/bs_roformer/bs_roformer.py:241-249."""
    outs = []
    for b, xs in enumerate(x.split(list(dim_inputs), dim=-1)):
        p = f'band_split.to_features.{b}.'
        h = rmsnorm(xs, sd[p + '0.gamma'])
        outs.append(F.linear(h, sd[p + '1.weight'], sd[p + '1.bias']))
    return torch.stack(outs, dim=-2)


def mask_estimator(x, sd, n, dim_inputs, n_linears):
    """models/bs_roformer/bs_roformer.py:252-310 (BS: ``depth`` Linears) and
    models/bs_roformer/mel_band_roformer.py:261-319 (Mel: ``depth + 1`` Linears); tanh between,
    GLU at the end, concatenated over bands."""
    outs = []
    for b, xb in enumerate(x.unbind(dim=-2)):
        h = xb
        for li in range(n_linears):
            p = f'mask_estimators.{n}.to_freqs.{b}.0.{2 * li}.'
            h = F.linear(h, sd[p + 'weight'], sd[p + 'bias'])
            if li != n_linears - 1:
                h = torch.tanh(h)
        outs.append(F.glu(h, dim=-1))
    return torch.cat(outs, dim=-1)


def axial_layers(x, sd, depth, t_depth, f_depth, heads, dim_head, norm_output, skip_connection=False):
    """models/bs_roformer/bs_roformer.py:506-546: x is (b, t, f, d)."""
    b, t, f, d = x.shape
    store = [None] * depth
    for i in range(depth):
        if skip_connection:
            for j in range(i):
                x = x + store[j]
        x = x.permute(0, 2, 1, 3).reshape(b * f, t, d)
        x = transformer(x, sd, f'layers.{i}.0.', t_depth, heads, dim_head, norm_output)
        x = x.reshape(b, f, t, d).permute(0, 2, 1, 3).reshape(b * t, f, d)
        x = transformer(x, sd, f'layers.{i}.1.', f_depth, heads, dim_head, norm_output)
        x = x.reshape(b, t, f, d)
        if skip_connection:
            store[i] = x
    return x


def _stft(raw_audio, n_fft, hop, win_length):
    """bs_roformer.py:470-494: (b, s, L) -> real view (b, f*s, t, 2), frequency-major with the
    channel interleaved inside frequency ('b s f t c -> b (f s) t c')."""
    b, s, L = raw_audio.shape
    window = torch.hann_window(win_length, device=raw_audio.device)
    z = torch.stft(raw_audio.reshape(b * s, L), n_fft=n_fft, hop_length=hop, win_length=win_length,
                   normalized=False, window=window, return_complex=True)
    z = torch.view_as_real(z).reshape(b, s, z.shape[-2], z.shape[-1], 2)
    return z.permute(0, 2, 1, 3, 4).reshape(b, -1, z.shape[3], 2), window


def _istft(spec, s, n_fft, hop, win_length, window, length):
    """bs_roformer.py:571-582: spec complex (b, n, f*s, t) -> (b, n, s, L)."""
    b, n, fs, t = spec.shape
    z = spec.reshape(b, n, fs // s, s, t).permute(0, 1, 3, 2, 4).reshape(b * n * s, fs // s, t)
    y = istft_exact(z, n_fft=n_fft, hop_length=hop, win_length=win_length, normalized=False,
                    window=window, return_complex=False, length=length)
    return y.reshape(b, n, s, -1)


def bs_roformer_forward(sd, cfg, raw_audio):
    """models/bs_roformer/bs_roformer.py:447-587 (inference branch).  ``cfg``: dict with the
    constructor kwargs (dim, depth, stereo, num_stems, time/freq_transformer_depth, freqs_per_bands,
    dim_head, heads, stft_n_fft, stft_hop_length, stft_win_length, mask_estimator_depth,
    skip_connection)."""
    if raw_audio.ndim == 2:
        raw_audio = raw_audio[:, None]
    s = 2 if cfg.get('stereo', False) else 1
    assert raw_audio.shape[1] == s
    heads, dh = cfg.get('heads', 8), cfg.get('dim_head', 64)
    n_fft, hop = cfg.get('stft_n_fft', 2048), cfg.get('stft_hop_length', 512)
    win = cfg.get('stft_win_length', 2048)
    fpb = tuple(cfg.get('freqs_per_bands', DEFAULT_FREQS_PER_BANDS))
    dim_inputs = tuple(2 * f * s for f in fpb)
    num_stems = cfg.get('num_stems', 1)

    stft_repr, window = _stft(raw_audio, n_fft, hop, win)                    # b (f s) t c
    b, fs, t, _ = stft_repr.shape
    x = stft_repr.permute(0, 2, 1, 3).reshape(b, t, fs * 2)                 # b t (f c)
    x = band_split(x, sd, dim_inputs)
    x = axial_layers(x, sd, cfg['depth'], cfg.get('time_transformer_depth', 2),
                     cfg.get('freq_transformer_depth', 2), heads, dh, False,
                     cfg.get('skip_connection', False))
    x = rmsnorm(x, sd['final_norm.gamma'])
    mask = torch.stack([mask_estimator(x, sd, n, dim_inputs, cfg.get('mask_estimator_depth', 2))
                        for n in range(num_stems)], dim=1)                   # b n t (f c)
    mask = mask.reshape(b, num_stems, t, fs, 2).permute(0, 1, 3, 2, 4).contiguous()
    spec = torch.view_as_complex(stft_repr.contiguous())[:, None] * torch.view_as_complex(mask)
    y = _istft(spec, s, n_fft, hop, win, window, raw_audio.shape[-1])
    return y[:, 0] if num_stems == 1 else y


def mel_band_index_maps(cfg):
    """models/bs_roformer/mel_band_roformer.py:405-443.  Returns (freq_indices[int64],
    num_freqs_per_band[list], num_bands_per_freq[int64 tensor], freqs_per_band[bool (bands, freqs)])."""
    n_fft = cfg.get('stft_n_fft', 2048)
    freqs = n_fft // 2 + 1
    num_bands = cfg.get('num_bands', 60)
    fb = torch.from_numpy(_mel(sr=cfg.get('sample_rate', 44100), n_fft=n_fft, n_mels=num_bands))
    fb[0][0] = 1.
    fb[-1, -1] = 1.
    freqs_per_band = fb > 0
    assert freqs_per_band.any(dim=0).all()
    rep = torch.arange(freqs)[None].expand(num_bands, freqs)
    freq_indices = rep[freqs_per_band]
    if cfg.get('stereo', False):
        freq_indices = (freq_indices[:, None] * 2 + torch.arange(2)).reshape(-1)
    return (freq_indices, freqs_per_band.sum(1).tolist(), freqs_per_band.sum(0), freqs_per_band)


def mel_band_roformer_forward(sd, cfg, raw_audio):
    """models/bs_roformer/mel_band_roformer.py:480-633 (inference branch)."""
    if raw_audio.ndim == 2:
        raw_audio = raw_audio[:, None]
    s = 2 if cfg.get('stereo', False) else 1
    assert raw_audio.shape[1] == s
    heads, dh = cfg.get('heads', 8), cfg.get('dim_head', 64)
    n_fft, hop = cfg.get('stft_n_fft', 2048), cfg.get('stft_hop_length', 512)
    win = cfg.get('stft_win_length', 2048)
    num_stems = cfg.get('num_stems', 1)
    freq_indices, nfpb, nbpf, _ = mel_band_index_maps(cfg)
    freq_indices = freq_indices.to(raw_audio.device)
    dim_inputs = tuple(2 * f * s for f in nfpb)
    length = raw_audio.shape[-1] if cfg.get('match_input_audio_length', False) else None

    stft_repr, window = _stft(raw_audio, n_fft, hop, win)                    # b (f s) t c
    b, fs, t, _ = stft_repr.shape
    x = stft_repr[:, freq_indices]                                          # :530
    x = x.permute(0, 2, 1, 3).reshape(b, t, -1)
    x = band_split(x, sd, dim_inputs)
    x = axial_layers(x, sd, cfg['depth'], cfg.get('time_transformer_depth', 2),
                     cfg.get('freq_transformer_depth', 2), heads, dh, True,
                     cfg.get('skip_connection', False))
    masks = torch.stack([mask_estimator(x, sd, n, dim_inputs, cfg.get('mask_estimator_depth', 1) + 1)
                         for n in range(num_stems)], dim=1)                  # b n t (f c)
    masks = masks.reshape(b, num_stems, t, -1, 2).permute(0, 1, 3, 2, 4).contiguous()
    masks = torch.view_as_complex(masks)
    spec = torch.view_as_complex(stft_repr.contiguous())[:, None]           # b 1 (f s) t
    idx = freq_indices[None, None, :, None].expand(b, num_stems, -1, t)
    summed = torch.zeros(b, num_stems, fs, t, dtype=spec.dtype, device=spec.device).scatter_add_(2, idx, masks)
    denom = nbpf.to(raw_audio.device).repeat_interleave(s)[:, None]          # '(f r) 1'
    spec = spec * (summed / denom.clamp(min=1e-8))
    y = _istft(spec, s, n_fft, hop, win, window, length)
    return y[:, 0] if num_stems == 1 else y
