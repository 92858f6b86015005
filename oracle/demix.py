"""CPU oracle: restatement of the chunked inference loop ``utils.demix`` (generic mode).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows /root/reference/utils.py:330-477 and
its copy inference_pytorch.py:55-186.  ``demix_schedule`` is the integer bookkeeping (pure Python);
``demix`` runs the loop with any callable ``model_fn(arr[B,C,L]) -> [B,(N,)C,L]``.
"""
import numpy as np
import torch
import torch.nn.functional as F


def windowing_array(window_size, fade_size):
    """utils.py:295-327 — ramps come from torch.linspace itself (its CPU kernel is not
    reproduced by start + i*step)."""
    w = torch.ones(window_size)
    w[-fade_size:] = torch.linspace(1, 0, fade_size)
    w[:fade_size] = torch.linspace(0, 1, fade_size)
    return w


def demix_schedule(length, chunk_size, num_overlap, batch_size):
    """Integer schedule of utils.py:382-437.  Returns dict(step, border, fade, padded, pad(bool),
    chunks=[(start, chunk_len, pad_mode, win)]), win in {'both','nofadein','nofadeout'} decided per
    FLUSH (utils.py:432-437) and shared by every chunk of that flush."""
    fade = chunk_size // 10
    step = chunk_size // num_overlap
    border = chunk_size - step
    pad = length > 2 * border and border > 0
    padded = length + 2 * border if pad else length
    chunks, batch = [], []
    i = 0
    while i < padded:
        clen = min(chunk_size, padded - i)
        batch.append((i, clen, 'reflect' if clen > chunk_size // 2 else 'constant'))
        i += step
        if len(batch) >= batch_size or i >= padded:
            if i - step == 0:
                win = 'nofadein'
            elif i >= padded:
                win = 'nofadeout'
            else:
                win = 'both'
            chunks += [(s, l, m, win) for (s, l, m) in batch]
            batch = []
    return dict(step=step, border=border, fade=fade, padded=padded, pad=pad, chunks=chunks)


def prefer_target_instrument(training):
    """utils.py:480-499."""
    ti = training.get('target_instrument', None)
    return [ti] if ti else list(training['instruments'])


def demix(mix, model_fn, chunk_size, num_overlap, batch_size, num_instruments, return_counter=False, max_chunks=None):
    """utils.py:369-464 (generic mode) restated: returns (num_instruments, C, len) float32 ndarray.
    ``max_chunks`` stops the loop after that many chunks: a BOUNDED TIMING SAMPLE of the same loop for bench.py's CPU
    baseline (the returned array is then incomplete and must not be used as a result)."""
    mix = torch.as_tensor(np.asarray(mix), dtype=torch.float32)
    sch = demix_schedule(mix.shape[-1], chunk_size, num_overlap, batch_size)
    fade, border = sch['fade'], sch['border']
    base = windowing_array(chunk_size, fade)
    if sch['pad']:
        mix = F.pad(mix[None], (border, border), mode='reflect')[0]
    result = torch.zeros((num_instruments,) + tuple(mix.shape), dtype=torch.float32)
    counter = torch.zeros_like(result)
    k = 0
    chunks = sch['chunks']
    while k < len(chunks):
        # a flush = maximal run sharing the schedule's batch grouping
        grp = chunks[k:k + batch_size]
        parts = []
        for (s, l, m, _) in grp:
            part = mix[:, s:s + l]
            if l < chunk_size:
                part = F.pad(part[None], (0, chunk_size - l), mode=m, **({'value': 0} if m == 'constant' else {}))[0]
            parts.append(part)
        x = model_fn(torch.stack(parts, 0))
        for j, (s, l, m, win) in enumerate(grp):
            w = base.clone()
            if win == 'nofadein':
                w[:fade] = 1
            elif win == 'nofadeout':
                w[-fade:] = 1
            result[..., s:s + l] += x[j, ..., :l].cpu() * w[:l]
            counter[..., s:s + l] += w[:l]
        k += len(grp)
        if max_chunks is not None and k >= max_chunks:
            break
    est = (result / counter).numpy()
    np.nan_to_num(est, copy=False, nan=0.0)
    cnt = counter.numpy()
    if sch['pad']:
        est = est[..., border:-border]
    return (est, cnt) if return_counter else est


def normalize_audio(audio):
    """utils.py:199-217."""
    mono = audio.mean(0)
    mean, std = mono.mean(), mono.std()
    return (audio - mean) / std, {'mean': mean, 'std': std}


def denormalize_audio(audio, p):
    """utils.py:220-238."""
    return audio * p['std'] + p['mean']


def apply_tta(mix, demix_fn, waveforms_orig):
    """utils.py:241-292: channel-swapped and polarity-inverted passes averaged with the original."""
    for i, aug in enumerate([mix[::-1].copy(), -1.0 * mix.copy()]):
        w = demix_fn(aug)
        for el in w:
            if i == 0:
                waveforms_orig[el] += w[el][::-1].copy()
            else:
                waveforms_orig[el] -= w[el]
    for el in waveforms_orig:
        waveforms_orig[el] /= 3
    return waveforms_orig
