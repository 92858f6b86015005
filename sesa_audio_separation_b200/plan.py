"""Integer bookkeeping of the chunked inference loop (host side, bit-exact by construction).

Mirrors the schedule of ``demix`` in the reference (utils.py:382-437 == inference_pytorch.py:83-162):
fade/step/border, the border reflect-pad condition, chunk starts/lengths, the right-pad mode of each
chunk and the per-FLUSH window rule.  The engine is free to run chunks in batches of any size; the
window kind of every chunk is still decided by the reference's ``config.inference.batch_size`` grouping.
"""
from dataclasses import dataclass
from typing import List

import torch

KIND_BOTH, KIND_NO_FADEIN, KIND_NO_FADEOUT = 0, 1, 2
PAD_ZERO, PAD_REFLECT = 0, 1


@dataclass
class DemixPlan:
    length: int
    chunk_size: int
    step: int
    border: int
    fade: int
    pad: bool
    padded: int
    starts: List[int]
    lens: List[int]
    modes: List[int]
    kinds: List[int]

    @property
    def n_chunks(self):
        return len(self.starts)


def make_plan(length: int, chunk_size: int, num_overlap: int, batch_size: int) -> DemixPlan:
    if chunk_size <= 0 or num_overlap <= 0 or batch_size <= 0:
        raise ValueError('chunk_size, num_overlap and batch_size must be positive')
    fade = chunk_size // 10
    step = chunk_size // num_overlap
    if step <= 0:
        raise ValueError('num_overlap larger than chunk_size')
    border = chunk_size - step
    pad = length > 2 * border and border > 0          # utils.py:392
    padded = length + 2 * border if pad else length
    starts, lens, modes, kinds = [], [], [], []
    pending = 0
    i = 0
    while i < padded:                                  # utils.py:413
        clen = min(chunk_size, padded - i)
        starts.append(i)
        lens.append(clen)
        modes.append(PAD_REFLECT if clen > chunk_size // 2 else PAD_ZERO)   # utils.py:417-420
        pending += 1
        i += step
        if pending >= batch_size or i >= padded:       # utils.py:428
            if i - step == 0:                          # utils.py:434
                kind = KIND_NO_FADEIN
            elif i >= padded:                          # utils.py:436
                kind = KIND_NO_FADEOUT
            else:
                kind = KIND_BOTH
            kinds += [kind] * pending
            pending = 0
    return DemixPlan(length, chunk_size, step, border, fade, pad, padded, starts, lens, modes, kinds)


def windowing_array(window_size: int, fade_size: int) -> torch.Tensor:
    """utils.py:295-327.  The ramps are taken from torch.linspace itself (host): its values are part
    of the bit-exact contract of the window-sum divisor."""
    w = torch.ones(window_size)
    if fade_size > 0:
        w[-fade_size:] = torch.linspace(1, 0, fade_size)
        w[:fade_size] = torch.linspace(0, 1, fade_size)
    return w


def shard_chunks(n_chunks: int, world: int, rank: int):
    """Contiguous chunk range [lo, hi) of ``rank``: ceil-sized blocks."""
    per = -(-n_chunks // world)
    lo = min(n_chunks, rank * per)
    return lo, min(n_chunks, lo + per)
