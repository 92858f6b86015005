"""Deterministic synthetic weights and inputs shared by the golden generator, the tests and bench.

TEST/BENCH INFRASTRUCTURE (see oracle/__init__.py).  Random-init weights come from a seeded CPU
``torch.Generator`` applied in sorted-key order, so the reference model (in the container that has
/root/reference) and the product model (on the GPU box) get bit-identical state_dicts without
shipping them.
"""
import math
import torch


def fill_state_dict(shapes, seed=0, dim_head=None):
    """shapes: {name: tuple}.  Returns {name: fp32 tensor}.  gamma / norm weight ~ 1 + 0.1 N(0,1);
    bias ~ 0.05 N(0,1); matrices ~ N(0,1)/sqrt(fan_in); ``rotary_embed.freqs`` keeps its analytic
    value 1/10000^(2j/d) (it is a state_dict entry in the reference, bs_roformer.py:384-385)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        r = torch.randn(shp, generator=g, dtype=torch.float32)
        if name.endswith('rotary_embed.freqs'):
            d = shp[0] * 2
            out[name] = 1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))
        elif name.endswith('gamma') or (len(shp) == 1 and name.endswith('weight')):
            out[name] = 1.0 + 0.1 * r
        elif name.endswith('bias'):
            out[name] = 0.05 * r
        else:
            fan_in = max(1, math.prod(shp[1:]))
            out[name] = r / math.sqrt(fan_in)
    return out


def synth_mix(length, channels=2, seed=1234, sr=44100):
    """SURVEY §8(d) synthetic input: 0.1*N(0,1) noise + 3 random sinusoids per channel,
    peak-normalised to 0.9; float32 (channels, length)."""
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(channels, length, generator=g)
    t = torch.arange(length, dtype=torch.float64) / sr
    for c in range(channels):
        for _ in range(3):
            f = 50.0 + 4000.0 * torch.rand(1, generator=g).item()
            a = 0.2 + 0.5 * torch.rand(1, generator=g).item()
            ph = 6.283185307179586 * torch.rand(1, generator=g).item()
            x[c] += (a * torch.sin(6.283185307179586 * f * t + ph)).float()
    x = x * (0.9 / x.abs().max())
    return x.numpy()
