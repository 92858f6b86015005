"""CPU oracle: functional fp32 restatement of MDX23C ``TFC_TDF_net.forward``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/models/mdx23c_tfc_tdf_v3.py; works on the reference state_dict.
``cfg`` = dict(audio=dict(n_fft, hop_length, dim_f, num_channels), model=dict(num_subbands,
num_scales, scale, num_blocks_per_scale, num_channels, growth, bottleneck_factor, norm, act),
num_target_instruments).  Only norm='InstanceNorm' and act='gelu' (the shipped vocals config)
plus act='relu' are restated.
"""
import torch
import torch.nn.functional as F

from .third_party import istft_exact


def _norm(x, sd, p):
    """get_norm 'InstanceNorm' -> nn.InstanceNorm2d(c, affine=True) (:47-59): per (b,c) stats over
    the last two dims, biased variance, eps 1e-5, no running stats."""
    return F.instance_norm(x, weight=sd[p + 'weight'], bias=sd[p + 'bias'], eps=1e-5)


def _act(x, act):
    if act == 'gelu':
        return F.gelu(x)
    if act == 'relu':
        return F.relu(x)
    raise ValueError(act)


def stft_fwd(x, n_fft, hop, dim_f):
    """STFT.__call__ (:14-30): (b, c, t) -> (b, c*2, dim_f, T) with channel order (c0_re, c0_im, ...)."""
    b, c, t = x.shape
    window = torch.hann_window(n_fft, periodic=True, device=x.device)
    z = torch.stft(x.reshape(-1, t), n_fft=n_fft, hop_length=hop, window=window, center=True,
                   return_complex=True)
    z = torch.view_as_real(z).permute(0, 3, 1, 2)
    z = z.reshape(b, c, 2, -1, z.shape[-1]).reshape(b, c * 2, -1, z.shape[-1])
    return z[..., :dim_f, :]


def stft_inv(x, n_fft, hop):
    """STFT.inverse (:32-44): (..., c*2, f, T) -> (..., 2, L); zero-pads f to n_fft/2+1."""
    window = torch.hann_window(n_fft, periodic=True, device=x.device)
    batch_dims = x.shape[:-3]
    c, f, t = x.shape[-3:]
    n = n_fft // 2 + 1
    x = torch.cat([x, torch.zeros([*batch_dims, c, n - f, t], device=x.device)], -2)
    x = x.reshape([*batch_dims, c // 2, 2, n, t]).reshape([-1, 2, n, t]).permute(0, 2, 3, 1)
    z = torch.complex(x[..., 0].contiguous(), x[..., 1].contiguous())
    y = istft_exact(z, n_fft=n_fft, hop_length=hop, window=window, center=True)
    return y.reshape([*batch_dims, 2, -1])


def tfc_tdf(x, sd, p, l, act):
    """TFC_TDF.forward (:131-138) with the block layout of :104-129."""
    for i in range(l):
        q = f'{p}blocks.{i}.'
        s = F.conv2d(x, sd[q + 'shortcut.weight'])
        x = F.conv2d(_act(_norm(x, sd, q + 'tfc1.0.'), act), sd[q + 'tfc1.2.weight'], padding=1)
        h = _act(_norm(x, sd, q + 'tdf.0.'), act)
        h = F.linear(h, sd[q + 'tdf.2.weight'])
        h = _act(_norm(h, sd, q + 'tdf.3.'), act)
        h = F.linear(h, sd[q + 'tdf.5.weight'])
        x = x + h
        x = F.conv2d(_act(_norm(x, sd, q + 'tfc2.0.'), act), sd[q + 'tfc2.2.weight'], padding=1)
        x = x + s
    return x


def mdx23c_forward(sd, cfg, x):
    """TFC_TDF_net.forward (:205-242)."""
    a, m = cfg['audio'], cfg['model']
    k = m['num_subbands']
    n, scale, l, act = m['num_scales'], tuple(m['scale']), m['num_blocks_per_scale'], m['act']
    nt = cfg['num_target_instruments']
    x = stft_fwd(x, a['n_fft'], a['hop_length'], a['dim_f'])
    b, c, f, t = x.shape
    mix = x = x.reshape(b, c, k, f // k, t).reshape(b, c * k, f // k, t)           # cac2cws :191-196
    first = x = F.conv2d(x, sd['first_conv.weight'])
    x = x.transpose(-1, -2)
    enc = []
    for i in range(n):
        x = tfc_tdf(x, sd, f'encoder_blocks.{i}.tfc_tdf.', l, act)
        enc.append(x)
        p = f'encoder_blocks.{i}.downscale.conv.'
        x = F.conv2d(_act(_norm(x, sd, p + '0.'), act), sd[p + '2.weight'], stride=scale)
    x = tfc_tdf(x, sd, 'bottleneck_block.', l, act)
    for i in range(n):
        p = f'decoder_blocks.{i}.upscale.conv.'
        x = F.conv_transpose2d(_act(_norm(x, sd, p + '0.'), act), sd[p + '2.weight'], stride=scale)
        x = torch.cat([x, enc.pop()], 1)
        x = tfc_tdf(x, sd, f'decoder_blocks.{i}.tfc_tdf.', l, act)
    x = x.transpose(-1, -2)
    x = x * first
    x = torch.cat([mix, x], 1)
    x = F.conv2d(_act(F.conv2d(x, sd['final_conv.0.weight']), act), sd['final_conv.2.weight'])
    b, c, f, t = x.shape
    x = x.reshape(b, c // k, k, f, t).reshape(b, c // k, f * k, t)                 # cws2cac :198-203
    if nt > 1:
        x = x.reshape(b, nt, -1, f * k, t)
    return stft_inv(x, a['n_fft'], a['hop_length'])
