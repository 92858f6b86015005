"""MDX23C TFC-TDF-v3 inference on the sm_100a kernel library.

Host-side mirror of ``TFC_TDF_net`` (models/mdx23c_tfc_tdf_v3.py:141-242): same constructor (a config object with
``audio`` / ``model`` / ``training`` sections), same state_dict key layout, same ``forward(x[B, C, L])`` contract
(returns ``[B, C, L']`` for one target instrument, ``[B, N, C, L']`` otherwise).

Execution plan (B200-first; nothing here is a translation of the reference's cuDNN/cuBLAS call sequence):
  * activations are channels-last fp32 ``x[b][t][f][c]`` (the reference's ``transpose(-1, -2)`` at :213/:226 is a
    no-op in this layout);
  * every ``norm -> act -> conv/Linear`` (:104-128, :74-97) = InstanceNorm statistics kernel + ONE elementwise pass
    that writes bf16 hi/lo planes + a tcgen05 GEMM: 3x3 convolutions as 9-tap implicit GEMMs (shifted TMA boxes,
    zero padding by TMA out-of-bounds fill), the stride-2 Downscale convolution as a 4-tap implicit GEMM with
    element-strided TMA boxes, the ConvTranspose2d Upscale as four GEMMs whose epilogue scatters rows to the 2x grid,
    1x1 convolutions and the TDF Linears as plain GEMMs (TDF operands are written channel-major by the prologue);
  * ``x + shortcut`` is the residual epilogue of the tfc2 GEMM; encoder outputs are written straight into the
    second half of the decoder's concat buffer (no ``torch.cat`` copy).
"""
import ctypes

import numpy as np
import torch

from . import _lib, tc
from ._lib import GemmEpilogue, call
from .config import prefer_target_instrument
from .module import KernelModule
from .roformer import _istft_envelope, _twiddle

ACT_CODES = {'gelu': _lib.ACT_GELU}
TAPS3 = [(kh - 1, kw - 1) for kh in range(3) for kw in range(3)]
TAPS2 = [(kh, kw) for kh in range(2) for kw in range(2)]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ep(act=0, residual=0):
    return GemmEpilogue(0, act, residual, 0, 0, 0, 1, 1, None)


def _r64(n):
    return (int(n) + 63) // 64 * 64


class TFC_TDF_net(KernelModule):
    """Drop-in for models.mdx23c_tfc_tdf_v3.TFC_TDF_net (constructor surface :142)."""

    def __init__(self, config, seed=0):
        super().__init__()
        self.config = config
        a, m = config.audio, config.model
        if m.norm != 'InstanceNorm':
            raise NotImplementedError("only norm='InstanceNorm' (the shipped MDX23C configs) is implemented")
        if m.act not in ACT_CODES:
            raise NotImplementedError(f"act={m.act!r} is not implemented (the shipped MDX23C configs use gelu)")
        self.act = ACT_CODES[m.act]
        self.num_target_instruments = len(prefer_target_instrument(config))
        self.k = int(m.num_subbands)
        self.audio_channels = int(a.num_channels)
        self.dim_c = self.k * self.audio_channels * 2
        self.n = int(m.num_scales)
        self.scale = tuple(int(s) for s in m.scale)
        if self.scale != (2, 2):
            raise NotImplementedError('only scale [2, 2] is implemented')
        self.l = int(m.num_blocks_per_scale)
        self.c0, self.g, self.bn = int(m.num_channels), int(m.growth), int(m.bottleneck_factor)
        self.n_fft, self.hop, self.dim_f = int(a.n_fft), int(a.hop_length), int(a.dim_f)
        self.fs = self.dim_f // self.k
        self.precision = 'fp32'
        self._ws = {}
        g = torch.Generator().manual_seed(seed)
        c, f = self.c0, self.fs
        self._register('first_conv.weight', (c, self.dim_c, 1, 1), 'linear_w', g)
        self.blocks = []   # (prefix, in_c, c, f) of every TFC_TDF in forward order
        for i in range(self.n):
            self._reg_tfc_tdf(f'encoder_blocks.{i}.tfc_tdf.', c, c, f, g)
            p = f'encoder_blocks.{i}.downscale.conv.'
            self._register(p + '0.weight', (c,), 'ones', g)
            self._register(p + '0.bias', (c,), 'zeros', g)
            self._register(p + '2.weight', (c + self.g, c, 2, 2), 'linear_w', g)
            f //= 2
            c += self.g
        self._reg_tfc_tdf('bottleneck_block.', c, c, f, g)
        for i in range(self.n):
            p = f'decoder_blocks.{i}.upscale.conv.'
            self._register(p + '0.weight', (c,), 'ones', g)
            self._register(p + '0.bias', (c,), 'zeros', g)
            self._register(p + '2.weight', (c, c - self.g, 2, 2), 'linear_w', g)
            f *= 2
            c -= self.g
            self._reg_tfc_tdf(f'decoder_blocks.{i}.tfc_tdf.', 2 * c, c, f, g)
        self._register('final_conv.0.weight', (c, c + self.dim_c, 1, 1), 'linear_w', g)
        self._register('final_conv.2.weight', (self.num_target_instruments * self.dim_c, c, 1, 1), 'linear_w', g)

    def _reg_tfc_tdf(self, p, in_c, c, f, g):
        for i in range(self.l):
            q = f'{p}blocks.{i}.'
            for name, ch in (('tfc1.0', in_c), ('tdf.0', c), ('tdf.3', c), ('tfc2.0', c)):
                self._register(q + name + '.weight', (ch,), 'ones', g)
                self._register(q + name + '.bias', (ch,), 'zeros', g)
            self._register(q + 'tfc1.2.weight', (c, in_c, 3, 3), 'linear_w', g)
            self._register(q + 'tdf.2.weight', (f // self.bn, f), 'linear_w', g)
            self._register(q + 'tdf.5.weight', (f, f // self.bn), 'linear_w', g)
            self._register(q + 'tfc2.2.weight', (c, c, 3, 3), 'linear_w', g)
            self._register(q + 'shortcut.weight', (c, in_c, 1, 1), 'linear_w', g)
            in_c = c

    def set_precision(self, precision):
        if precision not in ('fp32', 'bf16'):
            raise ValueError("MDX23C runs on the tensor-core path only: precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    # ------------------------------------------------------------------ weights
    def _conv_w(self, w):
        """[co][ci][kh][kw] -> bf16 planes of [co][tap][round_up(ci, 64)] (tap = kh*KW + kw)."""
        co, ci, kh, kw = w.shape
        cp = _r64(ci)
        wt = torch.zeros(co, kh * kw, cp, device=w.device)
        wt[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
        return tc.split_weight(wt.reshape(co, kh * kw * cp))

    def _prepare(self):
        _lib.require_cuda()
        if self._device.type != 'cuda':
            raise _lib.SesaError('model must be moved to a CUDA device before forward(); no CPU path exists')
        P, dev = self._params, self._device
        prep = {}
        for k, w in P.items():
            if k.endswith('tfc1.2.weight') or k.endswith('tfc2.2.weight') or k.endswith('downscale.conv.2.weight'):
                prep[k] = self._conv_w(w)
            elif k.endswith('upscale.conv.2.weight'):      # ConvTranspose2d: [ci][co][kh][kw] -> 4 x [co][ci]
                prep[k] = [tc.split_weight(w[:, :, kh, kw].t().contiguous()) for kh, kw in TAPS2]
            elif w.ndim == 4:                              # 1x1 convolutions
                prep[k] = tc.split_weight(w.reshape(w.shape[0], w.shape[1]))
            elif k.endswith('tdf.2.weight') or k.endswith('tdf.5.weight'):
                prep[k] = tc.split_weight(w)
            else:
                prep[k] = w.contiguous()
        win = torch.hann_window(self.n_fft)
        prep['window_cpu'] = win
        prep['window'] = win.to(dev)
        prep['twiddle'] = _twiddle(self.n_fft).to(dev)
        self._prepared = prep
        self._ws = {}
        return prep

    # ------------------------------------------------------------------ launch plan (built once per input shape)
    def _plan(self, B, L):
        key = (B, L)
        if key in self._ws:
            return self._ws[key]
        prep, dev = self._prepared, self._device
        T = 1 + L // self.hop
        Ffull = self.n_fft // 2 + 1
        C2 = self.audio_channels * 2
        f32 = dict(device=dev, dtype=torch.float32)
        ws = dict(T=T)
        steps = []          # closures executed in order by forward()
        keep = []           # buffers / tables referenced by raw pointers

        def buf(*shape):
            t = torch.zeros(*shape, **f32)
            keep.append(t)
            return t

        def planes(rows, cols):
            t = tc.alloc_planes(rows, cols, dev)
            keep.append(t)
            return t

        def gemm(problems, ep, block_n=256):
            tab = tc.TcGemmTable(problems, dev, block_n=block_n)
            keep.append(tab)
            steps.append(lambda: tab.run(ep, 3 if self.precision == 'fp32' else 1))

        stats_scratch = torch.zeros(2 * B * 2048, device=dev, dtype=torch.float64)
        keep.append(stats_scratch)

        def norm_planes(x, coff, C, stats_geo, split_geo, gamma, beta, out, st=None):
            """InstanceNorm statistics per (b, c), then norm + affine + act + bf16 split into `out` planes.
            stats_geo = (layout, n1, n2, ld) of sesa_instnorm_stats; split_geo = (mode, n1, n2, ld) of sesa_norm_act_split.
            ``st``: statistics already produced by the kernel that wrote x (no separate pass)."""
            have = st is not None
            st = st if have else buf(B, C, 2)
            xp = x.data_ptr() + 4 * coff
            (sl, sn1, sn2, sld), (mode, n1, n2, ld) = stats_geo, split_geo

            def run():
                if not have:
                    call('sesa_instnorm_stats', ctypes.c_void_p(xp), sl, B, sn1, C, sn2, sld, _ptr(stats_scratch), _ptr(st),
                         1e-5, _stream())
                call('sesa_norm_act_split', ctypes.c_void_p(xp), mode, B, n1, C, n2, ld, _ptr(st), _ptr(gamma), _ptr(beta),
                     self.act, _ptr(out), out.shape[-1], out.stride(0), _stream())
            steps.append(run)

        def cl(n_pos, ld):      # channels-last in, channels-last planes out
            return (0, n_pos, 1, ld), (0, n_pos, 1, ld)

        def conv_problem(a_planes, w_planes, cin, Tt, Ff, inT, inF, stride, taps, C_out, cptr, ldc):
            return dict(A=tc.planes_arg(a_planes), W=tc.planes_arg(w_planes), M=B * Tt * Ff, N=C_out,
                        K=len(taps) * _r64(cin), C=(cptr, ldc),
                        conv=dict(cin=cin, B=B, T=Tt, F=Ff, inT=inT, inF=inF, stride=stride, taps=taps))

        def tfc_tdf(prefix, x, x_ld, x_off, in_c, c, Tt, Ff, out, out_ld, out_off, x_raw, out_raw=None):
            """One TFC_TDF module (:100-138); x: (tensor, ld, channel offset); writes `out` likewise.  ``x_raw``: bf16 planes
            of the raw input (A operand of the first shortcut conv), written by whichever GEMM produced x; ``out_raw``
            (planes tensor, column offset) asks the module's last GEMM to leave the raw planes of ITS output there."""
            M = B * Tt * Ff
            J = Ff // self.bn
            for i in range(self.l):
                q = f'{prefix}blocks.{i}.'
                last = i == self.l - 1
                y, y_ld, y_off = (out, out_ld, out_off) if last else (buf(M, c), c, 0)
                yptr = y.data_ptr() + 4 * y_off
                # raw planes of y for its consumer's shortcut conv come out of the tfc2 epilogue below (no separate pass)
                y_raw = (planes(M, c), 0) if not last else out_raw
                # shortcut: 1x1 conv of the raw input -> y
                xr, xr_off = x_raw
                gemm([dict(A=tc.planes_arg(xr, xr_off), W=tc.planes_arg(prep[q + 'shortcut.weight']), M=M, N=c, K=in_c,
                           C=(yptr, y_ld))], _ep())
                # tfc1: norm -> act -> conv3x3 -> x1
                xa = planes(M, in_c)
                norm_planes(x, x_off, in_c, *cl(Tt * Ff, x_ld), prep[q + 'tfc1.0.weight'], prep[q + 'tfc1.0.bias'], xa)
                x1 = buf(M, c)
                gemm([conv_problem(xa, prep[q + 'tfc1.2.weight'], in_c, Tt, Ff, Tt, Ff, 1, TAPS3, c, x1.data_ptr(), c)], _ep())
                # tdf: norm -> act -> Linear(F -> F/bn) -> norm -> act -> Linear(F/bn -> F); x1 += tdf
                ta = planes(B * Tt * c, Ff)               # channel-major [b][t][c][f]
                norm_planes(x1, 0, c, (0, Tt * Ff, 1, c), (1, Tt, Ff, c), prep[q + 'tdf.0.weight'], prep[q + 'tdf.0.bias'], ta)
                hmid = buf(B * Tt * c, J)
                gemm([dict(A=tc.planes_arg(ta), W=tc.planes_arg(prep[q + 'tdf.2.weight']), M=B * Tt * c, N=J, K=Ff,
                           C=(hmid.data_ptr(), J))], _ep())
                tb = planes(B * Tt * c, J)
                norm_planes(hmid, 0, c, (1, Tt, J, 0), (2, Tt, J, 0), prep[q + 'tdf.3.weight'], prep[q + 'tdf.3.bias'], tb)
                gout = buf(B * Tt * c, Ff)
                gemm([dict(A=tc.planes_arg(tb), W=tc.planes_arg(prep[q + 'tdf.5.weight']), M=B * Tt * c, N=Ff, K=J,
                           C=(gout.data_ptr(), Ff))], _ep())
                # x1 += tdf(x1), with the statistics of the sum for tfc2's norm gathered in the same pass
                st2 = buf(B, c, 2)
                steps.append(lambda x1=x1, gout=gout, c=c, st2=st2: call(
                    'sesa_transpose_add_stats', _ptr(x1), _ptr(gout), B, Tt, Ff, c, c, _ptr(stats_scratch), _ptr(st2), 1e-5,
                    _stream()))
                # tfc2: norm -> act -> conv3x3, + shortcut (already in y) through the residual epilogue
                xb = planes(M, c)
                norm_planes(x1, 0, c, *cl(Tt * Ff, c), prep[q + 'tfc2.0.weight'], prep[q + 'tfc2.0.bias'], xb, st=st2)
                pr = conv_problem(xb, prep[q + 'tfc2.2.weight'], c, Tt, Ff, Tt, Ff, 1, TAPS3, c, yptr, y_ld)
                if y_raw is not None:
                    pr['P'] = tc.planes_arg(y_raw[0], y_raw[1])
                gemm([pr], _ep(residual=1))
                x, x_ld, x_off, in_c, x_raw = y, y_ld, y_off, c, y_raw
            return x, x_ld, x_off

        # ---- front end: STFT -> sub-band channels (cac2cws) -> first_conv
        spec = buf(B * T, Ffull * C2)
        M0 = B * T * self.fs
        mix = buf(M0, self.dim_c)
        mixp = planes(M0, self.dim_c)
        c = self.c0
        first = buf(M0, c)
        ws['spec'], ws['mix'] = spec, mix
        steps.append(lambda: call('sesa_mdx_pack', _ptr(spec), B * T, Ffull, self.fs, self.k, C2, _ptr(mix), _stream()))
        steps.append(lambda: call('sesa_norm_act_split', _ptr(mix), 0, B, T * self.fs, self.dim_c, 1, self.dim_c, None, None,
                                  None, 0, _ptr(mixp), mixp.shape[-1], mixp.stride(0), _stream()))
        x_raw = (planes(M0, c), 0)
        gemm([dict(A=tc.planes_arg(mixp), W=tc.planes_arg(prep['first_conv.weight']), M=M0, N=c, K=self.dim_c,
                   C=(first.data_ptr(), c), P=tc.planes_arg(x_raw[0]))], _ep())
        # ---- encoder
        x, x_ld, x_off = first, c, 0
        Tt, Ff = T, self.fs
        skips = []
        for i in range(self.n):
            cat = buf(B * Tt * Ff, 2 * c)      # decoder concat buffer: [upscaled | encoder output]
            catp = planes(B * Tt * Ff, 2 * c)  # its raw planes: both halves are written by the GEMMs that produce them
            x, x_ld, x_off = tfc_tdf(f'encoder_blocks.{i}.tfc_tdf.', x, x_ld, x_off, c, c, Tt, Ff, cat, 2 * c, c, x_raw,
                                     out_raw=(catp, c))
            skips.append((cat, catp, c, Tt, Ff))
            p = f'encoder_blocks.{i}.downscale.conv.'
            xa = planes(B * Tt * Ff, c)
            norm_planes(x, x_off, c, *cl(Tt * Ff, x_ld), prep[p + '0.weight'], prep[p + '0.bias'], xa)
            nxt = buf(B * (Tt // 2) * (Ff // 2), c + self.g)
            x_raw = (planes(B * (Tt // 2) * (Ff // 2), c + self.g), 0)
            pr = conv_problem(xa, prep[p + '2.weight'], c, Tt // 2, Ff // 2, Tt, Ff, 2, TAPS2, c + self.g, nxt.data_ptr(),
                              c + self.g)
            pr['P'] = tc.planes_arg(x_raw[0])
            gemm([pr], _ep())
            x, x_ld, x_off = nxt, c + self.g, 0
            Tt, Ff, c = Tt // 2, Ff // 2, c + self.g
        bott = buf(B * Tt * Ff, c)
        x, x_ld, x_off = tfc_tdf('bottleneck_block.', x, x_ld, x_off, c, c, Tt, Ff, bott, c, 0, x_raw)
        # ---- decoder
        for i in range(self.n):
            p = f'decoder_blocks.{i}.upscale.conv.'
            cat, catp, cs, Ts, Fs_ = skips.pop()
            xa = planes(B * Tt * Ff, c)
            norm_planes(x, x_off, c, *cl(Tt * Ff, x_ld), prep[p + '0.weight'], prep[p + '0.bias'], xa)
            probs = []
            for ti, (kh, kw) in enumerate(TAPS2):
                probs.append(dict(A=tc.planes_arg(xa), W=tc.planes_arg(prep[p + '2.weight'][ti]), M=B * Tt * Ff, N=c - self.g,
                                  K=c, C=(cat.data_ptr(), 2 * cs), P=tc.planes_arg(catp), row_map=(Ff, kh, kw)))
            gemm(probs, _ep())
            Tt, Ff, c = Tt * 2, Ff * 2, c - self.g
            assert (cs, Ts, Fs_) == (c, Tt, Ff)
            dec = buf(B * Tt * Ff, c)
            x, x_ld, x_off = tfc_tdf(f'decoder_blocks.{i}.tfc_tdf.', cat, 2 * c, 0, 2 * c, c, Tt, Ff, dec, c, 0, (catp, 0))
        # ---- head: x * first_conv_out, cat(mix, x), final_conv (1x1 -> act -> 1x1)
        fin = planes(M0, self.dim_c + c)
        steps.append(lambda x=x, ld=x_ld: call('sesa_mdx_final_concat', _ptr(mix), self.dim_c, _ptr(x), ld, _ptr(first), c, c, M0,
                                               _ptr(fin), fin.shape[-1], fin.stride(0), _stream()))
        hid = planes(M0, c)
        gemm([dict(A=tc.planes_arg(fin), W=tc.planes_arg(prep['final_conv.0.weight']), M=M0, N=c, K=self.dim_c + c,
                   P=tc.planes_arg(hid))], _ep(act=self.act))
        nout = self.num_target_instruments * self.dim_c
        yout = buf(M0, nout)
        gemm([dict(A=tc.planes_arg(hid), W=tc.planes_arg(prep['final_conv.2.weight']), M=M0, N=nout, K=c,
                   C=(yout.data_ptr(), nout))], _ep())
        out_len = self.hop * (T - 1)                      # torch.istft(length=None), :42
        ospec = buf(B * self.num_target_instruments * T, Ffull * C2)
        env = _istft_envelope(prep['window_cpu'], self.n_fft, self.hop, T, out_len).to(dev)
        steps.append(lambda: call('sesa_mdx_unpack', _ptr(yout), B, T, self.fs, self.k, C2, self.num_target_instruments, Ffull,
                                  _ptr(ospec), _stream()))
        ws.update(steps=steps, keep=keep, ospec=ospec, env=env, out_len=out_len)
        self._keep_workspace(key, ws)
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, x, out=None):
        if not isinstance(x, torch.Tensor) or x.device.type != 'cuda':
            raise _lib.SesaError('forward() needs a CUDA tensor; there is no CPU path')
        B, C, L = x.shape
        if C != self.audio_channels:
            raise AssertionError(f'expected {self.audio_channels} audio channels, got {C}')
        prep = self._prepared or self._prepare()
        audio = x.to(torch.float32).contiguous()
        ws = self._plan(B, L)
        T = ws['T']
        Ffull = self.n_fft // 2 + 1
        st = _stream()
        call('sesa_stft', _ptr(audio), _ptr(ws['spec']), _ptr(prep['window']), _ptr(prep['twiddle']), B, C, L, self.n_fft,
             self.hop, 0, Ffull, st)
        for step in ws['steps']:
            step()
        nt = self.num_target_instruments
        if out is None:
            out = torch.empty(B, nt, C, ws['out_len'], device=audio.device, dtype=torch.float32)
        else:
            assert out.is_contiguous() and out.numel() == B * nt * C * ws['out_len']
            out = out.view(B, nt, C, ws['out_len'])
        call('sesa_mask_istft', _ptr(ws['ospec']), None, None, None, _ptr(out), _ptr(prep['window']), _ptr(ws['env']),
             _ptr(prep['twiddle']), B, nt, C, self.n_fft, self.hop, T, ws['out_len'], 2, 0, st)
        return out[:, 0] if nt == 1 else out
