"""BS-RoFormer and Mel-Band-RoFormer inference on the sm_100a kernel library.

Host-side mirror of the reference model classes (same constructor keywords, same state_dict key
layout, same forward() contract):
  * BSRoformer       models/bs_roformer/bs_roformer.py:327-587
  * MelBandRoformer  models/bs_roformer/mel_band_roformer.py:324-633
The forward pass keeps the residual stream in ONE token-major buffer x[(b t f), d]; the reference's
rearranges between the time and band transformers (bs_roformer.py:526-543) become stride choices of
the attention kernel, RMSNorm is fused into the consuming GEMM (gamma*sqrt(d) folded into the weight,
1/||x|| applied in the epilogue), rotary embedding, biases, GELU/tanh/GLU, gating and residual adds are
GEMM / attention epilogues, and the 62 per-band Linears of BandSplit / MaskEstimator run as grouped
launches.  Training branches (target=..., losses), LinearAttention and use_torch_checkpoint are out
of scope and raise.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, tc
from ._lib import GEMM_GROUP_DTYPE, GemmEpilogue, call
from .module import KernelModule

PRECISIONS = ('fp32', 'bf16', 'fp32_simt')

DEFAULT_FREQS_PER_BANDS = (          # bs_roformer.py:315-324
    (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)
)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _twiddle(n):
    k = torch.arange(n, dtype=torch.float64)
    ang = -2.0 * math.pi * k / n
    return torch.stack([ang.cos(), ang.sin()], dim=-1).float().contiguous()


def _istft_envelope(window, n_fft, hop, n_frames, out_len):
    """Window-square envelope exactly as torch.istft builds it (col2im of window^2), trimmed to the
    kept region [n_fft/2, n_fft/2 + out_len)."""
    expected = n_fft + hop * (n_frames - 1)
    w2 = window.pow(2).view(1, n_fft, 1).expand(1, n_fft, n_frames)
    env = torch.nn.functional.fold(w2, output_size=(1, expected), kernel_size=(1, n_fft), stride=(1, hop))
    env = env.reshape(-1)[n_fft // 2:]
    if env.numel() < out_len:
        env = torch.cat([env, torch.ones(out_len - env.numel())])
    env = env[:out_len]
    if float(env.abs().min()) < 1e-11:
        raise RuntimeError('istft: window overlap-add envelope has zeros (NOLA violated)')
    return env.contiguous()


class _GroupTable:
    """Device-resident array of sesa_gemm_group records for one grouped launch."""

    def __init__(self, records, device):
        arr = np.zeros(len(records), dtype=GEMM_GROUP_DTYPE)
        for i, r in enumerate(records):
            arr[i] = (r['A'], r['W'], r.get('bias', 0) or 0, r['C'], r['M'], r['N'], r['K'], 0,
                      r['lda'], r['ldw'], r['ldc'])
        self.n = len(records)
        self.max_m = max(r['M'] for r in records)
        self.max_n = max(r['N'] for r in records)
        self.dev = torch.from_numpy(arr.view(np.uint8).copy()).to(device)


def _epilogue(rownorm=0, act=0, residual=0, glu=0, rot=None, rot_cols=0, rot_dim=0, pos_div=1, pos_mod=1):
    return GemmEpilogue(rownorm, act, residual, glu, rot_cols, rot_dim, pos_div, pos_mod,
                        rot.data_ptr() if rot is not None else None)


class _RoformerBase(KernelModule):
    """Shared engine of the two band-split RoFormers."""

    norm_output = False     # Mel: every Transformer ends with an RMSNorm (mel_band_roformer.py:218,226)
    has_final_norm = True   # BS: final_norm before the mask estimators (bs_roformer.py:548)
    mask_extra_linear = 0   # Mel's MLP has depth+1 Linears (mel_band_roformer.py:271)

    def _init_common(self, dim, depth, stereo, num_stems, time_transformer_depth, freq_transformer_depth,
                     linear_transformer_depth, dim_head, heads, flash_attn, stft_n_fft, stft_hop_length,
                     stft_win_length, stft_normalized, stft_window_fn, mask_estimator_depth,
                     mlp_expansion_factor, use_torch_checkpoint, skip_connection, dim_inputs, seed):
        super().__init__()
        if linear_transformer_depth > 0:
            raise NotImplementedError('linear_transformer_depth > 0 (LinearAttention) is out of scope')
        if stft_normalized:
            raise NotImplementedError('stft_normalized=True is not used by any shipped config')
        if stft_window_fn is not None:
            raise NotImplementedError('custom stft_window_fn is not supported (hann only)')
        if stft_win_length != stft_n_fft:
            raise NotImplementedError('stft_win_length != stft_n_fft is not supported')
        if dim_head != 64:
            raise NotImplementedError('the attention kernels are built for dim_head == 64')
        if heads > 8:
            raise NotImplementedError('the engine supports at most 8 attention heads (gate logits are stored 8 per token)')
        self.dim, self.depth, self.stereo = dim, depth, stereo
        self.audio_channels = 2 if stereo else 1
        self.num_stems = num_stems
        self.t_depth, self.f_depth = time_transformer_depth, freq_transformer_depth
        self.dim_head, self.heads = dim_head, heads
        self.inner = dim_head * heads
        self.n_fft, self.hop = stft_n_fft, stft_hop_length
        self.mask_estimator_depth = mask_estimator_depth
        self.mlp_expansion_factor = mlp_expansion_factor
        self.skip_connection = skip_connection
        self.dim_inputs = tuple(int(d) for d in dim_inputs)
        self.num_bands = len(self.dim_inputs)
        self.n_mask_linears = mask_estimator_depth + self.mask_extra_linear
        self.precision = 'fp32'   # 'fp32': split-bf16 (3 MMAs) tensor-core path at fp32-parity tolerance;
        self._ws = {}             # 'bf16': single bf16 MMA; 'fp32_simt': exact IEEE fp32 SIMT cross-check path
        g = torch.Generator().manual_seed(seed)
        D, inner, H = dim, self.inner, heads
        for i in range(depth):
            for a, sub_depth in ((0, time_transformer_depth), (1, freq_transformer_depth)):
                for j in range(sub_depth):
                    p = f'layers.{i}.{a}.layers.{j}.'
                    self._params[p + '0.rotary_embed.freqs'] = 1.0 / (10000 ** (torch.arange(0, dim_head, 2).float() / dim_head))
                    self._register(p + '0.norm.gamma', (D,), 'ones', g)
                    self._register(p + '0.to_qkv.weight', (3 * inner, D), 'linear_w', g)
                    self._register(p + '0.to_gates.weight', (H, D), 'linear_w', g)
                    self._register(p + '0.to_gates.bias', (H,), ('uniform', 1.0 / math.sqrt(D)), g)
                    self._register(p + '0.to_out.0.weight', (D, inner), 'linear_w', g)
                    self._register(p + '1.net.0.gamma', (D,), 'ones', g)
                    self._register(p + '1.net.1.weight', (4 * D, D), 'linear_w', g)
                    self._register(p + '1.net.1.bias', (4 * D,), ('uniform', 1.0 / math.sqrt(D)), g)
                    self._register(p + '1.net.4.weight', (D, 4 * D), 'linear_w', g)
                    self._register(p + '1.net.4.bias', (D,), ('uniform', 1.0 / math.sqrt(4 * D)), g)
                if self.norm_output:
                    self._register(f'layers.{i}.{a}.norm.gamma', (D,), 'ones', g)
        if self.has_final_norm:
            self._register('final_norm.gamma', (D,), 'ones', g)
        for b, din in enumerate(self.dim_inputs):
            p = f'band_split.to_features.{b}.'
            self._register(p + '0.gamma', (din,), 'ones', g)
            self._register(p + '1.weight', (D, din), 'linear_w', g)
            self._register(p + '1.bias', (D,), ('uniform', 1.0 / math.sqrt(din)), g)
        hidden = D * mlp_expansion_factor
        for n in range(num_stems):
            for b, din in enumerate(self.dim_inputs):
                dims = (D,) + (hidden,) * (self.n_mask_linears - 1) + (2 * din,)
                for li in range(self.n_mask_linears):
                    p = f'mask_estimators.{n}.to_freqs.{b}.0.{2 * li}.'
                    self._register(p + 'weight', (dims[li + 1], dims[li]), 'linear_w', g)
                    self._register(p + 'bias', (dims[li + 1],), ('uniform', 1.0 / math.sqrt(dims[li])), g)

    def set_precision(self, precision):
        """'fp32' (default): tensor cores with split-bf16 operands (3 MMAs per product, fp32 accumulate) — meets the
        fp32 parity gate; 'bf16': one bf16 MMA per product (SNR >= 40 dB gate); 'fp32_simt': exact IEEE fp32 SIMT
        kernels (the in-library cross-check)."""
        if precision not in PRECISIONS:
            raise ValueError(f'precision must be one of {PRECISIONS}, got {precision!r}')
        if precision != self.precision:
            self.precision = precision
            self._prepared = None
            self._ws = {}
        return self

    @property
    def _tc(self):
        return self.precision != 'fp32_simt'

    # ------------------------------------------------------------------ weight preparation
    def _prepare(self):
        """Fold norms / scales into the GEMM weights and build the device-side constant tables."""
        _lib.require_cuda()
        if self._device.type != 'cuda':
            raise _lib.SesaError('model must be moved to a CUDA device before forward(); no CPU path exists')
        P, dev = self._params, self._device
        D, inner, H, dh = self.dim, self.inner, self.heads, self.dim_head
        prep = {'layers': []}
        self.ld_qkv = (3 * inner + H + 3) // 4 * 4
        sD = math.sqrt(D)
        for i in range(self.depth):
            pair = []
            for a, sub_depth in ((0, self.t_depth), (1, self.f_depth)):
                subs = []
                for j in range(sub_depth):
                    p = f'layers.{i}.{a}.layers.{j}.'
                    g_attn = P[p + '0.norm.gamma'] * sD
                    wqkv = P[p + '0.to_qkv.weight'].clone()
                    wqkv[:inner] *= dh ** -0.5                     # softmax scale folded into q (attend.py:109,115)
                    w = torch.zeros(self.ld_qkv, D, device=dev)
                    w[:3 * inner] = wqkv
                    w[3 * inner:3 * inner + H] = P[p + '0.to_gates.weight']
                    w = (w * g_attn[None]).contiguous()
                    bias = torch.zeros(self.ld_qkv, device=dev)
                    bias[3 * inner:3 * inner + H] = P[p + '0.to_gates.bias']
                    g_ff = P[p + '1.net.0.gamma'] * sD
                    sub = dict(
                        wqkv=w, bqkv=bias, wo=P[p + '0.to_out.0.weight'].contiguous(),
                        w1=(P[p + '1.net.1.weight'] * g_ff[None]).contiguous(), b1=P[p + '1.net.1.bias'].contiguous(),
                        w2=P[p + '1.net.4.weight'].contiguous(), b2=P[p + '1.net.4.bias'].contiguous(),
                        freqs=P[p + '0.rotary_embed.freqs'].float().cpu())
                    if self._tc:   # bf16 hi/lo planes of the folded weights for the tcgen05 GEMMs
                        # to_qkv and to_gates share one GEMM: rows [q | k | v | gate logits]
                        sub.update(wqkv_p=tc.split_weight(w[:3 * inner + H]), bqkvg=bias[:3 * inner + H].contiguous(),
                                   wo_p=tc.split_weight(sub['wo']),
                                   w1_p=tc.split_weight(sub['w1']), w2_p=tc.split_weight(sub['w2']))
                    subs.append(sub)
                norm = P[f'layers.{i}.{a}.norm.gamma'].contiguous() if self.norm_output else None
                pair.append(dict(subs=subs, norm=norm))
            prep['layers'].append(pair)
        # band split: gamma*sqrt(dim_in) folded
        bs_w, bs_b = [], []
        for b, din in enumerate(self.dim_inputs):
            p = f'band_split.to_features.{b}.'
            bs_w.append((P[p + '1.weight'] * (P[p + '0.gamma'] * math.sqrt(din))[None]).contiguous())
            bs_b.append(P[p + '1.bias'].contiguous())
        prep['bs_w'], prep['bs_b'] = bs_w, bs_b
        if self._tc:
            prep['bs_wp'] = [tc.split_weight(w) for w in bs_w]
        # mask estimators: final_norm folded into the first Linear (BS), GLU rows interleaved in the last
        fin = (P['final_norm.gamma'] * sD) if self.has_final_norm else None
        me = []
        for n in range(self.num_stems):
            bands = []
            for b, din in enumerate(self.dim_inputs):
                lins = []
                for li in range(self.n_mask_linears):
                    p = f'mask_estimators.{n}.to_freqs.{b}.0.{2 * li}.'
                    w, bias = P[p + 'weight'], P[p + 'bias']
                    if li == 0 and fin is not None:
                        w = w * fin[None]
                    if li == self.n_mask_linears - 1:
                        w = torch.stack([w[:din], w[din:]], dim=1).reshape(2 * din, -1)
                        bias = torch.stack([bias[:din], bias[din:]], dim=1).reshape(2 * din)
                    lins.append((w.contiguous(), bias.contiguous()) +
                                ((tc.split_weight(w),) if self._tc else ()))
                bands.append(lins)
            me.append(bands)
        prep['mask'] = me
        win = torch.hann_window(self.n_fft)                          # periodic, as torch.hann_window default
        prep['window_cpu'] = win
        prep['window'] = win.to(dev)
        prep['twiddle'] = _twiddle(self.n_fft).to(dev)
        prep['rot'] = {}
        self._extra_prepare(prep)
        self._prepared = prep
        self._ws = {}
        return prep

    def _extra_prepare(self, prep):
        pass

    def _rot_table(self, prep, freqs, n):
        """(cos, sin) of pos*freq for pos < n — computed exactly like the reference's rotary module
        (fp32 outer product, then cos/sin), oracle/third_party.py."""
        key = (tuple(freqs.tolist()), n, self._tc)
        if key not in prep['rot']:
            ang = torch.einsum('i,j->ij', torch.arange(n, dtype=torch.float32), freqs)
            tab = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()           # [pos][dh/2][2]
            if self._tc:    # quad-major for the tensor-core epilogue: [dh/4][pos][4] (include/sesa_b200.h)
                tab = tab.reshape(n, -1, 4).permute(1, 0, 2).contiguous()
            prep['rot'][key] = tab.to(self._device)
        return prep['rot'][key]

    # ------------------------------------------------------------------ workspace per batch size
    def _workspace(self, B, L):
        key = (B, L)
        if key in self._ws:
            ws = self._ws.pop(key)          # most recently used last
            self._ws[key] = ws
            return ws
        prep, dev = self._prepared, self._device
        T = 1 + L // self.hop
        F = self.n_fft // 2 + 1
        C = self.audio_channels
        nb, D = self.num_bands, self.dim
        M = B * T * nb
        ftot = sum(self.dim_inputs)
        hidden = D * self.mlp_expansion_factor
        out_len = self._out_len(L, T)
        f32 = dict(device=dev, dtype=torch.float32)
        ws = dict(T=T, F=F, M=M, out_len=out_len, ftot=ftot)
        ws['spec'] = torch.empty(B * T, F * C * 2, **f32)
        ws['x'] = torch.empty(M, D, **f32)
        ws['mask'] = torch.empty(self.num_stems, B * T, ftot, **f32)
        ws['env'] = _istft_envelope(prep['window_cpu'], self.n_fft, self.hop, T, out_len).to(dev)
        if self.skip_connection:
            ws['store'] = [torch.empty(M, D, **f32) for _ in range(self.depth)]
        feat = self._features_buffer(ws, B, T)
        if self._tc:
            self._workspace_tc(ws, prep, B, T)
        else:
            self._workspace_simt(ws, prep, B, T, feat)
        self._keep_workspace(key, ws)
        return ws

    def _workspace_simt(self, ws, prep, B, T, feat):
        """fp32 activation buffers and group tables of the 'fp32_simt' cross-check mode only (the tensor-core modes never
        read them; for the 4-stem Mel model they would be ~10 GB of dead memory)."""
        dev = self._device
        nb, D, M, ftot = self.num_bands, self.dim, ws['M'], ws['ftot']
        hidden = D * self.mlp_expansion_factor
        f32 = dict(device=dev, dtype=torch.float32)
        ws['qkv'] = torch.empty(M, self.ld_qkv, **f32)
        ws['ao'] = torch.empty(M, self.inner, **f32)
        ws['h'] = torch.empty(M, 4 * D, **f32)
        ws['mh'] = [torch.empty(self.num_stems * nb, B * T, hidden, **f32)
                    for _ in range(min(2, self.n_mask_linears - 1))]
        # grouped GEMM tables
        offs = np.concatenate([[0], np.cumsum(self.dim_inputs)]).astype(np.int64)
        recs = []
        for b, din in enumerate(self.dim_inputs):
            recs.append(dict(A=feat.data_ptr() + 4 * int(offs[b]), W=prep['bs_w'][b].data_ptr(),
                             bias=prep['bs_b'][b].data_ptr(), C=ws['x'].data_ptr() + 4 * b * D,
                             M=B * T, N=D, K=din, lda=ftot, ldw=din, ldc=nb * D))
        ws['g_bandsplit'] = _GroupTable(recs, dev)
        ws['g_mask'] = []
        for li in range(self.n_mask_linears):
            recs = []
            for n in range(self.num_stems):
                for b, din in enumerate(self.dim_inputs):
                    w, bias = prep['mask'][n][b][li][:2]
                    if li == 0:
                        A, lda = ws['x'].data_ptr() + 4 * b * D, nb * D
                    else:
                        src = ws['mh'][(li - 1) % 2]
                        A, lda = src[n * nb + b].data_ptr(), hidden
                    if li == self.n_mask_linears - 1:
                        Cp, ldc = ws['mask'][n].data_ptr() + 4 * int(offs[b]), ftot
                    else:
                        dst = ws['mh'][li % 2]
                        Cp, ldc = dst[n * nb + b].data_ptr(), hidden
                    recs.append(dict(A=A, W=w.data_ptr(), bias=bias.data_ptr(), C=Cp, M=B * T, N=w.shape[0],
                                     K=w.shape[1], lda=lda, ldw=w.shape[1], ldc=ldc))
            ws['g_mask'].append(_GroupTable(recs, dev))

        def single(A, W, bias, Cm, M_, N_, K_, lda, ldc):
            return _GroupTable([dict(A=A.data_ptr(), W=W.data_ptr(), bias=bias.data_ptr() if bias is not None else 0,
                                     C=Cm.data_ptr(), M=M_, N=N_, K=K_, lda=lda, ldw=W.shape[1], ldc=ldc)], dev)
        ws['g_layers'] = []
        for pair in prep['layers']:
            gp = []
            for tr in pair:
                gs = []
                for s in tr['subs']:
                    gs.append(dict(
                        qkv=single(ws['x'], s['wqkv'], s['bqkv'], ws['qkv'], M, self.ld_qkv, D, D, self.ld_qkv),
                        out=single(ws['ao'], s['wo'], None, ws['x'], M, D, self.inner, self.inner, D),
                        ff1=single(ws['x'], s['w1'], s['b1'], ws['h'], M, 4 * D, D, D, 4 * D),
                        ff2=single(ws['h'], s['w2'], s['b2'], ws['x'], M, D, 4 * D, 4 * D, D)))
                gp.append(gs)
            ws['g_layers'].append(gp)

    def _workspace_tc(self, ws, prep, B, T):
        """Plane buffers and TMA tables of the tensor-core path (built once per batch shape)."""
        dev = self._device
        nb, D, inner, M = self.num_bands, self.dim, self.inner, ws['M']
        hidden = D * self.mlp_expansion_factor
        BT = B * T
        ws['xp'] = tc.alloc_planes(M, D, dev)          # normalised residual stream (A of qkv / ff1 / first mask Linear)
        ws['aop'] = tc.alloc_planes(M, inner, dev)     # attention output (A of to_out)
        ws['hp'] = tc.alloc_planes(M, 4 * D, dev)      # GELU(ff1) (A of ff2)
        ws['qkvp'] = tc.alloc_planes(M, 3 * inner, dev)  # rotated q (pre-scaled) | rotated k | v
        ws['gates'] = torch.zeros(M, 8, device=dev, dtype=torch.float32)   # to_gates logits
        x = ws['x']
        # per-row partial sums of squares of the residual stream, written by the residual GEMM epilogues (one slot per
        # column block and half) and consumed as the fused RMSNorm row scale of the next GEMM
        slots = tc.SS_SLOTS_PER_BLOCK * ((D + 255) // 256)
        ws['ss_slots'] = slots
        ws['ss'] = torch.zeros(M, slots, device=dev, dtype=torch.float32)
        ss_ptr = ws['ss'].data_ptr()

        def one(A, Wp, Mm, N, K, bias=None, C=None, Pl=None, **extra):
            return tc.TcGemmTable([dict(A=tc.planes_arg(A), W=tc.planes_arg(Wp), M=Mm, N=N, K=K,
                                        bias=bias.data_ptr() if bias is not None else 0,
                                        C=(C.data_ptr(), C.shape[-1]) if C is not None else None,
                                        P=tc.planes_arg(Pl) if Pl is not None else None, **extra)], dev)
        # BandSplit on the tensor cores: per-band normalised feature planes (bands at 16-byte aligned plane columns) feed
        # ONE grouped launch whose epilogue also emits the residual stream's planes and row sums of squares
        offs = np.concatenate([[0], np.cumsum(self.dim_inputs)]).astype(np.int32)
        poffs = np.concatenate([[0], np.cumsum([tc.round8(d) for d in self.dim_inputs])]).astype(np.int32)
        ws['bs_offs'] = torch.from_numpy(offs).to(dev)
        ws['bs_poffs'] = torch.from_numpy(poffs).to(dev)
        ws['featp'] = tc.alloc_planes(BT, int(poffs[-1]), dev)
        fp = ws['featp']
        probs = []
        for b, din in enumerate(self.dim_inputs):
            probs.append(dict(A=(fp.data_ptr() + 2 * int(poffs[b]), fp.shape[-1], fp.stride(0)), W=tc.planes_arg(prep['bs_wp'][b]),
                              M=BT, N=D, K=din, bias=prep['bs_b'][b].data_ptr(), C=(x.data_ptr() + 4 * b * D, nb * D),
                              P=(ws['xp'].data_ptr() + 2 * b * D, nb * D, ws['xp'].stride(0)),
                              ss_out=ss_ptr + 4 * b * slots, ss_ld=nb * slots))
        ws['t_bandsplit'] = tc.TcGemmTable(probs, dev)
        ws['t_layers'] = []
        for pair in prep['layers']:
            gp = []
            for tr in pair:
                gs = []
                for s in tr['subs']:
                    gs.append(dict(
                        qkv=one(ws['xp'], s['wqkv_p'], M, 3 * inner + self.heads, D, bias=s['bqkvg'], C=ws['gates'],
                                Pl=ws['qkvp'], rowss=ss_ptr, ss_slots=slots, p_cols=3 * inner, c_col0=3 * inner),
                        out=one(ws['aop'], s['wo_p'], M, D, inner, C=x, Pl=ws['xp'], ss_out=ss_ptr),
                        ff1=one(ws['xp'], s['w1_p'], M, 4 * D, D, bias=s['b1'], Pl=ws['hp'], rowss=ss_ptr, ss_slots=slots),
                        ff2=one(ws['hp'], s['w2_p'], M, D, 4 * D, bias=s['b2'], C=x, Pl=ws['xp'], ss_out=ss_ptr)))
                gp.append(gs)
            ws['t_layers'].append(gp)
        # mask estimators: grouped over (stem, band)
        nl = self.n_mask_linears
        ws['mhp'] = [torch.zeros(2, self.num_stems * nb * BT, hidden, device=dev, dtype=torch.bfloat16)
                     for _ in range(min(2, nl - 1))]
        offs = np.concatenate([[0], np.cumsum(self.dim_inputs)]).astype(np.int64)
        xp = ws['xp']
        ws['t_mask'] = []
        for li in range(nl):
            probs = []
            for n in range(self.num_stems):
                for b, din in enumerate(self.dim_inputs):
                    w, bias, wp = prep['mask'][n][b][li]
                    if li == 0:     # rows (b t) of band b inside the token-major stream: row stride nb*D
                        A = (xp.data_ptr() + 2 * b * D, nb * D, xp.stride(0))
                    else:
                        src = ws['mhp'][(li - 1) % 2]
                        A = (src.data_ptr() + 2 * (n * nb + b) * BT * hidden, hidden, src.stride(0))
                    pr = dict(A=A, W=tc.planes_arg(wp), M=BT, N=w.shape[0], K=w.shape[1], bias=bias.data_ptr())
                    if li == nl - 1:
                        pr['C'] = (ws['mask'][n].data_ptr() + 4 * int(offs[b]), ws['ftot'])
                    else:
                        dst = ws['mhp'][li % 2]
                        pr['P'] = (dst.data_ptr() + 2 * (n * nb + b) * BT * hidden, hidden, dst.stride(0))
                    probs.append(pr)
            ws['t_mask'].append(tc.TcGemmTable(probs, dev))

    def _features_buffer(self, ws, B, T):
        return ws['spec']

    def _out_len(self, L, T):
        return L

    # ------------------------------------------------------------------ forward
    def _gemm(self, table, ep):
        call('sesa_gemm_simt', _ptr(table.dev), table.n, table.max_m, table.max_n, ctypes.byref(ep), _stream())

    def _refresh_planes(self, ws):
        """(Re)build the bf16 planes and row sums of squares of the residual stream after a non-GEMM producer."""
        tc.prep_rows(ws['x'], ws['M'], self.dim, self.dim, ws['xp'], 2, rowinv=ws['ss'], ss_slots=ws['ss_slots'],
                     out_planes=2 if self.precision == 'fp32' else 1)

    def _transformer(self, ws, prep, gl, tr, axis, B, tgl=None):
        T, nb, D, H = ws['T'], self.num_bands, self.dim, self.heads
        M = ws['M']
        if axis == 0:   # time: sequences (b, f), positions t (row stride nb)
            n_seq, seq_len, inner_cnt, outer, inner_s, pos_s = B * nb, T, nb, T * nb, 1, nb
            pos_div, pos_mod = nb, T
        else:           # band: sequences (b, t), positions f
            n_seq, seq_len, inner_cnt, outer, inner_s, pos_s = B * T, nb, 1, nb, 0, 1
            pos_div, pos_mod = 1, nb
        nsplit = 3 if self.precision == 'fp32' else 1
        npl = 2 if nsplit == 3 else 1      # bf16 mode never reads the lo planes: do not write them either
        for si, (s, g) in enumerate(zip(tr['subs'], gl or [None] * len(tr['subs']))):
            rot = self._rot_table(prep, s['freqs'], seq_len)
            if self._tc:
                t = tgl[si]
                inner = self.inner
                # xp holds the raw residual stream as bf16 planes and ss its row sums of squares (both written by the
                # previous residual epilogue): RMSNorm is the consumer's row scale, gamma*sqrt(D) lives in the weights
                t['qkv'].run(_epilogue(rot=rot, rot_cols=2 * inner, rot_dim=self.dim_head, pos_div=pos_div,
                                       pos_mod=pos_mod), nsplit, npl)
                qp, ap = ws['qkvp'], ws['aop']
                call('sesa_attention_tc', _ptr(qp), qp.shape[-1], qp.stride(0), _ptr(ws['gates']), 8, _ptr(ap),
                     ap.shape[-1], ap.stride(0), H, self.dim_head, n_seq, seq_len, inner_cnt, outer, inner_s, pos_s,
                     T if axis == 1 else 0, nsplit, npl, _stream())
                t['out'].run(_epilogue(residual=1), nsplit, npl)
                t['ff1'].run(_epilogue(act=_lib.ACT_GELU), nsplit, npl)
                t['ff2'].run(_epilogue(residual=1), nsplit, npl)
                continue
            self._gemm(g['qkv'], _epilogue(rownorm=1, rot=rot, rot_cols=2 * self.inner, rot_dim=self.dim_head,
                                           pos_div=pos_div, pos_mod=pos_mod))
            call('sesa_attention_simt', _ptr(ws['qkv']), _ptr(ws['ao']), self.ld_qkv, self.inner, H, self.dim_head,
                 n_seq, seq_len, inner_cnt, outer, inner_s, pos_s, _stream())
            self._gemm(g['out'], _epilogue(residual=1))
            self._gemm(g['ff1'], _epilogue(rownorm=1, act=_lib.ACT_GELU))
            self._gemm(g['ff2'], _epilogue(residual=1))
        if tr['norm'] is not None:
            if self._tc:    # output RMSNorm + planes / row sums of squares of the new residual stream in one pass
                xp = ws['xp']
                call('sesa_rmsnorm_planes', _ptr(ws['x']), _ptr(tr['norm']), M, D, _ptr(xp), xp.shape[-1], xp.stride(0),
                     2 if self.precision == 'fp32' else 1, _ptr(ws['ss']), ws['ss_slots'], _stream())
            else:
                call('sesa_rmsnorm', _ptr(ws['x']), _ptr(tr['norm']), _ptr(ws['x']), M, D, _stream())

    def forward(self, raw_audio, target=None, return_loss_breakdown=False, out=None):
        if target is not None:
            raise NotImplementedError('training losses are out of scope for the inference engine')
        if not isinstance(raw_audio, torch.Tensor) or raw_audio.device.type != 'cuda':
            raise _lib.SesaError('forward() needs a CUDA tensor; there is no CPU path')
        if raw_audio.ndim == 2:
            raw_audio = raw_audio[:, None]
        B, C, L = raw_audio.shape
        if C != self.audio_channels:
            raise AssertionError('stereo needs to be set to True if passing in audio signal that is stereo '
                                 '(channel dimension of 2). also need to be False if mono (channel dimension of 1)')
        prep = self._prepared or self._prepare()
        audio = raw_audio.to(torch.float32).contiguous()
        ws = self._workspace(B, L)
        T, F = ws['T'], ws['F']
        st = _stream()
        call('sesa_stft', _ptr(audio), _ptr(ws['spec']), _ptr(prep['window']), _ptr(prep['twiddle']), B, C, L,
             self.n_fft, self.hop, 0, F, st)
        self._gather_features(ws, prep, B, T)
        if self._tc:
            nsplit = 3 if self.precision == 'fp32' else 1
            feat = ws.get('feat', ws['spec'])      # Mel: gathered rows; BS: the spectrogram itself
            fp = ws['featp']
            call('sesa_band_prep', _ptr(feat), feat.shape[-1], B * T, self.num_bands, _ptr(ws['bs_offs']), _ptr(ws['bs_poffs']),
                 _ptr(fp), fp.shape[-1], fp.stride(0), 2 if nsplit == 3 else 1, st)
            ws['t_bandsplit'].run(_epilogue(), nsplit, 2 if nsplit == 3 else 1)
        else:
            self._gemm(ws['g_bandsplit'], _epilogue(rownorm=1))
        for i, (pair, gp) in enumerate(zip(prep['layers'], ws.get('g_layers') or [(None, None)] * self.depth)):
            if self.skip_connection:
                for j in range(i):
                    call('sesa_add_inplace', _ptr(ws['x']), _ptr(ws['store'][j]), ws['x'].numel(), st)
                if self._tc and i > 0:
                    self._refresh_planes(ws)
            tgp = ws['t_layers'][i] if self._tc else (None, None)
            self._transformer(ws, prep, gp[0], pair[0], 0, B, tgp[0])
            self._transformer(ws, prep, gp[1], pair[1], 1, B, tgp[1])
            if self.skip_connection:
                ws['store'][i].copy_(ws['x'])
        nl = self.n_mask_linears
        if self._tc:
            nsplit = 3 if self.precision == 'fp32' else 1
            tc.prep_rows(ws['x'], ws['M'], self.dim, self.dim, ws['xp'], self.has_final_norm)
        for li in range(nl):
            last = li == nl - 1
            if self._tc:
                ws['t_mask'][li].run(_epilogue(act=0 if last else _lib.ACT_TANH, glu=1 if last else 0), nsplit,
                                     2 if nsplit == 3 else 1)
                continue
            self._gemm(ws['g_mask'][li], _epilogue(rownorm=1 if (li == 0 and self.has_final_norm) else 0,
                                                    act=0 if last else _lib.ACT_TANH, glu=1 if last else 0))
        out_len = ws['out_len']
        if out is None:
            out = torch.empty(B, self.num_stems, C, out_len, device=audio.device, dtype=torch.float32)
        else:
            assert out.is_contiguous() and out.numel() == B * self.num_stems * C * out_len
        self._mask_istft(ws, prep, out, B, C, T, out_len, st)
        out = out.view(B, self.num_stems, C, out_len)
        if self.num_stems == 1:
            out = out[:, 0]
        return out

    def _gather_features(self, ws, prep, B, T):
        pass

    def _mask_istft(self, ws, prep, out, B, C, T, out_len, st):
        call('sesa_mask_istft', _ptr(ws['spec']), _ptr(ws['mask']), None, None, _ptr(out), _ptr(prep['window']),
             _ptr(ws['env']), _ptr(prep['twiddle']), B, self.num_stems, C, self.n_fft, self.hop, T, out_len, 0, 0, st)


class BSRoformer(_RoformerBase):
    """Drop-in for models.bs_roformer.BSRoformer (bs_roformer.py:327-363 constructor surface)."""
    norm_output = False
    has_final_norm = True
    mask_extra_linear = 0

    def __init__(self, dim, *, depth, stereo=False, num_stems=1, time_transformer_depth=2,
                 freq_transformer_depth=2, linear_transformer_depth=0,
                 freqs_per_bands=DEFAULT_FREQS_PER_BANDS, dim_head=64, heads=8, attn_dropout=0.,
                 ff_dropout=0., flash_attn=True, dim_freqs_in=1025, stft_n_fft=2048, stft_hop_length=512,
                 stft_win_length=2048, stft_normalized=False, stft_window_fn=None, mask_estimator_depth=2,
                 multi_stft_resolution_loss_weight=1., multi_stft_resolutions_window_sizes=(4096, 2048, 1024, 512, 256),
                 multi_stft_hop_size=147, multi_stft_normalized=False, multi_stft_window_fn=None,
                 mlp_expansion_factor=4, use_torch_checkpoint=False, skip_connection=False, seed=0):
        if not isinstance(freqs_per_bands, tuple):       # @beartype Tuple[int, ...] in the reference
            raise TypeError('freqs_per_bands must be a tuple (use !!python/tuple in the YAML config)')
        freqs = stft_n_fft // 2 + 1
        assert len(freqs_per_bands) > 1
        assert sum(freqs_per_bands) == freqs, (f'the number of freqs in the bands must equal {freqs} based on '
                                               f'the STFT settings, but got {sum(freqs_per_bands)}')
        ch = 2 if stereo else 1
        self._init_common(dim, depth, stereo, num_stems, time_transformer_depth, freq_transformer_depth,
                          linear_transformer_depth, dim_head, heads, flash_attn, stft_n_fft, stft_hop_length,
                          stft_win_length, stft_normalized, stft_window_fn, mask_estimator_depth,
                          mlp_expansion_factor, use_torch_checkpoint, skip_connection,
                          tuple(2 * f * ch for f in freqs_per_bands), seed)


def mel_band_maps(sample_rate, n_fft, num_bands, stereo):
    """Band index maps of mel_band_roformer.py:405-443 (pure host integer work after the filter bank).
    The Slaney mel filter bank is restated from librosa.filters.mel (not installed; un-pinned upstream)."""
    freqs = n_fft // 2 + 1
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sample_rate / 2.0), num_bands + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((num_bands, freqs), dtype=np.float32)
    for i in range(num_bands):
        weights[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    weights *= (2.0 / (mel_f[2:num_bands + 2] - mel_f[:num_bands]))[:, None]
    weights[0, 0] = 1.0            # mel_band_roformer.py:416
    weights[-1, -1] = 1.0          # :421
    freqs_per_band = weights > 0
    assert freqs_per_band.any(axis=0).all(), 'all frequencies need to be covered by all bands for now'
    rep = np.broadcast_to(np.arange(freqs), (num_bands, freqs))
    freq_indices = rep[freqs_per_band].astype(np.int64)
    if stereo:
        freq_indices = (freq_indices[:, None] * 2 + np.arange(2)[None]).reshape(-1)
    return freq_indices, freqs_per_band.sum(1).astype(np.int64), freqs_per_band.sum(0).astype(np.int64), freqs_per_band


class MelBandRoformer(_RoformerBase):
    """Drop-in for models.bs_roformer.MelBandRoformer (mel_band_roformer.py:324-361 constructor surface)."""
    norm_output = True
    has_final_norm = False
    mask_extra_linear = 1

    def __init__(self, dim, *, depth, stereo=False, num_stems=1, time_transformer_depth=2,
                 freq_transformer_depth=2, linear_transformer_depth=0, num_bands=60, dim_head=64, heads=8,
                 attn_dropout=0.1, ff_dropout=0.1, flash_attn=True, dim_freqs_in=1025, sample_rate=44100,
                 stft_n_fft=2048, stft_hop_length=512, stft_win_length=2048, stft_normalized=False,
                 stft_window_fn=None, mask_estimator_depth=1, multi_stft_resolution_loss_weight=1.,
                 multi_stft_resolutions_window_sizes=(4096, 2048, 1024, 512, 256), multi_stft_hop_size=147,
                 multi_stft_normalized=False, multi_stft_window_fn=None, match_input_audio_length=False,
                 mlp_expansion_factor=4, use_torch_checkpoint=False, skip_connection=False, seed=0):
        fi, nfpb, nbpf, fpb = mel_band_maps(sample_rate, stft_n_fft, num_bands, stereo)
        ch = 2 if stereo else 1
        self.freq_indices = torch.from_numpy(fi)                 # non-persistent buffers in the reference (:436-443)
        self.freqs_per_band = torch.from_numpy(fpb)
        self.num_freqs_per_band = torch.from_numpy(nfpb)
        self.num_bands_per_freq = torch.from_numpy(nbpf)
        self.match_input_audio_length = match_input_audio_length
        self._init_common(dim, depth, stereo, num_stems, time_transformer_depth, freq_transformer_depth,
                          linear_transformer_depth, dim_head, heads, flash_attn, stft_n_fft, stft_hop_length,
                          stft_win_length, stft_normalized, stft_window_fn, mask_estimator_depth,
                          mlp_expansion_factor, use_torch_checkpoint, skip_connection,
                          tuple(2 * int(f) * ch for f in nfpb), seed)

    def _extra_prepare(self, prep):
        dev = self._device
        fi = self.freq_indices.numpy()
        C = self.audio_channels
        nfs = (self.n_fft // 2 + 1) * C
        inv = -np.ones((nfs, 2), dtype=np.int32)
        fill = np.zeros(nfs, dtype=np.int64)
        for j, fs in enumerate(fi):                              # ascending j = scatter_add order (:610)
            if fill[fs] >= 2:
                raise NotImplementedError('a frequency bin covered by more than two mel bands')
            inv[fs, fill[fs]] = j
            fill[fs] += 1
        cnt = np.repeat(self.num_bands_per_freq.numpy(), C).astype(np.float32)   # '(f r) 1' (:612)
        prep['freq_idx'] = torch.from_numpy(fi.astype(np.int32)).to(dev)
        prep['inv'] = torch.from_numpy(inv.reshape(-1)).to(dev)
        prep['cnt'] = torch.from_numpy(np.maximum(cnt, 1e-8)).to(dev)
        prep['J'] = int(len(fi))

    def _features_buffer(self, ws, B, T):
        ws['feat'] = torch.empty(B * T, self._prepared['J'] * 2, device=self._device, dtype=torch.float32)
        return ws['feat']

    def _out_len(self, L, T):
        return L if self.match_input_audio_length else self.hop * (T - 1)      # torch.istft length=None (:505,622)

    def _gather_features(self, ws, prep, B, T):
        nfs = (self.n_fft // 2 + 1) * self.audio_channels
        call('sesa_gather_rows', _ptr(ws['spec']), _ptr(prep['freq_idx']), _ptr(ws['feat']), B * T, nfs, prep['J'], 2,
             _stream())

    def _mask_istft(self, ws, prep, out, B, C, T, out_len, st):
        call('sesa_mask_istft', _ptr(ws['spec']), _ptr(ws['mask']), _ptr(prep['inv']), _ptr(prep['cnt']), _ptr(out),
             _ptr(prep['window']), _ptr(ws['env']), _ptr(prep['twiddle']), B, self.num_stems, C, self.n_fft,
             self.hop, T, out_len, 1, prep['J'], st)
