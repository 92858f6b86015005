"""GPU: each kernel of the C-ABI against a plain torch fp32 statement of the same op."""
import ctypes
import math

import numpy as np
import pytest
import torch

from conftest import max_rel, snr_db

pytestmark = pytest.mark.gpu


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope='module')
def lib():
    from sesa_audio_separation_b200 import _lib
    _lib.require_cuda()
    return _lib


def _tw(n, dev):
    from sesa_audio_separation_b200.roformer import _twiddle
    return _twiddle(n).to(dev)


@pytest.mark.parametrize('n_fft,hop,C,L', [(2048, 441, 2, 441 * 30), (2048, 512, 1, 20000), (1024, 256, 2, 256 * 31),
                                           (8192, 1024, 2, 1024 * 15), (256, 64, 2, 5000)])
def test_stft_matches_torch(lib, n_fft, hop, C, L):
    dev = 'cuda'
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, C, L, generator=g)
    win = torch.hann_window(n_fft)
    ref = torch.stft(x.reshape(-1, L), n_fft=n_fft, hop_length=hop, window=win, return_complex=True)
    T, F = ref.shape[-1], n_fft // 2 + 1
    ref = torch.view_as_real(ref).reshape(3, C, F, T, 2).permute(0, 3, 2, 1, 4).contiguous()  # b t f c ri
    xd, wd, tw = x.to(dev), win.to(dev), _tw(n_fft, dev)   # keep every device buffer referenced
    spec = torch.empty(3 * T, F * C * 2, device=dev)
    lib.call('sesa_stft', P(xd), P(spec), P(wd), P(tw), 3, C, L, n_fft, hop, 0, F, S())
    got = spec.cpu().reshape(3, T, F, C, 2)
    assert max_rel(ref.numpy(), got.numpy()) < 2e-6
    # layout 1 (MDX23C planes)
    dim_f = n_fft // 2
    spec1 = torch.empty(3, C * 2, dim_f, T, device=dev)
    lib.call('sesa_stft', P(xd), P(spec1), P(wd), P(tw), 3, C, L, n_fft, hop, 1, dim_f, S())
    ref1 = ref[:, :, :dim_f].permute(0, 3, 4, 2, 1).reshape(3, C * 2, dim_f, T)
    assert max_rel(ref1.numpy(), spec1.cpu().numpy()) < 2e-6


@pytest.mark.parametrize('n_fft,hop,C,L,nst', [(2048, 441, 2, 441 * 30, 1), (2048, 512, 1, 20000, 2), (1024, 256, 2, 256 * 31, 2)])
def test_mask_istft_matches_torch(lib, n_fft, hop, C, L, nst):
    from sesa_audio_separation_b200.roformer import _istft_envelope
    dev = 'cuda'
    g = torch.Generator().manual_seed(2)
    B = 2
    x = torch.randn(B, C, L, generator=g)
    win = torch.hann_window(n_fft)
    z = torch.stft(x.reshape(-1, L), n_fft=n_fft, hop_length=hop, window=win, return_complex=True)
    F, T = z.shape[-2:]
    z = z.reshape(B, C, F, T)
    mask = torch.randn(nst, B, T, F, C, 2, generator=g)
    mc = torch.view_as_complex(mask).permute(1, 0, 4, 3, 2)            # b n c f t
    prod = z[:, None] * mc
    ref = torch.istft(prod.reshape(-1, F, T), n_fft=n_fft, hop_length=hop, window=win, length=L).reshape(B, nst, C, L)
    spec = torch.view_as_real(z).permute(0, 3, 2, 1, 4).contiguous().to(dev)    # b t f c ri
    out = torch.empty(B, nst, C, L, device=dev)
    env = _istft_envelope(win, n_fft, hop, T, L).to(dev)
    md, wd, tw = mask.to(dev), win.to(dev), _tw(n_fft, dev)
    lib.call('sesa_mask_istft', P(spec), P(md), None, None, P(out), P(wd), P(env),
             P(tw), B, nst, C, n_fft, hop, T, L, 0, 0, S())
    assert max_rel(ref.numpy(), out.cpu().numpy()) < 3e-6
    assert snr_db(ref.numpy(), out.cpu().numpy()) > 110


@pytest.mark.parametrize('B,nst', [(5, 1), (3, 4)])
def test_stft_and_mask_istft_are_batch_invariant_and_reproducible(lib, B, nst):
    """A chunk's spectrogram and waveform must not depend on the batch it is launched in (engine batches and chunk-range
    shards regroup chunks), nor on the arrival order of the iSTFT's two-contributor atomics: bitwise equality of a batched
    launch, one-by-one launches and a repeated launch, at a frame count that spans many frame groups."""
    from sesa_audio_separation_b200.roformer import _istft_envelope
    dev = 'cuda'
    n_fft, hop, C, L = 2048, 441, 2, 441 * 400
    T, F = 1 + L // hop, n_fft // 2 + 1
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(B, C, L, device=dev, generator=g)
    mask = torch.randn(nst, B * T, F * C * 2, device=dev, generator=g)
    win = torch.hann_window(n_fft)
    env = _istft_envelope(win, n_fft, hop, T, L).to(dev)
    wd, tw = win.to(dev), _tw(n_fft, dev)
    spec = torch.empty(B * T, F, C, 2, device=dev)
    out = torch.empty(B, nst, C, L, device=dev)
    lib.call('sesa_stft', P(x), P(spec), P(wd), P(tw), B, C, L, n_fft, hop, 0, F, S())
    lib.call('sesa_mask_istft', P(spec), P(mask), None, None, P(out), P(wd), P(env), P(tw), B, nst, C, n_fft, hop, T, L, 0, 0, S())
    out2 = torch.empty_like(out)
    lib.call('sesa_mask_istft', P(spec), P(mask), None, None, P(out2), P(wd), P(env), P(tw), B, nst, C, n_fft, hop, T, L, 0, 0, S())
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    for b in range(B):
        sb = torch.empty(T, F, C, 2, device=dev)
        ob = torch.empty(1, nst, C, L, device=dev)
        xb = x[b:b + 1].contiguous()
        mb = mask[:, b * T:(b + 1) * T].contiguous()
        lib.call('sesa_stft', P(xb), P(sb), P(wd), P(tw), 1, C, L, n_fft, hop, 0, F, S())
        lib.call('sesa_mask_istft', P(sb), P(mb), None, None, P(ob), P(wd), P(env), P(tw), 1, nst, C, n_fft, hop, T, L, 0, 0, S())
        torch.cuda.synchronize()
        assert torch.equal(sb, spec[b * T:(b + 1) * T]), b
        assert torch.equal(ob[0], out[b]), (b, float((ob[0] - out[b]).abs().max()))


@pytest.mark.parametrize('hop,C', [(441, 2), (512, 2), (441, 1)])
def test_full_size_stft_istft_round_trip(lib, hop, C):
    """Size-independent property at the BASELINE chunk size (352 800 samples, n_fft 2048): STFT followed by the fused
    mask + iSTFT with a unit mask reconstructs the chunk (windowed overlap-add with the exact envelope is the identity),
    reflect-padded edges included."""
    from sesa_audio_separation_b200.roformer import _istft_envelope
    dev = 'cuda'
    n_fft, L, B = 2048, 352800, 2
    T, F = 1 + L // hop, n_fft // 2 + 1
    out_len = min(L, hop * (T - 1))        # samples covered by whole hops (L itself when hop divides L)
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(B, C, L, device=dev, generator=g)
    win = torch.hann_window(n_fft)
    env = _istft_envelope(win, n_fft, hop, T, out_len).to(dev)
    wd, tw = win.to(dev), _tw(n_fft, dev)
    spec = torch.empty(B * T, F, C, 2, device=dev)
    mask = torch.zeros(1, B * T, F, C, 2, device=dev)
    mask[..., 0] = 1.0
    out = torch.empty(B, 1, C, out_len, device=dev)
    lib.call('sesa_stft', P(x), P(spec), P(wd), P(tw), B, C, L, n_fft, hop, 0, F, S())
    lib.call('sesa_mask_istft', P(spec), P(mask), None, None, P(out), P(wd), P(env), P(tw), B, 1, C, n_fft, hop, T, out_len, 0, 0, S())
    torch.cuda.synchronize()
    err = float((out[:, 0] - x[..., :out_len]).abs().max())
    print('round trip hop', hop, 'C', C, 'max abs err', err)
    assert err < 5e-6        # |x| reaches ~5: a few fp32 ulps through two 2048-point transforms


def _run_gemm(lib, A, W, bias, Cbuf, ep_kwargs, M, N, K, lda, ldc):
    from sesa_audio_separation_b200.roformer import _GroupTable, _epilogue
    tab = _GroupTable([dict(A=A.data_ptr(), W=W.data_ptr(), bias=bias.data_ptr() if bias is not None else 0,
                            C=Cbuf.data_ptr(), M=M, N=N, K=K, lda=lda, ldw=W.shape[1], ldc=ldc)], A.device)
    ep = _epilogue(**ep_kwargs)
    lib.call('sesa_gemm_simt', P(tab.dev), tab.n, tab.max_m, tab.max_n, ctypes.byref(ep), S())
    torch.cuda.synchronize()


@pytest.mark.parametrize('M,N,K', [(300, 512, 512), (257, 130, 10), (1000, 1544, 64), (77, 16, 2048), (513, 1032, 96)])
def test_gemm_simt_plain_bias_act(lib, M, N, K):
    dev = 'cuda'
    g = torch.Generator().manual_seed(3)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    for act, fn in ((0, lambda v: v), (1, torch.nn.functional.gelu), (2, torch.tanh)):
        Cb = torch.zeros(M, N, device=dev)
        _run_gemm(lib, A, W, b, Cb, dict(act=act), M, N, K, K, N)
        ref = fn(A.double() @ W.double().T + b.double())
        assert max_rel(ref.cpu().numpy(), Cb.cpu().numpy()) < 3e-6, act


def test_gemm_simt_rownorm_residual_glu_rotary(lib):
    dev = 'cuda'
    g = torch.Generator().manual_seed(4)
    M, N, K = 333, 264, 48
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    # rownorm + residual
    C0 = torch.randn(M, N, generator=g).to(dev)
    Cb = C0.clone()
    _run_gemm(lib, A, W, b, Cb, dict(rownorm=1, residual=1), M, N, K, K, N)
    ref = C0.double() + torch.nn.functional.normalize(A.double(), dim=-1) @ W.double().T + b.double()
    assert max_rel(ref.cpu().numpy(), Cb.cpu().numpy()) < 3e-6
    # GLU with interleaved rows
    half = N // 2
    Wi = torch.stack([W[:half], W[half:]], 1).reshape(N, K).contiguous()
    bi = torch.stack([b[:half], b[half:]], 1).reshape(N).contiguous()
    Cg = torch.zeros(M, half, device=dev)
    _run_gemm(lib, A, Wi, bi, Cg, dict(glu=1), M, N, K, K, half)
    ref = torch.nn.functional.glu(A.double() @ W.double().T + b.double(), dim=-1)
    assert max_rel(ref.cpu().numpy(), Cg.cpu().numpy()) < 3e-6
    # rotary on the first 128 columns, positions (m // 3) % 37
    from oracle.third_party import apply_rotary
    freqs = 1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64))
    ang = torch.einsum('i,j->ij', torch.arange(37, dtype=torch.float32), freqs)
    rot = torch.stack([ang.cos(), ang.sin()], -1).contiguous().to(dev)
    Cr = torch.zeros(M, N, device=dev)
    _run_gemm(lib, A, W, None, Cr, dict(rot=rot, rot_cols=128, rot_dim=64, pos_div=3, pos_mod=37), M, N, K, K, N)
    plain = (A.double() @ W.double().T).float().cpu()
    pos = (torch.arange(M) // 3) % 37
    full = torch.einsum('i,j->ij', pos.float(), freqs).repeat_interleave(2, -1)
    ref = plain.clone()
    for h in range(2):
        blk = plain[:, h * 64:(h + 1) * 64]
        rh = torch.stack((-blk[:, 1::2], blk[:, 0::2]), -1).reshape(M, 64)
        ref[:, h * 64:(h + 1) * 64] = blk * full.cos() + rh * full.sin()
    assert max_rel(ref.numpy(), Cr.cpu().numpy()) < 3e-6


@pytest.mark.parametrize('axis', [0, 1])
def test_attention_simt(lib, axis):
    dev = 'cuda'
    g = torch.Generator().manual_seed(5)
    B, T, F, H, dh = 2, 150, 20, 3, 64
    inner = H * dh
    ld = (3 * inner + H + 3) // 4 * 4
    qkv = torch.randn(B * T * F, ld, generator=g)
    qkv[:, :inner] *= dh ** -0.5          # q arrives pre-scaled by dh^-0.5 (folded into Wq)
    qd = qkv.to(dev)
    out = torch.zeros(B * T * F, inner, device=dev)
    if axis == 0:
        args = (B * F, T, F, T * F, 1, F)
    else:
        args = (B * T, F, 1, F, 0, 1)
    lib.call('sesa_attention_simt', P(qd), P(out), ld, inner, H, dh, *args, S())
    x = qkv.reshape(B, T, F, ld)
    q = x[..., :inner].reshape(B, T, F, H, dh)
    k = x[..., inner:2 * inner].reshape(B, T, F, H, dh)
    v = x[..., 2 * inner:3 * inner].reshape(B, T, F, H, dh)
    gate = x[..., 3 * inner:3 * inner + H].sigmoid()
    if axis == 0:
        sim = torch.einsum('bifhd,bjfhd->bfhij', q.double(), k.double())
        o = torch.einsum('bfhij,bjfhd->bifhd', sim.softmax(-1), v.double())
    else:
        sim = torch.einsum('btihd,btjhd->bthij', q.double(), k.double())
        o = torch.einsum('bthij,btjhd->btihd', sim.softmax(-1), v.double())
    ref = (o * gate.double()[..., None]).reshape(B * T * F, inner)
    print('attention axis', axis, 'max_rel', max_rel(ref.numpy(), out.cpu().numpy()))
    assert max_rel(ref.numpy(), out.cpu().numpy()) < 5e-6


def test_rmsnorm_gather_add(lib):
    dev = 'cuda'
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1000, 384, generator=g).to(dev)
    gamma = torch.randn(384, generator=g).to(dev)
    y = torch.empty_like(x)
    lib.call('sesa_rmsnorm', P(x), P(gamma), P(y), 1000, 384, S())
    ref = torch.nn.functional.normalize(x, dim=-1) * math.sqrt(384) * gamma
    assert max_rel(ref.cpu().numpy(), y.cpu().numpy()) < 1e-6
    src = torch.randn(50, 30, 2, generator=g).to(dev)
    idx = torch.randint(0, 30, (45,), generator=g).to(torch.int32).to(dev)
    dst = torch.empty(50, 45, 2, device=dev)
    lib.call('sesa_gather_rows', P(src), P(idx), P(dst), 50, 30, 45, 2, S())
    assert torch.equal(dst, src[:, idx.long()])
    a = torch.randn(5000, generator=g).to(dev)
    b = torch.randn(5000, generator=g).to(dev)
    ref = a + b
    lib.call('sesa_add_inplace', P(a), P(b), 5000, S())
    assert torch.equal(a, ref)


# ------------------------------------------------------------------ tensor-core (tcgen05) GEMM
def _bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float64)


def _tc_case(lib, M, N, K, nsplit, block_n, bias=True, act=0, rowscale=False, residual=False, glu=False, rot=None,
             planes_out=False, seed=0, cg=1):
    from sesa_audio_separation_b200 import tc
    from sesa_audio_separation_b200._lib import GemmEpilogue
    dev = 'cuda'
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) if bias else None
    rs = (torch.rand(M, generator=g) + 0.5) if rowscale else None
    ad, wd = a.to(dev), w.to(dev)
    ap = tc.alloc_planes(M, K, dev)
    assert K % 4 == 0
    tc.prep_rows(ad, M, K, K, ap, normalize=False)
    wp = tc.split_weight(wd)
    n_out = N // 2 if glu else N
    ldc = n_out + 4
    c0 = torch.randn(M, ldc, generator=g)
    cd = c0.to(dev).clone()
    pout = tc.alloc_planes(M, n_out, dev) if planes_out else None
    bd = b.to(dev) if bias else None
    rsd = rs.to(dev) if rowscale else None
    prob = dict(A=tc.planes_arg(ap), W=tc.planes_arg(wp), M=M, N=N, K=K, bias=bd.data_ptr() if bias else 0,
                rowscale=rsd.data_ptr() if rowscale else 0, C=(cd.data_ptr(), ldc),
                P=tc.planes_arg(pout) if planes_out else None)
    tab = tc.TcGemmTable([prob], dev, block_n=block_n, cta_group=cg if block_n == 256 else 1)
    rot_d = None
    ep = GemmEpilogue(0, act, 1 if residual else 0, 1 if glu else 0, 0, 0, 1, 1, None)
    if rot is not None:
        rot_cols, rot_dim, pos_div, pos_mod = rot
        ang = torch.einsum('i,j->ij', torch.arange(pos_mod, dtype=torch.float32),
                           1.0 / (10000 ** (torch.arange(0, rot_dim, 2).float() / rot_dim)))
        rot_t = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
        rot_d = rot_t.reshape(pos_mod, -1, 4).permute(1, 0, 2).contiguous().to(dev)    # quad-major layout of sesa_gemm_tc
        ep = GemmEpilogue(0, act, 1 if residual else 0, 0, rot_cols, rot_dim, pos_div, pos_mod, rot_d.data_ptr())
    tab.run(ep, nsplit=nsplit, out_planes=2)
    torch.cuda.synchronize()
    # reference in fp64 (operands rounded to bf16 in bf16 mode)
    A64, W64 = a.double(), w.double()
    if nsplit == 1:
        A64, W64 = _bf16_round(a), _bf16_round(w)
    y = A64 @ W64.T
    if rowscale:
        y = y * rs.double()[:, None]
    if bias:
        y = y + b.double()[None]
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = y.tanh()
    elif act == 3:
        y = y.sigmoid()
    if rot is not None:
        pos = (torch.arange(M) // pos_div) % pos_mod
        cs = rot_t.double()[pos]                                   # M, rot_dim/2, 2
        yr = y[:, :rot_cols].reshape(M, rot_cols // rot_dim, rot_dim // 2, 2)
        x1, x2 = yr[..., 0], yr[..., 1]
        c, s = cs[:, None, :, 0], cs[:, None, :, 1]
        y = torch.cat([torch.stack([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).reshape(M, rot_cols), y[:, rot_cols:]], dim=1)
    if glu:
        y = y[:, 0::2] * y[:, 1::2].sigmoid()
    if residual:
        y = y + c0[:, :n_out].double()
    got = cd.cpu()
    assert torch.equal(got[:, n_out:], c0[:, n_out:]), 'wrote outside the output columns'
    err = max_rel(y.numpy(), got[:, :n_out].numpy())
    if planes_out:
        pr = (pout[0].float() + pout[1].float()).cpu()[:, :n_out]
        perr = max_rel(y.numpy(), pr.numpy())
        assert perr < 5e-5, perr
    return err


@pytest.mark.parametrize('block_n', [256, 128])
@pytest.mark.parametrize('nsplit', [3, 1])
def test_gemm_tc_plain(lib, nsplit, block_n):
    for (M, N, K) in [(128, 256, 64), (300, 512, 512), (1000, 1544, 512), (257, 2048, 128), (513, 512, 2048)]:
        err = _tc_case(lib, M, N, K, nsplit, block_n, seed=M)
        print('gemm_tc', M, N, K, nsplit, block_n, err)
        assert err < (3e-5 if nsplit == 3 else 2e-5), (M, N, K, err)   # bf16 mode is compared with bf16-rounded operands


@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_tc_ragged_and_epilogues(lib, cg):
    # ragged N / K tails (mask-estimator and band-split shapes), every epilogue; cg = 2: CTA pairs (cta_group::2)
    assert _tc_case(lib, 801, 16, 2048, 3, 256, act=0, glu=True, cg=cg) < 3e-5
    assert _tc_case(lib, 801, 1032, 2048, 3, 256, glu=True, seed=3, cg=cg) < 3e-5
    assert _tc_case(lib, 333, 96, 520, 3, 128 if cg == 1 else 256, act=2, seed=4, cg=cg) < 3e-5
    assert _tc_case(lib, 640, 2048, 512, 3, 256, act=1, rowscale=True, planes_out=True, seed=5, cg=cg) < 3e-5
    assert _tc_case(lib, 640, 512, 2048, 3, 256, residual=True, planes_out=True, seed=6, cg=cg) < 3e-5
    assert _tc_case(lib, 62 * 9, 1536, 512, 3, 256, bias=False, rot=(1024, 64, 62, 9), planes_out=True, seed=7, cg=cg) < 3e-5
    assert _tc_case(lib, 62 * 9, 1536, 512, 3, 256, bias=False, rot=(1024, 64, 1, 62), planes_out=True, seed=8, cg=cg) < 3e-5
    assert _tc_case(lib, 200, 24, 8, 3, 256, seed=8, cg=cg) < 3e-5


@pytest.mark.parametrize('nsplit', [3, 1])
def test_gemm_tc_cta_pairs(lib, nsplit):
    for (M, N, K) in [(128, 256, 64), (300, 512, 512), (1000, 1544, 512), (257, 2048, 128), (513, 512, 2048), (5000, 1024, 512)]:
        err = _tc_case(lib, M, N, K, nsplit, 256, seed=M, cg=2)
        print('gemm_tc 2-CTA', M, N, K, nsplit, err)
        assert err < (3e-5 if nsplit == 3 else 2e-5), (M, N, K, err)


def test_gemm_tc_grouped(lib):
    """62-band style grouped launch: different N per group, one launch."""
    from sesa_audio_separation_b200 import tc
    from sesa_audio_separation_b200._lib import GemmEpilogue
    dev = 'cuda'
    g = torch.Generator().manual_seed(11)
    M, K = 500, 256
    Ns = [16, 48, 512, 520, 96, 1032]
    a = [torch.randn(M, K, generator=g) for _ in Ns]
    w = [torch.randn(n, K, generator=g) / 16 for n in Ns]
    keep, probs, outs = [], [], []
    for ai, wi, n in zip(a, w, Ns):
        ap = tc.alloc_planes(M, K, dev)
        ad = ai.to(dev)
        tc.prep_rows(ad, M, K, K, ap, normalize=False)
        wp = tc.split_weight(wi.to(dev))
        c = torch.zeros(M, n, device=dev)
        keep += [ap, ad, wp]
        outs.append(c)
        probs.append(dict(A=tc.planes_arg(ap), W=tc.planes_arg(wp), M=M, N=n, K=K, C=(c.data_ptr(), n)))
    tab = tc.TcGemmTable(probs, dev)
    tab.run(GemmEpilogue(0, 0, 0, 0, 0, 0, 1, 1, None), nsplit=3)
    torch.cuda.synchronize()
    for ai, wi, c in zip(a, w, outs):
        assert max_rel((ai.double() @ wi.double().T).numpy(), c.cpu().numpy()) < 3e-5


def test_prep_rows(lib):
    from sesa_audio_separation_b200 import tc
    dev = 'cuda'
    g = torch.Generator().manual_seed(12)
    M, D, H = 777, 512, 8
    x = torch.randn(M, D, generator=g) * 3
    gw, gb = torch.randn(H, D, generator=g) / 20, torch.randn(H, generator=g)
    xd, gwd, gbd = x.to(dev), gw.to(dev), gb.to(dev)
    planes = tc.alloc_planes(M, D, dev)
    gates = torch.zeros(M, 12, device=dev)
    inv = torch.zeros(M, device=dev)
    tc.prep_rows(xd, M, D, D, planes, True, gwd, gbd, gates, 12, inv)
    xn = torch.nn.functional.normalize(x.double(), dim=-1)
    rec = (planes[0].float() + planes[1].float()).cpu().double()
    assert max_rel(xn.numpy(), rec.numpy()) < 1e-5
    assert max_rel((xn @ gw.double().T + gb.double()).numpy(), gates[:, :H].cpu().numpy()) < 1e-5
    assert max_rel((1 / x.double().norm(dim=-1)).numpy(), inv.cpu().numpy()) < 1e-6


@pytest.mark.parametrize('nsplit', [3, 1])
@pytest.mark.parametrize('axis,B,T,F', [(0, 2, 150, 20), (1, 2, 150, 20), (0, 1, 801, 3), (1, 3, 33, 62), (1, 1, 5, 60),
                                        (1, 2, 7, 100), (0, 1, 64, 2),
                                        # more work items than resident CTAs (persistent kernel walks several items per
                                        # CTA), with one, three and four KV blocks per item, and packed tiles
                                        (0, 2, 41, 62), (0, 4, 200, 40), (0, 2, 130, 70), (1, 40, 62, 20)])
def test_attention_tc(lib, axis, B, T, F, nsplit):
    from sesa_audio_separation_b200 import tc
    dev = 'cuda'
    g = torch.Generator().manual_seed(5)
    H, dh = 3, 64
    inner = H * dh
    M = B * T * F
    qkv = torch.randn(M, 3 * inner, generator=g)
    qkv[:, :inner] *= 0.35
    gl = torch.randn(M, 8, generator=g)
    qd, gd = qkv.to(dev), gl.to(dev)
    planes = tc.alloc_planes(M, 3 * inner, dev)
    tc.prep_rows(qd, M, 3 * inner, 3 * inner, planes, False)
    out = tc.alloc_planes(M, inner, dev)
    if axis == 0:
        args = (B * F, T, F, T * F, 1, F)
    else:
        args = (B * T, F, 1, F, 0, 1)
    lib.call('sesa_attention_tc', P(planes), planes.shape[-1], planes.stride(0), P(gd), 8, P(out), out.shape[-1],
             out.stride(0), H, dh, *args, T if axis == 1 else 0, nsplit, 2, S())
    torch.cuda.synchronize()
    src = qkv if nsplit == 3 else planes[0].float().cpu()
    x = src.reshape(B, T, F, 3 * inner).double()
    q = x[..., :inner].reshape(B, T, F, H, dh)
    k = x[..., inner:2 * inner].reshape(B, T, F, H, dh)
    v = x[..., 2 * inner:].reshape(B, T, F, H, dh)
    gate = gl[:, :H].double().sigmoid().reshape(B, T, F, H)
    if axis == 0:
        sim = torch.einsum('bifhd,bjfhd->bfhij', q, k)
        o = torch.einsum('bfhij,bjfhd->bifhd', sim.softmax(-1), v)
    else:
        sim = torch.einsum('btihd,btjhd->bthij', q, k)
        o = torch.einsum('bthij,btjhd->btihd', sim.softmax(-1), v)
    ref = (o * gate[..., None]).reshape(M, inner)
    got = (out[0].float() + out[1].float()).cpu()
    err = max_rel(ref.numpy(), got.numpy())
    print('attention_tc axis', axis, (B, T, F), 'nsplit', nsplit, 'max_rel', err)
    assert err < (3e-5 if nsplit == 3 else 1e-2)


# ------------------------------------------------------------------ MDX23C building blocks
def _cl_planes(lib, x):
    """channels-last fp32 [B, T, F, C] on the device -> bf16 planes [2, B*T*F, round8(C)]"""
    from sesa_audio_separation_b200 import tc
    B, T, F, C = x.shape
    pl = tc.alloc_planes(B * T * F, C, x.device)
    lib.call('sesa_norm_act_split', P(x), 0, B, T * F, C, 1, C, None, None, None, 0, P(pl), pl.shape[-1], pl.stride(0), S())
    return pl


@pytest.mark.parametrize('B,T,F,cin,cout', [(2, 8, 128, 64, 96), (1, 16, 32, 24, 16), (2, 4, 256, 128, 256), (1, 8, 64, 16, 40)])
def test_conv3x3_implicit_gemm(lib, B, T, F, cin, cout):
    from sesa_audio_separation_b200 import tc
    from sesa_audio_separation_b200._lib import GemmEpilogue
    from sesa_audio_separation_b200.mdx23c import TAPS3, _r64
    dev = 'cuda'
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, T, F, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1)     # B T F cout
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    pl = _cl_planes(lib, xd)
    cp = _r64(cin)
    wt = torch.zeros(cout, 9, cp)
    wt[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    wp = tc.split_weight(wt.reshape(cout, 9 * cp).to(dev))
    out = torch.zeros(B * T * F, cout, device=dev)
    tab = tc.TcGemmTable([dict(A=tc.planes_arg(pl), W=tc.planes_arg(wp), M=B * T * F, N=cout, K=9 * cp, C=(out.data_ptr(), cout),
                               conv=dict(cin=cin, B=B, T=T, F=F, inT=T, inF=F, stride=1, taps=TAPS3))], dev)
    tab.run(GemmEpilogue(0, 0, 0, 0, 0, 0, 1, 1, None), nsplit=3)
    torch.cuda.synchronize()
    err = max_rel(ref.numpy(), out.cpu().reshape(B, T, F, cout).numpy())
    print('conv3x3', (B, T, F, cin, cout), err)
    assert err < 3e-5


def test_downscale_and_upscale_convs(lib):
    from sesa_audio_separation_b200 import tc
    from sesa_audio_separation_b200._lib import GemmEpilogue
    from sesa_audio_separation_b200.mdx23c import TAPS2, _r64
    dev = 'cuda'
    g = torch.Generator().manual_seed(4)
    B, T, F, cin, cout = 2, 8, 64, 48, 40
    x = torch.randn(B, cin, T, F, generator=g)
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    pl = _cl_planes(lib, xd)
    ep = GemmEpilogue(0, 0, 0, 0, 0, 0, 1, 1, None)
    # Conv2d(kernel = stride = 2)
    w = torch.randn(cout, cin, 2, 2, generator=g) / math.sqrt(4 * cin)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), stride=2).permute(0, 2, 3, 1)
    cp = _r64(cin)
    wt = torch.zeros(cout, 4, cp)
    wt[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 4, cin)
    wp = tc.split_weight(wt.reshape(cout, 4 * cp).to(dev))
    out = torch.zeros(B * (T // 2) * (F // 2), cout, device=dev)
    tab = tc.TcGemmTable([dict(A=tc.planes_arg(pl), W=tc.planes_arg(wp), M=B * (T // 2) * (F // 2), N=cout, K=4 * cp,
                               C=(out.data_ptr(), cout),
                               conv=dict(cin=cin, B=B, T=T // 2, F=F // 2, inT=T, inF=F, stride=2, taps=TAPS2))], dev)
    tab.run(ep, nsplit=3)
    torch.cuda.synchronize()
    assert max_rel(ref.numpy(), out.cpu().reshape(B, T // 2, F // 2, cout).numpy()) < 3e-5
    # ConvTranspose2d(kernel = stride = 2), written into the first half of a concat buffer
    wT = torch.randn(cin, cout, 2, 2, generator=g) / math.sqrt(cin)
    refT = torch.nn.functional.conv_transpose2d(x.double(), wT.double(), stride=2).permute(0, 2, 3, 1)   # B 2T 2F cout
    cat = torch.full((B * 2 * T * 2 * F, 2 * cout), 7.0, device=dev)
    probs, keep = [], []
    for (kh, kw) in TAPS2:
        wpk = tc.split_weight(wT[:, :, kh, kw].t().contiguous().to(dev))
        keep.append(wpk)
        probs.append(dict(A=tc.planes_arg(pl), W=tc.planes_arg(wpk), M=B * T * F, N=cout, K=cin, C=(cat.data_ptr(), 2 * cout),
                          row_map=(F, kh, kw)))
    tc.TcGemmTable(probs, dev).run(ep, nsplit=3)
    torch.cuda.synchronize()
    got = cat.cpu().reshape(B, 2 * T, 2 * F, 2 * cout)
    assert max_rel(refT.numpy(), got[..., :cout].numpy()) < 3e-5
    assert torch.all(got[..., cout:] == 7.0)


def test_instnorm_and_norm_act_split(lib):
    from sesa_audio_separation_b200 import tc
    dev = 'cuda'
    g = torch.Generator().manual_seed(6)
    B, T, F, C = 2, 6, 40, 24
    x = torch.randn(B, T, F, C, generator=g) * 2 + 0.7
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xd, gd, bd = x.to(dev), gamma.to(dev), beta.to(dev)
    scratch = torch.zeros(2 * B * C, device=dev, dtype=torch.float64)
    stats = torch.zeros(B, C, 2, device=dev)
    lib.call('sesa_instnorm_stats', P(xd), 0, B, T * F, C, 1, C, P(scratch), P(stats), 1e-5, S())
    xn = torch.nn.functional.instance_norm(x.permute(0, 3, 1, 2).double(), weight=gamma.double(), bias=beta.double(), eps=1e-5)
    ref = torch.nn.functional.gelu(xn).permute(0, 2, 3, 1)          # B T F C
    pl = tc.alloc_planes(B * T * F, C, dev)
    lib.call('sesa_norm_act_split', P(xd), 0, B, T * F, C, 1, C, P(stats), P(gd), P(bd), 1, P(pl), pl.shape[-1], pl.stride(0), S())
    got = (pl[0].float() + pl[1].float()).cpu().reshape(B, T, F, C)
    assert max_rel(ref.numpy(), got.numpy()) < 2e-5
    # mode 1: channel-major planes [b][t][c][f]
    pt = tc.alloc_planes(B * T * C, F, dev)
    lib.call('sesa_norm_act_split', P(xd), 1, B, T, C, F, C, P(stats), P(gd), P(bd), 1, P(pt), pt.shape[-1], pt.stride(0), S())
    gott = (pt[0].float() + pt[1].float()).cpu().reshape(B, T, C, F).permute(0, 1, 3, 2)
    assert max_rel(ref.numpy(), gott.numpy()) < 2e-5
    # layout 1 statistics + mode 2 on channel-major data
    xc = x.permute(0, 1, 3, 2).contiguous().to(dev)                  # B T C F
    stats2 = torch.zeros(B, C, 2, device=dev)
    lib.call('sesa_instnorm_stats', P(xc), 1, B, T, C, F, 0, P(scratch), P(stats2), 1e-5, S())
    assert max_rel(stats.cpu().numpy(), stats2.cpu().numpy()) < 1e-5
    pc = tc.alloc_planes(B * T * C, F, dev)
    lib.call('sesa_norm_act_split', P(xc), 2, B, T, C, F, 0, P(stats2), P(gd), P(bd), 1, P(pc), pc.shape[-1], pc.stride(0), S())
    gotc = (pc[0].float() + pc[1].float()).cpu().reshape(B, T, C, F).permute(0, 1, 3, 2)
    assert max_rel(ref.numpy(), gotc.numpy()) < 2e-5
    # transpose-add
    gq = torch.randn(B * T, C, F, generator=g)
    x2 = xd.clone()
    gqd = gq.to(dev)
    lib.call('sesa_transpose_add', P(x2), P(gqd), B * T, F, C, C, S())
    assert torch.allclose(x2.cpu(), x + gq.reshape(B, T, C, F).permute(0, 1, 3, 2))


@pytest.mark.parametrize('length,L,ov,nc', [(4000, 1000, 1, 2), (40000, 4000, 4, 2), (40004, 4000, 4, 1), (52000, 8000, 2, 4),
                                             (8 * 4410 + 1236, 4416, 4, 2), (100000, 4800, 8, 2), (1000, 1000, 2, 2)])
def test_overlap_add_region_kernel_is_bit_identical_to_the_scalar_gather(lib, length, L, ov, nc):
    """The product path (no counter -> region kernel with block-uniform chunk lists) must reproduce the scalar gather
    bit for bit: chunk order, separate multiply and add, divide by the window sum (utils.py:439-464)."""
    from sesa_audio_separation_b200.plan import make_plan, windowing_array
    dev = 'cuda'
    plan = make_plan(length, L, ov, 1)
    g = torch.Generator(device=dev).manual_seed(length + L + ov)
    y = torch.randn(plan.n_chunks, nc, L, device=dev, generator=g)
    starts = torch.tensor(plan.starts, dtype=torch.int64, device=dev)
    lens = torch.tensor(plan.lens, dtype=torch.int64, device=dev)
    kinds = torch.tensor(plan.kinds, dtype=torch.int32, device=dev)
    window = windowing_array(L, plan.fade).to(dev)
    crop = plan.border if plan.pad else 0
    a = torch.full((nc, length), 7.0, device=dev)
    b = torch.full((nc, length), 9.0, device=dev)
    counter = torch.empty(plan.padded, device=dev)
    lib.call('sesa_overlap_add', P(y), P(starts), P(lens), P(kinds), plan.n_chunks, plan.step, L, plan.fade, P(window), 1, nc,
             plan.padded, crop, length, P(a), P(counter), S())
    lib.call('sesa_overlap_add', P(y), P(starts), P(lens), P(kinds), plan.n_chunks, plan.step, L, plan.fade, P(window), 1, nc,
             plan.padded, crop, length, P(b), None, S())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # and against a plain torch statement of the reference loop
    res = torch.zeros(nc, plan.padded, device=dev)
    cnt = torch.zeros(plan.padded, device=dev)
    for k in range(plan.n_chunks):
        s, n = plan.starts[k], plan.lens[k]
        w = window.clone()
        if plan.kinds[k] == 1:
            w[:plan.fade] = 1
        elif plan.kinds[k] == 2:
            w[-plan.fade:] = 1
        res[:, s:s + n] += y[k, :, :n] * w[:n]
        cnt[s:s + n] += w[:n]
    ref = torch.nan_to_num((res / cnt)[:, crop:crop + length], nan=0.0)
    assert torch.equal(ref, b)


@pytest.mark.parametrize('length,L,ov,bs,nc', [(4000, 1000, 1, 1, 2), (40000, 4000, 4, 2, 2), (40004, 4000, 4, 1, 1), (52000, 8000, 2, 3, 4),
                                               (30011, 1001, 3, 2, 2), (9999, 1000, 8, 4, 3), (300, 1000, 4, 2, 2),
                                               (60000, 1600, 16, 2, 1)])
def test_overlap_accumulate_streams_to_the_same_bits(lib, length, L, ov, bs, nc):
    """sesa_overlap_accumulate (the product path: one call per engine batch, running sums continued in `partial`) is
    bit-identical to the one-shot gather sesa_overlap_add for every batching, vectorised and scalar geometries, partial
    slabs and output windows that start mid-track, and never reads `partial` before writing it (NaN-filled here)."""
    from sesa_audio_separation_b200.plan import make_plan, windowing_array
    dev = 'cuda'
    plan = make_plan(length, L, ov, bs)
    g = torch.Generator(device=dev).manual_seed(length + L + ov)
    y = torch.randn(plan.n_chunks, nc, L, device=dev, generator=g)
    starts = torch.tensor(plan.starts, dtype=torch.int64, device=dev)
    lens = torch.tensor(plan.lens, dtype=torch.int64, device=dev)
    kinds = torch.tensor(plan.kinds, dtype=torch.int32, device=dev)
    window = windowing_array(L, plan.fade).to(dev)
    crop = plan.border if plan.pad else 0
    ref = torch.full((nc, length), 7.0, device=dev)
    counter = torch.empty(plan.padded, device=dev)
    lib.call('sesa_overlap_add', P(y), P(starts), P(lens), P(kinds), plan.n_chunks, plan.step, L, plan.fade, P(window), 1, nc,
             plan.padded, crop, length, P(ref), P(counter), S())
    span = -(-L // plan.step)
    ld = (length + 3) // 4 * 4
    for eb in (1, 3, plan.n_chunks):
        out = torch.full((nc, ld), float('nan'), device=dev)
        partial = torch.full((nc, (plan.padded + 3) // 4 * 4 + 4), float('nan'), device=dev)
        for k in range(0, plan.n_chunks, eb):
            nb = min(eb, plan.n_chunks - k)
            lib.call('sesa_overlap_accumulate', P(y[k:k + nb]), k, nb, P(starts), P(lens), P(kinds), plan.n_chunks, plan.step, L,
                     plan.fade, P(window), 1, nc, plan.padded, k, k + nb + span - 1, P(partial), partial.shape[1], 0, crop,
                     length, P(out), ld, 0, length, S())
        torch.cuda.synchronize()
        assert torch.equal(out[:, :length], ref), (eb, float((out[:, :length] - ref).abs().max()))
    # a shard's view: chunks [lo, n) only, partial slab and output window starting at the shard's first region, the
    # earlier chunks' raw sums handed in as the halo
    if plan.n_chunks >= 2 * span and plan.step % 4 == 0:
        lo = plan.n_chunks // 2
        p0 = plan.starts[lo]
        part_a = torch.full((nc, (plan.padded + 3) // 4 * 4 + 4), float('nan'), device=dev)
        out_a = torch.full((nc, ld), float('nan'), device=dev)
        lib.call('sesa_overlap_accumulate', P(y[:lo]), 0, lo, P(starts), P(lens), P(kinds), plan.n_chunks, plan.step, L, plan.fade,
                 P(window), 1, nc, plan.padded, 0, lo + span - 1, P(part_a), part_a.shape[1], 0, crop, length, P(out_a), ld, 0,
                 length, S())
        halo_p1 = min(plan.padded, (lo + span - 1) * plan.step)
        part_b = torch.full((nc, (plan.padded - p0 + 3) // 4 * 4 + 4), float('nan'), device=dev)
        part_b[:, :halo_p1 - p0] = part_a[:, p0:halo_p1]
        q0 = max(p0 - crop, 0)
        out_b = torch.full((nc, (length - q0 + 3) // 4 * 4), float('nan'), device=dev)
        lib.call('sesa_overlap_accumulate', P(y[lo:]), lo, plan.n_chunks - lo, P(starts), P(lens), P(kinds), plan.n_chunks,
                 plan.step, L, plan.fade, P(window), 1, nc, plan.padded, lo, plan.n_chunks + span, P(part_b), part_b.shape[1], p0,
                 crop, length, P(out_b), out_b.shape[1], q0, out_b.shape[1], S())
        torch.cuda.synchronize()
        assert torch.equal(out_a[:, :q0], ref[:, :q0])
        assert torch.equal(out_b[:, :length - q0], ref[:, q0:])


def test_pad_reflect_slice_equals_the_full_pad(lib):
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(3)
    C, length, border = 2, 5003, 750
    mix = torch.randn(C, length, device=dev, generator=g)
    full = torch.empty(C, length + 2 * border, device=dev)
    lib.call('sesa_pad_reflect', P(mix), P(full), C, length, border, border, S())
    assert torch.equal(full, torch.nn.functional.pad(mix[None], (border, border), mode='reflect')[0])
    for p0, p1 in ((0, 1000), (0, 2100), (1000, 3000), (4500, length + 2 * border), (5500, length + 2 * border), (0, length + 2 * border)):
        idx = [abs(j) if j < length else 2 * (length - 1) - j for j in (p0 - border, p1 - 1 - border)]
        m0, m1 = min(idx), max(idx)
        if p0 - border < 0:
            m0 = 0
        if p1 - 1 - border >= length:
            m1 = length - 1
        win = mix[:, m0:m1 + 1].contiguous()
        out = torch.empty(C, p1 - p0, device=dev)
        lib.call('sesa_pad_reflect_slice', P(win), win.shape[1], m0, P(out), C, length, border, p0, p1 - p0, S())
        assert torch.equal(out, full[:, p0:p1]), (p0, p1)
    with pytest.raises(lib.SesaError):      # a window that does not hold the needed samples is refused, not read out of bounds
        win = mix[:, 100:200].contiguous()
        lib.call('sesa_pad_reflect_slice', P(win), 100, 100, P(out), C, length, border, 0, 1000, S())


def test_tta_kernels_match_numpy_statement(lib):
    """utils.apply_tta's arithmetic (utils.py:271-290) on the device: augmented mixes and the += / -= / /= 3 sequence."""
    dev = 'cuda'
    rng = np.random.default_rng(1)
    mix = rng.standard_normal((2, 12345)).astype(np.float32)
    m = torch.from_numpy(mix).to(dev)
    sw, ng = torch.empty_like(m), torch.empty_like(m)
    lib.call('sesa_tta_variants', P(m), P(sw), P(ng), 2, mix.shape[1], S())
    assert np.array_equal(sw.cpu().numpy(), mix[::-1].copy()) and np.array_equal(ng.cpu().numpy(), -1.0 * mix.copy())
    est = [rng.standard_normal((3, 2, 12345)).astype(np.float32) for _ in range(3)]
    want = est[0].copy()
    for n in range(3):
        want[n] += est[1][n][::-1].copy()
        want[n] -= est[2][n]
        want[n] /= 3
    d = [torch.from_numpy(e).to(dev) for e in est]
    out = torch.empty_like(d[0])
    lib.call('sesa_tta_combine', P(d[0]), P(d[1]), P(d[2]), P(out), 3, 2, 12345, S())
    assert np.array_equal(out.cpu().numpy(), want)


def test_ensemble_wave_kernel_matches_numpy_on_float64(lib):
    """ensemble.py:172-183 on stems resident on the GPU: float64 accumulation in input order = numpy on the float64
    buffers the reference reads; the float32 result is the correctly rounded float64 one."""
    import sesa_audio_separation_b200 as sesa
    rng = np.random.default_rng(2)
    stems = [(rng.standard_normal((2, 50001)) * 0.3).astype(np.float32) for _ in range(4)]
    dev_stems = [torch.from_numpy(s).cuda() for s in stems]
    f64 = np.stack([s.astype(np.float64) for s in stems], 0)
    for n in (3, 4):
        got = sesa.ensemble_waveforms(dev_stems[:n], 'avg_wave').cpu().numpy()
        assert np.array_equal(got, np.mean(f64[:n], axis=0).astype(np.float32))
        w = [1.0, 2.0, 0.5, 3.0][:n]
        w32 = np.array(w, dtype=np.float32)
        w32 /= w32.sum()
        got = sesa.ensemble_waveforms(dev_stems[:n], 'avg_wave', w).cpu().numpy()
        assert np.array_equal(got, np.average(f64[:n], axis=0, weights=w32).astype(np.float32))
        assert np.array_equal(sesa.ensemble_waveforms(dev_stems[:n], 'median_wave').cpu().numpy(),
                              np.median(f64[:n], axis=0).astype(np.float32))
        assert np.array_equal(sesa.ensemble_waveforms(dev_stems[:n], 'max_wave').cpu().numpy(), np.max(f64[:n], axis=0).astype(np.float32))
        assert np.array_equal(sesa.ensemble_waveforms(dev_stems[:n], 'min_wave').cpu().numpy(), np.min(f64[:n], axis=0).astype(np.float32))


@pytest.mark.parametrize('dim,out_planes', [(384, 2), (512, 2), (64, 1)])
def test_rmsnorm_planes_equals_the_two_kernel_sequence(lib, dim, out_planes):
    """sesa_rmsnorm_planes (Mel per-transformer output norm + operand preparation in one pass) is bit-identical to
    sesa_rmsnorm followed by sesa_prep_rows(normalize=2), and matches the torch statement of RMSNorm."""
    from sesa_audio_separation_b200 import tc
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(dim)
    rows, slots = 1237, 4
    x = torch.randn(rows, dim, device=dev, generator=g) * 3
    gamma = 1 + 0.1 * torch.randn(dim, device=dev, generator=g)
    a = x.clone()
    pa = tc.alloc_planes(rows, dim, dev)
    sa = torch.full((rows, slots), 7.0, device=dev)
    lib.call('sesa_rmsnorm', P(a), P(gamma), P(a), rows, dim, S())
    tc.prep_rows(a, rows, dim, dim, pa, 2, rowinv=sa, ss_slots=slots, out_planes=out_planes)
    b = x.clone()
    pb = tc.alloc_planes(rows, dim, dev)
    sb = torch.full((rows, slots), 9.0, device=dev)
    lib.call('sesa_rmsnorm_planes', P(b), P(gamma), rows, dim, P(pb), pb.shape[-1], pb.stride(0), out_planes, P(sb), slots, S())
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(sa, sb)
    assert torch.equal(pa[:out_planes], pb[:out_planes])
    ref = torch.nn.functional.normalize(x, dim=-1) * dim ** 0.5 * gamma
    assert max_rel(ref.cpu().numpy(), b.cpu().numpy()) < 1e-6


@pytest.mark.parametrize('B,T,F,C', [(2, 8, 64, 96), (1, 16, 256, 128), (2, 4, 40, 72)])
def test_transpose_add_stats_equals_the_two_pass_sequence(lib, B, T, F, C):
    """x += tdf(x)^T with the InstanceNorm statistics of the sum gathered in the same pass: x is bit-identical to
    sesa_transpose_add, the (mean, rstd) pairs agree with sesa_instnorm_stats over the updated tensor and with torch."""
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(B * 1000 + F)
    x = torch.randn(B * T * F, C, device=dev, generator=g) * 2 + 0.3
    gt = torch.randn(B * T * C, F, device=dev, generator=g)
    a = x.clone()
    lib.call('sesa_transpose_add', P(a), P(gt), B * T, F, C, C, S())
    scratch = torch.zeros(2 * B * C, device=dev, dtype=torch.float64)
    sa = torch.empty(B, C, 2, device=dev)
    lib.call('sesa_instnorm_stats', P(a), 0, B, T * F, C, 1, C, P(scratch), P(sa), 1e-5, S())
    b = x.clone()
    sb = torch.empty(B, C, 2, device=dev)
    lib.call('sesa_transpose_add_stats', P(b), P(gt), B, T, F, C, C, P(scratch), P(sb), 1e-5, S())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.allclose(sa, sb, rtol=2e-6, atol=1e-7)
    ref = a.view(B, T * F, C).double()
    mean = ref.mean(1)
    rstd = 1.0 / torch.sqrt(ref.var(1, unbiased=False) + 1e-5)
    assert torch.allclose(sb[..., 0].double(), mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sb[..., 1].double(), rstd, rtol=1e-5)
