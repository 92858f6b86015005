"""GPU: model forward and demix parity against the committed golden vectors (produced by the
unmodified reference) and against the CPU oracle, all through the C-ABI."""
import os
import numpy as np
import pytest
import torch

from conftest import golden, max_rel, parity_report, snr_db
from oracle import demix as odemix
from oracle.cases import CASES, DEMIX_IDENTITY_CASES, DEMIX_MODEL_CASES, FLOW_CASES, make_input
from oracle.weights import fill_state_dict, synth_mix

pytestmark = pytest.mark.gpu

# fp32 parity gate of BASELINE.json's north_star
FP32_MAX_REL = 1e-4
FP32_SNR_DB = 60.0


def build(case):
    import sesa_audio_separation_b200 as sesa
    kind, cfg = case['kind'], dict(case['cfg'])
    if kind == 'bs_roformer':
        if 'freqs_per_bands' in cfg:
            cfg['freqs_per_bands'] = tuple(cfg['freqs_per_bands'])
        m = sesa.BSRoformer(**cfg)
    elif kind == 'mel_band_roformer':
        m = sesa.MelBandRoformer(**cfg)
    else:
        from sesa_audio_separation_b200.mdx23c import TFC_TDF_net
        m = TFC_TDF_net(sesa.ConfigDict(cfg))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, case['seed'])
    m.load_state_dict(sd)
    return m.eval().to('cuda'), sd


@pytest.mark.parametrize('name', list(CASES))
def test_forward_matches_reference_golden(manifest, name):
    case = CASES[name]
    model, sd = build(case)
    assert {k: list(v.shape) for k, v in sd.items()} == manifest[name]['shapes']
    y = model(make_input(case).cuda()).cpu().numpy()
    ref = golden(name)['y']
    assert y.shape == ref.shape
    print(name, 'max_rel', max_rel(ref, y), 'snr', snr_db(ref, y))
    assert max_rel(ref, y) <= FP32_MAX_REL
    assert snr_db(ref, y) >= FP32_SNR_DB


class Identity:
    pass


def _identity_model():
    from sesa_audio_separation_b200.module import KernelModule

    class Ident(KernelModule):
        def forward(self, x, **kw):
            return x.clone()
    return Ident()


def test_demix_identity_bit_exact_and_counter():
    import sesa_audio_separation_b200 as sesa
    g = golden('demix_identity')
    for i, (length, L, ov, bs) in enumerate(DEMIX_IDENTITY_CASES):
        cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=ov, batch_size=bs),
                                   training=dict(instruments=['a'], target_instrument='a')))
        mix = synth_mix(length, 2, seed=100 + i)
        for eb in (1, 3):
            eng = sesa.DemixEngine(cfg, _identity_model(), 'cuda', engine_batch=eb)
            est, cnt = eng.run(mix, return_counter=True)
            assert np.array_equal(est[0], g[f'case{i}']), (i, length, L, ov, bs, eb)
            _, ocnt = odemix.demix(mix, lambda a: a, L, ov, bs, 1, return_counter=True)
            assert np.array_equal(cnt, ocnt[0, 0]), ('counter', i)


@pytest.mark.parametrize('name', list(DEMIX_MODEL_CASES))
def test_demix_matches_reference_golden(name):
    import sesa_audio_separation_b200 as sesa
    dc = DEMIX_MODEL_CASES[name]
    case = CASES[dc['model']]
    model, _ = build(case)
    cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=dc['chunk_size']),
                               inference=dict(num_overlap=dc['num_overlap'], batch_size=dc['batch_size']),
                               training=dict(instruments=dc['instruments'], target_instrument=dc['target'])))
    mix = synth_mix(dc['length'], 2, seed=dc['seed'])
    res = sesa.demix(cfg, model, mix, 'cuda', case['kind'], engine_batch=3)
    g = golden(name)
    assert list(res.keys()) == list(g.keys())
    for k in res:
        print(name, k, 'max_rel', max_rel(g[k], res[k]), 'snr', snr_db(g[k], res[k]))
        assert res[k].shape == g[k].shape
        assert max_rel(g[k], res[k]) <= FP32_MAX_REL
        assert snr_db(g[k], res[k]) >= FP32_SNR_DB


def test_progress_protocol(capsys):
    import sesa_audio_separation_b200 as sesa
    cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=1000), inference=dict(num_overlap=2, batch_size=1),
                               training=dict(instruments=['a'], target_instrument='a')))
    backend = type('B', (), {'model': _identity_model()})()
    sesa.demix_pytorch_optimized(cfg, backend, synth_mix(5000, 2, seed=1), 'cuda')
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('[SESA_PROGRESS]')]
    vals = [int(l[len('[SESA_PROGRESS]'):]) for l in lines]
    assert vals[-1] == 100 and vals == sorted(vals)


def test_full_size_bs_roformer_chunk_vs_oracle_on_gpu():
    """BASELINE C2 model (dim 512, depth 12, 62 bands) on one full 352800-sample chunk against the oracle
    restatement evaluated in true fp32 on the same GPU (TF32 off) — the oracle itself is pinned on CPU."""
    import sesa_audio_separation_b200 as sesa
    from oracle import roformer as orof
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = dict(dim=512, depth=12, stereo=True, num_stems=1, time_transformer_depth=1, freq_transformer_depth=1,
               dim_head=64, heads=8, stft_n_fft=2048, stft_hop_length=441, stft_win_length=2048,
               mask_estimator_depth=2)
    model = sesa.BSRoformer(**cfg)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=3)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    x = torch.from_numpy(synth_mix(352800, 2, seed=9))[None].cuda()
    y = model(x).cpu().numpy()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = orof.bs_roformer_forward(sd_gpu, cfg, x).cpu().numpy()
    print('full-size BS chunk: max_rel', max_rel(ref, y), 'snr', snr_db(ref, y))
    assert max_rel(ref, y) <= FP32_MAX_REL
    assert snr_db(ref, y) >= FP32_SNR_DB


def test_full_size_mdx23c_chunk_vs_oracle_on_gpu():
    """BASELINE C1 model (MDX23C vocals, 112 M parameters) on one full 261120-sample chunk against the oracle
    restatement evaluated in true fp32 on the same GPU (TF32 off)."""
    import os
    import sesa_audio_separation_b200 as sesa
    from conftest import ROOT
    from oracle import mdx23c as omdx
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, cfg = sesa.get_model_from_config('mdx23c', os.path.join(ROOT, 'configs', 'config_vocals_mdx23c.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=21)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    x = torch.from_numpy(synth_mix(261120, 2, seed=22))[None].cuda()
    y = model(x).cpu().numpy()
    ocfg = dict(audio=dict(cfg.audio), model=dict(cfg.model), num_target_instruments=2)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = omdx.mdx23c_forward(sd_gpu, ocfg, x).cpu().numpy()
    print('full-size MDX23C chunk: max_rel', max_rel(ref, y), 'snr', snr_db(ref, y))
    assert y.shape == ref.shape == (1, 2, 2, 261120)
    assert max_rel(ref, y) <= FP32_MAX_REL
    assert snr_db(ref, y) >= FP32_SNR_DB


@pytest.mark.parametrize('name', ['bs_small', 'mel_small'])
def test_forward_is_batch_invariant(name):
    """A chunk's output must not depend on the batch it is launched in (engine_batch is a throughput knob only, and
    chunk-range sharding across GPUs regroups chunks): bitwise equality of batched vs one-by-one forwards."""
    case = CASES[name]
    model, _ = build(case)
    x = make_input(case).cuda()
    x = torch.cat([x, x.flip(0) * 0.5, x * 0.25], 0)
    y = model(x).clone()
    for i in range(x.shape[0]):
        yi = model(x[i:i + 1])
        assert torch.equal(yi[0], y[i]), (name, i, float((yi[0] - y[i]).abs().max()))


def test_cli_inference_end_to_end(tmp_path, capsys):
    """python -m sesa_audio_separation_b200.inference on a folder: the reference CLI flow (inference_pytorch.py:189-386)
    with its output naming, --extract_instrumental and the [SESA_PROGRESS] protocol."""
    import yaml
    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200.audio_io import load_audio, write_audio
    from sesa_audio_separation_b200.inference import proc_folder
    case = CASES['bs_small']
    mcfg = dict(case['cfg'])
    if 'freqs_per_bands' in mcfg:
        mcfg['freqs_per_bands'] = tuple(mcfg['freqs_per_bands'])
    L = 441 * 40
    cfg = dict(audio=dict(chunk_size=L, sample_rate=44100, num_channels=2), model=mcfg,
               training=dict(instruments=['vocals', 'other'], target_instrument='vocals', use_amp=False),
               inference=dict(batch_size=1, num_overlap=2))
    cfg_path = tmp_path / 'cfg.yaml'
    cfg_path.write_text(yaml.dump(cfg))
    model, config = sesa.get_model_from_config('bs_roformer', str(cfg_path))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=31)
    ckpt = tmp_path / 'model.ckpt'
    torch.save({'state_dict': sd}, str(ckpt))
    indir, outdir = tmp_path / 'in', tmp_path / 'out'
    indir.mkdir()
    mix = synth_mix(L * 3 + 500, 2, seed=32)
    write_audio(str(indir / 'song.wav'), mix.T, 44100, subtype='FLOAT')
    written = proc_folder(['--model_type', 'bs_roformer', '--config_path', str(cfg_path), '--start_check_point', str(ckpt),
                           '--input_folder', str(indir), '--store_dir', str(outdir), '--extract_instrumental',
                           '--export_format', 'wav FLOAT'])
    names = sorted(os.path.basename(w) for w in written)
    assert names == ['song.wav_instrumental.wav', 'song.wav_vocals.wav']
    out = capsys.readouterr().out
    assert '[SESA_PROGRESS]100' in out
    model.load_state_dict(sd)
    ref = sesa.demix(config, model.eval().to('cuda'), mix, 'cuda', 'bs_roformer')['vocals']
    voc, _ = load_audio(str(outdir / 'song.wav_vocals.wav'), 44100)
    ins, _ = load_audio(str(outdir / 'song.wav_instrumental.wav'), 44100)
    assert np.array_equal(voc, ref)
    assert np.allclose(ins, mix - ref, atol=1e-7)


@pytest.mark.parametrize('name', list(CASES))
def test_bf16_mode_snr_gate(name):
    """bf16 mode of BASELINE.json (single bf16 MMA per product, fp32 accumulate): SNR >= 40 dB vs the reference goldens."""
    case = CASES[name]
    model, _ = build(case)
    model.set_precision('bf16')
    y = model(make_input(case).cuda()).cpu().numpy()
    ref = golden(name)['y']
    print(name, 'bf16 mode: max_rel', max_rel(ref, y), 'snr', snr_db(ref, y))
    assert snr_db(ref, y) >= 40.0


def test_full_size_mel_band_roformer_chunk_vs_oracle_on_gpu():
    """BASELINE C3 model (Mel-Band-RoFormer dim 384, depth 6, 60 mel bands, 4 stems, 832.6 M parameters) on one
    352800-sample chunk against the oracle restatement evaluated on the CPU (the pinned oracle itself)."""
    import sesa_audio_separation_b200 as sesa
    from conftest import ROOT
    from oracle import roformer as orof
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, cfg = sesa.get_model_from_config('mel_band_roformer', os.path.join(ROOT, 'configs', 'config_mel_band_roformer_4stem.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=41)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    x = torch.from_numpy(synth_mix(352800, 2, seed=42))[None].cuda()
    y = model(x).cpu().numpy()
    mcfg = {k: v for k, v in dict(cfg.model).items()}
    torch.set_num_threads(os.cpu_count())
    with torch.inference_mode():
        ref = orof.mel_band_roformer_forward(sd, mcfg, x.cpu()).numpy()
    print('full-size Mel 4-stem chunk: max_rel', max_rel(ref, y), 'snr', snr_db(ref, y))
    assert y.shape == ref.shape == (1, 4, 2, 352800)
    assert max_rel(ref, y) <= FP32_MAX_REL
    assert snr_db(ref, y) >= FP32_SNR_DB


def test_demix_mdx23c_vs_oracle():
    """demix() end to end with the MDX23C model class (BASELINE C1 flow at a small size): CUDA path vs the oracle's
    demix restatement driving the oracle's MDX23C forward on the CPU."""
    import sesa_audio_separation_b200 as sesa
    from oracle import mdx23c as omdx
    case = CASES['mdx_small']
    model, sd = build(case)
    cfgd = case['cfg']
    L = cfgd['audio']['chunk_size']
    cfg = sesa.ConfigDict(dict(audio=cfgd['audio'], inference=dict(num_overlap=4, batch_size=1), training=cfgd['training']))
    mix = synth_mix(L * 2 + 1234, 2, seed=51)
    res = sesa.demix(cfg, model, mix, 'cuda', 'mdx23c', engine_batch=3)
    ocfg = dict(audio=cfgd['audio'], model=cfgd['model'], num_target_instruments=2)
    with torch.inference_mode():
        ref = odemix.demix(mix, lambda a: omdx.mdx23c_forward(sd, ocfg, a), L, 4, 1, 2)
    assert list(res.keys()) == ['vocals', 'other']
    for i, k in enumerate(res):
        print('demix mdx23c', k, 'max_rel', max_rel(ref[i], res[k]), 'snr', snr_db(ref[i], res[k]))
        assert max_rel(ref[i], res[k]) <= FP32_MAX_REL
        assert snr_db(ref[i], res[k]) >= FP32_SNR_DB


def _flow_config(fc):
    import sesa_audio_separation_b200 as sesa
    return sesa.ConfigDict(dict(audio=dict(chunk_size=fc['chunk_size'], sample_rate=44100, num_channels=2),
                                inference=dict(num_overlap=fc['num_overlap'], batch_size=fc['batch_size'],
                                               **({'normalize': True} if fc['normalize'] else {})),
                                training=dict(instruments=fc['instruments'], target_instrument=fc['target'])))


@pytest.mark.parametrize('name', list(FLOW_CASES))
def test_per_file_flow_matches_reference_golden(tmp_path, name):
    """normalize -> demix -> TTA -> DemudPhaseRemix -> instrumental -> denormalize (inference_pytorch.py:219-260) through
    the product's run_folder against the arrays the UNMODIFIED reference flow wrote (oracle/make_golden.py run_flows):
    covers utils.normalize_audio / denormalize_audio / apply_tta and both DemudPhaseRemix branches."""
    import argparse
    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200.audio_io import load_audio, write_audio
    from sesa_audio_separation_b200.inference import run_folder
    fc = FLOW_CASES[name]
    case = CASES[fc['model']]
    model, _ = build(case)
    config = _flow_config(fc)
    mix = synth_mix(fc['length'], 2, seed=fc['seed'])
    indir, outdir = tmp_path / 'in', tmp_path / 'out'
    indir.mkdir()
    write_audio(str(indir / 'song.wav'), mix.T, 44100, subtype='FLOAT')
    args = argparse.Namespace(input_folder=str(indir), store_dir=str(outdir), disable_detailed_pbar=True,
                              use_tta=fc['use_tta'], demud_phaseremix_inst=fc['demud'],
                              extract_instrumental=fc['extract_instrumental'], model_type=case['kind'],
                              export_format='wav FLOAT', flac_file=False, pcm_type='PCM_24')
    backend = sesa.create_inference_session(model, device='cuda', optimize_mode='default', enable_amp=False)
    written = run_folder(backend, args, config, 'cuda', model=model)
    g = golden(name)
    assert sorted(os.path.basename(w) for w in written) == sorted(g.keys())
    for fn in g.keys():
        got, _ = load_audio(str(outdir / fn), 44100)
        assert got.shape == g[fn].shape
        gl, fr, snr = parity_report(f'{name} {fn}', g[fn], got)
        assert gl <= FP32_MAX_REL and snr >= FP32_SNR_DB


def test_apply_tta_matches_reference_golden_and_fused_run():
    """utils.apply_tta (utils.py:241-292) against the reference golden of the flow that ends with TTA-only arithmetic
    (flow_bs_norm_tta_demud's vocals = denormalize(apply_tta(demix(normalize(mix))))), and the three forms the product
    offers — demix()+apply_tta() (two engine runs), demix_tta() (one engine run, combined on the device) — are
    bit-identical to each other."""
    import sesa_audio_separation_b200 as sesa
    fc = FLOW_CASES['flow_bs_norm_tta_demud']
    case = CASES[fc['model']]
    model, _ = build(case)
    config = _flow_config(fc)
    mix = synth_mix(fc['length'], 2, seed=fc['seed'])
    nmix, params = sesa.normalize_audio(mix)
    base = sesa.demix(config, model, nmix, 'cuda', case['kind'])
    two_step = sesa.apply_tta(config, model, nmix, {k: v.copy() for k, v in base.items()}, 'cuda', case['kind'])
    fused = sesa.demix_tta(config, model, nmix, 'cuda', case['kind'], engine_batch=3)
    assert np.array_equal(two_step['vocals'], fused['vocals'])
    ref = golden('flow_bs_norm_tta_demud')['song.wav_vocals.wav']
    got = sesa.denormalize_audio(fused['vocals'], params)
    gl, fr, snr = parity_report('apply_tta vs reference', ref, got)
    assert gl <= FP32_MAX_REL and snr >= FP32_SNR_DB
    # and the oracle's restatement of apply_tta applied to the product's three demix results gives the same bits
    ora = odemix.apply_tta(nmix, lambda m: sesa.demix(config, model, m, 'cuda', case['kind']), {k: v.copy() for k, v in base.items()})
    assert np.array_equal(ora['vocals'], fused['vocals'])


def test_normalize_denormalize_match_reference_statements():
    """utils.normalize_audio / denormalize_audio (utils.py:199-238) are host numpy in the reference and in the product:
    same bits as the oracle's restatement, and the full normalize -> demix -> denormalize flow is pinned by the
    'flow_bs_norm_only' reference golden above."""
    import sesa_audio_separation_b200 as sesa
    mix = synth_mix(30000, 2, seed=5) * 0.3 + 0.05
    a, pa = sesa.normalize_audio(mix)
    b, pb = odemix.normalize_audio(mix)
    assert np.array_equal(a, b) and pa['mean'] == pb['mean'] and pa['std'] == pb['std']
    assert np.array_equal(sesa.denormalize_audio(a, pa), odemix.denormalize_audio(b, pb))


def test_hop512_config_and_odd_lengths_vs_oracle():
    """The other RoFormer geometry named by the north star (hop 512, ZFTurbo-style generic config) through demix(), with a
    mix length that leaves ragged tail chunks; checked against the oracle's demix + forward on the CPU."""
    import sesa_audio_separation_b200 as sesa
    from oracle import roformer as orof
    cfg = dict(dim=64, depth=1, stereo=True, num_stems=1, time_transformer_depth=1, freq_transformer_depth=1, dim_head=64,
               heads=2, stft_n_fft=2048, stft_hop_length=512, stft_win_length=2048, mask_estimator_depth=2,
               mlp_expansion_factor=2)
    model = sesa.BSRoformer(**cfg)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=71)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    L = 512 * 24
    config = sesa.ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=4, batch_size=2),
                                  training=dict(instruments=['vocals', 'other'], target_instrument='vocals')))
    for length in (L * 3 + 1357, L // 3, 2 * L):
        mix = synth_mix(length, 2, seed=72 + length % 7)
        got = sesa.demix(config, model, mix, 'cuda', 'bs_roformer')['vocals']
        with torch.inference_mode():
            ref = odemix.demix(mix, lambda a: orof.bs_roformer_forward(sd, cfg, a), L, 4, 2, 1)[0]
        print('hop512 length', length, 'max_rel', max_rel(ref, got), 'snr', snr_db(ref, got))
        assert got.shape == ref.shape
        assert max_rel(ref, got) <= FP32_MAX_REL and snr_db(ref, got) >= FP32_SNR_DB


def test_full_size_identity_demix_is_bit_exact_and_engine_batch_invariant():
    """BASELINE C2 sizes (3-min 44.1 kHz stereo track, chunk 352 800, overlap 4 -> 96 chunks): with an identity model the
    whole demix bookkeeping (border padding, framing, per-flush window rule, overlap-add in ascending chunk order, divide,
    crop) must reproduce the oracle's restatement of utils.py:369-464 bit for bit, for every engine batch (the engine
    batch regroups chunks but must never change a sample), through the vectorised region overlap-add kernel."""
    import sesa_audio_separation_b200 as sesa
    length, L, ov = 180 * 44100, 352800, 4
    cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=ov, batch_size=1),
                               training=dict(instruments=['a'], target_instrument='a')))
    mix = synth_mix(length, 2, seed=4242)
    ref = odemix.demix(mix, lambda a: a, L, ov, 1, 1)[0]
    outs = []
    for eb in (1, 4, 7):
        eng = sesa.DemixEngine(cfg, _identity_model(), 'cuda', engine_batch=eb)
        est = eng.run(mix)
        assert eng.plan.n_chunks == 96
        outs.append(est[0])
        assert np.array_equal(est[0], ref), ('engine_batch', eb, float(np.abs(est[0] - ref).max()))
    # identity in, identity out (up to the rounding of w*x / w)
    assert np.abs(outs[0] - np.asarray(mix)).max() < 1e-6


def test_full_size_bs_roformer_forward_is_batch_invariant_and_reproducible():
    """BASELINE C2 model on full 352 800-sample chunks: a chunk's output is bit-identical whether it runs alone or inside
    a batch of three (engine batches and chunk-range shards regroup chunks), and a repeated launch reproduces every bit
    (persistent attention CTAs, grouped GEMM tile walks and the iSTFT's two-contributor atomics are all order-free)."""
    import sesa_audio_separation_b200 as sesa
    cfg = dict(dim=512, depth=12, stereo=True, num_stems=1, time_transformer_depth=1, freq_transformer_depth=1,
               dim_head=64, heads=8, stft_n_fft=2048, stft_hop_length=441, stft_win_length=2048,
               mask_estimator_depth=2)
    model = sesa.BSRoformer(**cfg)
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=3)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    x = torch.stack([torch.from_numpy(synth_mix(352800, 2, seed=20 + i)) for i in range(3)]).cuda()
    y = model(x).clone()
    y2 = model(x).clone()
    assert torch.equal(y, y2)
    for i in (0, 2):
        yi = model(x[i:i + 1])
        assert torch.equal(yi[0], y[i]), (i, float((yi[0] - y[i]).abs().max()))


# ----------------------------------------------------------------------------------------------------------------------
# Full BASELINE configurations end to end: the real model through sesa.demix() over a multi-chunk track (border pad,
# ragged tail chunks, several engine batches) against oracle.demix driving the oracle forward in true fp32.
def _oracle_on_gpu():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _check_stems(label, names, ref, res):
    for i, k in enumerate(names):
        assert res[k].shape == ref[i].shape
        gl, fr, snr = parity_report(f'{label} {k}', ref[i], res[k])
        assert gl <= FP32_MAX_REL and snr >= FP32_SNR_DB
        assert fr <= FP32_MAX_REL         # the stricter per-frame reading of the same gate


def test_full_config_c2_bs_roformer_demix_vs_oracle():
    """BASELINE C2: BS-RoFormer dim 512 depth 12 (159.8 M parameters), chunk 352 800, overlap 4, on a 37-s track ->
    22 chunks incl. the reflect-padded borders and the three ragged tail chunks (reflect / zero / zero padded),
    6 engine batches; oracle = oracle.demix + bs_roformer_forward on CUDA in fp32 (TF32 off)."""
    import sesa_audio_separation_b200 as sesa
    from conftest import ROOT
    from oracle import roformer as orof
    _oracle_on_gpu()
    model, config = sesa.get_model_from_config('bs_roformer', os.path.join(ROOT, 'configs', 'config_bs_roformer_vocals.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=3)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    L, ov, bs = int(config.audio.chunk_size), int(config.inference.num_overlap), int(config.inference.batch_size)
    length = 37 * 44100 + 1234
    mix = synth_mix(length, 2, seed=91)
    eng = sesa.DemixEngine(config, model, 'cuda', engine_batch=4)
    est = eng.run(mix)
    assert eng.plan.n_chunks >= 20 and eng.plan.pad and eng.plan.lens[-1] < L
    mcfg = dict(config.model)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = odemix.demix(mix, lambda a: orof.bs_roformer_forward(sd_gpu, mcfg, a.cuda()).cpu(), L, ov, bs, 1)
    _check_stems(f'C2 full config ({eng.plan.n_chunks} chunks)', ['vocals'], ref, {'vocals': est[0]})


def test_full_config_c3_mel_4stem_demix_vs_oracle():
    """BASELINE C3 model (Mel-Band-RoFormer 4 stems, 832.6 M parameters), chunk 352 800, overlap 2, on a 20-s track ->
    7 chunks = 2 engine batches of 4 + 3; oracle on CUDA in fp32 except torch.istft, which is only accurate on the CPU for
    this many signals (oracle/third_party.py:istft_exact, profiles/r2_torch_istft_cuda.md); the first chunk is
    cross-checked against the pinned all-CPU oracle."""
    import sesa_audio_separation_b200 as sesa
    from conftest import ROOT
    from oracle import roformer as orof
    _oracle_on_gpu()
    model, config = sesa.get_model_from_config('mel_band_roformer', os.path.join(ROOT, 'configs', 'config_mel_band_roformer_4stem.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=41)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    L, ov, bs = int(config.audio.chunk_size), int(config.inference.num_overlap), int(config.inference.batch_size)
    length = 20 * 44100
    mix = synth_mix(length, 2, seed=92)
    eng = sesa.DemixEngine(config, model, 'cuda', engine_batch=4)
    est = eng.run(mix)
    assert eng.plan.n_chunks >= 5
    mcfg = dict(config.model)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = odemix.demix(mix, lambda a: orof.mel_band_roformer_forward(sd_gpu, mcfg, a.cuda()).cpu(), L, ov, bs, 4)
        x0 = torch.from_numpy(mix[:, :L].copy())[None]
        torch.set_num_threads(os.cpu_count())
        cpu0 = orof.mel_band_roformer_forward(sd, mcfg, x0).numpy()
        gpu0 = orof.mel_band_roformer_forward(sd_gpu, mcfg, x0.cuda()).cpu().numpy()
    print('oracle on CUDA vs pinned CPU oracle, one chunk: max_rel', max_rel(cpu0, gpu0))
    assert max_rel(cpu0, gpu0) <= 2e-5
    names = list(sesa.prefer_target_instrument(config))
    _check_stems(f'C3 full config ({eng.plan.n_chunks} chunks)', names, ref, dict(zip(names, est)))


def test_full_config_c1_mdx23c_30s_demix_vs_oracle():
    """BASELINE C1: MDX23C vocals (112 M parameters), 30-s track, chunk 261 120, overlap 4 -> 27 chunks, the exact
    configuration BASELINE.json names; oracle.demix + mdx23c_forward on CUDA in fp32."""
    import sesa_audio_separation_b200 as sesa
    from conftest import ROOT
    from oracle import mdx23c as omdx
    _oracle_on_gpu()
    model, cfg = sesa.get_model_from_config('mdx23c', os.path.join(ROOT, 'configs', 'config_vocals_mdx23c.yaml'))
    sd = fill_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=21)
    model.load_state_dict(sd)
    model.eval().to('cuda')
    L, ov, bs = int(cfg.audio.chunk_size), int(cfg.inference.num_overlap), int(cfg.inference.batch_size)
    mix = synth_mix(30 * 44100, 2, seed=93)
    eng = sesa.DemixEngine(cfg, model, 'cuda', engine_batch=4)
    est = eng.run(mix)
    assert eng.plan.n_chunks == 27
    ocfg = dict(audio=dict(cfg.audio), model=dict(cfg.model), num_target_instruments=2)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.inference_mode():
        ref = odemix.demix(mix, lambda a: omdx.mdx23c_forward(sd_gpu, ocfg, a.cuda()).cpu(), L, ov, bs, 2)
    names = list(sesa.prefer_target_instrument(cfg))
    _check_stems('C1 full config (27 chunks)', names, ref, dict(zip(names, est)))


def test_in_memory_ensemble_of_three_models_matches_reference_arithmetic(tmp_path):
    """BASELINE C5 flow at a small size: BS-RoFormer + Mel-Band-RoFormer + MDX23C separate the same mix, the three vocals
    estimates are reduced on the device (sesa_ensemble_wave).  Reference arithmetic (ensemble.py:172-183 on the float64
    buffers soundfile hands it): np.mean / np.average / np.median over the model axis — here applied to each member's own
    demix() result, which the parity tests above pin to the reference."""
    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200.audio_io import load_audio
    from sesa_audio_separation_b200.ensemble import ensemble_separate, ensemble_tracks
    L = 441 * 32
    members, singles = [], []
    mix = synth_mix(L * 2 + 700, 2, seed=77)
    for name in ('bs_small', 'mel_small', 'mdx_small'):
        case = CASES[name]
        model, _ = build(case)
        if name == 'mdx_small':
            Lm = case['cfg']['audio']['chunk_size']
            cfg = sesa.ConfigDict(dict(audio=case['cfg']['audio'], inference=dict(num_overlap=2, batch_size=1),
                                       training=case['cfg']['training']))
        else:
            cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=2, batch_size=1),
                                       training=dict(instruments=['vocals', 'other'], target_instrument=None if name == 'mel_small' else 'vocals')))
        members.append((cfg, model))
        singles.append(sesa.demix(cfg, model, mix, 'cuda', case['kind'])['vocals'])
    f64 = np.stack([s.astype(np.float64) for s in singles], 0)
    got = ensemble_separate(members, mix, 'cuda', stem='vocals')
    assert np.array_equal(got, np.mean(f64, axis=0).astype(np.float32))
    w = np.array([2.0, 1.0, 1.0], dtype=np.float32)
    w /= w.sum()
    got_w = ensemble_separate(members, mix, 'cuda', stem='vocals', weights=[2.0, 1.0, 1.0])
    assert np.array_equal(got_w, np.average(f64, axis=0, weights=w).astype(np.float32))
    assert np.array_equal(ensemble_separate(members, mix, 'cuda', stem='vocals', method='median_wave'),
                          np.median(f64, axis=0).astype(np.float32))
    # track sharding (rank r of W takes tracks r, r+W, ...) and the PCM_24 output of ensemble.py:311
    tracks = [(f't{i}', mix * (1.0 - 0.1 * i)) for i in range(3)]
    done = ensemble_tracks(members, tracks, 'cuda', out_dir=str(tmp_path), rank=1, world=2)
    assert list(done) == ['t1']
    back, _ = load_audio(str(tmp_path / 't1_vocals_ensemble.wav'), 44100)
    assert np.abs(back - done['t1']).max() <= 2.0 ** -23 + 1e-9


class _FakeDist:
    """In-process stand-in for the torch.distributed point-to-point calls distributed.py uses: two 'ranks' run as two
    threads on the same GPU and exchange tensors through queues (NCCL refuses two ranks on one device; the real NCCL path is
    exercised by bench.py's `strong` record and tools/shard_check.py on 2-8 GPUs, the choreography by the gloo tests)."""

    def __init__(self, world):
        import queue
        import threading
        self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}
        self.local = threading.local()

    class P2POp:
        def __init__(self, op, tensor, peer, group=None):
            self.op, self.tensor, self.peer = op, tensor, peer

    class _Req:
        def __init__(self, fn):
            self.fn, self.done = fn, False

        def wait(self):
            if not self.done:
                self.fn()
                self.done = True

    def isend(self, tensor, dst, group=None):
        torch.cuda.synchronize()
        self.q[(self.local.rank, dst)].put(tensor.detach().clone())
        return self._Req(lambda: None)

    def irecv(self, tensor, src, group=None):
        rank = self.local.rank

        def finish():
            tensor.copy_(self.q[(src, rank)].get(timeout=120))
            torch.cuda.synchronize()
        return self._Req(finish)

    def batch_isend_irecv(self, ops):
        sends = [self.isend(o.tensor, o.peer) for o in ops if o.op == self.isend]
        recvs = [self.irecv(o.tensor, o.peer) for o in ops if o.op == self.irecv]
        return recvs + sends if recvs else sends

    def get_global_rank(self, group, r):
        return r


@pytest.mark.parametrize('name,world,ov', [('bs_small', 2, 4), ('mel_small', 3, 2)])
def test_chunk_range_sharding_on_the_gpu_path_is_bit_identical(monkeypatch, name, world, ov):
    """DemixEngine(world, rank) through the real CUDA path (slice upload + sesa_pad_reflect_slice, tail-first order,
    sesa_overlap_accumulate continuing from the received halo, grouped gather) with `world` ranks as threads on one GPU:
    the root's result equals the unsharded run bit for bit, and every rank uploads only its slice of the mix."""
    import threading
    import sesa_audio_separation_b200 as sesa
    from sesa_audio_separation_b200 import distributed
    case = CASES[name]
    L = 441 * 32
    cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=ov, batch_size=1),
                               training=dict(instruments=['vocals', 'other'],
                                             target_instrument=None if name == 'mel_small' else 'vocals')))
    mix = synth_mix(L * 6 + 1234, 2, seed=71)
    ref_model, _ = build(case)
    ref = sesa.DemixEngine(cfg, ref_model, 'cuda', engine_batch=3).run(mix)
    fake = _FakeDist(world)
    monkeypatch.setattr(distributed, 'dist', fake)
    out, err, stats = {}, [], {}

    def worker(rank):
        try:
            fake.local.rank = rank
            model, _ = build(case)
            eng = sesa.DemixEngine(cfg, model, 'cuda', engine_batch=3, world=world, rank=rank)
            out[rank] = eng.run(mix)
            stats[rank] = dict(eng.stats)
        except Exception as e:          # surface failures of either thread in the main thread
            import traceback
            err.append((rank, traceback.format_exc()))
    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not err, err
    assert all(out[r] is None for r in range(1, world))
    assert out[0].shape == ref.shape
    assert np.array_equal(out[0], ref)
    assert all(0 < stats[r]['h2d_bytes'] < mix.nbytes for r in range(world)), stats
