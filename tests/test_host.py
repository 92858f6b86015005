"""CPU: host logic, the C-ABI library's exported symbols, and the no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from oracle import demix as odemix
from oracle.cases import DEMIX_IDENTITY_CASES


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from sesa_audio_separation_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'sesa_b200.h')).read()
    declared = set(re.findall(r'\b(sesa_[a-z0-9_]+)\s*\(', header))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/sesa_b200.h but not exported'
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.sesa_abi_version() == 3


def test_plan_matches_oracle_schedule():
    from sesa_audio_separation_b200.plan import make_plan
    kinds = {'both': 0, 'nofadein': 1, 'nofadeout': 2}
    modes = {'constant': 0, 'reflect': 1}
    cases = list(DEMIX_IDENTITY_CASES) + [(7938000, 352800, 4, 1), (7938000, 352800, 4, 2), (1323000, 261120, 4, 1),
                                          (26460000, 352800, 2, 4)]
    for length, L, ov, bs in cases:
        p = make_plan(length, L, ov, bs)
        o = odemix.demix_schedule(length, L, ov, bs)
        assert (p.step, p.border, p.fade, p.padded, p.pad) == (o['step'], o['border'], o['fade'], o['padded'], o['pad'])
        assert list(zip(p.starts, p.lens, p.modes, p.kinds)) == [(s, l, modes[m], kinds[w]) for s, l, m, w in o['chunks']]
    p = make_plan(7938000, 352800, 4, 1)
    assert p.n_chunks == 96 and p.padded == 8467200 and p.lens[-3:] == [264600, 176400, 88200]
    assert make_plan(1323000, 261120, 4, 1).n_chunks == 27


def test_window_is_torch_linspace():
    from sesa_audio_separation_b200.plan import windowing_array
    w = windowing_array(352800, 35280)
    assert torch.equal(w, odemix.windowing_array(352800, 35280))


def test_state_dict_layout_matches_reference_manifest(manifest):
    import sesa_audio_separation_b200 as sesa
    from oracle.cases import CASES
    for name, case in CASES.items():
        cfg = dict(case['cfg'])
        if case['kind'] == 'bs_roformer':
            if 'freqs_per_bands' in cfg:
                cfg['freqs_per_bands'] = tuple(cfg['freqs_per_bands'])
            m = sesa.BSRoformer(**cfg)
        elif case['kind'] == 'mel_band_roformer':
            m = sesa.MelBandRoformer(**cfg)
        else:
            continue
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == manifest[name]['shapes']


def test_mel_band_maps_bit_exact():
    from sesa_audio_separation_b200.roformer import mel_band_maps
    g = golden('mel_small')
    fi, nfpb, nbpf, _ = mel_band_maps(44100, 2048, 60, True)
    assert np.array_equal(fi, g['freq_indices'])
    assert np.array_equal(nfpb, g['num_freqs_per_band'])
    assert np.array_equal(nbpf, g['num_bands_per_freq'])


def test_no_cpu_fallback():
    import sesa_audio_separation_b200 as sesa
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    m = sesa.BSRoformer(64, depth=1, stereo=True, heads=1, time_transformer_depth=1, freq_transformer_depth=1)
    with pytest.raises(sesa.SesaError):
        m(torch.zeros(1, 2, 4410))
    cfg = sesa.ConfigDict(dict(audio=dict(chunk_size=1000), inference=dict(num_overlap=2, batch_size=1),
                               training=dict(instruments=['a'], target_instrument='a')))
    with pytest.raises(sesa.SesaError):
        sesa.demix(cfg, m, np.zeros((2, 5000), np.float32), 'cpu', 'bs_roformer')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'sesa_audio_separation_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), fn


def test_config_loading_with_python_tuple(tmp_path):
    import sesa_audio_separation_b200 as sesa
    p = tmp_path / 'c.yaml'
    p.write_text('audio: {chunk_size: 352800}\nmodel:\n  dim: 64\n  depth: 1\n  stereo: true\n  heads: 1\n'
                 '  time_transformer_depth: 1\n  freq_transformer_depth: 1\n'
                 '  freqs_per_bands: !!python/tuple [512, 513]\n'
                 'training: {instruments: [vocals, other], target_instrument: vocals}\n'
                 'inference: {batch_size: 1, num_overlap: 4}\n')
    model, cfg = sesa.get_model_from_config('bs_roformer', str(p))
    assert cfg.audio.chunk_size == 352800 and model.num_bands == 2
    assert sesa.prefer_target_instrument(cfg) == ['vocals']
    with pytest.raises(ValueError):
        sesa.get_model_from_config('scnet', str(p))


def test_benchmark_yaml_configs_load_and_build():
    """configs/*.yaml (BASELINE configs 1-3) load through the reference-style registry and give the reference's
    parameter counts (SURVEY section 6 probes: BS 159.76 M, MDX23C 111.99 M, Mel 4-stem 832.6 M)."""
    import sesa_audio_separation_b200 as sesa
    want = {'bs_roformer': ('config_bs_roformer_vocals.yaml', 159_758_796, ['vocals']),
            'mel_band_roformer': ('config_mel_band_roformer_4stem.yaml', 832_599_084, ['bass', 'drums', 'other', 'vocals']),
            'mdx23c': ('config_vocals_mdx23c.yaml', 111_990_272, ['vocals', 'other'])}
    for mt, (fn, n, inst) in want.items():
        model, cfg = sesa.get_model_from_config(mt, os.path.join(ROOT, 'configs', fn))
        assert sum(v.numel() for v in model.state_dict().values()) == n
        assert list(sesa.prefer_target_instrument(cfg)) == inst
    with pytest.raises(ValueError):
        sesa.get_model_from_config('scnet', os.path.join(ROOT, 'configs', 'config_vocals_mdx23c.yaml'))


def test_audio_io_roundtrip_and_ensemble(tmp_path):
    """CLI support code on the host: WAV writer/reader (FLOAT, PCM_16, PCM_24), resampling entry, waveform ensemble modes
    against the reference's numpy statements (ensemble.py:172-183)."""
    from sesa_audio_separation_b200.audio_io import load_audio, write_audio
    from sesa_audio_separation_b200.ensemble import ensemble_waveforms, main as ens_main
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((2, 5000)) * 0.2).astype(np.float32)
    for subtype, tol in (('FLOAT', 0.0), ('PCM_16', 2.0 ** -15), ('PCM_24', 2.0 ** -23)):
        p = str(tmp_path / f'a_{subtype}.wav')
        write_audio(p, x.T, 44100, subtype=subtype)
        y, sr = load_audio(p, 44100)
        assert sr == 44100 and y.shape == x.shape and np.abs(y - x).max() <= tol
    y2, _ = load_audio(str(tmp_path / 'a_FLOAT.wav'), 22050)
    assert y2.shape == (2, 2500)
    stems = [x, x * 0.5 + 0.01, -x]
    assert np.allclose(ensemble_waveforms(stems, 'avg_wave'), np.mean(stems, axis=0))
    assert np.allclose(ensemble_waveforms(stems, 'avg_wave', [1, 2, 3]), np.average(stems, axis=0, weights=[1, 2, 3]))
    assert np.array_equal(ensemble_waveforms(stems, 'median_wave'), np.median(stems, axis=0))
    assert np.array_equal(ensemble_waveforms(stems, 'max_wave'), np.max(stems, axis=0))
    # torch inputs mean "stems still on the GPU" (sesa_ensemble_wave, tests/test_gpu_kernels.py); CPU tensors are refused
    with pytest.raises(RuntimeError):
        ensemble_waveforms([torch.from_numpy(s) for s in stems], 'avg_wave')
    files = []
    for i, s in enumerate(stems):
        f = str(tmp_path / f's{i}.wav')
        write_audio(f, s.T, 44100, subtype='FLOAT')
        files.append(f)
    out = str(tmp_path / 'ens.wav')
    with pytest.raises(SystemExit) as e:
        ens_main(['--files', *files, '--type', 'avg_wave', '--output', out])
    assert e.value.code == 0
    got, _ = load_audio(out, 44100)
    assert np.abs(got - np.mean(stems, axis=0)).max() < 2.0 ** -22
    with pytest.raises(SystemExit) as e:
        ens_main(['--files', *files, '--type', 'max_fft', '--output', out])
    assert e.value.code == 1


def test_cli_parser_accepts_reference_flags():
    from sesa_audio_separation_b200.inference import build_parser, shorten_filename
    a = build_parser().parse_args(['--model_type', 'bs_roformer', '--config_path', 'c.yaml', '--input_folder', 'in', '--store_dir',
                                   'out', '--device_ids', '0', '--extract_instrumental', '--export_format', 'wav FLOAT',
                                   '--pcm_type', 'PCM_16', '--use_tta', '--chunk_size', '485100', '--overlap', '2',
                                   '--optimize_mode', 'channels_last', '--enable_amp', '--disable_detailed_pbar'])
    assert a.model_type == 'bs_roformer' and a.device_ids == [0] and a.use_tta and a.enable_amp
    assert shorten_filename('a' * 40 + '.wav') == 'a' * 15 + '...' + 'a' * 10 + '.wav'
