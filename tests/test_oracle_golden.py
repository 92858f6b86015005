"""CPU: the oracle restatements reproduce what the unmodified reference produced (tests/golden)."""
import numpy as np
import pytest
import torch

from conftest import golden, max_rel, snr_db
from oracle import demix as odemix
from oracle import mdx23c as omdx
from oracle import roformer as orof
from oracle import flow as oflow
from oracle.cases import CASES, DEMIX_IDENTITY_CASES, DEMIX_MODEL_CASES, FLOW_CASES, make_input
from oracle.weights import fill_state_dict, synth_mix


def oracle_forward(case, sd, x):
    if case['kind'] == 'bs_roformer':
        return orof.bs_roformer_forward(sd, case['cfg'], x)
    if case['kind'] == 'mel_band_roformer':
        return orof.mel_band_roformer_forward(sd, case['cfg'], x)
    cfg = dict(case['cfg'])
    cfg['num_target_instruments'] = len(odemix.prefer_target_instrument(cfg['training']))
    return omdx.mdx23c_forward(sd, cfg, x)


def seeded_sd(manifest, name):
    m = manifest[name]
    sd = fill_state_dict({k: tuple(v) for k, v in m['shapes'].items()}, CASES[name]['seed'])
    csum = float(sum(v.double().sum().item() for v in sd.values()))
    assert csum == m['weight_checksum'], 'seeded weights differ from the ones the golden run used'
    return sd


@pytest.mark.parametrize('name', list(CASES))
def test_forward_matches_reference(manifest, name):
    case = CASES[name]
    sd = seeded_sd(manifest, name)
    with torch.inference_mode():
        y = oracle_forward(case, sd, make_input(case)).numpy()
    ref = golden(name)['y']
    assert y.shape == ref.shape
    assert max_rel(ref, y) <= 2e-5, max_rel(ref, y)
    assert snr_db(ref, y) >= 90


def test_mel_index_maps_bit_exact():
    g = golden('mel_small')
    fi, nfpb, nbpf, _ = orof.mel_band_index_maps(CASES['mel_small']['cfg'])
    assert np.array_equal(fi.numpy(), g['freq_indices'])
    assert np.array_equal(np.asarray(nfpb), g['num_freqs_per_band'])
    assert np.array_equal(nbpf.numpy(), g['num_bands_per_freq'])


def test_demix_identity_bit_exact():
    g = golden('demix_identity')
    for i, (length, L, ov, bs) in enumerate(DEMIX_IDENTITY_CASES):
        mix = synth_mix(length, 2, seed=100 + i)
        est = odemix.demix(mix, lambda a: a, L, ov, bs, 1)[0]
        assert np.array_equal(est, g[f'case{i}']), (i, length, L, ov, bs)


@pytest.mark.parametrize('name', list(DEMIX_MODEL_CASES))
def test_demix_model_matches_reference(manifest, name):
    dc = DEMIX_MODEL_CASES[name]
    case = CASES[dc['model']]
    sd = seeded_sd(manifest, dc['model'])
    instr = odemix.prefer_target_instrument(dict(instruments=dc['instruments'], target_instrument=dc['target']))
    mix = synth_mix(dc['length'], 2, seed=dc['seed'])
    with torch.inference_mode():
        est = odemix.demix(mix, lambda a: oracle_forward(case, sd, a), dc['chunk_size'], dc['num_overlap'],
                           dc['batch_size'], len(instr))
    g = golden(name)
    for k, e in zip(instr, est):
        assert e.shape == g[k].shape
        assert max_rel(g[k], e) <= 2e-5
        assert snr_db(g[k], e) >= 90


@pytest.mark.parametrize('name', list(FLOW_CASES))
def test_flow_matches_reference(manifest, name):
    """normalize -> demix -> TTA -> DemudPhaseRemix -> instrumental -> denormalize: the oracle's restatement
    (oracle/flow.py) against what the unmodified inference_pytorch.run_folder_pytorch_optimized wrote."""
    fc = FLOW_CASES[name]
    case = CASES[fc['model']]
    sd = seeded_sd(manifest, fc['model'])
    instr = odemix.prefer_target_instrument(dict(instruments=fc['instruments'], target_instrument=fc['target']))
    mix = synth_mix(fc['length'], 2, seed=fc['seed'])

    def demix_fn(m):
        with torch.inference_mode():
            est = odemix.demix(m, lambda a: oracle_forward(case, sd, a), fc['chunk_size'], fc['num_overlap'],
                               fc['batch_size'], len(instr))
        return {k: e for k, e in zip(instr, est)}
    out, names = oflow.separate_track(mix, demix_fn, fc['instruments'] if fc['target'] is None else instr,
                                      normalize=fc['normalize'], use_tta=fc['use_tta'], demud=fc['demud'],
                                      extract_instrumental=fc['extract_instrumental'])
    g = golden(name)
    assert sorted(f'song.wav_{n}.wav' for n in names) == sorted(g.keys())
    for n in names:
        ref = g[f'song.wav_{n}.wav']
        assert out[n].shape == ref.shape
        assert max_rel(ref, out[n]) <= 2e-5, (n, max_rel(ref, out[n]))
        assert snr_db(ref, out[n]) >= 90, (n, snr_db(ref, out[n]))
