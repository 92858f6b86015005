"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python -m oracle.make_golden
Third-party modules that are not installed are replaced by the stand-ins in oracle/shims.
Weights/inputs are produced by oracle.weights (seeded), so the fixtures hold only outputs,
the state_dict manifest (names+shapes) and a weight checksum.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
sys.path.insert(2, '/root/reference')
warnings.filterwarnings('ignore')

from oracle.weights import fill_state_dict, synth_mix  # noqa: E402
from oracle.cases import CASES, DEMIX_IDENTITY_CASES, DEMIX_MODEL_CASES, FLOW_CASES, make_input  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def build_reference_model(kind, cfg):
    if kind == 'bs_roformer':
        from models.bs_roformer.bs_roformer import BSRoformer
        kw = dict(cfg)
        if 'freqs_per_bands' in kw:
            kw['freqs_per_bands'] = tuple(kw['freqs_per_bands'])
        return BSRoformer(**kw)
    if kind == 'mel_band_roformer':
        from models.bs_roformer.mel_band_roformer import MelBandRoformer
        return MelBandRoformer(**cfg)
    if kind == 'mdx23c':
        from ml_collections import ConfigDict
        from models.mdx23c_tfc_tdf_v3 import TFC_TDF_net
        return TFC_TDF_net(ConfigDict(cfg))
    raise ValueError(kind)


def load_seeded(model, seed):
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if v.dtype.is_floating_point}
    new = fill_state_dict(shapes, seed)
    model.load_state_dict(new, strict=False)
    model.eval()
    csum = float(sum(v.double().sum().item() for v in new.values()))
    return shapes, csum


def run_flows():
    """The per-file flow (normalize / TTA / DemudPhaseRemix / instrumental / denormalize) through the unmodified
    ``inference_pytorch.run_folder_pytorch_optimized`` (:189-274).  Only file I/O is replaced: ``librosa.load`` hands
    back the seeded mix and ``soundfile.write`` captures the arrays that would have been written."""
    import argparse
    import tempfile
    import librosa
    import soundfile
    from ml_collections import ConfigDict

    box = {}
    librosa.load = lambda path, sr=None, mono=False: (box['mix'].copy(), sr)
    soundfile.write = lambda path, data, sr, subtype=None: box['out'].__setitem__(os.path.basename(path), np.array(data).T)
    import inference_pytorch as ref_inf
    from pytorch_backend import PyTorchBackend

    # what create_inference_session (pytorch_backend.py:492-536) does on a CPU device; the backend object is made once
    # because its constructor sets the interop thread count, which torch allows once per process
    backend = PyTorchBackend(device='cpu', optimize_mode='default')
    for name, fc in FLOW_CASES.items():
        case = CASES[fc['model']]
        model = build_reference_model(case['kind'], case['cfg'])
        load_seeded(model, case['seed'])
        cfg = ConfigDict(dict(audio=dict(chunk_size=fc['chunk_size'], sample_rate=44100),
                              inference=dict(num_overlap=fc['num_overlap'], batch_size=fc['batch_size'],
                                             **({'normalize': True} if fc['normalize'] else {})),
                              training=dict(instruments=fc['instruments'], target_instrument=fc['target'],
                                            use_amp=False)))
        box['mix'] = synth_mix(fc['length'], 2, seed=fc['seed'])
        box['out'] = {}
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, 'in'))
            open(os.path.join(tmp, 'in', 'song.wav'), 'wb').close()
            args = argparse.Namespace(input_folder=os.path.join(tmp, 'in'), store_dir=os.path.join(tmp, 'out'),
                                      disable_detailed_pbar=True, use_tta=fc['use_tta'],
                                      demud_phaseremix_inst=fc['demud'], extract_instrumental=fc['extract_instrumental'],
                                      model_type=case['kind'], export_format='wav FLOAT', flac_file=False,
                                      pcm_type='PCM_24')
            backend.optimize_model(model, use_amp=False)
            ref_inf.run_folder_pytorch_optimized(backend, args, cfg, 'cpu', model=model)
        np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **box['out'])
        print(name, {k: v.shape for k, v in box['out'].items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    manifest = {}
    for name, case in CASES.items():
        model = build_reference_model(case['kind'], case['cfg'])
        shapes, csum = load_seeded(model, case['seed'])
        x = make_input(case)
        with torch.inference_mode():
            y = model(x)
        extra = {}
        if case['kind'] == 'mel_band_roformer':
            extra = dict(freq_indices=model.freq_indices.numpy(),
                         num_freqs_per_band=model.num_freqs_per_band.numpy(),
                         num_bands_per_freq=model.num_bands_per_freq.numpy())
        np.savez_compressed(os.path.join(OUT, f'{name}.npz'), y=y.numpy(), **extra)
        manifest[name] = dict(shapes={k: list(v) for k, v in shapes.items()}, weight_checksum=csum,
                              out_shape=list(y.shape))
        print(name, tuple(y.shape), float(y.abs().max()))

    import utils as ref_utils
    from ml_collections import ConfigDict

    # identity-model demix: pins chunk schedule, window choice, counter, 0/0 quirk
    class Ident(torch.nn.Module):
        def forward(self, x):
            return x

    ident = {}
    for i, (length, L, ov, bs) in enumerate(DEMIX_IDENTITY_CASES):
        cfg = ConfigDict(dict(audio=dict(chunk_size=L), inference=dict(num_overlap=ov, batch_size=bs),
                              training=dict(instruments=['a'], target_instrument='a', use_amp=False)))
        mix = synth_mix(length, 2, seed=100 + i)
        res = ref_utils.demix(cfg, Ident(), mix, 'cpu', model_type='generic')
        ident[f'case{i}'] = res['a']
    np.savez_compressed(os.path.join(OUT, 'demix_identity.npz'), **ident)
    print('demix identity cases', len(ident))

    for name, dc in DEMIX_MODEL_CASES.items():
        case = CASES[dc['model']]
        model = build_reference_model(case['kind'], case['cfg'])
        load_seeded(model, case['seed'])
        instruments = dc['instruments']
        cfg = ConfigDict(dict(audio=dict(chunk_size=dc['chunk_size']),
                              inference=dict(num_overlap=dc['num_overlap'], batch_size=dc['batch_size']),
                              training=dict(instruments=instruments, target_instrument=dc['target'],
                                            use_amp=False)))
        mix = synth_mix(dc['length'], 2, seed=dc['seed'])
        res = ref_utils.demix(cfg, model, mix, 'cpu', model_type=case['kind'])
        np.savez_compressed(os.path.join(OUT, f'{name}.npz'), **{k: v for k, v in res.items()})
        print(name, {k: v.shape for k, v in res.items()})

    run_flows()

    with open(os.path.join(OUT, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=0, sort_keys=True)


if __name__ == '__main__':
    main()
