import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def manifest():
    with open(os.path.join(GOLDEN, 'manifest.json')) as f:
        return json.load(f)


def golden(name):
    return np.load(os.path.join(GOLDEN, f'{name}.npz'))


def snr_db(ref, est):
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    num = (ref ** 2).sum()
    den = ((ref - est) ** 2).sum()
    return 10 * np.log10(num / max(den, 1e-300))


def max_rel(ref, est):
    """max |ref-est| / max |ref| — the 'max relative error' of the fp32 parity gate."""
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    return np.abs(ref - est).max() / max(np.abs(ref).max(), 1e-300)


def max_rel_framewise(ref, est, frame=4096):
    """The stricter reading of 'max relative error': max over frames of (max |ref-est| in the frame) / (max |ref| in the
    frame), frames of ``frame`` samples along the last axis.  A frame whose reference peak is below 1e-3 of the global
    peak is normalised by that floor instead (a silent frame has no meaningful relative error)."""
    ref = np.asarray(ref, dtype=np.float64)
    est = np.asarray(est, dtype=np.float64)
    n = ref.shape[-1] // frame * frame
    if n == 0:
        return max_rel(ref, est)
    r = np.abs(ref[..., :n]).reshape(*ref.shape[:-1], -1, frame).max(-1)
    d = np.abs(ref[..., :n] - est[..., :n]).reshape(*ref.shape[:-1], -1, frame).max(-1)
    floor = 1e-3 * max(np.abs(ref).max(), 1e-300)
    return float((d / np.maximum(r, floor)).max())


def parity_report(label, ref, est):
    """Prints and returns (global max-rel, frame-wise max-rel, SNR dB)."""
    g, f, s = max_rel(ref, est), max_rel_framewise(ref, est), snr_db(ref, est)
    print(f'{label}: max_rel {g:.3e} (global), {f:.3e} (per 4096-sample frame), snr {s:.1f} dB')
    return g, f, s
