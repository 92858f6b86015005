"""CPU oracle: restatement of the per-file numeric flow around ``demix``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
/root/reference/inference_pytorch.py:219-260 (``run_folder_pytorch_optimized``; the same flow is in
inference.py:92-146): normalize -> demix -> TTA -> "DemudPhaseRemix" second pass -> instrumental by
subtraction -> denormalize.  File I/O, naming and progress printing are not part of the arithmetic and are
left out.  ``demix_fn(mix) -> {instrument: ndarray(C, len)}`` is the separation itself.
"""
from .demix import apply_tta, denormalize_audio, normalize_audio


def separate_track(mix, demix_fn, instruments, normalize=False, use_tta=False, demud=False,
                   extract_instrumental=False):
    """Returns ({name: ndarray(C, len)}, [names in output order])."""
    instruments = list(instruments)
    mix_orig = mix.copy()
    norm_params = None
    if normalize:                                                     # :221-223
        mix, norm_params = normalize_audio(mix)
    waveforms = demix_fn(mix)                                         # :226
    if use_tta:                                                       # :228-229
        waveforms = apply_tta(mix, demix_fn, waveforms)
    if demud:                                                         # :231-250
        lead = 'vocals' if 'vocals' in instruments else instruments[0]
        instruments.append('instrumental_phaseremix')
        if 'instrumental' not in instruments and 'Instrumental' not in instruments:
            modified = mix_orig - 2 * waveforms[lead]
            second = demix_fn(modified)
            if use_tta:
                second = apply_tta(modified, demix_fn, second)
            waveforms['instrumental_phaseremix'] = mix_orig + second[lead]
        else:
            modified = 2 * waveforms[lead] - mix_orig
            kept = modified.copy()
            second = demix_fn(modified)
            if use_tta:     # :247 hands the FIRST pass's dict to apply_tta: it is updated in place
                second = apply_tta(modified, demix_fn, waveforms)
            waveforms['instrumental_phaseremix'] = mix_orig + kept - second[lead]
    if extract_instrumental:                                          # :252-256
        lead = 'vocals' if 'vocals' in instruments else instruments[0]
        waveforms['instrumental'] = mix_orig - waveforms[lead]
        if 'instrumental' not in instruments:
            instruments.append('instrumental')
    out = {}
    for name in instruments:                                          # :258-261
        est = waveforms[name]
        if norm_params is not None:
            est = denormalize_audio(est, norm_params)
        out[name] = est
    return out, instruments
