"""Waveform-domain ensembling (reference: ensemble.py:172-183 process_waveform, CLI :409-438) — SURVEY §8f rank 1.

``ensemble_waveforms`` is the in-memory form (the reference round-trips every stem through disk and averages 32768-frame
buffers in numpy): stems that are still on the GPU are combined there with torch tensor ops on the device (a pure
HBM-bound elementwise pass), numpy inputs are combined on the host exactly like the reference.  The CLI keeps the
reference's argv (``--files --type --weights --output --buffer``) and PCM_24 output, exit status 0/1.
The spectral modes (``*_fft``) are out of scope.
"""
import argparse
import sys

import numpy as np
import torch

from .audio_io import load_audio, write_audio

WAVE_METHODS = ('avg_wave', 'median_wave', 'max_wave', 'min_wave')


def ensemble_waveforms(stems, method='avg_wave', weights=None):
    """stems: list of equal-shape arrays/tensors [channels, samples] -> same type as the inputs."""
    if method not in WAVE_METHODS:
        raise NotImplementedError(f'ensemble type {method!r} is not implemented (waveform modes: {WAVE_METHODS})')
    if len(stems) == 0:
        raise ValueError('no inputs')
    if weights is not None and len(weights) != len(stems):
        raise ValueError('one weight per input is required')
    if isinstance(stems[0], torch.Tensor):
        x = torch.stack([s.to(torch.float32) for s in stems], 0)
        if method == 'avg_wave':
            if weights is None:
                return x.mean(0)
            w = torch.tensor(weights, dtype=torch.float32, device=x.device)
            return (x * w.view(-1, 1, 1)).sum(0) / w.sum()
        if method == 'median_wave':      # np.median averages the two middle values for an even count
            s, _ = x.sort(0)
            n = x.shape[0]
            return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])
        return x.max(0).values if method == 'max_wave' else x.min(0).values
    chunks = np.stack([np.asarray(s) for s in stems], 0)
    if method == 'avg_wave':
        return np.average(chunks, axis=0, weights=weights) if weights is not None else np.mean(chunks, axis=0)
    if method == 'median_wave':
        return np.median(chunks, axis=0)
    return np.max(chunks, axis=0) if method == 'max_wave' else np.min(chunks, axis=0)


def run_ensemble(files, method, output_path, weights=None, buffer_size=32768):
    """ensemble.py:258-407 reduced to its arithmetic: common sample rate of the first file, common (minimum) length,
    PCM_24 output (:311).  ``buffer_size`` is accepted for argv compatibility (the whole file fits in memory)."""
    try:
        first, sr = load_audio(files[0], _probe_rate(files[0]))
        stems = [np.atleast_2d(first)]
        for f in files[1:]:
            a, _ = load_audio(f, sr)
            stems.append(np.atleast_2d(a))
        n = min(s.shape[1] for s in stems)
        stems = [s[:, :n] for s in stems]
        out = ensemble_waveforms(stems, method, weights)
        write_audio(output_path, np.asarray(out, dtype=np.float32).T, sr, subtype='PCM_24')
        print("[SESA_PROGRESS]100", flush=True)
        return True
    except Exception as e:
        print(f"\nError during processing: {e}", file=sys.stderr)
        return False


def _probe_rate(path):
    try:
        import soundfile as sf
        return sf.info(path).samplerate
    except ImportError:
        from scipy.io import wavfile
        return wavfile.read(path, mmap=True)[0]


def main(argv=None):
    p = argparse.ArgumentParser(description='Audio ensemble (waveform modes)', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--files', nargs='+', required=True)
    p.add_argument('--type', required=True, choices=['avg_wave', 'median_wave', 'max_wave', 'min_wave', 'max_fft', 'min_fft', 'median_fft'])
    p.add_argument('--weights', nargs='+', type=float)
    p.add_argument('--output', required=True)
    p.add_argument('--buffer', type=int, default=32768)
    args = p.parse_args(argv)
    ok = run_ensemble(args.files, args.type, args.output, args.weights, args.buffer)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
