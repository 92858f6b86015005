"""Waveform-domain ensembling (reference: ensemble.py:172-183 process_waveform, CLI :409-438) — SURVEY §8f rank 1.

``ensemble_waveforms`` is the in-memory form (the reference round-trips every stem through disk and averages 32768-frame
buffers in numpy): stems that are still on the GPU are combined there by one launch of ``sesa_ensemble_wave`` (a pure
HBM-bound elementwise pass, float64 accumulation in input order), numpy inputs are combined on the host exactly like
the reference.  The CLI keeps the
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
        return _ensemble_device(stems, method, weights)
    chunks = np.stack([np.asarray(s, dtype=np.float64) for s in stems], 0)     # sf.SoundFile.read -> float64 (:330)
    if method == 'avg_wave':
        if weights is not None:
            return np.average(chunks, axis=0, weights=_normalised_weights(weights))
        return np.mean(chunks, axis=0)
    if method == 'median_wave':
        return np.median(chunks, axis=0)
    return np.max(chunks, axis=0) if method == 'max_wave' else np.min(chunks, axis=0)


def _normalised_weights(weights):
    """ensemble.py:292-295: float32 weights divided by their float32 sum."""
    w = np.array(weights, dtype=np.float32)
    w /= w.sum()
    return w


def _ensemble_device(stems, method, weights):
    """Stems still resident on the GPU (e.g. DemixEngine.run(to_host=False) of several models): one pass of
    sesa_ensemble_wave, accumulating in float64 in input order like numpy does on the reference's float64 buffers."""
    import ctypes
    from ._lib import call, require_cuda
    require_cuda()
    dev = stems[0].device
    if dev.type != 'cuda':
        raise RuntimeError('torch inputs must be CUDA tensors (host arrays are combined with numpy like the reference)')
    xs = [s.to(device=dev, dtype=torch.float32).contiguous() for s in stems]
    if any(x.shape != xs[0].shape for x in xs):
        raise ValueError('all inputs must have the same shape')
    out = torch.empty_like(xs[0])
    ptrs = (ctypes.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    wts = None
    if weights is not None:
        if method != 'avg_wave':
            weights = None
        else:
            w = _normalised_weights(weights).astype(np.float64)
            wts = (ctypes.c_double * len(xs))(*w.tolist())
    with torch.cuda.device(dev):
        call('sesa_ensemble_wave', ctypes.cast(ptrs, ctypes.c_void_p), len(xs),
             ctypes.cast(wts, ctypes.c_void_p) if wts is not None else None, WAVE_METHODS.index(method), out.data_ptr(),
             out.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def ensemble_separate(members, mix, device, stem='vocals', method='avg_wave', weights=None, to_host=True, engine_batch=None):
    """The in-memory form of the GUI's ensemble flow (processing.py:798-1188 runs inference.py once per model, then
    ensemble.py:258-407 averages the files): every member ``(config, model)`` separates ``mix`` on ``device``, its ``stem``
    estimate STAYS on the device, and one sesa_ensemble_wave launch reduces them (ensemble.py:172-183).  One upload of the
    mix, one download of the ensembled stem; no intermediate files, so no intermediate PCM quantisation either (equal to
    the reference flow with 'wav FLOAT' exports).  Returns ndarray / CUDA tensor [channels, samples]."""
    from .demix import DemixEngine
    dev = torch.device(device)
    if isinstance(mix, torch.Tensor) and mix.device.type == 'cuda':
        mix_d = mix.to(device=dev, dtype=torch.float32)
    else:
        mix_d = torch.as_tensor(np.asarray(mix), dtype=torch.float32).to(dev)
    stems = []
    for config, model in members:
        eng = DemixEngine(config, model, dev, engine_batch=engine_batch)
        if stem not in eng.instruments:
            raise KeyError(f'{type(model).__name__} does not produce {stem!r} (it has {eng.instruments})')
        est = eng.run(mix_d, to_host=False)
        stems.append(est[eng.instruments.index(stem)])
    out = ensemble_waveforms(stems, method, weights)
    if not to_host:
        return out
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host.numpy()


def ensemble_tracks(members, tracks, device, out_dir=None, stem='vocals', method='avg_wave', weights=None, rank=0, world=1,
                    sample_rate=44100):
    """BASELINE config 5: many tracks x several models.  Tracks are sharded over the ranks round-robin (independent work:
    no collective); ``tracks`` is a list of (name, mix) pairs or audio paths.  Every ensembled stem is written as PCM_24
    like ensemble.py:311 when ``out_dir`` is given.  Returns {name: ndarray} for this rank's tracks."""
    import os
    done = {}
    for i, item in enumerate(tracks):
        if i % world != rank:
            continue
        if isinstance(item, str):
            name = os.path.splitext(os.path.basename(item))[0]
            mix, _ = load_audio(item, sample_rate)
            mix = np.atleast_2d(mix)
        else:
            name, mix = item
        est = ensemble_separate(members, mix, device, stem=stem, method=method, weights=weights)
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            write_audio(os.path.join(out_dir, f'{name}_{stem}_ensemble.wav'), est.T, sample_rate, subtype='PCM_24')
        done[name] = est
    return done


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
