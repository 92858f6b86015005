"""Audio file I/O for the CLI drivers (reference: librosa.load at inference_pytorch.py:213, sf.write at :272).

``librosa`` / ``soundfile`` are used when importable (they are the reference's own I/O: same resampler, FLAC); otherwise a
self-contained WAV path is used: scipy.io.wavfile for reading (PCM 8/16/24/32 and float), polyphase resampling to the
config sample rate, and an own RIFF writer for FLOAT / PCM_16 / PCM_24.  Host-side only; never on the timed path.
"""
import os
import struct

import numpy as np


def load_audio(path, sample_rate):
    """-> (mix float32 [channels, samples], sample_rate): the contract of librosa.load(path, sr=..., mono=False)."""
    try:
        # the reference's own call (inference_pytorch.py:213) whenever librosa is installed: identical decoding and the
        # soxr resampler it uses when the file's rate differs from the config's
        import librosa
        mix, sr = librosa.load(path, sr=sample_rate, mono=False)
        return np.ascontiguousarray(mix, dtype=np.float32), sr
    except ImportError:
        pass
    try:
        import soundfile as sf
        data, sr = sf.read(path, dtype='float32', always_2d=True)
        mix = data.T
    except ImportError:
        from scipy.io import wavfile
        sr, data = wavfile.read(path)
        if data.ndim == 1:
            data = data[:, None]
        if data.dtype == np.uint8:
            mix = (data.astype(np.float32) - 128.0) / 128.0
        elif np.issubdtype(data.dtype, np.integer):
            mix = data.astype(np.float32) / float(2 ** (8 * data.dtype.itemsize - 1))
        else:
            mix = data.astype(np.float32)
        mix = mix.T
    if sr != sample_rate:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(sr), int(sample_rate))
        mix = resample_poly(mix, int(sample_rate) // g, int(sr) // g, axis=1).astype(np.float32)
    if mix.shape[0] == 1:
        mix = mix[0]      # librosa returns a 1-D array for mono files
    return np.ascontiguousarray(mix, dtype=np.float32), sample_rate


def _write_wav(path, data, sr, subtype):
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    n, ch = data.shape
    if subtype == 'FLOAT':
        fmt, bits, payload = 3, 32, np.ascontiguousarray(data.astype('<f4')).tobytes()
    elif subtype == 'PCM_16':
        q = np.clip(np.rint(data.astype(np.float64) * 32768.0), -32768, 32767).astype('<i2')
        fmt, bits, payload = 1, 16, np.ascontiguousarray(q).tobytes()
    elif subtype == 'PCM_24':
        q = np.clip(np.rint(data.astype(np.float64) * 8388608.0), -8388608, 8388607).astype('<i4')
        b = np.ascontiguousarray(q).view(np.uint8).reshape(n, ch, 4)[:, :, :3]
        fmt, bits, payload = 1, 24, np.ascontiguousarray(b).tobytes()
    else:
        raise ValueError(f'unsupported WAV subtype {subtype!r}')
    block = ch * bits // 8
    with open(path, 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', 36 + len(payload)) + b'WAVE')
        f.write(b'fmt ' + struct.pack('<IHHIIHH', 16, fmt, ch, sr, sr * block, block, bits))
        f.write(b'data' + struct.pack('<I', len(payload)))
        f.write(payload)


def write_audio(path, data, sr, subtype='FLOAT'):
    """sf.write(path, data[samples, channels], sr, subtype=...)."""
    try:
        import soundfile as sf
        sf.write(path, data, sr, subtype=subtype)
        return
    except ImportError:
        pass
    if os.path.splitext(path)[1].lower() != '.wav':
        raise RuntimeError(f'writing {os.path.splitext(path)[1]} needs the soundfile package; use a wav export format')
    _write_wav(path, data, int(sr), subtype)
