"""Chunked separation inference: the B200-native ``demix`` (reference: utils.py:330-477 and its
copy inference_pytorch.py:55-186).

What changes relative to the reference loop: the mix is uploaded ONCE, border padding, chunk framing,
the model forward, the windowed overlap-add, the divide and the crop all run on the device, and the
result comes back in ONE device->host copy (the reference does an H2D and a D2H + sync per chunk and
accumulates on the CPU).  What does not change: the chunk schedule, pad modes, per-flush window rule
and the ascending-order accumulation, which are reproduced bit-exactly (plan.py, sesa_overlap_add).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import call
from .config import prefer_target_instrument
from .module import KernelModule
from .plan import make_plan, shard_chunks, windowing_array


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _unwrap(model):
    """Accept the engine model or a PyTorchBackend-style wrapper holding one."""
    inner = getattr(model, 'model', None)
    if isinstance(inner, KernelModule):
        return inner
    if isinstance(model, KernelModule):
        return model
    raise TypeError('demix() drives the sm_100a engine: pass a model built by '
                    'sesa_audio_separation_b200.get_model_from_config (got %r). There is no PyTorch fallback.'
                    % type(model).__name__)


class DemixEngine:
    """Device-resident demix of one track.  ``engine_batch`` chunks go through the model per launch
    group; it is a throughput knob only and never changes the result's bookkeeping."""

    def __init__(self, config, model, device, engine_batch=None, world=1, rank=0, progress=None, group=None):
        _lib.require_cuda()
        self.model = _unwrap(model)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.SesaError(f'demix needs a CUDA device, got {device!r}; there is no CPU path')
        self.model.to(self.device)
        self.chunk_size = int(config.audio.chunk_size)
        self.num_overlap = int(config.inference.num_overlap)
        self.batch_size = int(config.inference.batch_size)
        self.instruments = list(prefer_target_instrument(config))
        # chunks per launch group: a throughput knob only (results are batch-invariant); 4 fills the 148 SMs' tile
        # waves of every GEMM at the BASELINE model sizes and keeps the workspace under 40 GB
        self.engine_batch = int(engine_batch or 4)
        self.world, self.rank, self.group = world, rank, group
        self.progress = progress

    def run(self, mix, return_counter=False, to_host=True):
        dev = self.device
        with torch.cuda.device(dev):
            return self._run(mix, return_counter, to_host)

    def _run(self, mix, return_counter, to_host):
        dev = self.device
        if isinstance(mix, torch.Tensor) and mix.device.type == 'cuda':
            mix_d = mix.to(device=dev, dtype=torch.float32).contiguous()
        else:
            mix_h = torch.as_tensor(np.asarray(mix), dtype=torch.float32).contiguous()
            mix_d = mix_h.to(dev, non_blocking=True)
        if mix_d.ndim != 2:
            raise ValueError('mix must have shape (channels, time)')
        C, length = mix_d.shape
        L = self.chunk_size
        plan = make_plan(length, L, self.num_overlap, self.batch_size)
        st = _stream()
        if plan.pad:
            padded = torch.empty(C, plan.padded, device=dev, dtype=torch.float32)
            call('sesa_pad_reflect', _ptr(mix_d), _ptr(padded), C, length, plan.border, plan.border, st)
        else:
            padded = mix_d
        n_inst = len(self.instruments)
        lo, hi = shard_chunks(plan.n_chunks, self.world, self.rank)
        starts = torch.tensor(plan.starts, dtype=torch.int64).to(dev)
        lens = torch.tensor(plan.lens, dtype=torch.int64).to(dev)
        modes = torch.tensor(plan.modes, dtype=torch.int32).to(dev)
        kinds = torch.tensor(plan.kinds, dtype=torch.int32).to(dev)
        chunk_out = torch.empty(max(hi - lo, 1), n_inst, C, L, device=dev, dtype=torch.float32)
        EB = self.engine_batch
        chunks = torch.empty(EB, C, L, device=dev, dtype=torch.float32)
        k = lo
        last_pct = -1
        while k < hi:
            nb = min(EB, hi - k)
            call('sesa_frame_chunks', _ptr(padded), plan.padded, C, _ptr(starts) if k == 0 else ctypes.c_void_p(starts.data_ptr() + 8 * k),
                 ctypes.c_void_p(lens.data_ptr() + 8 * k), ctypes.c_void_p(modes.data_ptr() + 4 * k),
                 nb, L, _ptr(chunks), st)
            y = self.model.forward(chunks[:nb])
            y = y.reshape(nb, n_inst, C, -1)
            if y.shape[-1] != L:
                raise RuntimeError(f'model returned {y.shape[-1]} samples for a {L}-sample chunk '
                                   '(chunk_size must be a multiple of the hop for this model)')
            chunk_out[k - lo:k - lo + nb].copy_(y)
            k += nb
            if self.progress is not None:
                pct = int(min(1.0, (plan.starts[k - 1] + plan.step) / plan.padded) * 100)
                if pct > last_pct:
                    last_pct = pct
                    self.progress(pct)
        crop = plan.border if plan.pad else 0
        result = torch.empty(n_inst, C, length, device=dev, dtype=torch.float32)
        counter = torch.empty(plan.padded, device=dev, dtype=torch.float32) if return_counter else None
        window = windowing_array(L, plan.fade).to(dev)
        if self.world == 1:
            call('sesa_overlap_add', _ptr(chunk_out), _ptr(starts), _ptr(lens), _ptr(kinds), plan.n_chunks, plan.step,
                 L, plan.fade, _ptr(window), n_inst, C, plan.padded, crop, length,
                 _ptr(result), _ptr(counter), st)
        else:
            from .distributed import sharded_overlap_add
            nrows = n_inst * C

            def _range(p0, p1, init, init_p0, mode, out):
                call('sesa_overlap_add_range', _ptr(chunk_out), _ptr(starts), _ptr(lens), _ptr(kinds), plan.n_chunks,
                     lo, hi, plan.step, L, plan.fade, _ptr(window), n_inst, C, p0, p1, _ptr(init), init_p0,
                     init.shape[1] if init is not None else 0, mode, crop, length, _ptr(out), st)

            class _Ops:
                @staticmethod
                def raw(p0, p1):
                    out = torch.empty(nrows, p1 - p0, device=dev, dtype=torch.float32)
                    _range(p0, p1, None, 0, 1, out)
                    return out

                @staticmethod
                def final(p0, p1, init, init_p0):
                    full = torch.zeros(nrows, length, device=dev, dtype=torch.float32)
                    _range(p0, p1, init, init_p0, 0, full)
                    q0 = max(p0, crop) - crop
                    q1 = max(min(p1, crop + length) - crop, q0)
                    return full[:, q0:q1]
            res = sharded_overlap_add(plan, self.world, self.rank, nrows, _Ops, dev, group=self.group)
            self.plan = plan
            if res is None:
                return None
            result = res.view(n_inst, C, length)
            if return_counter:
                call('sesa_overlap_add', _ptr(chunk_out), _ptr(starts), _ptr(lens), _ptr(kinds), plan.n_chunks, plan.step,
                     L, plan.fade, _ptr(window), n_inst, C, plan.padded, crop, 0, None, _ptr(counter), st)
        self.plan = plan
        if not to_host:
            return (result, counter) if return_counter else result
        est = result.cpu().numpy()
        return (est, counter.cpu().numpy()) if return_counter else est


def demix(config, model, mix, device, model_type, pbar=False, engine_batch=None):
    """Drop-in for utils.demix (utils.py:330-477), generic mode.  Returns {instrument: ndarray(C, len)}."""
    if model_type == 'htdemucs':
        raise NotImplementedError('htdemucs (demucs mode of demix) is out of scope of the B200 hot path')
    eng = DemixEngine(config, model, device, engine_batch=engine_batch)
    est = eng.run(mix)
    return {k: v for k, v in zip(eng.instruments, est)}


def demix_pytorch_optimized(config, backend, mix, device, pbar=False, engine_batch=None):
    """Drop-in for inference_pytorch.demix_pytorch_optimized (:55-186): same result as demix() and the
    ``[SESA_PROGRESS]<int>`` stdout protocol the GUI parses (processing.py:345-359)."""
    eng = DemixEngine(config, backend, device, engine_batch=engine_batch,
                      progress=lambda p: print(f"[SESA_PROGRESS]{p}", flush=True))
    est = eng.run(mix)
    print("[SESA_PROGRESS]100", flush=True)
    return {k: v for k, v in zip(eng.instruments, est)}


def normalize_audio(audio):
    """utils.py:199-217."""
    mono = audio.mean(0)
    mean, std = mono.mean(), mono.std()
    return (audio - mean) / std, {"mean": mean, "std": std}


def denormalize_audio(audio, norm_params):
    """utils.py:220-238."""
    return audio * norm_params["std"] + norm_params["mean"]


# Test-time augmentations of utils.py:241-292 as (forward transform, inverse transform) pairs: channel swap and
# polarity inversion.  The result is the plain mean of the un-augmented estimate and the inverted augmented estimates,
# accumulated in the reference's order (+ swap, - polarity, / 3) so the arithmetic is identical.
_TTA_VARIANTS = (
    (lambda m: m[::-1].copy(), lambda w: w[::-1].copy(), +1.0),
    (lambda m: -1.0 * m.copy(), lambda w: w, -1.0),
)


def apply_tta(config, model, mix, waveforms_orig, device, model_type):
    """Drop-in for utils.apply_tta: averages ``waveforms_orig`` with the estimates of the augmented mixes (in place)."""
    for forward, inverse, sign in _TTA_VARIANTS:
        estimates = demix(config, model, forward(mix), device, model_type=model_type)
        for name, wave in estimates.items():
            if sign > 0:
                waveforms_orig[name] += inverse(wave)
            else:
                waveforms_orig[name] -= inverse(wave)
    scale = len(_TTA_VARIANTS) + 1
    for name in waveforms_orig:
        waveforms_orig[name] /= scale
    return waveforms_orig
