"""``PyTorchBackend`` surface (reference: pytorch_backend.py:19-332,492-536) on the sm_100a engine.

The reference's backend is the object ``demix_pytorch_optimized`` calls per batch.  Here it is a thin
seat for the kernel engine: ``optimize_mode`` values that select other execution back ends
(``compile``, ``jit``) are refused — the north-star forbids multi-backend dispatch — and ``use_amp``
selects the engine's arithmetic mode (True -> 'bf16' tensor-core mode, False -> 'fp32' parity mode).
"""
import torch

from . import _lib
from .module import KernelModule


class PyTorchBackend:
    def __init__(self, device='cuda:0', optimize_mode='channels_last'):
        _lib.require_cuda()
        if not str(device).startswith('cuda'):
            raise _lib.SesaError(f'the B200 engine runs on CUDA only (got device={device!r})')
        if optimize_mode not in ('channels_last', 'default'):
            raise ValueError(f"optimize_mode={optimize_mode!r} selects another back end; only 'channels_last' "
                             "and 'default' (both = the sm_100a kernel engine) are accepted")
        self.device = str(device)
        self.optimize_mode = optimize_mode
        self.model = None
        self.compiled_model = None
        self.use_amp = True

    def optimize_model(self, model, example_input=None, use_amp=True, use_channels_last=True):
        if not isinstance(model, KernelModule):
            raise TypeError('PyTorchBackend.optimize_model expects a model built by '
                            'sesa_audio_separation_b200.get_model_from_config')
        self.model = model.eval().to(self.device)
        self.use_amp = use_amp
        if hasattr(self.model, 'set_precision'):
            self.model.set_precision('bf16' if use_amp else 'fp32')
        self.compiled_model = self.model
        return self.compiled_model

    def __call__(self, x):
        if self.model is None:
            raise RuntimeError("Model not optimized. Call optimize_model first.")
        return self.model(x)


def create_inference_session(model, device='cuda:0', optimize_mode='channels_last', enable_amp=True,
                             enable_tf32=True, enable_cudnn_benchmark=True):
    """pytorch_backend.py:492-536.  enable_tf32 / enable_cudnn_benchmark are accepted and ignored:
    no cuBLAS/cuDNN call exists on this path."""
    backend = PyTorchBackend(device=device, optimize_mode=optimize_mode)
    backend.optimize_model(model, use_amp=enable_amp)
    return backend
