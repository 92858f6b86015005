"""Minimal parameter container mirroring the slice of the nn.Module surface the reference's callers
use on a model (inference_pytorch.py:324-369, pytorch_backend.py:95-110): state_dict /
load_state_dict(strict) / eval / to / parameters / __call__.  Tensors are plain torch tensors used
as device-memory handles; all arithmetic happens in the CUDA library."""
import math
from collections import OrderedDict

import torch


class KernelModule:
    def __init__(self):
        self._params = OrderedDict()   # name -> fp32 tensor (reference state_dict layout)
        self._device = torch.device('cpu')
        self._prepared = None
        self.training = False

    # -- construction helpers
    def _register(self, name, shape, kind, generator):
        shape = tuple(int(s) for s in shape)
        if kind == 'ones':
            t = torch.ones(shape)
        elif kind == 'zeros':
            t = torch.zeros(shape)
        elif kind == 'linear_w':   # nn.Linear / conv default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(max(1, math.prod(shape[1:])))
            t = (torch.rand(shape, generator=generator) * 2 - 1) * bound
        elif isinstance(kind, tuple) and kind[0] == 'uniform':
            t = (torch.rand(shape, generator=generator) * 2 - 1) * kind[1]
        else:
            raise ValueError(kind)
        self._params[name] = t

    # -- per-input-shape workspaces
    WORKSPACE_SLOTS = 2   # the full engine batch and a track's ragged tail batch: neither is rebuilt per track

    def _keep_workspace(self, key, ws):
        """Remember ``ws`` for input shape ``key``, evicting the least recently built one beyond WORKSPACE_SLOTS (every
        workspace holds multi-GB activation buffers and host-encoded TMA tables, so rebuilding one per launch is the
        expensive thing to avoid and holding many is the other)."""
        self._ws.pop(key, None)
        while len(self._ws) >= self.WORKSPACE_SLOTS:
            self._ws.pop(next(iter(self._ws)))
        self._ws[key] = ws

    # -- nn.Module-like surface
    def state_dict(self):
        return OrderedDict((k, v) for k, v in self._params.items())

    def load_state_dict(self, state_dict, strict=True):
        missing = [k for k in self._params if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._params]
        if strict and (missing or unexpected):
            raise RuntimeError(f'Error(s) in loading state_dict: missing {missing[:5]}..., unexpected {unexpected[:5]}...')
        for k, v in state_dict.items():
            if k not in self._params:
                continue
            if tuple(v.shape) != tuple(self._params[k].shape):
                raise RuntimeError(f'size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {tuple(self._params[k].shape)}')
            self._params[k] = v.detach().to(device=self._device, dtype=torch.float32).clone()
        self._prepared = None
        return missing, unexpected

    def parameters(self):
        return iter(self._params.values())

    def named_parameters(self):
        return iter(self._params.items())

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError('the B200 engine is inference-only')
        return self

    def to(self, *args, **kwargs):
        device = kwargs.get('device')
        for a in args:
            if isinstance(a, (str, torch.device)):
                device = a
        if kwargs.get('memory_format') is not None and device is None:
            return self   # channels_last request from PyTorchBackend.optimize_model: layouts are the engine's business
        if device is not None:
            device = torch.device(device)
            if device != self._device:
                self._params = OrderedDict((k, v.to(device)) for k, v in self._params.items())
                self._device = device
                self._prepared = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device('cuda', torch.cuda.current_device() if device is None else device))

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)
