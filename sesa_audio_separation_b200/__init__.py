"""sesa_audio_separation_b200 — B200-native (sm_100a) implementation of SESA's chunked separation
inference path behind the reference's own Python surface.  See DESIGN.md / INTEGRATION.md."""
from .config import ConfigDict, get_model_from_config, load_config, prefer_target_instrument  # noqa: F401
from .demix import (DemixEngine, apply_tta, demix, demix_pytorch_optimized, demix_tta, denormalize_audio,  # noqa: F401
                    normalize_audio)
from .ensemble import ensemble_waveforms  # noqa: F401
from .backend import PyTorchBackend, create_inference_session  # noqa: F401
from .roformer import BSRoformer, MelBandRoformer  # noqa: F401
from ._lib import SesaError  # noqa: F401

__all__ = ['ConfigDict', 'get_model_from_config', 'load_config', 'prefer_target_instrument', 'DemixEngine',
           'apply_tta', 'demix', 'demix_pytorch_optimized', 'demix_tta', 'ensemble_waveforms', 'normalize_audio', 'denormalize_audio',
           'PyTorchBackend', 'create_inference_session', 'BSRoformer', 'MelBandRoformer', 'SesaError']
