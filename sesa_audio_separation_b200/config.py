"""Model registry / YAML config loading for the hot path (reference: utils.py:26-161).

``ml_collections`` is not a dependency here: ``ConfigDict`` below provides the attribute/item access
the reference code relies on.  YAML is read with ``yaml.FullLoader`` like the reference (utils.py:54),
so ``!!python/tuple`` band tables load unchanged.
"""
import yaml


class ConfigDict(dict):
    """Attribute-style nested dict (the subset of ml_collections.ConfigDict the hot path uses)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def load_config(model_type: str, config_path: str) -> ConfigDict:
    """utils.py:26-59 (htdemucs/OmegaConf branch is out of scope)."""
    try:
        with open(config_path, 'r') as f:
            if model_type == 'htdemucs':
                raise NotImplementedError('htdemucs configs (OmegaConf) are out of scope of the B200 hot path')
            return ConfigDict(yaml.load(f, Loader=yaml.FullLoader))
    except FileNotFoundError:
        raise FileNotFoundError(f"Configuration file not found at {config_path}")
    except NotImplementedError:
        raise
    except Exception as e:
        raise ValueError(f"Error loading configuration: {e}")


def prefer_target_instrument(config):
    """utils.py:480-499."""
    if getattr(config.training, 'target_instrument', None):
        return [config.training.target_instrument]
    return config.training.instruments


SUPPORTED_MODEL_TYPES = ('bs_roformer', 'mel_band_roformer', 'mdx23c')


def build_model(model_type: str, config):
    if model_type == 'bs_roformer':
        from .roformer import BSRoformer
        return BSRoformer(**dict(config.model))
    if model_type == 'mel_band_roformer':
        from .roformer import MelBandRoformer
        return MelBandRoformer(**dict(config.model))
    if model_type == 'mdx23c':
        from .mdx23c import TFC_TDF_net
        return TFC_TDF_net(config)
    raise ValueError(f"Unknown model type: {model_type} (the B200 hot path covers {SUPPORTED_MODEL_TYPES})")


def get_model_from_config(model_type: str, config_path: str):
    """utils.py:62-161: returns (model, config)."""
    config = load_config(model_type, config_path)
    return build_model(model_type, config), config
