"""Stand-in for ml_collections.ConfigDict (absent here): attribute/item dict with nested wrapping.
Covers what the reference uses: attribute access, ``in``, ``[...]``, ``getattr(cfg, k, default)``,
``**dict(config.model)`` (utils.py:103,382-397,496)."""


class ConfigDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return ConfigDict(v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
