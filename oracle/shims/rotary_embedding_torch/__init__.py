from oracle.third_party import RotaryEmbedding  # noqa: F401
