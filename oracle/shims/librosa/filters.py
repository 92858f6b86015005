from oracle.third_party import mel  # noqa: F401
