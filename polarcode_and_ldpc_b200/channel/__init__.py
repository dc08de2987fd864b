from .awgn import AWGNChannel                 # noqa: F401
