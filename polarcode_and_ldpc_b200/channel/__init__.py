from .awgn import AWGNChannel                 # noqa: F401
from .bsc import BSCChannel                   # noqa: F401
from .fading import RayleighFadingChannel     # noqa: F401
