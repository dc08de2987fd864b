"""The decode-path part of the reference's src/polar/__init__.py:6 surface (INTEGRATION.md lists
what is not re-exported: the Gaussian-approximation construction)."""
from .decoder import SCDecoder, SCLDecoder    # noqa: F401
from .encoder import PolarEncoder             # noqa: F401
from .utils import bit_reverse, generate_frozen_bits, crc_encode, crc_check  # noqa: F401
from .construction import bhattacharyya_frozen_set, bhattacharyya_bounds, construct_polar_code  # noqa: F401
