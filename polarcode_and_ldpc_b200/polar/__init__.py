"""Same import surface as the reference's src/polar/__init__.py:6."""
from .decoder import SCDecoder, SCLDecoder    # noqa: F401
from .encoder import PolarEncoder             # noqa: F401
from .utils import bit_reverse, generate_frozen_bits, crc_encode, crc_check  # noqa: F401
from .construction import bhattacharyya_frozen_set  # noqa: F401
