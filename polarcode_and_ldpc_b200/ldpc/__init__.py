"""Same import surface as the reference's src/ldpc/__init__.py:6."""
from .decoder import BPDecoder, MSDecoder     # noqa: F401
from .encoder import LDPCEncoder              # noqa: F401
from .construction import gallager_parity_check, mackay_parity_check, generator_from_parity  # noqa: F401
