"""The decode-path part of the reference's src/ldpc/__init__.py:6 surface (INTEGRATION.md lists what
is not re-exported: peg_construction)."""
from .decoder import BPDecoder, MSDecoder     # noqa: F401
from .encoder import LDPCEncoder              # noqa: F401
from .construction import (gallager_parity_check, mackay_parity_check, generator_from_parity,  # noqa: F401
                           generate_ldpc_matrix, mackay_construction, create_tanner_graph, check_syndrome)
