from polarcode_and_ldpc_b200.ldpc import *  # noqa: F401,F403
from polarcode_and_ldpc_b200.ldpc import BPDecoder, MSDecoder, LDPCEncoder  # noqa: F401
