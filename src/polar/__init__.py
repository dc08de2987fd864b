from polarcode_and_ldpc_b200.polar import *  # noqa: F401,F403
from polarcode_and_ldpc_b200.polar import SCDecoder, SCLDecoder, PolarEncoder  # noqa: F401
