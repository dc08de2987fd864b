from polarcode_and_ldpc_b200.polar.utils import *  # noqa: F401,F403
