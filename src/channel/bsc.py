from polarcode_and_ldpc_b200.channel.bsc import *  # noqa: F401,F403
from polarcode_and_ldpc_b200.channel.bsc import BSCChannel  # noqa: F401
