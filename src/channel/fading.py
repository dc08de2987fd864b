from polarcode_and_ldpc_b200.channel.fading import *  # noqa: F401,F403
from polarcode_and_ldpc_b200.channel.fading import RayleighFadingChannel  # noqa: F401
