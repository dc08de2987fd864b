"""B200-native batched Polar (SC / SCL) and LDPC (BP / Min-Sum) decoders.

Drop-in for the decode hot path of B1ear/PolarCode_and_LDPC
(src/polar/decoder.py, src/ldpc/decoder.py): same classes, constructors and
decode(llr), plus decode_batch(llr[F, N]); all decoding runs in hand-written
sm_100a CUDA kernels behind the C ABI in include/pcl.h.  Importing the package
does not touch CUDA; constructing a decoder does, and fails loudly without it.
"""
from .polar.decoder import SCDecoder, SCLDecoder          # noqa: F401
from .polar.encoder import PolarEncoder                    # noqa: F401
from .polar.construction import bhattacharyya_frozen_set   # noqa: F401
from .ldpc.decoder import BPDecoder, MSDecoder             # noqa: F401
from .ldpc.encoder import LDPCEncoder                      # noqa: F401
from .ldpc.construction import gallager_parity_check, mackay_parity_check, generator_from_parity  # noqa: F401
from .channel import AWGNChannel, BSCChannel, RayleighFadingChannel  # noqa: F401
from .sweep import ErrorCounters, count_errors, shard_range  # noqa: F401
from .framegen import FrameGenerator                        # noqa: F401
from .simulate import make_ldpc_code, make_polar_code, simulate_point  # noqa: F401

__version__ = "0.1.0"
