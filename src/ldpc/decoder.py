"""Drop-in for the reference's src/ldpc/decoder.py (BPDecoder :11, MSDecoder :208)."""
from polarcode_and_ldpc_b200.ldpc.decoder import BPDecoder, MSDecoder  # noqa: F401
