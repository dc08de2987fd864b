"""Drop-in for the reference's src/polar/decoder.py (SCDecoder :12, SCLDecoder :176)."""
from polarcode_and_ldpc_b200.polar.decoder import SCDecoder, SCLDecoder  # noqa: F401
