"""Drop-in for the reference's src/channel/awgn.py plus transmit_batch (batched LLR feed)."""
from polarcode_and_ldpc_b200.channel.awgn import AWGNChannel  # noqa: F401
