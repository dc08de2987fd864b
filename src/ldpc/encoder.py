from polarcode_and_ldpc_b200.ldpc.encoder import LDPCEncoder  # noqa: F401
