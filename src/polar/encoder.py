from polarcode_and_ldpc_b200.polar.encoder import PolarEncoder  # noqa: F401
