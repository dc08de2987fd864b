from polarcode_and_ldpc_b200.channel import AWGNChannel, BSCChannel, RayleighFadingChannel  # noqa: F401
