from polarcode_and_ldpc_b200.channel.awgn import AWGNChannel  # noqa: F401
