from polarcode_and_ldpc_b200.ldpc.construction import generate_ldpc_matrix, mackay_construction  # noqa: F401
