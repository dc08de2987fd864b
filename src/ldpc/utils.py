from polarcode_and_ldpc_b200.ldpc.construction import create_tanner_graph, check_syndrome  # noqa: F401
