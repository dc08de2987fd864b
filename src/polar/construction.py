from polarcode_and_ldpc_b200.polar.construction import construct_polar_code, bhattacharyya_bounds  # noqa: F401
