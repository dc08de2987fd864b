"""Re-pointed benchmark callers: same entry points as the reference's
scripts; the shared Monte-Carlo loop against a host restatement of the reference's loop."""
import importlib.util
import os

import numpy as np
import pytest

import polarcode_and_ldpc_b200 as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location("bench_" + name, os.path.join(ROOT, "benchmarks", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_callers_keep_the_reference_entry_points():
    """Function names of /root/reference/benchmarks/*.py that other code imports or runs."""
    want = {"test_snr_curves": ["simulate_snr_curve", "test_multiple_rates", "analyze_snr_requirements", "main"],
            "ber_simulation": ["run_ber_simulation", "simulate_polar", "simulate_ldpc"],
            "test_code_parameters": ["test_code_lengths", "test_code_rates", "main"],
            "sc_vs_scl": ["simulate_sc_vs_scl"],
            "benchmark_scl": ["frame_error_rate_simulation"],
            "throughput_test": ["run_throughput_test"]}
    for mod, names in want.items():
        m = _load(mod)
        for n in names:
            assert callable(getattr(m, n)), f"{mod}.{n}"


@pytest.mark.gpu
def test_simulate_point_counts_like_the_reference_loop():
    """simulate_point's counters equal the reference loop's bookkeeping (np.sum(message !=
    decoded) per frame, test_snr_curves.py:133-141) evaluated on the host for the same frames."""
    N, K = 256, 128
    code = P.make_polar_code(N, K, 2.0)
    decs = {"sc": P.SCDecoder(N, K, frozen_bits=code["frozen_bits"]),
            "scl4": P.SCLDecoder(N, K, list_size=4, frozen_bits=code["frozen_bits"])}
    r = P.simulate_point(code, decs, 1.0, 3000, None, seed=5, first_chunk=1024)
    llr, msg, _ = code["gen"].generate(3000, 1.0, seed=5)
    for name, dec in decs.items():
        out = dec.decode_batch(llr).cpu().numpy()
        errs = (out != msg.cpu().numpy()).sum(axis=1)
        assert r[name]["frames_tested"] == 3000 and r[name]["total_bits"] == 3000 * K
        assert r[name]["error_bits"] == int(errs.sum()) and r[name]["frame_errors"] == int((errs > 0).sum())
    assert r["scl4"]["frame_errors"] <= r["sc"]["frame_errors"]
    # max_errors stops between chunks
    r2 = P.simulate_point(code, {"sc": decs["sc"]}, -2.0, 100000, 50, seed=1, first_chunk=512)
    assert r2["sc"]["frames_tested"] == 512 and r2["sc"]["frame_errors"] >= 50
    # LDPC: message bits sit at the generator's information positions
    lc = P.make_ldpc_code(504)
    r3 = P.simulate_point(lc, {"bp": P.BPDecoder(lc["H"], max_iter=20)}, 2.0, 2000, None, seed=2)
    assert r3["bp"]["total_bits"] == 2000 * lc["K"] and r3["bp"]["fer"] < 0.01
