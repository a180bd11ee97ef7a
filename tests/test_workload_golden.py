"""Committed golden vectors of whole backpropagations (tests/golden/workload_golden.json).

CPU: the oracle must still reproduce them exactly as generated — labels in the reference's row order,
coefficients, per-slice counters, errors.  GPU: the CUDA engine must match them within the parity bars
without anything under ``oracle/`` being imported.
"""

from __future__ import annotations

import json
import os

import numpy as np
import pytest

from conftest import COEFF_ATOL, COEFF_RTOL, ERROR_TOL
from qiskit_addon_obp_b200 import workloads as W

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "workload_golden.json")))
CASES = {c["name"]: c for c in GOLDEN["cases"]}
MAKERS = {
    "config1": lambda: W.config1_xy_chain(),
    "config2_small": lambda: W.config2_kicked_ising(theta_h=0.3, steps=2, max_paulis=3000),
    "config3_small": lambda: W.config3_square_lattice(side=3, steps=2, max_paulis=5000),
}
COUNTERS = ("raw_num_paulis", "num_unique_paulis", "num_duplicate_paulis", "num_trimmed_paulis",
            "num_truncated_paulis", "num_paulis")


def _check(case, out, rest, md, *, ordered: bool):
    ops = [out] if not isinstance(out, list) else out
    assert len(rest) == case["remaining_slices"]
    assert md.num_backpropagated_slices == case["num_backpropagated_slices"]
    assert len(ops) == len(case["observables"])
    for o, g in zip(ops, case["observables"]):
        want = {lab: complex(re, im) for lab, re, im in zip(g["labels"], g["re"], g["im"])}
        labels = o.paulis.to_labels()
        if ordered:
            assert labels == g["labels"]
        got = dict(zip(labels, o.coeffs))
        assert set(got) == set(want)
        for k in want:
            assert abs(got[k] - want[k]) <= COEFF_ATOL + COEFF_RTOL * abs(want[k]), k
    assert len(md.backpropagation_history) == len(case["history"])
    for s, g in zip(md.backpropagation_history, case["history"]):
        np.testing.assert_allclose(s.slice_errors, g["slice_errors"], rtol=0, atol=ERROR_TOL)
        for f in COUNTERS:
            assert list(getattr(s, f)) == g[f], f
        assert s.sum_paulis == g["sum_paulis"]
    for i, e in enumerate(case["accumulated_error"]):
        assert abs(md.accumulated_error(i) - e) <= ERROR_TOL


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_golden_vectors(name):
    from oracle import obp_oracle as oracle

    wl = MAKERS[name]()
    _check(CASES[name], *oracle.backpropagate(wl.observables, wl.slices, **wl.kwargs()), ordered=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_matches_the_golden_vectors(name):
    from qiskit_addon_obp_b200 import backpropagate

    wl = MAKERS[name]()
    _check(CASES[name], *backpropagate(wl.observables, wl.slices, **wl.kwargs()), ordered=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_with_merged_clifford_slices_matches_the_golden_vectors(name, monkeypatch):
    """OBP_B200_MERGE_CLIFFORD_SLICES=1: consecutive Clifford-only slices travel as one gate program, each slice's
    counters read back with obp_last_op_stats — every per-slice number must come out as slice by slice."""
    from qiskit_addon_obp_b200 import backpropagate

    monkeypatch.setenv("OBP_B200_MERGE_CLIFFORD_SLICES", "1")
    wl = MAKERS[name]()
    _check(CASES[name], *backpropagate(wl.observables, wl.slices, **wl.kwargs()), ordered=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_engine_row_order_mode_matches_the_golden_row_order(name, monkeypatch):
    from qiskit_addon_obp_b200 import backpropagate

    monkeypatch.setenv("OBP_B200_ROW_ORDER", "1")
    wl = MAKERS[name]()
    _check(CASES[name], *backpropagate(wl.observables, wl.slices, **wl.kwargs()), ordered=True)
