"""pytest configuration: the ``gpu`` marker and shared comparison helpers.

``-m "not gpu"`` runs the oracle against the reference's known answers, the host logic and the
C-ABI symbol check (no GPU needed); ``-m gpu`` runs the parity tests proper: the CUDA engine,
called through the C-ABI, against the oracle.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# north-star parity gates (BASELINE.json): key set bit-exact, coefficients 1e-10 rel / 1e-12 abs,
# error totals 1e-9; terms whose |c| lies within 1e-12 of a trim/truncation threshold may differ
COEFF_RTOL = 1e-10
COEFF_ATOL = 1e-12
ERROR_TOL = 1e-9
THRESHOLD_BAND = 1e-12


def assert_same_operator(got, want, *, thresholds=(1e-8,), what="operator"):
    """Bit-exact key set (modulo the threshold band) and fp64-tolerance coefficients."""
    assert got.num_qubits == want.num_qubits
    a, b = got.as_dict(), want.as_dict()
    assert len(a) == len(got), f"{what}: duplicate keys in the engine result"
    only_a = [k for k in a if k not in b]
    only_b = [k for k in b if k not in a]
    for k in only_a:
        assert any(abs(abs(a[k]) - t) <= THRESHOLD_BAND for t in thresholds), (
            f"{what}: key {k} only in engine result, |c|={abs(a[k])}"
        )
    for k in only_b:
        assert any(abs(abs(b[k]) - t) <= THRESHOLD_BAND for t in thresholds), (
            f"{what}: key {k} only in oracle result, |c|={abs(b[k])}"
        )
    for k in a:
        if k in b:
            assert abs(a[k] - b[k]) <= COEFF_ATOL + COEFF_RTOL * abs(b[k]), (
                f"{what}: coefficient of {k}: {a[k]} vs {b[k]}"
            )


def assert_same_metadata(got, want, *, exact_counters=True, allow_missing_counters=False):
    assert got.num_backpropagated_slices == want.num_backpropagated_slices
    assert len(got.backpropagation_history) == len(want.backpropagation_history)
    for s, (g, w) in enumerate(zip(got.backpropagation_history, want.backpropagation_history)):
        np.testing.assert_allclose(g.slice_errors, w.slice_errors, rtol=0, atol=ERROR_TOL, err_msg=f"slice {s}")
        assert g.num_paulis == w.num_paulis, f"slice {s} num_paulis"
        assert g.num_truncated_paulis == w.num_truncated_paulis, f"slice {s} num_truncated"
        assert g.sum_paulis == w.sum_paulis, f"slice {s} sum_paulis"
        assert g.num_qwc_groups == w.num_qwc_groups, f"slice {s} qwc"
        if exact_counters:
            assert g.raw_num_paulis == w.raw_num_paulis, f"slice {s} raw_num_paulis"
            for i in range(len(w.num_paulis)):
                if g.num_unique_paulis[i] is None:
                    # sharded table, slice ended on a Clifford-class gate: counters not reported
                    assert allow_missing_counters, f"slice {s}: simplify counters missing"
                    continue
                assert g.num_unique_paulis[i] == w.num_unique_paulis[i], f"slice {s} num_unique"
                assert g.num_duplicate_paulis[i] == w.num_duplicate_paulis[i], f"slice {s} num_duplicate"
                assert g.num_trimmed_paulis[i] == w.num_trimmed_paulis[i], f"slice {s} num_trimmed"
                assert abs(complex(g.sum_trimmed_coeffs[i]) - complex(w.sum_trimmed_coeffs[i])) <= 1e-12, (
                    f"slice {s} sum_trimmed"
                )
    for i in range(len(want.backpropagation_history[0].slice_errors) if want.backpropagation_history else 0):
        assert abs(got.accumulated_error(i) - want.accumulated_error(i)) <= ERROR_TOL
