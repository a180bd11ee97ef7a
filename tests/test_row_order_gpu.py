"""GPU parity tests of the row-order mode (``OBP_B200_ROW_ORDER=1``, ``obp_set_row_order``).

With the mode on the engine returns rows in EXACTLY the reference's order — each key at the place of
its first appearance among the gate's k*k raw rows (``utils/operations.py:103-105`` compose order,
``utils/simplify.py:126,160-165`` first-appearance merge) — so the reference's ordered KATs pass
literally and the row sequence equals the oracle's, not just the key set.
"""

from __future__ import annotations

import numpy as np
import pytest

import kats
from conftest import COEFF_ATOL, COEFF_RTOL, assert_same_metadata
from oracle import obp_oracle as oracle
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200.engine import EngineError, TermTable
from qiskit_addon_obp_b200.ir import (
    PauliLindbladError,
    PauliLindbladErrorInstruction,
    QuantumCircuit,
    SparsePauliOp,
)
from qiskit_addon_obp_b200.utils.simplify import OperatorBudget
from qiskit_addon_obp_b200.utils.truncating import setup_budget
from test_engine_gpu import random_observable, random_slices

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _row_order(monkeypatch):
    monkeypatch.setenv("OBP_B200_ROW_ORDER", "1")


def assert_same_rows(got: SparsePauliOp, want: SparsePauliOp, what="operator"):
    """Same rows in the same order; coefficients within the fp64 parity bar."""
    assert got.num_qubits == want.num_qubits
    assert len(got) == len(want), f"{what}: {len(got)} rows vs {len(want)}"
    gx, gz = got.packed()
    wx, wz = want.packed()
    bad = np.nonzero((gx != wx).any(axis=1) | (gz != wz).any(axis=1))[0]
    assert bad.size == 0, f"{what}: first differing row {bad[0]}: {got.paulis.to_labels()[bad[0]]} vs {want.paulis.to_labels()[bad[0]]}"
    np.testing.assert_allclose(got.coeffs, want.coeffs, rtol=COEFF_RTOL, atol=COEFF_ATOL, err_msg=what)


@pytest.mark.parametrize("kat", kats.ORDERED + kats.ORDER_FREE, ids=lambda f: f.__name__)
def test_reference_kat_in_reference_row_order(kat):
    if kat in kats.ORDERED:
        kat(backpropagate, ordered=True)
    else:
        kat(backpropagate)


def _compare(obs, slices, **kw):
    got, rest_g, md_g = backpropagate(obs, slices, **kw)
    want, rest_w, md_w = oracle.backpropagate(obs, slices, **kw)
    assert len(rest_g) == len(rest_w)
    single = isinstance(obs, SparsePauliOp)
    for k, (g, w) in enumerate(zip([got] if single else got, [want] if single else want)):
        assert_same_rows(g, w, f"observable {k}")
    assert_same_metadata(md_g, md_w)
    return got, md_g


@pytest.mark.parametrize("nq,seed", [(5, 0), (9, 1), (20, 2), (70, 3), (130, 4)])
def test_random_circuit_row_order(nq, seed):
    rng = np.random.default_rng(100 + seed)
    slices = random_slices(nq, 5, rng, p_rot=0.6)
    obs = random_observable(nq, 3, rng)
    _compare(obs, slices)


@pytest.mark.parametrize("seed", range(3))
def test_random_circuit_row_order_with_truncation(seed):
    rng = np.random.default_rng(200 + seed)
    nq = 8
    slices = random_slices(nq, 6, rng, p_rot=0.7, angles=[0.1, 0.2, 0.4, -0.3])
    obs = [random_observable(nq, 2, rng), random_observable(nq, 3, rng)]
    _compare(obs, slices, truncation_error_budget=setup_budget(max_error_per_slice=0.02))


def test_row_order_generic_gates_reset_and_noise():
    rng = np.random.default_rng(7)
    qc1 = QuantumCircuit(4)
    qc1.cp(0.37, 0, 1)
    qc1.crz(-0.8, 2, 3)
    qc1.h(0)
    qc2 = QuantumCircuit(4)
    qc2.reset(1)
    qc2.rxx(0.4, 0, 2)
    qc2.ecr(1, 3)
    gens = ["XI", "IZ", "YY", "ZX"]
    ple = PauliLindbladError(gens, rng.uniform(0.001, 0.05, size=len(gens)))
    qc3 = QuantumCircuit(4)
    qc3.append(PauliLindbladErrorInstruction(ple), [1, 2])
    qc3.ry(0.3, 1)
    qc3.cx(1, 2)
    obs = SparsePauliOp(["ZZII", "IXYI", "IIZZ", "XIIX"], [1.0, -0.5, 0.25, 0.7])
    _compare(obs, [qc1, qc2, qc3, qc1])
    _compare(obs, [qc3, qc2, qc1], truncation_error_budget=setup_budget(max_error_total=0.05, num_slices=3))


def test_row_order_unsimplified_input():
    """Duplicate and near-zero input rows reach the first gate as given (raw_num_paulis counts them)."""
    qc = QuantumCircuit(3)
    qc.rx(0.3, 0)
    qc.cx(0, 1)
    qc2 = QuantumCircuit(3)
    qc2.reset(0)
    qc2.rz(0.1, 1)
    obs = SparsePauliOp(["ZIX", "XYZ", "ZIX", "IIZ", "XYZ", "YYI"], [1.0, 0.5, 0.25, 1e-9, -0.2, 0.3])
    _compare(obs, [qc])
    _compare(obs, [qc2, qc])
    _compare(obs, [qc, qc2])


def test_row_order_operator_budget_stops_at_same_slice():
    rng = np.random.default_rng(11)
    slices = random_slices(6, 8, rng, p_rot=0.8)
    obs = random_observable(6, 2, rng)
    got, md = _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=40))
    assert md.num_backpropagated_slices < 8


def test_in_place_gate_kernels_refuse_row_order_mode():
    from qiskit_addon_obp_b200.gates import program_for

    t = TermTable.acquire(2, 1024)
    try:
        t.set_row_order(True)
        t.load(SparsePauliOp(["ZI"], [1.0]))
        qc = QuantumCircuit(2)
        qc.h(0)
        rec = program_for(qc.data[0]).template.copy()
        with pytest.raises(EngineError):
            t.apply_ops(rec, 1e-8)
    finally:
        t.set_row_order(False)
        t.release()
