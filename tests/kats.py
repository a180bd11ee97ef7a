"""Known-answer tests of the reference, restated against the neutral host IR.

Every function takes ``bp`` — a ``backpropagate`` implementation (the oracle on the CPU, the CUDA
engine on the GPU) — and asserts what the reference's own test-suite asserts for the same inputs.
Sources: ``/root/reference/test/test_backpropagation.py`` (line ranges in each docstring).
"""

from __future__ import annotations

import numpy as np
import pytest

from qiskit_addon_obp_b200.ir import (
    PauliLindbladError,
    PauliLindbladErrorInstruction,
    QuantumCircuit,
    SparsePauliOp,
    slice_by_barriers,
    slice_by_depth,
)
from qiskit_addon_obp_b200.utils.simplify import OperatorBudget
from qiskit_addon_obp_b200.utils.truncating import setup_budget

THETA = np.pi / 6


def _labels(op) -> set[str]:
    return set(op.paulis.to_labels())


def _ry_mat(theta):
    return np.array([[np.cos(theta / 2), -np.sin(theta / 2)], [np.sin(theta / 2), np.cos(theta / 2)]])


def kat_single_slice(bp):
    """test_backpropagation.py:34-64, 101-127 — RY on X; RZ inside / RY outside the light cone."""
    target = _ry_mat(THETA).T @ np.array([[0, 1], [1, 0]]) @ _ry_mat(THETA)
    qc = QuantumCircuit(1)
    qc.ry(THETA, 0)
    new_obs, rest, _ = bp([SparsePauliOp("X")], [qc])
    assert np.allclose(target, new_obs[0].to_matrix())
    assert len(rest) == 0
    new_obs, rest, _ = bp(SparsePauliOp("X"), [qc])  # non-list input -> non-list output
    assert isinstance(new_obs, SparsePauliOp)
    assert np.allclose(target, new_obs.to_matrix())
    assert rest == []

    qc = QuantumCircuit(2)
    qc.rz(THETA, 0)
    qc.ry(THETA, 1)
    new_obs, rest, _ = bp([SparsePauliOp("IX")], [qc])
    assert _labels(new_obs[0]) == {"IX", "IY"}
    assert rest == []


def kat_depth2_clifford(bp):
    """:65-73 — H then CX maps ZZ to ZI (order-sensitive equality)."""
    qc = QuantumCircuit(2)
    qc.h(0)
    qc.cx(0, 1)
    new_obs, rest, _ = bp(SparsePauliOp("ZZ"), [qc])
    assert new_obs == SparsePauliOp("ZI")
    assert rest == []


def kat_scattered_qargs(bp):
    """:74-99 — empty slice, gate outside the light cone, gate growing the support."""
    qc = QuantumCircuit(5)
    new_obs, rest, _ = bp(SparsePauliOp("XIYIZ"), [qc])
    assert new_obs == SparsePauliOp("XIYIZ") and rest == []
    qc = QuantumCircuit(5)
    qc.cx(2, 3)
    new_obs, rest, _ = bp(SparsePauliOp("ZIIIZ"), [qc])
    assert new_obs == SparsePauliOp("ZIIIZ")
    qc = QuantumCircuit(5)
    qc.cx(2, 4)
    new_obs, rest, _ = bp(SparsePauliOp("ZIIIZ"), [qc])
    assert new_obs == SparsePauliOp("ZIZIZ") and rest == []


def kat_multiple_slices(bp):
    """:129-149 — RZ(both) | CX against dense matrices."""
    Z = np.diag([1.0, -1.0])
    RZ = np.diag([np.exp(-1j * THETA / 2), np.exp(1j * THETA / 2)])
    CX = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]])
    RZRZ = np.kron(RZ, RZ)
    target = RZRZ.conj().T @ (CX.T @ np.kron(Z, Z) @ CX) @ RZRZ
    qc = QuantumCircuit(2)
    qc.rz(THETA, [0, 1])
    qc.barrier()
    qc.cx(0, 1)
    new_obs, rest, _ = bp([SparsePauliOp("ZZ")], slice_by_barriers(qc))
    assert np.allclose(target, new_obs[0].to_matrix())
    assert rest == []


def kat_max_error_total(bp):
    """:150-166 — Z coefficient 0.5 survives a 0.49 budget and is truncated at 0.5."""
    qc = QuantumCircuit(1)
    qc.ry(THETA, 0)
    obs = SparsePauliOp("X")
    new_obs, _, _ = bp([obs], [qc], truncation_error_budget=setup_budget(max_error_total=0.49))
    assert len(new_obs[0]) == 2
    new_obs, rest, _ = bp([obs], [qc], truncation_error_budget=setup_budget(max_error_total=0.5))
    assert len(new_obs[0]) == 1
    assert rest == []


def kat_per_slice_list(bp):
    """:167-195 — per-slice budgets as a list."""
    qc1, qc2 = QuantumCircuit(1), QuantumCircuit(1)
    qc1.ry(THETA, 0)
    qc2.rz(THETA, 0)
    obs = SparsePauliOp("X")
    for budgets, n in (([0.01, 0.01], 3), ([0.4, 0.5], 2), ([0.44, 0.5], 1)):
        new_obs, _, _ = bp(obs, [qc1, qc2], truncation_error_budget=setup_budget(max_error_per_slice=budgets))
        assert len(new_obs) == n, budgets


def kat_barrier_and_reset(bp):
    """:196-215 — barriers inside a slice are skipped; reset maps ZX to IX."""
    qc = QuantumCircuit(1)
    qc.ry(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    _, rest, _ = bp([SparsePauliOp("X")], [qc])
    assert rest == []
    qc = QuantumCircuit(2)
    qc.reset(1)
    new_obs, rest, _ = bp(SparsePauliOp("ZX"), [qc])
    assert rest == []
    assert new_obs == SparsePauliOp("IX")


def kat_reset_simplify(bp):
    """:216-225 (simplify=True branch) — X and Y vanish under reset, Z becomes I."""
    qc = QuantumCircuit(1)
    qc.reset(0)
    obs = SparsePauliOp(["X", "Y", "Z"])
    new_obs, _, _ = bp(obs, [qc], operator_budget=OperatorBudget(simplify=True))
    assert len(new_obs) == 1


def kat_reset_no_simplify(bp):
    """:216-221 (simplify=False branch) — all three rows are kept."""
    qc = QuantumCircuit(1)
    qc.reset(0)
    obs = SparsePauliOp(["X", "Y", "Z"])
    new_obs, _, _ = bp(obs, [qc], operator_budget=OperatorBudget(simplify=False))
    assert len(new_obs) == 3


def kat_total_and_per_slice(bp):
    """:226-293 — combinations of max_error_total and max_error_per_slice."""
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    obs = SparsePauliOp("X")
    slices = slice_by_barriers(qc)
    cases = [
        (dict(max_error_per_slice=0.44), {"X", "Z"}),
        (dict(max_error_total=0.44), {"X", "Z"}),
        (dict(max_error_total=1, max_error_per_slice=0.5), {"X"}),
        (dict(max_error_total=0.8, max_error_per_slice=0.5), {"X", "Y"}),
        (dict(max_error_total=0.001, max_error_per_slice=np.inf), {"X", "Y", "Z"}),
        (dict(max_error_total=np.inf, max_error_per_slice=0.001), {"X", "Y", "Z"}),
    ]
    for kw, want in cases:
        new_obs, rest, _ = bp([obs], slices, truncation_error_budget=setup_budget(**kw))
        assert _labels(new_obs[0]) == want, kw
        assert rest == []


def kat_slice_subsets(bp):
    """:294-332 — sub-lists of slices and the num_slices budget distribution."""
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    obs = SparsePauliOp("X")
    slices = slice_by_barriers(qc)
    new_obs, rest, _ = bp([obs], slices[-1:])
    assert _labels(new_obs[0]) == {"X", "Z"} and len(rest) == 0
    new_obs, rest, _ = bp([obs], slices[-2:])
    assert _labels(new_obs[0]) == {"X", "Y", "Z"} and rest == []
    new_obs, rest, _ = bp([obs], slices[-2:], truncation_error_budget=setup_budget(max_error_total=1.0, num_slices=2))
    assert _labels(new_obs[0]) == {"X"} and rest == []
    new_obs, rest, _ = bp([obs], slices[-2:], truncation_error_budget=setup_budget(max_error_total=0.9, num_slices=2))
    assert _labels(new_obs[0]) == {"X", "Z"} and rest == []


def kat_l2_distribution(bp):
    """:333-387 — L2 thresholds 0.49/0.5 and 0.66/0.67."""
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    qc.barrier()
    qc.rz(THETA, 0)
    obs = SparsePauliOp("X")
    slices = slice_by_barriers(qc)
    for sub, total, n, want in (
        (slices[-1:], 0.49, 1, {"X", "Y"}),
        (slices[-1:], 0.5, 1, {"X"}),
        (slices[-2:], 0.66, 2, {"X", "Y"}),
        (slices[-2:], 0.67, 2, {"X"}),
    ):
        new_obs, _, _ = bp([obs], sub, truncation_error_budget=setup_budget(max_error_total=total, num_slices=n, p_norm=2))
        assert _labels(new_obs[0]) == want, (total, n)


def _thirty_slices():
    qc = QuantumCircuit(2)
    for _ in range(10):
        qc.rz(THETA, 0)
        qc.barrier()
        qc.cx(0, 1)
        qc.barrier()
        qc.ry(THETA, 0)
        qc.barrier()
    return slice_by_barriers(qc)


def kat_numeric_l1(bp, ordered=True):
    """:389-415 — 30 single-gate slices, L1 budget: coefficients, row order, accumulated error."""
    slices = _thirty_slices()
    assert len(slices) == 30
    budget = setup_budget(max_error_total=0.1, num_slices=30, p_norm=1)
    new_obs, rest, md = bp([SparsePauliOp("IX")], slices, truncation_error_budget=budget)
    expected = SparsePauliOp(
        ["XY", "XX", "IY", "IX", "XZ", "IZ"],
        [-0.72261328, 0.08235629, 0.17949574, 0.43425712, 0.20675452, 0.45436588],
    )
    if ordered:
        assert new_obs[0] == expected
    else:
        assert new_obs[0].sorted_by_key() == expected.sorted_by_key()
    assert len(rest) == 0
    assert md.accumulated_error(0) == pytest.approx(0.05558522939071649, abs=1e-9)


def kat_numeric_l2(bp, ordered=True):
    """:417-443 — same circuit, L2 budget."""
    slices = _thirty_slices()
    budget = setup_budget(max_error_total=0.1, num_slices=30, p_norm=2)
    new_obs, rest, md = bp([SparsePauliOp("IX")], slices, truncation_error_budget=budget)
    expected = SparsePauliOp(
        ["XY", "XX", "IY", "IX", "XZ", "IZ"],
        [-0.72261328, 0.08235629, 0.1705233, 0.44979783, 0.17567308, 0.45436588],
    )
    if ordered:
        assert new_obs[0] == expected
    else:
        assert new_obs[0].sorted_by_key() == expected.sorted_by_key()
    assert len(rest) == 0
    assert md.accumulated_error(0) == pytest.approx(0.07992961136903981, abs=1e-9)


def kat_max_paulis(bp):
    """:445-477 — max_paulis stops the loop and returns the un-absorbed slice."""
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    obs = SparsePauliOp("X")
    slices = slice_by_barriers(qc)
    new_obs, rest, _ = bp([obs], slices, operator_budget=OperatorBudget(max_paulis=3))
    assert _labels(new_obs[0]) == {"X", "Y", "Z"} and rest == []
    new_obs, rest, md = bp([obs], slices, operator_budget=OperatorBudget(max_paulis=2))
    assert _labels(new_obs[0]) == {"X", "Z"}
    assert len(rest) == 1 and rest[0] is slices[0]
    assert rest[0].data[0].name == "rz"
    assert len(md.backpropagation_history) == md.num_backpropagated_slices + 1
    with pytest.raises(ValueError):
        bp(obs, slices, operator_budget=OperatorBudget(max_paulis=0))
    with pytest.raises(ValueError):
        bp(SparsePauliOp(["X", "Z"]), slices, operator_budget=OperatorBudget(max_paulis=1))


def kat_max_qwc_groups(bp):
    """:478-509 — max_qwc_groups on one-qubit operators."""
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    obs = SparsePauliOp("X")
    slices = slice_by_barriers(qc)
    new_obs, rest, _ = bp([obs], slices, operator_budget=OperatorBudget(max_qwc_groups=3))
    assert _labels(new_obs[0]) == {"X", "Y", "Z"} and len(rest) == 0
    new_obs, rest, _ = bp([obs], slices, operator_budget=OperatorBudget(max_qwc_groups=2))
    assert _labels(new_obs[0]) == {"X", "Z"} and len(rest) == 1
    with pytest.raises(ValueError):
        bp(obs, slices, operator_budget=OperatorBudget(max_qwc_groups=0))
    with pytest.raises(ValueError):
        bp(SparsePauliOp(["X", "Z"]), slices, operator_budget=OperatorBudget(max_qwc_groups=1))


def kat_single_slice_p2(bp):
    """:510-555 — five gates in one slice (light-cone skip), p=2, three budgets."""
    qc = QuantumCircuit(2)
    qc.ry(np.pi / 6, 0)
    qc.rz(np.pi / 3, 0)
    qc.cx(0, 1)
    qc.ry(np.pi / 6, 1)
    qc.rz(np.pi / 6, 1)
    obs = SparsePauliOp("IX")
    for total, want in ((0.1, {"XX", "XZ", "XY"}), (0.26, {"XX", "XY"}), (0.501, {"XY"})):
        new_obs, _, _ = bp(obs, [qc], truncation_error_budget=setup_budget(max_error_total=total, p_norm=2))
        assert _labels(new_obs) == want, total


def kat_error_messages(bp):
    """:556-593 — width validation messages."""
    with pytest.raises(ValueError) as e:
        bp(SparsePauliOp("X"), [QuantumCircuit(1), QuantumCircuit(2)])
    assert e.value.args[0] == (
        "All slices must be defined on the same number of qubits. "
        "slices[0] contains 1 qubits, but slices[1] contains 2 qubits."
    )
    with pytest.raises(ValueError) as e:
        bp([SparsePauliOp("X"), SparsePauliOp("XX")], [QuantumCircuit(1), QuantumCircuit(1)])
    assert e.value.args[0] == (
        "Input observables must all act on the same number of qubits. "
        "observables[0] acts on 1 qubits, but observables[1] acts on 2 qubits."
    )
    with pytest.raises(ValueError) as e:
        bp([SparsePauliOp("XX"), SparsePauliOp("XX")], [QuantumCircuit(1), QuantumCircuit(1)])
    assert e.value.args[0] == (
        "The input observables must be defined on the same number of qubits as the circuit slices."
    )
    with pytest.raises(ValueError) as e:
        bp(SparsePauliOp("II"), [QuantumCircuit(2)])
    assert e.value.args[0] == "Input operator may not be the identity operator."


def kat_pauli_lindblad(bp, ordered=True):
    """:595-631 — a Pauli-Lindblad layer after / before a CX."""
    lerr = PauliLindbladError(["IX"], [1e-3])
    obs = [SparsePauliOp(["XY", "ZI", "II"], [3, 2, 1])]
    qc1 = QuantumCircuit(2)
    qc1.h(0)
    for ple_first, zx in ((True, 1.996004), (False, 2.0)):
        qc2 = QuantumCircuit(2)
        if ple_first:
            qc2.append(PauliLindbladErrorInstruction(lerr), [0, 1])
            qc2.cx(0, 1)
        else:
            qc2.cx(0, 1)
            qc2.append(PauliLindbladErrorInstruction(lerr), [0, 1])
        new_obs, _, _ = bp(obs, [qc1, qc2])
        target = SparsePauliOp(["IY", "II", "ZX"], [-2.994006, 1.0, zx])
        if ordered:
            assert new_obs[0] == target
        else:
            assert new_obs[0].sorted_by_key() == target.sorted_by_key()


def kat_multi_observable(bp):
    """:633-700, 911-920 — several observables, light cones per observable, num_unique regression."""
    qc = QuantumCircuit(1)
    qc.ry(THETA, 0)
    m = _ry_mat(THETA)
    new_obs, rest, _ = bp([SparsePauliOp("X"), SparsePauliOp("Z")], [qc])
    assert np.allclose(m.T @ np.array([[0, 1], [1, 0]]) @ m, new_obs[0].to_matrix())
    assert np.allclose(m.T @ np.diag([1.0, -1.0]) @ m, new_obs[1].to_matrix())
    assert rest == []
    qc = QuantumCircuit(2)
    qc.rz(THETA, 0)
    qc.ry(THETA, 1)
    new_obs, rest, _ = bp([SparsePauliOp("IX"), SparsePauliOp("ZI")], [qc])
    assert _labels(new_obs[0]) == {"IX", "IY"}
    assert _labels(new_obs[1]) == {"ZI", "XI"}
    qc = QuantumCircuit(2)
    qc.x(0)
    new_obs, _, md = bp([SparsePauliOp("ZI"), SparsePauliOp(["XX"])], [qc])
    assert len(new_obs[0]) == 1 and len(new_obs[1]) == 1
    assert md.backpropagation_history[0].num_unique_paulis[0] == 1
    assert md.backpropagation_history[0].num_unique_paulis[1] == 1
    # max_paulis counts distinct keys over all observables (:409-431)
    qc = QuantumCircuit(1)
    qc.rz(THETA, 0)
    qc.barrier()
    qc.ry(THETA, 0)
    slices = slice_by_barriers(qc)
    obs = [SparsePauliOp("X"), SparsePauliOp("Z")]
    new_obs, rest, _ = bp(obs, slices, operator_budget=OperatorBudget(max_paulis=3))
    assert _labels(new_obs[0]) == {"X", "Y", "Z"} and rest == []
    new_obs, rest, _ = bp(obs, slices, operator_budget=OperatorBudget(max_paulis=2))
    assert _labels(new_obs[0]) == {"X", "Z"} and _labels(new_obs[1]) == {"X", "Z"}
    assert len(rest) == 1


def kat_zero_slices_quirk(bp):
    """backpropagation.py:300 — ``slices[:-0]``: nothing committed returns an empty slice list."""
    qc = QuantumCircuit(1)
    qc.ry(THETA, 0)
    new_obs, rest, md = bp(SparsePauliOp("X"), [qc], operator_budget=OperatorBudget(max_paulis=1))
    assert md.num_backpropagated_slices == 0
    assert rest == []
    assert new_obs == SparsePauliOp("X")


def kat_depth_slicing(bp):
    """:924-933 inputs — rx, ry, rz, cx sliced by depth; full backpropagation without budgets."""
    qc = QuantumCircuit(2)
    qc.rx(0.1, 0)
    qc.ry(0.1, 0)
    qc.rz(0.1, 0)
    qc.cx(0, 1)
    slices = slice_by_depth(qc, 1)
    assert len(slices) == 4
    new_obs, rest, md = bp(SparsePauliOp("IX"), slices)
    assert len(rest) == 0 and md.num_backpropagated_slices == 4
    full = np.eye(4, dtype=complex)
    for s in slices:
        for ins in s.data:
            m = ins.matrix
            if len(ins.qubits) == 1:
                u = np.kron(np.eye(2), m) if ins.qubits[0] == 0 else np.kron(m, np.eye(2))
            else:
                u = m
            full = u @ full
    want = full.conj().T @ SparsePauliOp("IX").to_matrix() @ full
    assert np.allclose(want, new_obs.to_matrix(), atol=1e-12)


#: KATs every implementation must pass; the ordered ones take ``ordered=False`` on the engine
ORDER_FREE = [
    kat_single_slice,
    kat_depth2_clifford,
    kat_scattered_qargs,
    kat_multiple_slices,
    kat_max_error_total,
    kat_per_slice_list,
    kat_barrier_and_reset,
    kat_reset_simplify,
    kat_total_and_per_slice,
    kat_slice_subsets,
    kat_l2_distribution,
    kat_max_paulis,
    kat_max_qwc_groups,
    kat_single_slice_p2,
    kat_error_messages,
    kat_multi_observable,
    kat_zero_slices_quirk,
    kat_depth_slicing,
]
ORDERED = [kat_numeric_l1, kat_numeric_l2, kat_pauli_lindblad]
