"""Pin the CPU oracle against the reference's own known answers (no GPU needed).

The oracle (``oracle/obp_oracle.py``) is only trusted by the GPU parity tests because it passes
every known-answer test the reference's test-suite holds for this path that can be stated without
Qiskit objects: ``test/test_backpropagation.py``, ``test/utils/test_simplify.py``,
``test/utils/test_operations.py``.
"""

from __future__ import annotations

import numpy as np
import pytest

import kats
from oracle import obp_oracle as oracle
from qiskit_addon_obp_b200.ir import SparsePauliOp


@pytest.mark.parametrize("kat", kats.ORDER_FREE + [kats.kat_reset_no_simplify], ids=lambda f: f.__name__)
def test_oracle_kat(kat):
    kat(oracle.backpropagate)


@pytest.mark.parametrize("kat", kats.ORDERED, ids=lambda f: f.__name__)
def test_oracle_kat_row_order(kat):
    """The oracle reproduces the reference's row order (order-sensitive ``==``)."""
    kat(oracle.backpropagate, ordered=True)


def test_accumulated_error_bit_identical():
    """the restated arithmetic reproduces the reference's error totals to the last bit."""
    from qiskit_addon_obp_b200.utils.truncating import setup_budget

    slices = kats._thirty_slices()
    for p, want in ((1, 0.05558522939071649), (2, 0.07992961136903981)):
        _, _, md = oracle.backpropagate(
            [SparsePauliOp("IX")], slices, truncation_error_budget=setup_budget(max_error_total=0.1, num_slices=30, p_norm=p)
        )
        assert abs(md.accumulated_error(0) - want) < 5e-16


# ---- utils/simplify.py KATs (test/utils/test_simplify.py) --------------------------------------
def test_simplify_merge_and_counters():
    """test_simplify.py:24-40."""
    coeffs = [3 + 1j, -3 - 1j, 0, 4, -5, 2.2, -1.1j]
    labels = ["IXI", "IXI", "ZZZ", "III", "III", "XXX", "XXX"]
    out, md = oracle.simplify(SparsePauliOp(labels, coeffs))
    assert out == SparsePauliOp(["III", "XXX"], [-1, 2.2 - 1.1j])
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis, md.sum_trimmed_coeffs) == (3, 3, 2, 0)


def test_simplify_zero_operator():
    """test_simplify.py:42-60."""
    coeffs = [3 + 1j, -3 - 1j, 0, 4, -5, 2.2, -1.1j]
    labels = ["IXI", "IXI", "ZZZ", "III", "III", "XXX", "XXX"]
    op = SparsePauliOp(labels, coeffs)
    out, md = oracle.simplify(op - op)
    np.testing.assert_array_equal(out.coeffs, [0])
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis, md.sum_trimmed_coeffs) == (3, 9, 5, 0)


def test_simplify_no_op_and_tolerances():
    """test_simplify.py:81-136."""
    op = SparsePauliOp(["IXI", "ZZZ"], [3, 2])
    out, md = oracle.simplify(op)
    assert out == op
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis) == (2, 0, 0)
    op = SparsePauliOp(["III", "ZII", "IZI", "IIZ"], [1.0, 0.1, 0.01, 0.001])
    out, md = oracle.simplify(op, atol=1e-2, rtol=1e-2)
    assert out == SparsePauliOp(["III", "ZII"], [1.0, 0.1])
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis) == (2, 0, 2)
    op = SparsePauliOp(["IXI", "IXI", "ZZZ", "XXX", "YYY"], [1e-8, 1e-8j, 1e-8 + 1e-8j, 0.1, 0.2])
    out, _ = oracle.simplify(op, atol=None, rtol=None)
    assert out == SparsePauliOp(["ZZZ", "XXX", "YYY"], [1e-8 + 1e-8j, 0.1, 0.2])
    out, _ = oracle.simplify(op, atol=0.01, rtol=1e-10)
    assert out == SparsePauliOp(["XXX", "YYY"], [0.1, 0.2])


# ---- utils/operations.py KATs (test/utils/test_operations.py) ----------------------------------
def test_apply_op_to_cx():
    """test_operations.py:38-49, 103-136 — CX conjugation incl. scattered qubits."""
    from qiskit_addon_obp_b200.ir import Instruction

    cx = oracle.pauli_decompose(Instruction("cx", (0, 1)).matrix)
    assert len(cx) == 4
    out = oracle.apply_op_to(SparsePauliOp("IX"), cx, [0, 1])
    out, _ = oracle.simplify(out)
    assert np.array_equal(out.to_matrix(), np.fliplr(np.identity(4)))
    # obs XX on qubits [0,1] of 3, CX on [1,2]
    obs = SparsePauliOp("IXX")
    out, _ = oracle.simplify(oracle.apply_op_to(obs, cx, [1, 2]))
    cxm = np.kron(np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]]), np.eye(2))
    assert np.allclose(out.to_matrix(), cxm.T @ obs.to_matrix() @ cxm)


def test_apply_reset_to():
    """test_operations.py:211-239."""
    op = SparsePauliOp(["XYZ", "YZX", "ZXY"], [1.0, 2.0, 3.0])
    for q, lbl, c in ((0, "XYI", 1.0), (1, "YIX", 2.0), (2, "IXY", 3.0)):
        out, _ = oracle.simplify(oracle.apply_reset_to(op, q))
        assert out == SparsePauliOp([lbl], [c])


def test_apply_ple_to():
    """test_operations.py:241-296 — global-width restatement of the seven cases."""
    cases = [
        (["Z"], [1.0], ["X"], [1e-3], [0], ["Z"], [0.998002]),
        (["ZX"], [1.0], ["XX"], [1e-3], [0, 1], ["ZX"], [0.998002]),
        (["ZI"], [1.0], ["X"], [1e-3], [0], ["ZI"], [1.0]),
        (["XZI"], [1.0], ["XX"], [1e-3], [0, 1], ["XZI"], [0.998002]),
        (["XZI"], [1.0], ["XX", "YY"], [1e-3, 1e-2], [0, 1], ["XZI"], [0.97824024]),
        (["XZI", "YYI"], [1.0, 1.0], ["YY"], [1e-2], [0, 1], ["XZI", "YYI"], [0.98019867, 1.0]),
        (["XZI", "YYI"], [1.0, 1.0], ["XX", "YY"], [1e-3, 1e-2], [0, 1], ["XZI", "YYI"], [0.97824024, 0.998002]),
    ]
    for labels, coeffs, gens, rates, qargs, want_l, want_c in cases:
        out = oracle.apply_ple_to(SparsePauliOp(labels, coeffs), gens, rates, qargs)
        assert out == SparsePauliOp(want_l, want_c), (labels, gens)


def test_k2_expansion_matches_dense_conjugation():
    """The restated compose/simplify equals dense U^dag O U on a small register."""
    from qiskit_addon_obp_b200.ir import Instruction

    rng = np.random.default_rng(7)
    n, nq = 40, 5
    x = rng.random((n, nq)) < 0.4
    z = rng.random((n, nq)) < 0.4
    op = SparsePauliOp((x, z), rng.normal(size=n) + 0j)
    for name, qs, th in (("rzz", (1, 3), 0.37), ("rx", (2,), 1.1), ("ryy", (0, 4), -0.6), ("t", (3,), None)):
        ins = Instruction(name, qs, () if th is None else (th,))
        gate = oracle.pauli_decompose(ins.matrix)
        out, _ = oracle.simplify(oracle.apply_op_to(op, gate, list(qs)))
        # dense check on the small system
        from functools import reduce

        if len(qs) == 1:
            mats = [np.eye(2)] * nq
            mats[qs[0]] = ins.matrix
            u = reduce(np.kron, mats[::-1])
        else:
            u = _embed_two(ins.matrix, qs, nq)
        want = u.conj().T @ op.to_matrix() @ u
        assert np.allclose(out.to_matrix(), want, atol=1e-12), name


def _embed_two(m, qs, nq):
    """Embed a little-endian 2-qubit matrix on qubits ``qs`` of an ``nq``-qubit register."""
    dim = 2**nq
    u = np.zeros((dim, dim), dtype=complex)
    a, b = qs
    for col in range(dim):
        ca, cb = (col >> a) & 1, (col >> b) & 1
        rest = col & ~((1 << a) | (1 << b))
        for ra in range(2):
            for rb in range(2):
                row = rest | (ra << a) | (rb << b)
                u[row, col] = m[ra + 2 * rb, ca + 2 * cb]
    return u


# ---- the oracle's own conventions against the product's host code (two independent statements) -------
def test_gate_tables_agree_and_match_closed_forms():
    """Every named gate of the host IR (``ir.py``: literal little-endian matrices) against the oracle's
    table (Pauli closed forms), plus textbook identities that fix the conventions themselves."""
    from qiskit_addon_obp_b200 import ir

    rng = np.random.default_rng(7)
    names = [(n, ()) for n in ir._FIXED_1Q + ir._FIXED_2Q]
    names += [(n, (float(rng.uniform(-3, 3)),)) for n in list(ir._ROT) + ["p", "cp", "crz"] for _ in range(3)]
    for name, params in names:
        nq = 1 if (name in ir._FIXED_1Q or name in ("rx", "ry", "rz", "p")) else 2
        m_ir = ir.Instruction(name, tuple(range(nq)), params).matrix
        m_or = oracle.gate_matrix(name, params)
        np.testing.assert_allclose(m_ir, m_or, atol=1e-15, err_msg=name)
        np.testing.assert_allclose(m_or @ m_or.conj().T, np.eye(2**nq), atol=1e-15, err_msg=f"{name} unitary")
    g = oracle.gate_matrix
    # index = q0 + 2*q1; cx: control = first argument flips the second: |q0=1,q1=0> (index 1) -> index 3
    assert g("cx")[3, 1] == 1 and g("cx")[1, 3] == 1 and g("cx")[0, 0] == 1 and g("cx")[2, 2] == 1
    assert g("cy")[3, 1] == 1j and g("cy")[1, 3] == -1j
    np.testing.assert_allclose(g("cz"), np.diag([1, 1, 1, -1]))
    np.testing.assert_allclose(g("swap")[[1, 2]][:, [2, 1]], np.eye(2))
    np.testing.assert_allclose(g("iswap"), np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]]))
    np.testing.assert_allclose(  # Qiskit's documented ECR matrix
        g("ecr"), np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]]) / np.sqrt(2)
    )
    np.testing.assert_allclose(g("sx") @ g("sx"), g("x"), atol=1e-15)
    np.testing.assert_allclose(g("s") @ g("s"), g("z"), atol=1e-15)
    np.testing.assert_allclose(g("t") @ g("t"), g("s"), atol=1e-15)
    np.testing.assert_allclose(g("h") @ g("z") @ g("h"), g("x"), atol=1e-15)
    np.testing.assert_allclose(g("rz", (0.3,)), np.diag([np.exp(-0.15j), np.exp(0.15j)]), atol=1e-15)
    np.testing.assert_allclose(g("crz", (0.3,)), np.diag([1, np.exp(-0.15j), 1, np.exp(0.15j)]), atol=1e-15)
    np.testing.assert_allclose(g("cp", (0.3,)), np.diag([1, 1, 1, np.exp(0.3j)]), atol=1e-15)
    # rzx(t) = exp(-i t/2 X (x) Z) with Z on the first argument (index bit 0), X on the second
    zx = np.kron(np.array([[0, 1], [1, 0]]), np.diag([1, -1]))
    np.testing.assert_allclose(g("rzx", (0.7,)), np.cos(0.35) * np.eye(4) - 1j * np.sin(0.35) * zx, atol=1e-15)


def test_gate_order_agrees_with_host_ir():
    from qiskit_addon_obp_b200.ir import QuantumCircuit, topological_op_order

    rng = np.random.default_rng(11)
    for _ in range(50):
        n = int(rng.integers(2, 9))
        qc = QuantumCircuit(n)
        for _ in range(int(rng.integers(1, 25))):
            if rng.random() < 0.5:
                qc.rx(0.1, int(rng.integers(n)))
            else:
                a, b = rng.choice(n, size=2, replace=False)
                qc.cx(int(a), int(b))
        got = oracle.gate_order(qc)
        want = topological_op_order(qc)
        assert [id(x) for x in got] == [id(x) for x in want]


def test_oracle_budget_replay_on_the_reference_fixture():
    """test/utils/test_metadata.py:183-265 on the oracle's own replay (the product's is pinned in test_host_logic)."""
    import json
    import os

    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    md = oracle.OBPMetadata.from_json(os.path.join(golden, "metadata_fixture.json"))
    want = json.load(open(os.path.join(golden, "metadata_expected.json")))
    for obs in (0, 1):
        n = len(md.backpropagation_history)
        np.testing.assert_allclose([md.accumulated_error(obs, i) for i in range(n)], want["accumulated_error"][obs], rtol=0, atol=5e-11)
        np.testing.assert_allclose([md.left_over_error_budget(obs, i) for i in range(n)], want["left_over_error_budget"][obs], rtol=0, atol=5e-11)


def test_oracle_imports_only_containers_from_the_product():
    import os
    import re

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "obp_oracle.py")).read()
    imports = re.findall(r"^from qiskit_addon_obp_b200[.\w]* import (.+)$", src, flags=re.M)
    assert imports == ["QuantumCircuit, SparsePauliOp"], imports
