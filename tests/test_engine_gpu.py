"""GPU parity tests: the CUDA engine (through the C-ABI) against the CPU oracle.

Bars (BASELINE.json north star): surviving Pauli key set bit-exact, coefficients within
1e-10 rel / 1e-12 abs, error totals within 1e-9, integer counters exact.
"""

from __future__ import annotations

import numpy as np
import pytest

import kats
from conftest import assert_same_metadata, assert_same_operator
from oracle import obp_oracle as oracle
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200.ir import (
    PauliLindbladError,
    PauliLindbladErrorInstruction,
    QuantumCircuit,
    SparsePauliOp,
)
from qiskit_addon_obp_b200.utils.simplify import OperatorBudget
from qiskit_addon_obp_b200.utils.truncating import TruncationErrorBudget, setup_budget

pytestmark = pytest.mark.gpu


# ---- the reference's own KATs through the engine ------------------------------------------------
@pytest.mark.parametrize("kat", kats.ORDER_FREE + [kats.kat_reset_no_simplify], ids=lambda f: f.__name__)
def test_reference_kat(kat):
    kat(backpropagate)


def test_three_and_four_qubit_unitaries():
    """Gates with 64 and 256 Pauli components run through the raw-row expansion kernel like cp / crz do
    (the reference decomposes any `Operator`, backpropagation.py:213)."""
    rng = np.random.default_rng(33)

    def haar(dim):
        a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        q, r = np.linalg.qr(a)
        return q * (np.diag(r) / np.abs(np.diag(r)))

    nq = 6
    qc1 = QuantumCircuit(nq)
    qc1.unitary(haar(8), [4, 0, 2])  # 64 components, qubits out of order
    qc1.rx(0.3, 1)
    qc2 = QuantumCircuit(nq)
    qc2.unitary(np.eye(8)[[0, 1, 2, 7, 4, 5, 6, 3]].astype(complex), [0, 1, 2], label="ccx")  # Toffoli: 8 components
    qc2.cx(2, 3)
    qc3 = QuantumCircuit(nq)
    qc3.unitary(haar(16), [5, 1, 3, 0])  # 256 components
    obs = SparsePauliOp(["ZIIIII", "IIXYII"], [1.0, 0.5])
    _compare(obs, [qc1, qc2])
    _compare(SparsePauliOp(["IIIIZI"], [1.0]), [qc3, qc2])


def test_simplify_false_matches_oracle():
    """OperatorBudget(simplify=False): raw rows multiply by k*k per gate, nothing merges (reference :220)."""
    qc1 = QuantumCircuit(3)
    qc1.rx(0.3, 0)
    qc1.cx(0, 1)
    qc2 = QuantumCircuit(3)
    qc2.h(1)
    qc2.reset(2)
    qc2.rzz(0.2, 0, 1)
    obs = SparsePauliOp(["ZIX", "XYZ", "ZIX"], [1.0, 0.5, 0.25])  # a duplicate row stays a duplicate
    budget = OperatorBudget(simplify=False)
    got, rest, md = backpropagate(obs, [qc1, qc2], operator_budget=budget)
    want, _, md_w = oracle.backpropagate(obs, [qc1, qc2], operator_budget=budget)
    assert rest == [] and len(got) == len(want)
    assert [s.num_paulis for s in md.backpropagation_history] == [s.num_paulis for s in md_w.backpropagation_history]
    assert [s.raw_num_paulis for s in md.backpropagation_history] == [s.raw_num_paulis for s in md_w.backpropagation_history]
    assert all(s.num_unique_paulis is None and s.sum_trimmed_coeffs is None for s in md.backpropagation_history)
    a, b = got.as_dict(), want.as_dict()  # sums per label
    assert set(a) == set(b)
    assert max(abs(a[k] - b[k]) for k in a) < 1e-12
    got, _, _ = backpropagate(obs, [qc1, qc2], operator_budget=budget, truncation_error_budget=setup_budget(max_error_per_slice=0.05))
    want, _, _ = oracle.backpropagate(obs, [qc1, qc2], operator_budget=budget, truncation_error_budget=setup_budget(max_error_per_slice=0.05))
    assert len(got) == len(want)


@pytest.mark.parametrize("kat", kats.ORDERED, ids=lambda f: f.__name__)
def test_reference_kat_numeric(kat):
    """Numeric KATs; row order is not part of the parity contract."""
    kat(backpropagate, ordered=False)


# ---- random circuits ------------------------------------------------------------------------------
_CLIFF1 = ["h", "s", "sdg", "x", "y", "z", "sx", "sxdg"]
_CLIFF2 = ["cx", "cz", "cy", "swap", "iswap", "ecr"]
_ROT1 = ["rx", "ry", "rz", "p"]
_ROT2 = ["rxx", "ryy", "rzz", "rzx"]
_FIX_ROT = ["t", "tdg"]


def random_slices(nq, n_slices, rng, *, p_rot=0.5, gates_per_slice=None, angles=None):
    slices = []
    for _ in range(n_slices):
        qc = QuantumCircuit(nq)
        perm = rng.permutation(nq)
        k = 0
        budget = gates_per_slice or nq
        while k < nq and budget > 0:
            budget -= 1
            two = k + 1 < nq and rng.random() < 0.4
            th = float(rng.uniform(-np.pi, np.pi)) if angles is None else float(rng.choice(angles))
            if two:
                a, b = int(perm[k]), int(perm[k + 1])
                k += 2
                if rng.random() < p_rot:
                    getattr(qc, str(rng.choice(_ROT2)))(th, a, b)
                else:
                    getattr(qc, str(rng.choice(_CLIFF2)))(a, b)
            else:
                a = int(perm[k])
                k += 1
                r = rng.random()
                if r < p_rot * 0.8:
                    getattr(qc, str(rng.choice(_ROT1)))(th, a)
                elif r < p_rot:
                    getattr(qc, str(rng.choice(_FIX_ROT)))(a)
                else:
                    getattr(qc, str(rng.choice(_CLIFF1)))(a)
        slices.append(qc)
    return slices


def random_observable(nq, n_terms, rng, weight=2):
    x = np.zeros((n_terms, nq), dtype=bool)
    z = np.zeros((n_terms, nq), dtype=bool)
    for i in range(n_terms):
        for q in rng.choice(nq, size=min(weight, nq), replace=False):
            c = rng.integers(1, 4)
            x[i, q] = bool(c & 1)
            z[i, q] = bool(c & 2)
    op = SparsePauliOp((x, z), rng.normal(size=n_terms) + 0j)
    out, _ = oracle.simplify(op)
    return out


def _compare(obs, slices, **kw):
    got, rest_g, md_g = backpropagate(obs, slices, **kw)
    want, rest_w, md_w = oracle.backpropagate(obs, slices, **kw)
    assert len(rest_g) == len(rest_w)
    single = isinstance(obs, SparsePauliOp)
    gl, wl = ([got], [want]) if single else (got, want)
    tb = kw.get("truncation_error_budget")
    for g, w in zip(gl, wl):
        assert_same_operator(g, w)
    assert_same_metadata(md_g, md_w)
    del tb
    return gl, md_g


@pytest.mark.parametrize("nq", [3, 6, 64, 70, 130, 200])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_circuit_no_truncation(nq, seed):
    rng = np.random.default_rng(1000 * nq + seed)
    slices = random_slices(nq, 6, rng, gates_per_slice=8)
    obs = random_observable(nq, 3, rng)
    _compare(obs, slices)


@pytest.mark.parametrize("nq,p", [(5, 1), (5, 2), (66, 1), (129, 2)])
def test_random_circuit_truncation(nq, p):
    rng = np.random.default_rng(77 + nq + p)
    slices = random_slices(nq, 8, rng, gates_per_slice=6, angles=[0.1, 0.3, -0.2, 0.7])
    obs = random_observable(nq, 2, rng)
    _compare(obs, slices, truncation_error_budget=TruncationErrorBudget([0.02, 0.0, 0.05], 0.2, p, 1e-8))
    _compare(obs, slices, truncation_error_budget=setup_budget(max_error_total=0.05, num_slices=8, p_norm=p))


@pytest.mark.parametrize("nq", [5, 40, 70, 127])
@pytest.mark.parametrize("mode", ["slice-in-smem", "tiny-partitions", "cut-into-pieces", "gate-by-gate"])
def test_rotation_layers(nq, mode, monkeypatch):
    """Slices made of rotations only (RX / RY / RZ / RZZ / RXX layers): the slice-in-shared-memory path
    (k_slice_smem: the table partitioned so that whole cosets live in one CTA, all gates between __syncthreads)
    and the gate-by-gate passes against the oracle.  Small angles and a loose atol put coefficients inside the
    trim band of the intermediate gates."""
    monkeypatch.setenv("OBP_B200_NO_SEGMENTS", "1" if mode == "gate-by-gate" else "0")
    # partitions of 24 rows: runs outgrow them, are retried with four times as many partitions and finally fall
    # back to the per-gate kernels (the input table must come through untouched)
    monkeypatch.setenv("OBP_B200_SEG_CAP_ROWS", {"tiny-partitions": "24", "cut-into-pieces": "48"}.get(mode, "0"))
    monkeypatch.setenv("OBP_B200_SEG_MIN_GATES", "2")
    # cut-into-pieces: a failed run is repeated in shorter pieces whatever the table size (default: only from 2^17 rows)
    monkeypatch.setenv("OBP_B200_SEG_PIECES_MIN_ROWS", "1" if mode == "cut-into-pieces" else str(1 << 17))
    rng = np.random.default_rng(4000 + nq)
    slices = []
    for k in range(6):
        qc = QuantumCircuit(nq)
        kind = k % 3
        if kind == 0:
            for q in range(nq):
                getattr(qc, ("rx", "ry", "rz")[int(rng.integers(3))])(float(rng.choice([0.1, 0.3, -0.7, 1e-3])), q)
        elif kind == 1:
            for q in range(k % 2, nq - 1, 2):
                getattr(qc, ("rzz", "rxx", "ryy")[int(rng.integers(3))])(float(rng.choice([0.2, -0.4, 2e-3])), q, q + 1)
        else:
            for q in range(0, nq - 1, 2):
                qc.cx(q, q + 1)
            for q in range(nq):
                qc.rx(0.25, q)
        slices.append(qc)
    obs = random_observable(nq, 3, rng, weight=min(nq, 4))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=20000))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=20000, atol=1e-4))
    _compare(obs, slices, truncation_error_budget=TruncationErrorBudget([0.01], 0.1, 2, 1e-8))
    if mode != "gate-by-gate":
        from qiskit_addon_obp_b200 import backpropagation as bpmod

        runs, cut, fell_back = bpmod.last_run.segments
        assert runs > 0, "the slice-in-shared-memory path did not run"
        if mode == "cut-into-pieces" and nq >= 40:
            assert cut > 0, "no run was cut into pieces: the case does not test what it says"


@pytest.mark.parametrize("nq", [6, 40, 70, 127])
@pytest.mark.parametrize("mode", ["slice-in-smem", "tiny-partitions", "cut-into-pieces", "rotations-only-runs"])
def test_mixed_runs(nq, mode, monkeypatch):
    """Slices that mix rotations with Clifford-class gates (one- and two-qubit, with sign flips): the whole slice is one
    run in shared memory — the Clifford gates are applied to the rows of each partition, the partition map vanishes on
    the generators pulled back through them — against the oracle, counters included; slices that end on a Clifford gate
    (its coset count goes through the program kernel) and on a rotation both occur."""
    monkeypatch.setenv("OBP_B200_SEG_MIN_GATES", "2")
    monkeypatch.setenv("OBP_B200_SEG_MIXED", "0" if mode == "rotations-only-runs" else "1")
    monkeypatch.setenv("OBP_B200_SEG_CAP_ROWS", {"tiny-partitions": "24", "cut-into-pieces": "48"}.get(mode, "0"))
    monkeypatch.setenv("OBP_B200_SEG_PIECES_MIN_ROWS", "1" if mode == "cut-into-pieces" else str(1 << 17))
    rng = np.random.default_rng(9000 + nq)
    slices = random_slices(nq, 6, rng, p_rot=0.45, angles=[0.1, 0.3, -0.7, 1e-3, 0.25])
    obs = random_observable(nq, 3, rng, weight=min(nq, 4))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=30000))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=30000, atol=1e-4))
    _compare(obs, slices, truncation_error_budget=TruncationErrorBudget([0.01], 0.1, 2, 1e-8))
    if mode != "rotations-only-runs" and nq >= 40:
        from qiskit_addon_obp_b200 import backpropagation as bpmod

        assert bpmod.last_run.segments[0] > 0, "no slice ran in shared memory"


def test_mixed_runs_config5(monkeypatch):
    """BASELINE config 5 (Clifford + RZ brickwork) at oracle size: every slice is a mixed run."""
    from qiskit_addon_obp_b200 import backpropagation as bpmod
    from qiskit_addon_obp_b200 import workloads as wl

    monkeypatch.setenv("OBP_B200_SEG_MIXED", "1")
    for seed in (0, 1):
        w = wl.config5_magic_sweep(theta=np.pi / 8, depth=14, max_paulis=4_000, seed=seed)
        _compare(w.observables, w.slices, **w.kwargs())
        assert bpmod.last_run.segments[0] > 0


def test_clifford_only_counters():
    """Clifford gates: keys move, nothing merges; counters come from the coset count kernel."""
    rng = np.random.default_rng(5)
    nq = 9
    slices = random_slices(nq, 10, rng, p_rot=0.0)
    obs = random_observable(nq, 40, rng, weight=4)
    _compare(obs, slices)


def test_multi_observable_and_max_paulis():
    rng = np.random.default_rng(11)
    nq = 8
    slices = random_slices(nq, 10, rng, gates_per_slice=6)
    obs = [random_observable(nq, 2, rng), random_observable(nq, 3, rng), SparsePauliOp("IIIIIIIZ")]
    _compare(obs, slices)
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=60))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=60, max_qwc_groups=1000))


def test_small_coefficients_trim_band():
    """Terms driven below atol by repeated small-angle branching are trimmed exactly like the reference."""
    nq = 4
    slices = []
    for k in range(12):
        qc = QuantumCircuit(nq)
        qc.rx(1e-3, k % nq)
        qc.rzz(2e-3, k % nq, (k + 1) % nq)
        slices.append(qc)
    obs = SparsePauliOp(["ZIII", "IIXY"], [1.0, 3e-6])
    _compare(obs, slices)
    _compare(obs, slices, operator_budget=OperatorBudget(atol=1e-5))


def test_reset_and_noise_layers():
    rng = np.random.default_rng(3)
    nq = 6
    slices = random_slices(nq, 4, rng, gates_per_slice=5)
    ple = PauliLindbladError(["IXX", "ZYI", "IIZ"], [1e-2, 3e-3, 5e-2])
    noisy = QuantumCircuit(nq)
    noisy.append(PauliLindbladErrorInstruction(ple), [1, 3, 4])
    noisy.cx(1, 3)
    rst = QuantumCircuit(nq)
    rst.reset(2)
    rst.h(0)
    obs = random_observable(nq, 6, rng, weight=3)
    _compare(obs, slices + [noisy] + slices[:2] + [rst] + slices[2:])


def test_generic_gates():
    """cp / crz / arbitrary unitaries: the raw k*k expansion + merge path, incl. its simplify counters."""
    rng = np.random.default_rng(21)
    nq = 7

    def haar(dim):
        a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
        q, r = np.linalg.qr(a)
        return q * (np.diag(r) / np.abs(np.diag(r)))

    slices = []
    for k in range(6):
        qc = QuantumCircuit(nq)
        qc.cp(0.3 + 0.1 * k, k % nq, (k + 2) % nq)
        qc.rx(0.4, (k + 1) % nq)
        qc.crz(-0.7, (k + 3) % nq, (k + 4) % nq)
        qc.unitary(haar(2), [(k + 5) % nq])
        if k % 2:
            qc.unitary(haar(4), [(k + 6) % nq, (k + 1) % nq])  # 16 Pauli components
        qc.h(k % nq)
        slices.append(qc)
    obs = random_observable(nq, 3, rng)
    _compare(obs, slices)
    _compare(obs, slices[:3] + [slices[4][:0] if False else slices[5]], truncation_error_budget=setup_budget(max_error_per_slice=0.01))
    # a slice ENDING on a generic gate reports the reference's counters of that gate
    last = QuantumCircuit(nq)
    last.h(0)
    last.cp(0.9, 0, 1)
    rev = QuantumCircuit(nq)
    rev.cp(0.9, 0, 1)
    rev.h(0)
    _compare(obs, slices[:2] + [last, rev])
    # 130 qubits: generic gates across word boundaries
    wide = []
    for k in range(3):
        qc = QuantumCircuit(130)
        qc.cp(0.5, 63, 64)
        qc.crz(0.2, 127, 128)
        qc.ry(0.3, 64)
        wide.append(qc)
    _compare(SparsePauliOp.from_sparse_list([("ZX", [63, 128], 1.0), ("Y", [64], 0.5)], 130), wide)


# ---- the named configurations at oracle-sized scale ----------------------------------------------
def test_config1_full():
    from qiskit_addon_obp_b200.workloads import config1_xy_chain

    wl = config1_xy_chain()
    gl, md = _compare(wl.observables, wl.slices, **wl.kwargs())
    assert md.num_backpropagated_slices == 18 and len(gl[0]) == 71


def test_config2_small():
    from qiskit_addon_obp_b200.workloads import config2_kicked_ising

    wl = config2_kicked_ising(theta_h=0.3, steps=2, max_paulis=3000)
    _compare(wl.observables, wl.slices, **wl.kwargs())


def test_config3_small():
    from qiskit_addon_obp_b200.workloads import config3_square_lattice

    wl = config3_square_lattice(side=10, steps=1, max_paulis=20000)
    _compare(wl.observables, wl.slices, **wl.kwargs())


def test_config4_small():
    from qiskit_addon_obp_b200.workloads import config4_near_clifford

    wl = config4_near_clifford(num_qubits=1000, depth=12, t_prob=0.01, max_paulis=4000)
    _compare(wl.observables, wl.slices, **wl.kwargs())
    wl = config4_near_clifford(num_qubits=1000, depth=12, t_prob=0.01, max_paulis=4000, per_slice=1e-4)
    _compare(wl.observables, wl.slices, **wl.kwargs())


def test_config5_small():
    from qiskit_addon_obp_b200.workloads import config5_magic_sweep

    wl = config5_magic_sweep(theta=np.pi / 16, depth=5, max_paulis=3000)
    _compare(wl.observables, wl.slices, **wl.kwargs())


# ---- utils-level entry points ----------------------------------------------------------------------
def test_simplify_device():
    from qiskit_addon_obp_b200.utils.simplify import simplify

    coeffs = [3 + 1j, -3 - 1j, 0, 4, -5, 2.2, -1.1j]
    labels = ["IXI", "IXI", "ZZZ", "III", "III", "XXX", "XXX"]
    op = SparsePauliOp(labels, coeffs)
    out, md = simplify(op)
    assert out.sorted_by_key() == SparsePauliOp(["III", "XXX"], [-1, 2.2 - 1.1j]).sorted_by_key()
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis, md.sum_trimmed_coeffs) == (3, 3, 2, 0)
    out, md = simplify(op - op)
    np.testing.assert_array_equal(out.coeffs, [0])
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis) == (3, 9, 5)
    op = SparsePauliOp(["IXI", "IXI", "ZZZ", "XXX", "YYY"], [1e-8, 1e-8j, 1e-8 + 1e-8j, 0.1, 0.2])
    out, _ = simplify(op)
    assert out.sorted_by_key() == SparsePauliOp(["ZZZ", "XXX", "YYY"], [1e-8 + 1e-8j, 0.1, 0.2]).sorted_by_key()
    out, _ = simplify(op, atol=0.01, rtol=1e-10)
    assert out.sorted_by_key() == SparsePauliOp(["XXX", "YYY"], [0.1, 0.2]).sorted_by_key()
    # random duplicates, wide keys
    rng = np.random.default_rng(0)
    nq, n = 150, 5000
    x = rng.random((n, nq)) < 0.02
    z = rng.random((n, nq)) < 0.02
    idx = rng.integers(0, n // 3, size=n)
    big = SparsePauliOp((x[idx], z[idx]), rng.normal(size=n) * (rng.random(n) < 0.9) + 1j * rng.normal(size=n))
    out, md = simplify(big)
    want, md_w = oracle.simplify(big)
    assert_same_operator(out, want)
    assert (md.num_unique_paulis, md.num_duplicate_paulis, md.num_trimmed_paulis) == (
        md_w.num_unique_paulis,
        md_w.num_duplicate_paulis,
        md_w.num_trimmed_paulis,
    )


@pytest.mark.parametrize("p", [1, 2, 3])
def test_truncate_binary_search_device(p):
    from qiskit_addon_obp_b200.utils.truncating import truncate_binary_search

    rng = np.random.default_rng(p)
    nq, n = 40, 20000
    x = rng.random((n, nq)) < 0.2
    z = rng.random((n, nq)) < 0.2
    op, _ = oracle.simplify(SparsePauliOp((x, z), rng.normal(size=n) * 10.0 ** rng.uniform(-6, 0, size=n)))
    for budget in (0.0, 1e-9, 1e-3, 0.5, 1e6):
        got, err = truncate_binary_search(op, budget, p_norm=p, tol=1e-8)
        want, err_w = oracle.truncate_binary_search(op, budget, p_norm=p, tol=1e-8)
        assert abs(err - err_w) <= 1e-9 * max(1.0, abs(err_w))
        assert_same_operator(got, want, thresholds=())


def test_operations_device():
    from qiskit_addon_obp_b200.ir import Instruction
    from qiskit_addon_obp_b200.utils.operations import apply_op_to, apply_ple_to, apply_reset_to

    new_op, qargs = apply_op_to(SparsePauliOp("IX"), [0, 1], Instruction("cx", (0, 1)), [0, 1], apply_as_transform=True)
    assert np.array_equal(new_op.to_matrix(), np.fliplr(np.identity(4))) and qargs == [0, 1]
    new_op, qargs = apply_op_to(SparsePauliOp("XX"), [2, 4], Instruction("cx", (0, 1)), [6, 8], apply_as_transform=True)
    assert qargs == [2, 4, 6, 8] and new_op == SparsePauliOp("IIXX")
    with pytest.raises(ValueError) as e:
        apply_op_to(SparsePauliOp("XYZ"), [1, 2], Instruction("x", (0,)), [1], apply_as_transform=True)
    assert e.value.args[0] == "The number of qubits in the operator (3) does not match the number of qargs (2)."
    op = SparsePauliOp(["XYZ", "YZX", "ZXY"], [1.0, 2.0, 3.0])
    for q, lbl, c in ((0, "XYI", 1.0), (1, "YIX", 2.0), (2, "IXY", 3.0)):
        assert apply_reset_to(op, q) == SparsePauliOp([lbl], [c])
    new_op, qargs = apply_ple_to(SparsePauliOp(["XZ", "YY"], [1.0, 1.0]), [1, 2], PauliLindbladError(["XX", "YY"], [1e-3, 1e-2]), [0, 1])
    assert qargs == [0, 1, 2]
    assert new_op.sorted_by_key() == SparsePauliOp(["XZI", "YYI"], [0.97824024, 0.998002]).sorted_by_key()


# ---- size-independent properties at BASELINE sizes -------------------------------------------------
def _inverse_slices(slices):
    inv = []
    for s in slices[::-1]:
        qc = QuantumCircuit(s.num_qubits)
        for ins in s.data[::-1]:
            qc.unitary(ins.matrix.conj().T, ins.qubits, label=ins.name + "_dg")
        inv.append(qc)
    return inv


def test_round_trip_config2_1e6():
    """U then U^dag returns the observable: Z_62 with coefficient 1 (127 qubits, ~5e5 terms mid-way).

    Backpropagating ``inverse + forward`` slices first absorbs ``forward`` (growing the operator to
    the ~1e6-term regime of BASELINE config 2) and then ``inverse``, which must merge everything
    back.  Trimming at atol=1e-8 loses at most (number of trimmed terms) * 1e-8 of weight.
    """
    from qiskit_addon_obp_b200.workloads import config2_kicked_ising

    wl = config2_kicked_ising(theta_h=0.3, steps=6, max_paulis=None)
    fwd = wl.slices
    from qiskit_addon_obp_b200 import backpropagation as bpmod

    out, rest, md = backpropagate(wl.observables, _inverse_slices(fwd) + fwd)
    assert rest == []
    peak = max(s.num_paulis[0] for s in md.backpropagation_history)
    assert peak > 100_000, peak
    assert bpmod.last_run.updates > 100_000_000
    d = out.as_dict()
    key = wl.observables.paulis.to_labels()[0]
    assert abs(d[key] - 1.0) < 1e-5
    rest_w = sum(abs(c) for k, c in d.items() if k != key)
    assert rest_w < 1e-2, rest_w


def test_norm_conservation_config5():
    """Unitary conjugation preserves sum |c|^2 (no truncation; trimmed weight is below atol per term)."""
    from qiskit_addon_obp_b200.workloads import config5_magic_sweep

    wl = config5_magic_sweep(theta=np.pi / 8, depth=12, max_paulis=None)
    out, rest, md = backpropagate(wl.observables, wl.slices[2:])  # the last 10 layers: ~1.4e6 terms
    assert rest == []
    assert len(out) > 1_000_000
    assert abs(np.sum(np.abs(out.coeffs) ** 2) - 1.0) < 1e-6
    assert len(out.as_dict()) == len(out)


def test_max_seconds_timeout():
    """backpropagation.py:130-142, 288-295 (test_backpropagation.py:923-977): SIGALRM stops the loop between
    device calls, the committed state is returned, the alarm is cleared."""
    import signal
    import time

    from qiskit_addon_obp_b200 import slice_by_depth

    qc = QuantumCircuit(2)
    qc.rx(0.1, 0)
    qc.ry(0.1, 0)
    qc.rz(0.1, 0)
    qc.cx(0, 1)
    slices = slice_by_depth(qc, 1)
    obs = SparsePauliOp("IX")
    t0 = time.perf_counter()
    new_obs, rest, md = backpropagate(obs, 100_000 * slices, max_seconds=1)
    assert time.perf_counter() - t0 < 10
    assert len(rest) > 0 and md.num_backpropagated_slices > 0
    assert len(rest) == 400_000 - md.num_backpropagated_slices
    assert len(new_obs) >= 1
    assert signal.alarm(0) == 0  # no alarm left pending
    new_obs, rest, _ = backpropagate(obs, slices, max_seconds=1)
    assert len(rest) == 0
    assert signal.alarm(0) == 0


def test_linearity_config5_large():
    """Conjugation is linear: backprop(a*O1 + b*O2) == a*backprop(O1) + b*backprop(O2) at ~1e6 terms
    (no truncation; terms within the trim threshold of either side may differ)."""
    from qiskit_addon_obp_b200.workloads import config5_magic_sweep

    wl = config5_magic_sweep(theta=np.pi / 8, depth=12, max_paulis=None)
    slices = wl.slices[3:]
    nq = wl.num_qubits
    o1 = SparsePauliOp.from_sparse_list([("Z", [nq // 2], 1.0)], nq)
    o2 = SparsePauliOp.from_sparse_list([("X", [nq // 2 + 1], 1.0)], nq)
    a, b = 0.75, -1.25
    both = SparsePauliOp((np.vstack([o1.x, o2.x]), np.vstack([o1.z, o2.z])), np.array([a, b], dtype=complex))
    r1, _, _ = backpropagate(o1, slices)
    r2, _, _ = backpropagate(o2, slices)
    r12, _, _ = backpropagate(both, slices)
    assert len(r12) > 200_000
    d = {}
    for op, w in ((r1, a), (r2, b)):
        xw, zw = op.packed()
        for k, c in zip(map(bytes, np.concatenate([xw, zw], axis=1)), op.coeffs * w):
            d[k] = d.get(k, 0j) + c
    xw, zw = r12.packed()
    got = dict(zip(map(bytes, np.concatenate([xw, zw], axis=1)), r12.coeffs))
    assert len(got) == len(r12)
    worst = 0.0
    for k in set(d) | set(got):
        diff = abs(d.get(k, 0j) - got.get(k, 0j))
        if k in d and k in got:
            worst = max(worst, diff)
        else:
            # present on one side only: a descendant of terms trimmed (|c| <= 1e-8, not a linear operation)
            # in one run and kept in the other
            assert diff < 1e-6, diff
    assert worst < 1e-6


# ---- edge cases ---------------------------------------------------------------------------------------
def test_table_growth_and_rollback(monkeypatch):
    """A table that overflows mid-slice is rolled back, grown and the slice re-run (no max_paulis)."""
    monkeypatch.setenv("OBP_B200_CAPACITY", "4096")
    rng = np.random.default_rng(4)
    nq = 12
    slices = random_slices(nq, 8, rng, p_rot=0.9, gates_per_slice=10)
    obs = random_observable(nq, 4, rng)
    gl, md = _compare(obs, slices)
    assert len(gl[0]) > 4096  # it did outgrow the first table


def test_everything_trimmed_zero_operator():
    """simplify.py:156-159: when every term is trimmed the operator becomes one identity row with coefficient 0,
    and the loop keeps going on it exactly like the reference."""
    qc = QuantumCircuit(2)
    qc.rx(0.3, 0)
    qc2 = QuantumCircuit(2)
    qc2.h(0)
    qc2.cx(0, 1)
    obs = SparsePauliOp(["ZI", "IX"], [3e-9, 5e-9])  # below atol after the first gate's row filter
    _compare(obs, [qc2, qc, qc2])
    # (a second truncated slice would make the reference itself fail: max() of an empty array, truncating.py:173)
    _compare(obs, [qc], truncation_error_budget=setup_budget(max_error_per_slice=0.1))
    _compare([obs, SparsePauliOp("ZZ")], [qc2, qc], operator_budget=OperatorBudget(max_paulis=10))


def test_duplicate_and_zero_rows_in_the_input():
    """Inputs are not simplified by the reference before the first gate touches them; untouched observables
    come back unchanged, touched ones are merged by the first simplify."""
    qc = QuantumCircuit(3)
    qc.ry(0.4, 0)
    obs = SparsePauliOp(["IIZ", "IIZ", "IXI", "IIX"], [0.5, 0.25, 0.0, 1.0])
    _compare(obs, [qc])
    untouched = SparsePauliOp(["ZII", "ZII"], [1.0, 2.0])
    got, _, _ = backpropagate([obs, untouched], [qc])
    assert got[1] == untouched


def test_many_observables_union_count():
    rng = np.random.default_rng(8)
    nq = 6
    slices = random_slices(nq, 6, rng, gates_per_slice=5)
    obs = [random_observable(nq, 2, rng) for _ in range(5)] + [SparsePauliOp("IIIIIZ"), SparsePauliOp("IIIIIZ")]
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=150))
    _compare(obs, slices, operator_budget=OperatorBudget(max_paulis=40))


def test_wide_generator_rotation_across_words():
    """rzz / rxx / ryy whose two qubits sit in different 64-bit words (generator spans two words)."""
    nq = 200
    qc = QuantumCircuit(nq)
    qc.rzz(0.3, 63, 64)
    qc.rxx(0.5, 10, 199)
    qc.ryy(-0.7, 127, 128)
    qc.rzx(0.2, 64, 0)
    obs = SparsePauliOp.from_sparse_list([("XZ", [63, 64], 1.0), ("Y", [199], 0.5), ("ZX", [127, 128], -0.25)], nq)
    _compare(obs, [qc, qc])


# ---- C-ABI protocol details -------------------------------------------------------------------------
def test_submit_wait_protocol_and_replace_keeps_the_commit():
    """obp_apply_ops_submit / _wait: nothing else may touch the context in between, and the pair returns what
    obp_apply_ops returns.  obp_replace: new content, the last snapshot stays restorable."""
    from qiskit_addon_obp_b200.engine import EngineError, TermTable
    from qiskit_addon_obp_b200.gates import program_for

    qc = QuantumCircuit(3)
    qc.rx(0.3, 0)
    qc.cx(0, 1)
    qc.ry(-0.4, 1)
    ops = np.concatenate([program_for(ins).template for ins in qc.data])
    ops["qubit"][0, 0], ops["qubit"][1, :2], ops["qubit"][2, 0] = 0, (0, 1), 1
    ops["flags"][-1] = 1
    obs = SparsePauliOp(["ZZI", "IXY"], [1.0, 0.5])
    a = TermTable.acquire(3, 1024)
    b = TermTable.acquire(3, 1024)
    try:
        a.load(obs)
        b.load(obs)
        want = a.apply_ops(ops, 1e-8)
        b.apply_ops_submit(ops, 1e-8)
        with pytest.raises(EngineError, match="pending"):
            b._check(b._lib.obp_snapshot(b._h), "obp_snapshot")  # (the raw call: the wrapper would drain first)
        got = b.apply_ops_wait()
        for f in ("n_in", "n_out", "raw_rows", "updates", "num_unique", "num_duplicate", "num_trimmed"):
            assert getattr(got, f) == getattr(want, f), f
        assert_same_operator(b.download(), a.download())
        # replace: the commit survives
        a.snapshot()
        before = a.download()
        a.load(SparsePauliOp(["XXX"], [2.0]), keep_snapshot=True)
        assert a.download().as_dict() == {"XXX": 2.0}
        a.restore()
        assert_same_operator(a.download(), before)
        a.load(SparsePauliOp(["XXX"], [2.0]))  # a plain load invalidates it
        with pytest.raises(EngineError):
            a.restore()
    finally:
        a.release()
        b.release()


def window_workload(nq, rounds):
    """RX / RZZ+CX / H+RY layers on a 12-qubit window (straddling the 64-bit word boundary when there is one):
    the observable grows 4, 7, 14, 25, 26, 63, 225, 227, 598, 2788, 3501, 7738, 27839, 34740, 86138 terms."""
    lo = max(0, min(nq, 70) - 12)
    win = list(range(lo, min(nq, lo + 12)))
    slices = []
    for k in range(rounds):
        qc = QuantumCircuit(nq)
        for q in win:
            qc.rx(0.5 + 0.01 * q, q)
        slices.append(qc)
        qc = QuantumCircuit(nq)
        for a, b in zip(win[k % 2 :: 2], win[k % 2 + 1 :: 2]):
            qc.rzz(0.8, a, b)
            qc.cx(a, b)
        slices.append(qc)
        qc = QuantumCircuit(nq)
        for q in win:
            qc.h(q)
            qc.ry(0.4, q)
        slices.append(qc)
    obs = SparsePauliOp.from_sparse_list([("Z", [win[5]], 1.0), ("XY", [win[2], win[3]], 0.7)], nq)
    return obs, slices


# ---- every code path of the gate kernels against the oracle ----------------------
@pytest.mark.parametrize("mode", ["no-program", "switch-at-64-rows", "switch-at-2000-rows"])
@pytest.mark.parametrize("nq", [6, 64, 70, 127])
def test_program_and_standalone_kernels_agree_with_oracle(nq, mode, monkeypatch):
    """Tables of <= 128 qubits run a slice either inside the persistent program kernel or as stand-alone
    k_rotate / k_clifford_reg launches (>= 2^18 rows by default).  Forced here: everything stand-alone, and a
    row threshold low enough that the run changes path part-way — all against the oracle, counters included."""
    if mode == "no-program":
        monkeypatch.setenv("OBP_B200_NO_PROGRAM", "1")
    else:
        monkeypatch.setenv("OBP_B200_PROG_MAX_ROWS", "64" if "64" in mode else "2000")
    obs, slices = window_workload(nq, 4)
    gl, md = _compare(obs, slices)
    assert len(gl[0]) > 3000 or nq == 6, f"the run must cross the row thresholds to mean anything ({len(gl[0])} terms)"
    _compare(obs, slices, truncation_error_budget=setup_budget(max_error_per_slice=1e-3))


def test_program_switch_named_workloads(monkeypatch):
    """configs 2, 3 and 5 (oracle-sized) with the path switch in the middle of the run."""
    from qiskit_addon_obp_b200 import workloads as wl

    monkeypatch.setenv("OBP_B200_PROG_MAX_ROWS", "1500")
    for w in (
        wl.config2_kicked_ising(theta_h=0.3, steps=4, max_paulis=20_000),
        wl.config3_square_lattice(steps=2, max_paulis=10_000),
        wl.config5_magic_sweep(theta=np.pi / 8, depth=12, max_paulis=3_000, seed=1),
    ):
        _compare(w.observables, w.slices, **w.kwargs())


# ---- a slice that outgrows the device table (ADVICE) ------------------------------
def _growing_workload(max_paulis):
    """Slices of four layers each: the observable grows 60 -> 1558 terms in the second slice."""
    nq = 14
    slices = []
    for k in range(4):
        qc = QuantumCircuit(nq)
        for q in range(nq):
            qc.rx(0.6 + 0.01 * q, q)
        for q in range(k % 2, nq - 1, 2):
            qc.rzz(0.9, q, q + 1)
        for q in range(nq):
            qc.ry(0.5 + 0.01 * q, q)
        for q in range((k + 1) % 2, nq - 1, 2):
            qc.rzz(0.7, q, q + 1)
        slices.append(qc)
    obs = SparsePauliOp.from_sparse_list([("Z", [5], 1.0), ("Y", [6], 0.5)], nq)
    return obs, slices, OperatorBudget(max_paulis=max_paulis)


def test_oversized_slice_exact_policy_matches_reference_metadata(monkeypatch):
    """OBP_B200_OVERSIZE_POLICY=exact: the table grows mid-slice as often as needed and the slice is judged on
    its exact final count — sum_paulis and num_backpropagated_slices as backpropagation.py:266-275, 427-431."""
    from qiskit_addon_obp_b200 import backpropagation as bp
    from qiskit_addon_obp_b200 import engine

    engine.clear_pool()
    monkeypatch.setenv("OBP_B200_CAPACITY", "600")
    monkeypatch.setenv("OBP_B200_OVERSIZE_POLICY", "exact")
    obs, slices, budget = _growing_workload(400)
    got, rest, md = backpropagate(obs, slices, operator_budget=budget)
    want, rest_w, md_w = oracle.backpropagate(obs, slices, operator_budget=budget)
    assert len(rest) == len(rest_w) and 0 < len(rest) < len(slices)
    assert_same_operator(got, want)
    assert_same_metadata(md, md_w)
    assert md.backpropagation_history[-1].sum_paulis > 3 * 400, "the rejected slice must really overflow the first table"
    assert md.device_capacity_stops == []
    # the attempts that overflowed are rolled back and not counted; the slice's final (complete) run is, as in the
    # reference, whose per-gate sizes (backpropagation.py:239-242) include the slice it then rejects
    tr = oracle.GateTrace()
    oracle.backpropagate(obs, slices, operator_budget=budget, trace=tr)
    assert bp.last_run.wasted_updates > 0 and bp.last_run.updates == tr.updates
    engine.clear_pool()


def test_oversized_slice_lookahead_policy(monkeypatch, caplog):
    """Look-ahead policy (default factor 3, here 1.5): a slice that reaches factor x max_paulis live terms is
    rejected without being finished.  Same operator, same slices absorbed, same history up to the rejected slice;
    that slice reports a lower bound of its size (> max_paulis, <= the exact count) and the stop is flagged in
    OBPMetadata.device_capacity_stops."""
    from qiskit_addon_obp_b200 import engine

    engine.clear_pool()
    monkeypatch.setenv("OBP_B200_CAPACITY", "1024")
    monkeypatch.setenv("OBP_B200_OVERSIZE_LOOKAHEAD", "1.5")
    obs, slices, budget = _growing_workload(400)
    import logging

    with caplog.at_level(logging.INFO, logger="qiskit_addon_obp_b200.backpropagation"):
        got, rest, md = backpropagate(obs, slices, operator_budget=budget)
    print("\n".join(r.getMessage() for r in caplog.records))
    want, rest_w, md_w = oracle.backpropagate(obs, slices, operator_budget=budget)
    assert len(rest) == len(rest_w)
    assert_same_operator(got, want)
    assert md.num_backpropagated_slices == md_w.num_backpropagated_slices
    last, last_w = md.backpropagation_history[-1], md_w.backpropagation_history[-1]
    assert 1.5 * 400 <= last.sum_paulis <= last_w.sum_paulis
    assert last.num_paulis == [last.sum_paulis]
    assert md.device_capacity_stops == [(len(md.backpropagation_history) - 1, 0, "lookahead")]
    md.backpropagation_history.pop()
    md_w.backpropagation_history.pop()
    assert_same_metadata(md, md_w)
    engine.clear_pool()


def test_reserve_failure_keeps_the_table(monkeypatch):
    """obp_reserve is failure-atomic: an impossible capacity raises MemoryError and leaves the
    table with its capacity, its rows and a working index."""
    from qiskit_addon_obp_b200.engine import TermTable

    rng = np.random.default_rng(8)
    op = random_observable(20, 300, rng, weight=4)
    t = TermTable(20, 1024)
    try:
        t.load(op)
        with pytest.raises(MemoryError):
            t.reserve(0xFFFFFFF0)  # ~4.3e9 rows x 100+ B: more than any device has
        assert t.capacity == 1024 and not t._lib.obp_is_broken(t._h)
        assert_same_operator(t.download(), op)
        qc = QuantumCircuit(20)
        qc.rx(0.4, 3)
        qc.cx(3, 4)
    finally:
        t.close()
    got, _, _ = backpropagate(op, [qc])
    want, _, _ = oracle.backpropagate(op, [qc])
    assert_same_operator(got, want)


def test_per_gate_debug_log_matches_reference_sizes(caplog):
    """backpropagation.py:239-242: at DEBUG level the reference logs the observable's size after every applied
    gate (this is how its term-gate update count is measured).  The engine then runs every gate
    as a pass of its own and logs the same (gate id, size) sequence; results and counters are unchanged."""
    import logging
    import re

    rng = np.random.default_rng(99)
    nq = 7
    slices = random_slices(nq, 5, rng, gates_per_slice=7)
    slices[2].barrier()
    slices[2].rx(0.2, 3)
    ple = PauliLindbladError(["XX", "ZY"], [1e-2, 3e-3])
    slices[3].append(PauliLindbladErrorInstruction(ple), [1, 3])
    slices[4].cp(0.4, 0, 6)
    obs = [random_observable(nq, 3, rng), SparsePauliOp("IIIZIII")]
    tr = oracle.GateTrace()
    want, _, md_w = oracle.backpropagate(obs, slices, trace=tr)
    with caplog.at_level(logging.DEBUG, logger="qiskit_addon_obp_b200.backpropagation"):
        got, _, md = backpropagate(obs, slices)
    for g, w in zip(got, want):
        assert_same_operator(g, w)
    assert_same_metadata(md, md_w)
    pat = re.compile(r"Size of the observable after backpropagating gate id (\d+) in the current layer: (\d+)")
    logged = [tuple(map(int, m.groups())) for r in caplog.records if (m := pat.search(r.getMessage()))]
    assert logged == [(gid, size) for _, _, gid, size in tr.gate_sizes]
    assert sum(1 for _ in logged) == tr.applied_gates


# ---- QWC groups on the device (backpropagation.py:433-440) --------------------------
@pytest.mark.parametrize("nq,n,dens", [(1, 3, 0.9), (5, 40, 0.4), (10, 700, 0.35), (64, 1500, 0.05), (70, 3000, 0.04),
                                       (127, 5000, 0.02), (300, 2000, 0.01), (1000, 1200, 0.004)])
def test_qwc_group_count_device_matches_host_and_oracle(nq, n, dens, monkeypatch):
    """obp_qwc_group_count (degrees over all pairs + greedy colouring by descending degree, ties in row order)
    against the host statement on packed words and the oracle's restatement on the bool arrays."""
    from qiskit_addon_obp_b200 import engine
    from qiskit_addon_obp_b200.utils.qwc import _qwc_group_count_host, qwc_group_count

    rng = np.random.default_rng(nq * 7 + n)
    x = rng.random((n, nq)) < dens
    z = rng.random((n, nq)) < dens
    op = SparsePauliOp((x, z), np.ones(n))
    want = _qwc_group_count_host(op)
    assert engine.qwc_group_count(op) == want
    if n <= 1500:
        assert oracle.qwc_group_count(op) == want
    # a different row order may legitimately give another count; the same order must not
    perm = rng.permutation(n)
    op2 = SparsePauliOp((x[perm], z[perm]), np.ones(n))
    assert engine.qwc_group_count(op2) == _qwc_group_count_host(op2)
    monkeypatch.setenv("OBP_B200_QWC_DEVICE_MIN", "0")
    assert qwc_group_count(op) == want


def test_max_qwc_groups_through_backpropagate_on_device(monkeypatch):
    """max_qwc_groups with every count taken on the device (threshold 0), against the oracle's history."""
    monkeypatch.setenv("OBP_B200_QWC_DEVICE_MIN", "0")
    obs, slices = window_workload(20, 3)
    for row_order in ("0", "1"):
        monkeypatch.setenv("OBP_B200_ROW_ORDER", row_order)
        got, rest, md = backpropagate(obs, slices, operator_budget=OperatorBudget(max_qwc_groups=60))
        want, rest_w, md_w = oracle.backpropagate(obs, slices, operator_budget=OperatorBudget(max_qwc_groups=60))
        assert len(rest) == len(rest_w) and 0 < len(rest) < len(slices)
        assert_same_operator(got, want)
        if row_order == "1":  # the reference's row order: the greedy count is then the reference's, slice by slice
            assert [s.num_qwc_groups for s in md.backpropagation_history] == [s.num_qwc_groups for s in md_w.backpropagation_history]


@pytest.mark.parametrize("nq", [130, 200, 1000])
def test_rotation_kernel_wide_tables(nq):
    """Stand-alone k_rotate on tables wider than 128 qubits (W = 3, 4, 16), incl. generators that straddle two words."""
    rng = np.random.default_rng(nq)
    slices = random_slices(nq, 6, rng, gates_per_slice=8)
    for k, qc in enumerate(slices):
        qc.rzz(0.37, 63, 64)  # generator on two words
        qc.rxx(-0.6, 127, 128)
        qc.rx(0.45, 64 + k)
    obs = SparsePauliOp.from_sparse_list([("ZY", [63, 64], 1.0), ("X", [128], 0.5), ("ZZ", [64, 127], -0.3)], nq)
    _compare(obs, slices)
    _compare(obs, slices, truncation_error_budget=setup_budget(max_error_per_slice=1e-3))
