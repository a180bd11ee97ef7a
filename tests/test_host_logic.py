"""Host-side logic (no GPU): budgets, metadata replay against the reference's fixture, gate
classification, the packed host IR and the synthetic workload generators."""

from __future__ import annotations

import json
import os

import numpy as np
import pytest

from qiskit_addon_obp_b200 import gates
from qiskit_addon_obp_b200.engine import OBP_OP_CLIFFORD, OBP_OP_ROTATION
from qiskit_addon_obp_b200.ir import (
    Instruction,
    QuantumCircuit,
    SparsePauliOp,
    pack_xz,
    slice_by_barriers,
    slice_by_depth,
    topological_op_order,
    unpack_xz,
)
from qiskit_addon_obp_b200.utils.metadata import OBPMetadata, SliceMetadata
from qiskit_addon_obp_b200.utils.qwc import qwc_group_count
from qiskit_addon_obp_b200.utils.simplify import OperatorBudget
from qiskit_addon_obp_b200.utils.truncating import TruncationErrorBudget, setup_budget

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- utils/truncating.py (reference test/utils/test_truncating.py:24-60) -----------------------
def test_setup_budget_branches():
    assert setup_budget(max_error_per_slice=0.1, max_error_total=1.0) == TruncationErrorBudget([0.1], 1.0, 1)
    assert setup_budget(max_error_total=1.0, num_slices=5) == TruncationErrorBudget([0.2], 1.0, 1)
    assert setup_budget(max_error_total=1.0, num_slices=5, p_norm=2) == TruncationErrorBudget([1.0 / 5], 1.0, 2)
    assert setup_budget(max_error_per_slice=0.1, max_error_total=None) == TruncationErrorBudget([0.1], np.inf)
    assert setup_budget(max_error_per_slice=[0.1, 0.2]) == TruncationErrorBudget([0.1, 0.2], np.inf)
    assert setup_budget(max_error_total=0.3) == TruncationErrorBudget([0.3], 0.3)
    with pytest.raises(ValueError):
        setup_budget(max_error_per_slice=None, max_error_total=None)
    with pytest.raises(TypeError):  # utils/truncating.py:121-124: a non-float scalar goes through list()
        setup_budget(max_error_per_slice=1)


def test_budget_is_active():
    assert not TruncationErrorBudget().is_active()
    assert not TruncationErrorBudget([0.0, 0.0], 1.0).is_active()
    assert not TruncationErrorBudget([0.1], 0.0).is_active()
    assert TruncationErrorBudget([0.0, 0.1], np.inf).is_active()
    assert not OperatorBudget().is_active()
    assert OperatorBudget(max_paulis=3).is_active() and OperatorBudget(max_qwc_groups=3).is_active()
    b = OperatorBudget()
    assert (b.max_paulis, b.max_qwc_groups, b.simplify, b.atol, b.rtol) == (None, None, True, None, None)


# ---- utils/metadata.py against the reference's fixture ------------------------------------------
def test_metadata_budget_replay_golden():
    """test/utils/test_metadata.py:183-265 — both sequences to 10 places for both observables."""
    md = OBPMetadata.from_json(os.path.join(GOLDEN, "metadata_fixture.json"))
    want = json.load(open(os.path.join(GOLDEN, "metadata_expected.json")))
    assert len(md.backpropagation_history) == 11 and md.num_backpropagated_slices == 10
    for obs in (0, 1):
        got_acc = [md.accumulated_error(obs, i) for i in range(len(md.backpropagation_history))]
        got_left = [md.left_over_error_budget(obs, i) for i in range(len(md.backpropagation_history))]
        np.testing.assert_allclose(got_acc, want["accumulated_error"][obs], rtol=0, atol=5e-11)
        np.testing.assert_allclose(got_left, want["left_over_error_budget"][obs], rtol=0, atol=5e-11)


def test_metadata_json_round_trip(tmp_path):
    """test/utils/test_metadata.py:169-181 — from_json then to_json reproduces the file byte for byte."""
    src = os.path.join(GOLDEN, "metadata_fixture.json")
    md = OBPMetadata.from_json(src)
    assert md.truncation_error_budget == TruncationErrorBudget([0.001], 0.01, 1, 1e-6)
    assert md.operator_budget == OperatorBudget(50, 12, True, 1e-10, 1e-10)
    assert isinstance(md.backpropagation_history[0], SliceMetadata)
    assert md.backpropagation_history[2].slice_errors == [0.0028387171342200146, 0.0028377078723825274]
    out = tmp_path / "md.json"
    md.to_json(str(out), indent=2)
    assert out.read_text() == open(src).read()


# ---- gate classification ---------------------------------------------------------------------------
def _conj_code(matrix, code, m):
    """(image code, sign) of U^dag s_code U for a Clifford U, by dense arithmetic."""
    basis = gates.pauli_basis(m)
    img = matrix.conj().T @ basis[code] @ matrix
    for q in range(4**m):
        ov = np.trace(basis[q] @ img).real / 2**m
        if abs(abs(ov) - 1) < 1e-9:
            return q, ov < 0
    raise AssertionError("not a Pauli")


@pytest.mark.parametrize("name", ["id", "x", "y", "z", "h", "s", "sdg", "sx", "sxdg", "cx", "cz", "cy", "swap", "iswap", "ecr"])
def test_clifford_lut(name):
    nq = 1 if name in ("id", "x", "y", "z", "h", "s", "sdg", "sx", "sxdg") else 2
    ins = Instruction(name, range(nq))
    prog = gates.program_for(ins)
    assert prog.kind == OBP_OP_CLIFFORD
    rec = prog.template[0]
    for code in range(4**nq):
        img, neg = _conj_code(ins.matrix, code, nq)
        assert (int(rec["lut"]) >> (4 * code)) & 15 == img
        assert bool((int(rec["sign_mask"]) >> code) & 1) == neg
    k = int(rec["n_components"])
    assert abs(float(rec["row_scale"]) - 1.0 / k) < 1e-12  # uniform |u_j| = 1/sqrt(k)


def test_rotation_classification():
    for name, label, m in (("rx", 0b01, 1), ("ry", 0b11, 1), ("rz", 0b10, 1), ("rxx", 0b0101, 2), ("ryy", 0b1111, 2),
                           ("rzz", 0b1010, 2), ("rzx", 0b0110, 2)):
        prog = gates.program_for(Instruction(name, range(m), (0.37,)))
        rec = prog.template[0]
        assert prog.kind == OBP_OP_ROTATION and int(rec["gen_local"]) == label, name
        assert np.allclose(rec["u0"], [np.cos(0.37 / 2), 0]) and np.allclose(rec["u1"], [0, -np.sin(0.37 / 2)])
    for name in ("t", "tdg"):
        prog = gates.program_for(Instruction(name, (0,)))
        assert prog.kind == OBP_OP_ROTATION and int(prog.template[0]["gen_local"]) == 0b10
    # Clifford angles of a rotation family are Clifford-class (signed permutation), like rzz(-pi/2) in config 2
    assert gates.program_for(Instruction("rzz", (0, 1), (-np.pi / 2,))).kind == OBP_OP_CLIFFORD
    assert gates.program_for(Instruction("rz", (0,), (np.pi,))).kind == OBP_OP_CLIFFORD
    # anything else is run as the literal k*k raw-row expansion
    from qiskit_addon_obp_b200.engine import OBP_OP_GENERIC

    prog = gates.program_for(Instruction("cp", (0, 1), (0.3,)))
    assert prog.kind == OBP_OP_GENERIC and prog.n_components == 4 and sorted(prog.codes.tolist()) == [0, 2, 8, 10]
    m = sum(u * gates.pauli_basis(2)[c] for c, u in zip(prog.codes, prog.coeffs))
    assert np.allclose(m, Instruction("cp", (0, 1), (0.3,)).matrix)


def test_decompose_matches_oracle():
    from oracle import obp_oracle

    for ins in (Instruction("cx", (0, 1)), Instruction("ecr", (0, 1)), Instruction("rzx", (0, 1), (0.4,)), Instruction("sx", (0,))):
        codes, us = gates.pauli_decompose(ins.matrix)
        op = obp_oracle.pauli_decompose(ins.matrix)
        m = len(ins.qubits)
        lbls = ["".join("IXZY"[(c >> (2 * j)) & 3] for j in range(m - 1, -1, -1)) for c in codes]
        assert lbls == op.paulis.to_labels()
        np.testing.assert_allclose(us, op.coeffs, atol=1e-15)


# ---- host IR -----------------------------------------------------------------------------------------
def test_pack_unpack_and_lazy_operator():
    rng = np.random.default_rng(0)
    for nq in (1, 63, 64, 65, 127, 130, 1000):
        x = rng.random((37, nq)) < 0.3
        z = rng.random((37, nq)) < 0.3
        xw, zw = pack_xz(x, z)
        assert xw.shape == (37, (nq + 63) // 64) and xw.dtype == np.dtype("<u8")
        x2, z2 = unpack_xz(xw, zw, nq)
        assert np.array_equal(x, x2) and np.array_equal(z, z2)
        for j in (0, nq - 1, nq // 2):
            assert np.array_equal((xw[:, j // 64] >> np.uint64(j % 64)) & np.uint64(1), x[:, j].astype(np.uint64))
        op = SparsePauliOp.from_packed(xw, zw, rng.normal(size=37), nq)
        assert len(op) == 37 and op.num_qubits == nq
        assert op.packed()[0] is op._packed[0]  # still packed: nothing was materialised
        assert np.array_equal(op.copy().x, x)
        assert np.array_equal(op.z, z)
        op.z[:, 0] = True  # in-place edits of the bool view invalidate the packed cache
        assert np.all((op.packed()[1][:, 0] & np.uint64(1)) == 1)


def test_labels_are_big_endian():
    op = SparsePauliOp(["IIXYZ"], [2.0])
    assert op.x[0].tolist() == [False, True, True, False, False]  # qubit 0 = rightmost character
    assert op.z[0].tolist() == [True, True, False, False, False]
    assert op.paulis.to_labels() == ["IIXYZ"]
    assert SparsePauliOp.from_sparse_list([("XZ", [3, 0], 1.5)], 5).paulis.to_labels() == ["IXIIZ"]
    assert np.allclose(SparsePauliOp("Y").to_matrix(), [[0, -1j], [1j, 0]])
    assert np.allclose(SparsePauliOp("XZ").to_matrix(), np.kron([[0, 1], [1, 0]], [[1, 0], [0, -1]]))


def test_topological_order_and_slicing():
    qc = QuantumCircuit(3)
    qc.h(2)
    qc.cx(0, 1)
    qc.rz(0.1, 0)
    qc.barrier()
    qc.x(1)
    order = [(i.name, i.qubits) for i in topological_op_order(qc)]
    # lexicographic by qubit indices among ready nodes
    assert order.index(("cx", (0, 1))) < order.index(("rz", (0,)))
    assert order.index(("cx", (0, 1))) < order.index(("h", (2,)))
    assert [len(s) for s in slice_by_barriers(qc)] == [3, 1]
    qc = QuantumCircuit(2)
    qc.rx(0.1, 0)
    qc.ry(0.1, 0)
    qc.rz(0.1, 0)
    qc.cx(0, 1)
    assert [[i.name for i in s.data] for s in slice_by_depth(qc, 1)] == [["rx"], ["ry"], ["rz"], ["cx"]]
    assert len(slice_by_depth(qc, 2)) == 2
    with pytest.raises(ValueError):
        QuantumCircuit(2).cx(0, 0)
    with pytest.raises(ValueError):
        QuantumCircuit(2).x(2)


def test_gate_matrices_are_unitary_and_little_endian():
    for name in ("h", "s", "t", "sx", "cx", "cz", "cy", "swap", "iswap", "ecr"):
        m = Instruction(name, range(1 if len(name) <= 2 and name not in ("cx", "cz", "cy") else 2)).matrix
        assert np.allclose(m.conj().T @ m, np.eye(m.shape[0]))
    cx = Instruction("cx", (0, 1)).matrix  # control = first argument = least significant bit
    assert np.allclose(cx @ np.array([0, 1, 0, 0]), [0, 0, 0, 1])
    rzx = Instruction("rzx", (0, 1), (0.3,)).matrix
    zx = np.kron([[0, 1], [1, 0]], [[1, 0], [0, -1]])  # X on qubit 1 (left factor), Z on qubit 0
    assert np.allclose(rzx, np.cos(0.15) * np.eye(4) - 1j * np.sin(0.15) * zx)


def test_qwc_group_count():
    assert qwc_group_count(SparsePauliOp(["X", "Y", "Z"])) == 3
    assert qwc_group_count(SparsePauliOp(["XI", "IX", "XX", "ZZ"])) == 2
    assert qwc_group_count(SparsePauliOp(["ZI", "IZ", "ZZ", "II"])) == 1


# ---- workload generators -------------------------------------------------------------------------------
def test_workload_shapes():
    from qiskit_addon_obp_b200 import workloads as W

    edges = W.heavy_hex_127_edges()
    assert len(edges) == 144 and len({tuple(sorted(e)) for e in edges}) == 144
    deg = np.bincount(np.array(edges).ravel(), minlength=127)
    assert deg.max() == 3 and deg.min() >= 1
    layers = W.edge_colouring(edges)
    assert len(layers) == 3
    wl = W.config2_kicked_ising(steps=20)
    assert len(wl.slices) == 80 and wl.num_qubits == 127
    assert wl.observables.paulis.to_labels()[0][::-1][62] == "Z"
    wl = W.config1_xy_chain()
    assert len(wl.slices) == 18 and wl.observables.paulis.to_labels() == ["IIIIIZIIII"]
    wl = W.config3_square_lattice()
    assert wl.num_qubits == 100 and len(wl.observables) == 100 and len(wl.slices) == 50
    assert sum(len(s) for s in wl.slices[:4]) == 180
    wl = W.config4_near_clifford(depth=4)
    assert wl.num_qubits == 1000 and len(wl.observables) == 8
    a = W.config4_near_clifford(depth=3)
    names = lambda w: [[(i.name, i.qubits) for i in s.data] for s in w.slices]  # noqa: E731
    assert names(a) == names(W.config4_near_clifford(depth=3))  # seeded: deterministic
    wl = W.config5_magic_sweep(depth=3)
    assert sum(1 for i in wl.slices[0].data if i.name == "rz") == 64


# ---- slice plans: light cone + device steps (backpropagation.py:183 of the reference) ----------------
def _reference_light_cone(slice_, support: set[int]) -> tuple[list[int], set[int]]:
    """The reference's rule, stated with sets: walk the slice's gates in application order; a gate acts
    iff it shares a qubit with the observable's support, and then widens it (a reset does not)."""
    from qiskit_addon_obp_b200.ir import topological_op_order

    nodes = [n for n in topological_op_order(slice_)[::-1] if n.name != "barrier"]
    support = set(support)
    applied = []
    for k, node in enumerate(nodes):
        if support.isdisjoint(node.qubits):
            continue
        applied.append(k)
        if node.name != "reset":
            support |= set(node.qubits)
    return applied, support


@pytest.mark.parametrize("seed", range(6))
def test_slice_plan_matches_the_reference_light_cone(seed):
    from qiskit_addon_obp_b200.backpropagation import SliceProgram

    rng = np.random.default_rng(seed)
    nq = 12
    qc = QuantumCircuit(nq)
    for _ in range(14):
        r = rng.random()
        a, b = (int(q) for q in rng.choice(nq, size=2, replace=False))
        if r < 0.3:
            qc.rx(float(rng.uniform(-1, 1)), a)
        elif r < 0.5:
            qc.cx(a, b)
        elif r < 0.65:
            qc.rzz(float(rng.uniform(-1, 1)), a, b)
        elif r < 0.75:
            qc.reset(a)
        elif r < 0.85:
            qc.cp(0.3, a, b)  # generic gate: its own device step
        elif r < 0.9:
            qc.barrier()
        else:
            qc.h(a)
    prog = SliceProgram(qc)
    for _ in range(5):
        support = {int(q) for q in rng.choice(nq, size=int(rng.integers(1, 4)), replace=False)}
        mask = sum(1 << q for q in support)
        plan = prog.plan(mask)
        want_applied, want_support = _reference_light_cone(qc, support)
        assert plan.applied == want_applied
        assert plan.support_after == sum(1 << q for q in want_support)
        assert prog.plan(mask) is plan  # cached per (slice, support)
        # the steps partition the applied gates in order; the slice's last gate asks for exact counters
        seen = []
        for kind, what, seg in plan.steps:
            if kind == "ops":
                assert len(what) > 0 and len(seg) == len(what)
                seen += [None] * len(what)
            else:
                assert kind == "node" and prog.special[what]
                seen.append(what)
        assert len(seen) == len(want_applied)
        if plan.steps and plan.steps[-1][0] == "ops":
            assert plan.steps[-1][1]["flags"][-1] == 1
            assert not plan.steps[-1][1]["flags"][:-1].any()


def test_slice_plan_full_cone_shortcut_and_unsupported_gate():
    from qiskit_addon_obp_b200.backpropagation import SliceProgram

    qc = QuantumCircuit(3)
    qc.h(0)
    qc.cx(0, 1)
    qc.rz(0.2, 2)
    prog = SliceProgram(qc)
    full = prog.plan(0b111)
    assert full.applied == [0, 1, 2] and full.support_after == 0b111
    assert prog.plan(0b100).applied == [k for k, n in enumerate(prog.nodes) if n.name == "rz"]
    # a gate the engine cannot express only raises when the light cone reaches it
    bad = QuantumCircuit(6)
    bad.unitary(np.linalg.qr(np.random.default_rng(0).normal(size=(32, 32)) + 0j)[0], [0, 1, 2, 3, 4])
    bad.h(5)
    prog = SliceProgram(bad)
    assert [k for k, _, _ in prog.plan(1 << 5).steps] == ["ops"]
    kinds = [k for k, _, _ in prog.plan(0b1).steps]
    assert kinds[-1] == "error"


@pytest.mark.parametrize("nq,n", [(1, 3), (5, 40), (70, 200), (130, 300)])
def test_qwc_group_count_on_packed_words_matches_the_bool_version(nq, n):
    from qiskit_addon_obp_b200.utils.qwc import _qwc_group_count_bool, _qwc_group_count_host

    rng = np.random.default_rng(nq + n)
    op = SparsePauliOp((rng.random((n, nq)) < 0.3, rng.random((n, nq)) < 0.3), np.ones(n))
    assert _qwc_group_count_host(op) == _qwc_group_count_bool(op)


# ---- compiled-slice cache: valid only for the very same instruction objects ---------------
def test_slice_program_cache_follows_in_place_edits():
    from qiskit_addon_obp_b200.backpropagation import _slice_program

    qc = QuantumCircuit(3)
    qc.rx(0.3, 0)
    qc.cx(0, 1)
    p1 = _slice_program(qc)
    assert _slice_program(qc) is p1
    # same length, different angle: a parameter sweep that reuses the slice object
    qc.data[0] = Instruction("rx", (0,), (0.7,))
    p2 = _slice_program(qc)
    assert p2 is not p1
    assert float(p2.recs["u0"][[n.name for n in p2.nodes].index("rx")][0]) == pytest.approx(np.cos(0.35))
    with pytest.raises(AttributeError):
        qc.data[0].params = (0.9,)  # instructions cannot change behind their identity
    m = np.eye(2, dtype=complex)
    ins = Instruction("unitary", (1,), (), m)
    m[0, 0] = -1  # the caller's array is not the instruction's
    assert ins.matrix[0, 0] == 1
    with pytest.raises(ValueError):
        ins.matrix[0, 0] = 5


def test_segment_partition_map_vanishes_on_every_generator():
    """The host step of the slice-in-shared-memory path (csrc/obp_segment.cuh): a GF(2)-linear map keys -> 16 bits
    that vanishes on every rotation generator of the run, so that P and P xor (any product of generators) share a
    partition — for generators on disjoint qubits (an RX / RZZ layer), for overlapping and for dependent ones — and
    that still tells keys apart that differ outside the generators' span."""
    import ctypes as C

    from qiskit_addon_obp_b200 import engine

    lib = engine.lib()
    rng = np.random.default_rng(7)

    def fmap(fm, x, z):
        v = 0
        for j in range(16):
            a = 0
            for w in range(len(x)):
                a ^= (int(x[w]) & int(fm[4 * j + 2 * w])) ^ (int(z[w]) & int(fm[4 * j + 2 * w + 1]))
            v |= (bin(a).count("1") & 1) << j
        return v

    for n_words, nq in ((1, 40), (2, 127)):
        layers = {
            "rx": [({q: "X"}) for q in range(nq)],
            "rzz": [({q: "Z", q + 1: "Z"}) for q in range(0, nq - 1, 2)],
            "xx+yy same bonds": [({q: p, q + 1: p}) for q in range(0, nq - 1, 2) for p in "XY"],
            "overlapping chain": [({q: "X", q + 1: "Z", q + 2: "Y"}) for q in range(0, nq - 2)],
            "random": [dict(zip(rng.choice(nq, 3, replace=False).tolist(), rng.choice(list("XYZ"), 3))) for _ in range(60)],
        }
        for name, gens in layers.items():
            gx = np.zeros((len(gens), n_words), dtype=np.uint64)
            gz = np.zeros((len(gens), n_words), dtype=np.uint64)
            for k, g in enumerate(gens):
                for q, p in g.items():
                    if p in "XY":
                        gx[k, q // 64] |= np.uint64(1 << (q % 64))
                    if p in "ZY":
                        gz[k, q // 64] |= np.uint64(1 << (q % 64))
            fm = np.zeros(64, dtype=np.uint64)
            rc = lib.obp_segment_partition_map(n_words, len(gens), gx.ctypes.data_as(C.c_void_p), gz.ctypes.data_as(C.c_void_p),
                                               fm.ctypes.data_as(C.c_void_p))
            assert rc == 0, name
            for k in range(len(gens)):
                assert fmap(fm, gx[k], gz[k]) == 0, (name, k)
            # linear: f(P xor G) == f(P); and not degenerate unless the generators span (almost) everything
            seen = set()
            for _ in range(200):
                x = rng.integers(0, 1 << 63, size=n_words, dtype=np.uint64)
                z = rng.integers(0, 1 << 63, size=n_words, dtype=np.uint64)
                k = int(rng.integers(len(gens)))
                assert fmap(fm, x ^ gx[k], z ^ gz[k]) == fmap(fm, x, z), name
                seen.add(fmap(fm, x, z))
            assert len(seen) > 150, (name, len(seen))


def test_clifford_only_slices_are_grouped_structurally():
    """``_merge_clifford_slices`` (the opt-in OBP_B200_MERGE_CLIFFORD_SLICES path): consecutive slices whose applied
    gates are all Clifford-class form a group that ends at the first slice with a rotation — decided on the compiled
    plans alone (light cone included), no device needed."""
    from qiskit_addon_obp_b200 import backpropagation as bp
    from qiskit_addon_obp_b200 import workloads as wl

    w = wl.config2_kicked_ising(theta_h=0.3, steps=2, max_paulis=1000)
    rev = w.slices[::-1]  # backpropagation order: RX layer first, then the three RZZ(-pi/2) colour layers
    support = (1 << w.num_qubits) - 1  # every qubit in the light cone
    assert bp._merge_clifford_slices(None, rev, 0, support, False) == []  # the RX slice is not Clifford-only
    group = bp._merge_clifford_slices(None, rev, 1, support, False)
    assert len(group) == 3
    for prog, plan in group:
        ops = plan.steps[0][1]
        assert (ops["kind"] == bp.OBP_OP_CLIFFORD).all()
        assert ops["flags"][-1] & bp.OBP_FLAG_EXACT_COUNTERS and not (ops["flags"][:-1] & bp.OBP_FLAG_EXACT_COUNTERS).any()
    assert len(bp._merge_clifford_slices(None, rev, 1, support, False, max_group=2)) == 2
    # a slice outside the light cone (no applied gate) ends a group
    one = 1 << 62
    g1 = bp._merge_clifford_slices(None, rev, 1, one, False)
    assert all(plan.applied for _, plan in g1)
