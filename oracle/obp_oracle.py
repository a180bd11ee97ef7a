"""CPU ORACLE — test infrastructure only, never the product path.

A plain-numpy restatement of the reference's operator-backpropagation hot path
(``/root/reference/qiskit_addon_obp``), written so that the CUDA engine in
``qiskit_addon_obp_b200`` can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.

Why a restatement: the reference is pure Python but imports ``qiskit``, ``qiskit._accelerate``
(Rust) and ``qiskit_ibm_runtime`` at module import (``backpropagation.py:23-25``,
``utils/simplify.py:21-23``, ``utils/operations.py:19-20``); none of them is installed or
installable here (no network).  The arithmetic that lives in those third-party packages
(qiskit ``>=1.2,<3`` per ``pyproject.toml:31``, no lock file) is restated from their published
semantics and marked **[EXT]** below.

Pinning: this oracle is checked (``tests/test_oracle_kat.py``) against every known-answer test the
reference's own test-suite holds for the path and that can be stated without Qiskit objects —
``test/test_backpropagation.py:33-631`` (incl. the numeric KATs at :389-443 with their row order
and accumulated errors), ``test/utils/test_simplify.py:24-136``, ``test/utils/test_operations.py``,
``test/utils/test_metadata.py:183-265`` — and against the executed notebook outputs the reference ships
(``docs/guides/truncate_operator_terms.ipynb`` cells 10-49: slices absorbed / terms / QWC groups of seven
runs; ``docs/guides/bound_error_using_p_norm.ipynb``: 116/10 and 84/6 terms/groups and the L2 expectation
value 9.317829770853422 to the last digit; ``tests/test_notebook_anchors.py``).

Every function cites the reference lines it follows.
"""

from __future__ import annotations

import copy
import itertools
from dataclasses import dataclass, field

import numpy as np

# The only things taken from the product package are two passive containers: the bool[n, N] x/z +
# complex128 operator (``SparsePauliOp`` work-alike) and the list-of-instructions circuit.  Everything
# that *decides* a result — gate matrices, the gate order inside a slice, the budget replay, the
# simplify/truncation arithmetic and the metadata records — is restated here, independently of the
# product's host code, so that a wrong convention there cannot cancel out in an engine-vs-oracle test.
from qiskit_addon_obp_b200.ir import QuantumCircuit, SparsePauliOp

ATOL_DEFAULT = 1e-8  # [EXT] SparsePauliOp.atol
RTOL_DEFAULT = 1e-5  # [EXT] SparsePauliOp.rtol


# ---------------------------------------------------------------------------------------------
# records (field-for-field: utils/simplify.py:26-77, utils/truncating.py:25-58, utils/metadata.py:25-222)
# ---------------------------------------------------------------------------------------------
@dataclass
class SimplifyMetadata:
    """``utils/simplify.py:26-40``."""

    num_unique_paulis: int
    num_duplicate_paulis: int
    num_trimmed_paulis: int
    sum_trimmed_coeffs: float


@dataclass
class OperatorBudget:
    """``utils/simplify.py:43-77``."""

    max_paulis: int | None = None
    max_qwc_groups: int | None = None
    simplify: bool = True
    atol: float | None = None
    rtol: float | None = None

    def is_active(self) -> bool:  # :73-77
        return self.max_paulis is not None or self.max_qwc_groups is not None


@dataclass
class TruncationErrorBudget:
    """``utils/truncating.py:25-58``."""

    per_slice_budget: list = field(default_factory=list)
    max_error_total: float = 0.0
    p_norm: int = 1
    tol: float = 1e-8

    def is_active(self) -> bool:  # :56-58
        return (
            any(budget > 0 for budget in self.per_slice_budget) and self.max_error_total > 0
        )


@dataclass
class SliceMetadata:
    """``utils/metadata.py:25-69``."""

    slice_errors: list
    raw_num_paulis: list
    num_unique_paulis: list | None
    num_duplicate_paulis: list | None
    num_trimmed_paulis: list | None
    sum_trimmed_coeffs: list | None
    num_truncated_paulis: list
    num_paulis: list
    sum_paulis: int | None
    num_qwc_groups: int | None


@dataclass
class OBPMetadata:
    """``utils/metadata.py:72-222`` — the two budget functions written out as the reference has them
    (``accumulated_error`` re-summed from zero inside every step of the replay, :152-157, :197-222)."""

    truncation_error_budget: object
    num_slices: int | None
    operator_budget: object
    backpropagation_history: list
    num_backpropagated_slices: int

    def accumulated_error(self, observable_idx: int, slice_idx: int | None = None) -> float:
        if slice_idx is None:
            slice_idx = self.num_backpropagated_slices
        accumulated_error = 0.0
        for i in range(slice_idx):  # :153-155
            accumulated_error += self.backpropagation_history[i].slice_errors[observable_idx]
        return accumulated_error

    def left_over_error_budget(self, observable_idx: int, slice_idx: int | None = None) -> float:
        if slice_idx is None:
            slice_idx = self.num_backpropagated_slices
        tb = self.truncation_error_budget
        left_over = 0.0
        for i in range(slice_idx + 1):  # :198
            left_over = left_over + tb.per_slice_budget[i % len(tb.per_slice_budget)]  # :202-207
            if i > 0:  # :210-213
                left_over -= self.backpropagation_history[i - 1].slice_errors[observable_idx]
            left_over = min(left_over, tb.max_error_total - self.accumulated_error(observable_idx, i))  # :216-220
        return left_over

    @classmethod
    def from_json(cls, json_file: str) -> "OBPMetadata":  # :224-243
        import json

        with open(json_file) as fh:
            data = json.load(fh)
        data["truncation_error_budget"] = TruncationErrorBudget(**data["truncation_error_budget"])
        data["operator_budget"] = OperatorBudget(**data["operator_budget"])
        data["backpropagation_history"] = [SliceMetadata(**s) for s in data["backpropagation_history"]]
        return cls(**data)


# ---------------------------------------------------------------------------------------------
# [EXT] Operator(gate).data — little-endian unitaries of Qiskit's standard gates, built here from
# their Pauli closed forms (qubit 0 of the gate = least-significant index bit).  Independent of the
# literal matrices in the product's ir.py; tests/test_oracle_kat.py compares the two tables.
# ---------------------------------------------------------------------------------------------
def _pauli_le(label: str) -> np.ndarray:
    """Matrix of a little-endian Pauli label: label[0] acts on the gate's qubit 0."""
    m = np.array([[1.0 + 0j]])
    for ch in label:
        m = np.kron(_P1[ch], m)
    return m


def _pauli_sum(terms: dict) -> np.ndarray:
    return sum(c * _pauli_le(lbl) for lbl, c in terms.items())


_ROT_GEN = {"rx": "X", "ry": "Y", "rz": "Z", "rxx": "XX", "ryy": "YY", "rzz": "ZZ", "rzx": "ZX"}  # rzx: Z on arg 0, X on arg 1


def gate_matrix(name: str, params=()) -> np.ndarray:
    """Unitary of a named Qiskit standard gate from its Pauli closed form."""
    r2 = 1.0 / np.sqrt(2.0)
    if name in _ROT_GEN:  # exp(-i theta/2 P)
        t = params[0] / 2
        g = _ROT_GEN[name]
        return _pauli_sum({"I" * len(g): np.cos(t), g: -1j * np.sin(t)})
    if name == "id":
        return _pauli_le("I")
    if name in ("x", "y", "z"):
        return _pauli_le(name.upper())
    if name == "h":
        return _pauli_sum({"X": r2, "Z": r2})
    phase_gates = {"s": np.pi / 2, "sdg": -np.pi / 2, "t": np.pi / 4, "tdg": -np.pi / 4}
    if name in phase_gates or name == "p":  # diag(1, e^{i lam}) = (1+e)/2 I + (1-e)/2 Z
        e = np.exp(1j * (params[0] if name == "p" else phase_gates[name]))
        return _pauli_sum({"I": (1 + e) / 2, "Z": (1 - e) / 2})
    if name in ("sx", "sxdg"):  # sqrt(X) = (1+i)/2 I + (1-i)/2 X
        a = (1 + 1j) / 2 if name == "sx" else (1 - 1j) / 2
        return _pauli_sum({"I": a, "X": np.conj(a)})
    ctrl = {"cx": "X", "cy": "Y", "cz": "Z"}
    if name in ctrl:  # |0><0|_0 (x) I + |1><1|_0 (x) P_1 = (II + Z_0 + P_1 - Z_0 P_1) / 2
        p = ctrl[name]
        return _pauli_sum({"II": 0.5, "ZI": 0.5, "I" + p: 0.5, "Z" + p: -0.5})
    if name == "swap":
        return _pauli_sum({"II": 0.5, "XX": 0.5, "YY": 0.5, "ZZ": 0.5})
    if name == "iswap":  # exp(i pi/4 (XX + YY))
        return _pauli_sum({"II": 0.5, "ZZ": 0.5, "XX": 0.5j, "YY": 0.5j})
    if name == "ecr":  # (IX - XY)/sqrt(2) in Qiskit's big-endian labels = (X_0 - Y_0 X_1)/sqrt(2)
        return _pauli_sum({"XI": r2, "YX": -r2})
    if name == "cp":  # II + (e - 1) |11><11|
        e = np.exp(1j * params[0]) - 1
        return _pauli_sum({"II": 1 + e / 4, "ZI": -e / 4, "IZ": -e / 4, "ZZ": e / 4})
    if name == "crz":  # |0><0|_0 (x) I + |1><1|_0 (x) RZ(theta)_1
        c, s = np.cos(params[0] / 2), -1j * np.sin(params[0] / 2)
        return _pauli_sum({"II": (1 + c) / 2, "ZI": (1 - c) / 2, "IZ": s / 2, "ZZ": -s / 2})
    raise KeyError(f"oracle: no closed form for gate {name!r}")


def _node_matrix(node) -> np.ndarray:
    explicit = getattr(node, "_matrix", None)
    return np.asarray(explicit, dtype=complex) if explicit is not None else gate_matrix(node.name, node.params)


# ---------------------------------------------------------------------------------------------
# [EXT] circuit_to_dag(slice).topological_op_nodes()  — backpropagation.py:170
# ---------------------------------------------------------------------------------------------
def gate_order(circuit) -> list:
    """Instructions in lexicographical topological order (key: the qubit indices, zero-padded and
    joined — tuples of ints compare the same way); a gate is ready once every earlier gate sharing a
    qubit with it has been emitted.  **Parity unpinned** (SURVEY §8c (i)): how Qiskit's wire-input
    nodes interleave with ready gates is not observable from any reference test; the order only
    decides which gate's counters a slice reports and the last bits of fp sums."""
    data = list(circuit.data)
    blockers: list[set] = []
    latest: dict = {}
    for idx, ins in enumerate(data):
        blockers.append({latest[q] for q in ins.qubits if q in latest})
        for q in ins.qubits:
            latest[q] = idx
    emitted: list[int] = []
    done: set = set()
    remaining = set(range(len(data)))
    while remaining:
        ready = [i for i in remaining if blockers[i] <= done]
        nxt = min(ready, key=lambda i: (tuple(data[i].qubits), i))
        emitted.append(nxt)
        done.add(nxt)
        remaining.discard(nxt)
    return [data[i] for i in emitted]

_P1 = {
    "I": np.eye(2, dtype=complex),
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
}


# ---------------------------------------------------------------------------------------------
# [EXT] SparsePauliOp.from_operator(Operator(gate))   — backpropagation.py:213
# ---------------------------------------------------------------------------------------------
def pauli_decompose(matrix: np.ndarray, atol: float = ATOL_DEFAULT) -> SparsePauliOp:
    """Dense 2^m x 2^m matrix -> Pauli sum, dropping components with ``|coeff| <= atol``.

    Restates [EXT] ``SparsePauliOp.from_operator`` (trace inner product with every m-qubit Pauli,
    ``np.isclose(coeff, 0, atol, rtol)`` test).  Output order: labels in I<X<Y<Z lexicographic
    order, highest qubit most significant.  The order Qiskit's Rust ``decompose_dense`` emits is
    not pinned by any reference test (SURVEY §8c "parity unpinned (iii)"); it only affects the
    order of the raw rows, never the merged result beyond the last fp bit.
    """
    m = int(round(np.log2(matrix.shape[0])))
    labels, coeffs = [], []
    for lbl in itertools.product("IXYZ", repeat=m):
        p = np.array([[1.0 + 0j]])
        for ch in lbl:  # big-endian label: leftmost = highest qubit = leftmost kron factor
            p = np.kron(p, _P1[ch])
        c = np.trace(p.conj().T @ matrix) / 2**m
        if not abs(c) <= atol:
            labels.append("".join(lbl))
            coeffs.append(c)
    return SparsePauliOp(labels, coeffs)


# ---------------------------------------------------------------------------------------------
# [EXT] SparsePauliOp.compose(other, qargs, front)  — utils/operations.py:103-107
# ---------------------------------------------------------------------------------------------
def _count_y(x: np.ndarray, z: np.ndarray) -> np.ndarray:
    return np.logical_and(x, z).sum(axis=1).astype(np.int64)


def compose(op: SparsePauliOp, other: SparsePauliOp, qargs: list[int], front: bool) -> SparsePauliOp:
    """Outer-product composition, rows ordered (term of ``op``)-major, then (term of ``other``).

    [EXT] Qiskit stores a Pauli as ``(-i)^phase Z^z X^x``; a Hermitian label has
    ``phase = count_y mod 4``.  ``compose`` adds the two phases plus ``2*|x1 & z2|`` (front) or
    ``2*|z1 & x2|`` (back), XORs x and z on ``qargs`` and multiplies the coefficients with
    ``np.multiply.outer``; the ``SparsePauliOp`` constructor then folds ``(-i)^(phase - count_y)``
    into the coefficients so that the stored Paulis are Hermitian labels again.
    """
    k = len(other)
    x1, z1 = op.x[:, qargs], op.z[:, qargs]
    x2, z2 = other.x, other.z
    m = other.num_qubits
    phase = np.add.outer(_count_y(op.x, op.z) % 4, _count_y(x2, z2) % 4).reshape(-1)
    if front:
        q = np.logical_and(x1[:, np.newaxis], z2).reshape((-1, m))
    else:
        q = np.logical_and(z1[:, np.newaxis], x2).reshape((-1, m))
    phase = phase + 2 * q.sum(axis=1)
    x3 = np.logical_xor(x1[:, np.newaxis], x2).reshape((-1, m))
    z3 = np.logical_xor(z1[:, np.newaxis], z2).reshape((-1, m))
    x4 = np.repeat(op.x, k, axis=0)
    z4 = np.repeat(op.z, k, axis=0)
    x4[:, qargs] = x3
    z4[:, qargs] = z3
    coeffs = np.multiply.outer(op.coeffs, other.coeffs).ravel()
    e = (phase - _count_y(x4, z4)) % 4
    coeffs = np.asarray((-1j) ** e * coeffs, dtype=complex)
    return SparsePauliOp((x4, z4), coeffs)


_DECOMPOSE_CACHE: dict = {}


def _decompose_cached(node) -> SparsePauliOp:
    """``from_operator`` per gate as in :213, memoised: the reference's Rust ``decompose_dense`` costs
    microseconds, this numpy restatement ~100 us — caching keeps the CPU baseline from being
    dominated by an artefact of the port."""
    explicit = getattr(node, "_matrix", None)
    key = (node.name, tuple(node.params)) if explicit is None else (node.name, np.asarray(explicit).tobytes())
    op = _DECOMPOSE_CACHE.get(key)
    if op is None:
        op = _DECOMPOSE_CACHE[key] = pauli_decompose(_node_matrix(node))
    return op


def adjoint(op: SparsePauliOp) -> SparsePauliOp:
    """[EXT] ``SparsePauliOp.adjoint``: Hermitian labels are kept, coefficients conjugated."""
    return SparsePauliOp((op.x.copy(), op.z.copy()), op.coeffs.conj())


# ---------------------------------------------------------------------------------------------
# utils/operations.py
# ---------------------------------------------------------------------------------------------
def apply_op_to(op: SparsePauliOp, gate: SparsePauliOp, qargs: list[int]) -> SparsePauliOp:
    """``op -> gate† op gate`` on global-width rows (``utils/operations.py:100-107``).

    The reference works on a lean operator restricted to ``op1_qargs`` and pads it with
    ``apply_layout`` when the support grows (:23-51); identity columns change nothing in the
    arithmetic, so the oracle keeps global width and tracks the support set separately.
    """
    return compose(compose(op, gate, qargs, front=True), adjoint(gate), qargs, front=False)


def apply_reset_to(op: SparsePauliOp, qubit: int) -> SparsePauliOp:
    """``utils/operations.py:216-225``: zero X/Y terms on ``qubit``, turn Z into I there."""
    op = op.copy()
    op.coeffs[op.x[:, qubit]] = 0.0
    op.z[:, qubit] = False
    op.x[:, qubit] = False
    return op


def apply_ple_to(op: SparsePauliOp, generators, rates, qargs: list[int]) -> SparsePauliOp:
    """``utils/operations.py:262-275``: multiply by exp(-2 rate) per anticommuting generator.

    ``generators`` are big-endian labels on ``len(qargs)`` qubits; qubit j of a generator acts on
    ``qargs[j]`` ([EXT] ``apply_layout``).
    """
    gen_op = SparsePauliOp(list(generators))
    gx = np.zeros((len(gen_op), op.num_qubits), dtype=bool)
    gz = np.zeros((len(gen_op), op.num_qubits), dtype=bool)
    gx[:, qargs] = gen_op.x
    gz[:, qargs] = gen_op.z
    new_coeffs = []
    for xr, zr, c in zip(op.x, op.z, op.coeffs):
        coeff = c
        for g in range(len(gen_op)):
            anti = (np.logical_and(xr, gz[g]).sum() + np.logical_and(zr, gx[g]).sum()) % 2
            if anti:
                coeff *= np.exp(-2 * rates[g])
        new_coeffs.append(coeff)
    return SparsePauliOp((op.x.copy(), op.z.copy()), new_coeffs)


def support_of(op: SparsePauliOp) -> list[int]:
    """``reduce_op`` (``utils/operations.py:176-181``): ascending non-identity columns."""
    qargs = [int(q) for q in np.logical_or(op.x, op.z).any(axis=0).nonzero()[0]]
    if qargs == []:
        raise ValueError("Input operator may not be the identity operator.")
    return qargs


# ---------------------------------------------------------------------------------------------
# utils/simplify.py:80-169
# ---------------------------------------------------------------------------------------------
def unordered_unique(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """[EXT] ``qiskit._accelerate.sparse_pauli_op.unordered_unique`` (``simplify.py:126``).

    ``indexes[g]`` = row of the first occurrence of group ``g``, groups numbered in
    first-appearance order; ``inverses[i]`` = group of row ``i``.
    """
    if keys.shape[0] == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    # one opaque item per row: np.unique then sorts fixed-size byte strings instead of doing a
    # column-by-column lexsort (several times faster; the Rust original hashes, which is faster still)
    rows = np.ascontiguousarray(keys).view(np.dtype((np.void, keys.dtype.itemsize * keys.shape[1]))).reshape(-1)
    _, first, inv = np.unique(rows, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(first, kind="stable")  # sorted-group id -> appearance rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return first[order], rank[inv]


def simplify(
    op: SparsePauliOp, *, atol: float | None = None, rtol: float | None = None
) -> tuple[SparsePauliOp, SimplifyMetadata]:
    """Line-by-line restatement of ``utils/simplify.py:102-169`` for complex128 coefficients."""
    if atol is None:
        atol = ATOL_DEFAULT
    if rtol is None:
        rtol = RTOL_DEFAULT
    non_zero = np.logical_not(np.isclose(op.coeffs, 0, atol=atol, rtol=rtol))  # :120
    px, pz, nz = op.x[non_zero], op.z[non_zero], op.coeffs[non_zero]  # :121-123
    array = np.concatenate(
        [np.packbits(px, axis=1), np.packbits(pz, axis=1)], axis=1
    )  # :125 (any injective packing)
    indexes, inverses = unordered_unique(array)  # :126
    md = SimplifyMetadata(
        num_unique_paulis=len(indexes),
        num_duplicate_paulis=len(nz) - len(indexes),
        num_trimmed_paulis=len(op.coeffs) - len(nz),
        sum_trimmed_coeffs=0.0,
    )  # :128-133
    if np.all(non_zero) and indexes.shape[0] == array.shape[0]:  # :135-137
        return op.copy(), md
    coeffs = np.zeros(indexes.shape[0], dtype=complex)
    np.add.at(coeffs, inverses, nz)  # :139-140
    is_zero = np.isclose(coeffs, 0, atol=atol, rtol=rtol)  # :147
    md.num_trimmed_paulis += int(sum(is_zero))  # :149
    md.sum_trimmed_coeffs += sum(coeffs[is_zero])  # :150
    if np.all(is_zero):  # :156-159
        x = np.zeros((1, op.num_qubits), dtype=bool)
        z = np.zeros((1, op.num_qubits), dtype=bool)
        coeffs = np.array([0j])
    else:  # :160-165
        keep = np.logical_not(is_zero)
        idx = indexes[keep]
        x, z, coeffs = px[idx], pz[idx], coeffs[keep]
    return SparsePauliOp((x, z), coeffs), md


# ---------------------------------------------------------------------------------------------
# utils/truncating.py:144-204
# ---------------------------------------------------------------------------------------------
def truncate_binary_search(
    op: SparsePauliOp, budget: float, *, p_norm: int = 1, tol: float = 1e-8
) -> tuple[SparsePauliOp, float]:
    abscs = np.abs(op.coeffs) ** p_norm  # :171
    upper_threshold = max(abscs)  # :173
    lower_threshold = 0.0
    upper_error = budget
    lower_error = 0.0
    while ((upper_threshold - lower_threshold) > tol) and not (
        np.isclose(upper_error, lower_error, atol=tol)
    ):  # :179-181
        mid_threshold = (upper_threshold + lower_threshold) / 2
        mid_error = np.power(np.sum(abscs[abscs < mid_threshold]), 1.0 / p_norm)  # :187
        if mid_error <= budget:
            lower_threshold = mid_threshold
            lower_error = mid_error
        else:
            upper_threshold = mid_threshold
            upper_error = mid_error
    to_keep = abscs > lower_threshold  # :199
    return SparsePauliOp((op.x[to_keep], op.z[to_keep]), op.coeffs[to_keep]), float(lower_error)


# ---------------------------------------------------------------------------------------------
# backpropagation.py:48-311
# ---------------------------------------------------------------------------------------------
def _union_size(observables: list[SparsePauliOp]) -> int:
    """``_observable_oversized`` (:419-428): distinct Pauli keys over all observables."""
    keys = np.concatenate(
        [
            np.concatenate([np.packbits(o.x, axis=1), np.packbits(o.z, axis=1)], axis=1)
            for o in observables
        ],
        axis=0,
    )
    return int(np.unique(keys, axis=0).shape[0])


def qwc_group_count(op: SparsePauliOp) -> int:
    """[EXT] ``len(op.group_commuting(qubit_wise=True))`` (:369, :434) — **parity unpinned**.

    Qiskit builds the qubit-wise non-commutation graph and colours it with rustworkx
    ``graph_greedy_color`` (nodes by descending degree, stable).  This is restated from the
    published algorithm; the reference's tests pin it only on 1-qubit operators
    (``test_backpropagation.py:478-509``).
    """
    n = len(op)
    code = op.x.astype(np.int8) + 2 * op.z.astype(np.int8)
    adj = np.zeros((n, n), dtype=bool)
    for i in range(n):
        both = (code[i] != 0) & (code != 0)
        adj[i] = np.any(both & (code[i] != code), axis=1)
    deg = adj.sum(axis=1)
    order = sorted(range(n), key=lambda k: -deg[k])
    colour = -np.ones(n, dtype=np.int64)
    for v in order:
        used = set(colour[adj[v]].tolist())
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    return int(colour.max() + 1) if n else 0


def _trunc_active(b) -> bool:
    """``TruncationErrorBudget.is_active`` (``utils/truncating.py:56-58``) from the fields of any budget object."""
    return any(x > 0 for x in b.per_slice_budget) and b.max_error_total > 0


def _op_budget_active(b) -> bool:
    """``OperatorBudget.is_active`` (``utils/simplify.py:73-77``)."""
    return b.max_paulis is not None or b.max_qwc_groups is not None


@dataclass
class GateTrace:
    """Per applied gate bookkeeping used by the benchmark (term-gate update count)."""

    updates: int = 0
    applied_gates: int = 0
    gate_sizes: list = field(default_factory=list)  # (slice position, observable, gate id, size after the gate), :239-242


def backpropagate(
    observables,
    slices,
    *,
    truncation_error_budget: TruncationErrorBudget | None = None,
    operator_budget: OperatorBudget | None = None,
    trace: GateTrace | None = None,
    max_updates: int | None = None,
):
    """Restatement of ``backpropagation.py:105-311`` (no ``max_seconds``; optional work bound).

    ``max_updates`` stops the loop after that many term-gate updates (used only to bound the
    CPU-baseline sample in ``bench.py``); it has no counterpart in the reference.
    """
    operator_budget = operator_budget or OperatorBudget()
    truncation_error_budget = truncation_error_budget or TruncationErrorBudget()
    single = isinstance(observables, SparsePauliOp)
    obs_in = [observables] if single else list(observables)
    num_qubits = slices[0].num_qubits
    _validate(obs_in, slices, operator_budget)

    qargs_out = qargs_tmp = [support_of(o) for o in obs_in]  # :114, :445-459
    observables_out = observables_tmp = [o.copy() for o in obs_in]  # :117
    qargs_tmp = copy.deepcopy(qargs_tmp)
    n_obs = len(obs_in)
    metadata = OBPMetadata(
        truncation_error_budget=truncation_error_budget,
        operator_budget=operator_budget,
        num_slices=len(slices),
        backpropagation_history=[],
        num_backpropagated_slices=0,
    )
    stop = False
    for slice_ in slices[::-1]:  # :146
        sm = SliceMetadata(
            slice_errors=[0.0] * n_obs,
            raw_num_paulis=[0] * n_obs,
            num_unique_paulis=[len(o) for o in observables_tmp] if operator_budget.simplify else None,
            num_duplicate_paulis=[0] * n_obs if operator_budget.simplify else None,
            num_trimmed_paulis=[0] * n_obs if operator_budget.simplify else None,
            sum_trimmed_coeffs=[0] * n_obs if operator_budget.simplify else None,
            num_truncated_paulis=[0] * n_obs,
            num_paulis=[0] * n_obs,
            sum_paulis=None,
            num_qwc_groups=None,
        )  # :148-161
        metadata.backpropagation_history.append(sm)
        observables_tmp = list(observables_tmp)
        for i in range(n_obs):  # :168
            non_trivial = False
            op_nodes = gate_order(slice_)[::-1]  # :170
            for op_idx, node in enumerate(op_nodes):
                if node.name == "barrier":  # :173
                    continue
                op_qargs = list(node.qubits)
                if not set(qargs_tmp[i]).intersection(op_qargs):  # :183
                    continue
                non_trivial = True
                if trace is not None:
                    trace.updates += len(observables_tmp[i])
                    trace.applied_gates += 1
                if node.name == "reset":  # :189-191
                    observables_tmp[i] = apply_reset_to(observables_tmp[i], op_qargs[0])
                elif node.name == "LayerError":  # :193-206
                    observables_tmp[i] = apply_ple_to(
                        observables_tmp[i], node.ple.generators, node.ple.rates, op_qargs
                    )
                    qargs_tmp[i] = sorted(set(qargs_tmp[i]) | set(op_qargs))
                else:  # :208-216
                    observables_tmp[i] = apply_op_to(
                        observables_tmp[i], _decompose_cached(node), op_qargs
                    )
                    qargs_tmp[i] = sorted(set(qargs_tmp[i]) | set(op_qargs))
                sm.raw_num_paulis[i] = len(observables_tmp[i])  # :218
                if operator_budget.simplify:  # :220-237
                    observables_tmp[i], smd = simplify(
                        observables_tmp[i], atol=operator_budget.atol, rtol=operator_budget.rtol
                    )
                    sm.num_unique_paulis[i] = smd.num_unique_paulis
                    sm.num_duplicate_paulis[i] = smd.num_duplicate_paulis
                    sm.num_trimmed_paulis[i] = smd.num_trimmed_paulis
                    sm.sum_trimmed_coeffs[i] = smd.sum_trimmed_coeffs
                if trace is not None:  # the DEBUG line of :239-242
                    trace.gate_sizes.append(
                        (len(metadata.backpropagation_history) - 1, i, len(op_nodes) - op_idx - 1, len(observables_tmp[i]))
                    )
                if max_updates is not None and trace is not None and trace.updates >= max_updates:
                    stop = True
                    break
            if stop:
                break
            if _trunc_active(truncation_error_budget) and non_trivial:  # :244-251
                prev = len(observables_tmp[i])
                sidx = metadata.num_backpropagated_slices
                budget = metadata.left_over_error_budget(i, sidx)  # :388
                observables_tmp[i], err = truncate_binary_search(
                    observables_tmp[i],
                    budget,
                    p_norm=truncation_error_budget.p_norm,
                    tol=truncation_error_budget.tol,
                )  # :391-396
                sm.slice_errors[i] = err  # :404
                sm.num_truncated_paulis[i] = prev - len(observables_tmp[i])
            sm.num_paulis[i] = len(observables_tmp[i])  # :253
        if stop:
            break
        if _op_budget_active(operator_budget):  # :266-275
            oversized = False
            if operator_budget.max_paulis is not None:
                sm.sum_paulis = _union_size(observables_tmp)
                oversized = sm.sum_paulis > operator_budget.max_paulis
            if not oversized and operator_budget.max_qwc_groups is not None:
                allp = _unique_union(observables_tmp)
                sm.num_qwc_groups = qwc_group_count(allp)
                oversized = sm.num_qwc_groups > operator_budget.max_qwc_groups
            if oversized:
                break
        observables_out = [o.copy() for o in observables_tmp]  # :279-280
        qargs_out = copy.deepcopy(qargs_tmp)
        metadata.num_backpropagated_slices += 1  # :285
    slices_out = slices[: -metadata.num_backpropagated_slices]  # :300 (incl. the [:-0] quirk)
    del qargs_out, num_qubits
    obs_out = observables_out[0] if single else observables_out
    return obs_out, slices_out, metadata


def _unique_union(observables: list[SparsePauliOp]) -> SparsePauliOp:
    """``all_paulis.simplify()`` of the coefficient-1 union (:419-425), first-appearance order."""
    x = np.concatenate([o.x for o in observables], axis=0)
    z = np.concatenate([o.z for o in observables], axis=0)
    keys = np.concatenate([np.packbits(x, axis=1), np.packbits(z, axis=1)], axis=1)
    idx, _ = unordered_unique(keys)
    return SparsePauliOp((x[idx], z[idx]), np.ones(len(idx)))


def _validate(obs, slices, operator_budget):
    """``_validate_input_options`` (:314-375) — same messages."""
    num_qubits = slices[0].num_qubits
    for i, s in enumerate(slices):
        if s.num_qubits != num_qubits:
            raise ValueError(
                "All slices must be defined on the same number of qubits. "
                f"slices[0] contains {num_qubits} qubits, but slices[{i}] contains "
                f"{s.num_qubits} qubits."
            )
    obs_qubits = obs[0].num_qubits
    for i, o in enumerate(obs):
        if o.num_qubits != obs_qubits:
            raise ValueError(
                "Input observables must all act on the same number of qubits. "
                f"observables[0] acts on {obs_qubits} qubits, but observables[{i}]"
                f" acts on {o.num_qubits} qubits."
            )
    if obs_qubits != num_qubits:
        raise ValueError(
            "The input observables must be defined on the same number of qubits as the "
            "circuit slices."
        )
    if operator_budget.max_paulis is not None and operator_budget.max_paulis < 1:
        raise ValueError("Limiting the number of Pauli terms to less than 1 does not make sense.")
    if operator_budget.max_qwc_groups is not None and operator_budget.max_qwc_groups < 1:
        raise ValueError(
            "Limiting the number of qubit-wise commmuting Pauli groups to less than 1 does not "
            "make sense."
        )
    if _op_budget_active(operator_budget):
        allp = _unique_union(obs)
        if operator_budget.max_paulis is not None and len(allp) > operator_budget.max_paulis:
            raise ValueError(
                f"You specified a maximum number of Pauli terms of {operator_budget.max_paulis}, but the "
                f"provided observables already exceed this threshold with a total of "
                f"{len(allp)} terms."
            )
        if operator_budget.max_qwc_groups is not None:
            g = qwc_group_count(allp)
            if g > operator_budget.max_qwc_groups:
                raise ValueError(
                    f"You specified a maximum number of qubit-wise commuting Pauli groups of "
                    f"{operator_budget.max_qwc_groups}, but the provided observables already exceed this threshold "
                    f"with a total of {g} terms."
                )


__all__ = [
    "GateTrace",
    "OBPMetadata",
    "OperatorBudget",
    "QuantumCircuit",
    "SimplifyMetadata",
    "SliceMetadata",
    "TruncationErrorBudget",
    "gate_matrix",
    "gate_order",
    "adjoint",
    "apply_op_to",
    "apply_ple_to",
    "apply_reset_to",
    "backpropagate",
    "compose",
    "pauli_decompose",
    "qwc_group_count",
    "simplify",
    "support_of",
    "truncate_binary_search",
    "unordered_unique",
]
