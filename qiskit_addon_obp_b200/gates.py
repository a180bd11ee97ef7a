"""Host-side gate analysis: Pauli decomposition and classification into device programs.

The reference turns every gate into a Pauli sum with ``SparsePauliOp.from_operator(Operator(op))``
(``backpropagation.py:213``) and conjugates the observable by that sum.  Here the same
decomposition is computed once per distinct gate (cached) and classified:

* **Clifford class** — the Pauli-transfer matrix ``M[q][p] = Tr(s_q U^dag s_p U) / 2^m`` is a signed
  permutation (h, s, x, y, z, sx, cx, cz, swap, iswap, ecr, rz(pi), rzz(+-pi/2), ...).  The device
  applies the permutation to the key bits in registers (``k_clifford_*``).
* **rotation class** — ``U = u0*I + u1*G`` with one Pauli generator ``G`` (rx, ry, rz, rxx, ryy,
  rzz, rzx, p, t, tdg, ...).  The device runs the branch-and-merge kernel (``k_rotate``).

* **generic** — anything else on up to 4 qubits, i.e. at most 256 Pauli components (``cp``, ``crz``,
  ``ccx``, arbitrary unitaries).  The device runs the reference's literal k*k raw-row expansion followed by the merge of
  a freshly loaded operator (``k_expand`` + ``k_index_build`` + ``k_trim``); no BASELINE configuration
  needs it, it is there for drop-in completeness.

Local Pauli code of an m-qubit Pauli: bit ``2j`` = x on the gate's qubit ``j``, bit ``2j+1`` = z.
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from .engine import OBP_OP_CLIFFORD, OBP_OP_GENERIC, OBP_OP_ROTATION, OP_DTYPE
from .ir import Instruction

#: [EXT] SparsePauliOp.atol — components of a gate with |u| <= this are dropped by from_operator
DECOMPOSE_ATOL = 1e-8
#: an entry of the Pauli-transfer matrix below this is treated as zero when classifying
MONOMIAL_TOL = 1e-14

_P1 = [
    np.eye(2, dtype=complex),  # code 0: I
    np.array([[0, 1], [1, 0]], dtype=complex),  # code 1: X
    np.array([[1, 0], [0, -1]], dtype=complex),  # code 2: Z
    np.array([[0, -1j], [1j, 0]], dtype=complex),  # code 3: Y
]
_BASIS_CACHE: dict[int, np.ndarray] = {}


def pauli_basis(m: int) -> np.ndarray:
    """``[4^m, 2^m, 2^m]`` Hermitian Pauli matrices indexed by local code (little-endian qubits)."""
    if m not in _BASIS_CACHE:
        mats = np.zeros((4**m, 2**m, 2**m), dtype=complex)
        for code in range(4**m):
            mat = np.array([[1.0 + 0j]])
            for j in range(m):  # qubit j is the j-th least significant -> kron on the left
                mat = np.kron(_P1[(code >> (2 * j)) & 3], mat)
            mats[code] = mat
        _BASIS_CACHE[m] = mats
    return _BASIS_CACHE[m]


def _label_rank(code: int, m: int) -> tuple:
    """Sort key reproducing label order I<X<Y<Z with the highest qubit most significant."""
    order = {0: 0, 1: 1, 3: 2, 2: 3}
    return tuple(order[(code >> (2 * j)) & 3] for j in range(m - 1, -1, -1))


def pauli_decompose(matrix: np.ndarray) -> tuple[list[int], np.ndarray]:
    """``U = sum_j u_j s_{code_j}``; components with ``|u_j| <= 1e-8`` dropped ([EXT] from_operator)."""
    m = int(round(np.log2(matrix.shape[0])))
    if matrix.shape != (2**m, 2**m):
        raise ValueError("gate matrix must be 2^m x 2^m")
    basis = pauli_basis(m)
    coeffs = np.einsum("kij,ji->k", basis, matrix) / 2**m  # Tr(P U) / 2^m, P Hermitian
    codes = [c for c in range(4**m) if not abs(coeffs[c]) <= DECOMPOSE_ATOL]
    codes.sort(key=lambda c: _label_rank(c, m))
    return codes, np.array([coeffs[c] for c in codes])


@dataclass
class GateProgram:
    """A classified gate: a template ``obp_op`` record plus what the driver needs to report."""

    kind: int
    n_qubits: int
    n_components: int
    template: np.ndarray  # one OP_DTYPE record with everything but the qubits filled in
    codes: np.ndarray | None = None  # generic gates: local Pauli codes ...
    coeffs: np.ndarray | None = None  # ... and their complex coefficients


def classify(matrix: np.ndarray, name: str = "gate") -> GateProgram:
    m = int(round(np.log2(matrix.shape[0])))
    codes, us = pauli_decompose(matrix)
    k = len(codes)
    if k == 0:
        raise ValueError(f"gate {name!r} has no Pauli component above tolerance")
    basis = pauli_basis(m)
    u_trunc = np.einsum("k,kij->ij", us, basis[codes])  # what the reference conjugates with
    # Pauli-transfer matrix M[q, p] = Tr(s_q U^dag s_p U) / 2^m
    conj = np.einsum("ab,pbc,cd->pad", u_trunc.conj().T, basis, u_trunc)
    ptm = np.einsum("qij,pji->qp", basis, conj).real / 2**m
    rec = np.zeros(1, dtype=OP_DTYPE)
    rec["n_qubits"] = m
    rec["n_components"] = k
    big = np.abs(ptm) > MONOMIAL_TOL
    mags = np.abs(us)
    monomial = bool(np.all(big.sum(axis=0) == 1)) and np.allclose(
        np.abs(ptm[big]), 1.0, atol=1e-9
    )
    uniform = bool(np.allclose(mags, mags[0], rtol=1e-12, atol=0.0))
    if monomial and uniform and m <= 2:
        lut, sign = 0, 0
        for p in range(4**m):
            q = int(np.argmax(np.abs(ptm[:, p])))
            lut |= q << (4 * p)
            if ptm[q, p] < 0:
                sign |= 1 << p
        group = sorted({a ^ b for a, b in itertools.product(codes, repeat=2)})
        closed = all((a ^ b) in group for a, b in itertools.product(group, repeat=2))
        if closed and len(group) <= 16:
            elems = 0
            for i, g in enumerate(x for x in group if x != 0):
                elems |= g << (4 * i)
            rec["kind"] = OBP_OP_CLIFFORD
            rec["lut"] = lut
            rec["sign_mask"] = sign
            rec["group_size"] = len(group)
            rec["group_elems"] = elems
            rec["row_scale"] = float(np.mean(mags) ** 2)
            return GateProgram(OBP_OP_CLIFFORD, m, k, rec)
    if k == 2 and codes[0] == 0 and m <= 4:
        rec["kind"] = OBP_OP_ROTATION
        rec["gen_local"] = codes[1]
        rec["u0"] = (us[0].real, us[0].imag)
        rec["u1"] = (us[1].real, us[1].imag)
        return GateProgram(OBP_OP_ROTATION, m, k, rec)
    if m <= 4:
        # anything else: the device runs the reference's literal k*k raw-row expansion + simplify
        rec["kind"] = OBP_OP_GENERIC
        return GateProgram(OBP_OP_GENERIC, m, k, rec, codes=np.array(codes, dtype=np.uint32), coeffs=np.array(us))
    raise NotImplementedError(
        f"gate {name!r} acts on {m} qubits ({k} Pauli components); the raw-row expansion kernel handles "
        "gates on up to 4 qubits"
    )


_CACHE: dict[tuple, GateProgram] = {}
_RAW_CACHE: dict[tuple, GateProgram] = {}


def raw_program_for(ins: Instruction) -> GateProgram:
    """The gate as its plain Pauli sum, for ``OperatorBudget(simplify=False)``: every gate is expanded
    into its k*k raw rows like the reference does (``backpropagation.py:210-220`` without simplify)."""
    key = (ins.name, ins.params) if ins._matrix is None else (ins.name, ins._matrix.shape, ins._matrix.tobytes())
    prog = _RAW_CACHE.get(key)
    if prog is None:
        codes, us = pauli_decompose(ins.matrix)
        if len(ins.qubits) > 4:
            raise NotImplementedError(f"gate {ins.name!r} acts on {len(ins.qubits)} qubits; at most 4 are supported")
        rec = np.zeros(1, dtype=OP_DTYPE)
        rec["kind"] = OBP_OP_GENERIC
        rec["n_qubits"] = len(ins.qubits)
        rec["n_components"] = len(codes)
        prog = GateProgram(OBP_OP_GENERIC, len(ins.qubits), len(codes), rec, np.array(codes, dtype=np.uint32), np.array(us))
        _RAW_CACHE[key] = prog
    return prog


def program_for(ins: Instruction) -> GateProgram:
    """Classified program of a circuit instruction, cached by gate identity."""
    if ins._matrix is None:
        key = (ins.name, ins.params)
    else:
        key = (ins.name, ins._matrix.shape, ins._matrix.tobytes())
    prog = _CACHE.get(key)
    if prog is None:
        prog = classify(ins.matrix, ins.name)
        _CACHE[key] = prog
    if prog.n_qubits != len(ins.qubits):
        raise ValueError(f"gate {ins.name!r} acts on {prog.n_qubits} qubits, got {len(ins.qubits)}")
    return prog
