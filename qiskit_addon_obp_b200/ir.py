"""Neutral host IR: Qiskit-free work-alikes of ``SparsePauliOp`` and ``QuantumCircuit``.

The reference (``/root/reference/qiskit_addon_obp/backpropagation.py:48-55``) takes
``qiskit.quantum_info.SparsePauliOp`` observables and ``qiskit.QuantumCircuit`` slices.  Qiskit is
not installable where this engine is built and benchmarked, so the host side works on these two
small classes, which keep Qiskit's conventions for everything the hot path can observe:

* Pauli labels are big-endian strings (leftmost character = highest qubit); ``x[:, j]`` / ``z[:, j]``
  are the symplectic bits of qubit ``j`` (little-endian columns), exactly the layout that
  ``utils/simplify.py:121-125`` of the reference reads.
* Pauli terms are Hermitian label products; any phase lives in ``coeffs`` (complex128).
* gate matrices are little-endian (qubit 0 = least-significant index bit), as in ``Operator(gate)``.

``qiskit_addon_obp_b200.qiskit_adapter`` converts real Qiskit objects to and from these classes
when Qiskit is importable; nothing else in the package touches Qiskit.
"""

from __future__ import annotations

from collections.abc import Iterable, Sequence

import numpy as np

_LABELS = "IXZY"  # index = x + 2*z


class PauliTableView:
    """Minimal stand-in for ``SparsePauliOp.paulis`` (``.x``, ``.z``, labels, iteration)."""

    def __init__(self, x: np.ndarray, z: np.ndarray):
        self.x = x
        self.z = z

    def __len__(self) -> int:
        return self.x.shape[0]

    def to_labels(self) -> list[str]:
        code = self.x.astype(np.uint8) + 2 * self.z.astype(np.uint8)
        lut = np.frombuffer(_LABELS.encode(), dtype="S1")
        chars = lut[code[:, ::-1]]
        return [b"".join(row).decode() for row in chars]

    def __iter__(self):
        return iter(self.to_labels())


class SparsePauliOp:
    """A sum of Pauli strings with complex coefficients (Qiskit-compatible subset)."""

    atol = 1e-8
    rtol = 1e-5

    def __init__(self, data, coeffs=None, *, num_qubits: int | None = None):
        if isinstance(data, SparsePauliOp):
            x, z = data.x.copy(), data.z.copy()
            if coeffs is None:
                coeffs = data.coeffs.copy()
        elif isinstance(data, str):
            x, z = _labels_to_xz([data])
        elif isinstance(data, tuple) and len(data) == 2 and isinstance(data[0], np.ndarray):
            x = np.ascontiguousarray(data[0], dtype=bool)
            z = np.ascontiguousarray(data[1], dtype=bool)
        else:
            labels = list(data)
            if len(labels) == 0:
                if num_qubits is None:
                    raise ValueError("num_qubits is required for an empty operator")
                x = np.zeros((0, num_qubits), dtype=bool)
                z = np.zeros((0, num_qubits), dtype=bool)
            else:
                x, z = _labels_to_xz(labels)
        if x.ndim != 2 or x.shape != z.shape:
            raise ValueError("x and z must be 2-D arrays of the same shape")
        if coeffs is None:
            coeffs = np.ones(x.shape[0], dtype=complex)
        coeffs = np.asarray(coeffs, dtype=complex).reshape(-1).copy()
        if coeffs.shape[0] != x.shape[0]:
            raise ValueError("number of coefficients does not match the number of Pauli terms")
        self._x = x
        self._z = z
        self._packed = None
        self._num_qubits = x.shape[1]
        self.coeffs = coeffs

    @classmethod
    def from_packed(cls, xw: np.ndarray, zw: np.ndarray, coeffs: np.ndarray, num_qubits: int) -> SparsePauliOp:
        """Wrap packed symplectic words (``uint64 [n, W]``, the engine's own format) without unpacking.

        The boolean ``x`` / ``z`` arrays of the Qiskit layout are materialised on first access.
        """
        self = cls.__new__(cls)
        self._x = self._z = None
        self._packed = (np.ascontiguousarray(xw, dtype="<u8"), np.ascontiguousarray(zw, dtype="<u8"))
        self._num_qubits = int(num_qubits)
        self.coeffs = np.asarray(coeffs, dtype=complex).reshape(-1)
        if self._packed[0].shape != self._packed[1].shape or self._packed[0].shape[0] != self.coeffs.shape[0]:
            raise ValueError("packed words and coefficients do not match")
        return self

    def _materialise(self) -> None:
        if self._x is None:
            self._x, self._z = unpack_xz(self._packed[0], self._packed[1], self._num_qubits)
        # the boolean arrays may now be mutated in place by the caller: they become the only truth
        self._packed = None

    @property
    def x(self) -> np.ndarray:
        self._materialise()
        return self._x

    @x.setter
    def x(self, value) -> None:
        self._materialise()
        self._x = np.ascontiguousarray(value, dtype=bool)

    @property
    def z(self) -> np.ndarray:
        self._materialise()
        return self._z

    @z.setter
    def z(self, value) -> None:
        self._materialise()
        self._z = np.ascontiguousarray(value, dtype=bool)

    def packed(self) -> tuple[np.ndarray, np.ndarray]:
        """``(x_words, z_words)`` as ``uint64 [n, W]``, qubit j -> bit j%64 of word j//64."""
        if self._packed is not None:
            return self._packed
        return pack_xz(self._x, self._z)

    # -- construction helpers -------------------------------------------------------------
    @classmethod
    def from_list(cls, obj: Iterable[tuple[str, complex]], *, num_qubits: int | None = None):
        obj = list(obj)
        if not obj:
            return cls([], num_qubits=num_qubits)
        return cls([lbl for lbl, _ in obj], [c for _, c in obj])

    @classmethod
    def from_sparse_list(cls, obj: Iterable[tuple[str, Sequence[int], complex]], num_qubits: int):
        obj = list(obj)
        x = np.zeros((len(obj), num_qubits), dtype=bool)
        z = np.zeros((len(obj), num_qubits), dtype=bool)
        coeffs = np.zeros(len(obj), dtype=complex)
        for i, (paulis, qubits, c) in enumerate(obj):
            for p, q in zip(paulis, qubits):
                k = _LABELS.index(p)
                x[i, q] ^= bool(k & 1)
                z[i, q] ^= bool(k & 2)
            coeffs[i] = c
        return cls((x, z), coeffs)

    # -- Qiskit-like surface --------------------------------------------------------------
    @property
    def num_qubits(self) -> int:
        return self._num_qubits

    @property
    def paulis(self) -> PauliTableView:
        return PauliTableView(self.x, self.z)

    @property
    def size(self) -> int:
        return self.coeffs.shape[0]

    def __len__(self) -> int:
        return self.coeffs.shape[0]

    def copy(self) -> SparsePauliOp:
        if self._packed is not None:
            return SparsePauliOp.from_packed(
                self._packed[0].copy(), self._packed[1].copy(), self.coeffs.copy(), self._num_qubits
            )
        return SparsePauliOp((self._x.copy(), self._z.copy()), self.coeffs.copy())

    def to_list(self) -> list[tuple[str, complex]]:
        return list(zip(self.paulis.to_labels(), self.coeffs.tolist()))

    def label_iter(self):
        return iter(self.to_list())

    def __add__(self, other: SparsePauliOp) -> SparsePauliOp:
        return SparsePauliOp(
            (np.vstack([self.x, other.x]), np.vstack([self.z, other.z])),
            np.concatenate([self.coeffs, other.coeffs]),
        )

    def __neg__(self) -> SparsePauliOp:
        return SparsePauliOp((self.x.copy(), self.z.copy()), -self.coeffs)

    def __sub__(self, other: SparsePauliOp) -> SparsePauliOp:
        return self + (-other)

    def __mul__(self, s: complex) -> SparsePauliOp:
        return SparsePauliOp((self.x.copy(), self.z.copy()), self.coeffs * s)

    __rmul__ = __mul__

    def __eq__(self, other) -> bool:
        """Entry-wise (order-sensitive) comparison, like ``SparsePauliOp.__eq__`` in Qiskit."""
        if not isinstance(other, SparsePauliOp):
            return NotImplemented
        return (
            self.x.shape == other.x.shape
            and np.array_equal(self.x, other.x)
            and np.array_equal(self.z, other.z)
            and np.allclose(self.coeffs, other.coeffs, atol=self.atol, rtol=self.rtol)
        )

    __hash__ = None  # type: ignore[assignment]

    def as_dict(self) -> dict[str, complex]:
        """``{label: summed coefficient}`` — the order-free view used by the parity tests."""
        out: dict[str, complex] = {}
        for lbl, c in self.to_list():
            out[lbl] = out.get(lbl, 0j) + c
        return out

    def equiv(self, other: SparsePauliOp, atol: float = 1e-8) -> bool:
        """Order-insensitive equality of the represented operators."""
        a, b = self.as_dict(), other.as_dict()
        return all(abs(a.get(k, 0j) - b.get(k, 0j)) <= atol for k in set(a) | set(b))

    def sorted_by_key(self) -> SparsePauliOp:
        """Rows in canonical order (ascending packed key); used to compare engine and oracle."""
        if len(self) == 0:
            return self.copy()
        xw, zw = self.packed()
        cols = [a[:, w] for w in range(xw.shape[1]) for a in (zw, xw)]
        order = np.lexsort(cols)  # last column is the primary key → highest word's x first
        return SparsePauliOp((self.x[order], self.z[order]), self.coeffs[order])

    def to_matrix(self) -> np.ndarray:
        n = self.num_qubits
        if n > 12:
            raise ValueError("to_matrix is meant for small test operators")
        mats = {
            0: np.eye(2, dtype=complex),
            1: np.array([[0, 1], [1, 0]], dtype=complex),
            2: np.array([[1, 0], [0, -1]], dtype=complex),
            3: np.array([[0, -1j], [1j, 0]], dtype=complex),
        }
        out = np.zeros((2**n, 2**n), dtype=complex)
        for xr, zr, c in zip(self.x, self.z, self.coeffs):
            m = np.array([[1.0 + 0j]])
            for q in range(n - 1, -1, -1):
                m = np.kron(m, mats[int(xr[q]) + 2 * int(zr[q])])
            out += c * m
        return out

    def __repr__(self) -> str:
        lbls = self.paulis.to_labels()
        if len(lbls) > 8:
            shown = ", ".join(repr(s) for s in lbls[:8]) + ", ..."
        else:
            shown = ", ".join(repr(s) for s in lbls)
        return f"SparsePauliOp([{shown}], coeffs={self.coeffs[:8]!r}{'...' if len(lbls) > 8 else ''})"


def _labels_to_xz(labels: Sequence[str]) -> tuple[np.ndarray, np.ndarray]:
    n = len(labels[0])
    raw = np.frombuffer("".join(labels).encode(), dtype=np.uint8)
    if raw.size != n * len(labels):
        raise ValueError("all Pauli labels must have the same length")
    raw = raw.reshape(len(labels), n)[:, ::-1]
    isx, isy, isz, isi = raw == ord("X"), raw == ord("Y"), raw == ord("Z"), raw == ord("I")
    if not np.all(isx | isy | isz | isi):
        raise ValueError("Pauli labels may only contain I, X, Y, Z")
    return np.ascontiguousarray(isx | isy), np.ascontiguousarray(isz | isy)


def pack_xz(x: np.ndarray, z: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Pack bool ``[n, N]`` symplectic arrays into ``uint64 [n, W]`` words, qubit j → bit j%64 of word j//64."""
    n, nq = x.shape
    w = max(1, (nq + 63) // 64)
    pad = w * 64 - nq

    def _pack(a):
        if pad:
            a = np.concatenate([a, np.zeros((n, pad), dtype=bool)], axis=1)
        b = np.packbits(a, axis=1, bitorder="little")
        return np.ascontiguousarray(b).view("<u8").reshape(n, w)

    return _pack(x), _pack(z)


def unpack_xz(xw: np.ndarray, zw: np.ndarray, num_qubits: int) -> tuple[np.ndarray, np.ndarray]:
    """Inverse of :func:`pack_xz`."""
    n = xw.shape[0]
    if n == 0:
        return np.zeros((0, num_qubits), dtype=bool), np.zeros((0, num_qubits), dtype=bool)

    def _unpack(a):
        b = np.ascontiguousarray(a, dtype="<u8").view(np.uint8).reshape(n, -1)
        bits = np.unpackbits(b, axis=1, count=num_qubits, bitorder="little")
        return bits.view(np.bool_)  # unpackbits yields 0/1 bytes: reinterpret, do not copy

    return _unpack(xw), _unpack(zw)


# ---------------------------------------------------------------------------------------------
# circuits
# ---------------------------------------------------------------------------------------------

_SQ2 = 1.0 / np.sqrt(2.0)


def _rot(label: str, theta: float) -> np.ndarray:
    """exp(-i theta/2 P) for a little-endian Pauli label (label[0] acts on the gate's qubit 0)."""
    p = np.array([[1.0 + 0j]])
    for ch in label:
        p = np.kron(_P1[ch], p)
    return np.cos(theta / 2) * np.eye(p.shape[0], dtype=complex) - 1j * np.sin(theta / 2) * p


_P1 = {
    "I": np.eye(2, dtype=complex),
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
}


def _fixed(name: str) -> np.ndarray:
    if name == "id":
        return np.eye(2, dtype=complex)
    if name in ("x", "y", "z"):
        return _P1[name.upper()].copy()
    if name == "h":
        return np.array([[1, 1], [1, -1]], dtype=complex) * _SQ2
    if name == "s":
        return np.diag([1, 1j]).astype(complex)
    if name == "sdg":
        return np.diag([1, -1j]).astype(complex)
    if name == "t":
        return np.diag([1, np.exp(1j * np.pi / 4)]).astype(complex)
    if name == "tdg":
        return np.diag([1, np.exp(-1j * np.pi / 4)]).astype(complex)
    if name == "sx":
        return np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex) / 2
    if name == "sxdg":
        return np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=complex) / 2
    # two-qubit gates, little-endian: index = q0 + 2*q1 with q0 the first qubit argument
    if name == "cx":  # control = first argument
        return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)
    if name == "cz":
        return np.diag([1, 1, 1, -1]).astype(complex)
    if name == "cy":
        return np.array([[1, 0, 0, 0], [0, 0, 0, -1j], [0, 0, 1, 0], [0, 1j, 0, 0]], dtype=complex)
    if name == "swap":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if name == "iswap":
        return np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if name == "ecr":
        return (
            np.array([[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]], dtype=complex)
            * _SQ2
        )
    raise KeyError(name)


_FIXED_1Q = ("id", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg")
_FIXED_2Q = ("cx", "cz", "cy", "swap", "iswap", "ecr")
_ROT = {"rx": "X", "ry": "Y", "rz": "Z", "rxx": "XX", "ryy": "YY", "rzz": "ZZ", "rzx": "XZ"}
# Qiskit's RZXGate(theta) = exp(-i theta/2 X⊗Z) with Z on the first argument, X on the second:
# little-endian label (first qubit first) is "ZX".
_ROT["rzx"] = "ZX"


class Instruction:
    """One circuit operation: a named gate on ``qubits`` with ``params``, or an explicit unitary.

    ``name`` follows Qiskit's gate names (``"rx"``, ``"cx"``, ``"barrier"``, ``"reset"``,
    ``"LayerError"``, ``"unitary"``...).  ``matrix`` is the little-endian unitary of the gate,
    i.e. what ``Operator(op_node.op).data`` is in ``backpropagation.py:213`` of the reference.

    Immutable: compiled slice programs are cached against the identity of their instructions
    (``backpropagation._slice_program``), so changing a gate means putting a new ``Instruction`` into
    ``QuantumCircuit.data``.
    """

    __slots__ = ("name", "qubits", "params", "_matrix", "ple")

    def __init__(self, name, qubits, params=(), matrix=None, ple=None):
        set_ = object.__setattr__
        set_(self, "name", name)
        set_(self, "qubits", tuple(int(q) for q in qubits))
        set_(self, "params", tuple(float(p) for p in params))
        if matrix is not None:
            matrix = np.array(matrix, dtype=complex)  # private copy, frozen
            matrix.setflags(write=False)
        set_(self, "_matrix", matrix)
        set_(self, "ple", ple)

    def __setattr__(self, key, value):
        raise AttributeError("Instruction is immutable: replace the entry of QuantumCircuit.data instead")

    __delattr__ = __setattr__

    @property
    def matrix(self) -> np.ndarray:
        if self._matrix is not None:
            return self._matrix
        if self.name in _ROT:
            return _rot(_ROT[self.name], self.params[0])
        if self.name == "p":
            return np.diag([1, np.exp(1j * self.params[0])]).astype(complex)
        if self.name == "cp":
            return np.diag([1, 1, 1, np.exp(1j * self.params[0])]).astype(complex)
        if self.name == "crz":
            t = self.params[0]
            return np.diag([1, np.exp(-1j * t / 2), 1, np.exp(1j * t / 2)]).astype(complex)
        return _fixed(self.name)

    def __repr__(self) -> str:
        return f"Instruction({self.name!r}, {self.qubits}, {self.params})"


class PauliLindbladError:
    """Work-alike of ``qiskit_ibm_runtime...PauliLindbladError``: Pauli generators and rates."""

    def __init__(self, generators, rates):
        if isinstance(generators, SparsePauliOp):
            generators = generators.paulis.to_labels()
        self.generators = [str(g) for g in generators]
        self.rates = np.asarray(rates, dtype=float)
        if len(self.generators) != len(self.rates):
            raise ValueError("generators and rates must have the same length")
        self.num_qubits = len(self.generators[0]) if self.generators else 0


class QuantumCircuit:
    """A flat list of instructions on ``num_qubits`` qubits (Qiskit-compatible builder subset)."""

    def __init__(self, num_qubits: int):
        self.num_qubits = int(num_qubits)
        self.data: list[Instruction] = []

    def _add(self, name, qubits, params=(), matrix=None, ple=None):
        qubits = [qubits] if isinstance(qubits, (int, np.integer)) else list(qubits)
        for q in qubits:
            if not 0 <= q < self.num_qubits:
                raise ValueError(f"qubit index {q} out of range for {self.num_qubits} qubits")
        if len(set(qubits)) != len(qubits):
            raise ValueError("duplicate qubit arguments")
        self.data.append(Instruction(name, qubits, params, matrix, ple))
        return self

    def _bcast1(self, name, qubit, params=()):
        if isinstance(qubit, (int, np.integer)):
            return self._add(name, [qubit], params)
        for q in qubit:
            self._add(name, [q], params)
        return self

    def unitary(self, matrix, qubits, label: str = "unitary"):
        return self._add(label, qubits, (), matrix)

    def barrier(self, *qubits):
        qs = list(qubits) if qubits else list(range(self.num_qubits))
        self.data.append(Instruction("barrier", qs))
        return self

    def reset(self, qubit):
        return self._bcast1("reset", qubit)

    def append(self, instruction, qargs):
        """Append a :class:`PauliLindbladErrorInstruction`-like object or an :class:`Instruction`."""
        ple = getattr(instruction, "ple", None)
        if ple is not None:
            return self._add("LayerError", qargs, (), None, ple)
        if isinstance(instruction, Instruction):
            return self._add(
                instruction.name, qargs, instruction.params, instruction._matrix, instruction.ple
            )
        raise TypeError(f"cannot append {type(instruction).__name__}")

    def compose(self, other: QuantumCircuit, inplace: bool = False):
        tgt = self if inplace else self.copy()
        tgt.data.extend(other.data)
        return tgt

    def copy(self) -> QuantumCircuit:
        qc = QuantumCircuit(self.num_qubits)
        qc.data = list(self.data)
        return qc

    def __len__(self) -> int:
        return len(self.data)

    def __repr__(self) -> str:
        return f"QuantumCircuit({self.num_qubits}, {len(self.data)} ops)"


def _make_1q(name):
    def f(self, qubit):
        return self._bcast1(name, qubit)

    f.__name__ = name
    return f


def _make_2q(name):
    def f(self, q0, q1):
        return self._add(name, [q0, q1])

    f.__name__ = name
    return f


def _make_rot(name):
    nq = len(_ROT[name])

    if nq == 1:

        def f(self, theta, qubit):
            return self._bcast1(name, qubit, (theta,))
    else:

        def f(self, theta, q0, q1):
            return self._add(name, [q0, q1], (theta,))

    f.__name__ = name
    return f


for _n in _FIXED_1Q:
    setattr(QuantumCircuit, _n, _make_1q(_n))
for _n in _FIXED_2Q:
    setattr(QuantumCircuit, _n, _make_2q(_n))
for _n in _ROT:
    setattr(QuantumCircuit, _n, _make_rot(_n))
QuantumCircuit.p = lambda self, theta, qubit: self._bcast1("p", qubit, (theta,))  # type: ignore[attr-defined]
QuantumCircuit.cp = lambda self, theta, q0, q1: self._add("cp", [q0, q1], (theta,))  # type: ignore[attr-defined]
QuantumCircuit.crz = lambda self, theta, q0, q1: self._add("crz", [q0, q1], (theta,))  # type: ignore[attr-defined]


class PauliLindbladErrorInstruction:
    """Carrier for a :class:`PauliLindbladError` inside a circuit (``utils/noise.py:23-75``)."""

    def __init__(self, ple: PauliLindbladError):
        self.ple = ple
        self.name = "LayerError"
        self.num_qubits = ple.num_qubits


# ---------------------------------------------------------------------------------------------
# gate order inside a slice and slicing helpers
# ---------------------------------------------------------------------------------------------


def topological_op_order(circuit: QuantumCircuit) -> list[Instruction]:
    """Instructions in the order of ``circuit_to_dag(c).topological_op_nodes()``.

    Qiskit sorts with a lexicographical topological sort whose key is the zero-padded qubit
    indices of each node ([EXT] ``DAGCircuit.topological_nodes``).  The reference reverses this
    list (``backpropagation.py:170``).
    """
    import heapq

    n_ops = len(circuit.data)
    last_on_qubit: dict[int, int] = {}
    succ: list[list[int]] = [[] for _ in range(n_ops)]
    indeg = [0] * n_ops
    for i, ins in enumerate(circuit.data):
        preds = {last_on_qubit[q] for q in ins.qubits if q in last_on_qubit}
        for p in preds:
            succ[p].append(i)
        indeg[i] = len(preds)
        for q in ins.qubits:
            last_on_qubit[q] = i
    # Qiskit's key is the zero-padded qubit indices joined by commas; tuples of ints order the same way
    key = [ins.qubits for ins in circuit.data]
    heap = [(key[i], i) for i in range(n_ops) if indeg[i] == 0]
    heapq.heapify(heap)
    out = []
    while heap:
        _, i = heapq.heappop(heap)
        out.append(circuit.data[i])
        for j in succ[i]:
            indeg[j] -= 1
            if indeg[j] == 0:
                heapq.heappush(heap, (key[j], j))
    return out


def slice_by_barriers(circuit: QuantumCircuit) -> list[QuantumCircuit]:
    """Split at barriers (work-alike of ``qiskit_addon_utils.slicing.slice_by_barriers``)."""
    slices = [QuantumCircuit(circuit.num_qubits)]
    for ins in circuit.data:
        if ins.name == "barrier":
            slices.append(QuantumCircuit(circuit.num_qubits))
        else:
            slices[-1].data.append(ins)
    return [s for s in slices if len(s.data)]


def slice_by_depth(circuit: QuantumCircuit, max_slice_depth: int = 1) -> list[QuantumCircuit]:
    """Greedy ASAP layering into slices of depth ``max_slice_depth`` (``slice_by_depth`` work-alike)."""
    level: dict[int, int] = {}
    layers: list[list[Instruction]] = []
    for ins in circuit.data:
        if ins.name == "barrier":
            continue
        d = max((level.get(q, 0) for q in ins.qubits), default=0)
        while len(layers) <= d:
            layers.append([])
        layers[d].append(ins)
        for q in ins.qubits:
            level[q] = d + 1
    out = []
    for i in range(0, len(layers), max_slice_depth):
        qc = QuantumCircuit(circuit.num_qubits)
        for lay in layers[i : i + max_slice_depth]:
            qc.data.extend(lay)
        out.append(qc)
    return out
