from __future__ import annotations

import numpy as np


class PauliList:
    """Symplectic table; ``phase`` is the label phase, always 0 inside a SparsePauliOp (as in Qiskit)."""

    def __init__(self, z, x, phase=None):
        self.z = np.array(z, dtype=bool)
        self.x = np.array(x, dtype=bool)
        self.phase = np.zeros(self.x.shape[0], dtype=int) if phase is None else np.asarray(phase, dtype=int)

    @classmethod
    def from_symplectic(cls, z, x, phase=0):
        z, x = np.atleast_2d(z), np.atleast_2d(x)
        return cls(z, x, np.zeros(x.shape[0], dtype=int) + phase)

    @classmethod
    def from_labels(cls, labels):
        n = len(labels[0])
        x = np.zeros((len(labels), n), dtype=bool)
        z = np.zeros((len(labels), n), dtype=bool)
        for i, lbl in enumerate(labels):
            for k, ch in enumerate(lbl):  # big-endian: leftmost character = highest qubit
                q = n - 1 - k
                x[i, q] = ch in "XY"
                z[i, q] = ch in "ZY"
        return cls(z, x)

    def __len__(self):
        return self.x.shape[0]

    @property
    def num_qubits(self):
        return self.x.shape[1]

    def to_labels(self):
        out = []
        for xr, zr in zip(self.x, self.z):
            out.append("".join("IXZY"[int(a) + 2 * int(b)] for a, b in zip(xr[::-1], zr[::-1])))
        return out


class SparsePauliOp:
    def __init__(self, data, coeffs=None, *, ignore_pauli_phase=False, copy=True):
        if isinstance(data, str):
            data = [data]
        if not isinstance(data, PauliList):
            data = PauliList.from_labels(list(data))
        self.paulis = data
        n = len(data)
        self.coeffs = np.ones(n, dtype=complex) if coeffs is None else np.array(coeffs, dtype=complex, copy=copy)
        del ignore_pauli_phase

    @property
    def num_qubits(self):
        return self.paulis.num_qubits

    def __len__(self):
        return len(self.paulis)

    def to_list(self):
        return list(zip(self.paulis.to_labels(), self.coeffs.tolist()))


class Operator:
    def __init__(self, obj):
        self.data = np.asarray(obj.to_matrix() if hasattr(obj, "to_matrix") else obj, dtype=complex)
