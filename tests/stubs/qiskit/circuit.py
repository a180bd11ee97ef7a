from __future__ import annotations

import numpy as np


class Qubit:
    __slots__ = ()


class _BitLocation:
    def __init__(self, index):
        self.index = index
        self.registers = []


class Gate:
    def __init__(self, name, num_qubits, params=(), matrix=None):
        self.name = name
        self.num_qubits = num_qubits
        self.params = list(params)
        self._matrix = matrix

    def to_matrix(self):
        if self._matrix is not None:
            return np.asarray(self._matrix, dtype=complex)
        from oracle.obp_oracle import gate_matrix

        return gate_matrix(self.name, tuple(self.params))


class Barrier(Gate):
    def __init__(self, n):
        super().__init__("barrier", n)


class Reset(Gate):
    def __init__(self):
        super().__init__("reset", 1)


class CircuitInstruction:
    def __init__(self, operation, qubits):
        self.operation = operation
        self.qubits = tuple(qubits)
        self.clbits = ()


class QuantumCircuit:
    def __init__(self, num_qubits):
        self.num_qubits = num_qubits
        self.qubits = [Qubit() for _ in range(num_qubits)]
        self._index = {id(q): k for k, q in enumerate(self.qubits)}
        self.data: list[CircuitInstruction] = []

    def find_bit(self, bit):
        return _BitLocation(self._index[id(bit)])

    def append(self, operation, qargs):
        self.data.append(CircuitInstruction(operation, [self.qubits[q] if isinstance(q, int) else q for q in qargs]))
        return self

    def _g(self, name, qubits, params=()):
        return self.append(Gate(name, len(qubits), params), list(qubits))

    def barrier(self):
        return self.append(Barrier(self.num_qubits), list(range(self.num_qubits)))

    def reset(self, q):
        return self.append(Reset(), [q])

    def unitary(self, matrix, qubits, label="unitary"):
        return self.append(Gate(label, len(qubits), (), np.asarray(matrix, dtype=complex)), list(qubits))


def _add_builders():
    for name in ("id", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg"):
        setattr(QuantumCircuit, name, (lambda n: lambda self, q: self._g(n, [q]))(name))
    for name in ("cx", "cz", "cy", "swap", "iswap", "ecr"):
        setattr(QuantumCircuit, name, (lambda n: lambda self, a, b: self._g(n, [a, b]))(name))
    for name in ("rx", "ry", "rz", "p"):
        setattr(QuantumCircuit, name, (lambda n: lambda self, t, q: self._g(n, [q], (t,)))(name))
    for name in ("rxx", "ryy", "rzz", "rzx", "cp", "crz"):
        setattr(QuantumCircuit, name, (lambda n: lambda self, t, a, b: self._g(n, [a, b], (t,)))(name))


_add_builders()
