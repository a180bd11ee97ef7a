"""Conversion between Qiskit objects and the neutral host IR (imported lazily, optional).

Qiskit is not installed where this engine is built, so this module is the only place that touches
it and it degrades to a pass-through when the inputs are already IR objects.  With Qiskit present,
``backpropagate`` accepts ``qiskit.quantum_info.SparsePauliOp`` / ``qiskit.QuantumCircuit`` and
returns ``SparsePauliOp`` objects of the caller's type, as the reference does
(``backpropagation.py:48-55``).  The remaining slices are returned as a slice of the caller's own
list (``:300``), never converted.
"""

from __future__ import annotations

import numpy as np

from .ir import Instruction, PauliLindbladError, QuantumCircuit, SparsePauliOp


def _is_ir_op(o) -> bool:
    return isinstance(o, SparsePauliOp)


def _op_from_qiskit(op) -> SparsePauliOp:
    # SparsePauliOp keeps Hermitian labels with the phase folded into coeffs (ignore_pauli_phase)
    x = np.asarray(op.paulis.x, dtype=bool)
    z = np.asarray(op.paulis.z, dtype=bool)
    phase = np.asarray(op.paulis.phase)
    coeffs = np.asarray(op.coeffs)
    if coeffs.dtype == object:
        raise TypeError("parameterized (object-dtype) coefficients are not supported on the device path")
    if np.any(phase != 0):
        coeffs = coeffs * (-1j) ** phase
    return SparsePauliOp((x.copy(), z.copy()), coeffs.astype(complex))


def _circuit_from_qiskit(qc) -> QuantumCircuit:
    from qiskit.quantum_info import Operator  # noqa: PLC0415

    out = QuantumCircuit(qc.num_qubits)
    for inst in qc.data:
        op = inst.operation
        qubits = [qc.find_bit(q).index for q in inst.qubits]
        if op.name == "barrier":
            out.data.append(Instruction("barrier", qubits))
        elif op.name == "reset":
            out.data.append(Instruction("reset", qubits))
        elif op.name == "LayerError":
            ple = getattr(op, "ple", None)
            ir_ple = None
            if ple is not None:
                ir_ple = PauliLindbladError(ple.generators.to_labels(), np.asarray(ple.rates, dtype=float))
            out.data.append(Instruction("LayerError", qubits, (), None, ir_ple))
        else:
            # the unitary the reference decomposes: Operator(op_node.op) (backpropagation.py:213)
            params = tuple(float(p) for p in getattr(op, "params", ()) if isinstance(p, (int, float)))
            out.data.append(Instruction(op.name, qubits, params, np.asarray(Operator(op).data)))
    return out


def to_ir(observables, slices):
    """Return ``(observables_ir, slices_ir, convert_back)``."""
    single = not isinstance(observables, (list, tuple))
    obs_list = [observables] if single else list(observables)
    if all(_is_ir_op(o) for o in obs_list) and all(isinstance(s, QuantumCircuit) for s in slices):
        return observables, list(slices), (lambda o: o)
    try:
        from qiskit.quantum_info import SparsePauliOp as QSparsePauliOp  # noqa: PLC0415
        from qiskit.quantum_info import PauliList  # noqa: PLC0415
    except ImportError as exc:  # pragma: no cover - Qiskit absent in the build container
        raise TypeError(
            "observables must be SparsePauliOp objects and slices QuantumCircuit objects "
            "(qiskit_addon_obp_b200.ir or, when Qiskit is installed, Qiskit's own)"
        ) from exc
    obs_ir = [o if _is_ir_op(o) else _op_from_qiskit(o) for o in obs_list]
    slices_ir = [s if isinstance(s, QuantumCircuit) else _circuit_from_qiskit(s) for s in slices]

    def back(o: SparsePauliOp):
        return QSparsePauliOp(
            PauliList.from_symplectic(o.z, o.x), o.coeffs, ignore_pauli_phase=True, copy=False
        )

    return (obs_ir[0] if single else obs_ir), slices_ir, back
