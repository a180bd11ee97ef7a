"""A minimal stand-in for the parts of Qiskit that ``qiskit_addon_obp_b200.qiskit_adapter`` touches.

TEST INFRASTRUCTURE.  Qiskit cannot be installed where this repository is built (no network, no wheel),
so the adapter would otherwise never execute.  The stub reproduces the object surface the adapter reads —
``QuantumCircuit.data / find_bit / num_qubits``, ``CircuitInstruction.operation / qubits``,
``Gate.name / params`` + ``Operator(gate).data``, ``SparsePauliOp.paulis.{x,z,phase} / coeffs``,
``PauliList.from_symplectic`` — with Qiskit's conventions (little-endian gate matrices, big-endian
labels, Hermitian labels with the phase in the coefficients).  Gate matrices come from the oracle's
closed forms, i.e. not from the product's own table.
"""

from .circuit import QuantumCircuit  # noqa: F401

__version__ = "0.0-stub"
