"""Operator-level helpers with the reference's names and argument meaning, on the device engine.

``apply_op_to`` / ``apply_reset_to`` / ``apply_ple_to`` / ``reduce_op`` / ``to_global_op`` mirror
``/root/reference/qiskit_addon_obp/utils/operations.py``.  The engine always works at global
width, so "lean operator + qargs" pairs are converted at the boundary of each call.
Unlike the reference, the device path merges duplicates and trims zeros as part of every gate
(the reference's loop calls ``simplify`` right after each of these functions,
``backpropagation.py:220-225``), so the returned operators are already simplified.
"""

from __future__ import annotations

import numpy as np

from ..ir import Instruction, PauliLindbladError, SparsePauliOp


def _validate_qargs(op: SparsePauliOp, qargs: list[int]) -> None:
    if op.num_qubits != len(set(qargs)):
        raise ValueError(
            f"The number of qubits in the operator ({op.num_qubits}) does not match "
            f"the number of qargs ({len(set(qargs))})."
        )


def _expand(op: SparsePauliOp, op_qargs: list[int], other_qargs: list[int]):
    """Grow ``op`` to the sorted union of qargs (``utils/operations.py:23-51``)."""
    if set(other_qargs) <= set(op_qargs):
        return op, list(op_qargs), [op_qargs.index(q) for q in other_qargs]
    out_qargs = sorted(set(op_qargs) | set(other_qargs))
    x = np.zeros((len(op), len(out_qargs)), dtype=bool)
    z = np.zeros((len(op), len(out_qargs)), dtype=bool)
    cols = [out_qargs.index(q) for q in op_qargs]
    x[:, cols] = op.x
    z[:, cols] = op.z
    return SparsePauliOp((x, z), op.coeffs.copy()), out_qargs, [out_qargs.index(q) for q in other_qargs]


def apply_op_to(op1, op1_qargs, op2, op2_qargs, *, apply_as_transform: bool = False, atol: float | None = None):
    """``op2^dag @ op1 @ op2`` for a gate ``op2`` given as an :class:`Instruction` or a unitary matrix.

    Only ``apply_as_transform=True`` (the mode ``backpropagate`` uses, ``backpropagation.py:215``)
    is implemented on the device.
    """
    from ..engine import OBP_FLAG_EXACT_COUNTERS, OBP_OP_GENERIC, TermTable
    from ..gates import classify, program_for

    if not apply_as_transform:
        raise NotImplementedError("only apply_as_transform=True runs on the device path")
    if isinstance(op2, Instruction):
        prog, m = program_for(op2), len(op2.qubits)
    else:
        mat = np.asarray(op2, dtype=complex)
        prog, m = classify(mat), int(round(np.log2(mat.shape[0])))
    _validate_qargs(op1, op1_qargs)
    if m != len(set(op2_qargs)):
        raise ValueError(
            f"The number of qubits in the operator ({m}) does not match "
            f"the number of qargs ({len(set(op2_qargs))})."
        )
    expanded, out_qargs, pos = _expand(op1, list(op1_qargs), list(op2_qargs))
    tol = SparsePauliOp.atol if atol is None else atol
    k2 = prog.n_components**2
    with TermTable.from_operator(expanded, capacity=max(1024, 2 * len(expanded) * (k2 + 1))) as table:
        if prog.kind == OBP_OP_GENERIC:
            table.apply_generic(pos, prog.codes, prog.coeffs, tol)
        else:
            rec = prog.template.copy()
            rec["qubit"][0, :m] = pos
            rec["flags"] = OBP_FLAG_EXACT_COUNTERS
            table.apply_ops(rec, tol)
        return table.download(), out_qargs


def to_global_op(op: SparsePauliOp, qargs: list[int], n_qubits: int) -> SparsePauliOp:
    """Pad a lean operator with identities (``utils/operations.py:112-136``)."""
    min_q, max_q = min(qargs), max(qargs)
    if min_q < 0:
        raise ValueError(f"qargs may not contain a negative qubit ID. Found: ({min_q}).")
    if n_qubits <= max_q:
        raise ValueError(
            f"qargs contains qubit ID ({max_q}), but the global system contains only ({n_qubits}) qubits."
        )
    x = np.zeros((len(op), n_qubits), dtype=bool)
    z = np.zeros((len(op), n_qubits), dtype=bool)
    x[:, qargs] = op.x
    z[:, qargs] = op.z
    return SparsePauliOp((x, z), op.coeffs.copy())


def reduce_op(global_op: SparsePauliOp) -> tuple[SparsePauliOp, list[int]]:
    """Drop all-identity qubits (``utils/operations.py:147-191``)."""
    qargs = [int(q) for q in np.logical_or(global_op.x, global_op.z).any(axis=0).nonzero()[0]]
    if qargs == []:
        raise ValueError("Input operator may not be the identity operator.")
    return SparsePauliOp((global_op.x[:, qargs], global_op.z[:, qargs]), global_op.coeffs.copy()), qargs


def apply_reset_to(op: SparsePauliOp, qubit_id: int, inplace: bool = False, atol: float | None = None):
    """``<0|op|0>`` on one qubit (``utils/operations.py:194-225``), simplified."""
    from ..engine import TermTable

    del inplace
    with TermTable.from_operator(op) as table:
        table.apply_reset(qubit_id, SparsePauliOp.atol if atol is None else atol)
        return table.download()


def apply_ple_to(op, op_qargs, ple: PauliLindbladError, ple_qargs, atol: float | None = None):
    """Pauli-Lindblad damping (``utils/operations.py:228-275``)."""
    from ..engine import TermTable

    _validate_qargs(op, op_qargs)
    expanded, out_qargs, pos = _expand(op, list(op_qargs), list(ple_qargs))
    gens = SparsePauliOp(list(ple.generators))
    gx = np.zeros((len(gens), expanded.num_qubits), dtype=bool)
    gz = np.zeros((len(gens), expanded.num_qubits), dtype=bool)
    gx[:, pos] = gens.x
    gz[:, pos] = gens.z
    with TermTable.from_operator(expanded) as table:
        table.apply_damping(gx, gz, np.exp(-2 * np.asarray(ple.rates, dtype=float)), SparsePauliOp.atol if atol is None else atol)
        return table.download(), out_qargs
