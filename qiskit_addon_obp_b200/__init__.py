"""B200-native operator backpropagation behind the ``qiskit_addon_obp`` API.

``backpropagate`` has the signature and return value of the reference's only public entry point
(``/root/reference/qiskit_addon_obp/__init__.py:16-20``, ``backpropagation.py:48-55``); the
per-gate compose/simplify and per-slice truncate work runs as sm_100a CUDA kernels reached through
the C-ABI in ``include/obp_b200.h``.
"""

from .ir import (
    PauliLindbladError,
    PauliLindbladErrorInstruction,
    QuantumCircuit,
    SparsePauliOp,
    slice_by_barriers,
    slice_by_depth,
)


def backpropagate(*args, **kwargs):
    """See :func:`qiskit_addon_obp_b200.backpropagation.backpropagate`."""
    from .backpropagation import backpropagate as _bp

    return _bp(*args, **kwargs)


__all__ = [
    "PauliLindbladError",
    "PauliLindbladErrorInstruction",
    "QuantumCircuit",
    "SparsePauliOp",
    "backpropagate",
    "slice_by_barriers",
    "slice_by_depth",
]
