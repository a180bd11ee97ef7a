"""Operator budget (host side) and the device-backed ``simplify``.

``OperatorBudget`` / ``SimplifyMetadata`` mirror ``utils/simplify.py:26-77`` of the reference.
``simplify`` has the reference's signature and counter semantics (:80-169) but merges duplicate
Paulis with the engine's hash-merge kernel instead of numpy + Rust ``unordered_unique``.
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass
class SimplifyMetadata:
    """Counters of one ``simplify`` call."""

    num_unique_paulis: int
    num_duplicate_paulis: int
    num_trimmed_paulis: int
    sum_trimmed_coeffs: float


@dataclass
class OperatorBudget:
    """Bounds on operator growth; backpropagation stops when one is exceeded."""

    max_paulis: int | None = None
    """Maximum number of distinct Pauli terms over all observables."""
    max_qwc_groups: int | None = None
    """Maximum number of qubit-wise commuting groups over all observables."""
    simplify: bool = True
    """Merge duplicates and trim zeros after every gate."""
    atol: float | None = None
    """Absolute zero tolerance (``None`` = ``SparsePauliOp.atol`` = 1e-8)."""
    rtol: float | None = None
    """Relative zero tolerance (``None`` = ``SparsePauliOp.rtol``; irrelevant against zero)."""

    def is_active(self) -> bool:
        return not (self.max_paulis is None and self.max_qwc_groups is None)


def simplify(operator, *, atol: float | None = None, rtol: float | None = None):
    """Trim near-zero rows, merge duplicate Paulis, trim near-zero sums (device path).

    Returns ``(operator, SimplifyMetadata)`` with the counter definitions of
    ``utils/simplify.py:128-152``.  ``rtol`` is accepted for API compatibility; against zero it
    does not enter ``np.isclose`` (:120, :147).
    """
    from ..engine import TermTable

    del rtol
    if atol is None:
        atol = operator.atol
    with TermTable.from_rows(operator, atol=atol) as (table, md):
        return table.to_operator(), md
