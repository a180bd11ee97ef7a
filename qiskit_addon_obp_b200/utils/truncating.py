"""Truncation error budget (host side) and the device-backed threshold search.

``TruncationErrorBudget`` / ``setup_budget`` mirror ``utils/truncating.py:25-141`` of the
reference (same fields, defaults and branch order, including the ``isinstance(x, float)`` test
at :121 that sends every non-float through ``list(...)``).  ``truncate_binary_search`` has the
reference's signature (:144-150) but runs on the GPU engine.
"""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass, field

import numpy as np


@dataclass
class TruncationErrorBudget:
    """How much truncation error may be spent, per slice and in total."""

    per_slice_budget: list[float] = field(default_factory=list)
    """Error grant per backpropagated slice; cycled through when shorter than the slice count."""
    max_error_total: float = 0.0
    """Hard cap on the accumulated truncation error (may be ``numpy.inf``)."""
    p_norm: int = 1
    """Lp norm in which truncation errors are measured."""
    tol: float = 1e-8
    """Absolute tolerance at which the threshold search stops."""

    def is_active(self) -> bool:
        return self.max_error_total > 0 and any(b > 0 for b in self.per_slice_budget)


def setup_budget(
    *,
    max_error_per_slice: float | Sequence[float] | None = None,
    max_error_total: float | None = None,
    num_slices: int | None = None,
    p_norm: int = 1,
) -> TruncationErrorBudget:
    """Build a :class:`TruncationErrorBudget` the way the reference does (:114-141)."""
    if max_error_per_slice is None and max_error_total is None:
        raise ValueError("max_error_per_slice and max_error_total may not both be None")
    if max_error_per_slice is not None:
        if isinstance(max_error_per_slice, float):
            grants = [max_error_per_slice]
        else:
            grants = list(max_error_per_slice)
    elif num_slices is not None:
        grants = [max_error_total / num_slices]
    else:
        grants = [max_error_total]
    return TruncationErrorBudget(
        grants, np.inf if max_error_total is None else max_error_total, p_norm
    )


def truncate_binary_search(observable, budget: float, *, p_norm: int = 1, tol: float = 1e-8):
    """Drop the lowest-weight terms whose Lp norm stays within ``budget`` (device path).

    Same contract as ``utils/truncating.py:144-204``: returns ``(truncated operator, error)``.
    """
    from ..engine import TermTable

    with TermTable.from_operator(observable) as table:
        err, _ = table.truncate(budget, p_norm, tol)
        return table.to_operator(), err
