"""Per-slice and whole-run metadata containers (host side).

Field-for-field mirrors of ``SliceMetadata`` / ``OBPMetadata`` in the reference
(``/root/reference/qiskit_addon_obp/utils/metadata.py:25-119``) including the JSON round trip
(:224-259).  The two budget functions feed the truncation threshold search, so they reproduce
the reference's floating-point evaluation order exactly (:152-157, :197-222): the error history is
summed left to right starting from ``0.0`` and the running budget is clipped after every slice.
"""

from __future__ import annotations

import json
from dataclasses import asdict, dataclass

from .simplify import OperatorBudget
from .truncating import TruncationErrorBudget


@dataclass
class SliceMetadata:
    """Counters recorded while one slice is backpropagated (one list entry per observable)."""

    slice_errors: list[float]
    """Truncation error incurred in this slice."""
    raw_num_paulis: list[int]
    """Number of raw rows (before simplification) produced by the last applied gate."""
    num_unique_paulis: list[int] | None
    """Distinct Pauli keys among the raw rows that passed the zero filter (``None`` without simplify)."""
    num_duplicate_paulis: list[int] | None
    """Raw rows merged into an already seen key (``None`` without simplify)."""
    num_trimmed_paulis: list[int] | None
    """Rows/keys removed because their coefficient was within ``atol`` of zero."""
    sum_trimmed_coeffs: list[float] | None
    """Sum of the merged coefficients that were trimmed."""
    num_truncated_paulis: list[int]
    """Terms removed by the error-budget truncation."""
    num_paulis: list[int]
    """Terms left at the end of the slice."""
    sum_paulis: int | None
    """Distinct Pauli keys over all observables (only when ``max_paulis`` is bounded)."""
    num_qwc_groups: int | None
    """Qubit-wise commuting groups over all observables (only when ``max_qwc_groups`` is bounded)."""


@dataclass
class OBPMetadata:
    """Everything :func:`~qiskit_addon_obp_b200.backpropagate` reports besides the operators."""

    truncation_error_budget: TruncationErrorBudget
    num_slices: int | None
    operator_budget: OperatorBudget
    backpropagation_history: list[SliceMetadata]
    num_backpropagated_slices: int

    def accumulated_error(self, observable_idx: int, slice_idx: int | None = None) -> float:
        """Sum of ``slice_errors[observable_idx]`` over the first ``slice_idx`` slices."""
        if slice_idx is None:
            slice_idx = self.num_backpropagated_slices
        total = 0.0
        for entry in self.backpropagation_history[:slice_idx]:
            total += entry.slice_errors[observable_idx]
        return total

    def left_over_error_budget(self, observable_idx: int, slice_idx: int | None = None) -> float:
        """Budget available to slice ``slice_idx``: cyclic per-slice grants minus what was spent,
        never more than ``max_error_total`` minus the error accumulated so far."""
        if slice_idx is None:
            slice_idx = self.num_backpropagated_slices
        grants = self.truncation_error_budget.per_slice_budget
        cap = self.truncation_error_budget.max_error_total
        left = 0.0
        spent = 0.0  # == accumulated_error(observable_idx, i), same summation order
        for i in range(slice_idx + 1):
            left = left + grants[i % len(grants)]
            if i > 0:
                used = self.backpropagation_history[i - 1].slice_errors[observable_idx]
                left -= used
                spent += used
            left = min(left, cap - spent)
        return left

    @classmethod
    def from_json(cls, json_file: str) -> OBPMetadata:
        with open(json_file) as fh:
            raw = json.load(fh)
        return cls(
            truncation_error_budget=TruncationErrorBudget(**raw["truncation_error_budget"]),
            num_slices=raw["num_slices"],
            operator_budget=OperatorBudget(**raw["operator_budget"]),
            backpropagation_history=[SliceMetadata(**s) for s in raw["backpropagation_history"]],
            num_backpropagated_slices=raw["num_backpropagated_slices"],
        )

    def to_json(self, json_file: str, **kwargs) -> None:
        with open(json_file, "w") as fh:
            json.dump(asdict(self), fh, **kwargs)
