"""``backpropagate`` — the reference's driver loop with the per-gate work on the GPU engine.

Mirrors ``/root/reference/qiskit_addon_obp/backpropagation.py:48-311``: same signature, same
validation messages, same slice/observable/gate order, same light-cone rule, same per-slice
metadata, same commit/rollback and stopping behaviour.  What changes is where the work runs: each
observable is a device-resident :class:`~qiskit_addon_obp_b200.engine.TermTable`; the applied gates
of a slice are classified on the host (:mod:`.gates`) and executed as one gate program
(``obp_apply_ops``), followed by the error-budget truncation (``obp_truncate``).
"""

from __future__ import annotations

import logging
import os
import signal
import sys
import time
from collections.abc import Sequence

import numpy as np

from . import engine
from .engine import (OBP_ATOL_RAW, OBP_FLAG_EXACT_COUNTERS, OBP_FLAG_OWN_SLOT, OBP_OP_CLIFFORD, OBP_OP_GENERIC, OP_DTYPE, CapacityError,
                     TermTable)
from .gates import program_for, raw_program_for
from .ir import QuantumCircuit, SparsePauliOp, topological_op_order
from .sharded import ShardedTable, communicator_from_env
from .utils.metadata import OBPMetadata, SliceMetadata
from .utils.simplify import OperatorBudget
from .utils.truncating import TruncationErrorBudget

LOGGER = logging.getLogger(__name__)


class TimeoutException(Exception):
    """Raised by the ``max_seconds`` alarm (``backpropagation.py:42``)."""


class RunInfo:
    """Work counters of the last :func:`backpropagate` call (used by ``bench.py``)."""

    updates = 0  # term-gate updates of committed work: sum over applied gates of the term count at gate entry
    wasted_updates = 0  # updates of slice attempts that overflowed the table and were rolled back
    applied_gates = 0
    device_ms = 0.0  # CUDA-event time from the first table load to just before the download
    h2d_bytes = 0  # packed keys + coefficients uploaded
    d2h_bytes = 0  # packed keys + coefficients downloaded
    profile: dict | None = None  # per kernel class: launches, ms, term_updates (PROFILE only)
    sharded = False  # the observable was a hash-sharded table: `updates` is already the global count
    exchanges = 0  # sharded: rotations that needed a record exchange
    exchanged_records = 0
    #: rotation-only gate programs applied in shared memory: (committed runs, failed runs repeated in shorter pieces,
    #: runs handed to the per-gate kernels)
    segments = (0, 0, 0)


#: set True (bench.py does) to bracket every kernel launch with CUDA events
PROFILE = False
#: print one line per gate program (sizes, wall time) — debugging aid
TRACE = bool(os.environ.get("OBP_B200_TRACE"))


last_run = RunInfo()


def _default_capacity(num_qubits: int, operator_budget: OperatorBudget, n_obs: int, device: int, comm=None) -> int:
    """Rows of one observable's table (the whole table when it is sharded over the ranks of ``comm``).

    With ``max_paulis`` set, 4 x max_paulis: a slice is given up once it holds 3 x max_paulis live terms
    (``_stop_on_overflow``) and compaction keeps the dense rows below 1.25 x the live ones.  Bounded by device
    memory: 60 % of what is free, divided by the bytes one term costs (table, snapshot, scratch, index — and the
    two exchange inboxes of a shard)."""
    env = os.environ.get("OBP_B200_CAPACITY")
    if env:
        return int(env)
    free, _ = engine.device_memory(device)
    per_term = int(engine.lib().obp_bytes_per_term(num_qubits))
    n_ranks = 1
    if comm is not None and hasattr(comm, "rank"):  # one GPU per rank
        n_ranks = comm.n_ranks
        if os.environ.get("OBP_B200_P2P", "1") not in ("0", ""):
            per_term += 2 * int(engine.lib().obp_record_bytes(num_qubits))
    limit = max(1 << 16, int(0.6 * free / per_term / max(1, n_obs)))
    if n_ranks > 1:
        limit = int(limit * n_ranks / 1.25)  # a rank holds 1.25 / n_ranks of the table (imbalance margin)
    if operator_budget.max_paulis is not None:
        want = max(1 << 20, 4 * int(operator_budget.max_paulis))
    else:
        want = 1 << 22
    return int(min(want, limit, 0xFFFF0000 * n_ranks))


def backpropagate(
    observables,
    slices: Sequence[QuantumCircuit],
    *,
    truncation_error_budget: TruncationErrorBudget | None = None,
    operator_budget: OperatorBudget | None = None,
    max_seconds: int | None = None,
):
    """Backpropagate ``slices`` (in reverse order) onto ``observables``.

    Returns ``(observables, remaining_slices, OBPMetadata)`` exactly like the reference; a single
    observable in gives a single observable out (``backpropagation.py:307-309``).
    """
    from . import qiskit_adapter

    observables, slices_ir, restore_types = qiskit_adapter.to_ir(observables, slices)

    operator_budget = operator_budget or OperatorBudget()
    truncation_error_budget = truncation_error_budget or TruncationErrorBudget()
    single = isinstance(observables, SparsePauliOp)
    obs_in: list[SparsePauliOp] = [observables] if single else list(observables)

    _validate_input_options(obs_in, slices_ir, operator_budget)
    raw = not operator_budget.simplify  # nothing merged or trimmed: every gate is a raw k*k expansion
    # OBP_B200_ROW_ORDER=1: rows come back in exactly the reference's order (first appearance among each
    # gate's k*k raw rows).  A slow parity mode: every gate runs as the literal expansion + ordered merge.
    ordered = os.environ.get("OBP_B200_ROW_ORDER", "0") == "1" and not raw
    expand_all = raw or ordered

    atol = SparsePauliOp.atol if operator_budget.atol is None else float(operator_budget.atol)
    num_qubits = slices_ir[0].num_qubits
    n_obs = len(obs_in)
    # reduce_op (utils/operations.py:176-181): support set, error for a pure identity observable
    support_mask = [sum(1 << q for q in _support(o)) for o in obs_in]  # qubit sets as integer bit masks
    device = engine.default_device()
    comm = communicator_from_env(device)  # None: one table per observable on this GPU
    capacity = _default_capacity(num_qubits, operator_budget, n_obs, device, comm)
    if comm is not None and ordered:
        raise NotImplementedError("OBP_B200_ROW_ORDER=1 is not supported on a sharded table")
    if comm is not None and hasattr(comm, "rank") and max_seconds is not None:
        # the alarm would fire at a different point of the loop on every rank and leave the peers inside a
        # collective; a timeout on a multi-process table needs a stop flag agreed at slice boundaries
        raise NotImplementedError("max_seconds is not supported on a table sharded over several processes")

    metadata = OBPMetadata(
        truncation_error_budget=truncation_error_budget,
        operator_budget=operator_budget,
        num_slices=len(slices_ir),
        backpropagation_history=[],
        num_backpropagated_slices=0,
    )
    # slices the engine rejected before finishing them: (history index, observable, "lookahead" | "memory");
    # a plain attribute, not a dataclass field, so the reference's JSON schema of OBPMetadata is unchanged
    metadata.device_capacity_stops = []
    info = RunInfo()
    tables: list[TermTable] = []
    lengths = [len(o) for o in obs_in]  # len(observables_tmp[i]) as the reference sees it
    # a table only exists once a gate touches its observable: until then the reference carries the
    # input rows along unchanged (duplicates and zeros included)
    loaded = [False] * n_obs
    loaded_committed = [False] * n_obs
    dirty = [_has_duplicates_or_zeros(o, atol) for o in obs_in]
    raw_first = [False] * n_obs  # the table holds the unsimplified input: its first gate runs as a raw expansion
    trunc_active = truncation_error_budget.is_active()

    if max_seconds is not None:
        if sys.platform == "win32":
            raise RuntimeError("The `max_seconds` argument is not available on Windows.")

        def handle_timeout(signum, frame):
            raise TimeoutException()

        signal.signal(signal.SIGALRM, handle_timeout)
        signal.alarm(max_seconds)

    try:
        for i in range(n_obs):
            cap_i = max(capacity, 2 * len(obs_in[i]))
            if comm is None:
                tables.append(TermTable.acquire(num_qubits, cap_i, device))
                tables[-1]._seg0 = tables[-1].segment_stats()  # (pooled contexts keep counting)
                tables[-1].set_row_order(ordered)
            else:
                tables.append(ShardedTable.acquire(num_qubits, cap_i, comm))
        try:
            reversed_slices = slices_ir[::-1]
            log_info = LOGGER.isEnabledFor(logging.INFO)
            merged_until = 0  # slices before this position were handled as part of a merged Clifford group
            # Several Clifford-only slices as one gate program (OBP_B200_MERGE_CLIFFORD_SLICES=1; off by default: it
            # halves the host round trips of the 127-qubit workload and was measured to gain nothing, the time of such a
            # slice is on the device side of its program): only where nothing between them can stop the run or
            # needs the operator on the host (one unsharded observable, no truncation, no QWC budget, no timeout, no
            # per-gate DEBUG log, fast path)
            can_merge = (
                n_obs == 1 and comm is None and not raw and not ordered and not trunc_active and max_seconds is None
                and operator_budget.max_qwc_groups is None and not LOGGER.isEnabledFor(logging.DEBUG)
                and os.environ.get("OBP_B200_MERGE_CLIFFORD_SLICES", "0") not in ("0", "")
            )
            for slice_pos, slice_ in enumerate(reversed_slices):
                if slice_pos < merged_until:
                    continue
                next_slice = reversed_slices[slice_pos + 1] if slice_pos + 1 < len(reversed_slices) else None
                if can_merge and loaded[0] and not tables[0].zero_operator:
                    group = _merge_clifford_slices(tables[0], reversed_slices, slice_pos, support_mask[0], expand_all)
                    if len(group) >= 2 and _run_clifford_group(tables[0], group, atol, metadata, operator_budget, info,
                                                               lengths, support_mask, log_info):
                        merged_until = slice_pos + len(group)
                        loaded_committed = list(loaded)
                        continue
                sm = SliceMetadata(
                    slice_errors=[0.0] * n_obs,
                    raw_num_paulis=[0] * n_obs,
                    num_unique_paulis=None if raw else list(lengths),
                    num_duplicate_paulis=None if raw else [0] * n_obs,
                    num_trimmed_paulis=None if raw else [0] * n_obs,
                    sum_trimmed_coeffs=None if raw else [0] * n_obs,
                    num_truncated_paulis=[0] * n_obs,
                    num_paulis=[0] * n_obs,
                    sum_paulis=None,
                    num_qwc_groups=None,
                )
                metadata.backpropagation_history.append(sm)
                prog = _slice_program(slice_, expand_all)
                overflow_peak: list[int | None] = [None] * n_obs
                for i in range(n_obs):
                    plan = prog.plan(support_mask[i])  # gates inside the light cone (backpropagation.py:183)
                    applied = plan.applied
                    non_trivial = bool(applied)
                    sidx = metadata.num_backpropagated_slices
                    if non_trivial:
                        if not loaded[i]:
                            # An input with duplicate or near-zero rows reaches the reference's first gate
                            # unsimplified (raw_num_paulis counts every given row): keep the rows as given and
                            # run that one gate as the literal k*k expansion + simplify.
                            first = applied[0]
                            if ordered:  # every gate is an expansion + merge anyway; so is the ordered reset
                                raw_first[i] = dirty[i] and prog.nodes[first].name != "LayerError"
                            else:
                                raw_first[i] = (
                                    dirty[i] and not raw and not prog.special[first] and first not in prog.errors
                                )
                            tables[i].load(obs_in[i], OBP_ATOL_RAW if (raw or raw_first[i]) else None)
                            info.h2d_bytes += len(obs_in[i]) * (16 * tables[i].words + 16)
                            tables[i].profile_enable(PROFILE)
                            tables[i].profile(reset=True)
                            tables[i].timer_start()
                            tables[i].snapshot()
                            loaded[i] = True
                        prefetch = None
                        if next_slice is not None and comm is None:
                            # while the device runs this slice: compile the next one and its light cone
                            def prefetch(ns=next_slice, after=plan.support_after):
                                _slice_program(ns, expand_all).plan(after)

                        if comm is not None:
                            # a table that is split into shards during this slice (overflow of the replicated copy)
                            # plans its shard function against the gates from this slice on
                            tables[i].lookahead = (reversed_slices, slice_pos, support_mask[i], expand_all)
                        committed_work = (info.updates, info.applied_gates)
                        while True:
                            try:
                                _run_slice(tables[i], prog, plan, atol, sm, i, info, raw, raw_first[i] and not ordered,
                                           prefetch)
                                prefetch = None
                                raw_first[i] = False
                                if trunc_active:
                                    budget = metadata.left_over_error_budget(i, sidx)
                                    err, removed = tables[i].truncate(
                                        budget, truncation_error_budget.p_norm, truncation_error_budget.tol
                                    )
                                    sm.slice_errors[i] = err
                                    sm.num_truncated_paulis[i] = int(removed)
                                break
                            except CapacityError as exc:
                                # the table overflowed mid-slice: roll back to the committed state; the work of the
                                # discarded attempt is not counted as useful work (RunInfo.wasted_updates)
                                info.wasted_updates += info.updates - committed_work[0]
                                info.updates, info.applied_gates = committed_work
                                prefetch = None
                                tables[i].restore()
                                peak = int(getattr(exc, "n_live", 0))
                                stop = _stop_on_overflow(operator_budget, trunc_active, peak)
                                LOGGER.info(
                                    f"[{metadata.num_backpropagated_slices:3}] device table of the {i}-th observable "
                                    f"overflowed mid-slice ({peak} live terms, capacity {tables[i].capacity}): "
                                    + ("slice rejected" if stop else "growing the table")
                                )
                                if stop:
                                    overflow_peak[i] = peak
                                    break
                                try:
                                    _grow(tables[i])
                                except MemoryError:
                                    if operator_budget.max_paulis is None or trunc_active or peak <= operator_budget.max_paulis:
                                        raise
                                    # device memory is exhausted while the slice already holds more terms than
                                    # max_paulis: reject the slice, flagged as a memory-limited stop
                                    overflow_peak[i] = peak
                                    metadata.device_capacity_stops.append((len(metadata.backpropagation_history) - 1, i, "memory"))
                                    break
                        if overflow_peak[i] is not None:
                            # the slice is oversized for this observable; its size is only known as a lower bound
                            sm.num_paulis[i] = overflow_peak[i]
                            if not any(s[:2] == (len(metadata.backpropagation_history) - 1, i) for s in metadata.device_capacity_stops):
                                metadata.device_capacity_stops.append((len(metadata.backpropagation_history) - 1, i, "lookahead"))
                            continue
                        support_mask[i] = plan.support_after
                        lengths[i] = len(tables[i])
                    if trunc_active and non_trivial:
                        lengths[i] = len(tables[i])
                        if log_info:
                            LOGGER.info(
                                f"[{sidx:3}] Accumulated error for the {i}-th observable: "
                                f"{metadata.accumulated_error(i, sidx + 1):.10f}"
                            )
                    sm.num_paulis[i] = lengths[i]
                    if log_info:
                        LOGGER.info(
                            f"[{metadata.num_backpropagated_slices:3}] size of the {i}-th observable: "
                            f"{sm.num_paulis[i]}"
                        )
                if any(p is not None for p in overflow_peak):
                    # Deviation from backpropagation.py:266-275, 419-431: the reference finishes the
                    # slice whatever its size and reports the exact count; the engine stops the slice for an
                    # observable once it holds OBP_B200_OVERSIZE_LOOKAHEAD x max_paulis live terms (or device memory
                    # is exhausted) and reports the count reached, a lower bound that already exceeds max_paulis.
                    sm.sum_paulis = max(p for p in overflow_peak if p is not None)
                    LOGGER.info(
                        f"[{metadata.num_backpropagated_slices:3}] Too many Pauli terms: at least "
                        f"{sm.sum_paulis} (slice stopped early, see OBPMetadata.device_capacity_stops)"
                    )
                    break
                if operator_budget.is_active() and _observable_oversized(
                    tables, loaded, obs_in, operator_budget, metadata.num_backpropagated_slices, sm, raw
                ):
                    break
                # commit (backpropagation.py:279-285): the snapshot is the deepcopy of the reference
                for i in range(n_obs):
                    if loaded[i]:
                        if comm is not None:  # (a split at this commit plans against the gates still to come)
                            tables[i].lookahead = (reversed_slices, slice_pos + 1, support_mask[i], expand_all)
                        tables[i].snapshot()
                loaded_committed = list(loaded)
                metadata.num_backpropagated_slices += 1
        except TimeoutException:
            LOGGER.warning("Reached the specified max_seconds timeout!")
        finally:
            if max_seconds is not None:
                signal.alarm(0)

        LOGGER.info(f"Backpropagated {metadata.num_backpropagated_slices} DAG multi-graph slices")
        n_done = metadata.num_backpropagated_slices
        slices_out = slices[:-n_done]  # same `[:-0]` quirk as backpropagation.py:300
        obs_out = []
        for i in range(n_obs):
            if not loaded_committed[i]:
                obs_out.append(obs_in[i].copy())
                continue
            # after a break or a timeout the table holds an uncommitted slice: roll it back
            if len(metadata.backpropagation_history) != n_done:
                tables[i].restore()
            info.device_ms = max(info.device_ms, tables[i].timer_stop())
            if PROFILE:
                prof = tables[i].profile()
                if info.profile is None:
                    info.profile = prof
                else:
                    for k, v in prof.items():
                        for f in v:
                            info.profile[k][f] += v[f]
            if os.environ.get("OBP_B200_DOWNLOAD", "1") == "0":
                # benchmark aid for tables of 1e8 terms (tens of GB): leave the result on the device(s)
                obs_out.append(SparsePauliOp([], num_qubits=num_qubits))
                continue
            obs_out.append(tables[i].download())
            info.d2h_bytes += len(obs_out[-1]) * (16 * tables[i].words + 16)
    finally:
        for t in tables:
            info.exchanges += getattr(t, "exchanges", 0)
            info.exchanged_records += getattr(t, "exchanged_records", 0)
            if hasattr(t, "_seg0"):
                info.segments = tuple(a + b - c for a, b, c in zip(info.segments, t.segment_stats(), t._seg0))
            t.profile_enable(False)
            t.release()

    last_run.updates = info.updates
    last_run.wasted_updates = info.wasted_updates
    last_run.applied_gates = info.applied_gates
    last_run.device_ms = info.device_ms
    last_run.h2d_bytes = info.h2d_bytes
    last_run.d2h_bytes = info.d2h_bytes
    last_run.profile = info.profile
    last_run.sharded = comm is not None
    last_run.exchanges = info.exchanges
    last_run.exchanged_records = info.exchanged_records
    last_run.segments = info.segments
    obs_out = [restore_types(o) for o in obs_out]
    return (obs_out[0] if single else obs_out), slices_out, metadata


def _clifford_only(plan: "SlicePlan") -> bool:
    """The slice's applied gates are one gate program of Clifford-class gates only (the term count cannot grow)."""
    if len(plan.steps) != 1 or plan.steps[0][0] != "ops":
        return False
    ops = plan.steps[0][1]
    cached = getattr(plan, "_clifford_only", None)
    if cached is None:
        cached = bool(len(ops)) and bool((ops["kind"] == OBP_OP_CLIFFORD).all())
        try:
            plan._clifford_only = cached
        except AttributeError:  # (__slots__: no cache)
            pass
    return cached


def _merge_clifford_slices(table, reversed_slices, slice_pos, support, expand_all, max_group: int = 8):
    """Consecutive slices from ``slice_pos`` on whose applied gates are all Clifford-class, with their plans.

    Such slices permute the terms (the raw-row filter can only remove some): none of them can exceed an operator
    budget the slice before them met, so they travel as ONE gate program with the last gate of each slice flagged
    for the reference's counters (``obp_last_op_stats``) instead of one host round trip each."""
    group = []
    pos = slice_pos
    while pos < len(reversed_slices) and len(group) < max_group:
        prog = _slice_program(reversed_slices[pos], expand_all)
        plan = prog.plan(support)
        if not plan.applied or not _clifford_only(plan):
            break
        group.append((prog, plan))
        support = plan.support_after
        pos += 1
    return group



def _run_clifford_group(table, group, atol, metadata, operator_budget, info, lengths, support_mask, log_info) -> bool:
    """Run a group of Clifford-only slices (one observable) as one gate program and book each slice exactly as the
    slice-by-slice loop would.  Returns False — with the table restored to the last commit and nothing booked — if
    anything unusual happens (the operator vanishes, the table overflows): the caller then takes the slices one by one."""
    parts = []
    ends = []
    for _, plan in group:
        ops = plan.steps[0][1]  # (its last gate already carries OBP_FLAG_EXACT_COUNTERS)
        parts.append(ops)
        ends.append((ends[-1] if ends else 0) + len(ops))
    merged = np.concatenate(parts)
    t0 = time.perf_counter() if TRACE else 0.0
    n_before = table.n
    try:
        table.apply_ops_submit(merged, atol)
        total = table.apply_ops_wait()
    except CapacityError:
        table.restore()
        return False
    stats = []
    try:
        stats = [table.last_op_stats(e - 1) for e in ends]
    except engine.EngineError:
        stats = []
    if not stats or any(int(st.n_out) == 0 for st in stats) or table.zero_operator:
        table.restore()  # (the all-trimmed operator has its own conventions: let the ordinary loop meet it)
        return False
    if TRACE:
        dt = time.perf_counter() - t0
        print(f"[obp trace] ops {len(merged):4d} (rot   0, {len(group)} Clifford-only slices) n {n_before:>10d} -> {int(total.n_out):>10d} "
              f"dense {int(total.n_dense):>10d} updates {int(total.updates):>12d} {dt * 1e3:9.3f} ms", flush=True)
    info.updates += int(total.updates)
    for (prog, plan), st in zip(group, stats):
        sm = SliceMetadata(
            slice_errors=[0.0],
            raw_num_paulis=[int(st.raw_rows)],
            num_unique_paulis=[int(st.num_unique)],
            num_duplicate_paulis=[int(st.num_duplicate)],
            num_trimmed_paulis=[int(st.num_trimmed)],
            sum_trimmed_coeffs=[0],
            num_truncated_paulis=[0],
            num_paulis=[int(st.n_out)],
            sum_paulis=None,
            num_qwc_groups=None,
        )
        c = complex(st.sum_trimmed[0], st.sum_trimmed[1])
        sm.sum_trimmed_coeffs[0] = c if c != 0 else 0.0
        metadata.backpropagation_history.append(sm)
        info.applied_gates += len(plan.applied)
        support_mask[0] = plan.support_after
        lengths[0] = int(st.n_out)
        if log_info:
            LOGGER.info(f"[{metadata.num_backpropagated_slices:3}] size of the 0-th observable: {sm.num_paulis[0]}")
        if operator_budget.max_paulis is not None:
            sm.sum_paulis = int(st.n_out)  # one observable: its keys are distinct; it cannot exceed what the slice before met
        metadata.num_backpropagated_slices += 1
    table.snapshot()  # the commit after the group's last slice (nothing can roll back to a point inside the group)
    return True



def _stop_on_overflow(operator_budget: OperatorBudget, trunc_active: bool, peak_live: int) -> bool:
    """Whether a mid-slice table overflow ends the backpropagation instead of growing the table.

    Without truncation a slice's term count does not come back down by a large factor (only exact
    cancellations remove terms), so once an observable holds ``OBP_B200_OVERSIZE_LOOKAHEAD`` (default 3)
    times ``max_paulis`` live terms inside a slice, the slice is declared oversized without being finished.
    ``OBP_B200_OVERSIZE_POLICY=exact`` turns this off: the table grows while device memory allows and the
    slice is judged on its exact final count, like ``backpropagation.py:266-275``."""
    if operator_budget.max_paulis is None or trunc_active:
        return False
    if os.environ.get("OBP_B200_OVERSIZE_POLICY", "lookahead") == "exact":
        return False
    factor = float(os.environ.get("OBP_B200_OVERSIZE_LOOKAHEAD", "3"))
    return peak_live >= factor * operator_budget.max_paulis


def _grow(table: TermTable) -> None:
    """Double the table, bounded by free device memory.  On a sharded table the verdict is agreed between the
    ranks first (one rank must not leave the collective ``reserve`` its peers are about to enter)."""
    free, _ = engine.device_memory(table.device)
    per_term = int(engine.lib().obp_bytes_per_term(table.num_qubits))
    new_cap = min(2 * table.capacity, 0xFFFF0000)
    comm = getattr(table, "comm", None)
    share = 1.25 / comm.n_ranks if comm is not None else 1.0
    refuse = new_cap <= table.capacity or (new_cap - table.capacity) * per_term * share > 0.9 * free
    if comm is not None:
        refuse = bool(comm.allreduce_max({r: [1.0 if refuse else 0.0] for r in table.tables})[0] > 0)
    if refuse:
        raise MemoryError(
            f"observable outgrew the device table ({table.capacity} terms) and it cannot be doubled"
        )
    table.reserve(new_cap)


class SliceProgram:
    """A slice compiled once: its gates in the reference's application order plus their device records.

    ``nodes`` are the instructions of ``circuit_to_dag(slice).topological_op_nodes()`` reversed
    (``backpropagation.py:170``) without barriers (:173); ``recs[k]`` is the classified ``obp_op``
    record of ``nodes[k]`` (``kind`` 0 for reset / LayerError, which have their own entry points).
    """

    __slots__ = ("nodes", "qubit_sets", "recs", "special", "errors", "data_ref", "generic", "touched", "all_idx",
                 "masks", "touched_mask", "is_reset", "plans", "gate_ids")

    def __init__(self, slice_: QuantumCircuit, raw: bool = False):
        self.data_ref = list(slice_.data)  # the instruction objects this program was compiled from (keeps them alive)
        classify = raw_program_for if raw else program_for
        ordered_all = topological_op_order(slice_)[::-1]
        self.nodes = [n for n in ordered_all if n.name != "barrier"]
        # the id the reference's DEBUG line gives a gate: len(op_nodes) - op_idx - 1, barriers counted (:239-242)
        self.gate_ids = [len(ordered_all) - idx - 1 for idx, n in enumerate(ordered_all) if n.name != "barrier"]
        self.qubit_sets = [frozenset(n.qubits) for n in self.nodes]
        self.touched = frozenset().union(*self.qubit_sets) if self.qubit_sets else frozenset()
        self.all_idx = list(range(len(self.nodes)))
        self.masks = [sum(1 << q for q in qs) for qs in self.qubit_sets]  # qubit sets as integer bit masks
        self.touched_mask = 0
        for m in self.masks:
            self.touched_mask |= m
        self.is_reset = [n.name == "reset" for n in self.nodes]
        self.plans: dict[int, SlicePlan] = {}
        self.recs = np.zeros(len(self.nodes), dtype=OP_DTYPE)
        self.special = [n.name in ("reset", "LayerError") for n in self.nodes]
        self.errors: dict[int, Exception] = {}
        self.generic: dict[int, object] = {}
        # classify each distinct gate once, then fill the records of all its occurrences in one go
        groups: dict[tuple, list[int]] = {}
        for k, node in enumerate(self.nodes):
            if not self.special[k]:
                key = (node.name, node.params) if node._matrix is None else (node.name, node._matrix.tobytes())
                groups.setdefault(key, []).append(k)
        qubit = self.recs["qubit"]
        for idxs in groups.values():
            try:
                prog = classify(self.nodes[idxs[0]])
            except (NotImplementedError, ValueError) as exc:  # only raised if the gate is applied
                for k in idxs:
                    self.errors[k] = exc
                continue
            if prog.kind == OBP_OP_GENERIC:  # own entry point, like reset / LayerError
                for k in idxs:
                    self.special[k] = True
                    self.generic[k] = prog
                continue
            self.recs[idxs] = prog.template[0]
            qubit[idxs, : prog.n_qubits] = np.array([self.nodes[k].qubits for k in idxs], dtype=np.int32)


    def plan(self, support: int) -> "SlicePlan":
        """The gates of this slice inside the light cone of ``support`` (a qubit bit mask) and their device steps.

        Purely structural (``backpropagation.py:183``: a gate acts iff it shares a qubit with the observable's
        current support, and then widens it — except ``reset``), so it is computed once per (slice, support)
        and can be prepared while the device is still busy with the previous slice.
        """
        plan = self.plans.get(support)
        if plan is None:
            if self.touched_mask & ~support == 0:
                applied, after = self.all_idx, support  # the light cone already covers the whole slice
            else:
                applied, after = [], support
                for k, m in enumerate(self.masks):
                    if after & m:
                        applied.append(k)
                        if not self.is_reset[k]:
                            after |= m
            if len(self.plans) >= 64:
                self.plans.clear()
            plan = self.plans[support] = SlicePlan(self, applied, after)
        return plan


class SlicePlan:
    """The applied gates of one slice for one light cone, cut into device steps.

    ``steps`` is a list of ``("ops", records)`` — a run of Clifford-class / rotation gates for
    ``obp_apply_ops``, the last gate of the slice flagged for exact counters — ``("node", k)`` — a gate
    with its own entry point (generic, reset, LayerError) — and ``("error", exc)`` — a gate the engine
    cannot express, raised when (and only when) it is reached.
    """

    __slots__ = ("applied", "support_after", "steps")

    def __init__(self, prog: SliceProgram, applied: list[int], support_after: int):
        self.applied = applied
        self.support_after = support_after
        self.steps: list[tuple] = []
        if not applied:
            return
        last = applied[-1]
        seg: list[int] = []

        def cut():
            nonlocal seg
            if seg:
                ops = prog.recs[seg]  # fancy index: a fresh array
                if seg[-1] == last:
                    ops["flags"][-1] = OBP_FLAG_EXACT_COUNTERS
                self.steps.append(("ops", ops, list(seg)))
                seg = []

        for k in applied:
            if not prog.special[k]:
                if k in prog.errors:
                    cut()
                    self.steps.append(("error", prog.errors[k], None))
                    return
                seg.append(k)
            else:
                cut()
                self.steps.append(("node", k, None))
        cut()


def _slice_program(slice_: QuantumCircuit, raw: bool = False) -> SliceProgram:
    """The compiled program of a slice, cached on the circuit object and revalidated on every call.

    The cache is valid only while ``slice_.data`` holds the very same instruction objects in the same order
    (``Instruction`` is immutable, so an instruction cannot change behind its identity): replacing an
    instruction in place, e.g. for a parameter sweep over the same slice objects, recompiles the slice.  The
    check is one C-level list comparison by identity, ~1 us per slice."""
    attr = "_obp_program_raw" if raw else "_obp_program"
    prog = getattr(slice_, attr, None)
    if prog is None or prog.data_ref != slice_.data:
        prog = SliceProgram(slice_, raw)
        setattr(slice_, attr, prog)
    return prog


def _run_slice(table: TermTable, prog: SliceProgram, plan: SlicePlan, atol: float, sm: SliceMetadata, i: int,
               info: RunInfo, raw: bool = False, first_as_expansion: bool = False, prefetch=None) -> None:
    """Execute the applied gates of one slice on one observable and record the last gate's counters.

    ``prefetch`` (optional) is host work for the NEXT slice; it runs between the submission of this slice's
    first gate program and the wait for it, i.e. while the device is busy.
    """
    stats = None
    applied = plan.applied
    if first_as_expansion:
        node = prog.nodes[applied[0]]
        g = raw_program_for(node)
        stats = table.apply_generic(node.qubits, g.codes, g.coeffs, atol, simplify=True)
        info.updates += int(stats.updates)
        info.applied_gates += 1
        if LOGGER.isEnabledFor(logging.DEBUG) and not isinstance(table, ShardedTable):
            _log_gate_size(prog.gate_ids[applied[0]], len(table))
        applied = applied[1:]
        plan = SlicePlan(prog, applied, plan.support_after)

    sharded = isinstance(table, ShardedTable)
    # the reference's per-gate DEBUG line (backpropagation.py:239-242): every gate then gets a pass of its own
    log_gates = LOGGER.isEnabledFor(logging.DEBUG) and not sharded
    for kind, what, seg in plan.steps:
        if kind == "ops":
            ops = what
            if log_gates:
                ops = ops.copy()
                ops["flags"] |= OBP_FLAG_OWN_SLOT
            t0 = time.perf_counter() if TRACE else 0.0
            n_before = table.n
            try:
                if sharded:
                    stats = table.apply_ops(ops, atol)
                else:
                    table.apply_ops_submit(ops, atol)
                    if prefetch is not None:
                        run, prefetch = prefetch, None
                        run()
                    stats = table.apply_ops_wait()
            except CapacityError as exc:
                info.updates += getattr(exc, "updates", 0)  # (moved to wasted_updates by the caller's rollback)
                raise
            info.updates += int(stats.updates)
            if log_gates:
                for k, size in zip(seg, table.last_op_sizes(len(ops))):
                    _log_gate_size(prog.gate_ids[k], size)
            if TRACE:
                dt = time.perf_counter() - t0
                kinds = ops["kind"]
                print(
                    f"[obp trace] ops {len(ops):4d} (rot {int((kinds == 2).sum()):3d}) n {n_before:>10d} -> {int(stats.n_out):>10d} "
                    f"dense {int(stats.n_dense):>10d} updates {int(stats.updates):>12d} {dt * 1e3:9.3f} ms "
                    f"{int(stats.updates) / max(dt, 1e-9) / 1e9:7.2f} G upd/s",
                    flush=True,
                )
            continue
        if kind == "error":
            raise what
        k = what
        node = prog.nodes[k]
        if k in prog.generic:
            g = prog.generic[k]
            stats = table.apply_generic(node.qubits, g.codes, g.coeffs, atol, simplify=not raw)
        elif node.name == "reset":
            stats = table.apply_reset(node.qubits[0], OBP_ATOL_RAW if raw else atol)
        else:
            ple = node.ple
            if ple is None:
                raise RuntimeError(
                    "Expected the LayerError circuit instruction to have a 'ple' "
                    "attribute but `None` was found."
                )
            gens = SparsePauliOp(list(ple.generators))
            gx = np.zeros((len(gens), table.num_qubits), dtype=bool)
            gz = np.zeros((len(gens), table.num_qubits), dtype=bool)
            gx[:, list(node.qubits)] = gens.x
            gz[:, list(node.qubits)] = gens.z
            stats = table.apply_damping(gx, gz, np.exp(-2 * np.asarray(ple.rates, dtype=float)), -1.0 if raw else atol)
        info.updates += int(stats.updates)
        if log_gates:
            _log_gate_size(prog.gate_ids[k], len(table))
    if prefetch is not None:
        prefetch()
    info.applied_gates += len(applied)
    # the slice reports the counters of its last applied gate (backpropagation.py:218-237)
    sm.raw_num_paulis[i] = int(stats.raw_rows)
    if raw:
        return  # simplify=False: the four simplify counters stay None (backpropagation.py:151-160)
    if getattr(stats, "counters_valid", True):
        sm.num_unique_paulis[i] = int(stats.num_unique)
        sm.num_duplicate_paulis[i] = int(stats.num_duplicate)
        sm.num_trimmed_paulis[i] = int(stats.num_trimmed)
        s = complex(stats.sum_trimmed[0], stats.sum_trimmed[1])
        sm.sum_trimmed_coeffs[i] = s if s != 0 else 0.0
    else:
        # sharded table, last gate Clifford-class: its coset count would need a cross-rank exchange
        sm.num_unique_paulis[i] = sm.num_duplicate_paulis[i] = sm.num_trimmed_paulis[i] = None
        sm.sum_trimmed_coeffs[i] = None


def future_ops(lookahead, max_slices: int = 48) -> np.ndarray | None:
    """The gate records the run applies next (light cone included), for a sharded table that plans its shard
    function (``obp_shard_split``).  Stops at the first gate with an entry point of its own (generic gates, reset,
    noise layers) and after ``max_slices`` slices; rotations beyond the plan still work, through the exchange path."""
    if lookahead is None:
        return None
    reversed_slices, start, support, expand_all = lookahead
    if expand_all:
        return None
    chunks = []
    for sl in reversed_slices[start : start + max_slices]:
        plan = _slice_program(sl, False).plan(support)
        for kind, what, _ in plan.steps:
            if kind != "ops":
                return np.concatenate(chunks) if chunks else None
            chunks.append(what)
        support = plan.support_after
    return np.concatenate(chunks) if chunks else None


def _log_gate_size(gate_id: int, size: int) -> None:
    LOGGER.debug(f"Size of the observable after backpropagating gate id {gate_id} in the current layer: {size}")


def _has_duplicates_or_zeros(op: SparsePauliOp, atol: float) -> bool:
    if len(op) == 0:
        return False
    if np.any(np.abs(op.coeffs) <= atol):
        return True
    xw, zw = op.packed()
    return len(np.unique(np.concatenate([xw, zw], axis=1), axis=0)) != len(op)


def _support(op: SparsePauliOp) -> list[int]:
    qargs = [int(q) for q in np.logical_or(op.x, op.z).any(axis=0).nonzero()[0]]
    if qargs == []:
        raise ValueError("Input operator may not be the identity operator.")
    return qargs


def _distinct_labels(ops: list[SparsePauliOp]) -> SparsePauliOp:
    """Coefficient-free union of the rows of several host operators, first-appearance order."""
    x = np.concatenate([o.x for o in ops], axis=0)
    z = np.concatenate([o.z for o in ops], axis=0)
    keys = np.concatenate([np.packbits(x, axis=1), np.packbits(z, axis=1)], axis=1)
    _, first = np.unique(keys, axis=0, return_index=True)
    first.sort()
    return SparsePauliOp((x[first], z[first]), np.ones(len(first)))


def _observable_oversized(tables, loaded, obs_in, operator_budget, slice_id, sm, raw=False) -> bool:
    """``backpropagation.py:409-442``: distinct keys / QWC groups over all observables."""
    from .utils.qwc import qwc_group_count

    max_paulis, max_qwc = operator_budget.max_paulis, operator_budget.max_qwc_groups
    if all(loaded):
        dev_tables = list(tables)
        host_ops: list[SparsePauliOp] = []
    else:
        dev_tables = [t for t, l in zip(tables, loaded) if l]
        host_ops = [o for o, l in zip(obs_in, loaded) if not l]
    union = None

    def union_op() -> SparsePauliOp:
        nonlocal union
        if union is None:
            union = _distinct_labels([t.download() for t in dev_tables] + host_ops)
        return union

    if max_paulis is not None:
        if raw:
            sm.sum_paulis = len(union_op())  # unmerged rows: count the distinct labels on the host
        elif not host_ops and len(dev_tables) == 1 and not dev_tables[0].zero_operator:
            sm.sum_paulis = dev_tables[0].n  # one observable: its keys are already distinct
        elif not host_ops and all(not t.zero_operator and isinstance(t, TermTable) for t in dev_tables):
            sm.sum_paulis = engine.union_count(dev_tables)
        # (several sharded observables: their hash frames differ, equal keys need not share a rank — the union
        # is counted on the gathered operators, like the general case below)
        else:
            sm.sum_paulis = len(union_op())
        if sm.sum_paulis > max_paulis:
            LOGGER.info(f"[{slice_id:3}] Too many Pauli terms: {sm.sum_paulis}")
            return True
    if max_qwc is not None:
        sm.num_qwc_groups = qwc_group_count(union_op())
        if sm.num_qwc_groups > max_qwc:
            LOGGER.info(
                f"[{slice_id:3}] Too many qubit-wise commuting Pauli groups: {sm.num_qwc_groups}"
            )
            return True
    return False


def _validate_input_options(obs, slices, operator_budget: OperatorBudget) -> None:
    """``backpropagation.py:314-375`` — same checks, same messages."""
    from .utils.qwc import qwc_group_count

    num_qubits = slices[0].num_qubits
    for i, s in enumerate(slices):
        if s.num_qubits != num_qubits:
            raise ValueError(
                "All slices must be defined on the same number of qubits. "
                f"slices[0] contains {num_qubits} qubits, but slices[{i}] contains "
                f"{s.num_qubits} qubits."
            )
    obs_qubits = obs[0].num_qubits
    for i, o in enumerate(obs):
        if o.num_qubits != obs_qubits:
            raise ValueError(
                "Input observables must all act on the same number of qubits. "
                f"observables[0] acts on {obs_qubits} qubits, but observables[{i}]"
                f" acts on {o.num_qubits} qubits."
            )
    if obs_qubits != num_qubits:
        raise ValueError(
            "The input observables must be defined on the same number of qubits as the "
            "circuit slices."
        )
    max_paulis, max_qwc = operator_budget.max_paulis, operator_budget.max_qwc_groups
    if max_paulis is not None and max_paulis < 1:
        raise ValueError("Limiting the number of Pauli terms to less than 1 does not make sense.")
    if max_qwc is not None and max_qwc < 1:
        raise ValueError(
            "Limiting the number of qubit-wise commmuting Pauli groups to less than 1 does not "
            "make sense."
        )
    if operator_budget.is_active():
        allp = _distinct_labels(list(obs))
        if max_paulis is not None and len(allp) > max_paulis:
            raise ValueError(
                f"You specified a maximum number of Pauli terms of {max_paulis}, but the "
                f"provided observables already exceed this threshold with a total of "
                f"{len(allp)} terms."
            )
        if max_qwc is not None:
            groups = qwc_group_count(allp)
            if groups > max_qwc:
                raise ValueError(
                    f"You specified a maximum number of qubit-wise commuting Pauli groups of "
                    f"{max_qwc}, but the provided observables already exceed this threshold "
                    f"with a total of {groups} terms."
                )
