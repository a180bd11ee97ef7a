"""GPU parity tests of the hash-sharded table: R virtual ranks on one B200 (``LocalComm``) run the
same kernels and exchange protocol as R GPUs, and must reproduce the oracle like the single table."""

from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_same_metadata, assert_same_operator
from oracle import obp_oracle as oracle
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200.ir import SparsePauliOp
from qiskit_addon_obp_b200.utils.simplify import OperatorBudget
from qiskit_addon_obp_b200.utils.truncating import TruncationErrorBudget, setup_budget
from test_engine_gpu import random_observable, random_slices

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _shard_from_the_first_commit(monkeypatch):
    """The tables of these tests are tiny: shard them at once (the default keeps a table replicated on every rank
    until it holds 2^20 terms, see test_late_sharding_* for that path)."""
    monkeypatch.setenv("OBP_B200_SHARD_MIN_TERMS", "0")


def _compare(monkeypatch, ranks, obs, slices, p2p=True, **kw):
    monkeypatch.setenv("OBP_B200_SHARDS", str(ranks))
    monkeypatch.setenv("OBP_B200_P2P", "1" if p2p else "0")
    got, rest_g, md_g = backpropagate(obs, slices, **kw)
    monkeypatch.delenv("OBP_B200_SHARDS")
    want, rest_w, md_w = oracle.backpropagate(obs, slices, **kw)
    assert len(rest_g) == len(rest_w)
    single = isinstance(obs, SparsePauliOp)
    for g, w in zip([got] if single else got, [want] if single else want):
        assert_same_operator(g, w)
    assert_same_metadata(md_g, md_w, allow_missing_counters=True)
    return got, md_g


@pytest.mark.parametrize("p2p", [True, False], ids=["inbox", "staged"])
@pytest.mark.parametrize("ranks", [2, 4, 8])
@pytest.mark.parametrize("nq", [6, 70, 130])
def test_sharded_random_circuit(monkeypatch, ranks, nq, p2p):
    """Both exchange paths: records stored straight into the peer's inbox, or staged and handed over."""
    rng = np.random.default_rng(31 * nq + ranks)
    slices = random_slices(nq, 6, rng, gates_per_slice=8)
    obs = random_observable(nq, 4, rng)
    _compare(monkeypatch, ranks, obs, slices, p2p=p2p)


@pytest.mark.parametrize("ranks,p", [(2, 1), (8, 2)])
def test_sharded_truncation(monkeypatch, ranks, p):
    rng = np.random.default_rng(5 + ranks)
    nq = 10
    slices = random_slices(nq, 8, rng, gates_per_slice=6, angles=[0.1, 0.3, -0.2, 0.7])
    obs = random_observable(nq, 3, rng)
    _compare(monkeypatch, ranks, obs, slices, truncation_error_budget=TruncationErrorBudget([0.02, 0.0, 0.05], 0.2, p, 1e-8))
    _compare(monkeypatch, ranks, obs, slices, truncation_error_budget=setup_budget(max_error_total=0.05, num_slices=8, p_norm=p))


def test_sharded_rotation_counters_exact(monkeypatch):
    """Slices ending on a rotation report the reference's simplify counters from the per-rank pieces."""
    rng = np.random.default_rng(9)
    nq = 7
    slices = random_slices(nq, 10, rng, p_rot=1.0, gates_per_slice=5)
    obs = random_observable(nq, 5, rng, weight=3)
    _, md = _compare(monkeypatch, 4, obs, slices)
    assert all(s.num_unique_paulis[0] is not None for s in md.backpropagation_history)


def test_sharded_configs(monkeypatch):
    from qiskit_addon_obp_b200.workloads import config1_xy_chain, config2_kicked_ising, config4_near_clifford

    wl = config1_xy_chain()
    _compare(monkeypatch, 8, wl.observables, wl.slices, **wl.kwargs())
    wl = config2_kicked_ising(theta_h=0.3, steps=2, max_paulis=3000)
    _compare(monkeypatch, 4, wl.observables, wl.slices, **wl.kwargs())
    wl = config4_near_clifford(num_qubits=1000, depth=10, t_prob=0.01, max_paulis=4000, per_slice=1e-4)
    _compare(monkeypatch, 8, wl.observables, wl.slices, **wl.kwargs())


def test_sharded_max_paulis_and_multi_observable(monkeypatch):
    rng = np.random.default_rng(2)
    nq = 8
    slices = random_slices(nq, 10, rng, gates_per_slice=6)
    obs = random_observable(nq, 3, rng)
    _compare(monkeypatch, 2, obs, slices, operator_budget=OperatorBudget(max_paulis=40))
    _compare(monkeypatch, 4, [obs, SparsePauliOp("IIIIIIIZ")], slices)
    # the operator budget over several sharded observables counts the union of their keys
    _, md = _compare(monkeypatch, 2, [obs, SparsePauliOp("IIIIIIIZ"), obs], slices, operator_budget=OperatorBudget(max_paulis=60))
    assert 0 < md.num_backpropagated_slices < 10


def test_sharded_round_trip_large(monkeypatch):
    """U then U^dag over 8 shards at ~5e5 terms returns Z_62 (no oracle at this size)."""
    from qiskit_addon_obp_b200.workloads import config2_kicked_ising
    from test_engine_gpu import _inverse_slices

    wl = config2_kicked_ising(theta_h=0.3, steps=6, max_paulis=None)
    monkeypatch.setenv("OBP_B200_SHARDS", "8")
    out, rest, md = backpropagate(wl.observables, _inverse_slices(wl.slices) + wl.slices)
    assert rest == []
    assert max(s.num_paulis[0] for s in md.backpropagation_history) > 100_000
    d = out.as_dict()
    key = wl.observables.paulis.to_labels()[0]
    assert abs(d[key] - 1.0) < 1e-5
    assert sum(abs(c) for k, c in d.items() if k != key) < 1e-2


def test_sharded_equals_single_table_large(monkeypatch):
    """8 shards and one table give the same 5e5-term operator (keys bit-exact, coefficients to 1e-12)."""
    from qiskit_addon_obp_b200.workloads import config2_kicked_ising

    wl = config2_kicked_ising(theta_h=0.3, steps=6, max_paulis=None)
    one, _, md1 = backpropagate(wl.observables, wl.slices)
    monkeypatch.setenv("OBP_B200_SHARDS", "8")
    many, _, md8 = backpropagate(wl.observables, wl.slices)
    assert len(one) == len(many) > 100_000
    assert [s.num_paulis for s in md1.backpropagation_history] == [s.num_paulis for s in md8.backpropagation_history]
    a, b = one.sorted_by_key(), many.sorted_by_key()
    assert np.array_equal(a.packed()[0], b.packed()[0]) and np.array_equal(a.packed()[1], b.packed()[1])
    assert np.max(np.abs(a.coeffs - b.coeffs)) < 1e-12


@pytest.mark.parametrize("ranks", [2, 8])
def test_sharded_reset_and_noise(monkeypatch, ranks):
    """reset moves the rows with Z on the qubit to the rank that owns the cleared key; damping is rank-local."""
    from qiskit_addon_obp_b200.ir import PauliLindbladError, PauliLindbladErrorInstruction, QuantumCircuit

    rng = np.random.default_rng(13 + ranks)
    nq = 7
    slices = random_slices(nq, 5, rng, gates_per_slice=6)
    ple = PauliLindbladError(["IXX", "ZYI", "IIZ"], [1e-2, 3e-3, 5e-2])
    noisy = QuantumCircuit(nq)
    noisy.append(PauliLindbladErrorInstruction(ple), [1, 3, 4])
    noisy.cx(1, 3)
    mixed = []
    for q in range(nq):  # a reset on every qubit: some keep their owner, most do not
        rst = QuantumCircuit(nq)
        rst.h((q + 1) % nq)
        rst.reset(q)
        mixed += [rst, slices[q % len(slices)]]
    last = QuantumCircuit(nq)
    last.rx(0.4, 2)
    last.reset(3)  # applied first (reverse order) ...
    first = QuantumCircuit(nq)
    first.reset(5)  # ... and a slice whose only (hence last) gate is the reset: its counters are reported
    obs = random_observable(nq, 6, rng, weight=3)
    _compare(monkeypatch, ranks, obs, slices + [noisy] + mixed + [last, first] + slices[:2])


@pytest.mark.parametrize("ranks", [2, 4])
def test_sharded_generic_gates(monkeypatch, ranks):
    """cp / crz / a random two-qubit unitary on a sharded table: gathered, expanded as one table, re-sharded;
    counters exact; a slice that breaks the operator budget after such a gate still rolls back to the commit."""
    from qiskit_addon_obp_b200.ir import QuantumCircuit

    rng = np.random.default_rng(77)
    nq = 8
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    q, r = np.linalg.qr(a)
    u = q * (np.diag(r) / np.abs(np.diag(r)))
    slices = []
    for k in range(5):
        qc = QuantumCircuit(nq)
        qc.rx(0.4 + 0.1 * k, k % nq)
        qc.cp(0.3 + 0.2 * k, k % nq, (k + 3) % nq)
        qc.cx((k + 1) % nq, (k + 2) % nq)
        if k == 2:
            qc.unitary(u, [(k + 4) % nq, (k + 6) % nq])
        qc.crz(-0.6, (k + 5) % nq, (k + 7) % nq)
        qc.ry(0.3, (k + 2) % nq)
        slices.append(qc)
    obs = random_observable(nq, 3, rng)
    _, md = _compare(monkeypatch, ranks, obs, slices)
    assert md.num_backpropagated_slices == 5
    _, md = _compare(monkeypatch, ranks, obs, slices, operator_budget=OperatorBudget(max_paulis=60))
    assert 0 < md.num_backpropagated_slices < 5
    _compare(monkeypatch, ranks, obs, slices, truncation_error_budget=setup_budget(max_error_per_slice=0.01))


@pytest.mark.parametrize("ranks", [2, 4])
def test_sharded_simplify_false(monkeypatch, ranks):
    """OperatorBudget(simplify=False) on a sharded table: every rank expands the rows it holds (k*k per gate),
    nothing merges; sizes, raw counts and the per-label sums match the oracle (row order is not defined)."""
    from qiskit_addon_obp_b200.ir import QuantumCircuit

    qc1 = QuantumCircuit(3)
    qc1.rx(0.3, 0)
    qc1.cx(0, 1)
    qc2 = QuantumCircuit(3)
    qc2.h(1)
    qc2.reset(2)
    qc2.rzz(0.2, 0, 1)
    obs = SparsePauliOp(["ZIX", "XYZ", "ZIX", "IYI", "XXI"], [1.0, 0.5, 0.25, -0.3, 0.7])
    budget = OperatorBudget(simplify=False)
    for kw in ({}, {"truncation_error_budget": setup_budget(max_error_per_slice=0.05)}):
        monkeypatch.setenv("OBP_B200_SHARDS", str(ranks))
        got, rest, md = backpropagate(obs, [qc1, qc2], operator_budget=budget, **kw)
        monkeypatch.delenv("OBP_B200_SHARDS")
        want, _, md_w = oracle.backpropagate(obs, [qc1, qc2], operator_budget=budget, **kw)
        assert rest == [] and len(got) == len(want)
        assert [s.num_paulis for s in md.backpropagation_history] == [s.num_paulis for s in md_w.backpropagation_history]
        assert [s.raw_num_paulis for s in md.backpropagation_history] == [s.raw_num_paulis for s in md_w.backpropagation_history]
        assert all(s.num_unique_paulis is None for s in md.backpropagation_history)
        if not kw:
            a, b = got.as_dict(), want.as_dict()  # sums per label
            assert set(a) == set(b)
            assert max(abs(a[k] - b[k]) for k in a) < 1e-12


@pytest.mark.parametrize("ranks", [2, 4])
def test_sharded_unsimplified_input(monkeypatch, ranks):
    """Duplicate and near-zero input rows reach the first gate as given: its counters (raw_num_paulis over every
    given row) match the oracle on a sharded table too."""
    from qiskit_addon_obp_b200.ir import QuantumCircuit

    qc = QuantumCircuit(3)
    qc.rx(0.3, 0)
    qc.cx(0, 1)
    qc2 = QuantumCircuit(3)
    qc2.rz(0.1, 1)
    qc2.ry(0.4, 2)
    obs = SparsePauliOp(["ZIX", "XYZ", "ZIX", "IIZ", "XYZ", "YYI"], [1.0, 0.5, 0.25, 1e-9, -0.2, 0.3])
    _, md = _compare(monkeypatch, ranks, obs, [qc])
    assert md.backpropagation_history[0].num_unique_paulis[0] is not None
    _compare(monkeypatch, ranks, obs, [qc2, qc])
    _compare(monkeypatch, ranks, obs, [qc, qc2])


def test_sharded_exact_counters_for_clifford_ending_slices(monkeypatch):
    """OBP_B200_SHARDED_EXACT_COUNTERS=1: slices that end on a Clifford-class gate report the reference's
    simplify counters too (computed on the gathered operator)."""
    rng = np.random.default_rng(19)
    nq = 7
    slices = random_slices(nq, 8, rng, p_rot=0.4, gates_per_slice=5)
    obs = random_observable(nq, 4, rng, weight=3)
    monkeypatch.setenv("OBP_B200_SHARDED_EXACT_COUNTERS", "1")
    monkeypatch.setenv("OBP_B200_SHARDS", "4")
    got, _, md_g = backpropagate(obs, slices)
    monkeypatch.delenv("OBP_B200_SHARDS")
    want, _, md_w = oracle.backpropagate(obs, slices)
    assert_same_operator(got, want)
    assert_same_metadata(md_g, md_w)  # no counter may be missing


# ---- late sharding: replicated while small, split at a commit or when a rank's copy overflows --------------
@pytest.mark.parametrize("plan", [False, True], ids=["hash-sharded", "planned"])
@pytest.mark.parametrize("min_terms", [40, 400, 10**9])
@pytest.mark.parametrize("ranks", [2, 8])
def test_late_sharding_splits_mid_run(monkeypatch, ranks, min_terms, plan):
    """The operator starts as an identical copy on every rank and is split into shards at the first commit at or
    above OBP_B200_SHARD_MIN_TERMS terms (never, for 10^9): same operator and metadata as the oracle either way,
    and — unlike a table sharded from the start — exact simplify counters for the slices run replicated."""
    from qiskit_addon_obp_b200 import backpropagation as bp

    monkeypatch.setenv("OBP_B200_SHARD_MIN_TERMS", str(min_terms))
    # planned: the shard function is drawn orthogonal to the rotations still to come — no exchange at all
    monkeypatch.setenv("OBP_B200_SHARD_PLAN", "1" if plan else "0")
    from test_engine_gpu import window_workload

    obs, slices = window_workload(70, 4)
    got, md = _compare(monkeypatch, ranks, obs, slices)
    sizes = [s.num_paulis[0] for s in md.backpropagation_history]
    assert sizes[-1] > 400, sizes
    if min_terms == 10**9:
        assert bp.last_run.exchanges == 0
        assert all(s.num_unique_paulis[0] is not None for s in md.backpropagation_history)
    elif not plan:
        assert bp.last_run.exchanges > 0
    # (planned: how many rotations the plan covers depends on how well it balances this periodic circuit, see
    #  test_planned_sharding_balance_and_fallback)
    _compare(monkeypatch, ranks, obs, slices, truncation_error_budget=setup_budget(max_error_per_slice=2e-3))


def test_late_sharding_split_on_overflow(monkeypatch):
    """A replicated copy that outgrows one rank's table is rolled back and split instead of grown."""
    from qiskit_addon_obp_b200 import backpropagation as bp
    from qiskit_addon_obp_b200 import engine, sharded

    sharded.close_all()
    engine.clear_pool()
    monkeypatch.setenv("OBP_B200_SHARD_PLAN", "0")
    monkeypatch.setenv("OBP_B200_SHARD_MIN_TERMS", "100000")  # never reached by size ...
    monkeypatch.setenv("OBP_B200_CAPACITY", "70000")          # ... but a rank holds at most max(2^16, ...) rows
    from test_engine_gpu import window_workload

    obs, slices = window_workload(70, 5)
    got, md = _compare(monkeypatch, 4, obs, slices)
    assert max(s.num_paulis[0] for s in md.backpropagation_history) > 2**16
    assert bp.last_run.exchanges > 0 and bp.last_run.wasted_updates > 0
    sharded.close_all()
    engine.clear_pool()


def test_planned_sharding_balance_and_fallback(monkeypatch):
    """The planned shard function: (1) on a near-Clifford brickwork (config 4: sparse T gates, fresh generators) it
    covers every rotation to come — no exchange at all — and still spreads the rows like a hash; (2) on a periodic
    Trotter-like circuit, whose future generators span the same space as the ones that made the rows, a full plan
    would put every row on one rank: the planner shortens its horizon until the split is balanced and the rotations
    outside the plan exchange records.  Same results as the oracle either way."""
    from qiskit_addon_obp_b200 import backpropagation as bp
    from qiskit_addon_obp_b200 import sharded
    from qiskit_addon_obp_b200.workloads import config4_near_clifford
    from test_engine_gpu import window_workload

    seen = {}
    orig = sharded.ShardedTable._split_now

    def spy(self):
        orig(self)
        seen["rows"] = [t.n for t in self.tables.values()]
        seen["planned"] = self.planned_rotations

    monkeypatch.setattr(sharded.ShardedTable, "_split_now", spy)
    # 1. near-Clifford brickwork, 5 % T gates: 4 ranks, split at >= 200 terms (the table then grows to 1362)
    monkeypatch.setenv("OBP_B200_SHARD_MIN_TERMS", "200")
    wl = config4_near_clifford(num_qubits=200, depth=24, t_prob=0.05, max_paulis=20_000, n_observable_terms=4)
    _, md = _compare(monkeypatch, 4, wl.observables, wl.slices, **wl.kwargs())
    assert seen and seen["planned"] > 0, seen
    assert bp.last_run.exchanges == 0, "every rotation after the split was in the plan"
    mean = sum(seen["rows"]) / 4
    assert max(seen["rows"]) <= 1.6 * mean + 40, seen
    assert max(s.num_paulis[0] for s in md.backpropagation_history) > 4 * 200
    # 2. periodic window circuit: balance forces a short (possibly empty) plan, exchanges happen
    seen.clear()
    monkeypatch.setenv("OBP_B200_SHARD_MIN_TERMS", "3000")
    obs, slices = window_workload(70, 4)
    _compare(monkeypatch, 4, obs, slices)
    mean = sum(seen["rows"]) / 4
    assert max(seen["rows"]) <= 1.6 * mean + 40, seen
    assert bp.last_run.exchanges > 0
