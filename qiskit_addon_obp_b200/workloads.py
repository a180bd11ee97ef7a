"""Synthetic circuits and observables of the five BASELINE.json configurations.

The reference ships no benchmark inputs; these generators restate the named workloads with the
neutral IR so that the engine, the oracle and ``bench.py`` all run on identical inputs.  Every
generator is deterministic (seeded where random) and returns a :class:`Workload`.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .ir import QuantumCircuit, SparsePauliOp, slice_by_depth
from .utils.simplify import OperatorBudget
from .utils.truncating import TruncationErrorBudget, setup_budget


@dataclass
class Workload:
    name: str
    observables: SparsePauliOp | list[SparsePauliOp]
    slices: list[QuantumCircuit]
    operator_budget: OperatorBudget = field(default_factory=OperatorBudget)
    truncation_error_budget: TruncationErrorBudget = field(default_factory=TruncationErrorBudget)

    @property
    def num_qubits(self) -> int:
        return self.slices[0].num_qubits

    def kwargs(self) -> dict:
        return {
            "operator_budget": self.operator_budget,
            "truncation_error_budget": self.truncation_error_budget,
        }


def _z_on(num_qubits: int, qubits, coeff: float = 1.0) -> SparsePauliOp:
    qubits = list(qubits)
    x = np.zeros((len(qubits), num_qubits), dtype=bool)
    z = np.zeros((len(qubits), num_qubits), dtype=bool)
    for i, q in enumerate(qubits):
        z[i, q] = True
    return SparsePauliOp((x, z), np.full(len(qubits), coeff, dtype=complex))


# ---------------------------------------------------------------------------------------------
# config 1: 10-qubit XY-model Trotter circuit (docs/guides/truncate_operator_terms.ipynb:80-112)
# ---------------------------------------------------------------------------------------------
def config1_xy_chain(num_qubits: int = 10, reps: int = 3, time: float = 2.0, max_paulis: int | None = 10_000,
                     per_slice: float = 1e-3) -> Workload:
    dt = time / reps
    qc = QuantumCircuit(num_qubits)
    bonds = [(i, i + 1) for i in range(0, num_qubits - 1, 2)] + [(i, i + 1) for i in range(1, num_qubits - 1, 2)]
    for _ in range(reps):
        for a, b in bonds:
            qc.rxx(2 * 0.05 * dt, a, b)
        for a, b in bonds:
            qc.ryy(2 * 0.02 * dt, a, b)
        for q in range(num_qubits):
            qc.rx(2 * 0.02 * dt, q)
        for q in range(num_qubits):
            qc.ry(2 * 0.08 * dt, q)
    slices = slice_by_depth(qc, 1)
    obs = _z_on(num_qubits, [num_qubits // 2 - 1])  # "IIIIIZIIII": Z on qubit 4
    return Workload(
        "config1: 10q XY chain, Z_4, 1e-3/slice, max_paulis=1e4",
        obs,
        slices,
        OperatorBudget(max_paulis=max_paulis),
        TruncationErrorBudget([per_slice], np.inf, 1, 1e-8),
    )


# ---------------------------------------------------------------------------------------------
# config 2: 127-qubit heavy-hex kicked Ising
# ---------------------------------------------------------------------------------------------
def heavy_hex_127_edges() -> list[tuple[int, int]]:
    rows = [range(0, 14), range(18, 33), range(37, 52), range(56, 71), range(75, 90), range(94, 109), range(113, 127)]
    edges = []
    for r in rows:
        r = list(r)
        edges += [(r[i], r[i + 1]) for i in range(len(r) - 1)]
    bridges = {
        14: (0, 18), 15: (4, 22), 16: (8, 26), 17: (12, 30),
        33: (20, 39), 34: (24, 43), 35: (28, 47), 36: (32, 51),
        52: (37, 56), 53: (41, 60), 54: (45, 64), 55: (49, 68),
        71: (58, 77), 72: (62, 81), 73: (66, 85), 74: (70, 89),
        90: (75, 94), 91: (79, 98), 92: (83, 102), 93: (87, 106),
        109: (96, 114), 110: (100, 118), 111: (104, 122), 112: (108, 126),
    }  # fmt: skip
    for b, (u, v) in bridges.items():
        edges += [(u, b), (b, v)]
    assert len(edges) == 144
    return edges


def edge_colouring(edges: list[tuple[int, int]]) -> list[list[tuple[int, int]]]:
    """Proper edge colouring of a bipartite graph with max-degree many colours (alternating paths)."""
    deg: dict[int, int] = {}
    for u, v in edges:
        deg[u] = deg.get(u, 0) + 1
        deg[v] = deg.get(v, 0) + 1
    ncol = max(deg.values())
    at: dict[int, dict[int, int]] = {n: {} for n in deg}  # node -> colour -> neighbour

    def free(n):
        return next(c for c in range(ncol) if c not in at[n])

    for u, v in edges:
        a, b = free(u), free(v)
        if a != b:
            # flip the a/b alternating path starting at v so that a becomes free at v
            path = []
            node, col = v, a
            while col in at[node]:
                nxt = at[node][col]
                path.append((node, nxt, col))
                node, col = nxt, (b if col == a else a)
            for x, y, c in path:
                del at[x][c]
                del at[y][c]
            for x, y, c in path:
                nc = b if c == a else a
                at[x][nc] = y
                at[y][nc] = x
        at[u][a] = v
        at[v][a] = u
    layers: list[list[tuple[int, int]]] = [[] for _ in range(ncol)]
    seen = set()
    for n, cols in at.items():
        for c, m in cols.items():
            e = (min(n, m), max(n, m))
            if e not in seen:
                seen.add(e)
                layers[c].append(e)
    for lay in layers:
        lay.sort()
        nodes = [q for e in lay for q in e]
        assert len(nodes) == len(set(nodes))
    assert sum(len(lay) for lay in layers) == len(edges)
    return layers


def config2_kicked_ising(theta_h: float = np.pi / 4, steps: int = 20, max_paulis: int | None = 1_000_000,
                         observable_qubit: int = 62) -> Workload:
    edges = heavy_hex_127_edges()
    layers = edge_colouring(edges)
    n = 127
    slices = []
    for _ in range(steps):
        for lay in layers:
            qc = QuantumCircuit(n)
            for a, b in lay:
                qc.rzz(-np.pi / 2, a, b)
            slices.append(qc)
        qc = QuantumCircuit(n)
        for q in range(n):
            qc.rx(theta_h, q)
        slices.append(qc)
    return Workload(
        f"config2: 127q heavy-hex kicked Ising, theta_h={theta_h:.4f}, Z_{observable_qubit}, {steps} steps, "
        f"max_paulis={max_paulis}",
        _z_on(n, [observable_qubit]),
        slices,
        OperatorBudget(max_paulis=max_paulis),
    )


# ---------------------------------------------------------------------------------------------
# config 3: L x L square-lattice transverse-field Ising Trotter evolution, sum-of-Z observable
# ---------------------------------------------------------------------------------------------
def config3_square_lattice(side: int = 10, steps: int = 10, dt: float = 0.1, coupling: float = 1.0, field: float = 1.0,
                           max_error_total: float = 1e-2, max_paulis: int | None = 1_000_000) -> Workload:
    n = side * side
    idx = lambda r, c: r * side + c  # noqa: E731
    colours = [
        [(idx(r, c), idx(r, c + 1)) for r in range(side) for c in range(0, side - 1, 2)],
        [(idx(r, c), idx(r, c + 1)) for r in range(side) for c in range(1, side - 1, 2)],
        [(idx(r, c), idx(r + 1, c)) for c in range(side) for r in range(0, side - 1, 2)],
        [(idx(r, c), idx(r + 1, c)) for c in range(side) for r in range(1, side - 1, 2)],
    ]
    slices = []
    for _ in range(steps):
        for lay in colours:
            qc = QuantumCircuit(n)
            for a, b in lay:
                qc.rzz(2 * coupling * dt, a, b)
            slices.append(qc)
        qc = QuantumCircuit(n)
        for q in range(n):
            qc.rx(2 * field * dt, q)
        slices.append(qc)
    return Workload(
        f"config3: {side}x{side} TFIM Trotter, sum-of-Z, total error {max_error_total}, max_paulis={max_paulis}",
        _z_on(n, range(n)),
        slices,
        OperatorBudget(max_paulis=max_paulis),
        setup_budget(max_error_total=max_error_total, num_slices=len(slices), p_norm=1),
    )


# ---------------------------------------------------------------------------------------------
# configs 4 and 5: random Clifford brickwork with sparse / dense non-Clifford rotations
# ---------------------------------------------------------------------------------------------
_ONE_Q = [(), ("h",), ("s",), ("s", "h"), ("h", "s"), ("h", "s", "h")]  # I, H, S, HS, SH, HSH (circuit order)
_TWO_Q = ["cx", "cz", "swap"]


def _brickwork(num_qubits: int, depth: int, rng: np.random.Generator, rot_prob: float, theta: float | None):
    """One slice per brickwork layer: random 1q Cliffords, a random 2q Clifford per bond, then T or RZ(theta)."""
    slices = []
    for layer in range(depth):
        qc = QuantumCircuit(num_qubits)
        for i in range(layer % 2, num_qubits - 1, 2):
            for q in (i, i + 1):
                for g in _ONE_Q[rng.integers(len(_ONE_Q))]:
                    getattr(qc, g)(q)
            getattr(qc, _TWO_Q[rng.integers(len(_TWO_Q))])(i, i + 1)
        hit = rng.random(num_qubits) < rot_prob
        for q in np.flatnonzero(hit):
            if theta is None:
                qc.t(int(q))
            else:
                qc.rz(theta, int(q))
        slices.append(qc)
    return slices


def config4_near_clifford(num_qubits: int = 1000, depth: int = 40, t_prob: float = 0.01, max_paulis: int | None = 100_000_000,
                          n_observable_terms: int = 8, per_slice: float | None = None, seed: int = 20261019) -> Workload:
    rng = np.random.default_rng(seed)
    slices = _brickwork(num_qubits, depth, rng, t_prob, None)
    qs = [int(round((k + 0.5) * num_qubits / n_observable_terms)) for k in range(n_observable_terms)]
    trunc = TruncationErrorBudget() if per_slice is None else TruncationErrorBudget([per_slice], np.inf, 1, 1e-8)
    return Workload(
        f"config4: {num_qubits}q near-Clifford brickwork depth {depth}, {t_prob:.0%} T, max_paulis={max_paulis}",
        _z_on(num_qubits, qs),
        slices,
        OperatorBudget(max_paulis=max_paulis),
        trunc,
    )


def config5_magic_sweep(theta: float = np.pi / 16, num_qubits: int = 64, depth: int = 40, max_paulis: int | None = 1_000_000,
                        seed: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    slices = _brickwork(num_qubits, depth, rng, 1.1, theta)  # RZ(theta) on every qubit after every layer
    return Workload(
        f"config5: {num_qubits}q Clifford+RZ({theta:.4f}) brickwork depth {depth}, max_paulis={max_paulis}, seed {seed}",
        _z_on(num_qubits, [num_qubits // 2]),
        slices,
        OperatorBudget(max_paulis=max_paulis),
    )
