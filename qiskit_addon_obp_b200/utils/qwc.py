"""Number of qubit-wise commuting groups (host side).

The reference calls ``len(op.group_commuting(qubit_wise=True))`` ([EXT] Qiskit + rustworkx,
``backpropagation.py:369,434``): build the graph whose edges join Paulis that do *not* commute
qubit-wise and colour it greedily, largest degree first (stable).  This stays on the host like
the reference's own graph colouring; it is O(n^2) (on packed 64-qubit words) and meant for the operators
for which ``max_qwc_groups`` is practical (up to ~10^5 terms).  Row order follows the engine's table order, so on operators where
the greedy colouring is order-sensitive the count can differ from Qiskit's.
"""

from __future__ import annotations

import numpy as np

from ..ir import SparsePauliOp


def qwc_group_count(op: SparsePauliOp) -> int:
    """Greedy colouring (largest degree first, stable) of the graph joining Paulis that do not commute qubit-wise.

    Two Paulis clash iff on some qubit both are non-identity and different.  On the packed words that is
    ``(x_i | z_i) & (x_j | z_j) & ((x_i ^ x_j) | (z_i ^ z_j)) != 0`` — 64 qubits per operation instead of
    one, which keeps the O(n^2) pass usable up to ~10^5 terms.
    """
    n = len(op)
    if n == 0:
        return 0
    xw, zw = op.packed()  # [n, W] uint64
    nzw = xw | zw
    adj = []
    degree = np.zeros(n, dtype=np.int64)
    block = max(1, (1 << 22) // max(1, n * xw.shape[1]))  # rows per block: ~32 MB of temporaries
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        both = nzw[i0:i1, None, :] & nzw[None, :, :]
        diff = (xw[i0:i1, None, :] ^ xw[None, :, :]) | (zw[i0:i1, None, :] ^ zw[None, :, :])
        clash = ((both & diff) != 0).any(axis=2)
        for r in range(i1 - i0):
            nb = np.flatnonzero(clash[r])
            adj.append(nb)
            degree[i0 + r] = len(nb)
    order = np.argsort(-degree, kind="stable")
    colour = np.full(n, -1, dtype=np.int64)
    for v in order:
        used = set(colour[adj[v]].tolist())
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    return int(colour.max()) + 1


def _qwc_group_count_bool(op: SparsePauliOp) -> int:
    """The same count from the unpacked bool arrays (the first implementation; kept as the test reference)."""
    n = len(op)
    if n == 0:
        return 0
    code = op.x.astype(np.int8) + 2 * op.z.astype(np.int8)
    nz = code != 0
    degree = np.zeros(n, dtype=np.int64)
    adj = []
    for i in range(n):
        clash = np.any(nz[i] & nz & (code != code[i]), axis=1)
        adj.append(np.flatnonzero(clash))
        degree[i] = len(adj[-1])
    order = np.argsort(-degree, kind="stable")
    colour = np.full(n, -1, dtype=np.int64)
    for v in order:
        used = set(colour[adj[v]].tolist())
        c = 0
        while c in used:
            c += 1
        colour[v] = c
    return int(colour.max()) + 1
