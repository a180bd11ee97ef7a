"""Number of qubit-wise commuting groups (host side).

The reference calls ``len(op.group_commuting(qubit_wise=True))`` ([EXT] Qiskit + rustworkx,
``backpropagation.py:369,434``): build the graph whose edges join Paulis that do *not* commute
qubit-wise and colour it greedily, largest degree first (stable).  This stays on the host like
the reference's own graph colouring; it is O(n^2) and only meant for the small operators for which
``max_qwc_groups`` is practical.  Row order follows the engine's table order, so on operators where
the greedy colouring is order-sensitive the count can differ from Qiskit's.
"""

from __future__ import annotations

import numpy as np

from ..ir import SparsePauliOp


def qwc_group_count(op: SparsePauliOp) -> int:
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
