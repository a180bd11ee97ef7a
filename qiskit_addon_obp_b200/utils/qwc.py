"""Number of qubit-wise commuting groups.

The reference calls ``len(op.group_commuting(qubit_wise=True))`` ([EXT] Qiskit + rustworkx,
``backpropagation.py:369,434``): build the graph whose edges join Paulis that do *not* commute
qubit-wise and colour it greedily, largest degree first, ties in row order.

Operators of ``OBP_B200_QWC_DEVICE_MIN`` (default 256) terms and more are counted on the GPU
(``obp_qwc_group_count``: all n^2 pair tests on packed words across the whole device, then the sequential
colouring in one CTA; ``csrc/obp_qwc.cuh``) and there is no CPU fallback for them.  Smaller ones — the input
validation of ``backpropagate`` sees a handful of terms before any device table exists — are counted right
here with the same algorithm on the same packed words.

Row order: ties of the colouring are broken by the order of the rows handed in.  In row-order mode
(``OBP_B200_ROW_ORDER=1``) that is exactly the reference's order; on the fast path it is the table's order, so
on an operator whose greedy colouring is order-sensitive the count can differ from Qiskit's (the reference's
notebook outputs — 10/10/14/20/9/9/10+8 and 10/6 groups — reproduce in both modes,
``tests/test_notebook_anchors.py``).
"""

from __future__ import annotations

import numpy as np

from ..ir import SparsePauliOp


def qwc_group_count(op: SparsePauliOp) -> int:
    """``len(op.group_commuting(qubit_wise=True))``: on the device from ``OBP_B200_QWC_DEVICE_MIN`` terms up."""
    import os

    if len(op) >= int(os.environ.get("OBP_B200_QWC_DEVICE_MIN", "256")):
        from .. import engine

        return engine.qwc_group_count(op)
    return _qwc_group_count_host(op)


def _qwc_group_count_host(op: SparsePauliOp) -> int:
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
