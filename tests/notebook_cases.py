"""The executed notebook outputs of the reference as end-to-end anchors.

``docs/guides/truncate_operator_terms.ipynb`` and ``docs/guides/bound_error_using_p_norm.ipynb`` of
``/root/reference`` print, for a 10-qubit XY chain, how many slices were absorbed, how many Pauli
terms the observable ends with, how many qubit-wise commuting groups those form and (p-norm guide)
the expectation value of the reduced circuit to 16 digits.  These are the only reference-held
vectors above 5 qubits; they pin the circuit generator conventions, compose/simplify, the truncation
search, the budget replay and the QWC count at once.

The circuit is restated without ``qiskit_addon_utils`` (not installable here):
``generate_xyz_hamiltonian(path_graph(10), (0.05, 0.02, 0), (0.02, 0.08, 0), InteractionThenColor)``
+ ``LieTrotter(reps=3)``, ``time=2.0`` + ``slice_by_gate_types`` = per repetition the six slices
RXX(even bonds), RXX(odd bonds), RYY(even), RYY(odd), RX(all), RY(all) with angle ``2*c*dt``
(``truncate_operator_terms.ipynb`` cell 4; "six slices that are repeated three times", cell 21).

One cell (28, ``max_error_total=0.006``) prints 67 terms / 20 groups; with the current default
``tol=1e-8`` the algorithm gives 65 / 19, with the pre-0.2 default ``tol=1e-10``
(``releasenotes/notes/0.2/trunc-tol-a1c03cf32154e58f.yaml``) it gives exactly 67 / 20: that cell's
stored output predates the tolerance change.  Both facts are asserted.
"""

from __future__ import annotations

import numpy as np

from qiskit_addon_obp_b200.ir import QuantumCircuit, SparsePauliOp

N = 10


def slices(n: int = N, reps: int = 3, time: float = 2.0) -> list[QuantumCircuit]:
    dt = time / reps
    even = [(i, i + 1) for i in range(0, n - 1, 2)]
    odd = [(i, i + 1) for i in range(1, n - 1, 2)]
    out = []
    for _ in range(reps):
        for name, c in (("rxx", 0.05), ("ryy", 0.02)):
            for bonds in (even, odd):
                qc = QuantumCircuit(n)
                for a, b in bonds:
                    getattr(qc, name)(2 * c * dt, a, b)
                out.append(qc)
        for name, c in (("rx", 0.02), ("ry", 0.08)):
            qc = QuantumCircuit(n)
            for q in range(n):
                getattr(qc, name)(2 * c * dt, q)
            out.append(qc)
    return out


def z5() -> SparsePauliOp:
    return SparsePauliOp(["IIIIIZIIII"], [1.0])


def x5() -> SparsePauliOp:
    return SparsePauliOp(["IIIIIXIIII"], [1.0])


def sum_z(n: int = N) -> SparsePauliOp:
    return SparsePauliOp.from_sparse_list([("Z", [i], 1.0) for i in range(n)], num_qubits=n)


def _tb(budget_cls, tol=None, **kw):
    """``setup_budget`` (utils/truncating.py:114-141) into whichever budget dataclass the caller uses."""
    per_slice, total, num_slices = kw.get("max_error_per_slice"), kw.get("max_error_total"), kw.get("num_slices")
    if per_slice is not None:
        grants = [per_slice] if isinstance(per_slice, float) else list(per_slice)
    elif num_slices is not None:
        grants = [total / num_slices]
    else:
        grants = [total]
    b = budget_cls(grants, np.inf if total is None else total, kw.get("p_norm", 1))
    if tol is not None:
        b.tol = tol
    return b


# (cell, observables, slice window, max_qwc_groups, setup_budget kwargs, tol override,
#  expected slices absorbed, expected [(terms, groups)] per observable)
TRUNCATE_CASES = [
    ("cell10", "z", None, 10, dict(max_error_per_slice=0.001), None, 11, [(29, 10)]),
    ("cell17", "z", None, 10, dict(max_error_per_slice=[0.0] * 3 + [0.001] * 15), None, 11, [(32, 10)]),
    ("cell23", "z", None, 20, dict(max_error_per_slice=[0.0] * 4 + [0.003] * 2), None, 13, [(49, 14)]),
    ("cell28_tol1e-10", "z", None, 20, dict(max_error_per_slice=[0.0] * 4 + [0.003] * 2, max_error_total=0.006), 1e-10, 10, [(67, 20)]),
    ("cell28_tol1e-8", "z", None, 20, dict(max_error_per_slice=[0.0] * 4 + [0.003] * 2, max_error_total=0.006), None, 10, [(65, 19)]),
    ("cell34", "z", None, 10, dict(max_error_total=0.018), None, 9, [(25, 9)]),
    ("cell42", "z", 12, 10, dict(max_error_total=0.018, num_slices=12, p_norm=1), None, 12, [(29, 9)]),
    ("cell49", "zx", None, 10, dict(max_error_per_slice=0.001), None, 11, [(29, 10), (23, 8)]),
]

# bound_error_using_p_norm.ipynb cells 8, 12, 14, 20, 22
EXACT_EXPECTATION = 9.318197859862146
PNORM_CASES = [
    ("l1", 1, (116, 10), 9.317869899338842),
    ("l2", 2, (84, 6), 9.317829770853422),
]


def run_truncate_case(case, backpropagate, operator_budget_cls, trunc_budget_cls, group_count):
    """Returns ``(slices absorbed, [(terms, groups)], metadata)`` for one TRUNCATE_CASES entry."""
    _, obs_kind, window, max_qwc, kw, tol, _, _ = case
    s = slices()
    if window is not None:
        s = s[-window:]
    obs = z5() if obs_kind == "z" else [z5(), x5()]
    out, rest, md = backpropagate(
        obs, s, operator_budget=operator_budget_cls(max_qwc_groups=max_qwc),
        truncation_error_budget=_tb(trunc_budget_cls, tol, **kw),
    )
    outs = out if isinstance(out, list) else [out]
    return len(s) - len(rest), [(len(o), group_count(o)) for o in outs], md


def expectation_on_zero_state(circuit_slices, op: SparsePauliOp, gate_matrix) -> float:
    """<0| U^dag O U |0> with U = the given slices in order, by a 2^n statevector (n = 10)."""
    n = op.num_qubits
    psi = np.zeros(2**n, dtype=complex)
    psi[0] = 1.0
    psi = psi.reshape([2] * n)  # axis k = qubit n-1-k (qubit 0 is the least-significant index bit)
    for qc in circuit_slices:
        for ins in qc.data:
            m = gate_matrix(ins.name, ins.params)
            k = len(ins.qubits)
            axes = [n - 1 - q for q in ins.qubits]  # gate qubit j <-> tensor axis
            # gate matrix index = sum_j bit_j << j: reshape to [out_{k-1}..out_0, in_{k-1}..in_0]
            g = m.reshape([2] * (2 * k))
            in_axes = list(range(2 * k - 1, k - 1, -1))  # in_0 is the last axis
            psi = np.tensordot(g, psi, axes=(in_axes, axes))
            # result axes: out_{k-1}..out_0 then the untouched ones; move out_j back to axis of qubit j
            psi = np.moveaxis(psi, list(range(k)), [axes[k - 1 - a] for a in range(k)])
    psi = psi.reshape(-1)
    total = 0.0
    idx = np.arange(2**n)
    for xr, zr, c in zip(op.x, op.z, op.coeffs):
        xm = int(sum(1 << q for q in range(n) if xr[q]))
        zm = int(sum(1 << q for q in range(n) if zr[q]))
        ny = int(np.sum(xr & zr))
        # P|b> = i^{ny} (-1)^{popcount(b & zm)} |b ^ xm>
        par = np.zeros(2**n, dtype=np.int64)
        v = idx & zm
        while np.any(v):
            par ^= v & 1
            v = v >> 1
        phase = (1j) ** ny * (1 - 2 * par)
        total += (c * np.vdot(psi[idx ^ xm], phase * psi)).real
    return float(total)
