"""``qiskit_adapter`` executed against a stub ``qiskit`` package (tests/stubs/qiskit).

CPU: conversion of Qiskit-shaped observables and circuits to the host IR and back.
GPU: ``backpropagate`` called with Qiskit-shaped inputs returns Qiskit-shaped operators, the caller's own
slice objects and the same numbers as the oracle (``backpropagation.py:48-55, 300, 307-309``)."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


@pytest.fixture
def qiskit_stub(monkeypatch):
    monkeypatch.syspath_prepend(STUBS)
    for m in [m for m in sys.modules if m == "qiskit" or m.startswith("qiskit.")]:
        monkeypatch.delitem(sys.modules, m)
    import qiskit
    import qiskit.quantum_info

    assert qiskit.__version__.endswith("stub")
    yield qiskit
    for m in [m for m in sys.modules if m == "qiskit" or m.startswith("qiskit.")]:
        sys.modules.pop(m, None)


def _circuits(qiskit):
    from qiskit_addon_obp_b200 import ir

    n = 5
    rng = np.random.default_rng(3)
    a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    u, _ = np.linalg.qr(a)
    both = []
    for cls in (qiskit.QuantumCircuit, ir.QuantumCircuit):
        s1 = cls(n)
        s1.rx(0.3, 2)
        s1.cx(2, 3)
        s1.rzz(0.4, 0, 1)
        s1.barrier()
        s1.h(4)
        s2 = cls(n)
        s2.ecr(3, 1)
        s2.unitary(u, [4, 2])
        s2.t(0)
        s3 = cls(n)
        s3.reset(1)
        s3.sx(2)
        both.append([s1, s2, s3])
    return both


def test_adapter_converts_qiskit_shaped_inputs(qiskit_stub):
    from qiskit.quantum_info import SparsePauliOp as QOp

    from qiskit_addon_obp_b200 import ir, qiskit_adapter

    q_slices, ir_slices = _circuits(qiskit_stub)
    q_obs = [QOp(["IIZII", "XYIII"], [1.0, 0.5 - 0.25j]), QOp("IIIIZ")]
    obs_ir, slices_ir, back = qiskit_adapter.to_ir(q_obs, q_slices)
    assert all(isinstance(o, ir.SparsePauliOp) for o in obs_ir) and all(isinstance(s, ir.QuantumCircuit) for s in slices_ir)
    assert obs_ir[0].to_list() == [("IIZII", 1.0 + 0j), ("XYIII", 0.5 - 0.25j)]
    for got, want in zip(slices_ir, ir_slices):
        assert [(g.name, g.qubits) for g in got.data] == [(w.name, w.qubits) for w in want.data]
        for g, w in zip(got.data, want.data):
            if g.name not in ("barrier", "reset"):
                np.testing.assert_allclose(g.matrix, w.matrix, atol=1e-15)
    out = back(obs_ir[0])
    assert isinstance(out, QOp) and out.to_list() == q_obs[0].to_list()
    # a single observable stays single; IR inputs pass straight through
    single, _, _ = qiskit_adapter.to_ir(q_obs[0], q_slices)
    assert isinstance(single, ir.SparsePauliOp)
    same, same_slices, ident = qiskit_adapter.to_ir(obs_ir, ir_slices)
    assert same is obs_ir and ident(obs_ir[0]) is obs_ir[0]
    # object-dtype (Parameter) coefficients are rejected, not silently cast (utils/simplify.py:108-118 is a host-only path)
    bad = QOp(["IIZII"])
    bad.coeffs = np.array([object()], dtype=object)
    with pytest.raises(TypeError):
        qiskit_adapter.to_ir(bad, q_slices)


@pytest.mark.gpu
def test_backpropagate_with_qiskit_shaped_objects(qiskit_stub):
    from qiskit.quantum_info import SparsePauliOp as QOp

    from conftest import assert_same_metadata, assert_same_operator
    from oracle import obp_oracle as oracle
    from qiskit_addon_obp_b200 import backpropagate, ir
    from qiskit_addon_obp_b200.utils.simplify import OperatorBudget

    q_slices, ir_slices = _circuits(qiskit_stub)
    q_obs = [QOp(["IIZII", "XYIII"], [1.0, 0.5]), QOp("IIIIZ")]
    ir_obs = [ir.SparsePauliOp(["IIZII", "XYIII"], [1.0, 0.5]), ir.SparsePauliOp("IIIIZ")]
    got, rest, md = backpropagate(q_obs, q_slices, operator_budget=OperatorBudget(max_paulis=30))
    want, rest_w, md_w = oracle.backpropagate(ir_obs, ir_slices, operator_budget=OperatorBudget(max_paulis=30))
    assert all(isinstance(o, QOp) for o in got)
    assert len(rest) == len(rest_w) and all(a is b for a, b in zip(rest, q_slices))  # the caller's own objects (:300)
    for g, w in zip(got, want):
        assert_same_operator(ir.SparsePauliOp((g.paulis.x, g.paulis.z), g.coeffs), w)
    assert_same_metadata(md, md_w)
    one, _, _ = backpropagate(q_obs[0], q_slices)
    assert isinstance(one, QOp)
