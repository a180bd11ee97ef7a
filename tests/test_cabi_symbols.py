"""The C-ABI library loads and exports every symbol ``include/obp_b200.h`` declares (no GPU calls)."""

from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "obp_b200.h")
LIB = os.path.join(ROOT, "qiskit_addon_obp_b200", "libobp_b200.so")


def declared_functions() -> list[str]:
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(obp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g

        g.build()
    return ctypes.CDLL(LIB)


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("obp_create", "obp_destroy", "obp_load", "obp_download", "obp_apply_ops", "obp_truncate", "obp_snapshot", "obp_restore"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in obp_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from qiskit_addon_obp_b200 import engine

    assert sorted(engine.EXPORTED_SYMBOLS) == declared_functions()


def test_op_record_layout_matches_header():
    """``obp_op`` is 104 bytes with the documented field offsets (numpy mirror in engine.py)."""
    from qiskit_addon_obp_b200.engine import OP_DTYPE, ObpProfile, ObpStats

    assert OP_DTYPE.itemsize == 104
    offs = {n: OP_DTYPE.fields[n][1] for n in OP_DTYPE.names}
    assert offs == {
        "kind": 0, "n_qubits": 4, "qubit": 8, "flags": 24, "n_components": 28, "lut": 32, "sign_mask": 40,
        "group_size": 44, "group_elems": 48, "row_scale": 56, "gen_local": 64, "reserved": 68, "u0": 72, "u1": 88,
    }  # fmt: skip
    assert ctypes.sizeof(ObpStats) == 80
    assert ctypes.sizeof(ObpProfile) == 8 * 8 * 4  # OBP_K_COUNT classes x four arrays


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the engine fails loudly instead of computing on the host."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    from qiskit_addon_obp_b200 import backpropagate
    from qiskit_addon_obp_b200.engine import EngineError
    from qiskit_addon_obp_b200.ir import QuantumCircuit, SparsePauliOp

    qc = QuantumCircuit(1)
    qc.rx(0.3, 0)
    with pytest.raises((EngineError, MemoryError)):
        backpropagate(SparsePauliOp("Z"), [qc])
    assert np.isfinite(1.0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it."""
    import re

    pkg = os.path.join(ROOT, "qiskit_addon_obp_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "obp_oracle" in text:
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert offenders == []


def test_option_and_class_constants_match_the_header():
    """``OBP_OPT_*`` / ``OBP_K_*`` / ``OBP_FLAG_*`` in engine.py are the header's values."""
    import re

    from qiskit_addon_obp_b200 import engine

    text = open(HEADER).read()
    defines = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(OBP_(?:OPT|FLAG|OP)_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)u?\b", text)}
    assert {"OBP_OPT_PROGRAM", "OBP_OPT_SEGMENTS", "OBP_OPT_SEGMENT_MIXED"} <= set(defines)
    for name, value in defines.items():
        if hasattr(engine, name):
            assert getattr(engine, name) == value, name
    for name in ("OBP_OPT_PROGRAM", "OBP_OPT_PROGRAM_MAX_ROWS", "OBP_OPT_SEGMENTS", "OBP_OPT_SEGMENT_MIN_GATES",
                 "OBP_OPT_SEGMENT_CAP_ROWS", "OBP_OPT_SEGMENT_PIECES_MIN_ROWS", "OBP_OPT_SEGMENT_MIXED"):
        assert getattr(engine, name) == defines[name], name
    enum = re.search(r"enum \{ OBP_K_ROTATE = 0,(.*?)\};", text, re.S).group(0)
    classes = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"OBP_K_([A-Z]+) = (\d+)", enum)}
    count = classes.pop("count")
    assert len(engine.OBP_K_NAMES) == count
    for name, k in classes.items():
        assert engine.OBP_K_NAMES[k] == name
