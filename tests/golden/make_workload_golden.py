"""Golden vectors of whole backpropagations, generated with the CPU oracle (run in the build container):

    python tests/golden/make_workload_golden.py

``workload_golden.json`` pins, for small deterministic workloads, the complete result — Pauli labels in the
reference's row order, coefficients, per-slice sizes, errors and simplify counters.  The oracle is itself pinned
against the reference's known-answer tests (``tests/test_oracle_kat.py``); these vectors freeze its output so
that a later change to the oracle or to the engine shows up against committed data, and so that the GPU tests
have fixtures that do not depend on any code under ``oracle/`` at run time.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

from oracle import obp_oracle as oracle  # noqa: E402
from qiskit_addon_obp_b200 import workloads as W  # noqa: E402


def record(name, wl):
    out, rest, md = oracle.backpropagate(wl.observables, wl.slices, **wl.kwargs())
    ops = [out] if not isinstance(out, list) else out
    return {
        "name": name,
        "workload": wl.name,
        "remaining_slices": len(rest),
        "num_backpropagated_slices": md.num_backpropagated_slices,
        "observables": [
            {"labels": o.paulis.to_labels(), "re": [float(c.real) for c in o.coeffs], "im": [float(c.imag) for c in o.coeffs]}
            for o in ops
        ],
        "history": [
            {
                "slice_errors": [float(e) for e in s.slice_errors],
                "raw_num_paulis": list(s.raw_num_paulis),
                "num_unique_paulis": list(s.num_unique_paulis),
                "num_duplicate_paulis": list(s.num_duplicate_paulis),
                "num_trimmed_paulis": list(s.num_trimmed_paulis),
                "num_truncated_paulis": list(s.num_truncated_paulis),
                "num_paulis": list(s.num_paulis),
                "sum_paulis": s.sum_paulis,
            }
            for s in md.backpropagation_history
        ],
        "accumulated_error": [float(md.accumulated_error(i)) for i in range(len(ops))],
    }


cases = [
    record("config1", W.config1_xy_chain()),
    record("config2_small", W.config2_kicked_ising(theta_h=0.3, steps=2, max_paulis=3000)),
    record("config3_small", W.config3_square_lattice(side=3, steps=2, max_paulis=5000)),
]
json.dump({"generator": "tests/golden/make_workload_golden.py (CPU oracle)", "cases": cases},
          open(os.path.join(HERE, "workload_golden.json"), "w"), indent=0)
print({c["name"]: (len(c["observables"][0]["labels"]), c["num_backpropagated_slices"]) for c in cases})
