"""Print the per-slice term counts the engine reaches on the named workloads (GPU; calibration aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W

def run(wl):
    t = time.perf_counter()
    out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
    dt = time.perf_counter() - t
    r = bpmod.last_run
    print(wl.name)
    print("  slices", md.num_backpropagated_slices, "of", len(wl.slices), "terms", len(out), "updates", r.updates,
          f"wall {dt:.3f}s device {r.device_ms:.1f}ms -> {r.updates / max(dt, 1e-9):.3e} upd/s e2e, {r.updates / max(r.device_ms * 1e-3, 1e-9):.3e} device")
    print("  num_paulis", [s.num_paulis[0] for s in md.backpropagation_history][-24:])

for th in (0.3, np.pi / 4):
    run(W.config2_kicked_ising(theta_h=th))
    run(W.config2_kicked_ising(theta_h=th))
for d in (10, 12, 14, 16):
    run(W.config5_magic_sweep(theta=np.pi / 8, depth=d, max_paulis=None))
run(W.config3_square_lattice())
run(W.config4_near_clifford(depth=60, max_paulis=1_000_000))
run(W.config5_magic_sweep(theta=np.pi / 16, depth=40, max_paulis=10_000_000))
