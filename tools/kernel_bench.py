"""Per-kernel-class device time on large tables (HBM-bound regime). GPU only.

usage: python tools/kernel_bench.py <config2|config5> <theta> <max_paulis> [reps]
Prints, per kernel class, launches / ms / term-gate updates and the algorithmic GB/s (B_tg per update).
"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W

name, theta, max_paulis = sys.argv[1], float(sys.argv[2]), int(float(sys.argv[3]))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if name == "config2":
    wl = W.config2_kicked_ising(theta_h=theta, max_paulis=max_paulis)
elif name == "config5":
    wl = W.config5_magic_sweep(theta=theta, depth=40, max_paulis=max_paulis)
elif name == "config3":
    wl = W.config3_square_lattice(max_paulis=max_paulis)
elif name == "config1":
    wl = W.config1_xy_chain()
elif name == "config4":
    wl = W.config4_near_clifford(depth=200, max_paulis=max_paulis)
else:
    raise SystemExit(name)
btg = 2 * (16 * ((wl.num_qubits + 63) // 64) + 16)
bpmod.PROFILE = True
for r in range(reps):
    t = time.perf_counter()
    out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
    dt = time.perf_counter() - t
    run = bpmod.last_run
    sizes = [s.num_paulis[0] for s in md.backpropagation_history]
    print(f"rep {r}: {wl.name}")
    print(f"  slices {md.num_backpropagated_slices} terms_out {len(out)} peak {max(sizes)} updates {run.updates:.4g} wall {dt*1e3:.1f} ms device {run.device_ms:.1f} ms"
          f" -> e2e {run.updates/dt:.3e} upd/s, device {run.updates/(run.device_ms*1e-3):.3e} upd/s")
    for k, v in run.profile.items():
        if v["launches"] or v["passes"]:
            gbs = v["term_updates"] * btg / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 and v["term_updates"] else 0.0
            print(f"  {k:9s} launches {v['launches']:6d} passes {v['passes']:6d} ms {v['ms']:10.3f} term_updates {v['term_updates']:.4g} -> {gbs:8.1f} GB/s algorithmic ({btg} B/update), {v['term_updates']/max(v['ms'],1e-9)/1e6:.2f} G upd/s")
