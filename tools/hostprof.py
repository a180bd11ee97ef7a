"""cProfile of one warm backpropagate() call (host-side overhead hunt): python tools/hostprof.py <config> <theta> <max_paulis>"""
import cProfile, pstats, sys, time
import numpy as np
sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W
name = sys.argv[1] if len(sys.argv) > 1 else "config2"
th = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
mp = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1_000_000
wl = {"config2": lambda: W.config2_kicked_ising(theta_h=th, max_paulis=mp),
      "config4": lambda: W.config4_near_clifford(depth=200, max_paulis=mp),
      "config1": lambda: W.config1_xy_chain(),
      "config3": lambda: W.config3_square_lattice(max_paulis=mp),
      "config5": lambda: W.config5_magic_sweep(theta=th, max_paulis=mp)}[name]()
backpropagate(wl.observables, wl.slices, **wl.kwargs())
for _ in range(2):
    t = time.perf_counter()
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
    print("wall", time.perf_counter() - t, "device_ms", bpmod.last_run.device_ms, "updates", bpmod.last_run.updates)
pr = cProfile.Profile()
pr.enable()
backpropagate(wl.observables, wl.slices, **wl.kwargs())
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
