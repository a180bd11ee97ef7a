"""cProfile of one warm backpropagate() call (host-side overhead hunt)."""
import cProfile, pstats, sys, time
import numpy as np
sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W
bpmod.PROFILE = True
th = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
wl = W.config2_kicked_ising(theta_h=th)
backpropagate(wl.observables, wl.slices, **wl.kwargs())
for _ in range(2):
    t = time.perf_counter()
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
    print("wall", time.perf_counter() - t, "device_ms", bpmod.last_run.device_ms, "updates", bpmod.last_run.updates)
    print(bpmod.last_run.profile)
pr = cProfile.Profile()
pr.enable()
backpropagate(wl.observables, wl.slices, **wl.kwargs())
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
