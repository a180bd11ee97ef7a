"""Two warm calls of backpropagate() on config 2 at one angle (target of ncu captures): python tools/one_run.py [theta] [max_paulis]"""
import sys

sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W

th = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
mp = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
wl = W.config2_kicked_ising(theta_h=th, max_paulis=mp)
for _ in range(2):
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
    print("updates", bpmod.last_run.updates, "device_ms", round(bpmod.last_run.device_ms, 3), "segments", bpmod.last_run.segments)
