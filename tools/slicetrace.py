"""Where does a warm backpropagate() spend its wall time?  python tools/slicetrace.py [theta] [max_paulis]

Runs config 2 with OBP_B200_TRACE_HOST=1 (the engine prints, per gate program, the host time spent
enqueueing and the time spent waiting for the device) and splits the wall time of the last warm call into
  enqueue  — C++ host work before the kernels are in the stream
  wait     — the host blocked on the stream (≈ device busy time of the slice)
  python   — everything else: the driver loop between the C calls (the device idles during it)
"""
import os, re, subprocess, sys

if os.environ.get("SLICETRACE_CHILD") != "1":
    env = dict(os.environ, SLICETRACE_CHILD="1", OBP_B200_TRACE_HOST="1")
    out = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    runs = out.stderr.split("[slicetrace] run\n")
    last = runs[-1]
    enq = [float(m.group(1)) for m in re.finditer(r"enqueue ([0-9.]+) us", last)]
    wait = [float(m.group(1)) for m in re.finditer(r"wait ([0-9.]+) us", last)]
    print(out.stdout.strip())
    wall_ms = float(re.search(r"wall_ms ([0-9.]+)", out.stdout.strip().splitlines()[-1]).group(1))
    print(f"programs {len(enq)}  enqueue {sum(enq) / 1e3:.2f} ms  wait {sum(wait) / 1e3:.2f} ms  "
          f"python+other {wall_ms - (sum(enq) + sum(wait)) / 1e3:.2f} ms  (wall {wall_ms:.2f} ms)")
    print(f"per program: enqueue {sum(enq) / max(len(enq), 1):.1f} us  wait {sum(wait) / max(len(wait), 1):.1f} us")
    sys.exit(out.returncode)

import gc, time
sys.path.insert(0, ".")
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W

th = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
mp = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
wl = W.config2_kicked_ising(theta_h=th, max_paulis=mp)
gc.collect()
gc.disable()
for _ in range(4):
    print("[slicetrace] run", file=sys.stderr, flush=True)
    t = time.perf_counter()
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
    wall = time.perf_counter() - t
    r = bpmod.last_run
    print(f"theta {th} updates {r.updates:.4g} device_ms {r.device_ms:.2f} wall_ms {wall * 1e3:.2f}", flush=True)
