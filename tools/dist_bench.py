"""Throughput of the sharded table over NCCL/P2P (torchrun, one rank per GPU): python ... tools/dist_bench.py <config> <theta> <max_paulis_per_rank>"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
os.environ["OBP_B200_DEVICE"] = str(local)
os.environ["OBP_B200_SHARDED"] = "1"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W
name, theta, mp = sys.argv[1], float(sys.argv[2]), int(float(sys.argv[3])) * world
wl = {"config2": lambda: W.config2_kicked_ising(theta_h=theta, max_paulis=mp),
      "config5": lambda: W.config5_magic_sweep(theta=theta, depth=40, max_paulis=mp),
      "config4": lambda: W.config4_near_clifford(depth=200, max_paulis=mp)}[name]()
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t
    r = bpmod.last_run
    if rank == 0:
        print(f"[dist_bench] {world} ranks {wl.name} rep {rep}: slices {md.num_backpropagated_slices} terms {len(out)} updates {r.updates:.4g} "
              f"wall {dt*1e3:.1f} ms -> {r.updates/dt:.3e} upd/s; exchanges {r.exchanges} records {r.exchanged_records:.4g}", flush=True)
if os.environ.get("OBP_PROFILE_HOST") and rank == 0:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
elif os.environ.get("OBP_PROFILE_HOST"):
    backpropagate(wl.observables, wl.slices, **wl.kwargs())
dist.destroy_process_group()
