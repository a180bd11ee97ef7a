"""Multi-GPU check of the sharded table over NCCL (launch with torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py [--parity-only]

1. parity against the oracle on small circuits (every rank gets the whole result back),
2. a large run (config 5 and config 2) for throughput with the exchange statistics.
"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, "."); sys.path.insert(0, "tests")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
os.environ["OBP_B200_DEVICE"] = str(local)
os.environ["OBP_B200_SHARDED"] = "1"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from conftest import assert_same_metadata, assert_same_operator
from oracle import obp_oracle as oracle
from qiskit_addon_obp_b200 import backpropagate
from qiskit_addon_obp_b200 import backpropagation as bpmod
from qiskit_addon_obp_b200 import workloads as W
from qiskit_addon_obp_b200.utils.truncating import TruncationErrorBudget
from test_engine_gpu import random_observable, random_slices

p2p = os.environ.get("OBP_B200_P2P", "1")


def log(*a):
    if rank == 0:
        print(*a, flush=True)

# ---- 1. parity -------------------------------------------------------------------------------------
os.environ["OBP_B200_SHARD_MIN_TERMS"] = "0"  # tiny tables: shard from the first commit (exchanges from the start)
for nq, seed in ((6, 0), (70, 1), (130, 2)):
    rng = np.random.default_rng(100 + seed)
    slices = random_slices(nq, 6, rng, gates_per_slice=8)
    obs = random_observable(nq, 4, rng)
    for kw in ({}, {"truncation_error_budget": TruncationErrorBudget([0.02, 0.0, 0.05], 0.2, 1, 1e-8)}):
        got, rest, md = backpropagate(obs, slices, **kw)
        want, rest_w, md_w = oracle.backpropagate(obs, slices, **kw)
        assert len(rest) == len(rest_w)
        assert_same_operator(got, want)
        assert_same_metadata(md, md_w, allow_missing_counters=True)
wl = W.config1_xy_chain()
got, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
want, _, md_w = oracle.backpropagate(wl.observables, wl.slices, **wl.kwargs())
assert_same_operator(got, want); assert_same_metadata(md, md_w, allow_missing_counters=True)
# late sharding: replicated at first, split into shards mid-run.  (a) a periodic circuit: the planner cannot cover
# its rotations without putting every row on one rank, records cross NVLink; (b) the same with the plain hash shard
# function; (c) a near-Clifford brickwork: the planned shard function covers every rotation, no exchange at all
from test_engine_gpu import window_workload
os.environ["OBP_B200_SHARD_MIN_TERMS"] = "300"
obs, slices = window_workload(70, 4)
want, rest_w, md_w = oracle.backpropagate(obs, slices)
for plan in ("1", "0"):
    os.environ["OBP_B200_SHARD_PLAN"] = plan
    got, rest, md = backpropagate(obs, slices)
    assert_same_operator(got, want); assert_same_metadata(md, md_w, allow_missing_counters=True)
    assert bpmod.last_run.exchanges > 0 and bpmod.last_run.exchanged_records > 0
os.environ["OBP_B200_SHARD_PLAN"] = "1"
os.environ["OBP_B200_SHARD_MIN_TERMS"] = "200"
wl = W.config4_near_clifford(num_qubits=200, depth=24, t_prob=0.05, max_paulis=20_000, n_observable_terms=4)
got, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
want, _, md_w = oracle.backpropagate(wl.observables, wl.slices, **wl.kwargs())
assert_same_operator(got, want); assert_same_metadata(md, md_w, allow_missing_counters=True)
assert bpmod.last_run.exchanges == 0, "planned shard function: no rotation after the split needs an exchange"
os.environ.pop("OBP_B200_SHARD_PLAN")
os.environ.pop("OBP_B200_SHARD_MIN_TERMS")
dist.barrier()
log(f"[dist_check] parity vs oracle OK on {world} ranks (NCCL, P2P inbox exchange = {p2p})")

if "--parity-only" in sys.argv:
    dist.destroy_process_group()
    sys.exit(0)

# ---- 2. throughput -----------------------------------------------------------------------------------
def run(wl, tag):
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier()
        t = time.perf_counter()
        out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t
        r = bpmod.last_run
        log(f"[dist_check] {tag} rep {rep}: slices {md.num_backpropagated_slices} terms {len(out)} updates {r.updates:.4g} "
            f"wall {dt*1e3:.1f} ms -> {r.updates/dt:.3e} upd/s; exchanges {r.exchanges} records {r.exchanged_records:.4g}")

run(W.config5_magic_sweep(theta=np.pi / 8, depth=40, max_paulis=4_000_000 * world), "config5 pi/8")
run(W.config2_kicked_ising(theta_h=0.3, max_paulis=4_000_000 * world), "config2 0.3")
run(W.config4_near_clifford(depth=120, max_paulis=200_000 * world), "config4")
dist.destroy_process_group()
