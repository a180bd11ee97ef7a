import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from test_engine_gpu import random_observable, random_slices
from qiskit_addon_obp_b200.sharded import ShardedTable, LocalComm
from qiskit_addon_obp_b200.backpropagation import _slice_program
from qiskit_addon_obp_b200.engine import TermTable
ranks, nq = 2, 6
rng = np.random.default_rng(31 * nq + ranks)
slices = random_slices(nq, 6, rng, gates_per_slice=8)
obs = random_observable(nq, 4, rng)
comm = LocalComm(ranks, 0)
st = ShardedTable(nq, 1 << 16, comm)
ref = TermTable(nq, 1 << 16, 0)
st.load(obs); ref.load(obs)
print("load", st.n, ref.n, [t.n for t in st.tables.values()])
for sl in slices[::-1]:
    prog = _slice_program(sl)
    ops = prog.recs[list(range(len(prog.nodes)))]
    ops["flags"][-1] = 1
    a = st.apply_ops(ops, 1e-8); b = ref.apply_ops(ops, 1e-8)
    da = st.download(); db = ref.download()
    flag = "" if (st.n == ref.n == len(da) == len(db)) else "  <<<<<<"
    print([ (n.name, n.qubits) for n in prog.nodes], "sharded n", st.n, [t.n for t in st.tables.values()], "dl", len(da), "ref", ref.n, len(db), "exch", st.exchanges, "upd", a.updates, b.updates, flag)
    print("   counters", (a.num_unique, a.num_duplicate, a.num_trimmed, a.counters_valid), (b.num_unique, b.num_duplicate, b.num_trimmed))
