"""Host-side logic of the multi-rank path on CPU: ``DistComm`` over gloo with world_size 2.

The CUDA kernels need a GPU, but the exchange protocol (counts first, then the payload, pairwise with
the peer ``rank ^ d``), the allreduces the budgets use and the operator gather are plain
``torch.distributed`` calls that run on the gloo backend with host tensors.
"""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qiskit_addon_obp_b200.ir import SparsePauliOp
        from qiskit_addon_obp_b200.sharded import DistComm

        comm = DistComm()
        assert comm.n_ranks == world and comm.local_ranks == [rank]
        # pairwise exchange of a ragged number of 5-word records (W = 1 -> 2W+3 = 5)
        n_send = 3 + 4 * rank
        send = torch.arange(n_send * 5, dtype=torch.int64).reshape(n_send, 5) + 1000 * rank
        got, n_got = comm.exchange({rank: send}, {rank: n_send}, 1)[rank]
        peer = rank ^ 1
        want = torch.arange((3 + 4 * peer) * 5, dtype=torch.int64).reshape(-1, 5) + 1000 * peer
        assert n_got == 3 + 4 * peer and torch.equal(got[:n_got], want)
        # an empty batch in one direction
        got, n_got = comm.exchange({rank: send}, {rank: 0 if rank == 0 else 2}, 1)[rank]
        assert n_got == (2 if rank == 0 else 0)
        # budget reductions
        assert comm.allreduce_sum({rank: [1.5 + rank, 2.0]}).tolist() == [4.0, 4.0]
        assert comm.allreduce_max({rank: [0.25 * (rank + 1)]}).tolist() == [0.5]
        assert comm.allreduce_sum_int({rank: [10 + rank, 7]}).tolist() == [21, 14]
        ti, tf = comm.allreduce_pair({rank: [1, rank, 5]}, {rank: [0.5, float(rank)]})
        assert ti.tolist() == [2, 1, 10] and tf.tolist() == [1.0, 1.0]
        # operator gather: rank order, packed words preserved
        xw = np.full((2 + rank, 1), rank + 1, dtype="<u8")
        op = SparsePauliOp.from_packed(xw, xw * 2, np.full(2 + rank, rank + 0.5), 5)
        full = comm.gather_operators({rank: op})
        assert len(full) == 5 and full.packed()[0][:, 0].tolist() == [1, 1, 2, 2, 2]
        assert full.coeffs.real.tolist() == [0.5, 0.5, 1.5, 1.5, 1.5]
        assert full.packed()[1][:, 0].tolist() == [2, 2, 4, 4, 4] and full.num_qubits == 5
        # one rank holds nothing
        empty = SparsePauliOp.from_packed(xw[:0], xw[:0], np.zeros(0), 5)
        full = comm.gather_operators({rank: op if rank == 1 else empty})
        assert len(full) == 3 and full.coeffs.real.tolist() == [1.5, 1.5, 1.5]
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


def test_dist_comm_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results


def test_local_comm_matches_protocol():
    from qiskit_addon_obp_b200.sharded import LocalComm

    comm = LocalComm(4)
    bufs = {r: f"buf{r}" for r in range(4)}
    counts = {r: 10 + r for r in range(4)}
    got = comm.exchange(bufs, counts, 3)
    assert got == {0: ("buf3", 13), 1: ("buf2", 12), 2: ("buf1", 11), 3: ("buf0", 10)}
    assert comm.allreduce_sum({r: [r, 1.0] for r in range(4)}).tolist() == [6.0, 4.0]
    assert comm.allreduce_max({r: [float(r)] for r in range(4)}).tolist() == [3.0]
    ti, tf = comm.allreduce_pair({r: [1, r] for r in range(4)}, {r: [0.25] for r in range(4)})
    assert ti.tolist() == [4, 6] and tf.tolist() == [1.0]
