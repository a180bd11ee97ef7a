"""One observable sharded over several GPUs by the hash of its Pauli keys.

Rank ``r`` of ``R`` (a power of two) owns the terms whose Clifford-invariant 64-bit hash has its top
``log2 R`` bits equal to ``r`` (``include/obp_b200.h``, "sharded table").  Duplicate keys therefore
co-locate; Clifford gates, Pauli-Lindblad damping and truncation kills are rank-local; a rotation
whose generator hash has owner bits ``d != 0`` exchanges one batch of records between the pairs
``r <-> r ^ d`` (the "all-to-all" of a branching gate is a pairwise permutation because the hash is
GF(2)-linear); budgets (term counts, masked sums, maxima) are combined with small allreduces.

The reference is a single process (``backpropagation.py:167`` only remarks that the loop "will
likely need to be parallelized"), so there is no reference counterpart of this module; its results
are checked against the same oracle as the single-table path.

Two communicators implement the exchange:

* :class:`DistComm` — one process per GPU over ``torch.distributed`` (NCCL send/recv on device
  buffers, NCCL allreduce); every rank runs the same driver loop (SPMD).
* :class:`LocalComm` — all ``R`` ranks inside one process on one GPU, records handed over as device
  pointers.  It exercises exactly the same kernels and protocol and is what the single-GPU parity
  tests use.
"""

from __future__ import annotations

import os
import time

import numpy as np

from .engine import OBP_ATOL_RAW, OBP_OP_ROTATION, CapacityError, TermTable, lib
from .ir import SparsePauliOp


class LocalComm:
    """``R`` virtual ranks in this process (one device)."""

    def __init__(self, n_ranks: int, device: int = 0):
        self.n_ranks = int(n_ranks)
        self.local_ranks = list(range(self.n_ranks))
        self.device = int(device)
        self.backend = f"local x{self.n_ranks}"

    def exchange(self, bufs: dict, counts: dict, peer_xor: int) -> dict:
        """Rank r receives what rank r ^ peer_xor emitted: ``{r: (buffer, n_records)}``."""
        return {r: (bufs[r ^ peer_xor], counts[r ^ peer_xor]) for r in self.local_ranks}

    def allreduce_sum(self, vals: dict) -> np.ndarray:
        return np.sum([np.asarray(vals[r], dtype=np.float64) for r in self.local_ranks], axis=0)

    def allreduce_max(self, vals: dict) -> np.ndarray:
        return np.max([np.asarray(vals[r], dtype=np.float64) for r in self.local_ranks], axis=0)

    def allreduce_sum_int(self, vals: dict) -> np.ndarray:
        return np.sum([np.asarray(vals[r], dtype=np.int64) for r in self.local_ranks], axis=0)

    def allreduce_pair(self, ints: dict, floats: dict) -> tuple[np.ndarray, np.ndarray]:
        """Sum an int64 vector and a float64 vector over the ranks (one round trip)."""
        return self.allreduce_sum_int(ints), self.allreduce_sum(floats)

    def gather_operators(self, ops: dict) -> SparsePauliOp:
        return _concat([ops[r] for r in self.local_ranks])

    def barrier(self) -> None:
        pass  # one thread runs all ranks: every emit has returned before the first merge starts

    def share_inboxes(self, exports: dict) -> dict:
        """``exports[r] = [(ipc handle, raw pointer)] * 2`` -> the same for every rank, IPC handle dropped."""
        return {r: [(None, ptr) for _, ptr in exports[r]] for r in self.local_ranks}


class DistComm:
    """One rank per process over ``torch.distributed`` (NCCL on GPUs; gloo works for host tests)."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.n_ranks = dist.get_world_size(group)
        self.local_ranks = [self.rank]
        self.backend = dist.get_backend(group)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        self.tdev = device
        self.device = device.index if device.type == "cuda" else 0
        if self.n_ranks & (self.n_ranks - 1):
            raise ValueError("the number of ranks must be a power of two")

    def exchange(self, bufs: dict, counts: dict, peer_xor: int) -> dict:
        torch, dist = self.torch, self.dist
        peer = self.rank ^ peer_xor
        send, n_send = bufs[self.rank], int(counts[self.rank])
        rec = send.shape[1]
        # 1. counts, 2. payload — both as one batched send/recv pair with the single peer
        c_out = torch.tensor([n_send], dtype=torch.int64, device=self.tdev)
        c_in = torch.zeros(1, dtype=torch.int64, device=self.tdev)
        for w in dist.batch_isend_irecv(
            [dist.P2POp(dist.isend, c_out, peer, self.group), dist.P2POp(dist.irecv, c_in, peer, self.group)]
        ):
            w.wait()
        n_recv = int(c_in.item())
        recv = torch.empty((max(n_recv, 1), rec), dtype=send.dtype, device=self.tdev)
        ops = []
        if n_send:
            ops.append(dist.P2POp(dist.isend, send[:n_send], peer, self.group))
        if n_recv:
            ops.append(dist.P2POp(dist.irecv, recv[:n_recv], peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.tdev.type == "cuda":
            torch.cuda.current_stream(self.tdev).synchronize()
        return {self.rank: (recv, n_recv)}

    def _allreduce(self, vals: dict, op, dtype):
        torch = self.torch
        t = torch.as_tensor(np.atleast_1d(np.asarray(vals[self.rank])), dtype=dtype).to(self.tdev)
        self.dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def allreduce_sum(self, vals: dict) -> np.ndarray:
        return self._allreduce(vals, self.dist.ReduceOp.SUM, self.torch.float64)

    def allreduce_max(self, vals: dict) -> np.ndarray:
        return self._allreduce(vals, self.dist.ReduceOp.MAX, self.torch.float64)

    def allreduce_sum_int(self, vals: dict) -> np.ndarray:
        return self._allreduce(vals, self.dist.ReduceOp.SUM, self.torch.int64)

    def allreduce_pair(self, ints: dict, floats: dict) -> tuple[np.ndarray, np.ndarray]:
        """Sum an int64 vector and a float64 vector over the ranks: both collectives in flight together."""
        torch = self.torch
        ti = torch.as_tensor(np.asarray(ints[self.rank], dtype=np.int64)).to(self.tdev, non_blocking=True)
        tf = torch.as_tensor(np.asarray(floats[self.rank], dtype=np.float64)).to(self.tdev, non_blocking=True)
        wi = self.dist.all_reduce(ti, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        wf = self.dist.all_reduce(tf, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
        wi.wait()
        wf.wait()
        return ti.cpu().numpy(), tf.cpu().numpy()

    def barrier(self) -> None:
        if self.tdev.type == "cuda":
            self.torch.cuda.synchronize(self.tdev)
        self.dist.barrier(group=self.group)

    def share_inboxes(self, exports: dict) -> dict:
        parts: list = [None] * self.n_ranks
        self.dist.all_gather_object(parts, exports[self.rank], group=self.group)
        return {r: [(h, 0) for h, _ in parts[r]] for r in range(self.n_ranks)}

    def gather_operators(self, ops: dict) -> SparsePauliOp:
        """Every rank gets the whole operator: one padded all_gather of the packed rows (no pickling)."""
        torch = self.torch
        mine = ops[self.rank]
        xw, zw = mine.packed()
        n, w = xw.shape[0], xw.shape[1]
        row = np.concatenate([xw, zw, mine.coeffs.view(np.float64).reshape(n, 2).view("<u8")], axis=1)  # [n, 2W+2] u64
        sizes = self.allreduce_sum_int({self.rank: [n if r == self.rank else 0 for r in range(self.n_ranks)]})
        n_max = int(max(sizes.max(), 1))
        send = torch.zeros((n_max, 2 * w + 2), dtype=torch.int64, device=self.tdev)
        if n:
            send[:n] = torch.from_numpy(row.view(np.int64)).to(self.tdev)
        recv = torch.empty((self.n_ranks * n_max, 2 * w + 2), dtype=torch.int64, device=self.tdev)
        self.dist.all_gather_into_tensor(recv, send, group=self.group)
        if self.tdev.type == "cuda":  # pinned staging: pageable D2H of a multi-100-MB result is ~5x slower
            host = torch.empty(recv.shape, dtype=recv.dtype, pin_memory=True)
            host.copy_(recv, non_blocking=True)
            torch.cuda.current_stream(self.tdev).synchronize()
        else:
            host = recv
        full = host.numpy().view("<u8").reshape(self.n_ranks, n_max, 2 * w + 2)
        n_tot = int(sizes.sum())
        xw_all, zw_all = np.empty((n_tot, w), dtype="<u8"), np.empty((n_tot, w), dtype="<u8")
        co_all = np.empty((n_tot, 2), dtype="<u8")
        at = 0
        for r in range(self.n_ranks):  # one strided copy per column block, straight into the result arrays
            k = int(sizes[r])
            xw_all[at : at + k] = full[r, :k, :w]
            zw_all[at : at + k] = full[r, :k, w : 2 * w]
            co_all[at : at + k] = full[r, :k, 2 * w :]
            at += k
        return SparsePauliOp.from_packed(xw_all, zw_all, co_all.view(np.float64).view(np.complex128).reshape(-1), mine.num_qubits)


def _concat(ops: list[SparsePauliOp]) -> SparsePauliOp:
    xs, zs = zip(*[o.packed() for o in ops])
    return SparsePauliOp.from_packed(
        np.concatenate(xs, axis=0), np.concatenate(zs, axis=0), np.concatenate([o.coeffs for o in ops]), ops[0].num_qubits
    )


class _Stats:
    """Whole-table view of per-rank counters, field-compatible with ``ObpStats`` where it matters."""

    def __init__(self):
        self.n_in = self.raw_rows = self.n_out = self.n_dense = self.updates = 0
        self.num_unique = self.num_duplicate = self.num_trimmed = 0
        self.sum_trimmed = [0.0, 0.0]
        self.counters_valid = True


class ShardedTable:
    """Drop-in for :class:`~qiskit_addon_obp_b200.engine.TermTable` in the driver loop."""

    def __init__(self, num_qubits: int, capacity: int, comm):
        """``capacity`` is the whole table's; every rank gets 1/R of it plus 25 % slack for imbalance."""
        self.comm = comm
        self.num_qubits = int(num_qubits)
        self.words = max(1, (self.num_qubits + 63) // 64)
        self.capacity = int(capacity)
        self.device = comm.device
        self.rec_u64 = int(lib().obp_record_bytes(self.num_qubits)) // 8
        per_rank = self._per_rank(capacity)
        self.tables = {r: TermTable.acquire(num_qubits, per_rank, comm.device) for r in comm.local_ranks}
        # Late sharding: while the operator holds fewer than `split_min` terms every rank carries an identical
        # copy of the WHOLE table (same deterministic kernels on the same input: no exchange, no allreduce, exact
        # counters); at the first commit at or above that size each rank drops the rows it does not own.
        self.split = False
        self.split_min = int(os.environ.get("OBP_B200_SHARD_MIN_TERMS", str(1 << 17)))
        self._split_on_restore = False
        self._skip_next_reserve = False
        self._inboxes_ready = False
        self.lookahead = None  # set by the driver: (reversed slices, next slice, support, expand_all), see _split_now
        self.planned_rotations = 0  # rotations the shard function was planned against (they need no exchange)
        self.p2p = os.environ.get("OBP_B200_P2P", "1") not in ("0", "")
        # one process per GPU: split rotations are ordered by device-side flags, the host never waits
        self.device_flags = (
            self.p2p and isinstance(comm, DistComm) and comm.backend == "nccl"
            and os.environ.get("OBP_B200_DEVICE_FLAGS", "1") not in ("0", "")
        )
        self.n = 0
        self.n_dense = 0
        self.zero_operator = False
        self._xstat0 = {r: t.shard_stats() for r, t in self.tables.items()}  # counters at the start of this run
        self._host_exchanges = 0  # host-ordered exchange paths (staged / barrier): counted here
        self._host_records = 0
        self.t_phase = {"local": 0.0, "emit": 0.0, "barrier": 0.0, "merge": 0.0, "reduce": 0.0}  # host seconds

    # -- plumbing ---------------------------------------------------------------------------------
    def _per_rank(self, capacity: int) -> int:
        # a rank's share of the table plus 25 % for imbalance — and room for the replicated phase
        return max(1 << 16, int(1.25 * capacity / self.comm.n_ranks) + 1, min(int(capacity), 4 * self.split_min_env()))

    @staticmethod
    def split_min_env() -> int:
        return int(os.environ.get("OBP_B200_SHARD_MIN_TERMS", str(1 << 17)))

    def _reset_replicated(self) -> None:
        """Back to the replicated phase (a recycled table starts a new run)."""
        self.split = False
        self._split_on_restore = self._skip_next_reserve = False
        for t in self.tables.values():
            t.set_shard(1, 0)
            t.set_shard_sync(False)
            t.track_zero = False

    def _split_now(self) -> None:
        """Every rank keeps its share (collective: all ranks reach this at the same commit).  The shard function is
        planned against the gates the run is about to apply (``OBP_B200_SHARD_PLAN=0`` turns that off): the owner
        bits of their rotation generators are zero, so those rotations never move a term to another rank."""
        ops = None
        t0 = time.perf_counter()
        if os.environ.get("OBP_B200_SHARD_PLAN", "1") not in ("0", ""):
            from .backpropagation import future_ops

            ops = future_ops(self.lookahead)
        t1 = time.perf_counter()
        for r, t in self.tables.items():
            _, self.planned_rotations = t.shard_split(self.comm.n_ranks, r, ops)
        t2 = time.perf_counter()
        self.t_phase["plan"] = self.t_phase.get("plan", 0.0) + (t1 - t0)
        self.t_phase["split"] = self.t_phase.get("split", 0.0) + (t2 - t1)
        if self.p2p and not self._inboxes_ready:
            self._setup_inboxes(max(t.capacity for t in self.tables.values()))
            self._inboxes_ready = True
        for t in self.tables.values():
            t.set_shard_sync(self.device_flags)
        self.split = True
        self._split_on_restore = False
        self._sync_counts()
        for t in self.tables.values():  # the commit the split invalidated
            t.snapshot()
        self._snap = (self.zero_operator, self.n)
        self.t_phase["split_total"] = self.t_phase.get("split_total", 0.0) + (time.perf_counter() - t0)
        if os.environ.get("OBP_B200_TRACE"):
            print(f"[obp trace] split into {self.comm.n_ranks} shards at {self.n} terms: local rows "
                  f"{[t.n for t in self.tables.values()]}, plan covers {self.planned_rotations} rotations "
                  f"({len(ops) if ops is not None else 0} gates looked ahead)", flush=True)

    def _maybe_split(self) -> None:
        if not self.split and not self.zero_operator and self.comm.n_ranks > 1 and self.n >= self.split_min:
            self._split_now()

    def _rep(self):
        """The replicated phase's representative table (all local copies are identical)."""
        return self.tables[min(self.tables)]

    def _rep_stats(self, stats_by_rank: dict) -> "_Stats":
        st = stats_by_rank[min(stats_by_rank)]
        out = _Stats()
        for f in ("n_in", "raw_rows", "n_out", "n_dense", "updates", "num_unique", "num_duplicate", "num_trimmed"):
            setattr(out, f, int(getattr(st, f)))
        out.sum_trimmed = [float(st.sum_trimmed[0]), float(st.sum_trimmed[1])]
        t = self._rep()
        self.n, self.n_dense = t.n, t.n_dense
        if self.n == 0:
            self.zero_operator = True
        return out

    @classmethod
    def acquire(cls, num_qubits, capacity, comm):
        """A recycled table of this communicator (device memory, inboxes and peer mappings stay alive)."""
        pool = comm.__dict__.setdefault("_idle_tables", [])
        for k, t in enumerate(pool):
            if t.num_qubits == int(num_qubits) and t.capacity >= int(capacity):
                pool.pop(k)
                t.n = t.n_dense = 0
                t._host_exchanges = t._host_records = 0
                t._xstat0 = {r: m.shard_stats() for r, m in t.tables.items()}
                t.zero_operator = False
                t.t_phase = dict.fromkeys(t.t_phase, 0.0)
                t.split_min = cls.split_min_env()
                t.lookahead, t.planned_rotations = None, 0
                t._reset_replicated()
                return t
        t = cls(num_qubits, capacity, comm)
        t._reset_replicated()
        return t

    def _setup_inboxes(self, per_rank: int) -> None:
        """Every rank allocates its two inboxes and maps those of all its peers (collective)."""
        for t in self.tables.values():
            t.inbox_create(per_rank)
        shared = self.comm.share_inboxes({r: [t.inbox_export(0), t.inbox_export(1)] for r, t in self.tables.items()})
        for r, t in self.tables.items():
            for peer, boxes in shared.items():
                if peer != r:
                    for which, (handle, ptr) in enumerate(boxes):
                        t.inbox_attach(peer, which, handle, ptr)
        self.comm.barrier()

    def _drop_inboxes(self) -> None:
        for t in self.tables.values():
            t.inbox_release(0)
        self.comm.barrier()
        for t in self.tables.values():
            t.inbox_release(1)

    def release(self):
        """Back to the communicator's pool (collective-free); :meth:`close` really frees everything."""
        if self.tables and os.environ.get("OBP_B200_TRACE"):
            print(f"[obp trace] sharded table: {self.exchanges} exchanges, {self.exchanged_records} records, host seconds "
                  + ", ".join(f"{k} {v:.3f}" for k, v in self.t_phase.items()), flush=True)
        pool = self.comm.__dict__.setdefault("_idle_tables", [])
        if not self.tables:
            return
        # keep the most recently used tables: the one just released is the likeliest to be asked for again (every rank
        # releases in the same order, so the collective close() of the evicted one lines up)
        pool.append(self)
        while len(pool) > 2:
            pool.pop(0).close()

    def close(self):
        if self.tables and self.p2p and self._inboxes_ready:
            self._drop_inboxes()
            self._inboxes_ready = False
        for t in self.tables.values():
            t.release()
        self.tables = {}

    def __len__(self) -> int:
        return 1 if self.zero_operator else self.n

    @property
    def exchanges(self) -> int:
        """Rotations of this run that needed a record exchange (per rank: every rank takes part in each of them)."""
        if self.device_flags and self.tables:
            r = min(self.tables)
            return self.tables[r].shard_stats()[0] - self._xstat0[r][0]
        return self._host_exchanges

    @property
    def exchanged_records(self) -> int:
        """Records emitted by the local rank(s) during this run."""
        if self.device_flags and self.tables:
            return sum(t.shard_stats()[1] - self._xstat0[r][1] for r, t in self.tables.items())
        return self._host_records

    def _sync_counts(self):
        if not self.split:
            t = self._rep()
            self.n, self.n_dense = int(t.n), int(t.n_dense)
            return
        tot = self.comm.allreduce_sum_int({r: [t.n, t.n_dense] for r, t in self.tables.items()})
        self.n, self.n_dense = int(tot[0]), int(tot[1])

    def _raise_if_any(self, failed: bool, what: str):
        flag = self.comm.allreduce_max({r: [1.0 if failed else 0.0] for r in self.tables})
        if flag[0] > 0:
            raise CapacityError(f"{what}: a shard overflowed its table")

    # -- I/O ----------------------------------------------------------------------------------------
    def load(self, op: SparsePauliOp, atol: float | None = None) -> _Stats:
        out = _Stats()
        self._reset_replicated()
        if atol != OBP_ATOL_RAW:
            # replicated start: every rank loads the whole operator and keeps all of it for now
            self._unsimplified = False
            for t in self.tables.values():
                if len(op) > t.capacity:
                    t.reserve(2 * len(op))
                t.load(op, atol)
            self._sync_counts()
            self.zero_operator = atol is not None and self.n == 0
            out.n_in, out.n_out = len(op), self.n
            self._maybe_split()
            return out
        self.split = True  # simplify=False deals the rows out at once (below): there is nothing to keep identical
        for r, t in self.tables.items():
            t.set_shard(self.comm.n_ranks, r)
        if self.p2p and not self._inboxes_ready:
            self._setup_inboxes(max(t.capacity for t in self.tables.values()))
            self._inboxes_ready = True
        for t in self.tables.values():
            t.set_shard_sync(self.device_flags)
        if atol == OBP_ATOL_RAW:
            # simplify=False: nothing ever merges, so ownership by key hash is moot — deal the rows out evenly
            R = self.comm.n_ranks
            xw, zw = op.packed()
            for r, t in self.tables.items():
                part = SparsePauliOp.from_packed(xw[r::R], zw[r::R], np.asarray(op.coeffs)[r::R], op.num_qubits)
                if len(part) > t.capacity:
                    t.reserve(2 * len(part))
                t.load(part, atol)
            self._sync_counts()
            self.zero_operator = False
            self._unsimplified = True  # rows as given (duplicates, zeros): a generic gate must see them as such
            out.n_in, out.n_out = len(op), self.n
            return out
        self._unsimplified = False
        for t in self.tables.values():  # every rank is given the whole operator and keeps its share
            if len(op) > t.capacity:
                t.reserve(2 * len(op))
            t.load(op, atol)
        self._sync_counts()
        self.zero_operator = atol is not None and self.n == 0
        out.n_in, out.n_out = len(op), self.n
        return out

    def download(self) -> SparsePauliOp:
        if self.zero_operator:
            z = np.zeros((1, self.num_qubits), dtype=bool)
            return SparsePauliOp((z.copy(), z), np.array([0j]))
        if not self.split:
            return self._rep().download()
        return self.comm.gather_operators({r: t.download() for r, t in self.tables.items()})

    to_operator = download

    # -- the hot path ---------------------------------------------------------------------------------
    def apply_ops(self, ops: np.ndarray, atol: float) -> _Stats:
        """A gate program on the sharded table.  The slice's last gate carries OBP_FLAG_EXACT_COUNTERS; when
        it is Clifford-class its counters need every member of a key's coset, which live on several ranks:
        they are reported as None unless OBP_B200_SHARDED_EXACT_COUNTERS=1 asks for the (whole-table gather,
        not scalable) exact path."""
        import torch

        out = _Stats()
        if self.zero_operator:
            k2 = int(ops[-1]["n_components"]) ** 2
            out.n_in, out.raw_rows, out.num_trimmed, out.n_out, out.updates = 1, k2, k2, 1, len(ops)
            return out
        if not self.split:
            try:
                return self._rep_stats({r: t.apply_ops(ops, atol) for r, t in self.tables.items()})
            except CapacityError:
                # the whole operator no longer fits one rank's table: split at the caller's rollback, retry sharded
                self._split_on_restore = self.comm.n_ranks > 1
                raise
        if (
            self.comm.n_ranks > 1 and int(ops[-1]["kind"]) != OBP_OP_ROTATION and int(ops[-1]["flags"]) & 1
            and os.environ.get("OBP_B200_SHARDED_EXACT_COUNTERS", "0") == "1"
        ):
            head, last_op = ops[:-1], ops[-1:]
            updates = 0
            if len(head):
                head = head.copy()
                updates = int(self.apply_ops(head, atol).updates)
                if self.zero_operator:
                    rest = self.apply_ops(last_op, atol)
                    rest.updates += updates
                    return rest
            # the last gate's counters from the gathered operator on one scratch table ...
            op = self.download()
            scratch = TermTable.acquire(self.num_qubits, max(1024, 2 * len(op)), self.device)
            try:
                scratch.track_zero = False
                scratch.load(op, None)
                exact = scratch.apply_ops(last_op, atol)
            finally:
                scratch.release()
            # ... and the gate itself on the shards
            plain = last_op.copy()
            plain["flags"] = 0
            out = self.apply_ops(plain, atol)
            out.updates += updates
            for f in ("n_in", "raw_rows", "num_unique", "num_duplicate", "num_trimmed"):
                setattr(out, f, int(getattr(exact, f)))
            out.sum_trimmed = [float(exact.sum_trimmed[0]), float(exact.sum_trimmed[1])]
            out.counters_valid = True
            return out
        pos, failed = 0, False
        updates = 0
        last: dict | None = None  # per-rank counters of the last executed op
        tdev = torch.device("cuda", self.device)
        if self.device_flags:
            # the whole program in one call: exchanges are ordered on the devices (inbox header flags)
            t0 = time.perf_counter()
            (r, t), = self.tables.items()
            st, done, peer, failed = t.apply_ops_sharded(ops, atol)
            if done != len(ops) or peer != 0:
                raise RuntimeError("device-flag mode: the program did not run to its end")
            updates = int(st.updates)
            split = t.last_split_counters()
            last = {"kind": "local", "stats": {r: st}, "op": ops[-1]} if split is None else {
                "kind": "split-fused", "stats": {r: st}, "cnt": {r: split}, "op": ops[-1]}
            pos = len(ops)
            self.t_phase["local"] += time.perf_counter() - t0
        while pos < len(ops):
            seg = ops[pos:]
            done = peer = 0
            seg_stats = {}
            t0 = time.perf_counter()
            for r, t in self.tables.items():
                st, done, peer, over = t.apply_ops_sharded(seg, atol)  # every rank stops at the same op
                failed |= over
                seg_stats[r] = st
                updates += int(st.updates)
            self.t_phase["local"] += time.perf_counter() - t0
            if done > 0:
                last = {"kind": "local", "stats": seg_stats, "op": seg[done - 1]}
            pos += done
            if peer == 0:
                break
            # ---- split rotation: emit, pairwise exchange, merge -----------------------------------
            op = ops[pos : pos + 1]
            emit, merge, counts = {}, {}, {}
            if self.p2p:
                # the emitting kernel stores its records straight into the peer's inbox over NVLink
                t0 = time.perf_counter()
                for r, t in self.tables.items():
                    emit[r], over = t.rotate_emit_p2p(op, atol, r ^ peer)
                    failed |= over
                    counts[r] = int(emit[r].n_records)
                    updates += int(emit[r].n_in)
                t1 = time.perf_counter()
                self.comm.barrier()
                t2 = time.perf_counter()
                for r, t in self.tables.items():
                    merge[r], over = t.rotate_merge_inbox()
                    failed |= over
                t3 = time.perf_counter()
                self.t_phase["emit"] += t1 - t0
                self.t_phase["barrier"] += t2 - t1
                self.t_phase["merge"] += t3 - t2
            else:
                bufs = {}
                for r, t in self.tables.items():
                    cap = max(int(t.n_dense), 1)
                    bufs[r] = torch.empty((cap, self.rec_u64), dtype=torch.int64, device=tdev)
                    emit[r], over = t.rotate_emit(op, atol, bufs[r].data_ptr(), cap)
                    failed |= over
                    counts[r] = min(int(emit[r].n_records), cap)
                    updates += int(emit[r].n_in)
                recv = self.comm.exchange(bufs, counts, peer)
                for r, t in self.tables.items():
                    buf, n_rec = recv[r]
                    merge[r], over = t.rotate_merge(buf.data_ptr() if n_rec else 0, n_rec)
                    failed |= over
                del bufs, recv
            self._host_exchanges += 1
            self._host_records += sum(counts.values())
            last = {"kind": "split", "emit": emit, "merge": merge, "op": op[0]}
            pos += 1
        # one combined reduction per program: overflow flag, row counts, update count, last-op counters
        t0 = time.perf_counter()
        first = min(self.tables)
        ints, floats = {}, {}
        for r, t in self.tables.items():
            c4, f2 = self._last_local(last, r)
            ints[r] = [1 if (failed and r == first) else 0, t.n, t.n_dense, updates if r == first else 0, *c4]
            floats[r] = f2
        ti, tf = self.comm.allreduce_pair(ints, floats)
        self.n, self.n_dense = int(ti[1]), int(ti[2])
        if ti[0] > 0:
            exc = CapacityError("sharded gate program: a shard overflowed its table")
            exc.updates, exc.n_live = int(ti[3]), self.n
            raise exc
        out.updates = int(ti[3])
        self._fill_last(out, last, ops[-1], [int(v) for v in ti[4:8]], [float(tf[0]), float(tf[1])])
        self.t_phase["reduce"] += time.perf_counter() - t0
        out.n_out, out.n_dense = self.n, self.n_dense
        if self.n == 0:
            self.zero_operator = True
        return out

    @staticmethod
    def _last_local(last, r) -> tuple[list[int], list[float]]:
        """This rank's four integer counters and the trimmed sum of the program's last op."""
        if last is None:
            return [0, 0, 0, 0], [0.0, 0.0]
        if last["kind"] == "local":
            st = last["stats"][r]
            return [int(st.n_in), int(st.num_unique), int(st.num_duplicate), int(st.num_trimmed)], list(st.sum_trimmed)
        if last["kind"] == "split-fused":
            c = last["cnt"][r]
            return [int(last["stats"][r].n_in), int(c.rows_surv), int(c.keys_present), int(c.cancelled)], list(c.sum_trimmed)
        e, m = last["emit"][r], last["merge"][r]
        return [int(e.n_in), int(e.rows_surv), int(e.keys_present + m.keys_present), int(m.cancelled)], list(m.sum_trimmed)

    def _fill_last(self, out: _Stats, last, last_op, tot: list[int], sums: list[float]) -> None:
        """Whole-table simplify counters of the program's last op from the summed per-rank pieces."""
        k2 = int(last_op["n_components"]) ** 2
        if last is None:
            out.counters_valid = False
            return
        out.sum_trimmed = sums
        if last["kind"] == "local":
            out.n_in, out.num_unique, out.num_duplicate, out.num_trimmed = tot
            # a Clifford-class gate's coset count sees only rank-local members of a coset
            out.counters_valid = int(last["op"]["kind"]) == OBP_OP_ROTATION or self.comm.n_ranks == 1
        else:  # a split rotation: rows counted where produced, keys where they live
            n_in, rows_surv, keys, cancelled = tot
            out.n_in = n_in
            out.num_unique = keys
            out.num_duplicate = rows_surv - keys
            out.num_trimmed = n_in * k2 - rows_surv + cancelled
        out.raw_rows = out.n_in * k2

    def apply_generic(self, qubits, codes, coeffs, atol: float, simplify: bool = True) -> _Stats:
        """A generic gate (cp, crz, arbitrary unitaries): its k*k raw rows land on any shard, so the operator is
        gathered, expanded and merged as ONE table on every rank (``k_expand`` + the merge of a fresh operator,
        exact counters), and every rank keeps its share of the result.  Correct but not scalable — whole-table
        host traffic per gate; no BASELINE configuration has such gates."""
        out = _Stats()
        k2 = len(codes) ** 2
        if not self.split and not self.zero_operator:
            return self._rep_stats({r: t.apply_generic(qubits, codes, coeffs, atol, simplify=simplify) for r, t in self.tables.items()})
        if not simplify:
            # OperatorBudget(simplify=False): every rank expands the rows it holds, nothing moves or merges
            n_in = self.n
            failed = False
            for t in self.tables.values():
                try:
                    t.apply_generic(qubits, codes, coeffs, atol, simplify=False)
                except CapacityError:
                    failed = True
            self._raise_if_any(failed, "sharded raw expansion")
            self._sync_counts()
            out.n_in, out.raw_rows, out.updates, out.n_out = n_in, n_in * k2, n_in, self.n
            return out
        if self.zero_operator:
            out.n_in, out.raw_rows, out.num_trimmed, out.n_out, out.updates = 1, k2, k2, 1, 1
            return out
        op = self.download()  # collective: every rank now holds the whole operator
        scratch = TermTable.acquire(self.num_qubits, max(1024, 2 * len(op) * (k2 + 1)), self.device)
        try:
            scratch.track_zero = False
            scratch.load(op, OBP_ATOL_RAW if getattr(self, "_unsimplified", False) else None)
            self._unsimplified = False
            st = scratch.apply_generic(qubits, codes, coeffs, atol, simplify=True)
            res = scratch.download() if scratch.n else SparsePauliOp([], num_qubits=self.num_qubits)
        finally:
            scratch.release()
        for t in self.tables.values():
            t.load(res, None, keep_snapshot=True)  # (a sharded table keeps the rows this rank owns)
        self._sync_counts()
        for f in ("n_in", "raw_rows", "updates", "num_unique", "num_duplicate", "num_trimmed"):
            setattr(out, f, int(getattr(st, f)))
        out.sum_trimmed = [float(st.sum_trimmed[0]), float(st.sum_trimmed[1])]
        out.n_out = self.n
        if self.n == 0:
            self.zero_operator = True
        return out

    def apply_reset(self, qubit: int, atol: float) -> _Stats:
        """``apply_reset_to`` + simplify: local when the cleared keys keep their owner, else the rows with Z
        on the qubit move to the peer rank (staged record exchange)."""
        import torch

        out = _Stats()
        if self.zero_operator:
            out.n_in = out.raw_rows = out.num_trimmed = out.n_out = out.updates = 1
            return out
        if not self.split:
            return self._rep_stats({r: t.apply_reset(qubit, atol) for r, t in self.tables.items()})
        n_in = self.n
        if atol == OBP_ATOL_RAW:  # simplify=False: zero / clear in place on every rank, nothing merges
            for t in self.tables.values():
                t.apply_reset(qubit, atol)
            self._sync_counts()
            out.n_in = out.raw_rows = out.updates = out.n_out = n_in
            return out
        tdev = torch.device("cuda", self.device)
        bufs, counts, emit, failed, peer = {}, {}, {}, False, 0
        for r, t in self.tables.items():
            cap = max(int(t.n_dense), 1)
            bufs[r] = torch.empty((cap, self.rec_u64), dtype=torch.int64, device=tdev)
            emit[r], peer, over = t.reset_emit(qubit, atol, bufs[r].data_ptr(), cap)
            failed |= over
            counts[r] = min(int(emit[r].n_records), cap)
        if peer == 0:  # every rank sees the same hash frame: all local
            stats = {r: t.apply_reset(qubit, atol) for r, t in self.tables.items()}
            ints = {r: [0, t.n, t.n_dense, 0, int(stats[r].n_in), int(stats[r].num_unique), int(stats[r].num_duplicate),
                        int(stats[r].num_trimmed)] for r, t in self.tables.items()}
            floats = {r: list(stats[r].sum_trimmed) for r in self.tables}
            ti, tf = self.comm.allreduce_pair(ints, floats)
            out.n_in, out.num_unique, out.num_duplicate, out.num_trimmed = (int(v) for v in ti[4:8])
        else:
            recv = self.comm.exchange(bufs, counts, peer)
            merge = {}
            for r, t in self.tables.items():
                buf, n_rec = recv[r]
                merge[r], over = t.reset_merge(atol, buf.data_ptr() if n_rec else 0, n_rec)
                failed |= over
            first = min(self.tables)
            ints = {r: [1 if (failed and r == first) else 0, t.n, t.n_dense, 0, int(emit[r].n_in), int(emit[r].rows_surv),
                        int(emit[r].keys_present + merge[r].keys_present), int(merge[r].cancelled)]
                    for r, t in self.tables.items()}
            floats = {r: [merge[r].sum_trimmed[0], merge[r].sum_trimmed[1]] for r in self.tables}
            ti, tf = self.comm.allreduce_pair(ints, floats)
            if ti[0] > 0:
                raise CapacityError("sharded reset: a shard overflowed its table")
            g_in, rows_surv, keys, cancelled = (int(v) for v in ti[4:8])
            out.n_in, out.num_unique, out.num_duplicate = g_in, keys, rows_surv - keys
            out.num_trimmed = g_in - rows_surv + cancelled
        self.n, self.n_dense = int(ti[1]), int(ti[2])
        out.sum_trimmed = [float(tf[0]), float(tf[1])]
        out.raw_rows = out.updates = n_in
        out.n_out = self.n
        if self.n == 0:
            self.zero_operator = True
        return out

    def apply_damping(self, gen_x, gen_z, factors, atol: float) -> _Stats:
        out = _Stats()
        if self.zero_operator:
            out.n_in = out.raw_rows = out.num_trimmed = out.n_out = out.updates = 1
            return out
        if not self.split:
            return self._rep_stats({r: t.apply_damping(gen_x, gen_z, factors, atol) for r, t in self.tables.items()})
        n_in = self.n
        for t in self.tables.values():
            t.apply_damping(gen_x, gen_z, factors, atol)
        self._sync_counts()
        out.n_in = out.raw_rows = out.updates = n_in
        out.num_unique, out.num_trimmed, out.n_out = self.n, n_in - self.n, self.n
        if self.n == 0:
            self.zero_operator = True
        return out

    def truncate(self, budget: float, p_norm: int, tol: float) -> tuple[float, int]:
        """``truncate_binary_search`` (``utils/truncating.py:171-204``) with allreduced maxima and sums."""
        if self.zero_operator:
            self.zero_operator = False
            self.n = 0
            return 0.0, 1
        if not self.split:
            res = {r: t.truncate(budget, p_norm, tol) for r, t in self.tables.items()}
            self._sync_counts()
            return res[min(res)]
        before = self.n
        upper = float(self.comm.allreduce_max({r: [t.trunc_begin(p_norm)] for r, t in self.tables.items()})[0])
        lower, upper_error, lower_error = 0.0, budget, 0.0
        while (upper - lower) > tol and not np.isclose(upper_error, lower_error, atol=tol):
            mid = (upper + lower) / 2
            s = float(self.comm.allreduce_sum({r: [t.trunc_masked_sum(mid)] for r, t in self.tables.items()})[0])
            mid_error = s if p_norm == 1 else float(np.power(s, 1.0 / p_norm))
            if mid_error <= budget:
                lower, lower_error = mid, mid_error
            else:
                upper, upper_error = mid, mid_error
        for t in self.tables.values():
            t.trunc_finish(lower)
        self._sync_counts()
        return float(lower_error), before - self.n

    def snapshot(self):
        """Commit.  The first commit at or above ``split_min`` terms also ends the replicated phase."""
        if not self.split and not self.zero_operator and self.comm.n_ranks > 1 and self.n >= self.split_min:
            self._split_now()  # (commits the shards itself)
            return
        self._snap = (self.zero_operator, self.n)
        for t in self.tables.values():
            t.snapshot()

    def restore(self):
        for t in self.tables.values():
            t.restore()
        self.zero_operator, self.n = self._snap
        if self._split_on_restore and not self.split:
            # the replicated table overflowed in the slice just rolled back: shard the committed state instead of
            # growing every rank's copy; the caller's reserve() that follows is then unnecessary
            self._split_now()
            self._skip_next_reserve = True

    def reserve(self, capacity: int):
        if self._skip_next_reserve:
            self._skip_next_reserve = False
            return
        had_inboxes = self.p2p and self._inboxes_ready
        if had_inboxes:
            self._drop_inboxes()
        for t in self.tables.values():
            t.reserve(self._per_rank(capacity))
        self.capacity = max(self.capacity, int(capacity))
        if had_inboxes:
            self._setup_inboxes(max(t.capacity for t in self.tables.values()))

    # -- measurement ------------------------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        for t in self.tables.values():
            t.profile_enable(on)

    def profile(self, reset: bool = False) -> dict:
        tot: dict = {}
        for t in self.tables.values():
            for k, v in t.profile(reset).items():
                a = tot.setdefault(k, {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0})
                for f in a:
                    a[f] += v[f]
        return tot

    def timer_start(self):
        for t in self.tables.values():
            t.timer_start()

    def timer_stop(self) -> float:
        return max(t.timer_stop() for t in self.tables.values())


def communicator_from_env(device: int):
    """``OBP_B200_SHARDS=R``: R virtual ranks on this GPU; ``OBP_B200_SHARDED=1``: torch.distributed ranks."""
    import os

    virt = int(os.environ.get("OBP_B200_SHARDS", "0") or 0)
    if virt > 1:
        key = ("local", virt, device)
        if key not in _COMMS:
            _COMMS[key] = LocalComm(virt, device)
        return _COMMS[key]
    if os.environ.get("OBP_B200_SHARDED", "") not in ("", "0"):
        import torch.distributed as dist

        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("OBP_B200_SHARDED=1 needs an initialised torch.distributed process group")
        if dist.get_world_size() > 1:
            key = ("dist", dist.get_world_size(), device)
            if key not in _COMMS:
                _COMMS[key] = DistComm()
            return _COMMS[key]
    return None


#: communicators are kept for the life of the process: they own the pool of idle sharded tables
_COMMS: dict = {}


def close_all() -> None:
    """Free every pooled sharded table (collective in the distributed case: all ranks must call it)."""
    for comm in _COMMS.values():
        for t in comm.__dict__.pop("_idle_tables", []):
            t.close()
