"""ctypes binding of ``libobp_b200.so`` (C-ABI: ``include/obp_b200.h``) and the ``TermTable`` wrapper.

There is no CPU fallback: if the shared library is missing or no sm_100a device is present, every
entry point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``.
"""

from __future__ import annotations

import atexit
import ctypes as C
import os
from contextlib import contextmanager

import numpy as np

from .ir import SparsePauliOp, pack_xz
from .utils.simplify import SimplifyMetadata

_LIB_PATH = os.environ.get(
    "OBP_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libobp_b200.so")
)  # the override selects another build of the same CUDA library (kernel A/B runs)

OBP_OK = 0
OBP_ERR_OOM = -3
OBP_ERR_CAPACITY = -4
OBP_OP_CLIFFORD = 1
OBP_OP_ROTATION = 2
OBP_OP_GENERIC = 3  # host-side tag only: run through obp_apply_generic, never inside an obp_op program
OBP_FLAG_EXACT_COUNTERS = 1
OBP_FLAG_OWN_SLOT = 2
OBP_ATOL_MERGE_ONLY = -1.0
OBP_ATOL_RAW = -2.0  # simplify=False: no merge, exact zeros stay
OBP_OPT_PROGRAM = 1
OBP_OPT_PROGRAM_MAX_ROWS = 2
OBP_OPT_SEGMENTS = 3
OBP_OPT_SEGMENT_MIN_GATES = 4
OBP_OPT_SEGMENT_CAP_ROWS = 5
OBP_OPT_SEGMENT_PIECES_MIN_ROWS = 6
OBP_OPT_SEGMENT_MIXED = 7
OBP_K_NAMES = ("rotate", "clifford", "truncate", "compact", "index", "other", "program", "segment")

#: numpy mirror of ``obp_op`` (104 bytes, natural C alignment)
OP_DTYPE = np.dtype(
    [
        ("kind", "<i4"),
        ("n_qubits", "<i4"),
        ("qubit", "<i4", (4,)),
        ("flags", "<u4"),
        ("n_components", "<u4"),
        ("lut", "<u8"),
        ("sign_mask", "<u4"),
        ("group_size", "<u4"),
        ("group_elems", "<u8"),
        ("row_scale", "<f8"),
        ("gen_local", "<u4"),
        ("reserved", "<u4"),
        ("u0", "<f8", (2,)),
        ("u1", "<f8", (2,)),
    ],
    align=True,
)
assert OP_DTYPE.itemsize == 104


class ObpStats(C.Structure):
    _fields_ = [
        ("n_in", C.c_uint64),
        ("raw_rows", C.c_uint64),
        ("num_unique", C.c_uint64),
        ("num_duplicate", C.c_uint64),
        ("num_trimmed", C.c_uint64),
        ("sum_trimmed", C.c_double * 2),
        ("n_out", C.c_uint64),
        ("n_dense", C.c_uint64),
        ("updates", C.c_uint64),
    ]


class ObpProfile(C.Structure):
    _fields_ = [
        ("launches", C.c_uint64 * 8),
        ("passes", C.c_uint64 * 8),
        ("ms", C.c_double * 8),
        ("term_updates", C.c_uint64 * 8),
    ]


class ObpShardCounters(C.Structure):
    _fields_ = [
        ("n_in", C.c_uint64),
        ("n_out", C.c_uint64),
        ("n_dense", C.c_uint64),
        ("rows_surv", C.c_uint64),
        ("keys_present", C.c_uint64),
        ("cancelled", C.c_uint64),
        ("n_records", C.c_uint64),
        ("sum_trimmed", C.c_double * 2),
    ]


class EngineError(RuntimeError):
    """A CUDA or engine failure reported through the C-ABI."""


class CapacityError(MemoryError):
    """The device term table (or its index) is full; the table must be restored and grown."""


_lib = None


def lib() -> C.CDLL:
    """Load ``libobp_b200.so`` once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: the CUDA extension is not built. There is no CPU fallback; "
            'run `python -c "import __graft_entry__ as g; g.build()"` at the repository root.'
        )
    L = C.CDLL(_LIB_PATH)
    vp, u64p, f64p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    sig = {
        "obp_create": ([C.c_int, C.c_int, C.c_uint64, C.POINTER(vp)], C.c_int),
        "obp_destroy": ([vp], None),
        "obp_last_error": ([vp], C.c_char_p),
        "obp_set_stream": ([vp, vp], C.c_int),
        "obp_device_memory": ([C.c_int, u64p, u64p], C.c_int),
        "obp_bytes_per_term": ([C.c_int], C.c_uint64),
        "obp_load": ([vp, vp, vp, vp, C.c_uint64, C.c_double, C.POINTER(ObpStats)], C.c_int),
        "obp_replace": ([vp, vp, vp, vp, C.c_uint64, C.c_double, C.POINTER(ObpStats)], C.c_int),
        "obp_count": ([vp, u64p], C.c_int),
        "obp_download": ([vp, vp, vp, vp, C.c_uint64, u64p], C.c_int),
        "obp_apply_ops": ([vp, vp, C.c_int32, C.c_double, C.POINTER(ObpStats)], C.c_int),
        "obp_apply_ops_submit": ([vp, vp, C.c_int32, C.c_double], C.c_int),
        "obp_apply_ops_wait": ([vp, C.POINTER(ObpStats)], C.c_int),
        "obp_last_op_sizes": ([vp, u64p, C.c_int32, C.POINTER(C.c_int32)], C.c_int),
        "obp_apply_generic": ([vp, C.c_int32, vp, C.c_int32, vp, vp, C.c_double, C.c_int32, C.POINTER(ObpStats)], C.c_int),
        "obp_apply_reset": ([vp, C.c_int32, C.c_double, C.POINTER(ObpStats)], C.c_int),
        "obp_apply_damping": ([vp, vp, vp, vp, C.c_int32, C.c_double, C.POINTER(ObpStats)], C.c_int),
        "obp_truncate": ([vp, C.c_double, C.c_int32, C.c_double, f64p, u64p, u64p], C.c_int),
        "obp_union_count": ([C.POINTER(vp), C.c_int32, u64p], C.c_int),
        "obp_qwc_group_count": ([C.c_int, C.c_int, vp, vp, C.c_uint64, u64p], C.c_int),
        "obp_snapshot": ([vp], C.c_int),
        "obp_restore": ([vp], C.c_int),
        "obp_reserve": ([vp, C.c_uint64], C.c_int),
        "obp_set_row_order": ([vp, C.c_int32], C.c_int),
        "obp_configure": ([vp, C.c_int32, C.c_int64], C.c_int),
        "obp_is_broken": ([vp], C.c_int),
        "obp_profile_enable": ([vp, C.c_int32], C.c_int),
        "obp_profile_get": ([vp, C.POINTER(ObpProfile), C.c_int32], C.c_int),
        "obp_set_shard": ([vp, C.c_int32, C.c_int32], C.c_int),
        "obp_record_bytes": ([C.c_int], C.c_uint64),
        "obp_shard_stats": ([vp, u64p, u64p], C.c_int),
        "obp_segment_stats": ([vp, u64p, u64p, u64p], C.c_int),
        "obp_last_op_stats": ([vp, C.c_int32, C.POINTER(ObpStats)], C.c_int),
        "obp_segment_partition_map": ([C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "obp_shard_split": ([vp, C.c_int32, C.c_int32, vp, C.c_int32, u64p, C.POINTER(C.c_int32)], C.c_int),
        "obp_apply_ops_sharded": (
            [vp, vp, C.c_int32, C.c_double, C.POINTER(ObpStats), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
            C.c_int,
        ),
        "obp_rotate_emit": ([vp, vp, C.c_double, vp, C.c_uint64, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_rotate_merge": ([vp, vp, C.c_uint64, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_inbox_create": ([vp, C.c_uint64], C.c_int),
        "obp_inbox_release": ([vp, C.c_int32], C.c_int),
        "obp_inbox_export": ([vp, C.c_int32, vp, u64p], C.c_int),
        "obp_inbox_attach": ([vp, C.c_int32, C.c_int32, vp, C.c_uint64], C.c_int),
        "obp_rotate_emit_p2p": ([vp, vp, C.c_double, C.c_int32, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_rotate_merge_inbox": ([vp, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_reset_emit": ([vp, C.c_int32, C.c_double, vp, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(ObpShardCounters)], C.c_int),
        "obp_reset_merge": ([vp, C.c_double, vp, C.c_uint64, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_set_shard_sync": ([vp, C.c_int32], C.c_int),
        "obp_last_split_counters": ([vp, C.POINTER(ObpShardCounters)], C.c_int),
        "obp_trunc_begin": ([vp, C.c_int32, f64p], C.c_int),
        "obp_trunc_masked_sum": ([vp, C.c_double, f64p], C.c_int),
        "obp_trunc_finish": ([vp, C.c_double, u64p, u64p], C.c_int),
        "obp_timer_start": ([vp], C.c_int),
        "obp_timer_stop": ([vp, f64p], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "obp_create obp_destroy obp_last_error obp_set_stream obp_device_memory obp_bytes_per_term "
    "obp_load obp_count obp_download obp_apply_ops obp_apply_generic obp_apply_reset obp_apply_damping obp_truncate "
    "obp_union_count obp_snapshot obp_restore obp_reserve obp_set_row_order obp_apply_ops_submit obp_apply_ops_wait obp_replace obp_profile_enable obp_profile_get "
    "obp_timer_start obp_timer_stop obp_set_shard obp_record_bytes obp_apply_ops_sharded obp_rotate_emit "
    "obp_rotate_merge obp_trunc_begin obp_trunc_masked_sum obp_trunc_finish obp_inbox_create obp_inbox_release "
    "obp_inbox_export obp_inbox_attach obp_rotate_emit_p2p obp_rotate_merge_inbox obp_set_shard_sync "
    "obp_last_split_counters obp_reset_emit obp_reset_merge obp_configure obp_is_broken obp_last_op_sizes obp_shard_split obp_qwc_group_count obp_shard_stats obp_segment_stats obp_segment_partition_map obp_last_op_stats"
).split()


def _host_buffer(shape, dtype) -> np.ndarray:
    """Result array for a download: page-locked (through torch's caching host allocator) when torch is
    there, so the device-to-host copy lands in it directly at full PCIe rate; plain numpy otherwise."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes >= (1 << 16) and os.environ.get("OBP_B200_PINNED_RESULT", "1") != "0":
        try:
            import torch

            if torch.cuda.is_available():
                return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy().view(dtype).reshape(shape)
        except Exception:  # noqa: BLE001 - torch missing or pinned allocation refused: pageable memory works too
            pass
    return np.empty(shape, dtype=dtype)


def device_memory(device: int = 0) -> tuple[int, int]:
    free, total = C.c_uint64(), C.c_uint64()
    rc = lib().obp_device_memory(device, C.byref(free), C.byref(total))
    if rc != OBP_OK:
        raise EngineError((lib().obp_last_error(None) or b"").decode())
    return free.value, total.value


def default_device() -> int:
    """One process per GPU: LOCAL_RANK selects the device (``torchrun`` sets it)."""
    return int(os.environ.get("OBP_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


#: idle tables kept for reuse, keyed by (device, num_qubits): device allocation and release cost
#: far more than a whole 1e6-term backpropagation, so contexts are recycled across calls
_POOL: dict[tuple[int, int], list["TermTable"]] = {}
POOL_MAX_PER_KEY = 4


def clear_pool() -> None:
    """Destroy all idle pooled tables (frees their device memory)."""
    for tabs in _POOL.values():
        for t in tabs:
            t.close()
    _POOL.clear()


atexit.register(clear_pool)


class TermTable:
    """One observable as a device-resident table of packed Pauli keys and complex128 coefficients."""

    @classmethod
    def acquire(cls, num_qubits: int, capacity: int, device: int | None = None) -> "TermTable":
        """A table with room for ``capacity`` terms: a recycled one when the pool has a fit."""
        device = default_device() if device is None else int(device)
        idle = _POOL.get((device, int(num_qubits)), [])
        for k, t in enumerate(idle):
            if t.capacity >= capacity:
                idle.pop(k)
                t.n = t.n_dense = 0
                t.zero_operator = False
                t.track_zero = True
                t.set_shard(1, 0)
                t.set_shard_sync(False)
                t.set_row_order(False)
                t._apply_env_options()
                return t
        t = cls(num_qubits, capacity, device)
        t._apply_env_options()
        return t

    def _apply_env_options(self) -> None:
        """Code-path switches from the environment, re-read at every acquire (pooled tables outlive a test):
        ``OBP_B200_NO_PROGRAM=1`` — one kernel per gate instead of the persistent program kernel;
        ``OBP_B200_PROG_MAX_ROWS=n`` — tables of at least n rows use the stand-alone kernels (default 2^18)."""
        no_prog = os.environ.get("OBP_B200_NO_PROGRAM", "0") not in ("", "0")
        self.configure(OBP_OPT_PROGRAM, 0 if no_prog else 1)
        self.configure(OBP_OPT_PROGRAM_MAX_ROWS, int(os.environ.get("OBP_B200_PROG_MAX_ROWS", str(1 << 18))))
        # runs of rotations applied in shared memory (OBP_B200_NO_SEGMENTS=1: one pass over the table per gate)
        self.configure(OBP_OPT_SEGMENTS, 0 if os.environ.get("OBP_B200_NO_SEGMENTS", "0") not in ("", "0") else 1)
        self.configure(OBP_OPT_SEGMENT_MIN_GATES, int(os.environ.get("OBP_B200_SEG_MIN_GATES", "6")))
        self.configure(OBP_OPT_SEGMENT_CAP_ROWS, int(os.environ.get("OBP_B200_SEG_CAP_ROWS", "0")))
        self.configure(OBP_OPT_SEGMENT_PIECES_MIN_ROWS, int(os.environ.get("OBP_B200_SEG_PIECES_MIN_ROWS", str(1 << 17))))
        self.configure(OBP_OPT_SEGMENT_MIXED, 1 if os.environ.get("OBP_B200_SEG_MIXED", "0") not in ("", "0") else 0)

    def configure(self, option: int, value: int) -> None:
        self._check(self._lib.obp_configure(self._h, int(option), int(value)), "obp_configure")

    def _drain(self) -> None:
        """Wait for (and discard the result of) a gate program that was submitted but never waited for —
        an exception (``max_seconds`` alarm, KeyboardInterrupt) can arrive between the two halves."""
        if getattr(self, "_submitted", False):
            self._submitted = False
            self._submitted_zero = None
            self._lib.obp_apply_ops_wait(self._h, None)  # a device error here resurfaces in the next call

    def release(self) -> None:
        """Hand the table back to the pool (or destroy it when the pool is full or the context is unusable)."""
        if not self._h.value:
            return
        self._drain()
        idle = _POOL.setdefault((self.device, self.num_qubits), [])
        if len(idle) >= POOL_MAX_PER_KEY or self._lib.obp_is_broken(self._h):
            self.close()
        else:
            idle.append(self)

    def __init__(self, num_qubits: int, capacity: int, device: int | None = None):
        self._lib = lib()
        self.num_qubits = int(num_qubits)
        self.words = max(1, (self.num_qubits + 63) // 64)
        self.capacity = int(capacity)
        self.device = default_device() if device is None else int(device)
        self._h = C.c_void_p()
        rc = self._lib.obp_create(self.device, self.num_qubits, self.capacity, C.byref(self._h))
        if rc != OBP_OK:
            msg = (self._lib.obp_last_error(None) or b"").decode()
            self._h = C.c_void_p()
            raise (MemoryError if rc == -3 else EngineError)(f"obp_create failed ({rc}): {msg}")
        self.n = 0  # live terms
        self.n_dense = 0
        self.zero_operator = False  # simplify.py:156-159: everything trimmed -> one identity row, coeff 0
        self.track_zero = True  # False for the members of a sharded table: an empty shard is not a zero operator

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc == OBP_OK:
            return
        msg = (self._lib.obp_last_error(self._h) or b"").decode()
        if rc == OBP_ERR_CAPACITY:
            raise CapacityError(f"{what}: {msg}")
        if rc == OBP_ERR_OOM:
            raise MemoryError(f"{what}: {msg}")
        raise EngineError(f"{what} failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._drain()
            self._lib.obp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        self._check(self._lib.obp_set_stream(self._h, C.c_void_p(cuda_stream or 0)), "obp_set_stream")

    def __len__(self) -> int:
        return 1 if self.zero_operator else self.n

    # -- I/O ----------------------------------------------------------------------------------
    def load(self, op: SparsePauliOp, atol: float | None = None, keep_snapshot: bool = False) -> ObpStats:
        """Upload ``op``; ``atol=None`` merges duplicate rows only, a float >= 0 runs the full simplify,
        ``OBP_ATOL_RAW`` keeps the rows exactly as given.  ``keep_snapshot``: the last commit stays restorable
        (``obp_replace``: the content is replaced in the middle of a slice)."""
        if op.num_qubits != self.num_qubits:
            raise ValueError("operator width does not match the table")
        if len(op) > self.capacity:
            self.reserve(2 * len(op))
        xw, zw = op.packed()
        coeffs = np.ascontiguousarray(op.coeffs, dtype=np.complex128)
        st = ObpStats()
        rc = (self._lib.obp_replace if keep_snapshot else self._lib.obp_load)(
            self._h,
            xw.ctypes.data_as(C.c_void_p),
            zw.ctypes.data_as(C.c_void_p),
            coeffs.ctypes.data_as(C.c_void_p),
            len(op),
            -1.0 if atol is None else float(atol),
            C.byref(st),
        )
        self._check(rc, "obp_load")
        self.n, self.n_dense = st.n_out, st.n_dense
        self.zero_operator = self.track_zero and atol is not None and atol >= 0 and self.n == 0
        return st

    def download(self) -> SparsePauliOp:
        self._drain()
        if self.zero_operator:
            z = np.zeros((1, self.num_qubits), dtype=bool)
            return SparsePauliOp((z.copy(), z), np.array([0j]))
        n = self.n
        xw = _host_buffer((max(n, 1), self.words), "<u8")
        zw = _host_buffer((max(n, 1), self.words), "<u8")
        co = _host_buffer((max(n, 1),), np.complex128)
        got = C.c_uint64()
        rc = self._lib.obp_download(
            self._h,
            xw.ctypes.data_as(C.c_void_p),
            zw.ctypes.data_as(C.c_void_p),
            co.ctypes.data_as(C.c_void_p),
            max(n, 1),
            C.byref(got),
        )
        self._check(rc, "obp_download")
        n = got.value
        self.n = self.n_dense = n
        # the result stays in the engine's packed format; bool x/z arrays are made on first access
        return SparsePauliOp.from_packed(xw[:n], zw[:n], co[:n], self.num_qubits)

    to_operator = download

    # -- hot path -----------------------------------------------------------------------------
    def apply_ops(self, ops: np.ndarray, atol: float) -> ObpStats:
        """Run a gate program (array of ``OP_DTYPE``) on the table."""
        st = ObpStats()
        if self.zero_operator:
            # the reference keeps conjugating the single zero row: every raw row is filtered
            last = ops[-1]
            k2 = int(last["n_components"]) ** 2
            st.n_in, st.raw_rows, st.num_trimmed, st.n_out = 1, k2, k2, 1
            st.updates = len(ops)
            return st
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        rc = self._lib.obp_apply_ops(
            self._h, ops.ctypes.data_as(C.c_void_p), len(ops), float(atol), C.byref(st)
        )
        try:
            self._check(rc, "obp_apply_ops")
        except CapacityError as exc:
            exc.updates = int(st.updates)  # the kernels ran before the overflow was detected
            exc.n_live = int(st.n_out)  # live rows at that moment
            raise
        self.n, self.n_dense = st.n_out, st.n_dense
        if self.n == 0:
            self.zero_operator = self.track_zero
        return st

    def apply_ops_submit(self, ops: np.ndarray, atol: float) -> None:
        """First half of :meth:`apply_ops`: enqueue the program and return without waiting for the device."""
        self._submitted_zero = None
        if self.zero_operator:
            last = ops[-1]
            self._submitted_zero = (int(last["n_components"]) ** 2, len(ops))
            return
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        self._submitted = True  # (set first: an asynchronous exception right after the call must still find it)
        rc = self._lib.obp_apply_ops_submit(self._h, ops.ctypes.data_as(C.c_void_p), len(ops), float(atol))
        if rc != 0:
            self._submitted = False
        self._check(rc, "obp_apply_ops_submit")

    def apply_ops_wait(self) -> ObpStats:
        """Second half of :meth:`apply_ops`: wait for the submitted program, return its statistics."""
        st = ObpStats()
        if getattr(self, "_submitted_zero", None) is not None:
            k2, n_ops = self._submitted_zero
            self._submitted_zero = None
            st.n_in, st.raw_rows, st.num_trimmed, st.n_out = 1, k2, k2, 1
            st.updates = n_ops
            return st
        rc = self._lib.obp_apply_ops_wait(self._h, C.byref(st))
        self._submitted = False
        try:
            self._check(rc, "obp_apply_ops_wait")
        except CapacityError as exc:
            exc.updates = int(st.updates)
            exc.n_live = int(st.n_out)
            raise
        self.n, self.n_dense = st.n_out, st.n_dense
        if self.n == 0:
            self.zero_operator = self.track_zero
        return st

    def last_op_stats(self, op_index: int) -> ObpStats:
        """Counters of one op of the program just waited for (it must have carried ``OBP_FLAG_EXACT_COUNTERS``)."""
        st = ObpStats()
        self._check(self._lib.obp_last_op_stats(self._h, int(op_index), C.byref(st)), "obp_last_op_stats")
        return st

    def last_op_sizes(self, n_ops: int) -> list[int]:
        """Live term count after every op of the program just waited for (exact per gate when the ops carried
        ``OBP_FLAG_OWN_SLOT``): what the reference logs at DEBUG level after each gate."""
        if self.zero_operator or n_ops == 0:
            return [1 if self.zero_operator else self.n] * n_ops
        buf = (C.c_uint64 * n_ops)()
        got = C.c_int32()
        self._check(self._lib.obp_last_op_sizes(self._h, buf, n_ops, C.byref(got)), "obp_last_op_sizes")
        if got.value != n_ops:  # e.g. the program ran while the operator was already the zero operator
            return [len(self)] * n_ops
        return [int(v) for v in buf]

    def apply_generic(self, qubits, codes, coeffs, atol: float, simplify: bool = True) -> ObpStats:
        """A gate given as its Pauli sum (local codes + complex coefficients): raw k*k expansion (+ simplify)."""
        st = ObpStats()
        k2 = len(codes) ** 2
        if self.zero_operator:
            st.n_in, st.raw_rows, st.num_trimmed, st.n_out, st.updates = 1, k2, k2, 1, 1
            return st
        q = np.ascontiguousarray(qubits, dtype=np.int32)
        cd = np.ascontiguousarray(codes, dtype=np.uint32)
        co = np.ascontiguousarray(np.asarray(coeffs, dtype=np.complex128))
        rc = self._lib.obp_apply_generic(
            self._h, len(q), q.ctypes.data_as(C.c_void_p), len(cd), cd.ctypes.data_as(C.c_void_p),
            co.ctypes.data_as(C.c_void_p), float(atol), int(bool(simplify)), C.byref(st),
        )
        self._check(rc, "obp_apply_generic")
        self.n, self.n_dense = st.n_out, st.n_dense
        if self.n == 0 and simplify:
            self.zero_operator = self.track_zero
        return st

    def apply_reset(self, qubit: int, atol: float) -> ObpStats:
        st = ObpStats()
        if self.zero_operator:
            st.n_in, st.raw_rows, st.num_trimmed, st.n_out, st.updates = 1, 1, 1, 1, 1
            return st
        self._check(self._lib.obp_apply_reset(self._h, int(qubit), float(atol), C.byref(st)), "obp_apply_reset")
        self.n, self.n_dense = st.n_out, st.n_dense
        if self.n == 0 and atol >= 0:
            self.zero_operator = self.track_zero
        return st

    def apply_damping(self, gen_x: np.ndarray, gen_z: np.ndarray, factors: np.ndarray, atol: float) -> ObpStats:
        st = ObpStats()
        if self.zero_operator:
            st.n_in, st.raw_rows, st.num_trimmed, st.n_out, st.updates = 1, 1, 1, 1, 1
            return st
        gxw, gzw = pack_xz(gen_x, gen_z)
        factors = np.ascontiguousarray(factors, dtype=np.float64)
        rc = self._lib.obp_apply_damping(
            self._h,
            gxw.ctypes.data_as(C.c_void_p),
            gzw.ctypes.data_as(C.c_void_p),
            factors.ctypes.data_as(C.c_void_p),
            len(factors),
            float(atol),
            C.byref(st),
        )
        self._check(rc, "obp_apply_damping")
        self.n, self.n_dense = st.n_out, st.n_dense
        if self.n == 0 and atol >= 0:
            self.zero_operator = self.track_zero
        return st

    def truncate(self, budget: float, p_norm: int, tol: float) -> tuple[float, int]:
        """Returns ``(slice_error, n_removed)``."""
        if self.zero_operator:
            # |0|**p == max == 0: the search loop never runs and `0 > 0` keeps nothing; the
            # reference then holds an empty operator
            self.zero_operator = False
            self.n = 0
            return 0.0, 1
        err, rem, out = C.c_double(), C.c_uint64(), C.c_uint64()
        rc = self._lib.obp_truncate(
            self._h, float(budget), int(p_norm), float(tol), C.byref(err), C.byref(rem), C.byref(out)
        )
        self._check(rc, "obp_truncate")
        self.n = self.n_dense = out.value
        return err.value, rem.value

    def snapshot(self):
        """Commit point: stream-ordered device copy of the table (no host synchronisation)."""
        self._drain()
        self._snap_zero = self.zero_operator
        self._snap_n = self.n
        self._check(self._lib.obp_snapshot(self._h), "obp_snapshot")

    def restore(self):
        self._drain()
        self._check(self._lib.obp_restore(self._h), "obp_restore")
        self.zero_operator = self._snap_zero
        self.n = self._snap_n

    def set_row_order(self, on: bool):
        """Keep rows in the reference's first-appearance order (every gate must then be `apply_generic`)."""
        self._check(self._lib.obp_set_row_order(self._h, int(bool(on))), "obp_set_row_order")
        self.row_order = bool(on)

    def reserve(self, capacity: int):
        self._check(self._lib.obp_reserve(self._h, int(capacity)), "obp_reserve")
        self.capacity = max(self.capacity, int(capacity))

    # -- measurement --------------------------------------------------------------------------
    def profile_enable(self, on: bool = True):
        self._check(self._lib.obp_profile_enable(self._h, int(on)), "obp_profile_enable")

    def profile(self, reset: bool = False) -> dict:
        p = ObpProfile()
        self._check(self._lib.obp_profile_get(self._h, C.byref(p), int(reset)), "obp_profile_get")
        return {
            name: {
                "launches": int(p.launches[k]),
                "passes": int(p.passes[k]),
                "ms": float(p.ms[k]),
                "term_updates": int(p.term_updates[k]),
            }
            for k, name in enumerate(OBP_K_NAMES)
        }

    # -- sharded table (see sharded.py) ----------------------------------------------------------
    def set_shard(self, n_ranks: int, rank: int):
        self._check(self._lib.obp_set_shard(self._h, int(n_ranks), int(rank)), "obp_set_shard")

    def segment_stats(self) -> tuple[int, int, int]:
        """``(runs of rotations applied in shared memory, failed runs repeated in shorter pieces, runs handed to the
        per-gate kernels)`` since the context was created."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.obp_segment_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "obp_segment_stats")
        return a.value, b.value, c.value

    def shard_stats(self) -> tuple[int, int]:
        """``(rotations that needed a record exchange, records emitted)`` since the context was created."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self._lib.obp_shard_stats(self._h, C.byref(a), C.byref(b)), "obp_shard_stats")
        return a.value, b.value

    def shard_split(self, n_ranks: int, rank: int, future_ops: np.ndarray | None = None) -> tuple[int, int]:
        """A replicated table keeps only the rows rank ``rank`` of ``n_ranks`` owns (late sharding).  With
        ``future_ops`` (the gate records the run applies next) the shard function is planned so that none of
        their rotations needs a record exchange.  Returns ``(rows kept, rotations covered by the plan)``."""
        out, cov = C.c_uint64(), C.c_int32()
        ops_p, n_ops = None, 0
        if future_ops is not None and len(future_ops):
            future_ops = np.ascontiguousarray(future_ops, dtype=OP_DTYPE)
            ops_p, n_ops = future_ops.ctypes.data_as(C.c_void_p), len(future_ops)
        self._check(
            self._lib.obp_shard_split(self._h, int(n_ranks), int(rank), ops_p, n_ops, C.byref(out), C.byref(cov)),
            "obp_shard_split",
        )
        self.n = self.n_dense = out.value
        return out.value, cov.value

    def apply_ops_sharded(self, ops: np.ndarray, atol: float) -> tuple[ObpStats, int, int, bool]:
        """Run ``ops`` up to the first rotation that needs a peer: ``(stats, n_done, peer_xor, overflowed)``.

        A table overflow is reported as a flag instead of an exception: all ranks must stay in
        lock-step through the exchange that follows, the caller raises once they have agreed.
        """
        st, done, peer = ObpStats(), C.c_int32(), C.c_int32()
        ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
        rc = self._lib.obp_apply_ops_sharded(
            self._h, ops.ctypes.data_as(C.c_void_p), len(ops), float(atol), C.byref(st), C.byref(done), C.byref(peer)
        )
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_apply_ops_sharded")
            self.n, self.n_dense = st.n_out, st.n_dense
        else:
            self.n = st.n_out  # live rows when the shard overflowed (the caller rolls back; used as a size lower bound)
        return st, done.value, peer.value, rc == OBP_ERR_CAPACITY

    def rotate_emit(self, op: np.ndarray, atol: float, d_records: int, cap_records: int) -> tuple[ObpShardCounters, bool]:
        out = ObpShardCounters()
        op = np.ascontiguousarray(op, dtype=OP_DTYPE)
        rc = self._lib.obp_rotate_emit(
            self._h, op.ctypes.data_as(C.c_void_p), float(atol), C.c_void_p(d_records), int(cap_records), C.byref(out)
        )
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_rotate_emit")
        self.n, self.n_dense = out.n_out, out.n_dense
        return out, rc == OBP_ERR_CAPACITY

    def rotate_merge(self, d_records: int, n_records: int) -> tuple[ObpShardCounters, bool]:
        out = ObpShardCounters()
        rc = self._lib.obp_rotate_merge(self._h, C.c_void_p(d_records), int(n_records), C.byref(out))
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_rotate_merge")
        self.n, self.n_dense = out.n_out, out.n_dense
        return out, rc == OBP_ERR_CAPACITY

    def inbox_create(self, cap_records: int):
        self._check(self._lib.obp_inbox_create(self._h, int(cap_records)), "obp_inbox_create")

    def inbox_release(self, stage: int):
        self._check(self._lib.obp_inbox_release(self._h, int(stage)), "obp_inbox_release")

    def inbox_export(self, which: int) -> tuple[bytes, int]:
        """``(64-byte CUDA IPC handle, raw device pointer)`` of inbox ``which``."""
        h, ptr = C.create_string_buffer(64), C.c_uint64()
        self._check(self._lib.obp_inbox_export(self._h, int(which), h, C.byref(ptr)), "obp_inbox_export")
        return h.raw, ptr.value

    def inbox_attach(self, peer_rank: int, which: int, handle: bytes | None, raw_ptr: int):
        buf = C.create_string_buffer(handle, 64) if handle is not None else None
        self._check(
            self._lib.obp_inbox_attach(self._h, int(peer_rank), int(which), buf, int(raw_ptr)), "obp_inbox_attach"
        )

    def rotate_emit_p2p(self, op: np.ndarray, atol: float, peer_rank: int) -> tuple[ObpShardCounters, bool]:
        out = ObpShardCounters()
        op = np.ascontiguousarray(op, dtype=OP_DTYPE)
        rc = self._lib.obp_rotate_emit_p2p(self._h, op.ctypes.data_as(C.c_void_p), float(atol), int(peer_rank), C.byref(out))
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_rotate_emit_p2p")
        self.n, self.n_dense = out.n_out, out.n_dense
        return out, rc == OBP_ERR_CAPACITY

    def rotate_merge_inbox(self) -> tuple[ObpShardCounters, bool]:
        out = ObpShardCounters()
        rc = self._lib.obp_rotate_merge_inbox(self._h, C.byref(out))
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_rotate_merge_inbox")
        self.n, self.n_dense = out.n_out, out.n_dense
        return out, rc == OBP_ERR_CAPACITY

    def reset_emit(self, qubit: int, atol: float, d_records: int, cap_records: int) -> tuple[ObpShardCounters, int, bool]:
        """Phase 1 of a reset whose cleared keys belong to another rank: ``(counters, peer_xor, overflowed)``."""
        out, peer = ObpShardCounters(), C.c_int32()
        rc = self._lib.obp_reset_emit(
            self._h, int(qubit), float(atol), C.c_void_p(d_records), int(cap_records), C.byref(peer), C.byref(out)
        )
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_reset_emit")
        if peer.value:
            self.n, self.n_dense = out.n_out, out.n_dense
        return out, peer.value, rc == OBP_ERR_CAPACITY

    def reset_merge(self, atol: float, d_records: int, n_records: int) -> tuple[ObpShardCounters, bool]:
        out = ObpShardCounters()
        rc = self._lib.obp_reset_merge(self._h, float(atol), C.c_void_p(d_records), int(n_records), C.byref(out))
        if rc != OBP_ERR_CAPACITY:
            self._check(rc, "obp_reset_merge")
        self.n, self.n_dense = out.n_out, out.n_dense
        return out, rc == OBP_ERR_CAPACITY

    def set_shard_sync(self, device_flags: bool):
        self._check(self._lib.obp_set_shard_sync(self._h, int(bool(device_flags))), "obp_set_shard_sync")

    def last_split_counters(self) -> ObpShardCounters | None:
        """Counters of the split rotation that ended the last program (device-flag mode), else ``None``."""
        out = ObpShardCounters()
        self._check(self._lib.obp_last_split_counters(self._h, C.byref(out)), "obp_last_split_counters")
        return out if out.n_in else None

    def trunc_begin(self, p_norm: int) -> float:
        mx = C.c_double()
        self._check(self._lib.obp_trunc_begin(self._h, int(p_norm), C.byref(mx)), "obp_trunc_begin")
        return mx.value

    def trunc_masked_sum(self, threshold: float) -> float:
        sm = C.c_double()
        self._check(self._lib.obp_trunc_masked_sum(self._h, float(threshold), C.byref(sm)), "obp_trunc_masked_sum")
        return sm.value

    def trunc_finish(self, threshold: float) -> int:
        rem, out = C.c_uint64(), C.c_uint64()
        self._check(self._lib.obp_trunc_finish(self._h, float(threshold), C.byref(rem), C.byref(out)), "obp_trunc_finish")
        self.n = self.n_dense = out.value
        return rem.value

    def timer_start(self):
        self._check(self._lib.obp_timer_start(self._h), "obp_timer_start")

    def timer_stop(self) -> float:
        """Milliseconds of device time since :meth:`timer_start` (CUDA events on the table's stream)."""
        ms = C.c_double()
        self._check(self._lib.obp_timer_stop(self._h, C.byref(ms)), "obp_timer_stop")
        return ms.value

    # -- constructors used by the utils API ------------------------------------------------
    @classmethod
    @contextmanager
    def from_operator(cls, op: SparsePauliOp, capacity: int | None = None):
        table = cls.acquire(op.num_qubits, capacity or max(1024, 2 * len(op)))
        try:
            table.load(op)
            yield table
        finally:
            table.release()

    @classmethod
    @contextmanager
    def from_rows(cls, op: SparsePauliOp, atol: float):
        """Load raw rows with the full simplify semantics; yields ``(table, SimplifyMetadata)``."""
        table = cls.acquire(op.num_qubits, max(1024, 2 * len(op)))
        try:
            st = table.load(op, atol=atol)
            s = complex(st.sum_trimmed[0], st.sum_trimmed[1])
            md = SimplifyMetadata(
                num_unique_paulis=int(st.num_unique),
                num_duplicate_paulis=int(st.num_duplicate),
                num_trimmed_paulis=int(st.num_trimmed),
                sum_trimmed_coeffs=s if s != 0 else 0.0,
            )
            yield table, md
        finally:
            table.release()


def qwc_group_count(op: SparsePauliOp, device: int | None = None) -> int:
    """``len(op.group_commuting(qubit_wise=True))`` on the device (``backpropagation.py:368-375, 433-440``)."""
    xw, zw = op.packed()
    xw = np.ascontiguousarray(xw, dtype="<u8")
    zw = np.ascontiguousarray(zw, dtype="<u8")
    out = C.c_uint64()
    rc = lib().obp_qwc_group_count(
        default_device() if device is None else int(device), op.num_qubits,
        xw.ctypes.data_as(C.c_void_p), zw.ctypes.data_as(C.c_void_p), len(op), C.byref(out),
    )
    if rc != OBP_OK:
        msg = (lib().obp_last_error(None) or b"").decode()
        raise (MemoryError if rc == OBP_ERR_OOM else EngineError)(f"obp_qwc_group_count failed ({rc}): {msg}")
    return int(out.value)


def union_count(tables: list[TermTable]) -> int:
    """Distinct Pauli keys over several tables (``backpropagation.py:419-428``)."""
    arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
    out = C.c_uint64()
    rc = lib().obp_union_count(arr, len(tables), C.byref(out))
    tables[0]._check(rc, "obp_union_count")
    return int(out.value)
