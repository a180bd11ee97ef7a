/*
 * obp_b200.h — C-ABI of the B200-native operator-backpropagation engine.
 *
 * The reference (Qiskit/qiskit-addon-obp, pure Python) has no FFI of its own: its hot path calls
 * numpy and the Rust extension `qiskit._accelerate` through Python.  The drop-in boundary is
 * therefore the body of the reference's gate loop (qiskit_addon_obp/backpropagation.py:171-258);
 * each entry point below names the reference call it replaces.  The Python driver in
 * qiskit_addon_obp_b200/backpropagation.py binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (0 = OK, <0 = error),
 *     never throws; obp_last_error() gives the message.
 *   - a context owns one observable: a device-resident term table of packed symplectic keys
 *     (W = ceil(n_qubits/64) words of x and of z per term, qubit j -> bit j%64 of word j/64, the
 *     little-endian column order of SparsePauliOp.paulis.x[:, j]) and complex128 coefficients.
 *   - host buffers are caller-owned and only read/written during the call.
 *   - every call synchronises its stream before returning unless stated otherwise, so a Python
 *     signal handler (max_seconds, backpropagation.py:138-142) runs between calls.
 */
#ifndef OBP_B200_H
#define OBP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OBP_OK 0
#define OBP_ERR_BADARG (-1)
#define OBP_ERR_CUDA (-2)
#define OBP_ERR_OOM (-3)
#define OBP_ERR_CAPACITY (-4) /* term table full: state is undefined, call obp_restore() */
#define OBP_ERR_UNSUPPORTED (-5)

#define OBP_MAX_WORDS 32 /* up to 2048 qubits */
/* atol values with a special meaning: merge duplicates but trim nothing / keep the rows exactly as
 * given or produced (OperatorBudget(simplify=False), backpropagation.py:220: no merge, zeros stay) */
#define OBP_ATOL_MERGE_ONLY (-1.0)
#define OBP_ATOL_RAW (-2.0)

typedef struct obp_ctx obp_ctx;

/* ---- gate programs -------------------------------------------------------------------------
 * One obp_op is one gate of a slice after the host has decomposed it the way the reference does
 * (SparsePauliOp.from_operator(Operator(gate)), backpropagation.py:213) and classified it.
 */
enum {
  OBP_OP_CLIFFORD = 1, /* Pauli-transfer matrix is a signed permutation (incl. Pauli gates)   */
  OBP_OP_ROTATION = 2  /* U = u0*I + u1*G with G a Pauli on <= 4 qubits (rx, rzz, t, p, ...)  */
};

#define OBP_FLAG_EXACT_COUNTERS 1u /* produce the reference's simplify counters for this op */
#define OBP_FLAG_OWN_SLOT 2u       /* do not fuse this Clifford-class gate with its neighbours: the term count
                                      after exactly this gate is then reported by obp_last_op_sizes()       */

typedef struct {
  int32_t kind;        /* OBP_OP_*                                                            */
  int32_t n_qubits;    /* qubits the gate acts on: 1..2 (Clifford), 1..4 (rotation)           */
  int32_t qubit[4];    /* global qubit indices                                                */
  uint32_t flags;      /* OBP_FLAG_*                                                          */
  uint32_t n_components; /* k = Pauli terms of the gate's decomposition (raw rows = n*k*k)    */
  /* Clifford: local Pauli code v = x0 | z0<<1 | x1<<2 | z1<<3 of a term on the gate's qubits   */
  uint64_t lut;        /* image code of v in bits [4v, 4v+4)                                  */
  uint32_t sign_mask;  /* bit v set: U^dag sigma_v U = -sigma_lut[v]                          */
  uint32_t group_size; /* |D|, D = {U_j xor U_l}: keys of the k*k raw rows of one term        */
  uint64_t group_elems;/* the group_size-1 non-identity members of D, 4-bit codes             */
  double row_scale;    /* |u_j * conj(u_l)| (uniform for a Clifford): raw-row magnitude / |c| */
  /* rotation */
  uint32_t gen_local;  /* generator code: bit 2j = x, bit 2j+1 = z on qubit[j]                */
  uint32_t reserved;
  double u0[2];        /* coefficient of I  (re, im), e.g. cos(theta/2)                       */
  double u1[2];        /* coefficient of G  (re, im), e.g. -i sin(theta/2)                    */
} obp_op;

/* Counters of the last op of a program, in the vocabulary of SimplifyMetadata
 * (qiskit_addon_obp/utils/simplify.py:26-40, 128-152). Valid when the op carried
 * OBP_FLAG_EXACT_COUNTERS; n_in / n_out / n_dense / updates are always valid. */
typedef struct {
  uint64_t n_in;          /* terms at entry of the last op                                    */
  uint64_t raw_rows;      /* n_in * k * k  (SliceMetadata.raw_num_paulis)                     */
  uint64_t num_unique;    /* distinct keys among the raw rows that passed the zero filter     */
  uint64_t num_duplicate; /* surviving raw rows - num_unique                                  */
  uint64_t num_trimmed;   /* filtered raw rows + merged keys with |sum| <= atol               */
  double sum_trimmed[2];  /* sum of the trimmed merged coefficients                           */
  uint64_t n_out;         /* live terms after the program                                     */
  uint64_t n_dense;       /* rows in the dense table incl. tombstones                         */
  uint64_t updates;       /* sum over ops of the live term count at op entry                  */
} obp_stats;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Create a context on CUDA device `device` for `n_qubits`-qubit observables with room for
 * `capacity` terms. */
int obp_create(int device, int n_qubits, uint64_t capacity, obp_ctx** out);
void obp_destroy(obp_ctx* ctx);
/* Message of the last failing call on ctx (ctx == NULL: of the last failing obp_create). */
const char* obp_last_error(const obp_ctx* ctx);
/* Run all work of ctx on an existing CUDA stream (cudaStream_t); NULL = the context's own. */
int obp_set_stream(obp_ctx* ctx, void* cuda_stream);
/* Free and total bytes of device memory (cudaMemGetInfo) — used to size the table. */
int obp_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes);
/* Bytes of device memory a context of this shape allocates per term of capacity. */
uint64_t obp_bytes_per_term(int n_qubits);

/* ---- table I/O ----------------------------------------------------------------------------- */

/* Upload n rows (x_words/z_words: [n][W] row-major, coeffs: [n][2]) and merge duplicate keys.
 * atol >= 0: full simplify semantics (utils/simplify.py:115-165: zero filter, merge, zero filter)
 * with counters in *stats; atol < 0: merge only, nothing trimmed.  Replaces the observable that
 * reduce_op() hands to the loop (backpropagation.py:114) and simplify() itself. */
int obp_load(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
             uint64_t n, double atol, obp_stats* stats);
/* Same as obp_load, but a snapshot taken before stays valid: the table's content is REPLACED in the middle of
 * a slice (a sharded table runs a generic gate on the gathered operator and hands every rank its share back),
 * obp_restore still returns to the last commit. */
int obp_replace(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
                uint64_t n, double atol, obp_stats* stats);
/* Number of live terms (len(observable)). */
int obp_count(obp_ctx* ctx, uint64_t* n);
/* Compact and copy the live terms to host arrays with room for max_rows rows. */
int obp_download(obp_ctx* ctx, uint64_t* x_words, uint64_t* z_words, double* coeffs,
                 uint64_t max_rows, uint64_t* n);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Conjugate the observable by n_ops gates in order, merging duplicates and trimming |c| <= atol
 * after every gate.  Replaces apply_op_to(...) + simplify(...) per gate
 * (backpropagation.py:210-237; utils/operations.py:100-107; utils/simplify.py:80-169). */
int obp_apply_ops(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats);

/* The same call in two halves, so that the caller can prepare the next slice (light cone, gate records)
 * while the device runs this one: obp_apply_ops_submit validates, copies the records and enqueues the
 * kernels without waiting; obp_apply_ops_wait waits for them and returns what obp_apply_ops would have.
 * Between the two no other call may touch the context (they return OBP_ERR_BADARG).  Single-GPU tables. */
int obp_apply_ops_submit(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol);
int obp_apply_ops_wait(obp_ctx* ctx, obp_stats* stats);
/* Live term count after every op of the last finished program (the size the reference logs at DEBUG level
 * after each gate, backpropagation.py:239-242).  Gates fused into one pass share that pass's count unless they
 * carried OBP_FLAG_OWN_SLOT.  Writes min(n_ops, max_ops) values, returns the number of ops in *n_ops. */
int obp_last_op_sizes(obp_ctx* ctx, uint64_t* sizes, int32_t max_ops, int32_t* n_ops);

/* The reference's simplify counters (and n_in / raw_rows / n_out) of op `op_index` of the gate program just finished,
 * for every op that carried OBP_FLAG_EXACT_COUNTERS — not only the last one: several slices whose gates are all
 * Clifford-class can then travel as ONE program, each slice's last gate flagged (the driver does that when the term
 * count cannot grow and nothing can be rejected in between).  Not available after a run in shared memory. */
int obp_last_op_stats(obp_ctx* ctx, int32_t op_index, obp_stats* out);

/* Any other gate on 1-4 qubits (cp, crz, arbitrary unitaries): U = sum_j u_j * Pauli(codes[j]) with
 * codes[j] = local Pauli code (bit 2q = x, bit 2q+1 = z on qubits[q]) and coeffs[j] = (re, im).  Runs
 * the reference's literal k*k raw-row expansion (utils/operations.py:103-105) followed by simplify
 * (utils/simplify.py:115-165); simplify = 0 keeps the raw rows unmerged (OperatorBudget(simplify=False)).
 * Needs room for n*k*k extra rows (OBP_ERR_CAPACITY otherwise). */
int obp_apply_generic(obp_ctx* ctx, int32_t n_qubits, const int32_t* qubits, int32_t n_comp, const uint32_t* codes,
                      const double* coeffs, double atol, int32_t simplify, obp_stats* stats);

/* <0|O|0> on one qubit then simplify: apply_reset_to (utils/operations.py:194-225).
 * atol == OBP_ATOL_RAW: without the simplify (X/Y rows keep their place with coefficient 0). */
int obp_apply_reset(obp_ctx* ctx, int32_t qubit, double atol, obp_stats* stats);

/* Pauli-Lindblad damping then simplify: apply_ple_to (utils/operations.py:228-275).
 * gen_x/gen_z: [n_gens][W] global-width generators, factors[g] = exp(-2*rate_g). */
int obp_apply_damping(obp_ctx* ctx, const uint64_t* gen_x, const uint64_t* gen_z,
                      const double* factors, int32_t n_gens, double atol, obp_stats* stats);

/* Error-budget truncation: truncate_binary_search (utils/truncating.py:144-204).
 * Returns the incurred error and the number of removed terms. */
int obp_truncate(obp_ctx* ctx, double budget, int32_t p_norm, double tol, double* slice_error,
                 uint64_t* n_removed, uint64_t* n_out);

/* Distinct keys over several observables: _observable_oversized (backpropagation.py:419-428). */
int obp_union_count(obp_ctx* const* ctxs, int32_t n_ctx, uint64_t* n_distinct);

/* Number of qubit-wise commuting groups of n Pauli terms given as packed words [n][W] (the union of all
 * observables' distinct keys, in the caller's row order): len(all_paulis.group_commuting(qubit_wise=True))
 * of backpropagation.py:368-375, 433-440 — greedy colouring of the qubit-wise non-commutation graph,
 * vertices by descending degree, ties in row order.  Needs no context; n <= 2^20. */
int obp_qwc_group_count(int device, int n_qubits, const uint64_t* x_words, const uint64_t* z_words, uint64_t n,
                        uint64_t* n_groups);

/* Commit / roll back: the deepcopy at backpropagation.py:279-280 and the discarded slice of
 * :266-275.  obp_snapshot copies the live terms aside; obp_restore brings them back. */
int obp_snapshot(obp_ctx* ctx);
int obp_restore(obp_ctx* ctx);

/* Grow the table to hold new_capacity terms (contents preserved).  Failure-atomic: when device
 * memory runs out the table keeps its old capacity and contents and OBP_ERR_OOM is returned (the
 * caller may go on without growing).  Only if the old state cannot be re-established either is the
 * context marked unusable: obp_is_broken() then returns 1, every other call fails, destroy it. */
int obp_reserve(obp_ctx* ctx, uint64_t new_capacity);
int obp_is_broken(const obp_ctx* ctx);

/* Run-time switches of one context.  They select between code paths with identical results
 * (utils/operations.py:103-105 + utils/simplify.py:115-165 either way); tests flip them to check
 * every path against the oracle.
 *   OBP_OPT_PROGRAM           1 (default): tables of <= 128 qubits run a slice's gate program in one
 *                             persistent cooperative launch; 0: one kernel per gate / Clifford batch.
 *   OBP_OPT_PROGRAM_MAX_ROWS  tables with at least this many dense rows use the stand-alone kernels
 *                             (default 2^18).
 *   OBP_OPT_SEGMENTS          1 (default): a call that consists of rotations only (an RX / RZ / RZZ ... layer of a
 *                             Trotter step; tables of <= 128 qubits) is applied in shared memory: the table is
 *                             partitioned so that every key a term can mix with shares its partition, one CTA
 *                             applies all gates to a partition and writes the survivors out once (obp_segment.cuh);
 *                             0: one pass over the table per gate.  Results are identical either way.
 *   OBP_OPT_SEGMENT_MIN_GATES shorter runs use the per-gate kernels (default 6).
 *   OBP_OPT_SEGMENT_CAP_ROWS  rows a partition may hold, 0 = what the shared memory allows (test hook: small values
 *                             force the cut into shorter pieces and the fall-back to the per-gate kernels).
 *   OBP_OPT_SEGMENT_PIECES_MIN_ROWS  a failed run on at least this many rows is repeated in shorter pieces, a failed
 *                             run on fewer rows goes to the per-gate kernels (default 2^17).
 *   OBP_OPT_SEGMENT_MIXED     1: a run may also hold Clifford-class gates between its rotations (applied to the rows
 *                             of each partition; the partition map vanishes on the generators pulled back through
 *                             them).  0 (default): rotation-only gate programs.  Results are identical either way. */
#define OBP_OPT_PROGRAM 1
#define OBP_OPT_PROGRAM_MAX_ROWS 2
#define OBP_OPT_SEGMENTS 3
#define OBP_OPT_SEGMENT_MIN_GATES 4
#define OBP_OPT_SEGMENT_CAP_ROWS 5
#define OBP_OPT_SEGMENT_PIECES_MIN_ROWS 6
#define OBP_OPT_SEGMENT_MIXED 7
int obp_configure(obp_ctx* ctx, int32_t option, int64_t value);

/* Runs of rotations applied in shared memory since the context was created: committed runs, failed runs repeated in
 * shorter pieces (too many rows shared a partition), runs handed to the per-gate kernels. */
int obp_segment_stats(obp_ctx* ctx, uint64_t* n_runs, uint64_t* n_retries, uint64_t* n_fallbacks);

/* The partition map of a run (host only, no device needed; exposed for tests): n_gen rotation generators as packed
 * words (gen_x / gen_z: n_gen x n_words, n_words = 1 or 2).  fmask (16 x 4 words) receives a GF(2)-linear map
 * keys -> 16 bits, bit j = parity of (x0 & fmask[4j]) ^ (z0 & fmask[4j+1]) ^ (x1 & fmask[4j+2]) ^ (z1 & fmask[4j+3]),
 * that vanishes on every generator: a key and the key xor any product of generators land in the same partition. */
int obp_segment_partition_map(int32_t n_words, int32_t n_gen, const uint64_t* gen_x, const uint64_t* gen_z, uint64_t* fmask);

/* Row-order mode (off by default).  The reference's simplify() returns rows in the order of each
 * key's first appearance among the gate's k*k raw rows (utils/simplify.py:126,160-165 on top of the
 * compose order of utils/operations.py:103-105); the in-place gate kernels of obp_apply_ops do not
 * keep that order.  With the mode on, obp_load / obp_apply_generic / obp_apply_reset /
 * obp_apply_damping / obp_truncate leave the table in exactly the reference's row order (every gate
 * must then go through obp_apply_generic; obp_apply_ops returns OBP_ERR_UNSUPPORTED).  Costs 12
 * bytes of scratch per term of capacity and the k*k expansion per gate.  Single-GPU tables only. */
int obp_set_row_order(obp_ctx* ctx, int32_t enabled);

/* ---- sharded table (multi-GPU) ---------------------------------------------------------------
 * One process per GPU; rank r of R (a power of two) holds the terms whose Clifford-invariant hash
 * has its top log2(R) bits equal to r, so duplicate keys co-locate and Clifford gates, damping and
 * truncation kills stay rank-local.  The reference has no counterpart (single process): these entry
 * points split the per-gate body of backpropagation.py:210-237 around the one exchange step a
 * branching gate needs.  All ranks must issue the same calls in the same order. */
typedef struct {
  uint64_t n_in, n_out, n_dense;
  uint64_t rows_surv;     /* raw rows that passed the zero filter (counted where they are produced)   */
  uint64_t keys_present;  /* distinct keys among them (counted where the key lives)                   */
  uint64_t cancelled;     /* merged keys with |sum| <= atol                                           */
  uint64_t n_records;     /* records emitted for the peer                                             */
  double sum_trimmed[2];
} obp_shard_counters;

int obp_set_shard(obp_ctx* ctx, int32_t n_ranks, int32_t rank);
/* Late sharding: while the operator is small every rank carries an identical copy of the whole table
 * (no gate needs an exchange); obp_shard_split turns that copy into rank `rank`'s shard by dropping the
 * rows other ranks own.  future_ops (optional, n_future records): the gates the run applies next, in
 * order.  The hash frame is then re-drawn such that the owner bits of the generator of every rotation
 * among them are zero — a rotation P -> P xor G keeps owner(P xor G) = owner(P) — so those rotations run
 * rank-local without any record exchange; *n_covered = how many rotations the plan covers.  Rotations
 * outside the plan still work through the exchange path.  Invalidates the snapshot (it held the whole
 * operator): call obp_snapshot again. */
int obp_shard_split(obp_ctx* ctx, int32_t n_ranks, int32_t rank, const obp_op* future_ops, int32_t n_future,
                    uint64_t* n_out, int32_t* n_covered);
/* Statistics since the context was created: rotations that needed a record exchange with a peer, and records
 * this rank emitted (the latter as of the last call that synchronised with the device). */
int obp_shard_stats(obp_ctx* ctx, uint64_t* n_exchange_rotations, uint64_t* n_records_emitted);
/* Bytes of one exchange record: W x/z word pairs, the hash, the complex128 coefficient. */
uint64_t obp_record_bytes(int n_qubits);
/* obp_apply_ops on a sharded table: stops before the first rotation whose partner keys live on
 * another rank.  *n_done = ops executed; *peer_xor = 0 if the program is finished, else the peer of
 * rank r for ops[*n_done] is r ^ *peer_xor. */
int obp_apply_ops_sharded(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats,
                          int32_t* n_done, int32_t* peer_xor);
/* Split rotation, part 1: update the rows' own keys and write the partner-key contributions as
 * records into the DEVICE buffer d_records (room for cap_records; out->n_records are produced). */
int obp_rotate_emit(obp_ctx* ctx, const obp_op* op, double atol, void* d_records, uint64_t cap_records,
                    obp_shard_counters* out);
/* Split rotation, part 2: merge the n_records the peer sent (DEVICE buffer) and run the second
 * zero filter.  Counters of emit and merge add up across ranks to the reference's simplify counters:
 * num_unique = sum keys_present, num_duplicate = sum rows_surv - num_unique,
 * num_trimmed = 4 * sum n_in(emit) - sum rows_surv + sum cancelled. */
int obp_rotate_merge(obp_ctx* ctx, const void* d_records, uint64_t n_records, obp_shard_counters* out);
/* apply_reset_to on a sharded table: rows with Z on the qubit move to the key with that Z cleared,
 * which belongs to rank ^ *peer_xor.  *peer_xor == 0: nothing was done, call obp_apply_reset (all
 * local).  Otherwise phase 1 ran (zero filter, moving rows written as records into d_records); exchange
 * the records with the peer and finish with obp_reset_merge.  Counters add up across ranks like those
 * of the split rotation (raw rows = sum n_in). */
int obp_reset_emit(obp_ctx* ctx, int32_t qubit, double atol, void* d_records, uint64_t cap_records, int32_t* peer_xor,
                   obp_shard_counters* out);
int obp_reset_merge(obp_ctx* ctx, double atol, const void* d_records, uint64_t n_records, obp_shard_counters* out);

/* The same split rotation with the transfer fused into the emitting kernel: every rank owns two
 * inboxes (ping-pong) in device memory that its peers map (CUDA IPC across processes, the raw
 * pointer inside one process); obp_rotate_emit_p2p stores the records with plain st.global into the
 * peer's inbox over NVLink and bumps the peer's cursor with a remote atomic, so no staging buffer,
 * no count exchange and no separate copy are needed.  The caller places one cross-rank barrier
 * between emit and merge (all emits done and visible), nothing else. */
int obp_inbox_create(obp_ctx* ctx, uint64_t cap_records);
/* stage 0: unmap the peers' inboxes; stage 1: free this rank's own (every rank finishes stage 0 first). */
int obp_inbox_release(obp_ctx* ctx, int32_t stage);
/* handle64: 64-byte CUDA IPC handle of inbox `which` (0/1) (may be NULL); raw_ptr: its device address. */
int obp_inbox_export(obp_ctx* ctx, int32_t which, void* handle64, uint64_t* raw_ptr);
/* Make rank peer_rank's inbox `which` writable from this context: by IPC handle (other process) or,
 * with handle64 == NULL, by the raw device pointer (same process). */
int obp_inbox_attach(obp_ctx* ctx, int32_t peer_rank, int32_t which, const void* handle64, uint64_t raw_ptr);
int obp_rotate_emit_p2p(obp_ctx* ctx, const obp_op* op, double atol, int32_t peer_rank, obp_shard_counters* out);
/* device_flags != 0 (one process per GPU only): obp_apply_ops_sharded no longer stops at a rotation
 * that needs a peer; it enqueues emit -> merge -> finish itself and the kernels of the two ranks are
 * ordered by flag words in the inbox headers (the merge kernel waits on the device until the peer's
 * emit kernel has raised `emitted`, the emit kernel waits until the peer has raised `merged` for the
 * inbox's previous use).  No host synchronisation or NCCL call per exchange. */
int obp_set_shard_sync(obp_ctx* ctx, int32_t device_flags);
/* Counters of the split rotation that ended the last obp_apply_ops_sharded program in device-flag
 * mode (out->n_in == 0: the program did not end on one). */
int obp_last_split_counters(obp_ctx* ctx, obp_shard_counters* out);
int obp_rotate_merge_inbox(obp_ctx* ctx, obp_shard_counters* out);

/* Stepwise truncation: truncate_binary_search (utils/truncating.py:171-204) with the maximum
 * (:173) and each masked sum (:187) exposed so that ranks can combine them with an allreduce. */
int obp_trunc_begin(obp_ctx* ctx, int32_t p_norm, double* local_max);
int obp_trunc_masked_sum(obp_ctx* ctx, double threshold, double* local_sum);
int obp_trunc_finish(obp_ctx* ctx, double threshold, uint64_t* n_removed, uint64_t* n_out);

/* ---- measurement --------------------------------------------------------------------------- */

enum { OBP_K_ROTATE = 0, OBP_K_CLIFFORD = 1, OBP_K_TRUNCATE = 2, OBP_K_COMPACT = 3,
       OBP_K_INDEX = 4, OBP_K_OTHER = 5, OBP_K_PROGRAM = 6, OBP_K_SEGMENT = 7, OBP_K_COUNT = 8 };

/* OBP_K_PROGRAM is the persistent cooperative kernel that runs a whole gate program in one launch
 * (tables of up to 128 qubits): it counts as ONE launch; the table passes it makes are counted per
 * class in `passes`, and their device time (from %globaltimer stamps taken inside the kernel at
 * each grid barrier) is attributed to the class of each pass in `ms`.
 * OBP_K_SEGMENT is a run of rotations applied in shared memory (five launches per run: partition count /
 * scan / scatter, the run itself, bookkeeping); `passes` counts its gates, `term_updates` the live terms at
 * the entry of every gate, `ms` the five kernels together (CUDA events). */
typedef struct {
  uint64_t launches[OBP_K_COUNT];  /* kernel launches per class since the last reset           */
  uint64_t passes[OBP_K_COUNT];    /* passes over the table per class (= launches for stand-alone kernels) */
  double ms[OBP_K_COUNT];          /* device time per class (only while profiling is on)       */
  uint64_t term_updates[OBP_K_COUNT]; /* live terms at entry summed over passes (x gates)      */
} obp_profile;

int obp_profile_enable(obp_ctx* ctx, int32_t on);
int obp_profile_get(obp_ctx* ctx, obp_profile* out, int32_t reset);

/* Device-side stopwatch on the context's stream (CUDA events): start records an event, stop records
 * a second one, synchronises and returns the elapsed milliseconds between them. */
int obp_timer_start(obp_ctx* ctx);
int obp_timer_stop(obp_ctx* ctx, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* OBP_B200_H */
