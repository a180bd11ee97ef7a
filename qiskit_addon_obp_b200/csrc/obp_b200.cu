// obp_b200.cu — host side of the C-ABI in include/obp_b200.h: context, gate programs, truncation,
// compaction, snapshot.  Kernels live in obp_kernels.cuh.  Built for sm_100a only.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "obp_kernels.cuh"
#include "obp_qwc.cuh"
#include "obp_segment.cuh"

namespace {

std::string g_create_error;

struct EventPair {
  cudaEvent_t a, b;
  int klass;
};

struct Snapshot {
  ulonglong2* xz[OBP_MAX_WORDS] = {nullptr};
  double2* coef = nullptr;
  u64* hash = nullptr;
  u64 cap = 0;
  u64 n = 0;       // dense rows held
  u64 n_live = 0;  // live rows among them
  bool valid = false;
  std::vector<u64> hx, hz;
};

}  // namespace

struct obp_ctx {
  int device = 0, nq = 0, W = 0, n_sm = 148;
  u64 cap = 0, T = 0;
  // The index is sized to the CAPACITY (T = next_pow2(2 cap) slots), not to the rows.  Measured on B200
  // (config 2, theta 0.1, tables of 10^3..10^4 rows): an index sized to the rows (load factor 1/8..1/2)
  // is slower (12.2 -> 15.8 ms of device time per run) — an absent partner costs one probe in a
  // near-empty index and ~2.5 dependent probes at load 1/2, and that chain, not the L2 residency of the
  // index, is what a small-table pass waits for.
  cudaStream_t stream = nullptr, own_stream = nullptr;
  ulonglong2* xz[OBP_MAX_WORDS] = {nullptr};
  double2* coef = nullptr;
  u64* hash = nullptr;
  u64* index = nullptr;
  ulonglong2* spare16 = nullptr;
  u64* spare8 = nullptr;
  unsigned* dest = nullptr;
  unsigned* blk = nullptr;
  // d_st, d_tbegin and d_ops live in ONE device block (and h_st, h_tbegin, h_ops in one pinned block):
  // [DevState | pad to 64 | t_begin | pad to 128 | OpStat x ops_alloc], so that a gate program brings
  // its state, time stamps and counters back with a single copy
  struct PendingOps* pending = nullptr;  // a submitted gate program nobody has waited for yet
  unsigned char* d_res = nullptr;
  unsigned char* h_res = nullptr;  // pinned and mapped: the persistent program kernel writes its results into it
  u64* h_res_dev = nullptr;        // device address of h_res
  DevState* d_st = nullptr;
  OpStat* d_ops = nullptr;
  int ops_alloc = 0;
  double* d_partial = nullptr;
  double* d_scalar = nullptr;
  u64* d_cols = nullptr;
  u64* d_gens = nullptr;
  double* d_factors = nullptr;
  int gens_alloc = 0;
  DevState* h_st = nullptr;    // pinned
  OpStat* h_ops = nullptr;     // pinned, ops_alloc entries
  double* h_scalar = nullptr;  // pinned
  // GF(2)-linear hash frame: h(key) = XOR of hx[q] over x bits and hz[q] over z bits.  Clifford
  // gates re-express the frame (columns of the touched qubits) so stored row hashes stay valid.
  std::vector<u64> hx, hz;
  bool cols_dirty = true;
  u64 epoch = 1;
  int n_ranks = 1, rank = 0, owner_shift = 64;  // sharding: owner = hash >> owner_shift
  ShardCnt* d_cnt = nullptr;
  ShardCnt* h_cnt = nullptr;  // pinned
  RotP shard_rot;             // rotation being split across obp_rotate_emit / obp_rotate_merge
  // P2P exchange: two inboxes (ping-pong) other ranks' emit kernels write into; [0] = cursor, records from +16
  u64* inbox[2] = {nullptr, nullptr};
  u64 inbox_cap = 0;
  std::vector<u64*> peer_inbox[2];     // per rank: its inbox as mapped into this process (nullptr = not attached)
  std::vector<bool> peer_is_ipc[2];
  int xchg_parity = 0;
  bool last_split_is_last_op = false;  // the program just run ended on a handshake split rotation (counters in h_cnt)
  bool flag_sync = false;  // split rotations are ordered by device-side flags, no host barrier
  u64 xseq = 0;            // split rotations run so far (identical on all ranks)
  u64 n_split_rot = 0;     // rotations that needed a record exchange (statistics)
  bool shard_rot_open = false;
  int rot_ctas_per_sm = 0;  // resident CTAs of k_rotate per SM (occupancy API)
  int prog_ctas_per_sm = 0; // resident CTAs of k_program per SM; 0 = cooperative launch unavailable
  unsigned char* d_prog = nullptr;  // persistent program: entries + parameter pool (device)
  unsigned char* h_prog = nullptr;  // pinned staging of the same
  size_t prog_alloc = 0;
  TruncState* d_ts = nullptr;
  TruncState* h_ts = nullptr;  // pinned
  int trunc_ctas_per_sm = 0;  // resident CTAs of k_trunc_search per SM; 0 = cooperative launch unavailable
  bool row_order = false;          // rows are kept in the reference's first-appearance order
  unsigned* ord_minidx = nullptr;  // row-order mode scratch: first raw position per key, flags, positions (cap u32 each)
  unsigned* ord_flag = nullptr;
  unsigned* ord_pos = nullptr;
  u64 ord_cap = 0;
  void* h_bounce = nullptr;  // pinned bounce buffer of downloads (OBP_BOUNCE_BYTES)
  u64* d_stage = nullptr;  // download staging for tables too wide for the compaction scratch
  size_t stage_bytes = 0;
  u64* d_tbegin = nullptr;
  u64* h_tbegin = nullptr;  // pinned
  std::vector<u64> last_op_sizes;  // live rows after every op of the last finished program
  std::vector<obp_op> last_ops;    // ... its ops, the counter slot of each and the row count it started from (obp_last_op_stats)
  std::vector<int> last_slot_of_op;
  int last_n_slots = 0;
  u64 last_n_start = 0;
  bool use_program = true;
  u64 prog_max_rows = 1ull << 18;  // tables of at least this many rows run their gates as stand-alone kernels
  bool broken = false;  // a failed obp_reserve could not restore the table: every entry point fails fast
  // ---- runs of rotations applied in shared memory (obp_segment.cuh) ----
  bool use_segments = true;
  bool seg_mixed = false;       // OBP_B200_SEG_MIXED=1: runs may hold Clifford-class gates between their rotations (parity-green; measured
                                // slower than the program kernel's fused Clifford batches on config 5, so off by default)
  int seg_min_gates = 6;        // shorter runs go through the per-gate kernels
  int seg_cap_rows = 0;         // test hook: rows a partition may hold (0: what the shared memory allows)
  u64 seg_pieces_min_rows = 1ull << 17;  // a failed run on at least this many rows is repeated in pieces (below: per-gate kernels)
  int seg_smem_max = 0;         // opt-in shared memory per CTA of k_seg_main (0: path unavailable)
  double seg_growth = 4.0;      // rows out / rows in of the last run (sizes the next run's partitions)
  int seg_max_gates = OBP_SEG_MAXG;  // longer runs are cut into pieces of at most this many gates (halved when a run fails:
                                     // fewer generators leave a finer partition map, i.e. smaller coset classes)
  u64 carry_updates = 0;             // term-gate updates of the pieces of a cut run that were finished before the last one
  std::vector<u64> carry_sizes;      // ... and their per-op row counts
  u64 seg_block_rows = ~0ull;   // a run on this many live rows fell back to the per-gate kernels: no attempts at or above it
  ulonglong2* alt_xz[2] = {nullptr, nullptr};  // the run's output columns (swapped with the table's on success)
  double2* alt_coef = nullptr;
  u64* alt_hash = nullptr;
  unsigned char* d_seg = nullptr;  // [cnt | off | cur] (OBP_SEG_MAXP + 1 u32 each) | SegRes
  bool index_touched = true;       // a program consulted the index since the last run committed: the next run fills it in its output stage
  u64 seg_T_new = 0;               // ... slots that run cleared (0: it did not touch the index)
  bool index_stale = false;        // the table's columns were replaced by a run: the index is rebuilt before its next use
  u64 T_eff = 0;                   // index slots in use (a power of two <= T): a lazy rebuild after a run sizes the index to the rows
  u64 n_seg_runs = 0, n_seg_retries = 0, n_seg_fallbacks = 0;
  bool trace_host = false;  // OBP_B200_TRACE_HOST: print host enqueue / wait time per gate program
  int prog_min_grid = 148;  // smallest grid of the persistent program kernel (grid barriers cost more with more CTAs)
  int prog_reach_shift = 6;  // the grid is sized for rows << min(#rotations, this)
  int prog_debug = 0;         // OBP_B200_PROG_DEBUG: timing experiments inside k_program (2: skip the table work)
  u64 n_live = 0, n_dense = 0;
  Snapshot snap;
  obp_ctx* scratch = nullptr;  // union-count scratch table
  // profiling
  bool prof_on = false;
  obp_profile prof;
  std::vector<EventPair> ev_pending, ev_free;
  EventPair run_ep{};
  bool run_open = false;
  cudaEvent_t tm_a = nullptr, tm_b = nullptr;
  std::string err;

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
};

#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ctx->fail(e__ == cudaErrorMemoryAllocation ? OBP_ERR_OOM : OBP_ERR_CUDA,          \
                       std::string(#call) + ": " + cudaGetErrorString(e__));                   \
  } while (0)

namespace {

inline u64 splitmix64(u64& s) {
  u64 z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

inline u64 next_pow2(u64 v) {
  u64 p = 1;
  while (p < v) p <<= 1;
  return p;
}

Tab make_tab(const obp_ctx* c) {
  Tab t;
  for (int w = 0; w < OBP_MAX_WORDS; ++w) t.xz[w] = c->xz[w];
  t.coef = c->coef;
  t.hash = c->hash;
  t.index = c->index;
  t.mask = (c->T_eff ? c->T_eff : c->T) - 1;
  t.cap = c->cap;
  t.W = c->W;
  t.st = c->d_st;
  t.ops = c->d_ops;
  return t;
}

inline int grid_for(const obp_ctx* c, u64 rows, int per_sm = 8) {
  u64 g = (rows + OBP_BS - 1) / OBP_BS;
  if (g < 1) g = 1;
  const u64 mx = (u64)c->n_sm * per_sm;
  return (int)std::min(g, mx);
}

// RAII launch bracket: counts the launch and, while profiling, times it with CUDA events.  Back to
// back launches of the same class share one event pair (the run is closed when the class changes
// or the stream is synchronised), so profiling adds two event records per run, not per launch.
void close_run(obp_ctx* c) {
  if (c->run_open) {
    cudaEventRecord(c->run_ep.b, c->stream);
    c->ev_pending.push_back(c->run_ep);
    c->run_open = false;
  }
}

struct Bracket {
  Bracket(obp_ctx* c, int k, int n_kernels = 1) {
    c->prof.launches[k] += (u64)n_kernels;
    c->prof.passes[k]++;
    if (!c->prof_on) return;
    if (c->run_open && c->run_ep.klass == k) return;
    close_run(c);
    EventPair ep{};
    if (!c->ev_free.empty()) {
      ep = c->ev_free.back();
      c->ev_free.pop_back();
    } else {
      cudaEventCreate(&ep.a);
      cudaEventCreate(&ep.b);
    }
    ep.klass = k;
    cudaEventRecord(ep.a, c->stream);
    c->run_ep = ep;
    c->run_open = true;
  }
};

void resolve_events(obp_ctx* c) {
  if (c->run_open) {  // callers close the run before they synchronise; this is the safety net
    close_run(c);
    cudaEventSynchronize(c->ev_pending.back().b);
  }
  for (auto& ep : c->ev_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) c->prof.ms[ep.klass] += ms;
    c->ev_free.push_back(ep);
  }
  c->ev_pending.clear();
}

constexpr size_t RES_HDR = OBP_RES_HDR;
static_assert(sizeof(DevState) <= 64, "result block header layout");

// Bring the table state back and wait for the stream.  n_ops > 0: the first n_ops counter slots (and the
// program's start stamp) come along in the same copy.
// written_by_kernel: the persistent program kernel already put all of that into h_res (mapped).
int sync_state(obp_ctx* ctx, int n_ops = 0, bool written_by_kernel = false) {
  close_run(ctx);
  const size_t bytes = n_ops > 0 ? RES_HDR + sizeof(OpStat) * (size_t)n_ops : sizeof(DevState);
  if (!written_by_kernel) CU(cudaMemcpyAsync(ctx->h_res, ctx->d_res, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  resolve_events(ctx);
  ctx->n_live = ctx->h_st->n_live;
  ctx->n_dense = ctx->h_st->n_scan;
  if (ctx->h_st->error) {
    const unsigned e = ctx->h_st->error;
    ctx->h_st->error = 0;
    cudaMemsetAsync(&ctx->d_st->error, 0, sizeof(unsigned), ctx->stream);
    if (e & OBP_ERR_BIT_PEER) return ctx->fail(OBP_ERR_CUDA, "sharded table: a peer rank's flag did not arrive within 10 s");
    return ctx->fail(OBP_ERR_CAPACITY, (e & OBP_ERR_BIT_CAP) ? "term table capacity exceeded"
                                                              : "hash index full");
  }
  return OBP_OK;
}

int ensure_ops(obp_ctx* ctx, int n) {
  if (n <= ctx->ops_alloc) return OBP_OK;
  const int na = std::max(n, std::max(256, ctx->ops_alloc * 2));
  const size_t bytes = RES_HDR + sizeof(OpStat) * (size_t)na;
  unsigned char *d = nullptr, *h = nullptr;
  CU(cudaMalloc(&d, bytes));
  CU(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
  CU(cudaMemset(d, 0, RES_HDR));
  memset(h, 0, RES_HDR);
  if (ctx->d_res) {  // the table state moves with the block
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(d, ctx->d_res, RES_HDR, cudaMemcpyDeviceToDevice));
    memcpy(h, ctx->h_res, RES_HDR);
    cudaFree(ctx->d_res);
    cudaFreeHost(ctx->h_res);
  }
  ctx->d_res = d;
  ctx->h_res = h;
  ctx->d_st = (DevState*)d;
  ctx->d_tbegin = (u64*)(d + 64);
  ctx->d_ops = (OpStat*)(d + RES_HDR);
  ctx->h_st = (DevState*)h;
  ctx->h_tbegin = (u64*)(h + 64);
  ctx->h_ops = (OpStat*)(h + RES_HDR);
  void* hd = nullptr;
  CU(cudaHostGetDevicePointer(&hd, h, 0));
  ctx->h_res_dev = (u64*)hd;
  ctx->ops_alloc = na;
  return OBP_OK;
}

int upload_cols(obp_ctx* ctx) {
  if (!ctx->cols_dirty) return OBP_OK;
  std::vector<u64> cols((size_t)2 * 64 * ctx->W, 0);
  for (int q = 0; q < ctx->nq; ++q) {
    cols[q] = ctx->hx[q];
    cols[(size_t)ctx->W * 64 + q] = ctx->hz[q];
  }
  CU(cudaMemcpyAsync(ctx->d_cols, cols.data(), cols.size() * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // cols is a stack-owned buffer
  ctx->cols_dirty = false;
  return OBP_OK;
}

void init_frame(obp_ctx* ctx) {
  u64 seed = 0x0B9B200C0FFEEull;
  ctx->hx.resize(ctx->nq);
  ctx->hz.resize(ctx->nq);
  for (int q = 0; q < ctx->nq; ++q) {
    ctx->hx[q] = splitmix64(seed);
    ctx->hz[q] = splitmix64(seed);
  }
  ctx->cols_dirty = true;
}

// (re)build the index from the live rows (no duplicates expected)
int rebuild_index(obp_ctx* ctx) {
  ctx->index_stale = false;
  ctx->T_eff = ctx->T;
  CU(cudaMemsetAsync(ctx->index, 0xFF, ctx->T * sizeof(u64), ctx->stream));
  Bracket b(ctx, OBP_K_INDEX);
  k_index_build<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), 0, -1.0, ctx->epoch, -1, nullptr);
  CU(cudaGetLastError());
  return OBP_OK;
}

// a run of rotations in shared memory replaces the table's columns and leaves the index behind
// need_rows: the rows the index must be able to hold before anybody rebuilds it again (0: the whole capacity).  The
// rebuild that follows a run clears and fills only as many slots as those rows want (a 10^4-row table in a context
// sized for 4 * 10^6 terms would otherwise pay a 64 MB memset per slice).
int ensure_index(obp_ctx* ctx, u64 need_rows = 0) {
  ctx->index_touched = true;
  if (need_rows == 0 || need_rows > ctx->cap) need_rows = ctx->cap;
  const u64 want = std::min<u64>(ctx->T, next_pow2(std::max<u64>(4 * need_rows, 1ull << 16)));
  if (!ctx->index_stale && (ctx->T_eff == 0 || ctx->T_eff >= std::min<u64>(ctx->T, 2 * need_rows))) return OBP_OK;
  ctx->index_stale = false;
  ctx->T_eff = want;
  CU(cudaMemsetAsync(ctx->index, 0xFF, want * sizeof(u64), ctx->stream));
  Bracket b(ctx, OBP_K_INDEX);
  k_index_build<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), 0, -1.0, ctx->epoch, -1, nullptr);
  CU(cudaGetLastError());
  return OBP_OK;
}

// move every row i with dest[i] != NONE to row dest[i] (all columns), then rebuild the index;
// st->scan_total holds the new row count
int scatter_by_dest(obp_ctx* ctx) {
  const int g = grid_for(ctx, ctx->n_dense);
  for (int w = 0; w < ctx->W; ++w) {
    k_scatter16<<<g, OBP_BS, 0, ctx->stream>>>(ctx->xz[w], ctx->spare16, ctx->dest, ctx->d_st);
    std::swap(ctx->xz[w], ctx->spare16);
  }
  k_scatter16<<<g, OBP_BS, 0, ctx->stream>>>((const ulonglong2*)ctx->coef, ctx->spare16, ctx->dest, ctx->d_st);
  {
    ulonglong2* old = (ulonglong2*)ctx->coef;
    ctx->coef = (double2*)ctx->spare16;
    ctx->spare16 = old;
  }
  k_scatter8<<<g, OBP_BS, 0, ctx->stream>>>(ctx->hash, ctx->spare8, ctx->dest, ctx->d_st);
  std::swap(ctx->hash, ctx->spare8);
  k_set_counts<<<1, 1, 0, ctx->stream>>>(ctx->d_st);
  CU(cudaGetLastError());
  ctx->n_dense = ctx->n_live;  // n_live is exact on the host at every call site
  return rebuild_index(ctx);
}

// stable compaction of all columns + index rebuild; leaves n_dense == n_live
int compact(obp_ctx* ctx) {
  if (ctx->n_dense == 0) return OBP_OK;
  const u64 nb = (ctx->n_dense + OBP_CHUNK - 1) / OBP_CHUNK;
  Tab t = make_tab(ctx);
  Bracket b(ctx, OBP_K_COMPACT, 5 + ctx->W + 2);
  k_live_count<<<(unsigned)nb, OBP_BS, 0, ctx->stream>>>(t, ctx->blk);
  k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->blk, nb, ctx->d_st);
  k_dest<<<(unsigned)nb, OBP_BS, 0, ctx->stream>>>(t, ctx->blk, ctx->dest);
  return scatter_by_dest(ctx);
}

// Row-order mode: like compact(), but the survivors land in the order of their keys' first
// appearance among the raw rows (ord_minidx was filled by k_index_build).
int compact_in_first_appearance_order(obp_ctx* ctx) {
  if (ctx->n_dense == 0) return OBP_OK;
  const u64 nb = (ctx->n_dense + OBP_CHUNK - 1) / OBP_CHUNK;
  Tab t = make_tab(ctx);
  Bracket b(ctx, OBP_K_COMPACT, 9 + ctx->W + 2);
  const int g = grid_for(ctx, ctx->n_dense);
  CU(cudaMemsetAsync(ctx->ord_flag, 0, ctx->n_dense * sizeof(unsigned), ctx->stream));
  k_order_flags<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->ord_minidx, ctx->ord_flag);
  k_flag_count<<<(unsigned)nb, OBP_BS, 0, ctx->stream>>>(ctx->ord_flag, ctx->d_st, ctx->blk);
  k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(ctx->blk, nb, ctx->d_st);
  k_flag_pos<<<(unsigned)nb, OBP_BS, 0, ctx->stream>>>(ctx->ord_flag, ctx->d_st, ctx->blk, ctx->ord_pos);
  k_order_dest<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->ord_minidx, ctx->ord_pos, ctx->dest);
  return scatter_by_dest(ctx);
}

// compaction after a merge of raw rows: stable, or into first-appearance order (row-order mode)
int compact_after_merge(obp_ctx* ctx) {
  return ctx->row_order ? compact_in_first_appearance_order(ctx) : compact(ctx);
}

// row-order mode: every row starts as its own first appearance
int reset_first_appearance(obp_ctx* ctx, u64 n_rows) {
  if (!ctx->row_order || !n_rows) return OBP_OK;
  CU(cudaMemsetAsync(ctx->ord_minidx, 0xFF, n_rows * sizeof(unsigned), ctx->stream));
  return OBP_OK;
}

int set_counts(obp_ctx* ctx, u64 n) {
  DevState s;
  memset(&s, 0, sizeof(s));
  s.n_scan = s.n_append = s.n_live = n;
  s.emitted_total = ctx->h_st->emitted_total;  // (a statistic, not table state)
  *ctx->h_st = s;
  CU(cudaMemcpyAsync(ctx->d_st, ctx->h_st, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->n_live = ctx->n_dense = n;
  return OBP_OK;
}

int alloc_table(obp_ctx* ctx, u64 cap) {
  ctx->cap = cap;
  ctx->T = next_pow2(std::max<u64>(2 * cap, 1024));
  ctx->T_eff = ctx->T;
  for (int w = 0; w < ctx->W; ++w) CU(cudaMalloc(&ctx->xz[w], cap * sizeof(ulonglong2)));
  CU(cudaMalloc(&ctx->coef, cap * sizeof(double2)));
  CU(cudaMalloc(&ctx->hash, cap * sizeof(u64)));
  CU(cudaMalloc(&ctx->index, ctx->T * sizeof(u64)));
  CU(cudaMalloc(&ctx->spare16, cap * sizeof(ulonglong2)));
  CU(cudaMalloc(&ctx->spare8, cap * sizeof(u64)));
  CU(cudaMalloc(&ctx->dest, cap * sizeof(unsigned)));
  CU(cudaMalloc(&ctx->blk, ((cap + OBP_CHUNK - 1) / OBP_CHUNK + 1) * sizeof(unsigned)));
  CU(cudaMemsetAsync(ctx->index, 0xFF, ctx->T * sizeof(u64), ctx->stream));
  ctx->index_stale = false;
  ctx->T_eff = ctx->T;
  return OBP_OK;
}

void free_seg_columns(obp_ctx* ctx) {
  for (int w = 0; w < 2; ++w) { cudaFree(ctx->alt_xz[w]); ctx->alt_xz[w] = nullptr; }
  cudaFree(ctx->alt_coef);
  cudaFree(ctx->alt_hash);
  ctx->alt_coef = nullptr;
  ctx->alt_hash = nullptr;
}

void free_table(obp_ctx* ctx) {
  free_seg_columns(ctx);
  for (int w = 0; w < OBP_MAX_WORDS; ++w) {
    if (ctx->xz[w]) cudaFree(ctx->xz[w]);
    ctx->xz[w] = nullptr;
  }
  cudaFree(ctx->coef);
  cudaFree(ctx->hash);
  cudaFree(ctx->index);
  cudaFree(ctx->spare16);
  cudaFree(ctx->spare8);
  cudaFree(ctx->dest);
  cudaFree(ctx->blk);
  ctx->coef = nullptr;
  ctx->hash = nullptr;
  ctx->index = nullptr;
  ctx->spare16 = nullptr;
  ctx->spare8 = nullptr;
  ctx->dest = nullptr;
  ctx->blk = nullptr;
}

void free_snapshot(Snapshot& s) {
  for (int w = 0; w < OBP_MAX_WORDS; ++w) {
    if (s.xz[w]) cudaFree(s.xz[w]);
    s.xz[w] = nullptr;
  }
  if (s.coef) cudaFree(s.coef);
  if (s.hash) cudaFree(s.hash);
  s.coef = nullptr;
  s.hash = nullptr;
  s.cap = 0;
  s.valid = false;
}

// ---- Clifford bookkeeping -----------------------------------------------------------------------

// local columns of the hash frame for a gate on qubits q[0..nq): order x0, z0, x1, z1
void frame_update(obp_ctx* ctx, const obp_op& op) {
  const int nb = 2 * op.n_qubits;
  u64 old_col[4], new_col[4];
  for (int j = 0; j < op.n_qubits; ++j) {
    old_col[2 * j] = ctx->hx[op.qubit[j]];
    old_col[2 * j + 1] = ctx->hz[op.qubit[j]];
  }
  // new column j = XOR of old columns over the bits of the preimage of basis vector e_j
  unsigned inv[16];
  for (unsigned v = 0; v < (1u << nb); ++v) inv[(op.lut >> (4 * v)) & 15ull] = v;
  for (int j = 0; j < nb; ++j) {
    const unsigned pre = inv[1u << j];
    u64 c = 0;
    for (int i = 0; i < nb; ++i)
      if (pre & (1u << i)) c ^= old_col[i];
    new_col[j] = c;
  }
  for (int j = 0; j < op.n_qubits; ++j) {
    ctx->hx[op.qubit[j]] = new_col[2 * j];
    ctx->hz[op.qubit[j]] = new_col[2 * j + 1];
  }
  ctx->cols_dirty = true;
}

bool lut_is_linear_bijection(const obp_op& op) {
  // circuits repeat a handful of distinct gates: remember the tables that already passed
  static thread_local std::vector<std::pair<u64, int>> ok_cache;
  for (const auto& e : ok_cache)
    if (e.first == op.lut && e.second == op.n_qubits) return true;
  const int nb = 2 * op.n_qubits;
  unsigned seen = 0;
  for (unsigned v = 0; v < (1u << nb); ++v) {
    const unsigned img = (unsigned)((op.lut >> (4 * v)) & 15ull);
    if (img >= (1u << nb)) return false;
    seen |= 1u << img;
    unsigned lin = 0;
    for (int i = 0; i < nb; ++i)
      if (v & (1u << i)) lin ^= (unsigned)((op.lut >> (4 * (1u << i))) & 15ull);
    if (lin != img) return false;
  }
  const bool ok = seen == ((nb == 4) ? 0xFFFFu : 0xFu);
  if (ok && ok_cache.size() < 256) ok_cache.emplace_back(op.lut, op.n_qubits);
  return ok;
}

struct PendingBatch {
  CliffBatch b;
  int n_ops = 0;
  void reset() {
    b.ng = 0;
    b.nslots = 0;
    b.scale = 1e300;
    n_ops = 0;
  }
  int slot_of_word(int w) {
    for (int s = 0; s < b.nslots; ++s)
      if (b.word_of_slot[s] == w) return s;
    b.word_of_slot[b.nslots] = w;
    return b.nslots++;
  }
};

// ---- host-side fusion of Clifford-class gates ------------------------------------------------------
// Inside a run of Clifford-class gates a one-qubit Clifford commutes with everything that does not
// touch its qubit, so it is held back (per qubit, composing with later one-qubit Cliffords on the same
// qubit) until a two-qubit Clifford absorbs it or the run ends (any rotation ends it).  The brickwork block [1q(a), 1q(b), 2q(a,b)]
// becomes ONE table-lookup gate; the device does a third of the bit work.
struct FusedGate {
  int nq = 0;
  int q[2] = {0, 0};
  u64 lut = 0;        // image of local code v in bits [4v, 4v+4)
  unsigned sign = 0;  // bit v: sign flip
  double scale = 1e300;
  int n_gates = 0;
};

inline unsigned lut_at(u64 lut, unsigned v) { return (unsigned)((lut >> (4 * v)) & 15ull); }

// `first` then `second`, both on the same qubits in the same order
FusedGate fuse_seq(const FusedGate& first, const FusedGate& second) {
  FusedGate r = first;
  r.lut = 0;
  r.sign = 0;
  const unsigned nv = 1u << (2 * first.nq);
  for (unsigned v = 0; v < nv; ++v) {
    const unsigned mid = lut_at(first.lut, v);
    r.lut |= (u64)lut_at(second.lut, mid) << (4 * v);
    r.sign |= (((first.sign >> v) ^ (second.sign >> mid)) & 1u) << v;
  }
  r.scale = std::min(first.scale, second.scale);
  r.n_gates = first.n_gates + second.n_gates;
  return r;
}

// a one-qubit gate as a two-qubit gate acting on position pos (0 or 1) of the pair (qa, qb)
FusedGate embed_1q(const FusedGate& g, int pos, int qa, int qb) {
  FusedGate r;
  r.nq = 2;
  r.q[0] = qa;
  r.q[1] = qb;
  r.scale = g.scale;
  r.n_gates = g.n_gates;
  for (unsigned v = 0; v < 16; ++v) {
    const unsigned mine = pos == 0 ? (v & 3u) : (v >> 2), other = pos == 0 ? (v >> 2) : (v & 3u);
    const unsigned img = lut_at(g.lut, mine);
    r.lut |= (u64)(pos == 0 ? (img | (other << 2)) : (other | (img << 2))) << (4 * v);
    r.sign |= ((g.sign >> mine) & 1u) << v;
  }
  return r;
}

FusedGate fused_from_op(const obp_op& op) {
  FusedGate g;
  g.nq = op.n_qubits;
  g.q[0] = op.qubit[0];
  g.q[1] = op.n_qubits == 2 ? op.qubit[1] : op.qubit[0];
  g.lut = op.lut;
  g.sign = op.sign_mask;
  g.scale = op.row_scale;
  g.n_gates = 1;
  return g;
}

// a gate program recorded for the persistent kernel instead of being launched op by op
struct ProgRec {
  std::vector<ProgEntry> entries;
  std::vector<unsigned char> pool;
  std::vector<int> klass;
  void add(int kind, int slot, const void* param, size_t bytes, int k) {
    const size_t off = (pool.size() + 15) & ~(size_t)15, padded = (bytes + 15) & ~(size_t)15;
    pool.resize(off + padded, 0);
    memcpy(pool.data() + off, param, bytes);
    entries.push_back(ProgEntry{kind, slot, (unsigned)off, (unsigned)padded});
    klass.push_back(k);
  }
};

#define OBP_BOUNCE_BYTES ((size_t)64 << 20)
#define OBP_INBOX_HDR 16  // u64 words before the first record (cursor + padding to 128 B)

void free_inboxes(obp_ctx* ctx) {
  for (int b = 0; b < 2; ++b) {
    for (size_t r = 0; r < ctx->peer_inbox[b].size(); ++r)
      if (ctx->peer_inbox[b][r] && ctx->peer_is_ipc[b][r]) cudaIpcCloseMemHandle(ctx->peer_inbox[b][r]);
    ctx->peer_inbox[b].clear();
    ctx->peer_is_ipc[b].clear();
    if (ctx->inbox[b]) cudaFree(ctx->inbox[b]);
    ctx->inbox[b] = nullptr;
  }
  ctx->inbox_cap = 0;
}

// kernel parameters of a rotation op under the current hash frame (epoch and slot are set by the caller)
RotP build_rotp(const obp_ctx* ctx, const obp_op& op, double atol) {
  RotP p;
  memset(&p, 0, sizeof(p));
  u64 hG = 0;
  int yG = 0;
  for (int j = 0; j < op.n_qubits; ++j) {
    const unsigned code = (op.gen_local >> (2 * j)) & 3u;
    if (!code) continue;
    const int w = op.qubit[j] / 64;
    const u64 bit = 1ull << (op.qubit[j] % 64);
    int sl = -1;
    for (int s = 0; s < p.ngw; ++s)
      if (p.gw[s] == w) sl = s;
    if (sl < 0) {
      sl = p.ngw++;
      p.gw[sl] = w;
    }
    if (code & 1u) { p.gx[sl] |= bit; hG ^= ctx->hx[op.qubit[j]]; }
    if (code & 2u) { p.gz[sl] |= bit; hG ^= ctx->hz[op.qubit[j]]; }
    if (code == 3u) yG++;
  }
  p.hG = hG;
  p.yG = yG;
  p.canon_slot = 0;
  {
    const u64 m = p.gx[0] | p.gz[0];
    const u64 low = m & (~m + 1);
    p.canon_bit = low;
    p.canon_z = (p.gx[0] & low) ? 0 : 1;
  }
  p.u0r = op.u0[0];
  p.u0i = op.u0[1];
  p.u1r = op.u1[0];
  p.u1i = op.u1[1];
  p.atol = atol;
  p.exact = (op.flags & OBP_FLAG_EXACT_COUNTERS) ? 1 : 0;
  p.simple = (op.u0[1] == 0.0 && op.u1[0] == 0.0) ? 1 : 0;
  return p;
}

int validate_op(obp_ctx* ctx, const obp_op& op) {
  const int maxq = op.kind == OBP_OP_CLIFFORD ? 2 : 4;
  if ((op.kind != OBP_OP_CLIFFORD && op.kind != OBP_OP_ROTATION) || op.n_qubits < 1 || op.n_qubits > maxq)
    return ctx->fail(OBP_ERR_UNSUPPORTED, "unsupported op kind or qubit count");
  for (int j = 0; j < op.n_qubits; ++j) {
    if (op.qubit[j] < 0 || op.qubit[j] >= ctx->nq) return ctx->fail(OBP_ERR_BADARG, "qubit out of range");
    for (int l = 0; l < j; ++l)
      if (op.qubit[l] == op.qubit[j]) return ctx->fail(OBP_ERR_BADARG, "duplicate qubit");
  }
  if (op.kind == OBP_OP_CLIFFORD && !lut_is_linear_bijection(op))
    return ctx->fail(OBP_ERR_BADARG, "Clifford table is not a linear bijection");
  if (op.kind == OBP_OP_ROTATION && (op.gen_local & ((1u << (2 * op.n_qubits)) - 1u)) == 0)
    return ctx->fail(OBP_ERR_BADARG, "rotation generator is the identity");
  if (op.kind == OBP_OP_CLIFFORD && op.group_size > 16) return ctx->fail(OBP_ERR_BADARG, "group_size > 16");
  return OBP_OK;
}

}  // namespace

// A gate program that has been enqueued but not yet waited for (obp_apply_ops_submit / _wait).
struct PendingOps {
  std::vector<obp_op> ops;
  ProgRec prog;
  std::vector<int> slot_of_op, slot_gates;
  int n_ops = 0, n_entries = 0, n_slots = 0, last_split = -1;
  u64 n_start = 0;
  bool prog_mode = false;
  bool segment = false;   // a run of rotations in shared memory (slot n_slots reports its outcome)
  std::vector<u64> frame_hx, frame_hz;  // the hash frame before the run, if the run holds Clifford-class gates (restored when it fails)
  double atol = 0.0;
  std::chrono::steady_clock::time_point t_host0;
};

#define OBP_NO_PENDING(ctx, what)                                                                                       \
  if ((ctx) && (ctx)->broken) return (ctx)->fail(OBP_ERR_OOM, what ": the context is unusable after a failed obp_reserve"); \
  if ((ctx) && (ctx)->pending) return (ctx)->fail(OBP_ERR_BADARG, what ": a submitted gate program is pending, call obp_apply_ops_wait first")


namespace {
template <int W>
int qwc_run(const ulonglong2* d_keys, ulonglong2* d_sorted, unsigned* d_degree, unsigned* d_order, unsigned* d_colour,
            unsigned* d_ncol, unsigned n, int n_sm, std::vector<unsigned>& order, unsigned* n_colours, std::string& err) {
  auto fail = [&](const char* what, cudaError_t e) {
    err = std::string("obp_qwc_group_count: ") + what + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? OBP_ERR_OOM : OBP_ERR_CUDA;
  };
  cudaError_t e;
  const unsigned nt = (n + OBP_QWC_TILE - 1) / OBP_QWC_TILE;
  const unsigned gy = std::max(1u, std::min(nt, (unsigned)(n_sm * 4 + nt - 1) / nt));
  k_qwc_degree<W><<<dim3(nt, gy), OBP_QWC_TILE>>>(d_keys, n, d_degree);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("degree kernel", e);
  std::vector<unsigned> degree(n);
  if ((e = cudaMemcpy(degree.data(), d_degree, n * sizeof(unsigned), cudaMemcpyDeviceToHost)) != cudaSuccess) return fail("degrees", e);
  // colouring order: descending degree, ties in row order (stable)
  order.resize(n);
  for (unsigned i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return degree[a] > degree[b]; });
  if ((e = cudaMemcpy(d_order, order.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("order", e);
  k_qwc_gather<<<std::min(nt * W, (unsigned)n_sm * 8), 256>>>(d_keys, d_order, n, W, d_sorted);
  const size_t smem = (size_t)((n + 31) / 32) * sizeof(unsigned);
  if ((e = cudaFuncSetAttribute(k_qwc_colour<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return fail("shared memory for the colour bit map", e);
  k_qwc_colour<W><<<1, 1024, smem>>>(d_sorted, n, d_colour, d_ncol);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail("colour kernel", e);
  if ((e = cudaMemcpy(n_colours, d_ncol, sizeof(unsigned), cudaMemcpyDeviceToHost)) != cudaSuccess) return fail("result", e);
  return OBP_OK;
}
}  // namespace

// ---- exchange-free sharding: a shard function orthogonal to the rotations to come -------------------
// owner(P) = top bits of the GF(2)-linear row hash, and a rotation with generator G moves a key to the
// owner  owner(P) ^ owner(G).  The hash frame is ours to choose: if the top bits of h(G_t) vanish for every
// rotation G_t the run is going to apply, no rotation ever sends a term to another rank and the shards evolve
// without any communication.  h(G_t) is a fixed linear function of the frame columns at split time (Clifford
// gates only re-express the frame), so the condition is a linear system on each top bit plane
// a in GF(2)^{2N}:  <a, G'_t> = 0  for the generators pulled back to the split-time frame.  It is solved
// without building a 2N x 2N tableau: run the frame forward over the future gates with random WIDE columns
// (OBP_PLAN_WORDS x 64 bits), collect the wide hashes s(G_t), take L in their null space, and set
// a[b] = parity(L . s_col[b]) — then <a, G'_t> = parity(L . s(G_t)) = 0 exactly, by linearity.
#define OBP_PLAN_WORDS 16  // 1024-bit columns: up to ~1000 planned rotations

struct WideCol {
  u64 w[OBP_PLAN_WORDS];
  WideCol& operator^=(const WideCol& o) {
    for (int k = 0; k < OBP_PLAN_WORDS; ++k) w[k] ^= o.w[k];
    return *this;
  }
};

void wide_frame_update(std::vector<WideCol>& sx, std::vector<WideCol>& sz, const obp_op& op) {
  const int nb = 2 * op.n_qubits;
  WideCol old_col[4], new_col[4];
  for (int j = 0; j < op.n_qubits; ++j) {
    old_col[2 * j] = sx[op.qubit[j]];
    old_col[2 * j + 1] = sz[op.qubit[j]];
  }
  unsigned inv[16];
  for (unsigned v = 0; v < (1u << nb); ++v) inv[(op.lut >> (4 * v)) & 15ull] = v;
  for (int j = 0; j < nb; ++j) {
    const unsigned pre = inv[1u << j];
    WideCol c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < nb; ++i)
      if (pre & (1u << i)) c ^= old_col[i];
    new_col[j] = c;
  }
  for (int j = 0; j < op.n_qubits; ++j) {
    sx[op.qubit[j]] = new_col[2 * j];
    sz[op.qubit[j]] = new_col[2 * j + 1];
  }
}

// Top-bit planes for `n_bits` owner bits: planes[i][2q] / [2q+1] = owner bit i of the x / z column of qubit q.
// Returns the number of rotations the plan covers (all of them unless the wide hash ran out of dimensions).
int plan_owner_planes(int nq, const obp_op* ops, int n_ops, int n_bits, int max_rotations,
                      std::vector<std::vector<unsigned char>>& planes) {
  u64 seed = 0x5EEDC0DE5EEDull;
  std::vector<WideCol> sx(nq), sz(nq);
  for (int q = 0; q < nq; ++q)
    for (int k = 0; k < OBP_PLAN_WORDS; ++k) {
      sx[q].w[k] = splitmix64(seed);
      sz[q].w[k] = splitmix64(seed);
    }
  const std::vector<WideCol> sx0 = sx, sz0 = sz;  // the columns at split time
  // Gaussian elimination of the rotation hashes as they come: basis rows in echelon form, pivot per row
  const int dims = 64 * OBP_PLAN_WORDS;
  std::vector<WideCol> basis;
  std::vector<int> pivot;
  int covered = 0;
  for (int k = 0; k < n_ops; ++k) {
    const obp_op& op = ops[k];
    if (op.kind == OBP_OP_CLIFFORD) {
      wide_frame_update(sx, sz, op);
      continue;
    }
    if (op.kind != OBP_OP_ROTATION) break;
    if (covered >= max_rotations) break;
    if ((int)basis.size() >= dims - n_bits - 8) break;  // keep a null space to choose from
    WideCol g;
    memset(&g, 0, sizeof(g));
    for (int j = 0; j < op.n_qubits; ++j) {
      const unsigned code = (op.gen_local >> (2 * j)) & 3u;
      if (code & 1u) g ^= sx[op.qubit[j]];
      if (code & 2u) g ^= sz[op.qubit[j]];
    }
    for (size_t r = 0; r < basis.size(); ++r)
      if ((g.w[pivot[r] >> 6] >> (pivot[r] & 63)) & 1ull) g ^= basis[r];
    int pv = -1;
    for (int b = 0; b < dims && pv < 0; ++b)
      if ((g.w[b >> 6] >> (b & 63)) & 1ull) pv = b;
    if (pv >= 0) {
      // keep the basis reduced: clear the new pivot from the older rows
      for (size_t r = 0; r < basis.size(); ++r)
        if ((basis[r].w[pv >> 6] >> (pv & 63)) & 1ull) basis[r] ^= g;
      basis.push_back(g);
      pivot.push_back(pv);
    }
    covered++;
  }
  // null space vectors: one per free coordinate f:  L = e_f + sum over basis rows r with bit f set of e_pivot[r]
  std::vector<char> is_pivot(dims, 0);
  for (int pv : pivot) is_pivot[pv] = 1;
  std::vector<int> free_coords;
  for (int b = 0; b < dims; ++b)
    if (!is_pivot[b]) free_coords.push_back(b);
  planes.assign(n_bits, std::vector<unsigned char>(2 * (size_t)nq, 0));
  for (int i = 0; i < n_bits; ++i) {
    // a random combination of free-coordinate null vectors (different per owner bit, never empty)
    WideCol L;
    memset(&L, 0, sizeof(L));
    int used = 0;
    for (size_t f = 0; f < free_coords.size(); ++f) {
      const bool take = (splitmix64(seed) & 3ull) == 0 || (int)f == i;
      if (!take) continue;
      const int fc = free_coords[f];
      L.w[fc >> 6] ^= 1ull << (fc & 63);
      for (size_t r = 0; r < basis.size(); ++r)
        if ((basis[r].w[fc >> 6] >> (fc & 63)) & 1ull) L.w[pivot[r] >> 6] ^= 1ull << (pivot[r] & 63);
      used++;
    }
    (void)used;
    auto par = [&](const WideCol& c) {
      u64 acc = 0;
      for (int k = 0; k < OBP_PLAN_WORDS; ++k) acc ^= c.w[k] & L.w[k];
      return (unsigned char)(__builtin_popcountll(acc) & 1);
    };
    for (int q = 0; q < nq; ++q) {
      planes[i][2 * q] = par(sx0[q]);
      planes[i][2 * q + 1] = par(sz0[q]);
    }
  }
  return covered;
}

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

uint64_t obp_bytes_per_term(int n_qubits) {
  const u64 W = (u64)std::max(1, (n_qubits + 63) / 64);
  // key columns + coef + hash, twice (table + per-slice snapshot), + spare16 + spare8 + dest +
  // index (T >= 2*cap, up to 4*cap)
  return 2 * (16 * W + 16 + 8) + 16 + 8 + 4 + 32;
}

int obp_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes) {
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "obp_device_memory: cudaSetDevice failed";
    return OBP_ERR_CUDA;
  }
  size_t f = 0, t = 0;
  if (cudaMemGetInfo(&f, &t) != cudaSuccess) {
    g_create_error = "obp_device_memory: cudaMemGetInfo failed";
    return OBP_ERR_CUDA;
  }
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return OBP_OK;
}

int obp_create(int device, int n_qubits, uint64_t capacity, obp_ctx** out) {
  if (!out || n_qubits < 1 || n_qubits > 64 * OBP_MAX_WORDS || capacity < 1 || capacity >= 0xffffffffull) {
    g_create_error = "obp_create: bad argument";
    return OBP_ERR_BADARG;
  }
  obp_ctx* ctx = new obp_ctx();
  auto bail = [&](int code) {
    g_create_error = ctx->err;
    obp_destroy(ctx);
    return code;
  };
  ctx->device = device;
  ctx->nq = n_qubits;
  ctx->W = (n_qubits + 63) / 64;
  memset(&ctx->prof, 0, sizeof(ctx->prof));
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    ctx->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return bail(OBP_ERR_CUDA);
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->n_sm = prop.multiProcessorCount;
  if (prop.major < 10) {
    ctx->err = "obp_b200 requires an sm_100a (Blackwell B200) device";
    return bail(OBP_ERR_UNSUPPORTED);
  }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    ctx->err = "cudaStreamCreate failed";
    return bail(OBP_ERR_CUDA);
  }
  ctx->stream = ctx->own_stream;
  auto body = [&]() -> int {
    int rc = ensure_ops(ctx, 256);  // also allocates the table state (d_st / h_st) and the start stamp
    if (rc) return rc;
    CU(cudaMallocHost(&ctx->h_scalar, sizeof(double) * 2));
    CU(cudaMalloc(&ctx->d_partial, sizeof(double) * (size_t)ctx->n_sm * 8 * OBP_TRUNC_NODES));
    CU(cudaMalloc(&ctx->d_scalar, sizeof(double) * 2));
    CU(cudaMalloc(&ctx->d_cols, sizeof(u64) * 2 * 64 * ctx->W));
    CU(cudaMalloc(&ctx->d_cnt, sizeof(ShardCnt)));
    CU(cudaMallocHost(&ctx->h_cnt, sizeof(ShardCnt)));
    rc = alloc_table(ctx, capacity);
    if (rc) return rc;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->rot_ctas_per_sm, k_rotate, OBP_BS, 0));
    if (ctx->rot_ctas_per_sm < 1) ctx->rot_ctas_per_sm = 1;
    if (ctx->W <= 2 && prop.cooperativeLaunch) {
      if (ctx->W == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prog_ctas_per_sm, k_program<1>, OBP_BS, 0));
      else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prog_ctas_per_sm, k_program<2>, OBP_BS, 0));
    }
    if (prop.cooperativeLaunch)
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->trunc_ctas_per_sm, k_trunc_search, OBP_BS, 0));
    if (ctx->W <= 2) {
      // the slice-in-shared-memory kernels opt in to (nearly) the whole shared memory of an SM
      int optin = 0;
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
      const int want = std::min(optin, 227 * 1024) - 1024;
      bool ok = want >= 64 * 1024;
      const void* fns[4] = {(const void*)k_seg_main<1, 128, 4>, (const void*)k_seg_main<2, 128, 4>,
                            (const void*)k_seg_main<1, 512, 1>, (const void*)k_seg_main<2, 512, 1>};
      for (int k = 0; k < 4 && ok; ++k) {
        ok = cudaFuncSetAttribute(fns[k], cudaFuncAttributeMaxDynamicSharedMemorySize, k < 2 ? want / 4 : want) == cudaSuccess &&
             cudaFuncSetAttribute(fns[k], cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess;
      }
      if (ok) ctx->seg_smem_max = want;
      else cudaGetLastError();
    }
    CU(cudaMalloc(&ctx->d_ts, sizeof(TruncState)));
    CU(cudaMallocHost(&ctx->h_ts, sizeof(TruncState)));
    {
      const char* env = getenv("OBP_B200_NO_PROGRAM");
      ctx->use_program = !(env && env[0] && env[0] != '0');
      if (const char* e1 = getenv("OBP_B200_TRACE_HOST")) ctx->trace_host = e1[0] && e1[0] != '0';
      if (const char* e2 = getenv("OBP_B200_PROG_MIN_GRID")) ctx->prog_min_grid = std::max(1, atoi(e2));
      if (const char* e3 = getenv("OBP_B200_PROG_REACH_SHIFT")) ctx->prog_reach_shift = std::max(0, atoi(e3));
      if (const char* e6 = getenv("OBP_B200_PROG_DEBUG")) ctx->prog_debug = atoi(e6) & ~1;
      if (const char* e7 = getenv("OBP_B200_PROG_MAX_ROWS")) ctx->prog_max_rows = (u64)std::max(0ll, atoll(e7));
      if (const char* e8 = getenv("OBP_B200_NO_SEGMENTS")) ctx->use_segments = !(e8[0] && e8[0] != '0');
      if (const char* e12 = getenv("OBP_B200_SEG_MIXED")) ctx->seg_mixed = e12[0] && e12[0] != '0';
      if (const char* e9 = getenv("OBP_B200_SEG_MIN_GATES")) ctx->seg_min_gates = std::max(1, atoi(e9));
      if (const char* e10 = getenv("OBP_B200_SEG_CAP_ROWS")) ctx->seg_cap_rows = std::max(0, atoi(e10));
      if (const char* e11 = getenv("OBP_B200_SEG_PIECES_MIN_ROWS")) ctx->seg_pieces_min_rows = (u64)std::max(0ll, atoll(e11));
    }
    init_frame(ctx);
    return set_counts(ctx, 0);
  };
  const int rc = body();
  if (rc) return bail(rc);
  *out = ctx;
  return OBP_OK;
}

void obp_destroy(obp_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch) obp_destroy(ctx->scratch);
  delete ctx->pending;
  free_table(ctx);
  free_snapshot(ctx->snap);
  cudaFree(ctx->d_res);
  cudaFree(ctx->d_partial);
  cudaFree(ctx->d_scalar);
  cudaFree(ctx->d_cols);
  cudaFree(ctx->d_cnt);
  cudaFree(ctx->d_prog);
  cudaFree(ctx->d_seg);
  cudaFree(ctx->d_stage);
  cudaFree(ctx->ord_minidx);
  cudaFree(ctx->ord_flag);
  cudaFree(ctx->ord_pos);
  if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
  cudaFree(ctx->d_ts);
  if (ctx->h_ts) cudaFreeHost(ctx->h_ts);
  if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
  free_inboxes(ctx);
  if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
  cudaFree(ctx->d_gens);
  cudaFree(ctx->d_factors);
  if (ctx->h_res) cudaFreeHost(ctx->h_res);
  if (ctx->h_scalar) cudaFreeHost(ctx->h_scalar);
  for (auto& ep : ctx->ev_pending) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  for (auto& ep : ctx->ev_free) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  if (ctx->tm_a) cudaEventDestroy(ctx->tm_a);
  if (ctx->tm_b) cudaEventDestroy(ctx->tm_b);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* obp_last_error(const obp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int obp_set_stream(obp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return OBP_OK;
}

int obp_count(obp_ctx* ctx, uint64_t* n) {
  if (!ctx || !n) return OBP_ERR_BADARG;
  *n = ctx->n_live;
  return OBP_OK;
}

static int load_impl(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
                     uint64_t n, double atol, obp_stats* stats, bool keep_snapshot);

int obp_load(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
             uint64_t n, double atol, obp_stats* stats) {
  return load_impl(ctx, x_words, z_words, coeffs, n, atol, stats, false);
}

int obp_replace(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
                uint64_t n, double atol, obp_stats* stats) {
  return load_impl(ctx, x_words, z_words, coeffs, n, atol, stats, true);
}

static int load_impl(obp_ctx* ctx, const uint64_t* x_words, const uint64_t* z_words, const double* coeffs,
                     uint64_t n, double atol, obp_stats* stats, bool keep_snapshot) {
  OBP_NO_PENDING(ctx, "obp_load");
  if (!ctx || (n && (!x_words || !z_words || !coeffs))) return OBP_ERR_BADARG;
  if (n > ctx->cap) return ctx->fail(OBP_ERR_CAPACITY, "obp_load: more rows than capacity");
  CU(cudaSetDevice(ctx->device));
  if (!keep_snapshot) ctx->snap.valid = false;  // (the snapshot carries its own hash frame: it survives a replace)
  if (!keep_snapshot) {  // a new observable: what the last one taught about runs in shared memory no longer holds
    ctx->seg_block_rows = ~0ull;
    ctx->seg_max_gates = OBP_SEG_MAXG;
    ctx->seg_growth = 4.0;
  }
  init_frame(ctx);
  int rc = upload_cols(ctx);
  if (rc) return rc;
  if (n) {
    std::vector<ulonglong2> col(n);
    for (int w = 0; w < ctx->W; ++w) {
      for (u64 i = 0; i < n; ++i) col[i] = make_ulonglong2(x_words[i * ctx->W + w], z_words[i * ctx->W + w]);
      CU(cudaMemcpy(ctx->xz[w], col.data(), n * sizeof(ulonglong2), cudaMemcpyHostToDevice));
    }
    // a row given with coefficient exactly 0 must stay distinguishable from a tombstone only until
    // the zero filter runs; with atol < 0 (no filter) such rows are dropped at load
    CU(cudaMemcpy(ctx->coef, coeffs, n * sizeof(double2), cudaMemcpyHostToDevice));
  }
  rc = set_counts(ctx, n);
  if (rc) return rc;
  rc = ensure_ops(ctx, 4);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat) * 2, ctx->stream));
  CU(cudaMemsetAsync(ctx->index, 0xFF, ctx->T * sizeof(u64), ctx->stream));
  ctx->index_stale = false;
  ctx->T_eff = ctx->T;
  Tab t = make_tab(ctx);
  if (n && atol == OBP_ATOL_RAW) {  // simplify=False: rows as given, duplicates and zeros included
    Bracket b(ctx, OBP_K_INDEX, 2);
    const int g = grid_for(ctx, n);
    k_hash_rows<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->d_cols, n);
    k_revive_zeros<<<g, OBP_BS, 0, ctx->stream>>>(t, 0, n);
    CU(cudaGetLastError());
  } else if (n) {
    Bracket b(ctx, OBP_K_INDEX);
    const int g = grid_for(ctx, n);
    k_hash_rows<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->d_cols, n);
    if (ctx->n_ranks > 1) k_shard_filter<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->owner_shift, (u64)ctx->rank);
    ctx->epoch++;
    rc = reset_first_appearance(ctx, n);
    if (rc) return rc;
    k_index_build<<<g, OBP_BS, 0, ctx->stream>>>(t, 1, atol, ctx->epoch, 0, ctx->row_order ? ctx->ord_minidx : nullptr);
    ctx->epoch++;
    k_trim<<<g, OBP_BS, 0, ctx->stream>>>(t, atol, ctx->epoch, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  if (rc) return rc;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    const OpStat& o = ctx->h_ops[0];
    // rows given as exact zeros count as filtered raw rows; rows of other ranks do not count at all
    const u64 n_mine = n - (ctx->n_ranks > 1 ? ctx->h_st->removed : 0);
    stats->n_in = n_mine;
    stats->raw_rows = n_mine;
    stats->num_unique = o.keys_present;
    stats->num_duplicate = o.rows_surv - o.keys_present;
    stats->num_trimmed = (n_mine - o.rows_surv) + o.cancelled;
    stats->sum_trimmed[0] = o.sum_re;
    stats->sum_trimmed[1] = o.sum_im;
    stats->n_out = ctx->n_live;
  }
  if (ctx->n_dense != ctx->n_live) {
    rc = atol == OBP_ATOL_RAW ? compact(ctx) : compact_after_merge(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  if (stats) stats->n_dense = ctx->n_dense;
  return OBP_OK;
}

int obp_download(obp_ctx* ctx, uint64_t* x_words, uint64_t* z_words, double* coeffs,
                 uint64_t max_rows, uint64_t* n_out) {
  OBP_NO_PENDING(ctx, "obp_download");
  if (!ctx || !n_out) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  if (ctx->n_dense != ctx->n_live) {
    int rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  const u64 n = ctx->n_live;
  *n_out = n;
  if (n == 0) return OBP_OK;
  if (n > max_rows || !x_words || !z_words || !coeffs) return ctx->fail(OBP_ERR_BADARG, "obp_download: buffer too small");
  // transpose to the caller's [n][W] layout on the device, then three plain copies
  // staging: the compaction scratch when the packed rows fit it, else a buffer kept with the context
  // (a cudaMalloc/cudaFree pair per download costs far more than the copy itself)
  u64* stage = nullptr;
  const size_t words = (size_t)n * ctx->W;
  if (words <= ctx->cap) {
    stage = (u64*)ctx->spare16;  // cap * 16 B = room for 2 * words u64
  } else {
    if (words * 2 * sizeof(u64) > ctx->stage_bytes) {
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_stage);
  cudaFree(ctx->ord_minidx);
  cudaFree(ctx->ord_flag);
  cudaFree(ctx->ord_pos);
  if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
      ctx->d_stage = nullptr;
      ctx->stage_bytes = 0;
      const size_t nb = words * 2 * sizeof(u64) + (words >> 2) * sizeof(u64);
      CU(cudaMalloc(&ctx->d_stage, nb));
      ctx->stage_bytes = nb;
    }
    stage = ctx->d_stage;
  }
  k_pack_rows<<<grid_for(ctx, n), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), stage, stage + words, n);
  CU(cudaGetLastError());
  // device -> pinned bounce buffer -> caller's (pageable) arrays, in chunks of at most 64 MiB: a
  // direct copy into freshly allocated pageable memory runs at a tenth of the pinned rate
  if (!ctx->h_bounce) {
    CU(cudaMallocHost(&ctx->h_bounce, OBP_BOUNCE_BYTES));
  }
  struct Part { const void* src; void* dst; size_t bytes; };
  const Part parts[3] = {{stage, x_words, words * sizeof(u64)},
                         {stage + words, z_words, words * sizeof(u64)},
                         {ctx->coef, coeffs, (size_t)n * sizeof(double2)}};
  for (const Part& p : parts) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p.dst) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      // the caller handed us page-locked memory (e.g. a torch pinned tensor): one direct copy
      CU(cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      continue;
    }
    cudaGetLastError();  // a pageable pointer makes cudaPointerGetAttributes report an error on some drivers
    for (size_t off = 0; off < p.bytes; off += OBP_BOUNCE_BYTES) {
      const size_t len = std::min<size_t>(OBP_BOUNCE_BYTES, p.bytes - off);
      CU(cudaMemcpyAsync(ctx->h_bounce, (const char*)p.src + off, len, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      memcpy((char*)p.dst + off, ctx->h_bounce, len);
    }
  }
  return OBP_OK;
}

// -------------------------------------------------------------------------------------------------
// Runs ops[0, n_ops) — or, on a sharded table, up to the first rotation whose partner keys live on
// another rank: *n_done = ops executed, *peer_xor = owner bits of that rotation's generator hash.
// Second half of a gate program: wait for the device, account, compact if needed, fill the statistics.

// ---- runs of rotations in shared memory (obp_segment.cuh) -------------------------------------------------------
// A gate program (or its prefix) that can run in shared memory: rotations and Clifford-class gates only, at least
// seg_min_gates rotations, no per-gate counter requests except the reference's counters on a LAST gate that is a rotation.
// Returns the number of leading ops the run covers (0: not eligible): everything, or everything up to and including the
// last rotation when Clifford-class gates with counters of their own follow it.
static int segment_span(const obp_ctx* ctx, const obp_op* ops, int n_ops) {
  if (!ctx->use_segments || ctx->seg_smem_max <= 0 || ctx->W > 2 || ctx->n_ranks != 1 || ctx->row_order || ctx->flag_sync) return 0;
  if (ctx->n_live == 0 || ctx->n_dense >= (1ull << 31) || ctx->n_live >= ctx->seg_block_rows) return 0;
  int n_rot = 0, last_rot = -1;
  for (int k = 0; k < n_ops; ++k) {
    if (ops[k].kind != OBP_OP_ROTATION && ops[k].kind != OBP_OP_CLIFFORD) return 0;
    const unsigned allowed = (k == n_ops - 1) ? OBP_FLAG_EXACT_COUNTERS : 0u;
    if (ops[k].flags & ~allowed) return 0;
    if (ops[k].kind == OBP_OP_ROTATION) { n_rot++; last_rot = k; }
  }
  if (n_rot < ctx->seg_min_gates) return 0;
  if (!ctx->seg_mixed && n_rot != n_ops) return 0;
  const bool tail_counts = ops[n_ops - 1].kind == OBP_OP_CLIFFORD && (ops[n_ops - 1].flags & OBP_FLAG_EXACT_COUNTERS);
  return tail_counts ? last_rot + 1 : n_ops;
}

// bytes of dynamic shared memory k_seg_main needs for `cap` rows per partition
static size_t seg_smem_bytes(int W, int cap, int n_index, int ng) {
  return (size_t)cap * (16 * W + 16 + 8 + 1) + (size_t)ng * sizeof(SegGate) + 2 * (size_t)((ng + 1) & ~1) * 4 + (size_t)n_index * 2 + 64;
}

// the largest partition (rows, multiple of 64) whose arrays and index (at most half full) fit `budget` bytes
static void seg_geometry(int W, int ng, size_t budget, int limit_rows, int* cap_rows, int* n_index) {
  int cap = 4096;
  while (cap > 64 && seg_smem_bytes(W, cap, (int)next_pow2((u64)2 * cap), ng) > budget) cap -= 64;
  if (limit_rows > 0) cap = std::min(cap, std::max(limit_rows, 8));
  *cap_rows = cap;
  *n_index = (int)next_pow2((u64)std::max(2 * cap, 64));
}

// The partition map of a run: a GF(2)-linear map keys -> OBP_SEG_FBITS bits that vanishes on every generator.
// Key bit positions: word w, plane (x = 0, z = 1), bit b -> 128 w + 64 plane + b.
static void seg_partition_map(const std::vector<SegGate>& gates, int W, u64 fmask[OBP_SEG_FBITS][4]) {
  struct V { u64 w[4]; };
  std::vector<V> basis;     // row echelon form of the generators: a row holds no pivot of the rows before it
  std::vector<int> pivot;
  int row_of_pivot[256];
  for (int b = 0; b < 256; ++b) row_of_pivot[b] = -1;
  for (const SegGate& g : gates) {
    V v{{g.gx[0], g.gz[0], W == 2 ? g.gx[1] : 0ull, W == 2 ? g.gz[1] : 0ull}};
    // reduce by the rows whose pivot v holds (generators on disjoint qubits: none — the common case costs O(weight))
    for (bool again = true; again;) {
      again = false;
      for (int j = 0; j < 4 && !again; ++j)
        for (u64 w = v.w[j]; w; w &= w - 1) {
          const int k = row_of_pivot[64 * j + __builtin_ctzll(w)];
          if (k >= 0) {
            for (int q = 0; q < 4; ++q) v.w[q] ^= basis[(size_t)k].w[q];
            again = true;
            break;
          }
        }
    }
    int pv = -1;
    for (int j = 0; j < 4 && pv < 0; ++j)
      if (v.w[j]) pv = 64 * j + __builtin_ctzll(v.w[j]);
    if (pv < 0) continue;  // dependent on the generators before it
    row_of_pivot[pv] = (int)basis.size();
    basis.push_back(v);
    pivot.push_back(pv);
  }
  // random weights on the free positions; the pivots follow from f(G) = 0, last row first (a row may hold pivots of
  // LATER rows only, and those are settled by then)
  unsigned r[256];
  u64 seed = 0x5E6B200D15EA5Eull;
  for (int b = 0; b < 256; ++b) r[b] = (unsigned)(splitmix64(seed) & ((1u << OBP_SEG_FBITS) - 1u));
  for (size_t k = basis.size(); k-- > 0;) {
    unsigned acc = 0;
    for (int j = 0; j < 4; ++j)
      for (u64 w = basis[k].w[j]; w; w &= w - 1) {
        const int b = 64 * j + __builtin_ctzll(w);
        if (b != pivot[k]) acc ^= r[b];
      }
    r[pivot[k]] = acc;
  }
  for (int j = 0; j < OBP_SEG_FBITS; ++j)
    for (int q = 0; q < 4; ++q) {
      u64 m = 0;
      for (int b = 0; b < 64; ++b)
        if ((r[64 * q + b] >> j) & 1u) m |= 1ull << b;
      fmask[j][q] = m;
    }
}

static int apply_ops_impl(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats,
                          int32_t* n_done, int32_t* peer_xor, bool defer);
static int apply_segment(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats, bool defer,
                         bool inner_piece);
static int apply_in_pieces(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats, bool defer, int piece);

static int finish_ops(obp_ctx* ctx, PendingOps& P, obp_stats* stats) {
  int rc;
  if (stats) memset(stats, 0, sizeof(*stats));
  const auto t_host1 = std::chrono::steady_clock::now();
  // state + start stamp + counters (+ the per-op barrier stamps while profiling) in one copy
  const bool by_kernel = (P.prog_mode && P.n_entries > 0) || P.segment;
  const int sync_rc = sync_state(ctx, P.n_slots + (P.segment ? 1 : 0) + (ctx->prof_on ? P.n_entries : 0), by_kernel);
  if (P.segment && !sync_rc) {
    // a run of rotations in shared memory: slot n_slots holds its outcome.  The input table is untouched until here.
    const unsigned ovf = (unsigned)ctx->h_ops[P.n_slots].n_out;
    if (ovf) {
      std::vector<obp_op> ops = P.ops;  // (P may be ctx->pending's storage)
      const int n_ops = P.n_ops;
      const double atol = P.atol;
      ctx->seg_growth = 16.0;
      if (!P.frame_hx.empty()) {  // the gates are applied again, by other kernels: back to the frame the run started in
        ctx->hx = P.frame_hx;
        ctx->hz = P.frame_hz;
        ctx->cols_dirty = true;
      }
      const bool can_cut = n_ops >= 2 * std::max(ctx->seg_min_gates, 2);
      if (can_cut) {
        // Too many rows share a partition: with fewer gates per run the partition map has fewer generators to
        // vanish on, so the classes shrink.  Later runs of this observable are cut into pieces a quarter as long.
        ctx->seg_max_gates = std::max(ctx->seg_min_gates, (n_ops + 3) / 4);
        if (P.n_start >= ctx->seg_pieces_min_rows) {  // a big table: worth repeating this very run in pieces
          ctx->n_seg_retries++;
          return apply_in_pieces(ctx, ops.data(), n_ops, atol, stats, false, ctx->seg_max_gates);
        }
      } else {
        // (even the shortest run grows classes no CTA can hold: stay with the per-gate kernels from half this size on)
        ctx->seg_block_rows = std::max<u64>(P.n_start / 2, 64);
      }
      ctx->n_seg_fallbacks++;
      const bool keep = ctx->use_segments;
      ctx->use_segments = false;
      rc = apply_ops_impl(ctx, ops.data(), n_ops, atol, stats, nullptr, nullptr, false);  // the per-gate kernels
      ctx->use_segments = keep;
      return rc;
    }
    // commit: the output columns become the table (the old columns are the next run's output)
    for (int w = 0; w < ctx->W; ++w) std::swap(ctx->xz[w], ctx->alt_xz[w]);
    std::swap(ctx->coef, ctx->alt_coef);
    std::swap(ctx->hash, ctx->alt_hash);
    ctx->index_stale = true;
    if (ctx->seg_T_new && ctx->h_ops[P.n_slots].cancelled == 0) {  // the run published every row it wrote
      ctx->index_stale = false;
      ctx->T_eff = ctx->seg_T_new;
    }
    ctx->index_touched = false;  // (set again by the next program that consults the index)
    ctx->n_seg_runs++;
    ctx->seg_growth = (double)ctx->n_live / (double)std::max<u64>(P.n_start, 1);
  }
  if (ctx->trace_host) {
    const auto t_host2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[obp host] ops %d slots %d: enqueue %.1f us, wait %.1f us\n", P.n_ops, P.n_slots,
            std::chrono::duration<double, std::micro>(t_host1 - P.t_host0).count(),
            std::chrono::duration<double, std::micro>(t_host2 - t_host1).count());
  }
  ctx->last_split_is_last_op = P.last_split >= 0 && P.last_split == P.n_ops - 1;
  if (sync_rc && sync_rc != OBP_ERR_CAPACITY) return sync_rc;
  if (P.prog_mode && ctx->prof_on) {  // per-class device time from the stamps taken at the grid barriers
    u64 t_prev = *ctx->h_tbegin;
    for (int k = 0; k < P.n_entries; ++k) {
      const u64 t_end = ctx->h_ops[P.n_slots + k].t_end;
      if (t_end >= t_prev) ctx->prof.ms[P.prog.klass[k]] += (double)(t_end - t_prev) * 1e-6;
      t_prev = t_end;
    }
  }

  // term-gate updates: live terms at entry of every op (also accounted when the table overflowed:
  // the kernels ran, the caller rolls the slice back)
  u64 updates = 0, n_in = P.n_start, n_in_last = P.n_start;
  for (int s = 0; s < P.n_slots; ++s) {
    updates += n_in * (u64)P.slot_gates[s];
    n_in_last = n_in;
    n_in = ctx->h_ops[s].n_out;
  }
  {
    // per-class term counts for the roofline
    u64 nin = P.n_start;
    for (int k = 0, s_prev = -1; k < P.n_ops; ++k) {
      const int s = P.slot_of_op[k];
      if (s != s_prev && s_prev >= 0) nin = ctx->h_ops[s_prev].n_out;
      ctx->prof.term_updates[P.segment ? OBP_K_SEGMENT : P.ops[k].kind == OBP_OP_ROTATION ? OBP_K_ROTATE : OBP_K_CLIFFORD] += nin;
      s_prev = s;
    }
  }
  ctx->last_op_sizes.assign((size_t)P.n_ops, 0);
  for (int k = 0; k < P.n_ops; ++k)
    if (P.slot_of_op[k] >= 0 && P.slot_of_op[k] < P.n_slots) ctx->last_op_sizes[k] = ctx->h_ops[P.slot_of_op[k]].n_out;
  if (!P.segment && ctx->carry_sizes.empty()) {
    ctx->last_ops = P.ops;
    ctx->last_slot_of_op = P.slot_of_op;
    ctx->last_n_slots = P.n_slots;
    ctx->last_n_start = P.n_start;
  } else {
    ctx->last_ops.clear();  // (runs in shared memory report the counters of their last gate only)
  }
  // a run that was cut into pieces: this is its last piece, the earlier ones are accounted here
  updates += ctx->carry_updates;
  if (!ctx->carry_sizes.empty()) ctx->last_op_sizes.insert(ctx->last_op_sizes.begin(), ctx->carry_sizes.begin(), ctx->carry_sizes.end());
  ctx->carry_updates = 0;
  ctx->carry_sizes.clear();
  if (sync_rc) {
    if (stats) {
      stats->updates = updates;      // work done before the overflow was noticed (the caller rolls the slice back)
      stats->n_out = ctx->n_live;    // live rows when the table overflowed: a lower bound of the slice's size
    }
    return sync_rc;
  }
  if (stats) {
    const obp_op& last = P.ops[P.n_ops - 1];
    const OpStat& o = ctx->h_ops[P.n_slots - 1];
    const u64 k2 = (u64)last.n_components * last.n_components;
    stats->n_in = n_in_last;
    stats->raw_rows = n_in_last * k2;
    stats->updates = updates;
    if (last.flags & OBP_FLAG_EXACT_COUNTERS) {
      if (last.kind == OBP_OP_ROTATION) {
        stats->num_unique = o.keys_present;
        stats->num_duplicate = o.rows_surv - o.keys_present;
        stats->num_trimmed = (stats->raw_rows - o.rows_surv) + o.cancelled;
        stats->sum_trimmed[0] = o.sum_re;
        stats->sum_trimmed[1] = o.sum_im;
      } else {
        // cosets hit = sum_c N_c / c ; keys present = |D| * cosets
        u64 cosets = 0;
        for (unsigned c = 1; c <= 16; ++c) cosets += o.ncnt[c] / c;
        const u64 rows = o.n_surv * k2;
        const u64 uniq = (u64)last.group_size * cosets;
        stats->num_unique = uniq;
        stats->num_duplicate = rows - uniq;
        stats->num_trimmed = (stats->raw_rows - rows) + (uniq - o.n_surv);
      }
    }
  }
  // keep tombstones below a quarter of the table
  if (ctx->n_dense - ctx->n_live > std::max<u64>(1024, ctx->n_live / 4)) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  if (stats) {
    stats->n_out = ctx->n_live;
    stats->n_dense = ctx->n_dense;
  }
  return OBP_OK;
}

// A run of rotations too long for one pass in shared memory (more gates than the parameter block holds, or a run that
// failed because too many rows shared a partition): pieces of at most `piece` gates, one after the other.  Every piece
// is a gate program of its own (its own partition map, one read and one write of the table); the statistics of the
// call are those of the last piece plus the update counts of the pieces before it.
static int apply_in_pieces(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats, bool defer, int piece) {
  // (called for a piece that failed and is cut again: what was carried for that piece is carried on)
  u64 updates = ctx->carry_updates;
  std::vector<u64> sizes = ctx->carry_sizes;
  const int stub = std::min(ctx->seg_min_gates, 8);  // no piece shorter than this: the one before it takes its gates
  piece = std::max(1, std::min(piece, OBP_SEG_MAXG - stub));
  for (int a = 0; a < n_ops;) {
    int b = std::min(n_ops, a + piece);
    if (n_ops - b > 0 && n_ops - b < stub) b = n_ops;
    const bool last = b == n_ops;
    if (last) {
      // the last piece carries the call: deferred if the caller asked for that, its statistics complete the call's
      ctx->carry_updates = updates;
      ctx->carry_sizes = sizes;
      return apply_segment(ctx, ops + a, b - a, atol, stats, defer, false);  // (finish_ops adds the carry)
    }
    std::vector<obp_op> part(ops + a, ops + b);
    for (obp_op& o : part) o.flags &= ~(unsigned)OBP_FLAG_EXACT_COUNTERS;
    obp_stats st;
    ctx->carry_updates = 0;
    ctx->carry_sizes.clear();
    const int rc = apply_segment(ctx, part.data(), b - a, atol, &st, false, true);
    if (rc) {
      if (stats) { memset(stats, 0, sizeof(*stats)); stats->updates = updates + st.updates; stats->n_out = st.n_out; }
      return rc;
    }
    updates += st.updates;
    sizes.insert(sizes.end(), ctx->last_op_sizes.begin(), ctx->last_op_sizes.end());
    a = b;
  }
  return OBP_OK;
}

// A run of rotations (the whole call) applied in shared memory: partition the table by a linear map that vanishes on
// every generator, one CTA per partition applies all gates, the survivors go to the second set of columns
// (obp_segment.cuh).  The table is committed in finish_ops once the outcome is known.
// inner_piece: a piece of a cut run that is not the last one (nobody looks keys up before the next piece).
static int apply_segment(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats, bool defer,
                         bool inner_piece) {
  const auto t_host0 = std::chrono::steady_clock::now();
  const int W = ctx->W, ng = n_ops;
  int rc = ensure_ops(ctx, ng + 2);
  if (rc) return rc;
  // second set of columns and the partition scratch: allocated on first use (a failed allocation turns the path off)
  if (!ctx->alt_coef) {
    bool ok = true;
    for (int w = 0; w < W && ok; ++w) ok = cudaMalloc(&ctx->alt_xz[w], ctx->cap * sizeof(ulonglong2)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->alt_coef, ctx->cap * sizeof(double2)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->alt_hash, ctx->cap * sizeof(u64)) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      free_seg_columns(ctx);
      ctx->use_segments = false;
      return apply_ops_impl(ctx, ops, n_ops, atol, stats, nullptr, nullptr, defer);
    }
  }
  const size_t part_bytes = (4 * (size_t)(OBP_SEG_MAXP + 1) * sizeof(unsigned) + 255) & ~(size_t)255;
  if (!ctx->d_seg) {
    CU(cudaMalloc(&ctx->d_seg, part_bytes + sizeof(SegRes)));
    CU(cudaMemsetAsync(ctx->d_seg, 0, part_bytes + sizeof(SegRes), ctx->stream));  // (every run leaves it clean again)
  }
  unsigned* d_cnt = (unsigned*)ctx->d_seg;
  unsigned* d_off = d_cnt + (OBP_SEG_MAXP + 1);
  unsigned* d_cur = d_off + (OBP_SEG_MAXP + 1);
  unsigned* d_redo = d_cur + (OBP_SEG_MAXP + 1);
  SegRes* d_res = (SegRes*)(ctx->d_seg + part_bytes);

  // gate parameters.  Clifford-class gates re-express the hash frame as they go (the generator hashes of the rotations
  // after them are taken in the frame of their place in the program) and, for the partition map, the generators are
  // pulled back through them to the keys the run starts from: a row P becomes C(P), a later rotation mixes C(P) with
  // C(P) xor G = C(P xor C^-1(G)), so the map has to vanish on C^-1(G) — the same column update as the hash frame,
  // on 256-bit columns that start as the identity.
  std::vector<SegGate> gates((size_t)ng), gens;
  bool has_cliff = false;
  for (int k = 0; k < ng; ++k) has_cliff = has_cliff || ops[k].kind == OBP_OP_CLIFFORD;
  std::vector<u64> hx0, hz0;
  std::vector<WideCol> sx, sz;
  if (has_cliff) {
    hx0 = ctx->hx;
    hz0 = ctx->hz;
    sx.resize((size_t)ctx->nq);
    sz.resize((size_t)ctx->nq);
    for (int q = 0; q < ctx->nq; ++q) {
      memset(&sx[(size_t)q], 0, sizeof(WideCol));
      memset(&sz[(size_t)q], 0, sizeof(WideCol));
      sx[(size_t)q].w[2 * (q / 64)] = 1ull << (q % 64);
      sz[(size_t)q].w[2 * (q / 64) + 1] = 1ull << (q % 64);
    }
  }
  auto build_gates = [&]() {
  for (int k = 0; k < ng; ++k) {
    SegGate& g = gates[k];
    memset(&g, 0, sizeof(g));
    if (ops[k].kind == OBP_OP_CLIFFORD) {
      const obp_op& op = ops[k];
      const int qa = op.qubit[0], qb = op.n_qubits == 2 ? op.qubit[1] : op.qubit[0];
      g.simple = OBP_SEG_CLIFFORD;
      g.gx[0] = op.lut;
      g.gx[1] = (u64)(op.sign_mask & 0xffffu) | ((u64)op.n_qubits << 16) | ((u64)(qa / 64) << 24) | ((u64)(qa % 64) << 32) |
                ((u64)(qb / 64) << 40) | ((u64)(qb % 64) << 48);
      g.u0r = op.row_scale;
      g.thr_self = g.thr_all = HUGE_VAL;
      frame_update(ctx, op);
      wide_frame_update(sx, sz, op);
      continue;
    }
    const RotP p = build_rotp(ctx, ops[k], atol);
    for (int q = 0; q < p.ngw; ++q) { g.gx[p.gw[q]] = p.gx[q]; g.gz[p.gw[q]] = p.gz[q]; }
    g.hG = p.hG;
    g.u0r = p.u0r; g.u0i = p.u0i; g.u1r = p.u1r; g.u1i = p.u1i;
    g.yG = p.yG;
    g.simple = p.simple;
    // no raw row of a term with |c|^2 above this can fall under the zero filter: the smallest factor is
    // min(|u0|^2, |u1|^2) (|u0||u1| lies between them)
    const double f0 = p.u0r * p.u0r + p.u0i * p.u0i, f1 = p.u1r * p.u1r + p.u1i * p.u1i, fmin = std::min(f0, f1);
    g.thr_self = g.thr_all = atol < 0.0 ? -1.0 : (fmin > 0.0 ? atol * atol * (1.0 + 1e-6) / (fmin * fmin) : HUGE_VAL);
    SegGate gen = g;
    if (has_cliff) {
      WideCol v;
      memset(&v, 0, sizeof(v));
      for (int j = 0; j < ops[k].n_qubits; ++j) {
        const unsigned code = (ops[k].gen_local >> (2 * j)) & 3u;
        if (code & 1u) v ^= sx[(size_t)ops[k].qubit[j]];
        if (code & 2u) v ^= sz[(size_t)ops[k].qubit[j]];
      }
      gen.gx[0] = v.w[0]; gen.gz[0] = v.w[1]; gen.gx[1] = v.w[2]; gen.gz[1] = v.w[3];
    }
    if (has_cliff) gens.push_back(gen);
  }
  };
  if (has_cliff) {
    build_gates();  // (the pulled-back generators come out of the same pass as the frames)
  } else {
    // rotations only: the partition map needs nothing but the generators, so the partition kernels are launched
    // before the host fills in the rest of the gate parameters (the device is idle until the first launch)
    for (int k = 0; k < ng; ++k) {
      SegGate gen;
      memset(&gen, 0, sizeof(gen));
      for (int j = 0; j < ops[k].n_qubits; ++j) {
        const unsigned code = (ops[k].gen_local >> (2 * j)) & 3u;
        const int w = ops[k].qubit[j] / 64;
        const u64 bit = 1ull << (ops[k].qubit[j] % 64);
        if (code & 1u) gen.gx[w] |= bit;
        if (code & 2u) gen.gz[w] |= bit;
      }
      gens.push_back(gen);
    }
  }
  SegP sp;
  memset(&sp, 0, sizeof(sp));
  sp.ng = ng;
  sp.exact_last = (ops[ng - 1].kind == OBP_OP_ROTATION && (ops[ng - 1].flags & OBP_FLAG_EXACT_COUNTERS)) ? 1 : 0;
  sp.slot0 = 0;
  if (const char* e = getenv("OBP_B200_SEG_DEBUG")) sp.dbg = atoi(e);
  sp.n_dense = ctx->n_dense;
  sp.out_cap = ctx->cap;
  sp.atol = atol;
  seg_partition_map(gens, W, sp.fmask);
  // Geometry.  First launch: 128-thread CTAs, four per SM (a quarter of the shared memory each), partitions sized for
  // ~256 rows after the run's growth; second launch: the partitions that outgrew those, one 512-thread CTA per SM.
  SegP sp2 = sp;
  seg_geometry(W, ng, std::min<size_t>((size_t)ctx->seg_smem_max / 4, 55 * 1024), ctx->seg_cap_rows, &sp.cap_rows, &sp.n_index);  // (228 KB per SM, 1 KB reserved per CTA)
  seg_geometry(W, ng, (size_t)ctx->seg_smem_max, 4 * ctx->seg_cap_rows, &sp2.cap_rows, &sp2.n_index);
  const size_t smem1 = seg_smem_bytes(W, sp.cap_rows, sp.n_index, ng), smem2 = seg_smem_bytes(W, sp2.cap_rows, sp2.n_index, ng);
  const int n_ctas1 = 4 * ctx->n_sm, n_ctas2 = ctx->n_sm;
  int P;
  {
    const double growth = std::min(32.0, std::max(1.5, ctx->seg_growth));
    double rows_per_part = std::min(256.0, 0.35 * sp.cap_rows);
    if (const char* e = getenv("OBP_B200_SEG_ROWS_PER_PART")) rows_per_part = std::max(8.0, std::min(atof(e), 0.8 * sp.cap_rows));  // (tuning experiment)
    const double want = (double)ctx->n_live * growth / rows_per_part;
    P = (int)std::min<double>((double)OBP_SEG_MAXP, std::max<double>((double)n_ctas1, std::ceil(want)));
  }
  P = std::max(1, std::min(P, OBP_SEG_MAXP));
  sp.P = sp2.P = P;
  sp.cap_rows_big = sp2.cap_rows_big = sp2.cap_rows;

  Tab t = make_tab(ctx);
  SegOut out;
  out.xz[0] = ctx->alt_xz[0];
  out.xz[1] = ctx->alt_xz[1];
  out.coef = ctx->alt_coef;
  out.hash = ctx->alt_hash;
  out.index = nullptr;
  out.index_mask = 0;
  ctx->seg_T_new = 0;
  if (ctx->index_touched && !inner_piece) {
    // the programs after a run of this observable look keys up: the run's output stage fills the index itself (cleared
    // here, sized to the rows expected; the input table's entries are gone either way once the run commits)
    const double growth = std::min(32.0, std::max(1.5, ctx->seg_growth));
    const u64 want = std::min<u64>(ctx->T, next_pow2(std::max<u64>((u64)(8.0 * growth * (double)ctx->n_live), 1ull << 16)));
    CU(cudaMemsetAsync(ctx->index, 0xFF, want * sizeof(u64), ctx->stream));
    ctx->index_stale = true;  // (until the run reports that every row was published)
    ctx->seg_T_new = want;
    out.index = ctx->index;
    out.index_mask = want - 1;
  }
  unsigned* part = ctx->dest;
  unsigned* perm = (unsigned*)ctx->spare8;
  {
    Bracket br(ctx, OBP_K_SEGMENT, 4);
    ctx->prof.passes[OBP_K_SEGMENT] += (u64)(ng - 1);  // (one pass per gate, in shared memory)
    const int g1 = grid_for(ctx, ctx->n_dense);
    const int gm = std::min(P, n_ctas1);
    if (W == 1) k_seg_count<1><<<g1, OBP_BS, 0, ctx->stream>>>(t, sp, part, d_cnt, d_off, d_cur, d_res);
    else k_seg_count<2><<<g1, OBP_BS, 0, ctx->stream>>>(t, sp, part, d_cnt, d_off, d_cur, d_res);
    k_seg_scatter<<<g1, OBP_BS, 0, ctx->stream>>>(part, d_off, d_cur, perm, ctx->n_dense);
    if (!has_cliff) build_gates();
    // upload the gates through the program staging buffer
    const size_t need = gates.size() * sizeof(SegGate);
    if (need > ctx->prog_alloc) {
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_prog);
      if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
      ctx->d_prog = ctx->h_prog = nullptr;
      ctx->prog_alloc = 0;
      const size_t na = std::max<size_t>(2 * need, 1 << 16);
      CU(cudaMalloc(&ctx->d_prog, na));
      CU(cudaMallocHost(&ctx->h_prog, na));
      ctx->prog_alloc = na;
    }
    memcpy(ctx->h_prog, gates.data(), need);
    CU(cudaMemcpyAsync(ctx->d_prog, ctx->h_prog, need, cudaMemcpyHostToDevice, ctx->stream));
    const SegGate* gs = (const SegGate*)ctx->d_prog;

    const unsigned* none = nullptr;
    SegFin fin0;
    memset(&fin0, 0, sizeof(fin0));
    SegFin fin = fin0;
    fin.st = ctx->d_st;
    fin.cnt = d_cnt;
    fin.n_live_in = ctx->n_live;
    fin.host_res = ctx->h_res_dev;
    fin.n_res_ops = ng + 1;
    if (W == 1) {
      k_seg_main<1, 128, 4><<<gm, 128, smem1, ctx->stream>>>(t, out, sp, gs, d_off, perm, d_res, none, none, d_redo, fin0);
      k_seg_main<1, 512, 1><<<n_ctas2, 512, smem2, ctx->stream>>>(t, out, sp2, gs, d_off, perm, d_res, d_redo, &d_res->n_redo, nullptr, fin);
    } else {
      k_seg_main<2, 128, 4><<<gm, 128, smem1, ctx->stream>>>(t, out, sp, gs, d_off, perm, d_res, none, none, d_redo, fin0);
      k_seg_main<2, 512, 1><<<n_ctas2, 512, smem2, ctx->stream>>>(t, out, sp2, gs, d_off, perm, d_res, d_redo, &d_res->n_redo, nullptr, fin);
    }
    CU(cudaGetLastError());
  }
  PendingOps P_;
  P_.ops.assign(ops, ops + n_ops);
  P_.slot_of_op.resize((size_t)n_ops);
  for (int k = 0; k < n_ops; ++k) P_.slot_of_op[k] = k;
  P_.slot_gates.assign((size_t)n_ops, 1);
  P_.n_ops = n_ops;
  P_.n_entries = 0;
  P_.n_slots = n_ops;
  P_.last_split = -1;
  P_.n_start = ctx->n_live;
  P_.prog_mode = false;
  P_.segment = true;
  P_.frame_hx = std::move(hx0);
  P_.frame_hz = std::move(hz0);
  P_.atol = atol;
  P_.t_host0 = t_host0;
  if (defer) {
    delete ctx->pending;
    ctx->pending = new PendingOps(std::move(P_));
    return OBP_OK;
  }
  return finish_ops(ctx, P_, stats);
}

static int apply_ops_impl(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats,
                          int32_t* n_done, int32_t* peer_xor, bool defer = false) {
  if (!ctx || (n_ops && !ops) || n_ops < 0) return OBP_ERR_BADARG;
  if (n_done) *n_done = n_ops;
  if (peer_xor) *peer_xor = 0;
  CU(cudaSetDevice(ctx->device));
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n_ops == 0) {
    if (stats) { stats->n_out = ctx->n_live; stats->n_dense = ctx->n_dense; }
    return OBP_OK;
  }
  // validate before enqueuing anything
  for (int k = 0; k < n_ops; ++k) {
    const int vrc = validate_op(ctx, ops[k]);
    if (vrc) return vrc;
  }
  if (!n_done) {
    const int span = segment_span(ctx, ops, n_ops);
    if (span > 0) {
      const int piece = std::max(ctx->seg_min_gates, std::min(ctx->seg_max_gates, OBP_SEG_MAXG - 8));
      if (span == n_ops) {
        if (n_ops <= piece) return apply_segment(ctx, ops, n_ops, atol, stats, defer, false);
        return apply_in_pieces(ctx, ops, n_ops, atol, stats, defer, piece);
      }
      // Clifford-class gates with counters of their own follow the last rotation: the run covers the gates before
      // them, they go through the program kernel (its coset count needs the global index) and complete the statistics
      obp_stats st;
      u64 carry_u = ctx->carry_updates;
      std::vector<u64> carry_s = ctx->carry_sizes;
      ctx->carry_updates = 0;
      ctx->carry_sizes.clear();
      const bool keep = ctx->use_segments;
      std::vector<obp_op> head(ops, ops + span);
      const int rc1 = span <= piece ? apply_segment(ctx, head.data(), span, atol, &st, false, false)
                                    : apply_in_pieces(ctx, head.data(), span, atol, &st, false, piece);
      if (rc1) {
        if (stats) { memset(stats, 0, sizeof(*stats)); stats->updates = carry_u + st.updates; stats->n_out = st.n_out; }
        return rc1;
      }
      carry_s.insert(carry_s.end(), ctx->last_op_sizes.begin(), ctx->last_op_sizes.end());
      ctx->carry_updates = carry_u + st.updates;
      ctx->carry_sizes = carry_s;
      ctx->use_segments = false;  // (the tail holds no rotation: it is not a run)
      const int rc2 = apply_ops_impl(ctx, ops + span, n_ops - span, atol, stats, nullptr, nullptr, defer);
      ctx->use_segments = keep;
      return rc2;
    }
  }
  {
    // rows this program can reach: every rotation may double the table
    int n_rot = 0;
    for (int k = 0; k < n_ops; ++k) n_rot += ops[k].kind == OBP_OP_ROTATION ? 1 : 0;
    const u64 reach = n_rot >= 40 ? ctx->cap : std::min<u64>(ctx->cap, std::max<u64>(ctx->n_dense, 1) << n_rot);
    const int irc = ensure_index(ctx, reach);
    if (irc) return irc;
  }
  const auto t_host0 = std::chrono::steady_clock::now();
  // tables of <= 128 qubits: record the program and run it with ONE persistent cooperative launch
  // (large tables keep the stand-alone kernels: a launch is cheap next to a pass over >= 2^18 rows and
  // k_rotate alone holds one more CTA per SM than the program kernel)
  const bool prog_mode = ctx->use_program && ctx->prog_ctas_per_sm > 0 && ctx->W <= 2 && ctx->n_dense < ctx->prog_max_rows && !ctx->flag_sync;
  ProgRec prog;
  int n_rot_ops = 0;
  int last_split = -1;  // index of the last split rotation run with the device-side handshake
  // wider tables: a run of Clifford-class gates becomes one stream of LOAD / gate / STORE micro-ops
  const bool stream_mode = ctx->W > 2;
  std::vector<CliffGate> stream;
  static thread_local CliffStream cs;  // 32 KB: kept off the stack
  int rc = ensure_ops(ctx, 3 * n_ops + 2);
  if (rc) return rc;
  if (!prog_mode) CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat) * (3 * n_ops + 2), ctx->stream));  // (a program zeroes its own slots)

  const u64 n_start = ctx->n_live;
  u64 bound = std::max<u64>(ctx->n_dense, 1);
  std::vector<int> slot_of_op(n_ops, -1);
  std::vector<int> slot_gates;  // gates accounted to each slot
  int n_slots = 0;
  PendingBatch pb;
  pb.reset();
  Tab t = make_tab(ctx);
  const size_t max_smem_slots = 2;  // register variant: a batch addresses at most the table's two words

  int stream_word = -1;  // word currently held in registers by the stream being built
  auto stream_switch = [&](int w) {
    if (stream_word == w) return;
    CliffGate m;
    memset(&m, 0, sizeof(m));
    m.nq = OBP_STREAM_SWITCH;
    m.ba = (unsigned char)w;
    stream.push_back(m);
    stream_word = w;
  };

  auto flush = [&]() -> int {
    if (stream_mode) {
      if (stream.empty()) return OBP_OK;
      stream_switch(OBP_MAX_WORDS);  // write the last word back
      stream_word = -1;
      cs.n = (int)stream.size();
      cs.slot = n_slots;
      cs.scale = pb.b.scale;
      cs.atol = atol;
      cs.epoch = ++ctx->epoch;
      memcpy(cs.g, stream.data(), stream.size() * sizeof(CliffGate));
      {
        Bracket br(ctx, OBP_K_CLIFFORD);
        k_clifford_stream<<<grid_for(ctx, bound), OBP_BS, 0, ctx->stream>>>(t, cs);
        CU(cudaGetLastError());
      }
      stream.clear();
      slot_gates.push_back(pb.n_ops);
      n_slots++;
      pb.reset();
      return OBP_OK;
    }
    if (pb.b.ng == 0) return OBP_OK;
    pb.b.atol = atol;
    pb.b.epoch = ++ctx->epoch;
    pb.b.slot = n_slots;
    bool identity_slots = ctx->W <= 2;
    if (identity_slots) {
      // register variant addresses words directly: remap slots to word ids
      for (int k = 0; k < pb.b.ng; ++k) {
        pb.b.g[k].sa = (unsigned char)pb.b.word_of_slot[pb.b.g[k].sa];
        pb.b.g[k].sb = (unsigned char)pb.b.word_of_slot[pb.b.g[k].sb];
      }
    }
    if (prog_mode) {
      prog.add(PROG_CLIFFORD, n_slots, &pb.b, offsetof(CliffBatch, g) + (size_t)pb.b.ng * sizeof(CliffGate), OBP_K_CLIFFORD);
      slot_gates.push_back(pb.n_ops);
      n_slots++;
      pb.reset();
      return OBP_OK;
    }
    Bracket br(ctx, OBP_K_CLIFFORD);
    const int g = grid_for(ctx, bound);
    if (identity_slots) {
      if (ctx->W == 1) k_clifford_reg<1><<<g, OBP_BS, 0, ctx->stream>>>(t, pb.b);
      else k_clifford_reg<2><<<g, OBP_BS, 0, ctx->stream>>>(t, pb.b);
    }
    CU(cudaGetLastError());
    slot_gates.push_back(pb.n_ops);
    n_slots++;
    pb.reset();
    return OBP_OK;
  };

  // one-qubit Cliffords held back per qubit (see FusedGate)
  std::vector<FusedGate> pend(ctx->nq);
  int n_pending = 0;  // qubits with a held-back gate

  // put one (possibly fused) Clifford gate into the current batch
  auto emit_gate = [&](const FusedGate& fg) -> int {
    if (stream_mode) {
      if (stream.size() + 8 > OBP_STREAM_MAX) {
        const int frc = flush();
        if (frc) return frc;
      }
      CliffGate g;
      memset(&g, 0, sizeof(g));
      g.lut = fg.lut;
      g.sign = (unsigned short)fg.sign;
      g.nq = (unsigned char)fg.nq;
      const int wa = fg.q[0] / 64, wb = fg.q[1] / 64;
      const int ba = fg.q[0] % 64, bb = fg.q[1] % 64;
      g.ba = (unsigned char)ba;
      g.bb = (unsigned char)(fg.nq == 2 ? bb : ba);
      if (fg.nq == 2 && wa != wb) {
        // the other word is touched in place; stay on (or move to) the word that is already loaded
        if (stream_word == wb) {  // swap the roles of the two qubits so that qubit a is in the current word
          u64 lut = 0;
          unsigned sign = 0;
          for (unsigned v = 0; v < 16; ++v) {
            const unsigned vs = ((v & 3u) << 2) | (v >> 2);  // code with a and b exchanged
            const unsigned img = lut_at(fg.lut, vs);
            lut |= (u64)(((img & 3u) << 2) | (img >> 2)) << (4 * v);
            sign |= ((fg.sign >> vs) & 1u) << v;
          }
          g.lut = lut;
          g.sign = (unsigned short)sign;
          g.ba = (unsigned char)bb;
          g.bb = (unsigned char)ba;
          g.sb = (unsigned char)wa;
        } else {
          stream_switch(wa);
          g.sb = (unsigned char)wb;
        }
        g.nq = OBP_STREAM_CROSS;
      } else {
        stream_switch(wa);
        if (fg.nq == 2 && (bb == ba + 1 || ba == bb + 1)) {
          // adjacent bits: pair layout, low bit first: code' = x_lo | x_hi<<1 | z_lo<<2 | z_hi<<3
          const bool a_low = ba < bb;
          auto to_std = [&](unsigned vp) {  // pair code -> x_a | z_a<<1 | x_b<<2 | z_b<<3
            const unsigned xlo = vp & 1u, xhi = (vp >> 1) & 1u, zlo = (vp >> 2) & 1u, zhi = (vp >> 3) & 1u;
            const unsigned xa = a_low ? xlo : xhi, za = a_low ? zlo : zhi, xb = a_low ? xhi : xlo, zb = a_low ? zhi : zlo;
            return xa | (za << 1) | (xb << 2) | (zb << 3);
          };
          auto to_pair = [&](unsigned vs) {
            const unsigned xa = vs & 1u, za = (vs >> 1) & 1u, xb = (vs >> 2) & 1u, zb = (vs >> 3) & 1u;
            const unsigned xlo = a_low ? xa : xb, xhi = a_low ? xb : xa, zlo = a_low ? za : zb, zhi = a_low ? zb : za;
            return xlo | (xhi << 1) | (zlo << 2) | (zhi << 3);
          };
          u64 lut = 0;
          unsigned sign = 0;
          for (unsigned vp = 0; vp < 16; ++vp) {
            const unsigned vs = to_std(vp);
            lut |= (u64)to_pair(lut_at(fg.lut, vs)) << (4 * vp);
            sign |= ((fg.sign >> vs) & 1u) << vp;
          }
          g.lut = lut;
          g.sign = (unsigned short)sign;
          g.ba = (unsigned char)std::min(ba, bb);
          g.nq = OBP_STREAM_PAIR;
        }
      }
      stream.push_back(g);
      pb.b.scale = std::min(pb.b.scale, fg.scale);
      pb.n_ops += fg.n_gates;
      obp_op fop;
      memset(&fop, 0, sizeof(fop));
      fop.n_qubits = fg.nq;
      fop.qubit[0] = fg.q[0];
      fop.qubit[1] = fg.q[1];
      fop.lut = fg.lut;
      frame_update(ctx, fop);
      return OBP_OK;
    }
    int need_new = 0;
    for (int j = 0; j < fg.nq; ++j) {
      bool have = false;
      for (int sl = 0; sl < pb.b.nslots; ++sl) have |= pb.b.word_of_slot[sl] == fg.q[j] / 64;
      need_new += !have;
    }
    if (pb.b.ng == OBP_CLIFF_BATCH || pb.b.nslots + need_new > (int)max_smem_slots) {
      const int frc = flush();
      if (frc) return frc;
    }
    CliffGate& g = pb.b.g[pb.b.ng++];
    memset(&g, 0, sizeof(g));
    g.lut = fg.lut;
    g.sign = (unsigned short)fg.sign;
    g.nq = (unsigned char)fg.nq;
    g.sa = (unsigned char)pb.slot_of_word(fg.q[0] / 64);
    g.ba = (unsigned char)(fg.q[0] % 64);
    if (fg.nq == 2) {
      g.sb = (unsigned char)pb.slot_of_word(fg.q[1] / 64);
      g.bb = (unsigned char)(fg.q[1] % 64);
    } else {
      g.sb = g.sa;
      g.bb = g.ba;
    }
    pb.b.scale = std::min(pb.b.scale, fg.scale);
    pb.n_ops += fg.n_gates;
    // the hash frame follows the keys: re-express it when the gate is really applied
    obp_op fop;
    memset(&fop, 0, sizeof(fop));
    fop.n_qubits = fg.nq;
    fop.qubit[0] = fg.q[0];
    fop.qubit[1] = fg.q[1];
    fop.lut = fg.lut;
    frame_update(ctx, fop);
    return OBP_OK;
  };
  auto release_qubit = [&](int q) -> int {  // apply the held-back gate of qubit q now
    if (pend[q].n_gates == 0) return OBP_OK;
    const FusedGate fg = pend[q];
    pend[q] = FusedGate();
    n_pending--;
    return emit_gate(fg);
  };
  auto release_all = [&]() -> int {  // descending qubit order: the order gates of a layer arrive in
    for (int q = ctx->nq - 1; q >= 0 && n_pending > 0; --q) {
      const int r = release_qubit(q);
      if (r) return r;
    }
    return OBP_OK;
  };

  for (int k = 0; k < n_ops; ++k) {
    const obp_op& op = ops[k];
    if (op.kind == OBP_OP_CLIFFORD) {
      const bool exact = (op.flags & OBP_FLAG_EXACT_COUNTERS) != 0;
      slot_of_op[k] = n_slots;
      if (!exact && (op.flags & OBP_FLAG_OWN_SLOT)) {
        // per-gate size log: this gate gets a pass (and a counter slot) of its own
        rc = release_all();
        if (rc) return rc;
        rc = flush();
        if (rc) return rc;
        slot_of_op[k] = n_slots;
        rc = emit_gate(fused_from_op(op));
        if (rc) return rc;
        rc = flush();
        if (rc) return rc;
        continue;
      }
      if (!exact && op.n_qubits == 1) {
        // hold it back: it commutes with everything that does not touch its qubit
        const int q = op.qubit[0];
        const FusedGate g = fused_from_op(op);
        if (pend[q].n_gates == 0) {
          pend[q] = g;
          n_pending++;
        } else {
          pend[q] = fuse_seq(pend[q], g);
        }
        continue;
      }
      if (!exact) {
        // two-qubit gate: absorb what is pending on its qubits
        const int qa = op.qubit[0], qb = op.qubit[1];
        FusedGate g = fused_from_op(op);
        if (pend[qb].n_gates) {
          g = fuse_seq(embed_1q(pend[qb], 1, qa, qb), g);
          pend[qb] = FusedGate();
          n_pending--;
        }
        if (pend[qa].n_gates) {
          g = fuse_seq(embed_1q(pend[qa], 0, qa, qb), g);
          pend[qa] = FusedGate();
          n_pending--;
        }
        rc = emit_gate(g);
        if (rc) return rc;
        continue;
      }
      // the slice's last gate with the reference's counters: everything held back is applied first,
      // so that this gate's batch is the last slot of the program
      rc = release_all();
      if (rc) return rc;
      rc = flush();
      if (rc) return rc;
      slot_of_op[k] = n_slots;
      {
        // counters come from the keys before the gate acts
        CosetP cp;
        memset(&cp, 0, sizeof(cp));
        cp.nd = (int)op.group_size - 1;
        cp.wa = op.qubit[0] / 64;
        cp.ba = op.qubit[0] % 64;
        cp.wb = op.n_qubits == 2 ? op.qubit[1] / 64 : -1;
        cp.bb = op.n_qubits == 2 ? op.qubit[1] % 64 : 0;
        for (int d = 0; d < cp.nd; ++d) {
          const unsigned code = (unsigned)((op.group_elems >> (4 * d)) & 15ull);
          cp.code[d] = code;
          u64 h = 0;
          if (code & 1u) h ^= ctx->hx[op.qubit[0]];
          if (code & 2u) h ^= ctx->hz[op.qubit[0]];
          if (op.n_qubits == 2) {
            if (code & 4u) h ^= ctx->hx[op.qubit[1]];
            if (code & 8u) h ^= ctx->hz[op.qubit[1]];
          }
          cp.hD[d] = h;
        }
        cp.scale = op.row_scale;
        cp.atol = atol;
        cp.slot = n_slots;  // the slot the gate's own batch will use
        if (prog_mode) {
          prog.add(PROG_COSET, n_slots, &cp, sizeof(cp), OBP_K_OTHER);
        } else {
          Bracket br(ctx, OBP_K_OTHER);
          k_coset_count<<<grid_for(ctx, bound), OBP_BS, 0, ctx->stream>>>(t, cp);
          CU(cudaGetLastError());
        }
      }
      rc = emit_gate(fused_from_op(op));
      if (rc) return rc;
      rc = flush();
      if (rc) return rc;
    } else {
      // A rotation changes coefficients, and the raw-row zero filter of a Clifford-class gate
      // (|c| * row_scale <= atol) must see the coefficients of ITS place in the program: nothing is
      // held back across a rotation, on whatever qubit.
      rc = release_all();
      if (rc) return rc;
      rc = flush();
      if (rc) return rc;
      RotP p = build_rotp(ctx, op, atol);
      if (ctx->n_ranks > 1 && (p.hG >> ctx->owner_shift) != 0 && ctx->flag_sync) {
        // Split rotation without leaving the device: emit into the peer's inbox, merge what the peer
        // emitted into ours, finish — the three kernels are ordered across GPUs by flags in the inbox
        // headers (cta_wait_flag / finish_kernel_flag), the host does not synchronise.
        const int peer_rank = ctx->rank ^ (int)(p.hG >> ctx->owner_shift);
        const u64 seq = ++ctx->xseq;
        ctx->n_split_rot++;
        const int b = (int)(seq & 1ull);
        if (!ctx->inbox[b] || !ctx->peer_inbox[b][peer_rank]) return ctx->fail(OBP_ERR_BADARG, "sharded table: peer inbox not attached");
        u64* peer = ctx->peer_inbox[b][peer_rank];
        u64* mine = ctx->inbox[b];
        p.epoch = ++ctx->epoch;
        p.slot = n_slots;
        CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
        {
          Bracket br(ctx, OBP_K_ROTATE, 3);
          k_rotate_emit<<<grid_for(ctx, bound, ctx->rot_ctas_per_sm), OBP_BS, 0, ctx->stream>>>(
              t, p, peer + OBP_INBOX_HDR, ctx->inbox_cap, peer + OBP_HDR_CURSOR, ctx->d_cnt,
              peer + OBP_HDR_MERGED, seq >= 2 ? seq - 2 : 0, peer + OBP_HDR_EMITTED, seq);
          k_merge_records<<<grid_for(ctx, std::min<u64>(ctx->inbox_cap, std::max<u64>(2 * bound, 1 << 14))), OBP_BS, 0,
                            ctx->stream>>>(t, mine + OBP_INBOX_HDR, 0, mine + OBP_HDR_CURSOR, ctx->inbox_cap, ctx->d_cnt,
                                           mine + OBP_HDR_EMITTED, seq);
          bound = std::min<u64>(ctx->cap, bound * 2);
          k_rotate_finish<<<grid_for(ctx, bound), OBP_BS, 0, ctx->stream>>>(t, p, ctx->d_cnt, mine, seq);
          CU(cudaGetLastError());
        }
        slot_of_op[k] = n_slots;
        slot_gates.push_back(1);
        n_slots++;
        n_rot_ops++;
        last_split = k;
        continue;
      }
      if (ctx->n_ranks > 1 && (p.hG >> ctx->owner_shift) != 0) {
        // partner keys of this rotation belong to rank ^ d: the caller runs the split rotation
        if (!n_done || !peer_xor) return ctx->fail(OBP_ERR_BADARG, "sharded table: use obp_apply_ops_sharded");
        *n_done = k;
        *peer_xor = (int32_t)(p.hG >> ctx->owner_shift);
        n_ops = k;
        break;  // (gates still held back are applied by the release_all() after the loop)
      }
      p.epoch = ++ctx->epoch;
      p.slot = n_slots;
      n_rot_ops++;
      if (prog_mode) {
        prog.add(PROG_ROTATE, n_slots, &p, sizeof(p), OBP_K_ROTATE);
      } else {
        Bracket br(ctx, OBP_K_ROTATE);
        k_rotate<<<grid_for(ctx, bound, ctx->rot_ctas_per_sm), OBP_BS, 0, ctx->stream>>>(t, p);  // one wave
        CU(cudaGetLastError());
      }
      slot_of_op[k] = n_slots;
      slot_gates.push_back(1);
      n_slots++;
      bound = std::min<u64>(ctx->cap, bound * 2);
    }
  }
  rc = release_all();
  if (rc) return rc;
  rc = flush();
  if (rc) return rc;
  if (n_ops == 0) {  // a sharded program that starts with an exchange rotation
    if (stats) { stats->n_in = stats->n_out = ctx->n_live; stats->n_dense = ctx->n_dense; }
    return OBP_OK;
  }

  const int n_entries = (int)prog.entries.size();
  if (prog_mode && n_entries) {
    // upload entries + parameter pool, then one cooperative launch sized to the rows the program can reach
    const size_t ent_bytes = (size_t)n_entries * sizeof(ProgEntry), need = ent_bytes + prog.pool.size();
    if (need > ctx->prog_alloc) {
      CU(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_prog);
      if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
      ctx->d_prog = ctx->h_prog = nullptr;
      ctx->prog_alloc = 0;
      const size_t na = std::max<size_t>(2 * need, 1 << 16);
      CU(cudaMalloc(&ctx->d_prog, na));
      CU(cudaMallocHost(&ctx->h_prog, na));
      ctx->prog_alloc = na;
    }
    memcpy(ctx->h_prog, prog.entries.data(), ent_bytes);
    memcpy(ctx->h_prog + ent_bytes, prog.pool.data(), prog.pool.size());
    CU(cudaMemcpyAsync(ctx->d_prog, ctx->h_prog, need, cudaMemcpyHostToDevice, ctx->stream));
    const u64 reach = std::min<u64>(ctx->cap, std::max<u64>(ctx->n_dense, 1) << std::min(n_rot_ops, ctx->prog_reach_shift));
    const int max_g = ctx->prog_ctas_per_sm * ctx->n_sm;
    int g = (int)std::min<u64>((reach + OBP_BS - 1) / OBP_BS, (u64)max_g);
    g = std::max(g, std::min(ctx->prog_min_grid, max_g));
    const ProgEntry* d_entries = (const ProgEntry*)ctx->d_prog;
    const unsigned char* d_pool = ctx->d_prog + ent_bytes;
    int ne = n_entries, ns = n_slots;
    int pflags = (ctx->prof_on ? 1 : 0) | ctx->prog_debug;
    u64* host_res = ctx->h_res_dev;
    int n_res_ops = n_slots + (ctx->prof_on ? n_entries : 0);
    void* args[] = {(void*)&t, (void*)&d_entries, (void*)&ne, (void*)&d_pool, (void*)&ns, (void*)&ctx->d_tbegin, (void*)&pflags, (void*)&host_res, (void*)&n_res_ops};
    ctx->prof.launches[OBP_K_PROGRAM]++;
    for (int k = 0; k < n_entries; ++k) ctx->prof.passes[prog.klass[k]]++;
    if (ctx->W == 1) {
      CU(cudaLaunchCooperativeKernel((void*)k_program<1>, dim3(g), dim3(OBP_BS), args, 0, ctx->stream));
    } else {
      CU(cudaLaunchCooperativeKernel((void*)k_program<2>, dim3(g), dim3(OBP_BS), args, 0, ctx->stream));
    }
  }
  if (last_split >= 0) CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  PendingOps P;
  P.ops.assign(ops, ops + n_ops);
  P.prog = std::move(prog);
  P.slot_of_op = std::move(slot_of_op);
  P.slot_gates = std::move(slot_gates);
  P.n_ops = n_ops;
  P.n_entries = n_entries;
  P.n_slots = n_slots;
  P.last_split = last_split;
  P.n_start = n_start;
  P.prog_mode = prog_mode;
  P.t_host0 = t_host0;
  if (defer) {  // obp_apply_ops_submit: the caller overlaps its own work with the device, then calls obp_apply_ops_wait
    delete ctx->pending;
    ctx->pending = new PendingOps(std::move(P));
    return OBP_OK;
  }
  return finish_ops(ctx, P, stats);
}

static const char* kRowOrderMsg =
    "obp_apply_ops: in-place gate kernels do not keep the reference's row order; use obp_apply_generic in row-order mode";
int obp_apply_ops(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats) {
  OBP_NO_PENDING(ctx, "obp_apply_ops");
  if (ctx && ctx->row_order) return ctx->fail(OBP_ERR_UNSUPPORTED, kRowOrderMsg);
  if (ctx) { ctx->carry_updates = 0; ctx->carry_sizes.clear(); }
  return apply_ops_impl(ctx, ops, n_ops, atol, stats, nullptr, nullptr);
}

int obp_apply_ops_submit(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol) {
  OBP_NO_PENDING(ctx, "obp_apply_ops_submit");
  if (ctx && ctx->row_order) return ctx->fail(OBP_ERR_UNSUPPORTED, kRowOrderMsg);
  if (ctx && ctx->n_ranks > 1) return ctx->fail(OBP_ERR_UNSUPPORTED, "obp_apply_ops_submit: not available on a sharded table");
  if (!ctx) return OBP_ERR_BADARG;
  ctx->carry_updates = 0;
  ctx->carry_sizes.clear();
  return apply_ops_impl(ctx, ops, n_ops, atol, nullptr, nullptr, nullptr, true);
}

int obp_apply_ops_wait(obp_ctx* ctx, obp_stats* stats) {
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->pending) {  // nothing was enqueued (an empty program)
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->n_in = stats->n_out = ctx->n_live;
      stats->n_dense = ctx->n_dense;
    }
    return OBP_OK;
  }
  PendingOps* P = ctx->pending;
  ctx->pending = nullptr;
  const int rc = finish_ops(ctx, *P, stats);
  delete P;
  return rc;
}

int obp_last_op_sizes(obp_ctx* ctx, uint64_t* sizes, int32_t max_ops, int32_t* n_ops) {
  if (!ctx || !n_ops) return OBP_ERR_BADARG;
  *n_ops = (int32_t)ctx->last_op_sizes.size();
  if (sizes)
    for (int k = 0; k < *n_ops && k < max_ops; ++k) sizes[k] = ctx->last_op_sizes[k];
  return OBP_OK;
}

int obp_last_op_stats(obp_ctx* ctx, int32_t op_index, obp_stats* out) {
  if (!ctx || !out) return OBP_ERR_BADARG;
  if (ctx->last_ops.empty() || op_index < 0 || op_index >= (int32_t)ctx->last_ops.size())
    return ctx->fail(OBP_ERR_BADARG, "obp_last_op_stats: no such op in the last gate program");
  const obp_op& op = ctx->last_ops[(size_t)op_index];
  const int s = ctx->last_slot_of_op[(size_t)op_index];
  if (!(op.flags & OBP_FLAG_EXACT_COUNTERS) || s < 0 || s >= ctx->last_n_slots)
    return ctx->fail(OBP_ERR_BADARG, "obp_last_op_stats: the op did not ask for exact counters");
  memset(out, 0, sizeof(*out));
  const OpStat& o = ctx->h_ops[s];
  const u64 n_in = s == 0 ? ctx->last_n_start : ctx->h_ops[s - 1].n_out;
  const u64 k2 = (u64)op.n_components * op.n_components;
  out->n_in = n_in;
  out->raw_rows = n_in * k2;
  out->n_out = o.n_out;
  if (op.kind == OBP_OP_ROTATION) {
    out->num_unique = o.keys_present;
    out->num_duplicate = o.rows_surv - o.keys_present;
    out->num_trimmed = (out->raw_rows - o.rows_surv) + o.cancelled;
    out->sum_trimmed[0] = o.sum_re;
    out->sum_trimmed[1] = o.sum_im;
  } else {
    u64 cosets = 0;
    for (unsigned c = 1; c <= 16; ++c) cosets += o.ncnt[c] / c;
    const u64 rows = o.n_surv * k2;
    const u64 uniq = (u64)op.group_size * cosets;
    out->num_unique = uniq;
    out->num_duplicate = rows - uniq;
    out->num_trimmed = (out->raw_rows - rows) + (uniq - o.n_surv);
  }
  return OBP_OK;
}

int obp_apply_ops_sharded(obp_ctx* ctx, const obp_op* ops, int32_t n_ops, double atol, obp_stats* stats,
                          int32_t* n_done, int32_t* peer_xor) {
  if (!n_done || !peer_xor) return OBP_ERR_BADARG;
  OBP_NO_PENDING(ctx, "obp_apply_ops_sharded");
  return apply_ops_impl(ctx, ops, n_ops, atol, stats, n_done, peer_xor);
}

// -------------------------------------------------------------------------------------------------
// sharded table: configuration, the split rotation, stepwise truncation
int obp_set_shard(obp_ctx* ctx, int32_t n_ranks, int32_t rank) {
  if (!ctx || n_ranks < 1 || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks || n_ranks > 1024)
    return OBP_ERR_BADARG;
  int bits = 0;
  while ((1 << bits) < n_ranks) ++bits;
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  ctx->owner_shift = 64 - bits;  // 64 for a single rank: never used as a shift then
  return OBP_OK;
}

// Late sharding: a table that so far held the WHOLE operator (every rank an identical copy) keeps only the
// rows this rank owns from now on.  `future_ops` (optional) are the gates the run is going to apply next, in
// order: the hash frame is re-drawn so that the owner bits of every rotation generator among them vanish
// (plan_owner_planes above) — those rotations then never need a record exchange — and every row is re-hashed
// under the new frame before the foreign ones are dropped.  *n_covered = rotations the plan covers.
int obp_shard_split(obp_ctx* ctx, int32_t n_ranks, int32_t rank, const obp_op* future_ops, int32_t n_future,
                    uint64_t* n_out, int32_t* n_covered) {
  OBP_NO_PENDING(ctx, "obp_shard_split");
  if (!ctx) return OBP_ERR_BADARG;
  if (ctx->n_ranks != 1) return ctx->fail(OBP_ERR_BADARG, "obp_shard_split: the table is already sharded");
  int rc = obp_set_shard(ctx, n_ranks, rank);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  if (n_covered) *n_covered = 0;
  int bits = 0;
  while ((1 << bits) < n_ranks) ++bits;
  if (future_ops && n_future > 0 && bits > 0) {
    for (int k = 0; k < n_future; ++k) {
      const int vrc = validate_op(ctx, future_ops[k]);
      if (vrc) return vrc;
    }
    if (bits > OBP_MAX_OWNER_BITS) return ctx->fail(OBP_ERR_UNSUPPORTED, "obp_shard_split: planning supports up to 64 ranks");
    // The plan must also spread the rows that exist now.  In a circuit whose future generators span the same space
    // as the past ones (periodic Trotter layers) a function that is blind to every future generator is blind to the
    // differences between the existing rows too: judge each candidate on the owner histogram of the current table
    // and halve the number of rotations it covers until no rank would get more than 1.5 x its share (a plan that
    // covers nothing is the plain hash shard function).
    std::vector<std::vector<unsigned char>> planes;
    int n_rot_future = 0;
    for (int k = 0; k < n_future; ++k) n_rot_future += future_ops[k].kind == OBP_OP_ROTATION;
    int covered = 0;
    u64* d_counts = ctx->d_cols;  // scratch: 2 * 64 * W u64, at least 128 entries
    std::vector<u64> counts((size_t)1 << bits);
    for (int target = n_rot_future;; target /= 2) {
      covered = plan_owner_planes(ctx->nq, future_ops, n_future, bits, target, planes);
      if (ctx->n_live == 0 || target == 0) break;
      OwnerPlanes pl;
      memset(&pl, 0, sizeof(pl));
      pl.bits = bits;
      for (int i = 0; i < bits; ++i)
        for (int q = 0; q < ctx->nq; ++q) {
          if (planes[i][2 * q]) pl.mask[i][q / 64].x |= 1ull << (q % 64);
          if (planes[i][2 * q + 1]) pl.mask[i][q / 64].y |= 1ull << (q % 64);
        }
      CU(cudaMemsetAsync(d_counts, 0, counts.size() * sizeof(u64), ctx->stream));
      k_owner_hist<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), pl, d_counts);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      const u64 mx = *std::max_element(counts.begin(), counts.end());
      const double share = (double)ctx->n_live / (double)counts.size();
      if ((double)mx <= 1.5 * share + 4.0 * std::sqrt(share) + 8.0) break;  // (small tables are allowed their scatter)
    }
    ctx->cols_dirty = true;  // d_cols served as scratch
    if (n_covered) *n_covered = covered;
    // a fresh frame: random low bits, planned owner bits on top (bit 63 - i = owner bit i counted from the top)
    // (the seed must be the same on every rank: records carry hashes across ranks — only quantities the replicated
    // copies share may enter it, never a context's own history such as its epoch counter)
    u64 seed = 0x0B9B200C0FFEEull ^ (0x9E3779B97F4A7C15ull * (ctx->n_live + 1));
    for (int q = 0; q < ctx->nq; ++q) {
      u64 cx = splitmix64(seed) >> bits, cz = splitmix64(seed) >> bits;
      for (int i = 0; i < bits; ++i) {
        cx |= (u64)planes[i][2 * q] << (63 - i);
        cz |= (u64)planes[i][2 * q + 1] << (63 - i);
      }
      ctx->hx[q] = cx;
      ctx->hz[q] = cz;
    }
    ctx->cols_dirty = true;
    rc = upload_cols(ctx);
    if (rc) return rc;
    if (ctx->n_dense) {
      k_hash_rows<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), ctx->d_cols, ctx->n_dense);
      CU(cudaGetLastError());
    }
  }
  if (ctx->n_dense && ctx->n_ranks > 1) {
    k_shard_drop<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), ctx->owner_shift, (u64)ctx->rank, ++ctx->epoch);
    CU(cudaGetLastError());
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  // rows moved or were re-hashed: dense again, index rebuilt under the current hashes
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
  } else {
    rc = rebuild_index(ctx);
  }
  if (rc) return rc;
  rc = sync_state(ctx);
  if (rc) return rc;
  ctx->snap.valid = false;  // the last commit held the whole operator: the caller commits again
  if (n_out) *n_out = ctx->n_live;
  return OBP_OK;
}

int obp_set_shard_sync(obp_ctx* ctx, int32_t device_flags) {
  if (!ctx) return OBP_ERR_BADARG;
  ctx->flag_sync = device_flags != 0;
  return OBP_OK;
}

int obp_shard_stats(obp_ctx* ctx, uint64_t* n_exchange_rotations, uint64_t* n_records_emitted) {
  if (!ctx) return OBP_ERR_BADARG;
  if (n_exchange_rotations) *n_exchange_rotations = ctx->n_split_rot;
  if (n_records_emitted) *n_records_emitted = ctx->h_st->emitted_total;  // as of the last synchronisation
  return OBP_OK;
}

int obp_last_split_counters(obp_ctx* ctx, obp_shard_counters* out) {
  if (!ctx || !out) return OBP_ERR_BADARG;
  memset(out, 0, sizeof(*out));
  if (!ctx->last_split_is_last_op) return OBP_OK;  // n_records stays 0, n_in 0: "no split rotation ended the program"
  const ShardCnt& c = *ctx->h_cnt;
  out->n_out = ctx->n_live;
  out->n_dense = ctx->n_dense;
  out->rows_surv = c.rows_surv;
  out->keys_present = c.keys_present;
  out->cancelled = c.cancelled;
  out->n_records = c.n_records;
  out->sum_trimmed[0] = c.sum_re;
  out->sum_trimmed[1] = c.sum_im;
  out->n_in = 1;  // marker: valid
  return OBP_OK;
}

uint64_t obp_record_bytes(int n_qubits) {
  const u64 W = (u64)std::max(1, (n_qubits + 63) / 64);
  return (2 * W + 3) * sizeof(u64);
}

static void fill_shard_stats(obp_ctx* ctx, obp_shard_counters* out, u64 n_in) {
  if (!out) return;
  const ShardCnt& c = *ctx->h_cnt;
  out->n_in = n_in;
  out->n_out = ctx->n_live;
  out->n_dense = ctx->n_dense;
  out->rows_surv = c.rows_surv;
  out->keys_present = c.keys_present;
  out->cancelled = c.cancelled;
  out->n_records = c.n_records;
  out->sum_trimmed[0] = c.sum_re;
  out->sum_trimmed[1] = c.sum_im;
}

int obp_rotate_emit(obp_ctx* ctx, const obp_op* op, double atol, void* d_records, uint64_t cap_records,
                    obp_shard_counters* out) {
  if (!ctx || !op || op->kind != OBP_OP_ROTATION || (cap_records && !d_records)) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  int rc = validate_op(ctx, *op);
  if (rc) return rc;
  RotP p = build_rotp(ctx, *op, atol);
  p.epoch = ++ctx->epoch;
  p.slot = -1;
  ctx->shard_rot = p;
  ctx->shard_rot_open = true;
  const u64 n_in = ctx->n_live;
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  {
    Bracket br(ctx, OBP_K_ROTATE);
    k_rotate_emit<<<grid_for(ctx, std::max<u64>(ctx->n_dense, 1), ctx->rot_ctas_per_sm), OBP_BS, 0, ctx->stream>>>(
        make_tab(ctx), p, (u64*)d_records, cap_records, &ctx->d_cnt->n_records, ctx->d_cnt, nullptr, 0, nullptr, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  ctx->prof.term_updates[OBP_K_ROTATE] += n_in;
  fill_shard_stats(ctx, out, n_in);
  if (rc == OBP_OK && ctx->h_cnt->n_records > cap_records)
    return ctx->fail(OBP_ERR_CAPACITY, "obp_rotate_emit: record buffer too small");
  return rc;
}

int obp_rotate_merge(obp_ctx* ctx, const void* d_records, uint64_t n_records, obp_shard_counters* out) {
  if (!ctx || (n_records && !d_records)) return OBP_ERR_BADARG;
  if (!ctx->shard_rot_open) return ctx->fail(OBP_ERR_BADARG, "obp_rotate_merge: no rotation in flight");
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  ctx->shard_rot_open = false;
  const u64 n_in = ctx->n_live;
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_ROTATE);
    if (n_records)
      k_merge_records<<<grid_for(ctx, n_records), OBP_BS, 0, ctx->stream>>>(t, (const u64*)d_records, n_records, nullptr, 0,
                                                                          ctx->d_cnt, nullptr, 0);
    const u64 bound = std::min<u64>(ctx->cap, std::max<u64>(ctx->n_dense + n_records, 1));
    k_rotate_finish<<<grid_for(ctx, bound), OBP_BS, 0, ctx->stream>>>(t, ctx->shard_rot, ctx->d_cnt, nullptr, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  int rc = sync_state(ctx);
  fill_shard_stats(ctx, out, n_in);
  if (rc) return rc;
  if (ctx->n_dense - ctx->n_live > std::max<u64>(1024, ctx->n_live / 4)) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
    if (out) out->n_dense = ctx->n_dense;
  }
  return OBP_OK;
}

// ---- reset on a sharded table (staged exchange) -------------------------------------------------------
// *peer_xor = owner bits of the hash column of Z on that qubit: 0 -> obp_apply_reset does everything
// locally; else phase 1 (filter + emit the rows that must move) here, phase 2 in obp_reset_merge.
int obp_reset_emit(obp_ctx* ctx, int32_t qubit, double atol, void* d_records, uint64_t cap_records, int32_t* peer_xor,
                   obp_shard_counters* out) {
  if (!ctx || qubit < 0 || qubit >= ctx->nq || !peer_xor || (cap_records && !d_records)) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  *peer_xor = ctx->n_ranks > 1 ? (int32_t)(ctx->hz[qubit] >> ctx->owner_shift) : 0;
  if (*peer_xor == 0) return OBP_OK;
  int rc = ensure_ops(ctx, 2);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  const u64 n_in = ctx->n_live;
  ResetP p;
  memset(&p, 0, sizeof(p));
  p.w = qubit / 64;
  p.bit = 1ull << (qubit % 64);
  p.hZ = ctx->hz[qubit];
  p.atol = atol;
  p.slot = 0;
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_OTHER, 2);
    const int g = grid_for(ctx, std::max<u64>(ctx->n_dense, 1));
    p.epoch = ++ctx->epoch;
    k_reset_filter<<<g, OBP_BS, 0, ctx->stream>>>(t, p);
    p.epoch = ++ctx->epoch;
    k_reset_emit<<<g, OBP_BS, 0, ctx->stream>>>(t, p, (u64*)d_records, cap_records, ctx->d_cnt);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  ctx->prof.term_updates[OBP_K_OTHER] += n_in;
  fill_shard_stats(ctx, out, n_in);
  if (out) out->rows_surv = ctx->h_ops[0].rows_surv;  // rows that passed the raw-row filter (k_reset_filter)
  return rc;
}

int obp_reset_merge(obp_ctx* ctx, double atol, const void* d_records, uint64_t n_records, obp_shard_counters* out) {
  if (!ctx || (n_records && !d_records)) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  const u64 n_in = ctx->n_live;
  int rc = ensure_ops(ctx, 2);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat), ctx->stream));
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_OTHER, 2);
    if (n_records)
      k_merge_records<<<grid_for(ctx, n_records), OBP_BS, 0, ctx->stream>>>(t, (const u64*)d_records, n_records, nullptr, 0,
                                                                          ctx->d_cnt, nullptr, 0);
    // second zero filter of simplify(): only merged sums can have become small
    k_trim<<<grid_for(ctx, std::min<u64>(ctx->cap, std::max<u64>(ctx->n_dense + n_records, 1))), OBP_BS, 0, ctx->stream>>>(
        t, atol, ++ctx->epoch, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  fill_shard_stats(ctx, out, n_in);
  if (out) {
    out->cancelled = ctx->h_ops[0].cancelled;
    out->sum_trimmed[0] = ctx->h_ops[0].sum_re;
    out->sum_trimmed[1] = ctx->h_ops[0].sum_im;
  }
  if (rc) return rc;
  // moved rows leave stale index entries behind: compact when anything died or moved
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
    if (out) out->n_dense = ctx->n_dense;
  }
  return OBP_OK;
}

// ---- P2P exchange: the emitting kernel stores its records straight into the peer's inbox -----------
int obp_inbox_create(obp_ctx* ctx, uint64_t cap_records) {
  if (!ctx || cap_records < 1) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  free_inboxes(ctx);
  const size_t words = OBP_INBOX_HDR + (size_t)cap_records * (2 * ctx->W + 3);
  for (int b = 0; b < 2; ++b) {
    CU(cudaMalloc(&ctx->inbox[b], words * sizeof(u64)));
    CU(cudaMemset(ctx->inbox[b], 0, OBP_INBOX_HDR * sizeof(u64)));
    ctx->peer_inbox[b].assign(ctx->n_ranks, nullptr);
    ctx->peer_is_ipc[b].assign(ctx->n_ranks, false);
  }
  ctx->inbox_cap = cap_records;
  ctx->xchg_parity = 0;
  ctx->xseq = 0;
  return OBP_OK;
}

int obp_inbox_release(obp_ctx* ctx, int32_t stage) {
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (stage == 0) {  // unmap the peers' inboxes (all ranks do this before anyone frees)
    for (int b = 0; b < 2; ++b)
      for (size_t r = 0; r < ctx->peer_inbox[b].size(); ++r) {
        if (ctx->peer_inbox[b][r] && ctx->peer_is_ipc[b][r]) cudaIpcCloseMemHandle(ctx->peer_inbox[b][r]);
        ctx->peer_inbox[b][r] = nullptr;
      }
  } else {
    free_inboxes(ctx);
  }
  return OBP_OK;
}

int obp_inbox_export(obp_ctx* ctx, int32_t which, void* handle64, uint64_t* raw_ptr) {
  if (!ctx || which < 0 || which > 1 || !ctx->inbox[which]) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  if (raw_ptr) *raw_ptr = (uint64_t)(uintptr_t)ctx->inbox[which];
  if (handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->inbox[which]));
    memcpy(handle64, &h, 64);
  }
  return OBP_OK;
}

int obp_inbox_attach(obp_ctx* ctx, int32_t peer_rank, int32_t which, const void* handle64, uint64_t raw_ptr) {
  if (!ctx || which < 0 || which > 1 || peer_rank < 0 || peer_rank >= ctx->n_ranks ||
      (size_t)peer_rank >= ctx->peer_inbox[which].size())
    return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  if (handle64) {  // another process: map its allocation (enables peer access lazily)
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_inbox[which][peer_rank] = (u64*)p;
    ctx->peer_is_ipc[which][peer_rank] = true;
  } else {  // same process: the device pointer itself
    ctx->peer_inbox[which][peer_rank] = (u64*)(uintptr_t)raw_ptr;
    ctx->peer_is_ipc[which][peer_rank] = false;
  }
  return OBP_OK;
}

int obp_rotate_emit_p2p(obp_ctx* ctx, const obp_op* op, double atol, int32_t peer_rank, obp_shard_counters* out) {
  if (!ctx || !op || op->kind != OBP_OP_ROTATION || peer_rank < 0 || peer_rank >= ctx->n_ranks) return OBP_ERR_BADARG;
  const int b = ctx->xchg_parity;
  if (!ctx->inbox[b] || !ctx->peer_inbox[b][peer_rank]) return ctx->fail(OBP_ERR_BADARG, "obp_rotate_emit_p2p: peer inbox not attached");
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  int rc = validate_op(ctx, *op);
  if (rc) return rc;
  RotP p = build_rotp(ctx, *op, atol);
  p.epoch = ++ctx->epoch;
  p.slot = -1;
  ctx->shard_rot = p;
  ctx->shard_rot_open = true;
  ctx->n_split_rot++;
  const u64 n_in = ctx->n_live;
  u64* peer = ctx->peer_inbox[b][peer_rank];
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  {
    Bracket br(ctx, OBP_K_ROTATE);
    k_rotate_emit<<<grid_for(ctx, std::max<u64>(ctx->n_dense, 1), ctx->rot_ctas_per_sm), OBP_BS, 0, ctx->stream>>>(
        make_tab(ctx), p, peer + OBP_INBOX_HDR, ctx->inbox_cap, peer, ctx->d_cnt, nullptr, 0, nullptr, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  ctx->prof.term_updates[OBP_K_ROTATE] += n_in;
  fill_shard_stats(ctx, out, n_in);
  return rc;
}

int obp_rotate_merge_inbox(obp_ctx* ctx, obp_shard_counters* out) {
  if (!ctx) return OBP_ERR_BADARG;
  if (!ctx->shard_rot_open) return ctx->fail(OBP_ERR_BADARG, "obp_rotate_merge_inbox: no rotation in flight");
  const int b = ctx->xchg_parity;
  if (!ctx->inbox[b]) return ctx->fail(OBP_ERR_BADARG, "obp_rotate_merge_inbox: no inbox");
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  ctx->shard_rot_open = false;
  ctx->xchg_parity ^= 1;
  const u64 n_in = ctx->n_live;
  CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(ShardCnt), ctx->stream));
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_ROTATE);
    // the peer emits at most one record per row it holds; shards are balanced, so size the grid on ours
    const u64 guess = std::min<u64>(ctx->inbox_cap, std::max<u64>(2 * ctx->n_dense, 1 << 14));
    k_merge_records<<<grid_for(ctx, guess), OBP_BS, 0, ctx->stream>>>(t, ctx->inbox[b] + OBP_INBOX_HDR, 0, ctx->inbox[b],
                                                                     ctx->inbox_cap, ctx->d_cnt, nullptr, 0);
    const u64 bound = std::min<u64>(ctx->cap, std::max<u64>(ctx->n_dense + guess, 1));
    k_rotate_finish<<<grid_for(ctx, bound), OBP_BS, 0, ctx->stream>>>(t, ctx->shard_rot, ctx->d_cnt, nullptr, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemsetAsync(ctx->inbox[b], 0, sizeof(u64), ctx->stream));  // cursor back to 0 for the rotation after next
  CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(ShardCnt), cudaMemcpyDeviceToHost, ctx->stream));
  int rc = sync_state(ctx);
  fill_shard_stats(ctx, out, n_in);
  if (rc) return rc;
  if (ctx->n_dense - ctx->n_live > std::max<u64>(1024, ctx->n_live / 4)) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
    if (out) out->n_dense = ctx->n_dense;
  }
  return OBP_OK;
}

// -------------------------------------------------------------------------------------------------
// Generic gate: raw k*k expansion + the merge of a freshly loaded operator (simplify.py:115-165).
int obp_apply_generic(obp_ctx* ctx, int32_t n_qubits, const int32_t* qubits, int32_t n_comp, const uint32_t* codes,
                      const double* coeffs, double atol, int32_t simplify, obp_stats* stats) {
  OBP_NO_PENDING(ctx, "obp_apply_generic");
  if (!ctx || !qubits || !codes || !coeffs || n_qubits < 1 || n_qubits > 4 || n_comp < 1 || n_comp > OBP_GEN_MAXK)
    return OBP_ERR_BADARG;
  // (without simplify nothing merges, so it does not matter which rank holds a row: the expansion is local)
  if (ctx->n_ranks > 1 && simplify)
    return ctx->fail(OBP_ERR_UNSUPPORTED, "obp_apply_generic: with simplify, not available on a sharded table");
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  GenP p;
  memset(&p, 0, sizeof(p));
  p.m = n_qubits;
  p.k = n_comp;
  for (int q = 0; q < n_qubits; ++q) {
    if (qubits[q] < 0 || qubits[q] >= ctx->nq) return ctx->fail(OBP_ERR_BADARG, "obp_apply_generic: qubit out of range");
    for (int l = 0; l < q; ++l)
      if (qubits[l] == qubits[q]) return ctx->fail(OBP_ERR_BADARG, "obp_apply_generic: duplicate qubit");
    p.w[q] = qubits[q] / 64;
    p.b[q] = qubits[q] % 64;
  }
  for (int j = 0; j < n_comp; ++j) {
    if (codes[j] >> (2 * n_qubits)) return ctx->fail(OBP_ERR_BADARG, "obp_apply_generic: code outside the gate's qubits");
    p.code[j] = codes[j];
    p.ur[j] = coeffs[2 * j];
    p.ui[j] = coeffs[2 * j + 1];
  }
  int rc;
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  const u64 n = ctx->n_live, k2 = (u64)n_comp * n_comp, raw = n * k2;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (n + raw > ctx->cap) return ctx->fail(OBP_ERR_CAPACITY, "obp_apply_generic: the k*k raw rows do not fit the table");
  rc = upload_cols(ctx);
  if (rc) return rc;
  rc = ensure_ops(ctx, 4);
  if (rc) return rc;
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_OTHER, 2);
    if (raw) k_expand<<<grid_for(ctx, raw), OBP_BS, 0, ctx->stream>>>(t, p, n);
    if (n) k_kill_prefix<<<grid_for(ctx, n), OBP_BS, 0, ctx->stream>>>(t, n, ++ctx->epoch);
    CU(cudaGetLastError());
  }
  // the table now is: n tombstones followed by `raw` fresh rows, all counted live until filtered
  DevState s;
  memset(&s, 0, sizeof(s));
  s.n_scan = s.n_append = n + raw;
  s.n_live = raw;
  *ctx->h_st = s;
  CU(cudaMemcpyAsync(ctx->d_st, ctx->h_st, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat) * 2, ctx->stream));
  CU(cudaMemsetAsync(ctx->index, 0xFF, ctx->T * sizeof(u64), ctx->stream));
  ctx->index_stale = false;
  ctx->T_eff = ctx->T;
  if (!simplify) {  // OperatorBudget(simplify=False): keep every raw row, zeros included
    if (raw) {
      Bracket br(ctx, OBP_K_OTHER);
      k_revive_zeros<<<grid_for(ctx, raw), OBP_BS, 0, ctx->stream>>>(t, n, n + raw);
      CU(cudaGetLastError());
    }
    rc = sync_state(ctx);
    if (rc) return rc;
    ctx->prof.term_updates[OBP_K_OTHER] += n;
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
    if (stats) {
      stats->n_in = n;
      stats->raw_rows = raw;
      stats->updates = n;
      stats->n_out = ctx->n_live;
      stats->n_dense = ctx->n_dense;
    }
    return OBP_OK;
  }
  if (raw) {
    Bracket br(ctx, OBP_K_INDEX, 3);
    const int g = grid_for(ctx, n + raw);
    k_hash_rows<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->d_cols, n + raw);
    rc = reset_first_appearance(ctx, n + raw);
    if (rc) return rc;
    k_index_build<<<g, OBP_BS, 0, ctx->stream>>>(t, 1, atol, ++ctx->epoch, 0, ctx->row_order ? ctx->ord_minidx : nullptr);
    k_trim<<<g, OBP_BS, 0, ctx->stream>>>(t, atol, ++ctx->epoch, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  if (rc) return rc;
  ctx->prof.term_updates[OBP_K_OTHER] += n;
  if (stats) {
    const OpStat& o = ctx->h_ops[0];
    stats->n_in = n;
    stats->raw_rows = raw;
    stats->updates = n;
    stats->num_unique = o.keys_present;
    stats->num_duplicate = o.rows_surv - o.keys_present;
    stats->num_trimmed = (raw - o.rows_surv) + o.cancelled;
    stats->sum_trimmed[0] = o.sum_re;
    stats->sum_trimmed[1] = o.sum_im;
  }
  rc = compact_after_merge(ctx);
  if (rc) return rc;
  rc = sync_state(ctx);
  if (rc) return rc;
  if (stats) { stats->n_out = ctx->n_live; stats->n_dense = ctx->n_dense; }
  return OBP_OK;
}

// -------------------------------------------------------------------------------------------------
// Row-order mode reset: the literal reference sequence — zero the X/Y rows and clear the qubit in
// place (utils/operations.py:221-223), then simplify the rows where they lie (backpropagation.py:191).
static int reset_in_row_order(obp_ctx* ctx, int32_t qubit, double atol, obp_stats* stats) {
  int rc;
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  const u64 n_in = ctx->n_live;
  rc = upload_cols(ctx);
  if (rc) return rc;
  rc = ensure_ops(ctx, 4);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat) * 2, ctx->stream));
  CU(cudaMemsetAsync(ctx->index, 0xFF, ctx->T * sizeof(u64), ctx->stream));
  ctx->index_stale = false;
  ctx->T_eff = ctx->T;
  Tab t = make_tab(ctx);
  if (n_in) {
    Bracket br(ctx, OBP_K_OTHER, 4);
    const int g = grid_for(ctx, n_in);
    k_reset_raw<<<g, OBP_BS, 0, ctx->stream>>>(t, qubit / 64, 1ull << (qubit % 64));
    k_hash_rows<<<g, OBP_BS, 0, ctx->stream>>>(t, ctx->d_cols, n_in);
    rc = reset_first_appearance(ctx, n_in);
    if (rc) return rc;
    k_index_build<<<g, OBP_BS, 0, ctx->stream>>>(t, 1, atol, ++ctx->epoch, 0, ctx->ord_minidx);
    k_trim<<<g, OBP_BS, 0, ctx->stream>>>(t, atol, ++ctx->epoch, 0);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  if (rc) return rc;
  ctx->prof.term_updates[OBP_K_OTHER] += n_in;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    const OpStat& o = ctx->h_ops[0];
    stats->n_in = stats->raw_rows = stats->updates = n_in;
    stats->num_unique = o.keys_present;
    stats->num_duplicate = o.rows_surv - o.keys_present;
    stats->num_trimmed = (n_in - o.rows_surv) + o.cancelled;
    stats->sum_trimmed[0] = o.sum_re;
    stats->sum_trimmed[1] = o.sum_im;
  }
  if (ctx->n_dense != ctx->n_live) {
    rc = compact_in_first_appearance_order(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  if (stats) { stats->n_out = ctx->n_live; stats->n_dense = ctx->n_dense; }
  return OBP_OK;
}

int obp_apply_reset(obp_ctx* ctx, int32_t qubit, double atol, obp_stats* stats) {
  OBP_NO_PENDING(ctx, "obp_apply_reset");
  if (!ctx || qubit < 0 || qubit >= ctx->nq) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  { const int irc = ensure_index(ctx); if (irc) return irc; }
  if (atol == OBP_ATOL_RAW) {  // simplify=False: zero the X/Y rows, clear the qubit, merge nothing
    const u64 n_rows = ctx->n_live;
    {
      Bracket br(ctx, OBP_K_OTHER);
      k_reset_raw<<<grid_for(ctx, std::max<u64>(ctx->n_dense, 1)), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), qubit / 64,
                                                                                             1ull << (qubit % 64));
      CU(cudaGetLastError());
    }
    const int src = sync_state(ctx);
    if (src) return src;
    ctx->prof.term_updates[OBP_K_OTHER] += n_rows;
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->n_in = stats->raw_rows = stats->updates = stats->n_out = n_rows;
      stats->n_dense = ctx->n_dense;
    }
    return OBP_OK;
  }
  if (ctx->row_order) return reset_in_row_order(ctx, qubit, atol, stats);
  int rc = ensure_ops(ctx, 2);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->d_ops, 0, sizeof(OpStat), ctx->stream));
  const u64 n_in = ctx->n_live;
  ResetP p;
  memset(&p, 0, sizeof(p));
  p.w = qubit / 64;
  p.bit = 1ull << (qubit % 64);
  p.hZ = ctx->hz[qubit];
  p.atol = atol;
  p.slot = 0;
  Tab t = make_tab(ctx);
  {
    Bracket br(ctx, OBP_K_OTHER);
    const int g = grid_for(ctx, std::max<u64>(ctx->n_dense, 1));
    p.epoch = ++ctx->epoch;
    k_reset_filter<<<g, OBP_BS, 0, ctx->stream>>>(t, p);
    p.epoch = ++ctx->epoch;
    k_reset_merge<<<g, OBP_BS, 0, ctx->stream>>>(t, p);
    CU(cudaGetLastError());
  }
  CU(cudaMemcpyAsync(ctx->h_ops, ctx->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, ctx->stream));
  rc = sync_state(ctx);
  if (rc) return rc;
  ctx->prof.term_updates[OBP_K_OTHER] += n_in;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    const OpStat& o = ctx->h_ops[0];
    stats->n_in = n_in;
    stats->raw_rows = n_in;
    stats->updates = n_in;
    stats->num_unique = o.keys_present;
    stats->num_duplicate = o.rows_surv - o.keys_present;
    stats->num_trimmed = (n_in - o.rows_surv) + o.cancelled;
    stats->sum_trimmed[0] = o.sum_re;
    stats->sum_trimmed[1] = o.sum_im;
  }
  // moved rows leave stale index entries behind: compact when anything died or moved
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  if (stats) { stats->n_out = ctx->n_live; stats->n_dense = ctx->n_dense; }
  return OBP_OK;
}

int obp_apply_damping(obp_ctx* ctx, const uint64_t* gen_x, const uint64_t* gen_z, const double* factors,
                      int32_t n_gens, double atol, obp_stats* stats) {
  OBP_NO_PENDING(ctx, "obp_apply_damping");
  if (!ctx || n_gens < 0 || (n_gens && (!gen_x || !gen_z || !factors))) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  const u64 n_in = ctx->n_live;
  if (n_gens > ctx->gens_alloc) {
    cudaFree(ctx->d_gens);
    cudaFree(ctx->d_factors);
    ctx->d_gens = nullptr;
    ctx->d_factors = nullptr;
    ctx->gens_alloc = 0;
    CU(cudaMalloc(&ctx->d_gens, sizeof(u64) * 2 * ctx->W * n_gens));
    CU(cudaMalloc(&ctx->d_factors, sizeof(double) * n_gens));
    ctx->gens_alloc = n_gens;
  }
  if (n_gens) {
    std::vector<u64> packed((size_t)n_gens * 2 * ctx->W);
    for (int g = 0; g < n_gens; ++g)
      for (int w = 0; w < ctx->W; ++w) {
        packed[(size_t)g * 2 * ctx->W + w] = gen_x[(size_t)g * ctx->W + w];
        packed[(size_t)g * 2 * ctx->W + ctx->W + w] = gen_z[(size_t)g * ctx->W + w];
      }
    CU(cudaMemcpy(ctx->d_gens, packed.data(), packed.size() * sizeof(u64), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_factors, factors, sizeof(double) * n_gens, cudaMemcpyHostToDevice));
  }
  {
    Bracket br(ctx, OBP_K_OTHER);
    k_damp<<<grid_for(ctx, std::max<u64>(ctx->n_dense, 1)), OBP_BS, 0, ctx->stream>>>(
        make_tab(ctx), ctx->d_gens, ctx->d_factors, n_gens, atol, ++ctx->epoch, -1);
    CU(cudaGetLastError());
  }
  int rc = sync_state(ctx);
  if (rc) return rc;
  ctx->prof.term_updates[OBP_K_OTHER] += n_in;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_in = n_in;
    stats->raw_rows = n_in;
    stats->updates = n_in;
    stats->num_unique = ctx->n_live;  // keys are already distinct: every surviving row is unique
    stats->num_duplicate = 0;
    stats->num_trimmed = n_in - ctx->n_live;
    stats->n_out = ctx->n_live;
    stats->n_dense = ctx->n_dense;
  }
  return OBP_OK;
}

// -------------------------------------------------------------------------------------------------
// Stepwise truncation (utils/truncating.py:171-204): begin caches |c|^p and returns its maximum,
// masked_sum returns sum(a[a < threshold]), finish keeps a > threshold.  obp_truncate() runs the
// reference's bisection on one table; a sharded table runs the same loop on the host with the
// per-rank maxima / sums combined by an allreduce between the steps.
int obp_trunc_begin(obp_ctx* ctx, int32_t p_norm, double* local_max) {
  if (!ctx || p_norm < 1 || !local_max) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  *local_max = 0.0;
  CU(cudaMemsetAsync(&ctx->d_st->removed, 0, 2 * sizeof(u64), ctx->stream));  // removed, amax_bits
  if (ctx->n_dense) {
    Bracket br(ctx, OBP_K_TRUNCATE);
    k_absp<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), (double*)ctx->spare16, p_norm);
    CU(cudaGetLastError());
  }
  const int rc = sync_state(ctx);
  if (rc) return rc;
  const u64 bits = ctx->h_st->amax_bits;
  memcpy(local_max, &bits, sizeof(double));
  return OBP_OK;
}

int obp_trunc_masked_sum(obp_ctx* ctx, double threshold, double* local_sum) {
  if (!ctx || !local_sum) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  *local_sum = 0.0;
  if (ctx->n_dense == 0) return OBP_OK;
  const int gs = std::min(grid_for(ctx, ctx->n_dense), ctx->n_sm * 8);
  {
    Bracket br(ctx, OBP_K_TRUNCATE);
    k_masked_sum<<<gs, OBP_BS, 0, ctx->stream>>>((const double*)ctx->spare16, ctx->d_st, threshold, ctx->d_partial);
    k_final_sum<<<1, OBP_BS, 0, ctx->stream>>>(ctx->d_partial, gs, ctx->d_scalar);
    CU(cudaGetLastError());
  }
  close_run(ctx);
  CU(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *local_sum = ctx->h_scalar[0];
  return OBP_OK;
}

int obp_trunc_finish(obp_ctx* ctx, double threshold, uint64_t* n_removed, uint64_t* n_out) {
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  const u64 before = ctx->n_live;
  int rc;
  if (ctx->n_dense) {
    Bracket br(ctx, OBP_K_TRUNCATE);
    k_trunc_kill<<<grid_for(ctx, ctx->n_dense), OBP_BS, 0, ctx->stream>>>(make_tab(ctx), (const double*)ctx->spare16,
                                                                         threshold, ++ctx->epoch);
    CU(cudaGetLastError());
  }
  rc = sync_state(ctx);
  if (rc) return rc;
  ctx->prof.term_updates[OBP_K_TRUNCATE] += before;
  if (n_removed) *n_removed = before - ctx->n_live;
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  if (n_out) *n_out = ctx->n_live;
  return OBP_OK;
}

int obp_truncate(obp_ctx* ctx, double budget, int32_t p_norm, double tol, double* slice_error,
                 uint64_t* n_removed, uint64_t* n_out) {
  OBP_NO_PENDING(ctx, "obp_truncate");
  if (!ctx || p_norm < 1) return OBP_ERR_BADARG;
  if (slice_error) *slice_error = 0.0;
  if (n_removed) *n_removed = 0;
  if (n_out) *n_out = ctx->n_live;
  if (ctx->n_live == 0) return OBP_OK;
  if (ctx->trunc_ctas_per_sm > 0 && ctx->use_program) {
    // the whole search in one cooperative launch
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(&ctx->d_st->removed, 0, 2 * sizeof(u64), ctx->stream));  // removed, amax_bits
    const u64 before = ctx->n_live;
    Tab t = make_tab(ctx);
    double* a = (double*)ctx->spare16;
    // two CTAs per SM at most (a pass ends in two grid barriers, which cost more with more CTAs); a small table
    // runs in ONE CTA: its barriers are __syncthreads
    int g = ctx->n_dense <= OBP_TRUNC_ONE_CTA ? 1
                                               : std::min(grid_for(ctx, ctx->n_dense), std::min(ctx->trunc_ctas_per_sm, 2) * ctx->n_sm);
    u64 ep = ++ctx->epoch;
    void* args[] = {(void*)&t, (void*)&a, (void*)&ctx->d_partial, (void*)&ctx->d_ts, (void*)&budget, (void*)&p_norm,
                    (void*)&tol, (void*)&ep};
    {
      Bracket br(ctx, OBP_K_TRUNCATE);
      if (g == 1)  // no grid barrier inside: an ordinary launch (a cooperative one costs several microseconds more)
        k_trunc_search<<<1, OBP_BS, 0, ctx->stream>>>(t, a, ctx->d_partial, ctx->d_ts, budget, p_norm, tol, ep);
      else
        CU(cudaLaunchCooperativeKernel((void*)k_trunc_search, dim3(g), dim3(OBP_BS), args, 0, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->h_ts, ctx->d_ts, sizeof(TruncState), cudaMemcpyDeviceToHost, ctx->stream));
    int rc1 = sync_state(ctx);
    if (rc1) return rc1;
    ctx->prof.term_updates[OBP_K_TRUNCATE] += before;
    if (slice_error) *slice_error = ctx->h_ts->lower_err;
    if (n_removed) *n_removed = before - ctx->n_live;
    if (ctx->n_dense != ctx->n_live) {
      rc1 = compact(ctx);
      if (rc1) return rc1;
      rc1 = sync_state(ctx);
      if (rc1) return rc1;
    }
    if (n_out) *n_out = ctx->n_live;
    return OBP_OK;
  }
  double upper = 0.0;
  int rc = obp_trunc_begin(ctx, p_norm, &upper);
  if (rc) return rc;
  // utils/truncating.py:173-196, evaluated on the host with device-side masked sums
  double lower = 0.0, upper_error = budget, lower_error = 0.0;
  auto isclose = [&](double x, double y) { return std::fabs(x - y) <= tol + 1e-5 * std::fabs(y); };
  while ((upper - lower) > tol && !isclose(upper_error, lower_error)) {
    const double mid = (upper + lower) / 2;
    double sum = 0.0;
    rc = obp_trunc_masked_sum(ctx, mid, &sum);
    if (rc) return rc;
    const double mid_error = p_norm == 1 ? sum : std::pow(sum, 1.0 / p_norm);
    if (mid_error <= budget) {
      lower = mid;
      lower_error = mid_error;
    } else {
      upper = mid;
      upper_error = mid_error;
    }
  }
  if (slice_error) *slice_error = lower_error;
  return obp_trunc_finish(ctx, lower, n_removed, n_out);
}

// -------------------------------------------------------------------------------------------------
int obp_snapshot(obp_ctx* ctx) {
  OBP_NO_PENDING(ctx, "obp_snapshot");
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  // Copies the dense rows (tombstones included) aside with stream-ordered D2D copies: no
  // compaction, no host synchronisation.  The buffers are sized once for the whole table so the
  // per-slice commit never allocates.
  Snapshot& s = ctx->snap;
  const u64 n = ctx->n_dense;
  if (s.cap < ctx->cap) {
    CU(cudaStreamSynchronize(ctx->stream));
    free_snapshot(s);
    for (int w = 0; w < ctx->W; ++w) CU(cudaMalloc(&s.xz[w], ctx->cap * sizeof(ulonglong2)));
    CU(cudaMalloc(&s.coef, ctx->cap * sizeof(double2)));
    CU(cudaMalloc(&s.hash, ctx->cap * sizeof(u64)));
    s.cap = ctx->cap;
  }
  if (n) {
    CopyP p;
    memset(&p, 0, sizeof(p));
    for (int w = 0; w < ctx->W; ++w) { p.src16[w] = ctx->xz[w]; p.dst16[w] = s.xz[w]; }
    p.src16[ctx->W] = (const ulonglong2*)ctx->coef;
    p.dst16[ctx->W] = (ulonglong2*)s.coef;
    p.n16 = ctx->W + 1;
    p.src8 = ctx->hash;
    p.dst8 = s.hash;
    ctx->prof.launches[OBP_K_OTHER]++;  // counted, not timed: an open event run would swallow the next program
    k_copy_table<<<grid_for(ctx, n), OBP_BS, 0, ctx->stream>>>(p, n);
    CU(cudaGetLastError());
  }
  s.n = n;
  s.n_live = ctx->n_live;
  s.hx = ctx->hx;
  s.hz = ctx->hz;
  s.valid = true;
  return OBP_OK;
}

int obp_restore(obp_ctx* ctx) {
  OBP_NO_PENDING(ctx, "obp_restore");
  if (!ctx) return OBP_ERR_BADARG;
  if (!ctx->snap.valid) return ctx->fail(OBP_ERR_BADARG, "obp_restore: no snapshot");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaMemsetAsync(&ctx->d_st->error, 0, sizeof(unsigned), ctx->stream);
  Snapshot& s = ctx->snap;
  if (s.n > ctx->cap) return ctx->fail(OBP_ERR_CAPACITY, "obp_restore: snapshot larger than table");
  if (s.n) {
    CopyP p;
    memset(&p, 0, sizeof(p));
    for (int w = 0; w < ctx->W; ++w) { p.src16[w] = s.xz[w]; p.dst16[w] = ctx->xz[w]; }
    p.src16[ctx->W] = (const ulonglong2*)s.coef;
    p.dst16[ctx->W] = (ulonglong2*)ctx->coef;
    p.n16 = ctx->W + 1;
    p.src8 = s.hash;
    p.dst8 = ctx->hash;
    ctx->prof.launches[OBP_K_OTHER]++;
    k_copy_table<<<grid_for(ctx, s.n), OBP_BS, 0, ctx->stream>>>(p, s.n);
    CU(cudaGetLastError());
  }
  ctx->hx = s.hx;
  ctx->hz = s.hz;
  ctx->cols_dirty = true;
  int rc = set_counts(ctx, s.n);
  if (rc) return rc;
  if (s.n_live != s.n) {
    // the snapshot carried tombstones: n_live differs from the dense row count
    ctx->h_st->n_live = s.n_live;
    CU(cudaMemcpyAsync(&ctx->d_st->n_live, &ctx->h_st->n_live, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  rc = rebuild_index(ctx);
  if (rc) return rc;
  return sync_state(ctx);
}

static int alloc_row_order_scratch(obp_ctx* ctx) {
  if (ctx->ord_cap >= ctx->cap) return OBP_OK;
  cudaFree(ctx->ord_minidx);
  cudaFree(ctx->ord_flag);
  cudaFree(ctx->ord_pos);
  ctx->ord_minidx = ctx->ord_flag = ctx->ord_pos = nullptr;
  ctx->ord_cap = 0;
  CU(cudaMalloc(&ctx->ord_minidx, ctx->cap * sizeof(unsigned)));
  CU(cudaMalloc(&ctx->ord_flag, ctx->cap * sizeof(unsigned)));
  CU(cudaMalloc(&ctx->ord_pos, ctx->cap * sizeof(unsigned)));
  ctx->ord_cap = ctx->cap;
  return OBP_OK;
}

// (re)allocate the buffers that hold no table data — index and compaction scratch — for `cap` rows
static int alloc_scratch(obp_ctx* ctx, u64 cap) {
  ctx->T = next_pow2(std::max<u64>(2 * cap, 1024));
  ctx->T_eff = ctx->T;
  CU(cudaMalloc(&ctx->index, ctx->T * sizeof(u64)));
  CU(cudaMalloc(&ctx->spare16, cap * sizeof(ulonglong2)));
  CU(cudaMalloc(&ctx->spare8, cap * sizeof(u64)));
  CU(cudaMalloc(&ctx->dest, cap * sizeof(unsigned)));
  CU(cudaMalloc(&ctx->blk, ((cap + OBP_CHUNK - 1) / OBP_CHUNK + 1) * sizeof(unsigned)));
  return OBP_OK;
}

static void free_scratch(obp_ctx* ctx) {
  cudaFree(ctx->index);
  cudaFree(ctx->spare16);
  cudaFree(ctx->spare8);
  cudaFree(ctx->dest);
  cudaFree(ctx->blk);
  ctx->index = nullptr;
  ctx->spare16 = nullptr;
  ctx->spare8 = nullptr;
  ctx->dest = nullptr;
  ctx->blk = nullptr;
}

// Grow the table.  Failure-atomic: the data columns move one at a time (new buffer, copy, free the old
// one — the peak footprint stays one column above the table), so at every point each column holds the n
// live rows and has room for at least the OLD capacity.  If an allocation fails the table therefore
// simply stays at its old capacity: the scratch buffers (which hold no data) are re-created at the old
// size, the index is rebuilt and OBP_ERR_OOM is returned.  Only if even that fails is the context marked
// broken; every entry point then fails fast and the Python wrapper destroys it instead of pooling it.
int obp_reserve(obp_ctx* ctx, uint64_t new_capacity) {
  OBP_NO_PENDING(ctx, "obp_reserve");
  if (!ctx || new_capacity >= 0xffffffffull) return OBP_ERR_BADARG;
  if (new_capacity <= ctx->cap) return OBP_OK;
  CU(cudaSetDevice(ctx->device));
  int rc;
  if (ctx->n_dense != ctx->n_live) {
    rc = compact(ctx);
    if (rc) return rc;
    rc = sync_state(ctx);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(ctx->stream));
  const u64 n = ctx->n_live;
  const u64 old_cap = ctx->cap, cap = new_capacity;
  free_scratch(ctx);
  free_seg_columns(ctx);  // (sized by the capacity; allocated again by the next run of rotations that wants them)
  // one data column: bigger buffer, copy the live rows, release the old buffer
  auto grow_column = [&](void** slot, size_t elem) -> int {
    void* p = nullptr;
    CU(cudaMalloc(&p, cap * elem));
    const cudaError_t e = cudaMemcpy(p, *slot, n * elem, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) {
      cudaFree(p);
      return ctx->fail(OBP_ERR_CUDA, std::string("obp_reserve copy: ") + cudaGetErrorString(e));
    }
    cudaFree(*slot);
    *slot = p;
    return OBP_OK;
  };
  rc = OBP_OK;
  for (int w = 0; w < ctx->W && !rc; ++w) rc = grow_column((void**)&ctx->xz[w], sizeof(ulonglong2));
  if (!rc) rc = grow_column((void**)&ctx->coef, sizeof(double2));
  if (!rc) rc = grow_column((void**)&ctx->hash, sizeof(u64));
  if (!rc) {
    rc = alloc_scratch(ctx, cap);
    if (rc) free_scratch(ctx);
  }
  if (!rc) {
    ctx->cap = cap;
  } else {
    // stay at the old capacity (every column still has room for it)
    const std::string why = ctx->err;
    cudaGetLastError();
    const int rc2 = alloc_scratch(ctx, old_cap);
    if (rc2) {
      free_scratch(ctx);
      ctx->broken = true;
      return ctx->fail(rc, "obp_reserve: out of device memory and the table could not be put back (" + why + "); the context is unusable");
    }
    ctx->cap = old_cap;
    const int rc3 = rebuild_index(ctx);
    const int rc4 = rc3 ? rc3 : sync_state(ctx);
    if (rc4) {
      ctx->broken = true;
      return rc4;
    }
    return ctx->fail(rc, "obp_reserve: " + why + " (the table keeps its capacity of " + std::to_string(old_cap) + " terms)");
  }
  if (ctx->row_order) {
    rc = alloc_row_order_scratch(ctx);
    if (rc) return rc;
  }
  rc = rebuild_index(ctx);
  if (rc) return rc;
  return sync_state(ctx);
}

// Run-time switches of one context (tests flip them to force each code path against the oracle).
int obp_configure(obp_ctx* ctx, int32_t option, int64_t value) {
  if (!ctx) return OBP_ERR_BADARG;
  switch (option) {
    case OBP_OPT_PROGRAM: ctx->use_program = value != 0; return OBP_OK;
    case OBP_OPT_PROGRAM_MAX_ROWS: ctx->prog_max_rows = value < 0 ? 0 : (u64)value; return OBP_OK;
    case OBP_OPT_SEGMENTS: ctx->use_segments = value != 0; return OBP_OK;
    case OBP_OPT_SEGMENT_MIN_GATES: ctx->seg_min_gates = value < 1 ? 1 : (int)std::min<int64_t>(value, 1 << 20); return OBP_OK;
    case OBP_OPT_SEGMENT_CAP_ROWS: ctx->seg_cap_rows = value < 0 ? 0 : (int)std::min<int64_t>(value, 4096); return OBP_OK;
    case OBP_OPT_SEGMENT_MIXED: ctx->seg_mixed = value != 0; return OBP_OK;
    case OBP_OPT_SEGMENT_PIECES_MIN_ROWS: ctx->seg_pieces_min_rows = value < 0 ? 0 : (u64)value; return OBP_OK;
    default: return ctx->fail(OBP_ERR_BADARG, "obp_configure: unknown option");
  }
}

int obp_segment_stats(obp_ctx* ctx, uint64_t* n_runs, uint64_t* n_retries, uint64_t* n_fallbacks) {
  if (!ctx) return OBP_ERR_BADARG;
  if (n_runs) *n_runs = ctx->n_seg_runs;
  if (n_retries) *n_retries = ctx->n_seg_retries;
  if (n_fallbacks) *n_fallbacks = ctx->n_seg_fallbacks;
  return OBP_OK;
}

int obp_segment_partition_map(int32_t n_words, int32_t n_gen, const uint64_t* gen_x, const uint64_t* gen_z, uint64_t* fmask) {
  if ((n_words != 1 && n_words != 2) || n_gen < 0 || (n_gen && (!gen_x || !gen_z)) || !fmask) return OBP_ERR_BADARG;
  std::vector<SegGate> gates((size_t)n_gen);
  for (int k = 0; k < n_gen; ++k) {
    memset(&gates[(size_t)k], 0, sizeof(SegGate));
    for (int w = 0; w < n_words; ++w) {
      gates[(size_t)k].gx[w] = gen_x[(size_t)k * n_words + w];
      gates[(size_t)k].gz[w] = gen_z[(size_t)k * n_words + w];
    }
  }
  u64 fm[OBP_SEG_FBITS][4];
  seg_partition_map(gates, n_words, fm);
  for (int j = 0; j < OBP_SEG_FBITS; ++j)
    for (int q = 0; q < 4; ++q) fmask[4 * j + q] = fm[j][q];
  return OBP_OK;
}

int obp_is_broken(const obp_ctx* ctx) { return ctx && ctx->broken ? 1 : 0; }

int obp_set_row_order(obp_ctx* ctx, int32_t enabled) {
  if (!ctx) return OBP_ERR_BADARG;
  if (enabled && ctx->n_ranks > 1) return ctx->fail(OBP_ERR_UNSUPPORTED, "obp_set_row_order: not available on a sharded table");
  CU(cudaSetDevice(ctx->device));
  ctx->row_order = enabled != 0;
  return ctx->row_order ? alloc_row_order_scratch(ctx) : OBP_OK;
}

// -------------------------------------------------------------------------------------------------
int obp_union_count(obp_ctx* const* ctxs, int32_t n_ctx, uint64_t* n_distinct) {
  if (!ctxs || n_ctx < 1 || !n_distinct) return OBP_ERR_BADARG;
  obp_ctx* ctx = ctxs[0];
  CU(cudaSetDevice(ctx->device));
  if (n_ctx == 1) {
    *n_distinct = ctx->n_live;
    return OBP_OK;
  }
  u64 total = 0;
  for (int k = 0; k < n_ctx; ++k) {
    obp_ctx* c = ctxs[k];
    if (c->nq != ctx->nq || c->device != ctx->device) return ctx->fail(OBP_ERR_BADARG, "obp_union_count: mismatching contexts");
    if (c->n_dense != c->n_live) {
      int rc = compact(c);
      if (rc) { ctx->err = c->err; return rc; }
      rc = sync_state(c);
      if (rc) { ctx->err = c->err; return rc; }
    }
    CU(cudaStreamSynchronize(c->stream));
    total += c->n_live;
  }
  if (total == 0) {
    *n_distinct = 0;
    return OBP_OK;
  }
  if (!ctx->scratch || ctx->scratch->cap < total) {
    if (ctx->scratch) obp_destroy(ctx->scratch);
    ctx->scratch = nullptr;
    const int rc = obp_create(ctx->device, ctx->nq, std::max<u64>(total + total / 2, 1024), &ctx->scratch);
    if (rc) return ctx->fail(rc, std::string("obp_union_count scratch: ") + obp_last_error(nullptr));
  }
  obp_ctx* sc = ctx->scratch;
  Tab ts = make_tab(sc);
  u64 off = 0;
  for (int k = 0; k < n_ctx; ++k) {
    obp_ctx* c = ctxs[k];
    if (c->n_live == 0) continue;
    k_copy_live_keys<<<grid_for(sc, c->n_live), OBP_BS, 0, sc->stream>>>(make_tab(c), ts, off);
    off += c->n_live;
  }
  if (cudaGetLastError() != cudaSuccess) return ctx->fail(OBP_ERR_CUDA, "k_copy_live_keys launch failed");
  int rc = set_counts(sc, total);
  if (rc) return ctx->fail(rc, sc->err);
  init_frame(sc);
  rc = upload_cols(sc);
  if (rc) return ctx->fail(rc, sc->err);
  cudaMemsetAsync(sc->d_ops, 0, sizeof(OpStat), sc->stream);
  cudaMemsetAsync(sc->index, 0xFF, sc->T * sizeof(u64), sc->stream);
  ts = make_tab(sc);
  k_hash_rows<<<grid_for(sc, total), OBP_BS, 0, sc->stream>>>(ts, sc->d_cols, total);
  k_index_build<<<grid_for(sc, total), OBP_BS, 0, sc->stream>>>(ts, 1, -1.0, ++sc->epoch, 0, nullptr);
  if (cudaGetLastError() != cudaSuccess) return ctx->fail(OBP_ERR_CUDA, "union count launch failed");
  cudaMemcpyAsync(sc->h_ops, sc->d_ops, sizeof(OpStat), cudaMemcpyDeviceToHost, sc->stream);
  rc = sync_state(sc);
  if (rc) return ctx->fail(rc, sc->err);
  *n_distinct = sc->h_ops[0].keys_present;
  return OBP_OK;
}

// -------------------------------------------------------------------------------------------------
// Qubit-wise commuting groups (backpropagation.py:368-375, 433-440), see obp_qwc.cuh.  Stand-alone: the
// rows are the union of all observables' keys in the caller's order (ties of the greedy colouring are
// broken by row order, like rustworkx does on Qiskit's row order).
int obp_qwc_group_count(int device, int n_qubits, const uint64_t* x_words, const uint64_t* z_words, uint64_t n,
                        uint64_t* n_groups) {
  if (!n_groups || n_qubits < 1 || n_qubits > 64 * OBP_MAX_WORDS || (n && (!x_words || !z_words))) {
    g_create_error = "obp_qwc_group_count: bad argument";
    return OBP_ERR_BADARG;
  }
  *n_groups = 0;
  if (n == 0) return OBP_OK;
  if (n > (1ull << 20)) {  // the colour bit map lives in shared memory; beyond this the O(n^2) pass is impractical anyway
    g_create_error = "obp_qwc_group_count: more than 2^20 terms";
    return OBP_ERR_UNSUPPORTED;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    g_create_error = "obp_qwc_group_count: cudaSetDevice failed";
    return OBP_ERR_CUDA;
  }
  cudaDeviceProp prop;
  int n_sm = 148;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    n_sm = prop.multiProcessorCount;
    if (prop.major < 10) {
      g_create_error = "obp_b200 requires an sm_100a (Blackwell B200) device";
      return OBP_ERR_UNSUPPORTED;
    }
  }
  const int W = (n_qubits + 63) / 64;
  int Wp = 1;
  while (Wp < W) Wp <<= 1;  // kernels are instantiated for 1, 2, 4, 8, 16 words: pad with identity words
  std::vector<ulonglong2> keys((size_t)n * Wp, make_ulonglong2(0ull, 0ull));
  for (u64 i = 0; i < n; ++i)
    for (int w = 0; w < W; ++w) keys[(size_t)i * Wp + w] = make_ulonglong2(x_words[i * W + w], z_words[i * W + w]);
  ulonglong2 *d_keys = nullptr, *d_sorted = nullptr;
  unsigned *d_u = nullptr;  // degree | order | colour | n_colours
  int rc = OBP_OK;
  cudaError_t e;
  const size_t kb = keys.size() * sizeof(ulonglong2);
  if ((e = cudaMalloc(&d_keys, kb)) != cudaSuccess || (e = cudaMalloc(&d_sorted, kb)) != cudaSuccess ||
      (e = cudaMalloc(&d_u, (3 * n + 1) * sizeof(unsigned))) != cudaSuccess) {
    g_create_error = std::string("obp_qwc_group_count: cudaMalloc: ") + cudaGetErrorString(e);
    rc = OBP_ERR_OOM;
  }
  if (!rc && ((e = cudaMemcpy(d_keys, keys.data(), kb, cudaMemcpyHostToDevice)) != cudaSuccess ||
              (e = cudaMemset(d_u, 0, (3 * n + 1) * sizeof(unsigned))) != cudaSuccess)) {
    g_create_error = std::string("obp_qwc_group_count: upload: ") + cudaGetErrorString(e);
    rc = OBP_ERR_CUDA;
  }
  unsigned ncol = 0;
  if (!rc) {
    std::vector<unsigned> order;
    unsigned *dg = d_u, *od = d_u + n, *co = d_u + 2 * n, *nc = d_u + 3 * n;
    const unsigned nn = (unsigned)n;
    switch (Wp) {
      case 1: rc = qwc_run<1>(d_keys, d_sorted, dg, od, co, nc, nn, n_sm, order, &ncol, g_create_error); break;
      case 2: rc = qwc_run<2>(d_keys, d_sorted, dg, od, co, nc, nn, n_sm, order, &ncol, g_create_error); break;
      case 4: rc = qwc_run<4>(d_keys, d_sorted, dg, od, co, nc, nn, n_sm, order, &ncol, g_create_error); break;
      case 8: rc = qwc_run<8>(d_keys, d_sorted, dg, od, co, nc, nn, n_sm, order, &ncol, g_create_error); break;
      default: rc = qwc_run<16>(d_keys, d_sorted, dg, od, co, nc, nn, n_sm, order, &ncol, g_create_error); break;
    }
  }
  cudaFree(d_keys);
  cudaFree(d_sorted);
  cudaFree(d_u);
  *n_groups = ncol;
  return rc;
}

// -------------------------------------------------------------------------------------------------
int obp_profile_enable(obp_ctx* ctx, int32_t on) {
  if (!ctx) return OBP_ERR_BADARG;
  ctx->prof_on = on != 0;
  return OBP_OK;
}

int obp_profile_get(obp_ctx* ctx, obp_profile* out, int32_t reset) {
  if (!ctx || !out) return OBP_ERR_BADARG;
  *out = ctx->prof;
  if (reset) memset(&ctx->prof, 0, sizeof(ctx->prof));
  return OBP_OK;
}

int obp_timer_start(obp_ctx* ctx) {
  if (!ctx) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->tm_a) {
    CU(cudaEventCreate(&ctx->tm_a));
    CU(cudaEventCreate(&ctx->tm_b));
  }
  CU(cudaEventRecord(ctx->tm_a, ctx->stream));
  return OBP_OK;
}

int obp_timer_stop(obp_ctx* ctx, double* ms) {
  if (!ctx || !ms || !ctx->tm_a) return OBP_ERR_BADARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->tm_b, ctx->stream));
  CU(cudaEventSynchronize(ctx->tm_b));
  float f = 0.f;
  CU(cudaEventElapsedTime(&f, ctx->tm_a, ctx->tm_b));
  *ms = (double)f;
  return OBP_OK;
}

}  // extern "C"
