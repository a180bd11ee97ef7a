// obp_device.cuh — device-side data layout and helpers of the operator-backpropagation engine.
//
// Term table (one per observable), all arrays indexed by the dense row id i < n_scan:
//   xz[w][i]  ulonglong2 {x word w, z word w}   16-B vector loads, coalesced over i   (W columns)
//   coef[i]   double2 {re, im}                  complex128 coefficient
//   hash[i]   u64                               GF(2)-linear hash of the key (invariant under
//                                               Clifford gates, see obp_b200.cu: HashFrame)
//   index[s]  u64  (hash>>32)<<32 | row         open-addressing table, linear probing, T = 2^k
//
// A row whose coefficient is {0.0, <integer bits below 2^52>} is a tombstone: the integer is the
// epoch (kernel sequence number) in which the row died.  Tombstones are never revived; lookups
// walk past them.  Compaction removes them and rebuilds the index.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "obp_b200.h"

typedef unsigned long long u64;

#define OBP_BS 256            // threads per CTA of every table kernel
#define OBP_EMPTY (~0ull)     // empty index slot
#define OBP_ERR_BIT_CAP 1u    // dense table full
#define OBP_ERR_BIT_INDEX 2u  // index probe wrapped around
#define OBP_ERR_BIT_PEER 4u   // a peer's flag did not arrive in time (sharded table, device-side handshake)
#define OBP_CLIFF_BATCH 96    // Clifford gates fused into one launch

struct DevState {
  u64 n_scan;    // rows visible to scans (stable while a kernel runs)
  u64 n_append;  // append cursor (atomic)
  u64 n_live;    // live rows
  u64 removed;   // truncation: rows removed by the last kill pass
  u64 amax_bits; // truncation: max |c|^p (bit pattern of a non-negative double)
  u64 scan_total;
  unsigned int ticket;  // last-block-done counter
  unsigned int error;   // OBP_ERR_BIT_*
  u64 emitted_total;    // sharded table: records this rank has emitted to peers so far (monotonic, statistics)
};

// Result block: [DevState | pad to 64 | program start stamp | pad to 128 | OpStat slots...], the same layout on
// the device and in the host's pinned mirror
#define OBP_RES_HDR 128

struct OpStat {  // one slot per launched op; zeroed at program start
  u64 n_out;         // live rows after the op
  u64 rows_surv;     // raw rows that passed the |row| <= atol filter
  u64 keys_present;  // distinct keys among them
  u64 cancelled;     // merged keys with |sum| <= atol
  double sum_re, sum_im;  // sum of the cancelled merged coefficients
  u64 n_surv;        // Clifford: live rows that passed the row filter
  u64 ncnt[17];      // Clifford: rows whose coset {P xor D} holds cnt live members
  u64 dlive;         // persistent program: change of the live-row count by this op (two's complement)
  u64 t_end;         // persistent program: %globaltimer (ns) when the op's grid barrier was passed
};

struct Tab {
  ulonglong2* xz[OBP_MAX_WORDS];
  double2* coef;
  u64* hash;
  u64* index;
  u64 mask;  // T - 1
  u64 cap;
  int W;
  DevState* st;
  OpStat* ops;
};

__device__ __forceinline__ bool is_live(double2 c) {
  return !(c.x == 0.0 && (((u64)__double_as_longlong(c.y)) >> 52) == 0ull);
}
__device__ __forceinline__ double2 tombstone(u64 epoch) {
  return make_double2(0.0, __longlong_as_double((long long)epoch));
}
__device__ __forceinline__ u64 tomb_epoch(double2 c) { return (u64)__double_as_longlong(c.y); }

// complex multiply without FMA contraction (numpy evaluates ar*br - ai*bi with separate roundings)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                      __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) {
  return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
}
// a * (-i)^e, exact
// (branch-free: the exponent differs from lane to lane, a switch would run its four cases one after the other)
__device__ __forceinline__ double2 mul_mi_pow(double2 a, int e) {
  const bool odd = (e & 1) != 0, neg = (e & 2) != 0;
  const double re = odd ? a.y : a.x;    // e = 1: (y, -x)
  const double im = odd ? -a.x : a.y;
  return make_double2(neg ? -re : re, neg ? -im : im);  // e = 2: (-x, -y), e = 3: (-y, x)
}
// np.isclose(c, 0, atol, rtol) == (|c| <= atol) with |c| = hypot(re, im); NaN compares false, i.e.
// the row is kept.  hypot() is ~60 instructions in fp64, so the squared magnitude decides unless it
// lies within 1e-9 (relative) of atol^2; only then the exact expression is evaluated.  The result is
// bit-identical to the plain hypot test.
__device__ __forceinline__ bool is_small(double2 a, double atol) {
  if (atol < 0.0) return false;
  const double m2 = fma(a.x, a.x, a.y * a.y);
  const double t2 = atol * atol;
  if (m2 > t2 * (1.0 + 1e-9)) return false;
  if (m2 < t2 * (1.0 - 1e-9)) return true;
  return hypot(a.x, a.y) <= atol;
}
// |c| * scale <= atol (raw-row filter of a Clifford-class gate), same two-stage evaluation
__device__ __forceinline__ bool is_small_scaled(double2 a, double scale, double atol) {
  const double m2 = fma(a.x, a.x, a.y * a.y) * (scale * scale);
  const double t2 = atol * atol;
  if (m2 > t2 * (1.0 + 1e-9)) return false;
  if (m2 < t2 * (1.0 - 1e-9)) return true;
  return hypot(a.x, a.y) * scale <= atol;
}

__device__ __forceinline__ u64 ld_volatile(const u64* p) { return *(const volatile u64*)p; }

// ---- warp / block reductions -----------------------------------------------------------------
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exclusive scan of a 0/1 flag over the CTA (OBP_BS threads).  Returns this thread's offset and
// the CTA total.  `warp_tot` is OBP_BS/32 ints of shared memory.  Contains two __syncthreads.
__device__ __forceinline__ unsigned block_scan_flag(bool flag, unsigned* warp_tot, unsigned& total) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const unsigned pre = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < OBP_BS / 32; ++w) {
    const unsigned v = warp_tot[w];
    if ((unsigned)w < warp) base += v;
    tot += v;
  }
  __syncthreads();
  total = tot;
  return base + pre;
}

// The last CTA to arrive publishes the append cursor as the next scan bound and records n_live.
__device__ __forceinline__ void finish_kernel(const Tab& t, int slot) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&t.st->ticket, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence();
      u64 na = atomicAdd(&t.st->n_append, 0ull);
      if (na > t.cap) na = t.cap;
      t.st->n_append = na;
      t.st->n_scan = na;
      if (slot >= 0) t.ops[slot].n_out = atomicAdd(&t.st->n_live, 0ull);
      t.st->ticket = 0;
    }
  }
}

// Publish row `row` (hash h) in the index.  Returns false if the probe wrapped around.
__device__ __forceinline__ bool index_insert(const Tab& t, u64 h, u64 row) {
  const u64 entry = ((h >> 32) << 32) | row;
  u64 s = h & t.mask;
  for (u64 probes = 0; probes <= t.mask; ++probes) {
    const u64 old = atomicCAS(&t.index[s], OBP_EMPTY, entry);
    if (old == OBP_EMPTY) return true;
    s = (s + 1) & t.mask;
  }
  return false;
}

// ---- device-side handshake between ranks (sharded table) ----------------------------------------
// Inbox header words: [0] record cursor, [1] emitted_seq (set by the SENDER's kernel when all its
// records are visible), [2] merged_seq (set by the OWNER when it has consumed the inbox).
#define OBP_HDR_CURSOR 0
#define OBP_HDR_EMITTED 1
#define OBP_HDR_MERGED 2

__device__ __forceinline__ u64 globaltimer_ns() {
  u64 v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

// Thread 0 of the CTA waits until *flag >= want (the flag may live in a peer GPU's memory), then the
// CTA proceeds.  The waited-for kernel runs on ANOTHER GPU; a 10 s timeout turns a lost peer into an
// error flag instead of a hang.
__device__ __forceinline__ void cta_wait_flag(const u64* flag, u64 want, unsigned* err) {
  if (flag != nullptr) {
    if (threadIdx.x == 0) {
      const u64 t0 = globaltimer_ns();
      // once a wait has timed out, later waits give up at once: the run is lost, it must not hang
      while (ld_volatile(flag) < want && !(*(volatile unsigned*)err & OBP_ERR_BIT_PEER)) {
        if (globaltimer_ns() - t0 > 10000000000ull) {
          atomicOr(err, OBP_ERR_BIT_PEER);
          break;
        }
        __nanosleep(100);
      }
      __threadfence_system();
    }
    __syncthreads();
  }
}

// finish_kernel() plus: the last CTA raises a flag word (possibly in a peer's memory) once every
// CTA's stores are visible system-wide; `also_zero` (optional) is cleared just before.
__device__ __forceinline__ void finish_kernel_flag(const Tab& t, int slot, u64* flag, u64 value, u64* also_zero) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&t.st->ticket, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      u64 na = atomicAdd(&t.st->n_append, 0ull);
      if (na > t.cap) na = t.cap;
      t.st->n_append = na;
      t.st->n_scan = na;
      if (slot >= 0) t.ops[slot].n_out = atomicAdd(&t.st->n_live, 0ull);
      t.st->ticket = 0;
      if (also_zero) *(volatile u64*)also_zero = 0ull;
      __threadfence_system();
      if (flag) *(volatile u64*)flag = value;
    }
  }
}
