// obp_segment.cuh — a whole run of Pauli rotations applied in shared memory ("slice in shared memory").
//
// Replaces, for a slice (or a run of gates of a slice) that consists of single-generator rotations only — the RX /
// RY / RZ layers and the RZZ / RXX / RYY layers of a Trotter step — the gate-by-gate loop
//     for gate in slice:  op = op.compose(U, front=True).compose(U^dag);  op = simplify(op)
// (backpropagation.py:208-237, utils/operations.py:103-105, utils/simplify.py:115-165) that k_rotate runs as one
// pass over the table per gate.  On the tables of the 127-qubit workload such a pass is a chain of dependent L2
// round trips between two grid barriers, 4-6 us per gate whatever the row count, and a slice has 127 of them.
//
// A rotation with generator G only ever mixes a key P with P^G.  The host therefore picks a GF(2)-linear map
// f: keys -> 16 bits that VANISHES on every generator of the run (a basis of the annihilator of span{G_g}, folded
// with random weights: for an RX layer f reads z bits only, for an RZZ layer it reads x bits and the parity z_a^z_b)
// and the table is partitioned by f: all keys a term can ever mix with during the run share its partition.  One CTA
// then owns a partition outright: it gathers the partition's rows into shared memory (keys, coefficients, stored
// hashes, a local open-addressing index of 16-bit row ids), applies ALL gates of the run there — one pass per gate
// with exactly the arithmetic of rotate_body (the same raw rows, the same two zero filters, the same merge order,
// exact simplify counters for the slice's last gate), separated by one __syncthreads — and writes the surviving rows
// out, compacted, to the table's second set of columns.  Global memory is read once and written once per RUN instead
// of once per GATE; the per-gate cost is a shared-memory pass over the partition.
//
// Kernels: k_seg_count (partition id per live row + histogram), k_seg_scan (offsets), k_seg_scatter (row ids in
// partition order), k_seg_main<W> (the run), k_seg_finish (per-gate row counts, table state, result block).
// The input table is never modified: if a partition outgrows its shared memory (or the output its capacity) the run
// reports it and the host applies the same gates through the per-gate kernels instead.
#pragma once
#include "obp_kernels.cuh"

#define OBP_SEG_THREADS 512
#define OBP_SEG_MAXG 192    // gates per run (their parameters are staged in shared memory)
#define OBP_SEG_MAXP 8192   // partitions per run
#define OBP_SEG_FBITS 16    // output bits of the partition map
#define OBP_SEG_ALIVE 255u

struct SegGate {  // 80 bytes
  u64 gx[2], gz[2];  // generator on table words 0 and 1 (zero where it is the identity)
  u64 hG;            // hash of the generator
  double u0r, u0i, u1r, u1i;
  int yG;
  int simple;
};

struct SegP {
  int ng;            // gates
  int P;             // partitions
  int cap_rows;      // rows a partition may hold in shared memory
  int n_index;       // local index slots (power of two)
  int exact_last;    // the last gate reports the reference's simplify counters
  int slot0;         // counter slot of gate 0 (gate g: slot0 + g)
  u64 n_dense;       // rows of the input table
  u64 out_cap;       // rows the output columns hold
  double atol;
  u64 fmask[OBP_SEG_FBITS][4];  // partition map: bit j = parity of (x0 & m[0]) ^ (z0 & m[1]) ^ (x1 & m[2]) ^ (z1 & m[3])
};

struct SegRes {  // device-side result of a run (zeroed before)
  u64 out_rows;          // output cursor
  unsigned overflow;     // 1: a partition outgrew shared memory, 2: the output outgrew its columns, 4: a local index filled up
  unsigned pad;
  long long delta[OBP_SEG_MAXG];  // change of the live-row count by every gate
};

struct SegOut {
  ulonglong2* xz[2];
  double2* coef;
  u64* hash;
};

template <int W>
__device__ __forceinline__ unsigned seg_part(const SegP& s, const ulonglong2* key) {
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < OBP_SEG_FBITS; ++j) {
    u64 a = (key[0].x & s.fmask[j][0]) ^ (key[0].y & s.fmask[j][1]);
    if (W == 2) a ^= (key[1].x & s.fmask[j][2]) ^ (key[1].y & s.fmask[j][3]);
    v |= (unsigned)(__popcll(a) & 1) << j;
  }
  return (v * (unsigned)s.P) >> OBP_SEG_FBITS;
}

template <int W>
__global__ void __launch_bounds__(OBP_BS) k_seg_count(Tab t, const __grid_constant__ SegP s, unsigned* __restrict__ part,
                                                      unsigned* __restrict__ cnt) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < s.n_dense; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    unsigned p = 0xffffffffu;
    if (is_live(c)) {
      ulonglong2 key[W];
#pragma unroll
      for (int w = 0; w < W; ++w) key[w] = t.xz[w][i];
      p = seg_part<W>(s, key);
      atomicAdd(&cnt[p], 1u);
    }
    part[i] = p;
  }
}

// exclusive scan of the partition sizes (P <= OBP_SEG_MAXP): off[p], off[P] = live rows; cur[p] = 0
__global__ void __launch_bounds__(1024) k_seg_scan(const unsigned* __restrict__ cnt, unsigned* __restrict__ off,
                                                   unsigned* __restrict__ cur, int P) {
  __shared__ unsigned s_tot[1024];
  const int per = (P + 1023) / 1024;
  const int p0 = threadIdx.x * per, p1 = min(P, p0 + per);
  unsigned local = 0;
  for (int p = p0; p < p1; ++p) local += cnt[p];
  s_tot[threadIdx.x] = local;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    const unsigned v = threadIdx.x >= (unsigned)d ? s_tot[threadIdx.x - d] : 0u;
    __syncthreads();
    s_tot[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = s_tot[threadIdx.x] - local;
  for (int p = p0; p < p1; ++p) {
    off[p] = run;
    cur[p] = 0;
    run += cnt[p];
  }
  if (threadIdx.x == 1023) off[P] = s_tot[1023];
}

__global__ void __launch_bounds__(OBP_BS) k_seg_scatter(const unsigned* __restrict__ part, const unsigned* __restrict__ off,
                                                        unsigned* __restrict__ cur, unsigned* __restrict__ perm, u64 n_dense) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n_dense; i += (u64)gridDim.x * OBP_BS) {
    const unsigned p = part[i];
    if (p != 0xffffffffu) perm[off[p] + atomicAdd(&cur[p], 1u)] = (unsigned)i;
  }
}

// shared-memory view of one partition
template <int W>
struct SegSmem {
  ulonglong2* xz[W];
  double2* coef;
  u64* hash;
  unsigned short* idx;
  unsigned char* died;  // OBP_SEG_ALIVE, or the gate that killed the row
  SegGate* gate;
  int* dg;              // per-gate change of the live-row count, summed over the CTA's partitions
};

// Row of the partition holding `key ^ flips` with hash h2 that was present when pass g started (alive, or killed
// during pass g), or -1.
template <int W>
__device__ __forceinline__ int seg_find(const SegSmem<W>& m, unsigned idx_mask, u64 h2, const ulonglong2* want, int n_pass,
                                        unsigned g, bool& fresh) {
  unsigned s = (unsigned)h2 & idx_mask;
  fresh = false;
  for (unsigned probes = 0; probes <= idx_mask; ++probes) {
    const unsigned v = m.idx[s];
    if (v == 0u) return -1;
    const int j = (int)v - 1;
    if (j < n_pass && m.hash[j] == h2) {
      bool eq = true;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const ulonglong2 kj = m.xz[w][j];
        eq = eq && kj.x == want[w].x && kj.y == want[w].y;
      }
      if (eq) {
        const unsigned d = m.died[j];
        if (d == OBP_SEG_ALIVE) return j;
        if (d == g) { fresh = true; return j; }
        // a stale tombstone of the same key: a live duplicate may follow
      }
    }
    s = (s + 1) & idx_mask;
  }
  return -1;
}

template <int W>
__device__ __forceinline__ bool seg_insert(const SegSmem<W>& m, unsigned idx_mask, u64 h, int row) {
  unsigned s = (unsigned)h & idx_mask;
  const unsigned short val = (unsigned short)(row + 1);
  for (unsigned probes = 0; probes <= idx_mask; ++probes) {
    const unsigned short old = atomicCAS(&m.idx[s], (unsigned short)0, val);
    if (old == 0) return true;
    s = (s + 1) & idx_mask;
  }
  return false;
}

template <int W>
__global__ void __launch_bounds__(OBP_SEG_THREADS, 1)
k_seg_main(Tab t, SegOut out, const __grid_constant__ SegP s, const SegGate* __restrict__ gates,
           const unsigned* __restrict__ off, const unsigned* __restrict__ perm, SegRes* __restrict__ res) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CAP = s.cap_rows;
  SegSmem<W> m;
  {
    unsigned char* p = smem_raw;
#pragma unroll
    for (int w = 0; w < W; ++w) { m.xz[w] = reinterpret_cast<ulonglong2*>(p); p += (size_t)CAP * 16; }
    m.coef = reinterpret_cast<double2*>(p); p += (size_t)CAP * 16;
    m.gate = reinterpret_cast<SegGate*>(p); p += (size_t)s.ng * sizeof(SegGate);
    m.hash = reinterpret_cast<u64*>(p); p += (size_t)CAP * 8;
    m.dg = reinterpret_cast<int*>(p); p += (size_t)((s.ng + 1) & ~1) * 4;
    m.idx = reinterpret_cast<unsigned short*>(p); p += (size_t)s.n_index * 2;
    m.died = p;
  }
  __shared__ int s_add[3];       // rows appended during a pass (rotating: read / written / cleared)
  __shared__ int s_live;         // live rows of the partition
  __shared__ int s_out;          // output cursor inside the partition
  __shared__ u64 s_base;         // first output row of the partition
  const unsigned idx_mask = (unsigned)s.n_index - 1u;
  const unsigned lane = threadIdx.x & 31u;
  const int nt = blockDim.x;
  // gate parameters and the per-gate counters: once per CTA
  for (int k = threadIdx.x; k < s.ng * (int)(sizeof(SegGate) / 8); k += nt)
    reinterpret_cast<u64*>(m.gate)[k] = reinterpret_cast<const u64*>(gates)[k];
  for (int k = threadIdx.x; k < s.ng; k += nt) m.dg[k] = 0;
  // exact counters of the last gate (thread-local over all partitions of this CTA)
  unsigned rows_surv = 0, keys_present = 0, cancelled = 0;
  double sum_re = 0.0, sum_im = 0.0;
  __syncthreads();

  for (int part = blockIdx.x; part < s.P; part += gridDim.x) {
    const unsigned o0 = off[part];
    const int n_in = (int)(off[part + 1] - o0);
    if (n_in == 0) continue;
    if (n_in > CAP) {
      if (threadIdx.x == 0) atomicOr(&res->overflow, 1u);
      continue;
    }
    // ---- gather the partition ----------------------------------------------------------------------------------
    for (int k = threadIdx.x; k < s.n_index / 4; k += nt) reinterpret_cast<u64*>(m.idx)[k] = 0ull;
    if (threadIdx.x == 0) {
      s_add[0] = 0; s_add[1] = 0; s_add[2] = 0;
      s_live = n_in;
      s_out = 0;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n_in; r += nt) {
      const u64 i = perm[o0 + r];
#pragma unroll
      for (int w = 0; w < W; ++w) m.xz[w][r] = t.xz[w][i];
      m.coef[r] = t.coef[i];
      const u64 h = t.hash[i];
      m.hash[r] = h;
      m.died[r] = (unsigned char)OBP_SEG_ALIVE;
      if (!seg_insert<W>(m, idx_mask, h, r)) atomicOr(&res->overflow, 4u);  // (cannot happen: n_index >= 2 * cap_rows)
    }
    __syncthreads();
    // ---- the gates --------------------------------------------------------------------------------------------
    int n_pass = n_in;
    bool failed = false;
    for (int g = 0; g < s.ng && !failed; ++g) {
      const SegGate& G = m.gate[g];
      RotP pg;  // the fields rot_rows() reads
      pg.u0r = G.u0r; pg.u0i = G.u0i; pg.u1r = G.u1r; pg.u1i = G.u1i;
      pg.atol = s.atol;
      pg.simple = G.simple;
      const bool exact = s.exact_last && g == s.ng - 1;
      const u64 hG = G.hG;
      u64 gx[W], gz[W];
#pragma unroll
      for (int w = 0; w < W; ++w) { gx[w] = G.gx[w]; gz[w] = G.gz[w]; }
      const int yG = G.yG;
      int* add_w = &s_add[(g + 1) % 3];
      if (threadIdx.x == 0) s_add[(g + 2) % 3] = 0;
      int dlive = 0;
      for (int base = (int)(threadIdx.x & ~31u); base < n_pass; base += nt) {
        const int r = base + (int)lane;
        bool want_append = false;
        double2 app_c = make_double2(0.0, 0.0);
        ulonglong2 pk[W];  // partner key
        u64 h2 = 0;
        if (r < n_pass && m.died[r] == OBP_SEG_ALIVE) {
          const double2 c = m.coef[r];
          int sp = 0, E = yG;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            const ulonglong2 key = m.xz[w][r];
            const u64 mm = gx[w] | gz[w];
            sp += __popcll(key.x & gz[w]) + __popcll(key.y & gx[w]);
            E += __popcll(key.x & key.y & mm) - __popcll((key.x ^ gx[w]) & (key.y ^ gz[w]) & mm) + 2 * __popcll(key.y & gx[w]);
            pk[w] = make_ulonglong2(key.x ^ gx[w], key.y ^ gz[w]);
          }
          const int sgn = sp & 1;
          E &= 3;
          if (!exact && !sgn) {
            // commuting row, fast mode: only its own key can change (row filter)
            const RowSet rp = rot_rows_self(c, sgn, pg);
            if (rp.nself < 2) {
              if (rp.nself == 0 || is_small(rp.self, s.atol)) {
                m.died[r] = (unsigned char)g;
                dlive--;
              } else {
                m.coef[r] = rp.self;
              }
            }
          } else {
            const RowSet rp = rot_rows(c, E, sgn, pg);
            h2 = m.hash[r] ^ hG;
            bool fresh;
            const int j = seg_find<W>(m, idx_mask, h2, pk, n_pass, (unsigned)g, fresh);
            if (j >= 0 && !fresh && r < j) {
              // both rows live, this thread owns the pair (the row with the smaller id)
              const double2 cq = m.coef[j];
              const RowSet rq = rot_rows(cq, (4 - E) & 3, sgn, pg);
              const bool pres_p = rp.nself > 0 || rq.ncross > 0;
              const bool pres_q = rq.nself > 0 || rp.ncross > 0;
              double2 np_ = rp.self, nq_ = rq.self;
              if (rq.ncross) np_ = rp.nself ? cadd(np_, rq.cross) : rq.cross;
              if (rp.ncross) nq_ = rq.nself ? cadd(nq_, rp.cross) : rp.cross;
              if (exact) {
                rows_surv += rp.nself + rp.ncross + rq.nself + rq.ncross;
                keys_present += (unsigned)pres_p + (unsigned)pres_q;
              }
              if (pres_p && !is_small(np_, s.atol)) {
                m.coef[r] = np_;
              } else {
                if (pres_p && exact) { cancelled++; sum_re += np_.x; sum_im += np_.y; }
                m.died[r] = (unsigned char)g;
                dlive--;
              }
              if (pres_q && !is_small(nq_, s.atol)) {
                m.coef[j] = nq_;
              } else {
                if (pres_q && exact) { cancelled++; sum_re += nq_.x; sum_im += nq_.y; }
                m.died[j] = (unsigned char)g;
                dlive--;
              }
            } else if (j < 0) {
              // partner key absent: update this row, append the partner if it survives
              if (exact) {
                rows_surv += rp.nself + rp.ncross;
                keys_present += (unsigned)(rp.nself > 0) + (unsigned)(rp.ncross > 0);
              }
              if (rp.nself > 0 && !is_small(rp.self, s.atol)) {
                m.coef[r] = rp.self;
              } else {
                if (rp.nself > 0 && exact) { cancelled++; sum_re += rp.self.x; sum_im += rp.self.y; }
                m.died[r] = (unsigned char)g;
                dlive--;
              }
              if (rp.ncross > 0) {
                if (!is_small(rp.cross, s.atol)) {
                  want_append = true;
                  app_c = rp.cross;
                } else if (exact) {
                  cancelled++; sum_re += rp.cross.x; sum_im += rp.cross.y;
                }
              }
            }
            // else: the pair is (or was) handled by its owner
          }
        }
        // ---- append the new keys of this warp: ballot, one shared-memory atomicAdd per warp ------------------
        const unsigned bal = __ballot_sync(0xffffffffu, want_append);
        if (bal) {
          int wbase = 0;
          if (lane == 0) wbase = atomicAdd(add_w, __popc(bal));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
          if (want_append) {
            const int row = n_pass + wbase + __popc(bal & ((1u << lane) - 1u));
            if (row < CAP) {
#pragma unroll
              for (int w = 0; w < W; ++w) m.xz[w][row] = pk[w];
              m.coef[row] = app_c;
              m.hash[row] = h2;
              m.died[row] = (unsigned char)OBP_SEG_ALIVE;
              if (!seg_insert<W>(m, idx_mask, h2, row)) atomicOr(&res->overflow, 4u);
              dlive++;
            }
            // (a row beyond the partition's capacity is not stored: n_pass > CAP after this pass ends the run)
          }
        }
      }
      const int dl = (int)warp_sum_i64((long long)dlive);
      if (lane == 0 && dl) {
        atomicAdd(&m.dg[g], dl);
        atomicAdd(&s_live, dl);
      }
      __syncthreads();
      n_pass += *add_w;
      failed = n_pass > CAP;  // (uniform: every thread reads the same, stable counter)
    }
    if (failed) {
      if (threadIdx.x == 0) atomicOr(&res->overflow, 1u);
      __syncthreads();
      continue;
    }
    // ---- write the survivors out, compacted ----------------------------------------------------------------------
    if (threadIdx.x == 0) s_base = atomicAdd(&res->out_rows, (u64)s_live);
    __syncthreads();
    const u64 obase = s_base;
    if (obase + (u64)s_live > s.out_cap) {
      if (threadIdx.x == 0) atomicOr(&res->overflow, 2u);
      __syncthreads();
      continue;
    }
    for (int base = (int)(threadIdx.x & ~31u); base < n_pass; base += nt) {
      const int r = base + (int)lane;
      const bool alive = r < n_pass && m.died[r] == OBP_SEG_ALIVE;
      const unsigned bal = __ballot_sync(0xffffffffu, alive);
      if (!bal) continue;
      int wbase = 0;
      if (lane == 0) wbase = atomicAdd(&s_out, __popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (alive) {
        const u64 o = obase + (u64)(wbase + __popc(bal & ((1u << lane) - 1u)));
#pragma unroll
        for (int w = 0; w < W; ++w) out.xz[w][o] = m.xz[w][r];
        out.coef[o] = m.coef[r];
        out.hash[o] = m.hash[r];
      }
    }
    __syncthreads();  // the next partition reuses the shared arrays
  }
  // ---- counters ---------------------------------------------------------------------------------------------------
  __syncthreads();
  for (int k = threadIdx.x; k < s.ng; k += nt)
    if (m.dg[k]) atomicAdd(reinterpret_cast<unsigned long long*>(&res->delta[k]), (unsigned long long)(long long)m.dg[k]);
  if (s.exact_last) {
    const u64 a = warp_sum_u64(rows_surv), b = warp_sum_u64(keys_present), cz = warp_sum_u64(cancelled);
    const double sr = warp_sum_f64(sum_re), si = warp_sum_f64(sum_im);
    if (lane == 0) {
      OpStat* o = &t.ops[s.slot0 + s.ng - 1];
      if (a) atomicAdd(&o->rows_surv, a);
      if (b) atomicAdd(&o->keys_present, b);
      if (cz) { atomicAdd(&o->cancelled, cz); atomicAdd(&o->sum_re, sr); atomicAdd(&o->sum_im, si); }
    }
  }
}

// Per-gate live-row counts, the table state (only if the run succeeded) and the result block for the host.
// Slot slot0 + ng reports the outcome: n_out = 0 ok / overflow bits, rows_surv = output rows.
__global__ void __launch_bounds__(256) k_seg_finish(DevState* st, OpStat* ops, const SegRes* __restrict__ res, int slot0, int ng,
                                                    u64 n_live_in, u64* __restrict__ host_res, int n_res_ops) {
  if (threadIdx.x == 0) {
    long long run = (long long)n_live_in;
    for (int g = 0; g < ng; ++g) {
      run += res->delta[g];
      ops[slot0 + g].n_out = (u64)run;
    }
    OpStat* f = &ops[slot0 + ng];
    f->n_out = res->overflow;
    f->rows_surv = res->out_rows;
    if (res->overflow == 0u) {
      st->n_scan = res->out_rows;
      st->n_append = res->out_rows;
      st->n_live = res->out_rows;
    }
    __threadfence();
  }
  __syncthreads();
  if (host_res) {
    const u64* src = reinterpret_cast<const u64*>(st);  // DevState | start stamp | counter slots: one block
    const int n_words = (int)((OBP_RES_HDR + (size_t)n_res_ops * sizeof(OpStat)) / sizeof(u64));
    for (int k = threadIdx.x; k < n_words; k += blockDim.x) host_res[k] = src[k];
  }
}
