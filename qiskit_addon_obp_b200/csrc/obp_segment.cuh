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
// A partition's pass is a chain of dependent shared-memory accesses and fp64 operations (~1 us per gate) whatever its
// size, so the run is spread over MANY SMALL CTAs — 128 threads, four per SM, partitions of a few hundred rows —
// which overlap each other's latencies; the few partitions that outgrow a small CTA's shared memory (a big coset
// class) are listed and run again by a second launch of one 512-thread CTA per SM with the whole shared memory.
//
// Four launches per run: k_seg_count (partition id per live row + histogram; its last CTA turns the histogram into
// offsets), k_seg_scatter (row ids in partition order), k_seg_main<W, 128, 4> (the run), k_seg_main<W, 512, 1> (the
// listed partitions again; its last CTA writes the per-gate row counts, the table state and the host's result
// block).  One small upload (the gates) sits between the partition kernels and the run; nothing is cleared in between
// (the table's hash index is, before the run, when the run is to publish its output rows in it).
// The input table is never modified: if a partition outgrows even the big CTA (or the output its capacity) the run
// reports it and the host repeats it in shorter pieces (fewer generators: a finer partition map) or applies the same
// gates through the per-gate kernels instead.
// Opt-in (OBP_OPT_SEGMENT_MIXED): a run may also hold Clifford-class gates; they are applied to the rows of each
// partition in groups without a barrier, and the partition map vanishes on the generators pulled back through them.
#pragma once
#include "obp_kernels.cuh"

#define OBP_SEG_MAXG 250    // gates per run (a row records the gate that killed it in one byte)
#define OBP_SEG_MAXP 65536  // partitions per run
#define OBP_SEG_FBITS 16    // output bits of the partition map
#define OBP_SEG_ALIVE 255u
#define OBP_SEG_CLIFFORD 2

struct SegGate {  // 96 bytes
  u64 gx[2], gz[2];  // generator on table words 0 and 1 (zero where it is the identity)
  u64 hG;            // hash of the generator
  double u0r, u0i, u1r, u1i;
  int yG;
  int simple;        // rotation: 1 = u0 real and u1 imaginary, 0 = general; OBP_SEG_CLIFFORD: a Clifford-class gate —
                     // gx[0] = its 16-entry table, gx[1] = sign mask | nq << 16 | word a << 24 | bit a << 32 | word b << 40 |
                     // bit b << 48, u0r = its row scale (k_clifford_reg's gate, applied to the rows of the partition)
  // |c|^2 above which no raw row of a term can fall under the zero filter (1e-6 relative margin; rows below take the
  // exact path): thr_self for the two rows of the term's own key (|c||u0|^2, |c||u1|^2), thr_all for all four
  double thr_self, thr_all;
};



static_assert(sizeof(SegGate) == 96, "SegGate is staged with 16-byte copies");

struct SegP {
  int ng;            // gates
  int P;             // partitions
  int cap_rows;      // rows a partition may hold in shared memory
  int n_index;       // local index slots (power of two)
  int exact_last;    // the last gate reports the reference's simplify counters
  int slot0;         // counter slot of gate 0 (gate g: slot0 + g)
  int cap_rows_big;  // rows a partition may hold in the big CTAs of the second launch (a bigger one cannot be run at all)
  int pad2;
  int dbg;           // timing experiments (OBP_B200_SEG_DEBUG; results are wrong): 1 no row work, 2 no partner path
  int pad;
  u64 n_dense;       // rows of the input table
  u64 out_cap;       // rows the output columns hold
  double atol;
  u64 fmask[OBP_SEG_FBITS][4];  // partition map: bit j = parity of (x0 & m[0]) ^ (z0 & m[1]) ^ (x1 & m[2]) ^ (z1 & m[3])
};

struct SegRes {  // device-side result of a run (zeroed before)
  u64 out_rows;          // output cursor
  unsigned overflow;     // 1: a partition outgrew shared memory, 2: the output outgrew its columns, 4: a local index filled up
  unsigned n_redo;       // partitions that outgrew the small CTAs and were run again by the big ones
  unsigned ticket;       // last-block-done counter (k_seg_count, second k_seg_main)
  unsigned index_skipped; // the output outgrew the cleared part of the index: rows were not published
  long long delta[OBP_SEG_MAXG];  // change of the live-row count by every gate
};

struct SegOut {
  ulonglong2* xz[2];
  double2* coef;
  u64* hash;
  u64* index;      // the table's hash index, cleared for this run: the output stage publishes the rows it writes
  u64 index_mask;  // (nullptr: the index is rebuilt later, if anybody needs it)
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

// Partition id of every live row + histogram; the last CTA to finish turns the histogram into offsets (off[p],
// off[P] = live rows) and clears the scatter cursors.  CTA 0 also zeroes the run's counter slots.
template <int W>
__global__ void __launch_bounds__(OBP_BS) k_seg_count(Tab t, const __grid_constant__ SegP s, unsigned* __restrict__ part,
                                                      unsigned* __restrict__ cnt, unsigned* __restrict__ off,
                                                      unsigned* __restrict__ cur, SegRes* __restrict__ res) {
  if (blockIdx.x == 0) {
    u64* w = reinterpret_cast<u64*>(t.ops + s.slot0);
    const int n_words = (s.ng + 1) * (int)(sizeof(OpStat) / sizeof(u64));
    for (int k = threadIdx.x; k < n_words; k += OBP_BS) w[k] = 0;
  }
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
  __shared__ bool s_last;
  __shared__ unsigned s_tot[OBP_BS];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&res->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int P = s.P;
  const int per = (P + OBP_BS - 1) / OBP_BS;
  const int p0 = threadIdx.x * per, p1 = min(P, p0 + per);
  unsigned local = 0, biggest = 0;
  for (int p = p0; p < p1; ++p) {
    const unsigned c = __ldcg(&cnt[p]);
    local += c;
    biggest = max(biggest, c);
  }
  if (biggest > (unsigned)s.cap_rows_big) atomicOr(&res->overflow, 8u);  // hopeless: the launches that follow return at once
  s_tot[threadIdx.x] = local;
  __syncthreads();
  for (int d = 1; d < OBP_BS; d <<= 1) {
    const unsigned v = threadIdx.x >= (unsigned)d ? s_tot[threadIdx.x - d] : 0u;
    __syncthreads();
    s_tot[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = s_tot[threadIdx.x] - local;
  for (int p = p0; p < p1; ++p) {
    off[p] = run;
    cur[p] = 0;
    run += __ldcg(&cnt[p]);
  }
  if (threadIdx.x == OBP_BS - 1) {
    off[P] = s_tot[OBP_BS - 1];
    res->ticket = 0;
  }
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
  int* dg;              // per-gate change of the live-row count, summed over the CTA's finished partitions
};

// Row of the partition holding `key ^ flips` with hash h2 that was present when pass g started (alive, or killed
// during pass g), or -1.
template <int W>
__device__ __forceinline__ int seg_find(const SegSmem<W>& m, unsigned idx_mask, u64 h2, const ulonglong2* want, int n_pass,
                                        unsigned g, bool& fresh, unsigned s) {
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

// One launch of the run.  plist == nullptr: every partition p < s.P (the first launch, small CTAs, several per SM);
// else the n = *plist_n partitions listed in plist (the second launch: the partitions that outgrew the small CTAs'
// shared memory, one big CTA per SM).  A partition that does not fit is appended to redo_list (first launch) or
// fails the run (second launch, redo_list == nullptr).
struct SegFin {  // what the last CTA of the second launch needs for the run's bookkeeping
  DevState* st;
  unsigned* cnt;     // partition histogram: cleared for the next run
  u64 n_live_in;
  u64* host_res;     // the host's mapped result block
  int n_res_ops;
};

template <int W, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_seg_main(Tab t, SegOut out, const __grid_constant__ SegP s, const SegGate* __restrict__ gates,
           const unsigned* __restrict__ off, const unsigned* __restrict__ perm, SegRes* __restrict__ res,
           const unsigned* __restrict__ plist, const unsigned* __restrict__ plist_n, unsigned* __restrict__ redo_list,
           const SegFin fin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CAP = s.cap_rows;
  const int ngp = (s.ng + 1) & ~1;
  SegSmem<W> m;
  int* dgp;  // per-gate change of the live-row count inside the current partition
  SegGate* sgate;  // the run's gates, staged once per CTA (a pass starts with broadcast loads from here)
  {
    unsigned char* p = smem_raw;
#pragma unroll
    for (int w = 0; w < W; ++w) { m.xz[w] = reinterpret_cast<ulonglong2*>(p); p += (size_t)CAP * 16; }
    m.coef = reinterpret_cast<double2*>(p); p += (size_t)CAP * 16;
    sgate = reinterpret_cast<SegGate*>(p); p += (size_t)s.ng * sizeof(SegGate);
    m.hash = reinterpret_cast<u64*>(p); p += (size_t)CAP * 8;
    m.dg = reinterpret_cast<int*>(p); p += (size_t)ngp * 4;
    dgp = reinterpret_cast<int*>(p); p += (size_t)ngp * 4;
    m.idx = reinterpret_cast<unsigned short*>(p); p += (size_t)s.n_index * 2;
    m.died = p;
  }
  __shared__ int s_add[3];       // rows appended during a pass (rotating: read / written / cleared)
  __shared__ int s_live;         // live rows of the partition
  __shared__ int s_out;          // output cursor inside the partition
  __shared__ u64 s_base;         // first output row of the partition
  const unsigned idx_mask = (unsigned)s.n_index - 1u;
  const unsigned lane = threadIdx.x & 31u;
  // exact counters of the last gate: thread-local, over all partitions this CTA completes
  unsigned rows_surv = 0, keys_present = 0, cancelled = 0;
  double sum_re = 0.0, sum_im = 0.0;
  // a run that has already failed (a partition too big for any CTA, or a partition of the second launch) is not continued;
  // one thread looks, so that the whole CTA takes the same decision
  __shared__ int s_abort, s_nlist;
  if (threadIdx.x == 0) {
    s_abort = __ldcg(&res->overflow) != 0u;
    s_nlist = s_abort ? 0 : (plist ? (int)__ldcg(plist_n) : s.P);
  }
  auto redo = [&](int part) {
    if (threadIdx.x == 0) {
      if (redo_list) redo_list[atomicAdd(&res->n_redo, 1u)] = (unsigned)part;
      else atomicOr(&res->overflow, 1u);
    }
  };
  __syncthreads();
  const int n_list = s_nlist;
  // a CTA without a partition (every CTA of the second launch when nothing outgrew the small ones) stages nothing
  const bool has_work = (int)blockIdx.x < n_list;
  if (has_work) {
    for (int k = threadIdx.x; k < ngp; k += NT) m.dg[k] = 0;
    for (int k = threadIdx.x; k < s.ng * (int)(sizeof(SegGate) / 16); k += NT)
      reinterpret_cast<uint4*>(sgate)[k] = reinterpret_cast<const uint4*>(gates)[k];
  }
  __syncthreads();

  for (int it = blockIdx.x; it < n_list; it += gridDim.x) {
    const int part = plist ? (int)plist[it] : it;
    if (plist) {
      if (threadIdx.x == 0) s_abort = __ldcg(&res->overflow) != 0u;
      __syncthreads();
      const bool ab = s_abort != 0;
      __syncthreads();  // (nobody rewrites the flag before everybody has read it)
      if (ab) break;
    }
    const unsigned o0 = off[part];
    const int n_in = (int)(off[part + 1] - o0);
    if (n_in == 0) continue;
    if (n_in > CAP) {
      redo(part);
      continue;
    }
    // ---- gather the partition ----------------------------------------------------------------------------------
    for (int k = threadIdx.x; k < s.n_index / 4; k += NT) reinterpret_cast<u64*>(m.idx)[k] = 0ull;
    for (int k = threadIdx.x; k < ngp; k += NT) dgp[k] = 0;
    if (threadIdx.x == 0) {
      s_add[0] = 0; s_add[1] = 0; s_add[2] = 0;
      s_live = n_in;
      s_out = 0;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n_in; r += NT) {
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
    unsigned p_rows_surv = 0, p_keys_present = 0, p_cancelled = 0;
    double p_sum_re = 0.0, p_sum_im = 0.0;
    int n_pass = n_in;
    int add_i = 1;  // s_add slot the current pass appends to (the next one is cleared meanwhile)
    bool failed = false;
    for (int g = 0; g < s.ng && !failed; ++g) {
      int* add_w = &s_add[add_i];
      if (threadIdx.x == 0) s_add[add_i == 2 ? 0 : add_i + 1] = 0;
      if (sgate[g].simple == OBP_SEG_CLIFFORD) {
        // A group of consecutive Clifford-class gates: they map every row on its own (key bits through a 16-entry
        // table, a sign, the raw-row zero filter |c| * row_scale <= atol, utils/operations.py:103-105 +
        // simplify for a signed-permutation gate), so the whole group is ONE pass without a barrier in between.
        // The rows' stored hashes stay valid: the host re-expresses the hash frame (frame_update).
        int g2 = g + 1;
        while (g2 < s.ng && sgate[g2].simple == OBP_SEG_CLIFFORD) ++g2;
        for (int r = threadIdx.x; r < n_pass; r += NT) {
          if (m.died[r] != OBP_SEG_ALIVE) continue;
          double2 c = m.coef[r];
          ulonglong2 key[W];
#pragma unroll
          for (int w = 0; w < W; ++w) key[w] = m.xz[w][r];
          unsigned flip = 0;
          int dead_at = -1;
          for (int gg = g; gg < g2; ++gg) {
            const SegGate& C = sgate[gg];
            if (s.atol >= 0.0 && is_small_scaled(c, C.u0r, s.atol)) { dead_at = gg; break; }
            const u64 lut = C.gx[0], pos = C.gx[1];
            const unsigned nq = (unsigned)(pos >> 16) & 0xffu;
            const unsigned wa = (unsigned)(pos >> 24) & 0xffu, ba = (unsigned)(pos >> 32) & 0xffu;
            const unsigned wb = (unsigned)(pos >> 40) & 0xffu, bb = (unsigned)(pos >> 48) & 0xffu;
            const ulonglong2 ka = (W == 2 && wa) ? key[W - 1] : key[0];
            const ulonglong2 kb = (W == 2 && wb) ? key[W - 1] : key[0];
            unsigned v = (unsigned)((ka.x >> ba) & 1ull) | ((unsigned)((ka.y >> ba) & 1ull) << 1);
            if (nq == 2) v |= ((unsigned)((kb.x >> bb) & 1ull) << 2) | ((unsigned)((kb.y >> bb) & 1ull) << 3);
            const unsigned d = ((unsigned)(lut >> (4 * v)) & 15u) ^ v;
            flip ^= (unsigned)(pos >> v) & 1u;
            const u64 fxa = (u64)(d & 1u) << ba, fza = (u64)((d >> 1) & 1u) << ba;
            const u64 fxb = (u64)((d >> 2) & 1u) << bb, fzb = (u64)((d >> 3) & 1u) << bb;
            if (W == 2) {
              key[0].x ^= (wa ? 0ull : fxa) ^ (wb ? 0ull : fxb);
              key[0].y ^= (wa ? 0ull : fza) ^ (wb ? 0ull : fzb);
              key[W - 1].x ^= (wa ? fxa : 0ull) ^ (wb ? fxb : 0ull);
              key[W - 1].y ^= (wa ? fza : 0ull) ^ (wb ? fzb : 0ull);
            } else {
              key[0].x ^= fxa ^ fxb;
              key[0].y ^= fza ^ fzb;
            }
          }
          if (dead_at >= 0) {
            m.died[r] = (unsigned char)dead_at;
            atomicAdd(&dgp[dead_at], -1);
            atomicAdd(&s_live, -1);
          } else {
#pragma unroll
            for (int w = 0; w < W; ++w) m.xz[w][r] = key[w];
            if (flip & 1u) m.coef[r] = make_double2(-c.x, -c.y);
          }
        }
        __syncthreads();
        n_pass += *add_w;  // (nothing is appended by a Clifford group: the slot was cleared by the pass before)
        add_i = add_i == 2 ? 0 : add_i + 1;
        g = g2 - 1;
        continue;
      }
      // a warp whose rows all lie beyond the partition only keeps the barrier company (most warps of most partitions)
      if ((int)(threadIdx.x & ~31u) < n_pass) {
      const SegGate& G = sgate[g];
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
      const double thr_self = G.thr_self, thr_all = G.thr_all;
      int dlive = 0;
      for (int base = (int)(threadIdx.x & ~31u); base < n_pass; base += NT) {
        const int r = base + (int)lane;
        bool want_append = false;
        double2 app_c = make_double2(0.0, 0.0);
        ulonglong2 pk[W];  // partner key
        u64 h2 = 0;
        if (s.dbg != 1 && r < n_pass) {
          // the row's own data: independent shared-memory loads, issued together
          const unsigned died_r = m.died[r];
          const double2 c = m.coef[r];
          ulonglong2 key[W];
#pragma unroll
          for (int w = 0; w < W; ++w) key[w] = m.xz[w][r];
          const u64 h_own = m.hash[r];
          // commutation parity: popc(a) + popc(b) = popc(a ^ b) mod 2, so one 32-bit popcount of the folded words
          // decides it (POPC issues at a quarter of the ALU rate)
          u64 fold = 0;
#pragma unroll
          for (int w = 0; w < W; ++w) fold ^= (key[w].x & gz[w]) ^ (key[w].y & gx[w]);
          const int sgn = s.dbg == 2 ? 0 : (__popc((unsigned)fold ^ (unsigned)(fold >> 32)) & 1);
          const double m2 = fma(c.x, c.x, c.y * c.y);
          if (died_r != OBP_SEG_ALIVE) {
            // dead row
          } else if (!exact && !sgn) {
            // commuting row, fast mode: only its own key can change (row filter) — and only a tiny coefficient does
            if (!(m2 > thr_self)) {
              const RowSet rp = rot_rows_self(c, sgn, pg);
              if (rp.nself < 2) {
                if (rp.nself == 0 || is_small(rp.self, s.atol)) {
                  m.died[r] = (unsigned char)g;
                  dlive--;
                } else {
                  m.coef[r] = rp.self;
                }
              }
            }
          } else {
            int E = yG;
#pragma unroll
            for (int w = 0; w < W; ++w) {
              const u64 mm = gx[w] | gz[w];
              E += __popcll(key[w].x & key[w].y & mm) - __popcll((key[w].x ^ gx[w]) & (key[w].y ^ gz[w]) & mm) +
                   2 * __popcll(key[w].y & gx[w]);
              pk[w] = make_ulonglong2(key[w].x ^ gx[w], key[w].y ^ gz[w]);
            }
            E &= 3;
            h2 = h_own ^ hG;
            // partner lookup: the first slot of the probe chain inline (the candidate's hash, key, state and coefficient
            // are independent loads), the rest of the chain only on a mismatch
            int j = -1;
            bool fresh = false;
            double2 cq = make_double2(0.0, 0.0);
            {
              const unsigned s0 = (unsigned)h2 & idx_mask;
              const unsigned v0 = m.idx[s0];
              if (v0 != 0u) {
                const int j0 = (int)v0 - 1;
                const u64 hj = m.hash[j0];
                ulonglong2 kj[W];
#pragma unroll
                for (int w = 0; w < W; ++w) kj[w] = m.xz[w][j0];
                const unsigned dj = m.died[j0];
                const double2 cj = m.coef[j0];
                bool eq = j0 < n_pass && hj == h2;
#pragma unroll
                for (int w = 0; w < W; ++w) eq = eq && kj[w].x == pk[w].x && kj[w].y == pk[w].y;
                if (eq && (dj == OBP_SEG_ALIVE || dj == (unsigned)g)) {
                  j = j0;
                  fresh = dj == (unsigned)g;
                  cq = cj;
                } else {
                  j = seg_find<W>(m, idx_mask, h2, pk, n_pass, (unsigned)g, fresh, (s0 + 1) & idx_mask);
                  if (j >= 0) cq = m.coef[j];
                }
              }
            }
            if (j >= 0 && !fresh && r < j) {
              // both rows live, this thread owns the pair (the row with the smaller id)
              const bool big = m2 > thr_all && fma(cq.x, cq.x, cq.y * cq.y) > thr_all;
              const RowSet rp = big ? rot_rows_nofilter(c, E, sgn, pg) : rot_rows(c, E, sgn, pg);
              const RowSet rq = big ? rot_rows_nofilter(cq, (4 - E) & 3, sgn, pg) : rot_rows(cq, (4 - E) & 3, sgn, pg);
              const bool pres_p = rp.nself > 0 || rq.ncross > 0;
              const bool pres_q = rq.nself > 0 || rp.ncross > 0;
              double2 np_ = rp.self, nq_ = rq.self;
              if (rq.ncross) np_ = rp.nself ? cadd(np_, rq.cross) : rq.cross;
              if (rp.ncross) nq_ = rq.nself ? cadd(nq_, rp.cross) : rp.cross;
              if (exact) {
                p_rows_surv += rp.nself + rp.ncross + rq.nself + rq.ncross;
                p_keys_present += (unsigned)pres_p + (unsigned)pres_q;
              }
              if (pres_p && !is_small(np_, s.atol)) {
                m.coef[r] = np_;
              } else {
                if (pres_p && exact) { p_cancelled++; p_sum_re += np_.x; p_sum_im += np_.y; }
                m.died[r] = (unsigned char)g;
                dlive--;
              }
              if (pres_q && !is_small(nq_, s.atol)) {
                m.coef[j] = nq_;
              } else {
                if (pres_q && exact) { p_cancelled++; p_sum_re += nq_.x; p_sum_im += nq_.y; }
                m.died[j] = (unsigned char)g;
                dlive--;
              }
            } else if (j < 0) {
              // partner key absent: update this row, append the partner if it survives
              const RowSet rp = m2 > thr_all ? rot_rows_nofilter(c, E, sgn, pg) : rot_rows(c, E, sgn, pg);
              if (exact) {
                p_rows_surv += rp.nself + rp.ncross;
                p_keys_present += (unsigned)(rp.nself > 0) + (unsigned)(rp.ncross > 0);
              }
              if (rp.nself > 0 && !is_small(rp.self, s.atol)) {
                m.coef[r] = rp.self;
              } else {
                if (rp.nself > 0 && exact) { p_cancelled++; p_sum_re += rp.self.x; p_sum_im += rp.self.y; }
                m.died[r] = (unsigned char)g;
                dlive--;
              }
              if (rp.ncross > 0) {
                if (!is_small(rp.cross, s.atol)) {
                  want_append = true;
                  app_c = rp.cross;
                } else if (exact) {
                  p_cancelled++; p_sum_re += rp.cross.x; p_sum_im += rp.cross.y;
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
            // (a row beyond the partition's capacity is not stored: n_pass > CAP after this pass ends the attempt)
          }
        }
      }
      const int dl = __reduce_add_sync(0xffffffffu, dlive);
      if (lane == 0 && dl) {
        atomicAdd(&dgp[g], dl);
        atomicAdd(&s_live, dl);
      }
      }
      __syncthreads();
      n_pass += *add_w;
      add_i = add_i == 2 ? 0 : add_i + 1;
      failed = n_pass > CAP;  // (uniform: every thread reads the same, stable counter)
    }
    if (failed) {
      redo(part);
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
    // publish the rows in the table's index while it stays at most half full
    const bool publish = out.index != nullptr && 2 * (obase + (u64)s_live) <= out.index_mask + 1;
    if (out.index != nullptr && !publish && threadIdx.x == 0) res->index_skipped = 1u;
    for (int base = (int)(threadIdx.x & ~31u); base < n_pass; base += NT) {
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
        const u64 h = m.hash[r];
        out.hash[o] = h;
        if (publish) {
          const u64 entry = ((h >> 32) << 32) | o;
          u64 sl = h & out.index_mask;
          for (u64 probes = 0; probes <= out.index_mask; ++probes) {
            if (atomicCAS(&out.index[sl], OBP_EMPTY, entry) == OBP_EMPTY) break;
            sl = (sl + 1) & out.index_mask;
          }
        }
      }
    }
    // the partition is done: its counters join the CTA's
    for (int k = threadIdx.x; k < s.ng; k += NT) m.dg[k] += dgp[k];
    rows_surv += p_rows_surv;
    keys_present += p_keys_present;
    cancelled += p_cancelled;
    sum_re += p_sum_re;
    sum_im += p_sum_im;
    __syncthreads();  // the next partition reuses the shared arrays
  }
  // ---- counters ---------------------------------------------------------------------------------------------------
  __syncthreads();
  if (has_work) {
    for (int k = threadIdx.x; k < s.ng; k += NT)
      if (m.dg[k]) atomicAdd(reinterpret_cast<unsigned long long*>(&res->delta[k]), (unsigned long long)(long long)m.dg[k]);
  }
  if (s.exact_last && has_work) {
    const u64 a = warp_sum_u64(rows_surv), b = warp_sum_u64(keys_present), cz = warp_sum_u64(cancelled);
    const double sr = warp_sum_f64(sum_re), si = warp_sum_f64(sum_im);
    if (lane == 0) {
      OpStat* o = &t.ops[s.slot0 + s.ng - 1];
      if (a) atomicAdd(&o->rows_surv, a);
      if (b) atomicAdd(&o->keys_present, b);
      if (cz) { atomicAdd(&o->cancelled, cz); atomicAdd(&o->sum_re, sr); atomicAdd(&o->sum_im, si); }
    }
  }
  if (fin.st == nullptr || NT < 256) return;  // (the bookkeeping below wants one thread per gate)
  // ---- second launch: the last CTA to finish does the run's bookkeeping ----------------------------------------------
  // Per-gate live-row counts, the table state (only if the run succeeded), the result block for the host; the
  // partition scratch is left clean for the next run.  Slot slot0 + ng reports the outcome: n_out = 0 ok / overflow
  // bits, rows_surv = output rows, keys_present = partitions that needed the big CTAs.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&res->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  OpStat* ops = t.ops + s.slot0;
  {
    // live rows after every gate: inclusive scan of the per-gate changes (one gate per thread, ng <= 250 < NT)
    long long* sc = reinterpret_cast<long long*>(smem_raw);
    long long v = 0;
    if ((int)threadIdx.x < s.ng) v = (long long)__ldcg(reinterpret_cast<const unsigned long long*>(&res->delta[threadIdx.x]));
    sc[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < NT; d <<= 1) {
      const long long a = (int)threadIdx.x >= d ? sc[threadIdx.x - d] : 0ll;
      __syncthreads();
      sc[threadIdx.x] += a;
      __syncthreads();
    }
    if ((int)threadIdx.x < s.ng) ops[threadIdx.x].n_out = (u64)((long long)fin.n_live_in + sc[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    const unsigned ovf = __ldcg(&res->overflow);
    const u64 n_rows = __ldcg(reinterpret_cast<const unsigned long long*>(&res->out_rows));
    OpStat* f = &ops[s.ng];
    f->n_out = ovf;
    f->rows_surv = n_rows;
    f->keys_present = __ldcg(&res->n_redo);
    f->cancelled = __ldcg(&res->index_skipped);
    if (ovf == 0u) {
      fin.st->n_scan = n_rows;
      fin.st->n_append = n_rows;
      fin.st->n_live = n_rows;
    }
  }
  __threadfence();
  __syncthreads();
  if (fin.host_res) {
    const u64* src = reinterpret_cast<const u64*>(fin.st);  // DevState | start stamp | counter slots: one block
    const int n_words = (int)((OBP_RES_HDR + (size_t)fin.n_res_ops * sizeof(OpStat)) / sizeof(u64));
    for (int k = threadIdx.x; k < n_words; k += NT) fin.host_res[k] = __ldcg(&src[k]);
  }
  for (int k = threadIdx.x; k <= s.P; k += NT) fin.cnt[k] = 0u;
  {
    u64* w = reinterpret_cast<u64*>(res);
    for (int k = threadIdx.x; k < (int)(sizeof(SegRes) / sizeof(u64)); k += NT) w[k] = 0ull;
  }
}

