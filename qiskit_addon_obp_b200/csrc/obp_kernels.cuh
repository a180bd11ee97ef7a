// obp_kernels.cuh — the sm_100a kernels of the operator-backpropagation hot path.
//
// All table kernels are persistent grid-stride kernels (grid = a multiple of the 148 SMs) that
// read the current row count from device memory, so a whole slice of gates is enqueued without a
// host round trip.  Every kernel ends with finish_kernel(): the last CTA publishes the append
// cursor as the next kernel's scan bound.
#pragma once
#include "obp_device.cuh"

// ================================================================================================
// k_rotate — U = u0*I + u1*G  (rx/ry/rz/rxx/ryy/rzz/t/p...), conjugation + merge + trim in one pass
//
// Replaces, per gate:  op.compose(U, front=True).compose(U^dag)  (utils/operations.py:103-105)
//                      followed by simplify()                     (utils/simplify.py:115-165).
//
// The reference materialises 4 raw rows per term P (coefficient c):
//   (0,0) key P      r00 = (c*u0)*conj(u0)
//   (1,1) key P      r11 = (-1)^s (c*u1)*conj(u1)            s = symplectic product <P,G>
//   (0,1) key P^G    r01 = ((c*u0)*conj(u1)) * (-i)^E        E = y(P)+y(G)-y(P^G)+2|zP&xG| mod 4
//   (1,0) key P^G    r10 = ((c*u1)*conj(u0)) * (-i)^(E+2s)
// drops every raw row with |r| <= atol, sums the rest per key, drops keys with |sum| <= atol.
// P and Q = P^G have the same s, and E(Q) = -E(P) mod 4, so a rotation is a set of independent
// 2x2 updates on (c_P, c_Q) pairs plus insertions of keys that were absent.  One thread per row:
//   commuting rows (s = 0) keep their key and (up to the row filter) their coefficient;
//   anticommuting rows look their partner up in the index.  If the partner is live the pair is
//   updated once, by the row whose canonical generator bit is 0 ("owner"); otherwise the row
//   updates itself and appends the partner key (warp ballot + block prefix scan, one atomicAdd
//   per CTA on the append cursor, then an atomicCAS slot claim in the index).
// No floating-point atomics: every coefficient is written by exactly one thread.
// With p.exact every live row does the partner lookup, which yields the reference's
// num_unique / num_duplicate / num_trimmed / sum_trimmed counters exactly.
// ================================================================================================
struct RotP {
  int ngw;          // generator words
  int gw[4];        // their word indices
  u64 gx[4], gz[4];
  u64 hG;           // hash of the generator
  int canon_slot;   // ownership bit: generator word slot, x or z plane, bit mask
  int canon_z;
  u64 canon_bit;
  int yG;           // |gx & gz|
  double u0r, u0i, u1r, u1i;
  double atol;
  u64 epoch;
  int exact;
  int slot;
  int simple;  // u0 real and u1 purely imaginary
};

struct RowSet {  // surviving raw rows of one term, already summed per key
  double2 self, cross;
  int nself, ncross;
};

// The four raw rows of one term.  SIMPLE: u0 is real and u1 purely imaginary (every exp(-i t/2 G)
// rotation): the products below are what cmul() yields with the zero components dropped, i.e. the
// same values bit for bit (up to the sign of a zero) at a third of the fp64 instructions.
// CROSS = false skips the two rows of key P^G (a commuting row in fast mode never needs them).
template <bool SIMPLE, bool CROSS>
__device__ __forceinline__ RowSet rot_rows_t(double2 c, int E, int s, const RotP& p) {
  double2 r00, r11, r01 = make_double2(0.0, 0.0), r10 = make_double2(0.0, 0.0);
  if (SIMPLE) {
    const double a = p.u0r, b = p.u1i;
    const double2 t0 = make_double2(__dmul_rn(c.x, a), __dmul_rn(c.y, a));
    const double2 t1 = make_double2(-__dmul_rn(c.y, b), __dmul_rn(c.x, b));
    r00 = make_double2(__dmul_rn(t0.x, a), __dmul_rn(t0.y, a));
    r11 = make_double2(__dmul_rn(t1.y, b), -__dmul_rn(t1.x, b));
    if (CROSS) {
      r01 = make_double2(__dmul_rn(t0.y, b), -__dmul_rn(t0.x, b));
      r10 = make_double2(__dmul_rn(t1.x, a), __dmul_rn(t1.y, a));
    }
  } else {
    const double2 u0 = make_double2(p.u0r, p.u0i), u1 = make_double2(p.u1r, p.u1i);
    const double2 t0 = cmul(c, u0), t1 = cmul(c, u1);
    r00 = cmul(t0, cconj(u0));
    r11 = cmul(t1, cconj(u1));
    if (CROSS) {
      r01 = cmul(t0, cconj(u1));
      r10 = cmul(t1, cconj(u0));
    }
  }
  if (s) r11 = make_double2(-r11.x, -r11.y);
  RowSet r;
  r.self = make_double2(0.0, 0.0);
  r.cross = make_double2(0.0, 0.0);
  r.nself = 0;
  r.ncross = 0;
  if (!is_small(r00, p.atol)) { r.self = r00; r.nself = 1; }
  if (!is_small(r11, p.atol)) { r.self = r.nself ? cadd(r.self, r11) : r11; r.nself++; }
  if (CROSS) {
    r01 = mul_mi_pow(r01, E);
    r10 = mul_mi_pow(r10, E + 2 * s);
    if (!is_small(r01, p.atol)) { r.cross = r01; r.ncross = 1; }
    if (!is_small(r10, p.atol)) { r.cross = r.ncross ? cadd(r.cross, r10) : r10; r.ncross++; }
  }
  return r;
}

// The same four raw rows when the caller knows that none of them can fall under the zero filter (|c| times the
// smallest factor is above atol with a margin): the same products and the same sums, no magnitude tests.
template <bool SIMPLE>
__device__ __forceinline__ RowSet rot_rows_nofilter_t(double2 c, int E, int s, const RotP& p) {
  double2 r00, r11, r01, r10;
  if (SIMPLE) {
    const double a = p.u0r, b = p.u1i;
    const double2 t0 = make_double2(__dmul_rn(c.x, a), __dmul_rn(c.y, a));
    const double2 t1 = make_double2(-__dmul_rn(c.y, b), __dmul_rn(c.x, b));
    r00 = make_double2(__dmul_rn(t0.x, a), __dmul_rn(t0.y, a));
    r11 = make_double2(__dmul_rn(t1.y, b), -__dmul_rn(t1.x, b));
    r01 = make_double2(__dmul_rn(t0.y, b), -__dmul_rn(t0.x, b));
    r10 = make_double2(__dmul_rn(t1.x, a), __dmul_rn(t1.y, a));
  } else {
    const double2 u0 = make_double2(p.u0r, p.u0i), u1 = make_double2(p.u1r, p.u1i);
    const double2 t0 = cmul(c, u0), t1 = cmul(c, u1);
    r00 = cmul(t0, cconj(u0));
    r11 = cmul(t1, cconj(u1));
    r01 = cmul(t0, cconj(u1));
    r10 = cmul(t1, cconj(u0));
  }
  if (s) r11 = make_double2(-r11.x, -r11.y);
  RowSet r;
  r.self = cadd(r00, r11);
  r.nself = 2;
  r.cross = cadd(mul_mi_pow(r01, E), mul_mi_pow(r10, E + 2 * s));
  r.ncross = 2;
  return r;
}
__device__ __forceinline__ RowSet rot_rows_nofilter(double2 c, int E, int s, const RotP& p) {
  return p.simple ? rot_rows_nofilter_t<true>(c, E, s, p) : rot_rows_nofilter_t<false>(c, E, s, p);
}

__device__ __forceinline__ RowSet rot_rows(double2 c, int E, int s, const RotP& p) {
  return p.simple ? rot_rows_t<true, true>(c, E, s, p) : rot_rows_t<false, true>(c, E, s, p);
}
__device__ __forceinline__ RowSet rot_rows_self(double2 c, int s, const RotP& p) {
  return p.simple ? rot_rows_t<true, false>(c, 0, s, p) : rot_rows_t<false, false>(c, 0, s, p);
}

// generator words for table word w (0 if G is identity there)
__device__ __forceinline__ void gen_word(const RotP& p, int w, u64& gx, u64& gz) {
  gx = 0; gz = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < p.ngw && p.gw[k] == w) { gx = p.gx[k]; gz = p.gz[k]; }
}

// Find the row holding key(i) ^ G.  Returns its row id or -1.  *fresh is set when the row was
// killed during this epoch (it was live when the kernel started).
__device__ __forceinline__ long long find_partner(const Tab& t, const RotP& p, u64 i, u64 h2,
                                                  u64 n_scan, double2& cq, bool& fresh) {
  const u64 fp = h2 >> 32;
  u64 s = h2 & t.mask;
  fresh = false;
  for (u64 probes = 0; probes <= t.mask; ++probes) {
    const u64 e = ld_volatile(&t.index[s]);
    if (e == OBP_EMPTY) return -1;
    if ((e >> 32) == fp) {
      const u64 j = e & 0xffffffffull;
      if (j < n_scan) {
        // the 32-bit fingerprint already matched: compare the full key (all words are independent
        // loads, issued together with the coefficient) instead of chasing hash[j] first
        const double2 cj = t.coef[j];
        bool eq = true;
        for (int w = 0; w < t.W && eq; ++w) {
          u64 gx, gz;
          gen_word(p, w, gx, gz);
          const ulonglong2 ki = t.xz[w][i], kj = t.xz[w][j];
          eq = (kj.x == (ki.x ^ gx)) && (kj.y == (ki.y ^ gz));
        }
        if (eq) {
          cq = cj;
          if (is_live(cq)) return (long long)j;
          if (tomb_epoch(cq) == p.epoch) { fresh = true; return (long long)j; }
          // stale tombstone of the same key: keep probing for a live duplicate
        }
      }
    }
    s = (s + 1) & t.mask;
  }
  return -1;
}

#ifndef OBP_ROT_MINB
#define OBP_ROT_MINB 4  // resident CTAs per SM the register allocation is held to
#endif
// Body of the rotation pass over rows [0, n); `live_ctr` receives the change of the live-row count.
__device__ __forceinline__ void rotate_body(const Tab& t, const RotP& p, const u64 n, u64* live_ctr) {
  const unsigned lane = threadIdx.x & 31u;
  // per-thread counters
  unsigned rows_surv = 0, keys_present = 0, cancelled = 0;
  int dlive = 0;
  double sum_re = 0.0, sum_im = 0.0;
  bool overflow = false;

  // Warps run independently (no CTA barrier): a warp whose rows need long partner probes does not
  // hold back the others.  Rows are taken in warp-sized, fully coalesced chunks.
  // The three streaming loads of a row (coefficient, first generator word, hash) are independent and
  // are issued together, one iteration ahead of their use (software pipelining): the dependent chain
  // of an anticommuting row is then index slot -> partner key/coefficient only.
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const ulonglong2* __restrict__ col0 = t.xz[p.gw[0]];
  u64 i_next = ((u64)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) + lane;
  double2 c_next = make_double2(0.0, 0.0);
  ulonglong2 k_next = make_ulonglong2(0ull, 0ull);
  u64 h_next = 0;
  // (first chunk: guarded by the capacity, not by n — inside the program kernel n has just been loaded from
  // device memory, and these loads then go out together with it instead of one round trip later; rows in
  // [n, cap) are allocated memory whose content is never used)
  if (i_next < t.cap) { c_next = t.coef[i_next]; k_next = col0[i_next]; h_next = t.hash[i_next]; }
  for (u64 base = ((u64)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)); base < n; base += stride) {
    const u64 i = base + lane;
    const double2 c = c_next;
    const ulonglong2 key0 = k_next;
    const u64 h_own = h_next;
    i_next = i + stride;
    if (i_next < n) { c_next = t.coef[i_next]; k_next = col0[i_next]; h_next = t.hash[i_next]; }
    bool want_append = false;
    double2 app_c = make_double2(0.0, 0.0);
    if (i < n) {
      if (is_live(c)) {
        // commutation parity and phase exponent from the generator's words only
        int sp = 0, E = p.yG, own_bit = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < p.ngw) {
            const ulonglong2 key = k == 0 ? key0 : t.xz[p.gw[k]][i];
            const u64 m = p.gx[k] | p.gz[k];
            sp += __popcll(key.x & p.gz[k]) + __popcll(key.y & p.gx[k]);
            E += __popcll(key.x & key.y & m) - __popcll((key.x ^ p.gx[k]) & (key.y ^ p.gz[k]) & m) +
                 2 * __popcll(key.y & p.gx[k]);
            if (k == p.canon_slot) own_bit = ((p.canon_z ? key.y : key.x) & p.canon_bit) != 0;
          }
        }
        const int s = sp & 1;
        E &= 3;
        if (!p.exact && !s) {
          // commuting row, fast mode: only its own key can change (row filter)
          const RowSet rp = rot_rows_self(c, s, p);
          if (rp.nself < 2) {
            if (rp.nself == 0 || is_small(rp.self, p.atol)) {
              t.coef[i] = tombstone(p.epoch);
              dlive--;
            } else {
              t.coef[i] = rp.self;
            }
          }
        } else {
          const RowSet rp = rot_rows(c, E, s, p);
          const u64 h2 = h_own ^ p.hG;
          double2 cq;
          bool fresh;
          const long long j = find_partner(t, p, i, h2, n, cq, fresh);
          if (j >= 0 && !fresh && !own_bit) {
            // both rows live, this thread owns the pair
            const RowSet rq = rot_rows(cq, (4 - E) & 3, s, p);
            const bool pres_p = rp.nself > 0 || rq.ncross > 0;
            const bool pres_q = rq.nself > 0 || rp.ncross > 0;
            double2 np_ = rp.self, nq_ = rq.self;
            if (rq.ncross) np_ = rp.nself ? cadd(np_, rq.cross) : rq.cross;
            if (rp.ncross) nq_ = rq.nself ? cadd(nq_, rp.cross) : rp.cross;
            rows_surv += rp.nself + rp.ncross + rq.nself + rq.ncross;
            keys_present += (unsigned)pres_p + (unsigned)pres_q;
            if (pres_p && !is_small(np_, p.atol)) {
              t.coef[i] = np_;
            } else {
              if (pres_p) { cancelled++; sum_re += np_.x; sum_im += np_.y; }
              t.coef[i] = tombstone(p.epoch);
              dlive--;
            }
            if (pres_q && !is_small(nq_, p.atol)) {
              t.coef[j] = nq_;
            } else {
              if (pres_q) { cancelled++; sum_re += nq_.x; sum_im += nq_.y; }
              t.coef[j] = tombstone(p.epoch);
              dlive--;
            }
          } else if (j < 0) {
            // partner key absent: update this row, append the partner if it survives
            rows_surv += rp.nself + rp.ncross;
            keys_present += (unsigned)(rp.nself > 0) + (unsigned)(rp.ncross > 0);
            if (rp.nself > 0 && !is_small(rp.self, p.atol)) {
              if (rp.self.x != c.x || rp.self.y != c.y) t.coef[i] = rp.self;
            } else {
              if (rp.nself > 0) { cancelled++; sum_re += rp.self.x; sum_im += rp.self.y; }
              t.coef[i] = tombstone(p.epoch);
              dlive--;
            }
            if (rp.ncross > 0) {
              if (!is_small(rp.cross, p.atol)) {
                want_append = true;
                app_c = rp.cross;
              } else {
                cancelled++; sum_re += rp.cross.x; sum_im += rp.cross.y;
              }
            }
          }
          // else: the pair is (or was) handled by its owner
        }
      }
    }
    // ---- append the new keys of this warp: ballot, one atomicAdd on the cursor per warp ---------
    const unsigned bal = __ballot_sync(0xffffffffu, want_append);
    if (bal) {
      u64 wbase = 0;
      if (lane == 0) wbase = atomicAdd(&t.st->n_append, (u64)__popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (want_append) {
        const u64 row = wbase + __popc(bal & ((1u << lane) - 1u));
        if (row < t.cap) {
          for (int w = 0; w < t.W; ++w) {
            u64 gx, gz;
            gen_word(p, w, gx, gz);
            ulonglong2 key = t.xz[w][i];
            key.x ^= gx;
            key.y ^= gz;
            t.xz[w][row] = key;
          }
          const u64 h2 = h_own ^ p.hG;
          t.coef[row] = app_c;
          t.hash[row] = h2;
          if (!index_insert(t, h2, row)) atomicOr(&t.st->error, OBP_ERR_BIT_INDEX);
          dlive++;
        } else {
          overflow = true;
        }
      }
    }
  }
  if (overflow) atomicOr(&t.st->error, OBP_ERR_BIT_CAP);
  // ---- counters: warp shuffle reduction, one atomic per warp ------------------------------------
  const long long dl = warp_sum_i64((long long)dlive);
  if (p.exact) {
    const u64 a = warp_sum_u64(rows_surv), b = warp_sum_u64(keys_present), cz = warp_sum_u64(cancelled);
    const double sr = warp_sum_f64(sum_re), si = warp_sum_f64(sum_im);
    if (lane == 0) {
      OpStat* o = &t.ops[p.slot];
      if (a) atomicAdd(&o->rows_surv, a);
      if (b) atomicAdd(&o->keys_present, b);
      if (cz) { atomicAdd(&o->cancelled, cz); atomicAdd(&o->sum_re, sr); atomicAdd(&o->sum_im, si); }
    }
  }
  if (lane == 0 && dl) atomicAdd(live_ctr, (u64)dl);
}

__global__ void __launch_bounds__(OBP_BS, OBP_ROT_MINB) k_rotate(Tab t, const __grid_constant__ RotP p) {
  rotate_body(t, p, t.st->n_scan, &t.st->n_live);
  finish_kernel(t, p.slot);
}

// ================================================================================================
// k_clifford_* — a batch of gates whose Pauli-transfer matrix is a signed permutation (h, s, x, y,
// z, sx, cx, cz, swap, iswap, ecr, rzz(+-pi/2), ...), applied to each term in registers in one
// pass (k_clifford_reg: tables of <= 128 qubits; k_clifford_stream further down: wider tables).
// Replaces compose+simplify for those gates: the key moves to its image, the coefficient picks up
// the sign; all k*k raw rows of a term have magnitude |c|*row_scale, so the term vanishes iff that
// is <= atol.  Bijection on keys: no merge is needed and the row's stored hash stays
// valid because the host re-expresses the hash frame (frame_update).
// ================================================================================================
struct CliffGate {  // 16 bytes
  u64 lut;
  unsigned short sign;           // bit v: U^dag sigma_v U = -sigma_lut[v]
  unsigned char sa, ba, sb, bb;  // word slot / bit of qubit a and b (stream LOAD/STORE: ba = word index)
  unsigned char nq;
  unsigned char pad;
};
static_assert(sizeof(CliffGate) == 16, "CliffGate must stay 16 bytes");

struct CliffBatch {
  int ng;
  int nslots;                       // touched words
  int word_of_slot[OBP_MAX_WORDS];
  double scale;                     // min row_scale over the batch
  double atol;
  u64 epoch;
  int slot;
  CliffGate g[OBP_CLIFF_BATCH];
};

// one gate on local bit values; returns the flip mask d = new ^ old and updates sgn
__device__ __forceinline__ unsigned cliff_apply_bits(const CliffGate& g, u64 xa, u64 za, u64 xb, u64 zb,
                                                     unsigned& sgn) {
  unsigned v = (unsigned)((xa >> g.ba) & 1ull) | ((unsigned)((za >> g.ba) & 1ull) << 1);
  if (g.nq == 2) v |= ((unsigned)((xb >> g.bb) & 1ull) << 2) | ((unsigned)((zb >> g.bb) & 1ull) << 3);
  const unsigned nv = (unsigned)(g.lut >> (4 * v)) & 15u;
  sgn ^= (g.sign >> v) & 1u;
  return nv ^ v;
}

template <int W>
__device__ __forceinline__ void clifford_reg_body(const Tab& t, const CliffBatch& b, const u64 n, u64* live_ctr) {
  static_assert(W == 1 || W == 2, "register variant handles W <= 2");
  int deaths = 0;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const double2 c = t.coef[i];
    if (!is_live(c)) continue;
    if (is_small_scaled(c, b.scale, b.atol)) {
      t.coef[i] = tombstone(b.epoch);
      deaths++;
      continue;
    }
    u64 x0, z0, x1 = 0, z1 = 0;
    { const ulonglong2 k0 = t.xz[0][i]; x0 = k0.x; z0 = k0.y; }
    if (W == 2) { const ulonglong2 k1 = t.xz[1][i]; x1 = k1.x; z1 = k1.y; }
    const u64 ox0 = x0, oz0 = z0, ox1 = x1, oz1 = z1;
    unsigned sgn = 0;
    for (int k = 0; k < b.ng; ++k) {
      const CliffGate g = b.g[k];
      const u64 xa = (W == 2 && g.sa) ? x1 : x0, za = (W == 2 && g.sa) ? z1 : z0;
      const u64 xb = (W == 2 && g.sb) ? x1 : x0, zb = (W == 2 && g.sb) ? z1 : z0;
      const unsigned d = cliff_apply_bits(g, xa, za, xb, zb, sgn);
      const u64 fxa = (u64)(d & 1u) << g.ba, fza = (u64)((d >> 1) & 1u) << g.ba;
      const u64 fxb = (u64)((d >> 2) & 1u) << g.bb, fzb = (u64)((d >> 3) & 1u) << g.bb;
      if (W == 2) {
        x0 ^= (g.sa ? 0ull : fxa) ^ (g.sb ? 0ull : fxb);
        z0 ^= (g.sa ? 0ull : fza) ^ (g.sb ? 0ull : fzb);
        x1 ^= (g.sa ? fxa : 0ull) ^ (g.sb ? fxb : 0ull);
        z1 ^= (g.sa ? fza : 0ull) ^ (g.sb ? fzb : 0ull);
      } else {
        x0 ^= fxa ^ fxb;
        z0 ^= fza ^ fzb;
      }
    }
    if (x0 != ox0 || z0 != oz0) t.xz[0][i] = make_ulonglong2(x0, z0);
    if (W == 2 && (x1 != ox1 || z1 != oz1)) t.xz[1][i] = make_ulonglong2(x1, z1);
    if (sgn) t.coef[i] = make_double2(-c.x, -c.y);
  }
  const long long dl = warp_sum_i64((long long)deaths);
  if ((threadIdx.x & 31) == 0 && dl) atomicAdd(live_ctr, (u64)(-dl));
}

template <int W>
__global__ void __launch_bounds__(OBP_BS) k_clifford_reg(Tab t, const __grid_constant__ CliffBatch b) {
  clifford_reg_body<W>(t, b, t.st->n_scan, &t.st->n_live);
  finish_kernel(t, b.slot);
}

// ================================================================================================
// k_coset_count — exact simplify counters of a Clifford-class gate (run before the gate).
// The k*k raw rows of a surviving term P carry the keys P ^ D, D in the group D of the gate's
// Pauli differences.  num_unique = |D| * (number of cosets of D hit by the surviving terms); each
// row counts how many live, surviving members its coset has, the host sums N_cnt / cnt.
// ================================================================================================
struct CosetP {
  int nd;             // |D| - 1
  int wa, ba, wb, bb; // word / bit of the gate's qubits (wb = -1 for one-qubit gates)
  unsigned code[15];  // D members: x_a | z_a<<1 | x_b<<2 | z_b<<3
  u64 hD[15];
  double scale, atol;
  int slot;
};

__device__ __forceinline__ void coset_body(const Tab& t, const CosetP& p, const u64 n) {
  __shared__ unsigned s_cnt[17];
  __shared__ unsigned s_surv;
  __syncthreads();  // the previous user of these counters (a persistent program) is done with them
  if (threadIdx.x < 17) s_cnt[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_surv = 0;
  __syncthreads();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const double2 c = t.coef[i];
    if (!is_live(c) || is_small_scaled(c, p.scale, p.atol)) continue;
    int cnt = 1;
    const u64 h = t.hash[i];
    for (int d = 0; d < p.nd; ++d) {
      const u64 mxa = (u64)(p.code[d] & 1u) << p.ba, mza = (u64)((p.code[d] >> 1) & 1u) << p.ba;
      const u64 mxb = p.wb >= 0 ? (u64)((p.code[d] >> 2) & 1u) << p.bb : 0ull;
      const u64 mzb = p.wb >= 0 ? (u64)((p.code[d] >> 3) & 1u) << p.bb : 0ull;
      const u64 h2 = h ^ p.hD[d];
      const u64 fp = h2 >> 32;
      u64 s = h2 & t.mask;
      for (u64 probes = 0; probes <= t.mask; ++probes) {
        const u64 e = ld_volatile(&t.index[s]);
        if (e == OBP_EMPTY) break;
        if ((e >> 32) == fp) {
          const u64 j = e & 0xffffffffull;
          if (j < n) {
            const double2 cj = t.coef[j];  // independent of the key words: issued together
            bool eq = true;
            for (int w = 0; w < t.W && eq; ++w) {
              u64 mx = 0, mz = 0;
              if (w == p.wa) { mx ^= mxa; mz ^= mza; }
              if (w == p.wb) { mx ^= mxb; mz ^= mzb; }
              const ulonglong2 ki = t.xz[w][i], kj = t.xz[w][j];
              eq = (kj.x == (ki.x ^ mx)) && (kj.y == (ki.y ^ mz));
            }
            if (eq) {
              if (is_live(cj)) {
                if (!is_small_scaled(cj, p.scale, p.atol)) cnt++;
                break;
              }
            }
          }
        }
        s = (s + 1) & t.mask;
      }
    }
    atomicAdd(&s_cnt[cnt], 1u);
    atomicAdd(&s_surv, 1u);
  }
  __syncthreads();
  if (threadIdx.x < 17 && s_cnt[threadIdx.x]) atomicAdd(&t.ops[p.slot].ncnt[threadIdx.x], (u64)s_cnt[threadIdx.x]);
  if (threadIdx.x == 0 && s_surv) atomicAdd(&t.ops[p.slot].n_surv, (u64)s_surv);
}

__global__ void __launch_bounds__(OBP_BS) k_coset_count(Tab t, const __grid_constant__ CosetP p) {
  coset_body(t, p, t.st->n_scan);
}

// ================================================================================================
// hashing and index (re)build
// ================================================================================================
// h = XOR of the hash-frame columns of the set key bits.  cols: [2*64*W] = x columns then z columns
__global__ void __launch_bounds__(OBP_BS) k_hash_rows(Tab t, const u64* __restrict__ cols, u64 n) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    u64 h = 0;
    for (int w = 0; w < t.W; ++w) {
      const ulonglong2 key = t.xz[w][i];
      u64 m = key.x;
      while (m) { const int bpos = __ffsll((long long)m) - 1; h ^= cols[w * 64 + bpos]; m &= m - 1; }
      m = key.y;
      while (m) { const int bpos = __ffsll((long long)m) - 1; h ^= cols[(t.W + w) * 64 + bpos]; m &= m - 1; }
    }
    t.hash[i] = h;
  }
}

// Insert every live row into the (cleared) index.  merge != 0: rows with equal keys are summed
// into the first published one (fp64 atomicAdd; only reached for caller-supplied duplicates) and
// rows with |c| <= atol (atol >= 0) are filtered first — the simplify() of a raw operator.
// minidx (optional, row-order mode): minidx[r] of the row r that ends up holding a key = the smallest
// row index among all rows with that key, i.e. the key's FIRST APPEARANCE among the raw rows.
__global__ void __launch_bounds__(OBP_BS) k_index_build(Tab t, int merge, double atol, u64 epoch, int slot,
                                                         unsigned* __restrict__ minidx) {
  const u64 n = t.st->n_scan;
  unsigned rows_surv = 0, published = 0;
  int deaths = 0;
  bool full = false;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    if (!is_live(c)) {
      // at load every given row was counted live; an exact zero is a filtered raw row.  Real
      // tombstones (epoch > 0: the rows a generic gate expanded) were never counted.
      if (merge && tomb_epoch(c) == 0) {
        deaths++;
        t.coef[i] = tombstone(epoch);
      }
      continue;
    }
    if (merge && atol >= 0.0 && is_small(c, atol)) {
      t.coef[i] = tombstone(epoch);
      deaths++;
      continue;
    }
    rows_surv++;
    const u64 h = t.hash[i];
    const u64 fp = h >> 32, entry = (fp << 32) | i;
    u64 s = h & t.mask;
    bool done = false;
    for (u64 probes = 0; probes <= t.mask && !done; ++probes) {
      u64 e = ld_volatile(&t.index[s]);
      if (e == OBP_EMPTY) {
        e = atomicCAS(&t.index[s], OBP_EMPTY, entry);
        if (e == OBP_EMPTY) {
          if (minidx) atomicMin(&minidx[i], (unsigned)i);
          published++;
          done = true;
          break;
        }
      }
      if (merge && (e >> 32) == fp) {
        const u64 j = e & 0xffffffffull;
        bool eq = t.hash[j] == h;
        for (int w = 0; w < t.W && eq; ++w) {
          const ulonglong2 ki = t.xz[w][i], kj = t.xz[w][j];
          eq = (ki.x == kj.x) && (ki.y == kj.y);
        }
        if (eq) {
          atomicAdd(&t.coef[j].x, c.x);
          atomicAdd(&t.coef[j].y, c.y);
          if (minidx) atomicMin(&minidx[j], (unsigned)i);
          t.coef[i] = tombstone(epoch);
          deaths++;
          done = true;
          break;
        }
      }
      s = (s + 1) & t.mask;
    }
    if (!done) full = true;
  }
  if (full) atomicOr(&t.st->error, OBP_ERR_BIT_INDEX);
  const long long dl = warp_sum_i64((long long)deaths);
  const u64 a = warp_sum_u64(rows_surv), b = warp_sum_u64(published);
  if ((threadIdx.x & 31) == 0) {
    if (dl) atomicAdd(&t.st->n_live, (u64)(-dl));
    if (slot >= 0) {
      if (a) atomicAdd(&t.ops[slot].rows_surv, a);
      if (b) atomicAdd(&t.ops[slot].keys_present, b);
    }
  }
  finish_kernel(t, slot);
}

// Second zero filter of simplify(): kill merged rows with |sum| <= atol, count and sum them.
// atol < 0: only sums that cancelled to an exact zero die (keeps n_live exact after a plain merge).
__global__ void __launch_bounds__(OBP_BS) k_trim(Tab t, double atol, u64 epoch, int slot) {
  const u64 n = t.st->n_scan;
  unsigned cancelled = 0;
  double sr = 0.0, si = 0.0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    // a merged sum that cancelled to exactly +0.0 + 0.0i looks like a tombstone of epoch 0
    const bool zero_sum = c.x == 0.0 && __double_as_longlong(c.y) == 0ll;
    if (!is_live(c) && !zero_sum) continue;
    if (zero_sum || is_small(c, atol)) {
      t.coef[i] = tombstone(epoch);
      cancelled++;
      sr += c.x;
      si += c.y;
    }
  }
  const u64 cz = warp_sum_u64(cancelled);
  sr = warp_sum_f64(sr);
  si = warp_sum_f64(si);
  if ((threadIdx.x & 31) == 0 && cz) {
    atomicAdd(&t.st->n_live, (u64)(-(long long)cz));
    if (slot >= 0) {
      atomicAdd(&t.ops[slot].cancelled, cz);
      atomicAdd(&t.ops[slot].sum_re, sr);
      atomicAdd(&t.ops[slot].sum_im, si);
    }
  }
  finish_kernel(t, slot);
}

// ================================================================================================
// reset and Pauli-Lindblad damping
// ================================================================================================
// apply_reset_to (utils/operations.py:221-223) + simplify: rows with X/Y on the qubit vanish, Z
// becomes I.  Clearing Z can collide with an existing key: the row that had Z looks the cleared key
// up; if it is live the coefficient moves there (the partner row adds it — pair owned by the row
// that has to move), else the row rewrites its own key/hash and is published under the new hash.
struct ResetP {
  int w;
  u64 bit;
  u64 hZ;  // hash column of z on that qubit
  double atol;
  u64 epoch;
  int slot;
};

// phase 1: the raw-row zero filter — X/Y rows (coefficient set to 0) and rows with |c| <= atol die
__global__ void __launch_bounds__(OBP_BS) k_reset_filter(Tab t, const __grid_constant__ ResetP p) {
  const u64 n = t.st->n_scan;
  unsigned rows_surv = 0;
  int deaths = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    if (!is_live(c)) continue;
    const ulonglong2 key = t.xz[p.w][i];
    if ((key.x & p.bit) || is_small(c, p.atol)) {
      t.coef[i] = tombstone(p.epoch);
      deaths++;
    } else {
      rows_surv++;
    }
  }
  const long long dl = warp_sum_i64((long long)deaths);
  const u64 a = warp_sum_u64(rows_surv);
  if ((threadIdx.x & 31) == 0) {
    if (dl) atomicAdd(&t.st->n_live, (u64)(-dl));
    if (a) atomicAdd(&t.ops[p.slot].rows_surv, a);
  }
  finish_kernel(t, -1);
}

// phase 2: rows with Z on the qubit move to the cleared key.  Only those rows act: they either add
// their coefficient to the live row that already holds the cleared key (exactly one Z row maps to
// it, so each coefficient has one writer) or rewrite their own key/hash and publish the new hash.
__global__ void __launch_bounds__(OBP_BS) k_reset_merge(Tab t, const __grid_constant__ ResetP p) {
  const u64 n = t.st->n_scan;
  unsigned keys_present = 0, cancelled = 0;
  int deaths = 0;
  double sr = 0.0, si = 0.0;
  bool full = false;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    ulonglong2 key = t.xz[p.w][i];
    if (!(key.y & p.bit)) {
      // identity on the qubit (X/Y rows died in phase 1): key stays, counted unless it is merged away
      const double2 ci = t.coef[i];  // live now, or cancelled by its Z partner in this phase
      if (is_live(ci) || tomb_epoch(ci) == p.epoch) keys_present++;
      continue;
    }
    const double2 c = t.coef[i];
    if (!is_live(c)) continue;
    const u64 h2 = t.hash[i] ^ p.hZ;
    const u64 fp = h2 >> 32;
    u64 s = h2 & t.mask;
    long long j = -1;
    for (u64 probes = 0; probes <= t.mask; ++probes) {
      const u64 e = ld_volatile(&t.index[s]);
      if (e == OBP_EMPTY) break;
      if ((e >> 32) == fp) {
        const u64 jj = e & 0xffffffffull;
        if (jj < n && jj != i && t.hash[jj] == h2) {
          bool eq = true;
          for (int w = 0; w < t.W && eq; ++w) {
            const ulonglong2 ki = t.xz[w][i], kj = t.xz[w][jj];
            const u64 mz = (w == p.w) ? p.bit : 0ull;
            eq = (kj.x == ki.x) && (kj.y == (ki.y ^ mz));
          }
          if (eq && is_live(t.coef[jj])) { j = (long long)jj; break; }
        }
      }
      s = (s + 1) & t.mask;
    }
    if (j >= 0) {
      const double2 sum = cadd(t.coef[j], c);
      t.coef[i] = tombstone(p.epoch);
      deaths++;
      if (is_small(sum, p.atol)) {
        // the partner row counted its key as present; it is present but cancels
        t.coef[j] = tombstone(p.epoch);
        deaths++;
        cancelled++;
        sr += sum.x;
        si += sum.y;
      } else {
        t.coef[j] = sum;
      }
    } else {
      key.y ^= p.bit;
      t.xz[p.w][i] = key;
      t.hash[i] = h2;
      if (!index_insert(t, h2, i)) full = true;
      keys_present++;
    }
  }
  if (full) atomicOr(&t.st->error, OBP_ERR_BIT_INDEX);
  const long long dl = warp_sum_i64((long long)deaths);
  const u64 b = warp_sum_u64(keys_present), cz = warp_sum_u64(cancelled);
  sr = warp_sum_f64(sr);
  si = warp_sum_f64(si);
  if ((threadIdx.x & 31) == 0) {
    if (dl) atomicAdd(&t.st->n_live, (u64)(-dl));
    OpStat* o = &t.ops[p.slot];
    if (b) atomicAdd(&o->keys_present, b);
    if (cz) { atomicAdd(&o->cancelled, cz); atomicAdd(&o->sum_re, sr); atomicAdd(&o->sum_im, si); }
  }
  finish_kernel(t, p.slot);
}

// apply_ple_to (utils/operations.py:267-273) + simplify: c *= prod_g factor[g] over anticommuting
// generators (applied in generator order, like the reference's loop), then the zero filter.
// gens: [n_gens][2*W] = x words then z words.
__global__ void __launch_bounds__(OBP_BS) k_damp(Tab t, const u64* __restrict__ gens,
                                                  const double* __restrict__ factors, int n_gens,
                                                  double atol, u64 epoch, int slot) {
  const u64 n = t.st->n_scan;
  int deaths = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    double2 c = t.coef[i];
    if (!is_live(c)) continue;
    bool changed = false;
    for (int g = 0; g < n_gens; ++g) {
      int par = 0;
      for (int w = 0; w < t.W; ++w) {
        const ulonglong2 key = t.xz[w][i];
        par += __popcll(key.x & gens[(size_t)g * 2 * t.W + t.W + w]) + __popcll(key.y & gens[(size_t)g * 2 * t.W + w]);
      }
      if (par & 1) {
        c.x = __dmul_rn(c.x, factors[g]);
        c.y = __dmul_rn(c.y, factors[g]);
        changed = true;
      }
    }
    if (is_small(c, atol)) {
      t.coef[i] = tombstone(epoch);
      deaths++;
    } else if (changed) {
      t.coef[i] = c;
    }
  }
  const long long dl = warp_sum_i64((long long)deaths);
  if ((threadIdx.x & 31) == 0 && dl) atomicAdd(&t.st->n_live, (u64)(-dl));
  finish_kernel(t, slot);
}

// ================================================================================================
// truncation (utils/truncating.py:171-204)
// ================================================================================================
// a[i] = |c_i|^p for live rows, -1 for tombstones; global max through atomicMax on the bit pattern
__global__ void __launch_bounds__(OBP_BS) k_absp(Tab t, double* __restrict__ a, int p_norm) {
  const u64 n = t.st->n_scan;
  double mx = 0.0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    double v = -1.0;
    if (is_live(c)) {
      const double r = hypot(c.x, c.y);
      v = p_norm == 1 ? r : (p_norm == 2 ? __dmul_rn(r, r) : pow(r, (double)p_norm));
      mx = fmax(mx, v);
    }
    a[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(&t.st->amax_bits, (u64)__double_as_longlong(mx));
}

// partial[b] = sum of a[i] with 0 <= a[i] < mid over the rows of CTA b (fixed reduction tree)
__global__ void __launch_bounds__(OBP_BS) k_masked_sum(const double* __restrict__ a, const DevState* st,
                                                        double mid, double* __restrict__ partial) {
  __shared__ double s_w[OBP_BS / 32];
  const u64 n = st->n_scan;
  double acc = 0.0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double v = a[i];
    if (v >= 0.0 && v < mid) acc += v;
  }
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < OBP_BS / 32; ++w) tot += s_w[w];
    partial[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(OBP_BS) k_final_sum(const double* __restrict__ partial, int nb,
                                                       double* __restrict__ out) {
  __shared__ double s_w[OBP_BS / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += OBP_BS) acc += partial[i];
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < OBP_BS / 32; ++w) tot += s_w[w];
    *out = tot;
  }
}

// keep a > thr (strict), tombstone the rest
__global__ void __launch_bounds__(OBP_BS) k_trunc_kill(Tab t, const double* __restrict__ a, double thr, u64 epoch) {
  const u64 n = t.st->n_scan;
  unsigned removed = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double v = a[i];
    if (v >= 0.0 && !(v > thr)) {
      t.coef[i] = tombstone(epoch);
      removed++;
    }
  }
  const u64 r = warp_sum_u64(removed);
  if ((threadIdx.x & 31) == 0 && r) {
    atomicAdd(&t.st->removed, r);
    atomicAdd(&t.st->n_live, (u64)(-(long long)r));
  }
}

// ================================================================================================
// stable compaction: per-CTA live counts -> exclusive scan -> destination ids -> column scatters
// ================================================================================================
#define OBP_CHUNK 1024  // rows per CTA (4 rounds of OBP_BS)

__global__ void __launch_bounds__(OBP_BS) k_live_count(Tab t, unsigned* __restrict__ blk_cnt) {
  __shared__ unsigned warp_tot[OBP_BS / 32];
  const u64 n = t.st->n_scan;
  const u64 base = (u64)blockIdx.x * OBP_CHUNK;
  unsigned cnt = 0;
#pragma unroll
  for (int r = 0; r < OBP_CHUNK / OBP_BS; ++r) {
    const u64 i = base + (u64)r * OBP_BS + threadIdx.x;
    if (i < n && is_live(t.coef[i])) cnt++;
  }
  cnt = (unsigned)warp_sum_u64(cnt);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < OBP_BS / 32; ++w) tot += warp_tot[w];
    blk_cnt[blockIdx.x] = tot;
  }
}

// single CTA: exclusive scan of nb counts in place, total -> st->scan_total
__global__ void __launch_bounds__(1024) k_scan_blocks(unsigned* __restrict__ blk, u64 nb, DevState* st) {
  __shared__ u64 s_part[1024];
  __shared__ u64 s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u64 base = 0; base < nb; base += 1024) {
    const u64 i = base + threadIdx.x;
    const u64 v = i < nb ? blk[i] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
      const u64 add = threadIdx.x >= (unsigned)off ? s_part[threadIdx.x - off] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    const u64 incl = s_part[threadIdx.x];
    if (i < nb) blk[i] = (unsigned)(s_carry + incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) st->scan_total = s_carry;
}

__global__ void __launch_bounds__(OBP_BS) k_dest(Tab t, const unsigned* __restrict__ blk_off,
                                                  unsigned* __restrict__ dest) {
  __shared__ unsigned warp_tot[OBP_BS / 32];
  const u64 n = t.st->n_scan;
  const u64 base = (u64)blockIdx.x * OBP_CHUNK;
  unsigned run = blk_off[blockIdx.x];
#pragma unroll 1
  for (int r = 0; r < OBP_CHUNK / OBP_BS; ++r) {
    const u64 i = base + (u64)r * OBP_BS + threadIdx.x;
    const bool live = i < n && is_live(t.coef[i]);
    unsigned total;
    const unsigned off = block_scan_flag(live, warp_tot, total);
    if (i < n) dest[i] = live ? run + off : 0xffffffffu;
    run += total;
  }
}

__global__ void __launch_bounds__(OBP_BS) k_scatter16(const ulonglong2* __restrict__ src, ulonglong2* __restrict__ dst,
                                                       const unsigned* __restrict__ dest, const DevState* st) {
  const u64 n = st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const unsigned d = dest[i];
    if (d != 0xffffffffu) dst[d] = src[i];
  }
}
__global__ void __launch_bounds__(OBP_BS) k_scatter8(const u64* __restrict__ src, u64* __restrict__ dst,
                                                      const unsigned* __restrict__ dest, const DevState* st) {
  const u64 n = st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const unsigned d = dest[i];
    if (d != 0xffffffffu) dst[d] = src[i];
  }
}
__global__ void k_set_counts(DevState* st) {
  const u64 tot = st->scan_total;
  st->n_scan = tot;
  st->n_append = tot;
  st->n_live = tot;
}

// ================================================================================================
// union count over several tables (backpropagation.py:419-428): copy keys into one scratch table
// ================================================================================================
__global__ void __launch_bounds__(OBP_BS) k_copy_live_keys(Tab src, Tab dst, u64 dst_off) {
  // src must be compact (no tombstones); copies keys and gives every row coefficient 1
  const u64 n = src.st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    for (int w = 0; w < src.W; ++w) dst.xz[w][dst_off + i] = src.xz[w][i];
    dst.coef[dst_off + i] = make_double2(1.0, 0.0);
  }
}

// ================================================================================================
// Sharded table (one rank = one GPU holds the terms whose hash has owner bits == rank)
//
// owner(term) = top log2(R) bits of its Clifford-invariant hash.  Clifford gates, damping and
// truncation never change a hash, so they stay rank-local.  A rotation with generator G maps key P
// to P^G with hash h ^ h(G): when the owner bits of h(G) are zero the partner lives on the same rank
// (plain k_rotate); otherwise ALL partners of this rank's rows live on the single peer
// rank ^ ownerbits(h(G)).  The rotation is then split in three local passes around one pairwise
// record exchange:
//   k_rotate_emit    every row keeps the part of the 2x2 update that stays on its own key (rows
//                    (0,0),(1,1)) and emits the part that lands on the partner key (rows (0,1),(1,0))
//                    as a record {key^G words, hash^h(G), coefficient};
//   k_merge_records  the receiver adds each record to the row holding that key or appends a new row;
//   k_rotate_finish  the second zero filter of simplify() on the rows the gate touched.
// A record is 2W+3 u64: x0,z0,...,x(W-1),z(W-1), hash, re, im.
// ================================================================================================
#define OBP_PENDING_IM 0x8000000000000000ull  // {+0.0, -0.0}: key has no surviving raw row of its own (yet)

__device__ __forceinline__ bool is_pending(double2 c) {
  return __double_as_longlong(c.x) == 0ll && (u64)__double_as_longlong(c.y) == OBP_PENDING_IM;
}

// A sum that cancelled to exactly +0.0 + 0.0i has the bit pattern of an (epoch 0) tombstone; while a
// split rotation is in flight such a value must stay a live row until the finish pass has counted
// it, so it is stored as -0.0 - 0.0i (live, not pending, arithmetically still zero).
__device__ __forceinline__ double2 keep_live(double2 c) {
  return is_live(c) && !is_pending(c) ? c : make_double2(-0.0, -0.0);
}

struct ShardCnt {  // per-call counters (device), summed over ranks by the host
  u64 rows_surv, keys_present, cancelled, n_records;
  double sum_re, sum_im;
};

// `rec` / `cursor` point either to a local staging buffer (exchange by NCCL send/recv) or straight
// into the PEER's inbox over NVLink (P2P stores + a remote atomic on the peer's cursor): in that case
// this one kernel is both the compute step and the transfer.
__global__ void __launch_bounds__(OBP_BS, OBP_ROT_MINB) k_rotate_emit(Tab t, const __grid_constant__ RotP p,
                                                                      u64* __restrict__ rec, u64 rec_cap,
                                                                      u64* __restrict__ cursor,
                                                                      ShardCnt* __restrict__ cnt,
                                                                      const u64* wait_flag, u64 wait_val,
                                                                      u64* done_flag, u64 done_val) {
  // device-side handshake: the peer must have consumed this inbox's previous contents
  cta_wait_flag(wait_flag, wait_val, &t.st->error);
  const u64 n = t.st->n_scan;
  const unsigned lane = threadIdx.x & 31u;
  const int RW = 2 * t.W + 3;
  unsigned rows_surv = 0, keys_present = 0, n_emit = 0;
  int dlive = 0;
  bool overflow = false;
  for (u64 base = ((u64)blockIdx.x * OBP_BS + (threadIdx.x & ~31u)); base < n; base += (u64)gridDim.x * OBP_BS) {
    const u64 i = base + lane;
    bool want_emit = false;
    double2 out_c = make_double2(0.0, 0.0);
    if (i < n) {
      const double2 c = t.coef[i];
      if (is_live(c)) {
        int sp = 0, E = p.yG;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < p.ngw) {
            const ulonglong2 key = t.xz[p.gw[k]][i];
            const u64 m = p.gx[k] | p.gz[k];
            sp += __popcll(key.x & p.gz[k]) + __popcll(key.y & p.gx[k]);
            E += __popcll(key.x & key.y & m) - __popcll((key.x ^ p.gx[k]) & (key.y ^ p.gz[k]) & m) +
                 2 * __popcll(key.y & p.gx[k]);
          }
        }
        const int s = sp & 1;
        E &= 3;
        if (!p.exact && !s) {
          const RowSet rp = rot_rows_self(c, s, p);
          if (rp.nself < 2) {
            if (rp.nself == 0 || is_small(rp.self, p.atol)) {
              t.coef[i] = tombstone(p.epoch);
              dlive--;
            } else {
              t.coef[i] = rp.self;
            }
          }
        } else {
          const RowSet rp = rot_rows(c, E, s, p);
          rows_surv += rp.nself + rp.ncross;
          keys_present += (unsigned)(rp.nself > 0);
          // the row stays live until the finish pass even when nothing of its own survives: a
          // record from the peer may still land on it
          t.coef[i] = rp.nself > 0 ? keep_live(rp.self) : make_double2(0.0, __longlong_as_double((long long)OBP_PENDING_IM));
          if (rp.ncross > 0) {
            want_emit = true;
            out_c = rp.cross;
          }
        }
      }
    }
    // ---- the warp's records leave together: one atomicAdd on the (possibly remote) cursor reserves a contiguous
    // run of the inbox, then ALL 32 lanes write that run word by word — consecutive lanes store consecutive u64,
    // so the records cross NVLink as full 128-byte writes instead of one 8-byte store per word per lane.
    const unsigned bal = __ballot_sync(0xffffffffu, want_emit);
    if (bal) {
      const unsigned cnt = (unsigned)__popc(bal);
      u64 wbase = 0;
      if (lane == 0) wbase = atomicAdd(cursor, (u64)cnt);
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (want_emit) n_emit++;
      const u64 h_out = want_emit ? (t.hash[i] ^ p.hG) : 0ull;
      unsigned fit = cnt;  // records of this warp that fit the inbox
      if (wbase + cnt > rec_cap) {
        fit = wbase < rec_cap ? (unsigned)(rec_cap - wbase) : 0u;
        overflow = true;
      }
      const unsigned total = fit * (unsigned)RW;
      u64* o = rec + wbase * RW;
      for (unsigned it = 0; it * 32u < total; ++it) {
        const unsigned flat = it * 32u + lane;
        const bool valid = flat < total;
        const unsigned k = valid ? flat / (unsigned)RW : 0u, idx = valid ? flat % (unsigned)RW : 0u;
        const int src = (int)__fns(bal, 0, (int)k + 1);  // lane that emitted the warp's k-th record
        const u64 sh = __shfl_sync(0xffffffffu, h_out, src);
        const u64 sre = __shfl_sync(0xffffffffu, (u64)__double_as_longlong(out_c.x), src);
        const u64 sim = __shfl_sync(0xffffffffu, (u64)__double_as_longlong(out_c.y), src);
        if (valid) {
          u64 v;
          if (idx < 2u * (unsigned)t.W) {
            const int w = (int)(idx >> 1);
            u64 gx, gz;
            gen_word(p, w, gx, gz);
            const u64* kw = reinterpret_cast<const u64*>(&t.xz[w][base + (u64)src]);
            v = (idx & 1u) ? (kw[1] ^ gz) : (kw[0] ^ gx);
          } else if (idx == 2u * (unsigned)t.W) {
            v = sh;
          } else {
            v = idx == 2u * (unsigned)t.W + 1u ? sre : sim;
          }
          o[flat] = v;
        }
      }
    }
  }
  if (overflow) atomicOr(&t.st->error, OBP_ERR_BIT_CAP);
  const long long dl = warp_sum_i64((long long)dlive);
  const u64 a = warp_sum_u64(rows_surv), b = warp_sum_u64(keys_present), ne = warp_sum_u64(n_emit);
  if (lane == 0) {
    if (a) atomicAdd(&cnt->rows_surv, a);
    if (b) atomicAdd(&cnt->keys_present, b);
    if (ne && cursor != &cnt->n_records) atomicAdd(&cnt->n_records, ne);
    if (ne) atomicAdd(&t.st->emitted_total, ne);
    if (dl) atomicAdd(&t.st->n_live, (u64)dl);
  }
  // records written into a peer's memory must be visible before the kernel ends / the flag is raised
  finish_kernel_flag(t, p.slot, done_flag, done_val, nullptr);
}

// One thread per incoming record: add to the live row with that key, or append it.
__global__ void __launch_bounds__(OBP_BS) k_merge_records(Tab t, const u64* __restrict__ rec, u64 n_rec,
                                                           const u64* __restrict__ n_rec_ptr, u64 rec_cap,
                                                           ShardCnt* __restrict__ cnt,
                                                           const u64* wait_flag, u64 wait_val) {
  // device-side handshake: the sender's kernel (on another GPU) has finished writing this inbox
  cta_wait_flag(wait_flag, wait_val, &t.st->error);
  if (n_rec_ptr) {  // inbox filled by the peer's kernel: the count is the inbox cursor
    n_rec = ld_volatile(n_rec_ptr);
    if (n_rec > rec_cap) {
      n_rec = rec_cap;
      if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(&t.st->error, OBP_ERR_BIT_CAP);
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) cnt->n_records = n_rec;
  }
  const u64 n = t.st->n_scan;
  const unsigned lane = threadIdx.x & 31u;
  const int RW = 2 * t.W + 3;
  unsigned keys_present = 0;
  int dlive = 0;
  bool overflow = false, full = false;
  for (u64 base = ((u64)blockIdx.x * OBP_BS + (threadIdx.x & ~31u)); base < n_rec; base += (u64)gridDim.x * OBP_BS) {
    const u64 r = base + lane;
    bool want_append = false;
    const u64* in = rec + r * RW;
    u64 h = 0;
    double2 c = make_double2(0.0, 0.0);
    if (r < n_rec) {
      h = in[2 * t.W];
      c = keep_live(make_double2(__longlong_as_double((long long)in[2 * t.W + 1]),
                                 __longlong_as_double((long long)in[2 * t.W + 2])));
      const u64 fp = h >> 32;
      u64 s = h & t.mask;
      long long j = -1;
      for (u64 probes = 0; probes <= t.mask; ++probes) {
        const u64 e = ld_volatile(&t.index[s]);
        if (e == OBP_EMPTY) break;
        if ((e >> 32) == fp) {
          const u64 jj = e & 0xffffffffull;
          if (jj < n) {
            const double2 cj = t.coef[jj];
            bool eq = true;
            for (int w = 0; w < t.W && eq; ++w) {
              const ulonglong2 kj = t.xz[w][jj];
              eq = (kj.x == in[2 * w]) && (kj.y == in[2 * w + 1]);
            }
            if (eq && is_live(cj)) {
              j = (long long)jj;
              if (is_pending(cj)) {
                t.coef[jj] = c;
                keys_present++;
              } else {
                t.coef[jj] = keep_live(cadd(cj, c));
              }
              break;
            }
          }
        }
        s = (s + 1) & t.mask;
      }
      if (j < 0) want_append = true;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, want_append);
    if (bal) {
      u64 wbase = 0;
      if (lane == 0) wbase = atomicAdd(&t.st->n_append, (u64)__popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (want_append) {
        const u64 row = wbase + __popc(bal & ((1u << lane) - 1u));
        if (row < t.cap) {
          for (int w = 0; w < t.W; ++w) t.xz[w][row] = make_ulonglong2(in[2 * w], in[2 * w + 1]);
          t.coef[row] = c;
          t.hash[row] = h;
          if (!index_insert(t, h, row)) full = true;
          keys_present++;
          dlive++;
        } else {
          overflow = true;
        }
      }
    }
  }
  if (overflow) atomicOr(&t.st->error, OBP_ERR_BIT_CAP);
  if (full) atomicOr(&t.st->error, OBP_ERR_BIT_INDEX);
  const long long dl = warp_sum_i64((long long)dlive);
  const u64 b = warp_sum_u64(keys_present);
  if (lane == 0) {
    if (b) atomicAdd(&cnt->keys_present, b);
    if (dl) atomicAdd(&t.st->n_live, (u64)dl);
  }
  finish_kernel(t, -1);
}

// Second zero filter on the rows the split rotation touched (anticommuting rows; all rows when
// the reference's counters are wanted).  A row still pending holds a key without any surviving raw
// row: it disappears without being counted.
__global__ void __launch_bounds__(OBP_BS) k_rotate_finish(Tab t, const __grid_constant__ RotP p,
                                                           ShardCnt* __restrict__ cnt, u64* inbox_hdr, u64 merged_val) {
  const u64 n = t.st->n_scan;
  unsigned cancelled = 0;
  int deaths = 0;
  double sr = 0.0, si = 0.0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    if (!is_live(c)) continue;
    if (!p.exact) {
      int sp = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < p.ngw) {
          const ulonglong2 key = t.xz[p.gw[k]][i];
          sp += __popcll(key.x & p.gz[k]) + __popcll(key.y & p.gx[k]);
        }
      }
      if (!(sp & 1)) continue;
    }
    if (is_pending(c)) {
      t.coef[i] = tombstone(p.epoch);
      deaths++;
    } else if (is_small(c, p.atol)) {
      t.coef[i] = tombstone(p.epoch);
      deaths++;
      cancelled++;
      sr += c.x;
      si += c.y;
    }
  }
  const long long dl = warp_sum_i64((long long)deaths);
  const u64 cz = warp_sum_u64(cancelled);
  sr = warp_sum_f64(sr);
  si = warp_sum_f64(si);
  if ((threadIdx.x & 31) == 0) {
    if (dl) atomicAdd(&t.st->n_live, (u64)(-dl));
    if (cz) { atomicAdd(&cnt->cancelled, cz); atomicAdd(&cnt->sum_re, sr); atomicAdd(&cnt->sum_im, si); }
  }
  // handshake mode: empty the inbox (cursor) and tell the senders that it is free again
  finish_kernel_flag(t, p.slot, inbox_hdr ? inbox_hdr + OBP_HDR_MERGED : nullptr, merged_val,
                     inbox_hdr ? inbox_hdr + OBP_HDR_CURSOR : nullptr);
}

// At load: rows whose hash belongs to another rank are dropped (every rank is given the full operator).
// They become exact zeros, which the index build that follows counts and buries like any other
// filtered raw row; st->removed tells the host how many rows were not this rank's.
__global__ void __launch_bounds__(OBP_BS) k_shard_filter(Tab t, int owner_shift, u64 rank) {
  const u64 n = t.st->n_scan;
  unsigned foreign = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    if ((t.hash[i] >> owner_shift) != rank) {
      t.coef[i] = make_double2(0.0, 0.0);
      foreign++;
    }
  }
  const u64 f = warp_sum_u64(foreign);
  if ((threadIdx.x & 31) == 0 && f) atomicAdd(&t.st->removed, f);
}

// A replicated table becomes this rank's shard: live rows owned by another rank die (late sharding: while
// the operator is small every rank carries all of it and no gate needs an exchange, see sharded.py).
__global__ void __launch_bounds__(OBP_BS) k_shard_drop(Tab t, int owner_shift, u64 rank, u64 epoch) {
  const u64 n = t.st->n_scan;
  int deaths = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    if ((t.hash[i] >> owner_shift) != rank && is_live(t.coef[i])) {
      t.coef[i] = tombstone(epoch);
      deaths++;
    }
  }
  const long long dl = warp_sum_i64((long long)deaths);
  if ((threadIdx.x & 31) == 0 && dl) atomicAdd(&t.st->n_live, (u64)(-dl));
  finish_kernel(t, -1);
}

// Rows per owner under a candidate shard function given as owner-bit planes on the KEY bits (the plan of
// obp_shard_split is judged on how evenly it spreads the rows before any hash is rewritten).
#define OBP_MAX_OWNER_BITS 6
struct OwnerPlanes {
  int bits;
  ulonglong2 mask[OBP_MAX_OWNER_BITS][OBP_MAX_WORDS];  // {x mask, z mask} of owner bit i on key word w
};

__global__ void __launch_bounds__(OBP_BS) k_owner_hist(Tab t, const __grid_constant__ OwnerPlanes pl, u64* __restrict__ counts) {
  __shared__ unsigned s_cnt[1 << OBP_MAX_OWNER_BITS];
  for (int k = threadIdx.x; k < (1 << pl.bits); k += OBP_BS) s_cnt[k] = 0;
  __syncthreads();
  const u64 n = t.st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    if (!is_live(t.coef[i])) continue;
    unsigned owner = 0;
    for (int b = 0; b < pl.bits; ++b) {
      int par = 0;
      for (int w = 0; w < t.W; ++w) {
        const ulonglong2 key = t.xz[w][i];
        par += __popcll(key.x & pl.mask[b][w].x) + __popcll(key.y & pl.mask[b][w].y);
      }
      owner |= (unsigned)(par & 1) << (pl.bits - 1 - b);  // bit 0 of the plan = most significant owner bit (hash bit 63)
    }
    atomicAdd(&s_cnt[owner], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < (1 << pl.bits); k += OBP_BS)
    if (s_cnt[k]) atomicAdd(&counts[k], (u64)s_cnt[k]);
}

// ================================================================================================
// k_program — one persistent cooperative launch for a whole gate program (W <= 2)
//
// A slice of the 127-qubit workload is ~50-130 gates, most of them on tables of a few thousand
// rows where a kernel launch costs more than the pass itself.  The program is therefore run by ONE
// cooperative kernel (grid = resident CTAs, never more than the rows need): every CTA loops over
// the ops, runs the same bodies as the stand-alone kernels and meets the others at a grid barrier
// before the next op.  Live-count changes go to per-op counters so that no CTA has to wait for a
// bookkeeping step; CTA 0 stamps %globaltimer after each barrier, which is how the per-class times
// of the roofline are measured inside this kernel.
// ================================================================================================
#include <cooperative_groups.h>

enum { PROG_ROTATE = 1, PROG_CLIFFORD = 2, PROG_COSET = 3 };

struct ProgEntry {
  int kind;
  int slot;
  unsigned offset;  // byte offset of the op's parameters in the parameter pool (16-byte aligned)
  unsigned bytes;
};

__device__ __forceinline__ u64 global_timer_ns() {
  u64 v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}

template <int W>
__global__ void __launch_bounds__(OBP_BS, 3)
k_program(Tab t, const ProgEntry* __restrict__ entries, int n_entries, const unsigned char* __restrict__ pool,
          int n_slots, u64* __restrict__ t_begin, int flags, u64* __restrict__ host_res, int n_res_ops) {
  namespace cg = cooperative_groups;
  const bool stamps = (flags & 1) != 0;  // per-op %globaltimer stamps (profiling only)
  // Grid barrier (includes the release/acquire that makes an op's stores visible to every CTA).  Measured:
  // ~0.8 us per op of a ~4 us small-table pass; running the program as ONE 16-CTA thread-block cluster
  // with the hardware cluster barrier instead gained < 5 % on 10^3..10^4-row tables and lost on the
  // rest (fewer threads in flight), so there is only this variant.
  auto barrier = [&]() { cg::this_grid().sync(); };
  constexpr unsigned PARAM_BYTES = ((sizeof(CliffBatch) > sizeof(RotP) ? sizeof(CliffBatch) : sizeof(RotP)) + 15u) & ~15u;  // uint4 copies
  __shared__ __align__(16) unsigned char s_param[2][PARAM_BYTES];  // double buffer: op k+1 is staged during op k
  static_assert(sizeof(CosetP) <= sizeof(CliffBatch), "parameter staging buffer too small");
  // the entry table of the program (16 B per op) is read once into shared memory
  constexpr int ENTRY_CACHE = 384;
  __shared__ ProgEntry s_entries[ENTRY_CACHE];
  for (int k = threadIdx.x; k < n_entries && k < ENTRY_CACHE; k += blockDim.x) s_entries[k] = entries[k];
  __syncthreads();
  auto entry = [&](int k) -> ProgEntry { return k < ENTRY_CACHE ? s_entries[k] : entries[k]; };
  {
    // the program zeroes its own counter slots (and stamp slots): one barrier instead of a memset in the stream
    u64* w = reinterpret_cast<u64*>(t.ops);
    const u64 n_words = (u64)(n_slots + n_entries) * (sizeof(OpStat) / sizeof(u64));
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < n_words; k += (u64)gridDim.x * blockDim.x) w[k] = 0;
    barrier();
  }
  if (stamps && blockIdx.x == 0 && threadIdx.x == 0) *t_begin = global_timer_ns();
  auto stage = [&](int k) {
    const ProgEntry e = entry(k);
    unsigned char* dst = s_param[k & 1];
    for (unsigned b = threadIdx.x * 16u; b < e.bytes; b += blockDim.x * 16u)
      *reinterpret_cast<uint4*>(dst + b) = *reinterpret_cast<const uint4*>(pool + e.offset + b);
  };
  if (n_entries > 0) stage(0);
  for (int k = 0; k < n_entries; ++k) {
    const ProgEntry e = entry(k);
    __syncthreads();  // op k's parameters are complete; everybody is done with op k-1's buffer
    if (k + 1 < n_entries) stage(k + 1);  // these loads are in flight while op k runs
    const unsigned char* par = s_param[k & 1];
    u64 n = ld_volatile(&t.st->n_append);  // stable: appends of the previous op ended at the barrier
    if (n > t.cap) n = t.cap;
    u64* live_ctr = &t.ops[e.slot].dlive;
    if (flags & 2) {
      // timing experiment (OBP_B200_PROG_DEBUG=2): barrier + staging only, no table work
    } else if (e.kind == PROG_ROTATE) {
      rotate_body(t, *reinterpret_cast<const RotP*>(par), n, live_ctr);
    } else if (e.kind == PROG_CLIFFORD) {
      clifford_reg_body<W>(t, *reinterpret_cast<const CliffBatch*>(par), n, live_ctr);
    } else {
      coset_body(t, *reinterpret_cast<const CosetP*>(par), n);
    }
    barrier();
    if (stamps && blockIdx.x == 0 && threadIdx.x == 0) t.ops[k + n_slots].t_end = global_timer_ns();  // stamps live behind the slots
  }
  if (blockIdx.x == 0) {
    // Bookkeeping by CTA 0: row counts, the live-row count after every slot (a warp-wide prefix sum of
    // the per-slot changes), then the state, the start stamp and the counter slots go straight into the
    // host's pinned result block (mapped): the host only waits for the kernel, no copy in the stream.
    __shared__ u64 s_live0;
    if (threadIdx.x == 0) {
      u64 n = t.st->n_append;
      if (n > t.cap) n = t.cap;
      t.st->n_append = n;
      t.st->n_scan = n;
      s_live0 = t.st->n_live;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x, per = (n_slots + 31) / 32;
      const int s0 = lane * per, s1 = min(n_slots, s0 + per);
      u64 local = 0;
      for (int s = s0; s < s1; ++s) local += t.ops[s].dlive;
      u64 incl = local;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u64 v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      u64 live = s_live0 + (incl - local);
      for (int s = s0; s < s1; ++s) {
        live += t.ops[s].dlive;
        t.ops[s].n_out = live;
      }
      if (lane == 31) t.st->n_live = s_live0 + incl;
    }
    __syncthreads();
    if (host_res) {
      const u64* src = reinterpret_cast<const u64*>(t.st);  // DevState | start stamp | counter slots: one block
      const int n_words = (int)((OBP_RES_HDR + (size_t)n_res_ops * sizeof(OpStat)) / sizeof(u64));
      for (int k = threadIdx.x; k < n_words; k += blockDim.x) host_res[k] = src[k];
    }
  }
}

// ================================================================================================
// k_expand — the literal k*k raw-row expansion of op.compose(U, front=True).compose(U^dag)
// (utils/operations.py:103-105) for a gate that is neither Clifford-class nor a single-generator
// rotation (cp, crz, arbitrary 1-4 qubit unitaries).  Row (i, j, l) gets key P_i ^ U_j ^ U_l and
// coefficient c_i * u_j * conj(u_l) * (-i)^e, written at n + i*k*k + j*k + l: the reference's raw-row
// order.  The host then runs the same merge as for a freshly loaded operator (k_hash_rows,
// k_index_build with the raw-row zero filter, k_trim), i.e. simplify() as the reference states it.
// ================================================================================================
#define OBP_GEN_MAXK 256  // every gate on <= 4 qubits (4^4 Pauli components); GenP travels as a 5 KB kernel parameter

struct GenP {
  int m;                   // gate qubits (<= 4)
  int k;                   // Pauli components of the gate
  int w[4];                // word of each qubit
  int b[4];                // bit of each qubit
  unsigned code[OBP_GEN_MAXK];  // local code: bit 2j = x, bit 2j+1 = z on the gate's qubit j
  double ur[OBP_GEN_MAXK], ui[OBP_GEN_MAXK];
  u64 epoch;
};

__device__ __forceinline__ int local_y(unsigned v) { return __popc(v & (v >> 1) & 0x55u); }
__device__ __forceinline__ unsigned local_x(unsigned v) { return v & 0x55u; }
__device__ __forceinline__ unsigned local_z(unsigned v) { return (v >> 1) & 0x55u; }

__global__ void __launch_bounds__(OBP_BS) k_expand(Tab t, const __grid_constant__ GenP p, u64 n) {
  const int k2 = p.k * p.k;
  for (u64 r = (u64)blockIdx.x * OBP_BS + threadIdx.x; r < n * (u64)k2; r += (u64)gridDim.x * OBP_BS) {
    const u64 i = r / (u64)k2;
    const int jl = (int)(r % (u64)k2), j = jl / p.k, l = jl % p.k;
    const double2 c = t.coef[i];
    const u64 row = n + r;
    if (!is_live(c)) {  // cannot happen after the compaction that precedes the expansion
      t.coef[row] = make_double2(0.0, 0.0);
      continue;
    }
    unsigned v = 0;
    for (int q = 0; q < p.m; ++q) {
      const ulonglong2 key = t.xz[p.w[q]][i];
      v |= (unsigned)((key.x >> p.b[q]) & 1ull) << (2 * q);
      v |= (unsigned)((key.y >> p.b[q]) & 1ull) << (2 * q + 1);
    }
    const unsigned cj = p.code[j], cl = p.code[l];
    const unsigned k1 = v ^ cj, k2c = k1 ^ cl;
    // compose(front=True): phases add, plus 2|x_P & z_U|; back (U^dag): plus 2|z_P' & x_U|
    int e = local_y(v) + local_y(cj) + 2 * __popc(local_x(v) & local_z(cj)) - local_y(k1);
    e += local_y(k1) + local_y(cl) + 2 * __popc(local_z(k1) & local_x(cl)) - local_y(k2c);
    const double2 uj = make_double2(p.ur[j], p.ui[j]), ul = make_double2(p.ur[l], -p.ui[l]);
    const double2 out = mul_mi_pow(cmul(cmul(c, uj), ul), e);
    const unsigned flip = v ^ k2c;
    for (int w = 0; w < t.W; ++w) {
      ulonglong2 key = t.xz[w][i];
      for (int q = 0; q < p.m; ++q)
        if (p.w[q] == w) {
          key.x ^= (u64)((flip >> (2 * q)) & 1u) << p.b[q];
          key.y ^= (u64)((flip >> (2 * q + 1)) & 1u) << p.b[q];
        }
      t.xz[w][row] = key;
    }
    t.coef[row] = out;
  }
}

// the expanded rows replace the originals
__global__ void __launch_bounds__(OBP_BS) k_kill_prefix(Tab t, u64 n, u64 epoch) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) t.coef[i] = tombstone(epoch);
}

// ---- OperatorBudget(simplify=False): nothing is merged or trimmed, rows with coefficient 0 stay ------
// an exact zero coefficient is stored as -0.0 - 0.0i so that the row does not look like a tombstone
__global__ void __launch_bounds__(OBP_BS) k_revive_zeros(Tab t, u64 first, u64 n) {
  for (u64 i = first + (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    if (!is_live(c) && tomb_epoch(c) == 0) t.coef[i] = make_double2(-0.0, -0.0);
  }
}

// apply_reset_to without the simplify that follows it in the loop (utils/operations.py:221-223)
__global__ void __launch_bounds__(OBP_BS) k_reset_raw(Tab t, int w, u64 bit) {
  const u64 n = t.st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    if (!is_live(t.coef[i])) continue;
    ulonglong2 key = t.xz[w][i];
    if (key.x & bit) t.coef[i] = make_double2(-0.0, -0.0);
    if ((key.x | key.y) & bit) {
      key.x &= ~bit;
      key.y &= ~bit;
      t.xz[w][i] = key;
    }
  }
}

// ================================================================================================
// k_clifford_stream — a whole run of Clifford-class gates on a table wider than 128 qubits in ONE
// pass, with the key words held in registers.  The host turns the run into a stream of micro-ops:
// LOAD word w into register slot s, gate on (slot, bit) operands, STORE slot s back to word w; it
// allocates the two slots (LRU), so a brickwork layer over 1000 qubits reads and writes each of the
// 16 word columns once and the coefficient once (528 B per term per layer) and no shared memory is
// involved.  The stream is read with warp-uniform loads.
// ================================================================================================
#define OBP_STREAM_SWITCH 0x10  // make word `ba` the current word (write the previous one back if it changed)
#define OBP_STREAM_PAIR 0x20    // two-qubit gate on ADJACENT bits ba, ba+1 of the current word (lut in pair layout)
#define OBP_STREAM_CROSS 0x30   // two-qubit gate: qubit a = bit ba of the current word, qubit b = bit bb of word sb
#define OBP_STREAM_MAX 1984     // micro-ops per launch: the stream travels as a kernel parameter (constant bank)

struct CliffStream {
  int n;
  int slot;
  double scale, atol;
  u64 epoch;
  CliffGate g[OBP_STREAM_MAX];
};
static_assert(sizeof(CliffStream) <= 32764, "kernel parameters are limited to 32764 bytes");

// nq = 1 / 2: gate on bit(s) ba (, bb) of the current word with the usual code x_a | z_a<<1 | x_b<<2 | z_b<<3.
// OBP_STREAM_PAIR: code = x_a | x_b<<1 | z_a<<2 | z_b<<3 (two adjacent bits of x, two of z): two
// shifts extract it and two shifts write the flip back — the common case of a brickwork bond.
__global__ void __launch_bounds__(OBP_BS) k_clifford_stream(Tab t, const __grid_constant__ CliffStream st) {
  const u64 n = t.st->n_scan;
  int deaths = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    if (!is_live(c)) continue;
    if (is_small_scaled(c, st.scale, st.atol)) {
      t.coef[i] = tombstone(st.epoch);
      deaths++;
      continue;
    }
    u64 x = 0, z = 0, ox = 0, oz = 0;
    int cur = -1;
    unsigned sgn = 0;
    for (int k = 0; k < st.n; ++k) {
      const CliffGate g = st.g[k];
      if (g.nq == OBP_STREAM_PAIR) {
        const unsigned v = (unsigned)((x >> g.ba) & 3ull) | ((unsigned)((z >> g.ba) & 3ull) << 2);
        const unsigned nv = (unsigned)(g.lut >> (4 * v)) & 15u;
        sgn ^= (g.sign >> v) & 1u;
        const unsigned d = nv ^ v;
        x ^= (u64)(d & 3u) << g.ba;
        z ^= (u64)(d >> 2) << g.ba;
      } else if (g.nq <= 2) {
        const unsigned dd = cliff_apply_bits(g, x, z, x, z, sgn);
        x ^= ((u64)(dd & 1u) << g.ba) ^ ((u64)((dd >> 2) & 1u) << g.bb);
        z ^= ((u64)((dd >> 1) & 1u) << g.ba) ^ ((u64)((dd >> 3) & 1u) << g.bb);
      } else if (g.nq == OBP_STREAM_SWITCH) {
        if (cur >= 0 && (x != ox || z != oz)) t.xz[cur][i] = make_ulonglong2(x, z);
        cur = g.ba;
        if (cur < OBP_MAX_WORDS) {
          const ulonglong2 key = t.xz[cur][i];
          x = ox = key.x;
          z = oz = key.y;
        } else {
          cur = -1;  // end of stream marker: only the write-back above
        }
      } else {  // OBP_STREAM_CROSS
        ulonglong2 kb = t.xz[g.sb][i];
        CliffGate h = g;
        h.nq = 2;
        const unsigned dd = cliff_apply_bits(h, x, z, kb.x, kb.y, sgn);
        x ^= (u64)(dd & 1u) << g.ba;
        z ^= (u64)((dd >> 1) & 1u) << g.ba;
        if (dd >> 2) {
          kb.x ^= (u64)((dd >> 2) & 1u) << g.bb;
          kb.y ^= (u64)((dd >> 3) & 1u) << g.bb;
          t.xz[g.sb][i] = kb;
        }
      }
    }
    if (sgn) t.coef[i] = make_double2(-c.x, -c.y);
  }
  const long long dl = warp_sum_i64((long long)deaths);
  if ((threadIdx.x & 31) == 0 && dl) atomicAdd(&t.st->n_live, (u64)(-dl));
  finish_kernel(t, st.slot);
}

// download: the word columns of row i as consecutive u64 (the caller's [n][W] row-major layout)
__global__ void __launch_bounds__(OBP_BS) k_pack_rows(Tab t, u64* __restrict__ out_x, u64* __restrict__ out_z, u64 n) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    for (int w = 0; w < t.W; ++w) {
      const ulonglong2 key = t.xz[w][i];
      out_x[i * t.W + w] = key.x;
      out_z[i * t.W + w] = key.y;
    }
  }
}

// ================================================================================================
// k_trunc_search — truncate_binary_search (utils/truncating.py:171-204) in ONE cooperative launch:
// |c|^p and its maximum, the whole bisection and the final kill pass.
//
// The bisection is a chain of ~27 dependent masked sums; each needs every CTA's partial sum, i.e. two grid
// barriers.  Instead of one step per pass, a pass evaluates the masked sums of ALL thresholds the next
// OBP_TRUNC_LEVELS steps can ask for: the 2^L - 1 nodes of the decision tree below the current
// (upper, lower), each `mid` computed with the reference's own expression (upper + lower) / 2 from its
// parent's bounds, so every threshold — and every sum, accumulated per threshold with the same fixed
// reduction tree — is bit-identical to what the one-step-per-pass search would produce.  CTA 0 then walks
// the tree with the reference's loop (`<`, `>`, the isclose exit, :179-196) for up to L steps.  6 passes
// instead of 27; a table of <= OBP_TRUNC_ONE_CTA rows runs in a single CTA with no grid barrier at all.
// ================================================================================================
#define OBP_TRUNC_LEVELS 5
#define OBP_TRUNC_NODES ((1 << OBP_TRUNC_LEVELS) - 1)
#define OBP_TRUNC_ONE_CTA 8192

struct TruncState {
  double upper, lower, upper_err, lower_err;
  u64 steps;
};

// np.isclose(upper_error, lower_error, atol=tol) with the default rtol = 1e-5 (:181)
__device__ __forceinline__ bool trunc_done(double upper, double lower, double ue, double le, double tol) {
  const bool close = fabs(ue - le) <= tol + 1e-5 * fabs(le);
  return !((upper - lower) > tol) || close;
}

__global__ void __launch_bounds__(OBP_BS) k_trunc_search(Tab t, double* __restrict__ a, double* __restrict__ partial,
                                                          TruncState* __restrict__ ts, double budget, int p_norm,
                                                          double tol, u64 epoch) {
  namespace cg = cooperative_groups;
  const bool one_cta = gridDim.x == 1;
  auto barrier = [&]() {
    if (one_cta) {
      __syncthreads();
    } else {
      __threadfence();
      cg::this_grid().sync();
    }
  };
  __shared__ double s_w[OBP_BS / 32][OBP_TRUNC_NODES];
  __shared__ double s_thr[OBP_TRUNC_NODES], s_tot[OBP_TRUNC_NODES];
  const u64 n = t.st->n_scan;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // ---- a = |c|^p, global maximum ----------------------------------------------------------------
  double mx = 0.0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double2 c = t.coef[i];
    double v = -1.0;
    if (is_live(c)) {
      const double r = hypot(c.x, c.y);
      v = p_norm == 1 ? r : (p_norm == 2 ? __dmul_rn(r, r) : pow(r, (double)p_norm));
      mx = fmax(mx, v);
    }
    a[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0 && mx > 0.0) atomicMax(&t.st->amax_bits, (u64)__double_as_longlong(mx));
  barrier();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ts->upper = __longlong_as_double((long long)ld_volatile(&t.st->amax_bits));
    ts->lower = 0.0;
    ts->upper_err = budget;
    ts->lower_err = 0.0;
    ts->steps = 0;
  }
  barrier();
  // ---- bisection, OBP_TRUNC_LEVELS steps per pass ---------------------------------------------------
  for (int round = 0; round < 4096; ++round) {
    const double upper = *(volatile double*)&ts->upper, lower = *(volatile double*)&ts->lower;
    const double ue = *(volatile double*)&ts->upper_err, le = *(volatile double*)&ts->lower_err;
    if (trunc_done(upper, lower, ue, le, tol)) break;  // every CTA evaluates the same test on the same state
    // decision tree in heap order: node k has bounds (up[k], lo[k]); child 2k+1 = "mid_error > budget"
    // (upper = mid), child 2k+2 = "mid_error <= budget" (lower = mid)
    if (threadIdx.x == 0) {
      double up[OBP_TRUNC_NODES], lo[OBP_TRUNC_NODES];
      up[0] = upper;
      lo[0] = lower;
#pragma unroll
      for (int k = 0; k < OBP_TRUNC_NODES; ++k) {
        const double mid = (up[k] + lo[k]) / 2;  // :185
        s_thr[k] = mid;
        if (2 * k + 2 < OBP_TRUNC_NODES) {
          up[2 * k + 1] = mid; lo[2 * k + 1] = lo[k];
          up[2 * k + 2] = up[k]; lo[2 * k + 2] = mid;
        }
      }
    }
    __syncthreads();
    double thr[OBP_TRUNC_NODES], acc[OBP_TRUNC_NODES];
#pragma unroll
    for (int k = 0; k < OBP_TRUNC_NODES; ++k) { thr[k] = s_thr[k]; acc[k] = 0.0; }
    for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
      const double v = a[i];
      if (v >= 0.0) {
#pragma unroll
        for (int k = 0; k < OBP_TRUNC_NODES; ++k)
          if (v < thr[k]) acc[k] += v;  // strict <, :187
      }
    }
#pragma unroll
    for (int k = 0; k < OBP_TRUNC_NODES; ++k) {
      const double r = warp_sum_f64(acc[k]);
      if (lane == 0) s_w[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < OBP_TRUNC_NODES) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < OBP_BS / 32; ++w) tot += s_w[w][threadIdx.x];
      if (one_cta) s_tot[threadIdx.x] = tot;
      else partial[(size_t)blockIdx.x * OBP_TRUNC_NODES + threadIdx.x] = tot;
    }
    barrier();
    if (blockIdx.x == 0) {
      if (!one_cta) {
        // fixed-order reduction of the per-CTA partial sums: warp w takes the thresholds w, w + 8, ...
        for (int k = (int)warp; k < OBP_TRUNC_NODES; k += OBP_BS / 32) {
          double sum = 0.0;
          for (unsigned b = lane; b < gridDim.x; b += 32u) sum += *(volatile double*)&partial[(size_t)b * OBP_TRUNC_NODES + k];
          sum = warp_sum_f64(sum);
          if (lane == 0) s_tot[k] = sum;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        double up = upper, lo = lower, uerr = ue, lerr = le;
        u64 steps = ts->steps;
        int k = 0;
        for (int lvl = 0; lvl < OBP_TRUNC_LEVELS; ++lvl) {
          if (trunc_done(up, lo, uerr, lerr, tol)) break;
          const double mid = s_thr[k], tot = s_tot[k];
          const double mid_err = p_norm == 1 ? tot : pow(tot, 1.0 / (double)p_norm);
          if (mid_err <= budget) { lo = mid; lerr = mid_err; k = 2 * k + 2; }
          else { up = mid; uerr = mid_err; k = 2 * k + 1; }
          ++steps;
        }
        ts->upper = up;
        ts->lower = lo;
        ts->upper_err = uerr;
        ts->lower_err = lerr;
        ts->steps = steps;
      }
    }
    barrier();
  }
  // ---- keep a > lower (:199) -----------------------------------------------------------------------
  const double thr_keep = *(volatile double*)&ts->lower;
  unsigned removed = 0;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    const double v = a[i];
    if (v >= 0.0 && !(v > thr_keep)) {
      t.coef[i] = tombstone(epoch);
      removed++;
    }
  }
  const u64 r = warp_sum_u64(removed);
  if (lane == 0 && r) {
    atomicAdd(&t.st->removed, r);
    atomicAdd(&t.st->n_live, (u64)(-(long long)r));
  }
}

// ---- reset on a sharded table whose cleared keys belong to the peer rank ---------------------------
// After k_reset_filter (X/Y rows and |c| <= atol rows are gone): rows with Z on the qubit move to the
// key with that Z cleared, whose hash has other owner bits — they leave as records; rows with I stay.
__global__ void __launch_bounds__(OBP_BS) k_reset_emit(Tab t, const __grid_constant__ ResetP p, u64* __restrict__ rec,
                                                        u64 rec_cap, ShardCnt* __restrict__ cnt) {
  const u64 n = t.st->n_scan;
  const unsigned lane = threadIdx.x & 31u;
  const int RW = 2 * t.W + 3;
  unsigned keys_present = 0;
  int deaths = 0;
  bool overflow = false;
  for (u64 base = ((u64)blockIdx.x * OBP_BS + (threadIdx.x & ~31u)); base < n; base += (u64)gridDim.x * OBP_BS) {
    const u64 i = base + lane;
    bool want_emit = false;
    double2 c = make_double2(0.0, 0.0);
    if (i < n) {
      c = t.coef[i];
      if (is_live(c)) {
        if (t.xz[p.w][i].y & p.bit) want_emit = true;
        else keys_present++;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, want_emit);
    if (bal) {
      u64 wbase = 0;
      if (lane == 0) wbase = atomicAdd(&cnt->n_records, (u64)__popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (want_emit) {
        const u64 r = wbase + __popc(bal & ((1u << lane) - 1u));
        if (r < rec_cap) {
          u64* o = rec + r * RW;
          for (int w = 0; w < t.W; ++w) {
            const ulonglong2 key = t.xz[w][i];
            o[2 * w] = key.x;
            o[2 * w + 1] = w == p.w ? (key.y & ~p.bit) : key.y;
          }
          o[2 * t.W] = t.hash[i] ^ p.hZ;
          o[2 * t.W + 1] = (u64)__double_as_longlong(c.x);
          o[2 * t.W + 2] = (u64)__double_as_longlong(c.y);
          t.coef[i] = tombstone(p.epoch);
          deaths++;
        } else {
          overflow = true;
        }
      }
    }
  }
  if (overflow) atomicOr(&t.st->error, OBP_ERR_BIT_CAP);
  const long long dl = warp_sum_i64((long long)deaths);
  const u64 b = warp_sum_u64(keys_present);
  if (lane == 0) {
    if (b) atomicAdd(&cnt->keys_present, b);
    if (dl) atomicAdd(&t.st->n_live, (u64)(-dl));
  }
  finish_kernel(t, -1);
}

// ================================================================================================
// Row-order mode: put the surviving rows in the order of their keys' first appearance among the raw
// rows — the row order of the reference's simplify() (unordered_unique keeps first-appearance
// order, utils/simplify.py:126, 160-165).  flag[p] = 1 iff some surviving key first appeared at raw
// position p; an exclusive scan of the flags is the new row index of that key.
// ================================================================================================
__global__ void __launch_bounds__(OBP_BS) k_order_flags(Tab t, const unsigned* __restrict__ minidx,
                                                         unsigned* __restrict__ flag) {
  const u64 n = t.st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS)
    if (is_live(t.coef[i])) flag[minidx[i]] = 1u;
}

__global__ void __launch_bounds__(OBP_BS) k_flag_count(const unsigned* __restrict__ flag, const DevState* st,
                                                        unsigned* __restrict__ blk_cnt) {
  __shared__ unsigned warp_tot[OBP_BS / 32];
  const u64 n = st->n_scan;
  const u64 base = (u64)blockIdx.x * OBP_CHUNK;
  unsigned cnt = 0;
#pragma unroll
  for (int r = 0; r < OBP_CHUNK / OBP_BS; ++r) {
    const u64 i = base + (u64)r * OBP_BS + threadIdx.x;
    if (i < n && flag[i]) cnt++;
  }
  cnt = (unsigned)warp_sum_u64(cnt);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < OBP_BS / 32; ++w) tot += warp_tot[w];
    blk_cnt[blockIdx.x] = tot;
  }
}

// pos[p] = number of flags before p (only meaningful where flag[p] is set)
__global__ void __launch_bounds__(OBP_BS) k_flag_pos(const unsigned* __restrict__ flag, const DevState* st,
                                                      const unsigned* __restrict__ blk_off, unsigned* __restrict__ pos) {
  __shared__ unsigned warp_tot[OBP_BS / 32];
  const u64 n = st->n_scan;
  const u64 base = (u64)blockIdx.x * OBP_CHUNK;
  unsigned run = blk_off[blockIdx.x];
#pragma unroll 1
  for (int r = 0; r < OBP_CHUNK / OBP_BS; ++r) {
    const u64 i = base + (u64)r * OBP_BS + threadIdx.x;
    const bool f = i < n && flag[i] != 0;
    unsigned total;
    const unsigned off = block_scan_flag(f, warp_tot, total);
    if (i < n) pos[i] = run + off;
    run += total;
  }
}

__global__ void __launch_bounds__(OBP_BS) k_order_dest(Tab t, const unsigned* __restrict__ minidx,
                                                        const unsigned* __restrict__ pos, unsigned* __restrict__ dest) {
  const u64 n = t.st->n_scan;
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS)
    dest[i] = is_live(t.coef[i]) ? pos[minidx[i]] : 0xffffffffu;
}

// ================================================================================================
// k_copy_table — commit / roll back (obp_snapshot / obp_restore): all columns of the first n rows in
// ONE launch instead of W + 2 stream copies (the per-slice commit is on the host's critical path).
// ================================================================================================
struct CopyP {
  const ulonglong2* src16[OBP_MAX_WORDS + 1];  // key columns, then the coefficient column
  ulonglong2* dst16[OBP_MAX_WORDS + 1];
  const u64* src8;
  u64* dst8;
  int n16;
};

__global__ void __launch_bounds__(OBP_BS) k_copy_table(const __grid_constant__ CopyP p, u64 n) {
  for (u64 i = (u64)blockIdx.x * OBP_BS + threadIdx.x; i < n; i += (u64)gridDim.x * OBP_BS) {
    for (int c = 0; c < p.n16; ++c) p.dst16[c][i] = p.src16[c][i];
    p.dst8[i] = p.src8[i];
  }
}
