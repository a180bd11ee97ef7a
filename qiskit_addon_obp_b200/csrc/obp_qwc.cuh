// obp_qwc.cuh — number of qubit-wise commuting groups on the device.
//
// Replaces  len(op.group_commuting(qubit_wise=True))  (backpropagation.py:368-375, 433-440): Qiskit builds
// the graph whose edges join Pauli terms that do NOT commute qubit-wise and rustworkx colours it greedily,
// vertices by descending degree, ties in row order.  Two terms clash iff on some qubit both are non-identity
// and different; on packed words:  (x_i|z_i) & (x_j|z_j) & ((x_i^x_j) | (z_i^z_j)) != 0  — 64 qubits per
// operation.  Two kernels: the degrees (all n^2 pairs, tiled through shared memory, every SM busy) and the
// colouring itself (inherently sequential over the vertices: ONE CTA, 1024 threads scan the already coloured
// vertices of each new vertex, the used colours are a bit map in shared memory).
#pragma once
#include "obp_device.cuh"

#define OBP_QWC_TILE 256
#define OBP_QWC_MAXW OBP_MAX_WORDS

// keys: [n][W] ulonglong2 {x word, z word} (row-major: a row's words are contiguous)
template <int W>
__device__ __forceinline__ bool qwc_clash(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ b) {
  u64 acc = 0;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const ulonglong2 p = a[w], q = b[w];
    acc |= (p.x | p.y) & (q.x | q.y) & ((p.x ^ q.x) | (p.y ^ q.y));
  }
  return acc != 0;
}

template <int W>
__global__ void __launch_bounds__(OBP_QWC_TILE) k_qwc_degree(const ulonglong2* __restrict__ keys, unsigned n,
                                                              unsigned* __restrict__ degree) {
  constexpr unsigned TJ = W <= 8 ? OBP_QWC_TILE : OBP_QWC_TILE / 2;  // column tile: at most 32 KB of shared memory
  __shared__ ulonglong2 s_keys[TJ * W];
  const unsigned i = blockIdx.x * OBP_QWC_TILE + threadIdx.x;
  ulonglong2 mine[W];
#pragma unroll
  for (int w = 0; w < W; ++w) mine[w] = i < n ? keys[(size_t)i * W + w] : make_ulonglong2(0ull, 0ull);
  unsigned deg = 0;
  // blockIdx.y strides over the column tiles so that small n still fills the GPU
  for (unsigned j0 = blockIdx.y * TJ; j0 < n; j0 += gridDim.y * TJ) {
    __syncthreads();
    for (unsigned k = threadIdx.x; k < TJ * W; k += OBP_QWC_TILE) {
      const size_t src = (size_t)j0 * W + k;
      s_keys[k] = src < (size_t)n * W ? keys[src] : make_ulonglong2(0ull, 0ull);  // all-identity rows clash with nothing
    }
    __syncthreads();
    const unsigned lim = min(TJ, n - j0);
    for (unsigned j = 0; j < lim; ++j) deg += qwc_clash<W>(mine, &s_keys[j * W]) ? 1u : 0u;
  }
  if (i < n && deg) atomicAdd(&degree[i], deg);
}

// keys_sorted: the keys in colouring order (descending degree, stable).  colour_sorted[k] = colour of the k-th
// vertex in that order.  *n_colours = number of colours used.
template <int W>
__global__ void __launch_bounds__(1024) k_qwc_colour(const ulonglong2* __restrict__ keys_sorted, unsigned n,
                                                      unsigned* __restrict__ colour_sorted, unsigned* __restrict__ n_colours) {
  extern __shared__ unsigned s_used[];  // bit c set: colour c is taken by a clashing neighbour; (n + 31) / 32 words
  __shared__ unsigned s_pick;
  __shared__ ulonglong2 s_v[W];
  const unsigned words = (n + 31u) / 32u;
  for (unsigned w = threadIdx.x; w < words; w += blockDim.x) s_used[w] = 0u;
  unsigned ncol = 0;  // identical in every thread
  __syncthreads();
  for (unsigned k = 0; k < n; ++k) {
    if (threadIdx.x < W) s_v[threadIdx.x] = keys_sorted[(size_t)k * W + threadIdx.x];
    if (threadIdx.x == 0) s_pick = 0xffffffffu;
    __syncthreads();
    for (unsigned j = threadIdx.x; j < k; j += blockDim.x)
      if (qwc_clash<W>(s_v, &keys_sorted[(size_t)j * W])) {
        const unsigned c = colour_sorted[j];
        atomicOr(&s_used[c >> 5], 1u << (c & 31u));
      }
    __syncthreads();
    // smallest colour not taken: among the first ncol colours, else a new one
    const unsigned cw = (ncol + 32u) / 32u;  // words that can hold a free colour <= ncol
    for (unsigned w = threadIdx.x; w < cw; w += blockDim.x) {
      const unsigned freebits = ~s_used[w];
      if (freebits) atomicMin(&s_pick, w * 32u + (unsigned)(__ffs((int)freebits) - 1));
    }
    __syncthreads();
    const unsigned c = s_pick;
    if (threadIdx.x == 0) colour_sorted[k] = c;
    if (c + 1u > ncol) ncol = c + 1u;
    for (unsigned w = threadIdx.x; w < cw; w += blockDim.x) s_used[w] = 0u;
    __syncthreads();  // (also orders colour_sorted[k] before the next vertex reads it)
  }
  if (threadIdx.x == 0) *n_colours = ncol;
}

__global__ void __launch_bounds__(256) k_qwc_gather(const ulonglong2* __restrict__ keys, const unsigned* __restrict__ order,
                                                     unsigned n, int W, ulonglong2* __restrict__ keys_sorted) {
  for (size_t t = (size_t)blockIdx.x * 256 + threadIdx.x; t < (size_t)n * W; t += (size_t)gridDim.x * 256) {
    const unsigned k = (unsigned)(t / W), w = (unsigned)(t % W);
    keys_sorted[t] = keys[(size_t)order[k] * W + w];
  }
}
