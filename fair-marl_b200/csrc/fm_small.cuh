// Small-team helpers shared by the env-tile (fm_tile.cu) and agent-warp (fm_aw.cu) kernels:
// lexifair assignment by enumeration for N <= 4, entirely in registers.
#pragma once
#include <utility>

#include "fm_device.cuh"

namespace fm {

// k-th permutation of 0..N-1 in lexicographic order, element i (compile-time foldable).
__host__ __device__ constexpr int perm_elem(int N, int k, int i) {
  int avail[4] = {0, 1, 2, 3};
  int fact = 1;
  for (int q = 2; q < N; ++q) fact *= q;             // (N-1)!
  int n = N, res = 0;
  for (int pos = 0; pos <= i; ++pos) {
    const int idx = k / fact;
    k -= idx * fact;
    res = avail[idx];
    for (int q = idx; q < n - 1; ++q) avail[q] = avail[q + 1];
    --n;
    if (n > 1) fact /= n;
  }
  return res;
}
__host__ __device__ constexpr int factorial(int n) { return n <= 1 ? 1 : n * factorial(n - 1); }

template <int N, int K, int I>
constexpr int perm_v = perm_elem(N, K, I);

// One candidate permutation K: descending-sorted ranks packed 4 bits each, largest most significant.
template <int N, int K, int... I>
__device__ __forceinline__ void perm_try(const int (&rank)[N * N], unsigned& best, int (&out)[N],
                                         std::integer_sequence<int, I...>) {
  int r[N] = {rank[I * N + perm_v<N, K, I>]...};
#pragma unroll
  for (int s = 0; s < N - 1; ++s)                    // bubble network: ascending in place
#pragma unroll
    for (int t = 0; t < N - 1 - s; ++t) {
      const int hi = max(r[t], r[t + 1]), lo = min(r[t], r[t + 1]);
      r[t + 1] = hi; r[t] = lo;
    }
  unsigned key = 0;
#pragma unroll
  for (int i = N - 1; i >= 0; --i) key = (key << 4) | (unsigned)r[i];   // ranks < 16 (N <= 4)
  const bool better = key < best;
  best = better ? key : best;
  ((out[I] = better ? perm_v<N, K, I> : out[I]), ...);
}

template <int N, int... K>
__device__ __forceinline__ void perm_all(const int (&rank)[N * N], unsigned& best, int (&out)[N],
                                         std::integer_sequence<int, K...>) {
  (perm_try<N, K>(rank, best, out, std::make_integer_sequence<int, N>{}), ...);
}

// Lexifair assignment by enumeration (marl_fair_assign.py:16-55; oracle/lexifair.py
// lexifair_bruteforce_batched): rank every entry in the total order (cost, i, j), then take the
// permutation whose descending-sorted rank vector is lexicographically smallest.  Equals the
// sorted threshold descent of lexifair_group<G> (fm_device.cuh) entry for entry, ties included.
template <int N>
__device__ __forceinline__ void lexifair_small(const double (&c)[N * N], int (&out)[N]) {
  int rank[N * N];
#pragma unroll
  for (int a = 0; a < N * N; ++a) rank[a] = 0;
#pragma unroll
  for (int a = 0; a < N * N; ++a)
#pragma unroll
    for (int b = a + 1; b < N * N; ++b) {
      const bool a_first = c[a] <= c[b];             // equal costs: lower flat index (i, j) first
      rank[b] += a_first ? 1 : 0;
      rank[a] += a_first ? 0 : 1;
    }
  unsigned best = 0xffffffffu;
  perm_all<N>(rank, best, out, std::make_integer_sequence<int, factorial(N)>{});
}

}  // namespace fm
