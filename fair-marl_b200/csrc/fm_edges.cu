// Policy-side edge list in ONE pass (TransformerConvNet.process_adj, onpolicy/algorithms/utils/gnn_new.py:381-413):
// mask (adj < max_edge_dist) & (adj > 0), edges in (b, i, j) order, every graph emitted `repeat` times consecutively.
//
// The three-kernel form (fm_kernels.cu: count / scan / emit) reads adj twice and writes each edge with three stores whose
// warps are mostly empty (7 % of the entries of a distance matrix are edges: 2-3 lanes per ballot, times `repeat` copies).
// Here a CTA of 8 warps owns 8 consecutive graphs:
//   1. each warp reads its graph once (8 loads in flight per lane) and compacts the edges -- ballot + popc, (b, i, j) order
//      -- into a shared-memory list (row, column, distance);
//   2. the CTA's edge count is published and its exclusive prefix fetched by a decoupled look-back over one 64-bit status
//      word per CTA (flag in the top two bits, count below: self-contained, no fence needed); tiles are claimed from a
//      counter, so a CTA only ever waits for CTAs that are already running;
//   3. each warp writes its graph's offsets and, per copy, the list with full consecutive lanes: 8-byte / 4-byte stores
//      to consecutive addresses.
// adj is read once, nothing is re-read from global memory, and there is no separate count or scan launch.
#include "fm_device.cuh"
#include "fm_launch.h"

namespace fm {

namespace {

constexpr int EF_WARPS = 8;
constexpr unsigned long long EF_AGG = 1ull << 62, EF_INC = 2ull << 62, EF_VAL = (1ull << 62) - 1ull;

__device__ __forceinline__ bool ef_pred(float d, float thr, int inclusive) {
  return (inclusive ? (d <= thr) : (d < thr)) && (d > 0.0f);
}
__device__ __forceinline__ unsigned long long ef_load(const unsigned long long* q) {
  return *reinterpret_cast<const volatile unsigned long long*>(q);
}
__device__ __forceinline__ void ef_store(unsigned long long* q, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long*>(q) = v;
}

__global__ void __launch_bounds__(EF_WARPS * 32, 6)
edge_fused_kernel(const float* __restrict__ adj, int num_graphs, int E, float thr, int inclusive, int repeat, long long capacity,
                  int num_tiles, unsigned long long* __restrict__ status, unsigned int* __restrict__ counter,
                  long long* __restrict__ graph_offsets, long long* __restrict__ edge_index, float* __restrict__ edge_attr,
                  long long* __restrict__ nnz_out) {
  extern __shared__ int2 lists[];                      // [EF_WARPS][E * E]: x = row << 16 | column, y = distance bits
  __shared__ int wc[EF_WARPS];
  __shared__ int s_tile;
  __shared__ long long s_excl;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(counter, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int g = tile * EF_WARPS + w;
  const int EE = E * E;
  int2* my = lists + (size_t)w * EE;
  int cnt = 0;
  if (g < num_graphs) {
    const float* a = adj + (size_t)g * EE;
    const unsigned magic = (E <= 100) ? ((1u << 20) + E - 1) / E : 0u;     // q / E == (q * magic) >> 20 for q < 2^20 / E
    constexpr int U = 8;
    for (int q0 = 0; q0 < EE; q0 += 32 * U) {
      float dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int q = q0 + u * 32 + lane; dv[u] = (q < EE) ? __ldcs(a + q) : 0.0f; }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u * 32 + lane;
        if (q0 + u * 32 >= EE) break;                                      // warp-uniform
        const bool pr = (q < EE) && ef_pred(dv[u], thr, inclusive);
        const unsigned b = __ballot_sync(FULL, pr);
        if (pr) {
          const int r = magic ? (int)(((unsigned)q * magic) >> 20) : q / E;
          my[cnt + __popc(b & ((1u << lane) - 1u))] = make_int2((r << 16) | (q - r * E), __float_as_int(dv[u]));
        }
        cnt += __popc(b);
      }
    }
  }
  if (lane == 0) wc[w] = cnt;
  __syncthreads();
  if (w == 0) {
    int v = lane < EF_WARPS ? wc[lane] : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    const unsigned long long agg = (unsigned long long)v;
    unsigned long long excl = 0;
    if (tile > 0) {
      if (lane == 0) ef_store(status + tile, EF_AGG | agg);
      int k = tile - 1;
      while (true) {                                                       // 32 predecessors per round, nearest in lane 0
        const int idx = k - lane;
        const unsigned long long word = idx >= 0 ? ef_load(status + idx) : EF_INC;
        const unsigned flag = (unsigned)(word >> 62);
        const unsigned inc = __ballot_sync(FULL, flag == 2u), none = __ballot_sync(FULL, flag == 0u);
        const int first = inc ? __ffs(inc) - 1 : 32;                       // the prefix ends at the first inclusive word
        const unsigned need = first >= 31 ? FULL : ((2u << first) - 1u);
        if (none & need) { __nanosleep(40); continue; }                    // a word on the way is not published yet
        unsigned long long part = (lane <= first) ? (word & EF_VAL) : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULL, part, off);
        excl += part;
        if (inc) break;
        k -= 32;
      }
    }
    if (lane == 0) {
      ef_store(status + tile, EF_INC | (excl + agg));
      s_excl = (long long)excl;
      if (tile == num_tiles - 1) {
        const long long total = (long long)(excl + agg) * repeat;
        graph_offsets[(size_t)num_graphs * repeat] = total;
        if (nnz_out) *nnz_out = total;
      }
    }
  }
  __syncthreads();
  if (g >= num_graphs) return;
  int before = 0;
#pragma unroll
  for (int k = 0; k < EF_WARPS; ++k) before += (k < w) ? wc[k] : 0;
  const long long base0 = (s_excl + before) * (long long)repeat;
  for (int cp = lane; cp < repeat; cp += 32) graph_offsets[(size_t)g * repeat + cp] = base0 + (long long)cp * cnt;
  for (int cp = 0; cp < repeat; ++cp) {
    const long long node0 = ((long long)g * repeat + cp) * E;
    const long long p0 = base0 + (long long)cp * cnt;
    for (int k = lane; k < cnt; k += 32) {
      const long long pos = p0 + k;
      if (pos < capacity) {
        const int2 e = my[k];
        __stcs(edge_index + pos, node0 + (e.x >> 16));
        __stcs(edge_index + capacity + pos, node0 + (e.x & 0xffff));
        __stcs(edge_attr + pos, __int_as_float(e.y));
      }
    }
  }
}

}  // namespace

// Shared memory the single-pass kernel needs for E entities; the caller falls back to the three-kernel form above 96 KB.
size_t edge_fused_smem(int E) { return (size_t)EF_WARPS * E * E * sizeof(int2); }
int edge_fused_tiles(int num_graphs) { return (num_graphs + EF_WARPS - 1) / EF_WARPS; }
// scratch: one status word per tile + the tile counter (zeroed here, stream-ordered)
size_t edge_fused_scratch_bytes(int num_graphs) { return sizeof(unsigned long long) * ((size_t)edge_fused_tiles(num_graphs) + 1); }

cudaError_t launch_edge_list_fused(const float* adj, int num_graphs, int E, float thr, int inclusive, int repeat,
                                   long long capacity, void* scratch, long long* graph_offsets, long long* edge_index,
                                   float* edge_attr, long long* nnz_out, cudaStream_t st) {
  const int tiles = edge_fused_tiles(num_graphs);
  if (tiles == 0) {
    cudaError_t e = cudaMemsetAsync(graph_offsets, 0, sizeof(long long), st);
    if (e == cudaSuccess && nnz_out) e = cudaMemsetAsync(nnz_out, 0, sizeof(long long), st);
    return e;
  }
  cudaError_t e = cudaMemsetAsync(scratch, 0, edge_fused_scratch_bytes(num_graphs), st);
  if (e != cudaSuccess) return e;
  unsigned long long* status = reinterpret_cast<unsigned long long*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(status + tiles);
  const size_t smem = edge_fused_smem(E);
  e = cudaFuncSetAttribute(edge_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  edge_fused_kernel<<<tiles, EF_WARPS * 32, smem, st>>>(adj, num_graphs, E, thr, inclusive, repeat, capacity, tiles, status, counter,
                                                        graph_offsets, edge_index, edge_attr, nnz_out);
  return cudaGetLastError();
}

}  // namespace fm
