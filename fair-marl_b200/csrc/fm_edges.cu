// Policy-side edge list (TransformerConvNet.process_adj, onpolicy/algorithms/utils/gnn_new.py:381-413): mask
// (adj < max_edge_dist) & (adj > 0), edges in (b, i, j) order, every graph emitted `repeat` times consecutively.
// Two forms live here (a third, round 1's count / one-block scan / emit, in fm_kernels.cu); all three are bit-identical and
// under test, FM_EDGE_FORM selects one (fm_abi.cu):
//   * the STREAMED form, the default (second half of this file): count -> offsets -> persistent emission;
//   * the SINGLE-PASS form (first half, opt-in), kept because it reads adj once.  A CTA of 8 warps owns 8 consecutive graphs:
//   1. each warp reads its graph once (8 loads in flight per lane) and compacts the edges -- ballot + popc, (b, i, j) order
//      -- into a shared-memory list (row, column, distance);
//   2. the CTA's edge count is published and its exclusive prefix fetched by a decoupled look-back over one 64-bit status
//      word per CTA (flag in the top two bits, count below: self-contained, no fence needed); tiles are claimed from a
//      counter, so a CTA only ever waits for CTAs that are already running;
//   3. each warp writes its graph's offsets and, per copy, the list with full consecutive lanes: 8-byte / 4-byte stores
//      to consecutive addresses.
//   Measured no faster than three kernels (448 us at config 3): 42 % of its stall samples are the CTA-wide barrier behind
//   the look-back (profiles/r02_j_edge_fused_kernel_ncu.txt).
#include <cstdlib>

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

// =============================================================================================================
// Streamed form (the default): count -> offsets -> emit, with nothing between the graphs of the emission.
//
// What the launch list and the source view of the two forms above say (profiles/r02_j_*): the emission of the three-kernel
// form moves 1.1 GB in 356 us (3.1 TB/s) with 32 768 short-lived CTAs, each of which walks counts -> barrier -> CTA offset
// -> adj loads -> stores as ONE dependent chain before it retires (37 waves of them per SM); the single-pass form keeps
// whole CTAs at the barrier behind its look-back (42 % of the stall samples).  Here
//   1. es_count_kernel: edges per graph, one warp per graph, nothing else;
//   2. es_offsets_kernel: one CTA per tile of ES_TILE graphs publishes the tile's edge total (one self-contained 64-bit
//      status word, tiles claimed from a counter so that a CTA only waits for CTAs that are already running), adds the
//      totals of ALL tiles in front of it (no chain: every total is published before its CTA waits for anything) and scans
//      its own counts: graph_offsets of EVERY graph copy, the end sentinel and nnz.  5 us where the one-block scan took 26;
//   3. es_emit_kernel: a persistent grid of warps walks (graph, 320-entry chunk) items with a grid stride; the NEXT item's
//      distances, offset and count are loaded into registers before the current item is compacted (ballot + popc, (b, i, j)
//      order) and stored, so every warp always has 1.3 KB of loads in flight and no warp ever waits for another one.
//      Specialised on the comparison and on the common call (one copy per graph, a list that fits its capacity by
//      construction): the first version took 857 warp instructions per graph at 77 % issue utilisation
//      (profiles/r02_es_emit_v1_ncu.txt), most of them 64-bit position / capacity arithmetic per copy and per entry.
// Same bits as the other two forms (tests/test_gpu_parity.py::test_edge_list_corner_cases runs all three).
constexpr int ES_TILE = 2048;                          // graphs per tile of the offsets kernel
constexpr int ES_WPB = 8;                              // warps per CTA of the count / emit kernels
constexpr int ES_U = 10;                               // loads in flight per lane: 320 entries cover E = 17 in one item
constexpr unsigned long long ES_FLAG = 1ull << 62;

__global__ void __launch_bounds__(ES_WPB * 32)
es_count_kernel(const float* __restrict__ adj, int num_graphs, int EE, float thr, int inclusive, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * ES_WPB + (threadIdx.x >> 5);
  if (g >= num_graphs) return;
  const float* a = adj + (size_t)g * EE;
  int c = 0;
#pragma unroll 8
  for (int q = lane; q < EE; q += 32) c += ef_pred(__ldg(a + q), thr, inclusive) ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(FULL, c, off);
  if (lane == 0) counts[g] = c;
}

__global__ void __launch_bounds__(256)
es_offsets_kernel(const int* __restrict__ counts, unsigned long long* __restrict__ status, unsigned int* __restrict__ counter,
                  int num_graphs, int repeat, long long* __restrict__ graph_offsets, long long* __restrict__ nnz_out) {
  constexpr int PER = ES_TILE / 256;                   // consecutive graphs per thread
  __shared__ long long wsum[8];
  __shared__ int wtot[8];
  __shared__ int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_tile = (int)atomicAdd(counter, 1u);
  __syncthreads();
  const int tile = s_tile;
  // this thread's PER counts, their sum, and its inclusive prefix inside the warp
  const int g0 = tile * ES_TILE + tid * PER;
  int c[PER], tot = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) { c[k] = (g0 + k < num_graphs) ? __ldg(counts + g0 + k) : 0; tot += c[k]; }
  int incl = tot;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += o; }
  if (lane == 31) wtot[w] = incl;
  __syncthreads();
  int wbefore = 0, tile_total = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { wbefore += (k < w) ? wtot[k] : 0; tile_total += wtot[k]; }
  if (tid == 0) ef_store(status + tile, ES_FLAG | (unsigned long long)tile_total);
  // edges in front of this tile: the totals of all earlier tiles (each is published before its CTA waits for anything)
  long long before = 0;
  for (int t = tid; t < tile; t += 256) {
    unsigned long long word;
    while (((word = ef_load(status + t)) & ES_FLAG) == 0) __nanosleep(20);
    before += (long long)(word & (ES_FLAG - 1ull));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) before += __shfl_xor_sync(FULL, before, off);
  if (lane == 0) wsum[w] = before;
  __syncthreads();
  long long base = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) base += wsum[k];
  long long excl = base + wbefore + (incl - tot);      // edges (of one copy) in front of graph g0
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int g = g0 + k;
    if (g < num_graphs) {
      const long long b0 = excl * repeat;
      for (int cp = 0; cp < repeat; ++cp) graph_offsets[(size_t)g * repeat + cp] = b0 + (long long)cp * c[k];
    }
    excl += c[k];
  }
  if (tile == gridDim.x - 1 && tid == 0) {
    const long long total = (base + tile_total) * repeat;
    graph_offsets[(size_t)num_graphs * repeat] = total;
    if (nnz_out) *nnz_out = total;
  }
}

// INCL: adj <= thr (update_graph) instead of adj < thr (process_adj).  FAST: one copy per graph and capacity >= num_graphs *
// E * E, i.e. no position can fall outside the list: no per-copy loop, no capacity tests.
template <bool INCL, bool FAST>
__global__ void __launch_bounds__(ES_WPB * 32, FAST ? 5 : 4)   // 48 / 64 registers
es_emit_kernel(const float* __restrict__ adj, int num_graphs, int E, float thr, int repeat, long long capacity,
               const int* __restrict__ counts, const long long* __restrict__ graph_offsets, long long* __restrict__ edge_index,
               float* __restrict__ edge_attr) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * ES_WPB;
  const int EE = E * E;
  const unsigned magic = ((1u << 20) + E - 1) / E;     // q / E == (q * magic) >> 20 for q < 2^20 / E (E <= 100: the launcher checks)
  unsigned lt;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
  int g = blockIdx.x * ES_WPB + (threadIdx.x >> 5), q0 = 0;
  if (g >= num_graphs) return;
  float nv[ES_U];
  {
    const float* a = adj + (size_t)g * EE + lane;
#pragma unroll
    for (int u = 0; u < ES_U; ++u) nv[u] = (u * 32 + lane < EE) ? __ldcs(a + u * 32) : 0.0f;
  }
  long long nbase = __ldg(graph_offsets + (size_t)g * repeat);
  int ncnt = __ldg(counts + g);
  // the current graph: where its first copy starts in the three output arrays, its first node id, edges so far
  long long* prow = edge_index;
  float* pattr = edge_attr;
  long long node_g = 0, lim = 0;
  int cnt = 0, run = 0;
  while (true) {
    float cv[ES_U];
#pragma unroll
    for (int u = 0; u < ES_U; ++u) cv[u] = nv[u];
    if (q0 == 0) {
      prow = edge_index + nbase; pattr = edge_attr + nbase;
      node_g = (long long)g * repeat * E;
      cnt = ncnt; run = 0;
      // positions relative to nbase are tested against `lim`: what is left of the capacity when the list is truncated
      if (!FAST) lim = (nbase + (long long)cnt * repeat <= capacity) ? 0x7fffffffffffffffLL : capacity - nbase;
    }
    // the next item: the next chunk of this graph, or the first chunk of this warp's next graph
    int g2 = g, q2 = q0 + 32 * ES_U;
    if (q2 >= EE) { g2 = g + stride; q2 = 0; }
    const bool more = g2 < num_graphs;
    if (more) {
      const float* a = adj + (size_t)g2 * EE + q2 + lane;
      const int left = EE - q2 - lane;                 // entries of the graph from this lane's first one on
#pragma unroll
      for (int u = 0; u < ES_U; ++u) nv[u] = (u * 32 < left) ? __ldcs(a + u * 32) : 0.0f;
      if (q2 == 0) { nbase = __ldg(graph_offsets + (size_t)g2 * repeat); ncnt = __ldg(counts + g2); }
    }
    if (cnt > 0) {                                     // warp-uniform: a graph without edges stores nothing
      const int left = EE - q0 - lane;
      unsigned q = (unsigned)(q0 + lane);
#pragma unroll
      for (int u = 0; u < ES_U; ++u) {
        if (q0 + u * 32 >= EE) break;                  // warp-uniform
        const float d = cv[u];
        const bool pr = (u * 32 < left) && (INCL ? (d <= thr) : (d < thr)) && (d > 0.0f);
        const unsigned b = __ballot_sync(FULL, pr);
        if (pr) {
          const int k = run + __popc(b & lt);
          const unsigned r = (q * magic) >> 20;
          const unsigned c = q - r * (unsigned)E;
          if (FAST) {
            __stcs(prow + k, node_g + r);
            __stcs(prow + capacity + k, node_g + c);
            __stcs(pattr + k, d);
          } else {
            long long pos = k, node0 = node_g;
            for (int cp = 0; cp < repeat; ++cp) {
              if (pos < lim) {
                __stcs(prow + pos, node0 + r);
                __stcs(prow + capacity + pos, node0 + c);
                __stcs(pattr + pos, d);
              }
              pos += cnt; node0 += E;
            }
          }
        }
        run += __popc(b);
        q += 32;
      }
    }
    if (!more) break;
    g = g2; q0 = q2;
  }
}

}  // namespace

// ---- streamed form: scratch = one status word per offsets tile + the tile counter (zeroed here), then the per-graph counts
int edge_stream_tiles(int num_graphs) { return (num_graphs + ES_TILE - 1) / ES_TILE; }
size_t edge_stream_scratch_bytes(int num_graphs) {
  return sizeof(unsigned long long) * ((size_t)edge_stream_tiles(num_graphs) + 1) + sizeof(int) * (size_t)(num_graphs > 0 ? num_graphs : 1);
}

cudaError_t launch_edge_list_stream(const float* adj, int num_graphs, int E, float thr, int inclusive, int repeat,
                                    long long capacity, void* scratch, long long* graph_offsets, long long* edge_index,
                                    float* edge_attr, long long* nnz_out, cudaStream_t st) {
  const int tiles = edge_stream_tiles(num_graphs);
  if (tiles == 0) {
    cudaError_t e = cudaMemsetAsync(graph_offsets, 0, sizeof(long long), st);
    if (e == cudaSuccess && nnz_out) e = cudaMemsetAsync(nnz_out, 0, sizeof(long long), st);
    return e;
  }
  unsigned long long* status = reinterpret_cast<unsigned long long*>(scratch);
  unsigned int* counter = reinterpret_cast<unsigned int*>(status + tiles);
  int* counts = reinterpret_cast<int*>(status + tiles + 1);
  cudaError_t e = cudaMemsetAsync(status, 0, sizeof(unsigned long long) * ((size_t)tiles + 1), st);
  if (e != cudaSuccess) return e;
  const int blocks = (num_graphs + ES_WPB - 1) / ES_WPB;
  es_count_kernel<<<blocks, ES_WPB * 32, 0, st>>>(adj, num_graphs, E * E, thr, inclusive, counts);
  es_offsets_kernel<<<tiles, 256, 0, st>>>(counts, status, counter, num_graphs, repeat, graph_offsets, nnz_out);
  static int sms = 0;                                  // persistent grid: the CTAs one device holds (5 or 4 per SM)
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    sms = n;
  }
  const bool fast = repeat == 1 && capacity >= (long long)num_graphs * E * E;
  const int resident = sms * (fast ? 5 : 4);
  const int grid = blocks < resident ? blocks : resident;
#define FM_ES_EMIT(INCL, FAST)                                                                                              \
  es_emit_kernel<INCL, FAST><<<grid, ES_WPB * 32, 0, st>>>(adj, num_graphs, E, thr, repeat, capacity, counts, graph_offsets, \
                                                           edge_index, edge_attr)
  if (inclusive) { if (fast) FM_ES_EMIT(true, true); else FM_ES_EMIT(true, false); }
  else { if (fast) FM_ES_EMIT(false, true); else FM_ES_EMIT(false, false); }
#undef FM_ES_EMIT
  return cudaGetLastError();
}

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
