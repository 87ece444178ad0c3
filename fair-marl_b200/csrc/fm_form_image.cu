// Formation family, split step path: the IMAGE kernel (see fm_formation.cu for the path as a whole).
#include "fm_form.cuh"

namespace fm {

#ifndef FM_FI_WARPS
#define FM_FI_WARPS 4                  // A-B knob: warps per image CTA (4 or 8)
#endif
constexpr int FI_WARPS = FM_FI_WARPS, FI_ENVS = 16, FI_PARTS = FI_WARPS * (32 / FI_ENVS);

// The share of PART in one env's images: every (ego, entity) row and adj pair whose running index is PART modulo the
// part count, with all indices compile-time values (f_row / f_adj_elem fold to immediate shared-memory offsets, the row
// type branches disappear) and independent of each other (the stores of one row overlap the arithmetic of the next).
template <int N, int O, int PART>
__device__ __forceinline__ void image_part(const float* __restrict__ r, float* __restrict__ node, float* __restrict__ adj) {
  constexpr int E = 2 * N + O;
  if (node) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int en = 0; en < E; ++en)
        if ((i * E + en) % FI_PARTS == PART) f_row(r, N, O, i, en, node + (i * E + en) * F_NODE);
    }
  }
  if (adj) {
    double X[E], Y[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { X[e] = (double)r[2 * e]; Y[e] = (double)r[2 * e + 1]; }
    int q = 0;
#pragma unroll
    for (int a = 0; a < E; ++a) {
      if (a % FI_PARTS == PART) adj[a * E + a] = 0.0f;
#pragma unroll
      for (int c = a + 1; c < E; ++c) {
        if (q % FI_PARTS == PART) { const float d = (float)dn(X[a] - X[c], Y[a] - Y[c]); adj[a * E + c] = d; adj[c * E + a] = d; }
        ++q;
      }
    }
  }
}

// One CTA = FI_ENVS (16) consecutive envs: a half tile, so that its images are 31 KB and 7 CTAs share an SM (the kernel
// is a chain recipe load -> rows -> bulk store per CTA: what hides the two memory latencies is CTAs in flight; with whole
// tiles, 62 KB, 3 CTAs / SM it ran at 30 us, profiles/r02_h).  Lane = (env, half): the two half-warps of a warp run
// different parts -- divergent, each part's instructions issue once for 16 lanes; the kernel is not issue bound.
template <int N, int OT>
__global__ void __launch_bounds__(FI_WARPS * 32) formation_image_kernel(const FormParams p) {
  constexpr int O = OT, E = 2 * N + O, NE = N * E, EE = E * E, NODE_W = NE * F_NODE;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int env0 = blockIdx.x * FI_ENVS;
  const int nenv = min(FI_ENVS, p.B - env0);
  const FormTile t = form_tile(N, OT);
  const int REC_W = FI_ENVS * t.rec_stride;                                 // floats; the block starts 16-byte aligned
  float* rec = smem;
  float* node = rec + ((REC_W + 3) & ~3);
  float* adj = node + FI_ENVS * NODE_W;
  if (p.ready) {                                                            // programmatic dependent launch: wait for this CTA's 16 envs
    if (tid == 0) {
      const int* f = p.ready + blockIdx.x;
      int v;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (!v) __nanosleep(200);
      } while (!v);
    }
    __syncthreads();
  }
  {                                                                         // this CTA's recipes: 16-byte loads, L2 hits
    const float4* src = reinterpret_cast<const float4*>(p.rec + (size_t)blockIdx.x * REC_W);
    float4* dst = reinterpret_cast<float4*>(rec);
#pragma unroll 2
    for (int k = tid; k < (REC_W + 3) / 4; k += FI_WARPS * 32) dst[k] = __ldcs(src + k);
  }
  __syncthreads();
  if (p.ready && tid == 0) p.ready[blockIdx.x] = 0;                         // consumed; the next step's producer runs after this grid
  float* g_node = p.out.node_obs ? p.out.node_obs + (size_t)env0 * NODE_W : nullptr;
  float* g_adj = p.out.adj ? p.out.adj + (size_t)env0 * EE : nullptr;
  const int el = lane & (FI_ENVS - 1);
  if (el < nenv) {
    const float* r = rec + el * t.rec_stride;
    float* nimg = g_node ? node + el * NODE_W : nullptr;
    float* aimg = g_adj ? adj + el * EE : nullptr;
    switch (w * (32 / FI_ENVS) + lane / FI_ENVS) {
      case 0: image_part<N, O, 0>(r, nimg, aimg); break;
      case 1: image_part<N, O, 1>(r, nimg, aimg); break;
      case 2: image_part<N, O, 2>(r, nimg, aimg); break;
      case 3: image_part<N, O, 3>(r, nimg, aimg); break;
      case 4: image_part<N, O, 4>(r, nimg, aimg); break;
      case 5: image_part<N, O, 5>(r, nimg, aimg); break;
      case 6: image_part<N, O, 6>(r, nimg, aimg); break;
      case 7: image_part<N, O, 7>(r, nimg, aimg); break;
      case 8: image_part<N, O, 8>(r, nimg, aimg); break;
      case 9: image_part<N, O, 9>(r, nimg, aimg); break;
      case 10: image_part<N, O, 10>(r, nimg, aimg); break;
      case 11: image_part<N, O, 11>(r, nimg, aimg); break;
      case 12: image_part<N, O, 12>(r, nimg, aimg); break;
      case 13: image_part<N, O, 13>(r, nimg, aimg); break;
      case 14: image_part<N, O, 14>(r, nimg, aimg); break;
      default: image_part<N, O, 15>(r, nimg, aimg); break;
    }
  }
  __syncthreads();
  const bool bulk = nenv == FI_ENVS && aligned16(g_node) && aligned16(g_adj);
  if (bulk) {
    if (tid == 0) {
      const uint64_t pol = evict_first_policy();
      fence_async_smem();
      if (g_node) bulk_store(g_node, node, (uint32_t)(FI_ENVS * NODE_W * 4), pol);
      if (g_adj) bulk_store(g_adj, adj, (uint32_t)(FI_ENVS * EE * 4), pol);
      bulk_commit();
      bulk_wait_read<0>();
    }
  } else {
    if (g_node) for (int k = tid; k < nenv * NODE_W; k += FI_WARPS * 32) __stcs(g_node + k, node[k]);
    if (g_adj) for (int k = tid; k < nenv * EE; k += FI_WARPS * 32) __stcs(g_adj + k, adj[k]);
  }
}

template <int N, int OT>
static cudaError_t launch_image(const FormParams& p, cudaStream_t st) {
  constexpr int E = 2 * N + OT;
  static_assert(FI_PARTS == 8 || FI_PARTS == 16, "formation_image_kernel dispatches up to 16 parts");
  const FormTile t = form_tile(N, OT);
  const size_t smem = (size_t)(((FI_ENVS * t.rec_stride + 3) & ~3) + FI_ENVS * (N * E * F_NODE + E * E)) * sizeof(float);
  static int attr_device = -1;                         // opt-in shared-memory size: once per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != attr_device) {
    e = cudaFuncSetAttribute(formation_image_kernel<N, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(formation_image_kernel<N, OT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_device = dev;
  }
  const int blocks = (p.B + FI_ENVS - 1) / FI_ENVS;
  if (!p.ready) {
    formation_image_kernel<N, OT><<<blocks, FI_WARPS * 32, smem, st>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(FI_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, formation_image_kernel<N, OT>, p);
}

template <int N>
static cudaError_t launch_image_n(const FormParams& p, cudaStream_t st) {
  switch (p.O) {
    case 0: return launch_image<N, 0>(p, st);
    case 1: return launch_image<N, 1>(p, st);
    case 2: return launch_image<N, 2>(p, st);
    case 3: return launch_image<N, 3>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

// Pending resets (fm_form.cuh): every env whose block was not drawn for its current episode key gets a fresh one.  Almost
// every thread leaves at once; the ones that stay walk the serial reset, off the step's critical path.
template <int N>
__global__ void __launch_bounds__(64) formation_prefetch_kernel(const FormParams p) {
  const int b = blockIdx.x * 64 + threadIdx.x;
  if (b >= p.B) return;
  const int key = p.st.episode[b];
  if (__float_as_int(p.pend[b]) == key) return;
  FEnv<N> e;
  e.episode = key;
  for (int i = 0; i < N; ++i) e.mint[i] = 0.0f;
  f_reset<N>(p, b, e);
  f_pending_write<N>(p, b, e, key);                                          // (tag last, behind a fence)
}

cudaError_t launch_formation_prefetch(const FormParams& p, cudaStream_t st) {
  const int blocks = (p.B + 63) / 64;
  switch (p.N) {
    case 2: formation_prefetch_kernel<2><<<blocks, 64, 0, st>>>(p); break;
    case 3: formation_prefetch_kernel<3><<<blocks, 64, 0, st>>>(p); break;
    case 4: formation_prefetch_kernel<4><<<blocks, 64, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// node_obs / adj of one step from the recipes the logic kernel left in p.rec (N <= 4, O <= 3)
cudaError_t launch_formation_image(const FormParams& p, cudaStream_t st) {
  switch (p.N) {
    case 2: return launch_image_n<2>(p, st);
    case 3: return launch_image_n<3>(p, st);
    case 4: return launch_image_n<4>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace fm
