// Agent-warp kernels, one-shot launches: one CTA per tile of 32 consecutive envs (layout and tile body: fm_aw.cuh).
//   aw_kernel<N, O, 0, NF>   one env step of the tiles [env_begin, env_end)   (FM_ROLL=0 diagnostic path; the product
//                            path for steps is the persistent rollout kernel, fm_roll.cu)
//   aw_kernel<N, O, 1, NF>   masked reset + observe (reset() path)
//
// Why this mapping (measured on B200, C2 = 65 536 envs x 3 agents, profiles/):
//   * group-per-env (fm_kernels.cu, G lanes per env, runtime N): 35 M warp instructions per step,
//     issue bound, 55 us / step (34 % of the HBM roofline).
//   * env-tile (round-1 experiment, removed): fine-grained work items spread over the warps of a CTA
//     through shared memory; 19.2 M warp instructions, 30 % of them in a gather-style output
//     emission, 38 us / step (49 %).
//   * here the compute phases are register resident (no item descriptors, no smem round trips), the
//     distances between static entities are computed once per episode and kept in the state block,
//     and every output is written lane = env into a shared-memory IMAGE of the API layout (odd strides:
//     conflict free) which one thread hands to the copy engine (TMA bulk stores).
#include <cstdlib>

#include "fm_aw.cuh"

namespace fm {

template <int N, int O, int MODE, int NF, int W = 0>
__global__ void __launch_bounds__((AwLayout<N, O, W>::THREADS), (AwLayout<N, O, W>::MIN_CTAS))
aw_kernel(const __grid_constant__ DevParams p) {
  extern __shared__ __align__(16) float smem[];
  // Programmatic dependent launch (fm_step with FM_STEP_PDL, MODE 0): the NEXT step's kernel may be scheduled as soon as every
  // CTA of this one has started -- its CTAs take the slots the last, partial wave leaves free and wait here until this grid
  // has completed and its memory is visible.  Without the launch attribute both instructions are no-ops.
  if (MODE == 0) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int env0 = p.env_begin + blockIdx.x * 32;
  AwRoll rs{};
  const FmOutputs out{p.o_obs, p.o_node, p.o_adj, p.o_rew, p.o_done, p.o_info};
  const AwIo io{p.act_idx, p.act_onehot, p.reset_mask, &out};
  aw_tile<N, O, MODE, NF, false, W>(p, io, env0, min(32, p.env_end - env0), smem, rs);
}

// =============================================================================================
// Distances between static entities from the SoA state (after fm_set_state injected positions).
__global__ void aw_static_kernel(const DevParams p, float* __restrict__ sdist) {
  const int M = p.N + p.O + p.W, SP = M * (M - 1) / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)p.Bp * SP) return;
  const int q = (int)(t / p.Bp), env = (int)(t % p.Bp);
  int a = 0, rem = q;
  while (rem >= M - 1 - a) { rem -= M - 1 - a; ++a; }
  const int b = a + 1 + rem;
  // static entity s: landmark, obstacle, or wall midpoint ((0, axis) for 'H', (axis, 0) for 'V')
  auto X = [&](int s) {
    if (s < p.N) return p.lx[(size_t)s * p.Bp + env];
    if (s < p.N + p.O) return p.ox[(size_t)(s - p.N) * p.Bp + env];
    const size_t wi = (size_t)(s - p.N - p.O) * p.Bp + env;
    return p.wor[wi] == 0 ? 0.0f : p.wax[wi];
  };
  auto Y = [&](int s) {
    if (s < p.N) return p.ly[(size_t)s * p.Bp + env];
    if (s < p.N + p.O) return p.oy[(size_t)(s - p.N) * p.Bp + env];
    const size_t wi = (size_t)(s - p.N - p.O) * p.Bp + env;
    return p.wor[wi] == 0 ? p.wax[wi] : 0.0f;
  };
  const size_t at = p.sd_env_stride ? (size_t)env * p.sd_env_stride + q : (size_t)q * p.Bp + env;
  sdist[at] = (float)dist64(X(a), Y(a), X(b), Y(b));
}

// =============================================================================================
// Host side.
template <int N, int O>
static cudaError_t aw_launch_no(const DevParams& p, cudaStream_t st, bool is_reset) {
  using L = AwLayout<N, O>;
  const int blocks = (p.env_end - p.env_begin + 31) / 32;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = (size_t)L::WORDS * sizeof(float);
  if (p.pdl && !is_reset && !p.feat_global) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(L::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, aw_kernel<N, O, 0, NODE_F>, p);
  }
  if (p.feat_global) {
    if (is_reset) aw_kernel<N, O, 1, NODE_F_GLOBAL><<<blocks, L::THREADS, smem, st>>>(p);
    else aw_kernel<N, O, 0, NODE_F_GLOBAL><<<blocks, L::THREADS, smem, st>>>(p);
  } else {
    if (is_reset) aw_kernel<N, O, 1, NODE_F><<<blocks, L::THREADS, smem, st>>>(p);
    else aw_kernel<N, O, 0, NODE_F><<<blocks, L::THREADS, smem, st>>>(p);
  }
  return cudaGetLastError();
}

template <int N, int O>
static cudaError_t aw_prepare_no() {
  using L = AwLayout<N, O>;
  const int smem = L::WORDS * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(aw_kernel<N, O, 0, NODE_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 1, NODE_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 0, NODE_F_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 1, NODE_F_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (const char* v = getenv("FM_CARVEOUT"); v && v[0] == '1') {      // A/B: largest shared-memory carve-out instead of the driver's pick
    if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 0, NODE_F>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 1, NODE_F>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  return e;
}

// Walls (N4): the same tile body with W wall entities behind the obstacles (relative node features only).
template <int N, int O, int W>
static cudaError_t aw_launch_w(const DevParams& p, cudaStream_t st, bool is_reset) {
  using L = AwLayout<N, O, W>;
  const int blocks = (p.env_end - p.env_begin + 31) / 32;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = (size_t)L::WORDS * sizeof(float);
  if (p.pdl && !is_reset) {                            // see aw_launch_no
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(L::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, aw_kernel<N, O, 0, NODE_F, W>, p);
  }
  if (is_reset) aw_kernel<N, O, 1, NODE_F, W><<<blocks, L::THREADS, smem, st>>>(p);
  else aw_kernel<N, O, 0, NODE_F, W><<<blocks, L::THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

template <int N, int O, int W>
static cudaError_t aw_prepare_w() {
  const int smem = AwLayout<N, O, W>::WORDS * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(aw_kernel<N, O, 0, NODE_F, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_kernel<N, O, 1, NODE_F, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  return e;
}

bool aw_supported(int N, int O, int W) {
  if (W > 0) {
#define X(n, o, w) if (N == n && O == o && W == w) return true;
    FM_AW_WALL_CASES(X)
#undef X
    return false;
  }
#define X(n, o) if (N == n && O == o) return true;
  FM_AW_CASES(X)
#undef X
  return false;
}

int aw_stats_rows(int B) { return (B + 31) / 32; }

cudaError_t aw_prepare(const DevParams& p) {
#define X(n, o, w) if (p.N == n && p.O == o && p.W == w) return aw_prepare_w<n, o, w>();
  FM_AW_WALL_CASES(X)
#undef X
  if (p.W > 0) return cudaErrorInvalidValue;
#define X(n, o) if (p.N == n && p.O == o) return aw_prepare_no<n, o>();
  FM_AW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t aw_launch(const DevParams& p, cudaStream_t st, bool is_reset) {
#define X(n, o, w) if (p.N == n && p.O == o && p.W == w) return aw_launch_w<n, o, w>(p, st, is_reset);
  FM_AW_WALL_CASES(X)
#undef X
  if (p.W > 0) return cudaErrorInvalidValue;
#define X(n, o) if (p.N == n && p.O == o) return aw_launch_no<n, o>(p, st, is_reset);
  FM_AW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t launch_static_dists(const DevParams& p, cudaStream_t st) {
  const int M = p.N + p.O + p.W, SP = M * (M - 1) / 2;
  if (SP == 0 || !p.sdist) return cudaSuccess;
  const long long total = (long long)p.Bp * SP;
  aw_static_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(p, p.sdist);
  return cudaGetLastError();
}

}  // namespace fm
