// Persistent rollout kernel of the agent-warp mapping: T consecutive env steps of a whole env range in ONE launch.
//
// Envs are independent and a step of tile k (32 consecutive envs) depends only on the previous step of the same tile,
// so a T-step rollout is a set of T x tiles work items (t, k) with chain dependencies (t - 1, k) -> (t, k).  The grid
// is one wave of resident CTAs (SMs x CTAs/SM); every CTA claims items from a global counter in STEP-MAJOR order and
// runs the tile body of fm_aw.cuh on them.  Item (t, k) may start once flags[k] >= t; with tiles >> resident CTAs the
// predecessor was claimed (tiles - resident) items earlier and the wait never spins.  There is no wave quantisation, no
// launch gap and no drain between steps: the copy engine is still streaming tile (t, k)'s node_obs image out of shared
// memory while the CTA loads the state of its next item.  Deadlock free: an item only waits for an item with a smaller
// index, which has been claimed by a CTA that is running.
//
// The control block (next item, exit count, per-tile flags) is zero between launches: the last CTA to leave resets it,
// so a launch captured in a CUDA graph can be replayed.  T = 1 is the single-step entry point (fm_step): same kernel,
// no flags.
#include "fm_aw.cuh"

namespace fm {

constexpr int ROLL_MAX_SMS = 256;
struct RollCtl {
  unsigned long long next;   // items claimed so far
  unsigned int done;         // CTAs that have left
  unsigned int pad;
  unsigned int sm_slot[ROLL_MAX_SMS];   // CTAs of this launch that have started on each SM (start-up stagger)
  int flags[1];              // [tiles] steps of this launch completed per tile
};

struct RollArgs {
  DevParams p;
  int T, ntiles, early;
  int stagger_ns;            // start-up delay per CTA slot of an SM (see the kernel)
  RollCtl* ctl;
  const int* act_idx;        // step t reads act_idx + t * act_stride ([B, N] each), or
  const float* act_onehot;   //             act_onehot + t * act_stride ([B, N, 5] each)
  long long act_stride;
  FmOutputs outs[FM_ROLL_MAX_STEPS];
};

template <int N, int O, int NF>
__global__ void __launch_bounds__(AwLayout<N, O>::THREADS, AwLayout<N, O>::MIN_CTAS)
aw_roll_kernel(const __grid_constant__ RollArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_next[2], s_last;
  __shared__ FmOutputs s_outs[FM_ROLL_MAX_STEPS];
  const DevParams& p = a.p;
  const int tid = threadIdx.x;
  const int total = a.T * a.ntiles;
  const bool late = a.T > 1 && a.early == 0;
  // Every item is claimed from the counter -- also the first one of a CTA -- so that "claimed" implies "held by a running
  // CTA": an item only ever waits for claimed items, whatever else occupies the GPU or however large the grid is.
  AwRoll rs;
  rs.flags = a.ctl->flags; rs.prev_tile = -1; rs.prev_t = 0; rs.early = a.early != 0; rs.multi = a.T > 1;
  rs.s_next = &s_next[0]; rs.nx = 0; rs.total = total; rs.ntiles = a.ntiles; rs.next_ready = false;
  for (int k = tid; k < a.T; k += blockDim.x) s_outs[k] = a.outs[k];
  if (tid == 0) {
    // Start-up stagger.  All CTAs of the wave start together and every item costs the same, so without it the CTAs of
    // an SM stay in phase for the whole launch: they all compute (HBM idle), then all store (SM idle).  The k-th CTA
    // to start on an SM waits k x stagger_ns; the phases stay spread because each CTA chains its items back to back.
    if (a.stagger_ns > 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const unsigned k = atomicAdd(&a.ctl->sm_slot[smid % ROLL_MAX_SMS], 1u);
      const unsigned long long wait = (unsigned long long)k * (unsigned)a.stagger_ns;
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do { __nanosleep(128); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (wait > 0 && t1 - t0 < wait);
    }
    s_next[1] = (int)atomicAdd(&a.ctl->next, 1ull);
  }
  __syncthreads();
  int item = s_next[1];
  int it = 0;
  while (item < total) {
    const int t = item / a.ntiles, tile = item - t * a.ntiles;
    rs.t = t; rs.tile = tile;
    rs.s_next = &s_next[it & 1];                  // double buffered: a slow reader of the previous item's word is never overtaken
    if (tid == 0) rs.nx = (int)atomicAdd(&a.ctl->next, 1ull);   // consumed at barrier #0 (latency hidden behind the state loads)
    if (late) {
      // Late release (some steps of the launch share output arrays): the previous item's outputs must be in global
      // memory before its tile's next step may run -- and BEFORE this CTA waits, since its next item may be that step.
      __syncthreads();                            // the previous item's vectorised stores (ragged tiles) are issued
      if (tid == 0 && rs.prev_tile >= 0) {
        bulk_wait_all();
        __threadfence();
        st_release_gpu(rs.flags + rs.prev_tile, rs.prev_t + 1);
        rs.prev_tile = -1;
      }
    }
    if (rs.multi && t > 0 && !rs.next_ready)       // (normally seen satisfied during the previous item, fm_aw.cuh barrier #1)
      while (ld_acquire_gpu(rs.flags + tile) < t) __nanosleep(64);
    rs.next_ready = false;
    AwIo io;
    io.act_idx = a.act_idx ? a.act_idx + (size_t)t * a.act_stride : nullptr;
    io.act_onehot = a.act_onehot ? a.act_onehot + (size_t)t * a.act_stride : nullptr;
    io.reset_mask = nullptr;
    io.out = &s_outs[t];
    const int env0 = p.env_begin + tile * 32;
    aw_tile<N, O, 0, NF, true>(p, io, env0, min(32, p.env_end - env0), smem, rs);
    rs.prev_tile = tile; rs.prev_t = t;
    item = s_next[it & 1];                        // written by thread 0 before barrier #0 of this item
    ++it;
  }
  // ---- leave: the last image has been read (or, late release, written); the last CTA resets the control block ----
  if (late) __syncthreads();
  if (tid == 0) {
    if (rs.prev_tile >= 0) {
      if (late) {
        bulk_wait_all();
        __threadfence();
        st_release_gpu(rs.flags + rs.prev_tile, rs.prev_t + 1);
      } else {
        bulk_wait_read<0>();
      }
    }
    s_last = (atomicAdd(&a.ctl->done, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    if (rs.multi)
      for (int k = tid; k < a.ntiles; k += blockDim.x) a.ctl->flags[k] = 0;
    if (a.stagger_ns > 0)
      for (int k = tid; k < ROLL_MAX_SMS; k += blockDim.x) a.ctl->sm_slot[k] = 0u;
    if (tid == 0) { a.ctl->next = 0ull; a.ctl->done = 0u; }
  }
}

// =============================================================================================
// Host side.
size_t roll_ctl_bytes(int B) { return sizeof(RollCtl) + sizeof(int) * (size_t)((B + 31) / 32); }

template <int N, int O>
static cudaError_t roll_prepare_no(const DevParams& p, int* ctas_per_sm) {
  using L = AwLayout<N, O>;
  const int smem = L::WORDS * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(aw_roll_kernel<N, O, NODE_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(aw_roll_kernel<N, O, NODE_F_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  if (p.feat_global) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, aw_roll_kernel<N, O, NODE_F_GLOBAL>, L::THREADS, smem);
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, aw_roll_kernel<N, O, NODE_F>, L::THREADS, smem);
}

template <int N, int O>
static cudaError_t roll_launch_no(const DevParams& p, const RollLaunch& r, cudaStream_t st) {
  using L = AwLayout<N, O>;
  RollArgs a;
  a.p = p;
  a.T = r.num_steps;
  a.ntiles = (p.env_end - p.env_begin + 31) / 32;
  a.early = r.early;
  a.stagger_ns = r.stagger_ns;
  a.ctl = reinterpret_cast<RollCtl*>(r.ctl);
  a.act_idx = r.act_idx; a.act_onehot = r.act_onehot; a.act_stride = r.act_stride;
  for (int t = 0; t < r.num_steps; ++t) a.outs[t] = r.outs[t];
  const long long total = (long long)a.T * a.ntiles;
  if (total <= 0) return cudaSuccess;
  const int grid = (int)(total < (long long)r.max_ctas ? total : (long long)r.max_ctas);
  const size_t smem = (size_t)L::WORDS * sizeof(float);
  if (p.feat_global) aw_roll_kernel<N, O, NODE_F_GLOBAL><<<grid, L::THREADS, smem, st>>>(a);
  else aw_roll_kernel<N, O, NODE_F><<<grid, L::THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t roll_prepare(const DevParams& p, int* ctas_per_sm) {
#define X(n, o) if (p.N == n && p.O == o) return roll_prepare_no<n, o>(p, ctas_per_sm);
  FM_AW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t roll_launch(const DevParams& p, const RollLaunch& r, cudaStream_t st) {
#define X(n, o) if (p.N == n && p.O == o) return roll_launch_no<n, o>(p, r, st);
  FM_AW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace fm
