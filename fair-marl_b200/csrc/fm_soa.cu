// SoA observation mode and the non-finite guard (BASELINE north_star (c): "coalesced,
// 16-byte-vectorised SoA buffers").
//
// The step kernels emit the API layout the reference's consumers expect (obs [B, N, 7], node_obs [B, N, E, F], adj
// [B, E, E]: an env's values are contiguous).  A device-side consumer that walks ENVS in lanes -- a policy kernel with
// lane = env, a statistics pass -- wants the transpose: one plane per value, envs fastest.  observe_soa_kernel builds
// exactly the values of fm_observe (Scenario.observation, navigation_graph.py:826-857; graph_observation /
// _get_entity_feat_relative, :1079-1124, or _get_entity_feat_global, :1058-1077; World.calculate_distances,
// core.py:204-228) from the internal SoA state ([row][Bp], envs fastest) into
//   obs  [N][7][S]      node_obs  [N][E][F][S]      adj  [E][E][S]        S = fm_soa_stride(h) = Bp, envs fastest
// with a thread per 4 consecutive envs: every state read is one aligned 16-byte load, every output one aligned 16-byte
// streaming store, and a warp's accesses are 512 contiguous bytes.  Same arithmetic as the API-layout kernels (fp32
// differences of fp32 state, float64 distances rounded once, float64 statistics): tests/test_gpu_vec_env.py compares the
// two layouts bit for bit.
//
// finite_guard_kernel: core.py:392 divides by the pair distance; two entities at the same point give 0/0 in the
// reference (a latent NaN).  The kernels here return a zero force for dist == 0, so a NaN can only come
// in through fm_set_state or an overflowing rollout; this pass flags the envs whose dynamic state holds a non-finite
// value and counts them, for a caller that wants to know before the values reach a policy.
#include "fm_device.cuh"
#include "fm_launch.h"

namespace fm {

namespace {

struct F4 { float v[4]; };
__device__ __forceinline__ F4 ld4(const float* base, size_t row, int Bp, int b4) {
  const float4 q = *reinterpret_cast<const float4*>(base + row * (size_t)Bp + b4);
  return F4{{q.x, q.y, q.z, q.w}};
}
__device__ __forceinline__ void st4(float* base, size_t plane, int Bp, int b4, const F4& a) {
  __stcs(reinterpret_cast<float4*>(base + plane * (size_t)Bp + b4), make_float4(a.v[0], a.v[1], a.v[2], a.v[3]));
}
__device__ __forceinline__ F4 sub4(const F4& a, const F4& b) {
  return F4{{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2], a.v[3] - b.v[3]}};
}
__device__ __forceinline__ F4 splat4(float x) { return F4{{x, x, x, x}}; }

// position of entity e (agents, landmarks, obstacles)
__device__ __forceinline__ void entity_pos(const DevParams& p, int e, int b4, F4& x, F4& y) {
  const int N = p.N;
  if (e < N) { x = ld4(p.px, e, p.Bp, b4); y = ld4(p.py, e, p.Bp, b4); }
  else if (e < 2 * N) { x = ld4(p.lx, e - N, p.Bp, b4); y = ld4(p.ly, e - N, p.Bp, b4); }
  else { x = ld4(p.ox, e - 2 * N, p.Bp, b4); y = ld4(p.oy, e - 2 * N, p.Bp, b4); }
}

__global__ void __launch_bounds__(128) observe_soa_kernel(const DevParams p, float* __restrict__ obs, float* __restrict__ node,
                                                          float* __restrict__ adj) {
  const int b4 = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (b4 >= p.Bp) return;                                                  // Bp is a multiple of 4; padding envs hold zeros
  const int N = p.N, O = p.O, E = 2 * N + O, Bp = p.Bp;
  const int F = p.feat_global ? NODE_F_GLOBAL : NODE_F;
  // ---- fairness observation (navigation_graph.py:846-851): mean / (std + 1e-4) of the travelled distances until the agent's
  // first info_callback of the episode, the running world statistics afterwards -- float64, rounded once
  double mean_p[4], std_p[4];
  {
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < N; ++j) { const F4 pd = ld4(p.pdist, j, Bp, b4); for (int c = 0; c < 4; ++c) sum[c] = __dadd_rn(sum[c], (double)pd.v[c]); }
    const double inv_n = 1.0 / N;
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < 4; ++c) mean_p[c] = __dmul_rn(sum[c], inv_n);
    for (int j = 0; j < N; ++j) {
      const F4 pd = ld4(p.pdist, j, Bp, b4);
      for (int c = 0; c < 4; ++c) q[c] = sq_acc(q[c], __dsub_rn((double)pd.v[c], mean_p[c]));
    }
    for (int c = 0; c < 4; ++c) std_p[c] = std_from_q(q[c], inv_n);
  }
  const F4 dmean = ld4(p.dmean, 0, Bp, b4), dstd = ld4(p.dstd, 0, Bp, b4);
  for (int i = 0; i < N; ++i) {
    const F4 x = ld4(p.px, i, Bp, b4), y = ld4(p.py, i, Bp, b4), vx = ld4(p.vx, i, Bp, b4), vy = ld4(p.vy, i, Bp, b4);
    const F4 dtg = ld4(p.dtg, i, Bp, b4);
    F4 gx, gy, fp;
    for (int c = 0; c < 4; ++c) {
      const int g = p.gm[(size_t)i * Bp + b4 + c];
      gx.v[c] = p.lx[(size_t)g * Bp + b4 + c]; gy.v[c] = p.ly[(size_t)g * Bp + b4 + c];
      fp.v[c] = (dtg.v[c] == -1.0f) ? ratio_eps(mean_p[c], std_p[c]) : ratio_eps((double)dmean.v[c], (double)dstd.v[c]);
    }
    if (obs) {
      const size_t o = (size_t)i * OBS_F;
      st4(obs, o + 0, Bp, b4, vx); st4(obs, o + 1, Bp, b4, vy); st4(obs, o + 2, Bp, b4, x); st4(obs, o + 3, Bp, b4, y);
      st4(obs, o + 4, Bp, b4, sub4(gx, x)); st4(obs, o + 5, Bp, b4, sub4(gy, y)); st4(obs, o + 6, Bp, b4, fp);
    }
    if (!node) continue;
    for (int e = 0; e < E; ++e) {
      F4 ex, ey, evx = splat4(0.0f), evy = splat4(0.0f), tx, ty;
      entity_pos(p, e, b4, ex, ey);
      tx = ex; ty = ey;                                                     // landmarks / obstacles: the goal is the entity itself
      float type = e < N ? 0.0f : (e < 2 * N ? 1.0f : 2.0f);
      if (e < N) {
        evx = ld4(p.vx, e, Bp, b4); evy = ld4(p.vy, e, Bp, b4);
        for (int c = 0; c < 4; ++c) {
          const int g = p.gm[(size_t)e * Bp + b4 + c];
          tx.v[c] = p.lx[(size_t)g * Bp + b4 + c]; ty.v[c] = p.ly[(size_t)g * Bp + b4 + c];
        }
      }
      const size_t r = ((size_t)i * E + e) * F;
      if (p.feat_global) {                                                  // [vel, pos, goal, type] (:1058-1077)
        st4(node, r + 0, Bp, b4, evx); st4(node, r + 1, Bp, b4, evy); st4(node, r + 2, Bp, b4, ex); st4(node, r + 3, Bp, b4, ey);
        st4(node, r + 4, Bp, b4, tx); st4(node, r + 5, Bp, b4, ty); st4(node, r + 6, Bp, b4, splat4(type));
      } else {                                                              // [rel vel, rel pos, rel goal, rel pos, rel pos, type] (:1079-1124)
        const F4 rx = sub4(ex, x), ry = sub4(ey, y);
        st4(node, r + 0, Bp, b4, sub4(evx, vx)); st4(node, r + 1, Bp, b4, sub4(evy, vy));
        st4(node, r + 2, Bp, b4, rx); st4(node, r + 3, Bp, b4, ry);
        st4(node, r + 4, Bp, b4, sub4(tx, x)); st4(node, r + 5, Bp, b4, sub4(ty, y));
        st4(node, r + 6, Bp, b4, rx); st4(node, r + 7, Bp, b4, ry); st4(node, r + 8, Bp, b4, rx); st4(node, r + 9, Bp, b4, ry);
        st4(node, r + 10, Bp, b4, splat4(type));
      }
    }
  }
  if (adj) {
    for (int a = 0; a < E; ++a) {
      F4 ax, ay;
      entity_pos(p, a, b4, ax, ay);
      st4(adj, (size_t)a * E + a, Bp, b4, splat4(0.0f));
      for (int c2 = a + 1; c2 < E; ++c2) {
        F4 bx, by, d;
        entity_pos(p, c2, b4, bx, by);
        for (int c = 0; c < 4; ++c) d.v[c] = (float)dist64(ax.v[c], ay.v[c], bx.v[c], by.v[c]);
        st4(adj, (size_t)a * E + c2, Bp, b4, d); st4(adj, (size_t)c2 * E + a, Bp, b4, d);
      }
    }
  }
}

__global__ void __launch_bounds__(256) finite_guard_kernel(const DevParams p, int* __restrict__ flags, int* __restrict__ count) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  bool bad = false;
  if (b < p.B) {
    for (int i = 0; i < p.N; ++i) {
      const size_t k = (size_t)i * p.Bp + b;
      bad = bad || !isfinite(p.px[k]) || !isfinite(p.py[k]) || !isfinite(p.vx[k]) || !isfinite(p.vy[k]) || !isfinite(p.pdist[k]);
    }
    bad = bad || !isfinite(p.dmean[b]) || !isfinite(p.dstd[b]);
    if (flags) flags[b] = bad ? 1 : 0;
  }
  const unsigned m = __ballot_sync(FULL, bad);
  if (count && (threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

}  // namespace

cudaError_t launch_observe_soa(const DevParams& p, float* obs, float* node, float* adj, cudaStream_t st) {
  const int threads = (p.Bp + 3) / 4;
  observe_soa_kernel<<<(threads + 127) / 128, 128, 0, st>>>(p, obs, node, adj);
  return cudaGetLastError();
}

cudaError_t launch_finite_guard(const DevParams& p, int* flags, int* count, cudaStream_t st) {
  if (count) { cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), st); if (e != cudaSuccess) return e; }
  finite_guard_kernel<<<(p.B + 255) / 256, 256, 0, st>>>(p, flags, count);
  return cudaGetLastError();
}

}  // namespace fm
