// Thread-per-env kernels for small teams (N <= 4), specialised at compile time on (N, O).
//
// Why a second mapping: at N = 3 one env is ~40 state words in and 479 words out, and the whole
// step is a few thousand scalar operations.  With one thread per env every loop over agents,
// entities and pairs unrolls into registers: no shuffles, no index arithmetic, no idle lanes.  The
// group-per-env kernel (fm_kernels.cu) needs ~530 warp instructions per env at N = 3 and is issue
// bound at 34 % of the HBM roofline; this one needs ~4x fewer.
//
//   lane  <-> env            32 consecutive envs per warp: every SoA state load / store is one
//                            fully coalesced 128-byte line per field.
//   outputs                  API layout is array-of-structs per env (297 + 81 + 21 words at N = 3), so
//                            a lane-per-env store would touch 32 different lines per instruction.
//                            Instead the warp stages through shared memory private to it:
//                            every lane drops its entity table (positions, velocities, goals) into
//                            smem; then ALL lanes build node_obs rows for a tile of TILE envs
//                            (row stride 11 words: conflict free) and the tile is streamed out as
//                            16-byte st.global.cs; adj and obs rows go register -> smem -> float4 too.
//   sync                     __syncwarp only; CTAs are 2 warps so that 65 536 envs give 1 024 CTAs
//                            (6.9 per SM: balanced over 148 SMs).
//
// Arithmetic is the same, operation for operation, as step_kernel<G> (same fp32 force terms, fp64
// force sums / integration / distances / statistics), so both mappings produce identical bits;
// tests/test_gpu_parity.py runs every case through both.
#include <utility>

#include "fm_device.cuh"
#include "fm_launch.h"
#include "fm_small.cuh"

namespace fm {

constexpr int TPE_THREADS = 64;
constexpr int TPE_TILE = 8;          // envs per node_obs staging tile

template <int N, int O>
struct TpeLayout {
  static constexpr int E = 2 * N + O;
  static constexpr int PAIRS = E * (E - 1) / 2;
  // entity table per env: P[E] V[N] G[N] as float2; an odd number of 8-byte words per env makes
  // the lane-per-env 64-bit accesses conflict free.
  static constexpr int TAB2 = E + 2 * N;
  static constexpr int TABW = 2 * (TAB2 | 1);
  static constexpr int ROWS = N * E;
  static constexpr int NODE_W = ROWS * NODE_F;
  static constexpr int ADJ_W = E * E;
  static constexpr int OBS_W = N * OBS_F;
  static constexpr int ADJ_ENVS = (32 * ADJ_W <= TPE_TILE * NODE_W) ? 32 : ((16 * ADJ_W <= TPE_TILE * NODE_W) ? 16 : 8);
  static constexpr int STAGE_RAW = (TPE_TILE * NODE_W > ADJ_ENVS * ADJ_W)
                                       ? (TPE_TILE * NODE_W > 32 * OBS_W ? TPE_TILE * NODE_W : 32 * OBS_W)
                                       : (ADJ_ENVS * ADJ_W > 32 * OBS_W ? ADJ_ENVS * ADJ_W : 32 * OBS_W);
  static constexpr int STAGE = (STAGE_RAW + 3) & ~3;
  static constexpr int TAB = (32 * TABW + 3) & ~3;
  static constexpr int PER_WARP = TAB + STAGE;     // floats
};

__host__ __device__ constexpr int pair_index(int a, int b, int E) { return a * E - a * (a + 1) / 2 + (b - a - 1); }   // a < b

// Per-thread env registers.
template <int N, int O>
struct EnvRegs {
  float px[N], py[N], vx[N], vy[N], lx[N], ly[N];
  float ox[O > 0 ? O : 1], oy[O > 0 ? O : 1];
  int gm[N];
};

template <int N, int O>
__device__ __forceinline__ float ent_x(const EnvRegs<N, O>& r, int e) {
  return e < N ? r.px[e] : (e < 2 * N ? r.lx[e - N] : r.ox[e - 2 * N]);
}
template <int N, int O>
__device__ __forceinline__ float ent_y(const EnvRegs<N, O>& r, int e) {
  return e < N ? r.py[e] : (e < 2 * N ? r.ly[e - N] : r.oy[e - 2 * N]);
}

// Distances that involve an agent (core.py:204-228 rows / columns 0..N-1), plus what reward() and
// info_callback() read off them (navigation_graph.py:650-661, :701-705, :773-782).
template <int N, int O>
__device__ __forceinline__ void agent_distances(const DevParams& p, const EnvRegs<N, O>& r,
                                                float (&adjv)[TpeLayout<N, O>::PAIRS], double (&dgoal)[N],
                                                int (&ncoll)[N], bool (&ocoll)[N]) {
  constexpr int E = 2 * N + O;
#pragma unroll
  for (int a = 0; a < N; ++a) { ncoll[a] = 0; ocoll[a] = false; dgoal[a] = 0.0; }
#pragma unroll
  for (int a = 0; a < N; ++a) {
#pragma unroll
    for (int b = a + 1; b < E; ++b) {
      const double d = dist64(ent_x(r, a), ent_y(r, a), ent_x(r, b), ent_y(r, b));
      adjv[pair_index(a, b, E)] = (float)d;
      if (b < N) {
        const int c = (d < p.dcoll) ? 1 : 0;
        ncoll[a] += c; ncoll[b] += c;
      } else if (b < 2 * N) {
        if (r.gm[a] == b - N) dgoal[a] = d;
      } else {
        ocoll[a] = ocoll[a] || (d < p.dcoll);
      }
    }
  }
}

// Landmark / obstacle block of the distance matrix (positions fixed within an episode).
template <int N, int O>
__device__ __forceinline__ void static_distances(const EnvRegs<N, O>& r, float (&adjv)[TpeLayout<N, O>::PAIRS]) {
  constexpr int E = 2 * N + O;
#pragma unroll
  for (int a = N; a < E; ++a)
#pragma unroll
    for (int b = a + 1; b < E; ++b)
      adjv[pair_index(a, b, E)] = (float)dist64(ent_x(r, a), ent_y(r, a), ent_x(r, b), ent_y(r, b));
}

// Randomised reset of this thread's env (navigation_graph.py:212-262, :264-570) + lexifair
// (:555-561).  `tb` is the thread's row of the shared entity table (P[e] at tb[2e], tb[2e+1]); it
// serves as dynamically indexable storage while placing.  Same Philox stream, draw order and
// acceptance rules as reset_group<G> (fm_device.cuh).
template <int N, int O>
__device__ __forceinline__ void tpe_reset(const DevParams& p, long long genv, uint32_t episode, float* __restrict__ tb,
                                          EnvRegs<N, O>& r, float (&mint)[N]) {
#pragma unroll 1
  for (int k = 0; k < O; ++k) {            // obstacles: 0.8 * U(-ws/2, ws/2)^2, draws 0..O-1 (:271-275)
    float x, y;
    draw_uniform2(p, genv, episode, (uint32_t)k, x, y);
    tb[2 * (2 * N + k)] = __fmul_rn(0.8f, x);
    tb[2 * (2 * N + k) + 1] = __fmul_rn(0.8f, y);
  }
  uint32_t d = (uint32_t)O;
#pragma unroll 1
  for (int slot = 0; slot < 2 * N; ++slot) {   // agents (:389-456) then goals (:472-535)
    const bool goal = slot >= N;
    const int base = goal ? N : 0;
    float x, y;
    while (true) {
      draw_uniform2(p, genv, episode, d, x, y);
      ++d;
      if (goal) { x = __fmul_rn(0.8f, x); y = __fmul_rn(0.8f, y); }
      bool bad = false;
#pragma unroll 1
      for (int k = 0; k < O; ++k) bad = bad || (dist64(tb[2 * (2 * N + k)], tb[2 * (2 * N + k) + 1], x, y) < p.dcoll);
#pragma unroll 1
      for (int j = base; j < slot; ++j) bad = bad || (dist64(tb[2 * j], tb[2 * j + 1], x, y) < p.dcoll);
      if (!bad || d >= (uint32_t)MAX_DRAWS) break;
    }
    tb[2 * slot] = x; tb[2 * slot + 1] = y;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    r.px[i] = tb[2 * i]; r.py[i] = tb[2 * i + 1];
    r.lx[i] = tb[2 * (N + i)]; r.ly[i] = tb[2 * (N + i) + 1];
    r.vx[i] = 0.f; r.vy[i] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < O; ++k) { r.ox[k] = tb[2 * (2 * N + k)]; r.oy[k] = tb[2 * (2 * N + k) + 1]; }
  double cost[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (p.has_max_speed) {                 // min_time with the PREVIOUS goal_match (:545-547, :719-728)
      const int og = N + r.gm[i];
      mint[i] = (float)(dist64(r.px[i], r.py[i], tb[2 * og], tb[2 * og + 1]) / p.max_speed);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) cost[i * N + j] = dist64(r.px[i], r.py[i], r.lx[j], r.ly[j]);   // cdist (:555)
  }
  lexifair_small<N>(cost, r.gm);
}

// Stream `nwords` floats of warp-private shared memory to global memory.
__device__ __forceinline__ void tpe_copy_out(float* __restrict__ dst, const float* __restrict__ src, int nwords, int lane) {
  warp_copy_out(dst, src, nwords, lane);
}

// Distance matrix completion + obs / node_obs / adj emission for the warp's 32 envs.
// navigation_graph.py:826-857 (observation), :941-1035 + :1079-1124 (graph_observation, relative
// features), core.py:204-228 (cached_dist_mag == adj).
template <int N, int O>
__device__ __forceinline__ void tpe_emit(const DevParams& p, const EnvRegs<N, O>& r, float (&adjv)[TpeLayout<N, O>::PAIRS],
                                         const float (&fobs)[N], float* __restrict__ tab, float* __restrict__ stage,
                                         int env0, int nenv, int lane) {
  using L = TpeLayout<N, O>;
  constexpr int E = L::E;
  float* tb = tab + lane * L::TABW;
  float2* tb2 = reinterpret_cast<float2*>(tb);
  // entity table row of this env: P[E], V[N], G[N].  The goal of agent i (landmark goal_match[i]) is
  // read back from the row just written: a dynamic index into shared memory instead of registers.
#pragma unroll
  for (int e = 0; e < E; ++e) tb2[e] = make_float2(ent_x(r, e), ent_y(r, e));
  float gx[N], gy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float2 g = tb2[N + r.gm[i]];
    gx[i] = g.x; gy[i] = g.y;
    tb2[E + i] = make_float2(r.vx[i], r.vy[i]);
    tb2[E + N + i] = g;
  }
  __syncwarp();

  // ---- obs [B, N, 7]
  if (p.o_obs) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float* o = stage + lane * L::OBS_W + i * OBS_F;
      o[0] = r.vx[i]; o[1] = r.vy[i]; o[2] = r.px[i]; o[3] = r.py[i];
      o[4] = gx[i] - r.px[i]; o[5] = gy[i] - r.py[i]; o[6] = fobs[i];
    }
    __syncwarp();
    tpe_copy_out(p.o_obs + (size_t)env0 * L::OBS_W, stage, nenv * L::OBS_W, lane);
    __syncwarp();
  }

  // ---- adj [B, E, E]: the thread's E x E matrix, staged ADJ_ENVS envs at a time
  if (p.o_adj) {
    static_distances<N, O>(r, adjv);
#pragma unroll 1
    for (int h = 0; h < 32; h += L::ADJ_ENVS) {
      if (h >= nenv) break;
      if (lane >= h && lane < h + L::ADJ_ENVS) {
        float* a = stage + (lane - h) * L::ADJ_W;
#pragma unroll
        for (int x = 0; x < E; ++x)
#pragma unroll
          for (int y = 0; y < E; ++y)
            a[x * E + y] = (x == y) ? 0.0f : adjv[x < y ? pair_index(x, y, E) : pair_index(y, x, E)];
      }
      __syncwarp();
      tpe_copy_out(p.o_adj + (size_t)(env0 + h) * L::ADJ_W, stage, min(L::ADJ_ENVS, nenv - h) * L::ADJ_W, lane);
      __syncwarp();
    }
  }

  // ---- node_obs [B, N, E, 11]: all lanes build rows of a TILE-env tile from the entity table
  if (p.o_node) {
#pragma unroll 1
    for (int t0 = 0; t0 < 32; t0 += TPE_TILE) {
      if (t0 >= nenv) break;
      const int tn = min(TPE_TILE, nenv - t0);
      const int rows = tn * L::ROWS;
#pragma unroll
      for (int r0 = 0; r0 < TPE_TILE * L::ROWS; r0 += 32) {
        const int row = r0 + lane;
        if (row < rows) {
          const int el = row / L::ROWS;
          const int rem = row - el * L::ROWS;
          const int a = rem / E;
          const int e = rem - a * E;
          const float2* t2 = reinterpret_cast<const float2*>(tab + (t0 + el) * L::TABW);
          const float2 pa = t2[a], va = t2[E + a], pe = t2[e];
          float2 ve = make_float2(0.f, 0.f), ge = pe;
          if (e < N) { ve = t2[E + e]; ge = t2[E + N + e]; }
          const float rpx = pe.x - pa.x, rpy = pe.y - pa.y;
          float* st = stage + row * NODE_F;
          st[0] = ve.x - va.x; st[1] = ve.y - va.y;
          st[2] = rpx; st[3] = rpy;
          st[4] = ge.x - pa.x; st[5] = ge.y - pa.y;
          st[6] = rpx; st[7] = rpy; st[8] = rpx; st[9] = rpy;
          st[10] = (e < N) ? 0.0f : ((e < 2 * N) ? 1.0f : 2.0f);
        }
      }
      __syncwarp();
      tpe_copy_out(p.o_node + (size_t)(env0 + t0) * L::NODE_W, stage, rows * NODE_F, lane);
      __syncwarp();
    }
  }
}

// =============================================================================================
// MODE 0: fused env step (MultiAgentGraphEnv.step, environment.py:816-877, + graphworker auto-reset,
//         env_wrappers.py:859-865).   MODE 1: masked reset + observe (environment.py:882-898).
template <int N, int O, int MODE>
__global__ void __launch_bounds__(TPE_THREADS) tpe_kernel(const __grid_constant__ DevParams p) {
  using L = TpeLayout<N, O>;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * (TPE_THREADS / 32) + wib;
  const int env0 = gw * 32;
  if (env0 >= p.B) return;                       // warp-uniform
  const int nenv = min(32, p.B - env0);
  const int env = env0 + lane;                   // < Bp: state arrays are padded, loads are always legal
  const bool venv = lane < nenv;
  float* tab = smem + (size_t)wib * L::PER_WARP;
  float* stage = tab + L::TAB;
  float* tb = tab + lane * L::TABW;
  const size_t Bp = (size_t)p.Bp;

  // ---- load state (SoA: one coalesced line per field per warp) ----------------------------------
  EnvRegs<N, O> r;
  float pd[N], dtg[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const size_t idx = i * Bp + env;
    r.px[i] = p.px[idx]; r.py[i] = p.py[idx]; r.vx[i] = p.vx[idx]; r.vy[i] = p.vy[idx];
    r.lx[i] = p.lx[idx]; r.ly[i] = p.ly[idx]; r.gm[i] = p.gm[idx];
    pd[i] = p.pdist[idx]; dtg[i] = p.dtg[idx];
  }
#pragma unroll
  for (int k = 0; k < O; ++k) { r.ox[k] = p.ox[k * Bp + env]; r.oy[k] = p.oy[k * Bp + env]; }
  const uint32_t episode = (uint32_t)p.episode[env];
  const float dmean = p.dmean[env], dstd = p.dstd[env];
  const long long genv = p.env_offset + env;

  float adjv[L::PAIRS];
  float fobs[N];
  double dgoal[N]; int ncoll[N]; bool ocoll[N];

  if (MODE == 1) {
    // ---- reset() / observe --------------------------------------------------------------------
    const bool do_reset = venv && (p.reset_mask ? (p.reset_mask[env] != 0) : true);
    if (do_reset) {
      float mint[N];
#pragma unroll
      for (int i = 0; i < N; ++i) mint[i] = p.mintime[i * Bp + env];
      tpe_reset<N, O>(p, genv, episode, tb, r, mint);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const size_t idx = i * Bp + env;
        pd[i] = 0.f; dtg[i] = -1.f;
        p.px[idx] = r.px[i]; p.py[idx] = r.py[i]; p.vx[idx] = 0.f; p.vy[idx] = 0.f; p.pdist[idx] = 0.f;
        p.dtg[idx] = -1.f; p.treq[idx] = -1.f; p.dleft[idx] = -1.f;
        p.gm[idx] = r.gm[i]; p.nac[idx] = 0; p.noc[idx] = 0; p.mintime[idx] = mint[i];
        p.lx[idx] = r.lx[i]; p.ly[idx] = r.ly[i];
      }
#pragma unroll
      for (int k = 0; k < O; ++k) { p.ox[k * Bp + env] = r.ox[k]; p.oy[k * Bp + env] = r.oy[k]; }
      p.step[env] = 0; p.episode[env] = (int)(episode + 1);
    }
    __syncwarp();
    // observation() on the current state (navigation_graph.py:826-857, :849-853)
    double sum_p = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) sum_p += (double)pd[j];
    const double mean_p = sum_p / N;
    double q_p = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) { const double dd = (double)pd[j] - mean_p; q_p += dd * dd; }
    const double std_p = sqrt(q_p / N);
#pragma unroll
    for (int i = 0; i < N; ++i)
      fobs[i] = (float)((dtg[i] == -1.0f) ? mean_p / (std_p + 0.0001) : (double)dmean / ((double)dstd + 0.0001));
    agent_distances<N, O>(p, r, adjv, dgoal, ncoll, ocoll);
    tpe_emit<N, O>(p, r, adjv, fobs, tab, stage, env0, nenv, lane);
    return;
  }

  // ---- step -----------------------------------------------------------------------------------
  float treq[N], dleft[N];
  int nac[N], noc[N];
  float ux[N], uy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const size_t idx = i * Bp + env;
    treq[i] = p.treq[idx]; dleft[i] = p.dleft[idx]; nac[i] = p.nac[idx]; noc[i] = p.noc[idx];
  }
  const int step = p.step[env];
  if (venv) {
    // action decode, environment.py:301-311: u = [a1 - a2, a3 - a4] * sensitivity (5.0)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (p.act_idx) {
        const int a = p.act_idx[(size_t)env * N + i];
        ux[i] = ((a == 1) ? 1.f : 0.f) - ((a == 2) ? 1.f : 0.f);
        uy[i] = ((a == 3) ? 1.f : 0.f) - ((a == 4) ? 1.f : 0.f);
      } else {
        const float* oh = p.act_onehot + ((size_t)env * N + i) * 5;
        ux[i] = oh[1] - oh[2];
        uy[i] = oh[3] - oh[4];
      }
      ux[i] *= 5.0f; uy[i] *= 5.0f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) { ux[i] = 0.f; uy[i] = 0.f; }
  }

  // World.step: forces (core.py:277-316, :370-404) from the positions at step entry; every agent
  // accumulates its partners in ascending entity index (pair (i, j) is visited with i ascending,
  // so agent j has received all i < j before its own row starts).
  double Fx[N], Fy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { Fx[i] = (double)ux[i]; Fy[i] = (double)uy[i]; }   // mass(1.0) * u + noise(0.0)
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      const float dx = r.px[i] - r.px[j], dy = r.py[i] - r.py[j];
      const float dist = sqrtf(dx * dx + dy * dy);
      const float pen = softplusf(-(dist - p.dist_min) / p.contact_margin) * p.contact_margin;
      const float tx = p.contact_force * dx / dist * pen, ty = p.contact_force * dy / dist * pen;
      Fx[i] = (double)tx + Fx[i]; Fy[i] = (double)ty + Fy[i];
      Fx[j] = (double)(-tx) + Fx[j]; Fy[j] = (double)(-ty) + Fy[j];
    }
#pragma unroll
    for (int k = 0; k < O; ++k) contact_force(p, r.px[i], r.py[i], r.ox[k], r.oy[k], Fx[i], Fy[i]);
  }
  // integrate_state (core.py:338-356) in float64; state is stored rounded to fp32.
  double pd64[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double v64x = (double)r.vx[i] * p.damping_keep + Fx[i] * p.dt;
    double v64y = (double)r.vy[i] * p.damping_keep + Fy[i] * p.dt;
    if (p.has_max_speed) {
      const double speed = sqrt(v64x * v64x + v64y * v64y);
      if (speed > p.max_speed) { v64x = v64x / speed * p.max_speed; v64y = v64y / speed * p.max_speed; }
    }
    const double sx = v64x * p.dt, sy = v64y * p.dt;
    pd64[i] = (double)pd[i] + sqrt(sx * sx + sy * sy);
    r.px[i] = (float)((double)r.px[i] + sx); r.py[i] = (float)((double)r.py[i] + sy);
    r.vx[i] = (float)v64x; r.vy[i] = (float)v64y;
  }
  const int nstep = step + 1;                    // environment.py:819, :823

  // calculate_distances (core.py:204-228) at the new positions: rows of the agents
  agent_distances<N, O>(p, r, adjv, dgoal, ncoll, ocoll);

  // per-agent loop of MultiAgentGraphEnv.step (environment.py:832-864): agent i's observation and
  // reward read world.dist_traveled_mean/stddev as left by agent i-1's info_callback
  // (navigation_graph.py:617-618); agent 0 reads last step's values.
  bool latched[N], reached[N];
  double dtg_prev[N], dtg_new[N], treq_prev[N], treq_new[N];
  float dleft_new[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    latched[i] = treq[i] != -1.0f;
    dtg_prev[i] = (double)dtg[i];
    dtg_new[i] = latched[i] ? dtg_prev[i] : pd64[i];
    reached[i] = dgoal[i] < p.min_dist_thresh;
    treq_prev[i] = (double)treq[i];
    treq_new[i] = (!latched[i] && reached[i]) ? (double)nstep * p.dt : treq_prev[i];   // :588
    dleft_new[i] = latched[i] ? dleft[i] : (float)dgoal[i];
  }
  double sum_p = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) sum_p += pd64[j];
  const double mean_p = sum_p / N;
  double q_p = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) { const double dd = pd64[j] - mean_p; q_p += dd * dd; }
  const double std_p = sqrt(q_p / N);
  // V_k = mean / std over [new_0 .. new_{k-1}, prev_k .. prev_{N-1}],  k = 1..N
  double vmean[N + 1], vstd[N + 1];
#pragma unroll
  for (int k = 1; k <= N; ++k) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s += (j < k) ? dtg_new[j] : dtg_prev[j];
    const double m = s / N;
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) { const double dd = ((j < k) ? dtg_new[j] : dtg_prev[j]) - m; q += dd * dd; }
    vmean[k] = m; vstd[k] = sqrt(q / N);
  }
  float rew[N], own_rew[N];
  double fparam[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (dtg[i] == -1.0f) fparam[i] = mean_p / (std_p + 0.0001);          // navigation_graph.py:764-769 / :849-853
    else if (i == 0) fparam[i] = (double)dmean / ((double)dstd + 0.0001);
    else fparam[i] = vmean[i] / (vstd[i] + 0.0001);
    // reward (navigation_graph.py:760-824)
    float rw = reached[i] ? p.goal_rew : -(float)dgoal[i];
    rw -= p.coll_rew * (float)ncoll[i];
    if (ocoll[i]) rw -= p.coll_rew;
    if (p.fairness_reward) {
      float fair = p.fair_rew * tanhf((float)(fparam[i] - p.zeroshift));
      if (fair < -2.0f) fair = -2.0f;
      rw += fair;
    }
    rw = fminf(fmaxf(rw, p.clip_lo), p.clip_hi);
    own_rew[i] = rw;
    nac[i] += ncoll[i];                          // :604-613
    noc[i] += ocoll[i] ? 1 : 0;                  // :602-603
  }
  if (p.collaborative) {                         // environment.py:866-870
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) tot += own_rew[j];
#pragma unroll
    for (int i = 0; i < N; ++i) rew[i] = tot;
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) rew[i] = own_rew[i];
  }
  const bool done = nstep >= p.episode_length;   // environment.py:237-247 (agent.status is never set)
  const bool do_reset = venv && done && (p.auto_reset != 0);

  // ---- info rows (navigation_graph.py:625-647) and episode statistics -------------------------
  const bool want_info = venv && (p.o_info != nullptr || p.stats != nullptr) && (done || p.info_every_step);
  const bool any_info = __any_sync(FULL, want_info);
  float mint[N];
#pragma unroll
  for (int i = 0; i < N; ++i) mint[i] = 0.f;
  if (any_info || __any_sync(FULL, do_reset)) {
#pragma unroll
    for (int i = 0; i < N; ++i) mint[i] = p.mintime[i * Bp + env];
  }
  if (p.stats) {
    double* row = p.stats + (size_t)gw * (15 * N + 2);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double v = venv ? (double)rew[i] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
      if (lane == 0) row[i] += v;
    }
    const unsigned termb = __ballot_sync(FULL, venv && done);
    if (lane == 0) { row[15 * N] += (double)__popc(termb); row[15 * N + 1] += (double)nenv; }
  }
  if (any_info) {
    double tacc = 0.0;                           // entity.state.time += dt per step (core.py:355)
    for (int k = 0; k < nstep; ++k) tacc += p.dt;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      // world-level means right after agent i's own info_callback: over [new_0..new_i, prev_i+1..]
      double st = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) st += (j <= i) ? treq_new[j] : treq_prev[j];
      const double mt = st / N;
      double qt = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) { const double dd = ((j <= i) ? treq_new[j] : treq_prev[j]) - mt; qt += dd * dd; }
      const double stv = sqrt(qt / N);
      const double md = vmean[i + 1], sdv = vstd[i + 1];
      float info[INFO_F];
      info[0] = own_rew[i]; info[1] = dleft_new[i]; info[2] = (float)treq_new[i]; info[3] = (float)nac[i];
      info[4] = (float)noc[i]; info[5] = (float)md; info[6] = (float)sdv; info[7] = (float)(md / (sdv + 0.0001));
      info[8] = (float)dtg_new[i]; info[9] = (float)tacc; info[10] = (float)mt; info[11] = (float)stv;
      info[12] = (float)(mt / (stv + 0.0001)); info[13] = mint[i];
      if (want_info && p.o_info) {
        float* o = p.o_info + ((size_t)env * N + i) * INFO_F;
#pragma unroll
        for (int k = 0; k < INFO_F; ++k) o[k] = info[k];
      }
      if (p.stats && __any_sync(FULL, venv && done)) {
        double* row = p.stats + (size_t)gw * (15 * N + 2);
#pragma unroll
        for (int k = 0; k < INFO_F; ++k) {
          double v = (venv && done) ? (double)info[k] : 0.0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
          if (lane == 0) row[N + i * INFO_F + k] += v;
        }
      }
    }
  }

  // ---- new state; graphworker auto-reset (env_wrappers.py:859-865): obs / node_obs / adj come from
  // the new episode, reward / done / info stay terminal -----------------------------------------
  float npd[N], ndtg[N], ntreq[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    fobs[i] = (float)fparam[i];
    npd[i] = (float)pd64[i]; ndtg[i] = (float)dtg_new[i]; ntreq[i] = (float)treq_new[i];
  }
  const float ndmean = (float)vmean[N], ndstd = (float)vstd[N];     // after the last agent's info_callback
  int nstep_store = nstep;
  uint32_t nepisode = episode;
  if (__any_sync(FULL, do_reset)) {
    if (do_reset) {
      tpe_reset<N, O>(p, genv, episode, tb, r, mint);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const size_t idx = i * Bp + env;
        npd[i] = 0.f; ndtg[i] = -1.f; ntreq[i] = -1.f; dleft_new[i] = -1.f; nac[i] = 0; noc[i] = 0;
        fobs[i] = 0.f;                           // mean(p_dist = 0) / (std + 1e-4)
        p.mintime[idx] = mint[i];
        p.lx[idx] = r.lx[i]; p.ly[idx] = r.ly[i];
      }
#pragma unroll
      for (int k = 0; k < O; ++k) { p.ox[k * Bp + env] = r.ox[k]; p.oy[k * Bp + env] = r.oy[k]; }
      nstep_store = 0; nepisode = episode + 1;
    }
    __syncwarp();
    agent_distances<N, O>(p, r, adjv, dgoal, ncoll, ocoll);
  }
  if (venv) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const size_t idx = i * Bp + env;
      p.px[idx] = r.px[i]; p.py[idx] = r.py[i]; p.vx[idx] = r.vx[i]; p.vy[idx] = r.vy[i]; p.pdist[idx] = npd[i];
      p.dtg[idx] = ndtg[i]; p.treq[idx] = ntreq[i]; p.dleft[idx] = dleft_new[i];
      p.gm[idx] = r.gm[i]; p.nac[idx] = nac[i]; p.noc[idx] = noc[i];
      if (p.o_rew) p.o_rew[(size_t)env * N + i] = rew[i];
      if (p.o_done) p.o_done[(size_t)env * N + i] = done ? 1 : 0;
    }
    p.step[env] = nstep_store; p.episode[env] = (int)nepisode; p.dmean[env] = ndmean; p.dstd[env] = ndstd;
  }
  tpe_emit<N, O>(p, r, adjv, fobs, tab, stage, env0, nenv, lane);
}

// =============================================================================================
template <int N, int O>
static cudaError_t tpe_launch_no(const DevParams& p, cudaStream_t st, bool is_reset) {
  using L = TpeLayout<N, O>;
  const int warps = (p.B + 31) / 32;
  const int blocks = (warps + TPE_THREADS / 32 - 1) / (TPE_THREADS / 32);
  const size_t smem = (size_t)L::PER_WARP * (TPE_THREADS / 32) * sizeof(float);
  if (is_reset) tpe_kernel<N, O, 1><<<blocks, TPE_THREADS, smem, st>>>(p);
  else tpe_kernel<N, O, 0><<<blocks, TPE_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

template <int N, int O>
static cudaError_t tpe_prepare_no() {
  using L = TpeLayout<N, O>;
  const int smem = L::PER_WARP * (TPE_THREADS / 32) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(tpe_kernel<N, O, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tpe_kernel<N, O, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// The (N, O) pairs compiled for this mapping.  Everything else runs the group-per-env kernels.
#define FM_TPE_CASES(X) X(1, 1) X(2, 0) X(3, 0) X(3, 3) X(4, 2)

bool tpe_supported(int N, int O) {
#define X(n, o) if (N == n && O == o) return true;
  FM_TPE_CASES(X)
#undef X
  return false;
}

int tpe_num_warps(int B) { return (B + 31) / 32; }

cudaError_t tpe_prepare(const DevParams& p) {
#define X(n, o) if (p.N == n && p.O == o) return tpe_prepare_no<n, o>();
  FM_TPE_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t tpe_launch(const DevParams& p, cudaStream_t st, bool is_reset) {
#define X(n, o) if (p.N == n && p.O == o) return tpe_launch_no<n, o>(p, st, is_reset);
  FM_TPE_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace fm
