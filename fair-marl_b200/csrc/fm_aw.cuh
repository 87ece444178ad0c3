// Agent-warp mapping, shared between the one-shot kernels (fm_aw.cu) and the persistent rollout kernel (fm_roll.cu):
// shared-memory layout, per-env reset, and the body that processes ONE tile of 32 consecutive envs for ONE step.
//
// One CTA owns a tile of 32 consecutive envs; lane <-> env, WARP <-> AGENT.  Thread (env, agent i) keeps agent i's
// whole state in registers from the first global load to the write-back: forces on agent i, integration, agent i's
// row of the distance matrix, its observation scalar, reward, goal latches and counters never leave the thread.  One
// extra "env warp" owns what belongs to the env rather than to an agent: static entity positions, the landmark /
// obstacle block of the distance matrix, auto-reset (placement + lexifair), episode counters.
// Specialised at compile time on (N, O); N <= 4 (lexifair by enumeration, whole-tile output staging).
//
// Shared memory per CTA (41.8 KB at N = 3, O = 3 -> 5 CTAs / SM):
//   tables   TP / TV / TG  positions, velocities, goals of the tile after the step ([row][32], lane = env)
//   staging  one region used twice: (1) adj | obs | reward | done of the 32 envs, with the scratch rows
//            the sequential-agent statistics exchange placed behind them; (2) node_obs of the 32 envs.
// Every output is written lane = env into a shared-memory IMAGE of the API layout (odd strides: conflict free); a
// full tile's slice of every output array is one contiguous, 16-byte aligned range that one thread hands to the
// copy engine (TMA bulk store); ragged / unaligned tiles use 16-byte st.global.cs by the whole CTA.
//
// Arithmetic is operation for operation that of step_kernel<G> / reset_kernel<G> (fm_kernels.cu, shared helpers in
// fm_device.cuh); tests/test_gpu_parity.py checks the mappings against each other bit for bit.
#pragma once
#include <utility>

#include "fm_device.cuh"
#include "fm_launch.h"
#include "fm_small.cuh"

namespace fm {

template <int N, int O, int W = 0>                 // W walls (0..2): entities 2N+O .. E-1, static like landmarks and obstacles
struct AwLayout {
  static constexpr int E = 2 * N + O + W, M = N + O + W, SP = M * (M - 1) / 2;
  static constexpr int WARPS = N + 1, THREADS = 32 * WARPS, ENVS = 32, RW = 32;
  // ---- global state rows ([row][Bp], fm_abi.cu fm_create order)
  static constexpr int PX = 0, PY = PX + N, VX = PY + N, VY = VX + N, PD = VY + N, DTG = PD + N, TREQ = DTG + N,
                       DLEFT = TREQ + N, MINT = DLEFT + N, GM = MINT + N, NAC = GM + N, NOC = NAC + N, LX = NOC + N,
                       LY = LX + N, OX = LY + N, OY = OX + O, DMEAN = OY + O, DSTD = DMEAN + 1, STEP = DSTD + 1,
                       EPIS = STEP + 1, SDIST = EPIS + 1;
  // ---- tables that live until the node_obs emission, rows of RW floats
  static constexpr int TP = 0,                 // [E][2] positions after the step (after the reset for envs that reset)
                       TV = TP + 2 * E,        // [N][2] velocities
                       TG = TV + 2 * N,        // [N][2] goal (assigned landmark) of agent i
                       TWA = TG + 2 * N,       // [W] wall axis, [W] orientation (0 = 'H'), [1] half-length (W > 0 only)
                       TWO = TWA + W, TWL = TWO + W,
                       T_ROWS = TWL + (W > 0 ? 1 : 0);
  static constexpr int OFF_STAGE = T_ROWS * RW;                       // multiple of 32 floats
  static constexpr int OBS_W = N * OBS_F, NODE_W = N * E * NODE_F, ADJ_W = E * E;
  // ---- staging, use 1: adj | obs | reward | done (bytes) | scratch
  static constexpr int S_ADJ = 0, S_OBS = S_ADJ + ENVS * ADJ_W, S_REW = S_OBS + ENVS * OBS_W, S_DONE = S_REW + ENVS * N,
                       SMALL_W = (S_DONE + (ENVS * N + 3) / 4 + 3) & ~3;
  // scratch rows (dead before the node_obs emission), RW floats each, relative to staging + SMALL_W
  static constexpr int DTGO = 0,               // [N] world.dists_to_goal at step entry
                       TREQO = DTGO + N,       // [N] world.times_required at step entry
                       NTREQ = TREQO + N,      // [N] ... after agent i's info_callback
                       OWN = NTREQ + N,        // [N] agent i's own reward
                       GMO = OWN + N,          // [N] goal_match at step entry (int bits)
                       RGM = GMO + N,          // [N] goal_match after a reset (int bits)
                       RMINT = RGM + N,        // [N] min_time after a reset
                       F_ROWS = (RMINT + N + 1) & ~1;
  static constexpr int PD64 = 0, SETM = PD64 + N, SETS = SETM + N + 1, D_ROWS = SETS + N + 1;   // rows of RW doubles
  static constexpr int OFF_SCR = SMALL_W, OFF_SCRD = OFF_SCR + F_ROWS * RW, SCR_END = OFF_SCRD + 2 * D_ROWS * RW;
  // ---- staging, use 2: node_obs
  static constexpr int STAGE_NODE = ENVS * NODE_W;
  static constexpr int STAGE_W = ((STAGE_NODE > SCR_END ? STAGE_NODE : SCR_END) + 3) & ~3;
  static constexpr int WORDS = OFF_STAGE + STAGE_W;
  static constexpr int FIT = (227 * 1024) / (WORDS * 4 + 1024);
  static constexpr int MIN_CTAS = FIT < 1 ? 1 : (FIT > 5 ? 5 : FIT);   // 5 x 128 threads -> up to 102 registers
};

// static pair (a, b), a < b < M, row-major  ->  SDIST row
__host__ __device__ constexpr int aw_spair(int a, int b, int M) { return a * M - a * (a + 1) / 2 + (b - a - 1); }

// float64 distance with one point already converted (same bits as dist64: the conversions are exact).
__device__ __forceinline__ double dist64_d(double ax, double ay, float bx, float by) {
  const double dx = __dsub_rn(ax, (double)bx);
  const double dy = __dsub_rn(ay, (double)by);
  return dsqrt_fast(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// CTA-wide copy of the staging image to global memory, 16-byte vectorised.  When the tile has all 32
// envs the trip count is a compile-time constant and the loads of 8 iterations are in flight before
// the first store; otherwise `nwords` is a runtime count.  The destination is 16-byte aligned whenever
// the output array is (tiles are 32 envs).
template <int THREADS, int WORDS_FULL>
__device__ __forceinline__ void cta_copy_out(float* __restrict__ dst, const float* __restrict__ src, int nwords, int tid) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    if (nwords == WORDS_FULL) {
      constexpr int N4 = WORDS_FULL >> 2;
      constexpr int FI = N4 / THREADS;               // iterations in which every thread moves 16 bytes
      constexpr int U = 6;
      const float4* sp = s4 + tid;
      float4* dp = d4 + tid;
#pragma unroll
      for (int k0 = 0; k0 < FI; k0 += U) {           // compile-time trip count: immediate offsets, no predicates
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (k0 + u < FI) v[u] = sp[(k0 + u) * THREADS];
#pragma unroll
        for (int u = 0; u < U; ++u) if (k0 + u < FI) __stcs(dp + (k0 + u) * THREADS, v[u]);
      }
      if (tid < N4 - FI * THREADS) __stcs(dp + FI * THREADS, sp[FI * THREADS]);
      for (int k = (N4 << 2) + tid; k < WORDS_FULL; k += THREADS) __stcs(dst + k, src[k]);
    } else {
      const int n4 = nwords >> 2;
      for (int k = tid; k < n4; k += THREADS) __stcs(d4 + k, s4[k]);
      for (int k = (n4 << 2) + tid; k < nwords; k += THREADS) __stcs(dst + k, src[k]);
    }
  } else {
    for (int k = tid; k < nwords; k += THREADS) __stcs(dst + k, src[k]);
  }
}

// Randomised reset of env `lane` by one thread of the env warp (navigation_graph.py:212-262,
// :264-570) + lexifair (:555-561).  Same Philox stream, draw order and acceptance rules as
// reset_group<G> (fm_device.cuh).  New positions go to the TP table, goal_match / min_time to the
// RGM / RMINT scratch rows; the distances between static entities are returned in sd[] (float).
template <int N, int O, int W = 0>
__device__ __noinline__ void aw_reset_env(const DevParams& p, long long genv, uint32_t episode, float* __restrict__ Tc,
                                          float* __restrict__ Sc, float* __restrict__ sd) {
  using L = AwLayout<N, O, W>;
  constexpr int RW = L::RW, M = L::M;
  auto PXY = [&](int e, int c) -> float& { return Tc[(L::TP + 2 * e + c) * RW]; };
  // The obstacles and the entities of the kind being placed are kept in registers as doubles (O <= 3, N <= 4: the loops
  // below unroll): the rejection tests are then independent float64 multiply-adds with no shared-memory load or conversion
  // in the chain, compared as SQUARES -- |c - q| < dcoll  <=>  |c - q|^2 < dcoll2_lt, the smallest double whose correctly
  // rounded root reaches dcoll (fm_create): the same decision without the root.  Terminal step 74 -> 59 us
  // (profiles/r02_sq_*).  Keeping the cost matrix and the static distances in registers as well made ptxas spill in the
  // CALLER (80 bytes in the step kernel's hot path: regular steps 27 -> 29 us) and was dropped.
  double ox[O > 0 ? O : 1], oy[O > 0 ? O : 1];
#pragma unroll
  for (int k = 0; k < O; ++k) {            // obstacles: 0.8 * U(-ws/2, ws/2)^2, draws 0..O-1 (:271-275)
    float x, y;
    draw_uniform2(p, genv, episode, (uint32_t)k, x, y);
    x = __fmul_rn(0.8f, x); y = __fmul_rn(0.8f, y);
    PXY(2 * N + k, 0) = x; PXY(2 * N + k, 1) = y;
    ox[k] = (double)x; oy[k] = (double)y;
  }
  auto too_close = [&](double qx, double qy, double cx, double cy) -> bool {   // dist64_sq(q, c) < dcoll2_lt
    const double dx = __dsub_rn(qx, cx), dy = __dsub_rn(qy, cy);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < p.dcoll2_lt;
  };
  // walls (navigation_graph.py:287-324): one draw for the axis offset U(0.2, 0.9) * ws / 2 (wall 0 at +, wall 1 at -), one
  // draw per wall for the orientation, draws O .. O + W (reset_group<G, true>, fm_device.cuh); the half-length is fixed per env
  float wax[W > 0 ? W : 1], wlen = 0.f;
  bool whz[W > 0 ? W : 1];
  if (W > 0) {
    float u0, u1;
    draw_u01(p, genv, episode, (uint32_t)O, u0, u1);
    const float wp = __fmul_rn(__fadd_rn(0.2f, __fmul_rn(0.7f, u0)), p.half_world);
    wlen = Tc[L::TWL * RW];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      draw_u01(p, genv, episode, (uint32_t)(O + 1 + k), u0, u1);
      whz[k] = !(u0 >= 0.5f);
      wax[k] = k == 0 ? wp : -wp;
      Tc[(L::TWA + k) * RW] = wax[k]; Tc[(L::TWO + k) * RW] = whz[k] ? 0.0f : 1.0f;
      PXY(2 * N + O + k, 0) = whz[k] ? 0.0f : wax[k]; PXY(2 * N + O + k, 1) = whz[k] ? wax[k] : 0.0f;   // midpoint
    }
  }
  uint32_t d = (uint32_t)(O + (W > 0 ? 1 + W : 0));
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {       // agents (:389-456) then goals (:472-535); entity index == pass * N + a
    double qx[N], qy[N];
#pragma unroll
    for (int a = 0; a < N; ++a) {
      float x, y;
      while (true) {
        draw_uniform2(p, genv, episode, d, x, y);
        ++d;
        if (pass) { x = __fmul_rn(0.8f, x); y = __fmul_rn(0.8f, y); }
        const double cx = (double)x, cy = (double)y;
        bool bad = false;
#pragma unroll
        for (int k = 0; k < O; ++k) bad |= too_close(ox[k], oy[k], cx, cy);
#pragma unroll
        for (int j = 0; j < a; ++j) bad |= too_close(qx[j], qy[j], cx, cy);
#pragma unroll
        for (int k = 0; k < W; ++k) bad |= in_wall_box(x, y, whz[k], wax[k], wlen);   // is_obstacle_collision's wall boxes (:670-683)
        if (!bad || d >= (uint32_t)MAX_DRAWS) break;
      }
      qx[a] = (double)x; qy[a] = (double)y;
      PXY(pass * N + a, 0) = x;
      PXY(pass * N + a, 1) = y;
    }
  }
  double cost[N * N];
  int gm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float ax = PXY(i, 0), ay = PXY(i, 1);
    if (p.has_max_speed) {                 // min_time with the PREVIOUS goal_match (:545-547, :719-728)
      const int og = __float_as_int(Sc[(L::GMO + i) * RW]);
      Sc[(L::RMINT + i) * RW] = (float)(dist64(ax, ay, PXY(N + og, 0), PXY(N + og, 1)) / p.max_speed);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) cost[i * N + j] = dist64(ax, ay, PXY(N + j, 0), PXY(N + j, 1));   // cdist (:555)
  }
  lexifair_small<N>(cost, gm);
#pragma unroll
  for (int i = 0; i < N; ++i) Sc[(L::RGM + i) * RW] = __int_as_float(gm[i]);
#pragma unroll 1
  for (int a = 0; a < M; ++a) {
    const double axx = (double)PXY(N + a, 0), ayy = (double)PXY(N + a, 1);
#pragma unroll
    for (int b = 0; b < M; ++b)                  // unrolled: the distances of a row are independent chains
      if (b > a) sd[aw_spair(a, b, M)] = (float)dist64_d(axx, ayy, PXY(N + b, 0), PXY(N + b, 1));
  }
}

// The same reset from the env's entry of the pending block (prefetch_kernel, fm_kernels.cu: placement + assignment of the
// NEXT episode drawn ahead of time from the same Philox stream; the caller has seen q_tag[env] == episode with an acquire
// load).  Same outputs as aw_reset_env, same bits: what is left to do here is min_time against the previous goal_match and
// the distances between the static entities.  The terminal step of an episode was 74 us against 26.5 for a regular one
// (profiles/r02_final_launches_driver_config.csv) -- 9 % of a rollout -- because every env warp walked the serial
// rejection sampling while its three agent warps waited.
template <int N, int O>
__device__ __forceinline__ void aw_reset_from_pending(const DevParams& p, int env, float* __restrict__ Tc, float* __restrict__ Sc,
                                                      float* __restrict__ sd) {
  using L = AwLayout<N, O>;
  constexpr int RW = L::RW, M = L::M;
  auto PXY = [&](int e, int c) -> float& { return Tc[(L::TP + 2 * e + c) * RW]; };
  const size_t Bp = (size_t)p.Bp;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    PXY(i, 0) = __ldcg(p.q_px + i * Bp + env); PXY(i, 1) = __ldcg(p.q_py + i * Bp + env);
    PXY(N + i, 0) = __ldcg(p.q_lx + i * Bp + env); PXY(N + i, 1) = __ldcg(p.q_ly + i * Bp + env);
  }
#pragma unroll
  for (int k = 0; k < O; ++k) { PXY(2 * N + k, 0) = __ldcg(p.q_ox + k * Bp + env); PXY(2 * N + k, 1) = __ldcg(p.q_oy + k * Bp + env); }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (p.has_max_speed) {                 // min_time with the PREVIOUS goal_match (:545-547, :719-728)
      const int og = __float_as_int(Sc[(L::GMO + i) * RW]);
      Sc[(L::RMINT + i) * RW] = (float)(dist64(PXY(i, 0), PXY(i, 1), PXY(N + og, 0), PXY(N + og, 1)) / p.max_speed);
    }
    Sc[(L::RGM + i) * RW] = __int_as_float(__ldcg(p.q_gm + i * Bp + env));
  }
#pragma unroll 1
  for (int a = 0; a < M; ++a)
#pragma unroll 1
    for (int b = a + 1; b < M; ++b)
      sd[aw_spair(a, b, M)] = (float)dist64(PXY(N + a, 0), PXY(N + a, 1), PXY(N + b, 0), PXY(N + b, 1));
}

// Per-launch (one-shot kernels) or per-step (rollout kernel) inputs and outputs of a tile.
// The output pointers are read where they are used (one-shot kernels: a local built from the kernel parameters, i.e.
// constant-bank operands; rollout kernel: the step's entry of a shared-memory table), so they hold no registers
// across the compute phases.
struct AwIo {
  const int* act_idx;        // [B,N] or null
  const float* act_onehot;   // [B,N,5] or null
  const uint8_t* reset_mask; // MODE 1 only; null = all
  const FmOutputs* out;
};

// Rollout-kernel scheduling state of a CTA (fm_roll.cu): work items are (step t, tile k), claimed in step-major order;
// item (t, k) may start once flags[k] >= t, i.e. once step t - 1 of the same tile has written its state back.
struct AwRoll {
  int* flags;        // [tiles] steps of this launch completed per tile (zero between launches)
  int tile, t;       // this item
  int prev_tile, prev_t;   // item whose node_obs bulk store may still be in flight (-1: none); released late if !early
  bool early;        // release the tile right after the state write-back (every step of the launch writes distinct outputs)
  bool multi;        // the launch has more than one step (flags are in use)
  int* s_next;       // shared-memory word: the CTA's next item, published by thread 0 at barrier #0
  int nx;            // thread 0: the item it claimed for the next iteration
  int total, ntiles; // items of the launch, tiles per step
  bool next_ready;   // this thread has already seen the next item's dependency flag satisfied (polled early)
};

// The (N, O) pairs compiled for this mapping.  Everything else runs the group-per-env kernels.
#define FM_AW_CASES(X) X(1, 1) X(2, 0) X(3, 0) X(3, 3) X(4, 2)
// (N, O, W) with walls: the 3-agent / 3-obstacle shape of the BASELINE configs with 1 or 2 walls, and the shape of the
// reference fixture n4_o2_w1
#define FM_AW_WALL_CASES(X) X(3, 3, 1) X(3, 3, 2) X(4, 2, 1)

__device__ __forceinline__ int ld_acquire_gpu(const int* q) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* q, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(q), "r"(v) : "memory");
}
// all bulk groups of this thread have COMPLETED (global writes performed, not only shared memory read)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// =============================================================================================
//   MODE 0: fused env step (MultiAgentGraphEnv.step, environment.py:816-877, + graphworker auto-reset,
//           env_wrappers.py:859-865).   MODE 1: masked reset + observe (environment.py:882-898).
// NF: node feature width, NODE_F (relative, 11) or NODE_F_GLOBAL (graph_feat_type = 'global', 7).
// ROLL: called from the persistent rollout kernel: the tile's previous user of the shared memory may still have a bulk
//       store in flight (waited for at barrier #0), the last bulk store of this call is left in flight, and the tile's
//       dependency flag is released (see AwRoll).
template <int N, int O, int MODE, int NF, bool ROLL, int W = 0>
__device__ __forceinline__ void aw_tile(const DevParams& p, const AwIo& io, const int env0, const int nenv,
                                        float* __restrict__ smem, AwRoll& rs) {
  using L = AwLayout<N, O, W>;
  constexpr int E = L::E, M = L::M, RW = L::RW, SP = L::SP;
  constexpr int NW = N * E * NF;                   // node_obs words per env (the staging region is sized for NF = 11)
  float* ST = smem + L::OFF_STAGE;
  const int tid = threadIdx.x, lane = tid & 31, role = tid >> 5;   // role < N: agent `role`;  role == N: env warp
  const int env = env0 + lane;                   // < Bp: the state block is padded to a multiple of 64 envs
  const bool venv = lane < nenv;
  const size_t Bp = (size_t)p.Bp;
  float* gs = p.px + env;                        // state row r of this env: gs[r * Bp]
  float* Tc = smem + lane;                       // table row r of this env: Tc[r * RW]
  float* Sc = ST + L::OFF_SCR + lane;            // scratch row r: Sc[r * RW]
  double* Dc = reinterpret_cast<double*>(ST + L::OFF_SCRD) + lane;
  float* adj = ST + L::S_ADJ + lane * L::ADJ_W;  // this env's adj image
  const long long genv = p.env_offset + env;
  const bool is_agent = role < N;
  const int i = role;

  // ---- registers of thread (env, agent i) ------------------------------------------------------
  float px = 0.f, py = 0.f, vx = 0.f, vy = 0.f, pd = 0.f, dtg = 0.f, treq = 0.f, dleft = 0.f, fobs = 0.f;
  int gm = 0, nac = 0, noc = 0;
  unsigned collbits = 0, reachbits = 0;          // bit e: float64 d(i, e) < collision distance / < goal threshold
  float dgoal_f = 0.f, own_rew = 0.f, rew_out = 0.f;
  // ---- env warp / common ---------------------------------------------------------------------------
  int step = 0, epis = 0, nstep = 0;
  bool do_reset = false, done = false;

  // Row (and column) i of the distance matrix at the positions in TP (core.py:204-228), written straight
  // into the adj image, + the predicate bits and the distance to the assigned goal.
  auto agent_distances = [&]() {
    const double ax = (double)px, ay = (double)py;
    collbits = 0; reachbits = 0;
    const int eg = N + gm;
    adj[i * E + i] = 0.0f;
#pragma unroll 4
    for (int k = 1; k < E; ++k) {                  // the E - 1 other entities, ascending from i + 1 (wrapping)
      const int e = (i + k >= E) ? i + k - E : i + k;
      const double dd = dist64_d(ax, ay, Tc[(L::TP + 2 * e) * RW], Tc[(L::TP + 2 * e + 1) * RW]);
      const float df = (float)dd;
      adj[i * E + e] = df;
      if (e >= N) adj[e * E + i] = df;
      collbits |= (dd < p.dcoll) ? (1u << e) : 0u;
      reachbits |= (dd < p.min_dist_thresh) ? (1u << e) : 0u;
      dgoal_f = (e == eg) ? df : dgoal_f;
    }
  };
  // Landmark/obstacle block of the adj image from the cached static distances (env warp).
  auto static_block = [&](const float* sd) {
#pragma unroll
    for (int x = 0; x < M; ++x) {
      adj[(N + x) * E + (N + x)] = 0.0f;
#pragma unroll
      for (int y = x + 1; y < M; ++y) {
        const float v = sd[aw_spair(x, y, M)];
        adj[(N + x) * E + (N + y)] = v;
        adj[(N + y) * E + (N + x)] = v;
      }
    }
  };
  // Static entities: state block -> registers (env warp), then registers -> TP table.
  auto load_static = [&](float* sxy, float* sd) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      sxy[2 * j] = __ldcg(gs + (size_t)(L::LX + j) * Bp);
      sxy[2 * j + 1] = __ldcg(gs + (size_t)(L::LY + j) * Bp);
    }
#pragma unroll
    for (int k = 0; k < O; ++k) {
      sxy[2 * (N + k)] = __ldcg(gs + (size_t)(L::OX + k) * Bp);
      sxy[2 * (N + k) + 1] = __ldcg(gs + (size_t)(L::OY + k) * Bp);
    }
#pragma unroll
    for (int k = 0; k < W; ++k) {                  // wall midpoint: (0, axis) for 'H', (axis, 0) for 'V' (navigation_graph.py:309-324)
      const float ax = __ldcg(p.wax + (size_t)k * Bp + env);
      const bool hz = __ldcg(p.wor + (size_t)k * Bp + env) == 0;
      sxy[2 * (N + O + k)] = hz ? 0.0f : ax;
      sxy[2 * (N + O + k) + 1] = hz ? ax : 0.0f;
      Tc[(L::TWA + k) * RW] = ax; Tc[(L::TWO + k) * RW] = hz ? 0.0f : 1.0f;
    }
    if (W > 0) Tc[L::TWL * RW] = __ldcg(p.wlen + env);
#pragma unroll
    for (int q = 0; q < SP; ++q) sd[q] = __ldcg(gs + (size_t)(L::SDIST + q) * Bp);
  };
  auto store_static = [&](const float* sxy) {
#pragma unroll
    for (int x = 0; x < M; ++x) {
      Tc[(L::TP + 2 * (N + x)) * RW] = sxy[2 * x];
      Tc[(L::TP + 2 * (N + x) + 1) * RW] = sxy[2 * x + 1];
    }
  };
  // Reset of this env by its env-warp thread: new placement -> tables, scratch and the state block.
  auto reset_static = [&](float* sd) {
    bool pend = false;
    if constexpr (W == 0) {                      // (the pending block carries no walls)
      if (p.q_tag != nullptr) pend = ld_acquire_gpu(p.q_tag + env) == epis;   // the entry's data is visible once its tag is
      if (pend) aw_reset_from_pending<N, O>(p, env, Tc, Sc, sd);
    }
    if (!pend) aw_reset_env<N, O, W>(p, genv, (uint32_t)epis, Tc, Sc, sd);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      p.wax[(size_t)k * Bp + env] = Tc[(L::TWA + k) * RW];
      p.wor[(size_t)k * Bp + env] = Tc[(L::TWO + k) * RW] == 0.0f ? 0 : 1;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      gs[(size_t)(L::LX + j) * Bp] = Tc[(L::TP + 2 * (N + j)) * RW];
      gs[(size_t)(L::LY + j) * Bp] = Tc[(L::TP + 2 * (N + j) + 1) * RW];
    }
#pragma unroll
    for (int k = 0; k < O; ++k) {
      gs[(size_t)(L::OX + k) * Bp] = Tc[(L::TP + 2 * (2 * N + k)) * RW];
      gs[(size_t)(L::OY + k) * Bp] = Tc[(L::TP + 2 * (2 * N + k) + 1) * RW];
    }
#pragma unroll
    for (int q = 0; q < SP; ++q) gs[(size_t)(L::SDIST + q) * Bp] = sd[q];
    gs[(size_t)L::EPIS * Bp] = __int_as_float(epis + 1);
    static_block(sd);
  };
  // Barrier #0 of the rollout kernel: the shared memory of this CTA is free again (the previous item's threads are
  // past their last table reads, the copy engine has read -- or, for a late release, written -- the last image).
  auto roll_barrier0 = [&]() {
    if (ROLL) {
      if (tid == 0) {
        if (rs.prev_tile >= 0) bulk_wait_read<0>();  // (a late release has already waited for the whole store, fm_roll.cu)
        *rs.s_next = rs.nx;
      }
      __syncthreads();
    }
  };

  if (MODE == 0) {
    // =========================================================================================
    double pd64 = 0.0;
    float dmean0 = 0.f, dstd0 = 0.f;
    // ---- P0: global loads into registers (own state + what the forces need, straight from the SoA state block) ----
    float qx[N + O], qy[N + O];                  // agent warps: partners (agents, own slot unused; obstacles)
    float sxy[2 * M > 0 ? 2 * M : 1], sd0[SP > 0 ? SP : 1];   // env warp: static positions, cached static distances
    float ux = 0.f, uy = 0.f;
    float wax[W > 0 ? W : 1], wlen = 0.f;          // agent warps: wall axis / orientation / half-length of THIS episode
    bool whz[W > 0 ? W : 1];
    if (is_agent) {
      const float* gi = gs + (size_t)i * Bp;
      px = __ldcg(gi + (size_t)L::PX * Bp); py = __ldcg(gi + (size_t)L::PY * Bp);
      vx = __ldcg(gi + (size_t)L::VX * Bp); vy = __ldcg(gi + (size_t)L::VY * Bp);
      pd = __ldcg(gi + (size_t)L::PD * Bp); dtg = __ldcg(gi + (size_t)L::DTG * Bp);
      treq = __ldcg(gi + (size_t)L::TREQ * Bp); dleft = __ldcg(gi + (size_t)L::DLEFT * Bp);
      gm = __float_as_int(__ldcg(gi + (size_t)L::GM * Bp));
      nac = __float_as_int(__ldcg(gi + (size_t)L::NAC * Bp));
      noc = __float_as_int(__ldcg(gi + (size_t)L::NOC * Bp));
      step = __float_as_int(__ldcg(gs + (size_t)L::STEP * Bp));
#pragma unroll
      for (int j = 0; j < N; ++j) { qx[j] = __ldcg(gs + (size_t)(L::PX + j) * Bp); qy[j] = __ldcg(gs + (size_t)(L::PY + j) * Bp); }
#pragma unroll
      for (int k = 0; k < O; ++k) { qx[N + k] = __ldcg(gs + (size_t)(L::OX + k) * Bp); qy[N + k] = __ldcg(gs + (size_t)(L::OY + k) * Bp); }
      if (i == 0) { dmean0 = __ldcg(gs + (size_t)L::DMEAN * Bp); dstd0 = __ldcg(gs + (size_t)L::DSTD * Bp); }
#pragma unroll
      for (int k = 0; k < W; ++k) { wax[k] = __ldcg(p.wax + (size_t)k * Bp + env); whz[k] = __ldcg(p.wor + (size_t)k * Bp + env) == 0; }
      if (W > 0) wlen = __ldcg(p.wlen + env);
      if (venv) {                                // environment.py:301-311: u = [a1 - a2, a3 - a4] * sensitivity (5.0)
        if (io.act_idx) {
          const int a = __ldg(io.act_idx + (size_t)env * N + i);
          ux = ((a == 1) ? 1.f : 0.f) - ((a == 2) ? 1.f : 0.f);
          uy = ((a == 3) ? 1.f : 0.f) - ((a == 4) ? 1.f : 0.f);
        } else {
          const float* oh = io.act_onehot + ((size_t)env * N + i) * 5;
          ux = __ldg(oh + 1) - __ldg(oh + 2);
          uy = __ldg(oh + 3) - __ldg(oh + 4);
        }
        ux *= 5.0f; uy *= 5.0f;
      }
    } else {
      load_static(sxy, sd0);
      step = __float_as_int(__ldcg(gs + (size_t)L::STEP * Bp));
      epis = __float_as_int(__ldcg(gs + (size_t)L::EPIS * Bp));
    }
    roll_barrier0();
    if (is_agent) {
      // ---- P1: forces on agent i (core.py:277-316, :370-404), partners in ascending entity index;
      // a pair (j, i), j < i, contributes -f(j, i) = f computed from agent i's side (IEEE sign symmetry).
      float cfx = 0.f, cfy = 0.f;
#pragma unroll
      for (int q = 0; q < N + O; ++q)
        if (q != i) contact_force(p, px, py, qx[q], qy[q], cfx, cfy);
      // walls: first as circle entities of the pair loop (they close world.entities), then the wall forces proper
      // (core.py:317-327, :407-462), both from the positions at step entry -- the order of step_kernel<G, true>
#pragma unroll
      for (int k = 0; k < W; ++k) contact_force_dmin(p, 0.15f, px, py, whz[k] ? 0.0f : wax[k], whz[k] ? wax[k] : 0.0f, cfx, cfy);
#pragma unroll
      for (int k = 0; k < W; ++k) wall_force(px, py, whz[k], wax[k], wlen, cfx, cfy);
      const double Fx = __dadd_rn((double)ux, (double)cfx), Fy = __dadd_rn((double)uy, (double)cfy);   // mass(1.0) * u + contact
      double v64x, v64y, sx, sy;                   // integrate_state (core.py:338-356)
      integrate64(p, vx, vy, Fx, Fy, pd, v64x, v64y, sx, sy, pd64);
      px = (float)__dadd_rn((double)px, sx); py = (float)__dadd_rn((double)py, sy);
      vx = (float)v64x; vy = (float)v64y; pd = (float)pd64;
      Tc[(L::TP + 2 * i) * RW] = px; Tc[(L::TP + 2 * i + 1) * RW] = py;
      Tc[(L::TV + 2 * i) * RW] = vx; Tc[(L::TV + 2 * i + 1) * RW] = vy;
      Sc[(L::DTGO + i) * RW] = dtg; Sc[(L::TREQO + i) * RW] = treq;
      Sc[(L::GMO + i) * RW] = __int_as_float(gm);
      Dc[(L::PD64 + i) * RW] = pd64;
    } else {
      store_static(sxy);
      static_block(sd0);
    }
    __syncthreads();                              // #1: new agent positions, static positions (env warp) visible
    if (ROLL && rs.multi) {
      // The next item of this CTA was published at barrier #0: look at its tile's flag NOW, so that the round trip of the
      // acquire load is hidden behind this item's compute instead of sitting in front of the next item's state loads.
      const int nitem = *rs.s_next;
      rs.next_ready = true;
      if (nitem < rs.total) {
        const int nt = nitem / rs.ntiles, ntile = nitem - nt * rs.ntiles;
        if (nt > 0) rs.next_ready = ld_acquire_gpu(rs.flags + ntile) >= nt;
      }
    }
    nstep = step + 1;                              // environment.py:819, :823
    done = nstep >= p.episode_length;              // environment.py:237-247 (agent.status is never set)
    do_reset = venv && done && (p.auto_reset != 0);
    if (is_agent) {
      // ---- P2: distances, statistic sets, observation scalar, reward, latches ---------------------
      Tc[(L::TG + 2 * i) * RW] = Tc[(L::TP + 2 * (N + gm)) * RW];
      Tc[(L::TG + 2 * i + 1) * RW] = Tc[(L::TP + 2 * (N + gm) + 1) * RW];
      agent_distances();
      // world.dists_to_goal as left by the previous agent's info_callback: set k over
      // [new_0..new_{k-1}, prev_k..] (navigation_graph.py:587-598, :617-618).  Agent i >= 1 needs set i;
      // agent 0 reads last step's value from the state and computes set N (the value after this step).
      const int kset = (i == 0) ? N : i;
      const bool first = dtg == -1.0f;             // first step of the episode: statistics of the new travelled distances
      double v[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const double pj = Dc[(L::PD64 + j) * RW];
        const double dj = (double)Sc[(L::DTGO + j) * RW];
        const bool lat = Sc[(L::TREQO + j) * RW] != -1.0f;
        v[j] = (j < kset && !lat) ? pj : dj;
      }
      double mk, sk;
      mean_std<N>(v, mk, sk);
      Dc[(L::SETM + kset) * RW] = mk; Dc[(L::SETS + kset) * RW] = sk;
      float fparam;                                // navigation_graph.py:764-769 / :849-853
      if (first) {
        double w[N];
#pragma unroll
        for (int j = 0; j < N; ++j) w[j] = Dc[(L::PD64 + j) * RW];
        double m0, s0;
        mean_std<N>(w, m0, s0);
        fparam = ratio_eps(m0, s0);
      } else if (i == 0) {
        fparam = ratio_eps((double)dmean0, (double)dstd0);
      } else {
        fparam = ratio_eps(mk, sk);
      }
      const bool reached = ((reachbits >> (N + gm)) & 1u) != 0;      // dgoal < min_dist_thresh (float64 compare)
      const int ncoll = __popc(collbits & ((1u << N) - 1u));
      bool ocoll = W > 0 ? ((collbits >> (2 * N)) & ((1u << O) - 1u)) != 0 : (collbits >> (2 * N)) != 0;   // obstacles only
#pragma unroll
      for (int k = 0; k < W; ++k) ocoll = ocoll || in_wall_box(px, py, whz[k], wax[k], wlen);   // :670-683, at the new position
      const bool latched = treq != -1.0f;
      const double treq_new = (!latched && reached) ? (double)nstep * p.dt : (double)treq;   // :588
      float rw = reached ? p.goal_rew : -dgoal_f;  // navigation_graph.py:760-824
      rw -= p.coll_rew * (float)ncoll;
      if (ocoll) rw -= p.coll_rew;
      if (p.fairness_reward) {
        float fair = p.fair_rew * tanhf(fparam - p.zeroshift_f);
        if (fair < -2.0f) fair = -2.0f;
        rw += fair;
      }
      rw = fminf(fmaxf(rw, p.clip_lo), p.clip_hi);
      nac += ncoll;                                // :604-613
      noc += ocoll ? 1 : 0;                        // :602-603
      own_rew = rw;
      fobs = fparam;
      dtg = latched ? dtg : pd;                    // pd == (float)pd64
      dleft = latched ? dleft : dgoal_f;
      Sc[(L::OWN + i) * RW] = rw;
      Sc[(L::NTREQ + i) * RW] = (float)treq_new;   // `treq` keeps the old value for the info pass
      if (i == 0 && venv) {                        // world.dist_traveled_mean / stddev after the last info_callback
        gs[(size_t)L::DMEAN * Bp] = (float)mk;
        gs[(size_t)L::DSTD * Bp] = (float)sk;
      }
    }
    const bool want_info = venv && ((float*)io.out->info != nullptr || p.stats != nullptr) && (done || p.info_every_step);
    double* stats_row = p.stats ? p.stats + (size_t)(env0 >> 5) * (15 * N + 2) : nullptr;
    const bool any_done = __syncthreads_or(venv && done) != 0;   // #2: OWN / NTREQ / SETM / SETS / TG visible
    const bool any_reset = any_done && (p.auto_reset != 0);
    const bool any_info = ((float*)io.out->info != nullptr || p.stats != nullptr) && (any_done || p.info_every_step);

    // ---- collaborative sum, episode statistics, info rows -------------------------------------------
    rew_out = own_rew;
    if (is_agent) {
      if (p.collaborative) {                       // environment.py:866-870
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) tot += Sc[(L::OWN + j) * RW];
        rew_out = tot;
      }
      if (stats_row) {
        double s = venv ? (double)rew_out : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
        if (lane == 0) atomicAdd(stats_row + i, s);                  // one add per (row, step): order is fixed
      }
      if (any_info) {
        // world-level time statistics right after agent i's own info_callback: new values of agents
        // j <= i, previous values of j > i (navigation_graph.py:620-621)
        double tacc = 0.0;                         // entity.state.time += dt per step (core.py:355)
#pragma unroll 1
        for (int k = 0; k < nstep; ++k) tacc += p.dt;
        double tv[N];                              // times_required as float64: a latch of THIS step is nstep * dt unrounded
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float told = Sc[(L::TREQO + j) * RW];
          const bool fresh = j <= i && told == -1.0f && Sc[(L::NTREQ + j) * RW] != -1.0f;
          tv[j] = fresh ? (double)nstep * p.dt : (double)told;
        }
        double mt, stv;
        mean_std<N>(tv, mt, stv);
        const double md = Dc[(L::SETM + i + 1) * RW];
        const double sdv = Dc[(L::SETS + i + 1) * RW];
        float info[INFO_F];
        info[0] = own_rew; info[1] = dleft; info[2] = Sc[(L::NTREQ + i) * RW];
        info[3] = (float)nac; info[4] = (float)noc;
        info[5] = (float)md; info[6] = (float)sdv; info[7] = ratio_eps(md, sdv);
        info[8] = dtg; info[9] = (float)tacc; info[10] = (float)mt; info[11] = (float)stv;
        info[12] = ratio_eps(mt, stv); info[13] = __ldcg(gs + (size_t)(L::MINT + i) * Bp);
        if (want_info && (float*)io.out->info) {
          float* o = (float*)io.out->info + ((size_t)env * N + i) * INFO_F;
#pragma unroll
          for (int k = 0; k < INFO_F; ++k) o[k] = info[k];
        }
        if (stats_row && __any_sync(FULL, venv && done)) {
#pragma unroll
          for (int k = 0; k < INFO_F; ++k) {
            double s = (venv && done) ? (double)info[k] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
            if (lane == 0) atomicAdd(stats_row + N + i * INFO_F + k, s);
          }
        }
      }
      treq = Sc[(L::NTREQ + i) * RW];
    } else if (stats_row) {
      const unsigned termb = __ballot_sync(FULL, venv && done);
      if (lane == 0) { atomicAdd(stats_row + 15 * N, (double)__popc(termb)); atomicAdd(stats_row + 15 * N + 1, (double)nenv); }
    }

    // ---- auto-reset (env_wrappers.py:859-865): obs / node_obs / adj come from the new episode, reward /
    // done / info stay terminal ------------------------------------------------------------------------
    if (any_reset) {
      if (!is_agent && do_reset) {
        float sd[SP > 0 ? SP : 1];
        reset_static(sd);
      }
      __syncthreads();                            // new positions / goal_match of the envs that reset
      if (is_agent) {
        if (do_reset) {
          px = Tc[(L::TP + 2 * i) * RW]; py = Tc[(L::TP + 2 * i + 1) * RW];
          vx = 0.f; vy = 0.f; pd = 0.f; dtg = -1.f; treq = -1.f; dleft = -1.f; nac = 0; noc = 0; fobs = 0.f;
          gm = __float_as_int(Sc[(L::RGM + i) * RW]);
          Tc[(L::TV + 2 * i) * RW] = 0.f; Tc[(L::TV + 2 * i + 1) * RW] = 0.f;
          Tc[(L::TG + 2 * i) * RW] = Tc[(L::TP + 2 * (N + gm)) * RW];
          Tc[(L::TG + 2 * i + 1) * RW] = Tc[(L::TP + 2 * (N + gm) + 1) * RW];
          gs[(size_t)(L::GM + i) * Bp] = __int_as_float(gm);
          if (p.has_max_speed) gs[(size_t)(L::MINT + i) * Bp] = Sc[(L::RMINT + i) * RW];
        }
        agent_distances();
      }
    }
    // ---- state write-back: agent rows from registers, one coalesced line per row --------------------
    if (venv) {
      if (is_agent) {
        float* gi = gs + (size_t)i * Bp;
        gi[(size_t)L::PX * Bp] = px; gi[(size_t)L::PY * Bp] = py;
        gi[(size_t)L::VX * Bp] = vx; gi[(size_t)L::VY * Bp] = vy;
        gi[(size_t)L::PD * Bp] = pd; gi[(size_t)L::DTG * Bp] = dtg;
        gi[(size_t)L::TREQ * Bp] = treq; gi[(size_t)L::DLEFT * Bp] = dleft;
        gi[(size_t)L::NAC * Bp] = __int_as_float(nac); gi[(size_t)L::NOC * Bp] = __int_as_float(noc);
      } else {
        gs[(size_t)L::STEP * Bp] = __int_as_float(do_reset ? 0 : nstep);
      }
    }
    if (is_agent) {
      ST[L::S_REW + lane * N + i] = rew_out;
      reinterpret_cast<uint8_t*>(ST + L::S_DONE)[lane * N + i] = done ? 1 : 0;
    }
  } else {
    // =========================================================================================
    // MODE 1: reset() / observe
    do_reset = venv && !p.observe_only && (io.reset_mask ? (io.reset_mask[env] != 0) : true);
    if (is_agent) {
      const float* gi = gs + (size_t)i * Bp;
      px = __ldcg(gi + (size_t)L::PX * Bp); py = __ldcg(gi + (size_t)L::PY * Bp);
      vx = __ldcg(gi + (size_t)L::VX * Bp); vy = __ldcg(gi + (size_t)L::VY * Bp);
      dtg = __ldcg(gi + (size_t)L::DTG * Bp);
      gm = __float_as_int(__ldcg(gi + (size_t)L::GM * Bp));
      double w[N];
#pragma unroll
      for (int j = 0; j < N; ++j) w[j] = (double)__ldcg(gs + (size_t)(L::PD + j) * Bp);
      const float dmean0 = __ldcg(gs + (size_t)L::DMEAN * Bp), dstd0 = __ldcg(gs + (size_t)L::DSTD * Bp);
      Tc[(L::TP + 2 * i) * RW] = px; Tc[(L::TP + 2 * i + 1) * RW] = py;
      Tc[(L::TV + 2 * i) * RW] = vx; Tc[(L::TV + 2 * i + 1) * RW] = vy;
      Sc[(L::GMO + i) * RW] = __int_as_float(gm);
      // observation() on the current state (navigation_graph.py:826-857, :849-853)
      double mean_p, std_p;
      mean_std<N>(w, mean_p, std_p);
      fobs = (dtg == -1.0f) ? ratio_eps(mean_p, std_p) : ratio_eps((double)dmean0, (double)dstd0);
    } else {
      float sxy[2 * M > 0 ? 2 * M : 1], sd[SP > 0 ? SP : 1];
      load_static(sxy, sd);
      store_static(sxy);
      epis = __float_as_int(__ldcg(gs + (size_t)L::EPIS * Bp));
      static_block(sd);
    }
    const bool any_reset = __syncthreads_or(do_reset) != 0;      // static positions / GMO visible
    if (any_reset) {
      if (!is_agent && do_reset) {
        float sd[SP > 0 ? SP : 1];
        reset_static(sd);
        gs[(size_t)L::STEP * Bp] = __int_as_float(0);
      }
      __syncthreads();
      if (is_agent && do_reset) {
        px = Tc[(L::TP + 2 * i) * RW]; py = Tc[(L::TP + 2 * i + 1) * RW];
        vx = 0.f; vy = 0.f; fobs = 0.f;           // mean(p_dist = 0) / (std + 1e-4)
        gm = __float_as_int(Sc[(L::RGM + i) * RW]);
        Tc[(L::TV + 2 * i) * RW] = 0.f; Tc[(L::TV + 2 * i + 1) * RW] = 0.f;
        float* gi = gs + (size_t)i * Bp;
        gi[(size_t)L::PX * Bp] = px; gi[(size_t)L::PY * Bp] = py;
        gi[(size_t)L::VX * Bp] = 0.f; gi[(size_t)L::VY * Bp] = 0.f; gi[(size_t)L::PD * Bp] = 0.f;
        gi[(size_t)L::DTG * Bp] = -1.f; gi[(size_t)L::TREQ * Bp] = -1.f; gi[(size_t)L::DLEFT * Bp] = -1.f;
        gi[(size_t)L::NAC * Bp] = __int_as_float(0); gi[(size_t)L::NOC * Bp] = __int_as_float(0);
        gi[(size_t)L::GM * Bp] = __int_as_float(gm);
        if (p.has_max_speed) gi[(size_t)L::MINT * Bp] = Sc[(L::RMINT + i) * RW];
      }
    }
    if (is_agent) {
      Tc[(L::TG + 2 * i) * RW] = Tc[(L::TP + 2 * (N + gm)) * RW];
      Tc[(L::TG + 2 * i + 1) * RW] = Tc[(L::TP + 2 * (N + gm) + 1) * RW];
      agent_distances();
    }
  }

  // =============================================================================================
  // Emission, use 1 of the staging region: adj (already in place) | obs | reward | done.
  if (is_agent) {
    float* o = ST + L::S_OBS + lane * L::OBS_W + i * OBS_F;       // navigation_graph.py:826-857
    const float gx = Tc[(L::TG + 2 * i) * RW], gy = Tc[(L::TG + 2 * i + 1) * RW];
    o[0] = vx; o[1] = vy; o[2] = px; o[3] = py; o[4] = gx - px; o[5] = gy - py; o[6] = fobs;
  }
  __syncthreads();                                // #3: image of the small outputs + TP / TV / TG complete; state written back
  // Step t of this tile is in the state block: step t + 1 may start.  Released by the first lane of the env warp (which has
  // nothing else to do here), so that thread 0 goes straight to the bulk stores instead of sitting in the fence.
  if (ROLL && rs.multi && rs.early && tid == N * 32) { __threadfence(); st_release_gpu(rs.flags + rs.tile, rs.t + 1); }
  // A full tile whose slices of the output arrays are 16-byte aligned (always, when the arrays are) goes out
  // as TMA bulk stores issued by one thread; ragged or unaligned tiles use vectorised st.global.cs.
  float* g_adj = (float*)io.out->adj ? (float*)io.out->adj + (size_t)env0 * L::ADJ_W : nullptr;
  float* g_obs = (float*)io.out->obs ? (float*)io.out->obs + (size_t)env0 * L::OBS_W : nullptr;
  float* g_rew = (MODE == 0 && (float*)io.out->reward) ? (float*)io.out->reward + (size_t)env0 * N : nullptr;
  uint8_t* g_done = (MODE == 0 && (uint8_t*)io.out->done) ? (uint8_t*)io.out->done + (size_t)env0 * N : nullptr;
  float* g_node = (float*)io.out->node_obs ? (float*)io.out->node_obs + (size_t)env0 * NW : nullptr;
  const bool bulk = nenv == 32 && aligned16(g_adj) && aligned16(g_obs) && aligned16(g_rew) && aligned16(g_done) &&
                    aligned16(g_node) && (32 * N) % 16 == 0;
  if (bulk) {
    if (tid == 0) {
      const uint64_t pol = evict_first_policy();
      fence_async_smem();
      if (g_adj) bulk_store(g_adj, ST + L::S_ADJ, 32 * L::ADJ_W * 4, pol);
      if (g_obs) bulk_store(g_obs, ST + L::S_OBS, 32 * L::OBS_W * 4, pol);
      if (g_rew) bulk_store(g_rew, ST + L::S_REW, 32 * N * 4, pol);
      if (g_done) bulk_store(g_done, ST + L::S_DONE, 32 * N, pol);
      bulk_commit();
      if (g_node || !ROLL) bulk_wait_read<0>();   // the node_obs image is about to overwrite these
    }
  } else {
    if (g_adj) cta_copy_out<L::THREADS, 32 * L::ADJ_W>(g_adj, ST + L::S_ADJ, nenv * L::ADJ_W, tid);
    if (g_obs) cta_copy_out<L::THREADS, 32 * L::OBS_W>(g_obs, ST + L::S_OBS, nenv * L::OBS_W, tid);
    if (g_rew) cta_copy_out<L::THREADS, 32 * N>(g_rew, ST + L::S_REW, nenv * N, tid);
    if (g_done) {
      const uint8_t* sdone = reinterpret_cast<const uint8_t*>(ST + L::S_DONE);
      for (int k = tid; k < nenv * N; k += L::THREADS) g_done[k] = sdone[k];
    }
  }
  if (g_node) {                                   // CTA-uniform
    __syncthreads();                              // #4: staging region fully read
    // ---- use 2: node_obs (navigation_graph.py:1079-1124, relative features): for ego agent i and entity e
    //   [v_e - v_i (2), p_e - p_i (2), goal_e - p_i (2), p_e - p_i (2), p_e - p_i (2), type]
    // with goal_e = assigned landmark for agents and = p_e otherwise, v_e = 0 for non-agents.  Agent warp i
    // writes ego i's E rows from its registers and the tables; lane = env, stride NODE_W (odd): conflict free.
    if (is_agent) {
      float* o = ST + lane * NW + i * E * NF;
      const float nvx = 0.f - vx, nvy = 0.f - vy;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float epx = Tc[(L::TP + 2 * e) * RW], epy = Tc[(L::TP + 2 * e + 1) * RW];
        const float ty = (e < N) ? 0.0f : ((e < 2 * N) ? 1.0f : ((e < 2 * N + O) ? 2.0f : 3.0f));
        if (NF == NODE_F_GLOBAL) {                    // [vel, pos, goal, type] in world coordinates, same rows for every ego agent
          float evx = 0.f, evy = 0.f, egx = epx, egy = epy;
          if (e < N) {
            evx = Tc[(L::TV + 2 * e) * RW]; evy = Tc[(L::TV + 2 * e + 1) * RW];
            egx = Tc[(L::TG + 2 * e) * RW]; egy = Tc[(L::TG + 2 * e + 1) * RW];
          }
          o[0] = evx; o[1] = evy; o[2] = epx; o[3] = epy; o[4] = egx; o[5] = egy; o[6] = ty;
        } else {
          const float rpx = epx - px, rpy = epy - py;
          float rvx = nvx, rvy = nvy, rgx = rpx, rgy = rpy;
          if (e < N) {
            rvx = Tc[(L::TV + 2 * e) * RW] - vx; rvy = Tc[(L::TV + 2 * e + 1) * RW] - vy;
            rgx = Tc[(L::TG + 2 * e) * RW] - px; rgy = Tc[(L::TG + 2 * e + 1) * RW] - py;
          }
          o[0] = rvx; o[1] = rvy; o[2] = rpx; o[3] = rpy; o[4] = rgx; o[5] = rgy;
          if (W > 0 && e >= 2 * N + O) {
            // wall row (navigation_graph.py:1108-1118): rel_goal = rel_pos, then the two corner offsets
            // (endpoints[0], axis + width / 2) - p_a and (endpoints[1], axis - width / 2) - p_a
            const float wa = Tc[(L::TWA + (e - 2 * N - O)) * RW], wl = Tc[L::TWL * RW];
            o[6] = -wl - px; o[7] = (wa + 0.05f) - py; o[8] = wl - px; o[9] = (wa - 0.05f) - py;
          } else {
            o[6] = rpx; o[7] = rpy; o[8] = rpx; o[9] = rpy;
          }
          o[10] = ty;
        }
        o += NF;
      }
    }
    __syncthreads();                              // #5
    if (bulk) {
      if (tid == 0) {
        fence_async_smem();
        bulk_store(g_node, ST, 32 * NW * 4, evict_first_policy());
        bulk_commit();
        if (!ROLL) bulk_wait_read<0>();           // the image must stay valid until the copy engine has read it
      }
    } else {
      cta_copy_out<L::THREADS, 32 * NW>(g_node, ST, nenv * NW, tid);
    }
  }
}

}  // namespace fm
