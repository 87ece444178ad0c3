// Env-tile kernels: one CTA owns a tile of 32 consecutive envs; lane <-> env, and the WORK ITEMS of
// each phase (force pairs, agents, distance pairs, statistic sets, output words) are spread over the
// CTA's warps.  Specialised at compile time on (N, O).
//
// Why this mapping (measured on B200, C2 = 65 536 envs x 3 agents):
//   * group-per-env (fm_kernels.cu, G lanes per env, runtime N): 35 M warp instructions per step,
//     issue bound, 55 us / step  (34 % of the HBM roofline).
//   * thread-per-env (everything of one env in one thread's registers): 11 M warp instructions, but
//     65 536 threads are only ~14 warps per SM: latency bound at IPC 0.23, 44 us / step.
//   * this file: same instruction economy (every item index is WARP-UNIFORM, so there is no per-lane
//     index arithmetic, no idle agent lanes, no shuffles in the hot phases), 4 threads per env, i.e.
//     ~55 warps per SM in flight, short dependency chains per item, ~2.5 k SASS instructions.
//
// Shared memory per CTA (all conflict free by construction):
//   S  [row][32]   SoA mirror of the global state block (same row order: one coalesced 128-byte line
//                  per row in, one out) + scratch rows (actions, force terms, rewards, flags)
//   SD [k][32]     doubles: new travelled distance per agent, mean / std sets
//   T  [32][TW]    per-env gather table for the outputs (positions, velocities, goals, fairness obs,
//                  constants 0/1/2), TW odd -> lane = env writes are conflict free
//   D  [32][DW]    per-env pair distances (float) + a zero slot for the diagonal, DW odd
//   FL [pair][32]  per pair predicate bits (d < collision distance, d < goal threshold), fp64 compares
//
// Output emission is a pure gather: every output word of obs / node_obs / adj is  T[s1] - T[s2]  (or
// D[s]) with (s1, s2) from a small per-(N, O) table built on the host, so lane = output word gives
// fully coalesced 128-byte stores straight from the table with no staging, no index arithmetic and
// no barrier; the 8 envs a warp emits sit at compile-time offsets from each other.
//
// Arithmetic is operation for operation that of step_kernel<G> / reset_kernel<G> (fm_kernels.cu):
// same fp32 force terms, fp64 force sums / integration / distances / statistics, same Philox draws;
// tests/test_gpu_parity.py checks the mappings against each other bit for bit.
#include <utility>
#include <vector>

#include "fm_device.cuh"
#include "fm_launch.h"
#include "fm_small.cuh"

namespace fm {

constexpr int TILE_WARPS = 4;
constexpr int TILE_THREADS = TILE_WARPS * 32;
constexpr int TILE_EPW = 32 / TILE_WARPS;     // envs emitted per warp

template <int N, int O>
struct TileLayout {
  static constexpr int E = 2 * N + O;
  static constexpr int PAIRS = E * (E - 1) / 2;
  static constexpr int NAA = N * (N - 1) / 2;          // agent-agent force items
  static constexpr int NF = NAA + N * O;               // force items
  // ---- S rows.  Rows [0, NROWS) mirror the global state block (fm_abi.cu fm_create order).
  static constexpr int PX = 0, PY = PX + N, VX = PY + N, VY = VX + N, PD = VY + N, DTG = PD + N, TREQ = DTG + N,
                       DLEFT = TREQ + N, MINT = DLEFT + N, GM = MINT + N, NAC = GM + N, NOC = NAC + N, LX = NOC + N,
                       LY = LX + N, OX = LY + N, OY = OX + O, DMEAN = OY + O, DSTD = DMEAN + 1, STEP = DSTD + 1,
                       EPIS = STEP + 1, NROWS = EPIS + 1;
  // scratch rows
  static constexpr int UX = NROWS, UY = UX + N, FT = UY + N, OWN = FT + 2 * NF, FOBS = OWN + N, NTREQ = FOBS + N,
                       NDMEAN = NTREQ + N,
                       NDSTD = NDMEAN + 1, RFLAG = NDSTD + 1, S_ROWS = RFLAG + 1;
  // ---- SD rows (doubles)
  static constexpr int PD64 = 0, VM = PD64 + N, VS = VM + N + 1, SD_ROWS = VS + N + 1;
  // ---- T fields
  static constexpr int TP = 0, TV = 2 * E, TG = TV + 2 * N, TF = TG + 2 * N, TZERO = TF + N, TONE = TZERO + 1,
                       TTWO = TONE + 1, TW = (TTWO + 1) | 1;
  static constexpr int DZERO = PAIRS, DW = (PAIRS + 1) | 1;
  // ---- carve-up in floats
  static constexpr int OFF_S = 0;
  static constexpr int OFF_SD = OFF_S + ((S_ROWS * 32 + 1) & ~1);
  static constexpr int OFF_T = OFF_SD + 2 * SD_ROWS * 32;
  static constexpr int OFF_D = OFF_T + TW * 32;
  static constexpr int OFF_FL = OFF_D + DW * 32;
  static constexpr int WORDS = OFF_FL + PAIRS * 8;     // PAIRS * 32 bytes
  static constexpr int OBS_W = N * OBS_F, NODE_W = N * E * NODE_F, ADJ_W = E * E;
};

struct TileLuts {
  const uint32_t* obs;     // [7N]      s1 | s2 << 16 into T
  const uint32_t* node;    // [11 N E]  s1 | s2 << 16 into T
  const uint32_t* adj;     // [E E]     index into D
};

__host__ __device__ constexpr int tile_pair_index(int a, int b, int E) { return a * E - a * (a + 1) / 2 + (b - a - 1); }   // a < b

// S row holding the x / y coordinate of entity e (agents, landmarks, obstacles).
template <int N, int O>
__device__ __forceinline__ int row_x(int e) {
  using L = TileLayout<N, O>;
  return e < N ? L::PX + e : (e < 2 * N ? L::LX + (e - N) : L::OX + (e - 2 * N));
}
template <int N, int O>
__device__ __forceinline__ int row_y(int e) {
  using L = TileLayout<N, O>;
  return e < N ? L::PY + e : (e < 2 * N ? L::LY + (e - N) : L::OY + (e - 2 * N));
}

// ---------------------------------------------------------------------------------------------
// Compile-time loops: static_for<B, E, S>(f) calls f(std::integral_constant<int, i>) for i = B, B+S, ... < E,
// so that every item index is a constant expression (immediate shared-memory offsets, no index math).
template <int B, int S, int... I, class F>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, B + S * I>{}), ...);
}
template <int B, int E, int S = 1, class F>
__device__ __forceinline__ void static_for(F&& f) {
  constexpr int count = (E > B) ? (E - B + S - 1) / S : 0;
  static_for_impl<B, S>(static_cast<F&&>(f), std::make_integer_sequence<int, count>{});
}

// pair q (row-major over a < b) -> a, b
__host__ __device__ constexpr int pair_a(int q, int E) { int a = 0; while (q >= E - 1 - a) { q -= E - 1 - a; ++a; } return a; }
__host__ __device__ constexpr int pair_b(int q, int E) { int a = 0; while (q >= E - 1 - a) { q -= E - 1 - a; ++a; } return a + 1 + q; }

// ---------------------------------------------------------------------------------------------
// Phase: distances of the entity pairs q = W, W + 4, ... at the current positions (core.py:204-228)
// + predicate bits.  W is the warp's role: all indices are compile-time constants.
template <int N, int O, int W>
__device__ __forceinline__ void tile_distances(const DevParams& p, const float* __restrict__ S, float* __restrict__ D,
                                               uint8_t* __restrict__ FL, int lane) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  static_for<W, L::PAIRS, TILE_WARPS>([&](auto qc) {
    constexpr int q = decltype(qc)::value;
    constexpr int a = pair_a(q, E), b = pair_b(q, E);
    const double d = dist64(S[row_x<N, O>(a) * 32 + lane], S[row_y<N, O>(a) * 32 + lane],
                            S[row_x<N, O>(b) * 32 + lane], S[row_y<N, O>(b) * 32 + lane]);
    D[lane * L::DW + q] = (float)d;
    FL[q * 32 + lane] = (uint8_t)(((d < p.dcoll) ? 1 : 0) | ((d < p.min_dist_thresh) ? 2 : 0));
  });
  if (W == 0) D[lane * L::DW + L::DZERO] = 0.0f;
}

// Phase: per-env gather table for the outputs (items it = W, W + 4, ...).
template <int N, int O, int W>
__device__ __forceinline__ void tile_fill_table(const float* __restrict__ S, float* __restrict__ T, int lane) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  float* t = T + lane * L::TW;
  static_for<W, E + N + 1, TILE_WARPS>([&](auto ic) {
    constexpr int it = decltype(ic)::value;
    if constexpr (it < E) {                         // position of entity `it`
      t[L::TP + 2 * it] = S[row_x<N, O>(it) * 32 + lane];
      t[L::TP + 2 * it + 1] = S[row_y<N, O>(it) * 32 + lane];
    } else if constexpr (it < E + N) {              // velocity, goal (landmark goal_match[i]) and fairness obs of agent i
      constexpr int i = it - E;
      t[L::TV + 2 * i] = S[(L::VX + i) * 32 + lane];
      t[L::TV + 2 * i + 1] = S[(L::VY + i) * 32 + lane];
      const int g = __float_as_int(S[(L::GM + i) * 32 + lane]);
      t[L::TG + 2 * i] = S[(L::LX + g) * 32 + lane];
      t[L::TG + 2 * i + 1] = S[(L::LY + g) * 32 + lane];
      t[L::TF + i] = S[(L::FOBS + i) * 32 + lane];
    } else {
      t[L::TZERO] = 0.0f; t[L::TONE] = 1.0f; t[L::TTWO] = 2.0f;
    }
  });
}

// Gather descriptors of this lane, fetched once at kernel entry (registers): word w = lane + 32 k.
template <int N, int O>
struct LaneLuts {
  using L = TileLayout<N, O>;
  static constexpr int KN = (L::NODE_W + 31) / 32, KA = (L::ADJ_W + 31) / 32, KO = (L::OBS_W + 31) / 32;
  uint32_t node[KN], adj[KA], obs[KO];
  __device__ __forceinline__ void load(const TileLuts& luts, int lane) {
#pragma unroll
    for (int k = 0; k < KN; ++k) node[k] = (lane + 32 * k < L::NODE_W) ? __ldg(luts.node + lane + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < KA; ++k) adj[k] = (lane + 32 * k < L::ADJ_W) ? __ldg(luts.adj + lane + 32 * k) : 0u;
#pragma unroll
    for (int k = 0; k < KO; ++k) obs[k] = (lane + 32 * k < L::OBS_W) ? __ldg(luts.obs + lane + 32 * k) : 0u;
  }
};

// Phase: emission.  out[(env0 + el) * W + w] = T[el][s1(w)] - T[el][s2(w)]; a warp emits TILE_EPW envs
// that sit at compile-time offsets from each other (immediate offsets for loads and stores).
template <int TW, int WORDS, int K>
__device__ __forceinline__ void tile_gather_sub(float* __restrict__ out, const uint32_t (&lut)[K],
                                                const float* __restrict__ T, int env0, int nenv, int lane, int warp) {
  const int el0 = warp * TILE_EPW;
  if (el0 >= nenv) return;
  float* o = out + (size_t)(env0 + el0) * WORDS + lane;
  const float* t = T + el0 * TW;
  const int ne = min(TILE_EPW, nenv - el0);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (lane + 32 * k < WORDS) {
      const int s1 = (int)(lut[k] & 0xffffu), s2 = (int)(lut[k] >> 16);
      if (ne == TILE_EPW) {
        float v[TILE_EPW];
#pragma unroll
        for (int e = 0; e < TILE_EPW; ++e) v[e] = t[e * TW + s1] - t[e * TW + s2];
#pragma unroll
        for (int e = 0; e < TILE_EPW; ++e) __stcs(o + e * WORDS + 32 * k, v[e]);
      } else {
        for (int e = 0; e < ne; ++e) __stcs(o + e * WORDS + 32 * k, t[e * TW + s1] - t[e * TW + s2]);
      }
    }
  }
}

template <int DW, int WORDS, int K>
__device__ __forceinline__ void tile_gather_adj(float* __restrict__ out, const uint32_t (&lut)[K],
                                                const float* __restrict__ D, int env0, int nenv, int lane, int warp) {
  const int el0 = warp * TILE_EPW;
  if (el0 >= nenv) return;
  float* o = out + (size_t)(env0 + el0) * WORDS + lane;
  const float* d = D + el0 * DW;
  const int ne = min(TILE_EPW, nenv - el0);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if (lane + 32 * k < WORDS) {
      const int s = (int)lut[k];
      if (ne == TILE_EPW) {
        float v[TILE_EPW];
#pragma unroll
        for (int e = 0; e < TILE_EPW; ++e) v[e] = d[e * DW + s];
#pragma unroll
        for (int e = 0; e < TILE_EPW; ++e) __stcs(o + e * WORDS + 32 * k, v[e]);
      } else {
        for (int e = 0; e < ne; ++e) __stcs(o + e * WORDS + 32 * k, d[e * DW + s]);
      }
    }
  }
}

template <int N, int O>
__device__ __forceinline__ void tile_emit(const DevParams& p, const LaneLuts<N, O>& ll, const float* __restrict__ T,
                                          const float* __restrict__ D, int env0, int nenv, int lane, int warp) {
  using L = TileLayout<N, O>;
  if (p.o_node) tile_gather_sub<L::TW, L::NODE_W>(p.o_node, ll.node, T, env0, nenv, lane, warp);
  if (p.o_adj) tile_gather_adj<L::DW, L::ADJ_W>(p.o_adj, ll.adj, D, env0, nenv, lane, warp);
  if (p.o_obs) tile_gather_sub<L::TW, L::OBS_W>(p.o_obs, ll.obs, T, env0, nenv, lane, warp);
}

// Randomised reset of env `lane` (navigation_graph.py:212-262, :264-570) + lexifair (:555-561), by
// one thread, with the S rows of the env as dynamically indexable storage.  Same Philox stream,
// draw order and acceptance rules as reset_group<G> (fm_device.cuh).
template <int N, int O>
__device__ __forceinline__ void tile_reset_env(const DevParams& p, long long genv, uint32_t episode, float* __restrict__ S,
                                               int lane) {
  using L = TileLayout<N, O>;
#pragma unroll 1
  for (int k = 0; k < O; ++k) {            // obstacles: 0.8 * U(-ws/2, ws/2)^2, draws 0..O-1 (:271-275)
    float x, y;
    draw_uniform2(p, genv, episode, (uint32_t)k, x, y);
    S[(L::OX + k) * 32 + lane] = __fmul_rn(0.8f, x);
    S[(L::OY + k) * 32 + lane] = __fmul_rn(0.8f, y);
  }
  uint32_t d = (uint32_t)O;
#pragma unroll 1
  for (int slot = 0; slot < 2 * N; ++slot) {   // agents (:389-456) then goals (:472-535); entity index == slot
    const bool goal = slot >= N;
    const int base = goal ? N : 0;
    float x, y;
    while (true) {
      draw_uniform2(p, genv, episode, d, x, y);
      ++d;
      if (goal) { x = __fmul_rn(0.8f, x); y = __fmul_rn(0.8f, y); }
      bool bad = false;
#pragma unroll 1
      for (int k = 0; k < O; ++k)
        bad = bad || (dist64(S[(L::OX + k) * 32 + lane], S[(L::OY + k) * 32 + lane], x, y) < p.dcoll);
#pragma unroll 1
      for (int j = base; j < slot; ++j)
        bad = bad || (dist64(S[row_x<N, O>(j) * 32 + lane], S[row_y<N, O>(j) * 32 + lane], x, y) < p.dcoll);
      if (!bad || d >= (uint32_t)MAX_DRAWS) break;
    }
    S[row_x<N, O>(slot) * 32 + lane] = x;
    S[row_y<N, O>(slot) * 32 + lane] = y;
  }
  double cost[N * N];
  int gm[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float ax = S[(L::PX + i) * 32 + lane], ay = S[(L::PY + i) * 32 + lane];
    if (p.has_max_speed) {                 // min_time with the PREVIOUS goal_match (:545-547, :719-728)
      const int og = __float_as_int(S[(L::GM + i) * 32 + lane]);
      S[(L::MINT + i) * 32 + lane] = (float)(dist64(ax, ay, S[(L::LX + og) * 32 + lane], S[(L::LY + og) * 32 + lane]) / p.max_speed);
    }
#pragma unroll
    for (int j = 0; j < N; ++j)             // costs = cdist(agent_pos, goal_pos) (:555)
      cost[i * N + j] = dist64(ax, ay, S[(L::LX + j) * 32 + lane], S[(L::LY + j) * 32 + lane]);
  }
  lexifair_small<N>(cost, gm);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    S[(L::GM + i) * 32 + lane] = __int_as_float(gm[i]);
    S[(L::VX + i) * 32 + lane] = 0.f; S[(L::VY + i) * 32 + lane] = 0.f; S[(L::PD + i) * 32 + lane] = 0.f;
    S[(L::DTG + i) * 32 + lane] = -1.f; S[(L::TREQ + i) * 32 + lane] = -1.f; S[(L::DLEFT + i) * 32 + lane] = -1.f;
    S[(L::NAC + i) * 32 + lane] = __int_as_float(0); S[(L::NOC + i) * 32 + lane] = __int_as_float(0);
    S[(L::FOBS + i) * 32 + lane] = 0.f;     // mean(p_dist = 0) / (std + 1e-4)
  }
  S[L::STEP * 32 + lane] = __int_as_float(0);
  S[L::EPIS * 32 + lane] = __int_as_float((int)(episode + 1));
}

// =============================================================================================
// Body of the kernels for warp role W (the warp's index in the CTA): every phase handles the items
// W, W + 4, W + 8, ... of that phase with compile-time indices.  All warps execute the same sequence
// of barriers.
//   MODE 0: fused env step (MultiAgentGraphEnv.step, environment.py:816-877, + graphworker auto-reset,
//           env_wrappers.py:859-865).   MODE 1: masked reset + observe (environment.py:882-898).
template <int N, int O, int MODE, int W>
__device__ __forceinline__ void tile_role(const DevParams& p, const TileLuts& luts, float* __restrict__ smem) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  float* S = smem + L::OFF_S;
  double* SD = reinterpret_cast<double*>(smem + L::OFF_SD);
  float* T = smem + L::OFF_T;
  float* D = smem + L::OFF_D;
  uint8_t* FL = reinterpret_cast<uint8_t*>(smem + L::OFF_FL);
  const int lane = threadIdx.x & 31;
  const int env0 = blockIdx.x * 32;
  const int nenv = min(32, p.B - env0);
  const int env = env0 + lane;                   // < Bp: the state block is padded to a multiple of 32 envs
  const bool venv = lane < nenv;
  const size_t Bp = (size_t)p.Bp;
  float* gstate = p.px + env;                    // [NROWS][Bp], rows in TileLayout order
  const long long genv = p.env_offset + env;

  // ---- A: state block -> S (one coalesced line per row; all loads in flight before the first store),
  // gather descriptors -> registers, action decode ------------------------------------------------
  {
    constexpr int NR = (L::NROWS - W + TILE_WARPS - 1) / TILE_WARPS;
    float tmp[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) tmp[k] = __ldcg(gstate + (size_t)(W + TILE_WARPS * k) * Bp);
    if (MODE == 0) {
      static_for<W, N, TILE_WARPS>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        float ux = 0.f, uy = 0.f;
        if (venv) {                              // environment.py:301-311: u = [a1 - a2, a3 - a4] * sensitivity (5.0)
          if (p.act_idx) {
            const int a = __ldg(p.act_idx + (size_t)env * N + i);
            ux = ((a == 1) ? 1.f : 0.f) - ((a == 2) ? 1.f : 0.f);
            uy = ((a == 3) ? 1.f : 0.f) - ((a == 4) ? 1.f : 0.f);
          } else {
            const float* oh = p.act_onehot + ((size_t)env * N + i) * 5;
            ux = __ldg(oh + 1) - __ldg(oh + 2);
            uy = __ldg(oh + 3) - __ldg(oh + 4);
          }
          ux *= 5.0f; uy *= 5.0f;
        }
        S[(L::UX + i) * 32 + lane] = ux; S[(L::UY + i) * 32 + lane] = uy;
      });
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) S[(W + TILE_WARPS * k) * 32 + lane] = tmp[k];
  }
  LaneLuts<N, O> ll;
  ll.load(luts, lane);
  __syncthreads();

  if (MODE == 1) {
    // ---- reset() / observe --------------------------------------------------------------------
    const bool do_reset = venv && (p.reset_mask ? (p.reset_mask[env] != 0) : true);
    const bool any_reset = __syncthreads_or(do_reset) != 0;
    if (W == 0) {
      S[L::RFLAG * 32 + lane] = __int_as_float(do_reset ? 1 : 0);
      if (do_reset) {
        tile_reset_env<N, O>(p, genv, (uint32_t)__float_as_int(S[L::EPIS * 32 + lane]), S, lane);
      } else {
        // observation() on the current state (navigation_graph.py:826-857, :849-853)
        double sum_p = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) sum_p += (double)S[(L::PD + j) * 32 + lane];
        const double mean_p = sum_p / N;
        double q_p = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) { const double dd = (double)S[(L::PD + j) * 32 + lane] - mean_p; q_p += dd * dd; }
        const double std_p = sqrt(q_p / N);
        const double dm = (double)S[L::DMEAN * 32 + lane], ds = (double)S[L::DSTD * 32 + lane];
#pragma unroll
        for (int i = 0; i < N; ++i)
          S[(L::FOBS + i) * 32 + lane] =
              (float)((S[(L::DTG + i) * 32 + lane] == -1.0f) ? mean_p / (std_p + 0.0001) : dm / (ds + 0.0001));
      }
    }
    __syncthreads();
    tile_distances<N, O, W>(p, S, D, FL, lane);
    tile_fill_table<N, O, W>(S, T, lane);
    if (any_reset && venv && __float_as_int(S[L::RFLAG * 32 + lane])) {   // rows a reset changes -> state block
      static_for<W, L::NROWS, TILE_WARPS>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        if constexpr (r != L::DMEAN && r != L::DSTD) gstate[(size_t)r * Bp] = S[r * 32 + lane];
      });
    }
    __syncthreads();
    tile_emit<N, O>(p, ll, T, D, env0, nenv, lane, W);
    return;
  }

  // ---- B: force terms (core.py:277-316, :370-404) from the positions at step entry -------------
  // item f < NAA: agent pair (i, j), i < j;  item NAA + i * O + k: agent i vs obstacle k.  fp32 terms as
  // contact_force() (fm_device.cuh); the ordered fp64 accumulation happens in C.
  static_for<W, L::NF, TILE_WARPS>([&](auto fc) {
    constexpr int f = decltype(fc)::value;
    constexpr bool aa = f < L::NAA;
    constexpr int i = aa ? pair_a(f, N) : (f - L::NAA) / (O > 0 ? O : 1);
    constexpr int rbx = aa ? L::PX + pair_b(aa ? f : 0, N) : L::OX + (f - L::NAA) - i * O;
    constexpr int rby = aa ? L::PY + pair_b(aa ? f : 0, N) : L::OY + (f - L::NAA) - i * O;
    const float dx = S[(L::PX + i) * 32 + lane] - S[rbx * 32 + lane], dy = S[(L::PY + i) * 32 + lane] - S[rby * 32 + lane];
    const float dist = sqrtf(dx * dx + dy * dy);
    const float pen = softplusf(-(dist - p.dist_min) / p.contact_margin) * p.contact_margin;
    S[(L::FT + 2 * f) * 32 + lane] = p.contact_force * dx / dist * pen;
    S[(L::FT + 2 * f + 1) * 32 + lane] = p.contact_force * dy / dist * pen;
  });
  __syncthreads();

  // ---- C: ordered force sum + integrate_state (core.py:338-356), float64; state rounded to fp32 ---
  static_for<W, N, TILE_WARPS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    double Fx = (double)S[(L::UX + i) * 32 + lane], Fy = (double)S[(L::UY + i) * 32 + lane];   // mass(1.0) * u + noise(0.0)
    static_for<0, N>([&](auto jc) {              // partners in ascending entity index (core.py:311-316)
      constexpr int j = decltype(jc)::value;
      if constexpr (j != i) {
        constexpr int f = tile_pair_index(j < i ? j : i, j < i ? i : j, N);
        float tx = S[(L::FT + 2 * f) * 32 + lane], ty = S[(L::FT + 2 * f + 1) * 32 + lane];
        if (j < i) { tx = -tx; ty = -ty; }       // agent i is the `b` of pair (j, i): -force
        Fx = (double)tx + Fx; Fy = (double)ty + Fy;
      }
    });
#pragma unroll
    for (int k = 0; k < O; ++k) {
      Fx = (double)S[(L::FT + 2 * (L::NAA + i * O + k)) * 32 + lane] + Fx;
      Fy = (double)S[(L::FT + 2 * (L::NAA + i * O + k) + 1) * 32 + lane] + Fy;
    }
    double v64x = (double)S[(L::VX + i) * 32 + lane] * p.damping_keep + Fx * p.dt;
    double v64y = (double)S[(L::VY + i) * 32 + lane] * p.damping_keep + Fy * p.dt;
    if (p.has_max_speed) {
      const double speed = sqrt(v64x * v64x + v64y * v64y);
      if (speed > p.max_speed) { v64x = v64x / speed * p.max_speed; v64y = v64y / speed * p.max_speed; }
    }
    const double sx = v64x * p.dt, sy = v64y * p.dt;
    const double pd64 = (double)S[(L::PD + i) * 32 + lane] + sqrt(sx * sx + sy * sy);
    SD[(L::PD64 + i) * 32 + lane] = pd64;
    S[(L::PX + i) * 32 + lane] = (float)((double)S[(L::PX + i) * 32 + lane] + sx);
    S[(L::PY + i) * 32 + lane] = (float)((double)S[(L::PY + i) * 32 + lane] + sy);
    S[(L::VX + i) * 32 + lane] = (float)v64x; S[(L::VY + i) * 32 + lane] = (float)v64y;
    S[(L::PD + i) * 32 + lane] = (float)pd64;
  });
  __syncthreads();

  // ---- D: calculate_distances (core.py:204-228) at the new positions --------------------------
  tile_distances<N, O, W>(p, S, D, FL, lane);

  // ---- E: statistic sets.  k = 0: mean / std of the new travelled distances; k = 1..N: mean / std of
  // world.dists_to_goal as left by agent k-1's info_callback, i.e. over [new_0..new_{k-1}, prev_k..]
  // (navigation_graph.py:587-598, :617-618).  Independent of D, so no barrier in between.
  static_for<W, N + 1, TILE_WARPS>([&](auto kc) {
    constexpr int k = decltype(kc)::value;
    double v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double pj = SD[(L::PD64 + j) * 32 + lane];
      if (k == 0) v[j] = pj;
      else if (j < k) v[j] = (S[(L::TREQ + j) * 32 + lane] != -1.0f) ? (double)S[(L::DTG + j) * 32 + lane] : pj;
      else v[j] = (double)S[(L::DTG + j) * 32 + lane];
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s += v[j];
    const double m = s / N;
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) { const double dd = v[j] - m; q += dd * dd; }
    SD[(L::VM + k) * 32 + lane] = m;
    SD[(L::VS + k) * 32 + lane] = sqrt(q / N);
  });
  __syncthreads();

  // ---- F: per-agent observation scalar, reward, latches (environment.py:832-864) ----------------
  const int nstep = __float_as_int(S[L::STEP * 32 + lane]) + 1;      // environment.py:819, :823
  const bool done = nstep >= p.episode_length;   // environment.py:237-247 (agent.status is never set)
  const bool do_reset = venv && done && (p.auto_reset != 0);
  const bool want_info = venv && (p.o_info != nullptr || p.stats != nullptr) && (done || p.info_every_step);
  double* stats_row = p.stats ? p.stats + (size_t)blockIdx.x * (15 * N + 2) : nullptr;
  static_for<W, N, TILE_WARPS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    const float dtg = S[(L::DTG + i) * 32 + lane], treq = S[(L::TREQ + i) * 32 + lane];
    const int gmi = __float_as_int(S[(L::GM + i) * 32 + lane]);
    const int qg = tile_pair_index(i, N, E) + gmi;                   // pair (i, N + gm)
    const float dgoal_f = D[lane * L::DW + qg];                      // (float)dgoal
    const bool reached = (FL[qg * 32 + lane] & 2) != 0;              // dgoal < min_dist_thresh (float64 compare)
    int ncoll = 0;
    static_for<0, N>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if constexpr (j != i) ncoll += FL[tile_pair_index(j < i ? j : i, j < i ? i : j, E) * 32 + lane] & 1;
    });
    bool ocoll = false;
#pragma unroll
    for (int k = 0; k < O; ++k) ocoll = ocoll || ((FL[tile_pair_index(i, 2 * N + k, E) * 32 + lane] & 1) != 0);
    const bool latched = treq != -1.0f;
    const double pd64 = SD[(L::PD64 + i) * 32 + lane];
    const double dtg_new = latched ? (double)dtg : pd64;
    const double treq_new = (!latched && reached) ? (double)nstep * p.dt : (double)treq;   // :588
    const float dleft_new = latched ? S[(L::DLEFT + i) * 32 + lane] : dgoal_f;
    double fparam;                               // navigation_graph.py:764-769 / :849-853
    if (dtg == -1.0f) fparam = SD[(L::VM + 0) * 32 + lane] / (SD[(L::VS + 0) * 32 + lane] + 0.0001);
    else if (i == 0) fparam = (double)S[L::DMEAN * 32 + lane] / ((double)S[L::DSTD * 32 + lane] + 0.0001);
    else fparam = SD[(L::VM + i) * 32 + lane] / (SD[(L::VS + i) * 32 + lane] + 0.0001);
    float rw = reached ? p.goal_rew : -dgoal_f;  // navigation_graph.py:760-824
    rw -= p.coll_rew * (float)ncoll;
    if (ocoll) rw -= p.coll_rew;
    if (p.fairness_reward) {
      float fair = p.fair_rew * tanhf((float)(fparam - p.zeroshift));
      if (fair < -2.0f) fair = -2.0f;
      rw += fair;
    }
    rw = fminf(fmaxf(rw, p.clip_lo), p.clip_hi);
    const int nac = __float_as_int(S[(L::NAC + i) * 32 + lane]) + ncoll;          // :604-613
    const int noc = __float_as_int(S[(L::NOC + i) * 32 + lane]) + (ocoll ? 1 : 0); // :602-603
    S[(L::OWN + i) * 32 + lane] = rw;
    S[(L::FOBS + i) * 32 + lane] = (float)fparam;
    S[(L::DTG + i) * 32 + lane] = (float)dtg_new;
    S[(L::NTREQ + i) * 32 + lane] = (float)treq_new;   // TREQ keeps the old value for the info pass (G)
    S[(L::DLEFT + i) * 32 + lane] = dleft_new;
    S[(L::NAC + i) * 32 + lane] = __int_as_float(nac);
    S[(L::NOC + i) * 32 + lane] = __int_as_float(noc);
    if (i == N - 1) {                            // world.dist_traveled_mean / stddev after the last info_callback
      S[L::NDMEAN * 32 + lane] = (float)SD[(L::VM + N) * 32 + lane];
      S[L::NDSTD * 32 + lane] = (float)SD[(L::VS + N) * 32 + lane];
    }
    if (venv && p.o_done) p.o_done[(size_t)env * N + i] = done ? 1 : 0;
  });
  __syncthreads();

  // ---- G: collaborative sum, reward output, episode statistics, info rows -----------------------
  static_for<W, N, TILE_WARPS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    float rew = S[(L::OWN + i) * 32 + lane];
    if (p.collaborative) {                       // environment.py:866-870
      float tot = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) tot += S[(L::OWN + j) * 32 + lane];
      rew = tot;
    }
    if (venv && p.o_rew) p.o_rew[(size_t)env * N + i] = rew;
    if (stats_row) {
      double v = venv ? (double)rew : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
      if (lane == 0) stats_row[i] += v;
    }
  });
  if (stats_row && W == TILE_WARPS - 1) {
    const unsigned termb = __ballot_sync(FULL, venv && done);
    if (lane == 0) { stats_row[15 * N] += (double)__popc(termb); stats_row[15 * N + 1] += (double)nenv; }
  }
  if (__syncthreads_or(want_info)) {
    static_for<W, N, TILE_WARPS>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      // world-level time statistics right after agent i's own info_callback: new values of agents
      // j <= i, previous values of j > i (navigation_graph.py:620-621)
      double tacc = 0.0;                         // entity.state.time += dt per step (core.py:355)
      for (int k = 0; k < nstep; ++k) tacc += p.dt;
      double tv[N];                              // times_required as float64: a latch of THIS step is nstep * dt unrounded
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float told = S[(L::TREQ + j) * 32 + lane];
        const bool fresh = j <= i && told == -1.0f && S[(L::NTREQ + j) * 32 + lane] != -1.0f;
        tv[j] = fresh ? (double)nstep * p.dt : (double)told;
      }
      double st = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) st += tv[j];
      const double mt = st / N;
      double qt = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) { const double dd = tv[j] - mt; qt += dd * dd; }
      const double stv = sqrt(qt / N);
      const double md = SD[(L::VM + i + 1) * 32 + lane], sdv = SD[(L::VS + i + 1) * 32 + lane];
      float info[INFO_F];
      info[0] = S[(L::OWN + i) * 32 + lane]; info[1] = S[(L::DLEFT + i) * 32 + lane]; info[2] = S[(L::NTREQ + i) * 32 + lane];
      info[3] = (float)__float_as_int(S[(L::NAC + i) * 32 + lane]); info[4] = (float)__float_as_int(S[(L::NOC + i) * 32 + lane]);
      info[5] = (float)md; info[6] = (float)sdv; info[7] = (float)(md / (sdv + 0.0001));
      info[8] = S[(L::DTG + i) * 32 + lane]; info[9] = (float)tacc; info[10] = (float)mt; info[11] = (float)stv;
      info[12] = (float)(mt / (stv + 0.0001)); info[13] = S[(L::MINT + i) * 32 + lane];
      if (want_info && p.o_info) {
        float* o = p.o_info + ((size_t)env * N + i) * INFO_F;
#pragma unroll
        for (int k = 0; k < INFO_F; ++k) o[k] = info[k];
      }
      if (stats_row && __any_sync(FULL, venv && done)) {
#pragma unroll
        for (int k = 0; k < INFO_F; ++k) {
          double v = (venv && done) ? (double)info[k] : 0.0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
          if (lane == 0) stats_row[N + i * INFO_F + k] += v;
        }
      }
    });
  }

  // ---- auto-reset (env_wrappers.py:859-865): obs / node_obs / adj come from the new episode, reward /
  // done / info stay terminal --------------------------------------------------------------------
  const bool any_reset = __syncthreads_or(do_reset) != 0;
  if (W == 0) {
    S[L::RFLAG * 32 + lane] = __int_as_float(do_reset ? 1 : 0);
    if (!do_reset) S[L::STEP * 32 + lane] = __int_as_float(nstep);
#pragma unroll
    for (int i = 0; i < N; ++i) S[(L::TREQ + i) * 32 + lane] = S[(L::NTREQ + i) * 32 + lane];
    S[L::DMEAN * 32 + lane] = S[L::NDMEAN * 32 + lane];
    S[L::DSTD * 32 + lane] = S[L::NDSTD * 32 + lane];
    if (do_reset) tile_reset_env<N, O>(p, genv, (uint32_t)__float_as_int(S[L::EPIS * 32 + lane]), S, lane);
  }
  __syncthreads();
  if (any_reset) tile_distances<N, O, W>(p, S, D, FL, lane);
  tile_fill_table<N, O, W>(S, T, lane);
  // ---- I: state block write-back (rows that change every step; the rest only for envs that reset) ----
  if (venv) {
    const bool was_reset = any_reset && __float_as_int(S[L::RFLAG * 32 + lane]);
    static_for<W, L::NROWS, TILE_WARPS>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      constexpr bool every_step = r < L::MINT || (r >= L::NAC && r < L::LX) || r == L::DMEAN || r == L::DSTD || r == L::STEP;
      if (every_step || was_reset) gstate[(size_t)r * Bp] = S[r * 32 + lane];
    });
  }
  __syncthreads();
  tile_emit<N, O>(p, ll, T, D, env0, nenv, lane, W);
}

template <int N, int O, int MODE>
__global__ void __launch_bounds__(TILE_THREADS) tile_kernel(const __grid_constant__ DevParams p, const TileLuts luts) {
  extern __shared__ __align__(16) float smem[];
  static_assert(TILE_WARPS == 4, "role dispatch below is written for 4 warps");
  switch (threadIdx.x >> 5) {                    // warp role: compile-time item indices per role
    case 0: tile_role<N, O, MODE, 0>(p, luts, smem); break;
    case 1: tile_role<N, O, MODE, 1>(p, luts, smem); break;
    case 2: tile_role<N, O, MODE, 2>(p, luts, smem); break;
    default: tile_role<N, O, MODE, 3>(p, luts, smem); break;
  }
}

// =============================================================================================
// Host side: gather tables, launch.
template <int N, int O>
static void tile_build_luts_no(std::vector<uint32_t>& obs, std::vector<uint32_t>& node, std::vector<uint32_t>& adj) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  auto pk = [](int s1, int s2) { return (uint32_t)s1 | ((uint32_t)s2 << 16); };
  const int Z = L::TZERO;
  obs.assign(L::OBS_W, 0); node.assign(L::NODE_W, 0); adj.assign(L::ADJ_W, 0);
  for (int i = 0; i < N; ++i) {             // navigation_graph.py:826-857: [vel, pos, goal - pos, fairness_param]
    uint32_t* o = &obs[i * OBS_F];
    o[0] = pk(L::TV + 2 * i, Z); o[1] = pk(L::TV + 2 * i + 1, Z);
    o[2] = pk(L::TP + 2 * i, Z); o[3] = pk(L::TP + 2 * i + 1, Z);
    o[4] = pk(L::TG + 2 * i, L::TP + 2 * i); o[5] = pk(L::TG + 2 * i + 1, L::TP + 2 * i + 1);
    o[6] = pk(L::TF + i, Z);
  }
  for (int a = 0; a < N; ++a)               // navigation_graph.py:1079-1124 (relative features)
    for (int e = 0; e < E; ++e) {
      uint32_t* o = &node[(a * E + e) * NODE_F];
      const int pa = L::TP + 2 * a, va = L::TV + 2 * a, pe = L::TP + 2 * e;
      const int vex = e < N ? L::TV + 2 * e : Z, vey = e < N ? L::TV + 2 * e + 1 : Z;
      const int gex = e < N ? L::TG + 2 * e : pe, gey = e < N ? L::TG + 2 * e + 1 : pe + 1;
      o[0] = pk(vex, va); o[1] = pk(vey, va + 1);
      o[2] = pk(pe, pa); o[3] = pk(pe + 1, pa + 1);
      o[4] = pk(gex, pa); o[5] = pk(gey, pa + 1);
      o[6] = pk(pe, pa); o[7] = pk(pe + 1, pa + 1); o[8] = pk(pe, pa); o[9] = pk(pe + 1, pa + 1);
      o[10] = pk(e < N ? Z : (e < 2 * N ? L::TONE : L::TTWO), Z);
    }
  for (int x = 0; x < E; ++x)
    for (int y = 0; y < E; ++y)
      adj[x * E + y] = (x == y) ? (uint32_t)L::DZERO : (uint32_t)tile_pair_index(x < y ? x : y, x < y ? y : x, E);
}

template <int N, int O>
static cudaError_t tile_launch_no(const DevParams& p, const TileLuts& luts, cudaStream_t st, bool is_reset) {
  using L = TileLayout<N, O>;
  const int blocks = (p.B + 31) / 32;
  const size_t smem = (size_t)L::WORDS * sizeof(float);
  if (is_reset) tile_kernel<N, O, 1><<<blocks, TILE_THREADS, smem, st>>>(p, luts);
  else tile_kernel<N, O, 0><<<blocks, TILE_THREADS, smem, st>>>(p, luts);
  return cudaGetLastError();
}

template <int N, int O>
static cudaError_t tile_prepare_no() {
  using L = TileLayout<N, O>;
  const int smem = L::WORDS * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(tile_kernel<N, O, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(tile_kernel<N, O, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// The (N, O) pairs compiled for this mapping.  Everything else runs the group-per-env kernels.
#define FM_TILE_CASES(X) X(1, 1) X(2, 0) X(3, 0) X(3, 3) X(4, 2)

bool tile_supported(int N, int O) {
#define X(n, o) if (N == n && O == o) return true;
  FM_TILE_CASES(X)
#undef X
  return false;
}

int tile_num_ctas(int B) { return (B + 31) / 32; }

cudaError_t tile_prepare(const DevParams& p) {
#define X(n, o) if (p.N == n && p.O == o) return tile_prepare_no<n, o>();
  FM_TILE_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

void tile_build_luts(int N, int O, std::vector<uint32_t>& obs, std::vector<uint32_t>& node, std::vector<uint32_t>& adj) {
#define X(n, o) if (N == n && O == o) return tile_build_luts_no<n, o>(obs, node, adj);
  FM_TILE_CASES(X)
#undef X
}

cudaError_t tile_launch(const DevParams& p, cudaStream_t st, bool is_reset) {
  TileLuts luts{p.lut_obs, p.lut_node, p.lut_adj};
#define X(n, o) if (p.N == n && p.O == o) return tile_launch_no<n, o>(p, luts, st, is_reset);
  FM_TILE_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace fm
