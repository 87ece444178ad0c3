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
  static constexpr int OBS_W = N * OBS_F, NODE_W = N * E * NODE_F, ADJ_W = E * E;
  static constexpr int LUT_W = OBS_W + NODE_W + ADJ_W;  // gather descriptors: obs | node | adj
  static constexpr int OFF_LUT = OFF_FL + PAIRS * 8;   // FL: PAIRS * 32 bytes
  static constexpr int WORDS = OFF_LUT + LUT_W;
};

struct TileLuts {
  const uint32_t* all;     // obs [7N] | node [11 N E]: s1 | s2 << 16 into T;  adj [E E]: index into D
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
// Item descriptors, passed by value in the kernel-parameter constant bank: shared-memory word offsets
// (row * 32) so that an item needs no index arithmetic beyond `offset + lane`.
template <int N, int O>
struct TileTables {
  using L = TileLayout<N, O>;
  uint16_t pair[L::PAIRS][4];                    // x_a, y_a, x_b, y_b rows of distance pair q (a < b, row-major)
  uint16_t force[L::NF > 0 ? L::NF : 1][4];      // x_i, y_i, x_partner, y_partner rows of force item f
  uint16_t ent[L::E][2];                         // x, y rows of entity e
  uint8_t every_step[L::NROWS];                  // 1: state row is written back every step; 0: only after a reset
};

// pair q (row-major over a < b) -> a, b
__host__ __device__ constexpr int pair_a(int q, int E) { int a = 0; while (q >= E - 1 - a) { q -= E - 1 - a; ++a; } return a; }
__host__ __device__ constexpr int pair_b(int q, int E) { int a = 0; while (q >= E - 1 - a) { q -= E - 1 - a; ++a; } return a + 1 + q; }

// Phase: distances of the entity pairs q = warp, warp + 4, ... at the current positions
// (core.py:204-228) + predicate bits.
template <int N, int O>
__device__ __forceinline__ void tile_distances(const DevParams& p, const TileTables<N, O>& tb, const float* __restrict__ S,
                                               float* __restrict__ D, uint8_t* __restrict__ FL, int lane, int warp) {
  using L = TileLayout<N, O>;
  const float* Sl = S + lane;
  float* Dq = D + lane * L::DW + warp;
  uint8_t* Fq = FL + warp * 32 + lane;
#pragma unroll 3
  for (int k = 0; k < (L::PAIRS + TILE_WARPS - 1) / TILE_WARPS; ++k) {
    const int q = warp + TILE_WARPS * k;
    if (q < L::PAIRS) {
      const uint2 o = *reinterpret_cast<const uint2*>(tb.pair[q]);
      const double d = dist64(Sl[o.x & 0xffffu], Sl[o.x >> 16], Sl[o.y & 0xffffu], Sl[o.y >> 16]);
      Dq[TILE_WARPS * k] = (float)d;
      Fq[TILE_WARPS * 32 * k] = (uint8_t)(((d < p.dcoll) ? 1 : 0) | ((d < p.min_dist_thresh) ? 2 : 0));
    }
  }
  if (warp == 0) D[lane * L::DW + L::DZERO] = 0.0f;
}

// Phase: per-env gather table for the outputs (items it = warp, warp + 4, ...).
template <int N, int O>
__device__ __forceinline__ void tile_fill_table(const TileTables<N, O>& tb, const float* __restrict__ S, float* __restrict__ T,
                                                int lane, int warp) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  float* t = T + lane * L::TW;
  for (int it = warp; it < E + N + 1; it += TILE_WARPS) {
    if (it < E) {                                   // position of entity `it`
      const uint32_t o = *reinterpret_cast<const uint32_t*>(tb.ent[it]);
      t[L::TP + 2 * it] = S[(o & 0xffffu) + lane];
      t[L::TP + 2 * it + 1] = S[(o >> 16) + lane];
    } else if (it < E + N) {                        // velocity, goal (landmark goal_match[i]) and fairness obs of agent i
      const int i = it - E;
      const float* Si = S + i * 32 + lane;
      t[L::TV + 2 * i] = Si[L::VX * 32];
      t[L::TV + 2 * i + 1] = Si[L::VY * 32];
      const int g = __float_as_int(Si[L::GM * 32]);
      t[L::TG + 2 * i] = S[(L::LX + g) * 32 + lane];
      t[L::TG + 2 * i + 1] = S[(L::LY + g) * 32 + lane];
      t[L::TF + i] = Si[L::FOBS * 32];
    } else {
      t[L::TZERO] = 0.0f; t[L::TONE] = 1.0f; t[L::TTWO] = 2.0f;
    }
  }
}

// Phase: emission.  out[(env0 + el) * WORDS + w] = T[el][s1(w)] - T[el][s2(w)] (SUB) or D[el][s(w)]; a warp
// emits TILE_EPW envs that sit at compile-time offsets from each other (immediate offsets for loads and
// stores); lane = output word, so every store instruction writes 128 contiguous bytes.
template <int STRIDE, int WORDS, bool SUB>
__device__ __forceinline__ void tile_gather(float* __restrict__ out, const uint32_t* __restrict__ lut,
                                            const float* __restrict__ tab, int env0, int nenv, int lane, int warp) {
  const int el0 = warp * TILE_EPW;
  if (el0 >= nenv) return;
  float* o = out + (size_t)(env0 + el0) * WORDS + lane;
  const float* t = tab + el0 * STRIDE;
  const int ne = min(TILE_EPW, nenv - el0);
  uint32_t u = lut[lane < WORDS ? lane : 0];
#pragma unroll 1
  for (int w = lane; w < WORDS; w += 32) {
    const int s1 = (int)(u & 0xffffu), s2 = (int)(u >> 16);
    u = lut[w + 32 < WORDS ? w + 32 : 0];          // descriptor of the next iteration
    if (ne == TILE_EPW) {
      float v[TILE_EPW];
#pragma unroll
      for (int e = 0; e < TILE_EPW; ++e) v[e] = SUB ? t[e * STRIDE + s1] - t[e * STRIDE + s2] : t[e * STRIDE + s1];
#pragma unroll
      for (int e = 0; e < TILE_EPW; ++e) __stcs(o + e * WORDS, v[e]);
    } else {
      for (int e = 0; e < ne; ++e) __stcs(o + e * WORDS, SUB ? t[e * STRIDE + s1] - t[e * STRIDE + s2] : t[e * STRIDE + s1]);
    }
    o += 32;
  }
}

template <int N, int O>
__device__ __forceinline__ void tile_emit(const DevParams& p, const uint32_t* __restrict__ LUT, const float* __restrict__ T,
                                          const float* __restrict__ D, int env0, int nenv, int lane, int warp) {
  using L = TileLayout<N, O>;
  if (p.o_node) tile_gather<L::TW, L::NODE_W, true>(p.o_node, LUT + L::OBS_W, T, env0, nenv, lane, warp);
  if (p.o_adj) tile_gather<L::DW, L::ADJ_W, false>(p.o_adj, LUT + L::OBS_W + L::NODE_W, D, env0, nenv, lane, warp);
  if (p.o_obs) tile_gather<L::TW, L::OBS_W, true>(p.o_obs, LUT, T, env0, nenv, lane, warp);
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
// One code path for all warps (the binary stays ~4 k instructions: a role-specialised variant with
// compile-time items per warp executed 40 % fewer instructions but was instruction-fetch bound).
//   MODE 0: fused env step (MultiAgentGraphEnv.step, environment.py:816-877, + graphworker auto-reset,
//           env_wrappers.py:859-865).   MODE 1: masked reset + observe (environment.py:882-898).
template <int N, int O, int MODE>
__global__ void __launch_bounds__(TILE_THREADS) tile_kernel(const __grid_constant__ DevParams p, const TileLuts luts,
                                                            const __grid_constant__ TileTables<N, O> tb) {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  extern __shared__ __align__(16) float smem[];
  float* S = smem + L::OFF_S;
  double* SD = reinterpret_cast<double*>(smem + L::OFF_SD);
  float* T = smem + L::OFF_T;
  float* D = smem + L::OFF_D;
  uint8_t* FL = reinterpret_cast<uint8_t*>(smem + L::OFF_FL);
  uint32_t* LUT = reinterpret_cast<uint32_t*>(smem + L::OFF_LUT);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int env0 = blockIdx.x * 32;
  const int nenv = min(32, p.B - env0);
  const int env = env0 + lane;                   // < Bp: the state block is padded to a multiple of 32 envs
  const bool venv = lane < nenv;
  const size_t Bp = (size_t)p.Bp;
  float* gstate = p.px + env;                    // [NROWS][Bp], rows in TileLayout order
  const long long genv = p.env_offset + env;
  float* Sl = S + lane;

  // ---- A: state block -> S (one coalesced line per row; all loads in flight before the first store),
  // gather descriptors -> smem, action decode -----------------------------------------------------
  {
    constexpr int NRK = (L::NROWS + TILE_WARPS - 1) / TILE_WARPS;
    float tmp[NRK];
    const float* gw = gstate + (size_t)warp * Bp;
#pragma unroll
    for (int k = 0; k < NRK; ++k)
      if (TILE_WARPS * k + TILE_WARPS <= L::NROWS || warp + TILE_WARPS * k < L::NROWS) tmp[k] = __ldcg(gw + (size_t)(TILE_WARPS * k) * Bp);
    for (int w = threadIdx.x; w < L::LUT_W; w += TILE_THREADS) LUT[w] = __ldg(luts.all + w);
    if (MODE == 0) {
      for (int i = warp; i < N; i += TILE_WARPS) {
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
        Sl[(L::UX + i) * 32] = ux; Sl[(L::UY + i) * 32] = uy;
      }
    }
    float* Sw = Sl + warp * 32;
#pragma unroll
    for (int k = 0; k < NRK; ++k)
      if (TILE_WARPS * k + TILE_WARPS <= L::NROWS || warp + TILE_WARPS * k < L::NROWS) Sw[TILE_WARPS * k * 32] = tmp[k];
  }
  __syncthreads();

  if (MODE == 1) {
    // ---- reset() / observe --------------------------------------------------------------------
    const bool do_reset = venv && (p.reset_mask ? (p.reset_mask[env] != 0) : true);
    const bool any_reset = __syncthreads_or(do_reset) != 0;
    if (warp == 0) {
      Sl[L::RFLAG * 32] = __int_as_float(do_reset ? 1 : 0);
      if (do_reset) {
        tile_reset_env<N, O>(p, genv, (uint32_t)__float_as_int(Sl[L::EPIS * 32]), S, lane);
      } else {
        // observation() on the current state (navigation_graph.py:826-857, :849-853)
        double sum_p = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) sum_p += (double)Sl[(L::PD + j) * 32];
        const double mean_p = sum_p / N;
        double q_p = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j) { const double dd = (double)Sl[(L::PD + j) * 32] - mean_p; q_p = sq_acc(q_p, dd); }
        const double std_p = sqrt(q_p / N);
        const double dm = (double)Sl[L::DMEAN * 32], ds = (double)Sl[L::DSTD * 32];
#pragma unroll
        for (int i = 0; i < N; ++i)
          Sl[(L::FOBS + i) * 32] = (float)((Sl[(L::DTG + i) * 32] == -1.0f) ? mean_p / (std_p + 0.0001) : dm / (ds + 0.0001));
      }
    }
    __syncthreads();
    tile_distances<N, O>(p, tb, S, D, FL, lane, warp);
    tile_fill_table<N, O>(tb, S, T, lane, warp);
    if (any_reset && venv && __float_as_int(Sl[L::RFLAG * 32])) {   // rows a reset changes -> state block
      for (int r = warp; r < L::NROWS; r += TILE_WARPS)
        if (r != L::DMEAN && r != L::DSTD) gstate[(size_t)r * Bp] = Sl[r * 32];
    }
    __syncthreads();
    tile_emit<N, O>(p, LUT, T, D, env0, nenv, lane, warp);
    return;
  }

  // ---- B: force terms (core.py:277-316, :370-404) from the positions at step entry -------------
  // item f < NAA: agent pair (i, j), i < j;  item NAA + i * O + k: agent i vs obstacle k.  fp32 terms as
  // contact_force() (fm_device.cuh); the ordered fp64 accumulation happens in C.
#pragma unroll 3
  for (int k = 0; k < (L::NF + TILE_WARPS - 1) / TILE_WARPS; ++k) {
    const int f = warp + TILE_WARPS * k;
    if (f < L::NF) {
      const uint2 o = *reinterpret_cast<const uint2*>(tb.force[f]);
      const float dx = Sl[o.x & 0xffffu] - Sl[o.y & 0xffffu], dy = Sl[o.x >> 16] - Sl[o.y >> 16];
      const float dist = sqrtf(dx * dx + dy * dy);
      const float pen = softplusf(-(dist - p.dist_min) / p.contact_margin) * p.contact_margin;
      Sl[(L::FT + 2 * f) * 32] = p.contact_force * dx / dist * pen;
      Sl[(L::FT + 2 * f + 1) * 32] = p.contact_force * dy / dist * pen;
    }
  }
  __syncthreads();

  // ---- C: ordered force sum + integrate_state (core.py:338-356), float64; state rounded to fp32 ---
  for (int i = warp; i < N; i += TILE_WARPS) {
    float* Si = Sl + i * 32;
    double Fx = (double)Si[L::UX * 32], Fy = (double)Si[L::UY * 32];   // mass(1.0) * u + noise(0.0)
#pragma unroll
    for (int j = 0; j < N; ++j) {                // partners in ascending entity index (core.py:311-316)
      if (j != i) {
        const int f = j < i ? tile_pair_index(j, i, N) : tile_pair_index(i, j, N);
        float tx = Sl[(L::FT + 2 * f) * 32], ty = Sl[(L::FT + 2 * f + 1) * 32];
        if (j < i) { tx = -tx; ty = -ty; }       // agent i is the `b` of pair (j, i): -force
        Fx = (double)tx + Fx; Fy = (double)ty + Fy;
      }
    }
    const float* Fo = Sl + (L::FT + 2 * (L::NAA + i * O)) * 32;
#pragma unroll
    for (int k = 0; k < O; ++k) { Fx = (double)Fo[2 * k * 32] + Fx; Fy = (double)Fo[(2 * k + 1) * 32] + Fy; }
    double v64x, v64y, sx, sy, pd64;
    integrate64(p, Si[L::VX * 32], Si[L::VY * 32], Fx, Fy, Si[L::PD * 32], v64x, v64y, sx, sy, pd64);
    SD[(L::PD64 + i) * 32 + lane] = pd64;
    Si[L::PX * 32] = (float)__dadd_rn((double)Si[L::PX * 32], sx);
    Si[L::PY * 32] = (float)__dadd_rn((double)Si[L::PY * 32], sy);
    Si[L::VX * 32] = (float)v64x; Si[L::VY * 32] = (float)v64y;
    Si[L::PD * 32] = (float)pd64;
  }
  __syncthreads();

  // ---- D: calculate_distances (core.py:204-228) at the new positions --------------------------
  tile_distances<N, O>(p, tb, S, D, FL, lane, warp);

  // ---- E: statistic sets.  k = 0: mean / std of the new travelled distances; k = 1..N: mean / std of
  // world.dists_to_goal as left by agent k-1's info_callback, i.e. over [new_0..new_{k-1}, prev_k..]
  // (navigation_graph.py:587-598, :617-618).  Independent of D, so no barrier in between.
  for (int k = warp; k <= N; k += TILE_WARPS) {
    double v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double pj = SD[(L::PD64 + j) * 32 + lane];
      const double dj = (double)Sl[(L::DTG + j) * 32];
      const bool latched = Sl[(L::TREQ + j) * 32] != -1.0f;
      v[j] = (k == 0) ? pj : ((j < k && !latched) ? pj : dj);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s += v[j];
    const double m = s / N;
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) { const double dd = v[j] - m; q = sq_acc(q, dd); }
    SD[(L::VM + k) * 32 + lane] = m;
    SD[(L::VS + k) * 32 + lane] = sqrt(q / N);
  }
  __syncthreads();

  // ---- F: per-agent observation scalar, reward, latches (environment.py:832-864) ----------------
  const int nstep = __float_as_int(Sl[L::STEP * 32]) + 1;            // environment.py:819, :823
  const bool done = nstep >= p.episode_length;   // environment.py:237-247 (agent.status is never set)
  const bool do_reset = venv && done && (p.auto_reset != 0);
  const bool want_info = venv && (p.o_info != nullptr || p.stats != nullptr) && (done || p.info_every_step);
  double* stats_row = p.stats ? p.stats + (size_t)blockIdx.x * (15 * N + 2) : nullptr;
  for (int i = warp; i < N; i += TILE_WARPS) {
    float* Si = Sl + i * 32;
    const float dtg = Si[L::DTG * 32], treq = Si[L::TREQ * 32];
    const int gmi = __float_as_int(Si[L::GM * 32]);
    const int qi = tile_pair_index(i, i + 1, E);                     // first pair of row i: (i, i + 1)
    const int qg = qi + (N - i - 1) + gmi;                           // pair (i, N + gm)
    const float dgoal_f = D[lane * L::DW + qg];                      // (float)dgoal
    const bool reached = (FL[qg * 32 + lane] & 2) != 0;              // dgoal < min_dist_thresh (float64 compare)
    int ncoll = 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j != i) ncoll += FL[(j < i ? tile_pair_index(j, i, E) : tile_pair_index(i, j, E)) * 32 + lane] & 1;
    bool ocoll = false;
#pragma unroll
    for (int k = 0; k < O; ++k) ocoll = ocoll || ((FL[(qi + (2 * N - i - 1) + k) * 32 + lane] & 1) != 0);
    const bool latched = treq != -1.0f;
    const double pd64 = SD[(L::PD64 + i) * 32 + lane];
    const double dtg_new = latched ? (double)dtg : pd64;
    const double treq_new = (!latched && reached) ? (double)nstep * p.dt : (double)treq;   // :588
    const float dleft_new = latched ? Si[L::DLEFT * 32] : dgoal_f;
    double fparam;                               // navigation_graph.py:764-769 / :849-853
    if (dtg == -1.0f) fparam = SD[(L::VM + 0) * 32 + lane] / (SD[(L::VS + 0) * 32 + lane] + 0.0001);
    else if (i == 0) fparam = (double)Sl[L::DMEAN * 32] / ((double)Sl[L::DSTD * 32] + 0.0001);
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
    const int nac = __float_as_int(Si[L::NAC * 32]) + ncoll;          // :604-613
    const int noc = __float_as_int(Si[L::NOC * 32]) + (ocoll ? 1 : 0); // :602-603
    Si[L::OWN * 32] = rw;
    Si[L::FOBS * 32] = (float)fparam;
    Si[L::DTG * 32] = (float)dtg_new;
    Si[L::NTREQ * 32] = (float)treq_new;         // TREQ keeps the old value for the info pass (G)
    Si[L::DLEFT * 32] = dleft_new;
    Si[L::NAC * 32] = __int_as_float(nac);
    Si[L::NOC * 32] = __int_as_float(noc);
    if (i == N - 1) {                            // world.dist_traveled_mean / stddev after the last info_callback
      Sl[L::NDMEAN * 32] = (float)SD[(L::VM + N) * 32 + lane];
      Sl[L::NDSTD * 32] = (float)SD[(L::VS + N) * 32 + lane];
    }
    if (venv && p.o_done) p.o_done[(size_t)env * N + i] = done ? 1 : 0;
  }
  __syncthreads();

  // ---- G: collaborative sum, reward output, episode statistics, info rows -----------------------
  for (int i = warp; i < N; i += TILE_WARPS) {
    float rew = Sl[(L::OWN + i) * 32];
    if (p.collaborative) {                       // environment.py:866-870
      float tot = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) tot += Sl[(L::OWN + j) * 32];
      rew = tot;
    }
    if (venv && p.o_rew) p.o_rew[(size_t)env * N + i] = rew;
    if (stats_row) {
      double v = venv ? (double)rew : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
      if (lane == 0) stats_row[i] += v;
    }
  }
  if (stats_row && warp == TILE_WARPS - 1) {
    const unsigned termb = __ballot_sync(FULL, venv && done);
    if (lane == 0) { stats_row[15 * N] += (double)__popc(termb); stats_row[15 * N + 1] += (double)nenv; }
  }
  if (__syncthreads_or(want_info)) {
    for (int i = warp; i < N; i += TILE_WARPS) {
      // world-level time statistics right after agent i's own info_callback: new values of agents
      // j <= i, previous values of j > i (navigation_graph.py:620-621)
      double tacc = 0.0;                         // entity.state.time += dt per step (core.py:355)
      for (int k = 0; k < nstep; ++k) tacc += p.dt;
      double tv[N];                              // times_required as float64: a latch of THIS step is nstep * dt unrounded
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float told = Sl[(L::TREQ + j) * 32];
        const bool fresh = j <= i && told == -1.0f && Sl[(L::NTREQ + j) * 32] != -1.0f;
        tv[j] = fresh ? (double)nstep * p.dt : (double)told;
      }
      double st = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) st += tv[j];
      const double mt = st / N;
      double qt = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) { const double dd = tv[j] - mt; qt = sq_acc(qt, dd); }
      const double stv = sqrt(qt / N);
      const double md = SD[(L::VM + i + 1) * 32 + lane], sdv = SD[(L::VS + i + 1) * 32 + lane];
      const float* Si = Sl + i * 32;
      float info[INFO_F];
      info[0] = Si[L::OWN * 32]; info[1] = Si[L::DLEFT * 32]; info[2] = Si[L::NTREQ * 32];
      info[3] = (float)__float_as_int(Si[L::NAC * 32]); info[4] = (float)__float_as_int(Si[L::NOC * 32]);
      info[5] = (float)md; info[6] = (float)sdv; info[7] = (float)(md / (sdv + 0.0001));
      info[8] = Si[L::DTG * 32]; info[9] = (float)tacc; info[10] = (float)mt; info[11] = (float)stv;
      info[12] = (float)(mt / (stv + 0.0001)); info[13] = Si[L::MINT * 32];
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
    }
  }

  // ---- auto-reset (env_wrappers.py:859-865): obs / node_obs / adj come from the new episode, reward /
  // done / info stay terminal --------------------------------------------------------------------
  const bool any_reset = __syncthreads_or(do_reset) != 0;
  if (warp == 0) {
    Sl[L::RFLAG * 32] = __int_as_float(do_reset ? 1 : 0);
    if (!do_reset) Sl[L::STEP * 32] = __int_as_float(nstep);
#pragma unroll
    for (int i = 0; i < N; ++i) Sl[(L::TREQ + i) * 32] = Sl[(L::NTREQ + i) * 32];
    Sl[L::DMEAN * 32] = Sl[L::NDMEAN * 32];
    Sl[L::DSTD * 32] = Sl[L::NDSTD * 32];
    if (do_reset) tile_reset_env<N, O>(p, genv, (uint32_t)__float_as_int(Sl[L::EPIS * 32]), S, lane);
  }
  __syncthreads();
  if (any_reset) tile_distances<N, O>(p, tb, S, D, FL, lane, warp);
  tile_fill_table<N, O>(tb, S, T, lane, warp);
  // ---- I: state block write-back (rows that change every step; the rest only for envs that reset) ----
  if (venv) {
    const bool was_reset = any_reset && __float_as_int(Sl[L::RFLAG * 32]);
    constexpr int NRK = (L::NROWS + TILE_WARPS - 1) / TILE_WARPS;
    float* gw = gstate + (size_t)warp * Bp;
    const float* Sw = Sl + warp * 32;
#pragma unroll
    for (int k = 0; k < NRK; ++k) {
      const int r = warp + TILE_WARPS * k;
      if ((TILE_WARPS * k + TILE_WARPS <= L::NROWS || r < L::NROWS) && (tb.every_step[r] || was_reset))
        gw[(size_t)(TILE_WARPS * k) * Bp] = Sw[TILE_WARPS * k * 32];
    }
  }
  __syncthreads();
  tile_emit<N, O>(p, LUT, T, D, env0, nenv, lane, warp);
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
static TileTables<N, O> tile_make_tables() {
  using L = TileLayout<N, O>;
  constexpr int E = L::E;
  TileTables<N, O> tb{};
  auto rx = [](int e) { return e < N ? L::PX + e : (e < 2 * N ? L::LX + (e - N) : L::OX + (e - 2 * N)); };
  auto ry = [](int e) { return e < N ? L::PY + e : (e < 2 * N ? L::LY + (e - N) : L::OY + (e - 2 * N)); };
  for (int q = 0; q < L::PAIRS; ++q) {
    const int a = pair_a(q, E), b = pair_b(q, E);
    tb.pair[q][0] = (uint16_t)(rx(a) * 32); tb.pair[q][1] = (uint16_t)(ry(a) * 32);
    tb.pair[q][2] = (uint16_t)(rx(b) * 32); tb.pair[q][3] = (uint16_t)(ry(b) * 32);
  }
  for (int f = 0; f < L::NF; ++f) {
    int i, partner;
    if (f < L::NAA) { i = pair_a(f, N); partner = pair_b(f, N); }
    else { i = (f - L::NAA) / (O > 0 ? O : 1); partner = 2 * N + (f - L::NAA) - i * O; }
    tb.force[f][0] = (uint16_t)(rx(i) * 32); tb.force[f][1] = (uint16_t)(ry(i) * 32);
    tb.force[f][2] = (uint16_t)(rx(partner) * 32); tb.force[f][3] = (uint16_t)(ry(partner) * 32);
  }
  for (int e = 0; e < E; ++e) { tb.ent[e][0] = (uint16_t)(rx(e) * 32); tb.ent[e][1] = (uint16_t)(ry(e) * 32); }
  for (int r = 0; r < L::NROWS; ++r)
    tb.every_step[r] = (r < L::MINT || (r >= L::NAC && r < L::LX) || r == L::DMEAN || r == L::DSTD || r == L::STEP) ? 1 : 0;
  return tb;
}

template <int N, int O>
static cudaError_t tile_launch_no(const DevParams& p, const TileLuts& luts, cudaStream_t st, bool is_reset) {
  using L = TileLayout<N, O>;
  static const TileTables<N, O> tb = tile_make_tables<N, O>();
  const int blocks = (p.B + 31) / 32;
  const size_t smem = (size_t)L::WORDS * sizeof(float);
  if (is_reset) tile_kernel<N, O, 1><<<blocks, TILE_THREADS, smem, st>>>(p, luts, tb);
  else tile_kernel<N, O, 0><<<blocks, TILE_THREADS, smem, st>>>(p, luts, tb);
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
  TileLuts luts{p.lut_obs};   // obs | node | adj are contiguous (fm_abi.cu)
#define X(n, o) if (p.N == n && p.O == o) return tile_launch_no<n, o>(p, luts, st, is_reset);
  FM_TILE_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace fm
