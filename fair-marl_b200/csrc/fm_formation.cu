// Formation-family scenarios on the device: a first, correctness-first CUDA path.
//
// Reference: multiagent/custom_scenarios/nav_fairassign_fairrew_formation_graph.py (FA+FR) and
// nav_fairassign_nofairrew_formation_graph.py (FA), driven by MultiAgentGraphEnv.step (environment.py:816-877) over
// World.step (core.py:250-404).  oracle/formation.py is the float64 restatement these kernels are tested against
// (pinned to 755 steps of the unmodified reference).
//
// Mapping: ONE THREAD PER ENV.  The per-agent loop of MultiAgentGraphEnv.step is inherently sequential in this family
// (agent i's observation rewrites the goal-occupancy table agent i + 1 reads, agent 0's reward call re-solves the
// assignment, a latching agent's velocity is zeroed between its observation and its node rows), and teams are small
// (N <= 4 here: lexifair by enumeration in registers, fm_small.cuh), so a thread walks the reference's own order with
// the whole env in registers / local memory; lanes of a warp are 32 consecutive envs.  All arithmetic is float64 like
// the reference; the state is stored as float32 in API layout (no transposes: fm_formation_get/set_state are plain
// copies).  Outputs are written straight in API layout (per-thread contiguous runs; staging them through shared memory
// for coalesced stores is the next step for this kernel -- it is not tuned).
#include "fm_device.cuh"
#include "fm_launch.h"
#include "fm_small.cuh"

namespace fm {

constexpr int F_OBS = FM_FORMATION_OBS_DIM, F_NODE = FM_FORMATION_NODE_FEAT_DIM, F_MAXO = FM_FORMATION_MAX_OBSTACLES;

template <int N>
struct FEnv {
  double px[N], py[N], vx[N], vy[N], pd[N], lx[N], ly[N], ox[F_MAXO], oy[F_MAXO];
  double occ[N], hist[N], reached[N], dtg[N], treq[N], dleft[N], mint[N], nac[N], noc[N];
  int gm[N];
  bool status[N];
  double dmean, dstd;
  int step, episode;
};

__device__ __forceinline__ double dn(double dx, double dy) {      // sqrt(dx*dx + dy*dy), no contraction (numpy has none)
  return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

template <int N>
__device__ void f_load(const FormParams& p, int b, FEnv<N>& e) {
  for (int i = 0; i < N; ++i) {
    const size_t k = (size_t)b * N + i;
    e.px[i] = p.st.pos[2 * k]; e.py[i] = p.st.pos[2 * k + 1]; e.vx[i] = p.st.vel[2 * k]; e.vy[i] = p.st.vel[2 * k + 1];
    e.pd[i] = p.st.p_dist[k]; e.lx[i] = p.st.landmark_pos[2 * k]; e.ly[i] = p.st.landmark_pos[2 * k + 1];
    e.gm[i] = p.st.goal_match[k]; e.dtg[i] = p.st.dists_to_goal[k]; e.treq[i] = p.st.times_required[k];
    e.dleft[i] = p.st.dist_left_to_goal[k]; e.nac[i] = p.st.num_agent_collisions[k]; e.noc[i] = p.st.num_obstacle_collisions[k];
    e.mint[i] = p.st.min_time[k]; e.status[i] = p.st.status[k] != 0; e.reached[i] = p.st.goal_reached[k];
    e.occ[i] = p.st.occupied[k]; e.hist[i] = p.st.goal_history[k];
  }
  for (int k = 0; k < p.O; ++k) { e.ox[k] = p.st.obstacle_pos[((size_t)b * p.O + k) * 2]; e.oy[k] = p.st.obstacle_pos[((size_t)b * p.O + k) * 2 + 1]; }
  e.dmean = p.st.dist_traveled_mean[b]; e.dstd = p.st.dist_traveled_stddev[b]; e.step = p.st.step[b]; e.episode = p.st.episode[b];
}

template <int N>
__device__ void f_store(const FormParams& p, int b, const FEnv<N>& e, bool statics) {
  for (int i = 0; i < N; ++i) {
    const size_t k = (size_t)b * N + i;
    p.st.pos[2 * k] = (float)e.px[i]; p.st.pos[2 * k + 1] = (float)e.py[i]; p.st.vel[2 * k] = (float)e.vx[i]; p.st.vel[2 * k + 1] = (float)e.vy[i];
    p.st.p_dist[k] = (float)e.pd[i]; p.st.goal_match[k] = e.gm[i]; p.st.dists_to_goal[k] = (float)e.dtg[i];
    p.st.times_required[k] = (float)e.treq[i]; p.st.dist_left_to_goal[k] = (float)e.dleft[i];
    p.st.num_agent_collisions[k] = (float)e.nac[i]; p.st.num_obstacle_collisions[k] = (float)e.noc[i];
    p.st.min_time[k] = (float)e.mint[i]; p.st.status[k] = e.status[i] ? 1 : 0; p.st.goal_reached[k] = (float)e.reached[i];
    p.st.occupied[k] = (float)e.occ[i]; p.st.goal_history[k] = (float)e.hist[i];
    if (statics) { p.st.landmark_pos[2 * k] = (float)e.lx[i]; p.st.landmark_pos[2 * k + 1] = (float)e.ly[i]; }
  }
  if (statics)
    for (int k = 0; k < p.O; ++k) { p.st.obstacle_pos[((size_t)b * p.O + k) * 2] = (float)e.ox[k]; p.st.obstacle_pos[((size_t)b * p.O + k) * 2 + 1] = (float)e.oy[k]; }
  p.st.dist_traveled_mean[b] = (float)e.dmean; p.st.dist_traveled_stddev[b] = (float)e.dstd; p.st.step[b] = e.step; p.st.episode[b] = e.episode;
}

// is_obstacle_collision (:576-586, no walls): closer than 2.0 * (size + size) to any obstacle.
template <int N>
__device__ bool f_obstacle_hit(const FormParams& p, const FEnv<N>& e, double x, double y) {
  const double dmin = 2.0 * (0.05 + 0.05);
  bool hit = false;
  for (int k = 0; k < p.O; ++k) hit = hit || (dn(e.ox[k] - x, e.oy[k] - y) < dmin);
  return hit;
}

// np.mean / np.std (population) of a length-N vector.
template <int N>
__device__ void f_mean_std(const double (&v)[N], double& mean, double& sd) {
  double s = 0.0;
  for (int j = 0; j < N; ++j) s = __dadd_rn(s, v[j]);
  mean = s / N;
  double q = 0.0;
  for (int j = 0; j < N; ++j) { const double d = v[j] - mean; q = __dadd_rn(q, __dmul_rn(d, d)); }
  sd = sqrt(q / N);
}

// Far branch shared by observation (:933-956) and the agent rows of the node features (:1256-1270): nearest goal not
// marked occupied (== 1); if every goal is, the entity itself and a cleared table.
template <int N>
__device__ void f_pick_goal(FEnv<N>& e, double qx, double qy, int slot, double& gx, double& gy, double& occ, double& hist) {
  int best = -1;
  double bd = 0.0;
  for (int g = 0; g < N; ++g) {
    if (e.occ[g] == 1.0) continue;
    const double d = dn(qx - e.lx[g], qy - e.ly[g]);
    if (best < 0 || d < bd) { best = g; bd = d; }
  }
  if (best >= 0) { gx = e.lx[best]; gy = e.ly[best]; occ = e.occ[best]; hist = e.hist[best]; return; }
  for (int g = 0; g < N; ++g) e.occ[g] = 0.0;                              // :951 / :1266
  gx = qx; gy = qy; occ = e.occ[slot]; hist = e.hist[slot];
}

// Scenario.observation (:840-1015) of agent i: 11 values, updates the occupancy table and the goal history.
template <int N>
__device__ void f_observation(const FormParams& p, FEnv<N>& e, int i, float* __restrict__ o) {
  const double x = e.px[i], y = e.py[i];
  double d[N];
  int first = 0;
  for (int g = 0; g < N; ++g) { d[g] = dn(x - e.lx[g], y - e.ly[g]); if (d[g] < d[first]) first = g; }
  int second = first == 0 ? 1 : 0;                                         // np.argsort(dists)[1]
  for (int g = 0; g < N; ++g) if (g != first && d[g] < d[second]) second = g;
  const double sgx = e.lx[second], sgy = e.ly[second], socc = e.occ[second];   // read before the updates below
  const double mind = d[first], th = p.min_dist_thresh;
  double gx, gy, gocc, ghist;
  if (mind < p.min_obs_dist) {
    int chosen = first;
    gx = e.lx[chosen]; gy = e.ly[chosen];
    for (int g = 0; g < N; ++g) {                                          // :866-879 nearby goals marked occupied
      if (!(d[g] < p.min_obs_dist) || e.occ[g] != 1.0) continue;
      bool any = false; double mn = 0.0;
      for (int j = 0; j < N; ++j) { const double q = dn(e.lx[g] - e.px[j], e.ly[g] - e.py[j]); any = any || (q < th); mn = j == 0 ? q : fmin(mn, q); }
      if (!any) e.occ[g] = mn;
    }
    if (mind < th) {                                                       // :882-885
      e.occ[chosen] = 1.0; e.hist[chosen] = (double)i;
    } else {
      bool any = false; double closest = 0.0;
      for (int j = 0; j < N; ++j) { const double q = dn(gx - e.px[j], gy - e.py[j]); any = any || (q < th); closest = j == 0 ? q : fmin(closest, q); }
      if (e.occ[chosen] == 1.0 && any) {                                   // :908-923: nearest FREE goal; `chosen` becomes its
        int k = 0, bestk = -1, bestg = -1; double bd = 0.0;                //   index in the free SUBSET (reference quirk, kept)
        for (int g = 0; g < N; ++g) {
          if (e.occ[g] == 1.0) continue;
          const double q = dn(x - e.lx[g], y - e.ly[g]);
          if (bestk < 0 || q < bd) { bestk = k; bestg = g; bd = q; }
          ++k;
        }
        if (bestk >= 0) { chosen = bestk; gx = e.lx[bestg]; gy = e.ly[bestg]; }   // (no free goal: the reference raises)
      } else {
        e.occ[chosen] = 1.0 - closest;
      }
    }
    gocc = e.occ[chosen]; ghist = e.hist[chosen];                          // :930-931
  } else {
    f_pick_goal<N>(e, x, y, i, gx, gy, gocc, ghist);
  }
  if (o) {
    o[0] = (float)e.vx[i]; o[1] = (float)e.vy[i]; o[2] = (float)x; o[3] = (float)y; o[4] = (float)(gx - x); o[5] = (float)(gy - y);
    o[6] = (float)gocc; o[7] = (float)ghist; o[8] = (float)(sgx - x); o[9] = (float)(sgy - y); o[10] = (float)socc;
  }
}

// graph_observation + _get_entity_feat_relative (:1083-1178, :1222-1340) for ego agent i: [E, 13].
template <int N>
__device__ void f_node_rows(const FormParams& p, FEnv<N>& e, int i, float* __restrict__ rows) {
  const double x = e.px[i], y = e.py[i], vx = e.vx[i], vy = e.vy[i];
  auto put = [&](int r, double rvx, double rvy, double rx, double ry, double gx, double gy, double occ, double hist, double type) {
    if (!rows) return;
    float* q = rows + (size_t)r * F_NODE;
    q[0] = (float)rvx; q[1] = (float)rvy; q[2] = (float)rx; q[3] = (float)ry; q[4] = (float)gx; q[5] = (float)gy;
    q[6] = (float)occ; q[7] = (float)hist; q[8] = (float)rx; q[9] = (float)ry; q[10] = (float)rx; q[11] = (float)ry; q[12] = (float)type;
  };
  for (int a = 0; a < N; ++a) {
    const double qx = e.px[a], qy = e.py[a];
    int first = 0; double mind = 0.0;
    for (int g = 0; g < N; ++g) { const double d = dn(qx - e.lx[g], qy - e.ly[g]); if (g == 0 || d < mind) { first = g; mind = d; } }
    double gx, gy, occ, hist;
    if (mind < p.min_obs_dist) { gx = e.lx[first]; gy = e.ly[first]; occ = e.occ[first]; hist = e.hist[first]; }
    else f_pick_goal<N>(e, qx, qy, a, gx, gy, occ, hist);
    put(a, e.vx[a] - vx, e.vy[a] - vy, qx - x, qy - y, gx - x, gy - y, occ, hist, 0.0);
  }
  for (int a = 0; a < N; ++a)                                              // landmarks: occupied 1, history = landmark id
    put(N + a, 0.0 - vx, 0.0 - vy, e.lx[a] - x, e.ly[a] - y, e.lx[a] - x, e.ly[a] - y, 1.0, (double)a, 1.0);
  for (int k = 0; k < p.O; ++k)                                            // obstacles: id None -> history 0
    put(2 * N + k, 0.0 - vx, 0.0 - vy, e.ox[k] - x, e.oy[k] - y, e.ox[k] - x, e.oy[k] - y, 1.0, 0.0, 2.0);
}

// cached_dist_mag (core.py:204-228) -> adj [E, E].
template <int N>
__device__ void f_adj(const FormParams& p, const FEnv<N>& e, float* __restrict__ adj) {
  if (!adj) return;
  const int E = 2 * N + p.O;
  auto X = [&](int s) { return s < N ? e.px[s] : (s < 2 * N ? e.lx[s - N] : e.ox[s - 2 * N]); };
  auto Y = [&](int s) { return s < N ? e.py[s] : (s < 2 * N ? e.ly[s - N] : e.oy[s - 2 * N]); };
  for (int a = 0; a < E; ++a) {
    adj[a * E + a] = 0.0f;
    for (int c = a + 1; c < E; ++c) { const float d = (float)dn(X(a) - X(c), Y(a) - Y(c)); adj[a * E + c] = d; adj[c * E + a] = d; }
  }
}

template <int N>
__device__ void f_assign(FEnv<N>& e) {                                     // cdist + lexifair (:704-721, :481-486)
  double c[N * N];
  for (int a = 0; a < N; ++a)
    for (int g = 0; g < N; ++g) c[a * N + g] = dn(e.px[a] - e.lx[g], e.py[a] - e.ly[g]);
  lexifair_small<N>(c, e.gm);
}

// env.reset()'s observation pass (environment.py:882-898): obs_i, then node rows_i, per agent.
template <int N>
__device__ void f_observe(const FormParams& p, int b, FEnv<N>& e) {
  const int E = 2 * N + p.O;
  for (int i = 0; i < N; ++i) {
    f_observation<N>(p, e, i, p.out.obs ? p.out.obs + ((size_t)b * N + i) * F_OBS : nullptr);
    f_node_rows<N>(p, e, i, p.out.node_obs ? p.out.node_obs + ((size_t)b * N + i) * E * F_NODE : nullptr);
  }
  f_adj<N>(p, e, p.out.adj ? p.out.adj + (size_t)b * E * E : nullptr);
}

// reset_world + random_scenario (:217-487) with the Philox draw scheme of the navigation kernels: draw counter per
// (seed, global env, episode); obstacles 0.8 * U, agents U rejected vs obstacles (2.0x) / placed agents (1.05x), goals
// 0.8 * U rejected vs obstacles (2.0x) / placed goals (1.2x).  Positions are float32 values, predicates float64.
template <int N>
__device__ void f_reset(const FormParams& p, int b, FEnv<N>& e) {
  const long long genv = p.env_offset + b;
  const float ws = (float)p.world_size, half = (float)(p.world_size / 2);
  uint32_t d = 0;
  auto draw = [&](float& x, float& y) {
    uint32_t c0 = d, c1 = (uint32_t)e.episode, c2 = (uint32_t)((unsigned long long)genv & 0xffffffffull),
             c3 = (uint32_t)((unsigned long long)genv >> 32);
    philox4x32_10(c0, c1, c2, c3, p.seed_lo, p.seed_hi);
    x = __fsub_rn(__fmul_rn(ws, u01_24(c0)), half); y = __fsub_rn(__fmul_rn(ws, u01_24(c1)), half);
    ++d;
  };
  for (int k = 0; k < p.O; ++k) { float x, y; draw(x, y); e.ox[k] = (double)__fmul_rn(0.8f, x); e.oy[k] = (double)__fmul_rn(0.8f, y); }
  const double r2 = 0.05 + 0.05;
  for (int pass = 0; pass < 2; ++pass) {
    double* X = pass ? e.lx : e.px; double* Y = pass ? e.ly : e.py;
    const double dsame = pass ? 1.2 * r2 : 1.05 * r2;
    for (int a = 0; a < N; ++a) {
      while (true) {
        float fx, fy; draw(fx, fy);
        if (pass) { fx = __fmul_rn(0.8f, fx); fy = __fmul_rn(0.8f, fy); }
        const double x = fx, y = fy;
        bool bad = f_obstacle_hit<N>(p, e, x, y);
        for (int j = 0; j < a; ++j) bad = bad || (dn(X[j] - x, Y[j] - y) < dsame);
        if (!bad || d >= (uint32_t)MAX_DRAWS) { X[a] = x; Y[a] = y; break; }
      }
    }
  }
  for (int i = 0; i < N; ++i) {
    e.vx[i] = e.vy[i] = 0.0; e.pd[i] = 0.0; e.status[i] = false; e.treq[i] = e.dtg[i] = e.dleft[i] = -1.0;
    e.noc[i] = e.nac[i] = 0.0; e.hist[i] = -1.0; e.reached[i] = -1.0; e.occ[i] = 0.0;
    if (p.has_max_speed) e.mint[i] = dn(e.px[i] - e.lx[i], e.py[i] - e.ly[i]) / p.max_speed;   // goal_match = arange here (:229, :474-476)
  }
  e.step = 0;
  f_assign<N>(e);
  e.episode += 1;
}

template <int N>
__global__ void __launch_bounds__(128) formation_reset_kernel(const FormParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  FEnv<N> e;
  f_load<N>(p, b, e);
  const bool doit = !p.mask || p.mask[b] != 0;
  if (doit) f_reset<N>(p, b, e);
  f_observe<N>(p, b, e);
  f_store<N>(p, b, e, doit);
}

// MultiAgentGraphEnv.step (environment.py:816-877) + graphworker auto-reset (env_wrappers.py:856-865).
template <int N>
__global__ void __launch_bounds__(128) formation_step_kernel(const FormParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  FEnv<N> e;
  f_load<N>(p, b, e);
  const int O = p.O, E = 2 * N + O;
  e.step += 1;                                                             // :819, :823
  // ---- World.step: action force (core.py:277-298), pair forces from the positions at step entry (:301-316, :370-404)
  double Fx[N], Fy[N];
  for (int i = 0; i < N; ++i) {
    const int a = p.actions[(size_t)b * N + i];
    Fx[i] = (((a == 1) ? 1.0 : 0.0) - ((a == 2) ? 1.0 : 0.0)) * 5.0;       // environment.py:301-311
    Fy[i] = (((a == 3) ? 1.0 : 0.0) - ((a == 4) ? 1.0 : 0.0)) * 5.0;
  }
  const double km = 0.02, dmin = 0.05 + 0.05;
  auto pair_force = [&](double dx, double dy, double& fx, double& fy) {
    const double dist = dn(dx, dy);
    const double z = -(dist - dmin) / km;
    const double pen = (fmax(z, 0.0) + log1p(exp(-fabs(z)))) * km;         // np.logaddexp(0, z) * k  (:391)
    fx = 300.0 * dx / dist * pen; fy = 300.0 * dy / dist * pen;            // :392
  };
  for (int a = 0; a < N; ++a) {
    for (int c = a + 1; c < N; ++c) {
      double fx, fy; pair_force(e.px[a] - e.px[c], e.py[a] - e.py[c], fx, fy);
      if (!e.status[a]) { Fx[a] = fx + Fx[a]; Fy[a] = fy + Fy[a]; }        // core.py:397
      if (!e.status[c]) { Fx[c] = -fx + Fx[c]; Fy[c] = -fy + Fy[c]; }      // core.py:398
    }
    for (int k = 0; k < O; ++k) {                                          // obstacles: whatever the status (:401)
      double fx, fy; pair_force(e.px[a] - e.ox[k], e.py[a] - e.oy[k], fx, fy);
      Fx[a] = fx + Fx[a]; Fy[a] = fy + Fy[a];
    }
  }
  for (int i = 0; i < N; ++i) {                                            // integrate_state (:338-356): every agent
    double vx = __dmul_rn(e.vx[i], 0.75), vy = __dmul_rn(e.vy[i], 0.75);
    vx = __dadd_rn(vx, __dmul_rn(Fx[i], 0.1)); vy = __dadd_rn(vy, __dmul_rn(Fy[i], 0.1));
    if (p.has_max_speed) {
      const double sp = dn(vx, vy);
      if (sp > p.max_speed) { vx = __dmul_rn(vx / sp, p.max_speed); vy = __dmul_rn(vy / sp, p.max_speed); }
    }
    e.vx[i] = vx; e.vy[i] = vy;
    const double sx = __dmul_rn(vx, 0.1), sy = __dmul_rn(vy, 0.1);
    e.px[i] = __dadd_rn(e.px[i], sx); e.py[i] = __dadd_rn(e.py[i], sy);
    e.pd[i] = __dadd_rn(e.pd[i], dn(sx, sy));
  }
  f_adj<N>(p, e, p.out.adj ? p.out.adj + (size_t)b * E * E : nullptr);

  // ---- per-agent loop (environment.py:832-864): observation, reward, node rows, done, info -- in this order
  double rew[N];
  bool done[N], all_done = true;
  const double th = p.min_dist_thresh, dcoll = 1.05 * (0.05 + 0.05);
  for (int i = 0; i < N; ++i) {
    f_observation<N>(p, e, i, p.out.obs ? p.out.obs + ((size_t)b * N + i) * F_OBS : nullptr);
    // reward (:691-802)
    double fairness;
    if (e.dtg[i] == -1.0) { double m, s; f_mean_std<N>(e.pd, m, s); fairness = m / (s + 0.0001); }
    else fairness = e.dmean / (e.dstd + 0.0001);
    if (i == 0) f_assign<N>(e);                                            // :704-721: re-assignment every step
    const double x = e.px[i], y = e.py[i];
    const double dg = dn(x - e.lx[e.gm[i]], y - e.ly[e.gm[i]]);
    double r = 0.0;
    if (dg < th) {                                                         // :725-733
      if (!e.status[i]) { e.status[i] = true; e.vx[i] = 0.0; e.vy[i] = 0.0; r += p.goal_rew; }
    } else {
      r -= dg;
    }
    int hits = 0;
    for (int j = 0; j < N; ++j) if (j != i && dn(e.px[j] - x, e.py[j] - y) < dcoll) { r -= p.collision_rew; ++hits; }
    const bool ohit = f_obstacle_hit<N>(p, e, x, y);
    if (ohit) r -= p.collision_rew;
    if (p.fairness_reward) {                                               // :770-786
      double fair = p.fair_rew * tanh(fairness - p.zeroshift);
      if (fair < -p.fair_rew) fair = -p.fair_rew;
      r += fair;
    }
    r = fmin(fmax(r, -2.0 * p.collision_rew), p.goal_rew + p.fair_rew);
    rew[i] = r;
    f_node_rows<N>(p, e, i, p.out.node_obs ? p.out.node_obs + ((size_t)b * N + i) * E * F_NODE : nullptr);
    done[i] = e.status[i] || e.step >= p.episode_length;                   // environment.py:237-247
    all_done = all_done && done[i];
    // info_callback (:489-575)
    {
      int near = 0; double d = 0.0;
      for (int g = 0; g < N; ++g) { const double q = dn(x - e.lx[g], y - e.ly[g]); if (g == 0 || q < d) { near = g; d = q; } }
      const double now = (double)e.step * 0.1, nr = (double)near;
      if (d < th && (nr != e.reached[i] && e.reached[i] != -1.0)) { e.reached[i] = nr; e.dleft[i] = d; }          // :497-499
      if (d < th && e.treq[i] == -1.0) { e.treq[i] = now; e.dtg[i] = e.pd[i]; e.dleft[i] = d; e.reached[i] = nr; }   // :501-505
      if (e.treq[i] == -1.0) { e.dtg[i] = e.pd[i]; e.dleft[i] = d; }                                                // :507-509
      if (d > th && e.treq[i] != -1.0) { e.dtg[i] = e.pd[i]; e.treq[i] = now; e.dleft[i] = d; }                      // :511-514
      if (d < th && nr == e.reached[i]) { e.dleft[i] = d; e.reached[i] = nr; }                                      // :516-518
      if (ohit) e.noc[i] += 1.0;                                                                                    // :521-523
      e.nac[i] += (double)hits;
      f_mean_std<N>(e.dtg, e.dmean, e.dstd);                                                                        // :534-535
      if (p.out.info) {
        double tm, ts; f_mean_std<N>(e.treq, tm, ts);
        float* q = p.out.info + ((size_t)b * N + i) * INFO_F;
        q[0] = (float)r; q[1] = (float)e.dleft[i]; q[2] = (float)e.treq[i]; q[3] = (float)e.nac[i]; q[4] = (float)e.noc[i];
        q[5] = (float)e.dmean; q[6] = (float)e.dstd; q[7] = (float)(e.dmean / (e.dstd + 0.0001)); q[8] = (float)e.dtg[i];
        q[9] = (float)e.treq[i]; q[10] = (float)tm; q[11] = (float)ts; q[12] = (float)(tm / (ts + 0.0001)); q[13] = (float)e.mint[i];
      }
    }
  }
  double total = 0.0;
  for (int i = 0; i < N; ++i) total += rew[i];
  for (int i = 0; i < N; ++i) {
    if (p.out.reward) p.out.reward[(size_t)b * N + i] = (float)(p.collaborative ? total : rew[i]);   // environment.py:867-870
    if (p.out.done) p.out.done[(size_t)b * N + i] = done[i] ? 1 : 0;
  }
  const bool reset = p.auto_reset && all_done;                             // env_wrappers.py:859-865
  if (reset) { f_reset<N>(p, b, e); f_observe<N>(p, b, e); }
  f_store<N>(p, b, e, reset);
}

cudaError_t launch_formation(const FormParams& p, bool is_reset, cudaStream_t st) {
  const int blocks = (p.B + 127) / 128;
#define FM_F_CASE(n)                                                            \
  case n:                                                                       \
    if (is_reset) formation_reset_kernel<n><<<blocks, 128, 0, st>>>(p);         \
    else formation_step_kernel<n><<<blocks, 128, 0, st>>>(p);                   \
    break;
  switch (p.N) {
    FM_F_CASE(2) FM_F_CASE(3) FM_F_CASE(4)
    default: return cudaErrorInvalidValue;
  }
#undef FM_F_CASE
  return cudaGetLastError();
}

}  // namespace fm
