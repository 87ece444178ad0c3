// Fused graph-network forward of the rollout policy: ONE kernel does, per graph,
//   process_adj (gnn_new.py:381-413)  ->  EmbedConv (gnn_new.py:23-141)  ->  act(TransformerConv) x (1 + gnn_layer_N)
//   (gnn_new.py:252-271, PyG TransformerConv(heads, concat=False, beta=False, edge_dim=1, root_weight=True))
//   ->  node gather (graph_aggr='node') or global mean / max / add pool (gnn_new.py:555-575)
// reading `node_obs` and `adj` exactly as the simulator's step kernel wrote them (adj once per env, shared by the N ego
// graphs of that env).  It replaces ~400 small torch kernels over [graphs, E, E, 16] edge-message tensors; the dense
// head behind it (MLPBase, GRU step, action / value layer: plain GEMMs over [graphs, 64]) stays with cuBLAS.
//
// Mapping: one WARP per graph; the weights of the network (~36 KB) live in shared memory for the whole (persistent)
// CTA, everything of a graph lives in a warp-private shared-memory scratch and in registers; no block barrier after the
// weights are loaded.  The graphs are tiny and dense (E = 2N + O + W nodes, E <= 32; ~half of the E(E-1) directed pairs
// are edges), so there is no edge list: lane (t, j) owns target node t and every LPT-th source of it (LPT = 32 / E
// lanes per target), skips the pairs that are not edges, and the LPT partial results of a target are combined with
// shuffles in a fixed order (deterministic, no atomics).
//   stage A0  hn[r]   = W_n [x_r | embed(type_r)] + b           per node      (lin1 splits into a node and an edge part)
//   stage A   x0[c]   = sum_{r -> c} LN(act(W_h LN(act(hn[r] + w_d d_rc)) + b_h))    per edge, 16-wide in registers
//   stage B1  [q|k|v|skip][n] = W_l x[n] + b_l                    dense 16 -> 160, lane = 5 output columns
//   stage B2  alpha[h,t,s] = softmax_s((q_th . k_sh + d_st (q_th . w_e,h)) / 4)      in registers, per target
//   stage B3  x'[t] = act(mean_h(sum_s alpha (v_sh + d_st w_e,h)) + skip[t])
// Supported shape family (the configuration of every shipped model_weights/*/config.yaml): embed_hidden_size =
// gnn_hidden_size = 16, gnn_num_heads = 3, gnn_concat_heads = False; any node feature width <= 17, embed_layer_N <= 2,
// any gnn_layer_N, ReLU or Tanh, with or without LayerNorm.  fp32 throughout (the parity bar is 1e-5 against the
// reference's fp32 torch modules; tests/test_gpu_policy.py).
#include <cstdio>

#include "fm_device.cuh"
#include "fm_launch.h"

namespace fm {

constexpr int GH = 16;            // embed_hidden_size == gnn_hidden_size
constexpr int GHEADS = 3;
constexpr int GHC = GHEADS * GH;  // 48
constexpr int GQW = 3 * GHC + GH; // 160 columns: q | k | v | skip
constexpr int GQS = GQW + 4;      // row stride of the q|k|v|skip scratch: 41 x 16 bytes (odd: LDS.128 conflict free over rows)
constexpr int GXS = GH + 4;       // row stride of the node-state scratch: 5 x 16 bytes
constexpr int G_WARPS = 8;

// packed weight blob (floats), see fair_marl_b200/policy.py::pack_gnn_weights
constexpr int GW_WN = 0;                       // [16 feature rows][16]  lin1 weight, node-feature part (rows >= NF-1 are zero)
constexpr int GW_TYPE = GW_WN + 16 * GH;       // [4][16]   lin1 . embed(type) + lin1 bias
constexpr int GW_WD = GW_TYPE + 4 * GH;        // [16]      lin1 weight, edge-attribute column
constexpr int GW_LN1 = GW_WD + GH;             // [16] gamma, [16] beta
constexpr int GW_EMBED_END = GW_LN1 + 2 * GH;
constexpr int GW_HID = GH * GH + GH + 2 * GH;  // per hidden embed layer: W[f][g] | b | gamma | beta
constexpr int GW_CONV = GH * GQW + GQW + GHC;  // per conv: W[k][160] | b[160] | w_e[48]

__host__ __device__ inline int gnn_weight_floats(int embed_layers, int conv_layers) {
  return GW_EMBED_END + embed_layers * GW_HID + conv_layers * GW_CONV;
}

struct GnnArgs {
  const float* w;        // packed weights
  const float* node;     // [M, E, NF]
  const float* adj;      // [M / rep, E, E]
  const int* agent_id;   // [M] node index gathered by aggr 0 (null: m % rep)
  float* out;            // [M, 16]
  int M, rep, NF, embed_layers, conv_layers, aggr, relu, ln;
  float max_edge_dist;
  int wfloats;
};

template <bool RELU>
__device__ __forceinline__ float g_act(float x) { return RELU ? fmaxf(x, 0.0f) : tanhf(x); }

// LayerNorm over 16 registers (torch: biased variance of the centred values, eps = 1e-5 inside the root)
__device__ __forceinline__ void g_layer_norm(float (&h)[GH], const float* __restrict__ gb, bool ln) {
  if (!ln) return;
  float s = 0.f;
#pragma unroll
  for (int f = 0; f < GH; ++f) s += h[f];
  const float mean = s * (1.0f / GH);
  float q = 0.f;
#pragma unroll
  for (int f = 0; f < GH; ++f) { h[f] -= mean; q = fmaf(h[f], h[f], q); }
  const float inv = rsqrtf(q * (1.0f / GH) + 1e-5f);
#pragma unroll
  for (int f = 0; f < GH; f += 4) {
    const float4 g = *reinterpret_cast<const float4*>(gb + f), b = *reinterpret_cast<const float4*>(gb + GH + f);
    h[f] = fmaf(h[f] * inv, g.x, b.x); h[f + 1] = fmaf(h[f + 1] * inv, g.y, b.y);
    h[f + 2] = fmaf(h[f + 2] * inv, g.z, b.z); h[f + 3] = fmaf(h[f + 3] * inv, g.w, b.w);
  }
}

template <int E>
struct GnnLayout {
  static constexpr int LPT = 32 / E;                          // lanes per target node
  static constexpr int SPL = (E + LPT - 1) / LPT;             // sources per lane
  static constexpr int NFP = 20;                              // padded feature row (NF <= 17)
  static constexpr int ADJ = 0, FEAT = (E * E + 3) & ~3, X = FEAT + E * NFP, Q = X + E * GXS, END = Q + E * GQS;
  static constexpr int WORDS = (END + 3) & ~3;                // floats per warp
};

template <int E, bool RELU>
__global__ void __launch_bounds__(G_WARPS * 32, 2) gnn_kernel(const __grid_constant__ GnnArgs a) {
  using L = GnnLayout<E>;
  constexpr int LPT = L::LPT, SPL = L::SPL;
  extern __shared__ __align__(16) float smem[];
  float* W = smem;                                            // weights
  const int wpad = (a.wfloats + 3) & ~3;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* S = smem + wpad + wib * L::WORDS;
  for (int k = threadIdx.x; k < a.wfloats; k += blockDim.x) W[k] = __ldg(a.w + k);
  __syncthreads();
  float* s_adj = S + L::ADJ;
  float* s_feat = S + L::FEAT;
  float* s_x = S + L::X;
  float* s_q = S + L::Q;
  const bool ln = a.ln != 0;
  const int NF = a.NF, KF = NF - 1;
  // lane (t, j): target t = lane / LPT, sources s = j, j + LPT, ...
  const int t = lane / LPT, j = lane - t * LPT;
  const bool tl = t < E;                                      // this lane owns a target
  const unsigned tmask = __ballot_sync(FULL, tl);             // the lanes that do (shuffles inside `if (tl)` name only these)

  for (int m = blockIdx.x * G_WARPS + wib; m < a.M; m += gridDim.x * G_WARPS) {
    // ---- inputs -> scratch ---------------------------------------------------------------------------------
    {
      const float* ga = a.adj + (size_t)(m / a.rep) * (E * E);
      for (int k = lane; k < E * E; k += 32) s_adj[k] = __ldg(ga + k);
      const float* gn = a.node + (size_t)m * E * NF;
      for (int k = lane; k < E * NF; k += 32) { const int n = k / NF; s_feat[n * L::NFP + (k - n * NF)] = __ldcs(gn + k); }
    }
    __syncwarp();
    // ---- A0: hn[n][c] = sum_k Wn[k][c] feat[n][k] + Ttype[type_n][c]  -> s_x ------------------------------
    for (int o = lane; o < E * GH; o += 32) {
      const int n = o >> 4, c = o & 15;
      const float* f = s_feat + n * L::NFP;
      const int ty = min(max((int)f[KF], 0), 3);
      float acc = W[GW_TYPE + ty * GH + c];
      for (int k = 0; k < KF; ++k) acc = fmaf(W[GW_WN + k * GH + c], f[k], acc);
      s_x[n * GXS + c] = acc;
    }
    __syncwarp();
    // ---- A: edge messages, summed at the target -------------------------------------------------------------
    {
      float acc[GH];
#pragma unroll
      for (int f = 0; f < GH; ++f) acc[f] = 0.f;
      if (tl) {
#pragma unroll 1
        for (int r = j; r < E; r += LPT) {
          const float d = s_adj[r * E + t];                   // edge r -> t
          if (!(d < a.max_edge_dist && d > 0.0f)) continue;   // process_adj: strict <, > 0
          float h[GH];
#pragma unroll
          for (int f = 0; f < GH; f += 4) {
            const float4 hn = *reinterpret_cast<const float4*>(s_x + r * GXS + f);
            const float4 wd = *reinterpret_cast<const float4*>(W + GW_WD + f);
            h[f] = g_act<RELU>(fmaf(wd.x, d, hn.x)); h[f + 1] = g_act<RELU>(fmaf(wd.y, d, hn.y));
            h[f + 2] = g_act<RELU>(fmaf(wd.z, d, hn.z)); h[f + 3] = g_act<RELU>(fmaf(wd.w, d, hn.w));
          }
          g_layer_norm(h, W + GW_LN1, ln);
          for (int l = 0; l < a.embed_layers; ++l) {
            const float* wl = W + GW_EMBED_END + l * GW_HID;
            float o[GH];
#pragma unroll
            for (int g = 0; g < GH; g += 4) {
              const float4 b = *reinterpret_cast<const float4*>(wl + GH * GH + g);
              o[g] = b.x; o[g + 1] = b.y; o[g + 2] = b.z; o[g + 3] = b.w;
            }
#pragma unroll
            for (int f = 0; f < GH; ++f) {
#pragma unroll
              for (int g = 0; g < GH; g += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wl + f * GH + g);   // same address in every lane: broadcast
                o[g] = fmaf(w4.x, h[f], o[g]); o[g + 1] = fmaf(w4.y, h[f], o[g + 1]);
                o[g + 2] = fmaf(w4.z, h[f], o[g + 2]); o[g + 3] = fmaf(w4.w, h[f], o[g + 3]);
              }
            }
#pragma unroll
            for (int g = 0; g < GH; ++g) h[g] = g_act<RELU>(o[g]);
            g_layer_norm(h, wl + GH * GH + GH, ln);
          }
#pragma unroll
          for (int f = 0; f < GH; ++f) acc[f] += h[f];
        }
      }
      __syncwarp();                                           // every lane has read hn before s_x is overwritten
      // the LPT partial sums of a target, in lane order
#pragma unroll
      for (int f = 0; f < GH; ++f) {
        float v = acc[f];
#pragma unroll
        for (int o = 1; o < LPT; ++o) { const float u = __shfl_down_sync(FULL, acc[f], o); if (j + o < LPT) v += u; }
        acc[f] = v;
      }
      if (tl && j == 0) {
#pragma unroll
        for (int f = 0; f < GH; f += 4) *reinterpret_cast<float4*>(s_x + t * GXS + f) = make_float4(acc[f], acc[f + 1], acc[f + 2], acc[f + 3]);
      }
    }
    __syncwarp();
    // ---- convs ----------------------------------------------------------------------------------------------
    for (int layer = 0; layer < a.conv_layers; ++layer) {
      const float* wc = W + GW_EMBED_END + a.embed_layers * GW_HID + layer * GW_CONV;
      const float* bc = wc + GH * GQW;
      const float* we = bc + GQW;
      // B1: [q | k | v | skip][n][col] for col = lane + 32 u, u < 5
      {
        constexpr int NB = E < 9 ? E : 9;                     // nodes per pass (accumulators: NB x 5)
#pragma unroll 1
        for (int n0 = 0; n0 < E; n0 += NB) {
          float acc[NB][5];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const float b = bc[lane + 32 * u];
#pragma unroll
            for (int n = 0; n < NB; ++n) acc[n][u] = b;
          }
#pragma unroll
          for (int k0 = 0; k0 < GH; k0 += 4) {
            float w[4][5];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int u = 0; u < 5; ++u) w[kk][u] = wc[(k0 + kk) * GQW + lane + 32 * u];
#pragma unroll
            for (int n = 0; n < NB; ++n) {
              if (n0 + n < E) {
                const float4 x4 = *reinterpret_cast<const float4*>(s_x + (n0 + n) * GXS + k0);
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                  acc[n][u] = fmaf(w[0][u], x4.x, acc[n][u]); acc[n][u] = fmaf(w[1][u], x4.y, acc[n][u]);
                  acc[n][u] = fmaf(w[2][u], x4.z, acc[n][u]); acc[n][u] = fmaf(w[3][u], x4.w, acc[n][u]);
                }
              }
            }
          }
#pragma unroll
          for (int n = 0; n < NB; ++n)
            if (n0 + n < E) {
#pragma unroll
              for (int u = 0; u < 5; ++u) s_q[(n0 + n) * GQS + lane + 32 * u] = acc[n][u];
            }
        }
      }
      __syncwarp();
      // B2 + B3: attention of target t over its sources, per head, in registers
      {
        float out[GH];
#pragma unroll
        for (int c = 0; c < GH; ++c) out[c] = 0.f;
        float adsum[GHEADS] = {0.f, 0.f, 0.f};
        if (tl) {
          float sc[GHEADS][SPL], dd[SPL];
          float qe[GHEADS];
          // scores
#pragma unroll
          for (int h = 0; h < GHEADS; ++h) {
            float q[GH];
#pragma unroll
            for (int c = 0; c < GH; c += 4) {
              const float4 q4 = *reinterpret_cast<const float4*>(s_q + t * GQS + h * GH + c);
              q[c] = q4.x; q[c + 1] = q4.y; q[c + 2] = q4.z; q[c + 3] = q4.w;
            }
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < GH; ++c) e = fmaf(q[c], we[h * GH + c], e);
            qe[h] = e;
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
              const int s = j + i * LPT;
              float v = -INFINITY;
              if (s < E) {
                const float d = s_adj[s * E + t];             // edge s -> t
                if (h == 0) dd[i] = d;
                if (d < a.max_edge_dist && d > 0.0f) {
                  float dot = 0.f;
#pragma unroll
                  for (int c = 0; c < GH; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(s_q + s * GQS + GHC + h * GH + c);
                    dot = fmaf(q[c], k4.x, dot); dot = fmaf(q[c + 1], k4.y, dot);
                    dot = fmaf(q[c + 2], k4.z, dot); dot = fmaf(q[c + 3], k4.w, dot);
                  }
                  v = (dot + qe[h] * d) * 0.25f;               // 1 / sqrt(16)
                }
              } else if (h == 0) {
                dd[i] = 0.f;
              }
              sc[h][i] = v;
            }
          }
          // softmax over the sources of (h, t): LPT lanes x SPL entries
#pragma unroll
          for (int h = 0; h < GHEADS; ++h) {
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < SPL; ++i) mx = fmaxf(mx, sc[h][i]);
#pragma unroll
            for (int o = 1; o < LPT; ++o) {                   // all-to-all inside the target's lane group
              const float u = __shfl_sync(tmask, mx, t * LPT + ((j + o) % LPT));
              mx = fmaxf(mx, u);
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < SPL; ++i) { const float ex = sc[h][i] == -INFINITY ? 0.f : expf(sc[h][i] - mx); sc[h][i] = ex; sum += ex; }
            float tot = 0.f;                                   // fixed order: lane j = 0, 1, ... of the group
#pragma unroll
            for (int o = 0; o < LPT; ++o) tot += __shfl_sync(tmask, sum, t * LPT + o);
            const float inv = tot > 0.f ? 1.0f / tot : 0.f;   // targets without incoming edges: alpha = 0 (nan_to_num)
#pragma unroll
            for (int i = 0; i < SPL; ++i) { sc[h][i] *= inv; adsum[h] = fmaf(sc[h][i], dd[i], adsum[h]); }
          }
          // weighted values
#pragma unroll
          for (int i = 0; i < SPL; ++i) {
            const int s = j + i * LPT;
            if (s < E) {
#pragma unroll
              for (int h = 0; h < GHEADS; ++h) {
                const float al = sc[h][i];
#pragma unroll
                for (int c = 0; c < GH; c += 4) {
                  const float4 v4 = *reinterpret_cast<const float4*>(s_q + s * GQS + 2 * GHC + h * GH + c);
                  out[c] = fmaf(al, v4.x, out[c]); out[c + 1] = fmaf(al, v4.y, out[c + 1]);
                  out[c + 2] = fmaf(al, v4.z, out[c + 2]); out[c + 3] = fmaf(al, v4.w, out[c + 3]);
                }
              }
            }
          }
        }
        // combine the LPT lanes of a target (lane order), add the edge term, mean over heads, skip, activation
#pragma unroll
        for (int c = 0; c < GH; ++c) {
          float v = out[c];
#pragma unroll
          for (int o = 1; o < LPT; ++o) { const float u = __shfl_down_sync(FULL, out[c], o); if (j + o < LPT) v += u; }
          out[c] = v;
        }
#pragma unroll
        for (int h = 0; h < GHEADS; ++h) {
          float v = adsum[h];
#pragma unroll
          for (int o = 1; o < LPT; ++o) { const float u = __shfl_down_sync(FULL, adsum[h], o); if (j + o < LPT) v += u; }
          adsum[h] = v;
        }
        __syncwarp();                                         // s_x (layer input) is dead: every lane is past B1
        if (tl && j == 0) {
#pragma unroll
          for (int c = 0; c < GH; ++c) {
            float v = out[c];
#pragma unroll
            for (int h = 0; h < GHEADS; ++h) v = fmaf(adsum[h], we[h * GH + c], v);
            v = v * (1.0f / GHEADS) + s_q[t * GQS + 3 * GHC + c];
            s_x[t * GXS + c] = g_act<RELU>(v);
          }
        }
      }
      __syncwarp();
    }
    // ---- aggregation ----------------------------------------------------------------------------------------
    if (lane < GH) {
      float v;
      if (a.aggr == 0) {
        const int node = a.agent_id ? a.agent_id[m] : (m % a.rep);
        v = s_x[min(max(node, 0), E - 1) * GXS + lane];
      } else {
        v = s_x[lane];
        for (int n = 1; n < E; ++n) { const float u = s_x[n * GXS + lane]; v = a.aggr == 2 ? fmaxf(v, u) : v + u; }
        if (a.aggr == 1) v *= (1.0f / E);
      }
      a.out[(size_t)m * GH + lane] = v;
    }
    __syncwarp();
  }
}

template <int E>
static cudaError_t gnn_launch_e(const GnnArgs& a, cudaStream_t st) {
  using L = GnnLayout<E>;
  const size_t smem = (size_t)(((a.wfloats + 3) & ~3) + G_WARPS * L::WORDS) * sizeof(float);
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const int want = (a.M + G_WARPS - 1) / G_WARPS;
  const int grid = want < 2 * sms ? want : 2 * sms;
  if (a.relu) {
    e = cudaFuncSetAttribute(gnn_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gnn_kernel<E, true><<<grid, G_WARPS * 32, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(gnn_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    gnn_kernel<E, false><<<grid, G_WARPS * 32, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// E values compiled: the navigation / formation shapes at N = 2..7 with O <= 3 and up to 2 walls
#define FM_GNN_CASES(X) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(17) X(19)

bool gnn_supported_entities(int E) {
#define X(e) if (E == e) return true;
  FM_GNN_CASES(X)
#undef X
  return false;
}

cudaError_t launch_gnn(const FmGnnConfig& c, const float* weights, const float* node, const float* adj, const int* agent_id,
                       float* out, cudaStream_t st) {
  GnnArgs a;
  a.w = weights; a.node = node; a.adj = adj; a.agent_id = agent_id; a.out = out;
  a.M = c.num_graphs; a.rep = c.graphs_per_adj; a.NF = c.node_feat_dim; a.embed_layers = c.embed_layers;
  a.conv_layers = c.conv_layers; a.aggr = c.aggr; a.relu = c.relu; a.ln = c.layer_norm;
  a.max_edge_dist = (float)c.max_edge_dist;
  a.wfloats = gnn_weight_floats(c.embed_layers, c.conv_layers);
  if (a.M <= 0) return cudaSuccess;
#define X(e) if (c.num_entities == e) return gnn_launch_e<e>(a, st);
  FM_GNN_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

int gnn_weight_count(int embed_layers, int conv_layers) { return gnn_weight_floats(embed_layers, conv_layers); }

// =====================================================================================================================
// Fused policy HEAD (everything behind the graph network): GR_Actor.forward / GR_Critic.forward after gnn_base
// (onpolicy/algorithms/graph_actor_critic.py:150-178, :380-397):
//   x = [obs | nbd]  ->  MLPBase (utils/mlp.py: LayerNorm(in), then (Linear, act, LayerNorm) x (1 + layer_N))
//     ->  RNNLayer, one step (utils/rnn.py:23-28, :57: GRU on h * mask, LayerNorm)
//     ->  actor: Categorical head (utils/act.py, distributions.py:14-28): log-softmax, mode or an inverse-CDF draw from a
//         caller-supplied uniform, log-prob of the action;   critic: v_out (a Linear; PopArt's forward is the same map).
// One WARP per row (graph); hidden_size = 64: lane l owns hidden units l and l + 32.  All weights of the head (~32 K
// floats) live in shared memory for the whole persistent CTA, stored k-quad interleaved ([k / 4][column][4]) so that one
// LDS.128 feeds four FMAs and consecutive lanes read consecutive 16-byte words; the activation vector of the row sits in
// a warp-private 64-float scratch and is read as broadcast LDS.128.  Replaces ~30 torch launches per network and step
// (LayerNorm kernels with one CTA per 64-float row, GEMVs for the 5- and 1-wide output layers, multinomial).
constexpr int HH = 64;                 // hidden_size
constexpr int HD = 32;                 // padded input width (obs_dim + 16 <= 32)
constexpr int HEAD_WARPS = 16;
constexpr int HW_FN = 0;                               // feature_norm gamma[32], beta[32]
constexpr int HW_FC1 = HW_FN + 2 * HD;                 // [8 quads][64][4], then b[64], ln gamma[64], beta[64]
constexpr int HW_FC1_SZ = HD * HH + 3 * HH;
constexpr int HW_FC_SZ = HH * HH + 3 * HH;             // per extra layer: [16 quads][64][4], b, gamma, beta
constexpr int HW_GRU_SZ = 2 * HH * 3 * HH + 2 * 3 * HH + 2 * HH;   // Wih [16][192][4] | Whh | b_ih[192] | b_hh[192] | ln gamma, beta
constexpr int HW_OUT_SZ = 8 * HH + 8;                  // Wo[8 rows][64] | bo[8]

__host__ __device__ inline int head_weight_floats(int layers, int recurrent) {
  return HW_FC1 + HW_FC1_SZ + layers * HW_FC_SZ + (recurrent ? HW_GRU_SZ : 0) + HW_OUT_SZ;
}

struct HeadArgs {
  const float* w;
  const float* obs;        // [M, obs_dim] or null (obs_dim 0)
  const float* nbd;        // [M, 16]
  const float* rnn_in;     // [M, 64]   (recurrent)
  const float* mask;       // [M]       (recurrent)
  const float* u;          // [M] uniforms in [0, 1) for the categorical draw; null: mode
  float* rnn_out;          // [M, 64]
  float* logp;             // [M]   actor
  long long* action;       // [M]   actor
  float* value;            // [M]   critic
  int M, obs_dim, layers, recurrent, feat_norm, relu, A, wfloats;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// LayerNorm over the 64 values held two per lane
__device__ __forceinline__ void head_ln(float& a, float& b, const float* __restrict__ g, int lane) {
  const float mean = warp_sum(a + b) * (1.0f / HH);
  a -= mean; b -= mean;
  const float var = warp_sum(fmaf(a, a, b * b)) * (1.0f / HH);
  const float inv = rsqrtf(var + 1e-5f);
  a = fmaf(a * inv, g[lane], g[HH + lane]);
  b = fmaf(b * inv, g[lane + 32], g[HH + lane + 32]);
}

// acc[c] += sum_k W[k][col_c] x[k] for NC columns of this lane; W k-quad interleaved with `ncol` columns, x in shared memory
template <int NC>
__device__ __forceinline__ void head_matvec(const float* __restrict__ Wq, int ncol, int quads, const float* __restrict__ xs,
                                            const int (&col)[NC], float (&acc)[NC]) {
#pragma unroll 4
  for (int q = 0; q < quads; ++q) {
    const float4 x4 = *reinterpret_cast<const float4*>(xs + 4 * q);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 w4 = *reinterpret_cast<const float4*>(Wq + ((size_t)q * ncol + col[c]) * 4);
      acc[c] = fmaf(w4.x, x4.x, acc[c]); acc[c] = fmaf(w4.y, x4.y, acc[c]);
      acc[c] = fmaf(w4.z, x4.z, acc[c]); acc[c] = fmaf(w4.w, x4.w, acc[c]);
    }
  }
}

template <bool RELU>
__global__ void __launch_bounds__(HEAD_WARPS * 32, 1) head_kernel(const __grid_constant__ HeadArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* W = smem;
  const int wpad = (a.wfloats + 3) & ~3;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* xs = smem + wpad + wib * (2 * HH);                  // activation vector | masked hidden state
  float* hs = xs + HH;
  for (int k = threadIdx.x; k < a.wfloats; k += blockDim.x) W[k] = __ldg(a.w + k);
  __syncthreads();
  const int D = a.obs_dim + GH;
  const float* w_gru = W + HW_FC1 + HW_FC1_SZ + a.layers * HW_FC_SZ;
  const float* w_out = w_gru + (a.recurrent ? HW_GRU_SZ : 0);
  const int c2[2] = {lane, lane + 32};
  for (int m = blockIdx.x * HEAD_WARPS + wib; m < a.M; m += gridDim.x * HEAD_WARPS) {
    // ---- input [obs | nbd], feature LayerNorm over D values --------------------------------------------------------
    float x = 0.f;
    if (lane < a.obs_dim) x = __ldcs(a.obs + (size_t)m * a.obs_dim + lane);
    else if (lane < D) x = __ldcs(a.nbd + (size_t)m * GH + (lane - a.obs_dim));
    if (a.feat_norm) {
      const float mean = warp_sum(x) / (float)D;
      const float d = lane < D ? x - mean : 0.f;
      const float var = warp_sum(d * d) / (float)D;
      x = lane < D ? fmaf(d * rsqrtf(var + 1e-5f), W[HW_FN + lane], W[HW_FN + HD + lane]) : 0.f;
    }
    xs[lane] = x;
    __syncwarp();
    // ---- fc1 and the layer_N hidden layers: Linear, act, LayerNorm ----------------------------------------------------
    float h0, h1;
    {
      const float* wl = W + HW_FC1;
      float acc[2] = {wl[HD * HH + lane], wl[HD * HH + lane + 32]};
      head_matvec<2>(wl, HH, HD / 4, xs, c2, acc);
      h0 = g_act<RELU>(acc[0]); h1 = g_act<RELU>(acc[1]);
      head_ln(h0, h1, wl + HD * HH + HH, lane);
    }
    __syncwarp();
    xs[lane] = h0; xs[lane + 32] = h1;
    __syncwarp();
    for (int l = 0; l < a.layers; ++l) {
      const float* wl = W + HW_FC1 + HW_FC1_SZ + l * HW_FC_SZ;
      float acc[2] = {wl[HH * HH + lane], wl[HH * HH + lane + 32]};
      head_matvec<2>(wl, HH, HH / 4, xs, c2, acc);
      h0 = g_act<RELU>(acc[0]); h1 = g_act<RELU>(acc[1]);
      head_ln(h0, h1, wl + HH * HH + HH, lane);
      __syncwarp();
      xs[lane] = h0; xs[lane + 32] = h1;
      __syncwarp();
    }
    // ---- one GRU step on h * mask (torch.nn.GRU gate order r | z | n), LayerNorm ----------------------------------------
    if (a.recurrent) {
      const float mk = a.mask[m];
      const float p0 = __ldcs(a.rnn_in + (size_t)m * HH + lane) * mk, p1 = __ldcs(a.rnn_in + (size_t)m * HH + lane + 32) * mk;
      hs[lane] = p0; hs[lane + 32] = p1;
      __syncwarp();
      const int c6[6] = {lane, lane + 32, HH + lane, HH + lane + 32, 2 * HH + lane, 2 * HH + lane + 32};
      const float* bih = w_gru + 2 * HH * 3 * HH;
      const float* bhh = bih + 3 * HH;
      float gi[6], gh[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { gi[c] = bih[c6[c]]; gh[c] = bhh[c6[c]]; }
      head_matvec<6>(w_gru, 3 * HH, HH / 4, xs, c6, gi);
      head_matvec<6>(w_gru + HH * 3 * HH, 3 * HH, HH / 4, hs, c6, gh);
      const float r0 = 1.0f / (1.0f + expf(-(gi[0] + gh[0]))), r1 = 1.0f / (1.0f + expf(-(gi[1] + gh[1])));
      const float z0 = 1.0f / (1.0f + expf(-(gi[2] + gh[2]))), z1 = 1.0f / (1.0f + expf(-(gi[3] + gh[3])));
      const float n0 = tanhf(gi[4] + r0 * gh[4]), n1 = tanhf(gi[5] + r1 * gh[5]);
      h0 = (1.0f - z0) * n0 + z0 * p0; h1 = (1.0f - z1) * n1 + z1 * p1;
      a.rnn_out[(size_t)m * HH + lane] = h0; a.rnn_out[(size_t)m * HH + lane + 32] = h1;
      head_ln(h0, h1, bhh + 3 * HH, lane);
    }
    // ---- output layer -------------------------------------------------------------------------------------------------
    float lg[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float v = 0.f;
      if (o < a.A) v = warp_sum(fmaf(w_out[o * HH + lane], h0, w_out[o * HH + lane + 32] * h1)) + w_out[8 * HH + o];
      lg[o] = v;
    }
    if (a.value) {
      if (lane == 0) a.value[m] = lg[0];
    } else {
      float mx = -INFINITY;
#pragma unroll
      for (int o = 0; o < 8; ++o) if (o < a.A) mx = fmaxf(mx, lg[o]);
      float se = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) if (o < a.A) se += expf(lg[o] - mx);
      const float lse = mx + logf(se);
      int act = 0;
      if (a.u) {                                             // inverse CDF of softmax(logits)
        const float u = a.u[m];
        float cdf = 0.f;
        act = a.A - 1;
        bool set = false;
#pragma unroll
        for (int o = 0; o < 8; ++o) if (o < a.A) { cdf += expf(lg[o] - lse); if (!set && u < cdf) { act = o; set = true; } }
      } else {                                               // FixedCategorical.mode: first maximum
        float best = -INFINITY;
#pragma unroll
        for (int o = 0; o < 8; ++o) if (o < a.A && lg[o] > best) { best = lg[o]; act = o; }
      }
      float sel = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) if (o == act) sel = lg[o];
      if (lane == 0) { a.action[m] = act; a.logp[m] = sel - lse; }
    }
    __syncwarp();
  }
}

int head_weight_count(int layers, int recurrent) { return head_weight_floats(layers, recurrent); }

cudaError_t launch_head(const FmHeadConfig& c, const float* weights, const float* obs, const float* nbd, const float* rnn_in,
                        const float* mask, const float* u, float* rnn_out, float* logp, long long* action, float* value,
                        cudaStream_t st) {
  HeadArgs a;
  a.w = weights; a.obs = obs; a.nbd = nbd; a.rnn_in = rnn_in; a.mask = mask; a.u = u; a.rnn_out = rnn_out; a.logp = logp;
  a.action = action; a.value = value;
  a.M = c.num_rows; a.obs_dim = c.obs_dim; a.layers = c.layers; a.recurrent = c.recurrent; a.feat_norm = c.feature_norm;
  a.relu = c.relu; a.A = c.num_outputs; a.wfloats = head_weight_floats(c.layers, c.recurrent);
  if (a.M <= 0) return cudaSuccess;
  const size_t smem = (size_t)(((a.wfloats + 3) & ~3) + HEAD_WARPS * 2 * HH) * sizeof(float);
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const int want = (a.M + HEAD_WARPS - 1) / HEAD_WARPS;
  const int grid = want < sms ? want : sms;
  if (a.relu) {
    e = cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    head_kernel<true><<<grid, HEAD_WARPS * 32, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    head_kernel<false><<<grid, HEAD_WARPS * 32, smem, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace fm
