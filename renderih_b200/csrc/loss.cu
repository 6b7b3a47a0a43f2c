// Fused training loss of the hot path: GraphLoss.calc_loss for both hands (reference core/Loss.py:103-162 + the weighting of
// calc_loss_GCN, :236-275) as ONE forward and ONE backward kernel, one CTA per (image, hand).
// The reference evaluates it with ~200 small torch kernels per direction (face gathers, normalisations, a 21x778 matmul, SmoothL1 / MSE
// reductions, an index_put with a radix sort in backward): 2 ms of a 36 ms training step at batch 64.
//   terms per hand (sums; the finalize kernel turns them into means and the weighted total):
//     0 vert3d  SmoothL1(v3d_pred, v3d_gt)                         778 x 3
//     1 vert2d  MSE(v2d_pred / S * 2 - 1, v2d_gt / S * 2 - 1)      778 x 2
//     2 joint   SmoothL1(J21 v3d_pred, J21 v3d_gt)                  21 x 3
//     3 normal  SmoothL1(<normalize(e_pred), n_gt>, 0)               F x 3     (n_gt = normalize(e_gt0 x e_gt1))
//     4 edge    SmoothL1(|e_pred|, |e_gt|)                           F x 3
//     5 coarse3 SmoothL1(v3d_252, avgpool4(v3d_gt[perm]))          252 x 3
//     6 coarse2 MSE(v2d_252 / S * 2 - 1, avgpool4(v2d_gt[perm]) / S * 2 - 1)   252 x 2
#include "common.cuh"
using namespace rih;

constexpr int GL_THREADS = 256, GL_V = 778, GL_J = 21, GL_TERMS = 7, GL_MAXVC = 256;

struct GraphLossHand {
  const float* v3p; const float* v2p;     // predictions [B,778,3], [B,778,2]
  const float* v3g; const float* v2g;     // labels
  const float* root_rel;                  // [B,3] added to v3g (right hand, Loss.py:215) or NULL
  const float* v3c; const float* v2c;     // coarse predictions [B,Vc,3], [B,Vc,2]
  const float* J21;                       // [21,778]
  const int* faces;                       // [F,3]
  const int* perm;                        // [pool*Vc] vert_to_GCN permutation
  float* g_v3p; float* g_v2p; float* g_v3c; float* g_v2c;   // gradients (backward only)
};
struct GraphLossArgs { GraphLossHand h[2]; int B, F, Vc, pool; float img; };

__device__ __forceinline__ float sl1(float d) { const float a = fabsf(d); return a < 1.f ? 0.5f * d * d : a - 0.5f; }
__device__ __forceinline__ float sl1_grad(float d) { return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f); }

__device__ __forceinline__ void block_reduce_terms(float (&acc)[GL_TERMS], float* s_red, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < GL_TERMS; ++t) {
    float v = warp_sum(acc[t]);
    if (lane == 0) s_red[warp * GL_TERMS + t] = v;
  }
  __syncthreads();
  if (threadIdx.x < GL_TERMS) {
    float v = 0.f;
    for (int w = 0; w < GL_THREADS / 32; ++w) v += s_red[w * GL_TERMS + threadIdx.x];
    out[threadIdx.x] = v;
  }
}

// BWD = false: partial[hand][b][7] term sums.  BWD = true: gradients; coef[hand-independent 7] = d total / d (term sum), already including
// the upstream gradient, the 1/2 hand average, the term weight and the 1/count of the mean.
template <bool BWD>
__global__ void __launch_bounds__(GL_THREADS)
graph_loss_kernel(GraphLossArgs a, float* __restrict__ partial, const float* __restrict__ coef) {
  pdl_sync();
  __shared__ float s_vp[GL_V * 3], s_vg[GL_V * 3];
  __shared__ float s_gv[BWD ? GL_V * 3 : 1];
  __shared__ float s_jd[GL_J * 3];
  __shared__ float s_red[(GL_THREADS / 32) * GL_TERMS];
  const int b = blockIdx.x, hand = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const GraphLossHand& h = a.h[hand];
  const float inv_img2 = 2.f / a.img;
  float acc[GL_TERMS];
#pragma unroll
  for (int t = 0; t < GL_TERMS; ++t) acc[t] = 0.f;
  float c[GL_TERMS];
#pragma unroll
  for (int t = 0; t < GL_TERMS; ++t) c[t] = BWD ? coef[t] : 0.f;

  // ---- vertices: load, vert3d / vert2d terms
  const float* v3p = h.v3p + (size_t)b * GL_V * 3;
  const float* v3g = h.v3g + (size_t)b * GL_V * 3;
  float rr[3] = {0.f, 0.f, 0.f};
  if (h.root_rel) { rr[0] = h.root_rel[b * 3]; rr[1] = h.root_rel[b * 3 + 1]; rr[2] = h.root_rel[b * 3 + 2]; }
  for (int i = tid; i < GL_V * 3; i += GL_THREADS) {
    const float p = v3p[i], g = v3g[i] + rr[i % 3];
    s_vp[i] = p; s_vg[i] = g;
    const float d = p - g;
    acc[0] += sl1(d);
    if (BWD) s_gv[i] = c[0] * sl1_grad(d);
  }
  const float* v2p = h.v2p + (size_t)b * GL_V * 2;
  const float* v2g = h.v2g + (size_t)b * GL_V * 2;
  for (int i = tid; i < GL_V * 2; i += GL_THREADS) {
    const float d = (v2p[i] * inv_img2 - 1.f) - (v2g[i] * inv_img2 - 1.f);
    acc[1] += d * d;
    if (BWD) h.g_v2p[(size_t)b * GL_V * 2 + i] = c[1] * 2.f * d * inv_img2;
  }
  __syncthreads();
  // ---- joints: warp per (joint, component pair): d = J21[j] . (vp - vg)
  for (int j = warp; j < GL_J; j += GL_THREADS / 32) {
    const float* Jr = h.J21 + (size_t)j * GL_V;
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
    for (int v = lane; v < GL_V; v += 32) {
      const float w = __ldg(Jr + v);
      p0 = fmaf(w, s_vp[v * 3], p0); p1 = fmaf(w, s_vp[v * 3 + 1], p1); p2 = fmaf(w, s_vp[v * 3 + 2], p2);
      g0 = fmaf(w, s_vg[v * 3], g0); g1 = fmaf(w, s_vg[v * 3 + 1], g1); g2 = fmaf(w, s_vg[v * 3 + 2], g2);
    }
    p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2); g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
    if (lane == 0) {
      const float d0 = p0 - g0, d1 = p1 - g1, d2 = p2 - g2;
      acc[2] += sl1(d0) + sl1(d1) + sl1(d2);
      if (BWD) { s_jd[j * 3] = c[2] * sl1_grad(d0); s_jd[j * 3 + 1] = c[2] * sl1_grad(d1); s_jd[j * 3 + 2] = c[2] * sl1_grad(d2); }
    }
  }
  if (BWD) {
    __syncthreads();
    for (int v = tid; v < GL_V; v += GL_THREADS) {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
      for (int j = 0; j < GL_J; ++j) {
        const float w = __ldg(h.J21 + (size_t)j * GL_V + v);
        g0 = fmaf(w, s_jd[j * 3], g0); g1 = fmaf(w, s_jd[j * 3 + 1], g1); g2 = fmaf(w, s_jd[j * 3 + 2], g2);
      }
      s_gv[v * 3] += g0; s_gv[v * 3 + 1] += g1; s_gv[v * 3 + 2] += g2;      // own vertex only: no hazard
    }
    __syncthreads();
  }
  // ---- faces: normal and edge terms
  for (int f = tid; f < a.F; f += GL_THREADS) {
    const int i0 = h.faces[f * 3], i1 = h.faces[f * 3 + 1], i2 = h.faces[f * 3 + 2];
    const int idx[3] = {i0, i1, i2};
    float P[3][3], G[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) { P[k][d] = s_vp[idx[k] * 3 + d]; G[k][d] = s_vg[idx[k] * 3 + d]; }
    float ep[3][3], eg[3][3];      // edge k = v_k - v_{k+1}
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) { ep[k][d] = P[k][d] - P[(k + 1) % 3][d]; eg[k][d] = G[k][d] - G[(k + 1) % 3][d]; }
    float n[3] = {eg[0][1] * eg[1][2] - eg[0][2] * eg[1][1], eg[0][2] * eg[1][0] - eg[0][0] * eg[1][2], eg[0][0] * eg[1][1] - eg[0][1] * eg[1][0]};
    const float nn = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
    n[0] /= nn; n[1] /= nn; n[2] /= nn;
    float ge[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float lp = sqrtf(ep[k][0] * ep[k][0] + ep[k][1] * ep[k][1] + ep[k][2] * ep[k][2]);
      const float lg = sqrtf(eg[k][0] * eg[k][0] + eg[k][1] * eg[k][1] + eg[k][2] * eg[k][2]);
      const float lpc = fmaxf(lp, 1e-12f);
      const float u0 = ep[k][0] / lpc, u1 = ep[k][1] / lpc, u2 = ep[k][2] / lpc;
      const float t = u0 * n[0] + u1 * n[1] + u2 * n[2];
      acc[3] += sl1(t);
      acc[4] += sl1(lp - lg);
      if (BWD) {
        const float gt = c[3] * sl1_grad(t) / lpc;             // d t / d ep = (n - t u) / |ep|
        const float gl = c[4] * sl1_grad(lp - lg);             // d |ep| / d ep = u
        ge[k][0] = gt * (n[0] - t * u0) + gl * u0;
        ge[k][1] = gt * (n[1] - t * u1) + gl * u1;
        ge[k][2] = gt * (n[2] - t * u2) + gl * u2;
      }
    }
    if (BWD) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          atomicAdd(&s_gv[idx[k] * 3 + d], ge[k][d]);                // + on the edge's first vertex
          atomicAdd(&s_gv[idx[(k + 1) % 3] * 3 + d], -ge[k][d]);     // - on its second
        }
    }
  }
  // ---- coarse level (252 vertices): GT = mean of `pool` consecutive permuted label vertices
  const float invp = 1.f / (float)a.pool;
  for (int u = tid; u < a.Vc; u += GL_THREADS) {
    float g3[3] = {0.f, 0.f, 0.f}, g2[2] = {0.f, 0.f};
    for (int k = 0; k < a.pool; ++k) {
      const int v = h.perm[u * a.pool + k];
      g3[0] += s_vg[v * 3]; g3[1] += s_vg[v * 3 + 1]; g3[2] += s_vg[v * 3 + 2];
      g2[0] += v2g[v * 2]; g2[1] += v2g[v * 2 + 1];
    }
    const size_t o3 = ((size_t)b * a.Vc + u) * 3, o2 = ((size_t)b * a.Vc + u) * 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float dd = h.v3c[o3 + d] - g3[d] * invp;
      acc[5] += sl1(dd);
      if (BWD) h.g_v3c[o3 + d] = c[5] * sl1_grad(dd);
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float dd = (h.v2c[o2 + d] * inv_img2 - 1.f) - (g2[d] * invp * inv_img2 - 1.f);
      acc[6] += dd * dd;
      if (BWD) h.g_v2c[o2 + d] = c[6] * 2.f * dd * inv_img2;
    }
  }
  if (BWD) {
    __syncthreads();
    float* gout = h.g_v3p + (size_t)b * GL_V * 3;
    for (int i = tid; i < GL_V * 3; i += GL_THREADS) gout[i] = s_gv[i];
  } else {
    block_reduce_terms(acc, s_red, partial + ((size_t)hand * a.B + b) * GL_TERMS);
  }
}

// out[0] = total; out[1 + hand*7 + t] = mean of term t for `hand`.  w[7] = term weights (LABEL_3D, LABEL_2D, LABEL_3D, NORMAL, alpha*EDGE,
// LABEL_3D, LABEL_2D); coef[7] (for backward) = d total / d (term sum) = w[t] / (2 * count[t]) * upstream.
__global__ void graph_loss_finalize_kernel(const float* __restrict__ partial, int B, int F, int Vc, float w0, float w1, float w2, float w3, float w4,
                                           float w5, float w6, float* __restrict__ out, float* __restrict__ coef) {
  pdl_sync();
  const float w[GL_TERMS] = {w0, w1, w2, w3, w4, w5, w6};
  const float cnt[GL_TERMS] = {(float)B * GL_V * 3, (float)B * GL_V * 2, (float)B * GL_J * 3, (float)B * F * 3, (float)B * F * 3, (float)B * Vc * 3, (float)B * Vc * 2};
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int t = 0; t < GL_TERMS; ++t) {
      float m[2];
      for (int hnd = 0; hnd < 2; ++hnd) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += partial[((size_t)hnd * B + b) * GL_TERMS + t];
        m[hnd] = cnt[t] > 0.f ? s / cnt[t] : 0.f;          // Vc == 0: no coarse level (ManoLoss, core/Loss_mano.py:184-208)
        out[1 + hnd * GL_TERMS + t] = m[hnd];
      }
      total += w[t] * (m[0] + m[1]) * 0.5f;
      coef[t] = cnt[t] > 0.f ? w[t] * 0.5f / cnt[t] : 0.f;
    }
    out[0] = total;
  }
}
__global__ void graph_loss_scale_coef_kernel(const float* __restrict__ coef, const float* __restrict__ upstream, float* __restrict__ out) {
  pdl_sync();
  if (threadIdx.x < GL_TERMS) out[threadIdx.x] = coef[threadIdx.x] * upstream[0];
}

static int fill_args(GraphLossArgs& a, const float* const* ptrs, const int* const* iptrs, float* const* gptrs, int B, int F, int Vc, int pool, float img) {
  RIH_REQUIRE(B > 0 && F > 0 && Vc >= 0 && Vc <= 1024 && pool >= 1, "graph_loss: bad sizes B=%d F=%d Vc=%d pool=%d", B, F, Vc, pool);
  for (int hnd = 0; hnd < 2; ++hnd) {
    GraphLossHand& h = a.h[hnd];
    const float* const* p = ptrs + hnd * 8;
    h.v3p = p[0]; h.v2p = p[1]; h.v3g = p[2]; h.v2g = p[3]; h.root_rel = p[4]; h.v3c = p[5]; h.v2c = p[6]; h.J21 = p[7];
    h.faces = iptrs[hnd * 2]; h.perm = iptrs[hnd * 2 + 1];
    h.g_v3p = gptrs ? gptrs[hnd * 4] : nullptr; h.g_v2p = gptrs ? gptrs[hnd * 4 + 1] : nullptr;
    h.g_v3c = gptrs ? gptrs[hnd * 4 + 2] : nullptr; h.g_v2c = gptrs ? gptrs[hnd * 4 + 3] : nullptr;
  }
  a.B = B; a.F = F; a.Vc = Vc; a.pool = pool; a.img = img;
  return 0;
}

// Forward.  ptrs: 2 x 8 float pointers {v3d_pred, v2d_pred, v3d_gt, v2d_gt, root_rel | NULL, v3d_coarse, v2d_coarse, J21}; iptrs: 2 x 2 int
// pointers {faces [F,3], graph_perm [pool*Vc]}; weights[7]; partial: scratch [2,B,7]; out: [15] (total, 7 means per hand); coef: [7].
// Vc == 0 (coarse pointers / perm may be NULL) drops the coarse terms: the mesh part of ManoLoss.calc_mano_loss (core/Loss_mano.py:157-182).
// reference: GraphLoss.calc_loss + calc_loss_GCN, core/Loss.py:103-162, 201-277
RIH_API int rih_graph_loss_fwd(const float* const* ptrs, const int* const* iptrs, int B, int F, int Vc, int pool, float img, const float* weights_host,
                               float* partial, float* out, float* coef, cudaStream_t s) {
  GraphLossArgs a;
  if (int e = fill_args(a, ptrs, iptrs, nullptr, B, F, Vc, pool, img)) return e;
  launch_k(graph_loss_kernel<false>, dim3(B, 2), GL_THREADS, 0, s, a, partial, nullptr);
  if (int e = check_launch("graph_loss_fwd")) return e;
  const float* w = weights_host;
  launch_k(graph_loss_finalize_kernel, 1, 32, 0, s, partial, B, F, Vc, w[0], w[1], w[2], w[3], w[4], w[5], w[6], out, coef);
  return check_launch("graph_loss_finalize");
}

// Backward: gptrs = 2 x 4 gradient pointers {d v3d_pred, d v2d_pred, d v3d_coarse, d v2d_coarse}; upstream = device pointer to d(total);
// coef from the forward call; coef_scaled: scratch [7].
RIH_API int rih_graph_loss_bwd(const float* const* ptrs, const int* const* iptrs, float* const* gptrs, int B, int F, int Vc, int pool, float img,
                               const float* coef, const float* upstream, float* coef_scaled, cudaStream_t s) {
  GraphLossArgs a;
  if (int e = fill_args(a, ptrs, iptrs, gptrs, B, F, Vc, pool, img)) return e;
  launch_k(graph_loss_scale_coef_kernel, 1, 32, 0, s, coef, upstream, coef_scaled);
  if (int e = check_launch("graph_loss_scale_coef")) return e;
  launch_k(graph_loss_kernel<true>, dim3(B, 2), GL_THREADS, 0, s, a, nullptr, coef_scaled);
  return check_launch("graph_loss_bwd");
}
