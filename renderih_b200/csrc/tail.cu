// Decoder tail as one kernel per direction: avg_head + params_head (scale, trans2d), coord_head (252-vertex mesh),
// dense 252->778 mesh upsample (unsample_layer) and the orthographic projection of both meshes.
// reference: models/decoder.py:139-159 ; projection_batch utils/manoutils.py:26-44
// One CTA per batch element; everything for one hand stays in shared memory.
#include "common.cuh"
using namespace rih;

constexpr int TAIL_THREADS = 1024;   // one CTA per image is latency bound: 32 warps walk the 778 up-sample rows / 252 vertices 4x faster than 8
constexpr int TAIL_MAXV = 256, TAIL_MAXF = 128, TAIL_MAXN = 800;

struct TailParams {
  const float* avg_w; const float* avg_b; const float* par_w; const float* par_b; const float* coord_w; const float* coord_b; const float* U;
};

__global__ void __launch_bounds__(TAIL_THREADS)
tail_fwd_kernel(TailParams P, const float* __restrict__ Lf, int ldl, int V, int F, int Nv, float img,
                float* __restrict__ prm_out /*[B,3] raw scale,tx,ty*/, float* __restrict__ temp_out /*[B,F]*/,
                float* __restrict__ v3c, float* __restrict__ v2c, float* __restrict__ v3, float* __restrict__ v2) {
  pdl_sync();
  __shared__ float s_temp[TAIL_MAXF];
  __shared__ float s_prm[3];
  __shared__ float s_v3[TAIL_MAXV][3];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* L = Lf + (size_t)b * V * ldl;
  for (int f = tid; f < F; f += TAIL_THREADS) {
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc = fmaf(L[(size_t)v * ldl + f], P.avg_w[v], acc);
    acc += P.avg_b[0];
    s_temp[f] = acc;
    temp_out[(size_t)b * F + f] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    for (int c = 0; c < 3; ++c) {
      float acc = 0.f;
      for (int f = lane; f < F; f += 32) acc = fmaf(s_temp[f], P.par_w[c * F + f], acc);
      acc = warp_sum(acc);
      if (lane == 0) { acc += P.par_b[c]; s_prm[c] = acc; prm_out[(size_t)b * 3 + c] = acc; }
    }
  }
  for (int v = warp; v < V; v += TAIL_THREADS / 32) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int f = lane; f < F; f += 32) {
      float x = L[(size_t)v * ldl + f];
      a0 = fmaf(x, P.coord_w[f], a0); a1 = fmaf(x, P.coord_w[F + f], a1); a2 = fmaf(x, P.coord_w[2 * F + f], a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { s_v3[v][0] = a0 + P.coord_b[0]; s_v3[v][1] = a1 + P.coord_b[1]; s_v3[v][2] = a2 + P.coord_b[2]; }
  }
  __syncthreads();
  const float sc = s_prm[0] * img, tx = s_prm[1] * img / 2.f + img / 2.f, ty = s_prm[2] * img / 2.f + img / 2.f;
  for (int v = tid; v < V; v += TAIL_THREADS) {
    float x = s_v3[v][0], y = s_v3[v][1], z = s_v3[v][2];
    float* o3 = v3c + ((size_t)b * V + v) * 3; o3[0] = x; o3[1] = y; o3[2] = z;
    float* o2 = v2c + ((size_t)b * V + v) * 2; o2[0] = sc * x + tx; o2[1] = sc * y + ty;
  }
  for (int n = warp; n < Nv; n += TAIL_THREADS / 32) {
    const float* u = P.U + (size_t)n * V;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int v = lane; v < V; v += 32) {
      float w = u[v];
      a0 = fmaf(w, s_v3[v][0], a0); a1 = fmaf(w, s_v3[v][1], a1); a2 = fmaf(w, s_v3[v][2], a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
      float* o3 = v3 + ((size_t)b * Nv + n) * 3; o3[0] = a0; o3[1] = a1; o3[2] = a2;
      float* o2 = v2 + ((size_t)b * Nv + n) * 2; o2[0] = sc * a0 + tx; o2[1] = sc * a1 + ty;
    }
  }
}

RIH_API int rih_tail_fwd(const float* const* params /*7 ptrs: TailParams order*/, const float* Lf, int ldl, int B, int V, int F, int Nv, float img,
                         float* prm, float* temp, float* v3c, float* v2c, float* v3, float* v2, cudaStream_t s) {
  RIH_REQUIRE(V <= TAIL_MAXV && F <= TAIL_MAXF && Nv <= TAIL_MAXN, "tail_fwd: shape V=%d F=%d Nv=%d exceeds limits", V, F, Nv);
  if (B == 0) return 0;
  TailParams P{params[0], params[1], params[2], params[3], params[4], params[5], params[6]};
  launch_k(tail_fwd_kernel, B, TAIL_THREADS, 0, s, P, Lf, ldl, V, F, Nv, img, prm, temp, v3c, v2c, v3, v2);
  return check_launch("tail_fwd");
}

struct TailGrads {
  float* d_avg_w; float* d_avg_b; float* d_par_w; float* d_par_b; float* d_coord_w; float* d_coord_b; float* d_U;  // accumulated (atomics)
};

__global__ void __launch_bounds__(TAIL_THREADS)
tail_bwd_kernel(TailParams P, TailGrads G, const float* __restrict__ Lf, int ldl, int V, int F, int Nv, float img,
                const float* __restrict__ prm, const float* __restrict__ temp, const float* __restrict__ v3c, const float* __restrict__ v3,
                const float* __restrict__ d_scale, const float* __restrict__ d_trans, const float* __restrict__ d_v3c, const float* __restrict__ d_v2c,
                const float* __restrict__ d_v3, const float* __restrict__ d_v2, float* __restrict__ dLf, int lddl) {
  pdl_sync();
  __shared__ float s_gup[TAIL_MAXN][3];
  __shared__ float s_gc[TAIL_MAXV][3];
  __shared__ float s_dtemp[TAIL_MAXF];
  __shared__ float s_red[3];    // d_s, d_tx, d_ty (w.r.t. pixel-space scale / translation)
  __shared__ float s_dprm[3];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* L = Lf + (size_t)b * V * ldl;
  const float sc = prm[(size_t)b * 3] * img;
  if (tid < 3) s_red[tid] = 0.f;
  __syncthreads();
  float ds = 0.f, dtx = 0.f, dty = 0.f;
  for (int n = tid; n < Nv; n += TAIL_THREADS) {
    size_t o3 = ((size_t)b * Nv + n) * 3, o2 = ((size_t)b * Nv + n) * 2;
    float g0 = d_v3 ? d_v3[o3] : 0.f, g1 = d_v3 ? d_v3[o3 + 1] : 0.f, g2 = d_v3 ? d_v3[o3 + 2] : 0.f;
    if (d_v2) {
      float e0 = d_v2[o2], e1 = d_v2[o2 + 1];
      g0 = fmaf(sc, e0, g0); g1 = fmaf(sc, e1, g1);
      ds += e0 * v3[o3] + e1 * v3[o3 + 1]; dtx += e0; dty += e1;
    }
    s_gup[n][0] = g0; s_gup[n][1] = g1; s_gup[n][2] = g2;
  }
  __syncthreads();
  for (int v = tid; v < V; v += TAIL_THREADS) {
    size_t o3 = ((size_t)b * V + v) * 3, o2 = ((size_t)b * V + v) * 2;
    float g0 = d_v3c ? d_v3c[o3] : 0.f, g1 = d_v3c ? d_v3c[o3 + 1] : 0.f, g2 = d_v3c ? d_v3c[o3 + 2] : 0.f;
    if (d_v2c) {
      float e0 = d_v2c[o2], e1 = d_v2c[o2 + 1];
      g0 = fmaf(sc, e0, g0); g1 = fmaf(sc, e1, g1);
      ds += e0 * v3c[o3] + e1 * v3c[o3 + 1]; dtx += e0; dty += e1;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int n = 0; n < Nv; ++n) {
      float w = P.U[(size_t)n * V + v];
      a0 = fmaf(w, s_gup[n][0], a0); a1 = fmaf(w, s_gup[n][1], a1); a2 = fmaf(w, s_gup[n][2], a2);
    }
    s_gc[v][0] = g0 + a0; s_gc[v][1] = g1 + a1; s_gc[v][2] = g2 + a2;
  }
  ds = warp_sum(ds); dtx = warp_sum(dtx); dty = warp_sum(dty);
  if (lane == 0) { atomicAdd(&s_red[0], ds); atomicAdd(&s_red[1], dtx); atomicAdd(&s_red[2], dty); }
  __syncthreads();
  if (tid < 3) {
    float g = (tid == 0) ? (d_scale ? d_scale[b] : 0.f) + s_red[0] * img
                         : (d_trans ? d_trans[(size_t)b * 2 + tid - 1] : 0.f) + s_red[tid] * img / 2.f;
    s_dprm[tid] = g;
    if (G.d_par_b) atomicAdd(G.d_par_b + tid, g);
  }
  __syncthreads();
  for (int f = tid; f < F; f += TAIL_THREADS) {
    float acc = 0.f;
    for (int c = 0; c < 3; ++c) {
      acc = fmaf(P.par_w[c * F + f], s_dprm[c], acc);
      if (G.d_par_w) atomicAdd(G.d_par_w + c * F + f, s_dprm[c] * temp[(size_t)b * F + f]);
    }
    s_dtemp[f] = acc;
  }
  __syncthreads();
  // dLf and avg_w gradient (warp per vertex)
  for (int v = warp; v < V; v += TAIL_THREADS / 32) {
    float aw = P.avg_w[v], g0 = s_gc[v][0], g1 = s_gc[v][1], g2 = s_gc[v][2];
    float dw = 0.f;
    for (int f = lane; f < F; f += 32) {
      float val = aw * s_dtemp[f] + g0 * P.coord_w[f] + g1 * P.coord_w[F + f] + g2 * P.coord_w[2 * F + f];
      dLf[((size_t)b * V + v) * lddl + f] = val;
      dw = fmaf(s_dtemp[f], L[(size_t)v * ldl + f], dw);
    }
    dw = warp_sum(dw);
    if (lane == 0 && G.d_avg_w) atomicAdd(G.d_avg_w + v, dw);
  }
  // coord_w / coord_b / avg_b gradients
  for (int o = tid; o < 3 * F; o += TAIL_THREADS) {
    int c = o / F, f = o - c * F;
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc = fmaf(s_gc[v][c], L[(size_t)v * ldl + f], acc);
    if (G.d_coord_w) atomicAdd(G.d_coord_w + o, acc);
  }
  if (tid < 3 && G.d_coord_b) {
    float acc = 0.f;
    for (int v = 0; v < V; ++v) acc += s_gc[v][tid];
    atomicAdd(G.d_coord_b + tid, acc);
  }
  if (tid == 32 && G.d_avg_b) {
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc += s_dtemp[f];
    atomicAdd(G.d_avg_b, acc);
  }
  if (G.d_U) {
    for (int i = tid; i < Nv * V; i += TAIL_THREADS) {
      int n = i / V, v = i - n * V;
      const float* c3 = v3c + ((size_t)b * V + v) * 3;
      atomicAdd(G.d_U + i, s_gup[n][0] * c3[0] + s_gup[n][1] * c3[1] + s_gup[n][2] * c3[2]);
    }
  }
}

RIH_API int rih_tail_bwd(const float* const* params, float* const* grads /*7 ptrs, may be null*/, const float* Lf, int ldl, int B, int V, int F, int Nv, float img,
                         const float* prm, const float* temp, const float* v3c, const float* v3,
                         const float* d_scale, const float* d_trans, const float* d_v3c, const float* d_v2c, const float* d_v3, const float* d_v2,
                         float* dLf, int lddl, cudaStream_t s) {
  RIH_REQUIRE(V <= TAIL_MAXV && F <= TAIL_MAXF && Nv <= TAIL_MAXN, "tail_bwd: shape exceeds limits");
  if (B == 0) return 0;
  TailParams P{params[0], params[1], params[2], params[3], params[4], params[5], params[6]};
  TailGrads G{grads[0], grads[1], grads[2], grads[3], grads[4], grads[5], grads[6]};
  launch_k(tail_bwd_kernel, B, TAIL_THREADS, 0, s, P, G, Lf, ldl, V, F, Nv, img, prm, temp, v3c, v3, d_scale, d_trans, d_v3c, d_v2c, d_v3, d_v2, dLf, lddl);
  return check_launch("tail_bwd");
}

// ============================================================== row gather: y[b, i, :] = x[b, idx[i], :]   (GCN_to_vert / vert_to_GCN, graph_upsample p)
// reference: GCN_vert_convert models/model_zoo/__init__.py:85-96 ; decoder.py:165-172 (graph_upsample(p=4) then GCN_to_vert)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, float* __restrict__ y, int B, int Vin, int Vout, int C, int div) {
  pdl_sync();
  long long total = (long long)B * Vout * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long t = i / C; int v = (int)(t % Vout); int b = (int)(t / Vout);
    y[i] = x[((size_t)b * Vin + idx[v] / div) * C + c];
  }
}
__global__ void scatter_rows_add_kernel(const float* __restrict__ dy, const int* __restrict__ idx, float* __restrict__ dx, int B, int Vin, int Vout, int C, int div) {
  pdl_sync();
  long long total = (long long)B * Vout * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long t = i / C; int v = (int)(t % Vout); int b = (int)(t / Vout);
    atomicAdd(dx + ((size_t)b * Vin + idx[v] / div) * C + c, dy[i]);
  }
}
RIH_API int rih_gather_rows(const float* x, const int* idx, float* y, int B, int Vin, int Vout, int C, int div, cudaStream_t s) {
  long long total = (long long)B * Vout * C;
  if (total == 0) return 0;
  int grid = (int)min(ew_ctas(s), (total + 255) / 256);
  launch_k(gather_rows_kernel, grid, 256, 0, s, x, idx, y, B, Vin, Vout, C, div);
  return check_launch("gather_rows");
}
// dx must be zero-initialised (or hold a gradient to accumulate into)
RIH_API int rih_scatter_rows_add(const float* dy, const int* idx, float* dx, int B, int Vin, int Vout, int C, int div, cudaStream_t s) {
  long long total = (long long)B * Vout * C;
  if (total == 0) return 0;
  int grid = (int)min(ew_ctas(s), (total + 255) / 256);
  launch_k(scatter_rows_add_kernel, grid, 256, 0, s, dy, idx, dx, B, Vin, Vout, C, div);
  return check_launch("scatter_rows_add");
}
