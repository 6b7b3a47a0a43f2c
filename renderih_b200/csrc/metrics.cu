// Evaluation metrics of the hot path's caller (reference apps/eval_interhand.py:300-438 loop body, utils/eval_metrics.py:36-50):
// joint regression, root alignment, bone-length rescale, per-joint / per-vertex L2 errors, Procrustes-aligned errors (joints and
// mesh, single hand and the reference's zero-padded "double" variant), relative root translation error and contact deviation.
// The reference spends ~150 small torch kernels + two batched cuSOLVER SVDs + a pytorch3d kNN per batch on this; here it is
//   eval_metrics_kernel  one CTA per (sample, hand), everything in shared memory, 3 block reductions
//   eval_pair_kernel     one CTA per sample: "mrrpe" and the 778 x 778 nearest-vertex contact deviation
// HBM-bound in principle (37 KB in per sample, ~13 KB out); in practice latency bound (a 4x4 Jacobi eigen-solve per CTA).
//
// Procrustes: the reference takes R = V Z U^T from the SVD of K = X1 X2^T (Kabsch).  The same proper rotation is the unit quaternion
// that is the dominant eigenvector of Horn's symmetric 4x4 matrix N(K), and tr(R K) is its eigenvalue -- a 4x4 cyclic Jacobi sweep in
// fp64 on one lane instead of a 3x3 SVD; identical result wherever the optimum is unique.
#include "common.cuh"
using namespace rih;

constexpr int EM_THREADS = 256, EM_V = 778, EM_J = 21, EM_WARPS = EM_THREADS / 32;
constexpr int EM_SAMPLE = 8;    // floats per (hand, sample): ori_j, ori_v, scaled_j, scaled_v, pa_j, pa_v, double_pa_j, double_pa_v (means over points)

struct EvalHand { const float* vp; const float* vg; const float* J21; };
struct EvalArgs {
  EvalHand h[2];
  int B;
  float* sample;      // [2,B,8]
  float* per_joint;   // [2,B,2,21]  (ori, scaled) or NULL
  float* per_vert;    // [2,B,2,778] or NULL
  float* roots;       // [2,B,2,3]   (pred root joint, gt root joint), absolute
};

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* s_red) {   // every thread ends up with the block totals
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                       // s_red may still be read from the previous call
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float r = warp_sum(v[i]);
    if (lane == 0) s_red[warp * N + i] = r;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < EM_WARPS; ++w) r += s_red[w * N + i];
    v[i] = r;
  }
}

// Similarity transform from the centred moments: K[a][b] = sum x1_a x2_b, var1 = sum |x1|^2 (both about the means mu1, mu2).
// out[13] = R (row major 9), scale, t (3).
__device__ void solve_similarity(const double K[9], double var1, const double mu1[3], const double mu2[3], float* out) {
  const double Sxx = K[0], Sxy = K[1], Sxz = K[2], Syx = K[3], Syy = K[4], Syz = K[5], Szx = K[6], Szy = K[7], Szz = K[8];
  double A[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double Q[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 4; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j]; }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (A[p][q] == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 4; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 4; ++k) { const double qkp = Q[k][p], qkq = Q[k][q]; Q[k][p] = c * qkp - s * qkq; Q[k][q] = s * qkp + c * qkq; }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (A[i][i] > A[best][best]) best = i;
  double w = Q[0][best], x = Q[1][best], y = Q[2][best], z = Q[3][best];
  const double n = 1.0 / sqrt(w * w + x * x + y * y + z * z);
  w *= n; x *= n; y *= n; z *= n;
  const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                       2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                       2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
  double tr = 0.0;                                       // tr(R K) = sum_ij R_ij K_ji
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tr += R[i * 3 + j] * K[j * 3 + i];
  const double scale = tr / var1;
  for (int i = 0; i < 9; ++i) out[i] = (float)R[i];
  out[9] = (float)scale;
  for (int i = 0; i < 3; ++i) out[10 + i] = (float)(mu2[i] - scale * (R[i * 3] * mu1[0] + R[i * 3 + 1] * mu1[1] + R[i * 3 + 2] * mu1[2]));
}

__device__ __forceinline__ float aligned_err(const float* T, const float* x1, const float* x2) {
  float e = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float d = T[9] * (T[i * 3] * x1[0] + T[i * 3 + 1] * x1[1] + T[i * 3 + 2] * x1[2]) + T[10 + i] - x2[i];
    e += d * d;
  }
  return sqrtf(e);
}

__global__ void __launch_bounds__(EM_THREADS) eval_metrics_kernel(EvalArgs a) {
  pdl_sync();
  __shared__ float s_vp[EM_V * 3], s_vg[EM_V * 3];         // absolute on load, root-relative after step 2
  __shared__ float s_jp[EM_J * 3], s_jg[EM_J * 3];
  __shared__ float s_red[EM_WARPS * 20];
  __shared__ float s_T[4][13];                             // similarity transforms: joints, verts, double joints, double verts
  const int b = blockIdx.x, hand = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const EvalHand& h = a.h[hand];
  const float* vp = h.vp + (size_t)b * EM_V * 3;
  const float* vg = h.vg + (size_t)b * EM_V * 3;
  for (int i = tid; i < EM_V * 3; i += EM_THREADS) { s_vp[i] = vp[i]; s_vg[i] = vg[i]; }
  __syncthreads();
  // ---- 1. joints = J21 @ verts (Jr.__call__, eval_interhand.py:169-170): one warp per (set, joint) row
  for (int r = warp; r < 2 * EM_J; r += EM_WARPS) {
    const int j = r % EM_J;
    const float* src = r < EM_J ? s_vp : s_vg;
    const float* Jrow = h.J21 + (size_t)j * EM_V;
    float x = 0.f, y = 0.f, z = 0.f;
    for (int v = lane; v < EM_V; v += 32) { const float w = Jrow[v]; x += w * src[v * 3]; y += w * src[v * 3 + 1]; z += w * src[v * 3 + 2]; }
    x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
    if (lane == 0) { float* dst = (r < EM_J ? s_jp : s_jg) + j * 3; dst[0] = x; dst[1] = y; dst[2] = z; }
  }
  __syncthreads();
  // ---- 2. roots, bone length (joint 1 - joint 0), scale (:321-344)
  const float rp[3] = {s_jp[0], s_jp[1], s_jp[2]}, rg[3] = {s_jg[0], s_jg[1], s_jg[2]};
  float lp = 0.f, lg = 0.f;
#pragma unroll
  for (int d = 0; d < 3; ++d) { const float dp = s_jp[3 + d] - rp[d], dg = s_jg[3 + d] - rg[d]; lp += dp * dp; lg += dg * dg; }
  const float scale = sqrtf(lg) / sqrtf(lp);
  __syncthreads();                                         // everyone has read the roots before they are zeroed
  if (tid < 3 && a.roots) { float* r = a.roots + ((size_t)hand * a.B + b) * 6; r[tid] = rp[tid]; r[3 + tid] = rg[tid]; }
  for (int i = tid; i < EM_V * 3; i += EM_THREADS) { s_vp[i] -= rp[i % 3]; s_vg[i] -= rg[i % 3]; }
  if (tid < EM_J * 3) { s_jp[tid] -= rp[tid % 3]; s_jg[tid] -= rg[tid % 3]; }
  __syncthreads();
  // ---- 3. first moments of the four point sets (joints pred / gt, verts pred / gt)
  float m[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = 0.f;
  if (tid < EM_J) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { m[d] = s_jp[tid * 3 + d]; m[3 + d] = s_jg[tid * 3 + d]; }
  }
  for (int v = tid; v < EM_V; v += EM_THREADS) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { m[6 + d] += s_vp[v * 3 + d]; m[9 + d] += s_vg[v * 3 + d]; }
  }
  block_sum<12>(m, s_red);
#pragma unroll
  for (int d = 0; d < 6; ++d) { m[d] *= 1.f / EM_J; m[6 + d] *= 1.f / EM_V; }
  // ---- 4. centred second moments: [0..8] K, [9] var1 for joints; [10..19] for verts
  float c[20];
#pragma unroll
  for (int i = 0; i < 20; ++i) c[i] = 0.f;
  if (tid < EM_J) {
    float x1[3], x2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { x1[d] = s_jp[tid * 3 + d] - m[d]; x2[d] = s_jg[tid * 3 + d] - m[3 + d]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c[9] += x1[i] * x1[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) c[i * 3 + j] = x1[i] * x2[j];
    }
  }
  for (int v = tid; v < EM_V; v += EM_THREADS) {
    float x1[3], x2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { x1[d] = s_vp[v * 3 + d] - m[6 + d]; x2[d] = s_vg[v * 3 + d] - m[9 + d]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c[19] += x1[i] * x1[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) c[10 + i * 3 + j] += x1[i] * x2[j];
    }
  }
  block_sum<20>(c, s_red);
  // ---- 5. four eigen-solves on lane 0 of warps 0..3.  The "double" sets (:405-424) are the right hand's points plus as many all-zero
  //         points in both sets: mean' = mean / 2, K' = K + n mu1 mu2^T / 2, var1' = var1 + n |mu1|^2 / 2.
  if (lane == 0 && warp < 4 && (warp < 2 || hand == 1)) {
    const bool verts = warp & 1, dbl = warp >= 2;
    const float* cc = c + (verts ? 10 : 0);
    const float* mm = m + (verts ? 6 : 0);
    const double n = verts ? (double)EM_V : (double)EM_J;
    double K[9], var1 = cc[9], mu1[3], mu2[3];
    for (int i = 0; i < 9; ++i) K[i] = cc[i];
    for (int d = 0; d < 3; ++d) { mu1[d] = mm[d]; mu2[d] = mm[3 + d]; }
    if (dbl) {
      for (int i = 0; i < 3; ++i) { var1 += 0.5 * n * mu1[i] * mu1[i]; for (int j = 0; j < 3; ++j) K[i * 3 + j] += 0.5 * n * mu1[i] * mu2[j]; }
      for (int d = 0; d < 3; ++d) { mu1[d] *= 0.5; mu2[d] *= 0.5; }
    }
    solve_similarity(K, var1, mu1, mu2, s_T[warp]);
  }
  __syncthreads();
  // ---- 6. errors: [0] ori joint, [1] ori vert, [2] scaled joint, [3] scaled vert, [4] pa joint, [5] pa vert, [6] double pa joint, [7] double pa vert
  float e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) e[i] = 0.f;
  const bool dbl = hand == 1;
  if (tid < EM_J) {
    const float* x1 = s_jp + tid * 3;
    const float* x2 = s_jg + tid * 3;
    float eo = 0.f, es = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { const float o = x1[d] - x2[d], s = x1[d] * scale - x2[d]; eo += o * o; es += s * s; }
    eo = sqrtf(eo); es = sqrtf(es);
    e[0] = eo; e[2] = es;
    e[4] = aligned_err(s_T[0], x1, x2);
    if (dbl) e[6] = aligned_err(s_T[2], x1, x2);
    if (a.per_joint) { float* o = a.per_joint + ((size_t)hand * a.B + b) * 2 * EM_J; o[tid] = eo; o[EM_J + tid] = es; }
  }
  for (int v = tid; v < EM_V; v += EM_THREADS) {
    const float* x1 = s_vp + v * 3;
    const float* x2 = s_vg + v * 3;
    float eo = 0.f, es = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { const float o = x1[d] - x2[d], s = x1[d] * scale - x2[d]; eo += o * o; es += s * s; }
    eo = sqrtf(eo); es = sqrtf(es);
    e[1] += eo; e[3] += es;
    e[5] += aligned_err(s_T[1], x1, x2);
    if (dbl) e[7] += aligned_err(s_T[3], x1, x2);
    if (a.per_vert) { float* o = a.per_vert + ((size_t)hand * a.B + b) * 2 * EM_V; o[v] = eo; o[EM_V + v] = es; }
  }
  block_sum<8>(e, s_red);
  if (tid == 0) {
    float* o = a.sample + ((size_t)hand * a.B + b) * EM_SAMPLE;
    o[0] = e[0] / EM_J; o[1] = e[1] / EM_V; o[2] = e[2] / EM_J; o[3] = e[3] / EM_V; o[4] = e[4] / EM_J; o[5] = e[5] / EM_V;
    if (dbl) {   // the zero points map to t: each contributes |t|
      const float tj = sqrtf(s_T[2][10] * s_T[2][10] + s_T[2][11] * s_T[2][11] + s_T[2][12] * s_T[2][12]);
      const float tv = sqrtf(s_T[3][10] * s_T[3][10] + s_T[3][11] * s_T[3][11] + s_T[3][12] * s_T[3][12]);
      o[6] = (e[6] + EM_J * tj) / (2 * EM_J); o[7] = (e[7] + EM_V * tv) / (2 * EM_V);
    } else { o[6] = 0.f; o[7] = 0.f; }
  }
}

// One CTA per sample.  mrrpe[b][3] = |(root_left_pred - root_right_pred) - (root_left_gt - root_right_gt)| per component (the reference's
// `.sum(axis=1)` at :470 runs over a singleton axis); cdev[b] = mean over the GT-right vertices whose nearest GT-left vertex is within
// `contact` of |pred_left[nn] - pred_right|, NaN (0/0) without contact (utils/eval_metrics.py:36-50).
__global__ void __launch_bounds__(EM_THREADS)
eval_pair_kernel(const float* __restrict__ pl, const float* __restrict__ pr, const float* __restrict__ gl, const float* __restrict__ gr,
                 const float* __restrict__ roots, int B, float contact, float* __restrict__ mrrpe, float* __restrict__ cdev) {
  pdl_sync();
  __shared__ float s_gl[EM_V * 3];
  __shared__ float s_red[EM_WARPS * 2];
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t base = (size_t)b * EM_V * 3;
  for (int i = tid; i < EM_V * 3; i += EM_THREADS) s_gl[i] = gl[base + i];
  if (tid < 3) {
    const float* rl = roots + (size_t)b * 6;
    const float* rr = roots + ((size_t)B + b) * 6;
    mrrpe[b * 3 + tid] = fabsf((rl[tid] - rr[tid]) - (rl[3 + tid] - rr[3 + tid]));
  }
  __syncthreads();
  float acc[2] = {0.f, 0.f};
  for (int v = tid; v < EM_V; v += EM_THREADS) {
    const float x = gr[base + v * 3], y = gr[base + v * 3 + 1], z = gr[base + v * 3 + 2];
    float best = 3.4e38f;
    int arg = 0;
    for (int u = 0; u < EM_V; ++u) {
      const float dx = x - s_gl[u * 3], dy = y - s_gl[u * 3 + 1], dz = z - s_gl[u * 3 + 2];
      const float d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best) { best = d2; arg = u; }
    }
    if (sqrtf(best) <= contact) {
      const float dx = pl[base + arg * 3] - pr[base + v * 3], dy = pl[base + arg * 3 + 1] - pr[base + v * 3 + 1], dz = pl[base + arg * 3 + 2] - pr[base + v * 3 + 2];
      acc[0] += sqrtf(dx * dx + dy * dy + dz * dz);
      acc[1] += 1.f;
    }
  }
  block_sum<2>(acc, s_red);
  if (tid == 0) cdev[b] = acc[0] / acc[1];
}

// ptrs: {pred_left, gt_left, J21_left, pred_right, gt_right, J21_right}; verts [B,778,3] metres, J21 [21,778].
// sample [2,B,8]; per_joint [2,B,2,21] | NULL; per_vert [2,B,2,778] | NULL; roots [2,B,2,3]; mrrpe [B,3]; cdev [B].
// reference: apps/eval_interhand.py:300-438 (loop body), utils/eval_metrics.py:36-50 (compute_cdev)
RIH_API int rih_eval_metrics(const float* const* ptrs, int B, float contact_dist, float* sample, float* per_joint, float* per_vert, float* roots,
                             float* mrrpe, float* cdev, cudaStream_t s) {
  RIH_REQUIRE(B > 0 && ptrs && sample && roots && mrrpe && cdev, "eval_metrics: bad arguments (B=%d)", B);
  EvalArgs a;
  for (int hnd = 0; hnd < 2; ++hnd) {
    a.h[hnd].vp = ptrs[hnd * 3]; a.h[hnd].vg = ptrs[hnd * 3 + 1]; a.h[hnd].J21 = ptrs[hnd * 3 + 2];
    RIH_REQUIRE(a.h[hnd].vp && a.h[hnd].vg && a.h[hnd].J21, "eval_metrics: null input pointer (hand %d)", hnd);
  }
  a.B = B; a.sample = sample; a.per_joint = per_joint; a.per_vert = per_vert; a.roots = roots;
  launch_k(eval_metrics_kernel, dim3(B, 2), EM_THREADS, 0, s, a);
  RIH_CUDA(cudaGetLastError());
  launch_k(eval_pair_kernel, B, EM_THREADS, 0, s, a.h[0].vp, a.h[1].vp, a.h[0].vg, a.h[1].vg, roots, B, contact_dist, mrrpe, cdev);
  RIH_CUDA(cudaGetLastError());
  return 0;
}
