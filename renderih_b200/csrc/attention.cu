// Fused multi-head attention core (SIMT fp32, exact softmax): O = dropout(softmax(Q K^T * scale)) V
// One CTA per (batch, head): the whole K/V of a head (<= 512 keys, d <= 64) lives in shared memory.
// reference: SelfAttn.self_attn  models/model_attn/self_attn.py:63-76 ; inter_attn.inter_attn  inter_attn.py:73-123
// Tensor addressing: element (b, row, h, dd) of T is  T + b*T_bs + row*ldT + h*d + dd.
#include "common.cuh"
using namespace rih;

constexpr int ATT_MAXJ = 16;   // keys per lane -> Sk <= 512
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(const float* __restrict__ q, long long q_bs, int ldq, const float* __restrict__ k, long long k_bs, int ldk,
                const float* __restrict__ v, long long v_bs, int ldv, float* __restrict__ o, long long o_bs, int ldo,
                float* __restrict__ lse, int H, int Sq, int Sk, int d, float scale, int rows_per_cta,
                const unsigned long long* __restrict__ seed_ptr, unsigned long long site, uint32_t thresh, float inv_keep) {
  const unsigned long long seed = thresh ? (*seed_ptr + site * 0xD1B54A32D192ED03ull) : 0ull;
  extern __shared__ float smem[];
  const int dp = d + 1;
  float* Ks = smem;                    // [Sk][d+1]
  float* Vs = Ks + (size_t)Sk * dp;    // [Sk][d]
  float* qs = Vs + (size_t)Sk * d;     // [W][d]
  float* ps = qs + ATT_WARPS * d;      // [W][Sk]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* kb = k + (size_t)b * k_bs + h * d;
  const float* vb = v + (size_t)b * v_bs + h * d;
  for (int i = tid; i < Sk * d; i += blockDim.x) {
    int j = i / d, dd = i - j * d;
    Ks[j * dp + dd] = kb[(size_t)j * ldk + dd];
    Vs[j * d + dd] = vb[(size_t)j * ldv + dd];
  }
  __syncthreads();
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(Sq, r0 + rows_per_cta);
  float* myq = qs + warp * d;
  float* myp = ps + (size_t)warp * Sk;
  for (int i = r0 + warp; i < r1; i += ATT_WARPS) {
    const float* qrow = q + (size_t)b * q_bs + (size_t)i * ldq + h * d;
    for (int dd = lane; dd < d; dd += 32) myq[dd] = qrow[dd];
    __syncwarp();
    float s[ATT_MAXJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      int j = lane + 32 * jj;
      s[jj] = -INFINITY;
      if (j < Sk) {
        float acc = 0.f;
        const float* kr = Ks + j * dp;
        for (int dd = 0; dd < d; ++dd) acc = fmaf(myq[dd], kr[dd], acc);
        s[jj] = acc * scale;
        mx = fmaxf(mx, s[jj]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      int j = lane + 32 * jj;
      if (j < Sk) { s[jj] = expf(s[jj] - mx); sum += s[jj]; }
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    size_t drop_base = ((size_t)bh * Sq + i) * Sk;
#pragma unroll
    for (int jj = 0; jj < ATT_MAXJ; ++jj) {
      int j = lane + 32 * jj;
      if (j < Sk) {
        float p = s[jj] * inv;
        if (thresh) p *= dropout_scale(seed, drop_base + j, thresh, inv_keep);
        myp[j] = p;
      }
    }
    if (lane == 0 && lse) lse[(size_t)bh * Sq + i] = mx + logf(sum);
    __syncwarp();
    float* orow = o + (size_t)b * o_bs + (size_t)i * ldo + h * d;
    for (int dd = lane; dd < d; dd += 32) {
      float acc = 0.f;
      for (int j = 0; j < Sk; ++j) acc = fmaf(myp[j], Vs[j * d + dd], acc);
      orow[dd] = acc;
    }
    __syncwarp();
  }
}

RIH_API int rih_attn_fwd(const float* q, long long q_bs, int ldq, const float* k, long long k_bs, int ldk,
                         const float* v, long long v_bs, int ldv, float* o, long long o_bs, int ldo, float* lse,
                         int B, int H, int Sq, int Sk, int d, float scale, float dropout_p, const unsigned long long* seed_ptr, unsigned long long site, cudaStream_t s) {
  RIH_REQUIRE(Sk > 0 && Sk <= 32 * ATT_MAXJ, "attn_fwd: Sk=%d unsupported (max %d)", Sk, 32 * ATT_MAXJ);
  RIH_REQUIRE(d > 0 && d <= 128, "attn_fwd: head dim %d unsupported", d);
  if (B * H == 0 || Sq == 0) return 0;
  size_t smem = sizeof(float) * ((size_t)Sk * (d + 1) + (size_t)Sk * d + ATT_WARPS * d + (size_t)ATT_WARPS * Sk);
  RIH_REQUIRE(smem <= 227 * 1024, "attn_fwd: shared memory %zu too large", smem);
  static bool attr_set = false;
  if (!attr_set) { RIH_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; }
  int ysplit = 1;
  while (B * H * ysplit < 148 * 2 && Sq / (ysplit * 2) >= 16) ysplit *= 2;
  int rows_per_cta = cdiv(Sq, ysplit);
  dim3 grid(B * H, cdiv(Sq, rows_per_cta));
  uint32_t thresh = dropout_p > 0.f ? dropout_thresh(dropout_p) : 0u;
  attn_fwd_kernel<<<grid, ATT_WARPS * 32, smem, s>>>(q, q_bs, ldq, k, k_bs, ldk, v, v_bs, ldv, o, o_bs, ldo, lse, H, Sq, Sk, d, scale,
                                                     rows_per_cta, seed_ptr, site, thresh, dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f);
  return check_launch("attn_fwd");
}

// ------------------------------------------------------------------ backward
// Phase A (warp per query): dQ.  Phase B (warp per key): dK, dV.  No atomics, deterministic.
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_kernel(const float* __restrict__ q, long long q_bs, int ldq, const float* __restrict__ k, long long k_bs, int ldk,
                const float* __restrict__ v, long long v_bs, int ldv, const float* __restrict__ o, long long o_bs, int ldo,
                const float* __restrict__ dout, long long do_bs, int lddo, const float* __restrict__ lse,
                float* __restrict__ dq, long long dq_bs, int lddq, float* __restrict__ dk, long long dk_bs, int lddk,
                float* __restrict__ dv, long long dv_bs, int lddv,
                int H, int Sq, int Sk, int d, float scale, const unsigned long long* __restrict__ seed_ptr, unsigned long long site, uint32_t thresh, float inv_keep) {
  const unsigned long long seed = thresh ? (*seed_ptr + site * 0xD1B54A32D192ED03ull) : 0ull;
  extern __shared__ float smem[];
  const int dp = d + 1;
  const int Smax = max(Sq, Sk);
  float* Qs = smem;                       // [Sq][dp]
  float* dOs = Qs + (size_t)Sq * dp;      // [Sq][dp]
  float* Ks = dOs + (size_t)Sq * dp;      // [Sk][dp]
  float* Vs = Ks + (size_t)Sk * dp;       // [Sk][dp]
  float* Dl = Vs + (size_t)Sk * dp;       // [Sq]
  float* Ls = Dl + Sq;                    // [Sq]
  float* w1 = Ls + Sq;                    // [W][Smax]
  float* w2 = w1 + (size_t)ATT_WARPS * Smax;  // [W][Smax]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* qb = q + (size_t)b * q_bs + h * d;
  const float* dob = dout + (size_t)b * do_bs + h * d;
  const float* kb = k + (size_t)b * k_bs + h * d;
  const float* vb = v + (size_t)b * v_bs + h * d;
  for (int i = tid; i < Sq * d; i += blockDim.x) {
    int r = i / d, dd = i - r * d;
    Qs[r * dp + dd] = qb[(size_t)r * ldq + dd];
    dOs[r * dp + dd] = dob[(size_t)r * lddo + dd];
  }
  for (int i = tid; i < Sk * d; i += blockDim.x) {
    int r = i / d, dd = i - r * d;
    Ks[r * dp + dd] = kb[(size_t)r * ldk + dd];
    Vs[r * dp + dd] = vb[(size_t)r * ldv + dd];
  }
  // D_i = sum_dd dO_i * O_i
  for (int i = warp; i < Sq; i += ATT_WARPS) {
    const float* orow = o + (size_t)b * o_bs + (size_t)i * ldo + h * d;
    const float* drow = dob + (size_t)i * lddo;
    float acc = 0.f;
    for (int dd = lane; dd < d; dd += 32) acc += orow[dd] * drow[dd];
    acc = warp_sum(acc);
    if (lane == 0) { Dl[i] = acc; Ls[i] = lse[(size_t)bh * Sq + i]; }
  }
  __syncthreads();
  float* my1 = w1 + (size_t)warp * Smax;
  float* my2 = w2 + (size_t)warp * Smax;
  // ---- phase A: dQ_i = scale * sum_j dS_ij K_j
  for (int i = warp; i < Sq; i += ATT_WARPS) {
    const float* qr = Qs + i * dp;
    const float* dr = dOs + i * dp;
    float Li = Ls[i], Di = Dl[i];
    size_t drop_base = ((size_t)bh * Sq + i) * Sk;
    for (int j = lane; j < Sk; j += 32) {
      const float* kr = Ks + j * dp;
      const float* vr = Vs + j * dp;
      float sc = 0.f, dpt = 0.f;
      for (int dd = 0; dd < d; ++dd) { sc = fmaf(qr[dd], kr[dd], sc); dpt = fmaf(dr[dd], vr[dd], dpt); }
      float p = expf(sc * scale - Li);
      float m = thresh ? dropout_scale(seed, drop_base + j, thresh, inv_keep) : 1.f;
      my1[j] = p * (m * dpt - Di);
    }
    __syncwarp();
    float* dqr = dq + (size_t)b * dq_bs + (size_t)i * lddq + h * d;
    for (int dd = lane; dd < d; dd += 32) {
      float acc = 0.f;
      for (int j = 0; j < Sk; ++j) acc = fmaf(my1[j], Ks[j * dp + dd], acc);
      dqr[dd] = acc * scale;
    }
    __syncwarp();
  }
  // ---- phase B: dV_j = sum_i Pt_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
  for (int j = warp; j < Sk; j += ATT_WARPS) {
    const float* kr = Ks + j * dp;
    const float* vr = Vs + j * dp;
    for (int i = lane; i < Sq; i += 32) {
      const float* qr = Qs + i * dp;
      const float* dr = dOs + i * dp;
      float sc = 0.f, dpt = 0.f;
      for (int dd = 0; dd < d; ++dd) { sc = fmaf(qr[dd], kr[dd], sc); dpt = fmaf(dr[dd], vr[dd], dpt); }
      float p = expf(sc * scale - Ls[i]);
      float m = thresh ? dropout_scale(seed, ((size_t)bh * Sq + i) * Sk + j, thresh, inv_keep) : 1.f;
      my1[i] = p * m;
      my2[i] = p * (m * dpt - Dl[i]);
    }
    __syncwarp();
    float* dkr = dk + (size_t)b * dk_bs + (size_t)j * lddk + h * d;
    float* dvr = dv + (size_t)b * dv_bs + (size_t)j * lddv + h * d;
    for (int dd = lane; dd < d; dd += 32) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < Sq; ++i) { av = fmaf(my1[i], dOs[i * dp + dd], av); ak = fmaf(my2[i], Qs[i * dp + dd], ak); }
      dkr[dd] = ak * scale;
      dvr[dd] = av;
    }
    __syncwarp();
  }
}

RIH_API int rih_attn_bwd(const float* q, long long q_bs, int ldq, const float* k, long long k_bs, int ldk,
                         const float* v, long long v_bs, int ldv, const float* o, long long o_bs, int ldo,
                         const float* dout, long long do_bs, int lddo, const float* lse,
                         float* dq, long long dq_bs, int lddq, float* dk, long long dk_bs, int lddk, float* dv, long long dv_bs, int lddv,
                         int B, int H, int Sq, int Sk, int d, float scale, float dropout_p, const unsigned long long* seed_ptr, unsigned long long site, cudaStream_t s) {
  RIH_REQUIRE(d > 0 && d <= 128, "attn_bwd: head dim %d unsupported", d);
  if (B * H == 0 || Sq == 0 || Sk == 0) return 0;
  int Smax = Sq > Sk ? Sq : Sk;
  size_t smem = sizeof(float) * ((size_t)(2 * Sq + 2 * Sk) * (d + 1) + 2 * (size_t)Sq + 2 * (size_t)ATT_WARPS * Smax);
  RIH_REQUIRE(smem <= 227 * 1024, "attn_bwd: shared memory %zu too large (Sq=%d Sk=%d d=%d)", smem, Sq, Sk, d);
  static bool attr_set = false;
  if (!attr_set) { RIH_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; }
  uint32_t thresh = dropout_p > 0.f ? dropout_thresh(dropout_p) : 0u;
  attn_bwd_kernel<<<B * H, ATT_WARPS * 32, smem, s>>>(q, q_bs, ldq, k, k_bs, ldk, v, v_bs, ldv, o, o_bs, ldo, dout, do_bs, lddo, lse,
                                                      dq, dq_bs, lddq, dk, dk_bs, lddk, dv, dv_bs, lddv, H, Sq, Sk, d, scale, seed_ptr, site, thresh,
                                                      dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f);
  return check_launch("attn_bwd");
}
