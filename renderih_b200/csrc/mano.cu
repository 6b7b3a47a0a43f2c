// MANO forward kinematics + linear blend skinning as ONE kernel (one CTA per hand).
// reference: ManoLayer.forward  models/manolayer.py:250-322 ; rodrigues_batch :32-48 ; pca2axis :163-166
// Stage plan inside the CTA (warp roles):
//   S1  all warps : PCA->axis-angle (45 dots), shape blend v_shaped = T + S*beta (2334 outputs, coalesced [10][2334] table)
//   S2  warps 0   : Rodrigues x15 ; warps 1.. : joint regression j_tpose = Jreg @ v_shaped (warp per joint row)
//   S3  warp 0    : serial FK chain over the kinematic tree ; warps 1.. : pose blend v_tpose += P*(R-I) ([135][2334] table)
//   S4  all warps : LBS per vertex (16 joint weights as 4x float4), results kept in shared memory
//   S5  warp 0    : joints (FK joints + 5 finger tips), reorder, centre
//   S6  all warps : centre / scale / trans, coalesced stores
#include "common.cuh"
using namespace rih;

struct ManoConsts {
  const float* comps;       // [45,45]  hands_components
  const float* hands_mean;  // [45]
  const float* shapedirsT;  // [10][2334]
  const float* posedirsT;   // [135][2334]
  const float* v_template;  // [2334]
  const float* jreg;        // [16,778]
  const float* weights;     // [778,16]
  int parent[16];
  int new_order[21];
  int tips[5];
};

constexpr int MANO_V = 778, MANO_V3 = 2334, MANO_THREADS = 512;

__global__ void __launch_bounds__(MANO_THREADS)
mano_fwd_kernel(ManoConsts mc, const float* __restrict__ root_rot, const float* __restrict__ pose, int use_pca, int ncomps,
                const float* __restrict__ shape, const float* __restrict__ trans, const float* __restrict__ scale,
                int center_idx, int new_skel, float* __restrict__ v_out, float* __restrict__ j_out) {
  __shared__ float s_v[MANO_V3];        // v_shaped -> v_tpose -> v_output
  __shared__ float s_axis[48];
  __shared__ float s_R[16][9];          // [0] = root, [1..15] = pose rotations
  __shared__ float s_ps[136];           // vec(R - I)
  __shared__ float s_beta[10];
  __shared__ float s_j[16][3];          // j_tpose
  __shared__ float s_T[16][12];         // global SE3 (top 3 rows)
  __shared__ float s_jo[21][3];         // joints before reorder
  __shared__ float s_jf[21][3];         // final joints
  __shared__ float s_center[3];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = MANO_THREADS / 32;

  if (tid < 10) s_beta[tid] = shape[(size_t)b * 10 + tid];
  if (tid < 9) s_R[0][tid] = root_rot[(size_t)b * 9 + tid];
  __syncthreads();
  // ---- S1
  if (use_pca) {
    if (tid < 45) {
      float acc = 0.f;
      for (int i = 0; i < ncomps; ++i) acc = fmaf(pose[(size_t)b * ncomps + i], mc.comps[i * 45 + tid], acc);
      s_axis[tid] = acc + mc.hands_mean[tid];
    }
  } else {
    if (tid < 135) s_R[1 + tid / 9][tid % 9] = pose[(size_t)b * 135 + tid];
  }
  for (int o = tid; o < MANO_V3; o += MANO_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) acc = fmaf(mc.shapedirsT[i * MANO_V3 + o], s_beta[i], acc);
    s_v[o] = mc.v_template[o] + acc;
  }
  __syncthreads();
  // ---- S2
  if (warp == 0) {
    if (use_pca && lane < 15) {
      float ax = s_axis[lane * 3], ay = s_axis[lane * 3 + 1], az = s_axis[lane * 3 + 2];
      float angle = sqrtf(ax * ax + ay * ay + az * az) + 1e-8f;   // manolayer.py:37 (eps added after the norm)
      float x = ax / angle, y = ay / angle, z = az / angle;
      float sn = sinf(angle), cs = cosf(angle), oc = 1.f - cs;
      // L = [[0,-z,y],[z,0,-x],[-y,x,0]] ; R = I + sin*L + (1-cos)*L@L
      float* R = s_R[1 + lane];
      R[0] = 1.f + oc * (-(z * z) - (y * y)); R[1] = -sn * z + oc * (x * y);        R[2] = sn * y + oc * (x * z);
      R[3] = sn * z + oc * (x * y);         R[4] = 1.f + oc * (-(z * z) - (x * x)); R[5] = -sn * x + oc * (y * z);
      R[6] = -sn * y + oc * (x * z);        R[7] = sn * x + oc * (y * z);           R[8] = 1.f + oc * (-(y * y) - (x * x));
    }
  } else {
    for (int o = warp - 1; o < 48; o += nwarps - 1) {
      int jt = o / 3, c = o % 3;
      float acc = 0.f;
      for (int vtx = lane; vtx < MANO_V; vtx += 32) acc = fmaf(mc.jreg[jt * MANO_V + vtx], s_v[vtx * 3 + c], acc);
      acc = warp_sum(acc);
      if (lane == 0) s_j[jt][c] = acc;
    }
  }
  __syncthreads();
  if (tid < 135) { int k = tid % 9; s_ps[tid] = s_R[1 + tid / 9][k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f); }
  __syncthreads();
  // ---- S3
  if (warp == 0) {
    if (lane == 0) {
      // local transform i: R_i, t_i = (I - R_i) j_i ; global = global[parent] o local  (manolayer.py:274-283)
      for (int i = 0; i < 16; ++i) {
        const float* R = s_R[i];
        float jx = s_j[i][0], jy = s_j[i][1], jz = s_j[i][2];
        float t[3];
        t[0] = (1.f - R[0]) * jx + (-R[1]) * jy + (-R[2]) * jz;
        t[1] = (-R[3]) * jx + (1.f - R[4]) * jy + (-R[5]) * jz;
        t[2] = (-R[6]) * jx + (-R[7]) * jy + (1.f - R[8]) * jz;
        if (i == 0) {
          for (int r = 0; r < 3; ++r) { s_T[0][r * 4 + 0] = R[r * 3]; s_T[0][r * 4 + 1] = R[r * 3 + 1]; s_T[0][r * 4 + 2] = R[r * 3 + 2]; s_T[0][r * 4 + 3] = t[r]; }
        } else {
          const float* P = s_T[mc.parent[i]];
          for (int r = 0; r < 3; ++r) {
            float p0 = P[r * 4], p1 = P[r * 4 + 1], p2 = P[r * 4 + 2], p3 = P[r * 4 + 3];
            s_T[i][r * 4 + 0] = p0 * R[0] + p1 * R[3] + p2 * R[6];
            s_T[i][r * 4 + 1] = p0 * R[1] + p1 * R[4] + p2 * R[7];
            s_T[i][r * 4 + 2] = p0 * R[2] + p1 * R[5] + p2 * R[8];
            s_T[i][r * 4 + 3] = p0 * t[0] + p1 * t[1] + p2 * t[2] + p3;
          }
        }
      }
    }
  } else {
    for (int o = tid - 32; o < MANO_V3; o += MANO_THREADS - 32) {
      float acc = 0.f;
#pragma unroll 9
      for (int i = 0; i < 135; ++i) acc = fmaf(mc.posedirsT[i * MANO_V3 + o], s_ps[i], acc);
      s_v[o] += acc;
    }
  }
  __syncthreads();
  // ---- S4: LBS
  for (int vtx = tid; vtx < MANO_V; vtx += MANO_THREADS) {
    const float4* w4 = reinterpret_cast<const float4*>(mc.weights + (size_t)vtx * 16);
    float w[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) { float4 t = __ldg(w4 + i); w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w; }
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
    for (int jt = 0; jt < 16; ++jt)
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = fmaf(w[jt], s_T[jt][e], T[e]);
    float x = s_v[vtx * 3], y = s_v[vtx * 3 + 1], z = s_v[vtx * 3 + 2];
    float ox = T[0] * x + T[1] * y + T[2] * z + T[3];
    float oy = T[4] * x + T[5] * y + T[6] * z + T[7];
    float oz = T[8] * x + T[9] * y + T[10] * z + T[11];
    s_v[vtx * 3] = ox; s_v[vtx * 3 + 1] = oy; s_v[vtx * 3 + 2] = oz;   // own element only: no hazard
  }
  __syncthreads();
  // ---- S5: joints
  if (tid < 21) {
    float jx, jy, jz;
    if (tid == 0) { jx = s_j[0][0]; jy = s_j[0][1]; jz = s_j[0][2]; }
    else if (tid < 16) {
      const float* P = s_T[mc.parent[tid]];
      float x = s_j[tid][0], y = s_j[tid][1], z = s_j[tid][2];
      jx = P[0] * x + P[1] * y + P[2] * z + P[3];
      jy = P[4] * x + P[5] * y + P[6] * z + P[7];
      jz = P[8] * x + P[9] * y + P[10] * z + P[11];
    } else {
      int vtx = mc.tips[tid - 16];
      jx = s_v[vtx * 3]; jy = s_v[vtx * 3 + 1]; jz = s_v[vtx * 3 + 2];
    }
    s_jo[tid][0] = jx; s_jo[tid][1] = jy; s_jo[tid][2] = jz;
  }
  __syncthreads();
  if (tid < 21) {
    int src = mc.new_order[tid];
    s_jf[tid][0] = s_jo[src][0]; s_jf[tid][1] = s_jo[src][1]; s_jf[tid][2] = s_jo[src][2];
  }
  __syncthreads();
  if (tid < 3) s_center[tid] = center_idx >= 0 ? s_jf[center_idx][tid] : 0.f;
  __syncthreads();
  // ---- S6: outputs
  float sc = scale ? scale[b] : 1.f;
  float tr[3] = {0.f, 0.f, 0.f};
  if (trans) { tr[0] = trans[(size_t)b * 3]; tr[1] = trans[(size_t)b * 3 + 1]; tr[2] = trans[(size_t)b * 3 + 2]; }
  for (int o = tid; o < MANO_V3; o += MANO_THREADS) {
    int c = o % 3;
    float val = s_v[o];
    if (center_idx >= 0) val = val - s_center[c];
    if (scale) val = val * sc;
    if (trans) val = val + tr[c];
    s_v[o] = val;
    v_out[(size_t)b * MANO_V3 + o] = val;
  }
  __syncthreads();
  if (tid < 63) {
    int jt = tid / 3, c = tid % 3;
    float val = s_jf[jt][c];
    if (center_idx >= 0) val = val - s_center[c];
    if (scale) val = val * sc;
    if (trans) val = val + tr[c];
    if (new_skel) {   // manolayer.py:316-320
      if (jt == 5) val = (s_v[63 * 3 + c] + s_v[144 * 3 + c]) / 2.f;
      else if (jt == 9) val = (s_v[271 * 3 + c] + s_v[220 * 3 + c]) / 2.f;
      else if (jt == 13) val = (s_v[148 * 3 + c] + s_v[290 * 3 + c]) / 2.f;
      else if (jt == 17) val = (s_v[770 * 3 + c] + s_v[83 * 3 + c]) / 2.f;
    }
    j_out[(size_t)b * 63 + tid] = val;
  }
}

// consts: 7 device pointers in the order of ManoConsts; parent[16]
RIH_API int rih_mano_fwd(const float* const* const_ptrs, const int* parent, const float* root_rot, const float* pose, int use_pca, int ncomps,
                         const float* shape, const float* trans, const float* scale, int center_idx, int new_skel,
                         float* v_out, float* j_out, int B, cudaStream_t s) {
  RIH_REQUIRE(B >= 0, "mano_fwd: negative batch");
  RIH_REQUIRE(!use_pca || (ncomps >= 0 && ncomps <= 45), "mano_fwd: ncomps=%d out of range", ncomps);
  RIH_REQUIRE(center_idx < 21, "mano_fwd: center_idx=%d out of range", center_idx);
  if (B == 0) return 0;
  ManoConsts mc;
  mc.comps = const_ptrs[0]; mc.hands_mean = const_ptrs[1]; mc.shapedirsT = const_ptrs[2]; mc.posedirsT = const_ptrs[3];
  mc.v_template = const_ptrs[4]; mc.jreg = const_ptrs[5]; mc.weights = const_ptrs[6];
  for (int i = 0; i < 16; ++i) {
    mc.parent[i] = parent[i];
    RIH_REQUIRE(i == 0 || (parent[i] >= 0 && parent[i] < i), "mano_fwd: kinematic parent table is not topologically ordered");
  }
  static const int order[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};  // manolayer.py:110-115
  static const int tips[5] = {745, 317, 444, 556, 673};                                                    // manolayer.py:296
  for (int i = 0; i < 21; ++i) mc.new_order[i] = order[i];
  for (int i = 0; i < 5; ++i) mc.tips[i] = tips[i];
  mano_fwd_kernel<<<B, MANO_THREADS, 0, s>>>(mc, root_rot, pose, use_pca, ncomps, shape, trans, scale, center_idx, new_skel, v_out, j_out);
  return check_launch("mano_fwd");
}
