// Host launchers + C-ABI for the tcgen05 TF32 GEMM / implicit-GEMM convolution (see gemm_tc.cuh).
#include "gemm_tc.cuh"
#include "bgemm.cuh"

namespace rih {
namespace tc {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4)) { set_error("tensor map: base/ld not 16-byte aligned"); return 1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);   // this thread (e.g. an autograd worker) has no driver context bound yet: let the runtime bind the primary one
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld); return 1; }
  return 0;
}

int make_tmap_2d_grouped(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int groups) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) || (cols % 32)) { set_error("tensor map (grouped): base/ld not 16-byte aligned or cols %% 32 != 0"); return 1; }
  cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128u};
  cuuint32_t box[3] = {32u, 32u, (cuuint32_t)groups};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(grouped) failed (%d) rows=%lld cols=%lld ld=%lld groups=%d", (int)r, rows, cols, ld, groups); return 1; }
  return 0;
}

int make_tmap_nhwc(CUtensorMap* map, const float* base, int N, int H, int W, int C, long long ld, int bw, int bh, int bn, bool atom32, int estride) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4)) { set_error("tensor map: base/ld not 16-byte aligned"); return 1; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
  // with a traversal stride e the box extent is given in traversed elements: bw * e positions yield bw fetched pixels
  cuuint32_t box[4] = {32u, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), (cuuint32_t)bn};
  cuuint32_t estr[4] = {1u, (cuuint32_t)estride, (cuuint32_t)estride, 1u};
  if (box[1] > 256u || box[2] > 256u) { set_error("tensor map (nhwc, element stride %d): box %u x %u exceeds 256", estride, box[1], box[2]); return 1; }
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(4d) failed (%d) N=%d H=%d W=%d C=%d ld=%lld box=%d,%d,%d", (int)r, N, H, W, C, ld, bw, bh, bn); return 1; }
  return 0;
}

int make_tmap_nhwc_grouped(CUtensorMap* map, const float* base, int N, int H, int W, int C, long long ld, int bw, int bh, int estride, int groups) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) || (C % 32)) { set_error("tensor map (nhwc grouped): base/ld not 16-byte aligned or C %% 32 != 0"); return 1; }
  cuuint64_t dims[5] = {32u, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / 32)};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4, 128u};
  cuuint32_t box[5] = {32u, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), 1u, (cuuint32_t)groups};
  cuuint32_t estr[5] = {1u, (cuuint32_t)estride, (cuuint32_t)estride, 1u, 1u};
  if (box[1] > 256u || box[2] > 256u) { set_error("tensor map (nhwc grouped, element stride %d): box %u x %u exceeds 256", estride, box[1], box[2]); return 1; }
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(5d grouped) failed (%d) N=%d H=%d W=%d C=%d ld=%lld box=%d,%d groups=%d", (int)r, N, H, W, C, ld, bw, bh, groups); return 1; }
  return 0;
}

int make_tmap_stem(CUtensorMap* map, const float* base, int N, int Hp, int Wp, int Wo, int box_ow, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if (reinterpret_cast<uintptr_t>(base) & 15) { set_error("tensor map (stem): base not 16-byte aligned"); return 1; }
  cuuint64_t dims[4] = {32u, (cuuint64_t)Wo, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {32u, (cuuint64_t)Wp * 16u, (cuuint64_t)Hp * Wp * 16u};     // bytes: 2 padded pixels, one padded row, one padded image
  cuuint32_t box[4] = {32u, (cuuint32_t)box_ow, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(stem, overlapping strides) failed (%d) N=%d Hp=%d Wp=%d Wo=%d", (int)r, N, Hp, Wp, Wo); return 1; }
  return 0;
}

// wgrad output dW[Cout][taps][Cin] as a 3-D map, box {32 channels, 1 tap, 128 filters}: a channel chunk is clipped at its own tap
int make_tmap_wgrad_out(CUtensorMap* map, const float* base, int Cout, int taps, int Cin) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (Cin % 4)) { set_error("tensor map (wgrad out): base/Cin not 16-byte aligned"); return 1; }
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)taps, (cuuint64_t)Cout};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)taps * Cin * 4};
  cuuint32_t box[3] = {32u, 1u, (cuuint32_t)BM};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad out) failed (%d) Cout=%d taps=%d Cin=%d", (int)r, Cout, taps, Cin); return 1; }
  return 0;
}

int g_stats_fused = 0;            // set by the launcher when the epilogue accumulated ep.stats (fused BatchNorm statistics)
static int g_wide_tiles = 1;     // 128 x 256 output tiles when N % 256 == 0
void set_wide_tiles(int on) { g_wide_tiles = on ? 1 : 0; }
static int g_wgrad_wide = 3;     // 256-wide multi-tap N tiles for the convolution weight gradients, also over padded channel counts (rih_set_wgrad_wide)
void set_wgrad_wide(int on) { g_wgrad_wide = on & 3; }     // bit 0: multi-tap 256-wide tiles, bit 1: also for channel counts that are not multiples of 32
static int g_narrow_small = 1;   // 64-wide N tiles for GEMMs whose 128-wide tiling would leave half of the SMs idle (RIH_NARROW_TILES=0 to compare)
void set_narrow_small(int on) { g_narrow_small = on ? 1 : 0; }
static int g_persistent = 1;     // 0 = one CTA per tile (non-persistent kernel; debug / comparison)
void set_persistent(int on) { g_persistent = on ? 1 : 0; }
static int g_tma_epilogue = 1;   // 0 = per-thread global stores (debug / comparison)
void set_tma_epilogue(int on) { g_tma_epilogue = on ? 1 : 0; }
long long g_launch_counts[3] = {0, 0, 0};   // tensor-core launches with the TMA-store epilogue / with per-thread stores / SIMT GEMM launches (rih_gemm_launch_counts)
static int g_tma_grouped = 7;    // grouped rank-3 tensor maps for MN-major operands (rih_set_tma_grouped)
void set_tma_grouped(int on) { g_tma_grouped = on & 7; }       // bit 0: rank-3 grouped maps for dense MN-major operands / dY, bit 1: rank-5 grouped input maps of the conv wgrad, bit 2: grouped weight boxes of the conv dgrad
static int g_epi_opt = 7;        // Epilogue::opt of every launch (rih_set_epilogue_opt)
void set_epilogue_opt(int v) { g_epi_opt = v & 7; }
static int g_tma_res = 1;        // residual rows read by the TMA-store epilogue (0: GEMMs with a residual use per-thread global stores, as before)
void set_tma_res(int on) { g_tma_res = on ? 1 : 0; }
template <int BN, bool A_MN, bool B_MN, class Producer, int NSPLIT>
static int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, Epilogue ep, const Producer& prod, int M, int N, int num_kb, int splits,
                      int kb_per_split, cudaStream_t s, const CUtensorMap* c_map = nullptr) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, Producer, NSPLIT>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, NSPLIT>()) != cudaSuccess) {
      set_error("gemm_tc: cannot raise dynamic shared memory to %d", smem_bytes<BN, NSPLIT>());
      return 2;
    }
    attr = true;
  }
  dim3 grid(cdiv(N, BN), cdiv(M, BM), splits);
  // TMA-store epilogue whenever the output is a 16-byte aligned matrix (a residual is read back row-wise by the epilogue threads when it is aligned too)
  CUtensorMap tc_;
  int tma_epi = 0;
  if (c_map) {       // caller-built output map (3-D wgrad view); only the persistent TMA-store epilogue understands it
    if (!g_persistent) { set_error("gemm_tc: output-map override needs the persistent kernel"); return 1; }
    tc_ = *c_map;
    tma_epi = 1;
  } else if ((!ep.res || (g_tma_res && splits == 1 && ((reinterpret_cast<uintptr_t>(ep.res) & 15) == 0) && ep.ldres % 4 == 0 && N % 4 == 0)) &&
             ((reinterpret_cast<uintptr_t>(ep.c) & 15) == 0) && (ep.ldc % 4 == 0) && g_tma_epilogue) {
    if (make_tmap_2d(&tc_, ep.c, M, N, ep.ldc, BM)) return 1;
    tma_epi = 1;
  } else {
    tc_ = ta;
  }
  if (ep.col_scale && ep.bias) { set_error("gemm_tc: a folded per-column affine and a bias cannot be combined (fold the bias into the shift)"); return 1; }
  g_launch_counts[tma_epi ? 0 : 1]++;
  if (g_persistent) {
    auto pk = gemm_tc_persistent_kernel<BN, A_MN, B_MN, Producer, NSPLIT>;
    static bool pattr = false;
    if (!pattr) {
      if (cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, NSPLIT>()) != cudaSuccess) {
        set_error("gemm_tc(persistent): cannot raise dynamic shared memory to %d", smem_bytes<BN, NSPLIT>());
        return 2;
      }
      pattr = true;
    }
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
    const int tiles_n = cdiv(N, BN), tiles_m = cdiv(M, BM);
    const long long total = (long long)tiles_n * tiles_m * splits;
    const int cap = stream_cta_limit(s, num_sms);
    const int ctas = (int)(total < cap ? total : cap);
    if (ep.stats && !(tma_epi && splits == 1)) ep.stats = nullptr;
    g_stats_fused = ep.stats != nullptr;
    launch_k(pk, dim3(ctas), dim3(persistent_threads<NSPLIT>()), (size_t)smem_bytes<BN, NSPLIT>(), s, ta, tb, tc_, ep, prod, num_kb, kb_per_split, tiles_m, tiles_n, splits, tma_epi);
    return check_launch("gemm_tc_persistent");
  }
  g_stats_fused = 0;
  launch_k(kern, grid, dim3(THREADS), (size_t)smem_bytes<BN, NSPLIT>(), s, ta, tb, tc_, ep, prod, num_kb, kb_per_split, tma_epi);
  return check_launch("gemm_tc");
}
static int g_nsplit = 1;   // set per call by the dispatchers below (1 = TF32, 2 = TF32 rn, 3 = 3xTF32)
static float g_scale = 1.f; // accumulator scale (bias-compensated truncating TF32)
void set_acc_scale(float s) { g_scale = s; }
template <int BN, bool A_MN, bool B_MN, class Producer>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, Epilogue ep, const Producer& prod, int M, int N, int num_kb, int splits,
                      int kb_per_split, cudaStream_t s, const CUtensorMap* c_map = nullptr) {
  ep.scale = g_scale;
  ep.reverse = g_reverse;
  ep.kb_rotate = g_kb_rotate;
  ep.a_policy = g_l2_hints ? 1ull : 0ull;
  ep.opt = g_epi_opt;
  if (g_nsplit == 3) return launch_one<BN, A_MN, B_MN, Producer, 3>(ta, tb, ep, prod, M, N, num_kb, splits, kb_per_split, s, c_map);
  if (g_nsplit == 2 && g_persistent) return launch_one<BN, A_MN, B_MN, Producer, 2>(ta, tb, ep, prod, M, N, num_kb, splits, kb_per_split, s, c_map);
  return launch_one<BN, A_MN, B_MN, Producer, 1>(ta, tb, ep, prod, M, N, num_kb, splits, kb_per_split, s, c_map);
}
void set_nsplit(int n) { g_nsplit = (n == 3) ? 3 : (n == 2 ? 2 : 1); }

// split-K planning over k-blocks of 32; switches the epilogue to atomic accumulation when splitting
static void plan_splitk(Epilogue& ep, int M, int N, int BN, int num_kb, int allow, int& splits, int& kb_per_split, cudaStream_t s, int N_grid = 0) {
  // N = real width of the output rows (what a split-K run has to zero); N_grid = width of the tile grid when it differs (padded wgrad taps)
  long long tiles = (long long)cdiv(M, BM) * cdiv(N_grid ? N_grid : N, BN);
  splits = 1;
  if (allow && tiles < 148 && num_kb >= 32) {
    splits = (int)((296 + tiles - 1) / tiles);
    int maxs = num_kb / 4;      // >= 4 k-blocks (128 reduction rows) per slice: these GEMMs are latency bound, so many short slices win
    if (splits > maxs) splits = maxs;
    if (splits > 128) splits = 128;
    if (splits < 1) splits = 1;
  }
  kb_per_split = cdiv(num_kb, splits);
  splits = cdiv(num_kb, kb_per_split);
  if (splits > 1) {
    if (ep.mode == 0) {
      if (ep.ldc == N) cudaMemsetAsync(ep.c, 0, (size_t)M * N * sizeof(float), s);
      else cudaMemset2DAsync(ep.c, (size_t)ep.ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s);
    }
    ep.mode = 2;
  }
}

// a: K-major [M,K] (a_mn=0) or MN-major [K,M] (a_mn=1); b likewise with N.
int gemm_tf32(const float* a, long long lda, int a_mn, const float* b, long long ldb, int b_mn, Epilogue ep, int M, int N, int K,
              int allow_splitk, cudaStream_t s) {
  if (M <= 0 || N <= 0) return 0;
  CUtensorMap ta, tb;
  int BN = (N > 64) ? 128 : 64;
  // 128 x 256 tiles (full 512 TMEM columns with double buffering) when N allows: each A tile is fetched once per 256 columns
  if (g_persistent && g_wide_tiles && N % 256 == 0 && (long long)cdiv(M, BM) * (N / 256) >= 120) BN = 256;
  // Latency-bound GEMMs of the token decoder (a few dozen 128-row tiles, K <= 512): when 128-wide tiles would occupy at most half of the SMs,
  // 64-wide tiles put twice as many CTAs to work -- each splits / multiplies half the B tile and drains half the accumulator, which is what a
  // single-tile-per-CTA dependent chain (TMA -> split -> MMA -> epilogue) is made of.
  if (BN == 128 && g_persistent && g_narrow_small && (long long)cdiv(M, BM) * cdiv(N, 128) <= 74) BN = 64;
  // MN-major operands with a multiple of 32 columns: ONE grouped box per k-block instead of BM / 32 (BN / 32) four-KB boxes (rih_set_tma_grouped)
  const int a_grp = (a_mn && (g_tma_grouped & 1) && M % 32 == 0) ? 1 : 0, b_grp = (b_mn && (g_tma_grouped & 1) && N % 32 == 0) ? 1 : 0;
  if (a_mn ? (a_grp ? make_tmap_2d_grouped(&ta, a, K, M, lda, BM / 32) : make_tmap_2d(&ta, a, K, M, lda, 32, true)) : make_tmap_2d(&ta, a, M, K, lda, BM)) return 1;
  if (b_mn ? (b_grp ? make_tmap_2d_grouped(&tb, b, K, N, ldb, BN / 32) : make_tmap_2d(&tb, b, K, N, ldb, 32, true)) : make_tmap_2d(&tb, b, N, K, ldb, BN)) return 1;
  int num_kb = cdiv(K, BK), splits, kps;
  plan_splitk(ep, M, N, BN, num_kb, allow_splitk, splits, kps, s);
#define RIH_TC_CASE(bn, am, bm)                                                              \
  if (BN == bn && a_mn == am && b_mn == bm) {                                                \
    DenseProducer<bn, am != 0, bm != 0> prod{0, 0ull, a_grp, b_grp};                                             \
    return launch_cfg<bn, am != 0, bm != 0>(ta, tb, ep, prod, M, N, num_kb, splits, kps, s); \
  }
  RIH_TC_CASE(256, 0, 0) RIH_TC_CASE(256, 0, 1) RIH_TC_CASE(256, 1, 1)
  RIH_TC_CASE(128, 0, 0) RIH_TC_CASE(64, 0, 0)
  RIH_TC_CASE(128, 0, 1) RIH_TC_CASE(64, 0, 1)
  RIH_TC_CASE(128, 1, 1) RIH_TC_CASE(64, 1, 1)
#undef RIH_TC_CASE
  set_error("gemm_tf32: unsupported configuration");
  return 1;
}

// ---------------------------------------------------------------- batched per-head GEMMs of the attention core
static int make_tmap_batched(CUtensorMap* map, const BOperand& o, int B, int H, int box_rows, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return 1; }
  if ((reinterpret_cast<uintptr_t>(o.p) & 15) || (o.ld % 4) || (o.tok && o.cols % 4)) { set_error("attention operand: base / stride / head dim not 16-byte aligned"); return 1; }
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1u, 1u, 1u, 1u};
  if (o.tok) {   // element (b, h, s, k) at p + (b * S + s) * ld + h * d + k
    dims[0] = (cuuint64_t)o.cols; dims[1] = (cuuint64_t)H; dims[2] = (cuuint64_t)o.rows; dims[3] = (cuuint64_t)B;
    strides[0] = (cuuint64_t)o.cols * 4; strides[1] = (cuuint64_t)o.ld * 4; strides[2] = (cuuint64_t)o.rows * o.ld * 4;
    box[0] = 32u; box[1] = 1u; box[2] = (cuuint32_t)box_rows; box[3] = 1u;
  } else {       // element (z, r, c) at p + (z * rows + r) * ld + c
    dims[0] = (cuuint64_t)o.cols; dims[1] = (cuuint64_t)o.rows; dims[2] = (cuuint64_t)B * H; dims[3] = 1;
    strides[0] = (cuuint64_t)o.ld * 4; strides[1] = (cuuint64_t)o.rows * o.ld * 4; strides[2] = (cuuint64_t)B * H * o.rows * o.ld * 4;
    box[0] = 32u; box[1] = (cuuint32_t)box_rows; box[2] = 1u; box[3] = 1u;
  }
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(o.p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(batched, tok=%d) failed (%d) rows=%d cols=%d ld=%lld B=%d H=%d", o.tok, (int)r, o.rows, o.cols, o.ld, B, H); return 1; }
  return 0;
}

// C[z] = scale * op(A[z]) op(B[z])^T for every head z = b * H + h, as ONE persistent launch: the tile loop runs over (n, m, z).
//   a_mn / b_mn = 0: operand stored [rows = M|N, cols = K] (K-major); 1: stored [rows = K, cols = M|N] (MN-major).
int bgemm_tf32(const BOperand& a, int a_mn, const BOperand& b, int b_mn, const BOperand& c, int B, int H, int M, int N, int K, float scale, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0 || B * H <= 0) return 0;
  if (!g_persistent) { set_error("bgemm_tf32 needs the persistent kernel"); return 1; }
  const int BN = (N > 64) ? 128 : 64;
  CUtensorMap ta, tb, tcm;
  if (make_tmap_batched(&ta, a, B, H, a_mn ? 32 : BM, a_mn != 0)) return 1;
  if (make_tmap_batched(&tb, b, B, H, b_mn ? 32 : BN, b_mn != 0)) return 1;
  if (make_tmap_batched(&tcm, c, B, H, BM, false)) return 1;
  Epilogue ep = make_epilogue(const_cast<float*>(c.p), (int)c.ld, M, N, nullptr, 0, 0);
  ep.batch_heads = H | (c.tok ? 0 : (1 << 30));
  const int num_kb = cdiv(K, BK);
  const float keep = g_scale;
  g_scale = scale * g_scale;     // launch_cfg copies g_scale into ep.scale
  int rc = 1;
  bool matched = false;
#define RIH_BG_CASE(bn, am, bm)                                                                            \
  if (BN == bn && a_mn == am && b_mn == bm) {                                                              \
    BatchedProducer<bn, am != 0, bm != 0> prod{H, a.tok, b.tok}; matched = true;                            \
    rc = launch_cfg<bn, am != 0, bm != 0>(ta, tb, ep, prod, M, N, num_kb, B * H, num_kb, s, &tcm);           \
  }
  RIH_BG_CASE(128, 0, 0) RIH_BG_CASE(64, 0, 0) RIH_BG_CASE(128, 0, 1) RIH_BG_CASE(64, 0, 1) RIH_BG_CASE(128, 1, 1) RIH_BG_CASE(64, 1, 1)
#undef RIH_BG_CASE
  g_scale = keep;
  if (!matched) set_error("bgemm_tf32: unsupported operand majors (a_mn=%d b_mn=%d)", a_mn, b_mn);
  return rc;
}

static int g_s2_direct = 1;      // stride-2 convolutions through element-strided tensor maps (no parity-stacked / zero-inserted copies); rih_set_s2_direct
void set_s2_direct(int on) { g_s2_direct = on ? 1 : 0; }
int s2_direct() { return g_s2_direct; }
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Can the stride-1 convolution run on the tcgen05 implicit-GEMM path?
// stride 2 is supported for even H, W through a parity-stacked copy of the input (fwd, wgrad) and a zero-inserted copy of
// dY (dgrad); both need caller-provided workspace (conv_tc_workspace).
bool conv_tc_supported(const ConvGeom& g, int which /*0 fwd, 1 dgrad, 2 wgrad*/) {
  if (g.stride != 1 && g.stride != 2) return false;
  if (g.stride == 2 && ((g.H | g.W) & 1 || g.Ho * 2 != g.H || g.Wo * 2 != g.W || g.R != g.S || (g.R != 1 && g.R != 3) || g.pad != g.R / 2)) return false;
  // channel counts: multiples of 16 from 32 up (HRNet-w48's 48 / 96-wide branches included): a K block / N chunk that sticks out of the
  // tensor is zero-filled on load and clipped on store by the TMA unit
  if (g.Cin % 16 || g.Cout % 16 || g.Cin < 32 || g.Cout < 32) return false;
  if (g.ldx % 4 || g.ldy % 4) return false;
  if (which == 0) return pow2(g.Wo) && pow2(g.Ho) && g.Wo <= 128 && ((long long)g.N * g.Ho * g.Wo) % BM == 0 && (g.Wo * g.Ho >= BM || BM % (g.Wo * g.Ho) == 0);
  if (which == 1 && g.stride == 2 && g_s2_direct)     // four parity-class GEMMs over the half-resolution (= output) grid
    return pow2(g.Wo) && pow2(g.Ho) && g.Wo <= 128 && ((long long)g.N * g.Ho * g.Wo) % BM == 0 && (g.Wo * g.Ho >= BM || BM % (g.Wo * g.Ho) == 0);
  if (which == 1) return pow2(g.W) && pow2(g.H) && g.W <= 128 && ((long long)g.N * g.H * g.W) % BM == 0 && (g.W * g.H >= BM || BM % (g.W * g.H) == 0);
  return pow2(g.Wo) && pow2(g.Ho) && (g.Ho * g.Wo) % 32 == 0 && (g.Cin % 64 == 0 || g_persistent);
}

long long conv_tc_workspace(const ConvGeom& g, int which) {
  if (g.stride != 2 || g_s2_direct) return 0;
  if (which == 1) return (long long)g.N * g.H * g.W * g.Cout;   // zero-inserted dY
  return (long long)g.N * g.H * g.W * g.Cin;                    // parity-stacked x
}
extern "C" int rih_parity_stack(const float* x, int ldx, float* xp, int N, int H, int W, int C, cudaStream_t s);
extern "C" int rih_dilate2x(const float* y, int ldy, float* yd, int N, int Ho, int Wo, int C, cudaStream_t s);

int conv_fwd_tf32(const float* x, const float* w, Epilogue ep, const ConvGeom& g, float* ws, cudaStream_t s) {
  const int M = g.N * g.Ho * g.Wo;
  int BN = (g.Cout > 64) ? 128 : 64;
  if (g_persistent && g_wide_tiles && g.Cout % 256 == 0 && (long long)(M / BM) * (g.Cout / 256) >= 120) BN = 256;
  const int tile_h = (BM / g.Wo) < g.Ho ? (BM / g.Wo) : g.Ho;
  const int tile_n = BM / (g.Wo * tile_h);
  CUtensorMap ta, tb;
  if (g.stride == 2 && g_s2_direct) {
    if (make_tmap_nhwc(&ta, x, g.N, g.H, g.W, g.Cin, g.ldx, g.Wo, tile_h, tile_n, false, 2)) return 1;
  } else if (g.stride == 2) {
    if (!ws) { set_error("conv_fwd_tf32: stride-2 path needs workspace"); return 1; }
    if (int e = rih_parity_stack(x, g.ldx, ws, g.N, g.H, g.W, g.Cin, s)) return e;
    if (make_tmap_nhwc(&ta, ws, 4 * g.N, g.H / 2, g.W / 2, g.Cin, g.Cin, g.Wo, tile_h, tile_n, false)) return 1;
  } else {
    if (make_tmap_nhwc(&ta, x, g.N, g.H, g.W, g.Cin, g.ldx, g.Wo, tile_h, tile_n, false)) return 1;
  }
  if (make_tmap_2d(&tb, w, g.Cout, (long long)g.R * g.S * g.Cin, (long long)g.R * g.S * g.Cin, BN)) return 1;
  ConvTcGeom cg{g.Cin, g.Cout, g.R, g.S, g.pad, g.H, g.W, g.Ho, g.Wo, tile_h, g.stride == 2 ? (g_s2_direct ? -1 : g.N) : 0, g.Cin};
  const int num_kb = g.R * g.S * cdiv(g.Cin, BK);
  if (BN == 256) { ConvFwdProducer<256> p{cg, 0ull}; return launch_cfg<256, false, false>(ta, tb, ep, p, M, g.Cout, num_kb, 1, num_kb, s); }
  if (BN == 128) { ConvFwdProducer<128> p{cg, 0ull}; return launch_cfg<128, false, false>(ta, tb, ep, p, M, g.Cout, num_kb, 1, num_kb, s); }
  ConvFwdProducer<64> p{cg, 0ull};
  return launch_cfg<64, false, false>(ta, tb, ep, p, M, g.Cout, num_kb, 1, num_kb, s);
}

// stride-2 dgrad as four dense parity-class GEMMs over the half-resolution grid (ConvDgradS2Producer), stored / accumulated in place through an
// element-strided map of dx.  Classes without a contributing tap (1x1 convolutions: three of four) are zero-filled unless the caller accumulates.
static int conv_dgrad_s2_direct(const float* dy, const float* w, Epilogue ep, const ConvGeom& g, cudaStream_t s) {
  const int H2 = g.Ho, W2 = g.Wo;                       // class grid == output grid (even H, W)
  const int M = g.N * H2 * W2;
  int BN = (g.Cin > 64) ? 128 : 64;
  if (g_persistent && g_wide_tiles && g.Cin % 256 == 0 && (long long)(M / BM) * (g.Cin / 256) >= 120) BN = 256;
  const int tile_h = (BM / W2) < H2 ? (BM / W2) : H2;
  const int tile_n = BM / (W2 * tile_h);
  CUtensorMap ta, tb, tcm;
  if (make_tmap_nhwc(&ta, dy, g.N, g.Ho, g.Wo, g.Cout, g.ldy, W2, tile_h, tile_n, false)) return 1;
  // weights [Cout][R*S*Cin] as the MN-major B operand: one grouped box per k-block when Cin % 32 == 0 (rih_set_tma_grouped bit 2)
  const int b_grp = ((g_tma_grouped & 4) && g.Cin % 32 == 0) ? 1 : 0;
  if (b_grp ? make_tmap_2d_grouped(&tb, w, g.Cout, (long long)g.R * g.S * g.Cin, (long long)g.R * g.S * g.Cin, BN / 32)
            : make_tmap_2d(&tb, w, g.Cout, (long long)g.R * g.S * g.Cin, (long long)g.R * g.S * g.Cin, 32, true)) return 1;
  if (make_tmap_nhwc(&tcm, ep.c, g.N, g.H, g.W, g.Cin, ep.ldc, W2, tile_h, tile_n, false, 2)) return 1;
  const int accumulate = ep.mode != 0;
  if (accumulate) { set_error("conv_dgrad_s2_direct: accumulation through an element-strided map is not supported (use the copy-based path)"); return 1; }
  bool zeroed = false;
  for (int ph = 0; ph < 2; ++ph) for (int pw = 0; pw < 2; ++pw) {
    int rs[2], nr = 0, ss[2], ns = 0;
    for (int r = 0; r < g.R; ++r) if (((ph + g.pad - r) & 1) == 0) rs[nr++] = r;
    for (int q = 0; q < g.S; ++q) if (((pw + g.pad - q) & 1) == 0) ss[ns++] = q;
    if (nr * ns == 0) {
      if (!accumulate && !zeroed) {      // this class receives no gradient: zero the whole tensor once, the other classes overwrite their pixels
        if (ep.ldc == g.Cin) cudaMemsetAsync(ep.c, 0, (size_t)g.N * g.H * g.W * g.Cin * sizeof(float), s);
        else cudaMemset2DAsync(ep.c, (size_t)ep.ldc * sizeof(float), 0, (size_t)g.Cin * sizeof(float), (size_t)g.N * g.H * g.W, s);
        zeroed = true;
      }
      continue;
    }
  }
  for (int ph = 0; ph < 2; ++ph) for (int pw = 0; pw < 2; ++pw) {
    ConvTcGeom cg{g.Cin, g.Cout, g.R, g.S, g.pad, H2, W2, g.Ho, g.Wo, tile_h, 0, g.Cin};
    int nt = 0, tap[4], dr[4], ds[4];
    for (int r = 0; r < g.R; ++r) if (((ph + g.pad - r) & 1) == 0)
      for (int q = 0; q < g.S; ++q) if (((pw + g.pad - q) & 1) == 0) {
        if (nt >= 4) { set_error("conv_dgrad_s2: more than 4 taps per parity class (kernel %dx%d)", g.R, g.S); return 1; }
        tap[nt] = r * g.S + q; dr[nt] = (ph + g.pad - r) / 2; ds[nt] = (pw + g.pad - q) / 2; ++nt;
      }
    if (nt == 0) continue;
    Epilogue e2 = ep;
    e2.M = M; e2.s2_w2 = W2; e2.s2_h2 = H2; e2.s2_ph = ph; e2.s2_pw = pw;
    const int num_kb = nt * cdiv(g.Cout, BK);
    int rc;
#define RIH_S2_CASE(bn)                                                                                                        \
    { ConvDgradS2Producer<bn> p{cg, nt, {tap[0], tap[1], tap[2], tap[3]}, {dr[0], dr[1], dr[2], dr[3]}, {ds[0], ds[1], ds[2], ds[3]}, b_grp}; \
      rc = launch_cfg<bn, false, true>(ta, tb, e2, p, M, g.Cin, num_kb, 1, num_kb, s, &tcm); }
    if (BN == 256) RIH_S2_CASE(256) else if (BN == 128) RIH_S2_CASE(128) else RIH_S2_CASE(64)
#undef RIH_S2_CASE
    if (rc) return rc;
  }
  return 0;
}

int conv_dgrad_tf32(const float* dy, const float* w, Epilogue ep, const ConvGeom& g0, float* ws, cudaStream_t s) {
  ConvGeom g = g0;
  if (g0.stride == 2 && g_s2_direct) return conv_dgrad_s2_direct(dy, w, ep, g0, s);
  if (g0.stride == 2) {   // zero-insert dY, then it is a stride-1 dgrad over an H x W "output"
    if (!ws) { set_error("conv_dgrad_tf32: stride-2 path needs workspace"); return 1; }
    if (int e = rih_dilate2x(dy, g0.ldy, ws, g0.N, g0.Ho, g0.Wo, g0.Cout, s)) return e;
    dy = ws; g.stride = 1; g.Ho = g0.H; g.Wo = g0.W; g.ldy = g0.Cout;
  }
  const int M = g.N * g.H * g.W;
  int BN = (g.Cin > 64) ? 128 : 64;
  if (g_persistent && g_wide_tiles && g.Cin % 256 == 0 && (long long)(M / BM) * (g.Cin / 256) >= 120) BN = 256;
  const int tile_h = (BM / g.W) < g.H ? (BM / g.W) : g.H;
  const int tile_n = BM / (g.W * tile_h);
  CUtensorMap ta, tb;
  if (make_tmap_nhwc(&ta, dy, g.N, g.Ho, g.Wo, g.Cout, g.ldy, g.W, tile_h, tile_n, false)) return 1;
  // weights [Cout][R*S*Cin] as the MN-major B operand: one grouped box per k-block when Cin % 32 == 0 (rih_set_tma_grouped bit 2)
  const int b_grp = ((g_tma_grouped & 4) && g.Cin % 32 == 0) ? 1 : 0;
  if (b_grp ? make_tmap_2d_grouped(&tb, w, g.Cout, (long long)g.R * g.S * g.Cin, (long long)g.R * g.S * g.Cin, BN / 32)
            : make_tmap_2d(&tb, w, g.Cout, (long long)g.R * g.S * g.Cin, (long long)g.R * g.S * g.Cin, 32, true)) return 1;
  ConvTcGeom cg{g.Cin, g.Cout, g.R, g.S, g.pad, g.H, g.W, g.Ho, g.Wo, tile_h, 0, g.Cin};
  const int num_kb = g.R * g.S * cdiv(g.Cout, BK);
  if (BN == 256) { ConvDgradProducer<256> p{cg, b_grp}; return launch_cfg<256, false, true>(ta, tb, ep, p, M, g.Cin, num_kb, 1, num_kb, s); }
  if (BN == 128) { ConvDgradProducer<128> p{cg, b_grp}; return launch_cfg<128, false, true>(ta, tb, ep, p, M, g.Cin, num_kb, 1, num_kb, s); }
  ConvDgradProducer<64> p{cg, b_grp};
  return launch_cfg<64, false, true>(ta, tb, ep, p, M, g.Cin, num_kb, 1, num_kb, s);
}

// RGB stem as an implicit GEMM over the zero-bordered NHWC4 image xp[N][H+6][W+8][4] (see StemFwdProducer): y[N*Ho*Wo, 64] = X'[., 224] . w224^T
int stem_fwd_tf32(const float* xp, const float* w224, Epilogue ep, int N, int H, int W, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2, M = N * Ho * Wo;
  if (Wo != BM || (H & 1) || (W & 1)) { set_error("stem_fwd_tf32: needs an even image with W / 2 == 128 (got %d x %d)", H, W); return 1; }
  CUtensorMap ta, tb;
  if (make_tmap_stem(&ta, xp, N, H + 6, W + 8, Wo, BM, false)) return 1;
  if (make_tmap_2d(&tb, w224, 64, 224, 224, 64)) return 1;
  StemFwdProducer<64> p{Ho, Wo};
  return launch_cfg<64, false, false>(ta, tb, ep, p, M, 64, 7, 1, 7, s);
}
// dw224[64, 224] (+)= dY[N*Ho*Wo, 64]^T . X'
int stem_wgrad_tf32(const float* dy, int lddy, const float* xp, Epilogue ep, int N, int H, int W, cudaStream_t s) {
  const int Ho = H / 2, Wo = W / 2, P = N * Ho * Wo;
  if (Wo % 32 || (H & 1) || (W & 1)) { set_error("stem_wgrad_tf32: needs an even image with W / 2 a multiple of 32"); return 1; }
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, dy, P, 64, lddy, 32, true)) return 1;
  if (make_tmap_stem(&tb, xp, N, H + 6, W + 8, Wo, 32, true)) return 1;
  int num_kb = P / BK, splits, kps;
  plan_splitk(ep, 64, 224, 256, num_kb, 1, splits, kps, s);
  StemWgradProducer<256> p{Ho, Wo};
  return launch_cfg<256, true, true>(ta, tb, ep, p, 64, 224, num_kb, splits, kps, s);
}

int conv_wgrad_tf32(const float* dy, const float* x, Epilogue ep, const ConvGeom& g, float* ws, cudaStream_t s) {
  const int P = g.N * g.Ho * g.Wo, taps = g.R * g.S, Nn = taps * g.Cin;
  // Cin a multiple of 32: the 32-column chunks of an N tile never straddle a tap, so 256-wide tiles may span several taps (the dY operand is
  // re-read once per N tile: 3 times instead of 9 for Cin = 64, 9 instead of 18 for Cin = 256) -- RIH_WGRAD_WIDE=0 keeps one tap per tile
  // Channel counts that are not multiples of 32 (HRNet's 48, 80, 208 ...) take the wide tiles too, on a "virtual" grid of cin32 = ceil32(Cin)
  // columns per tap (rih_set_wgrad_wide bit 1): 48-channel convolutions ran nine 64-wide tiles (dY re-read 9 times, 240 us vs 53 us for the
  // 96-channel convolutions of equal flops in the HRNet-w48 step); three 256-wide tiles over 9 x 64 virtual columns instead.
  const int cin32 = cdiv(g.Cin, 32) * 32;
  const bool wide_pad = (g_wgrad_wide & 2) && g.Cin % 32 != 0 && g.Cin % 4 == 0 && taps * cin32 >= 256;
  const bool wide = g_wgrad_wide && g_persistent && (g.Cin % 32 == 0 || wide_pad) && Nn >= 256;
  const int BN = wide ? 256 : ((g.Cin % 128 == 0) ? 128 : 64);
  // N tiles never straddle a tap: each tap owns ceil(Cin / BN) tiles; when BN does not divide Cin (48, 96, ...) the tile grid is
  // "virtual" (cin_pad columns per tap) and the epilogue stores through a 3-D map [Cout][taps][Cin] that clips at the tap's edge
  const int cin_pad = wide ? cin32 : cdiv(g.Cin, BN) * BN, Ngrid = taps * cin_pad;
  const int bw = g.Wo < 32 ? g.Wo : 32, bh = 32 / bw;
  CUtensorMap ta, tb, tcm;
  const int a_grp = ((g_tma_grouped & 1) && g.Cout % 32 == 0) ? 1 : 0;
  if (a_grp ? make_tmap_2d_grouped(&ta, dy, P, g.Cout, g.ldy, BM / 32) : make_tmap_2d(&ta, dy, P, g.Cout, g.ldy, 32, true)) return 1;
  // input chunks of one tap merged into grouped rank-5 boxes (rih_set_tma_grouped bit 1): b_grp = gcd(chunks per tap, 8) consecutive chunks per box
  int b_grp = 0;
  if ((g_tma_grouped & 2) && wide && !wide_pad && (g.stride == 1 || g_s2_direct)) {
    const int per_tap = g.Cin / 32;
    b_grp = (per_tap % 8 == 0) ? 8 : ((per_tap % 4 == 0) ? 4 : ((per_tap % 2 == 0) ? 2 : 0));
  }
  if (b_grp) {
    if (make_tmap_nhwc_grouped(&tb, x, g.N, g.H, g.W, g.Cin, g.ldx, bw, bh, g.stride == 2 ? 2 : 1, b_grp)) return 1;
  } else if (g.stride == 2 && g_s2_direct) {
    if (make_tmap_nhwc(&tb, x, g.N, g.H, g.W, g.Cin, g.ldx, bw, bh, 1, true, 2)) return 1;
  } else if (g.stride == 2) {
    if (!ws) { set_error("conv_wgrad_tf32: stride-2 path needs workspace"); return 1; }
    if (int e = rih_parity_stack(x, g.ldx, ws, g.N, g.H, g.W, g.Cin, s)) return e;
    if (make_tmap_nhwc(&tb, ws, 4 * g.N, g.H / 2, g.W / 2, g.Cin, g.Cin, bw, bh, 1, true)) return 1;
  } else {
    if (make_tmap_nhwc(&tb, x, g.N, g.H, g.W, g.Cin, g.ldx, bw, bh, 1, true)) return 1;
  }
  ConvTcGeom cg{g.Cin, g.Cout, g.R, g.S, g.pad, g.H, g.W, g.Ho, g.Wo, 0, g.stride == 2 ? (g_s2_direct ? -1 : g.N) : 0, cin_pad};
  int num_kb = P / BK, splits, kps;
  plan_splitk(ep, g.Cout, Nn, BN, num_kb, 1, splits, kps, s, Ngrid);
  const CUtensorMap* cmap = nullptr;
  if (cin_pad != g.Cin) {
    if (ep.ldc != Nn) { set_error("conv_wgrad_tf32: padded-tap path needs a dense dW"); return 1; }
    if (make_tmap_wgrad_out(&tcm, ep.c, g.Cout, taps, g.Cin)) return 1;
    cmap = &tcm; ep.nv_pad = cin_pad; ep.nv_real = g.Cin;
  }
  if (BN == 256) { ConvWgradProducer<256> p{cg, a_grp, b_grp}; return launch_cfg<256, true, true>(ta, tb, ep, p, g.Cout, Ngrid, num_kb, splits, kps, s, cmap); }
  if (BN == 128) { ConvWgradProducer<128> p{cg, a_grp, 0}; return launch_cfg<128, true, true>(ta, tb, ep, p, g.Cout, Ngrid, num_kb, splits, kps, s, cmap); }
  ConvWgradProducer<64> p{cg, a_grp, 0};
  return launch_cfg<64, true, true>(ta, tb, ep, p, g.Cout, Ngrid, num_kb, splits, kps, s, cmap);
}

}  // namespace tc
}  // namespace rih

using namespace rih;

// Raw tcgen05 GEMM entry point (testing / benchmarking of the tensor-core path in isolation):
//   c[M,N] (+)= op(a) * op(b)^T ; a_mn/b_mn select MN-major ([K,M] / [K,N] row-major) operands.
RIH_API int rih_gemm_tf32(const float* a, long long lda, int a_mn, const float* b, long long ldb, int b_mn, float* c, int ldc,
                          int M, int N, int K, const float* bias, int relu, int accumulate, int allow_splitk, int nsplit, cudaStream_t stream) {
  RIH_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tf32: bad shape");
  // nsplit: 1 = TF32, 3 = 3xTF32.  Debug variants: negative = per-thread-store epilogue, +10 = non-persistent kernel.
  int v = nsplit < 0 ? -nsplit : nsplit;
  const int narrow = v >= 20;
  if (narrow) v -= 20;
  const int nonpersistent = v >= 10;
  if (nonpersistent) v -= 10;
  tc::set_wide_tiles(!narrow);
  RIH_REQUIRE(v == 1 || v == 2 || v == 3, "gemm_tf32: nsplit must be 1 (TF32, truncating), 2 (TF32, round-to-nearest) or 3 (3xTF32)");
  tc::set_tma_epilogue(nsplit > 0);
  tc::set_persistent(!nonpersistent);
  nsplit = v;
  tc::set_nsplit(nsplit);
  Epilogue ep = make_epilogue(c, ldc, M, N, bias, relu, accumulate ? 1 : 0);
  int rc = tc::gemm_tf32(a, lda, a_mn, b, ldb, b_mn, ep, M, N, K, allow_splitk, stream);
  tc::set_tma_epilogue(1);
  tc::set_persistent(1);
  tc::set_wide_tiles(1);
  return rc;
}
