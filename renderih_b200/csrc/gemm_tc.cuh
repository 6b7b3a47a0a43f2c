// tcgen05 (5th-gen tensor core) TF32 GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma with the
// fp32 accumulator in TMEM -> tcgen05.ld epilogue.  Hand-written PTX, no CUTLASS.
//   D[M,N] (+)= A(M x K) * B(N x K)^T,  fp32 storage, TF32 multiplicands, fp32 accumulation.
// Operand tiles are [rows x 128 bytes] TMA boxes with SWIZZLE_128B; the same box format serves
//   K-major  operands (rows = M/N index, 32 consecutive k per row)          and
//   MN-major operands (rows = k index, 32 consecutive m/n per row; one box per 32-wide m/n chunk).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_simt.cuh"   // Epilogue

namespace rih {
namespace tc {

constexpr int BM = 128;          // UMMA M
constexpr int BK = 32;           // fp32 elements per stage row = 128 bytes = one swizzle span
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
// L2 eviction-priority hints for streamed operands: an activation tile a GEMM reads once should not push the output of the previous
// kernel (which the next kernel is about to read) out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_hint(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// TMA store of a [rows x 128 B] shared-memory tile (plain or += reduction); bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// 3-D variants (wgrad output viewed as [Cout][taps][Cin] so that a channel chunk is clipped at ITS tap's edge)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar_sync(int wg = 0) { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// tcgen05 shared-memory matrix descriptor (PTX ISA "matrix descriptor", sm_100 version bit set), SWIZZLE_128B.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1     bits [61,64) layout type (2 = SWIZZLE_128B)
//   layout type 1 = SWIZZLE_128B_BASE32B (32-byte swizzle atoms: the only layout tcgen05 accepts for MN-major TF32 operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n, bool a_mn, bool b_mn) {
  return (1u << 4)            // D format F32
       | (2u << 7)            // A format TF32
       | (2u << 10)           // B format TF32
       | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
       | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// NSPLIT = 1: single-pass TF32 (hardware truncates the fp32 operands).  NSPLIT = 2: single-pass TF32 with the operands
// rounded to nearest in shared memory first (unbiased, what cuDNN/cuBLAS TF32 do).  NSPLIT = 3: error-compensated "3xTF32": every operand tile is split in shared memory into
// hi = value rounded to TF32 and lo = value - hi (exact in fp32), and D += hi*hi + lo*hi + hi*lo  (fp32-faithful).
template <int BN, int NSPLIT> __host__ __device__ constexpr int stage_bytes() { return (BM * 128 + BN * 128) * (NSPLIT == 3 ? 2 : 1); }
// Build-time tuning knobs (A/B variants of the library are built by renderih_b200/_build.build_variant):
//   RIH_EPI_WGS   epilogue warpgroups of the single-pass TF32 persistent kernels (default 2)
//   RIH_DEEP      1 = as many shared-memory stages as fit next to the staging tiles for the single-pass TF32 kernels
#ifndef RIH_EPI_WGS
#define RIH_EPI_WGS 2
#endif
#ifndef RIH_DEEP
#define RIH_DEEP 1      // measured on the ResNet-50 trunk (tools/trunk_bench.py): 20.85 -> 20.53 ms fwd+bwd
#endif
template <int BN, int NSPLIT> __host__ __device__ constexpr int num_stages() {
  if (NSPLIT == 1 && RIH_DEEP) {
    // 227 KB per CTA - staging (32 KB per epilogue warpgroup) - barriers, in units of one stage (24 / 32 / 48 KB for BN = 64 / 128 / 256)
    return (227 * 1024 - 32 * 1024 * RIH_EPI_WGS - 2048) / ((BM + BN) * 128);
  }
  return (BN >= 256 && NSPLIT == 3) ? 2 : (BN >= 128 ? 3 : 4);
}
constexpr int EPI_STAGING_BYTES = 2 * BM * 128;   // two [128 x 32 fp32] staging tiles for the TMA-store epilogue (per epilogue warpgroup)
// Epilogue warpgroups of the persistent kernel.  The single-pass TF32 kernels (NSPLIT = 1) have no splitter warps and room in shared memory, so
// they run TWO epilogue warpgroups that take alternate 32-column chunks of the accumulator (each with its own pair of staging tiles, its own
// named barrier and its own TMA-store bulk groups): the store-bound 1x1 convolutions (K = 64 ... 256, 128 KB of output per 128 x 256 tile)
// were limited by the serial tcgen05.ld -> st.shared -> barrier -> TMA-store chain of one warpgroup (3.6 TB/s of 6.5).
template <int NSPLIT> __host__ __device__ constexpr int epi_wgs() { return NSPLIT == 1 ? RIH_EPI_WGS : 1; }
template <int BN, int NSPLIT> __host__ __device__ constexpr int smem_base_bytes() {
  return num_stages<BN, NSPLIT>() * stage_bytes<BN, NSPLIT>() + EPI_STAGING_BYTES * epi_wgs<NSPLIT>() + 1024 + 256;
}
// Per-column epilogue vectors of the tile's BN columns cached in shared memory (persistent kernel): [0] = multiplicative (folded BatchNorm scale),
// [1] = additive (bias or folded BatchNorm shift).  Every epilogue thread needs all 32 values of a chunk: 16 broadcast LDS.128 instead of 64
// broadcast LDG.32 per chunk.  Only when it fits next to the stage ring (it does not for 3xTF32 128 x 256 tiles: global loads there).
template <int BN, int NSPLIT> __host__ __device__ constexpr int colvec_bytes() {
  return (smem_base_bytes<BN, NSPLIT>() + 2 * BN * 4 <= 227 * 1024) ? 2 * BN * 4 : 0;
}
template <int BN, int NSPLIT> __host__ __device__ constexpr int smem_bytes() { return smem_base_bytes<BN, NSPLIT>() + colvec_bytes<BN, NSPLIT>(); }

// ---------------------------------------------------------------- producers (TMA issue logic, one elected lane)
// Each producer loads, for k-block `kb`, the A tile (BM x 32) to `sa` and the B tile (BN x 32) to `sb`.
// Tiles are made of 128-byte rows; MN-major operands are split into 32-wide chunks of BK rows (4 KB apart).
template <int BN, bool A_MN, bool B_MN>
struct DenseProducer {
  // WIDE_OK: an MN-major operand is fetched as BM / 32 (or BN / 32) four-KB boxes per k-block; load<true> issues them from different lanes of
  // the producer warp (lanes 0 .. 3: A boxes, lanes 4 ..: B boxes) instead of one after the other from a single thread.  load<false>(.., 0) is the
  // single-thread form (fully unrolled).  K-major operands are one box each and stay with lane 0 / lane 1.
  static constexpr bool WIDE_OK = A_MN || B_MN;
  int kbeg;
  unsigned long long a_policy;   // 0 = no hint
  // a_grp / b_grp != 0: the MN-major operand comes through a "grouped" rank-3 tensor map {32 columns of a group, k rows, column groups} whose box
  // {32, 32, BM/32 or BN/32} lands in shared memory as the same [group][k row][32] chunks -- ONE TMA operation per operand and k-block
  // instead of one per 32-column chunk (make_tmap_2d_grouped; needs a column count that is a multiple of 32)
  int a_grp, b_grp;
  __device__ __forceinline__ void set_policy(unsigned long long p) { a_policy = p; }
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int k0 = kb * BK;
    if constexpr (!A_MN) { if (!W || lane == 0) { if (a_policy) tma_load_2d_hint(sa, ta, k0, m0, bar, a_policy); else tma_load_2d(sa, ta, k0, m0, bar); } }
    else if (a_grp) { if (!W || lane == 0) tma_load_3d(sa, ta, 0, k0, m0 >> 5, bar); }
    else if constexpr (W) { if (lane < BM / 32) tma_load_2d(sa + lane * (BK * 128), ta, m0 + lane * 32, k0, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), ta, m0 + c * 32, k0, bar);
    }
    if constexpr (!B_MN) { if (!W || lane == BM / 32) tma_load_2d(sb, tb, k0, n0, bar); }
    else if (b_grp) { if (!W || lane == BM / 32) tma_load_3d(sb, tb, 0, k0, n0 >> 5, bar); }
    else if constexpr (W) { const int c = lane - BM / 32; if (c >= 0 && c < BN / 32) tma_load_2d(sb + c * (BK * 128), tb, n0 + c * 32, k0, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), tb, n0 + c * 32, k0, bar);
    }
  }
};

// Batched (per attention head) GEMM operands as 4-D tensors, tile batch index z = b * H + h:
//   token matrix  T[B*S, ld], head slice of width d : dims {d, H, S, B}    -> coordinates (col, z % H, row, z / H)
//   score matrix  P[B*H, R, C] (row stride padded)   : dims {C, R, B*H, 1}  -> coordinates (col, row, z, 0)
// Rows / columns of a 128 x BN tile that stick out of ONE head's matrix are zero-filled by the TMA unit (they never alias the next
// head, because head and batch are separate tensor dimensions), so Sq = 63 ... 316 and d = 16 need no padding copies.
template <int BN, bool A_MN, bool B_MN>
struct BatchedProducer {
  static constexpr bool WIDE_OK = A_MN || B_MN;
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  int H;          // heads per batch image
  int a_tok, b_tok;   // operand is a token matrix (1) or a score matrix (0)
  // token matrix map dims {d, H, S, B} (strides ascending), score matrix map dims {C, R, B*H, 1}
  __device__ __forceinline__ void ld4(void* dst, const CUtensorMap* m, int tok, int col, int row, int z, uint64_t* bar) const {
    if (tok) tma_load_4d(dst, m, col, z % H, row, z / H, bar);
    else tma_load_4d(dst, m, col, row, z, 0, bar);
  }
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int k0 = kb * BK;
    if constexpr (!A_MN) { if (!W || lane == 0) ld4(sa, ta, a_tok, k0, m0, z, bar); }
    else if constexpr (W) { if (lane < BM / 32) ld4(sa + lane * (BK * 128), ta, a_tok, m0 + lane * 32, k0, z, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BM / 32; ++c) ld4(sa + c * (BK * 128), ta, a_tok, m0 + c * 32, k0, z, bar);
    }
    if constexpr (!B_MN) { if (!W || lane == BM / 32) ld4(sb, tb, b_tok, k0, n0, z, bar); }
    else if constexpr (W) { const int c = lane - BM / 32; if (c >= 0 && c < BN / 32) ld4(sb + c * (BK * 128), tb, b_tok, n0 + c * 32, k0, z, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) ld4(sb + c * (BK * 128), tb, b_tok, n0 + c * 32, k0, z, bar);
    }
  }
};

struct ConvTcGeom {
  int Cin, Cout, R, S, pad;
  int H, W, Ho, Wo;        // input / output spatial size (stride 1 only)
  int tile_h;              // fwd/dgrad: rows of the 128-pixel tile (tile_w == full width); tile images = 128/(tile_w*tile_h)
  int s2_images;           // > 0: stride-2 convolution reading the parity-stacked input [4*N, H/2, W/2, C]; value = N
                           // -1 : stride-2 convolution reading the input IN PLACE through a tensor map with element strides {1, 2, 2, 1}
                           //      (the TMA unit fetches every second pixel of a box): coordinates are plain input pixels, no copy of the input
  int cin_pad;             // wgrad: columns per tap in the (virtual) N tile grid = ceil(Cin / BN) * BN (== Cin when BN divides Cin)
};
// stride 2: input row 2*o + r - pad = 2*(o + shift) + parity
__device__ __forceinline__ void s2_tap(int r, int pad, int& parity, int& shift) {
  const int t = r - pad;
  parity = t & 1;
  shift = (t - parity) >> 1;
}

// forward: A = shifted NHWC input boxes (4-D map {C,W,H,N}), B = weights [Cout][R*S*Cin] (K-major 2-D map)
template <int BN>
struct ConvFwdProducer {
  static constexpr bool WIDE_OK = false;     // one or two large boxes per k-block (+ a few weight chunks): issued by lane 0
  ConvTcGeom g;
  unsigned long long a_policy;
  __device__ __forceinline__ void set_policy(unsigned long long p) { a_policy = p; }
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    // channel blocks per tap; when Cin % 32 != 0 the last block of a tap is partly out of bounds in the activation map (TMA zero
    // fill), which also cancels whatever the weight box picks up from the next tap's columns
    const int cpb = (g.Cin + BK - 1) / BK;
    const int tap = kb / cpb, c0 = (kb - tap * cpb) * BK;
    const int r = tap / g.S, s = tap - r * g.S;
    const int P = g.Ho * g.Wo;
    const int n = m0 / P, oh0 = (m0 - n * P) / g.Wo;
    if (g.s2_images < 0) {
      tma_load_4d(sa, ta, c0, s - g.pad, 2 * oh0 + r - g.pad, n, bar);
    } else if (g.s2_images > 0) {
      int ph, dh, pw, dw_;
      s2_tap(r, g.pad, ph, dh);
      s2_tap(s, g.pad, pw, dw_);
      tma_load_4d(sa, ta, c0, dw_, oh0 + dh, (ph * 2 + pw) * g.s2_images + n, bar);
    } else if (a_policy && g.R == 1) {
      tma_load_4d_hint(sa, ta, c0, s - g.pad, oh0 + r - g.pad, n, bar, a_policy);    // single-tap only: a 3x3 re-reads every input tile nine times
    } else {
      tma_load_4d(sa, ta, c0, s - g.pad, oh0 + r - g.pad, n, bar);
    }
    tma_load_2d(sb, tb, tap * g.Cin + c0, n0, bar);
  }
};
// dgrad (stride 1): A = shifted dY boxes (4-D map over [N,Ho,Wo,Cout]), B = weights as MN-major chunks {32 c, 32 co}
template <int BN>
struct ConvDgradProducer {
  static constexpr bool WIDE_OK = false;     // one or two large boxes per k-block (+ a few weight chunks): issued by lane 0
  ConvTcGeom g;
  int b_grp;                                 // weights through a grouped rank-3 map: the BN / 32 chunks of a k-block in ONE box (Cin % 32 == 0)
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int cpb = (g.Cout + BK - 1) / BK;
    const int tap = kb / cpb, co0 = (kb - tap * cpb) * BK;
    const int r = tap / g.S, s = tap - r * g.S;
    const int P = g.H * g.W;
    const int n = m0 / P, ih0 = (m0 - n * P) / g.W;
    tma_load_4d(sa, ta, co0, g.pad - s, ih0 + g.pad - r, n, bar);
    if (b_grp) { tma_load_3d(sb, tb, 0, co0, (tap * g.Cin + n0) >> 5, bar); return; }
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), tb, tap * g.Cin + n0 + c * 32, co0, bar);
  }
};
// dgrad of a stride-2 convolution, one PARITY CLASS of input pixels (ih, iw) = (2 i + ph, 2 j + pw): only the taps r = ph + pad (mod 2),
// s = pw + pad (mod 2) reach those pixels, from output pixel (i + dr, j + ds) with dr = (ph + pad - r) / 2.  Each class is a dense
// stride-1 problem over the half-resolution grid: A = dY boxes shifted by (dr, ds), B = that tap's weights (MN-major chunks), K = ntaps * Cout;
// the four classes together do a quarter of the multiply-adds of "zero-insert dY, then a stride-1 dgrad" and need no dilated copy.
template <int BN>
struct ConvDgradS2Producer {
  static constexpr bool WIDE_OK = false;     // one or two large boxes per k-block (+ a few weight chunks): issued by lane 0
  ConvTcGeom g;            // H, W = half-resolution grid of the class (= Ho, Wo of the convolution); tile_h as usual
  int ntaps;
  int tap[4], dr[4], ds[4];
  int b_grp;               // see ConvDgradProducer
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int cpb = (g.Cout + BK - 1) / BK;
    const int ti = kb / cpb, co0 = (kb - ti * cpb) * BK;
    const int P = g.H * g.W;
    const int n = m0 / P, i0 = (m0 - n * P) / g.W;
    tma_load_4d(sa, ta, co0, ds[ti], i0 + dr[ti], n, bar);
    if (b_grp) { tma_load_3d(sb, tb, 0, co0, (tap[ti] * g.Cin + n0) >> 5, bar); return; }
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) tma_load_2d(sb + c * (BK * 128), tb, tap[ti] * g.Cin + n0 + c * 32, co0, bar);
  }
};
// wgrad: A = dY as MN-major chunks {32 co, 32 pixels}, B = shifted input boxes of 32 pixels as MN-major chunks {32 c, 32 pixels}.
// The N axis enumerates (tap, channel) in cin_pad-wide groups; every 32-column chunk of the N tile finds its OWN tap, so a tile may span
// several taps (Cin = 64: a 256-wide tile covers 4 taps and the dY tile is fetched 3 times instead of 9).
template <int BN>
struct ConvWgradProducer {
  static constexpr bool WIDE_OK = true;      // 4 dY boxes + BN / 32 input boxes (own tap, own coordinates each) of 4 KB per k-block: one lane per box
  ConvTcGeom g;
  int a_grp;                                 // dY through a grouped rank-3 map (one box per k-block), see DenseProducer
  int b_grp;                                 // > 0: the input through a grouped rank-5 map {32, W, H, N, Cin / 32}: one box fetches b_grp consecutive 32-channel
                                             // chunks of ONE tap (b_grp divides both the chunks per tap and the 8 chunks of a tile): BN / 32 / b_grp boxes per k-block
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int p0 = kb * BK;
    const int P = g.Ho * g.Wo;
    const int n = p0 / P, rem = p0 - n * P;
    const int oh = rem / g.Wo, ow = rem - oh * g.Wo;
    const int taps = g.R * g.S;
    if (a_grp) { if (!W || lane == 0) tma_load_3d(sa, ta, 0, p0, m0 >> 5, bar); }
    else if constexpr (W) { if (lane < BM / 32) tma_load_2d(sa + lane * (BK * 128), ta, m0 + lane * 32, p0, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), ta, m0 + c * 32, p0, bar);
    }
    const int G = b_grp > 0 ? b_grp : 1;                  // chunks per box
#pragma unroll
    for (int c0 = 0; c0 < (W ? 1 : BN / 32); ++c0) {
      const int c = (W ? lane - BM / 32 : c0) * (W ? G : 1);
      if (W && (c < 0 || c >= BN / 32)) break;
      if (!W && (c % G) != 0) continue;
      const int col = n0 + c * 32;
      int tap = col / g.cin_pad, cbase = col - tap * g.cin_pad;
      if (tap >= taps) { tap = 0; cbase = g.cin_pad + 64; }      // beyond the last tap: a channel coordinate outside the tensor -> the TMA unit zero-fills
      const int r = tap / g.S, s = tap - r * g.S;
      int cw = ow + s - g.pad, ch = oh + r - g.pad, cn = n;
      if (g.s2_images < 0) { cw = 2 * ow + s - g.pad; ch = 2 * oh + r - g.pad; }
      else if (g.s2_images > 0) {
        int ph, dh, pw, dw_;
        s2_tap(r, g.pad, ph, dh);
        s2_tap(s, g.pad, pw, dw_);
        cw = ow + dw_; ch = oh + dh; cn = (ph * 2 + pw) * g.s2_images + n;
      }
      if (b_grp > 0) tma_load_5d(sb + c * (BK * 128), tb, 0, cw, ch, cn, cbase >> 5, bar);
      else tma_load_4d(sb + c * (BK * 128), tb, cbase, cw, ch, cn, bar);
    }
  }
};

// ---- RGB stem (conv1 7x7 / stride 2 / pad 3, 3 -> 64; torchvision ResNet, models/encoder.py:108) as an implicit GEMM without an im2col buffer.
// The image is stored once as zero-bordered NHWC4 P[n][H+6][W+8][4] (3 rows / columns of padding in front, RGB + one zero channel).  The 8
// horizontally adjacent padded pixels that filter row r sees for output pixel (oh, ow) -- columns 2 ow .. 2 ow + 7 of padded row 2 oh + r, i.e.
// taps s = 0 .. 6 plus one extra column that meets a zero weight -- are 32 CONTIGUOUS floats.  A 4-D tensor map with the (overlapping) strides
// {4 B, 32 B (= 2 pixels), row pitch, image pitch} and dims {32, Wo, H + 6, N} therefore presents them as a "virtual" NHWC tensor with 32
// channels per (row, ow), and the convolution becomes a 7-tap (R = 7, S = 1) implicit GEMM with K = 7 x 32 = 224 against weights repacked
// as [64][7][8][4]: one TMA box {32, 128 ow, 1, 1} per k-block, 128-byte rows, SWIZZLE_128B -- the same machinery as every other
// convolution.  An output tile is one output row (Wo = 128).
template <int BN>
struct StemFwdProducer {
  static constexpr bool WIDE_OK = false;     // one or two large boxes per k-block (+ a few weight chunks): issued by lane 0
  int Ho, Wo;
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int P = Ho * Wo;
    const int n = m0 / P, oh = (m0 - n * P) / Wo;
    tma_load_4d(sa, ta, 0, 0, 2 * oh + kb, n, bar);       // filter row r = kb
    tma_load_2d(sb, tb, kb * 32, n0, bar);
  }
};
// weight gradient of the stem: dW[64][7 x 32] = sum over pixels dY^T . X', k-block = 32 consecutive output pixels of one output row
template <int BN>
struct StemWgradProducer {
  static constexpr bool WIDE_OK = true;
  int Ho, Wo;
  __device__ __forceinline__ void set_policy(unsigned long long) {}
  template <bool W>
  __device__ __forceinline__ void load(const CUtensorMap* ta, const CUtensorMap* tb, int kb, int m0, int n0, int z, uint8_t* sa, uint8_t* sb, uint64_t* bar, int lane) const {
    const int p0 = kb * BK;
    const int P = Ho * Wo;
    const int n = p0 / P, rem = p0 - n * P;
    const int oh = rem / Wo, ow = rem - oh * Wo;
    if constexpr (W) { if (lane < BM / 32) tma_load_2d(sa + lane * (BK * 128), ta, m0 + lane * 32, p0, bar); }
    else {
#pragma unroll
      for (int c = 0; c < BM / 32; ++c) tma_load_2d(sa + c * (BK * 128), ta, m0 + c * 32, p0, bar);
    }
#pragma unroll
    for (int c0 = 0; c0 < (W ? 1 : BN / 32); ++c0) {
      const int c = W ? lane - BM / 32 : c0;
      if (W && (c < 0 || c >= BN / 32)) break;
      const int r = (n0 >> 5) + c;                         // filter row of this 32-column chunk; rows >= 7 do not exist -> out-of-range row coordinate, zero fill
      tma_load_4d(sb + c * (BK * 128), tb, 0, ow, r < 7 ? 2 * oh + r : (1 << 20), n, bar);
    }
  }
};

// Epilogue arithmetic of one 32-column chunk held in registers (row m, columns nb .. nb+31): the same sequence as Epilogue::store4.
//   pre  : the residual row chunk already in registers (prefetched by the caller) or nullptr = load it here
//   cmul / cadd : this chunk's 32 multiplicative / additive column values in shared memory (see colvec_bytes) or nullptr = global loads
__device__ __forceinline__ void epilogue_math(float (&v)[32], const Epilogue& ep, int m, int nb, unsigned long long dseed,
                                              const float4* pre = nullptr, const float* cmul = nullptr, const float* cadd = nullptr) {
  if (ep.scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= ep.scale;
  }
  const bool affine = ep.col_scale != nullptr;
  if (cadd) {                       // cached column vectors: columns beyond N hold 1 / 0
    if (affine && !ep.affine_post) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(cmul + 4 * j), b = *reinterpret_cast<const float4*>(cadd + 4 * j);
        v[4 * j] = fmaf(v[4 * j], a.x, b.x); v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
        v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z); v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
      }
    } else if (ep.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(cadd + 4 * j);
        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
  } else {
    if (affine && !ep.affine_post) {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (nb + j < ep.N) v[j] = fmaf(v[j], __ldg(ep.col_scale + nb + j), __ldg(ep.col_shift + nb + j));
    }
    if (ep.bias) {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (nb + j < ep.N) v[j] += __ldg(ep.bias + nb + j);
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (affine && ep.affine_post) {
    if (cadd) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(cmul + 4 * j), b = *reinterpret_cast<const float4*>(cadd + 4 * j);
        v[4 * j] = fmaf(v[4 * j], a.x, b.x); v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
        v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z); v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (nb + j < ep.N) v[j] = fmaf(v[j], __ldg(ep.col_scale + nb + j), __ldg(ep.col_shift + nb + j));
    }
  }
  if (ep.thresh) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= dropout_scale(dseed, (uint64_t)m * ep.N + nb + j, ep.thresh, ep.inv_keep);
  }
  if (pre) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[4 * j] += pre[j].x; v[4 * j + 1] += pre[j].y; v[4 * j + 2] += pre[j].z; v[4 * j + 3] += pre[j].w; }
  } else if (ep.res && m < ep.M) {        // residual row: 128 contiguous bytes per thread (host checked 16-byte alignment and N % 4 == 0)
    const float4* r4 = reinterpret_cast<const float4*>(ep.res + (size_t)m * ep.ldres + nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) if (nb + 4 * j < ep.N) {
      const float4 q = __ldg(r4 + j);
      v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
    }
  }
  if (ep.relu_post) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}
// residual row chunk (row m, columns nb .. nb+31) into registers; zeros outside the matrix
__device__ __forceinline__ void load_res_chunk(float4 (&r)[8], const Epilogue& ep, int m, int nb) {
  if (m < ep.M) {
    const float4* p = reinterpret_cast<const float4*>(ep.res + (size_t)m * ep.ldres + nb);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (nb + 4 * j < ep.N) ? __ldg(p + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}


// ---------------------------------------------------------------- the kernel
// tma_epi != 0: the epilogue stages 32-column chunks in 128B-swizzled shared memory and writes them with TMA stores
// (cp.reduce.async.bulk ... .add for accumulate / split-K), fully coalesced and clipped at the tensor edge by the hardware.
template <int BN, bool A_MN, bool B_MN, class Producer, int NSPLIT>
__global__ void __launch_bounds__(THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_c,
               Epilogue ep, Producer prod, int num_kb_total, int kb_per_split, int tma_epi) {
  constexpr int STAGES = num_stages<BN, NSPLIT>();
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, AB_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGE_BYTES = stage_bytes<BN, NSPLIT>();   // [A | B] (+ [A_lo | B_lo] when NSPLIT == 3)
  constexpr uint32_t IDESC = make_idesc_tf32(BN, A_MN, B_MN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_STAGING_BYTES);
  uint64_t* full = bars;                 // [STAGES] TMA bytes landed
  uint64_t* empty = bars + STAGES;       // [STAGES] MMAs finished reading the stage
  uint64_t* ready = bars + 2 * STAGES;   // [STAGES] hi/lo split written (NSPLIT == 3 only)
  uint64_t* tmem_full = bars + 3 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);

  pdl_launch_dependents();     // the stream successor may start its own prologue now; it blocks in its pdl_wait() until this grid is done
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kb_beg = blockIdx.z * kb_per_split;
  const int num_kb = min(num_kb_total - kb_beg, kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (tma_epi) tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 128); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // prologue done (barriers, TMEM, descriptors): from here on the kernel reads what its stream predecessor wrote

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], AB_BYTES);
        uint8_t* sa = smem + s * STAGE_BYTES;
        prod.template load<false>(&tmap_a, &tmap_b, kb_beg + kb, m0, n0, 0, sa, sa + A_BYTES, &full[s], 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(NSPLIT == 3 ? &ready[s] : &full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major (SWIZZLE_128B): 8-row groups are 1024 B apart (SBO), a K step of 8 tf32 = +32 B inside the swizzle span.
          // MN-major (SWIZZLE_128B_BASE32B, TMA 128B_ATOM_32B): 32-wide m/n chunks are BK*128 B apart (LBO), 4-row k groups
          // are 512 B apart (SBO), a K step of 8 rows = +1024 B.
          uint64_t ad = A_MN ? make_smem_desc(sa + k * 1024, BK * 128, 512, 1) : make_smem_desc(sa + k * 32, 16, 1024, 2);
          uint64_t bd = B_MN ? make_smem_desc(sb + k * 1024, BK * 128, 512, 1) : make_smem_desc(sb + k * 32, 16, 1024, 2);
          umma_tf32(tmem_base, ad, bd, IDESC, (kb | k) ? 1u : 0u);
          if constexpr (NSPLIT == 3) {
            // descriptors address in 16-byte units: the lo copies sit AB_BYTES after the hi tiles
            const uint64_t lo_off = (uint64_t)(AB_BYTES >> 4);
            umma_tf32(tmem_base, ad + lo_off, bd, IDESC, 1u);   // lo(A) * hi(B)
            umma_tf32(tmem_base, ad, bd + lo_off, IDESC, 1u);   // hi(A) * lo(B)
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // warps 2..5: (NSPLIT == 3) hi/lo splitter during the main loop, then the epilogue.
    if constexpr (NSPLIT == 3) {
      const int t = threadIdx.x - 64;   // 0..127
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        uint4* hi = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + AB_BYTES);
#pragma unroll 4
        for (int i = t; i < AB_BYTES / 16; i += 128) {
          uint4 v = hi[i];
          // hi = round-to-nearest TF32 (10-bit mantissa): unbiased, so the dropped lo*lo term and the hardware's truncation of lo
          // do not accumulate coherently over K
          uint4 h = make_uint4((v.x + 0x1000u) & 0xFFFFE000u, (v.y + 0x1000u) & 0xFFFFE000u, (v.z + 0x1000u) & 0xFFFFE000u, (v.w + 0x1000u) & 0xFFFFE000u);
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[i] = h;
          lo[i] = l;
        }
        fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&ready[s]);
      }
    }
    // TMEM lane quarters (warp % 4)
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int rr = q * 32 + lane;       // row inside the tile
    const int m = m0 + rr;
    if (num_kb > 0 && tma_epi) {
      const bool elected = (threadIdx.x == 64);
      const unsigned long long dseed = ep.thresh ? (*ep.seed_ptr + ep.site * 0xD1B54A32D192ED03ull) : 0ull;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
        const int nb = n0 + c * 32;
        epilogue_math(v, ep, m, nb, dseed);
        if (c >= 2) {                    // staging buffer (c & 1) must have been read out by its previous TMA store
          if (elected) tma_store_wait_read<1>();
          epi_bar_sync();
        }
        uint8_t* buf = staging + (c & 1) * (BM * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)      // 16-byte chunk j of row rr lives at chunk (j ^ (rr & 7)): the SWIZZLE_128B pattern
          *reinterpret_cast<float4*>(buf + rr * 128 + ((j ^ (rr & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        epi_bar_sync();
        if (elected) {
          if (ep.mode == 0) tma_store_2d(&tmap_c, buf, nb, m0);
          else tma_reduce_add_2d(&tmap_c, buf, nb, m0);
          tma_store_commit();
        }
      }
      if (elected) tma_store_wait<0>();
    } else if (num_kb > 0) {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ep.store4(m, n0 + c * 32 + j * 4, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ---------------------------------------------------------------- persistent variant
// One CTA per SM loops over output tiles (static round-robin).  The shared-memory stage ring runs continuously across
// tiles, the accumulator is double-buffered in TMEM (2 x BN columns) and the epilogue has its own four warps, so the
// TMA stores of tile i overlap the loads / MMAs of tile i+1 and the per-CTA prologue (TMEM allocation, barrier init,
// descriptor fetch) is paid once per SM instead of once per tile.
//   warp 0 = TMA producer | warp 1 = MMA issuer + TMEM owner | warps 2..9 = hi/lo splitter (NSPLIT >= 2 only) |
//   last 4 warps = epilogue (TMEM -> registers -> swizzled smem -> TMA store / reduce-add)
constexpr int SPLIT_WARPS = 8;   // hi/lo splitter warps of the persistent kernel (NSPLIT >= 2): the split pass (32-48 KB of shared memory per k-block) is what
                                 // bounds the many small, latency-bound 3xTF32 GEMMs of the decoder -- 8 warps halve it against the original 4
template <int NSPLIT> __host__ __device__ constexpr int persistent_threads() { return NSPLIT >= 2 ? (2 + SPLIT_WARPS + 4) * 32 : (2 + 4 * epi_wgs<NSPLIT>()) * 32; }

template <int BN, bool A_MN, bool B_MN, class Producer, int NSPLIT>
__global__ void __launch_bounds__(persistent_threads<NSPLIT>())
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_c,
                          Epilogue ep, Producer prod, int num_kb_total, int kb_per_split, int tiles_m, int tiles_n, int splits, int tma_epi) {
  constexpr int STAGES = num_stages<BN, NSPLIT>();
  constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, AB_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGE_BYTES = stage_bytes<BN, NSPLIT>();
  constexpr uint32_t IDESC = make_idesc_tf32(BN, A_MN, B_MN);
  constexpr int EPI_WARP0 = (NSPLIT >= 2) ? 2 + SPLIT_WARPS : 2;
  constexpr int EPI_WG = epi_wgs<NSPLIT>();
  constexpr uint32_t TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_STAGING_BYTES * EPI_WG);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* ready = bars + 2 * STAGES;
  uint64_t* tmem_full = bars + 3 * STAGES;       // [2]
  uint64_t* tmem_empty = bars + 3 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);
  constexpr bool COLVEC = colvec_bytes<BN, NSPLIT>() > 0;
  float* const colvec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2][BN] when COLVEC

  pdl_launch_dependents();     // the stream successor may start its own prologue now; it blocks in its pdl_wait() until this grid is done
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = tiles_m * tiles_n * splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (tma_epi) tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], SPLIT_WARPS * 32); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128 * EPI_WG); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // prologue done (barriers, TMEM, descriptors): from here on the kernel reads what its stream predecessor wrote

  // third tile coordinate z: split-K slice (kb_beg = z * kb_per_split), or -- batched GEMM (ep.batch_heads != 0) -- the batch index
  const bool batched = ep.batch_heads != 0;
  auto tile_coords = [&](int t, int& m0, int& n0, int& kb_beg, int& nkb, int& z) {
    if (ep.reverse) t = total_tiles - 1 - t;          // serpentine traversal across consecutive kernels (rih_set_traversal)
    const int ni = t % tiles_n;
    const int r = t / tiles_n;
    const int mi = r % tiles_m;
    z = r / tiles_m;
    m0 = mi * BM; n0 = ni * BN;
    kb_beg = batched ? 0 : z * kb_per_split;
    nkb = batched ? num_kb_total : min(num_kb_total - kb_beg, kb_per_split);
  };

  if (warp == 0) {
    // Producers with many small boxes per k-block (MN-major operands: the weight-gradient GEMMs) produce with the whole warp: lane 0 waits for
    // the stage and arms its barrier, then every lane issues its box (Producer::load<true>).  ep.opt bit 2 = 0 keeps the single-thread issue.
    // (Measured: ConvWgrad kernels 3.38 -> 2.15 ms per step; producers with one or two large boxes were SLOWER warp-wide and stay single-thread.)
    const bool wide_issue = Producer::WIDE_OK && (ep.opt & 4) != 0;
    if (lane == 0 || wide_issue) {
      uint32_t it = 0;
      Producer pr = prod;
      pr.set_policy(ep.a_policy ? l2_policy_evict_first() : 0ull);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb_beg, nkb, z;
        tile_coords(t, m0, n0, kb_beg, nkb, z);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if (lane == 0) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], AB_BYTES);
          }
          if (wide_issue) __syncwarp();
          uint8_t* sa = smem + s * STAGE_BYTES;
          // k-block rotation: the sum over k does not care about the order, so each tile starts at its own k-block and wraps around --
          // CTAs that run in lockstep then fetch DIFFERENT 128-byte slices of their (1 KB-pitch) rows at any instant instead of all hitting
          // the same address bits [7, 10) of every row
          int kk = kb;
          if (ep.kb_rotate) { kk = kb + (t % nkb); if (kk >= nkb) kk -= nkb; }
          if constexpr (Producer::WIDE_OK) {
            if (wide_issue) pr.template load<true>(&tmap_a, &tmap_b, kb_beg + kk, m0, n0, z, sa, sa + A_BYTES, &full[s], lane);
            else pr.template load<false>(&tmap_a, &tmap_b, kb_beg + kk, m0, n0, z, sa, sa + A_BYTES, &full[s], 0);
          } else {
            pr.template load<false>(&tmap_a, &tmap_b, kb_beg + kk, m0, n0, z, sa, sa + A_BYTES, &full[s], 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tc) {
        int m0, n0, kb_beg, nkb, z;
        tile_coords(t, m0, n0, kb_beg, nkb, z);
        const uint32_t acc = tc & 1, aph = (tc >> 1) & 1;
        mbar_wait(&tmem_empty[acc], aph ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(NSPLIT >= 2 ? &ready[s] : &full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            uint64_t ad = A_MN ? make_smem_desc(sa + k * 1024, BK * 128, 512, 1) : make_smem_desc(sa + k * 32, 16, 1024, 2);
            uint64_t bd = B_MN ? make_smem_desc(sb + k * 1024, BK * 128, 512, 1) : make_smem_desc(sb + k * 32, 16, 1024, 2);
            umma_tf32(tmem_d, ad, bd, IDESC, (kb | k) ? 1u : 0u);
            if constexpr (NSPLIT == 3) {
              const uint64_t lo_off = (uint64_t)(AB_BYTES >> 4);
              umma_tf32(tmem_d, ad + lo_off, bd, IDESC, 1u);
              umma_tf32(tmem_d, ad, bd + lo_off, IDESC, 1u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (NSPLIT >= 2 && warp < EPI_WARP0) {
    // hi/lo splitter warps 2 .. 2 + SPLIT_WARPS - 1
    const int tsp = threadIdx.x - 64;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m0, n0, kb_beg, nkb, z;
      tile_coords(t, m0, n0, kb_beg, nkb, z);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        uint4* hi = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(smem + s * STAGE_BYTES + AB_BYTES);
#pragma unroll 4
        for (int i = tsp; i < AB_BYTES / 16; i += SPLIT_WARPS * 32) {
          uint4 v = hi[i];
          uint4 h = make_uint4((v.x + 0x1000u) & 0xFFFFE000u, (v.y + 0x1000u) & 0xFFFFE000u, (v.z + 0x1000u) & 0xFFFFE000u, (v.w + 0x1000u) & 0xFFFFE000u);
          uint4 l;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[i] = h;
          if constexpr (NSPLIT == 3) lo[i] = l;
        }
        fence_proxy_async();
        mbar_arrive(&ready[s]);
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int q = warp & 3;                         // TMEM lane quarter this warp may read (warp index modulo 4)
    const int wg = (warp - EPI_WARP0) >> 2;         // epilogue warpgroup: takes the chunks c with c % EPI_WG == wg
    const int rr = q * 32 + lane;
    const bool elected = (threadIdx.x == (EPI_WARP0 + 4 * wg) * 32);
    uint8_t* const wg_staging = staging + wg * EPI_STAGING_BYTES;
    const unsigned long long dseed = ep.thresh ? (*ep.seed_ptr + ep.site * 0xD1B54A32D192ED03ull) : 0ull;
    constexpr int NCHUNK = BN / 32;
    const int last_c = ((NCHUNK - 1 - wg) / EPI_WG) * EPI_WG + wg;    // last chunk of this warpgroup
    // fused BatchNorm statistics: per-thread fp64 accumulators (lane = column of a chunk, warp = row quarter), carried ACROSS the tiles of this
    // persistent CTA and flushed with one atomic pair per column when the CTA moves to another column block / at the end.  Flushing per
    // tile put ~10^6 fp64 atomics per convolution onto 2 N addresses: same-address serialisation in the L2 (~10 ns each) DOUBLED the time
    // of the memory-bound 1x1 convolutions in train mode (146 us in the step vs 67 us alone for 256 -> 64 at 64x64).
    constexpr int NCH_WG = (NCHUNK + EPI_WG - 1) / EPI_WG;
    double racc1[NCH_WG], racc2[NCH_WG];
#pragma unroll
    for (int k = 0; k < NCH_WG; ++k) { racc1[k] = 0.0; racc2[k] = 0.0; }
    int acc_n0 = -1;
    auto flush_stats = [&]() {
      if (acc_n0 >= 0) {
#pragma unroll
        for (int k = 0; k < NCH_WG; ++k) {
          const int col = acc_n0 + (k * EPI_WG + wg) * 32 + lane;
          if (col < ep.N && (k * EPI_WG + wg) < NCHUNK) { atomicAdd(ep.stats + col, racc1[k]); atomicAdd(ep.stats + ep.N + col, racc2[k]); }
          racc1[k] = 0.0; racc2[k] = 0.0;
        }
      }
    };
    uint32_t tc = 0, cc = 0;
    // residual rows are prefetched one chunk ahead into registers (the first chunk of a tile before the wait for its accumulator): a load issued
    // inside the chunk's own arithmetic exposed one DRAM latency per chunk (conv3 + identity of a bottleneck: 230 us for 603 MB)
    const bool use_res = tma_epi && ep.res != nullptr && !batched && !ep.nv_pad && (ep.opt & 1);
    const bool use_cv = COLVEC && tma_epi && !batched && !ep.nv_pad && (ep.bias != nullptr || ep.col_scale != nullptr) && (ep.opt & 2);
    // single-pass TF32 kernels (320 threads, 204 registers each) double-buffer the prefetch: chunk c + 1 is requested before chunk c's accumulator
    // is read; the operand-splitting kernels (448 threads, 128 registers) re-use one buffer and request the next chunk at the end of the current one
    constexpr bool RES_DB = (NSPLIT == 1);
    float4 rcur[8], rnext[RES_DB ? 8 : 1];
    int cv_n0 = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tc) {
      int m0, n0, kb_beg, nkb, z;
      tile_coords(t, m0, n0, kb_beg, nkb, z);
      const uint32_t acc = tc & 1, aph = (tc >> 1) & 1;
      const int m = m0 + rr;
      if (use_res) load_res_chunk(rcur, ep, m, n0 + wg * 32);
      if (use_cv && n0 != cv_n0) {       // this warpgroup's chunks of the tile's column vectors -> shared memory (all its threads are past the last read)
        const int tw = threadIdx.x - (EPI_WARP0 + 4 * wg) * 32;      // 0..127 inside the warpgroup
        for (int i = tw; i < NCH_WG * 32; i += 128) {
          const int col = n0 + ((i >> 5) * EPI_WG + wg) * 32 + (i & 31);
          const bool in = col < ep.N;
          float mul = 1.f, add = 0.f;
          if (ep.col_scale) { if (in) { mul = __ldg(ep.col_scale + col); add = __ldg(ep.col_shift + col); } }
          else if (in) add = __ldg(ep.bias + col);
          colvec[wg * NCH_WG * 32 + i] = mul;
          colvec[BN + wg * NCH_WG * 32 + i] = add;
        }
        cv_n0 = n0;
        epi_bar_sync(wg);
      }
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      if (ep.stats && n0 != acc_n0) { flush_stats(); acc_n0 = n0; }
#pragma unroll 1
      for (int c = wg; c < NCHUNK; c += EPI_WG) {
        if constexpr (RES_DB) {
          if (use_res) {
            if (c != wg) {
#pragma unroll
              for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
            }
            if (c + EPI_WG < NCHUNK) load_res_chunk(rnext, ep, m, n0 + (c + EPI_WG) * 32);
          }
        }
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (c == last_c) {             // all TMEM reads of this accumulator by this thread are done: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        const int nb = n0 + c * 32;
        int sc0 = nb, sc1 = 0;
        if (ep.nv_pad) {               // virtual wgrad column -> (channel, tap); chunks wholly beyond Cin are skipped (warp-uniform)
          sc1 = nb / ep.nv_pad; sc0 = nb - sc1 * ep.nv_pad;
          if (sc0 >= ep.nv_real || sc1 * ep.nv_real >= ep.N) continue;      // beyond the tap's channels / beyond the last tap (wide tiles over a padded grid)
        }
        if (batched && nb >= ep.N) continue;   // e.g. head dim 16 in a 64-wide tile: nothing to store
        if (tma_epi) {
          const float* cmul = use_cv ? colvec + (wg * NCH_WG + (c - wg) / EPI_WG) * 32 : nullptr;
          epilogue_math(v, ep, m, nb, dseed, use_res ? rcur : nullptr, cmul, use_cv ? cmul + BN : nullptr);
          if (cc >= 2) {
            if (elected) tma_store_wait_read<1>();
            epi_bar_sync(wg);
          }
          uint8_t* buf = wg_staging + (cc & 1) * (BM * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(buf + rr * 128 + ((j ^ (rr & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          epi_bar_sync(wg);
          if (elected) {
            if (batched) {
              const int hh = ep.batch_heads & 0x3FFFFFFF;
              if (ep.batch_heads & (1 << 30)) tma_store_4d(&tmap_c, buf, nb, m0, z, 0);     // score matrix [B*H][M][N]
              else tma_store_4d(&tmap_c, buf, nb, z % hh, m0, z / hh);                        // head slice of a token matrix
            } else if (ep.s2_w2) {           // stride-2 dgrad parity class: rows (n, i, j) -> input pixels (2 i + ph, 2 j + pw), element-strided 4-D map
              const int P2 = ep.s2_h2 * ep.s2_w2;
              const int n_ = m0 / P2, i0 = (m0 - n_ * P2) / ep.s2_w2;
              if (ep.mode == 0) tma_store_4d(&tmap_c, buf, nb, ep.s2_pw, 2 * i0 + ep.s2_ph, n_);
              else tma_reduce_add_4d(&tmap_c, buf, nb, ep.s2_pw, 2 * i0 + ep.s2_ph, n_);
            } else if (ep.nv_pad) {
              if (ep.mode == 0) tma_store_3d(&tmap_c, buf, sc0, sc1, m0);
              else tma_reduce_add_3d(&tmap_c, buf, sc0, sc1, m0);
            } else if (ep.mode == 0) tma_store_2d(&tmap_c, buf, nb, m0);
            else tma_reduce_add_2d(&tmap_c, buf, nb, m0);
            tma_store_commit();
          }
          if (ep.stats) {
            // fused BatchNorm statistics: column sums / sums of squares of the chunk over this warp's 32 rows.  Each lane holds one ROW
            // (32 columns) in registers: a halving butterfly over the lanes (16 + 8 + 4 + 2 + 1 = 31 exchanges per quantity) leaves lane l with
            // the totals of column l -- no shared-memory pass (the serial 32-row LDS loop this replaces cost ~0.5 us per chunk, a third of
            // the epilogue of the store-bound 1x1 convolutions).  Then one fp64 atomic pair per column per warp.
            const int rows_valid = min(32, ep.M - (m0 + q * 32));
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float x = (lane < rows_valid) ? v[j] : 0.f; s1[j] = x; s2[j] = x * x; }
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1) {
              const bool up = (lane & w) != 0;
#pragma unroll
              for (int j = 0; j < w; ++j) {
                // keep the half of the columns whose bit `w` equals this lane's bit `w`, hand the other half to the partner lane
                const float keep1 = up ? s1[j + w] : s1[j], give1 = up ? s1[j] : s1[j + w];
                const float keep2 = up ? s2[j + w] : s2[j], give2 = up ? s2[j] : s2[j + w];
                s1[j] = keep1 + __shfl_xor_sync(0xffffffffu, give1, w);
                s2[j] = keep2 + __shfl_xor_sync(0xffffffffu, give2, w);
              }
            }
            // after the butterfly lane l holds column l of the chunk (bit w of the column index = bit w of the lane)
            const int ci = (c - wg) / EPI_WG;
#pragma unroll
            for (int k = 0; k < NCH_WG; ++k) if (k == ci) { racc1[k] += (double)s1[0]; racc2[k] += (double)s2[0]; }
          }
          ++cc;
          if constexpr (!RES_DB) {
            if (use_res && c + EPI_WG < NCHUNK) load_res_chunk(rcur, ep, m, n0 + (c + EPI_WG) * 32);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ep.store4(m, nb + j * 4, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      }
    }
    if (ep.stats) flush_stats();
    if (elected && tma_epi) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

// 2-D fp32 tensor map over a row-major [rows, cols] matrix with row stride ld (elements); box = {32 cols, box_rows}.
int make_tmap_2d(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_rows, bool atom32 = false);
// MN-major operand [rows (k), cols] with cols % 32 == 0 as a rank-3 map {32, rows, cols / 32}: box {32, 32, groups} = `groups` chunks in one operation
int make_tmap_2d_grouped(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int groups);
// 4-D fp32 tensor map over an NHWC tensor [N,H,W,C] with pixel stride ld; box = {32 channels, bw, bh, bn}.
// estride = 2: the box visits every second pixel in W and H (stride-2 convolutions read / write the full-resolution tensor in place);
// bw / bh stay the numbers of pixels FETCHED
int make_tmap_nhwc(CUtensorMap* map, const float* base, int N, int H, int W, int C, long long ld, int bw, int bh, int bn, bool atom32, int estride = 1);
// NHWC tensor with C % 32 == 0 as a rank-5 map {32, W, H, N, C / 32}: box {32, bw, bh, 1, groups} = `groups` consecutive 32-channel chunks of one pixel window
int make_tmap_nhwc_grouped(CUtensorMap* map, const float* base, int N, int H, int W, int C, long long ld, int bw, int bh, int estride, int groups);
// overlapping-stride view of the zero-bordered NHWC4 stem input (see StemFwdProducer); box = {32, box_ow, 1, 1}
int make_tmap_stem(CUtensorMap* map, const float* base, int N, int Hp, int Wp, int Wo, int box_ow, bool atom32);

}  // namespace tc
}  // namespace rih
