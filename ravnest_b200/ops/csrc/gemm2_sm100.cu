// K3, 2-CTA variant: bf16 GEMM with `tcgen05.mma.cta_group::2` — a CTA PAIR (thread-block cluster of 2, two SMs of one TPC) owns a
// 256 x BN output tile.  Each CTA stages ITS 128 rows of A and ITS half (BN/2 rows) of B per k-block; the leader CTA issues one UMMA
// of shape 256 x BN x 16 that reads both halves of B out of both CTAs' shared memory and writes rows 0..127 of the accumulator into
// the leader's TMEM and rows 128..255 into the peer's.  Per CTA and k-block that is 16 KB of A + 16 KB of B for a 128 x 256 slice of
// the product (the 1-CTA kernel needs 16 + 32 KB): shared-memory bandwidth per FLOP is halved, which is what caps the 1-CTA kernel
// at ~0.8 of cuBLAS on large problems, and the smaller stage buys a 6-deep pipeline.
//
//   warp 0 (both CTAs) : TMA producer, `cp.async.bulk.tensor ... .cta_group::2` — completion bytes of BOTH CTAs land on the LEADER's
//                        full barrier (mbarrier address with the pair-rank bit cleared)
//   warp 1 (leader)    : MMA issuer; `tcgen05.commit ... .multicast::cluster` releases the smem stage / publishes the accumulator in
//                        both CTAs
//   warp 2 (both)      : TMEM allocation (`tcgen05.alloc.cta_group::2`)
//   warps 4-11 (both)  : epilogue of the CTA's own 128 rows (tcgen05.ld -> bias / activation -> bf16 -> swizzled smem -> TMA store);
//                        both CTAs hand the accumulator stage back by arriving on the LEADER's tmem_empty barrier
//
// Epilogue features: bias + activation, or accumulate-into-output through TMA reduce-add (the large forward GEMMs and the LM-head weight
// gradient).  GEMMs that need the auxiliary tile paths (activation-gradient multiply, pre-activation store, K2 pull) or split-K stay on
// the 1-CTA kernel in gemm_sm100.cu; binding.cpp picks per call with the wave model (pair_tiling_cost vs choose_tiling).
// Measured on B200 (profiles/ncu_gemm2_r2.md, gemm_sweep_r2i.json): 8192^3 722 us at 98 % tensor-pipe activity (1-CTA kernel: 776 us, 84 %;
// cuBLAS 644 us); LM head 4096 x 30528 x 768: 125 us (1-CTA 147 us, cuBLAS 141 us).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace rvn {

namespace g2 {

constexpr int BLOCK_M = 128;         // rows per CTA (256 per pair)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int SLAB_N = 64;
constexpr int SLAB_BYTES = BLOCK_M * SLAB_N * 2;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the pair-rank bit of a shared::cluster address -> the leader CTA's copy

struct Params {
  int M, N, K;
  const __nv_bfloat16* bias;
  int act;
  int reduce_out;                 // D += acc by TMA reduce-add (gradient accumulation into the bucket) instead of a store
  int num_m2_blocks, num_n_blocks, num_k_blocks;
  float* stats;                   // [2][N] fp32 or nullptr: column sums / sums of squares of the bf16 output are ADDED here (BatchNorm)
};

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;            // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;           // this CTA's half of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 7;
  static constexpr int OPERAND_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = 2 * SLAB_BYTES;
  static constexpr int BAR_OFFSET = OPERAND_BYTES + STAGING_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[both CTAs' smem] * B[both CTAs' smem]; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// when the MMAs issued so far retire, arrive on the barrier at this smem offset in EVERY CTA of ``mask``
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// TMA tile load into THIS CTA's smem whose completion bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the LEADER's copy of a barrier (local arrive for the leader itself)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == 3) return fmaxf(x, 0.0f);
  if (act == 1) {                     // exact-erf GELU (Abramowitz-Stegun erfc, same constants as gemm_sm100.cu)
    const float ax = fabsf(x);
    const float e = exp2f(-0.72134752044448170f * x * x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_erfc = 0.5f * poly * t * e;
    return x * (x >= 0.0f ? 1.0f - half_erfc : half_erfc);
  }
  if (act == 2) {
    const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
    return 0.5f * x * (1.0f + tanhf(u));
  }
  return x;
}

template <int BN, bool A_MN, bool B_MN, bool STATS = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
rvn_gemm_bf16_tcgen05_2cta(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                           const __grid_constant__ CUtensorMap map_d, const Params p) {
  using L = Smem<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int BH = BN / 2;                                   // B rows staged by one CTA
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;   // two accumulator stages
  constexpr uint32_t IDESC = make_idesc_bf16(256, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::OPERAND_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = p.num_m2_blocks * p.num_n_blocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                 // leader's copy is the one in use: its own arrive.expect_tx
      mbar_init(&empty_bar[s], 1);                // one multicast commit per phase
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * (EPI_THREADS / 32));      // the epilogue warps of BOTH CTAs arrive on the leader's copy
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();                               // barrier inits and the TMEM allocation of BOTH CTAs are visible
  __syncthreads();                                  // (implied by the cluster barrier; spelled out for compute-sanitizer racecheck)
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile % p.num_m2_blocks) * 256 + static_cast<int>(rank) * BLOCK_M;      // this CTA's rows of A / D
        const int n0 = (tile / p.num_m2_blocks) * BN + static_cast<int>(rank) * BH;            // this CTA's half of B
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);       // both CTAs' bytes complete on this barrier
          const int k0 = kb * BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_2d_pair(sa, &map_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d_pair(sa + j * (BLOCK_K * 128), &map_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &map_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BH / 64; ++j) tma_load_2d_pair(sb + j * (BLOCK_K * 128), &map_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
            const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
              const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BLOCK_K * 128, 1024)
                                       : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
              umma2_bf16(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
            }
            umma2_commit_mc(&empty_bar[stage], 3);                 // frees this smem stage in both CTAs
            if (kb == p.num_k_blocks - 1) umma2_commit_mc(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue of this CTA's 128 rows (both CTAs)
    const int ewarp = (warp - 4) & 3;
    const int half = (warp - 4) >> 2;
    const int row_in_tile = ewarp * 32 + lane_id();
    const bool issuer = threadIdx.x == 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    ColStatsAcc<BN> cstats;           // STATS instantiations only (see gemm_sm100.cu)
    int stats_n0 = -1;
    if constexpr (STATS) cstats.clear();
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile % p.num_m2_blocks) * 256 + static_cast<int>(rank) * BLOCK_M;
      const int n0 = (tile / p.num_m2_blocks) * BN;
      if constexpr (STATS) {
        if (n0 != stats_n0) {                             // the pair moves to another column block: publish what this CTA has summed
          if (stats_n0 >= 0) cstats.flush(p.stats, p.N, stats_n0, threadIdx.x - 128);
          stats_n0 = n0;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int slab = 0; slab < BN / SLAB_N; ++slab) {
        const int col0 = n0 + slab * SLAB_N;
        if (col0 >= p.N) break;
        uint8_t* buf = staging + (slab_counter & 1) * SLAB_BYTES;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1, EPI_THREADS);
        {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + acc * BN + slab * SLAB_N + half * 32;
          tmem_ld_32x32(taddr, v);
          tmem_ld_wait();
          const int c_base = col0 + half * 32;
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (c_base + i < p.N) {
                const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(p.bias + c_base + i));
                const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 bf = __bfloat1622float2(b2[j]);
                  f[i + 2 * j] += bf.x;
                  f[i + 2 * j + 1] += bf.y;
                }
              }
            }
          }
          if (p.act != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = act_apply(f[i], p.act);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const int chunk = half * 4 + (i >> 3);
            uint4 o;
            o.x = pack_bf16x2(f[i], f[i + 1]); o.y = pack_bf16x2(f[i + 2], f[i + 3]);
            o.z = pack_bf16x2(f[i + 4], f[i + 5]); o.w = pack_bf16x2(f[i + 6], f[i + 7]);
            *reinterpret_cast<uint4*>(buf + row_in_tile * 128 + ((chunk ^ (row_in_tile & 7)) << 4)) = o;
          }
        }
        if (slab == BN / SLAB_N - 1 || col0 + SLAB_N >= p.N) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane_id() == 0) mbar_arrive_leader(&tmem_empty[acc]);     // hand the accumulator stage back to the leader's MMA warp
        }
        fence_proxy_async_smem();
        named_bar_sync(1, EPI_THREADS);
        if (issuer) {
          if (p.reduce_out) tma_reduce_add_2d(&map_d, buf, col0, m0); else tma_store_2d(&map_d, buf, col0, m0);
          tma_store_commit();
        }
        if constexpr (STATS) {  // column sums of the staged slab; rows beyond M are zero (TMA zero-fills the A tail, no bias / activation)
          float ds, dq;
          slab_col_stats<EPI_THREADS>(buf, threadIdx.x - 128, ds, dq);
          cstats.add(slab, ds, dq);
        }
        ++slab_counter;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (STATS) {
      if (stats_n0 >= 0) cstats.flush(p.stats, p.N, stats_n0, threadIdx.x - 128);
    }
    if (issuer) tma_store_wait_all();
  }

  tcgen05_fence_before();
  cluster_sync_all();                               // nobody frees TMEM / exits while the peer may still touch this CTA's smem or TMEM
  if (warp == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
  static EncodeTiledFn4 enc = nullptr;
  if (enc == nullptr) {
    cudaFree(nullptr);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || ptr == nullptr) return -1;
    enc = reinterpret_cast<EncodeTiledFn4>(ptr);
  }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

template <int BN, bool A_MN, bool B_MN, bool STATS = false>
static int launch(const void* A, int lda, const void* B, int ldb, void* D, int ldd, Params p, int num_sms, cudaStream_t stream) {
  using L = Smem<BN>;
  constexpr int BH = BN / 2;
  CUtensorMap ma, mb, md;
  int e = 0;
  e |= A_MN ? make_map(&ma, A, p.M, p.K, lda, 64, BLOCK_K) : make_map(&ma, A, p.K, p.M, lda, BLOCK_K, BLOCK_M);
  e |= B_MN ? make_map(&mb, B, p.N, p.K, ldb, 64, BLOCK_K) : make_map(&mb, B, p.K, p.N, ldb, BLOCK_K, BH);
  e |= make_map(&md, D, p.N, p.M, ldd, SLAB_N, BLOCK_M);
  if (e != 0) return 1000 + e;
  p.num_m2_blocks = (p.M + 255) / 256;
  p.num_n_blocks = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  auto kern = rvn_gemm_bf16_tcgen05_2cta<BN, A_MN, B_MN, STATS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return 2000;
    configured = true;
  }
  const int tiles = p.num_m2_blocks * p.num_n_blocks;
  int pairs = num_sms / 2;
  if (pairs > tiles) pairs = tiles;
  kern<<<2 * pairs, NUM_THREADS, L::TOTAL, stream>>>(ma, mb, md, p);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

}  // namespace g2
}  // namespace rvn

// 2-CTA GEMM entry: D[M,N] = act(op(A) op(B)^T + bias); layouts as in rvn_gemm_bf16.  block_n in {128, 256}.
extern "C" int rvn_gemm_bf16_2cta(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* D, int ldd, int M, int N, int K,
                                  const void* bias, int act, int reduce_out, float* stats, int block_n, int num_sms, cudaStream_t stream) {
  rvn::g2::Params p{};
  if (stats != nullptr && (bias != nullptr || act != 0 || reduce_out || a_mn || b_mn)) return -7;     // forward layout, plain store only
  p.stats = stats;
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.act = act;
  p.reduce_out = reduce_out;
  if (stats != nullptr) {
    if (block_n == 128) return rvn::g2::launch<128, false, false, true>(A, lda, B, ldb, D, ldd, p, num_sms, stream);
    return rvn::g2::launch<256, false, false, true>(A, lda, B, ldb, D, ldd, p, num_sms, stream);
  }
#define RVN_D2(BN_)                                                                                              \
  if (!a_mn && !b_mn) return rvn::g2::launch<BN_, false, false>(A, lda, B, ldb, D, ldd, p, num_sms, stream);       \
  if (!a_mn && b_mn) return rvn::g2::launch<BN_, false, true>(A, lda, B, ldb, D, ldd, p, num_sms, stream);         \
  if (a_mn && b_mn) return rvn::g2::launch<BN_, true, true>(A, lda, B, ldb, D, ldd, p, num_sms, stream);           \
  return rvn::g2::launch<BN_, true, false>(A, lda, B, ldb, D, ldd, p, num_sms, stream);
  if (block_n == 128) { RVN_D2(128) }
  RVN_D2(256)
#undef RVN_D2
}
