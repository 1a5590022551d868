// K3: bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, cta_group::1, UMMA 128 x BN x 16), operands staged in
// shared memory by TMA (128B swizzle), fp32 accumulators in TMEM (double-buffered so the epilogue of tile i
// overlaps the MMAs of tile i+1), persistent CTAs (one per SM), warp-specialised:
//   warp 0 : TMA producer      warp 1 : MMA issuer (one elected lane)     warp 2 : TMEM allocator
//   warps 4-11: epilogue  (tcgen05.ld -> bias / activation / activation-gradient -> bf16 -> swizzled smem -> TMA store); two warps
//               per TMEM lane quadrant, each taking one 32-column half of the 64-column slab (GELU epilogues are issue-bound)
//
//   D[M,N] = epi( A * B ),  A is M x K, B is N x K ("K-major", row-major [rows, K]) or their transposes stored
//   row-major [K, rows] ("MN-major") — so forward (x W^T), dgrad (dY W) and wgrad (dY^T x) all run on the same
//   kernel without materialising a transpose.
//
// Reference call sites this replaces: every nn.Linear fwd/dgrad/wgrad of the BERT / GPT / CNN / fc layers
//  which upstream go through cuBLAS via torch 2.2.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace rvn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int EPI_THREADS = 256;
constexpr int SLAB_N = 64;           // epilogue staging slab: 128 rows x 64 cols bf16 = 16 KB
constexpr int SLAB_BYTES = BLOCK_M * SLAB_N * 2;

enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_GELU_TANH = 2, ACT_RELU = 3 };

struct GemmParams {
  int M, N, K;
  const __nv_bfloat16* bias;      // [N] or nullptr
  const __nv_bfloat16* dact_in;   // [M, ld_aux] auxiliary input tile or nullptr
  int ld_aux;
  int aux_mode;                   // 1: D = acc * act'(aux)   2: D = acc + aux (gradient accumulation)   3: D = acc * aux
  int act;                        // Act
  int store_preact;               // also write through map_aux: 1 the pre-activation z, 2 the activation derivative act'(z)
  int b_copy;                     // K2: B lives in a PEER's HBM.  The (otherwise idle) epilogue warps of all CTAs first pull it
  const __nv_bfloat16* b_peer;    //     tile by tile over NVLink into the local buffer (every SM issues peer loads: one SM alone gets
  __nv_bfloat16* b_local;         //     ~5 GB/s, the whole chip ~750 GB/s) and publish flags[n_blk * num_k_blocks + kb] = epoch;
  int ld_bpeer, ld_blocal;        //     TMA producers wait on the flag of the tile they need and read the LOCAL copy.
  unsigned* bcopy_flags;
  unsigned bcopy_epoch;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int split_k;                    // >1: work unit = (tile, k-slice); only with reduce_out (partials are summed by the L2)
  int reduce_out;                 // D += acc through TMA reduce-add (cp.reduce.async.bulk.tensor .add on the bf16 map) instead of
                                  // store: gradient accumulation into the flat bucket without reading it, and split-K safe
  float* stats;                   // [2][N] fp32 or nullptr: column sums / sums of squares of the bf16 output are ADDED here (the batch
                                  // statistics of a following BatchNorm; plain store epilogues without split-K only)
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int OPERAND_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_BYTES = 2 * SLAB_BYTES;
  static constexpr int BAR_OFFSET = OPERAND_BYTES + STAGING_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // + barriers + slack for manual 1024 B alignment
};

// exact-erf GELU and its derivative from ONE exponential: with t = 1 / (1 + p |x|/sqrt2) (Abramowitz-Stegun 7.1.26, |err| < 1.5e-7),
//   erfc(|x|/sqrt2) = poly(t) * exp(-x^2/2)   and   pdf(x) = exp(-x^2/2) / sqrt(2 pi)
// so cdf, gelu = x cdf and gelu' = cdf + x pdf cost one ex2, one rcp and ~12 FMAs (libdevice erff alone is ~2x that and branches).
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf_x) {
  const float ax = fabsf(x);
  const float e = exp2f(-0.72134752044448170f * x * x);              // exp(-x^2/2)
  const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * e;                        // 0.5 * erfc(|x|/sqrt2)
  cdf = x >= 0.0f ? 1.0f - half_erfc : half_erfc;
  pdf_x = 0.39894228040143268f * x * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, px;
  gelu_parts(x, cdf, px);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float cdf, px;
  gelu_parts(x, cdf, px);
  return cdf + px;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float dgelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(x);
    case ACT_GELU_TANH: return gelu_tanh(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}
__device__ __forceinline__ float apply_dact(float z, int act) {
  switch (act) {
    case ACT_GELU: return dgelu_erf(z);
    case ACT_GELU_TANH: return dgelu_tanh(z);
    case ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}

template <int BN, bool A_MN, bool B_MN, bool STATS = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rvn_gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_aux,
                      const __grid_constant__ CUtensorMap map_bcopy, const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);   // two accumulator stages, power of two
  constexpr uint32_t IDESC = make_idesc_bf16(BLOCK_M, BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::OPERAND_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* aux_bar = tmem_empty + 2;                     // TMA load of the auxiliary input slab (epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_units = num_tiles * p.split_k;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
    if (p.store_preact || p.dact_in != nullptr) tma_prefetch_desc(&map_aux);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_THREADS / 32);
    }
    mbar_init(aux_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit % num_tiles;
        const int ks = unit / num_tiles;
        const int kb_begin = ks * p.num_k_blocks / p.split_k, kb_end = (ks + 1) * p.num_k_blocks / p.split_k;
        const int m0 = (tile % p.num_m_blocks) * BLOCK_M;
        const int n0 = (tile / p.num_m_blocks) * BN;
        if (!B_MN && p.b_copy) {
          // K2: wait until the epilogue warps (of up to num_k_blocks different CTAs, in parallel) have pulled this n-block of B
          // over NVLink into local HBM; ONE proxy fence per tile, then every B load of the tile reads the local copy
          // (peer memory bypasses L2, so each weight byte must cross the link exactly once)
          const unsigned* flags = p.bcopy_flags + (tile / p.num_m_blocks) * p.num_k_blocks;
          for (int kb = 0; kb < p.num_k_blocks; ++kb) {
            unsigned seen;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + kb) : "memory");
            } while (seen < p.bcopy_epoch);
          }
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          const int k0 = kb * BLOCK_K;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &map_a, &full_bar[stage], k0, m0);                 // box {64 (K), 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)                             // box {64 (M), 64 K-rows} per atom column
              tma_load_2d(sa + j * (BLOCK_K * 128), &map_a, &full_bar[stage], m0 + j * 64, k0);
          }
          if constexpr (!B_MN) {
            if (p.b_copy) {
              tma_load_2d(sb, &map_bcopy, &full_bar[stage], k0, n0);          // local copy (its landing was awaited per tile)
            } else {
              tma_load_2d(sb, &map_b, &full_bar[stage], k0, n0);               // box {64 (K), BN rows}
            }
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (BLOCK_K * 128), &map_b, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int ks = unit / num_tiles;
      const int kb_begin = ks * p.num_k_blocks / p.split_k, kb_end = (ks + 1) * p.num_k_blocks / p.split_k;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major : atoms are 8 rows x 128 B, SBO = 1024 B between 8-row groups, advance 32 B per UMMA_K
            // MN-major: atoms are 8 K-rows x 128 B (64 MN elements), SBO = 1024 B along K, LBO = BLOCK_K*128 B along MN,
            //           advance 2 atoms (2048 B) per UMMA_K
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16(tmem_d, da, db, IDESC, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem stage when these MMAs retire
          if (kb == kb_end - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue (256 threads: one row x one 32-column half each)
    const int ewarp = (warp - 4) & 3;           // TMEM lane quadrant == warp index % 4
    const int half = (warp - 4) >> 2;           // which 32-column half of every 64-column slab this warp owns
    const int row_in_tile = ewarp * 32 + lane_id();
    const bool issuer = threadIdx.x == 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    uint32_t aux_phase = 0;
    if (!B_MN && p.b_copy) {
      // ---- K2 cooperative pull: B tile t = (n_blk, kb) is fetched from the leader by CTA (t mod grid); 16 B per load,
      //      BN*8/128 independent loads per thread in flight, landed with plain stores, then a release flag
      const int et = threadIdx.x - 128;
      const int n_pull = p.num_n_blocks * p.num_k_blocks;
      for (int t = blockIdx.x; t < n_pull; t += gridDim.x) {
        const int r0 = (t / p.num_k_blocks) * BN;
        const int c0 = (t % p.num_k_blocks) * BLOCK_K;
#pragma unroll 4
        for (int ch = et; ch < BN * 8; ch += EPI_THREADS) {
          const int r = r0 + (ch >> 3), c = c0 + (ch & 7) * 8;
          if (r < p.N && c < p.K) {
            const uint4 v = *reinterpret_cast<const uint4*>(p.b_peer + static_cast<size_t>(r) * p.ld_bpeer + c);
            *reinterpret_cast<uint4*>(p.b_local + static_cast<size_t>(r) * p.ld_blocal + c) = v;
          }
        }
        __threadfence();
        named_bar_sync(1, EPI_THREADS);
        if (et == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.bcopy_flags + t), "r"(p.bcopy_epoch) : "memory");
      }
    }
    const bool has_aux_in = p.dact_in != nullptr;           // never together with reduce_out / split_k (host guarantees it)
    const bool one_buffer = p.store_preact || has_aux_in;     // staging[1] is taken by the aux slab (in or out)
    if (has_aux_in && issuer && static_cast<int>(blockIdx.x) < num_tiles) {   // aux tile of the very first slab: overlaps the mainloop
      mbar_expect_tx(aux_bar, SLAB_BYTES);
      tma_load_2d(staging + SLAB_BYTES, &map_aux, aux_bar, (blockIdx.x / p.num_m_blocks) * BN, (blockIdx.x % p.num_m_blocks) * BLOCK_M);
    }
    // STATS instantiations (forward layout only) also collect the column statistics of the output: a separate kernel, so the
    // epilogue of every other GEMM keeps its register allocation and instruction stream (the runtime-flag version cost BERT 5 %)
    ColStatsAcc<BN> cstats;
    int stats_n0 = -1;
    if constexpr (STATS) cstats.clear();
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit % num_tiles;
      const int m0 = (tile % p.num_m_blocks) * BLOCK_M;
      const int n0 = (tile / p.num_m_blocks) * BN;
      if constexpr (STATS) {
        if (n0 != stats_n0) {                             // the CTA moves to another column block: publish what it has summed
          if (stats_n0 >= 0) cstats.flush(p.stats, p.N, stats_n0, threadIdx.x - 128);
          stats_n0 = n0;
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int slab = 0; slab < BN / SLAB_N; ++slab) {
        const int col0 = n0 + slab * SLAB_N;
        if (col0 >= p.N) break;                                   // whole slab out of range (N tail)
        uint8_t* buf_main = staging + (one_buffer ? 0 : (slab_counter & 1)) * SLAB_BYTES;
        uint8_t* buf_aux = staging + SLAB_BYTES;
        // the TMA store that last read this staging buffer must have finished reading it
        if (issuer) {
          if (one_buffer) tma_store_wait_read0(); else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
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
              if (c_base + i < p.N) {      // N is a multiple of 8 (16 B rows), so 8-wide groups are all-in or all-out
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
          if (p.store_preact) {
            float g[32];
            if (p.store_preact == 2) {
#pragma unroll
              for (int i = 0; i < 32; ++i) g[i] = apply_dact(f[i], p.act);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) g[i] = f[i];
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              const int chunk = half * 4 + (i >> 3);
              uint4 o;
              o.x = pack_bf16x2(g[i], g[i + 1]); o.y = pack_bf16x2(g[i + 2], g[i + 3]);
              o.z = pack_bf16x2(g[i + 4], g[i + 5]); o.w = pack_bf16x2(g[i + 6], g[i + 7]);
              *reinterpret_cast<uint4*>(buf_aux + row_in_tile * 128 + ((chunk ^ (row_in_tile & 7)) << 4)) = o;
            }
          }
          if (has_aux_in) {
            // auxiliary input slab (pre-activation z, or the gradient tile being accumulated into) arrives through TMA
            // (coalesced 128 B rows into swizzled smem); its load was issued one slab ahead, so this wait is usually free
            mbar_wait(aux_bar, aux_phase);
            aux_phase ^= 1;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              const int chunk = half * 4 + (i >> 3);
              const uint4 z4 = *reinterpret_cast<const uint4*>(buf_aux + row_in_tile * 128 + ((chunk ^ (row_in_tile & 7)) << 4));
              const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&z4);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 zf = __bfloat1622float2(z2[j]);
                if (p.aux_mode == 2) {
                  f[i + 2 * j] += zf.x;
                  f[i + 2 * j + 1] += zf.y;
                } else if (p.aux_mode == 3) {
                  f[i + 2 * j] *= zf.x;
                  f[i + 2 * j + 1] *= zf.y;
                } else {
                  f[i + 2 * j] *= apply_dact(zf.x, p.act);
                  f[i + 2 * j + 1] *= apply_dact(zf.y, p.act);
                }
              }
            }
          } else if (p.act != ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], p.act);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            const int chunk = half * 4 + (i >> 3);       // 16-byte chunk index inside the 128-byte slab row
            uint4 o;
            o.x = pack_bf16x2(f[i], f[i + 1]); o.y = pack_bf16x2(f[i + 2], f[i + 3]);
            o.z = pack_bf16x2(f[i + 4], f[i + 5]); o.w = pack_bf16x2(f[i + 6], f[i + 7]);
            *reinterpret_cast<uint4*>(buf_main + row_in_tile * 128 + ((chunk ^ (row_in_tile & 7)) << 4)) = o;
          }
        }
        if (slab == BN / SLAB_N - 1 || col0 + SLAB_N >= p.N) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, EPI_THREADS);
        if (issuer) {
          if (p.reduce_out) tma_reduce_add_2d(&map_d, buf_main, col0, m0); else tma_store_2d(&map_d, buf_main, col0, m0);
          if (p.store_preact) tma_store_2d(&map_aux, buf_aux, col0, m0);
          tma_store_commit();
          if (has_aux_in) {
            // every thread has consumed buf_aux (it passed the barrier above): prefetch the NEXT slab's aux tile now,
            // either the next slab of this tile or the first slab of this CTA's next tile
            int nt = tile, ns = slab + 1;
            if (ns == BN / SLAB_N || n0 + ns * SLAB_N >= p.N) { nt = tile + gridDim.x; ns = 0; }     // split_k == 1 here: unit == tile
            if (nt < num_tiles) {
              mbar_expect_tx(aux_bar, SLAB_BYTES);
              tma_load_2d(buf_aux, &map_aux, aux_bar, (nt / p.num_m_blocks) * BN + ns * SLAB_N, (nt % p.num_m_blocks) * BLOCK_M);
            }
          }
        }
        if constexpr (STATS) {  // column sums of the staged slab; rows beyond M are zero (TMA zero-fills the A tail, no bias / activation)
          float ds, dq;
          slab_col_stats<EPI_THREADS>(buf_main, threadIdx.x - 128, ds, dq);
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
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    cudaFree(nullptr);             // binds the primary context to THIS thread (autograd worker threads may not have one yet)
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (err != cudaSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2D bf16 tensor map: ``inner`` contiguous elements, ``outer`` rows with ``ld`` elements between rows.
static int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return -1;
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.  Threads that have only
  // ever been handed work by the autograd engine may not have touched the runtime yet -> CUDA_ERROR_INVALID_CONTEXT.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

template <int BN, bool A_MN, bool B_MN, bool STATS = false>
static int launch(const void* A, int lda, const void* B, int ldb, void* D, int ldd, void* aux_out, int ld_auxout,
                  void* b_copy_dst, int ld_bcopy, GemmParams p, int num_sms, cudaStream_t stream) {
  if (p.bcopy_flags == nullptr) b_copy_dst = nullptr;
  if (p.reduce_out) {                    // the accumulated-into tensor is never read: no aux slab
    p.dact_in = nullptr;
    p.aux_mode = 0;
  }
  if (p.dact_in != nullptr) {            // aux INPUT tile (mutually exclusive with the pre-activation output)
    aux_out = const_cast<__nv_bfloat16*>(p.dact_in);
    ld_auxout = p.ld_aux;
    p.store_preact = 0;
  }
  using L = SmemLayout<BN>;
  CUtensorMap ma, mb, md, mx, mc;
  int e = 0;
  p.b_copy = (b_copy_dst != nullptr && !B_MN) ? 1 : 0;
  p.b_peer = static_cast<const __nv_bfloat16*>(B);
  p.b_local = static_cast<__nv_bfloat16*>(b_copy_dst);
  p.ld_bpeer = ldb;
  p.ld_blocal = ld_bcopy;
  // A: K-major -> [M rows, K inner]; MN-major -> stored [K rows, M inner]
  e |= A_MN ? make_map(&ma, A, p.M, p.K, lda, 64, BLOCK_K) : make_map(&ma, A, p.K, p.M, lda, BLOCK_K, BLOCK_M);
  e |= B_MN ? make_map(&mb, B, p.N, p.K, ldb, 64, BLOCK_K) : make_map(&mb, B, p.K, p.N, ldb, BLOCK_K, BN);
  e |= make_map(&md, D, p.N, p.M, ldd, SLAB_N, BLOCK_M);
  e |= make_map(&mx, aux_out ? aux_out : D, p.N, p.M, aux_out ? ld_auxout : ldd, SLAB_N, BLOCK_M);
  e |= make_map(&mc, p.b_copy ? b_copy_dst : B, p.K, p.N, p.b_copy ? ld_bcopy : ldb, BLOCK_K, BN);
  if (e != 0) return 1000 + e;
  p.num_m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  auto kern = rvn_gemm_bf16_tcgen05<BN, A_MN, B_MN, STATS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (ce != cudaSuccess) return 2000 + static_cast<int>(ce);
    configured = true;
  }
  if (p.split_k < 1 || !p.reduce_out) p.split_k = 1;
  if (p.split_k > p.num_k_blocks) p.split_k = p.num_k_blocks;
  const int tiles = p.num_m_blocks * p.num_n_blocks * p.split_k;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, L::TOTAL, stream>>>(ma, mb, md, mx, mc, p);
  cudaError_t le = cudaGetLastError();
  return le == cudaSuccess ? 0 : 3000 + static_cast<int>(le);
}

}  // namespace rvn

// C entry point used by binding.cpp.  Layout flags: a_mn / b_mn = operand stored [K, rows] instead of [rows, K].
extern "C" int rvn_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* D, int ldd, int M, int N,
                             int K, const void* bias, int act, const void* dact_in, int ld_dact, int aux_mode, void* preact_out,
                             int ld_preact, int preact_mode, void* b_copy_dst, int ld_bcopy, unsigned* bcopy_flags, unsigned bcopy_epoch, int block_n,
                             int split_k, int reduce_out, float* stats, int num_sms, cudaStream_t stream) {
  rvn::GemmParams p{};
  p.stats = (split_k <= 1 && !reduce_out && bias == nullptr && act == 0 && dact_in == nullptr && preact_out == nullptr && !a_mn && !b_mn &&
             b_copy_dst == nullptr) ? stats : nullptr;
  if (stats != nullptr && p.stats == nullptr) return -7;     // column statistics: forward layout, plain store epilogue only
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.dact_in = static_cast<const __nv_bfloat16*>(dact_in);
  p.ld_aux = ld_dact;
  p.aux_mode = dact_in != nullptr ? aux_mode : 0;
  p.act = act;
  p.store_preact = preact_out != nullptr ? (preact_mode == 2 ? 2 : 1) : 0;
  p.bcopy_flags = bcopy_flags;
  p.bcopy_epoch = bcopy_epoch;
  p.split_k = split_k;
  p.reduce_out = reduce_out;
  if (p.stats != nullptr) {
    if (block_n == 256) return rvn::launch<256, false, false, true>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);
    if (block_n == 192) return rvn::launch<192, false, false, true>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);
    return rvn::launch<128, false, false, true>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);
  }
#define RVN_DISPATCH(BN_)                                                                                              \
  if (!a_mn && !b_mn) return rvn::launch<BN_, false, false>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream); \
  if (!a_mn && b_mn) return rvn::launch<BN_, false, true>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);   \
  if (a_mn && b_mn) return rvn::launch<BN_, true, true>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);     \
  return rvn::launch<BN_, true, false>(A, lda, B, ldb, D, ldd, preact_out, ld_preact, b_copy_dst, ld_bcopy, p, num_sms, stream);
  if (block_n == 256) { RVN_DISPATCH(256) }
  if (block_n == 192) { RVN_DISPATCH(192) }
  RVN_DISPATCH(128)
#undef RVN_DISPATCH
}
