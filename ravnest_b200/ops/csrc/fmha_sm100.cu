// K6: fused multi-head attention forward / backward on tcgen05 + TMEM (head_dim 64, bf16, fp32 softmax).
//
// Upstream materialises q@k^T, masks, softmaxes, drops out and multiplies by v as separate library calls
// (examples/sorter/mingpt/model_without_padding_mask.py:83-107; HF BertSelfAttention).  Here one CTA owns a
// 128-row query tile of one (batch, head):
//   control warp : TMA loads of Q / K_j / V_j from the packed [B*S, 3H] qkv activation, issues
//                  S = Q K_j^T  (UMMA 128x128x16, both operands K-major)  into TMEM columns [0,128)
//                  O_j = P V_j  (UMMA 128x64x16, P K-major from smem, V MN-major) into TMEM columns [128,192)
//   4 softmax warps (thread == query row == TMEM lane): tcgen05.ld S, scale + padding/causal mask,
//                  online softmax in the log2 domain, P -> bf16 -> 128B-swizzled smem (UMMA A operand),
//                  accumulate O in registers with the running rescale, finally O / l and LSE.
// S and P never touch HBM.  Backward recomputes S/P per tile and produces dQ, dK, dV with five UMMAs per
// (q tile, kv tile) pair; P and dS are written once to smem and read both K-major (dQ) and MN-major (dK, dV).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"
#include "sm100_ptx.cuh"

namespace rvn {

constexpr int FM_BLOCK = 128;        // query rows per CTA and keys per inner tile
constexpr int FM_D = 64;             // head dim
constexpr int FM_TILE_BYTES = FM_BLOCK * FM_D * 2;     // 16 KB: one Q / K / V tile (128 rows x 128 B)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct FmhaParams {
  int B, S, heads, H;                // H = heads * 64 (hidden); qkv row stride = 3H
  float scale_log2;                  // (1/sqrt(D)) * log2(e)
  const float* kbias;                // [B, S] additive key bias (0 / -inf) or nullptr
  int causal;
  __nv_bfloat16* out;                // [B*S, H]
  float* lse;                        // [B, heads, S] natural-log LSE of the scaled scores
  const long long* rng;              // {seed, step} on the device (dropout) or nullptr
  float dropout_p;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// smem: Q | K | V | P(2 x 16 KB) | barriers
struct FwdSmem {
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = FM_TILE_BYTES;
  static constexpr int V_OFF = 2 * FM_TILE_BYTES;
  static constexpr int P_OFF = 3 * FM_TILE_BYTES;
  static constexpr int BAR_OFF = 5 * FM_TILE_BYTES;
  static constexpr int XCHG_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = XCHG_OFF + 2048 + 64 + 1024;
};

constexpr int FM_SOFTMAX_THREADS = 256;      // 8 warps: thread = (query row, key half); warp w -> TMEM lanes 32*(w%4), half w/4
constexpr int FM_THREADS = FM_SOFTMAX_THREADS + 32;

// masked, scaled score in the log2 domain
__device__ __forceinline__ float fm_score(uint32_t raw, float scale_log2, float kbias_log2, bool masked) {
  const float s = fmaf(__uint_as_float(raw), scale_log2, kbias_log2);
  return masked ? -INFINITY : s;
}

__global__ void __launch_bounds__(FM_THREADS, 2)
rvn_fmha_fwd_tcgen05(const __grid_constant__ CUtensorMap map_qkv, const FmhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR_OFF);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_p = bars + 4;
  uint64_t* bar_o = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* xchg = reinterpret_cast<float*>(smem + FwdSmem::XCHG_OFF);      // [2][128] per-half row maxima / sums
  float* kb_s = xchg + 256;                                               // [2][128] additive key bias (log2 domain) of tile j / j+1
  int* kb_any = reinterpret_cast<int*>(kb_s + 256);                       // [2][4] "this tile has a non-zero bias", one word per staging warp

  const int warp = threadIdx.x >> 5;
  const int q0 = blockIdx.x * FM_BLOCK;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int row_base = b * p.S;                       // first row of this batch in the [B*S, 3H] view
  int n_tiles = (p.S + FM_BLOCK - 1) / FM_BLOCK;
  if (p.causal) {
    const int last = (q0 + FM_BLOCK - 1) / FM_BLOCK + 1;
    n_tiles = n_tiles < last ? n_tiles : last;
  }

  if (threadIdx.x == FM_SOFTMAX_THREADS) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, FM_SOFTMAX_THREADS);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 8) {
    // ============================================================== control: TMA + MMA issue
    if (elect_one()) {
      constexpr uint32_t IDESC_S = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t IDESC_O = make_idesc_bf16(128, 64, false, true);
      const uint32_t q_addr = smem_u32(smem + FwdSmem::Q_OFF);
      const uint32_t k_addr = smem_u32(smem + FwdSmem::K_OFF);
      const uint32_t v_addr = smem_u32(smem + FwdSmem::V_OFF);
      const uint32_t p_addr = smem_u32(smem + FwdSmem::P_OFF);
      mbar_expect_tx(bar_q, FM_TILE_BYTES);
      tma_load_2d(smem + FwdSmem::Q_OFF, &map_qkv, bar_q, h * FM_D, row_base + q0);
      mbar_expect_tx(bar_k, FM_TILE_BYTES);
      tma_load_2d(smem + FwdSmem::K_OFF, &map_qkv, bar_k, p.H + h * FM_D, row_base);
      mbar_expect_tx(bar_v, FM_TILE_BYTES);
      tma_load_2d(smem + FwdSmem::V_OFF, &map_qkv, bar_v, 2 * p.H + h * FM_D, row_base);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(bar_k, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < FM_D / 16; ++k)
          umma_bf16(tmem_s, make_smem_desc_sw128(q_addr + k * 32, 0, 1024), make_smem_desc_sw128(k_addr + k * 32, 0, 1024),
                    IDESC_S, k != 0);
        umma_commit(bar_s);
        mbar_wait(bar_s, ph);                       // S done -> K buffer is free: prefetch the next K tile
        if (j + 1 < n_tiles) {
          mbar_expect_tx(bar_k, FM_TILE_BYTES);
          tma_load_2d(smem + FwdSmem::K_OFF, &map_qkv, bar_k, p.H + h * FM_D, row_base + (j + 1) * FM_BLOCK);
        }
        mbar_wait(bar_p, ph);                       // softmax warps wrote P_j
        mbar_wait(bar_v, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < FM_BLOCK / 16; ++kk)
          umma_bf16(tmem_o, make_smem_desc_sw128(p_addr + (kk >> 2) * FM_TILE_BYTES + (kk & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(v_addr + kk * 2048, FM_TILE_BYTES, 1024), IDESC_O, kk != 0);
        umma_commit(bar_o);
        mbar_wait(bar_o, ph);                       // O_j done -> V and P buffers are free
        if (j + 1 < n_tiles) {
          mbar_expect_tx(bar_v, FM_TILE_BYTES);
          tma_load_2d(smem + FwdSmem::V_OFF, &map_qkv, bar_v, 2 * p.H + h * FM_D, row_base + (j + 1) * FM_BLOCK);
        }
      }
    }
    __syncwarp();
  } else {
    // ============================================================== softmax / accumulate
    const int half = warp >> 2;                     // which 64 keys of the tile / which 32 output dims
    const int r = (warp & 3) * 32 + lane_id();      // query row in the tile == TMEM lane
    const int q = q0 + r;
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint8_t* p_row = smem + FwdSmem::P_OFF + half * FM_TILE_BYTES + r * 128;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    const float* kb = p.kbias ? p.kbias + static_cast<size_t>(b) * p.S : nullptr;
    const bool use_drop = p.rng != nullptr && p.dropout_p > 0.f;
    DropoutRng drng;
    if (use_drop) drng.init(p.rng, p.dropout_p);
    // dropout groups: 8 consecutive keys of one (batch, head, query) row
    const unsigned long long drow = (static_cast<unsigned long long>(blockIdx.y) * p.S + q) * ((p.S + 7) / 8);
    // additive key bias of a tile, staged in smem once per tile: user padding bias (log2 domain) and -inf for keys >= S
    // ... and a per-tile flag: tiles whose keys all carry a zero bias (no padding in them — with padding at the end of the sequence
    // that is every tile but the last one or two) take the bias-free fast path of the softmax
    auto stage_bias = [&](int tile) {
      if (half == 0) {
        const int k = tile * FM_BLOCK + r;
        const float v = k >= p.S ? -INFINITY : (kb != nullptr ? kb[k] * LOG2E : 0.f);
        kb_s[(tile & 1) * 128 + r] = v;
        const int any = __any_sync(0xffffffffu, v != 0.f);
        if (lane_id() == 0) kb_any[(tile & 1) * 4 + (warp & 3)] = any;
      }
    };
    stage_bias(0);
    named_bar_sync(1, FM_SOFTMAX_THREADS);
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int kbase = j * FM_BLOCK + half * 64;    // first key handled by this thread
      const int* anyp = kb_any + (j & 1) * 4;
      const bool biased = (anyp[0] | anyp[1] | anyp[2] | anyp[3]) != 0;
      const bool diag = p.causal && (j * FM_BLOCK + FM_BLOCK - 1 > q0);
      const bool edge = biased || diag;
      const float* kbt = kb_s + (j & 1) * 128 + half * 64;
      if (j + 1 < n_tiles) stage_bias(j + 1);        // ordered before its first read by this tile's exchange barrier
      mbar_wait(bar_s, ph);
      tcgen05_fence_after();
      // ---- pass 1: partial row max over this thread's 64 keys
      float tmax = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_sel + half * 64 + c * 32, v);
        tmem_ld_wait();
        if (!edge) {
#pragma unroll
          for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(v[i]) * p.scale_log2);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int k = kbase + c * 32 + i;
            tmax = fmaxf(tmax, fm_score(v[i], p.scale_log2, kbt[c * 32 + i], diag && k > q));
          }
        }
      }
      xchg[half * 128 + r] = tmax;
      named_bar_sync(1, FM_SOFTMAX_THREADS);
      const float m_new = fmaxf(m, fmaxf(tmax, xchg[(half ^ 1) * 128 + r]));
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m - m_use);       // m == -inf -> 0
      // ---- pass 2: probabilities -> bf16 -> swizzled smem; partial row sum
      float rsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + lane_sel + half * 64 + c * 32, v);
        tmem_ld_wait();
        float pr[32];
        if (!edge) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pr[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_use));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int k = kbase + c * 32 + i;
            pr[i] = fast_exp2(fm_score(v[i], p.scale_log2, kbt[c * 32 + i], diag && k > q) - m_use);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rsum += pr[i];
        if (use_drop) {                                // the row sum stays un-dropped; P fed to the PV product is dropped
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float sc[8];
            drng.scales8(drow + ((kbase + c * 32 + i) >> 3), sc);
#pragma unroll
            for (int t = 0; t < 8; ++t) pr[i + t] *= sc[t];
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const int chunk = c * 4 + (i >> 3);          // 16-byte chunk 0..7 of this half-row (64 keys = one 128 B swizzle row)
          uint4 w;
          w.x = pack_bf16x2(pr[i], pr[i + 1]); w.y = pack_bf16x2(pr[i + 2], pr[i + 3]);
          w.z = pack_bf16x2(pr[i + 4], pr[i + 5]); w.w = pack_bf16x2(pr[i + 6], pr[i + 7]);
          *reinterpret_cast<uint4*>(p_row + ((chunk ^ (r & 7)) << 4)) = w;
        }
      }
      l = l * alpha + rsum;
      m = m_new;
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] *= alpha;
      tcgen05_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      // ---- O[:, half*32 .. +32] += (P V_j)
      mbar_wait(bar_o, ph);
      tcgen05_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32(tmem_o + lane_sel + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v[i]);
      }
      tcgen05_fence_before();
    }
    // combine the two halves' row sums
    xchg[half * 128 + r] = l;
    named_bar_sync(1, FM_SOFTMAX_THREADS);
    l += xchg[(half ^ 1) * 128 + r];
    if (q < p.S) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* dst = p.out + static_cast<size_t>(row_base + q) * p.H + h * FM_D + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 w;
        w.x = pack_bf16x2(o[i] * inv, o[i + 1] * inv); w.y = pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        w.z = pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv); w.w = pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = w;
      }
      if (half == 0) p.lse[(static_cast<size_t>(b) * p.heads + h) * p.S + q] = l > 0.f ? (m + log2f(l)) * LN2 : INFINITY;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
}

// ================================================================================================ backward
struct FmhaBwdParams {
  int B, S, heads, H;
  float scale, scale_log2;
  const float* kbias;
  int causal;
  const float* lse;                  // [B, heads, S]
  const float* delta;                // [B, heads, S]  rowsum(dO * O)
  float* dq_accum;                   // [B*S, H] fp32, zero-initialised (atomic accumulation over kv tiles)
  __nv_bfloat16* dqkv;               // [B*S, 3H]  (dK, dV written here; dQ cast in afterwards)
  const long long* rng;
  float dropout_p;
};

// delta[b,h,q] = sum_d dO[q, h*64+d] * O[q, h*64+d]; one thread per (row, head)
__global__ void __launch_bounds__(256) rvn_fmha_bwd_prep(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                                         float* __restrict__ delta, int B, int S, int heads) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S * heads) return;
  const int h = idx % heads;
  const int row = idx / heads;
  const size_t off = static_cast<size_t>(row) * heads * FM_D + h * FM_D;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < FM_D; i += 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + off + i);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + off + i);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(a2[j]);
      const float2 y = __bfloat1622float2(g2[j]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  const int b = row / S, q = row % S;
  delta[(static_cast<size_t>(b) * heads + h) * S + q] = acc;
}

// dqkv[:, 0:H] = bf16(dq_accum)
__global__ void __launch_bounds__(256) rvn_fmha_bwd_cast_dq(const float* __restrict__ dq, __nv_bfloat16* __restrict__ dqkv, size_t rows, int H) {
  const size_t n4 = rows * (H / 4);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / (H / 4);
    const int c = static_cast<int>(i % (H / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(dq + row * H + c);
    uint2 w;
    w.x = pack_bf16x2(v.x, v.y);
    w.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dqkv + row * 3 * H + c) = w;
  }
}

// smem: K | V | Q | dO | P (2 x 16 KB) | dS (2 x 16 KB) | dQ staging (2 x 16 KB fp32, one 128 x 32 box per dim half) | barriers
struct BwdSmem {
  static constexpr int K_OFF = 0;
  static constexpr int V_OFF = FM_TILE_BYTES;
  static constexpr int Q_OFF = 2 * FM_TILE_BYTES;
  static constexpr int DO_OFF = 3 * FM_TILE_BYTES;
  static constexpr int P_OFF = 4 * FM_TILE_BYTES;
  static constexpr int DS_OFF = 6 * FM_TILE_BYTES;
  static constexpr int DQ_OFF = 8 * FM_TILE_BYTES;
  static constexpr int BAR_OFF = 10 * FM_TILE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 512 + 64 + 1024;
};

__global__ void __launch_bounds__(FM_THREADS, 1)
rvn_fmha_bwd_tcgen05(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                     const __grid_constant__ CUtensorMap map_dq, const FmhaBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BAR_OFF);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_qd = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_g = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int k0 = blockIdx.x * FM_BLOCK;
  const int b = blockIdx.y / p.heads;
  const int h = blockIdx.y % p.heads;
  const int row_base = b * p.S;
  const int n_q_tiles = (p.S + FM_BLOCK - 1) / FM_BLOCK;
  const int i_first = p.causal ? blockIdx.x : 0;          // causal: only q tiles at or after this kv tile

  if (threadIdx.x == FM_SOFTMAX_THREADS) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_dq);
    mbar_init(bar_kv, 1);
    mbar_init(bar_qd, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, FM_SOFTMAX_THREADS);
    mbar_init(bar_g, 1);
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320, tm_dq = tmem_base + 384;

  if (warp == 8) {
    if (elect_one()) {
      constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, false, false);   // S, dP : both K-major
      constexpr uint32_t IDESC_TT = make_idesc_bf16(128, 64, true, true);      // dV, dK: both MN-major
      constexpr uint32_t IDESC_DQ = make_idesc_bf16(128, 64, false, true);     // dQ    : dS K-major, K MN-major
      const uint32_t k_addr = smem_u32(smem + BwdSmem::K_OFF), v_addr = smem_u32(smem + BwdSmem::V_OFF);
      const uint32_t q_addr = smem_u32(smem + BwdSmem::Q_OFF), do_addr = smem_u32(smem + BwdSmem::DO_OFF);
      const uint32_t p_addr = smem_u32(smem + BwdSmem::P_OFF), ds_addr = smem_u32(smem + BwdSmem::DS_OFF);
      mbar_expect_tx(bar_kv, 2 * FM_TILE_BYTES);
      tma_load_2d(smem + BwdSmem::K_OFF, &map_qkv, bar_kv, p.H + h * FM_D, row_base + k0);
      tma_load_2d(smem + BwdSmem::V_OFF, &map_qkv, bar_kv, 2 * p.H + h * FM_D, row_base + k0);
      mbar_expect_tx(bar_qd, 2 * FM_TILE_BYTES);
      tma_load_2d(smem + BwdSmem::Q_OFF, &map_qkv, bar_qd, h * FM_D, row_base + i_first * FM_BLOCK);
      tma_load_2d(smem + BwdSmem::DO_OFF, &map_do, bar_qd, h * FM_D, row_base + i_first * FM_BLOCK);
      mbar_wait(bar_kv, 0);
      int it = 0;
      for (int i = i_first; i < n_q_tiles; ++i, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(bar_qd, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < FM_D / 16; ++k)
          umma_bf16(tm_s, make_smem_desc_sw128(q_addr + k * 32, 0, 1024), make_smem_desc_sw128(k_addr + k * 32, 0, 1024), IDESC_QK, k != 0);
#pragma unroll
        for (int k = 0; k < FM_D / 16; ++k)
          umma_bf16(tm_dp, make_smem_desc_sw128(do_addr + k * 32, 0, 1024), make_smem_desc_sw128(v_addr + k * 32, 0, 1024), IDESC_QK, k != 0);
        umma_commit(bar_s);
        mbar_wait(bar_p, ph);                        // P and dS are in smem
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < FM_BLOCK / 16; ++kk) {   // reduction over the 128 query rows, 16 at a time
          umma_bf16(tm_dv, make_smem_desc_sw128(p_addr + kk * 2048, FM_TILE_BYTES, 1024),
                    make_smem_desc_sw128(do_addr + kk * 2048, FM_TILE_BYTES, 1024), IDESC_TT, (it | kk) != 0);
          umma_bf16(tm_dk, make_smem_desc_sw128(ds_addr + kk * 2048, FM_TILE_BYTES, 1024),
                    make_smem_desc_sw128(q_addr + kk * 2048, FM_TILE_BYTES, 1024), IDESC_TT, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < FM_BLOCK / 16; ++kk)     // reduction over the 128 keys
          umma_bf16(tm_dq, make_smem_desc_sw128(ds_addr + (kk >> 2) * FM_TILE_BYTES + (kk & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(k_addr + kk * 2048, FM_TILE_BYTES, 1024), IDESC_DQ, kk != 0);
        umma_commit(bar_g);
        mbar_wait(bar_g, ph);                        // all five products retired: Q / dO / P / dS buffers reusable
        if (i + 1 < n_q_tiles) {
          mbar_expect_tx(bar_qd, 2 * FM_TILE_BYTES);
          tma_load_2d(smem + BwdSmem::Q_OFF, &map_qkv, bar_qd, h * FM_D, row_base + (i + 1) * FM_BLOCK);
          tma_load_2d(smem + BwdSmem::DO_OFF, &map_do, bar_qd, h * FM_D, row_base + (i + 1) * FM_BLOCK);
        }
      }
    }
    __syncwarp();
  } else {
    const int half = warp >> 2;                     // which 64 keys (P / dS) and which 32 dims (dQ, dK, dV) this thread owns
    const int r = (warp & 3) * 32 + lane_id();
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row_off = half * FM_TILE_BYTES + r * 128;
    uint8_t* p_smem = smem + BwdSmem::P_OFF + row_off;
    uint8_t* ds_smem = smem + BwdSmem::DS_OFF + row_off;
    const float* kb = p.kbias ? p.kbias + static_cast<size_t>(b) * p.S : nullptr;
    const size_t stat_base = (static_cast<size_t>(b) * p.heads + h) * p.S;
    const bool use_drop = p.rng != nullptr && p.dropout_p > 0.f;
    DropoutRng drng;
    if (use_drop) drng.init(p.rng, p.dropout_p);
    const int kbase = k0 + half * 64;
    float* kb_s = reinterpret_cast<float*>(smem + BwdSmem::BAR_OFF + 128);      // [128] key bias of this CTA's kv tile
    int* kb_any = reinterpret_cast<int*>(kb_s + 128);                           // [4] non-zero bias seen by staging warp w
    if (half == 0) {
      const int k = k0 + r;
      const float v = k >= p.S ? -INFINITY : (kb != nullptr ? kb[k] * LOG2E : 0.f);
      kb_s[r] = v;
      const int any = __any_sync(0xffffffffu, v != 0.f);
      if (lane_id() == 0) kb_any[warp & 3] = any;
    }
    named_bar_sync(1, FM_SOFTMAX_THREADS);
    const float* kbt = kb_s + half * 64;
    const bool biased = (kb_any[0] | kb_any[1] | kb_any[2] | kb_any[3]) != 0;     // no padded key in this kv tile: bias-free fast path
    int it = 0;
    for (int i = i_first; i < n_q_tiles; ++i, ++it) {
      const uint32_t ph = it & 1;
      const int q = i * FM_BLOCK + r;
      const bool q_ok = q < p.S;
      const float lse2 = q_ok ? p.lse[stat_base + q] * LOG2E : INFINITY;
      const float dlt = q_ok ? p.delta[stat_base + q] : 0.f;
      const unsigned long long drow = (static_cast<unsigned long long>(blockIdx.y) * p.S + q) * ((p.S + 7) / 8);
      const bool diag = p.causal && (k0 + FM_BLOCK - 1 > i * FM_BLOCK);
      const bool edge = biased || diag || (i * FM_BLOCK + FM_BLOCK > p.S);
      mbar_wait(bar_s, ph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tm_s + lane_sel + half * 64 + c * 32, sv);
        tmem_ld_32x32(tm_dp + lane_sel + half * 64 + c * 32, dv);
        tmem_ld_wait();
        float pr[32], ds[32];
        if (!edge) {
#pragma unroll
          for (int j = 0; j < 32; ++j) pr[j] = fast_exp2(fmaf(__uint_as_float(sv[j]), p.scale_log2, -lse2));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int k = kbase + c * 32 + j;
            const bool masked = !q_ok || (diag && k > q);
            pr[j] = masked ? 0.f : fast_exp2(fmaf(__uint_as_float(sv[j]), p.scale_log2, kbt[c * 32 + j]) - lse2);
          }
        }
        if (use_drop) {                                // Z = P o M feeds dV; dP = dZ o M
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float sc[8];
            drng.scales8(drow + ((kbase + c * 32 + j) >> 3), sc);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float pj = pr[j + t];
              ds[j + t] = pj * (__uint_as_float(dv[j + t]) * sc[t] - dlt) * p.scale;
              pr[j + t] = pj * sc[t];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) ds[j] = pr[j] * (__uint_as_float(dv[j]) - dlt) * p.scale;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int off = ((c * 4 + (j >> 3)) ^ (r & 7)) << 4;
          uint4 w;
          w.x = pack_bf16x2(pr[j], pr[j + 1]); w.y = pack_bf16x2(pr[j + 2], pr[j + 3]);
          w.z = pack_bf16x2(pr[j + 4], pr[j + 5]); w.w = pack_bf16x2(pr[j + 6], pr[j + 7]);
          *reinterpret_cast<uint4*>(p_smem + off) = w;
          w.x = pack_bf16x2(ds[j], ds[j + 1]); w.y = pack_bf16x2(ds[j + 2], ds[j + 3]);
          w.z = pack_bf16x2(ds[j + 4], ds[j + 5]); w.w = pack_bf16x2(ds[j + 6], ds[j + 7]);
          *reinterpret_cast<uint4*>(ds_smem + off) = w;
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      mbar_wait(bar_g, ph);
      tcgen05_fence_after();
      // dQ_i[:, half*32 .. +32] += dS K_j (this kv tile's contribution): the 128 x 32 fp32 tile of this dim half is staged in
      // swizzled smem and added into the fp32 accumulator by ONE bulk reduction (cp.reduce.async.bulk.tensor .add at the L2)
      // instead of 32 scalar atomics per thread (8192 RED instructions per tile pair: the L2 atomic rate was the bottleneck)
      {
        uint8_t* dq_s = smem + BwdSmem::DQ_OFF + half * FM_TILE_BYTES;
        const bool dq_issuer = (warp & 3) == 0 && lane_id() == 0;
        if (dq_issuer) tma_store_wait_read0();             // the previous pair's reduction has finished reading the staging tile
        named_bar_sync(2 + half, 128);
        uint32_t v[32];
        tmem_ld_32x32(tm_dq + lane_sel + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          uint4 w;
          w.x = q_ok ? v[j] : 0u; w.y = q_ok ? v[j + 1] : 0u; w.z = q_ok ? v[j + 2] : 0u; w.w = q_ok ? v[j + 3] : 0u;
          *reinterpret_cast<uint4*>(dq_s + r * 128 + (((j >> 2) ^ (r & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();
        named_bar_sync(2 + half, 128);
        if (dq_issuer) {
          tma_reduce_add_2d(&map_dq, dq_s, h * FM_D + half * 32, row_base + i * FM_BLOCK);
          tma_store_commit();
        }
      }
      tcgen05_fence_before();
    }
    if ((warp & 3) == 0 && lane_id() == 0) tma_store_wait_all();      // the last dQ reductions have landed before the CTA exits
    // dK_j, dV_j: thread == (key row, 32-dim half)
    const int k = k0 + r;
    if (it > 0) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(row_base + k) * 3 * p.H + (which == 0 ? p.H : 2 * p.H) + h * FM_D + half * 32;
        uint32_t v[32];
        tmem_ld_32x32((which == 0 ? tm_dk : tm_dv) + lane_sel + half * 32, v);
        tmem_ld_wait();
        if (k < p.S) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            w.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            w.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            w.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
            *reinterpret_cast<uint4*>(dst + j) = w;
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                       uint32_t box_outer) {
  static EncodeTiledFn2 enc = nullptr;
  if (enc == nullptr) {
    cudaFree(nullptr);             // binds the primary context to THIS thread (autograd worker threads may not have one yet)
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || ptr == nullptr) return -1;
    enc = reinterpret_cast<EncodeTiledFn2>(ptr);
  }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {               // driver call below needs a context current on this thread
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

// fp32 [rows, H] accumulator, 128 x 32 boxes (128 B rows, 128B swizzle) for the dQ bulk reductions
static int make_map_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                        uint32_t box_outer) {
  CUtensorMap tmp;
  if (make_map_2d(&tmp, base, 8, 8, 8, 8, 8) != 0) return -1;      // resolves the driver entry point / binds the context
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || ptr == nullptr) return -1;
  EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(ptr);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

}  // namespace rvn

// qkv: [B*S, 3H] bf16 contiguous; out: [B*S, H]; lse: [B, heads, S]; kbias: [B, S] fp32 or nullptr
extern "C" int rvn_fmha_forward(const void* qkv, void* out, float* lse, const float* kbias, int B, int S, int heads, int causal,
                                float scale, const long long* rng, float dropout_p, cudaStream_t stream) {
  using namespace rvn;
  const int H = heads * FM_D;
  CUtensorMap map;
  if (make_map_2d(&map, qkv, 3ull * H, static_cast<uint64_t>(B) * S, 3ull * H, FM_D, FM_BLOCK) != 0) return 1000;
  FmhaParams p{};
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.scale_log2 = scale * LOG2E;
  p.kbias = kbias;
  p.causal = causal;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.rng = rng;
  p.dropout_p = dropout_p;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(rvn_fmha_fwd_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL) != cudaSuccess) return 2000;
    configured = true;
  }
  dim3 grid((S + FM_BLOCK - 1) / FM_BLOCK, B * heads);
  rvn_fmha_fwd_tcgen05<<<grid, FM_THREADS, FwdSmem::TOTAL, stream>>>(map, p);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

// qkv [B*S,3H], o / d_o [B*S,H] bf16; lse [B,heads,S]; dqkv [B*S,3H] (fully written); dq_accum [B*S,H] fp32 zeroed; delta scratch
extern "C" int rvn_fmha_backward(const void* qkv, const void* o, const void* d_o, const float* lse, const float* kbias, float* delta,
                                 float* dq_accum, void* dqkv, int B, int S, int heads, int causal, float scale, const long long* rng,
                                 float dropout_p, int num_sms, cudaStream_t stream) {
  using namespace rvn;
  const int H = heads * FM_D;
  CUtensorMap map_qkv, map_do, map_dq;
  if (make_map_2d(&map_qkv, qkv, 3ull * H, static_cast<uint64_t>(B) * S, 3ull * H, FM_D, FM_BLOCK) != 0) return 1000;
  if (make_map_2d(&map_do, d_o, H, static_cast<uint64_t>(B) * S, H, FM_D, FM_BLOCK) != 0) return 1001;
  if (make_map_f32(&map_dq, dq_accum, H, static_cast<uint64_t>(B) * S, H, 32, FM_BLOCK) != 0) return 1002;
  const int n = B * S * heads;
  rvn_fmha_bwd_prep<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o),
                                                       delta, B, S, heads);
  FmhaBwdParams p{};
  p.B = B; p.S = S; p.heads = heads; p.H = H;
  p.scale = scale;
  p.scale_log2 = scale * LOG2E;
  p.kbias = kbias;
  p.causal = causal;
  p.lse = lse;
  p.delta = delta;
  p.dq_accum = dq_accum;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.rng = rng;
  p.dropout_p = dropout_p;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(rvn_fmha_bwd_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL) != cudaSuccess) return 2000;
    configured = true;
  }
  dim3 grid((S + FM_BLOCK - 1) / FM_BLOCK, B * heads);
  rvn_fmha_bwd_tcgen05<<<grid, FM_THREADS, BwdSmem::TOTAL, stream>>>(map_qkv, map_do, map_dq, p);
  const size_t rows = static_cast<size_t>(B) * S;
  int blocks = num_sms * 4;
  rvn_fmha_bwd_cast_dq<<<blocks, 256, 0, stream>>>(dq_accum, static_cast<__nv_bfloat16*>(dqkv), rows, H);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}
