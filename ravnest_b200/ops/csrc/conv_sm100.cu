// K4: NHWC bf16 convolution as an IMPLICIT GEMM on tcgen05 (no im2col buffer ever exists).
//
//   y[n, ho, wo, co] = sum_{r, s, ci} x[n, ho*stride + r*dil - pad, wo*stride + s*dil - pad, ci] * w[co, r, s, ci]
//
// GEMM view: M = output pixels, N = Cout, K = (r, s, ci).  A CTA tile is a TN x TH x TW patch of output pixels (<= 128
// rows).  For filter tap (r, s) and a 64-channel slice, the A operand of that patch is the SAME patch of input pixels
// shifted by the tap — one 4-D TMA box load {64 ch, TW, TH, TN} (element strides = conv stride) whose out-of-bounds
// elements TMA zero-fills: padding costs nothing and there is no per-element address arithmetic anywhere.
// B (weights, [Cout, R*S, Cin]) is a 3-D TMA box {64 ch, 1 tap, BN}.  The K loop walks taps x channel slices.
// Everything downstream of the operand fetch is the GEMM pipeline of gemm_sm100.cu: smem ring -> tcgen05.mma (UMMA
// 128 x BN x 16) -> double-buffered TMEM accumulators -> tcgen05.ld epilogue -> bf16 -> swizzled smem -> 4-D TMA store
// of the output patch (clipped at the image borders by TMA).
//
// The same kernel computes dgrad of stride-1 convolutions (dx = conv(dy, flipped taps) with the weight read MN-major
// through the [Cin, R*S, Cout] view), see rvn_conv_forward's `transposed` flag.
// Reference call sites: every nn.Conv2d of ResNet-50 / Inception-V3 / CNN_Net, cuDNN upstream.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace rvn {

constexpr int CV_BLOCK_M = 128;
constexpr int CV_BLOCK_K = 64;
constexpr int CV_THREADS = 256;
constexpr int CV_EPI_THREADS = 128;
constexpr int CV_SLAB_N = 64;
constexpr int CV_SLAB_BYTES = CV_BLOCK_M * CV_SLAB_N * 2;

struct ConvParams {
  int N, H, W, Cin, Ho, Wo, Cout;
  int R, S, stride, pad, dil;
  int pad_w;                      // horizontal padding (pad = vertical): rectangular kernels (1x7, 7x1, 1x3, 3x1)
  int TW, TH, TN;                 // output patch of one CTA tile (TW*TH*TN <= 128)
  int tiles_w, tiles_h, tiles_n;  // patches per dimension
  int num_n_blocks;               // Cout / BN (rounded up)
  int cin_blocks;                 // ceil(Cin / 64)
  int flip;                       // 1: walk taps mirrored (dgrad)
  const __nv_bfloat16* bias;      // [Cout] or nullptr
  int relu;
  long long xs_w, xs_h, xs_n;     // host only: element strides of the input's W / H / N dimensions (0 = dense NHWC).  They may
                                  // OVERLAP (xs_w < Cin): the small-Cin stem reads 8 neighbouring 8-channel pixels as one 64-wide K block
  float* stats;                   // [2][Cout] fp32 or nullptr: per-channel sum / sum of squares of the (bf16) output are ADDED here
                                  // by the epilogue — the batch statistics of the BatchNorm that follows, without a pass over y
  // one output-parity phase (a, b) of a STRIDE-2 data gradient (rvn_conv_dgrad_s2): dx[2h'+a, 2w'+b] only receives the filter taps
  // r = r0 + 2i, s = s0 + 2j (r0 = (a + pad) % 2, ...), read from dy[h' + ch - i, w' + cw - j] — a stride-1 convolution of dy with a
  // sub-sampled filter, written through a tensor map that addresses every second pixel of dx
  int sub;                        // 1: phase mode
  int r0, s0, Ra, Sb, ch, cw;
  long long ys_w, ys_h, ys_n;     // host only: element strides of the output's W / H / N dimensions (0 = dense NHWC)
};

// 4-D / 3-D TMA loads and the 4-D store
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// shared -> global bulk reduction (element-wise add at the L2), 3-D tile, element type from the tensor map
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int BN>
struct ConvSmem {
  static constexpr int A_BYTES = CV_BLOCK_M * CV_BLOCK_K * 2;
  static constexpr int B_BYTES = BN * CV_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int OPERAND_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = OPERAND_BYTES + 2 * CV_SLAB_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, bool B_MN, bool STATS = false>
__global__ void __launch_bounds__(CV_THREADS, 1)
rvn_conv_implicit_gemm_tcgen05(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                               const __grid_constant__ CUtensorMap map_y, const ConvParams p) {
  using L = ConvSmem<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = make_idesc_bf16(CV_BLOCK_M, BN, false, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::OPERAND_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_tiles = m_tiles * p.num_n_blocks;
  const int num_k_blocks = (p.sub ? p.Ra * p.Sb : p.R * p.S) * p.cin_blocks;
  const int rows = p.TW * p.TH * p.TN;                    // valid rows of an A tile / output patch
  const uint32_t a_bytes = static_cast<uint32_t>(rows) * 128u;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], CV_EPI_THREADS / 32);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 4 && rows < CV_BLOCK_M) {
    // rows beyond the patch are never loaded by TMA: zero them once so the MMA never multiplies stale NaN patterns
    for (int s = 0; s < STAGES; ++s) {
      uint4* dst = reinterpret_cast<uint4*>(smem + s * L::STAGE_BYTES + rows * 128);
      for (int i = threadIdx.x - 128; i < (CV_BLOCK_M - rows) * 8; i += CV_EPI_THREADS) dst[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int& wo0, int& ho0, int& n0, int& co0) {
    const int mt = tile % m_tiles;
    co0 = (tile / m_tiles) * BN;
    wo0 = (mt % p.tiles_w) * p.TW;
    ho0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
    n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
  };

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int wo0, ho0, n0, co0;
        decode(tile, wo0, ho0, n0, co0);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          int tap = kb / p.cin_blocks;
          const int c0 = (kb % p.cin_blocks) * CV_BLOCK_K;
          int xw, xh;
          if (p.sub) {                                           // stride-2 dgrad phase: sub-sampled filter, see ConvParams
            const int i = tap / p.Sb, j = tap % p.Sb;
            tap = (p.r0 + 2 * i) * p.S + (p.s0 + 2 * j);
            xh = ho0 + p.ch - i;
            xw = wo0 + p.cw - j;
          } else {
            int r = tap / p.S, s = tap % p.S;
            if (p.flip) { r = p.R - 1 - r; s = p.S - 1 - s; }
            xw = wo0 * p.stride + s * p.dil - p.pad_w;
            xh = ho0 * p.stride + r * p.dil - p.pad;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], a_bytes + L::B_BYTES);
          // input patch of this tap: coordinates may be negative / beyond the image -> TMA zero fill == padding
          tma_load_4d(sa, &map_x, &full_bar[stage], c0, xw, xh, n0);
          if constexpr (!B_MN) {
            tma_load_3d(sb, &map_w, &full_bar[stage], c0, tap, co0);                       // {64 ci, 1 tap, BN co}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)                                               // weight viewed [ci(K rows)][tap][co]
              tma_load_3d(sb + j * (CV_BLOCK_K * 128), &map_w, &full_bar[stage], co0 + j * 64, tap, c0);
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
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < CV_BLOCK_K / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, CV_BLOCK_K * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16(tmem_d, da, db, IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue
    const int ewarp = warp - 4;
    const int row_in_tile = ewarp * 32 + lane_id();
    const bool issuer = threadIdx.x == 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_counter = 0;
    const int et = threadIdx.x - 128;
    ColStatsAcc<BN> cstats;           // STATS instantiations only: per-channel sum / sum of squares of the output (BatchNorm)
    int stats_co0 = -1;
    if constexpr (STATS) cstats.clear();
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int wo0, ho0, n0, co0;
      decode(tile, wo0, ho0, n0, co0);
      bool row_ok = true;
      if constexpr (STATS) {
        if (co0 != stats_co0) {                              // the CTA moves to another channel block: publish what it has summed
          if (stats_co0 >= 0) cstats.flush(p.stats, p.Cout, stats_co0, et);
          stats_co0 = co0;
        }
        // patch rows outside the output image (clipped by the TMA store) may hold non-zero accumulators: keep them out of the sums
        const int tw = row_in_tile % p.TW, th = (row_in_tile / p.TW) % p.TH, tn = row_in_tile / (p.TW * p.TH);
        row_ok = row_in_tile < rows && wo0 + tw < p.Wo && ho0 + th < p.Ho && n0 + tn < p.N;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int slab = 0; slab < BN / CV_SLAB_N; ++slab) {
        const int col0 = co0 + slab * CV_SLAB_N;
        if (col0 >= p.Cout) break;
        uint8_t* buf = staging + (slab_counter & 1) * CV_SLAB_BYTES;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        named_bar_sync(1, CV_EPI_THREADS);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + acc * BN + slab * CV_SLAB_N + half * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          const int c_base = col0 + half * 32;
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              if (c_base + i < p.Cout) {
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
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if constexpr (STATS) {
            if (!row_ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = 0.f;
            }
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
        if (slab == BN / CV_SLAB_N - 1 || col0 + CV_SLAB_N >= p.Cout) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, CV_EPI_THREADS);
        if (issuer) {
          tma_store_4d(&map_y, buf, col0, wo0, ho0, n0);        // rows of the patch in (n, h, w) order; clipped at borders
          tma_store_commit();
        }
        if constexpr (STATS) {                                  // column sums of the staged slab (the store only reads it too)
          float ds, dq;
          slab_col_stats<CV_EPI_THREADS>(buf, et, ds, dq);
          cstats.add(slab, ds, dq);
        }
        ++slab_counter;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (STATS) {
      if (stats_co0 >= 0) cstats.flush(p.stats, p.Cout, stats_co0, et);
    }
    if (issuer) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn3 conv_encode() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn == nullptr) {
    cudaFree(nullptr);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn3>(ptr);
  }
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  return fn;
}

// generic bf16 tensor map: dims / strides given innermost-first (strides in ELEMENTS for dims 1..rank-1)
static int make_map_nd(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                       const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn3 enc = conv_encode();
  if (enc == nullptr) return -1;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_elems[i] * 2;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// fp32 variant (weight-gradient accumulator): same encoder, 4-byte elements
static int make_map_nd_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                           const uint32_t* box) {
  EncodeTiledFn3 enc = conv_encode();
  if (enc == nullptr) return -1;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_elems[i] * 4;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// pick the output patch (TW, TH, TN), TW*TH*TN <= 128, that wastes the fewest MMA rows for this layer
static void pick_patch(int Wo, int Ho, int N, int& TW, int& TH, int& TN) {
  double best = -1.0;
  for (int tw = 1; tw <= 128 && tw <= Wo + 0; ++tw) {
    if (tw > 64) break;
    for (int th = 1; tw * th <= 128 && th <= Ho; ++th) {
      int tn = 128 / (tw * th);
      if (tn > N) tn = N;
      if (tn < 1) continue;
      const double tiles = static_cast<double>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((N + tn - 1) / tn);
      const double eff = static_cast<double>(Wo) * Ho * N / (tiles * 128.0);
      if (eff > best + 1e-9) { best = eff; TW = tw; TH = th; TN = tn; }
    }
  }
}

template <int BN, bool B_MN, bool STATS = false>
static int conv_launch(const void* x, const void* w, void* y, ConvParams p, int num_sms, cudaStream_t stream) {
  using L = ConvSmem<BN>;
  CUtensorMap mx, mw, my;
  pick_patch(p.Wo, p.Ho, p.N, p.TW, p.TH, p.TN);
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.tiles_n = (p.N + p.TN - 1) / p.TN;
  p.num_n_blocks = (p.Cout + BN - 1) / BN;
  p.cin_blocks = (p.Cin + CV_BLOCK_K - 1) / CV_BLOCK_K;
  int e = 0;
  {   // x: [N, H, W, Cin] -> dims (Cin, W, H, N); box walks `stride` elements per output pixel
    const uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t st[3] = {(uint64_t)p.Cin, (uint64_t)p.W * p.Cin, (uint64_t)p.H * p.W * p.Cin};
    if (p.xs_w > 0) { st[0] = (uint64_t)p.xs_w; st[1] = (uint64_t)p.xs_h; st[2] = (uint64_t)p.xs_n; }
    const uint32_t box[4] = {64u, (uint32_t)((p.TW - 1) * p.stride + 1), (uint32_t)((p.TH - 1) * p.stride + 1), (uint32_t)p.TN};
    const uint32_t es[4] = {1u, (uint32_t)p.stride, (uint32_t)p.stride, 1u};
    e |= make_map_nd(&mx, x, 4, dims, st, box, es);
  }
  if (!B_MN) {   // w: [Cout, R*S, Cin] -> dims (Cin, RS, Cout), box {64, 1, BN}
    const uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.R * p.S, (uint64_t)p.Cout};
    const uint64_t st[2] = {(uint64_t)p.Cin, (uint64_t)p.R * p.S * p.Cin};
    const uint32_t box[3] = {64u, 1u, (uint32_t)BN};
    const uint32_t es[3] = {1u, 1u, 1u};
    e |= make_map_nd(&mw, w, 3, dims, st, box, es);
  } else {       // dgrad: weight stored [K_out = old Cout... see rvn_conv_forward]: dims (Cout_gemm, RS, Cin_gemm), box {64, 1, 64}
    const uint64_t dims[3] = {(uint64_t)p.Cout, (uint64_t)p.R * p.S, (uint64_t)p.Cin};
    const uint64_t st[2] = {(uint64_t)p.Cout, (uint64_t)p.R * p.S * p.Cout};
    const uint32_t box[3] = {64u, 1u, 64u};
    const uint32_t es[3] = {1u, 1u, 1u};
    e |= make_map_nd(&mw, w, 3, dims, st, box, es);
  }
  {   // y: [N, Ho, Wo, Cout] -> dims (Cout, Wo, Ho, N), box {64, TW, TH, TN}
    const uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    uint64_t st[3] = {(uint64_t)p.Cout, (uint64_t)p.Wo * p.Cout, (uint64_t)p.Ho * p.Wo * p.Cout};
    if (p.ys_w > 0) { st[0] = (uint64_t)p.ys_w; st[1] = (uint64_t)p.ys_h; st[2] = (uint64_t)p.ys_n; }
    const uint32_t box[4] = {64u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    const uint32_t es[4] = {1u, 1u, 1u, 1u};
    e |= make_map_nd(&my, y, 4, dims, st, box, es);
  }
  if (e != 0) return 1000 + e;
  auto kern = rvn_conv_implicit_gemm_tcgen05<BN, B_MN, STATS>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return 2000;
    configured = true;
  }
  const int tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.num_n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, CV_THREADS, L::TOTAL, stream>>>(mx, mw, my, p);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

}  // namespace rvn

// x [N,H,W,Cin], w [Cout,R,S,Cin] (or, transposed=1, the SAME buffer consumed as a map from Cout_in -> Cin_out for dgrad:
// then `Cin`/`Cout` below are the GEMM K / N channel counts, i.e. Cin = channels of the incoming gradient), y [N,Ho,Wo,Cout].
// xs_w / xs_h / xs_n: element strides of x's W / H / N dimensions (0 = dense NHWC); the channel stride is always 1.
extern "C" int rvn_conv_forward_strided(const void* x, const void* w, void* y, const void* bias, int relu, int N, int H, int W, int Cin,
                                        int Ho, int Wo, int Cout, int R, int S, int stride, int pad, int pad_w, int dil, int transposed,
                                        long long xs_w, long long xs_h, long long xs_n, float* stats, int num_sms,
                                        cudaStream_t stream) {
  rvn::ConvParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.R = R; p.S = S; p.stride = stride; p.pad = pad; p.pad_w = pad_w; p.dil = dil;
  p.xs_w = xs_w; p.xs_h = xs_h; p.xs_n = xs_n;
  p.stats = transposed ? nullptr : stats;
  p.flip = transposed;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.relu = relu;
  const long long m_tiles = (static_cast<long long>(N) * Ho * Wo + 127) / 128;
  const bool wide = Cout >= 256 && m_tiles * ((Cout + 255) / 256) >= num_sms;
  if (transposed) {
    if (wide) return rvn::conv_launch<256, true>(x, w, y, p, num_sms, stream);
    return rvn::conv_launch<128, true>(x, w, y, p, num_sms, stream);
  }
  if (stats != nullptr) {
    if (relu) return -7;               // the statistics are those of the convolution output itself
    if (wide) return rvn::conv_launch<256, false, true>(x, w, y, p, num_sms, stream);
    return rvn::conv_launch<128, false, true>(x, w, y, p, num_sms, stream);
  }
  if (wide) return rvn::conv_launch<256, false>(x, w, y, p, num_sms, stream);
  return rvn::conv_launch<128, false>(x, w, y, p, num_sms, stream);
}

// Data gradient of a STRIDE-2 convolution y = conv(x, w; stride 2, pad, pad_w), dilation 1:
//   dx[n, h, w, ci] = sum over (r, s, co) with (h + pad - r), (w + pad_w - s) even of dy[n, (h + pad - r) / 2, (w + pad_w - s) / 2, co] * w[co, r, s, ci]
// Split by the parity (a, b) of (h, w): every phase is a stride-1 convolution of dy with the filter taps of matching parity, i.e. one
// launch of the implicit-GEMM kernel (transposed / MN-major weights, as the stride-1 dgrad) whose output tensor map steps over every
// second pixel of dx.  No zero-insertion, no scatter: each dx element is written exactly once (phases without taps — 1x1 filters —
// stay zero: the caller zero-fills dx when R == 1 or S == 1).  dy [N,Ho,Wo,Cout], w [Cout,R,S,Cin], dx [N,H,W,Cin].
extern "C" int rvn_conv_dgrad_s2(const void* dy, const void* w, void* dx, int N, int Ho, int Wo, int Cout, int H, int W, int Cin, int R,
                                 int S, int pad, int pad_w, int num_sms, cudaStream_t stream) {
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      const int Ha = (H - a + 1) / 2, Wb = (W - b + 1) / 2;
      if (Ha <= 0 || Wb <= 0) continue;
      rvn::ConvParams p{};
      p.sub = 1;
      p.r0 = (a + pad) % 2;
      p.s0 = (b + pad_w) % 2;
      p.Ra = p.r0 < R ? (R - p.r0 + 1) / 2 : 0;
      p.Sb = p.s0 < S ? (S - p.s0 + 1) / 2 : 0;
      if (p.Ra == 0 || p.Sb == 0) continue;                 // no tap reaches this phase
      p.ch = (a + pad - p.r0) / 2;
      p.cw = (b + pad_w - p.s0) / 2;
      p.N = N; p.H = Ho; p.W = Wo; p.Cin = Cout;            // GEMM K = channels of dy
      p.Ho = Ha; p.Wo = Wb; p.Cout = Cin;                   // GEMM N = channels of dx
      p.R = R; p.S = S; p.stride = 1; p.pad = 0; p.pad_w = 0; p.dil = 1;
      p.ys_w = 2LL * Cin; p.ys_h = 2LL * W * Cin; p.ys_n = static_cast<long long>(H) * W * Cin;
      void* base = static_cast<__nv_bfloat16*>(dx) + (static_cast<long long>(a) * W + b) * Cin;
      const long long m_tiles = (static_cast<long long>(N) * Ha * Wb + 127) / 128;
      const bool wide = Cin >= 256 && m_tiles * ((Cin + 255) / 256) >= num_sms;
      const int rc = wide ? rvn::conv_launch<256, true>(dy, w, base, p, num_sms, stream) : rvn::conv_launch<128, true>(dy, w, base, p, num_sms, stream);
      if (rc != 0) return rc;
    }
  return 0;
}

extern "C" int rvn_conv_forward(const void* x, const void* w, void* y, const void* bias, int relu, int N, int H, int W, int Cin, int Ho,
                                int Wo, int Cout, int R, int S, int stride, int pad, int pad_w, int dil, int transposed, int num_sms,
                                cudaStream_t stream) {
  return rvn_conv_forward_strided(x, w, y, bias, relu, N, H, W, Cin, Ho, Wo, Cout, R, S, stride, pad, pad_w, dil, transposed, 0, 0, 0, nullptr,
                                  num_sms, stream);
}

// ================================================================================================ weight gradient
//   dW[co, (r,s), ci] = sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*stride + r*dil - pad, wo*stride + s*dil - pad, ci]
// GEMM with the PIXELS as the reduction dimension: A = dy patch (128 pixels x 128 co, MN-major), B = the tap-shifted x patch
// (128 pixels x BN ci, MN-major) — both are the 4-D TMA patch loads of the forward kernel (zero fill = padding, and
// zero-filled dy rows outside the image contribute nothing).  One CTA owns an output tile (co block, tap, ci block) and a
// slice of the pixel patches (split-K: early layers have only a handful of output tiles); partial sums are added into
// an fp32 [Cout, R*S, Cin] buffer with red.global.add.f32.
namespace rvn {

constexpr int WG_PIX = 128;                     // pixels per K block (power-of-two patch, zero-filled by TMA)
constexpr int WG_A_BYTES = 2 * WG_PIX * 128;    // 128 co = two 64-wide MN atoms columns
constexpr int WG_THREADS = 192;                 // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue

struct WgradParams {
  int N, H, W, Cin, Ho, Wo, Cout;
  int R, S, stride, pad, dil, pad_w;
  int TW, TH, TN, tiles_w, tiles_h, tiles_n;
  int co_blocks, ci_blocks, splits, patches_per_split;
  float* dw;                                    // [Cout, R*S, Cin] fp32, accumulated into (zero it first for a plain gradient)
  long long xs_w, xs_h, xs_n;                   // host only: element strides of x's W / H / N dimensions (0 = dense NHWC)
};

template <int BN>
struct WgradSmem {
  static constexpr int B_BYTES = (BN / 64) * WG_PIX * 128;
  static constexpr int STAGE_BYTES = WG_A_BYTES + B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 128 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
rvn_conv_wgrad_tcgen05(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                       const __grid_constant__ CUtensorMap map_dw, const WgradParams p) {
  using L = WgradSmem<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr uint32_t IDESC = make_idesc_bf16(128, BN, true, true);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  // tile decode: blockIdx.x = ((split * co_blocks + co_blk) * RS + tap) * ci_blocks + ci_blk
  int t = blockIdx.x;
  const int ci_blk = t % p.ci_blocks; t /= p.ci_blocks;
  const int tap = t % (p.R * p.S); t /= (p.R * p.S);
  const int co_blk = t % p.co_blocks;
  const int split = t / p.co_blocks;
  const int n_patches = p.tiles_w * p.tiles_h * p.tiles_n;
  const int patch0 = split * p.patches_per_split;
  const int patch1 = min(n_patches, patch0 + p.patches_per_split);
  const int r = tap / p.S, s = tap % p.S;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dw);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = patch0; pt < patch1; ++pt) {
        const int wo0 = (pt % p.tiles_w) * p.TW;
        const int ho0 = ((pt / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = (pt / (p.tiles_w * p.tiles_h)) * p.TN;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + WG_A_BYTES;
        mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_4d(sa + j * (WG_PIX * 128), &map_dy, &full_bar[stage], co_blk * 128 + j * 64, wo0, ho0, n0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(sb + j * (WG_PIX * 128), &map_x, &full_bar[stage], ci_blk * BN + j * 64, wo0 * p.stride + s * p.dil - p.pad_w,
                      ho0 * p.stride + r * p.dil - p.pad, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = patch0; pt < patch1; ++pt) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t b_addr = a_addr + WG_A_BYTES;
#pragma unroll
        for (int k = 0; k < WG_PIX / 16; ++k)
          umma_bf16(tmem_d, make_smem_desc_sw128(a_addr + k * 2048, WG_PIX * 128, 1024),
                    make_smem_desc_sw128(b_addr + k * 2048, WG_PIX * 128, 1024), IDESC, (pt != patch0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (pt == patch1 - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (patch1 > patch0) {
    // epilogue: thread == output-channel row.  The split-K partial tile goes into dW[co, tap, ci] as BULK reductions: every 128 x 32
    // fp32 chunk is staged (128B-swizzled) in the now idle operand stages and added at the L2 by one cp.reduce.async.bulk.tensor,
    // instead of 32 scalar red.global.add.f32 per thread and chunk (16 K atomics per CTA: the wgrad kernels spent most of their
    // time in the atomic unit — tensor pipe 22 %).  Cout / Cin tails are clipped by the tensor map.
    const int quad = warp & 3;                            // warps 2,3,4,5 -> TMEM lane quadrants 2,3,0,1
    const int row = quad * 32 + lane_id();
    const bool issuer = threadIdx.x == 64;                // first epilogue thread
    mbar_wait(done_bar, 0);                               // every MMA has retired: the smem stages are free
    tcgen05_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint8_t* buf = smem + (c & 1) * (128 * 128);        // two 16 KB staging tiles inside stage 0
      if (c >= 2) {
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the reduction that last read this tile is done
        named_bar_sync(1, 128);
      }
      uint32_t v[32];
      tmem_ld_32x32(tmem_d + (static_cast<uint32_t>(quad * 32) << 16) + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        uint4 w4;
        w4.x = v[j]; w4.y = v[j + 1]; w4.z = v[j + 2]; w4.w = v[j + 3];
        *reinterpret_cast<uint4*>(buf + row * 128 + (((j >> 2) ^ (row & 7)) << 4)) = w4;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {
        tma_reduce_add_3d(&map_dw, buf, ci_blk * BN + c * 32, tap, co_blk * 128);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, BN < 32 ? 32 : BN);
}

// power-of-two patch with TW*TH*TN == 128 (zero-filled beyond the image) maximising useful pixels
static void pick_patch_pow2(int Wo, int Ho, int N, int& TW, int& TH, int& TN) {
  double best = -1.0;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; tw * th <= 128; th *= 2) {
      const int tn = 128 / (tw * th);
      const double tiles = static_cast<double>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((N + tn - 1) / tn);
      const double eff = static_cast<double>(Wo) * Ho * N / (tiles * 128.0);
      if (eff > best + 1e-9) { best = eff; TW = tw; TH = th; TN = tn; }
    }
}

template <int BN>
static int wgrad_launch(const void* dy, const void* x, WgradParams p, int num_sms, cudaStream_t stream) {
  using L = WgradSmem<BN>;
  pick_patch_pow2(p.Wo, p.Ho, p.N, p.TW, p.TH, p.TN);
  p.tiles_w = (p.Wo + p.TW - 1) / p.TW;
  p.tiles_h = (p.Ho + p.TH - 1) / p.TH;
  p.tiles_n = (p.N + p.TN - 1) / p.TN;
  p.co_blocks = (p.Cout + 127) / 128;
  p.ci_blocks = (p.Cin + BN - 1) / BN;
  const int base_tiles = p.co_blocks * p.R * p.S * p.ci_blocks;
  const int n_patches = p.tiles_w * p.tiles_h * p.tiles_n;
  int splits = (2 * num_sms + base_tiles - 1) / base_tiles;
  if (splits > n_patches) splits = n_patches;
  if (splits < 1) splits = 1;
  p.patches_per_split = (n_patches + splits - 1) / splits;
  p.splits = (n_patches + p.patches_per_split - 1) / p.patches_per_split;
  CUtensorMap mdy, mx;
  int e = 0;
  {
    const uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.N};
    const uint64_t st[3] = {(uint64_t)p.Cout, (uint64_t)p.Wo * p.Cout, (uint64_t)p.Ho * p.Wo * p.Cout};
    const uint32_t box[4] = {64u, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    const uint32_t es[4] = {1u, 1u, 1u, 1u};
    e |= make_map_nd(&mdy, dy, 4, dims, st, box, es);
  }
  {
    const uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.N};
    uint64_t st[3] = {(uint64_t)p.Cin, (uint64_t)p.W * p.Cin, (uint64_t)p.H * p.W * p.Cin};
    if (p.xs_w > 0) { st[0] = (uint64_t)p.xs_w; st[1] = (uint64_t)p.xs_h; st[2] = (uint64_t)p.xs_n; }
    const uint32_t box[4] = {64u, (uint32_t)((p.TW - 1) * p.stride + 1), (uint32_t)((p.TH - 1) * p.stride + 1), (uint32_t)p.TN};
    const uint32_t es[4] = {1u, (uint32_t)p.stride, (uint32_t)p.stride, 1u};
    e |= make_map_nd(&mx, x, 4, dims, st, box, es);
  }
  CUtensorMap mdw;
  {
    const uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)(p.R * p.S), (uint64_t)p.Cout};
    const uint64_t st[2] = {(uint64_t)p.Cin, (uint64_t)p.R * p.S * p.Cin};
    const uint32_t box[3] = {32u, 1u, 128u};
    e |= make_map_nd_f32(&mdw, p.dw, 3, dims, st, box);
  }
  if (e != 0) return 1000 + e;
  auto kern = rvn_conv_wgrad_tcgen05<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return 2000;
    configured = true;
  }
  kern<<<base_tiles * p.splits, WG_THREADS, L::TOTAL, stream>>>(mdy, mx, mdw, p);
  return cudaGetLastError() == cudaSuccess ? 0 : 3000;
}

}  // namespace rvn

// dy [N,Ho,Wo,Cout], x [N,H,W,Cin] bf16 NHWC (x's W / H / N element strides may be given, 0 = dense); dw [Cout,R,S,Cin] fp32 is
// ACCUMULATED into: a zeroed buffer gives the plain gradient, a persistent fp32 accumulator collects several passes
extern "C" int rvn_conv_wgrad_strided(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Ho, int Wo, int Cout,
                                      int R, int S, int stride, int pad, int pad_w, int dil, long long xs_w, long long xs_h,
                                      long long xs_n, int num_sms, cudaStream_t stream) {
  rvn::WgradParams p{};
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.R = R; p.S = S; p.stride = stride; p.pad = pad; p.pad_w = pad_w; p.dil = dil;
  p.dw = dw;
  p.xs_w = xs_w; p.xs_h = xs_h; p.xs_n = xs_n;
  if (Cin > 64) return rvn::wgrad_launch<128>(dy, x, p, num_sms, stream);
  return rvn::wgrad_launch<64>(dy, x, p, num_sms, stream);
}

extern "C" int rvn_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int R,
                              int S, int stride, int pad, int pad_w, int dil, int num_sms, cudaStream_t stream) {
  return rvn_conv_wgrad_strided(dy, x, dw, N, H, W, Cin, Ho, Wo, Cout, R, S, stride, pad, pad_w, dil, 0, 0, 0, num_sms, stream);
}
