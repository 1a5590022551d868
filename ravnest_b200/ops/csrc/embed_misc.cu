// K10 / K11 / K12: the small memory-bound ops around the GEMMs that the reference leaves to ATen.
//
//  K10 embedding   forward : out[i, :] = W[idx[i], :]                       (one warp per row, 16 B vectors)
//                  backward: dW[r, :] += sum_{i: idx[i] == r} dy[i, :]      deterministic, atomic-free, sort-free:
//                            CTA i decides whether i is the FIRST occurrence of its index (an O(n) scan of the index list that
//                            lives in L2), the owner gathers every matching row in fp32 registers and adds the sum into the
//                            gradient row once.  ATen's path is radix sort + segment offsets + two reduction kernels
//                            (~0.25 ms per BERT step, profiles/bert_step_kernels_r1a.txt); n <= 16384 indices here.
//  K11 dropout     y = x * mask / (1 - p) with the mask regenerated from Philox(seed, step, position) in backward — no mask
//                  tensor, bit-identical between forward, recompute and backward (reference: saved CPU+CUDA RNG state per
//                  forward id, compute.py:63-68).
//  K12 pack        activation "compression" before a send: fp32 -> fp16 with clamp to +-65504, or fp32 -> bf16, one pass
//                  (reference utils.py:184-194 does clamp and cast as separate ops on the CPU).
// Reference call sites: nn.Embedding in minGPT (model_without_padding_mask.py:193-194) and BERT, nn.Dropout in models.py:9,:15,
// :20,:188 and GPT, compress_tensor before trigger_send (communication.py:86-95,:110-111).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace rvn {

// ------------------------------------------------------------------------------------------------ K10 forward
template <typename T>
__global__ void __launch_bounds__(256) rvn_embedding_fwd(const int64_t* __restrict__ idx, const T* __restrict__ w, T* __restrict__ out,
                                                         long long n, int H, long long V) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  long long r = idx[row];
  if (r < 0 || r >= V) r = 0;                       // out-of-range indices are a caller bug; never read outside the table
  constexpr int VEC = 16 / sizeof(T);
  const T* src = w + r * H;
  T* dst = out + row * H;
  if (H % VEC == 0) {
    for (int c = lane * VEC; c < H; c += 32 * VEC) *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<const uint4*>(src + c);
  } else {
    for (int c = lane; c < H; c += 32) dst[c] = src[c];
  }
}

// ------------------------------------------------------------------------------------------------ K10 backward
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float v) { d = __float2bfloat16_rn(v); }

// grid = n CTAs of 256 threads; dynamic smem = n * 4 bytes (match list, ascending positions -> fixed summation order).
// dW rows are read-modify-written by exactly one CTA (the owner of the index), so the caller may pass the live gradient bucket
// (accumulation) or a zero-filled tensor.
template <typename T>
__global__ void __launch_bounds__(256) rvn_embedding_bwd(const int64_t* __restrict__ idx, const T* __restrict__ dy, T* __restrict__ dw,
                                                         int n, int H, long long V, long long padding_idx) {
  extern __shared__ int matches[];
  __shared__ float red[8][256];
  __shared__ int warp_count[8];
  __shared__ int not_first;
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long r = idx[i];
  if (r < 0 || r >= V || r == padding_idx) return;          // uniform across the CTA
  if (threadIdx.x == 0) not_first = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < i; j += blockDim.x)
    if (idx[j] == r) not_first = 1;                           // benign race: every writer stores 1
  __syncthreads();
  if (not_first) return;
  // owner: ordered list of every occurrence at or after i (ballot prefix per 256-candidate chunk)
  int m = 0;
  for (int base = i; base < n; base += 256) {
    const int j = base + threadIdx.x;
    const bool hit = j < n && idx[j] == r;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_count[w] = __popc(ballot);
    __syncthreads();
    int before = m;
    for (int q = 0; q < w; ++q) before += warp_count[q];
    if (hit) matches[before + __popc(ballot & ((1u << lane) - 1))] = j;
    int total = 0;
    for (int q = 0; q < 8; ++q) total += warp_count[q];
    m += total;
    __syncthreads();
  }
  // warp w sums matches w, w + 8, ... ; lanes cover 256 columns per pass (8 per lane, 32-element coalesced segments)
  for (int chunk = 0; chunk < H; chunk += 256) {
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int k = w; k < m; k += 8) {
      const T* row = dy + static_cast<size_t>(matches[k]) * H + chunk;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = q * 32 + lane;
        if (chunk + c < H) acc[q] += to_f(row[c]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) red[w][q * 32 + lane] = acc[q];
    __syncthreads();
    const int c = chunk + threadIdx.x;
    if (c < H) {
      float total = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) total += red[q][threadIdx.x];
      T* dst = dw + static_cast<size_t>(r) * H + c;
      from_f(*dst, to_f(*dst) + total);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ K11 dropout
// y = x * keep_scale(position); the same kernel is the backward (dx = dy * keep_scale).  8 bf16 / 4 fp32 per thread.
__global__ void __launch_bounds__(256) rvn_dropout_bf16(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n8,
                                                        const long long* __restrict__ rng, float p) {
  DropoutRng d;
  d.init(rng, p);
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n8; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(x + g * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float s[8];
    d.scales8(g, s);
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      oh[k] = __floats2bfloat162_rn(f.x * s[2 * k], f.y * s[2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(y + g * 8) = o;
  }
}
__global__ void __launch_bounds__(256) rvn_dropout_f32(const float* __restrict__ x, float* __restrict__ y, size_t n4,
                                                       const long long* __restrict__ rng, float p) {
  DropoutRng d;
  d.init(rng, p);
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(x + g * 4);
    float s[4];
    d.scales4(g, s);
    v.x *= s[0]; v.y *= s[1]; v.z *= s[2]; v.w *= s[3];
    *reinterpret_cast<float4*>(y + g * 4) = v;
  }
}

// snapshot = state; state.step += 1 — the per-consumer RNG bookkeeping as ONE single-thread launch (graph-capturable)
__global__ void rvn_rng_next_kernel(long long* __restrict__ state, long long* __restrict__ snap) {
  snap[0] = state[0];
  snap[1] = state[1];
  state[1] += 1;
}

// ------------------------------------------------------------------------------------------------ K12 pack
// mode 0: fp32 -> fp16 with clamp to the fp16 range (the reference's wire format), mode 1: fp32 -> bf16
__global__ void __launch_bounds__(256) rvn_pack_f32(const float* __restrict__ x, uint16_t* __restrict__ y, size_t n4, size_t n, int mode) {
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n4; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float f[4];
    if (g * 4 + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + g * 4);
      f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) f[k] = g * 4 + k < n ? x[g * 4 + k] : 0.f;
    }
    uint16_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (mode == 0) {
        const __half hv = __float2half_rn(fminf(fmaxf(f[k], -65504.f), 65504.f));
        o[k] = *reinterpret_cast<const uint16_t*>(&hv);
      } else {
        const __nv_bfloat16 bv = __float2bfloat16_rn(f[k]);
        o[k] = *reinterpret_cast<const uint16_t*>(&bv);
      }
    }
    if (g * 4 + 4 <= n) {
      *reinterpret_cast<uint2*>(y + g * 4) = make_uint2(o[0] | (static_cast<uint32_t>(o[1]) << 16), o[2] | (static_cast<uint32_t>(o[3]) << 16));
    } else {
      for (int k = 0; k < 4 && g * 4 + k < n; ++k) y[g * 4 + k] = o[k];
    }
  }
}

}  // namespace rvn

static int stream_grid(size_t items, int num_sms) {
  size_t blocks = (items + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks ? blocks : 1);
}

extern "C" int rvn_embedding_forward(const int64_t* idx, const void* w, void* out, long long n, int H, long long V, int is_bf16,
                                     cudaStream_t stream) {
  if (n == 0) return 0;
  const int grid = static_cast<int>((n + 7) / 8);
  if (is_bf16)
    rvn::rvn_embedding_fwd<__nv_bfloat16><<<grid, 256, 0, stream>>>(idx, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, n, H, V);
  else
    rvn::rvn_embedding_fwd<float><<<grid, 256, 0, stream>>>(idx, (const float*)w, (float*)out, n, H, V);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_embedding_backward(const int64_t* idx, const void* dy, void* dw, int n, int H, long long V, long long padding_idx,
                                      int is_bf16, cudaStream_t stream) {
  if (n == 0) return 0;
  if (n > 16384) return -2;
  const size_t smem = static_cast<size_t>(n) * sizeof(int);
  if (is_bf16) {
    auto k = rvn::rvn_embedding_bwd<__nv_bfloat16>;
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4); configured = true; }
    k<<<n, 256, smem, stream>>>(idx, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dw, n, H, V, padding_idx);
  } else {
    auto k = rvn::rvn_embedding_bwd<float>;
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4); configured = true; }
    k<<<n, 256, smem, stream>>>(idx, (const float*)dy, (float*)dw, n, H, V, padding_idx);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_dropout(const void* x, void* y, size_t n, const long long* rng, float p, int is_bf16, int num_sms, cudaStream_t stream) {
  if (n == 0) return 0;
  if (is_bf16) {
    if (n % 8 != 0) return -2;
    rvn::rvn_dropout_bf16<<<stream_grid(n / 8, num_sms), 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8, rng, p);
  } else {
    if (n % 4 != 0) return -2;
    rvn::rvn_dropout_f32<<<stream_grid(n / 4, num_sms), 256, 0, stream>>>((const float*)x, (float*)y, n / 4, rng, p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_rng_next(long long* state, long long* snap, cudaStream_t stream) {
  rvn::rvn_rng_next_kernel<<<1, 1, 0, stream>>>(state, snap);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_pack(const float* x, void* y, size_t n, int mode, int num_sms, cudaStream_t stream) {
  if (n == 0) return 0;
  const size_t n4 = (n + 3) / 4;
  rvn::rvn_pack_f32<<<stream_grid(n4, num_sms), 256, 0, stream>>>(x, (uint16_t*)y, n4, n, mode);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
