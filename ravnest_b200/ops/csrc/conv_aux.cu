// K4 helpers (streaming kernels around the implicit-GEMM convolution):
//
//  * rvn_stem_pack: the first convolution of every vision model reads 1 - 4 channel images (ResNet / Inception: RGB, CNN_Net: one
//    channel), whose 2 - 8 byte pixel pitch TMA cannot address (16-byte granularity).  One pass turns the image batch — any layout,
//    bf16 or fp32 — into a zero-bordered [N, H + 2*pad, W + 2*pad (+ slack), 8] bf16 tensor.  The convolution then reads EIGHT
//    neighbouring pixels (8 x 8 channels = one 64-wide K block = one 128-byte swizzle row) per tap row through a tensor map whose
//    W stride (16 B) is smaller than its innermost extent (128 B): an overlapping-stride im2col done by the TMA unit, no patch
//    matrix in memory (ops/native.py: _ConvSmallCin; reference call sites: conv1 of torchvision ResNet-50, models.py Inception /
//    CNN_Net stems).
//  * rvn_wgrad_flush: adds the stage's fp32 convolution-weight-gradient accumulator into the (bf16 or fp32) flat gradient bucket
//    and leaves the accumulator zero — once per optimizer step instead of a fill + cast-add pair per convolution and micro-step.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvn {

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// thread == one output pixel (16-byte store); input strides in elements, so NCHW, channels_last and NHWC sources all work
template <typename T>
__global__ void __launch_bounds__(256)
rvn_stem_pack(const T* __restrict__ x, uint4* __restrict__ out, int N, int C, int H, int W, int Hp, int Wp, int pad_h, int pad_w, long long sn,
              long long sc, long long sh, long long sw) {
  const size_t total = static_cast<size_t>(N) * Hp * Wp;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(g % Wp);
    const int hp = static_cast<int>((g / Wp) % Hp);
    const long long n = static_cast<long long>(g / (static_cast<size_t>(Wp) * Hp));
    const int h = hp - pad_h, w = wp - pad_w;
    float f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const T* src = x + n * sn + h * sh + w * sw;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) f[c] = ld_as_float(src + c * sc);
    }
    uint4 o;
    o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
    out[g] = o;
  }
}

// 8 elements per thread and iteration; untouched (all-zero) accumulator groups cost one 32-byte read and nothing else
template <typename G>
__global__ void __launch_bounds__(256) rvn_wgrad_flush(float* __restrict__ acc, G* __restrict__ grad, size_t n8) {
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < n8; g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4* a = reinterpret_cast<float4*>(acc) + 2 * g;
    const float4 a0 = a[0], a1 = a[1];
    if (a0.x == 0.f && a0.y == 0.f && a0.z == 0.f && a0.w == 0.f && a1.x == 0.f && a1.y == 0.f && a1.z == 0.f && a1.w == 0.f) continue;
    if constexpr (sizeof(G) == 2) {
      uint4* gp = reinterpret_cast<uint4*>(grad) + g;
      const uint4 gv = *gp;
      const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
      const float2 p0 = __bfloat1622float2(g2[0]), p1 = __bfloat1622float2(g2[1]), p2 = __bfloat1622float2(g2[2]),
                   p3 = __bfloat1622float2(g2[3]);
      uint4 o;
      o.x = pack2(p0.x + a0.x, p0.y + a0.y); o.y = pack2(p1.x + a0.z, p1.y + a0.w);
      o.z = pack2(p2.x + a1.x, p2.y + a1.y); o.w = pack2(p3.x + a1.z, p3.y + a1.w);
      *gp = o;
    } else {
      float4* gp = reinterpret_cast<float4*>(grad) + 2 * g;
      float4 g0 = gp[0], g1 = gp[1];
      g0.x += a0.x; g0.y += a0.y; g0.z += a0.z; g0.w += a0.w;
      g1.x += a1.x; g1.y += a1.y; g1.z += a1.z; g1.w += a1.w;
      gp[0] = g0; gp[1] = g1;
    }
    a[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    a[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

static int aux_grid(size_t items, int num_sms) {
  size_t blocks = (items + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks ? blocks : 1);
}

}  // namespace rvn

// x: element strides (sn, sc, sh, sw) of an [N, C, H, W]-shaped source, C <= 8; out: dense [N, Hp, Wp, 8] bf16
extern "C" int rvn_stem_pack(const void* x, int x_is_bf16, void* out, int N, int C, int H, int W, int Hp, int Wp, int pad_h, int pad_w,
                             long long sn, long long sc, long long sh, long long sw, int num_sms, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(N) * Hp * Wp;
  if (total == 0) return 0;
  if (C < 1 || C > 8) return 1;
  const int grid = rvn::aux_grid(total, num_sms);
  if (x_is_bf16)
    rvn::rvn_stem_pack<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint4*>(out), N, C, H, W, Hp,
                                                               Wp, pad_h, pad_w, sn, sc, sh, sw);
  else
    rvn::rvn_stem_pack<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<uint4*>(out), N, C, H, W, Hp, Wp, pad_h, pad_w,
                                                       sn, sc, sh, sw);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// acc: fp32 [n]; grad: bf16 or fp32 [n]; n % 8 == 0, both 32-byte aligned
extern "C" int rvn_wgrad_flush(float* acc, void* grad, int grad_is_bf16, size_t n, int num_sms, cudaStream_t stream) {
  if (n == 0) return 0;
  if (n % 8 != 0) return 1;
  const size_t n8 = n / 8;
  const int grid = rvn::aux_grid(n8, num_sms);
  if (grad_is_bf16)
    rvn::rvn_wgrad_flush<__nv_bfloat16><<<grid, 256, 0, stream>>>(acc, static_cast<__nv_bfloat16*>(grad), n8);
  else
    rvn::rvn_wgrad_flush<float><<<grid, 256, 0, stream>>>(acc, static_cast<float*>(grad), n8);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
