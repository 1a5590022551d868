// K7: training-mode BatchNorm on NHWC bf16 activations, fused with the ReLU and the residual add that follow it in
// ResNet / Inception blocks:      y = relu( (x - mean) * invstd * gamma + beta  (+ residual) )
//
// Measured motivation (profiles/resnet50_step_baseline_r1.txt): in the stock bf16 channels_last ResNet-50 step ATen's four
// BatchNorm kernels take 12.4 ms of 24 ms while all convolutions together take ~3.5 ms.  These kernels are pure streams:
// the activation is viewed as a flat array of 16-byte vectors and the launch is shaped so that every thread always lands on
// the SAME 8-channel group (total thread count is a multiple of C/8) — per-channel parameters and partial sums live in
// registers, loads are perfectly coalesced, no index arithmetic per element.
//   forward : rvn_bn_stats (sum, sum of squares) -> rvn_bn_finalize (mean, invstd, running stats) -> rvn_bn_apply
//   backward: rvn_bn_bwd_reduce (sum dz, sum dz*xhat with dz = dy * relu_mask) -> rvn_bn_bwd_apply (dx, dz for the skip branch)
// Reference call sites: nn.BatchNorm2d in torchvision resnet50 / models.py BasicConv2d (models.py:383-392), F.relu, `out += identity`.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvn {

struct v8 {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void to_float(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

// per-channel parameters / statistics arrive in the module's own dtype (bf16 under a bf16 Node, fp32 otherwise): no cast launches
__device__ __forceinline__ float ldp(const void* p, int i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void stp(void* p, int i, float v, int is_bf16) {
  if (is_bf16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// sums[0..C) += sum_r x[r,c];  sums[C..2C) += sum_r x[r,c]^2      (total threads % (C/8) == 0)
__global__ void __launch_bounds__(256) rvn_bn_stats(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums, size_t nvec, int C) {
  extern __shared__ float sh[];                     // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int c0 = static_cast<int>(tid % (C / 8)) * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  size_t i = tid;
  for (; i + 3 * stride < nvec; i += 4 * stride) {  // four independent loads in flight
    v8 a, b, c, d;
    float fa[8], fb[8], fc[8], fd[8];
    a.load(x + i * 8);
    b.load(x + (i + stride) * 8);
    c.load(x + (i + 2 * stride) * 8);
    d.load(x + (i + 3 * stride) * 8);
    a.to_float(fa);
    b.to_float(fb);
    c.to_float(fc);
    d.to_float(fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += (fa[j] + fb[j]) + (fc[j] + fd[j]);
      q[j] += (fa[j] * fa[j] + fb[j] * fb[j]) + (fc[j] * fc[j] + fd[j] * fd[j]);
    }
  }
  for (; i + stride < nvec; i += 2 * stride) {
    v8 a, b;
    float fa[8], fb[8];
    a.load(x + i * 8);
    b.load(x + (i + stride) * 8);
    a.to_float(fa);
    b.to_float(fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += fa[j] + fb[j]; q[j] += fa[j] * fa[j] + fb[j] * fb[j]; }
  }
  if (i < nvec) {
    v8 a;
    float fa[8];
    a.load(x + i * 8);
    a.to_float(fa);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += fa[j]; q[j] += fa[j] * fa[j]; }
  }
  // narrow layers (C = 64 / 128: the LARGEST activations) put 32 / (C/8) lanes of a warp on the same channel group: fold those with
  // shuffles first, otherwise 32 - 256 threads of the CTA hammer the same shared-memory word with (CAS-emulated) float atomics
  const int groups = C / 8;
  bool owner = true;
  if (groups < 32 && (32 % groups) == 0) {
    for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
      }
    }
    owner = (threadIdx.x & 31) < groups;
  }
  if (owner) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[c0 + j], s[j]);
      atomicAdd(&sh[C + c0 + j], q[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) atomicAdd(&sums[c], sh[c]);
}

// mean / invstd from the sums; running statistics updated like torch (unbiased variance, momentum)
__global__ void rvn_bn_finalize(const float* __restrict__ sums, float* __restrict__ mean, float* __restrict__ invstd,
                                void* __restrict__ running_mean, void* __restrict__ running_var, long long* __restrict__ batches_tracked,
                                float count, float eps, float momentum, int C, int stats_bf16) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && batches_tracked != nullptr) *batches_tracked += 1;
  if (c >= C) return;
  const float m = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - m * m, 0.f);
  mean[c] = m;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    stp(running_mean, c, (1.f - momentum) * ldp(running_mean, c, stats_bf16) + momentum * m, stats_bf16);
    stp(running_var, c, (1.f - momentum) * ldp(running_var, c, stats_bf16) + momentum * var * (count / fmaxf(count - 1.f, 1.f)), stats_bf16);
  }
}

// y = act((x - mean) * invstd * gamma + beta + residual)
__global__ void __launch_bounds__(256, 5) rvn_bn_apply(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                    const void* __restrict__ gamma, const void* __restrict__ beta,
                                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                                    __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ relu_mask, size_t nvec, int C,
                                                    int relu, int param_bf16) {
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int c0 = static_cast<int>(tid % (C / 8)) * 8;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = ldp(gamma, c0 + j, param_bf16) * invstd[c0 + j];
    sf[j] = ldp(beta, c0 + j, param_bf16) - mean[c0 + j] * sc[j];
  }
  auto finish = [&](size_t i, const v8& a, const v8& r) {
    v8 o;
    float f[8];
    a.to_float(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sf[j]);
    if (res != nullptr) {
      float rf[8];
      r.to_float(rf);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += rf[j];
    }
    if (relu) {
      if (relu_mask != nullptr) {                   // one bit per element: what backward needs of y (1/16 of its bytes)
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (f[j] > 0.f ? 1u : 0u) << j;
        relu_mask[i] = static_cast<uint8_t>(m);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    o.from_float(f);
    o.store(y + i * 8);
  };
  size_t i = tid;
  for (; i + stride < nvec; i += 2 * stride) {      // all loads of two vectors in flight before either is consumed (pure stream:
    v8 a0, a1, r0, r1;                              // bytes in flight per SM decide the bandwidth)
    a0.load(x + i * 8);
    a1.load(x + (i + stride) * 8);
    if (res != nullptr) {
      r0.load(res + i * 8);
      r1.load(res + (i + stride) * 8);
    }
    finish(i, a0, r0);
    finish(i + stride, a1, r1);
  }
  if (i < nvec) {
    v8 a0, r0;
    a0.load(x + i * 8);
    if (res != nullptr) r0.load(res + i * 8);
    finish(i, a0, r0);
  }
}

// sums[0..C) += sum dz ; sums[C..2C) += sum dz * xhat      with dz = dy * (y > 0) when relu
__global__ void __launch_bounds__(256, 4) rvn_bn_bwd_reduce(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                         const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd, float* __restrict__ sums, size_t nvec, int C,
                                                         int relu, const uint8_t* __restrict__ relu_mask) {
  extern __shared__ float sh[];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int c0 = static_cast<int>(tid % (C / 8)) * 8;
  float mu[8], is[8], s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j]; s[j] = 0.f; q[j] = 0.f; }
  auto accumulate_vec = [&](const v8& a, const v8& b, const v8& c, unsigned m) {
    float g[8], xv[8];
    a.to_float(g);
    b.to_float(xv);
    if (relu) {
      if (relu_mask != nullptr) {                   // the forward's bitmask: 1 byte instead of the 16-byte y vector
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = ((m >> j) & 1u) ? g[j] : 0.f;
      } else {
        float yv[8];
        c.to_float(yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += g[j];
      q[j] += g[j] * (xv[j] - mu[j]) * is[j];
    }
  };
  const bool use_y = relu && relu_mask == nullptr;
  const bool use_m = relu && relu_mask != nullptr;
  size_t i = tid;
  for (; i + stride < nvec; i += 2 * stride) {      // two vectors: 4 - 6 independent 16-byte loads in flight per thread
    v8 a0, b0, c0v, a1, b1, c1v;
    unsigned m0 = 0, m1 = 0;
    a0.load(dy + i * 8);
    b0.load(x + i * 8);
    a1.load(dy + (i + stride) * 8);
    b1.load(x + (i + stride) * 8);
    if (use_y) { c0v.load(y + i * 8); c1v.load(y + (i + stride) * 8); }
    if (use_m) { m0 = relu_mask[i]; m1 = relu_mask[i + stride]; }
    accumulate_vec(a0, b0, c0v, m0);
    accumulate_vec(a1, b1, c1v, m1);
  }
  if (i < nvec) {
    v8 a0, b0, c0v;
    unsigned m0 = 0;
    a0.load(dy + i * 8);
    b0.load(x + i * 8);
    if (use_y) c0v.load(y + i * 8);
    if (use_m) m0 = relu_mask[i];
    accumulate_vec(a0, b0, c0v, m0);
  }
  // narrow layers (C = 64 / 128: the LARGEST activations) put 32 / (C/8) lanes of a warp on the same channel group: fold those with
  // shuffles first, otherwise 32 - 256 threads of the CTA hammer the same shared-memory word with (CAS-emulated) float atomics
  const int groups = C / 8;
  bool owner = true;
  if (groups < 32 && (32 % groups) == 0) {
    for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
      }
    }
    owner = (threadIdx.x & 31) < groups;
  }
  if (owner) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[c0 + j], s[j]);
      atomicAdd(&sh[C + c0 + j], q[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) atomicAdd(&sums[c], sh[c]);
}

// dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat));  dres = dz (gradient of the fused residual branch)
__global__ void __launch_bounds__(256, 3) rvn_bn_bwd_apply(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                        const __nv_bfloat16* __restrict__ x, const void* __restrict__ gamma,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        const float* __restrict__ sums, __nv_bfloat16* __restrict__ dx,
                                                        __nv_bfloat16* __restrict__ dres, void* __restrict__ dgamma,
                                                        void* __restrict__ dbeta, size_t nvec, int C, float count, int relu,
                                                        int param_bf16, int accumulate, const uint8_t* __restrict__ relu_mask) {
  if (blockIdx.x == 0 && dgamma != nullptr) {       // parameter gradients in the parameter dtype, optionally accumulated in place
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float dg = sums[C + c] + (accumulate ? ldp(dgamma, c, param_bf16) : 0.f);
      const float db = sums[c] + (accumulate ? ldp(dbeta, c, param_bf16) : 0.f);
      stp(dgamma, c, dg, param_bf16);
      stp(dbeta, c, db, param_bf16);
    }
  }
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int c0 = static_cast<int>(tid % (C / 8)) * 8;
  float mu[8], is[8], gi[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[c0 + j];
    is[j] = invstd[c0 + j];
    gi[j] = ldp(gamma, c0 + j, param_bf16) * is[j];
    m1[j] = sums[c0 + j] / count;
    m2[j] = sums[C + c0 + j] / count;
  }
  auto finish = [&](size_t i, const v8& a, const v8& b, const v8& c, unsigned m) {
    v8 o;
    float g[8], xv[8];
    a.to_float(g);
    b.to_float(xv);
    if (relu) {
      if (relu_mask != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = ((m >> j) & 1u) ? g[j] : 0.f;
      } else {
        float yv[8];
        c.to_float(yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
      }
    }
    if (dres != nullptr) {
      o.from_float(g);
      o.store(dres + i * 8);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = gi[j] * (g[j] - m1[j] - (xv[j] - mu[j]) * is[j] * m2[j]);
    o.from_float(g);
    o.store(dx + i * 8);
  };
  const bool use_y = relu && relu_mask == nullptr;
  const bool use_m = relu && relu_mask != nullptr;
  size_t i = tid;
  for (; i + stride < nvec; i += 2 * stride) {
    v8 a0, b0, c0v, a1, b1, c1v;
    unsigned k0 = 0, k1 = 0;
    a0.load(dy + i * 8);
    b0.load(x + i * 8);
    a1.load(dy + (i + stride) * 8);
    b1.load(x + (i + stride) * 8);
    if (use_y) { c0v.load(y + i * 8); c1v.load(y + (i + stride) * 8); }
    if (use_m) { k0 = relu_mask[i]; k1 = relu_mask[i + stride]; }
    finish(i, a0, b0, c0v, k0);
    finish(i + stride, a1, b1, c1v, k1);
  }
  if (i < nvec) {
    v8 a0, b0, c0v;
    unsigned k0 = 0;
    a0.load(dy + i * 8);
    b0.load(x + i * 8);
    if (use_y) c0v.load(y + i * 8);
    if (use_m) k0 = relu_mask[i];
    finish(i, a0, b0, c0v, k0);
  }
}

}  // namespace rvn

// grid such that (grid * 256) % (C/8) == 0 and there are ~`waves` x num_sms CTAs
static int bn_grid(size_t nvec, int C, int num_sms, int waves) {
  const int groups = C / 8;
  int g = 1;                                      // smallest grid multiple that keeps threads pinned to one channel group
  while ((static_cast<long long>(g) * 256) % groups != 0) ++g;
  long long want = static_cast<long long>(num_sms) * waves;
  const long long max_useful = static_cast<long long>((nvec + 255) / 256);
  if (want > max_useful) want = max_useful;
  long long grid = (want / g) * g;
  if (grid < g) grid = g;
  return static_cast<int>(grid);
}

extern "C" int rvn_batchnorm_forward(const void* x, const void* res, const void* gamma, const void* beta, void* running_mean,
                                     void* running_var, long long* batches_tracked, float* mean, float* invstd, float* sums, void* y,
                                     long long M, int C, float eps, float momentum, int relu, int training, int param_bf16,
                                     int stats_bf16, void* relu_mask, int sums_ready, int num_sms, cudaStream_t stream) {
  if (C % 8 != 0 || C > 4096) return -2;
  const size_t nvec = static_cast<size_t>(M) * C / 8;
  using T = const __nv_bfloat16*;
  if (training) {
    if (!sums_ready) {       // (sums_ready: the producing GEMM / convolution epilogue already collected sum and sum of squares)
      cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream);
      rvn::rvn_bn_stats<<<bn_grid(nvec, C, num_sms, 4), 256, 2 * C * sizeof(float), stream>>>((T)x, sums, nvec, C);
    }
    rvn::rvn_bn_finalize<<<(C + 127) / 128, 128, 0, stream>>>(sums, mean, invstd, running_mean, running_var, batches_tracked,
                                                               static_cast<float>(M), eps, momentum, C, stats_bf16);
  }
  rvn::rvn_bn_apply<<<bn_grid(nvec, C, num_sms, 8), 256, 0, stream>>>((T)x, (T)res, gamma, beta, mean, invstd, (__nv_bfloat16*)y,
                                                                      static_cast<uint8_t*>(relu_mask), nvec, C, relu, param_bf16);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// sums (2*C fp32) holds {dbeta = sum dz, dgamma = sum dz*xhat} on return
extern "C" int rvn_batchnorm_backward(const void* dy, const void* y, const void* x, const void* gamma, const float* mean,
                                      const float* invstd, float* sums, void* dx, void* dres, void* dgamma, void* dbeta, long long M,
                                      int C, int relu, int param_bf16, int accumulate, const void* relu_mask, int num_sms,
                                      cudaStream_t stream) {
  if (C % 8 != 0 || C > 4096) return -2;
  const size_t nvec = static_cast<size_t>(M) * C / 8;
  using T = const __nv_bfloat16*;
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream);
  rvn::rvn_bn_bwd_reduce<<<bn_grid(nvec, C, num_sms, 4), 256, 2 * C * sizeof(float), stream>>>((T)dy, (T)y, (T)x, mean, invstd, sums, nvec, C,
                                                                                            relu, static_cast<const uint8_t*>(relu_mask));
  rvn::rvn_bn_bwd_apply<<<bn_grid(nvec, C, num_sms, 8), 256, 0, stream>>>((T)dy, (T)y, (T)x, gamma, mean, invstd, sums, (__nv_bfloat16*)dx,
                                                                          (__nv_bfloat16*)dres, dgamma, dbeta, nvec, C, static_cast<float>(M),
                                                                          relu, param_bf16, accumulate, static_cast<const uint8_t*>(relu_mask));
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
