// Row-wise memory-bound kernels: LayerNorm (+fused residual) fwd/bwd (K5), softmax cross-entropy fwd/bwd with
// ignore_index (K9), column sum for bias gradients.  bf16 I/O, fp32 math, 16-byte vector accesses.
// Reference call sites: nn.LayerNorm in BERT / minGPT (mingpt/model_without_padding_mask.py:119-121,:197),
// CrossEntropyLoss in examples/{resnet50,sorter,bert}/provider.py.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace rvn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct bf16x8 {
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

// ------------------------------------------------------------------------------------------------ LayerNorm
// y = LN(dropout(x) + residual).  One warp per row; each lane owns VPL 8-element vectors.  The pre-norm sum s is
// written (bf16) whenever it differs from x, because backward re-reads exactly what was normalised.
template <int VPL>
__global__ void __launch_bounds__(256) rvn_layernorm_fwd(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                                                         const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                                         __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ sum_out,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
                                                         float eps, const long long* __restrict__ rng, float dropout_p) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  const bool use_drop = rng != nullptr && dropout_p > 0.f;
  DropoutRng drng;
  if (use_drop) drng.init(rng, dropout_p);
  float v[VPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
      bf16x8 a;
      a.load(x + static_cast<size_t>(row) * H + vec * 8);
      a.to_float(v[i]);
      if (use_drop) {
        const unsigned long long g = (static_cast<unsigned long long>(row) * H + vec * 8) >> 2;
        float sc[4];
        drng.scales4(g, sc);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] *= sc[j];
        drng.scales4(g + 1, sc);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][4 + j] *= sc[j];
      }
      if (res != nullptr) {
        bf16x8 r;
        float rf[8];
        r.load(res + static_cast<size_t>(row) * H + vec * 8);
        r.to_float(rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] += rf[j];
      }
      if (sum_out != nullptr) {
        bf16x8 o;
        o.from_float(v[i]);
        o.store(sum_out + static_cast<size_t>(row) * H + vec * 8);
        o.to_float(v[i]);            // normalise exactly what backward will re-read
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
      bf16x8 g, b, o;
      float gf[8], bfv[8], of[8];
      g.load(gamma + vec * 8);
      b.load(beta + vec * 8);
      g.to_float(gf);
      b.to_float(bfv);
#pragma unroll
      for (int j = 0; j < 8; ++j) of[j] = (v[i][j] - mean) * rstd * gf[j] + bfv[j];
      o.from_float(of);
      o.store(y + static_cast<size_t>(row) * H + vec * 8);
    }
  }
}

// dsum = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat));  dx = dsum o dropout-mask (regenerated), dres = dsum.
// Pure streaming kernel: one warp per row, nothing shared between rows (parameter gradients are a separate column reduce).
template <int VPL>
__global__ void __launch_bounds__(256) rvn_layernorm_bwd_dx(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dsum_out,
                                                            __nv_bfloat16* __restrict__ dx_drop_out, int rows, int H,
                                                            const long long* __restrict__ rng, float dropout_p) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = H >> 3;
  const bool use_drop = rng != nullptr && dropout_p > 0.f && dx_drop_out != nullptr;
  DropoutRng drng;
  if (use_drop) drng.init(rng, dropout_p);
  const float mean = mean_in[row], rstd = rstd_in[row];
  float gdy[VPL][8], xh[VPL][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
      bf16x8 a, b, g;
      float gf[8];
      a.load(dy + static_cast<size_t>(row) * H + vec * 8);
      b.load(x + static_cast<size_t>(row) * H + vec * 8);
      g.load(gamma + vec * 8);
      a.to_float(gdy[i]);
      b.to_float(xh[i]);
      g.to_float(gf);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xh[i][j] - mean) * rstd;
        gdy[i][j] *= gf[j];
        s1 += gdy[i][j];
        s2 += gdy[i][j] * xh[i][j];
      }
    }
  }
  s1 = warp_sum(s1) / H;
  s2 = warp_sum(s2) / H;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = lane + i * 32;
    if (vec < nvec) {
      float of[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) of[j] = rstd * (gdy[i][j] - s1 - xh[i][j] * s2);
      bf16x8 o;
      o.from_float(of);
      o.store(dsum_out + static_cast<size_t>(row) * H + vec * 8);
      if (use_drop) {
        const unsigned long long g = (static_cast<unsigned long long>(row) * H + vec * 8) >> 2;
        float sc[4];
        drng.scales4(g, sc);
#pragma unroll
        for (int j = 0; j < 4; ++j) of[j] *= sc[j];
        drng.scales4(g + 1, sc);
#pragma unroll
        for (int j = 0; j < 4; ++j) of[4 + j] *= sc[j];
        o.from_float(of);
        o.store(dx_drop_out + static_cast<size_t>(row) * H + vec * 8);
      }
    }
  }
}

// Self-cleaning split reduction shared by the column reducers below: partial sums are added with fp32 atomics into a
// workspace that is ZERO AT REST; the last CTA of a column block (ticket counter) converts the totals into the output
// (optionally accumulating into it: the flat gradient bucket), then zeroes the workspace slots and its counter again.
// No memset, no finalize kernel, no cast / AccumulateGrad launch per parameter.
__device__ __forceinline__ bool last_block_of_column(unsigned* counter, unsigned parts) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == parts - 1;
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}
__device__ __forceinline__ void finalize_column(float* ws, __nv_bfloat16* out, int c, int accumulate) {
  const float total = __ldcg(ws + c);
  ws[c] = 0.f;
  out[c] = __float2bfloat16(accumulate ? total + __bfloat162float(out[c]) : total);
}

// dgamma[c] (+)= sum_r dy[r,c] * xhat[r,c];  dbeta[c] (+)= sum_r dy[r,c]   (column reduce; same tiling as rvn_colsum)
__global__ void __launch_bounds__(256) rvn_layernorm_bwd_params(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                                float* __restrict__ acc, unsigned* __restrict__ counters,
                                                                __nv_bfloat16* __restrict__ dgamma, __nv_bfloat16* __restrict__ dbeta,
                                                                int accumulate, int rows, int H, int rows_per_block) {
  __shared__ float sh[2][8][256];
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int row0 = blockIdx.y * rows_per_block;
  const int row1 = min(rows, row0 + rows_per_block);
  float dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { dg[j] = 0.f; db[j] = 0.f; }
  if (c0 < H) {
    int r = row0 + rl;
    for (; r + 8 < row1; r += 16) {                     // two rows (four independent 16-byte loads) in flight per thread
      bf16x8 a0, b0, a1, b1;
      a0.load(dy + static_cast<size_t>(r) * H + c0);
      b0.load(x + static_cast<size_t>(r) * H + c0);
      a1.load(dy + static_cast<size_t>(r + 8) * H + c0);
      b1.load(x + static_cast<size_t>(r + 8) * H + c0);
      float g0[8], x0[8], g1[8], x1[8];
      a0.to_float(g0); b0.to_float(x0); a1.to_float(g1); b1.to_float(x1);
      const float m0 = mean_in[r], s0 = rstd_in[r], m1 = mean_in[r + 8], s1 = rstd_in[r + 8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dg[j] += g0[j] * (x0[j] - m0) * s0 + g1[j] * (x1[j] - m1) * s1;
        db[j] += g0[j] + g1[j];
      }
    }
    for (; r < row1; r += 8) {
      bf16x8 a, b;
      float g[8], xv[8];
      a.load(dy + static_cast<size_t>(r) * H + c0);
      b.load(x + static_cast<size_t>(r) * H + c0);
      a.to_float(g);
      b.to_float(xv);
      const float mean = mean_in[r], rstd = rstd_in[r];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dg[j] += g[j] * (xv[j] - mean) * rstd;
        db[j] += g[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[0][rl][cg * 8 + j] = dg[j];
    sh[1][rl][cg * 8 + j] = db[j];
  }
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < H) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += sh[0][i][threadIdx.x]; t1 += sh[1][i][threadIdx.x]; }
    atomicAdd(&acc[c], t0);
    atomicAdd(&acc[H + c], t1);
  }
  if (last_block_of_column(counters + blockIdx.x, gridDim.y)) {
    if (c < H) {
      finalize_column(acc, dgamma, c, accumulate);
      finalize_column(acc + H, dbeta, c, accumulate);
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;
  }
}

// ------------------------------------------------------------------------------------------------ cross entropy
// one CTA per row.  loss_row = logsumexp(row) - row[target]; rows with target == ignore_index contribute 0.
// stats[0] += loss_row, stats[1] += 1 (valid rows).  lse is kept for the backward pass.
__global__ void __launch_bounds__(256) rvn_ce_fwd(const __nv_bfloat16* __restrict__ logits, int ld, const int64_t* __restrict__ target,
                                                  float* __restrict__ lse_out, float* __restrict__ stats, int V, int64_t ignore_index) {
  const int row = blockIdx.x;
  const int64_t t = target[row];
  if (t == ignore_index) {
    if (threadIdx.x == 0) lse_out[row] = 0.f;
    return;
  }
  const __nv_bfloat16* p = logits + static_cast<size_t>(row) * ld;
  __shared__ float sh[8];
  float m = -INFINITY;
  const int nvec = V >> 3;
  for (int vec = threadIdx.x; vec < nvec; vec += blockDim.x) {
    bf16x8 a;
    float f[8];
    a.load(p + vec * 8);
    a.to_float(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, f[j]);
  }
  for (int c = nvec * 8 + threadIdx.x; c < V; c += blockDim.x) m = fmaxf(m, __bfloat162float(p[c]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  m = sh[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
  __syncthreads();
  float s = 0.f;
  for (int vec = threadIdx.x; vec < nvec; vec += blockDim.x) {
    bf16x8 a;
    float f[8];
    a.load(p + vec * 8);
    a.to_float(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
  }
  for (int c = nvec * 8 + threadIdx.x; c < V; c += blockDim.x) s += __expf(__bfloat162float(p[c]) - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sh[i];
    const float lse = m + __logf(tot);
    lse_out[row] = lse;
    atomicAdd(&stats[0], lse - __bfloat162float(p[t]));
    atomicAdd(&stats[1], 1.0f);
  }
}

// dlogits[row, c] = (exp(logit - lse) - [c == target]) * gscale[0] / max(stats[1], 1); zero for ignored rows and pad columns
__global__ void __launch_bounds__(256) rvn_ce_bwd(const __nv_bfloat16* __restrict__ logits, int ld, const int64_t* __restrict__ target,
                                                  const float* __restrict__ lse_in, const float* __restrict__ stats,
                                                  const float* __restrict__ gscale, __nv_bfloat16* __restrict__ dlogits, int ld_out,
                                                  int V, int64_t ignore_index) {
  const int row = blockIdx.x;
  const int64_t t = target[row];
  __nv_bfloat16* o = dlogits + static_cast<size_t>(row) * ld_out;
  const int nvec_out = ld_out >> 3;
  if (t == ignore_index) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int vec = threadIdx.x; vec < nvec_out; vec += blockDim.x) *reinterpret_cast<uint4*>(o + vec * 8) = z;
    return;
  }
  const __nv_bfloat16* p = logits + static_cast<size_t>(row) * ld;
  const float lse = lse_in[row];
  const float scale = gscale[0] / fmaxf(stats[1], 1.0f);
  for (int vec = threadIdx.x; vec < nvec_out; vec += blockDim.x) {
    float f[8];
    const int c0 = vec * 8;
    if (c0 + 8 <= V) {
      bf16x8 a;
      a.load(p + c0);
      a.to_float(f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (c0 + j < V) ? __bfloat162float(p[c0 + j]) : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = (c0 + j < V) ? __expf(f[j] - lse) : 0.f;
      if (c0 + j == t) g -= 1.0f;
      f[j] = g * scale;
    }
    bf16x8 r;
    r.from_float(f);
    r.store(o + c0);
  }
}

// loss = stats[0] / max(stats[1], 1)
__global__ void rvn_ce_finalize(const float* __restrict__ stats, float* __restrict__ loss) {
  loss[0] = stats[0] / fmaxf(stats[1], 1.0f);
}

// ------------------------------------------------------------------------------------------------ column sum (bias grad)
// out[c] (+)= sum_r x[r, c]; x bf16 [rows, ld] (ld % 8 == 0).  256 threads = 32 column groups (8 columns, one 16 B load each)
// x 8 row lanes; grid = (ceil(C/256), row_chunks); fp32 atomics into the zero-at-rest workspace, last CTA per column block
// writes the bf16 result (see last_block_of_column).
__global__ void __launch_bounds__(256) rvn_colsum(const __nv_bfloat16* __restrict__ x, int ld, float* __restrict__ ws,
                                                  unsigned* __restrict__ counters, __nv_bfloat16* __restrict__ out, int accumulate,
                                                  int rows, int C, int rows_per_block) {
  __shared__ float sh[8][256];
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cg * 8;
  const int row0 = blockIdx.y * rows_per_block;
  const int row1 = min(rows, row0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 + 8 <= C) {
    int r = row0 + rl;
    // four independent 16-byte loads in flight per thread: the kernel is a pure stream and was latency-bound with one
    // (8 us for the 6.3 MB of a 4096 x 768 gradient: 0.8 TB/s)
    for (; r + 24 < row1; r += 32) {
      bf16x8 a0, a1, a2, a3;
      a0.load(x + static_cast<size_t>(r) * ld + c0);
      a1.load(x + static_cast<size_t>(r + 8) * ld + c0);
      a2.load(x + static_cast<size_t>(r + 16) * ld + c0);
      a3.load(x + static_cast<size_t>(r + 24) * ld + c0);
      float f0[8], f1[8], f2[8], f3[8];
      a0.to_float(f0); a1.to_float(f1); a2.to_float(f2); a3.to_float(f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; r < row1; r += 8) {
      bf16x8 a;
      float f[8];
      a.load(x + static_cast<size_t>(r) * ld + c0);
      a.to_float(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  } else if (c0 < C) {
    for (int r = row0 + rl; r < row1; r += 8)
      for (int j = 0; j < 8 && c0 + j < C; ++j) acc[j] += __bfloat162float(x[static_cast<size_t>(r) * ld + c0 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(&ws[c], t);
  }
  if (last_block_of_column(counters + blockIdx.x, gridDim.y)) {
    if (c < C) finalize_column(ws, out, c, accumulate);
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;
  }
}

}  // namespace rvn

#define RVN_LN_DISPATCH(FN, ...)                       \
  if (vpl <= 1) FN<1> __VA_ARGS__;                     \
  else if (vpl <= 2) FN<2> __VA_ARGS__;                \
  else if (vpl <= 3) FN<3> __VA_ARGS__;                \
  else if (vpl <= 4) FN<4> __VA_ARGS__;                \
  else if (vpl <= 8) FN<8> __VA_ARGS__;                \
  else if (vpl <= 16) FN<16> __VA_ARGS__;              \
  else return -1;

extern "C" int rvn_layernorm_forward(const void* x, const void* res, const void* gamma, const void* beta, void* y, void* sum_out,
                                     float* mean, float* rstd, int rows, int H, float eps, const long long* rng, float dropout_p,
                                     cudaStream_t stream) {
  if (H % 8 != 0) return -2;
  const int vpl = (H / 8 + 31) / 32;
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
  using T = const __nv_bfloat16*;
  RVN_LN_DISPATCH(rvn::rvn_layernorm_fwd, <<<grid, block, 0, stream>>>((T)x, (T)res, (T)gamma, (T)beta, (__nv_bfloat16*)y,
                                                                        (__nv_bfloat16*)sum_out, mean, rstd, rows, H, eps, rng, dropout_p))
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// acc: >= 2*H fp32 zero-at-rest workspace, counters: >= ceil(H/256) zero-at-rest tickets (both left zeroed again on completion);
// accumulate != 0 adds the parameter gradients into dgamma / dbeta instead of overwriting them
extern "C" int rvn_layernorm_backward(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                                      void* dsum, void* dx_drop, void* dgamma, void* dbeta, float* acc, unsigned* counters,
                                      int accumulate, int num_sms, int rows, int H, const long long* rng, float dropout_p,
                                      cudaStream_t stream) {
  if (H % 8 != 0) return -2;
  const int vpl = (H / 8 + 31) / 32;
  using T = const __nv_bfloat16*;
  const int wpb = 8;
  dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
  RVN_LN_DISPATCH(rvn::rvn_layernorm_bwd_dx, <<<grid, block, 0, stream>>>((T)dy, (T)x, (T)gamma, mean, rstd, (__nv_bfloat16*)dsum,
                                                                          (__nv_bfloat16*)dx_drop, rows, H, rng, dropout_p))
  const int col_blocks = (H + 255) / 256;
  int chunks = (2 * num_sms + col_blocks - 1) / col_blocks;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks < 1) chunks = 1;
  const int rpb = (rows + chunks - 1) / chunks;
  dim3 pgrid(col_blocks, (rows + rpb - 1) / rpb);
  rvn::rvn_layernorm_bwd_params<<<pgrid, 256, 0, stream>>>((T)dy, (T)x, mean, rstd, acc, counters, (__nv_bfloat16*)dgamma,
                                                           (__nv_bfloat16*)dbeta, accumulate, rows, H, rpb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_ce_forward(const void* logits, int ld, const int64_t* target, float* lse, float* stats, float* loss, int rows,
                              int V, int64_t ignore_index, cudaStream_t stream) {
  rvn::rvn_ce_fwd<<<rows, 256, 0, stream>>>((const __nv_bfloat16*)logits, ld, target, lse, stats, V, ignore_index);
  rvn::rvn_ce_finalize<<<1, 1, 0, stream>>>(stats, loss);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_ce_backward(const void* logits, int ld, const int64_t* target, const float* lse, const float* stats,
                               const float* gscale, void* dlogits, int ld_out, int rows, int V, int64_t ignore_index,
                               cudaStream_t stream) {
  rvn::rvn_ce_bwd<<<rows, 256, 0, stream>>>((const __nv_bfloat16*)logits, ld, target, lse, stats, gscale, (__nv_bfloat16*)dlogits,
                                            ld_out, V, ignore_index);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_column_sum(const void* x, int ld, float* ws, unsigned* counters, void* out, int accumulate, int rows, int C,
                              int num_sms, cudaStream_t stream) {
  const int col_blocks = (C + 255) / 256;
  int chunks = (2 * num_sms + col_blocks - 1) / col_blocks;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks < 1) chunks = 1;
  const int rpb = (rows + chunks - 1) / chunks;
  dim3 grid(col_blocks, (rows + rpb - 1) / rpb);
  rvn::rvn_colsum<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, ld, ws, counters, (__nv_bfloat16*)out, accumulate, rows, C, rpb);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ activation backward
// dz = dy * act'(z)  (act: 1 = erf-GELU, 2 = tanh-GELU, 3 = ReLU); flat, 8 elements per thread
namespace rvn {
__global__ void __launch_bounds__(256) rvn_act_bwd(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                   __nv_bfloat16* __restrict__ dz, size_t n8, int act) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    bf16x8 a, b, o;
    float g[8], x[8];
    a.load(dy + i * 8);
    b.load(z + i * 8);
    a.to_float(g);
    b.to_float(x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d;
      if (act == 1) {
        const float cdf = 0.5f * (1.0f + erff(x[j] * 0.70710678118654752f));
        d = cdf + x[j] * 0.39894228040143268f * __expf(-0.5f * x[j] * x[j]);
      } else if (act == 2) {
        const float u = 0.7978845608028654f * (x[j] + 0.044715f * x[j] * x[j] * x[j]);
        const float t = tanhf(u);
        d = 0.5f * (1.0f + t) + 0.5f * x[j] * (1.0f - t * t) * 0.7978845608028654f * (1.0f + 0.134145f * x[j] * x[j]);
      } else {
        d = x[j] > 0.f ? 1.f : 0.f;
      }
      g[j] *= d;
    }
    o.from_float(g);
    o.store(dz + i * 8);
  }
}
}  // namespace rvn

extern "C" int rvn_activation_backward(const void* dy, const void* z, void* dz, size_t n, int act, int num_sms, cudaStream_t stream) {
  if (n % 8 != 0) return -2;
  const size_t n8 = n / 8;
  size_t blocks = (n8 + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms) * 8) blocks = static_cast<size_t>(num_sms) * 8;
  rvn::rvn_act_bwd<<<static_cast<int>(blocks ? blocks : 1), 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)z,
                                                                              (__nv_bfloat16*)dz, n8, act);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
