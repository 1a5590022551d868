// K1: fused (DP all-reduce over NVLink peer memory) + gradient scale + optimizer update + weight re-publish.
//
// One launch over the stage's FLAT buckets (runtime/param_store.py):
//     g      = scale * sum_{peer in DP group} grad_peer[i]        <- P2P loads straight from every replica's HBM
//     (m, v) = moments update, w32 = master update                 <- SGD-momentum / Adam(W) / LAMB
//     w_out  = bf16(w32)  -> next weight-version slot              <- what the next forward reads
// so the gradient never makes an extra HBM round trip between "all-reduce", "unscale", "optimizer" and "cast"
// (upstream: ring all-reduce on CPU over gRPC, then optimizer.step(), then param copy —
//  ravnest/communication.py:160-277, compute.py:180-185, utils.py:97-109).
//
// Cross-GPU synchronisation is done on the device with system-scope release/acquire flags that live in the
// symmetric-memory signal pad of every rank (no host round trip, CUDA-graph capturable):
//   barrier A (entry): everybody's gradients are complete and visible
//   barrier B (exit) : everybody finished reading my gradients (so the next backward may overwrite them)
// With n_peers == 1 the same kernels are the single-GPU fused optimizers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvn {

constexpr int MAX_PEERS = 8;

struct PeerPtrs {
  const void* grad[MAX_PEERS];     // flat gradient buffer of every DP rank (index 0..n-1, own rank included)
  uint32_t* signal[MAX_PEERS];     // signal pad of every DP rank (uint32 slots)
  int n;
  int rank;
};

struct OptArgs {
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  float grad_scale;                // 1 / (n_replicas) for averaging, or user loss scaling
  float bias_c1, bias_c2;          // Adam/LAMB bias corrections (1 if disabled)
  float clamp_value;               // LAMB ||w|| clamp
  int nesterov, first_step, decoupled_wd, lamb_adam_mode;
};

// ------------------------------------------------------------------------------------------------ device barrier
// Every CTA 0 thread p < n signals peer p and waits for peer p's signal.  ``epoch`` increases monotonically so the
// flags never need resetting.  Other CTAs wait on a local (device-scope) flag released by CTA 0.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// slot layout in each rank's signal pad: [base + peer] written by ``peer``; [base + MAX_PEERS] local broadcast flag
__device__ __forceinline__ void cross_gpu_barrier(const PeerPtrs& pp, int base, uint32_t epoch) {
  if (pp.n <= 1) return;
  if (blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x < pp.n) {
      __threadfence_system();
      st_release_sys(pp.signal[threadIdx.x] + base + pp.rank, epoch);            // tell peer "rank arrived"
      while (ld_acquire_sys(pp.signal[pp.rank] + base + threadIdx.x) < epoch) {  // wait for the peer's arrival
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(pp.signal[pp.rank] + base + MAX_PEERS, epoch);
  } else {
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(pp.signal[pp.rank] + base + MAX_PEERS) < epoch) {
      }
    }
    __syncthreads();
  }
}

// grid-wide "all CTAs done" then CTA 0 runs the exit barrier with the peers
__device__ __forceinline__ void exit_barrier(const PeerPtrs& pp, int base, uint32_t epoch, uint32_t* done_counter) {
  if (pp.n <= 1) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {          // last CTA: every CTA of this rank finished reading peer gradients
      *done_counter = 0;
      __threadfence_system();
      for (int p = 0; p < pp.n; ++p) st_release_sys(pp.signal[p] + base + pp.rank, epoch);
      for (int p = 0; p < pp.n; ++p)
        while (ld_acquire_sys(pp.signal[pp.rank] + base + p) < epoch) {
        }
    }
  }
}

__device__ __forceinline__ void exit_barrier_sys(const PeerPtrs& pp, int base, uint32_t epoch, uint32_t* done_counter);

// ------------------------------------------------------------------------------------------------ gradient fetch
template <typename G>
__device__ __forceinline__ void load4(const G* p, size_t i, float (&f)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, size_t i, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p + i);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, size_t i, float (&f)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p + i);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
template <typename W>
__device__ __forceinline__ void store4(W* p, size_t i, const float (&f)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, size_t i, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p + i) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, size_t i, const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p + i) = v;
}

template <typename G>
__device__ __forceinline__ void reduced_grad4(const PeerPtrs& pp, size_t i, float scale, float (&g)[4]) {
  load4<G>(static_cast<const G*>(pp.grad[pp.rank]), i, g);
  for (int p = 1; p < pp.n; ++p) {                       // start at rank+1 so replicas spread over the links
    int q = pp.rank + p;
    if (q >= pp.n) q -= pp.n;
    float t[4];
    load4<G>(static_cast<const G*>(pp.grad[q]), i, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] += t[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) g[j] *= scale;
}

// U vectors of 4 elements per thread (idx = i0 + u * step), ALL of one peer's loads issued before any is consumed: peer memory
// is ~2.5 us away and is not cached in the local L2, so link bandwidth = bytes in flight / latency.  Measured on 2 GPUs with one
// 8-byte load per thread per peer: 172 GB/s; U = 4 puts 148 x 512 x 32 B = 2.4 MB in flight per peer.
constexpr int GRAD_U = 8;
template <typename G>
__device__ __forceinline__ void reduced_grad_tile(const PeerPtrs& pp, size_t i0, size_t step, size_t end, float scale,
                                                  float (&g)[GRAD_U][4]) {
  const G* own = static_cast<const G*>(pp.grad[pp.rank]);
#pragma unroll
  for (int u = 0; u < GRAD_U; ++u) {
    const size_t idx = i0 + u * step;
    if (idx < end) {
      load4<G>(own, idx, g[u]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[u][j] = 0.f;
    }
  }
  for (int p = 1; p < pp.n; ++p) {                       // start at rank+1 so replicas spread over the links
    int q = pp.rank + p;
    if (q >= pp.n) q -= pp.n;
    const G* src = static_cast<const G*>(pp.grad[q]);
    float t[GRAD_U][4];
#pragma unroll
    for (int u = 0; u < GRAD_U; ++u) {
      const size_t idx = i0 + u * step;
      if (idx < end) {
        load4<G>(src, idx, t[u]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) t[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < GRAD_U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) g[u][j] += t[u][j];
  }
#pragma unroll
  for (int u = 0; u < GRAD_U; ++u)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[u][j] *= scale;
}

// ------------------------------------------------------------------------------------------------ SGD / Adam
// kind: 0 = SGD(momentum, wd, nesterov), 1 = Adam (L2 wd) / AdamW (decoupled wd)
template <typename G, typename W, int KIND>
__global__ void __launch_bounds__(512) rvn_fused_allreduce_update(PeerPtrs pp, OptArgs a, float* __restrict__ master,
                                                                  float* __restrict__ m, float* __restrict__ v,
                                                                  W* __restrict__ w_out, G* __restrict__ own_grad_zero, size_t n,
                                                                  uint32_t epoch, uint32_t* done_counter) {
  cross_gpu_barrier(pp, 0, epoch);
  const size_t step = static_cast<size_t>(blockDim.x) * 4;
  const size_t tile = step * GRAD_U;
  for (size_t base = static_cast<size_t>(blockIdx.x) * tile; base < n; base += static_cast<size_t>(gridDim.x) * tile) {
   const size_t i0 = base + threadIdx.x * 4;
   float gt[GRAD_U][4];
   reduced_grad_tile<G>(pp, i0, step, n, a.grad_scale, gt);
#pragma unroll
   for (int u = 0; u < GRAD_U; ++u) {
    const size_t i = i0 + u * step;
    if (i >= n) break;
    float g[4], w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = gt[u][j];
    load4<float>(master, i, w);
    if (KIND == 0) {
      float mom[4];
      if (a.momentum != 0.f) load4<float>(m, i, mom);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = g[j] + a.weight_decay * w[j];
        if (a.momentum != 0.f) {
          mom[j] = a.first_step ? d : a.momentum * mom[j] + (1.f - a.dampening) * d;
          d = a.nesterov ? d + a.momentum * mom[j] : mom[j];
        }
        w[j] -= a.lr * d;
      }
      if (a.momentum != 0.f) store4<float>(m, i, mom);
    } else {
      float m1[4], v1[4];
      load4<float>(m, i, m1);
      load4<float>(v, i, v1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gj = g[j];
        if (!a.decoupled_wd) gj += a.weight_decay * w[j];
        else w[j] *= 1.f - a.lr * a.weight_decay;
        m1[j] = a.beta1 * m1[j] + (1.f - a.beta1) * gj;
        v1[j] = a.beta2 * v1[j] + (1.f - a.beta2) * gj * gj;
        const float denom = sqrtf(v1[j]) / a.bias_c2 + a.eps;        // bias_c2 = sqrt(1 - beta2^t)
        w[j] -= (a.lr / a.bias_c1) * m1[j] / denom;                  // bias_c1 = 1 - beta1^t
      }
      store4<float>(m, i, m1);
      store4<float>(v, i, v1);
    }
    store4<float>(master, i, w);
    store4<W>(w_out, i, w);
   }
  }
  exit_barrier(pp, MAX_PEERS + 1, epoch, done_counter);
  (void)own_grad_zero;
}

// ------------------------------------------------------------------------------------------------ LAMB (two phases)
// phase 1: reduce gradient, update moments, write the reduced+scaled gradient-derived update direction implicitly
//          (recomputed in phase 2), accumulate per-parameter ||w||^2 and ||u||^2.
// seg_of_block[b] = parameter index of CTA b's chunk; chunk_start/chunk_end delimit it (chunks never straddle params).
template <typename G>
__global__ void __launch_bounds__(512) rvn_fused_allreduce_lamb1(PeerPtrs pp, OptArgs a, const float* __restrict__ master,
                                                                 float* __restrict__ m, float* __restrict__ v,
                                                                 const int* __restrict__ chunk_param, const long long* __restrict__ chunk_start,
                                                                 const long long* __restrict__ chunk_end, float* __restrict__ norms,
                                                                 uint32_t epoch, uint32_t* done_counter) {
  cross_gpu_barrier(pp, 0, epoch);
  const int param = chunk_param[blockIdx.x];
  const size_t s = chunk_start[blockIdx.x], e = chunk_end[blockIdx.x];
  float wsq = 0.f, usq = 0.f;
  const size_t step = static_cast<size_t>(blockDim.x) * 4;
  for (size_t base = s; base < e; base += step * GRAD_U) {
   const size_t i0 = base + threadIdx.x * 4;
   float gt[GRAD_U][4];
   reduced_grad_tile<G>(pp, i0, step, e, a.grad_scale, gt);
#pragma unroll
   for (int u = 0; u < GRAD_U; ++u) {
    const size_t i = i0 + u * step;
    if (i >= e) break;
    float g[4], w[4], m1[4], v1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = gt[u][j];
    load4<float>(master, i, w);
    load4<float>(m, i, m1);
    load4<float>(v, i, v1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m1[j] = a.beta1 * m1[j] + (1.f - a.beta1) * g[j];
      v1[j] = a.beta2 * v1[j] + (1.f - a.beta2) * g[j] * g[j];
      const float u = m1[j] / (sqrtf(v1[j]) + a.eps) + a.weight_decay * w[j];   // debias is folded into lr (torch_optimizer.Lamb)
      if (i + j < e) {                                   // padding lanes (zeros) contribute nothing anyway
        wsq += w[j] * w[j];
        usq += u * u;
      }
    }
    store4<float>(m, i, m1);
    store4<float>(v, i, v1);
   }
  }
  __shared__ float red[2][16];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wsq += __shfl_xor_sync(0xffffffffu, wsq, o);
    usq += __shfl_xor_sync(0xffffffffu, usq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = wsq;
    red[1][threadIdx.x >> 5] = usq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { a0 += red[0][i]; a1 += red[1][i]; }
    atomicAdd(&norms[2 * param], a0);
    atomicAdd(&norms[2 * param + 1], a1);
  }
  exit_barrier_sys(pp, MAX_PEERS + 1, epoch, done_counter);   // system scope: in two-shot mode the peers read these norms next
}

template <typename W>
__global__ void __launch_bounds__(512) rvn_fused_lamb2(OptArgs a, float* __restrict__ master, const float* __restrict__ m,
                                                       const float* __restrict__ v, W* __restrict__ w_out,
                                                       const int* __restrict__ chunk_param, const long long* __restrict__ chunk_start,
                                                       const long long* __restrict__ chunk_end, const float* __restrict__ norms) {
  const int param = chunk_param[blockIdx.x];
  const size_t s = chunk_start[blockIdx.x], e = chunk_end[blockIdx.x];
  const float wn = fminf(sqrtf(norms[2 * param]), a.clamp_value);
  const float un = sqrtf(norms[2 * param + 1]);
  const float trust = (a.lamb_adam_mode || wn == 0.f || un == 0.f) ? 1.f : wn / un;
  for (size_t i = s + threadIdx.x * 4; i < e; i += blockDim.x * 4) {
    float w[4], m1[4], v1[4];
    load4<float>(master, i, w);
    load4<float>(m, i, m1);
    load4<float>(v, i, v1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = m1[j] / (sqrtf(v1[j]) + a.eps) + a.weight_decay * w[j];
      w[j] -= a.lr * trust * u;
    }
    store4<float>(master, i, w);
    store4<W>(w_out, i, w);
  }
}


// ------------------------------------------------------------------------------------------------ two-shot variants (K1 v2)
// One-shot (above): every rank pulls every peer's WHOLE gradient bucket ((N-1) x bucket over its links) and every rank redoes the
// full optimizer update (whole master / m / v through HBM).  Two-shot: rank r owns the contiguous shard r of the flat bucket;
//   shot 1 (reduce-scatter): pull only shard r of every peer's gradients,
//   update only shard r of master / m / v (1/N of the optimizer's HBM traffic),
//   shot 2 (all-gather): PUSH the new compute-dtype weights of shard r into every peer's landing buffer with posted P2P stores.
// Per rank (N-1)/N x bucket crosses the link in each direction — at 8 replicas 8x less than one-shot — and the kernel is still ONE
// launch with the same device-side barriers (entry: gradients complete everywhere; exit: every peer has finished reading my
// gradients AND writing its shard into my landing buffer).
struct WeightPeers {
  void* w[MAX_PEERS];              // landing buffer (compute dtype, whole bucket) of every DP rank
};

// like exit_barrier, but every CTA publishes its P2P stores at system scope before it counts itself done
__device__ __forceinline__ void exit_barrier_sys(const PeerPtrs& pp, int base, uint32_t epoch, uint32_t* done_counter) {
  if (pp.n <= 1) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t prev = atomicAdd(done_counter, 1u);
    if (prev == gridDim.x - 1) {
      *done_counter = 0;
      __threadfence_system();
      for (int p = 0; p < pp.n; ++p) st_release_sys(pp.signal[p] + base + pp.rank, epoch);
      for (int p = 0; p < pp.n; ++p)
        while (ld_acquire_sys(pp.signal[pp.rank] + base + p) < epoch) {
        }
    }
  }
}

template <typename G, typename W, int KIND>
__global__ void __launch_bounds__(512) rvn_fused_twoshot_update(PeerPtrs pp, WeightPeers wp, OptArgs a, float* __restrict__ master,
                                                                float* __restrict__ m, float* __restrict__ v, size_t shard_begin,
                                                                size_t shard_end, uint32_t epoch, uint32_t* done_counter) {
  cross_gpu_barrier(pp, 0, epoch);
  const size_t step = static_cast<size_t>(blockDim.x) * 4;
  const size_t tile = step * GRAD_U;
  for (size_t base = shard_begin + static_cast<size_t>(blockIdx.x) * tile; base < shard_end; base += static_cast<size_t>(gridDim.x) * tile) {
    const size_t i0 = base + threadIdx.x * 4;
    float gt[GRAD_U][4];
    reduced_grad_tile<G>(pp, i0, step, shard_end, a.grad_scale, gt);
#pragma unroll
    for (int u = 0; u < GRAD_U; ++u) {
      const size_t i = i0 + u * step;
      if (i >= shard_end) break;
      float g[4], w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = gt[u][j];
      load4<float>(master, i, w);
      if (KIND == 0) {
        float mom[4];
        if (a.momentum != 0.f) load4<float>(m, i, mom);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = g[j] + a.weight_decay * w[j];
          if (a.momentum != 0.f) {
            mom[j] = a.first_step ? d : a.momentum * mom[j] + (1.f - a.dampening) * d;
            d = a.nesterov ? d + a.momentum * mom[j] : mom[j];
          }
          w[j] -= a.lr * d;
        }
        if (a.momentum != 0.f) store4<float>(m, i, mom);
      } else {
        float m1[4], v1[4];
        load4<float>(m, i, m1);
        load4<float>(v, i, v1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float gj = g[j];
          if (!a.decoupled_wd) gj += a.weight_decay * w[j];
          else w[j] *= 1.f - a.lr * a.weight_decay;
          m1[j] = a.beta1 * m1[j] + (1.f - a.beta1) * gj;
          v1[j] = a.beta2 * v1[j] + (1.f - a.beta2) * gj * gj;
          const float denom = sqrtf(v1[j]) / a.bias_c2 + a.eps;
          w[j] -= (a.lr / a.bias_c1) * m1[j] / denom;
        }
        store4<float>(m, i, m1);
        store4<float>(v, i, v1);
      }
      store4<float>(master, i, w);
      for (int q = 0; q < pp.n; ++q) {                     // all-gather by posted stores, own copy first, then rank+1, ...
        int t = pp.rank + q;
        if (t >= pp.n) t -= pp.n;
        store4<W>(static_cast<W*>(wp.w[t]), i, w);
      }
    }
  }
  exit_barrier_sys(pp, MAX_PEERS + 1, epoch, done_counter);
}

// LAMB two-shot, phase 1: like rvn_fused_allreduce_lamb1 but only over the chunks of this rank's shard (chunk arrays are passed
// already offset); the partial squared norms of the shard land in THIS rank's symmetric norm buffer.
// phase 2: trust ratio from the sum of every rank's partial norms (P2P loads of 2 floats per parameter), update of the shard,
// push of the new weights into every peer's landing buffer.
struct NormPeers {
  const float* norms[MAX_PEERS];
};

template <typename W>
__global__ void __launch_bounds__(512) rvn_fused_twoshot_lamb2(PeerPtrs pp, WeightPeers wp, NormPeers np, OptArgs a,
                                                               float* __restrict__ master, const float* __restrict__ m,
                                                               const float* __restrict__ v, const int* __restrict__ chunk_param,
                                                               const long long* __restrict__ chunk_start,
                                                               const long long* __restrict__ chunk_end, uint32_t epoch,
                                                               uint32_t* done_counter) {
  const int param = chunk_param[blockIdx.x];
  const size_t s = chunk_start[blockIdx.x], e = chunk_end[blockIdx.x];
  float wsq = 0.f, usq = 0.f;
  for (int q = 0; q < pp.n; ++q) {
    wsq += np.norms[q][2 * param];
    usq += np.norms[q][2 * param + 1];
  }
  const float wn = fminf(sqrtf(wsq), a.clamp_value);
  const float un = sqrtf(usq);
  const float trust = (a.lamb_adam_mode || wn == 0.f || un == 0.f) ? 1.f : wn / un;
  for (size_t i = s + threadIdx.x * 4; i < e; i += blockDim.x * 4) {
    float w[4], m1[4], v1[4];
    load4<float>(master, i, w);
    load4<float>(m, i, m1);
    load4<float>(v, i, v1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = m1[j] / (sqrtf(v1[j]) + a.eps) + a.weight_decay * w[j];
      w[j] -= a.lr * trust * u;
    }
    store4<float>(master, i, w);
    for (int q = 0; q < pp.n; ++q) {
      int t = pp.rank + q;
      if (t >= pp.n) t -= pp.n;
      store4<W>(static_cast<W*>(wp.w[t]), i, w);
    }
  }
  exit_barrier_sys(pp, 2 * MAX_PEERS + 2, epoch, done_counter);
}

// parameter averaging (reference semantics of parallel_ring_reduce): master = mean over replicas of master_peer
template <typename W>
__global__ void __launch_bounds__(512) rvn_fused_average(PeerPtrs pp, float* __restrict__ master_out, W* __restrict__ w_out, size_t n,
                                                         uint32_t epoch, uint32_t* done_counter) {
  cross_gpu_barrier(pp, 0, epoch);
  const float inv = 1.f / pp.n;
  const size_t step = static_cast<size_t>(blockDim.x) * 4;
  const size_t tile = step * GRAD_U;
  for (size_t base = static_cast<size_t>(blockIdx.x) * tile; base < n; base += static_cast<size_t>(gridDim.x) * tile) {
    const size_t i0 = base + threadIdx.x * 4;
    float gt[GRAD_U][4];
    reduced_grad_tile<float>(pp, i0, step, n, inv, gt);
#pragma unroll
    for (int u = 0; u < GRAD_U; ++u) {
      const size_t i = i0 + u * step;
      if (i >= n) break;
      store4<float>(master_out, i, gt[u]);
      store4<W>(w_out, i, gt[u]);
    }
  }
  exit_barrier(pp, MAX_PEERS + 1, epoch, done_counter);
}

}  // namespace rvn

using rvn::OptArgs;
using rvn::PeerPtrs;

static int grid_for(size_t n, int num_sms) {
  size_t blocks = (n / 4 + 511) / 512;
  size_t cap = static_cast<size_t>(num_sms) * 2;
  return static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// grad_is_bf16 / out_is_bf16 select the template; kind: 0 SGD, 1 Adam
extern "C" int rvn_fused_update(const PeerPtrs* pp, const OptArgs* a, int kind, float* master, float* m, float* v, void* w_out,
                                int grad_is_bf16, int out_is_bf16, size_t n, uint32_t epoch, uint32_t* done_counter, int num_sms,
                                cudaStream_t stream) {
  // with peers every rank must run the SAME number of co-resident CTAs that spin on flags -> cap at num_sms
  int grid = grid_for(n, num_sms);
  if (pp->n > 1 && grid > num_sms) grid = num_sms;
#define RVN_LAUNCH(G, W, K)                                                                                          \
  rvn::rvn_fused_allreduce_update<G, W, K><<<grid, 512, 0, stream>>>(*pp, *a, master, m, v, static_cast<W*>(w_out),   \
                                                                     nullptr, n, epoch, done_counter)
  if (kind == 0) {
    if (grad_is_bf16 && out_is_bf16) RVN_LAUNCH(__nv_bfloat16, __nv_bfloat16, 0);
    else if (!grad_is_bf16 && !out_is_bf16) RVN_LAUNCH(float, float, 0);
    else if (grad_is_bf16) RVN_LAUNCH(__nv_bfloat16, float, 0);
    else RVN_LAUNCH(float, __nv_bfloat16, 0);
  } else {
    if (grad_is_bf16 && out_is_bf16) RVN_LAUNCH(__nv_bfloat16, __nv_bfloat16, 1);
    else if (!grad_is_bf16 && !out_is_bf16) RVN_LAUNCH(float, float, 1);
    else if (grad_is_bf16) RVN_LAUNCH(__nv_bfloat16, float, 1);
    else RVN_LAUNCH(float, __nv_bfloat16, 1);
  }
#undef RVN_LAUNCH
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_fused_lamb(const PeerPtrs* pp, const OptArgs* a, float* master, float* m, float* v, void* w_out, int grad_is_bf16,
                              int out_is_bf16, const int* chunk_param, const long long* chunk_start, const long long* chunk_end,
                              int n_chunks, float* norms, int n_params, uint32_t epoch, uint32_t* done_counter, cudaStream_t stream) {
  cudaMemsetAsync(norms, 0, sizeof(float) * 2 * n_params, stream);
  if (grad_is_bf16)
    rvn::rvn_fused_allreduce_lamb1<__nv_bfloat16><<<n_chunks, 512, 0, stream>>>(*pp, *a, master, m, v, chunk_param, chunk_start,
                                                                              chunk_end, norms, epoch, done_counter);
  else
    rvn::rvn_fused_allreduce_lamb1<float><<<n_chunks, 512, 0, stream>>>(*pp, *a, master, m, v, chunk_param, chunk_start, chunk_end,
                                                                      norms, epoch, done_counter);
  if (out_is_bf16)
    rvn::rvn_fused_lamb2<__nv_bfloat16><<<n_chunks, 512, 0, stream>>>(*a, master, m, v, static_cast<__nv_bfloat16*>(w_out),
                                                                    chunk_param, chunk_start, chunk_end, norms);
  else
    rvn::rvn_fused_lamb2<float><<<n_chunks, 512, 0, stream>>>(*a, master, m, v, static_cast<float*>(w_out), chunk_param, chunk_start,
                                                            chunk_end, norms);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}


using rvn::NormPeers;
using rvn::WeightPeers;

// two-shot SGD / Adam(W): rank owns [shard_begin, shard_end) of the flat bucket; wp = landing buffers (compute dtype) of all ranks
extern "C" int rvn_fused_twoshot_update(const PeerPtrs* pp, const WeightPeers* wp, const OptArgs* a, int kind, float* master, float* m,
                                        float* v, int grad_is_bf16, int out_is_bf16, size_t shard_begin, size_t shard_end, uint32_t epoch,
                                        uint32_t* done_counter, int num_sms, cudaStream_t stream) {
  int grid = grid_for(shard_end - shard_begin, num_sms);
  if (grid > num_sms) grid = num_sms;
#define RVN_LAUNCH2(G, W, K)                                                                                           \
  rvn::rvn_fused_twoshot_update<G, W, K><<<grid, 512, 0, stream>>>(*pp, *wp, *a, master, m, v, shard_begin, shard_end, epoch, done_counter)
  if (kind == 0) {
    if (grad_is_bf16 && out_is_bf16) RVN_LAUNCH2(__nv_bfloat16, __nv_bfloat16, 0);
    else if (!grad_is_bf16 && !out_is_bf16) RVN_LAUNCH2(float, float, 0);
    else if (grad_is_bf16) RVN_LAUNCH2(__nv_bfloat16, float, 0);
    else RVN_LAUNCH2(float, __nv_bfloat16, 0);
  } else {
    if (grad_is_bf16 && out_is_bf16) RVN_LAUNCH2(__nv_bfloat16, __nv_bfloat16, 1);
    else if (!grad_is_bf16 && !out_is_bf16) RVN_LAUNCH2(float, float, 1);
    else if (grad_is_bf16) RVN_LAUNCH2(__nv_bfloat16, float, 1);
    else RVN_LAUNCH2(float, __nv_bfloat16, 1);
  }
#undef RVN_LAUNCH2
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// two-shot LAMB over the n_chunks chunks of this rank's shard (chunk arrays already offset to the shard's first chunk);
// norms = this rank's symmetric partial-norm buffer (zeroed here), np = every rank's
extern "C" int rvn_fused_twoshot_lamb(const PeerPtrs* pp, const WeightPeers* wp, const NormPeers* np, const OptArgs* a, float* master,
                                      float* m, float* v, int grad_is_bf16, int out_is_bf16, const int* chunk_param,
                                      const long long* chunk_start, const long long* chunk_end, int n_chunks, float* norms, int n_params,
                                      uint32_t epoch, uint32_t* done_counter, cudaStream_t stream) {
  cudaMemsetAsync(norms, 0, sizeof(float) * 2 * n_params, stream);
  if (n_chunks <= 0) return 2;
  if (grad_is_bf16)
    rvn::rvn_fused_allreduce_lamb1<__nv_bfloat16><<<n_chunks, 512, 0, stream>>>(*pp, *a, master, m, v, chunk_param, chunk_start,
                                                                              chunk_end, norms, epoch, done_counter);
  else
    rvn::rvn_fused_allreduce_lamb1<float><<<n_chunks, 512, 0, stream>>>(*pp, *a, master, m, v, chunk_param, chunk_start, chunk_end,
                                                                      norms, epoch, done_counter);
  if (out_is_bf16)
    rvn::rvn_fused_twoshot_lamb2<__nv_bfloat16><<<n_chunks, 512, 0, stream>>>(*pp, *wp, *np, *a, master, m, v, chunk_param, chunk_start,
                                                                            chunk_end, epoch, done_counter);
  else
    rvn::rvn_fused_twoshot_lamb2<float><<<n_chunks, 512, 0, stream>>>(*pp, *wp, *np, *a, master, m, v, chunk_param, chunk_start,
                                                                    chunk_end, epoch, done_counter);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int rvn_fused_param_average(const PeerPtrs* pp, float* master_out, void* w_out, int out_is_bf16, size_t n, uint32_t epoch,
                                       uint32_t* done_counter, int num_sms, cudaStream_t stream) {
  int grid = grid_for(n, num_sms);
  if (grid > num_sms) grid = num_sms;
  if (out_is_bf16)
    rvn::rvn_fused_average<__nv_bfloat16><<<grid, 512, 0, stream>>>(*pp, master_out, static_cast<__nv_bfloat16*>(w_out), n, epoch,
                                                                  done_counter);
  else
    rvn::rvn_fused_average<float><<<grid, 512, 0, stream>>>(*pp, master_out, static_cast<float*>(w_out), n, epoch, done_counter);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
