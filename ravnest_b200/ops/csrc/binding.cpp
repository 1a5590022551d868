// pybind11 glue between torch tensors and the extern "C" kernel launchers (compiled by nvcc without torch headers).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <cstdlib>

#include <map>
#include <mutex>
#include <vector>

extern "C" {
int rvn_gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* D, int ldd, int M, int N, int K,
                  const void* bias, int act, const void* dact_in, int ld_dact, int aux_mode, void* preact_out, int ld_preact,
                  int preact_mode, void* b_copy_dst, int ld_bcopy, unsigned* bcopy_flags, unsigned bcopy_epoch, int block_n, int split_k,
                  int reduce_out, float* stats, int num_sms,
                  cudaStream_t stream);
int rvn_gemm_bf16_2cta(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* D, int ldd, int M, int N, int K,
                       const void* bias, int act, int reduce_out, float* stats, int block_n, int num_sms, cudaStream_t stream);
int rvn_layernorm_forward(const void* x, const void* res, const void* gamma, const void* beta, void* y, void* sum_out, float* mean,
                          float* rstd, int rows, int H, float eps, const long long* rng, float dropout_p, cudaStream_t stream);
int rvn_layernorm_backward(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, void* dsum,
                           void* dx_drop, void* dgamma, void* dbeta, float* acc, unsigned* counters, int accumulate, int num_sms, int rows,
                           int H, const long long* rng, float dropout_p, cudaStream_t stream);
int rvn_ce_forward(const void* logits, int ld, const int64_t* target, float* lse, float* stats, float* loss, int rows, int V,
                   int64_t ignore_index, cudaStream_t stream);
int rvn_ce_backward(const void* logits, int ld, const int64_t* target, const float* lse, const float* stats, const float* gscale,
                    void* dlogits, int ld_out, int rows, int V, int64_t ignore_index, cudaStream_t stream);
int rvn_column_sum(const void* x, int ld, float* ws, unsigned* counters, void* out, int accumulate, int rows, int C, int num_sms,
                   cudaStream_t stream);
int rvn_fmha_forward(const void* qkv, void* out, float* lse, const float* kbias, int B, int S, int heads, int causal, float scale,
                     const long long* rng, float dropout_p, cudaStream_t stream);
int rvn_fmha_backward(const void* qkv, const void* o, const void* d_o, const float* lse, const float* kbias, float* delta, float* dq_accum,
                      void* dqkv, int B, int S, int heads, int causal, float scale, const long long* rng, float dropout_p, int num_sms,
                      cudaStream_t stream);
int rvn_conv_forward(const void* x, const void* w, void* y, const void* bias, int relu, int N, int H, int W, int Cin, int Ho, int Wo,
                     int Cout, int R, int S, int stride, int pad, int pad_w, int dil, int transposed, int num_sms, cudaStream_t stream);
int rvn_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int R, int S,
                   int stride, int pad, int pad_w, int dil, int num_sms, cudaStream_t stream);
int rvn_conv_forward_strided(const void* x, const void* w, void* y, const void* bias, int relu, int N, int H, int W, int Cin, int Ho, int Wo,
                             int Cout, int R, int S, int stride, int pad, int pad_w, int dil, int transposed, long long xs_w, long long xs_h,
                             long long xs_n, float* stats, int num_sms, cudaStream_t stream);
int rvn_conv_wgrad_strided(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int R, int S,
                           int stride, int pad, int pad_w, int dil, long long xs_w, long long xs_h, long long xs_n, int num_sms,
                           cudaStream_t stream);
int rvn_conv_dgrad_s2(const void* dy, const void* w, void* dx, int N, int Ho, int Wo, int Cout, int H, int W, int Cin, int R, int S, int pad,
                      int pad_w, int num_sms, cudaStream_t stream);
int rvn_stem_pack(const void* x, int x_is_bf16, void* out, int N, int C, int H, int W, int Hp, int Wp, int pad_h, int pad_w, long long sn,
                  long long sc, long long sh, long long sw, int num_sms, cudaStream_t stream);
int rvn_wgrad_flush(float* acc, void* grad, int grad_is_bf16, size_t n, int num_sms, cudaStream_t stream);
int rvn_batchnorm_forward(const void* x, const void* res, const void* gamma, const void* beta, void* running_mean, void* running_var,
                          long long* batches_tracked, float* mean, float* invstd, float* sums, void* y, long long M, int C, float eps,
                          float momentum, int relu, int training, int param_bf16, int stats_bf16, void* relu_mask, int sums_ready, int num_sms,
                          cudaStream_t stream);
int rvn_batchnorm_backward(const void* dy, const void* y, const void* x, const void* gamma, const float* mean, const float* invstd,
                           float* sums, void* dx, void* dres, void* dgamma, void* dbeta, long long M, int C, int relu, int param_bf16,
                           int accumulate, const void* relu_mask, int num_sms, cudaStream_t stream);
int rvn_pool_forward(const void* x, void* y, void* arg, int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad, int is_max,
                     int num_sms, cudaStream_t stream);
int rvn_pool_backward(const void* dy, const void* arg, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                      int is_max, int num_sms, cudaStream_t stream);
int rvn_activation_backward(const void* dy, const void* z, void* dz, size_t n, int act, int num_sms, cudaStream_t stream);
int rvn_embedding_forward(const int64_t* idx, const void* w, void* out, long long n, int H, long long V, int is_bf16, cudaStream_t stream);
int rvn_embedding_backward(const int64_t* idx, const void* dy, void* dw, int n, int H, long long V, long long padding_idx, int is_bf16,
                           cudaStream_t stream);
int rvn_dropout(const void* x, void* y, size_t n, const long long* rng, float p, int is_bf16, int num_sms, cudaStream_t stream);
int rvn_pack(const float* x, void* y, size_t n, int mode, int num_sms, cudaStream_t stream);
int rvn_rng_next(long long* state, long long* snap, cudaStream_t stream);
}

struct PeerPtrs {
  const void* grad[8];
  uint32_t* signal[8];
  int n;
  int rank;
};
struct OptArgs {
  float lr, beta1, beta2, eps, weight_decay, momentum, dampening;
  float grad_scale;
  float bias_c1, bias_c2;
  float clamp_value;
  int nesterov, first_step, decoupled_wd, lamb_adam_mode;
};

extern "C" {
int rvn_fused_update(const PeerPtrs* pp, const OptArgs* a, int kind, float* master, float* m, float* v, void* w_out, int grad_is_bf16,
                     int out_is_bf16, size_t n, uint32_t epoch, uint32_t* done_counter, int num_sms, cudaStream_t stream);
int rvn_fused_lamb(const PeerPtrs* pp, const OptArgs* a, float* master, float* m, float* v, void* w_out, int grad_is_bf16,
                   int out_is_bf16, const int* chunk_param, const long long* chunk_start, const long long* chunk_end, int n_chunks,
                   float* norms, int n_params, uint32_t epoch, uint32_t* done_counter, cudaStream_t stream);
struct WeightPeers { void* w[8]; };
struct NormPeers { const float* norms[8]; };
int rvn_fused_twoshot_update(const PeerPtrs* pp, const WeightPeers* wp, const OptArgs* a, int kind, float* master, float* m, float* v,
                             int grad_is_bf16, int out_is_bf16, size_t shard_begin, size_t shard_end, uint32_t epoch, uint32_t* done_counter,
                             int num_sms, cudaStream_t stream);
int rvn_fused_twoshot_lamb(const PeerPtrs* pp, const WeightPeers* wp, const NormPeers* np, const OptArgs* a, float* master, float* m, float* v,
                           int grad_is_bf16, int out_is_bf16, const int* chunk_param, const long long* chunk_start, const long long* chunk_end,
                           int n_chunks, float* norms, int n_params, uint32_t epoch, uint32_t* done_counter, cudaStream_t stream);
int rvn_fused_param_average(const PeerPtrs* pp, float* master_out, void* w_out, int out_is_bf16, size_t n, uint32_t epoch,
                            uint32_t* done_counter, int num_sms, cudaStream_t stream);
}

namespace {

int sm_count() {
  static int n = 0;
  if (n == 0) n = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  return n;
}

cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void check(int code, const char* what) {
  TORCH_CHECK(code == 0, "ravnest_b200 kernel launch failed: ", what, " (code ", code, ")");
}

int ld_of(const torch::Tensor& t) {
  TORCH_CHECK(t.dim() == 2 && t.stride(1) == 1, "expected a 2-D tensor with unit inner stride");
  TORCH_CHECK(t.stride(0) % 8 == 0 || t.size(0) == 1, "row stride must be a multiple of 8 elements (16 B) for TMA");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0, "tensor base must be 16-byte aligned");
  return static_cast<int>(t.stride(0));
}

// Tile shape (128 x BN, BN in {128, 192, 256}) and split-K factor from a wave model of the persistent kernel: one CTA per SM runs
// ceil(units / SMs) work units back to back; a unit costs (k-blocks / split) * BN / eff(BN) plus one exposed epilogue.  eff = achieved
// fraction of the tensor peak per tile shape (128-wide tiles are shared-memory-bandwidth bound: A + B = 128 B/clk).  Explicit
// block_n / split_k > 0 (or RVN_GEMM_BN / RVN_GEMM_SPLITK) override the model.
static double choose_tiling(int M, int N, int K, bool can_split, int sms, int64_t* block_n, int64_t* split_k) {
  static const int env_bn = getenv("RVN_GEMM_BN") ? atoi(getenv("RVN_GEMM_BN")) : 0;
  static const int env_sk = getenv("RVN_GEMM_SPLITK") ? atoi(getenv("RVN_GEMM_SPLITK")) : 0;
  if (*block_n <= 0 && env_bn > 0) *block_n = env_bn;
  if (*split_k <= 0 && env_sk > 0) *split_k = env_sk;
  if (!can_split) *split_k = 1;
  const int mb = (M + 127) / 128, nkb = (K + 63) / 64;
  const int bns[3] = {128, 192, 256};
  const double eff[3] = {0.76, 0.86, 0.92};
  const int splits[6] = {1, 2, 3, 4, 6, 8};
  double best = 1e30;
  int best_bn = 128, best_s = 1;
  for (int i = 0; i < 3; ++i) {
    const int bn = bns[i];
    if (*block_n > 0 && bn != *block_n) continue;
    if (bn > 128 && N < bn) continue;
    const int nb = (N + bn - 1) / bn;
    for (int j = 0; j < 6; ++j) {
      const int s = splits[j];
      if (*split_k > 0 && s != *split_k) continue;
      if (s > 1 && (!can_split || nkb / s < 8)) continue;
      const long units = static_cast<long>(mb) * nb * s;
      const long waves = (units + sms - 1) / sms;
      const double unit_cost = (static_cast<double>(nkb) / s + 1.5) * bn / eff[i];     // + pipeline fill / barrier turn-around per unit
      const double cost = waves * unit_cost + 5.0 * bn + (s > 1 ? 2.0 * bn : 0.0);    // + one exposed epilogue (+ L2 reduce traffic)
      if (cost < best) { best = cost; best_bn = bn; best_s = s; }
    }
  }
  if (*block_n <= 0 || (*block_n != 128 && *block_n != 192 && *block_n != 256)) *block_n = best_bn;
  if (*split_k <= 0) *split_k = best_s;
  return best;
}

// Same wave model for the CTA-pair kernel (gemm2_sm100.cu): 256 x BN tiles on sms/2 pairs; per-tile efficiency measured on B200
// (8192^3: tensor pipe 98 % vs 84 % for the 1-CTA kernel; LM head 4096 x 30528 x 768: 125 us vs 147 us).  Returns the cost, sets *bn2.
static double pair_tiling_cost(int M, int N, int K, int sms, int* bn2) {
  static const int off = getenv("RVN_GEMM_2CTA") ? (atoi(getenv("RVN_GEMM_2CTA")) == 0) : 0;
  if (off) return 1e30;
  const int m2 = (M + 255) / 256, nkb = (K + 63) / 64, pairs = sms / 2;
  const int bns[2] = {128, 256};
  const double eff[2] = {0.72, 1.0};
  double best = 1e30;
  for (int i = 0; i < 2; ++i) {
    const int bn = bns[i];
    if (bn > 128 && N < bn) continue;
    const long tiles = static_cast<long>(m2) * ((N + bn - 1) / bn);
    const long waves = (tiles + pairs - 1) / pairs;
    const double cost = waves * (nkb + 1.5) * bn / eff[i] + 5.0 * bn;      // same units as choose_tiling (a pair tile is two CTA tiles)
    if (cost < best) { best = cost; *bn2 = bn; }
  }
  return best;
}

// D[M,N] = epi(op(A) * op(B)^T).  A: [M,K] (a_mn=false) or [K,M] (a_mn=true); B: [N,K] (b_mn=false) or [K,N] (b_mn=true).
// store_preact: 0 none, 1 also return the pre-activation z, 2 also return act'(z).  aux_mode (with dact_in): 1 D = acc * act'(aux),
// 2 D = acc + aux, 3 D = acc * aux.  Returns {D} or {D, preact}.  D is allocated with a row stride rounded up to 8 elements; a narrowed view is returned.
std::vector<torch::Tensor> gemm(const torch::Tensor& A, bool a_mn, const torch::Tensor& B, bool b_mn,
                                const c10::optional<torch::Tensor>& bias, int64_t act,
                                const c10::optional<torch::Tensor>& dact_in, int64_t store_preact, int64_t block_n,
                                int64_t aux_mode, const c10::optional<torch::Tensor>& out,
                                const c10::optional<torch::Tensor>& b_copy_dst, const c10::optional<torch::Tensor>& bcopy_flags,
                                int64_t bcopy_epoch, int64_t split_k, const c10::optional<torch::Tensor>& col_stats) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && A.scalar_type() == torch::kBFloat16 && B.scalar_type() == torch::kBFloat16,
              "gemm expects CUDA bf16 operands");
  c10::cuda::CUDAGuard guard(A.device());
  const int M = a_mn ? A.size(1) : A.size(0);
  const int K = a_mn ? A.size(0) : A.size(1);
  const int N = b_mn ? B.size(1) : B.size(0);
  TORCH_CHECK((b_mn ? B.size(0) : B.size(1)) == K, "inner dimensions differ");
  const int lda = ld_of(A), ldb = ld_of(B);
  int ldd = (N + 7) / 8 * 8;
  torch::Tensor D;
  if (out.has_value() && out->defined()) {           // write (or accumulate, aux_mode 2) straight into a caller-owned buffer
    TORCH_CHECK(out->dim() == 2 && out->size(0) == M && out->size(1) == N && out->scalar_type() == torch::kBFloat16, "bad out tensor");
    ldd = ld_of(*out);
    D = *out;
  } else {
    D = torch::empty({M, ldd}, A.options());
  }
  torch::Tensor P;
  const void* bias_ptr = nullptr;
  if (bias.has_value() && bias->defined()) {
    TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->is_contiguous() && bias->numel() == N, "bad bias");
    bias_ptr = bias->data_ptr();
  }
  const void* dact_ptr = nullptr;
  int ld_dact = 0;
  if (dact_in.has_value() && dact_in->defined()) {
    TORCH_CHECK(dact_in->size(0) == M && dact_in->size(1) == N, "dact_in must be [M,N]");
    ld_dact = ld_of(*dact_in);
    dact_ptr = dact_in->data_ptr();
  }
  void* preact_ptr = nullptr;
  if (store_preact) {
    P = torch::empty({M, (N + 7) / 8 * 8}, A.options());
    preact_ptr = P.data_ptr();
  }
  // gradient accumulation D += A*B (aux_mode 2 with the aux tile being D itself): done by TMA reduce-add, D is never read and the
  // K range may be split over several CTAs (wgrad shapes have few output tiles and a long K)
  const bool reduce_out = dact_ptr != nullptr && aux_mode == 2 && out.has_value() && out->defined() &&
                          dact_ptr == out->data_ptr() && ld_dact == ldd && !store_preact && !(bias.has_value() && bias->defined()) && act == 0;
  if (reduce_out) dact_ptr = nullptr;
  // column statistics (sum, sum of squares of the bf16 output, ADDED into a [2, N] fp32 tensor): the BatchNorm that follows a
  // pointwise convolution gets its batch statistics from this epilogue instead of reading the activation again
  float* stats_ptr = nullptr;
  if (col_stats.has_value() && col_stats->defined()) {
    TORCH_CHECK(col_stats->scalar_type() == torch::kFloat32 && col_stats->is_contiguous() && col_stats->numel() == 2 * static_cast<int64_t>(N) &&
                    col_stats->device() == A.device(), "col_stats must be a contiguous fp32 [2, N] tensor");
    TORCH_CHECK(!reduce_out && dact_ptr == nullptr && !store_preact && bias_ptr == nullptr && act == 0 &&
                    !(b_copy_dst.has_value() && b_copy_dst->defined()), "col_stats needs a plain store epilogue");
    stats_ptr = col_stats->data_ptr<float>();
    split_k = 1;
  }
  // RVN_STATS_PAIR=0: GEMMs that collect column statistics stay on the single-CTA kernel (its statistics pass is cheaper per tile)
  static const bool stats_on_pair = [] { const char* e = std::getenv("RVN_STATS_PAIR"); return e == nullptr || std::atoi(e) != 0; }();
  const bool explicit_tiling = block_n > 0 || (split_k > 0 && stats_ptr == nullptr) || (stats_ptr != nullptr && !stats_on_pair);
  const double cost1 = choose_tiling(M, N, K, reduce_out, sm_count(), &block_n, &split_k);
  // plain epilogues (bias / activation, or accumulate-into-output) may run on the CTA-pair kernel when the wave model prefers it
  if (!explicit_tiling && dact_ptr == nullptr && !store_preact && !(b_copy_dst.has_value() && b_copy_dst->defined()) && M >= 256) {
    int bn2 = 256;
    const double cost2 = pair_tiling_cost(M, N, K, sm_count(), &bn2);
    if (cost2 < 0.97 * cost1) {
      check(rvn_gemm_bf16_2cta(A.data_ptr(), lda, a_mn, B.data_ptr(), ldb, b_mn, D.data_ptr(), ldd, M, N, K, bias_ptr, static_cast<int>(act),
                               reduce_out ? 1 : 0, stats_ptr, bn2, sm_count(), cur_stream()),
            "gemm_bf16_2cta");
      std::vector<torch::Tensor> res2;
      res2.push_back(D.size(1) == N ? D : D.narrow(1, 0, N));
      return res2;
    }
  }
  void* bcopy_ptr = nullptr;
  int ld_bcopy = 0;
  if (b_copy_dst.has_value() && b_copy_dst->defined()) {   // K2: B lives in a peer's HBM; land it locally while multiplying
    TORCH_CHECK(!b_mn && b_copy_dst->sizes() == B.sizes() && b_copy_dst->scalar_type() == torch::kBFloat16, "bad b_copy_dst");
    ld_bcopy = ld_of(*b_copy_dst);
    bcopy_ptr = b_copy_dst->data_ptr();
    TORCH_CHECK(bcopy_flags.has_value() && bcopy_flags->defined() && bcopy_flags->scalar_type() == torch::kInt32 &&
                bcopy_flags->numel() >= static_cast<int64_t>((N + 127) / 128) * ((K + 63) / 64), "bcopy_flags too small");
  }
  unsigned* flags_ptr = bcopy_ptr != nullptr ? reinterpret_cast<unsigned*>(bcopy_flags->data_ptr<int>()) : nullptr;
  check(rvn_gemm_bf16(A.data_ptr(), lda, a_mn, B.data_ptr(), ldb, b_mn, D.data_ptr(), ldd, M, N, K, bias_ptr, static_cast<int>(act),
                      dact_ptr, ld_dact, static_cast<int>(aux_mode), preact_ptr, store_preact ? static_cast<int>(P.stride(0)) : ldd,
                      static_cast<int>(store_preact), bcopy_ptr, ld_bcopy, flags_ptr, static_cast<unsigned>(bcopy_epoch), static_cast<int>(block_n),
                      static_cast<int>(split_k), reduce_out ? 1 : 0, stats_ptr, sm_count(),
                      cur_stream()),
        "gemm_bf16");
  std::vector<torch::Tensor> res;
  res.push_back(D.size(1) == N ? D : D.narrow(1, 0, N));
  if (store_preact) res.push_back(P.size(1) == N ? P : P.narrow(1, 0, N));
  return res;
}

// would gemm() route a plain M x N x K product to the CTA-pair kernel?  (the caller may prefer not to fuse extra epilogue work there)
bool gemm_prefers_pair(int64_t M, int64_t N, int64_t K) {
  if (M < 256) return false;
  int64_t bn = 0, sk = 0;
  int bn2 = 256;
  const double cost1 = choose_tiling(static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), false, sm_count(), &bn, &sk);
  const double cost2 = pair_tiling_cost(static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), sm_count(), &bn2);
  return cost2 < 0.97 * cost1;
}

// 2-CTA (cta_group::2, 256 x BN tile per CTA pair) GEMM: D = act(op(A) op(B)^T + bias).  Plain epilogue only (see gemm2_sm100.cu).
torch::Tensor gemm_2cta(const torch::Tensor& A, bool a_mn, const torch::Tensor& B, bool b_mn, const c10::optional<torch::Tensor>& bias,
                        int64_t act, int64_t block_n) {
  TORCH_CHECK(A.is_cuda() && B.is_cuda() && A.scalar_type() == torch::kBFloat16 && B.scalar_type() == torch::kBFloat16, "gemm_2cta expects CUDA bf16");
  c10::cuda::CUDAGuard guard(A.device());
  const int M = a_mn ? A.size(1) : A.size(0);
  const int K = a_mn ? A.size(0) : A.size(1);
  const int N = b_mn ? B.size(1) : B.size(0);
  TORCH_CHECK((b_mn ? B.size(0) : B.size(1)) == K, "inner dimensions differ");
  const int lda = ld_of(A), ldb = ld_of(B);
  const int ldd = (N + 7) / 8 * 8;
  torch::Tensor D = torch::empty({M, ldd}, A.options());
  const void* bias_ptr = nullptr;
  if (bias.has_value() && bias->defined()) {
    TORCH_CHECK(bias->scalar_type() == torch::kBFloat16 && bias->is_contiguous() && bias->numel() == N, "bad bias");
    bias_ptr = bias->data_ptr();
  }
  check(rvn_gemm_bf16_2cta(A.data_ptr(), lda, a_mn, B.data_ptr(), ldb, b_mn, D.data_ptr(), ldd, M, N, K, bias_ptr, static_cast<int>(act), 0, nullptr,
                           block_n == 128 ? 128 : 256, sm_count(), cur_stream()),
        "gemm_bf16_2cta");
  return D.size(1) == N ? D : D.narrow(1, 0, N);
}

const long long* rng_ptr(const c10::optional<torch::Tensor>& rng) {
  if (!rng.has_value() || !rng->defined()) return nullptr;
  TORCH_CHECK(rng->scalar_type() == torch::kInt64 && rng->numel() >= 2 && rng->is_cuda(), "rng state must be a CUDA int64[2] tensor");
  return reinterpret_cast<const long long*>(rng->data_ptr<int64_t>());
}

// y = LN(dropout(x) + res).  Returns {y, mean, rstd} (+ {sum} when a residual or dropout makes it differ from x)
std::vector<torch::Tensor> layernorm_fwd(const torch::Tensor& x, const c10::optional<torch::Tensor>& res, const torch::Tensor& gamma,
                                         const torch::Tensor& beta, double eps, const c10::optional<torch::Tensor>& rng, double dropout_p) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.scalar_type() == torch::kBFloat16, "layernorm expects contiguous bf16");
  const int H = x.size(-1);
  const int rows = x.numel() / H;
  auto y = torch::empty_like(x);
  auto mean = torch::empty({rows}, x.options().dtype(torch::kFloat32));
  auto rstd = torch::empty({rows}, x.options().dtype(torch::kFloat32));
  torch::Tensor sum;
  const void* res_ptr = nullptr;
  void* sum_ptr = nullptr;
  const bool has_res = res.has_value() && res->defined();
  const long long* rp = dropout_p > 0.0 ? rng_ptr(rng) : nullptr;
  if (has_res) {
    TORCH_CHECK(res->is_contiguous() && res->sizes() == x.sizes(), "residual must match x");
    res_ptr = res->data_ptr();
  }
  if (has_res || rp != nullptr) {
    sum = torch::empty_like(x);
    sum_ptr = sum.data_ptr();
  }
  check(rvn_layernorm_forward(x.data_ptr(), res_ptr, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), sum_ptr,
                              mean.data_ptr<float>(), rstd.data_ptr<float>(), rows, H, static_cast<float>(eps), rp,
                              static_cast<float>(dropout_p), cur_stream()),
        "layernorm_fwd");
  if (sum.defined()) return {y, mean, rstd, sum};
  return {y, mean, rstd};
}

// Zero-at-rest reduction workspace of the column reducers (rowwise.cu: last_block_of_column), one per device.  The engine issues
// all compute of a process on ONE stream (eagerly or as graph replays), which is what makes sharing it safe; it must exist before
// a CUDA-graph capture starts (every capture in this code base is preceded by eager warm-up steps, which create it).
struct ReduceWs {
  torch::Tensor sums;       // fp32 [kReduceCap]
  torch::Tensor counters;   // int32 [kReduceCap / 256 + 1]
};
static constexpr int64_t kReduceCap = 1 << 17;
static ReduceWs& reduce_ws(const torch::Device& dev, int64_t need) {
  TORCH_CHECK(need <= kReduceCap, "column reduction wider than the workspace (", need, " > ", kReduceCap, ")");
  static std::mutex mu;
  static std::map<int, ReduceWs> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(static_cast<int>(dev.index()));
  if (it == cache.end()) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(cur_stream(), &st);
    TORCH_CHECK(st == cudaStreamCaptureStatusNone, "reduction workspace must be created before CUDA-graph capture (run one eager step first)");
    ReduceWs w;
    w.sums = torch::zeros({kReduceCap}, torch::TensorOptions().device(dev).dtype(torch::kFloat32));
    w.counters = torch::zeros({kReduceCap / 256 + 1}, torch::TensorOptions().device(dev).dtype(torch::kInt32));
    it = cache.emplace(static_cast<int>(dev.index()), std::move(w)).first;
  }
  return it->second;
}

// returns {dsum (grad of the normalised sum == grad of the residual), dx (== dsum, or dsum o mask with dropout), dgamma, dbeta};
// with grad_gamma / grad_beta given the parameter gradients are accumulated into them in place (undefined tensors returned)
std::vector<torch::Tensor> layernorm_bwd(const torch::Tensor& dy, const torch::Tensor& x, const torch::Tensor& gamma,
                                         const torch::Tensor& mean, const torch::Tensor& rstd, const c10::optional<torch::Tensor>& rng,
                                         double dropout_p, const c10::optional<torch::Tensor>& grad_gamma,
                                         const c10::optional<torch::Tensor>& grad_beta) {
  c10::cuda::CUDAGuard guard(x.device());
  const int H = x.size(-1);
  const int rows = x.numel() / H;
  auto dsum = torch::empty_like(x);
  const long long* rp = dropout_p > 0.0 ? rng_ptr(rng) : nullptr;
  torch::Tensor dx = rp != nullptr ? torch::empty_like(x) : dsum;
  const bool acc_mode = grad_gamma.has_value() && grad_gamma->defined() && grad_beta.has_value() && grad_beta->defined();
  torch::Tensor dgamma, dbeta;
  void *dg_ptr, *db_ptr;
  if (acc_mode) {
    TORCH_CHECK(grad_gamma->scalar_type() == torch::kBFloat16 && grad_beta->scalar_type() == torch::kBFloat16 && grad_gamma->is_contiguous() &&
                    grad_beta->is_contiguous() && grad_gamma->numel() == H && grad_beta->numel() == H, "bad LayerNorm gradient accumulators");
    dg_ptr = grad_gamma->data_ptr();
    db_ptr = grad_beta->data_ptr();
  } else {
    dgamma = torch::empty({H}, x.options());
    dbeta = torch::empty({H}, x.options());
    dg_ptr = dgamma.data_ptr();
    db_ptr = dbeta.data_ptr();
  }
  ReduceWs& ws = reduce_ws(x.device(), 2 * static_cast<int64_t>(H));
  check(rvn_layernorm_backward(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr<float>(), rstd.data_ptr<float>(),
                               dsum.data_ptr(), rp != nullptr ? dx.data_ptr() : nullptr, dg_ptr, db_ptr, ws.sums.data_ptr<float>(),
                               reinterpret_cast<unsigned*>(ws.counters.data_ptr<int>()), acc_mode ? 1 : 0, sm_count(), rows, H, rp,
                               static_cast<float>(dropout_p), cur_stream()),
        "layernorm_bwd");
  return {dsum, dx, dgamma, dbeta};
}

// returns {loss (scalar fp32), lse [rows], stats [2]}
std::vector<torch::Tensor> ce_fwd(const torch::Tensor& logits, const torch::Tensor& target, int64_t ignore_index) {
  c10::cuda::CUDAGuard guard(logits.device());
  TORCH_CHECK(logits.dim() == 2 && logits.stride(1) == 1 && logits.scalar_type() == torch::kBFloat16, "ce expects 2-D bf16 logits");
  TORCH_CHECK(logits.stride(0) % 8 == 0 && reinterpret_cast<uintptr_t>(logits.data_ptr()) % 16 == 0, "logit rows must be 16 B aligned");
  const int rows = logits.size(0), V = logits.size(1);
  auto f32 = logits.options().dtype(torch::kFloat32);
  auto lse = torch::empty({rows}, f32);
  auto stats = torch::zeros({2}, f32);
  auto loss = torch::empty({}, f32);
  check(rvn_ce_forward(logits.data_ptr(), logits.stride(0), target.data_ptr<int64_t>(), lse.data_ptr<float>(), stats.data_ptr<float>(),
                       loss.data_ptr<float>(), rows, V, ignore_index, cur_stream()),
        "ce_fwd");
  return {loss, lse, stats};
}

torch::Tensor ce_bwd(const torch::Tensor& logits, const torch::Tensor& target, const torch::Tensor& lse, const torch::Tensor& stats,
                     const torch::Tensor& gscale, int64_t ignore_index) {
  c10::cuda::CUDAGuard guard(logits.device());
  const int rows = logits.size(0), V = logits.size(1);
  const int ld_out = (V + 7) / 8 * 8;
  auto d = torch::empty({rows, ld_out}, logits.options());
  check(rvn_ce_backward(logits.data_ptr(), logits.stride(0), target.data_ptr<int64_t>(), lse.data_ptr<float>(), stats.data_ptr<float>(),
                        gscale.data_ptr<float>(), d.data_ptr(), ld_out, rows, V, ignore_index, cur_stream()),
        "ce_bwd");
  return ld_out == V ? d : d.narrow(1, 0, V);
}

// bf16 column sums of a 2-D bf16 matrix (bias gradients).  With ``into`` given the sums are ADDED to it in place (the flat gradient
// bucket) and it is returned; otherwise a fresh bf16 [C] tensor is returned.
torch::Tensor column_sum(const torch::Tensor& x, const c10::optional<torch::Tensor>& into) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.dim() == 2 && x.stride(1) == 1 && x.scalar_type() == torch::kBFloat16, "column_sum expects 2-D bf16");
  const int64_t C = x.size(1);
  const bool acc_mode = into.has_value() && into->defined();
  torch::Tensor out;
  if (acc_mode) {
    TORCH_CHECK(into->scalar_type() == torch::kBFloat16 && into->is_contiguous() && into->numel() == C, "bad column_sum accumulator");
    out = *into;
  } else {
    out = torch::empty({C}, x.options());
  }
  ReduceWs& ws = reduce_ws(x.device(), C);
  check(rvn_column_sum(x.data_ptr(), x.stride(0), ws.sums.data_ptr<float>(), reinterpret_cast<unsigned*>(ws.counters.data_ptr<int>()),
                       out.data_ptr(), acc_mode ? 1 : 0, x.size(0), C, sm_count(), cur_stream()),
        "column_sum");
  return out;
}

torch::Tensor act_bwd(const torch::Tensor& dy, const torch::Tensor& z, int64_t act) {
  c10::cuda::CUDAGuard guard(dy.device());
  TORCH_CHECK(dy.is_contiguous() && z.is_contiguous() && dy.numel() == z.numel(), "act_bwd expects contiguous same-size tensors");
  auto dz = torch::empty_like(dy);
  check(rvn_activation_backward(dy.data_ptr(), z.data_ptr(), dz.data_ptr(), dy.numel(), static_cast<int>(act), sm_count(), cur_stream()),
        "act_bwd");
  return dz;
}

// ---------------------------------------------------------------------------------------------- K10 / K11 / K12
static int float_kind(const torch::Tensor& t, const char* what) {
  TORCH_CHECK(t.scalar_type() == torch::kBFloat16 || t.scalar_type() == torch::kFloat32, what, " must be bf16 or fp32");
  return t.scalar_type() == torch::kBFloat16;
}

// out[..., :] = weight[idx[...], :]
torch::Tensor embedding_fwd(const torch::Tensor& idx, const torch::Tensor& weight) {
  c10::cuda::CUDAGuard guard(weight.device());
  TORCH_CHECK(idx.scalar_type() == torch::kInt64 && idx.is_cuda() && weight.dim() == 2 && weight.is_contiguous(), "embedding expects int64 indices and a contiguous table");
  const int bf = float_kind(weight, "embedding table");
  auto ids = idx.contiguous();
  auto shape = ids.sizes().vec();
  shape.push_back(weight.size(1));
  auto out = torch::empty(shape, weight.options());
  check(rvn_embedding_forward(ids.data_ptr<int64_t>(), weight.data_ptr(), out.data_ptr(), ids.numel(), static_cast<int>(weight.size(1)),
                              weight.size(0), bf, cur_stream()),
        "embedding_fwd");
  return out;
}

// dW[r] += sum of dy rows whose index is r.  ``into``: accumulate into this [V, H] tensor (the gradient bucket view) and return it;
// otherwise a zero-filled tensor is created.  Deterministic (no atomics).  Requires numel(idx) <= 16384.
torch::Tensor embedding_bwd(const torch::Tensor& idx, const torch::Tensor& dy, int64_t num_rows, int64_t padding_idx,
                            const c10::optional<torch::Tensor>& into) {
  c10::cuda::CUDAGuard guard(dy.device());
  const int bf = float_kind(dy, "embedding gradient");
  auto ids = idx.contiguous();
  auto g = dy.contiguous();
  const int H = g.size(-1);
  TORCH_CHECK(ids.numel() * H == g.numel() && ids.numel() <= 16384, "embedding_bwd: shape mismatch or more than 16384 indices");
  torch::Tensor dw;
  if (into.has_value() && into->defined()) {
    TORCH_CHECK(into->is_contiguous() && into->scalar_type() == g.scalar_type() && into->size(0) == num_rows && into->size(1) == H, "bad accumulator");
    dw = *into;
  } else {
    dw = torch::zeros({num_rows, H}, g.options());
  }
  check(rvn_embedding_backward(ids.data_ptr<int64_t>(), g.data_ptr(), dw.data_ptr(), static_cast<int>(ids.numel()), H, num_rows, padding_idx, bf,
                               cur_stream()),
        "embedding_bwd");
  return dw;
}

// y = x * Philox keep-scale(rng snapshot, position); the backward applies the same call to dy
torch::Tensor dropout_apply(const torch::Tensor& x, const torch::Tensor& rng, double p) {
  c10::cuda::CUDAGuard guard(x.device());
  const int bf = float_kind(x, "dropout input");
  TORCH_CHECK(x.is_contiguous() && x.numel() % (bf ? 8 : 4) == 0, "dropout expects a contiguous tensor with numel % 8 == 0 (bf16) / % 4 == 0 (fp32)");
  auto y = torch::empty_like(x);
  check(rvn_dropout(x.data_ptr(), y.data_ptr(), x.numel(), rng_ptr(rng), static_cast<float>(p), bf, sm_count(), cur_stream()), "dropout");
  return y;
}

// returns a copy of the {seed, step} state and advances step by one
torch::Tensor rng_next(torch::Tensor state) {
  c10::cuda::CUDAGuard guard(state.device());
  TORCH_CHECK(state.scalar_type() == torch::kInt64 && state.numel() >= 2 && state.is_contiguous(), "rng state must be int64[2]");
  auto snap = torch::empty({2}, state.options());
  check(rvn_rng_next(reinterpret_cast<long long*>(state.data_ptr<int64_t>()), reinterpret_cast<long long*>(snap.data_ptr<int64_t>()), cur_stream()),
        "rng_next");
  return snap;
}

// fp32 -> fp16 (clamped to the fp16 range; mode 0) or bf16 (mode 1) in one pass
torch::Tensor pack_f32(const torch::Tensor& x, int64_t mode) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.scalar_type() == torch::kFloat32, "pack expects contiguous fp32");
  auto y = torch::empty(x.sizes(), x.options().dtype(mode == 0 ? torch::kFloat16 : torch::kBFloat16));
  check(rvn_pack(x.data_ptr<float>(), y.data_ptr(), x.numel(), static_cast<int>(mode), sm_count(), cur_stream()), "pack");
  return y;
}

// qkv: [B, S, 3H] bf16 contiguous (q | k | v, heads of 64); kbias: optional [B, S] fp32 additive key bias.  Returns {out [B,S,H], lse [B,heads,S]}
std::vector<torch::Tensor> fmha_fwd(const torch::Tensor& qkv, int64_t heads, const c10::optional<torch::Tensor>& kbias, bool causal,
                                    double scale, const c10::optional<torch::Tensor>& rng, double dropout_p) {
  c10::cuda::CUDAGuard guard(qkv.device());
  TORCH_CHECK(qkv.is_contiguous() && qkv.scalar_type() == torch::kBFloat16 && qkv.dim() == 3, "fmha expects contiguous bf16 [B,S,3H]");
  const int B = qkv.size(0), S = qkv.size(1), H = qkv.size(2) / 3;
  TORCH_CHECK(H == heads * 64, "fmha kernel is specialised for head_dim 64");
  auto out = torch::empty({B, S, H}, qkv.options());
  auto lse = torch::empty({B, heads, S}, qkv.options().dtype(torch::kFloat32));
  const float* kb = nullptr;
  if (kbias.has_value() && kbias->defined()) {
    TORCH_CHECK(kbias->is_contiguous() && kbias->scalar_type() == torch::kFloat32 && kbias->numel() == static_cast<int64_t>(B) * S, "bad kbias");
    kb = kbias->data_ptr<float>();
  }
  check(rvn_fmha_forward(qkv.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), kb, B, S, static_cast<int>(heads), causal,
                         static_cast<float>(scale), rng_ptr(rng), static_cast<float>(dropout_p), cur_stream()),
        "fmha_fwd");
  return {out, lse};
}

torch::Tensor fmha_bwd(const torch::Tensor& qkv, const torch::Tensor& out, const torch::Tensor& d_out, const torch::Tensor& lse,
                       int64_t heads, const c10::optional<torch::Tensor>& kbias, bool causal, double scale,
                       const c10::optional<torch::Tensor>& rng, double dropout_p) {
  c10::cuda::CUDAGuard guard(qkv.device());
  TORCH_CHECK(qkv.is_contiguous() && out.is_contiguous() && d_out.is_contiguous(), "fmha_bwd expects contiguous tensors");
  const int B = qkv.size(0), S = qkv.size(1), H = qkv.size(2) / 3;
  auto dqkv = torch::empty_like(qkv);
  auto f32 = qkv.options().dtype(torch::kFloat32);
  auto dq_accum = torch::zeros({B, S, H}, f32);
  auto delta = torch::empty({B, heads, S}, f32);
  const float* kb = nullptr;
  if (kbias.has_value() && kbias->defined()) kb = kbias->data_ptr<float>();
  check(rvn_fmha_backward(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr<float>(), kb, delta.data_ptr<float>(),
                          dq_accum.data_ptr<float>(), dqkv.data_ptr(), B, S, static_cast<int>(heads), causal, static_cast<float>(scale),
                          rng_ptr(rng), static_cast<float>(dropout_p), sm_count(), cur_stream()),
        "fmha_bwd");
  return dqkv;
}

// x: [N,H,W,Cin] bf16 contiguous (NHWC); w: [Cout,R,S,Cin] contiguous.  transposed=false: y = conv(x, w) -> [N,Ho,Wo,Cout].
// transposed=true (dgrad, stride 1): x is dy [N,Ho,Wo,Cout], result is dx [N,H,W,Cin]; `out_hw` gives the result's H, W.
// element strides (W, H, N) of an NHWC activation view for the TMA descriptor: {0, 0, 0} when dense.  Views with a unit channel
// stride and 16-byte-multiple outer strides are consumed in place — including OVERLAPPING ones (the small-Cin stem, conv_aux.cu)
static void nhwc_strides(const torch::Tensor& x, long long* st) {
  st[0] = st[1] = st[2] = 0;
  if (x.is_contiguous()) return;
  TORCH_CHECK(x.dim() == 4 && x.stride(3) == 1 && x.stride(2) % 8 == 0 && x.stride(1) % 8 == 0 && x.stride(0) % 8 == 0 && x.stride(2) > 0 &&
                  reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0,
              "strided NHWC input: channel stride must be 1, the other strides multiples of 8 elements, the base 16-byte aligned");
  st[0] = x.stride(2); st[1] = x.stride(1); st[2] = x.stride(0);
}

torch::Tensor conv_nhwc(const torch::Tensor& x, const torch::Tensor& w, const c10::optional<torch::Tensor>& bias, int64_t stride,
                        int64_t pad, int64_t dil, bool relu, bool transposed, int64_t out_h, int64_t out_w, int64_t pad_w,
                        const c10::optional<torch::Tensor>& col_stats) {
  if (pad_w < 0) pad_w = pad;          // square padding unless a horizontal one is given (rectangular kernels)
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(w.is_contiguous() && x.dim() == 4 && w.dim() == 4 && x.scalar_type() == torch::kBFloat16 &&
              w.scalar_type() == torch::kBFloat16, "conv_nhwc expects bf16 NHWC activations and contiguous [Cout,R,S,Cin] weights");
  long long xs[3];
  nhwc_strides(x, xs);
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3);
  const int R = w.size(1), S = w.size(2);
  const int k_ch = transposed ? w.size(0) : w.size(3);      // channels reduced over
  const int n_ch = transposed ? w.size(3) : w.size(0);      // channels produced
  TORCH_CHECK(C == k_ch && k_ch % 8 == 0 && n_ch % 8 == 0, "channel counts must match and be multiples of 8");
  int Ho, Wo;
  if (transposed) {
    TORCH_CHECK(stride == 1, "native dgrad handles stride 1");
    Ho = out_h; Wo = out_w;
  } else {
    Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
    Wo = (W + 2 * pad_w - dil * (S - 1) - 1) / stride + 1;
  }
  auto y = torch::empty({N, Ho, Wo, n_ch}, x.options());
  const void* bias_ptr = (bias.has_value() && bias->defined()) ? bias->data_ptr() : nullptr;
  const int eff_pad = transposed ? static_cast<int>(dil * (R - 1) - pad) : static_cast<int>(pad);
  const int eff_pad_w = transposed ? static_cast<int>(dil * (S - 1) - pad_w) : static_cast<int>(pad_w);
  float* stats_ptr = nullptr;          // per-channel sum / sum of squares of y, ADDED into a [2, Cout] fp32 tensor (BatchNorm statistics)
  if (col_stats.has_value() && col_stats->defined()) {
    TORCH_CHECK(col_stats->scalar_type() == torch::kFloat32 && col_stats->is_contiguous() && col_stats->numel() == 2 * static_cast<int64_t>(n_ch) &&
                    col_stats->device() == x.device() && !relu, "col_stats must be a contiguous fp32 [2, Cout] tensor (no fused ReLU)");
    stats_ptr = col_stats->data_ptr<float>();
  }
  check(rvn_conv_forward_strided(x.data_ptr(), w.data_ptr(), y.data_ptr(), bias_ptr, relu, N, H, W, k_ch, Ho, Wo, n_ch, R, S,
                                 static_cast<int>(stride), eff_pad, eff_pad_w, static_cast<int>(dil), transposed, xs[0], xs[1], xs[2], stats_ptr,
                                 sm_count(), cur_stream()),
        "conv_forward");
  return y;
}

// data gradient of a stride-2 convolution: dy [N,Ho,Wo,Cout], w [Cout,R,S,Cin] -> dx [N,H,W,Cin] (four parity phases, conv_sm100.cu)
torch::Tensor conv_dgrad_s2(const torch::Tensor& dy, const torch::Tensor& w, int64_t H, int64_t W, int64_t pad, int64_t pad_w) {
  if (pad_w < 0) pad_w = pad;
  c10::cuda::CUDAGuard guard(dy.device());
  TORCH_CHECK(dy.is_contiguous() && w.is_contiguous() && dy.dim() == 4 && w.dim() == 4 && dy.scalar_type() == torch::kBFloat16 &&
              w.scalar_type() == torch::kBFloat16 && dy.size(3) == w.size(0), "conv_dgrad_s2 expects contiguous bf16 dy [N,Ho,Wo,Cout], w [Cout,R,S,Cin]");
  const int N = dy.size(0), Ho = dy.size(1), Wo = dy.size(2), Cout = dy.size(3);
  const int R = w.size(1), S = w.size(2), Cin = w.size(3);
  TORCH_CHECK(Cin % 8 == 0 && Cout % 8 == 0, "channel counts must be multiples of 8");
  TORCH_CHECK((H + 2 * pad - R) / 2 + 1 == Ho && (W + 2 * pad_w - S) / 2 + 1 == Wo, "dy does not match a stride-2 convolution of an H x W input");
  auto dx = (R == 1 || S == 1) ? torch::zeros({N, H, W, Cin}, dy.options()) : torch::empty({N, H, W, Cin}, dy.options());
  check(rvn_conv_dgrad_s2(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), N, Ho, Wo, Cout, static_cast<int>(H), static_cast<int>(W), Cin, R, S,
                          static_cast<int>(pad), static_cast<int>(pad_w), sm_count(), cur_stream()),
        "conv_dgrad_s2");
  return dx;
}

// dy [N,Ho,Wo,Cout], x [N,H,W,Cin] (NHWC bf16, x possibly a strided view) -> dW [Cout,R,S,Cin] fp32.  With `acc` (a contiguous fp32
// tensor of that many elements, e.g. the stage's persistent weight-gradient accumulator) the gradient is ADDED into it and returned
torch::Tensor conv_wgrad_nhwc(const torch::Tensor& dy, const torch::Tensor& x, int64_t R, int64_t S, int64_t stride, int64_t pad, int64_t dil,
                              int64_t pad_w, const c10::optional<torch::Tensor>& acc) {
  if (pad_w < 0) pad_w = pad;
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(dy.is_contiguous() && x.dim() == 4 && dy.scalar_type() == torch::kBFloat16 && x.scalar_type() == torch::kBFloat16,
              "conv_wgrad_nhwc expects bf16 NHWC tensors (dy contiguous)");
  long long xs[3];
  nhwc_strides(x, xs);
  const int N = x.size(0), H = x.size(1), W = x.size(2), Cin = x.size(3);
  const int Ho = dy.size(1), Wo = dy.size(2), Cout = dy.size(3);
  TORCH_CHECK(Cin % 8 == 0 && Cout % 8 == 0, "channel counts must be multiples of 8");
  torch::Tensor dw;
  if (acc.has_value() && acc->defined()) {
    TORCH_CHECK(acc->is_contiguous() && acc->scalar_type() == torch::kFloat32 && acc->numel() == static_cast<int64_t>(Cout) * R * S * Cin &&
                    acc->device() == x.device() && reinterpret_cast<uintptr_t>(acc->data_ptr()) % 16 == 0,
                "acc must be a contiguous, 16-byte aligned fp32 tensor of Cout*R*S*Cin elements");
    dw = *acc;
  } else {
    dw = torch::zeros({Cout, R, S, Cin}, x.options().dtype(torch::kFloat32));
  }
  check(rvn_conv_wgrad_strided(dy.data_ptr(), x.data_ptr(), dw.data_ptr<float>(), N, H, W, Cin, Ho, Wo, Cout, static_cast<int>(R),
                               static_cast<int>(S), static_cast<int>(stride), static_cast<int>(pad), static_cast<int>(pad_w),
                               static_cast<int>(dil), xs[0], xs[1], xs[2], sm_count(), cur_stream()),
        "conv_wgrad");
  return dw;
}

// [N, C <= 8, H, W]-shaped images (any strides, bf16 or fp32) -> zero-bordered dense [N, Hp, Wp, 8] bf16 (conv_aux.cu)
torch::Tensor stem_pack(const torch::Tensor& x, int64_t pad_h, int64_t pad_w, int64_t Hp, int64_t Wp) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_cuda() && x.dim() == 4 && x.size(1) >= 1 && x.size(1) <= 8 &&
                  (x.scalar_type() == torch::kBFloat16 || x.scalar_type() == torch::kFloat32),
              "stem_pack expects an [N, C <= 8, H, W] bf16 / fp32 CUDA tensor");
  TORCH_CHECK(Hp >= x.size(2) + pad_h && Wp >= x.size(3) + pad_w && pad_h >= 0 && pad_w >= 0, "padded extent smaller than the image");
  auto out = torch::empty({x.size(0), Hp, Wp, 8}, x.options().dtype(torch::kBFloat16));
  check(rvn_stem_pack(x.data_ptr(), x.scalar_type() == torch::kBFloat16, out.data_ptr(), x.size(0), x.size(1), x.size(2), x.size(3),
                      static_cast<int>(Hp), static_cast<int>(Wp), static_cast<int>(pad_h), static_cast<int>(pad_w), x.stride(0), x.stride(1),
                      x.stride(2), x.stride(3), sm_count(), cur_stream()),
        "stem_pack");
  return out;
}

// grad += acc; acc = 0   (fp32 convolution-weight-gradient accumulator -> flat gradient bucket, conv_aux.cu)
void wgrad_flush(torch::Tensor acc, torch::Tensor grad) {
  c10::cuda::CUDAGuard guard(acc.device());
  TORCH_CHECK(acc.is_cuda() && acc.is_contiguous() && grad.is_contiguous() && acc.scalar_type() == torch::kFloat32 &&
                  acc.numel() == grad.numel() && acc.numel() % 8 == 0 &&
                  (grad.scalar_type() == torch::kBFloat16 || grad.scalar_type() == torch::kFloat32) &&
                  reinterpret_cast<uintptr_t>(acc.data_ptr()) % 32 == 0 && reinterpret_cast<uintptr_t>(grad.data_ptr()) % 16 == 0,
              "wgrad_flush expects a contiguous fp32 accumulator and a bf16 / fp32 gradient bucket of the same (multiple of 8) length");
  check(rvn_wgrad_flush(acc.data_ptr<float>(), grad.data_ptr(), grad.scalar_type() == torch::kBFloat16, static_cast<size_t>(acc.numel()),
                        sm_count(), cur_stream()),
        "wgrad_flush");
}

// x: [..., C] bf16 contiguous (NHWC flattened); gamma / beta in the parameter dtype (bf16 or fp32), mean / invstd fp32 [C].
// training: batch statistics are computed (mean, invstd returned) and, when given, running_mean / running_var / num_batches_tracked
// are updated by the finalize kernel; eval: the given mean / invstd are applied.  Returns {y, mean, invstd}
static int is_bf16_or_f32(const torch::Tensor& t, const char* what) {
  TORCH_CHECK(t.scalar_type() == torch::kBFloat16 || t.scalar_type() == torch::kFloat32, what, " must be bf16 or fp32");
  TORCH_CHECK(t.is_contiguous(), what, " must be contiguous");
  return t.scalar_type() == torch::kBFloat16;
}

std::vector<torch::Tensor> batchnorm_fwd(const torch::Tensor& x, const c10::optional<torch::Tensor>& res, const torch::Tensor& gamma,
                                         const torch::Tensor& beta, const c10::optional<torch::Tensor>& mean_in,
                                         const c10::optional<torch::Tensor>& invstd_in, double eps, bool relu, bool training,
                                         const c10::optional<torch::Tensor>& running_mean, const c10::optional<torch::Tensor>& running_var,
                                         const c10::optional<torch::Tensor>& batches_tracked, double momentum, bool want_mask,
                                         const c10::optional<torch::Tensor>& sums_in) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.scalar_type() == torch::kBFloat16, "batchnorm expects contiguous bf16 NHWC input");
  const int pbf = is_bf16_or_f32(gamma, "gamma");
  TORCH_CHECK(beta.scalar_type() == gamma.scalar_type() && beta.is_contiguous(), "beta must match gamma");
  const int C = x.size(-1);
  const long long M = x.numel() / C;
  auto f32 = x.options().dtype(torch::kFloat32);
  auto y = torch::empty_like(x);
  torch::Tensor mean = training ? torch::empty({C}, f32) : mean_in.value();
  torch::Tensor invstd = training ? torch::empty({C}, f32) : invstd_in.value();
  // sums_in: the per-channel sum / sum of squares of x, already collected by the epilogue of the kernel that produced x
  const bool sums_ready = training && sums_in.has_value() && sums_in->defined();
  if (sums_ready)
    TORCH_CHECK(sums_in->scalar_type() == torch::kFloat32 && sums_in->is_contiguous() && sums_in->numel() == 2 * static_cast<int64_t>(C) &&
                    sums_in->device() == x.device(), "sums_in must be a contiguous fp32 [2, C] tensor");
  auto sums = sums_ready ? *sums_in : torch::empty({2 * static_cast<int64_t>(C)}, f32);
  const void* res_ptr = (res.has_value() && res->defined()) ? res->data_ptr() : nullptr;
  if (res_ptr) TORCH_CHECK(res->is_contiguous() && res->numel() == x.numel(), "residual must match x");
  void *rm = nullptr, *rv = nullptr;
  long long* nbt = nullptr;
  int sbf = 0;
  if (training && running_mean.has_value() && running_mean->defined()) {
    sbf = is_bf16_or_f32(*running_mean, "running_mean");
    TORCH_CHECK(running_var.has_value() && running_var->scalar_type() == running_mean->scalar_type() && running_var->is_contiguous(),
                "running_var must match running_mean");
    rm = running_mean->data_ptr();
    rv = running_var->data_ptr();
    if (batches_tracked.has_value() && batches_tracked->defined()) {
      TORCH_CHECK(batches_tracked->scalar_type() == torch::kInt64, "num_batches_tracked must be int64");
      nbt = reinterpret_cast<long long*>(batches_tracked->data_ptr());
    }
  }
  torch::Tensor mask;                        // ReLU sign bits (one per element) for the backward pass: it then never reads y
  if (want_mask && relu) mask = torch::empty({static_cast<int64_t>(x.numel() / 8)}, x.options().dtype(torch::kUInt8));
  check(rvn_batchnorm_forward(x.data_ptr(), res_ptr, gamma.data_ptr(), beta.data_ptr(), rm, rv, nbt, mean.data_ptr<float>(),
                              invstd.data_ptr<float>(), sums.data_ptr<float>(), y.data_ptr(), M, C, static_cast<float>(eps),
                              static_cast<float>(momentum), relu, training, pbf, sbf, mask.defined() ? mask.data_ptr() : nullptr,
                              sums_ready ? 1 : 0, sm_count(), cur_stream()),
        "batchnorm_fwd");
  if (mask.defined()) return {y, mean, invstd, mask};
  return {y, mean, invstd};
}

// returns {dx, dres (dz, defined only when want_dres), dgamma, dbeta} — dgamma / dbeta in gamma's dtype; when grad_gamma / grad_beta
// are given the parameter gradients are accumulated into them in place and undefined tensors are returned instead
std::vector<torch::Tensor> batchnorm_bwd(const torch::Tensor& dy, const c10::optional<torch::Tensor>& y, const torch::Tensor& x,
                                         const torch::Tensor& gamma, const torch::Tensor& mean, const torch::Tensor& invstd, bool relu,
                                         bool want_dres, const c10::optional<torch::Tensor>& grad_gamma,
                                         const c10::optional<torch::Tensor>& grad_beta, const c10::optional<torch::Tensor>& relu_mask) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(dy.is_contiguous() && x.is_contiguous(), "batchnorm_bwd expects contiguous tensors");
  const int pbf = is_bf16_or_f32(gamma, "gamma");
  const int C = x.size(-1);
  const long long M = x.numel() / C;
  auto sums = torch::empty({2 * static_cast<int64_t>(C)}, x.options().dtype(torch::kFloat32));
  auto dx = torch::empty_like(x);
  torch::Tensor dres, dgamma, dbeta;
  if (want_dres) dres = torch::empty_like(x);
  const bool acc = grad_gamma.has_value() && grad_gamma->defined() && grad_beta.has_value() && grad_beta->defined();
  void *dg_ptr, *db_ptr;
  if (acc) {
    TORCH_CHECK(grad_gamma->scalar_type() == gamma.scalar_type() && grad_beta->scalar_type() == gamma.scalar_type() &&
                    grad_gamma->is_contiguous() && grad_beta->is_contiguous() && grad_gamma->numel() == C && grad_beta->numel() == C,
                "gradient accumulators must match gamma");
    dg_ptr = grad_gamma->data_ptr();
    db_ptr = grad_beta->data_ptr();
  } else {
    dgamma = torch::empty_like(gamma);
    dbeta = torch::empty_like(gamma);
    dg_ptr = dgamma.data_ptr();
    db_ptr = dbeta.data_ptr();
  }
  const void* mask_ptr = nullptr;
  if (relu && relu_mask.has_value() && relu_mask->defined()) {
    TORCH_CHECK(relu_mask->scalar_type() == torch::kUInt8 && relu_mask->numel() == x.numel() / 8, "bad relu mask");
    mask_ptr = relu_mask->data_ptr();
  }
  const void* y_ptr = (relu && mask_ptr == nullptr) ? y.value().data_ptr() : nullptr;
  check(rvn_batchnorm_backward(dy.data_ptr(), y_ptr, x.data_ptr(), gamma.data_ptr(), mean.data_ptr<float>(), invstd.data_ptr<float>(),
                               sums.data_ptr<float>(), dx.data_ptr(), want_dres ? dres.data_ptr() : nullptr, dg_ptr, db_ptr, M, C, relu, pbf,
                               acc ? 1 : 0, mask_ptr, sm_count(), cur_stream()),
        "batchnorm_bwd");
  return {dx, dres, dgamma, dbeta};
}

// x [N,H,W,C] bf16 NHWC -> {y [N,Ho,Wo,C], argmax (uint8, max pooling only)}
std::vector<torch::Tensor> pool_fwd(const torch::Tensor& x, int64_t k, int64_t stride, int64_t pad, bool is_max) {
  c10::cuda::CUDAGuard guard(x.device());
  TORCH_CHECK(x.is_contiguous() && x.dim() == 4 && x.scalar_type() == torch::kBFloat16 && x.size(3) % 8 == 0, "pool expects NHWC bf16, C % 8 == 0");
  const int N = x.size(0), H = x.size(1), W = x.size(2), C = x.size(3);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  auto y = torch::empty({N, Ho, Wo, C}, x.options());
  torch::Tensor arg = is_max ? torch::empty({N, Ho, Wo, C}, x.options().dtype(torch::kUInt8)) : torch::Tensor();
  check(rvn_pool_forward(x.data_ptr(), y.data_ptr(), is_max ? arg.data_ptr() : nullptr, N, H, W, C, Ho, Wo, static_cast<int>(k),
                         static_cast<int>(stride), static_cast<int>(pad), is_max, sm_count(), cur_stream()),
        "pool_fwd");
  if (is_max) return {y, arg};
  return {y};
}

torch::Tensor pool_bwd(const torch::Tensor& dy, const c10::optional<torch::Tensor>& arg, int64_t H, int64_t W, int64_t k, int64_t stride,
                       int64_t pad, bool is_max) {
  c10::cuda::CUDAGuard guard(dy.device());
  TORCH_CHECK(dy.is_contiguous() && dy.dim() == 4, "pool_bwd expects contiguous NHWC dy");
  const int N = dy.size(0), Ho = dy.size(1), Wo = dy.size(2), C = dy.size(3);
  auto dx = torch::empty({N, H, W, C}, dy.options());
  check(rvn_pool_backward(dy.data_ptr(), is_max ? arg.value().data_ptr() : nullptr, dx.data_ptr(), N, static_cast<int>(H), static_cast<int>(W), C,
                          Ho, Wo, static_cast<int>(k), static_cast<int>(stride), static_cast<int>(pad), is_max, sm_count(), cur_stream()),
        "pool_bwd");
  return dx;
}

// ---------------------------------------------------------------------------------------------- fused optimizers
PeerPtrs make_peers(const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& signal_ptrs, int64_t rank) {
  PeerPtrs pp{};
  TORCH_CHECK(grad_ptrs.size() >= 1 && grad_ptrs.size() <= 8, "1..8 DP peers supported");
  pp.n = static_cast<int>(grad_ptrs.size());
  pp.rank = static_cast<int>(rank);
  for (int i = 0; i < pp.n; ++i) {
    pp.grad[i] = reinterpret_cast<const void*>(grad_ptrs[i]);
    pp.signal[i] = i < static_cast<int>(signal_ptrs.size()) ? reinterpret_cast<uint32_t*>(signal_ptrs[i]) : nullptr;
  }
  return pp;
}

OptArgs make_args(const std::vector<double>& f, const std::vector<int64_t>& flags) {
  OptArgs a{};
  a.lr = f[0]; a.beta1 = f[1]; a.beta2 = f[2]; a.eps = f[3]; a.weight_decay = f[4]; a.momentum = f[5]; a.dampening = f[6];
  a.grad_scale = f[7]; a.bias_c1 = f[8]; a.bias_c2 = f[9]; a.clamp_value = f[10];
  a.nesterov = flags[0]; a.first_step = flags[1]; a.decoupled_wd = flags[2]; a.lamb_adam_mode = flags[3];
  return a;
}

void fused_update(int64_t kind, const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& signal_ptrs, int64_t rank,
                  const std::vector<double>& f, const std::vector<int64_t>& flags, torch::Tensor master, torch::Tensor m,
                  torch::Tensor v, torch::Tensor w_out, bool grad_is_bf16, int64_t epoch, torch::Tensor done_counter) {
  c10::cuda::CUDAGuard guard(master.device());
  PeerPtrs pp = make_peers(grad_ptrs, signal_ptrs, rank);
  OptArgs a = make_args(f, flags);
  check(rvn_fused_update(&pp, &a, static_cast<int>(kind), master.data_ptr<float>(), m.defined() && m.numel() ? m.data_ptr<float>() : nullptr,
                         v.defined() && v.numel() ? v.data_ptr<float>() : nullptr, w_out.data_ptr(), grad_is_bf16,
                         w_out.scalar_type() == torch::kBFloat16, static_cast<size_t>(master.numel()), static_cast<uint32_t>(epoch),
                         reinterpret_cast<uint32_t*>(done_counter.data_ptr()), sm_count(), cur_stream()),
        "fused_update");
}

void fused_lamb(const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& signal_ptrs, int64_t rank, const std::vector<double>& f,
                const std::vector<int64_t>& flags, torch::Tensor master, torch::Tensor m, torch::Tensor v, torch::Tensor w_out,
                bool grad_is_bf16, torch::Tensor chunk_param, torch::Tensor chunk_start, torch::Tensor chunk_end, torch::Tensor norms,
                int64_t epoch, torch::Tensor done_counter) {
  c10::cuda::CUDAGuard guard(master.device());
  PeerPtrs pp = make_peers(grad_ptrs, signal_ptrs, rank);
  OptArgs a = make_args(f, flags);
  check(rvn_fused_lamb(&pp, &a, master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), w_out.data_ptr(), grad_is_bf16,
                       w_out.scalar_type() == torch::kBFloat16, chunk_param.data_ptr<int>(),
                       reinterpret_cast<const long long*>(chunk_start.data_ptr<int64_t>()),
                       reinterpret_cast<const long long*>(chunk_end.data_ptr<int64_t>()), static_cast<int>(chunk_param.numel()),
                       norms.data_ptr<float>(), static_cast<int>(norms.numel() / 2), static_cast<uint32_t>(epoch),
                       reinterpret_cast<uint32_t*>(done_counter.data_ptr()), cur_stream()),
        "fused_lamb");
}

// K1 v2 (two-shot): this rank reduces + updates only [shard_begin, shard_end) and pushes the new weights into every rank's landing buffer
void fused_twoshot_update(int64_t kind, const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& signal_ptrs,
                          const std::vector<int64_t>& landing_ptrs, int64_t rank, const std::vector<double>& f,
                          const std::vector<int64_t>& flags, torch::Tensor master, torch::Tensor m, torch::Tensor v, bool grad_is_bf16,
                          bool out_is_bf16, int64_t shard_begin, int64_t shard_end, int64_t epoch, torch::Tensor done_counter) {
  c10::cuda::CUDAGuard guard(master.device());
  PeerPtrs pp = make_peers(grad_ptrs, signal_ptrs, rank);
  TORCH_CHECK(landing_ptrs.size() == grad_ptrs.size() && landing_ptrs.size() <= 8, "one landing buffer per DP rank (<= 8)");
  WeightPeers wp{};
  for (size_t i = 0; i < landing_ptrs.size(); ++i) wp.w[i] = reinterpret_cast<void*>(landing_ptrs[i]);
  OptArgs a = make_args(f, flags);
  check(rvn_fused_twoshot_update(&pp, &wp, &a, static_cast<int>(kind), master.data_ptr<float>(),
                                 m.defined() && m.numel() ? m.data_ptr<float>() : nullptr, v.defined() && v.numel() ? v.data_ptr<float>() : nullptr,
                                 grad_is_bf16, out_is_bf16, static_cast<size_t>(shard_begin), static_cast<size_t>(shard_end),
                                 static_cast<uint32_t>(epoch), reinterpret_cast<uint32_t*>(done_counter.data_ptr()), sm_count(), cur_stream()),
        "fused_twoshot_update");
}

void fused_twoshot_lamb(const std::vector<int64_t>& grad_ptrs, const std::vector<int64_t>& signal_ptrs,
                        const std::vector<int64_t>& landing_ptrs, const std::vector<int64_t>& norm_ptrs, int64_t rank,
                        const std::vector<double>& f, const std::vector<int64_t>& flags, torch::Tensor master, torch::Tensor m,
                        torch::Tensor v, bool grad_is_bf16, bool out_is_bf16, torch::Tensor chunk_param, torch::Tensor chunk_start,
                        torch::Tensor chunk_end, int64_t first_chunk, int64_t n_chunks, torch::Tensor norms, int64_t epoch,
                        torch::Tensor done_counter) {
  c10::cuda::CUDAGuard guard(master.device());
  PeerPtrs pp = make_peers(grad_ptrs, signal_ptrs, rank);
  TORCH_CHECK(landing_ptrs.size() == grad_ptrs.size() && norm_ptrs.size() == grad_ptrs.size() && grad_ptrs.size() <= 8, "bad peer lists");
  WeightPeers wp{};
  NormPeers np{};
  for (size_t i = 0; i < landing_ptrs.size(); ++i) {
    wp.w[i] = reinterpret_cast<void*>(landing_ptrs[i]);
    np.norms[i] = reinterpret_cast<const float*>(norm_ptrs[i]);
  }
  OptArgs a = make_args(f, flags);
  TORCH_CHECK(first_chunk >= 0 && first_chunk + n_chunks <= chunk_param.numel(), "chunk range out of bounds");
  check(rvn_fused_twoshot_lamb(&pp, &wp, &np, &a, master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), grad_is_bf16, out_is_bf16,
                               chunk_param.data_ptr<int>() + first_chunk,
                               reinterpret_cast<const long long*>(chunk_start.data_ptr<int64_t>()) + first_chunk,
                               reinterpret_cast<const long long*>(chunk_end.data_ptr<int64_t>()) + first_chunk, static_cast<int>(n_chunks),
                               norms.data_ptr<float>(), static_cast<int>(norms.numel() / 2), static_cast<uint32_t>(epoch),
                               reinterpret_cast<uint32_t*>(done_counter.data_ptr()), cur_stream()),
        "fused_twoshot_lamb");
}

void fused_param_average(const std::vector<int64_t>& src_ptrs, const std::vector<int64_t>& signal_ptrs, int64_t rank, torch::Tensor master_out,
                         torch::Tensor w_out, int64_t epoch, torch::Tensor done_counter) {
  c10::cuda::CUDAGuard guard(master_out.device());
  PeerPtrs pp = make_peers(src_ptrs, signal_ptrs, rank);
  check(rvn_fused_param_average(&pp, master_out.data_ptr<float>(), w_out.data_ptr(), w_out.scalar_type() == torch::kBFloat16,
                                static_cast<size_t>(master_out.numel()), static_cast<uint32_t>(epoch),
                                reinterpret_cast<uint32_t*>(done_counter.data_ptr()), sm_count(), cur_stream()),
        "fused_param_average");
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("gemm", &gemm, "bf16 tcgen05 GEMM with fused epilogue", py::arg("A"), py::arg("a_mn"), py::arg("B"), py::arg("b_mn"),
        py::arg("bias") = py::none(), py::arg("act") = 0, py::arg("aux_in") = py::none(), py::arg("store_preact") = false,
        py::arg("block_n") = 0, py::arg("aux_mode") = 1, py::arg("out") = py::none(), py::arg("b_copy_dst") = py::none(),
        py::arg("bcopy_flags") = py::none(), py::arg("bcopy_epoch") = 0, py::arg("split_k") = 0, py::arg("col_stats") = py::none());
  m.def("gemm_prefers_pair", &gemm_prefers_pair);
  m.def("gemm_2cta", &gemm_2cta, "bf16 tcgen05 GEMM on CTA pairs (cta_group::2)", py::arg("A"), py::arg("a_mn"), py::arg("B"), py::arg("b_mn"),
        py::arg("bias") = py::none(), py::arg("act") = 0, py::arg("block_n") = 256);
  m.def("layernorm_fwd", &layernorm_fwd);
  m.def("layernorm_bwd", &layernorm_bwd, py::arg("dy"), py::arg("x"), py::arg("gamma"), py::arg("mean"), py::arg("rstd"), py::arg("rng"),
        py::arg("dropout_p"), py::arg("grad_gamma") = py::none(), py::arg("grad_beta") = py::none());
  m.def("ce_fwd", &ce_fwd);
  m.def("ce_bwd", &ce_bwd);
  m.def("column_sum", &column_sum, py::arg("x"), py::arg("into") = py::none());
  m.def("act_bwd", &act_bwd);
  m.def("embedding_fwd", &embedding_fwd);
  m.def("embedding_bwd", &embedding_bwd, py::arg("idx"), py::arg("dy"), py::arg("num_rows"), py::arg("padding_idx") = -1, py::arg("into") = py::none());
  m.def("dropout_apply", &dropout_apply);
  m.def("pack_f32", &pack_f32);
  m.def("rng_next", &rng_next);
  m.def("conv_nhwc", &conv_nhwc, py::arg("x"), py::arg("w"), py::arg("bias"), py::arg("stride"), py::arg("pad"), py::arg("dil"), py::arg("relu"),
        py::arg("transposed"), py::arg("out_h"), py::arg("out_w"), py::arg("pad_w") = -1, py::arg("col_stats") = py::none());
  m.def("pool_fwd", &pool_fwd);
  m.def("pool_bwd", &pool_bwd);
  m.def("batchnorm_fwd", &batchnorm_fwd, py::arg("x"), py::arg("res"), py::arg("gamma"), py::arg("beta"), py::arg("mean_in"), py::arg("invstd_in"),
        py::arg("eps"), py::arg("relu"), py::arg("training"), py::arg("running_mean"), py::arg("running_var"), py::arg("batches_tracked"),
        py::arg("momentum"), py::arg("want_mask") = false, py::arg("sums_in") = py::none());
  m.def("batchnorm_bwd", &batchnorm_bwd, py::arg("dy"), py::arg("y"), py::arg("x"), py::arg("gamma"), py::arg("mean"), py::arg("invstd"),
        py::arg("relu"), py::arg("want_dres"), py::arg("grad_gamma") = py::none(), py::arg("grad_beta") = py::none(),
        py::arg("relu_mask") = py::none());
  m.def("conv_wgrad_nhwc", &conv_wgrad_nhwc, py::arg("dy"), py::arg("x"), py::arg("R"), py::arg("S"), py::arg("stride"), py::arg("pad"),
        py::arg("dil"), py::arg("pad_w") = -1, py::arg("acc") = py::none());
  m.def("conv_dgrad_s2", &conv_dgrad_s2, "data gradient of a stride-2 convolution (four parity phases of the implicit-GEMM kernel)");
  m.def("stem_pack", &stem_pack, "zero-bordered 8-channel NHWC bf16 copy of an image batch");
  m.def("wgrad_flush", &wgrad_flush, "grad += acc; acc = 0");
  m.def("fmha_fwd", &fmha_fwd);
  m.def("fmha_bwd", &fmha_bwd);
  m.def("fused_update", &fused_update);
  m.def("fused_lamb", &fused_lamb);
  m.def("fused_twoshot_update", &fused_twoshot_update);
  m.def("fused_twoshot_lamb", &fused_twoshot_lamb);
  m.def("fused_param_average", &fused_param_average);
}
