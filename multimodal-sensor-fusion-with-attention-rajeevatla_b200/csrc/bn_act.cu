// BatchNorm1d -> ReLU -> Dropout behind a Linear layer: the hidden blocks of SimpleMLPEncoder
// (src/encoders.py:339-397: `Linear, BatchNorm1d, ReLU, Dropout` per layer, BatchNorm at :374-375) as three small
// HBM-bound kernels instead of four eager framework ops.
//
//   statistics (training): per column of the Linear output y (rows x cols) the sum and the sum of squares over the
//              batch, accumulated in fp64 (a thread owns a column, a block a slab of rows; one fp64 atomic pair
//              per column and block) -> mean, biased variance, invstd; running_mean / running_var follow
//              nn.BatchNorm1d (momentum, UNBIASED variance into running_var).  Eval mode uses the running stats.
//   forward:   out = drop(relu((y - mean) * invstd * gamma + beta)), Philox dropout keyed by (seed, row, col / 8)
//   backward:  g  = dout * [out != 0] * scale          (ReLU and dropout gate recovered from the saved output)
//              dgamma = sum_r g * xhat, dbeta = sum_r g
//              training: dy = gamma * invstd * (g - dbeta / rows - xhat * dgamma / rows)
//              eval:     dy = gamma * invstd * g
// Under data parallelism the statistics are per shard, as the reference's nn.BatchNorm1d (no SyncBatchNorm in
// src/train.py) computes them per process.
#include "msf_common.cuh"

namespace msf {
namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_ROWS_PER_BLOCK = 256;

// grid (ceil(cols / 32), ceil(rows / BN_ROWS_PER_BLOCK)); block 32 x 8: lane = column, warp strides the slab's rows
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ y, long long rows, int cols,
                                                              double* __restrict__ acc) {
  __shared__ double sh[2][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const long long r0 = (long long)blockIdx.y * BN_ROWS_PER_BLOCK;
  const long long r1 = min(rows, r0 + BN_ROWS_PER_BLOCK);
  double s = 0.0, q = 0.0;
  if (c < cols)
    for (long long r = r0 + warp; r < r1; r += 8) {
      const double v = (double)__ldg(y + r * cols + c);
      s += v;
      q += v * v;
    }
  sh[0][warp][lane] = s;
  sh[1][warp][lane] = q;
  __syncthreads();
  if (warp == 0 && c < cols) {
    for (int w = 1; w < 8; ++w) {
      s += sh[0][w][lane];
      q += sh[1][w][lane];
    }
    atomicAdd(acc + c, s);
    atomicAdd(acc + cols + c, q);
  }
}

// one thread per column: mean / invstd of the batch, running statistics moved on; clears the accumulators
__global__ void bn_finalize_kernel(double* __restrict__ acc, long long rows, int cols, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const double n = (double)rows;
  const double mean = acc[c] / n;
  double var = acc[cols + c] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  acc[c] = 0.0;
  acc[cols + c] = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var != nullptr) {
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     int cols, float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  save_mean[c] = running_mean[c];
  save_invstd[c] = rsqrtf(running_var[c] + eps);
}

// 8 consecutive columns per thread (cols % 8 == 0) or one column per thread
template <bool VEC8>
__global__ void __launch_bounds__(BN_THREADS) bn_act_fwd_kernel(const float* __restrict__ y, float* __restrict__ out,
                                                                long long rows, int cols,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, int relu, DropCfg drop,
                                                                int sub) {
  const long long per_row = VEC8 ? cols >> 3 : cols;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int g = (int)(i - r * per_row);
    if (VEC8) {
      const int c0 = g * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(y + r * cols + c0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(y + r * cols + c0) + 1);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (drop.active) drop8(drop, SITE_CLS, sub, r, g, dm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        float z = (v[j] - __ldg(mean + c)) * __ldg(invstd + c) * (gamma ? __ldg(gamma + c) : 1.0f) + (beta ? __ldg(beta + c) : 0.0f);
        if (relu) z = fmaxf(z, 0.0f);
        v[j] = z * dm[j];
      }
      float4* o = reinterpret_cast<float4*>(out + r * cols + c0);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      const int c = g;
      float z = (__ldg(y + r * cols + c) - __ldg(mean + c)) * __ldg(invstd + c) * (gamma ? __ldg(gamma + c) : 1.0f) +
                (beta ? __ldg(beta + c) : 0.0f);
      if (relu) z = fmaxf(z, 0.0f);
      out[r * cols + c] = z * drop1(drop, SITE_CLS, sub, r, c);
    }
  }
}

// column sums of g and g * xhat: same decomposition as bn_stats_kernel
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const float* __restrict__ dout,
                                                                   const float* __restrict__ y,
                                                                   const float* __restrict__ out, long long rows, int cols,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd, int gated, float scale,
                                                                   double* __restrict__ acc) {
  __shared__ double sh[2][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const long long r0 = (long long)blockIdx.y * BN_ROWS_PER_BLOCK;
  const long long r1 = min(rows, r0 + BN_ROWS_PER_BLOCK);
  double s = 0.0, q = 0.0;
  if (c < cols) {
    const float mu = __ldg(mean + c), is = __ldg(invstd + c);
    for (long long r = r0 + warp; r < r1; r += 8) {
      const long long e = r * cols + c;
      float g = __ldg(dout + e);
      if (gated) g = (__ldg(out + e) != 0.0f) ? g * scale : 0.0f;
      const float xhat = (__ldg(y + e) - mu) * is;
      s += (double)g;
      q += (double)g * (double)xhat;
    }
  }
  sh[0][warp][lane] = s;
  sh[1][warp][lane] = q;
  __syncthreads();
  if (warp == 0 && c < cols) {
    for (int w = 1; w < 8; ++w) {
      s += sh[0][w][lane];
      q += sh[1][w][lane];
    }
    atomicAdd(acc + c, s);
    atomicAdd(acc + cols + c, q);
  }
}

__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float* __restrict__ dout,
                                                                  const float* __restrict__ y,
                                                                  const float* __restrict__ out, long long rows, int cols,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, int gated, float scale,
                                                                  int training, const double* __restrict__ acc,
                                                                  float* __restrict__ dy) {
  const long long total = rows * cols;
  const double inv_n = 1.0 / (double)rows;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cols);
    float g = __ldg(dout + e);
    if (gated) g = (__ldg(out + e) != 0.0f) ? g * scale : 0.0f;
    const float is = __ldg(invstd + c), ga = gamma ? __ldg(gamma + c) : 1.0f;
    float r = g;
    if (training) {
      const float xhat = (__ldg(y + e) - __ldg(mean + c)) * is;
      r = g - (float)(acc[c] * inv_n) - xhat * (float)(acc[cols + c] * inv_n);
    }
    dy[e] = ga * is * r;
  }
}

__global__ void bn_bwd_params_kernel(double* __restrict__ acc, int cols, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  if (dbeta != nullptr) dbeta[c] = (float)acc[c];
  if (dgamma != nullptr) dgamma[c] = (float)acc[cols + c];
  acc[c] = 0.0;
  acc[cols + c] = 0.0;
}

int grid_for(long long work) {
  const long long blocks = ceil_div(work, BN_THREADS);
  return (int)(blocks < 1 ? 1 : blocks > 148 * 8 ? 148 * 8 : blocks);
}

}  // namespace
}  // namespace msf

extern "C" {

int msf_bn_act_forward(const float* y, float* out, int64_t rows, int32_t cols, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int32_t training,
                       int32_t relu, float dropout_p, uint64_t seed, float* save_mean, float* save_invstd,
                       double* scratch, void* stream) {
  using namespace msf;
  MSF_REQUIRE(y && out && save_mean && save_invstd && scratch && rows >= 1 && cols >= 1, "msf_bn_act_forward: bad arguments");
  MSF_REQUIRE(dropout_p >= 0.0f && dropout_p < 1.0f, "msf_bn_act_forward: dropout_p must be in [0, 1)");
  MSF_REQUIRE(training || (running_mean && running_var), "msf_bn_act_forward: eval mode needs the running statistics");
  MSF_REQUIRE(!training || rows > 1, "Expected more than 1 value per channel when training, got input size [%lld, %d]",
              (long long)rows, cols);
  cudaStream_t st = (cudaStream_t)stream;
  if (training) {
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, BN_ROWS_PER_BLOCK));
    bn_stats_kernel<<<grid, BN_THREADS, 0, st>>>(y, rows, cols, scratch);
    MSF_LAUNCH_CHECK();
    bn_finalize_kernel<<<(unsigned)ceil_div(cols, 128), 128, 0, st>>>(scratch, rows, cols, eps, momentum, running_mean,
                                                                     running_var, save_mean, save_invstd);
    MSF_LAUNCH_CHECK();
  } else {
    bn_eval_stats_kernel<<<(unsigned)ceil_div(cols, 128), 128, 0, st>>>(running_mean, running_var, cols, eps, save_mean,
                                                                       save_invstd);
    MSF_LAUNCH_CHECK();
  }
  DropCfg drop;
  memset(&drop, 0, sizeof(drop));
  drop.seed = seed;
  drop.p = dropout_p;
  drop.scale = 1.0f / (1.0f - dropout_p);
  drop.active = (training && dropout_p > 0.0f) ? 1 : 0;
  const bool vec = cols % 8 == 0 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec)
    bn_act_fwd_kernel<true><<<grid_for(rows * (cols / 8)), BN_THREADS, 0, st>>>(y, out, rows, cols, gamma, beta, save_mean,
                                                                                save_invstd, relu, drop, 0);
  else
    bn_act_fwd_kernel<false><<<grid_for(rows * (long long)cols), BN_THREADS, 0, st>>>(y, out, rows, cols, gamma, beta,
                                                                                      save_mean, save_invstd, relu, drop, 0);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_bn_act_backward(const float* dout, const float* y, const float* out, int64_t rows, int32_t cols,
                        const float* gamma, const float* save_mean, const float* save_invstd, int32_t training,
                        int32_t relu, float dropout_p, float* dy, float* dgamma, float* dbeta, double* scratch,
                        void* stream) {
  using namespace msf;
  MSF_REQUIRE(dout && y && out && save_mean && save_invstd && dy && scratch && rows >= 1 && cols >= 1,
              "msf_bn_act_backward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int gated = (relu || (training && dropout_p > 0.0f)) ? 1 : 0;
  const float scale = (training && dropout_p > 0.0f) ? 1.0f / (1.0f - dropout_p) : 1.0f;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, BN_ROWS_PER_BLOCK));
  bn_bwd_reduce_kernel<<<grid, BN_THREADS, 0, st>>>(dout, y, out, rows, cols, save_mean, save_invstd, gated, scale, scratch);
  MSF_LAUNCH_CHECK();
  bn_bwd_apply_kernel<<<grid_for(rows * (long long)cols), BN_THREADS, 0, st>>>(dout, y, out, rows, cols, gamma, save_mean,
                                                                               save_invstd, gated, scale, training, scratch, dy);
  MSF_LAUNCH_CHECK();
  bn_bwd_params_kernel<<<(unsigned)ceil_div(cols, 128), 128, 0, st>>>(scratch, cols, dgamma, dbeta);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // extern "C"
