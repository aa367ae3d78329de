// Per-modality LayerNorm between the encoders and HybridFusion (src/train.py:170-171,267-268:
// `encoded[m] = layer_norms[m](encoders[m](features[m]))`, nn.LayerNorm(output_dim), eps 1e-5, biased variance).
//
// On the fused tensor-core path the forward normalisation is part of proj_kernel's input phase (proj_gemm.cu: the
// fp32 tile is already in shared memory there).  These stand-alone kernels serve every other path (fp32 parity
// mode, shapes outside proj_kernel) and the backward pass: one warp per row, the row held in registers, mean and
// centred sum of squares as two warp reductions (no E[x^2] - mean^2 cancellation).
//
//   forward : y = (x - mean) * rstd * gamma + beta
//   backward: xhat = (x - mean) * rstd;  g = dy * gamma
//             dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += dy * xhat;  dbeta += dy
// dgamma / dbeta are summed per block in shared memory and added to global memory with one atomic per column and
// block (the entry point clears them first).
#include "msf_common.cuh"

namespace msf {
namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_MAX_PER_LANE = 16;   // D <= 512

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(LN_THREADS) layer_norm_fwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    float* __restrict__ y, long long rows, int D,
                                                                    float eps) {
  const int lane = threadIdx.x & 31, wpb = LN_THREADS / 32;
  const float inv_d = 1.0f / (float)D;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const float* xr = x + row * D;
    float v[LN_MAX_PER_LANE];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < D ? __ldg(xr + c) : 0.0f;
      s += v[i];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      const float d = c < D ? v[i] - mean : 0.0f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    float* yr = y + row * D;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      if (c < D) yr[c] = (v[i] - mean) * rstd * (gamma ? __ldg(gamma + c) : 1.0f) + (beta ? __ldg(beta + c) : 0.0f);
    }
  }
}

__global__ void __launch_bounds__(LN_THREADS) layer_norm_bwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ dy,
                                                                    float* __restrict__ dx, float* __restrict__ dgamma,
                                                                    float* __restrict__ dbeta, long long rows, int D,
                                                                    float eps) {
  extern __shared__ float part[];   // [2][D]: dgamma, dbeta of this block
  for (int c = threadIdx.x; c < 2 * D; c += LN_THREADS) part[c] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = LN_THREADS / 32;
  const float inv_d = 1.0f / (float)D;
  float acc_g[LN_MAX_PER_LANE], acc_b[LN_MAX_PER_LANE];
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) acc_g[i] = acc_b[i] = 0.0f;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const float* xr = x + row * D;
    const float* gr = dy + row * D;
    float v[LN_MAX_PER_LANE], g[LN_MAX_PER_LANE];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < D ? __ldg(xr + c) : 0.0f;
      g[i] = c < D ? __ldg(gr + c) : 0.0f;
      s += v[i];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      const float d = c < D ? v[i] - mean : 0.0f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    float sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      const float xhat = c < D ? (v[i] - mean) * rstd : 0.0f;
      acc_g[i] += g[i] * xhat;
      acc_b[i] += g[i];
      const float gg = g[i] * ((gamma && c < D) ? __ldg(gamma + c) : 1.0f);
      v[i] = xhat;
      g[i] = gg;
      sg += gg;
      sgx += gg * xhat;
    }
    const float mg = warp_sum(sg) * inv_d, mgx = warp_sum(sgx) * inv_d;
    if (dx != nullptr) {
      float* dr = dx + row * D;
#pragma unroll
      for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < D) dr[c] = rstd * (g[i] - mg - v[i] * mgx);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      atomicAdd(&part[c], acc_g[i]);
      atomicAdd(&part[D + c], acc_b[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += LN_THREADS) {
    if (dgamma != nullptr) atomicAdd(dgamma + c, part[c]);
    if (dbeta != nullptr) atomicAdd(dbeta + c, part[D + c]);
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_layer_norm_forward(const float* x, const float* gamma, const float* beta, float* y, int64_t rows,
                                      int32_t dim, float eps, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && y && rows >= 0 && dim >= 1 && dim <= 32 * LN_MAX_PER_LANE, "msf_layer_norm_forward: bad arguments (dim %d)",
              (int)dim);
  if (rows == 0) return MSF_OK;
  long long blocks = ceil_div(rows, LN_THREADS / 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  layer_norm_fwd_kernel<<<(unsigned)blocks, LN_THREADS, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, rows, dim, eps);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

extern "C" int msf_layer_norm_backward(const float* x, const float* gamma, const float* dy, float* dx, float* dgamma,
                                       float* dbeta, int64_t rows, int32_t dim, float eps, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && dy && rows >= 0 && dim >= 1 && dim <= 32 * LN_MAX_PER_LANE,
              "msf_layer_norm_backward: bad arguments (dim %d)", (int)dim);
  cudaStream_t st = (cudaStream_t)stream;
  if (dgamma) MSF_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * dim, st));
  if (dbeta) MSF_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * dim, st));
  if (rows == 0) return MSF_OK;
  long long blocks = ceil_div(rows, LN_THREADS / 32);
  if (blocks > 148 * 2) blocks = 148 * 2;
  layer_norm_bwd_kernel<<<(unsigned)blocks, LN_THREADS, 2 * sizeof(float) * dim, st>>>(x, gamma, dy, dx, dgamma, dbeta, rows,
                                                                                     dim, eps);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
