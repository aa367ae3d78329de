// Temporal pooling of FrameEncoder (src/encoders.py:258-336: attention / average / max pooling over the frames of a
// clip, optional frame mask) as ONE kernel per direction instead of a chain of elementwise ops:
//
//   attention:  s_t = x_t . w + b;  masked frames -> -inf;  p = softmax_t(s) (a clip with no valid frame: p = 0);
//               pooled = sum_t p_t x_t                                    (encoders.py:312-336)
//   average  :  pooled = sum_t m_t x_t / max(sum_t m_t, 1e-8)             (mean over t without a mask)
//   max      :  pooled_j = max_t x_tj over the valid frames (0 for a clip with no valid frame)
//
// One block per clip: pass 1 (attention) computes the T scores with one warp per frame (lanes along the hidden
// dimension: coalesced rows, shuffle reduction), the softmax runs over shared memory, pass 2 accumulates the pooled row
// with one thread per hidden column (coalesced across the block, the clip's frames come back from the L2).  HBM-bound:
// the clip is read once from DRAM (T*H*4 bytes), everything else is O(T + H).
//
// Backward (attention), with g_t = x_t . d pooled:  d s_t = p_t (g_t - sum_u p_u g_u),
//   d x_t = p_t d pooled + d s_t w,   d w = sum_{b,t} d s_t x_t,   d b = sum_{b,t} d s_t
// d w / d b are written as one partial row per clip and summed over the clips in fixed order by a second kernel.
#include <float.h>

#include "msf_common.cuh"

namespace msf {
namespace {

constexpr int FP_THREADS = 256;
constexpr int FP_MAX_FRAMES = 8192;   // scores of one clip live in shared memory

__device__ __forceinline__ float fp_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float fp_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reduction of one float per thread (sum or max); `red` holds FP_THREADS / 32 floats
template <bool MAX>
__device__ __forceinline__ float fp_block_reduce(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = MAX ? fp_warp_max(v) : fp_warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = MAX ? -FLT_MAX : 0.0f;
  for (int i = 0; i < FP_THREADS / 32; ++i) r = MAX ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// mode 0 attention, 1 average, 2 max.  weights_out (B, T): attention / average weights p_t (what the backward pass
// needs); argmax_out (B, H) int32 for mode 2.
__global__ void __launch_bounds__(FP_THREADS) frame_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                                    const float* __restrict__ w, const float* __restrict__ bias,
                                                                    int T, int H, int mode, float* __restrict__ pooled,
                                                                    float* __restrict__ weights_out, int* __restrict__ argmax_out) {
  extern __shared__ float sc[];   // T scores / weights, then 8 floats of reduction scratch
  float* red = sc + T;
  const long long b = blockIdx.x;
  const float* xb = x + b * (long long)T * H;
  const float* mb = mask != nullptr ? mask + b * (long long)T : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mode == 0) {
    const float b0 = bias != nullptr ? __ldg(bias) : 0.0f;
    for (int t = warp; t < T; t += FP_THREADS / 32) {   // one warp per frame
      float s = 0.0f;
      for (int j = lane; j < H; j += 32) s = fmaf(__ldg(xb + (long long)t * H + j), __ldg(w + j), s);
      s = fp_warp_sum(s) + b0;
      if (lane == 0) sc[t] = (mb != nullptr && mb[t] == 0.0f) ? -INFINITY : s;
    }
    __syncthreads();
    float mx = -FLT_MAX;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) mx = fmaxf(mx, sc[t]);
    mx = fp_block_reduce<true>(mx, red);
    float sum = 0.0f;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) {
      const float e = (sc[t] == -INFINITY) ? 0.0f : expf(sc[t] - mx);
      sc[t] = e;
      sum += e;
    }
    sum = fp_block_reduce<false>(sum, red);
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;   // no valid frame: softmax gives NaN, the reference maps it to 0
    for (int t = threadIdx.x; t < T; t += FP_THREADS) {
      sc[t] *= inv;
      weights_out[b * (long long)T + t] = sc[t];
    }
    __syncthreads();
  } else if (mode == 1) {
    float cnt = 0.0f;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) cnt += mb != nullptr ? mb[t] : 1.0f;
    cnt = fp_block_reduce<false>(cnt, red);
    const float inv = mb != nullptr ? 1.0f / fmaxf(cnt, 1e-8f) : 1.0f / (float)T;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) {
      sc[t] = (mb != nullptr ? mb[t] : 1.0f) * inv;
      weights_out[b * (long long)T + t] = sc[t];
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < H; j += FP_THREADS) {   // one thread per hidden column
    if (mode == 2) {
      float best = -INFINITY;
      int arg = -1;
      for (int t = 0; t < T; ++t) {
        if (mb != nullptr && mb[t] == 0.0f) continue;
        const float v = __ldg(xb + (long long)t * H + j);
        if (v > best || arg < 0) { best = v; arg = t; }   // first maximum, like torch.max
      }
      pooled[b * H + j] = arg < 0 ? 0.0f : best;
      argmax_out[b * H + j] = arg;
    } else {
      float acc = 0.0f;
      for (int t = 0; t < T; ++t) acc = fmaf(sc[t], __ldg(xb + (long long)t * H + j), acc);
      pooled[b * H + j] = acc;
    }
  }
}

__global__ void __launch_bounds__(FP_THREADS) frame_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ weights, const int* __restrict__ argmax,
                                                                    const float* __restrict__ d_pooled, int T, int H, int mode,
                                                                    float* __restrict__ dx, float* __restrict__ dw_partial,
                                                                    float* __restrict__ db_partial) {
  extern __shared__ float sc[];   // T values d s_t (attention), then reduction scratch
  float* red = sc + T;
  const long long b = blockIdx.x;
  const float* xb = x + b * (long long)T * H;
  const float* pb = weights != nullptr ? weights + b * (long long)T : nullptr;
  const float* dp = d_pooled + b * (long long)H;
  float* dxb = dx + b * (long long)T * H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mode == 0) {
    for (int t = warp; t < T; t += FP_THREADS / 32) {   // g_t = x_t . d pooled
      float g = 0.0f;
      for (int j = lane; j < H; j += 32) g = fmaf(__ldg(xb + (long long)t * H + j), __ldg(dp + j), g);
      g = fp_warp_sum(g);
      if (lane == 0) sc[t] = g;
    }
    __syncthreads();
    float dot = 0.0f;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) dot = fmaf(pb[t], sc[t], dot);
    dot = fp_block_reduce<false>(dot, red);
    float dsum = 0.0f;
    for (int t = threadIdx.x; t < T; t += FP_THREADS) {
      const float ds = pb[t] * (sc[t] - dot);
      sc[t] = ds;
      dsum += ds;
    }
    dsum = fp_block_reduce<false>(dsum, red);
    if (threadIdx.x == 0) db_partial[b] = dsum;
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += FP_THREADS) {
      const float dpj = __ldg(dp + j), wj = __ldg(w + j);
      float dwj = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float xv = __ldg(xb + (long long)t * H + j);
        dxb[(long long)t * H + j] = fmaf(pb[t], dpj, sc[t] * wj);
        dwj = fmaf(sc[t], xv, dwj);
      }
      dw_partial[b * H + j] = dwj;
    }
  } else if (mode == 1) {
    for (int j = threadIdx.x; j < H; j += FP_THREADS) {
      const float dpj = __ldg(dp + j);
      for (int t = 0; t < T; ++t) dxb[(long long)t * H + j] = pb[t] * dpj;
    }
  } else {
    for (int j = threadIdx.x; j < H; j += FP_THREADS) {
      const int arg = argmax[b * H + j];
      const float dpj = __ldg(dp + j);
      for (int t = 0; t < T; ++t) dxb[(long long)t * H + j] = (t == arg) ? dpj : 0.0f;
    }
  }
}

// d w[j] = sum_b dw_partial[b][j], d b = sum_b db_partial[b]: fixed order over the clips
__global__ void frame_pool_reduce_kernel(const float* __restrict__ dw_partial, const float* __restrict__ db_partial, long long B,
                                         int H, float* __restrict__ dw, float* __restrict__ db) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    float s = 0.0f;
    for (long long b = 0; b < B; ++b) s += dw_partial[b * H + j];
    dw[j] = s;
  } else if (j == H && db != nullptr) {
    float s = 0.0f;
    for (long long b = 0; b < B; ++b) s += db_partial[b];
    db[0] = s;
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_frame_pool_forward(const float* x, const float* mask, const float* w, const float* bias, int64_t batch,
                                      int32_t frames, int32_t hidden, int32_t mode, float* pooled, float* weights,
                                      int32_t* argmax, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && pooled && batch >= 0 && frames >= 1 && frames <= FP_MAX_FRAMES && hidden >= 1 && mode >= 0 && mode <= 2,
              "msf_frame_pool_forward: bad arguments (1..%d frames, mode 0 attention / 1 average / 2 max)", FP_MAX_FRAMES);
  MSF_REQUIRE(mode != 0 || w != nullptr, "msf_frame_pool_forward: attention pooling needs the scoring weights");
  MSF_REQUIRE((mode == 2 ? argmax != nullptr : weights != nullptr), "msf_frame_pool_forward: null tape (weights / argmax)");
  MSF_REQUIRE(batch < (1ll << 31), "msf_frame_pool_forward: too many clips");
  if (batch == 0) return MSF_OK;
  const size_t smem = (size_t)(frames + 8) * sizeof(float);
  frame_pool_fwd_kernel<<<(unsigned)batch, FP_THREADS, smem, (cudaStream_t)stream>>>(x, mask, w, bias, frames, hidden, mode, pooled,
                                                                                    weights, argmax);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

extern "C" int msf_frame_pool_backward(const float* x, const float* w, const float* weights, const int32_t* argmax,
                                       const float* d_pooled, int64_t batch, int32_t frames, int32_t hidden, int32_t mode,
                                       float* dx, float* dw, float* db, float* scratch, void* stream) {
  using namespace msf;
  MSF_REQUIRE(x && d_pooled && dx && batch >= 0 && frames >= 1 && frames <= FP_MAX_FRAMES && hidden >= 1 && mode >= 0 && mode <= 2,
              "msf_frame_pool_backward: bad arguments");
  MSF_REQUIRE(mode != 0 || (w && weights && dw && scratch), "msf_frame_pool_backward: attention pooling needs w, weights, dw, scratch");
  MSF_REQUIRE(mode != 1 || weights != nullptr, "msf_frame_pool_backward: average pooling needs the weights of the forward pass");
  MSF_REQUIRE(mode != 2 || argmax != nullptr, "msf_frame_pool_backward: max pooling needs the arg-max of the forward pass");
  MSF_REQUIRE(batch < (1ll << 31), "msf_frame_pool_backward: too many clips");
  if (batch == 0) {
    if (mode == 0) {
      MSF_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)hidden, (cudaStream_t)stream));
      if (db) MSF_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float), (cudaStream_t)stream));
    }
    return MSF_OK;
  }
  float* dw_partial = scratch;                                            // [B][H]
  float* db_partial = scratch ? scratch + (size_t)batch * hidden : nullptr;   // [B]
  const size_t smem = (size_t)(frames + 8) * sizeof(float);
  frame_pool_bwd_kernel<<<(unsigned)batch, FP_THREADS, smem, (cudaStream_t)stream>>>(x, w, weights, argmax, d_pooled, frames, hidden,
                                                                                    mode, dx, dw_partial, db_partial);
  MSF_LAUNCH_CHECK();
  if (mode == 0) {
    frame_pool_reduce_kernel<<<(unsigned)ceil_div((long long)hidden + 1, 256), 256, 0, (cudaStream_t)stream>>>(
        dw_partial, db_partial, batch, hidden, dw, db);
    MSF_LAUNCH_CHECK();
  }
  return MSF_OK;
}
