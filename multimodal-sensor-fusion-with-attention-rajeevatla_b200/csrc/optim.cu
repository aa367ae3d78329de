// Flat-arena optimizer pieces: gradient square-norm, AdamW with global-norm
// clipping, and gather/scatter between per-tensor storage and the arena.
//   src/train.py:378-382   torch.optim.AdamW(lr, weight_decay)
//   src/train.py:416-430   clip_gradients(..., gradient_clip_algorithm="norm")
#include <string.h>

#include "msf_common.cuh"

namespace msf {
namespace {

__global__ void __launch_bounds__(256) sq_norm_kernel(const float* __restrict__ g, long long n,
                                                      double* __restrict__ out) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)__ldg(g + i);
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n,
                                                    float lr, float beta1, float beta2, float eps, float wd,
                                                    float bc1, float sqrt_bc2, float grad_scale,
                                                    float max_norm, const double* __restrict__ sq_norm,
                                                    const unsigned long long* __restrict__ train_state) {
  if (train_state != nullptr) {  // step lives on the device (CUDA-graph replay)
    const double step = (double)train_state[2];
    bc1 = (float)(1.0 - pow((double)beta1, step));
    sqrt_bc2 = (float)sqrt(1.0 - pow((double)beta2, step));
  }
  float gs = grad_scale;
  if (max_norm > 0.0f && sq_norm != nullptr) {
    const float total = (float)sqrt(*sq_norm) * grad_scale;
    const float coef = max_norm / (total + 1e-6f);
    gs *= fminf(coef, 1.0f);
  }
  lr = resolve_lr(lr, train_state);
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

// table: n_tensors x {ptr, arena_offset, numel}
template <bool GATHER>
__global__ void __launch_bounds__(256) arena_copy_kernel(const long long* __restrict__ table, int n_tensors,
                                                         float* __restrict__ arena) {
  const int t = blockIdx.y;
  if (t >= n_tensors) return;
  float* ptr = reinterpret_cast<float*>(table[3 * t]);
  const long long off = table[3 * t + 1], numel = table[3 * t + 2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    if (GATHER) arena[off + i] = ptr[i];
    else ptr[i] = arena[off + i];
  }
}

// ---- layout-aware optimizer over the HybridFusion master arena -------------------------------------
// The query/key projections of every pair module are dead inside HybridFusion: their gradient and both
// Adam moments are identically zero, so AdamW reduces to the weight-decay multiply p *= 1 - lr*wd
// (m = v = 0 gives a zero Adam update exactly).  Dead segments therefore touch 8 bytes per parameter
// instead of 28 and are skipped by the gradient norm.
struct ArenaSeg {
  long long begin, count;
  int dead;
};
constexpr int OPT_MAX_SEGS = 2 * MSF_MAX_MODALITIES * (MSF_MAX_MODALITIES - 1) + 2;
struct ArenaSegs {
  ArenaSeg s[OPT_MAX_SEGS];
  int n;
};

__global__ void __launch_bounds__(256) seg_sq_norm_kernel(const __grid_constant__ ArenaSegs segs,
                                                          const float* __restrict__ g, double* __restrict__ out) {
  const ArenaSeg sg = segs.s[blockIdx.y];
  if (sg.dead) return;
  double s = 0.0;
  const float* p = g + sg.begin;
  if ((sg.begin & 3) == 0) {
    const long long n4 = sg.count >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (sg.count & 3)) {
      const double v = (double)__ldg(p + (n4 << 2) + threadIdx.x);
      s += v * v;
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < sg.count; i += (long long)gridDim.x * blockDim.x) {
      const double v = (double)__ldg(p + i);
      s += v * v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    if (t != 0.0) atomicAdd(out, t);
  }
}

struct AdamCfg {
  float lr, beta1, beta2, eps, wd, grad_scale, max_norm;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCfg& c, float gs, float decay,
                                         float step_size, float inv_sqrt_bc2) {
  const float gi = g * gs;
  float pi = p * decay;
  m = c.beta1 * m + (1.0f - c.beta1) * gi;
  v = c.beta2 * v + (1.0f - c.beta2) * gi * gi;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + c.eps;
  p = pi - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) seg_adamw_kernel(const __grid_constant__ ArenaSegs segs, const AdamCfg c,
                                                        float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const double* __restrict__ sq_norm,
                                                        const unsigned long long* __restrict__ train_state) {
  const ArenaSeg sg = segs.s[blockIdx.y];
  const double step = (double)train_state[2];
  const float bc1 = (float)(1.0 - pow((double)c.beta1, step));
  const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)c.beta2, step));
  float gs = c.grad_scale;
  if (c.max_norm > 0.0f && sq_norm != nullptr) {
    const float total = (float)sqrt(*sq_norm) * c.grad_scale;
    gs *= fminf(c.max_norm / (total + 1e-6f), 1.0f);
  }
  const float lr = resolve_lr(c.lr, train_state);
  const float step_size = lr / bc1, decay = 1.0f - lr * c.wd;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool vec = (sg.begin & 3) == 0;
  const long long n4 = vec ? (sg.count >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(p + sg.begin);
  if (sg.dead) {  // g = m = v = 0: only the decoupled weight decay acts
    for (long long i = t0; i < n4; i += stride) {
      float4 x = p4[i];
      x.x *= decay; x.y *= decay; x.z *= decay; x.w *= decay;
      p4[i] = x;
    }
    for (long long i = (n4 << 2) + t0; i < sg.count; i += stride) p[sg.begin + i] *= decay;
    return;
  }
  const float4* g4 = reinterpret_cast<const float4*>(g + sg.begin);
  float4* m4 = reinterpret_cast<float4*>(m + sg.begin);
  float4* v4 = reinterpret_cast<float4*>(v + sg.begin);
  const float inv_sqrt_bc2 = 1.0f / sqrt_bc2;
  for (long long i = t0; i < n4; i += stride) {
    float4 x = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    // same operation order as adamw_kernel: sqrt(v) / sqrt(bc2) + eps
    const float gx = gg.x * gs, gy = gg.y * gs, gz = gg.z * gs, gw = gg.w * gs;
    mm.x = c.beta1 * mm.x + (1.0f - c.beta1) * gx; vv.x = c.beta2 * vv.x + (1.0f - c.beta2) * gx * gx;
    mm.y = c.beta1 * mm.y + (1.0f - c.beta1) * gy; vv.y = c.beta2 * vv.y + (1.0f - c.beta2) * gy * gy;
    mm.z = c.beta1 * mm.z + (1.0f - c.beta1) * gz; vv.z = c.beta2 * vv.z + (1.0f - c.beta2) * gz * gz;
    mm.w = c.beta1 * mm.w + (1.0f - c.beta1) * gw; vv.w = c.beta2 * vv.w + (1.0f - c.beta2) * gw * gw;
    x.x = x.x * decay - step_size * (mm.x / (sqrtf(vv.x) / sqrt_bc2 + c.eps));
    x.y = x.y * decay - step_size * (mm.y / (sqrtf(vv.y) / sqrt_bc2 + c.eps));
    x.z = x.z * decay - step_size * (mm.z / (sqrtf(vv.z) / sqrt_bc2 + c.eps));
    x.w = x.w * decay - step_size * (mm.w / (sqrtf(vv.w) / sqrt_bc2 + c.eps));
    p4[i] = x; m4[i] = mm; v4[i] = vv;
  }
  (void)inv_sqrt_bc2;
  for (long long i = (n4 << 2) + t0; i < sg.count; i += stride) {
    const long long e = sg.begin + i;
    const float gi = g[e] * gs;
    const float mi = c.beta1 * m[e] + (1.0f - c.beta1) * gi;
    const float vi = c.beta2 * v[e] + (1.0f - c.beta2) * gi * gi;
    p[e] = p[e] * decay - step_size * (mi / (sqrtf(vi) / sqrt_bc2 + c.eps));
    m[e] = mi;
    v[e] = vi;
  }
}

__global__ void train_state_advance_kernel(unsigned long long* state) {
  state[1] += 1ull;
  state[2] += 1ull;
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(DropCfg d, int site, int sub, long long rows,
                                                           long long cols, float* __restrict__ out) {
  const long long quads = (cols + 3) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < rows * quads;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / quads;
    const int c4 = (int)(i % quads);
    float v[4];
    drop4(d, site, sub, row, c4, v);
    for (int j = 0; j < 4; ++j)
      if (c4 * 4 + j < cols) out[row * cols + c4 * 4 + j] = v[j];
  }
}

static void live_segments(const Layout& L, ArenaSegs* segs) {
  memset(segs, 0, sizeof(*segs));
  auto add = [&](long long begin, long long count, int dead) {
    if (count <= 0) return;
    segs->s[segs->n].begin = begin; segs->s[segs->n].count = count; segs->s[segs->n].dead = dead;
    ++segs->n;
  };
  const long long half = 2 * ((long long)L.H * L.H + L.H);
  add(0, L.pair_base, 0);
  for (int p = 0; p < L.num_pairs(); ++p) {
    add(L.pair_base + p * L.pair_stride, half, 1);         // query_proj + key_proj: dead
    add(L.pair_base + p * L.pair_stride + half, half, 0);  // value_proj + out_proj
  }
  const long long tail = L.pair_base + (long long)L.num_pairs() * L.pair_stride;
  add(tail, L.total - tail, 0);
}

}  // namespace

// sq_norm[0] = sum of g^2 over the live slots (the dead query/key slots hold exact zeros)
int fusion_live_sq_norm(const Layout& L, const float* grad, double* sq_norm, cudaStream_t st) {
  ArenaSegs segs;
  live_segments(L, &segs);
  MSF_CHECK_CUDA(cudaMemsetAsync(sq_norm, 0, sizeof(double), st));
  dim3 grid(24, (unsigned)segs.n);
  seg_sq_norm_kernel<<<grid, 256, 0, st>>>(segs, grad, sq_norm);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

bool fusion_bf16_eligible(const Layout& L);
int fusion_bf16_opt_pack(const Layout& L, float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                         uint64_t* train_state, float lr, float beta1, float beta2, float eps, float wd,
                         float grad_scale, float max_norm, double* sq_norm, void* arena_v, int advance,
                         cudaStream_t st);

}  // namespace msf

extern "C" {

int msf_grad_sq_norm(const float* grad, int64_t n, double* sq_norm, void* stream) {
  MSF_REQUIRE(grad && sq_norm && n >= 0, "msf_grad_sq_norm: bad arguments");
  if (n == 0) return MSF_OK;
  long long blocks = msf::ceil_div(n, 256 * 8);
  if (blocks > 592) blocks = 592;
  msf::sq_norm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad, n, sq_norm);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_adamw_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, float max_norm, const double* sq_norm, void* stream) {
  MSF_REQUIRE(params && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "msf_adamw_step: bad arguments");
  if (n == 0) return MSF_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = msf::ceil_div(n, 256 * 4);
  if (blocks > 1184) blocks = 1184;
  msf::adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
      (float)sqrt(bc2), grad_scale, max_norm, sq_norm, nullptr);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_adamw_step_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float grad_scale, float max_norm, const double* sq_norm,
                       void* stream) {
  MSF_REQUIRE(params && grad && exp_avg && exp_avg_sq && train_state && n >= 0, "msf_adamw_step_dev: bad arguments");
  if (n == 0) return MSF_OK;
  long long blocks = msf::ceil_div(n, 256 * 4);
  if (blocks > 1184) blocks = 1184;
  msf::adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, 1.0f, 1.0f, grad_scale,
      max_norm, sq_norm, reinterpret_cast<const unsigned long long*>(train_state));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_fusion_optimizer_step(const msf_fusion_shape* shape, float* params, const float* grad, float* exp_avg,
                              float* exp_avg_sq, const uint64_t* train_state, float lr, float beta1, float beta2,
                              float eps, float weight_decay, float grad_scale, float max_norm, double* sq_norm,
                              void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(params && grad && exp_avg && exp_avg_sq && train_state && sq_norm, "msf_fusion_optimizer_step: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = msf::fusion_live_sq_norm(L, grad, sq_norm, st))) return rc;
  msf::ArenaSegs segs;
  msf::live_segments(L, &segs);
  msf::AdamCfg c{lr, beta1, beta2, eps, weight_decay, grad_scale, max_norm};
  dim3 grid2(48, (unsigned)segs.n);
  msf::seg_adamw_kernel<<<grid2, 256, 0, st>>>(segs, c, params, grad, exp_avg, exp_avg_sq, sq_norm,
                                                reinterpret_cast<const unsigned long long*>(train_state));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_fusion_optimizer_step_packed(const msf_fusion_shape* shape, float* params, const float* grad,
                                     float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                     float max_norm, double* sq_norm, void* params_bf16, int32_t advance_state,
                                     void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(params && grad && exp_avg && exp_avg_sq && train_state && sq_norm && params_bf16,
              "msf_fusion_optimizer_step_packed: null pointer");
  if (!msf::fusion_bf16_eligible(L)) {
    msf::set_error("shape not eligible for the tensor-core path");
    return MSF_E_UNSUPPORTED;
  }
  return msf::fusion_bf16_opt_pack(L, params, grad, exp_avg, exp_avg_sq, train_state, lr, beta1, beta2, eps,
                                   weight_decay, grad_scale, max_norm, sq_norm, params_bf16, advance_state,
                                   (cudaStream_t)stream);
}

int msf_train_state_advance(uint64_t* train_state, void* stream) {
  MSF_REQUIRE(train_state != nullptr, "msf_train_state_advance: null state");
  msf::train_state_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<unsigned long long*>(train_state));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_arena_gather(const int64_t* table, int32_t n_tensors, int64_t total, float* arena, void* stream) {
  MSF_REQUIRE(table && arena && n_tensors >= 0 && total >= 0, "msf_arena_gather: bad arguments");
  if (n_tensors == 0) return MSF_OK;
  dim3 grid(64, (unsigned)n_tensors);
  msf::arena_copy_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(table), n_tensors, arena);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_arena_scatter(const int64_t* table, int32_t n_tensors, int64_t total, const float* arena, void* stream) {
  MSF_REQUIRE(table && arena && n_tensors >= 0 && total >= 0, "msf_arena_scatter: bad arguments");
  if (n_tensors == 0) return MSF_OK;
  dim3 grid(64, (unsigned)n_tensors);
  msf::arena_copy_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(table), n_tensors, const_cast<float*>(arena));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

int msf_dropout_mask(uint64_t seed, uint64_t offset, int32_t site, int32_t sub, int64_t rows, int64_t cols,
                     float p, float* out, void* stream) {
  MSF_REQUIRE(out && rows >= 0 && cols >= 0 && p >= 0.0f && p < 1.0f, "msf_dropout_mask: bad arguments");
  if (rows * cols == 0) return MSF_OK;
  msf::DropCfg d;
  d.seed = seed; d.offset = offset; d.p = p; d.scale = 1.0f / (1.0f - p); d.active = 1; d.state = nullptr;
  long long blocks = msf::ceil_div(rows * ((cols + 3) / 4), 256);
  if (blocks > 2048) blocks = 2048;
  msf::dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d, site, sub, rows, cols, out);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // extern "C"


// Gradient accumulation over micro-batches (config/base.yaml:75, src/train.py:519-521): acc = (first ? 0 : acc) + grad,
// and the running mean of the micro-batch losses beside it.
namespace msf {
namespace {
__global__ void grad_accumulate_kernel(float* __restrict__ acc, const float* __restrict__ grad, long long n, int first,
                                       float* __restrict__ loss_acc, const float* __restrict__ loss, float loss_scale) {
  const long long n4 = n >> 2;
  const bool vec = ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (long long i = tid; i < n4; i += stride) {
      float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
      if (!first) {
        const float4 a = reinterpret_cast<const float4*>(acc)[i];
        g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
      }
      reinterpret_cast<float4*>(acc)[i] = g;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) acc[i] = (first ? 0.0f : acc[i]) + __ldg(grad + i);
  } else {
    for (long long i = tid; i < n; i += stride) acc[i] = (first ? 0.0f : acc[i]) + __ldg(grad + i);
  }
  if (tid == 0 && loss_acc != nullptr && loss != nullptr) *loss_acc = (first ? 0.0f : *loss_acc) + loss_scale * *loss;
}
}  // namespace
}  // namespace msf

extern "C" int msf_grad_accumulate(float* acc, const float* grad, int64_t n, float* loss_acc, const float* loss,
                                   float loss_scale, int32_t first, void* stream) {
  using namespace msf;
  MSF_REQUIRE(acc && grad && n >= 1, "msf_grad_accumulate: bad arguments");
  const long long blocks = ceil_div(n >> 2, 256);
  grad_accumulate_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks > 1184 ? 1184 : blocks), 256, 0, (cudaStream_t)stream>>>(
      acc, grad, n, first, loss_acc, loss, loss_scale);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
