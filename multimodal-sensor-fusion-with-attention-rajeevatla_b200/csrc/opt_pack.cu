// Clip + AdamW over the HybridFusion master arena fused with the re-pack of the bf16 compute arena
// (src/train.py:378-382,416-430 followed by what msf_fusion_pack_bf16 does) and the advance of the
// device-side train state.  One launch updates every live parameter, writes its bf16 copy (and, for
// the GEMM weights, the transposed copy through a shared-memory tile) while the value is still in
// registers, so the master arena is not read a second time and two launches disappear from the step.
// The arithmetic is seg_adamw_kernel's (optim.cu), operation for operation.
#include <math.h>
#include <string.h>

#include "fusion_bf16_layout.cuh"

namespace msf {

int fusion_live_sq_norm(const Layout& L, const float* grad, double* sq_norm, cudaStream_t st);  // optim.cu

namespace {

struct OptJob {
  long long begin, batch_stride;   // master-arena element offset of batch 0 / between batches
  long long dst, dstT, dst_batch;  // compute-arena element offsets (matrix jobs)
  int kind;                        // 0 live vector, 1 dead slot (weight decay only), 2 live matrix + bf16 copies
  int rows, cols;                  // matrix: rows x cols row-major; vector / dead: cols elements
  int batch;
  int dst_ld, dstT_ld;
  int unit_begin, units_per_batch;
};
constexpr int OPT_MAX_JOBS = 2 * MSF_MAX_MODALITIES + 16;
struct OptList {
  OptJob j[OPT_MAX_JOBS];
  int count, total_units;
};
struct OptCfg {
  float lr, beta1, beta2, eps, wd, grad_scale, max_norm;
  int advance;
};
constexpr int OPT_VEC_UNIT = 1024, OPT_DEAD_UNIT = 4096;

__device__ unsigned int g_opt_ticket = 0;

__global__ void __launch_bounds__(256) opt_pack_kernel(const __grid_constant__ OptList list, const OptCfg c,
                                                       float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       const double* __restrict__ sq_norm,
                                                       unsigned long long* __restrict__ train_state,
                                                       bf16* __restrict__ arena) {
  __shared__ float tile[32][33];
  const double step = (double)train_state[2];
  const float bc1 = (float)(1.0 - pow((double)c.beta1, step));
  const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)c.beta2, step));
  float gs = c.grad_scale;
  if (c.max_norm > 0.0f) {
    const float total = (float)sqrt(*sq_norm) * c.grad_scale;
    gs *= fminf(c.max_norm / (total + 1e-6f), 1.0f);
  }
  const float step_size = c.lr / bc1, decay = 1.0f - c.lr * c.wd;
  const float ob1 = 1.0f - c.beta1, ob2 = 1.0f - c.beta2;
  auto adam = [&](long long e) -> float {
    const float gi = __ldg(g + e) * gs;
    const float mi = c.beta1 * m[e] + ob1 * gi;
    const float vi = c.beta2 * v[e] + ob2 * gi * gi;
    const float x = p[e] * decay - step_size * (mi / (sqrtf(vi) / sqrt_bc2 + c.eps));
    p[e] = x;
    m[e] = mi;
    v[e] = vi;
    return x;
  };
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int unit = blockIdx.x; unit < list.total_units; unit += gridDim.x) {
    int ji = 0;
    while (ji + 1 < list.count && unit >= list.j[ji + 1].unit_begin) ++ji;
    const OptJob& J = list.j[ji];
    int local = unit - J.unit_begin;
    const int b = local / J.units_per_batch;
    local -= b * J.units_per_batch;
    const long long base = J.begin + (long long)b * J.batch_stride;
    if (J.kind == 2) {
      const int tc = (J.cols + 31) >> 5;
      const int r0 = (local / tc) << 5, c0 = (local % tc) << 5;
      bf16* dst = arena + J.dst + (long long)b * J.dst_batch;
      bf16* dstT = arena + J.dstT + (long long)b * J.dst_batch;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + 8 * i, cc = c0 + tx;
        float x = 0.0f;
        if (r < J.rows && cc < J.cols) {
          x = adam(base + (long long)r * J.cols + cc);
          dst[(long long)r * J.dst_ld + cc] = __float2bfloat16_rn(x);
        }
        tile[ty + 8 * i][tx] = x;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cc = c0 + ty + 8 * i, r = r0 + tx;   // destination row = source column
        if (cc < J.cols && r < J.rows) dstT[(long long)cc * J.dstT_ld + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
      }
      __syncthreads();
    } else if (J.kind == 0) {
      const int e1 = min(J.cols, (local + 1) * OPT_VEC_UNIT);
      for (int e = local * OPT_VEC_UNIT + threadIdx.x; e < e1; e += 256) adam(base + e);
    } else {  // g = m = v = 0: only the decoupled weight decay acts
      const int e0 = local * OPT_DEAD_UNIT, e1 = min(J.cols, e0 + OPT_DEAD_UNIT);
      if (((base + e0) & 3) == 0) {
        float4* p4 = reinterpret_cast<float4*>(p + base + e0);
        const int n4 = (e1 - e0) >> 2;
        for (int i = threadIdx.x; i < n4; i += 256) {
          float4 x = p4[i];
          x.x *= decay; x.y *= decay; x.z *= decay; x.w *= decay;
          p4[i] = x;
        }
        for (int e = e0 + (n4 << 2) + threadIdx.x; e < e1; e += 256) p[base + e] *= decay;
      } else {
        for (int e = e0 + threadIdx.x; e < e1; e += 256) p[base + e] *= decay;
      }
    }
  }
  if (c.advance) {  // the last CTA to finish moves the train state on: {seed, offset + 1, step + 1}
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&g_opt_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
      train_state[1] += 1ull;
      train_state[2] += 1ull;
      g_opt_ticket = 0;
    }
  }
}

}  // namespace

int fusion_bf16_opt_pack(const Layout& L, float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                         uint64_t* train_state, float lr, float beta1, float beta2, float eps, float wd,
                         float grad_scale, float max_norm, double* sq_norm, void* arena_v, int advance,
                         cudaStream_t st) {
  const ArenaBf16 A = arena_layout(L);
  OptList list;
  memset(&list, 0, sizeof(list));
  const long long H = L.H;
  auto push = [&](OptJob J, int units) {
    if (J.batch <= 0 || units <= 0) return;
    J.unit_begin = list.total_units;
    J.units_per_batch = units;
    list.total_units += J.batch * units;
    list.j[list.count++] = J;
  };
  auto matrix = [&](long long begin, long long stride, int batch, int rows, int cols, size_t dst, int dst_ld,
                    size_t dstT, int dstT_ld, long long dst_batch) {
    OptJob J;
    memset(&J, 0, sizeof(J));
    J.kind = 2; J.begin = begin; J.batch_stride = stride; J.batch = batch; J.rows = rows; J.cols = cols;
    J.dst = (long long)dst; J.dstT = (long long)dstT; J.dst_batch = dst_batch; J.dst_ld = dst_ld; J.dstT_ld = dstT_ld;
    push(J, ((rows + 31) / 32) * ((cols + 31) / 32));
  };
  auto vec = [&](long long begin, long long stride, int batch, long long count, int dead) {
    OptJob J;
    memset(&J, 0, sizeof(J));
    J.kind = dead ? 1 : 0; J.begin = begin; J.batch_stride = stride; J.batch = batch; J.rows = 1; J.cols = (int)count;
    const int unit = dead ? OPT_DEAD_UNIT : OPT_VEC_UNIT;
    push(J, (int)((count + unit - 1) / unit));
  };
  const int pairs = L.num_pairs();
  for (int m = 0; m < L.M; ++m) {
    matrix(L.proj_w[m], 0, 1, L.H, L.D[m], A.wp[m], L.D[m], A.wpT[m], L.H, 0);
    vec(L.proj_b[m], 0, 1, H, 0);
  }
  if (pairs > 0) {
    vec(L.pair_w(0, 0), L.pair_stride, pairs, 2 * (H * H + H), 1);   // query_proj + key_proj: dead
    matrix(L.pair_w(0, 2), L.pair_stride, pairs, L.H, L.H, A.wv, L.H, A.wvT, L.H, H * H);
    vec(L.pair_b(0, 2), L.pair_stride, pairs, H, 0);
    matrix(L.pair_w(0, 3), L.pair_stride, pairs, L.H, L.H, A.wo, L.H, A.woT, L.H, H * H);
    vec(L.pair_b(0, 3), L.pair_stride, pairs, H, 0);
  }
  vec(L.gate_w[0], 0, 1, L.cls_w1 - L.gate_w[0], 0);   // gating layers: M x (H weights + 1 bias), contiguous
  matrix(L.cls_w1, 0, 1, L.H, L.H, A.w1, L.H, A.w1T, L.H, 0);
  vec(L.cls_b1, 0, 1, H, 0);
  matrix(L.cls_w2, 0, 1, L.C, L.H, A.w2, L.H, A.w2T, A.Cp, 0);   // w2T padding columns stay zero
  vec(L.cls_b2, 0, 1, L.C, 0);
  MSF_REQUIRE(list.count <= OPT_MAX_JOBS, "opt_pack: job table overflow");

  int rc = fusion_live_sq_norm(L, grad, sq_norm, st);
  if (rc) return rc;
  OptCfg c{lr, beta1, beta2, eps, wd, grad_scale, max_norm, advance};
  const int grid = list.total_units < 1184 ? list.total_units : 1184;
  opt_pack_kernel<<<grid, 256, 0, st>>>(list, c, params, grad, exp_avg, exp_avg_sq, sq_norm,
                                        reinterpret_cast<unsigned long long*>(train_state),
                                        reinterpret_cast<bf16*>(arena_v));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // namespace msf
