// Clip + AdamW over the HybridFusion master arena fused with the re-pack of the bf16 compute arena
// (src/train.py:378-382,416-430 followed by what msf_fusion_pack_bf16 does) and the advance of the
// device-side train state.  One launch updates every live parameter, writes its bf16 copy (and, for
// the GEMM weights, the transposed copy through a shared-memory tile) while the value is still in
// registers, so the master arena is not read a second time and two launches disappear from the step.
// The arithmetic is seg_adamw_kernel's (optim.cu), operation for operation.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "dp_signals.cuh"
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
  int fold_norm;   // 1: compute the gradient norm in this launch (needs all CTAs resident: grid barrier);
                   // 2: sq_norm[1] already holds it (MSF_OPT_NORM_GIVEN): no norm phase, no barrier
  // data-parallel mode (after dp_reduce_kernel): the gradient is this rank's reduced arena, filled by peer
  // writes; wait for every rank's "slice reduced" flag, form the norm from the published partials and close
  // the communicator epoch at the end
  unsigned long long* dp_sig;
  int dp_world;
};
template <bool DP>
__device__ __forceinline__ float ld_grad(const float* p) { return DP ? ld_peer1(p) : __ldg(p); }
template <bool DP>
__device__ __forceinline__ float4 ld_grad4(const float* p) {
  return DP ? ld_peer4(p) : __ldg(reinterpret_cast<const float4*>(p));
}
constexpr int OPT_VEC_UNIT = 1024, OPT_DEAD_UNIT = 2048;

__device__ unsigned int g_opt_ticket = 0;
__device__ unsigned int g_opt_barrier = 0;   // grid barrier between the norm phase and the update phase
__device__ double g_opt_sq = 0.0;            // sum of g^2 of the running launch (reset by its last CTA)

struct AdamConsts {
  float gs, step_size, decay, sqrt_bc2, beta1, beta2, ob1, ob2, eps;
};
__device__ __forceinline__ float adam_elem(const AdamConsts& k, float g, float& m, float& v, float p) {
  const float gi = g * k.gs;
  m = k.beta1 * m + k.ob1 * gi;
  v = k.beta2 * v + k.ob2 * gi * gi;
  return p * k.decay - k.step_size * (m / (sqrtf(v) / k.sqrt_bc2 + k.eps));
}

// All CTAs are resident (grid <= SMs x occupancy, nothing else runs on the stream): a counter barrier.
__device__ __forceinline__ void opt_grid_barrier() {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&g_opt_barrier, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(&g_opt_barrier) < gridDim.x) {
      if (clock64() - t0 > 4000000000ll) {
        printf("msf_b200 opt_pack: grid barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <bool DP>
__global__ void __launch_bounds__(256, 4) opt_pack_kernel(const __grid_constant__ OptList list, const OptCfg c,
                                                          float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          double* __restrict__ sq_norm,
                                                          unsigned long long* __restrict__ train_state,
                                                          bf16* __restrict__ arena) {
  TL_KERNEL(0);
  __shared__ float tile[32][33];
  __shared__ double red[8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  pdl_wait();
  TL_WAITED(0);
  pdl_launch();
  unsigned long long dp_epoch = 0ull;
  double dp_total = 0.0;
  if (DP) {
    dp_epoch = c.dp_sig[SIG_EPOCH] + 1ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) c.dp_sig[SIG_TIME + 3] = gtime();
    wait_all(c.dp_sig, SIG_DONE_R, c.dp_world, dp_epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) c.dp_sig[SIG_TIME + 4] = gtime();
    for (int r = 0; r < c.dp_world; ++r)
      dp_total += __longlong_as_double((long long)ld_acquire_sys(c.dp_sig + SIG_NORM + r));
  }

  if (!DP && c.fold_norm == 1) {  // ---- phase 1: sum of g^2 over the live slots (dead slots hold exact zeros) ----
    double sq = 0.0;
    for (int unit = blockIdx.x; unit < list.total_units; unit += gridDim.x) {
      int ji = 0;
      while (ji + 1 < list.count && unit >= list.j[ji + 1].unit_begin) ++ji;
      const OptJob& J = list.j[ji];
      if (J.kind == 1) continue;
      int local = unit - J.unit_begin;
      const int b = local / J.units_per_batch;
      local -= b * J.units_per_batch;
      const long long base = J.begin + (long long)b * J.batch_stride;
      if (J.kind == 2) {
        const int tc = (J.cols + 31) >> 5;
        const int r0 = (local / tc) << 5, c0 = (local % tc) << 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + ty + 8 * i, cc = c0 + tx;
          if (r < J.rows && cc < J.cols) {
            const double x = (double)__ldg(g + base + (long long)r * J.cols + cc);
            sq += x * x;
          }
        }
      } else {
        const int e1 = min(J.cols, (local + 1) * OPT_VEC_UNIT);
        for (int e = local * OPT_VEC_UNIT + threadIdx.x; e < e1; e += 256) {
          const double x = (double)__ldg(g + base + e);
          sq += x * x;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (tx == 0) red[ty] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += red[i];
      if (t != 0.0) atomicAdd(&g_opt_sq, t);
    }
    TL_MARK(0, 6);   // norm phase done (before the barrier)
    opt_grid_barrier();
    TL_MARK(0, 7);
  }

  // ---- phase 2: clip + AdamW + bf16 copies ----
  const double sq_total = DP ? dp_total
                             : c.fold_norm == 2 ? __ldcg(sq_norm + 1)
                             : c.fold_norm == 1 ? *reinterpret_cast<volatile double*>(&g_opt_sq) : *sq_norm;
  __shared__ AdamConsts ks;
  if (threadIdx.x == 0) {   // two fp64 pow() per CTA instead of per thread
    const double step = (double)train_state[2];
    AdamConsts t;
    const float bc1 = (float)(1.0 - pow((double)c.beta1, step));
    t.sqrt_bc2 = (float)sqrt(1.0 - pow((double)c.beta2, step));
    t.gs = c.grad_scale;
    if (c.max_norm > 0.0f) {
      const float total = (float)sqrt(sq_total) * c.grad_scale;
      t.gs *= fminf(c.max_norm / (total + 1e-6f), 1.0f);
    }
    const float lr = resolve_lr(c.lr, reinterpret_cast<const unsigned long long*>(train_state));
    t.step_size = lr / bc1;
    t.decay = 1.0f - lr * c.wd;
    t.beta1 = c.beta1; t.beta2 = c.beta2; t.ob1 = 1.0f - c.beta1; t.ob2 = 1.0f - c.beta2; t.eps = c.eps;
    ks = t;
  }
  __syncthreads();
  const AdamConsts k = ks;
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
      const bool vec4 = ((base & 3) == 0) && ((J.cols & 3) == 0) && ((J.dst_ld & 3) == 0) && ((J.dstT_ld & 3) == 0) &&
                        ((J.rows & 3) == 0);
      if (vec4) {  // thread = (row, 4 consecutive columns): 128-bit accesses on all four fp32 arrays
        const int r = r0 + (threadIdx.x >> 3), cc = c0 + (threadIdx.x & 7) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < J.rows && cc < J.cols) {
          const long long e = base + (long long)r * J.cols + cc;
          const float4 gg = ld_grad4<DP>(g + e);
          float4 mm = *reinterpret_cast<float4*>(m + e), vv = *reinterpret_cast<float4*>(v + e);
          x = *reinterpret_cast<float4*>(p + e);
          x.x = adam_elem(k, gg.x, mm.x, vv.x, x.x);
          x.y = adam_elem(k, gg.y, mm.y, vv.y, x.y);
          x.z = adam_elem(k, gg.z, mm.z, vv.z, x.z);
          x.w = adam_elem(k, gg.w, mm.w, vv.w, x.w);
          *reinterpret_cast<float4*>(p + e) = x;
          *reinterpret_cast<float4*>(m + e) = mm;
          *reinterpret_cast<float4*>(v + e) = vv;
          uint2 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
          h[0] = __floats2bfloat162_rn(x.x, x.y);
          h[1] = __floats2bfloat162_rn(x.z, x.w);
          *reinterpret_cast<uint2*>(dst + (long long)r * J.dst_ld + cc) = pk;
        }
        const int lr = threadIdx.x >> 3, lc = (threadIdx.x & 7) * 4;
        tile[lr][lc] = x.x; tile[lr][lc + 1] = x.y; tile[lr][lc + 2] = x.z; tile[lr][lc + 3] = x.w;
        __syncthreads();
        {  // destination row = source column; 4 consecutive destination columns = 4 source rows
          const int tcn = c0 + (threadIdx.x >> 3), tr = r0 + (threadIdx.x & 7) * 4;
          if (tcn < J.cols && tr < J.rows) {
            const int sc = threadIdx.x >> 3, sr = (threadIdx.x & 7) * 4;
            uint2 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
            h[0] = __floats2bfloat162_rn(tile[sr][sc], tile[sr + 1][sc]);
            h[1] = __floats2bfloat162_rn(tile[sr + 2][sc], tile[sr + 3][sc]);
            *reinterpret_cast<uint2*>(dstT + (long long)tcn * J.dstT_ld + tr) = pk;
          }
        }
        __syncthreads();
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + ty + 8 * i, cc = c0 + tx;
          float x = 0.0f;
          if (r < J.rows && cc < J.cols) {
            const long long e = base + (long long)r * J.cols + cc;
            float mm = m[e], vv = v[e];
            x = adam_elem(k, ld_grad<DP>(g + e), mm, vv, p[e]);
            p[e] = x; m[e] = mm; v[e] = vv;
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
      }
    } else if (J.kind == 0) {
      const int e1 = min(J.cols, (local + 1) * OPT_VEC_UNIT);
      for (int e = local * OPT_VEC_UNIT + threadIdx.x; e < e1; e += 256) {
        const long long ee = base + e;
        float mm = m[ee], vv = v[ee];
        p[ee] = adam_elem(k, ld_grad<DP>(g + ee), mm, vv, p[ee]);
        m[ee] = mm; v[ee] = vv;
      }
    } else {  // g = m = v = 0: only the decoupled weight decay acts
      const int e0 = local * OPT_DEAD_UNIT, e1 = min(J.cols, e0 + OPT_DEAD_UNIT);
      if (((base + e0) & 3) == 0) {
        float4* p4 = reinterpret_cast<float4*>(p + base + e0);
        const int n4 = (e1 - e0) >> 2;
        for (int i = threadIdx.x; i < n4; i += 256) {
          float4 x = p4[i];
          x.x *= k.decay; x.y *= k.decay; x.z *= k.decay; x.w *= k.decay;
          p4[i] = x;
        }
        for (int e = e0 + (n4 << 2) + threadIdx.x; e < e1; e += 256) p[base + e] *= k.decay;
      } else {
        for (int e = e0 + threadIdx.x; e < e1; e += 256) p[base + e] *= k.decay;
      }
    }
  }
  // the last CTA to finish publishes the norm, resets the launch-scoped globals and moves the train state on
  {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&g_opt_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
      if (DP) {   // close the communicator epoch
        if (sq_norm != nullptr) *sq_norm = sq_total;
        c.dp_sig[SIG_EPOCH] = dp_epoch;
        c.dp_sig[SIG_TIME + 5] = gtime();
      } else if (c.fold_norm == 1) {
        *sq_norm = sq_total;
        g_opt_sq = 0.0;
        g_opt_barrier = 0;
      } else if (c.fold_norm == 2) {
        *sq_norm = sq_total;
      }
      if (c.advance) {   // {seed, offset + 1, step + 1}
        train_state[1] += 1ull;
        train_state[2] += 1ull;
      }
      g_opt_ticket = 0;
    }
  }
}

}  // namespace

static int build_jobs(const Layout& L, OptList& list) {
  const ArenaBf16 A = arena_layout(L);
  memset(&list, 0, sizeof(list));
  const long long H = L.H;
  auto push = [&](OptJob J, int units) {
    if (J.batch <= 0 || units <= 0) return;
    J.unit_begin = list.total_units;
    J.units_per_batch = units;
    list.total_units += J.batch * units;
    list.j[list.count++] = J;
  };
  // Units are dealt to the CTAs round-robin in table order, so the table lists the heavy units first (matrix tiles:
  // four fp32 arrays plus two bf16 copies per element), then the live vectors, and the light weight-decay-only
  // units last: every CTA gets the same number of tiles (+-1) and the light units fill the remainder.
  int want = 2;
  auto matrix = [&](long long begin, long long stride, int batch, int rows, int cols, size_t dst, int dst_ld,
                    size_t dstT, int dstT_ld, long long dst_batch) {
    if (want != 2) return;
    OptJob J;
    memset(&J, 0, sizeof(J));
    J.kind = 2; J.begin = begin; J.batch_stride = stride; J.batch = batch; J.rows = rows; J.cols = cols;
    J.dst = (long long)dst; J.dstT = (long long)dstT; J.dst_batch = dst_batch; J.dst_ld = dst_ld; J.dstT_ld = dstT_ld;
    push(J, ((rows + 31) / 32) * ((cols + 31) / 32));
  };
  auto vec = [&](long long begin, long long stride, int batch, long long count, int dead) {
    if (want != (dead ? 1 : 0)) return;
    OptJob J;
    memset(&J, 0, sizeof(J));
    J.kind = dead ? 1 : 0; J.begin = begin; J.batch_stride = stride; J.batch = batch; J.rows = 1; J.cols = (int)count;
    const int unit = dead ? OPT_DEAD_UNIT : OPT_VEC_UNIT;
    push(J, (int)((count + unit - 1) / unit));
  };
  const int pairs = L.num_pairs();
  const int order[3] = {2, 0, 1};
  for (int pass = 0; pass < 3; ++pass) {
    want = order[pass];
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
  }
  MSF_REQUIRE(list.count <= OPT_MAX_JOBS, "opt_pack: job table overflow");
  return MSF_OK;
}

int fusion_bf16_opt_pack(const Layout& L, float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                         uint64_t* train_state, float lr, float beta1, float beta2, float eps, float wd,
                         float grad_scale, float max_norm, double* sq_norm, void* arena_v, int flags,
                         cudaStream_t st) {
  const int advance = flags & 1;
  const bool given = (flags & MSF_OPT_NORM_GIVEN) != 0;
  OptList list;
  int rc0 = build_jobs(L, list);
  if (rc0) return rc0;
  // One launch when every CTA can be resident at once (the norm phase ends in a grid barrier);
  // MSF_OPT_TWO_PASS=1 keeps the separate norm kernel.
  static int resident = -1;
  if (resident < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MSF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, opt_pack_kernel<false>, 256, 0));
    resident = sms * per_sm;
  }
  const bool fold = given || (resident >= 148 && !getenv("MSF_OPT_TWO_PASS"));
  int rc = MSF_OK;
  if (!fold && (rc = fusion_live_sq_norm(L, grad, sq_norm, st))) return rc;
  OptCfg c{lr, beta1, beta2, eps, wd, grad_scale, max_norm, advance, given ? 2 : fold ? 1 : 0, nullptr, 0};
  int grid = list.total_units < 1184 ? list.total_units : 1184;
  if (fold && grid > resident) grid = resident;
  MSF_CHECK_CUDA(launch_pdl(opt_pack_kernel<false>, dim3(grid), dim3(256), 0, st, list, c, params, grad, exp_avg,
                            exp_avg_sq, sq_norm, reinterpret_cast<unsigned long long*>(train_state),
                            reinterpret_cast<bf16*>(arena_v)));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

// Data-parallel variant: `reduced` is this rank's reduced-gradient arena (pushed by the slice owners in
// dp_reduce_kernel), `sig` its signal block.  Must follow dp_reduce_kernel on the same stream.
int fusion_bf16_opt_pack_dp(const Layout& L, float* params, const float* reduced, unsigned long long* sig, int world,
                            float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr, float beta1,
                            float beta2, float eps, float wd, float grad_scale, float max_norm, void* arena_v,
                            int advance, cudaStream_t st) {
  OptList list;
  int rc = build_jobs(L, list);
  if (rc) return rc;
  OptCfg c{lr, beta1, beta2, eps, wd, grad_scale, max_norm, advance, 0, sig, world};
  const int grid = list.total_units < 592 ? list.total_units : 592;
  opt_pack_kernel<true><<<grid, 256, 0, st>>>(list, c, params, reduced, exp_avg, exp_avg_sq, nullptr,
                                              reinterpret_cast<unsigned long long*>(train_state),
                                              reinterpret_cast<bf16*>(arena_v));
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

}  // namespace msf
