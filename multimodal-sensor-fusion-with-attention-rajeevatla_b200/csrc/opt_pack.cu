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
  int live_units;   // units [0, live_units) carry gradients (matrix tiles, then vectors); the rest are dead slots
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
// Sharded ("owner computes") data parallelism, msf_dpz_optimizer_step_packed: live unit u belongs to rank u % world.
// The owner sums the unit's gradient over the ranks (dpz_reduce_kernel), applies clip + AdamW to it and pushes the
// bf16 copies of the updated weights into every rank's compute arena; master weights and Adam moments of a unit
// live on its owner only.  Dead slots (weight decay only) are updated by every rank.
struct ZCfg {
  int rank, world;
  bf16* arenas[8];                 // peer-mapped compute arenas, rank order
  float* params[8];                // peer-mapped master arenas: updated VECTOR slots (biases, gating layers) are pushed
                                   // to every rank, the forward kernels read them from the fp32 master
  float* stages[8];                // peer-mapped staging arenas: world x total floats, [contributor][arena offset]
  unsigned long long* sigs[8];     // peer-mapped signal blocks
  // optional NVLink multicast addresses of the ranks' gradient / compute / master arenas (all three or none): the
  // owner then reads a unit's gradient summed by the switch (no staging pushes) and stores updated weights once
  const float* mc_grad;
  bf16* mc_arena;
  float* mc_params;
  long long total;                 // stride of a staging arena's per-contributor slices: master-arena elements rounded up to 4
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

template <bool DP, bool Z>
__global__ void __launch_bounds__(256, 4) opt_pack_kernel(const __grid_constant__ OptList list, const OptCfg c,
                                                          float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          double* __restrict__ sq_norm,
                                                          unsigned long long* __restrict__ train_state,
                                                          bf16* __restrict__ arena, const __grid_constant__ ZCfg z) {
  TL_KERNEL(0);
  __shared__ float tile[32][33];
  __shared__ double red[8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  pdl_wait();
  TL_WAITED(0);
  pdl_launch();
  unsigned long long dp_epoch = 0ull;
  double dp_total = 0.0;
  if (DP || Z) {
    dp_epoch = c.dp_sig[SIG_EPOCH] + 1ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) c.dp_sig[SIG_TIME + 3] = gtime();
    wait_all(c.dp_sig, SIG_DONE_R, c.dp_world, dp_epoch);
    if (blockIdx.x == 0 && threadIdx.x == 0) c.dp_sig[SIG_TIME + 4] = gtime();
    for (int r = 0; r < c.dp_world; ++r)
      dp_total += __longlong_as_double((long long)ld_acquire_sys(c.dp_sig + SIG_NORM + r));
  }

  if (!DP && !Z && c.fold_norm == 1) {  // ---- phase 1: sum of g^2 over the live slots (dead slots hold exact zeros) ----
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
  const double sq_total = (DP || Z) ? dp_total
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
  // Z: owned live units (unit % world == rank), then every dead unit; otherwise every unit
  const int ndst = Z ? z.world : 1;
  const int live_n = Z ? (list.live_units - z.rank + z.world - 1) / z.world : list.total_units;
  const int iter_n = Z ? live_n + (list.total_units - list.live_units) : list.total_units;
  for (int it = blockIdx.x; it < iter_n; it += gridDim.x) {
    const int unit = !Z ? it : (it < live_n ? z.rank + it * z.world : list.live_units + (it - live_n));
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
      const long long dst_off = J.dst + (long long)b * J.dst_batch, dstT_off = J.dstT + (long long)b * J.dst_batch;
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
          if (Z && z.mc_arena != nullptr) mm_st_bf16x4(z.mc_arena + dst_off + (long long)r * J.dst_ld + cc, pk);
          else
            for (int d = 0; d < ndst; ++d)
              *reinterpret_cast<uint2*>((Z ? z.arenas[d] : arena) + dst_off + (long long)r * J.dst_ld + cc) = pk;
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
            if (Z && z.mc_arena != nullptr) mm_st_bf16x4(z.mc_arena + dstT_off + (long long)tcn * J.dstT_ld + tr, pk);
            else
              for (int d = 0; d < ndst; ++d)
                *reinterpret_cast<uint2*>((Z ? z.arenas[d] : arena) + dstT_off + (long long)tcn * J.dstT_ld + tr) = pk;
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
            for (int d = 0; d < ndst; ++d)
              ((Z ? z.arenas[d] : arena) + dst_off)[(long long)r * J.dst_ld + cc] = __float2bfloat16_rn(x);
          }
          tile[ty + 8 * i][tx] = x;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = c0 + ty + 8 * i, r = r0 + tx;   // destination row = source column
          if (cc < J.cols && r < J.rows)
            for (int d = 0; d < ndst; ++d)
              ((Z ? z.arenas[d] : arena) + dstT_off)[(long long)cc * J.dstT_ld + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
        }
        __syncthreads();
      }
    } else if (J.kind == 0) {
      const int e1 = min(J.cols, (local + 1) * OPT_VEC_UNIT);
      for (int e = local * OPT_VEC_UNIT + threadIdx.x; e < e1; e += 256) {
        const long long ee = base + e;
        float mm = m[ee], vv = v[ee];
        const float x = adam_elem(k, ld_grad<DP>(g + ee), mm, vv, p[ee]);
        m[ee] = mm; v[ee] = vv;
        if (Z && z.mc_params != nullptr) {
          mm_st_f32(z.mc_params + ee, x);
        } else if (Z) {
          for (int d = 0; d < ndst; ++d) z.params[d][ee] = x;   // every rank's forward reads these from the master
        } else {
          p[ee] = x;
        }
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
    // one fence per block, by the thread that reports in, after the block barrier (fences are cumulative: the
    // other threads' writes are ordered before it through the barrier)
    __syncthreads();
    if (threadIdx.x == 0) {
      if (Z) __threadfence_system();   // this block's pushed bf16 weights are visible system-wide before it reports in
      else __threadfence();
      last = (atomicAdd(&g_opt_ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (Z && last) {   // "my share of the compute arena has landed everywhere", then wait for everybody else's
      if ((int)threadIdx.x < z.world) {
        __threadfence_system();
        st_release_sys(z.sigs[threadIdx.x] + SIG_DONE_W + z.rank, dp_epoch);
      }
      wait_all(c.dp_sig, SIG_DONE_W, z.world, dp_epoch);
    }
    if (last && threadIdx.x == 0) {
      if (DP || Z) {   // close the communicator epoch
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

// ---------------------------------------------------------------------------
// Sharded step, first kernel: reduce-scatter of the live gradient units over NVLink peer memory.
//   phase 1  every rank pushes its values of every unit it does NOT own into the owner's staging arena
//            (slot [contributor rank][arena offset]: no index translation on either side);
//   phase 2  (after a cross-GPU flag barrier) the owner sums its units over the ranks in fixed rank order, in place
//            in its own gradient arena, accumulates the units' square norm and publishes it with its "reduced" flag.
// opt_pack_kernel<false, true> follows on the same stream.
// ---------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void unit_elems(const OptJob& J, int local, long long base, F f) {
  // f(arena offset, 4) for 16-byte aligned groups of 4 consecutive elements, f(offset, 1) otherwise
  if (J.kind == 2) {
    const int tc = (J.cols + 31) >> 5;
    const int r0 = (local / tc) << 5, c0 = (local % tc) << 5;
    if (((base & 3) == 0) && ((J.cols & 3) == 0)) {
      const int r = r0 + (threadIdx.x >> 3), cc = c0 + (threadIdx.x & 7) * 4;
      if (r < J.rows && cc < J.cols) f(base + (long long)r * J.cols + cc, 4);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + (threadIdx.x >> 5) + 8 * i, cc = c0 + (threadIdx.x & 31);
        if (r < J.rows && cc < J.cols) f(base + (long long)r * J.cols + cc, 1);
      }
    }
  } else {
    const int e0 = local * OPT_VEC_UNIT, e1 = min(J.cols, e0 + OPT_VEC_UNIT);
    if (((base + e0) & 3) == 0) {
      const int e = e0 + threadIdx.x * 4;
      if (e + 3 < e1) f(base + e, 4);
      else
        for (int j = e; j < e1; ++j) f(base + j, 1);
    } else {
      for (int e = e0 + threadIdx.x; e < e1; e += 256) f(base + e, 1);
    }
  }
}

__global__ void __launch_bounds__(256, 4) dpz_reduce_kernel(const __grid_constant__ OptList list,
                                                            const __grid_constant__ ZCfg z, float* __restrict__ grad) {
  unsigned long long* sig = z.sigs[z.rank];
  const unsigned long long epoch = sig[SIG_EPOCH] + 1ull;   // same on every rank: all ranks run the same steps
  if (blockIdx.x == 0 && threadIdx.x == 0) sig[SIG_TIME + 0] = gtime();
  __shared__ bool last;
  auto locate = [&](int unit, int& local, long long& base) -> const OptJob& {
    int ji = 0;
    while (ji + 1 < list.count && unit >= list.j[ji + 1].unit_begin) ++ji;
    const OptJob& J = list.j[ji];
    local = unit - J.unit_begin;
    const int b = local / J.units_per_batch;
    local -= b * J.units_per_batch;
    base = J.begin + (long long)b * J.batch_stride;
    return J;
  };

  // ---- phase 1: my values of the units other ranks own -> their staging arenas ----
  for (int unit = blockIdx.x; z.mc_grad == nullptr && unit < list.live_units; unit += gridDim.x) {
    const int owner = unit % z.world;
    if (owner == z.rank) continue;
    int local;
    long long base;
    const OptJob& J = locate(unit, local, base);
    float* dst = z.stages[owner] + (long long)z.rank * z.total;
    unit_elems(J, local, base, [&](long long e, int w) {
      if (w == 4) *reinterpret_cast<float4*>(dst + e) = __ldg(reinterpret_cast<const float4*>(grad + e));
      else dst[e] = __ldg(grad + e);
    });
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();   // cumulative: the block's pushes are visible system-wide before it reports in
    last = atomicAdd(sig + SIG_TICKET, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < z.world) {
    __threadfence_system();
    st_release_sys(z.sigs[threadIdx.x] + SIG_DONE_B + z.rank, epoch);   // "my contributions have landed"
  }
  wait_all(sig, SIG_DONE_B, z.world, epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) sig[SIG_TIME + 1] = gtime();

  // ---- phase 2: sum my units in rank order (bit-reproducible), in place; their square norm ----
  const float* stage = z.stages[z.rank];
  double sq = 0.0;
  for (int unit = z.rank + blockIdx.x * z.world; unit < list.live_units; unit += gridDim.x * z.world) {
    int local;
    long long base;
    const OptJob& J = locate(unit, local, base);
    unit_elems(J, local, base, [&](long long e, int w) {
      if (w == 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (z.mc_grad != nullptr) {
          acc = mm_ld_reduce4(z.mc_grad + e);   // summed over the ranks' arenas by the switch
        } else {
          for (int pr = 0; pr < z.world; ++pr) {
            const float4 gg = (pr == z.rank) ? *reinterpret_cast<const float4*>(grad + e)
                                             : ld_peer4(stage + (long long)pr * z.total + e);
            acc.x += gg.x; acc.y += gg.y; acc.z += gg.z; acc.w += gg.w;
          }
        }
        *reinterpret_cast<float4*>(grad + e) = acc;
        sq += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
      } else {
        float acc = 0.0f;
        if (z.mc_grad != nullptr) acc = mm_ld_reduce1(z.mc_grad + e);
        else
          for (int pr = 0; pr < z.world; ++pr) acc += (pr == z.rank) ? grad[e] : ld_peer1(stage + (long long)pr * z.total + e);
        grad[e] = acc;
        sq += (double)acc * acc;
      }
    });
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) sig[SIG_TIME + 6] = gtime();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(reinterpret_cast<double*>(sig + SIG_ACC), t);
    __threadfence();
    last = atomicAdd(sig + SIG_TICKET2, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < z.world) {   // the last block publishes: norm first, then the release flag
    const unsigned long long bits = ld_acquire_sys(sig + SIG_ACC);
    unsigned long long* peer = z.sigs[threadIdx.x];
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peer + SIG_NORM + z.rank), "l"(bits) : "memory");
    __threadfence_system();
    st_release_sys(peer + SIG_DONE_R + z.rank, epoch);
  }
  if (last) {
    __syncthreads();
    if (threadIdx.x == 0) {
      sig[SIG_ACC] = 0ull;
      sig[SIG_TICKET] = 0ull;
      sig[SIG_TICKET2] = 0ull;
      sig[SIG_TIME + 2] = gtime();
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
  list.live_units = list.total_units;
  for (int i = list.count - 1; i >= 0 && list.j[i].kind == 1; --i) list.live_units = list.j[i].unit_begin;
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
    MSF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, opt_pack_kernel<false, false>, 256, 0));
    resident = sms * per_sm;
  }
  const bool fold = given || (resident >= 148 && !getenv("MSF_OPT_TWO_PASS"));
  int rc = MSF_OK;
  if (!fold && (rc = fusion_live_sq_norm(L, grad, sq_norm, st))) return rc;
  OptCfg c{lr, beta1, beta2, eps, wd, grad_scale, max_norm, advance, given ? 2 : fold ? 1 : 0, nullptr, 0};
  int grid = list.total_units < 1184 ? list.total_units : 1184;
  if (fold && grid > resident) grid = resident;
  ZCfg z;
  memset(&z, 0, sizeof(z));
  MSF_CHECK_CUDA(launch_pdl(opt_pack_kernel<false, false>, dim3(grid), dim3(256), 0, st, list, c, params, grad, exp_avg,
                            exp_avg_sq, sq_norm, reinterpret_cast<unsigned long long*>(train_state),
                            reinterpret_cast<bf16*>(arena_v), z));
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
  ZCfg z;
  memset(&z, 0, sizeof(z));
  opt_pack_kernel<true, false><<<grid, 256, 0, st>>>(list, c, params, reduced, exp_avg, exp_avg_sq, nullptr,
                                                     reinterpret_cast<unsigned long long*>(train_state),
                                                     reinterpret_cast<bf16*>(arena_v), z);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

// Sharded data-parallel optimizer step (include/msf_b200.h: msf_dpz_optimizer_step_packed).
int fusion_bf16_opt_pack_dpz(const Layout& L, const msf_dpz_comm* comm, float* params, float* grad, float* exp_avg,
                             float* exp_avg_sq, uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                             float wd, float grad_scale, float max_norm, int advance, cudaStream_t st) {
  OptList list;
  int rc = build_jobs(L, list);
  if (rc) return rc;
  ZCfg z;
  memset(&z, 0, sizeof(z));
  z.rank = comm->rank; z.world = comm->world;
  z.total = (L.total + 3) & ~3ll;   // stride between the contributors' slices of a staging arena: keeps float4 groups aligned
  for (int r = 0; r < comm->world; ++r) {
    MSF_REQUIRE((comm->stages[r] || comm->mc_grad) && comm->arenas_bf16[r] && comm->params[r] && comm->sigs[r],
                "msf_dpz_optimizer_step_packed: null peer pointer (rank %d)", r);
    z.stages[r] = comm->stages[r];
    z.arenas[r] = reinterpret_cast<bf16*>(comm->arenas_bf16[r]);
    z.params[r] = comm->params[r];
    z.sigs[r] = reinterpret_cast<unsigned long long*>(comm->sigs[r]);
  }
  MSF_REQUIRE(params == comm->params[comm->rank], "msf_dpz_optimizer_step_packed: params must be this rank's entry of comm->params");
  const int n_mc = (comm->mc_grad != nullptr) + (comm->mc_arena_bf16 != nullptr) + (comm->mc_params != nullptr);
  MSF_REQUIRE(n_mc == 0 || n_mc == 3, "msf_dpz_optimizer_step_packed: give all three multicast addresses or none");
  z.mc_grad = comm->mc_grad;
  z.mc_arena = reinterpret_cast<bf16*>(comm->mc_arena_bf16);
  z.mc_params = comm->mc_params;
  // Every block waits inside the kernels for the peers' flags, which depend on ALL blocks of every rank: both
  // grids must be co-resident (occupancy 4 x SMs of 256 threads) or the ranks deadlock.
  static int resident = -1;
  if (resident < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MSF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dpz_reduce_kernel, 256, 0));
    int per_sm2 = 0;
    MSF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, opt_pack_kernel<false, true>, 256, 0));
    resident = sms * (per_sm < per_sm2 ? per_sm : per_sm2);
  }
  MSF_REQUIRE(resident >= 1, "msf_dpz_optimizer_step_packed: kernels cannot be resident");
  int grid = list.live_units < 592 ? list.live_units : 592;
  if (grid > resident) grid = resident;
  if (grid < 1) grid = 1;
  dpz_reduce_kernel<<<grid, 256, 0, st>>>(list, z, grad);
  MSF_LAUNCH_CHECK();
  OptCfg c{lr, beta1, beta2, eps, wd, grad_scale, max_norm, advance, 0, z.sigs[z.rank], z.world};
  const int owned = (list.live_units - z.rank + z.world - 1) / z.world + (list.total_units - list.live_units);
  int grid2 = owned < 592 ? owned : 592;
  if (grid2 > resident) grid2 = resident;
  if (grid2 < 1) grid2 = 1;
  opt_pack_kernel<false, true><<<grid2, 256, 0, st>>>(list, c, params, grad, exp_avg, exp_avg_sq, nullptr,
                                                      reinterpret_cast<unsigned long long*>(train_state),
                                                      z.arenas[z.rank], z);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

// owner[e] = rank that owns master-arena element e under the sharded step (unit % world), -1 for replicated elements
// (dead query / key slots, updated by every rank; vector slots are owned but their values are pushed to everyone).
int fusion_bf16_dpz_owner_map(const Layout& L, int world, signed char* owner_host) {
  OptList list;
  int rc = build_jobs(L, list);
  if (rc) return rc;
  for (long long e = 0; e < L.total; ++e) owner_host[e] = -1;
  for (int unit = 0; unit < list.live_units; ++unit) {
    int ji = 0;
    while (ji + 1 < list.count && unit >= list.j[ji + 1].unit_begin) ++ji;
    const OptJob& J = list.j[ji];
    int local = unit - J.unit_begin;
    const int b = local / J.units_per_batch;
    local -= b * J.units_per_batch;
    const long long base = J.begin + (long long)b * J.batch_stride;
    const signed char own = (signed char)(unit % world);
    if (J.kind == 2) {
      const int tc = (J.cols + 31) >> 5;
      const int r0 = (local / tc) << 5, c0 = (local % tc) << 5;
      for (int r = r0; r < r0 + 32 && r < J.rows; ++r)
        for (int cc = c0; cc < c0 + 32 && cc < J.cols; ++cc) owner_host[base + (long long)r * J.cols + cc] = own;
    } else {
      const int e0 = local * OPT_VEC_UNIT, e1 = J.cols < e0 + OPT_VEC_UNIT ? J.cols : e0 + OPT_VEC_UNIT;
      for (int e = e0; e < e1; ++e) owner_host[base + e] = -2 - own;   // vector slot: owner's moments, everybody's value
    }
  }
  return MSF_OK;
}

}  // namespace msf
