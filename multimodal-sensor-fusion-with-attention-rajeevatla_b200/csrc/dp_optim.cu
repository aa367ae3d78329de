// Data-parallel gradient reduction fused with clip + AdamW, over NVLink peer memory.
//
// One process per GPU; every rank holds its gradient arena, a staging arena, a "reduced" arena and a
// small signal block in symmetric (peer-mapped) memory.  Instead of a library all-reduce followed by
// an optimizer pass, two kernels do both, moving data with posted peer WRITES only (peer reads were
// measured at ~200 GB/s on this box; stores stream at link rate):
//
//   dp_reduce_kernel  phase 1  every rank pushes, for each peer r, its gradient values of r's slice of the
//                              live gradient elements into r's staging arena;
//                     phase 2  (after a cross-GPU flag barrier) rank r sums its slice over all ranks in fixed
//                              rank order (deterministic), pushes the sums into every rank's reduced arena
//                              and publishes the slice's square-norm;
//   dp_adamw_kernel            (after the second flag barrier) every rank forms the global norm from the
//                              published partials and applies clip + AdamW to its replica from its local
//                              reduced arena — identical inputs on every rank, so replicas stay bit-identical.
//
// The dead query/key slots (zero gradient on every rank) are neither sent nor reduced: 47 % of the
// arena never crosses NVLink.  Cross-GPU ordering uses system-scope release/acquire counters that
// carry a per-communicator epoch; every wait is bounded (trap instead of hang).
//   src/train.py:378-382,416-430 (AdamW + global-norm clip) under batch-sharded data parallelism.
#include <stdlib.h>
#include <string.h>

#include "dp_signals.cuh"
#include "msf_common.cuh"

namespace msf {
namespace {

constexpr int DP_MAX_RANKS = 8;
constexpr int DP_MAX_SEGS = 2 * MSF_MAX_MODALITIES * (MSF_MAX_MODALITIES - 1) + 2;

struct DpSeg {
  long long begin, count;   // arena range
  long long live_begin;     // position of its first element in the concatenated live index space (-1: dead)
};
struct DpArgs {
  DpSeg seg[DP_MAX_SEGS];
  int nseg;
  long long n_live, chunk;            // live elements, slice length per rank
  int rank, world;
  const float* grads[DP_MAX_RANKS];   // peer-mapped gradient arenas, rank order
  float* stages[DP_MAX_RANKS];        // peer-mapped staging arenas: [contributor rank][position in the owner's slice]
  float* reds[DP_MAX_RANKS];          // peer-mapped reduced arenas
  unsigned long long* sigs[DP_MAX_RANKS];
  const unsigned long long* train_state;  // {seed, offset, step}: step drives the Adam bias correction
};

// ---------------------------------------------------------------------------
// reduce-scatter + norm, push based
// ---------------------------------------------------------------------------
// Calls fn(arena_offset, live_position, count) for every maximal run of the live range [l0, l1) that is
// contiguous in the arena.
template <typename F>
__device__ __forceinline__ void for_each_run(const DpArgs& a, long long l0, long long l1, F fn) {
  for (int s = 0; s < a.nseg; ++s) {
    const long long lb = a.seg[s].live_begin;
    if (lb < 0) continue;
    const long long x0 = max(l0, lb), x1 = min(l1, lb + a.seg[s].count);
    if (x0 < x1) fn(a.seg[s].begin + (x0 - lb), x0, x1 - x0);
  }
}

__global__ void __launch_bounds__(256) dp_reduce_kernel(const __grid_constant__ DpArgs a) {
  unsigned long long* sig = a.sigs[a.rank];
  const unsigned long long epoch = sig[SIG_EPOCH] + 1ull;   // same on every rank: all ranks run the same steps
  if (blockIdx.x == 0 && threadIdx.x == 0) sig[SIG_TIME + 0] = gtime();
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const float* mine = a.grads[a.rank];
  __shared__ bool last;

  // ---- phase 1: push my values of every peer's slice into that peer's staging arena ----
  for (int r = 0; r < a.world; ++r) {
    if (r == a.rank) continue;
    const long long lo = a.chunk * r, hi = min(a.n_live, lo + a.chunk);
    float* dst = a.stages[r] + (long long)a.rank * a.chunk;
    for_each_run(a, lo, hi, [&](long long e0, long long l0, long long n) {
      const long long d0 = l0 - lo;
      if (((e0 | d0) & 3) == 0) {
        const long long n4 = n >> 2;
        for (long long i = t0; i < n4; i += stride)
          *reinterpret_cast<float4*>(dst + d0 + 4 * i) = __ldg(reinterpret_cast<const float4*>(mine + e0) + i);
        for (long long i = (n4 << 2) + t0; i < n; i += stride) dst[d0 + i] = __ldg(mine + e0 + i);
      } else {
        for (long long i = t0; i < n; i += stride) dst[d0 + i] = __ldg(mine + e0 + i);
      }
    });
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();   // cumulative: the block's pushes are visible system-wide before it reports in
    last = atomicAdd(sig + SIG_TICKET, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.sigs[threadIdx.x] + SIG_DONE_B + a.rank, epoch);   // "my contributions have landed"
  }
  wait_all(sig, SIG_DONE_B, a.world, epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) sig[SIG_TIME + 1] = gtime();

  // ---- phase 2: sum my slice in rank order, push the sums to every rank's reduced arena ----
  const long long lo = a.chunk * a.rank, hi = min(a.n_live, lo + a.chunk);
  const float* stage = a.stages[a.rank];
  double sq = 0.0;
  for_each_run(a, lo, hi, [&](long long e0, long long l0, long long n) {
    const long long d0 = l0 - lo;
    const bool vec = ((e0 | d0) & 3) == 0;
    const long long n4 = vec ? (n >> 2) : 0;
    for (long long i = t0; i < n4; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {   // fixed rank order: bit-reproducible
        const float4 g = (p == a.rank) ? __ldg(reinterpret_cast<const float4*>(mine + e0) + i)
                                       : ld_peer4(stage + (long long)p * a.chunk + d0 + 4 * i);
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      }
      for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4*>(a.reds[r] + e0 + 4 * i) = acc;
      sq += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) {
      float acc = 0.0f;
      for (int p = 0; p < a.world; ++p)
        acc += (p == a.rank) ? __ldg(mine + e0 + i) : ld_peer1(stage + (long long)p * a.chunk + d0 + i);
      for (int r = 0; r < a.world; ++r) a.reds[r][e0 + i] = acc;
      sq += (double)acc * acc;
    }
  });
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
    __threadfence_system();   // the block's pushed sums and its norm share are visible system-wide
    last = atomicAdd(sig + SIG_TICKET2, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < a.world) {   // the last block publishes: norm first, then the release flag
    __threadfence_system();
    const unsigned long long bits = ld_acquire_sys(sig + SIG_ACC);
    unsigned long long* peer = a.sigs[threadIdx.x];
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(peer + SIG_NORM + a.rank), "l"(bits) : "memory");
    __threadfence_system();
    st_release_sys(peer + SIG_DONE_R + a.rank, epoch);
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

// ---------------------------------------------------------------------------
// all-gather + clip + AdamW
// ---------------------------------------------------------------------------
struct DpAdam {
  float lr, beta1, beta2, eps, wd, grad_scale, max_norm;
};

__global__ void __launch_bounds__(256) dp_adamw_kernel(const __grid_constant__ DpArgs a, const DpAdam c,
                                                       float* __restrict__ p, float* __restrict__ m,
                                                       float* __restrict__ v) {
  unsigned long long* sig = a.sigs[a.rank];
  const unsigned long long epoch = sig[SIG_EPOCH] + 1ull;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sig[SIG_TIME + 3] = gtime();
  wait_all(sig, SIG_DONE_R, a.world, epoch);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) sig[SIG_TIME + 4] = gtime();
  double total = 0.0;
  for (int r = 0; r < a.world; ++r) total += __longlong_as_double((long long)ld_acquire_sys(sig + SIG_NORM + r));
  const double step = (double)a.train_state[2];
  const float bc1 = (float)(1.0 - pow((double)c.beta1, step));
  const float sqrt_bc2 = (float)sqrt(1.0 - pow((double)c.beta2, step));
  float gs = c.grad_scale;
  if (c.max_norm > 0.0f) gs *= fminf(c.max_norm / ((float)sqrt(total) * c.grad_scale + 1e-6f), 1.0f);
  const float lr = resolve_lr(c.lr, reinterpret_cast<const unsigned long long*>(a.train_state));
  const float step_size = lr / bc1, decay = 1.0f - lr * c.wd;

  const DpSeg sg = a.seg[blockIdx.y];
  const float* red = a.reds[a.rank];
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool vec = (sg.begin & 3) == 0 && (sg.live_begin < 0 || (sg.live_begin & 3) == 0);
  const long long n4 = vec ? (sg.count >> 2) : 0;
  if (sg.live_begin < 0) {  // dead slots: g = m = v = 0, only the decoupled weight decay acts
    float4* p4 = reinterpret_cast<float4*>(p + sg.begin);
    for (long long i = t0; i < n4; i += stride) {
      float4 x = p4[i];
      x.x *= decay; x.y *= decay; x.z *= decay; x.w *= decay;
      p4[i] = x;
    }
    for (long long i = (n4 << 2) + t0; i < sg.count; i += stride) p[sg.begin + i] *= decay;
  } else {
    for (long long i = t0; i < n4; i += stride) {
      const long long e = sg.begin + 4 * i;
      const float4 g = ld_peer4(red + e);   // pushed by the slice's owner before its release flag
      float4 x = *reinterpret_cast<float4*>(p + e), mm = *reinterpret_cast<float4*>(m + e),
             vv = *reinterpret_cast<float4*>(v + e);
      const float gx = g.x * gs, gy = g.y * gs, gz = g.z * gs, gw = g.w * gs;
      mm.x = c.beta1 * mm.x + (1.0f - c.beta1) * gx; vv.x = c.beta2 * vv.x + (1.0f - c.beta2) * gx * gx;
      mm.y = c.beta1 * mm.y + (1.0f - c.beta1) * gy; vv.y = c.beta2 * vv.y + (1.0f - c.beta2) * gy * gy;
      mm.z = c.beta1 * mm.z + (1.0f - c.beta1) * gz; vv.z = c.beta2 * vv.z + (1.0f - c.beta2) * gz * gz;
      mm.w = c.beta1 * mm.w + (1.0f - c.beta1) * gw; vv.w = c.beta2 * vv.w + (1.0f - c.beta2) * gw * gw;
      x.x = x.x * decay - step_size * (mm.x / (sqrtf(vv.x) / sqrt_bc2 + c.eps));
      x.y = x.y * decay - step_size * (mm.y / (sqrtf(vv.y) / sqrt_bc2 + c.eps));
      x.z = x.z * decay - step_size * (mm.z / (sqrtf(vv.z) / sqrt_bc2 + c.eps));
      x.w = x.w * decay - step_size * (mm.w / (sqrtf(vv.w) / sqrt_bc2 + c.eps));
      *reinterpret_cast<float4*>(p + e) = x;
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
    }
    for (long long i = (n4 << 2) + t0; i < sg.count; i += stride) {
      const long long e = sg.begin + i;
      const float gi = ld_peer1(red + e) * gs;
      const float mi = c.beta1 * m[e] + (1.0f - c.beta1) * gi;
      const float vi = c.beta2 * v[e] + (1.0f - c.beta2) * gi * gi;
      p[e] = p[e] * decay - step_size * (mi / (sqrtf(vi) / sqrt_bc2 + c.eps));
      m[e] = mi;
      v[e] = vi;
    }
  }
  // the last block to finish closes the epoch
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(sig + SIG_TICKET3, 1ull) == (unsigned long long)gridDim.x * gridDim.y - 1;
    if (last) {
      sig[SIG_TICKET3] = 0ull;
      sig[SIG_EPOCH] = epoch;
      sig[SIG_TIME + 5] = gtime();
    }
  }
}

}  // namespace
}  // namespace msf

namespace msf {
bool fusion_bf16_eligible(const Layout& L);
int fusion_bf16_opt_pack_dp(const Layout& L, float* params, const float* reduced, unsigned long long* sig, int world,
                            float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr, float beta1,
                            float beta2, float eps, float wd, float grad_scale, float max_norm, void* arena_v,
                            int advance, cudaStream_t st);
}

static int dp_step(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params, float* exp_avg,
                   float* exp_avg_sq, const uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float grad_scale, float max_norm, void* params_bf16, int advance,
                   void* stream);

namespace msf {
int fusion_bf16_opt_pack_dpz(const Layout& L, const msf_dpz_comm* comm, float* params, float* grad, float* exp_avg,
                             float* exp_avg_sq, uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                             float wd, float grad_scale, float max_norm, int advance, cudaStream_t st);
int fusion_bf16_dpz_owner_map(const Layout& L, int world, signed char* owner_host);
}

extern "C" int msf_dpz_optimizer_step_packed(const msf_fusion_shape* shape, const msf_dpz_comm* comm, float* params,
                                             float* grad, float* exp_avg, float* exp_avg_sq, uint64_t* train_state,
                                             float lr, float beta1, float beta2, float eps, float weight_decay,
                                             float grad_scale, float max_norm, int32_t advance_state, void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(comm && params && grad && exp_avg && exp_avg_sq && train_state, "msf_dpz_optimizer_step_packed: null pointer");
  MSF_REQUIRE(comm->world >= 1 && comm->world <= msf::DP_MAX_RANKS && comm->rank >= 0 && comm->rank < comm->world,
              "msf_dpz_optimizer_step_packed: rank %d / world %d out of range", comm->rank, comm->world);
  if (!msf::fusion_bf16_eligible(L)) {
    msf::set_error("shape not eligible for the tensor-core path");
    return MSF_E_UNSUPPORTED;
  }
  return msf::fusion_bf16_opt_pack_dpz(L, comm, params, grad, exp_avg, exp_avg_sq, train_state, lr, beta1, beta2, eps,
                                       weight_decay, grad_scale, max_norm, advance_state, (cudaStream_t)stream);
}

extern "C" int msf_dpz_owner_map(const msf_fusion_shape* shape, int32_t world, int8_t* owner_host) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(owner_host != nullptr && world >= 1 && world <= msf::DP_MAX_RANKS, "msf_dpz_owner_map: bad arguments");
  return msf::fusion_bf16_dpz_owner_map(L, world, reinterpret_cast<signed char*>(owner_host));
}

extern "C" int msf_dp_optimizer_step(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params,
                                     float* exp_avg, float* exp_avg_sq, const uint64_t* train_state, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                     float max_norm, void* stream) {
  return dp_step(shape, comm, params, exp_avg, exp_avg_sq, train_state, lr, beta1, beta2, eps, weight_decay,
                 grad_scale, max_norm, nullptr, 0, stream);
}

extern "C" int msf_dp_optimizer_step_packed(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params,
                                            float* exp_avg, float* exp_avg_sq, uint64_t* train_state, float lr,
                                            float beta1, float beta2, float eps, float weight_decay,
                                            float grad_scale, float max_norm, void* params_bf16,
                                            int32_t advance_state, void* stream) {
  MSF_REQUIRE(params_bf16 != nullptr, "msf_dp_optimizer_step_packed: null params_bf16");
  return dp_step(shape, comm, params, exp_avg, exp_avg_sq, train_state, lr, beta1, beta2, eps, weight_decay,
                 grad_scale, max_norm, params_bf16, advance_state, stream);
}

static int dp_step(const msf_fusion_shape* shape, const msf_dp_comm* comm, float* params, float* exp_avg,
                   float* exp_avg_sq, const uint64_t* train_state, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float grad_scale, float max_norm, void* params_bf16, int advance,
                   void* stream) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(comm && params && exp_avg && exp_avg_sq && train_state, "msf_dp_optimizer_step: null pointer");
  MSF_REQUIRE(comm->world >= 1 && comm->world <= msf::DP_MAX_RANKS && comm->rank >= 0 && comm->rank < comm->world,
              "msf_dp_optimizer_step: rank %d / world %d out of range", comm->rank, comm->world);
  msf::DpArgs a;
  memset(&a, 0, sizeof(a));
  long long live = 0;
  auto add = [&](long long begin, long long count, bool dead) {
    if (count <= 0) return;
    a.seg[a.nseg].begin = begin;
    a.seg[a.nseg].count = count;
    a.seg[a.nseg].live_begin = dead ? -1 : live;
    if (!dead) live += count;
    ++a.nseg;
  };
  const long long half = 2 * ((long long)L.H * L.H + L.H);
  add(0, L.pair_base, false);
  for (int p = 0; p < L.num_pairs(); ++p) {
    add(L.pair_base + p * L.pair_stride, half, true);
    add(L.pair_base + p * L.pair_stride + half, half, false);
  }
  const long long tail = L.pair_base + (long long)L.num_pairs() * L.pair_stride;
  add(tail, L.total - tail, false);
  a.n_live = live;
  a.chunk = (msf::ceil_div(live, comm->world) + 3) / 4 * 4;
  a.rank = comm->rank;
  a.world = comm->world;
  for (int r = 0; r < comm->world; ++r) {
    MSF_REQUIRE(comm->grads[r] && comm->stages[r] && comm->reds[r] && comm->sigs[r],
                "msf_dp_optimizer_step: null peer pointer (rank %d)", r);
    a.grads[r] = comm->grads[r];
    a.stages[r] = comm->stages[r];
    a.reds[r] = comm->reds[r];
    a.sigs[r] = reinterpret_cast<unsigned long long*>(comm->sigs[r]);
  }
  a.train_state = reinterpret_cast<const unsigned long long*>(train_state);
  cudaStream_t st = (cudaStream_t)stream;
  // Every block waits inside the kernel for the peers' flags, and those depend on ALL blocks of every rank
  // having pushed: the grid must be co-resident (148 SMs x 8 blocks of 256 threads) or the ranks deadlock.
  // The bound comes from the device this call runs on (occupancy x SM count), not from a fixed part.
  static int resident[64] = {0};   // co-resident blocks of dp_reduce_kernel per device
  int dev = 0;
  MSF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && resident[dev] == 0) {
    int per_sm = 0, sms = 0;
    MSF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msf::dp_reduce_kernel, 256, 0));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    resident[dev] = per_sm * sms;
  }
  const int cap = (dev >= 0 && dev < 64) ? resident[dev] : 0;
  if (cap < 1) {
    msf::set_error("msf_dp_optimizer_step: the exchange kernel cannot be made co-resident on this device");
    return MSF_E_UNSUPPORTED;
  }
  long long blocks = msf::ceil_div(msf::ceil_div(a.n_live, 4), 256);
  const long long lo = cap < 148 ? cap : 148, hi = cap < 592 ? cap : 592;
  if (blocks < lo) blocks = lo;
  if (blocks > hi) blocks = hi;
  msf::dp_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  MSF_LAUNCH_CHECK();
  if (params_bf16 != nullptr) {   // clip + AdamW + bf16 re-pack + state advance in one launch
    if (!msf::fusion_bf16_eligible(L)) {
      msf::set_error("shape not eligible for the tensor-core path");
      return MSF_E_UNSUPPORTED;
    }
    return msf::fusion_bf16_opt_pack_dp(L, params, a.reds[a.rank], a.sigs[a.rank], a.world, exp_avg, exp_avg_sq,
                                        const_cast<uint64_t*>(train_state), lr, beta1, beta2, eps, weight_decay,
                                        grad_scale, max_norm, params_bf16, advance, st);
  }
  msf::DpAdam c{lr, beta1, beta2, eps, weight_decay, grad_scale, max_norm};
  dim3 grid(48, (unsigned)a.nseg);
  msf::dp_adamw_kernel<<<grid, 256, 0, st>>>(a, c, params, exp_avg, exp_avg_sq);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}
