// cta_group::2 probe: D[M, 256] = A[M, K] . B[256, K]^T (bf16 in, fp32 out) with one CTA PAIR per 256-row tile.
// Each CTA of the pair stages its own 128 rows of A and HALF of the B tile (128 of the 256 rows); the leader
// CTA issues tcgen05.mma.cta_group::2 (M = 256 across the pair), which reads both CTAs' shared memory, so the
// B operand crosses each SM's shared-memory port only once per pair.  Not on the product path yet: this is the
// building block for the round-2 chain / weight-gradient kernels, kept testable through msf_debug_pair_gemm.
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace msf {
namespace {

constexpr int PG_THREADS = 256;     // warp 0 TMA, warp 1 MMA (leader CTA only), warp 2 TMEM, warp 3 idle, warps 4..7 epilogue
constexpr int PG_STAGES = 4;
constexpr uint32_t PG_A_BYTES = 128 * 64 * 2, PG_B_BYTES = 128 * 64 * 2;
constexpr uint32_t PG_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: rank 0 of the pair

struct PairLaunch {
  CUtensorMap map_a, map_b;
  float* d;
  int M, K;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes go to the LEADER CTA's mbarrier (both CTAs of the pair issue it)
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar & PG_PEER_MASK)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the same barrier offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PG_THREADS, 1)
    pair_gemm_kernel(const __grid_constant__ PairLaunch L) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t base = (off0 + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = a_base + PG_STAGES * PG_A_BYTES, bar_base = b_base + PG_STAGES * PG_B_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PG_STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * PG_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * PG_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile = blockIdx.x >> 1;                 // 256-row tile of the pair
  const int m0 = tile * 256 + (int)rank * 128;      // this CTA's 128 rows
  const int KB = L.K >> 6;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.map_a);
    tma_prefetch_desc(&L.map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // leader: one expect_tx arrival covering both CTAs' bytes
      mbar_init(empty_bar(s), 1);   // one multicast commit arrival per use, in each CTA
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();     // both CTAs' barriers are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {   // TMA producer (both CTAs): own A rows, own half of B; bytes are counted on the leader's barrier
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (PG_A_BYTES + PG_B_BYTES));
        tma_load_3d_2sm(a_base + stage * PG_A_BYTES, &L.map_a, kb * 64, m0, 0, full_bar(stage));
        tma_load_3d_2sm(b_base + stage * PG_B_BYTES, &L.map_b, kb * 64, (int)rank * 128, 0, full_bar(stage));
        if (++stage == PG_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // MMA issuer: leader CTA only
      // instruction descriptor: bf16 x bf16 -> fp32, K-major, N = 256, M = 256 (pair)
      uint32_t idesc = 0;
      idesc |= 1u << 4;
      idesc |= 1u << 7;
      idesc |= 1u << 10;
      idesc |= (uint32_t)(256 >> 3) << 17;
      idesc |= (uint32_t)(256 >> 4) << 24;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = a_base + stage * PG_A_BYTES, b_addr = b_base + stage * PG_B_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16_2sm(tmem_base, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit_2sm(empty_bar(stage));
        if (++stage == PG_STAGES) { stage = 0; phase ^= 1u; }
      }
      tc_commit_2sm(acc_full);
    }
  } else if (warp >= 4) {   // epilogue: thread = accumulator row of this CTA's half
    const int lq = warp & 3;
    const int row = m0 + lq * 32 + lane;
    mbar_wait(acc_full, 0u);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 256; c += 16) {
      uint32_t acc[16];
      tmem_ld16_issue(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c, acc);
      tmem_wait16(acc);
      if (row < L.M) {
        float4* dst = reinterpret_cast<float4*>(L.d + (long long)row * 256 + c);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                               __uint_as_float(acc[4 * q + 3]));
      }
    }
    tc_fence_before();
  }
  cluster_sync_all();     // neither CTA frees TMEM / exits while the pair's MMAs or the peer's reads are in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_debug_pair_gemm(const void* a_bf16, const void* b_bf16, float* d, int64_t m, int64_t k, void* stream) {
  using namespace msf;
  MSF_REQUIRE(a_bf16 && b_bf16 && d && m >= 1 && k >= 64 && k % 64 == 0, "msf_debug_pair_gemm: bad arguments");
  PairLaunch L;
  memset(&L, 0, sizeof(L));
  int rc;
  if ((rc = tc_encode_map(&L.map_a, a_bf16, m, k, k, 1, 0, 64, 128))) return rc;
  if ((rc = tc_encode_map(&L.map_b, b_bf16, 256, k, k, 1, 0, 64, 128))) return rc;
  L.d = d; L.M = (int)m; L.K = (int)k;
  const size_t smem = 1024 + PG_STAGES * (PG_A_BYTES + PG_B_BYTES) + 8 * (2 * PG_STAGES + 2);
  MSF_CHECK_CUDA(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (int)ceil_div(m, 256);
  pair_gemm_kernel<<<2 * tiles, PG_THREADS, smem, (cudaStream_t)stream>>>(L);
  MSF_LAUNCH_CHECK();
  return MSF_OK;
}

// ---------------------------------------------------------------------------
// Issue-rate probe: one thread per CTA issues `reps` tcgen05.mma (M = 128, N = n, K = 16, bf16, both operands
// resident in shared memory — no TMA, no epilogue) and reports clock64 cycles per MMA.  Separates what the
// tensor pipe can do from what the chained kernels achieve (profiles/README.md).
// ---------------------------------------------------------------------------
namespace msf {
namespace {
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int reps, int ksteps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 4 * 16384, bar = b_base + 4 * 32768, slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  for (uint32_t i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;   // bf16 small values
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = instr_desc(n, false, false);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int st = r & 3;   // walk over 4 resident operand stages like a ring
      for (int k = 0; k < ksteps; ++k)
        tc_mma_bf16(tmem + (uint32_t)((r & 1) * 256), smem_desc(a_base + st * 16384 + k * 32, 16, 1024),
                    smem_desc(b_base + st * 32768 + k * 32, 16, 1024), idesc, (k > 0) ? 1u : 0u);
    }
    tc_commit(bar);
    mbar_wait(bar, 0u);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
}  // namespace
}  // namespace msf

extern "C" int msf_debug_mma_rate(int32_t n, int32_t reps, int32_t ksteps, int32_t ctas, int64_t* cycles_out, void* stream) {
  using namespace msf;
  MSF_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && reps >= 1 && ksteps >= 1 && ksteps <= 4 && ctas >= 1 && cycles_out,
              "msf_debug_mma_rate: bad arguments");
  long long* dev = nullptr;
  MSF_CHECK_CUDA(cudaMalloc(&dev, sizeof(long long)));
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
  MSF_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(n, reps, ksteps, dev);
  MSF_LAUNCH_CHECK();
  MSF_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  long long v = 0;
  MSF_CHECK_CUDA(cudaMemcpy(&v, dev, sizeof(v), cudaMemcpyDeviceToHost));
  MSF_CHECK_CUDA(cudaFree(dev));
  *cycles_out = v;
  return MSF_OK;
}
