// chain3_kernel: the chained pair GEMMs of chain_gemm.cuh, rebuilt around the resource that bounded chain2_kernel:
// L2 -> SM operand traffic.  chain2_kernel streamed 1.15 MB per (128-window tile, outer modality) item from L2 (every
// CTA fetched every weight block itself, and the half-pair pipeline fetched each A block twice): 147 MB per launch at
// B = 4096, i.e. 23 us at the ~6.3 TB/s the L2 delivers to 128 SMs — the kernel's whole duration, at every batch size.
//
//   * A CLUSTER of C CTAs (C = 2 / 4 / 8) takes C consecutive 128-window tiles of ONE outer modality; all of them
//     need the same weight blocks, so each CTA fetches 1/C of every block (H/C rows) and TMA-multicasts it into the
//     same shared-memory offset of all C CTAs.  Weight traffic from L2 drops C-fold; the A operand is read once.
//   * Every tcgen05.mma has N = H (no N = 128 half-pair GEMM1 with its 95-cycle floor and doubled A reads).
//   * TMEM holds one intermediate T (H columns) and ACC (H columns).  The pipeline is k-block granular instead:
//     epilogue1 drains T one 64-column k-block at a time into a 3-slot ring of 16 KB staging blocks; GEMM2's k-block j
//     is issued as soon as staging block j is written, and GEMM1 of the NEXT pair is issued ahead of GEMM2's last
//     k-block (T is drained by then), so the tensor pipe works under the epilogue warps and the epilogue warps work
//     under the tensor pipe:   G1(0) | e(0) ‖ G2(0)[0..KB-2] | G1(1) G2(0)[KB-1] | e(1) ‖ G2(1)[0..KB-2] | ...
//   * Operands arrive through two rings with their own producer threads: 4 weight slots (H x 64 bf16, filled by
//     multicast from the whole cluster) and 3 A slots (128 x 64 bf16, private).  The first weight blocks are
//     requested BEFORE griddepcontrol.wait: weights were written at least two kernels earlier (msf_common.cuh).
//   * TMEM loads of the epilogue are software-pipelined (the next 32 columns are in flight while the current 32 are
//     converted), staging-block TMA stores are released one block late (wait_group.read 1).
//
// Roles: warp 0 weight producer, warp 1 MMA issuer, warp 2 TMEM allocator + A producer, warp 3 TMA store,
// warps 4..11 epilogue (thread = accumulator row, two column groups).  All mbarrier waits are bounded.
// Cross-CTA signalling is hardware-only: multicast TMA completes bytes on every destination CTA's "full" barrier,
// tcgen05.commit.multicast arrives on every CTA's "empty" barrier (count C) once this CTA's MMAs have read a slot.
#include "chain_gemm.cuh"

#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int C3_THREADS = 384;
constexpr int C3_EPI_WARPS = 8;
#ifndef C3_CFG_W
#define C3_CFG_W 4
#define C3_CFG_A 3
#define C3_CFG_U 3
#endif
constexpr int C3_W_SLOTS = C3_CFG_W, C3_A_SLOTS = C3_CFG_A, C3_U_SLOTS = C3_CFG_U;
constexpr uint32_t C3_BLK = 128 * 64 * 2;  // one k-block of a 128-row operand: 128 rows x 64 bf16
constexpr size_t C3_SMEM_LIMIT = 232448;
constexpr int C3_NBAR = (2 * C3_W_SLOTS + 2 * C3_A_SLOTS + 3 * C3_U_SLOTS + 7 + 1) & ~1;   // even: the bias row behind the barriers is read as float4

struct Ring {
  int slot, n;
  uint32_t phase;
  __device__ __forceinline__ explicit Ring(int n_) : slot(0), n(n_), phase(0u) {}
  __device__ __forceinline__ void next() {
    if (++slot == n) { slot = 0; phase ^= 1u; }
  }
};

__device__ __forceinline__ uint32_t c3_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void c3_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one slice of a weight block -> the same shared-memory offset of every CTA in `mask`; each destination CTA's
// barrier at the same offset receives the bytes
__device__ __forceinline__ void c3_tma_load_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "h"(mask)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once all MMAs issued so far have completed
__device__ __forceinline__ void c3_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void c3_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// issue only: the registers are not valid until c3_tmem_wait32() on the same array
__device__ __forceinline__ void c3_tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// waits for every outstanding tcgen05.ld of this thread; the "+r" operands make later uses of r depend on it
__device__ __forceinline__ void c3_tmem_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
  asm volatile(""
               : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// 32 fp32 values of this thread's row -> 32 TMEM columns (complete after c3_tmem_wait_st)
__device__ __forceinline__ void c3_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void c3_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Raw32 {
  uint4 q[4];
};
__device__ __forceinline__ Raw32 ld_row32(const __nv_bfloat16* p, bool ok) {
  Raw32 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.q[i] = make_uint4(0u, 0u, 0u, 0u);
  if (ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r.q[i] = __ldg(reinterpret_cast<const uint4*>(p) + i);
  }
  return r;
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// 8 results of row `trow`, columns [c8*8, c8*8+8) of a 64-column k-block -> K-major 128B-swizzled block
__device__ __forceinline__ void st_swz8(unsigned char* blk, int trow, int c8, const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  *reinterpret_cast<uint4*>(blk + trow * 128 + ((c8 ^ (trow & 7)) << 4)) = pk;
}

// wait-time accounting of CTA 0 (clock64 cycles; timeline builds only): msf_debug_chain_stamps()
//  [0] kernel start  [1] MMA: waits for operand slots  [2] MMA: waits for T drained  [3] MMA: waits for staged blocks
//  [4] MMA loop end  [5] epilogue warp 4: waits for G1  [6] epilogue: waits for a free staging block
//  [7] epilogue: first block start  [8] epilogue: waits for ACC  [9] epilogue end  [10] W producer: waits for free slots
//  [11] W producer end  [12] first operands landed (MMA thread)
#ifdef MSF_TIMELINE
__device__ long long g_chain3_stamps[32];
#define C3_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 128 && item_cnt == 0) g_chain3_stamps[i] = clock64(); } while (0)
#define C3_T0() const long long _t0 = clock64()
#define C3_ACC(i) do { dbg_acc[(i) & 3] += clock64() - _t0; } while (0)
#define C3_FLUSH(i) do { if (blockIdx.x == 0) g_chain3_stamps[i] = dbg_acc[(i) & 3]; } while (0)
#define C3_SET(i) do { if (blockIdx.x == 0) g_chain3_stamps[i] = clock64(); } while (0)
#else
#define C3_T0()
#define C3_ACC(i)
#define C3_FLUSH(i)
#define C3_SET(i)
#define C3_STAMP(i)
#endif

template <int MODE, int KBT>
__global__ void __launch_bounds__(C3_THREADS, 1) chain3_kernel(const __grid_constant__ ChainLaunch L) {
  TL_KERNEL(MODE);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t base = (off0 + 1023u) & ~1023u;   // identical in every CTA of the cluster (same kernel, same layout)
  constexpr int KB = KBT;                          // 64-column k-blocks of the hidden size: 2 or 4
  const int H = L.H, M = L.M, C = L.cluster;
  const uint32_t WB = (uint32_t)H * 128u;           // one weight k-block: H rows x 64 bf16
  const uint32_t w_base = base;
  const uint32_t a_base = w_base + C3_W_SLOTS * WB;
  const uint32_t u_base = a_base + C3_A_SLOTS * C3_BLK;
  const uint32_t bar_base = u_base + C3_U_SLOTS * C3_BLK;
  auto full_w = [&](int s) { return bar_base + 8u * s; };
  auto empty_w = [&](int s) { return bar_base + 8u * (C3_W_SLOTS + s); };
  auto full_a = [&](int s) { return bar_base + 8u * (2 * C3_W_SLOTS + s); };
  auto empty_a = [&](int s) { return bar_base + 8u * (2 * C3_W_SLOTS + C3_A_SLOTS + s); };
  auto u_full = [&](int s) { return bar_base + 8u * (2 * C3_W_SLOTS + 2 * C3_A_SLOTS + s); };              // block written
  auto u_free = [&](int s) { return bar_base + 8u * (2 * C3_W_SLOTS + 2 * C3_A_SLOTS + C3_U_SLOTS + s); }; // block reusable
  auto x_full = [&](int s) { return bar_base + 8u * (2 * C3_W_SLOTS + 2 * C3_A_SLOTS + 2 * C3_U_SLOTS + s); };            // aux block landed (TMA)
  const uint32_t bar_misc = bar_base + 8u * (2 * C3_W_SLOTS + 2 * C3_A_SLOTS + 3 * C3_U_SLOTS);
  const uint32_t t_full = bar_misc, t_empty = bar_misc + 8u, acc_full = bar_misc + 16u, acc_ready = bar_misc + 24u;
  const uint32_t tmem_slot = bar_misc + 32u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * C3_NBAR - off0));   // H floats, restaged per phase
  unsigned char* u_smem = smem_raw + (u_base - off0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (C > 1) ? c3_ctarank() : 0u;
  const uint16_t mc_mask = (uint16_t)((1u << C) - 1u);
  const int cluster_id = (int)blockIdx.x / C, n_clusters = (int)gridDim.x / C;
  const int tile_groups = (L.row_tiles + C - 1) / C;
  const int n_citems = tile_groups * L.n_active;    // cluster items: (tile group, outer modality)
  const uint32_t tmem_cols = (2 * H <= 256) ? 256u : 512u;
  constexpr int NX = (MODE == 1 ? 2 : 1) * KBT;   // aux k-blocks per item: the aux tile (+ backward: the P tile)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.map_a1);
    tma_prefetch_desc(&L.map_w1);
    tma_prefetch_desc(&L.map_w2);
    tma_prefetch_desc(&L.map_out1);
    tma_prefetch_desc(&L.map_out);
    if (MODE == 1) tma_prefetch_desc(&L.map_aux2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C3_W_SLOTS; ++s) {
      mbar_init(full_w(s), 1);
      mbar_init(empty_w(s), (uint32_t)C);   // one multicast commit arrival from every CTA of the cluster
    }
    for (int s = 0; s < C3_A_SLOTS; ++s) {
      mbar_init(full_a(s), 1);
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < C3_U_SLOTS; ++s) {
      mbar_init(u_full(s), C3_EPI_WARPS / 2);   // the 4 warps that own the block's k-block
      mbar_init(u_free(s), 2);              // GEMM2 (or the store warp a second time) + the store warp
      mbar_init(x_full(s), 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_empty, C3_EPI_WARPS);
    mbar_init(acc_full, 1);
    mbar_init(acc_ready, C3_EPI_WARPS);   // ACC holds the item's aux tile (and the previous item's ACC has been read)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (C > 1) c3_cluster_sync();   // every CTA's barriers exist before anything is multicast
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
#ifdef MSF_TIMELINE
  long long dbg_acc[4] = {0, 0, 0, 0};
  if (blockIdx.x == 0 && threadIdx.x == 0) g_chain3_stamps[0] = clock64();
#endif

  auto item_outer = [&](int ci) { return (int)L.active[ci % L.n_active]; };
  auto item_m0 = [&](int ci) { return ((ci / L.n_active) * C + (int)rank) * 128; };   // may lie beyond the batch: TMA
                                                                                      // zero-fills loads, clips stores
  if (warp == 0) {
    // =========================== weight producer ===========================
    // Does not wait for the predecessor grid: the weight arena was written at least two kernels ago.
    if (lane == 0) {
      Ring w(C3_W_SLOTS);
      const int slice_rows = H / C;
      const uint32_t slice_off = rank * (uint32_t)slice_rows * 128u;
      auto load_w = [&](const CUtensorMap* map, int kb, int zp) {
        { C3_T0(); mbar_wait(empty_w(w.slot), w.phase ^ 1u); C3_ACC(10); }
        mbar_expect_tx(full_w(w.slot), WB);
        const uint32_t dst = w_base + (uint32_t)w.slot * WB + slice_off;
        if (C > 1) c3_tma_load_mc(dst, map, kb * 64, (int)rank * slice_rows, zp, full_w(w.slot), mc_mask);
        else tma_load_3d(dst, map, kb * 64, 0, zp, full_w(w.slot));
        w.next();
      };
      for (int ci = cluster_id; ci < n_citems; ci += n_clusters) {
        const ChainOuter& O = L.outer[item_outer(ci)];
        const int n = O.n;
        if (n == 0) continue;
        for (int kb = 0; kb < KB; ++kb) load_w(&L.map_w1, kb, O.pair[0]);
        for (int i = 0; i < n; ++i) {      // the MMA issuer consumes in exactly this order
          for (int kb = 0; kb + 1 < KB; ++kb) load_w(&L.map_w2, kb, O.pair[i]);
          if (i + 1 < n)
            for (int kb = 0; kb < KB; ++kb) load_w(&L.map_w1, kb, O.pair[i + 1]);
          load_w(&L.map_w2, KB - 1, O.pair[i]);
        }
      }
      C3_FLUSH(10);
      C3_SET(11);
    }
  } else {
    pdl_wait();     TL_WAITED(MODE);  // the prologue overlapped the previous kernel; its outputs are visible from here on
    // The forward chain lets its successor (the head kernel) become resident now; the backward chain does not
    // (its successor's early blocks would squat the SMs the column sums beside it are meant to use).
    if (MODE != 1) pdl_launch();

    if (warp == 2) {
      // =========================== A producer ===============================
      if (lane == 0) {
        Ring a(C3_A_SLOTS);
        for (int ci = cluster_id; ci < n_citems; ci += n_clusters) {
          const ChainOuter& O = L.outer[item_outer(ci)];
          const int m0 = item_m0(ci);
          for (int i = 0; i < O.n; ++i)
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(empty_a(a.slot), a.phase ^ 1u);
              mbar_expect_tx(full_a(a.slot), C3_BLK);
              tma_load_3d(a_base + (uint32_t)a.slot * C3_BLK, &L.map_a1, kb * 64, m0, O.inner[i], full_a(a.slot));
              a.next();
            }
        }
      }
    } else if (warp == 1) {
      // =========================== MMA issuer ===============================
      if (lane == 0) {
        const uint32_t idesc = instr_desc(H, false, false);
        const uint32_t t_addr = tmem_base, acc_addr = tmem_base + (uint32_t)H;
        Ring w(C3_W_SLOTS), a(C3_A_SLOTS), u(C3_U_SLOTS);
        uint32_t t_cnt = 0, item_cnt = 0;
        auto free_w = [&](int s) {
          if (C > 1) c3_commit_mc(empty_w(s), mc_mask);
          else tc_commit(empty_w(s));
        };
        auto g1 = [&]() {   // T = A1 . W1^T, N = H
          { C3_T0(); mbar_wait(t_empty, (t_cnt & 1u) ^ 1u); C3_ACC(2); }
          ++t_cnt;
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            { C3_T0(); mbar_wait(full_a(a.slot), a.phase); mbar_wait(full_w(w.slot), w.phase); C3_ACC(1); }
            tc_fence_after();
#ifdef MSF_TIMELINE
            if (t_cnt == 1 && kb == 0) C3_SET(12);
#endif
            const uint32_t a_addr = a_base + (uint32_t)a.slot * C3_BLK, b_addr = w_base + (uint32_t)w.slot * WB;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(t_addr, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_a(a.slot));
            free_w(w.slot);
            a.next();
            w.next();
          }
          tc_commit(t_full);
        };
        auto g2 = [&]() {   // ACC += staged block . W2[:, k-block]^T, N = H (ACC starts as the item's aux tile)
          { C3_T0(); mbar_wait(u_full(u.slot), u.phase); C3_ACC(3); }
          { C3_T0(); mbar_wait(full_w(w.slot), w.phase); C3_ACC(1); }
          tc_fence_after();
          const uint32_t a_addr = u_base + (uint32_t)u.slot * C3_BLK, b_addr = w_base + (uint32_t)w.slot * WB;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(acc_addr, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                        1u);
          free_w(w.slot);
          tc_commit(u_free(u.slot));   // 1 of 2: GEMM2 no longer reads the staging block
          w.next();
          u.next();
        };
        for (int ci = cluster_id; ci < n_citems; ci += n_clusters, ++item_cnt) {
          const int n = L.outer[item_outer(ci)].n;
          for (int x = 0; x < NX; ++x) u.next();      // the item's aux blocks pass through the staging ring first
          if (n > 0) {
            g1();
            for (int i = 0; i < n; ++i) {
              if (i == 0) {   // the epilogue warps have read the previous item's ACC and written this item's aux tile
                mbar_wait(acc_ready, item_cnt & 1u);
                tc_fence_after();
              }
              for (int kb = 0; kb + 1 < KB; ++kb) g2();
              if (i + 1 < n) g1();      // T is drained before the last staging block is written
              g2();
            }
          }
          tc_commit(acc_full);
          for (int kb = 0; kb < KB; ++kb) u.next();   // the final tile's staging blocks have no GEMM2 reader
        }
        C3_FLUSH(1); C3_FLUSH(2); C3_FLUSH(3);
        C3_SET(4);
      }
    } else if (warp == 3) {
      // =========================== TMA store ================================
      if (lane == 0) {
        Ring u(C3_U_SLOTS);
        uint32_t item_cnt = 0;
        int pend = -1, pend_times = 0;   // staging block whose store has been committed but not yet released
        auto release_pending = [&]() {
          for (int t = 0; t < pend_times; ++t) mbar_arrive(u_free(pend));
          pend = -1;
        };
        for (int ci = cluster_id; ci < n_citems; ci += n_clusters, ++item_cnt) {
          const int o = item_outer(ci), m0 = item_m0(ci);
          const ChainOuter& O = L.outer[o];
          // the item's aux tile (forward P_q, backward dS_k, then P_k) streams through the staging ring: the
          // epilogue warps move it into ACC / reduce it to ReLU bits while the first GEMM1 is being fetched
          if (pend >= 0) { tma_store_wait_read(); release_pending(); }   // never wait for a slot this thread still holds
          for (int x = 0; x < NX; ++x) {
            mbar_wait(u_free(u.slot), u.phase ^ 1u);
            mbar_expect_tx(x_full(u.slot), C3_BLK);
            tma_load_3d(u_base + (uint32_t)u.slot * C3_BLK, (MODE == 1 && x >= KB) ? &L.map_aux2 : &L.map_a1,
                        (x % KB) * 64, m0, o, x_full(u.slot));
            u.next();
          }
          // A parity wait cannot tell "two phases ago": do not look at a slot's "written" barrier before the
          // aux use of that slot has been closed by the epilogue warps (they do that before signalling acc_ready).
          mbar_wait(acc_ready, item_cnt & 1u);
          for (int i = 0; i < O.n; ++i)
            for (int kb = 0; kb < KB; ++kb) {
              mbar_wait(u_full(u.slot), u.phase);
              if (L.store1) {
                tma_store_3d(&L.map_out1, u_base + (uint32_t)u.slot * C3_BLK, kb * 64, m0, O.pair[i]);
                tma_store_commit();
                if (pend >= 0) { c3_store_wait_read1(); release_pending(); }
                pend = u.slot; pend_times = 1;
              } else {
                mbar_arrive(u_free(u.slot));   // 2 of 2
              }
              u.next();
            }
          for (int kb = 0; kb < KB; ++kb) {    // the item's output tile, k-block by k-block
            mbar_wait(u_full(u.slot), u.phase);
            tma_store_3d(&L.map_out, u_base + (uint32_t)u.slot * C3_BLK, kb * 64, m0, o);
            tma_store_commit();
            if (pend >= 0) { c3_store_wait_read1(); release_pending(); }
            pend = u.slot; pend_times = 2;     // no GEMM2 reader: both arrivals come from here
            u.next();
          }
        }
        if (pend >= 0) { tma_store_wait_read(); release_pending(); }
        tma_store_wait_all();
      }
    } else if (warp >= 4) {
      // =========================== epilogue =================================
      // Warp (lq, cg) owns rows [32 lq, 32 lq + 32) of the k-blocks kb = cg, cg + 2: whole 64-column staging blocks,
      // so a block is signalled by 4 warps and every warp runs NB = KB / 2 block hand-offs per pass.  Everything a
      // pass needs from global memory (attention gates, aux rows, ReLU masks) is fetched BEFORE the wait for the
      // tensor pipe; the pass itself only touches TMEM, registers and shared memory.
      constexpr int NB = KBT / 2;        // k-blocks per warp per pass
      constexpr int NG = 2 * NB;         // 32-column groups per thread per pass
      const DropCfg drop = resolve_drop(L.drop);
      const int lq = warp & 3, cg = (warp - 4) >> 2;
      const int trow = lq * 32 + lane;                 // row inside the tile = TMEM lane
      const int et = (int)threadIdx.x - 128;
      const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
      auto grp_kb = [&](int j) { return cg + 2 * (j >> 1); };                  // k-block of group j
      auto grp_col = [&](int j) { return grp_kb(j) * 64 + (j & 1) * 32; };     // first column of group j in the tile
      uint32_t blk0 = 0;                 // staging blocks handed off before this pass (all roles count the same way)
      uint32_t xpar = 0u;                // per staging slot: parity of its next aux-block (TMA) use
      uint32_t t_cnt = 0, item_cnt = 0;
      uint32_t r0[32], r1[32];
      for (int ci = cluster_id; ci < n_citems; ci += n_clusters, ++item_cnt) {
        const int o = item_outer(ci), m0 = item_m0(ci);
        const ChainOuter& O = L.outer[o];
        const int n = O.n;
        const long long row = (long long)m0 + trow;
        const bool row_ok = row < L.rows;

        // ---- aux blocks: the item's aux tile (forward P_q, backward dS_k) goes straight into the ACC columns of
        // TMEM (GEMM2 accumulates on top of it; the final pass needs no global loads), backward also the P tile,
        // reduced to one ReLU bit per column.  The store warp streams them through the staging ring by TMA; a
        // thread reads its row of a block with conflict-free 16-byte shared-memory loads (128B swizzle).  This
        // thread's cells of ACC were last read by this thread (previous item's final pass): program order suffices.
        uint32_t relu_bits[NG];
#pragma unroll
        for (int x = 0; x < NX; ++x) {
          const uint32_t idx = blk0 + (uint32_t)x, slot = idx % C3_U_SLOTS;
          const uint32_t par = (xpar >> slot) & 1u;
          xpar ^= 1u << slot;
          const int kb = x % KBT;
          if ((kb & 1) != cg) continue;                 // the other column group's block
          mbar_wait(x_full(slot), par);
          const unsigned char* xb = u_smem + slot * C3_BLK + trow * 128;
          uint4 ch[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) ch[c] = *reinterpret_cast<const uint4*>(xb + ((c ^ (trow & 7)) << 4));
          if (x < KBT) {   // aux tile -> fp32 -> ACC columns [kb*64, kb*64 + 64)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const uint32_t w[4] = {ch[4 * h + c8].x, ch[4 * h + c8].y, ch[4 * h + c8].z, ch[4 * h + c8].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  r0[8 * c8 + 2 * e] = row_ok ? (w[e] << 16) : 0u;              // bf16 -> fp32 bit patterns
                  r0[8 * c8 + 2 * e + 1] = row_ok ? (w[e] & 0xffff0000u) : 0u;
                }
              }
              c3_tmem_st32(tmem_base + lane_base + (uint32_t)(H + kb * 64 + 32 * h), r0);
            }
          } else {         // P tile -> ReLU bits of groups 2*(kb>>1), 2*(kb>>1) + 1
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t bits = 0u;
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const uint32_t w[4] = {ch[4 * h + c8].x, ch[4 * h + c8].y, ch[4 * h + c8].z, ch[4 * h + c8].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  bits |= (bf_lo(w[e]) > 0.0f ? 1u : 0u) << (8 * c8 + 2 * e);
                  bits |= (bf_hi(w[e]) > 0.0f ? 1u : 0u) << (8 * c8 + 2 * e + 1);
                }
              }
              relu_bits[2 * (kb >> 1) + h] = bits;
            }
          }
          // the row has been consumed (its values were used above): once the block's four reader warps are
          // through, the slot goes back to the ring (both arrivals: an aux block has no GEMM2 / store reader)
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cg) : "memory");
          if (lq == 0 && lane == 0) {
            // every use of a slot completes one phase of BOTH its barriers (the ring positions of all roles carry
            // one parity per lap): close the "written" phase nobody waits for, then free the slot
            for (int t = 0; t < C3_EPI_WARPS / 2; ++t) mbar_arrive(u_full(slot));
            mbar_arrive(u_free(slot));
            mbar_arrive(u_free(slot));
          }
        }
        blk0 += (uint32_t)NX;
        c3_tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_ready);
        C3_STAMP(16);

        for (int i = 0; i < n; ++i) {
          const int zp = O.pair[i];
          if (MODE == 0) {   // value_proj bias of this pair -> shared memory (overlaps GEMM1)
            asm volatile("bar.sync 1, %0;" ::"n"(32 * C3_EPI_WARPS) : "memory");   // previous readers are done
            const float* bsrc = L.bias1[zp];
            for (int e = et; e < H; e += 32 * C3_EPI_WARPS) bias_s[e] = bsrc ? __ldg(bsrc + e) : 0.0f;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * C3_EPI_WARPS) : "memory");
          }
          // attention gates of this thread's 16-column groups (a group lies inside one head), before the wait
          float gate[2 * NG];
          {
            float gate_mask = 1.0f;
            if (MODE == 0 && L.mask != nullptr && row_ok) gate_mask = __ldg(L.mask + row * M + O.mask_col[i]);
            const int sub = O.sub[i];
            float* gate_out = (MODE == 0 && L.gate_out) ? L.gate_out + ((long long)zp * L.rows + row) * L.heads : nullptr;
            const float* gate_in = (MODE == 1) ? L.gate_in + ((long long)zp * L.rows + row) * L.heads : nullptr;
            int cur_head = -1;
            float cur_gate = 0.0f;
#pragma unroll
            for (int g = 0; g < 2 * NG; ++g) {
              const int ca = grp_col(g >> 1) + 16 * (g & 1);
              const int head = L.head_shift >= 0 ? (ca >> L.head_shift) : ca / L.head_dim;
              if (head != cur_head) {
                cur_head = head;
                if (MODE == 0) {
                  cur_gate = (gate_mask != 0.0f) ? 1.0f : 0.0f;
                  if (drop.active) cur_gate *= drop1(drop, SITE_ATTN, sub, row, head);
                  if (gate_out != nullptr && row_ok && ca == head * L.head_dim) gate_out[head] = cur_gate;
                } else {
                  cur_gate = row_ok ? __ldg(gate_in + head) : 0.0f;
                }
              }
              gate[g] = cur_gate;
            }
          }

          // 32 columns of T (registers) -> half of a staging block
          auto half1 = [&](uint32_t (&acc)[32], int j) {
            const uint32_t idx = blk0 + (uint32_t)grp_kb(j), slot = idx % C3_U_SLOTS, phase = (idx / C3_U_SLOTS) & 1u;
            if ((j & 1) == 0) { C3_T0(); mbar_wait(u_free(slot), phase ^ 1u); C3_ACC(6); }
            unsigned char* ub = u_smem + slot * C3_BLK;
            const int cb = grp_col(j);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              float v[8];
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (MODE == 0) {
                b0 = *reinterpret_cast<const float4*>(bias_s + cb + 8 * c8);
                b1 = *reinterpret_cast<const float4*>(bias_s + cb + 8 * c8 + 4);
              }
              const float g = gate[2 * j + (c8 >> 1)];
              v[0] = (__uint_as_float(acc[8 * c8 + 0]) + b0.x) * g;
              v[1] = (__uint_as_float(acc[8 * c8 + 1]) + b0.y) * g;
              v[2] = (__uint_as_float(acc[8 * c8 + 2]) + b0.z) * g;
              v[3] = (__uint_as_float(acc[8 * c8 + 3]) + b0.w) * g;
              v[4] = (__uint_as_float(acc[8 * c8 + 4]) + b1.x) * g;
              v[5] = (__uint_as_float(acc[8 * c8 + 5]) + b1.y) * g;
              v[6] = (__uint_as_float(acc[8 * c8 + 6]) + b1.z) * g;
              v[7] = (__uint_as_float(acc[8 * c8 + 7]) + b1.w) * g;
              st_swz8(ub, trow, (j & 1) * 4 + c8, v);
            }
            if (j & 1) {   // the warp's rows of this block are complete
              fence_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(u_full(slot));
            }
          };

          { C3_T0(); mbar_wait(t_full, t_cnt & 1u); C3_ACC(5); }
          ++t_cnt;
          tc_fence_after();
#ifdef MSF_TIMELINE
          if (threadIdx.x == 128 && t_cnt == 1) C3_SET(7);
          if (i < 3) C3_STAMP(17 + 2 * i);
#endif
          const uint32_t t_col = tmem_base + lane_base;
          c3_tmem_ld32_issue(t_col + (uint32_t)grp_col(0), r0);
#pragma unroll
          for (int j = 0; j < NG; ++j) {
            if (j & 1) {
              c3_tmem_wait32(r1);
              if (j + 1 < NG) c3_tmem_ld32_issue(t_col + (uint32_t)grp_col(j + 1), r0);
            } else {
              c3_tmem_wait32(r0);
              c3_tmem_ld32_issue(t_col + (uint32_t)grp_col(j + 1), r1);   // NG is even
            }
            if (j + 1 == NG) {   // T is in registers: GEMM1 of the next pair may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(t_empty);
            }
            if (j & 1) half1(r1, j);
            else half1(r0, j);
          }
          blk0 += (uint32_t)KBT;
#ifdef MSF_TIMELINE
          if (i < 3) C3_STAMP(18 + 2 * i);
#endif
        }

        // ---- final epilogue over ACC (columns [H, 2H) of TMEM), through the same staging ring ----
        if (MODE == 0) {   // summed out_proj biases of this item -> shared memory
          asm volatile("bar.sync 1, %0;" ::"n"(32 * C3_EPI_WARPS) : "memory");
          for (int e = et; e < H; e += 32 * C3_EPI_WARPS) {
            float val = 0.0f;
            for (int j = 0; j < n; ++j) {
              const float* bsrc = L.bias2[O.pair[j]];
              if (bsrc) val += __ldg(bsrc + e);
            }
            for (int j = 0; j < O.nb; ++j) {
              const float* bsrc = L.bias2[O.bias_only[j]];
              if (bsrc) val += __ldg(bsrc + e);
            }
            bias_s[e] = val;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * C3_EPI_WARPS) : "memory");
        }
        float rscale = 1.0f;
        if (MODE == 0) rscale = L.inv_cnt[o] * ((L.mask != nullptr && row_ok) ? __ldg(L.mask + row * M + o) : 1.0f);
        auto half2 = [&](uint32_t (&acc)[32], int j) {
          const uint32_t idx = blk0 + (uint32_t)grp_kb(j), slot = idx % C3_U_SLOTS, phase = (idx / C3_U_SLOTS) & 1u;
          if ((j & 1) == 0) { C3_T0(); mbar_wait(u_free(slot), phase ^ 1u); C3_ACC(6); }
          unsigned char* ub = u_smem + slot * C3_BLK;
          const int cb = grp_col(j);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a_lo = __uint_as_float(acc[8 * c8 + 2 * e]), a_hi = __uint_as_float(acc[8 * c8 + 2 * e + 1]);
              if (MODE == 0) {
                const float2 bb = *reinterpret_cast<const float2*>(bias_s + cb + 8 * c8 + 2 * e);
                v[2 * e] = rscale != 0.0f ? (a_lo + bb.x) * rscale : 0.0f;       // masked row: exact 0
                v[2 * e + 1] = rscale != 0.0f ? (a_hi + bb.y) * rscale : 0.0f;
              } else {
                const uint32_t bits = relu_bits[j] >> (8 * c8 + 2 * e);
                v[2 * e] = a_lo * ((bits & 1u) ? L.scale : 0.0f);
                v[2 * e + 1] = a_hi * ((bits & 2u) ? L.scale : 0.0f);
              }
            }
            st_swz8(ub, trow, (j & 1) * 4 + c8, v);
          }
          if (j & 1) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(u_full(slot));
          }
        };

        { C3_T0(); mbar_wait(acc_full, item_cnt & 1u); C3_ACC(8); }   // n == 0: committed without MMAs, ACC = aux
        tc_fence_after();
        C3_STAMP(23);
        {
          const uint32_t a_col = tmem_base + lane_base + (uint32_t)H;
          c3_tmem_ld32_issue(a_col + (uint32_t)grp_col(0), r0);
#pragma unroll
          for (int j = 0; j < NG; ++j) {
            if (j & 1) {
              c3_tmem_wait32(r1);
              if (j + 1 < NG) c3_tmem_ld32_issue(a_col + (uint32_t)grp_col(j + 1), r0);
            } else {
              c3_tmem_wait32(r0);
              c3_tmem_ld32_issue(a_col + (uint32_t)grp_col(j + 1), r1);
            }
            if (j & 1) half2(r1, j);
            else half2(r0, j);
          }
        }
        blk0 += (uint32_t)KBT;
#ifdef MSF_TIMELINE
        if (threadIdx.x == 128) { C3_FLUSH(5); C3_FLUSH(6); C3_FLUSH(8); C3_SET(9); }
#endif
      }
    }
  }

  tc_fence_before();
  if (C > 1) c3_cluster_sync();   // no CTA exits while a peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

size_t chain3_smem(int H) {
  return 1024 + (size_t)C3_W_SLOTS * H * 128 + (size_t)(C3_A_SLOTS + C3_U_SLOTS) * C3_BLK + 8 * C3_NBAR + (size_t)H * 4;
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster,
                               Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace

bool chain3_eligible(int H, int M, int head_dim) {
  return (H == 128 || H == 256) && M >= 1 && M <= MSF_MAX_MODALITIES && head_dim >= 16 && head_dim % 16 == 0 &&
         chain3_smem(H) <= C3_SMEM_LIMIT;
}

// CTAs per cluster = 128-window tiles that share one fetch of every weight block.  MSF_CHAIN_CLUSTER = 1 / 2 / 4 / 8
// overrides; small batches use the smallest power of two that covers their tiles (a cluster's surplus CTAs run on
// out-of-range rows: loads are zero-filled, stores clipped).
int chain3_cluster(int H, long long rows) {
  int c = 4;
  if (const char* e = getenv("MSF_CHAIN_CLUSTER")) {
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) c = v;
  }
  const long long tiles = ceil_div(rows < 1 ? 1 : rows, 128);
  while (c > 1 && c / 2 >= tiles) c /= 2;
  while (c > 1 && (H / c) % 8 != 0) c /= 2;
  return c;
}

int chain3_launch(ChainLaunch& L, cudaStream_t stream, const char* label) {
  MSF_REQUIRE(chain3_eligible(L.H, L.M, L.head_dim), "chain3_gemm: hidden %d / modalities %d / head_dim %d not supported",
              L.H, L.M, L.head_dim);
  MSF_REQUIRE(L.rows >= 1, "chain3_gemm: empty batch");
  L.cluster = chain3_cluster(L.H, L.rows);
  L.row_tiles = (int)ceil_div(L.rows, 128);
  if (L.n_active <= 0) {   // default: every modality is an outer modality
    L.n_active = L.M;
    for (int m = 0; m < L.M; ++m) L.active[m] = (short)m;
  }
  L.items = L.row_tiles * L.n_active;
  L.stages = C3_W_SLOTS;
  const int C = L.cluster;
  const int n_citems = (int)ceil_div(L.row_tiles, C) * L.n_active;
  const size_t smem = chain3_smem(L.H);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // One cluster per (tile group, outer modality).  chain_launch only picks this kernel for launches of one wave:
  // the item loops are written for persistent CTAs, but both carrying the ring / barrier state from one item to the
  // next and launching more clusters than fit at once still fail on the device (chain_gemm.cu: chain_variant).
  const int clusters = n_citems;
  const int grid = clusters * C;
  if (prof_enabled()) {
    double pairs = 0.0;
    for (int i = 0; i < L.n_active; ++i) pairs += L.outer[L.active[i]].n;
    prof_begin(label, 2.0 * 2.0 * (double)L.rows * L.H * L.H * pairs, stream, prof_repeat());
  }
  // the kernel only reads its operands and overwrites its outputs: repeating it (profiling) changes nothing
  void (*kern)(const ChainLaunch) = nullptr;
  if (L.H == 256) kern = L.mode == 0 ? chain3_kernel<0, 4> : chain3_kernel<1, 4>;
  else kern = L.mode == 0 ? chain3_kernel<0, 2> : chain3_kernel<1, 2>;
  MSF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = prof_repeat(); rep > 0; --rep)
    MSF_CHECK_CUDA(launch_pdl_cluster(kern, dim3(grid), dim3(C3_THREADS), smem, stream, C, L));
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  return MSF_OK;
}

int chain3_debug_stamps(long long* out16) {
#ifdef MSF_TIMELINE
  MSF_CHECK_CUDA(cudaDeviceSynchronize());
  MSF_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_chain3_stamps, sizeof(long long) * 32));
  return MSF_OK;
#else
  (void)out16;
  set_error("chain3 stamps are compiled into the timeline build only (MSF_BUILD_VARIANT=timeline)");
  return MSF_E_UNSUPPORTED;
#endif
}

}  // namespace msf
