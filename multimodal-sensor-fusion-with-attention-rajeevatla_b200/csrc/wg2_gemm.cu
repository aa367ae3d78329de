// wg2_kernel: every weight gradient of a training pass, dW[M, N] = dY^T . X with the contraction over the windows,
// on CTA PAIRS.  The single-CTA grouped GEMM (tc_gemm_kernel<1>, 128 x 128 tiles) was bound by the L2: every tile
// streams (128 + 128) x rows operand elements, 220 MB per launch at B = 4096 for 14.5 GFLOP (~8 TB/s, 28 us).
//
//   * A cluster of two CTAs owns one 256 x 256 tile of dW (tcgen05.mma.cta_group::2, M = 256 over the pair): each
//     CTA stages its own 128 columns of dY and HALF of the X columns, so a 256 x 256 gradient reads each operand
//     element exactly once: 4 MB per matrix instead of 8.
//   * The contraction is split in two halves over two clusters (60 clusters of work for 30 matrices on 74 CTA
//     pairs).  The halves meet in the destination: each cluster finishes one half of the tile's columns and
//     sends its partial sums of the other half (stored into the destination, then a flag per (tile, CTA, lane
//     quarter, column half)); the finisher waits for the flag, adds its own sums and stores the result.  Two
//     addends commute, so the result is bit-reproducible, the destination needs no clearing, and the finisher
//     sees the finished values: it also adds their sum of squares to the gradient norm the optimizer is given.
//   * Operands are MN-major (the windows are the rows of both operands): 64 (k) x 64 (mn) boxes, 128B swizzle, the
//     same shared-memory layout and descriptors as tc_gemm_kernel<1>.  5 stages of 32 KB per CTA.
//
// Roles per CTA: warp 0 TMA producer (own operand halves; bytes are counted on the leader's barrier), warp 1 MMA
// issuer (leader CTA only), warp 2 TMEM allocator, warps 4..11 epilogue (thread = accumulator row; 32-column
// chunks through a padded staging block so that global traffic is whole 128-byte rows).  Every wait is bounded.
#include "wg2_gemm.cuh"

#include <stdlib.h>

#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int WG2_THREADS = 384;
constexpr int WG2_EPI_WARPS = 8;
constexpr int WG2_STAGES = 5;
constexpr uint32_t WG2_A_BYTES = 128 * 64 * 2;   // this CTA's 128 dY columns x 64 windows
constexpr uint32_t WG2_B_BYTES = 128 * 64 * 2;   // up to 128 X columns x 64 windows
constexpr int WG2_STG_PITCH = 32 * 4 + 16;       // one staged row: 32 fp32 + 16 B (conflict-free float4 rows)
constexpr uint32_t WG2_STG_BYTES = WG2_EPI_WARPS * 32 * WG2_STG_PITCH;
constexpr uint32_t WG2_PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the pair's leader
constexpr size_t WG2_SMEM = 1024 + WG2_STAGES * (WG2_A_BYTES + WG2_B_BYTES) + WG2_STG_BYTES + 8 * (2 * WG2_STAGES + 4);

__device__ unsigned g_wg2_flags[WG2_MAX_FLAG_TILES * 16];

__device__ __forceinline__ uint32_t wg2_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void wg2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory whose bytes are counted on the LEADER's mbarrier
__device__ __forceinline__ void wg2_tma_load(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar & WG2_PEER_MASK)
      : "memory");
}
__device__ __forceinline__ void wg2_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void wg2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3)
               : "memory");
}
__device__ __forceinline__ void wg2_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & WG2_PEER_MASK) : "memory");
}
__device__ __forceinline__ unsigned wg2_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct Wg2Tile {
  int problem, tm, tn;
};
__device__ __forceinline__ Wg2Tile wg2_locate(const Wg2Launch& L, int tile) {
  int pi = 0;
  while (pi + 1 < L.count && tile >= L.p[pi + 1].tile_begin) ++pi;
  const int local = tile - L.p[pi].tile_begin;
  const int tiles_n = (L.p[pi].N + 255) >> 8;
  Wg2Tile t;
  t.problem = pi;
  t.tm = local / tiles_n;
  t.tn = local - t.tm * tiles_n;
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG2_THREADS, 1)
    wg2_kernel(const __grid_constant__ Wg2Launch L) {
  TL_KERNEL(0);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t base = (off0 + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = a_base + WG2_STAGES * WG2_A_BYTES;
  const uint32_t stg_base = b_base + WG2_STAGES * WG2_B_BYTES;
  const uint32_t bar_base = stg_base + WG2_STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG2_STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * WG2_STAGES);
  const uint32_t acc_empty = bar_base + 8u * (2 * WG2_STAGES + 1);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WG2_STAGES + 2);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  unsigned char* stage_smem = smem_raw + (stg_base - off0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = wg2_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int units = L.total_tiles * L.splits;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < L.nmaps; ++i) tma_prefetch_desc(&L.maps[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // leader: one expect_tx arrival covering both CTAs' bytes
      mbar_init(empty_bar(s), 1);   // one multicast commit arrival per use, in each CTA
    }
    mbar_init(acc_full, 1);                   // multicast commit of the tile's last MMA
    mbar_init(acc_empty, 2 * WG2_EPI_WARPS);  // leader: every epilogue warp of the pair has drained its rows
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  wg2_cluster_sync();   // both CTAs' barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();     TL_WAITED(0);
  pdl_launch();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 dY columns, own half of the X columns =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters) {
        const int tile = unit / L.splits, split = unit - tile * L.splits;
        const Wg2Tile t = wg2_locate(L, tile);
        const Wg2Problem& P = L.p[t.problem];
        const CUtensorMap* amap = &L.maps[P.a_map];
        const CUtensorMap* bmap = &L.maps[P.b_map];
        const int ntile = min(256, P.N - t.tn * 256);
        const int half = ntile >> 1;
        const int m0 = t.tm * 256 + (int)rank * 128;
        const int n0 = t.tn * 256 + (int)rank * half;
        const uint32_t bytes = WG2_A_BYTES + (uint32_t)half * 128u;
        const int kb0 = split * L.kb_split, kb1 = min(kb0 + L.kb_split, L.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * bytes);
          const uint32_t a_dst = a_base + stage * WG2_A_BYTES, b_dst = b_base + stage * WG2_B_BYTES;
          wg2_tma_load(a_dst, amap, m0, kb * 64, P.a_z, full_bar(stage));
          wg2_tma_load(a_dst + 8192u, amap, m0 + 64, kb * 64, P.a_z, full_bar(stage));
          for (int j = 0; j < (half >> 6); ++j)
            wg2_tma_load(b_dst + j * 8192u, bmap, n0 + 64 * j, kb * 64, P.b_z, full_bar(stage));
          if (++stage == WG2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only =====
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = cluster_id; unit < units; unit += n_clusters, ++it) {
        const int tile = unit / L.splits, split = unit - tile * L.splits;
        const Wg2Tile t = wg2_locate(L, tile);
        const int ntile = min(256, L.p[t.problem].N - t.tn * 256);
        // bf16 x bf16 -> fp32, both operands MN-major, N = ntile, M = 256 over the pair
        uint32_t idesc = 0;
        idesc |= 1u << 4;
        idesc |= 1u << 7;
        idesc |= 1u << 10;
        idesc |= 1u << 15;
        idesc |= 1u << 16;
        idesc |= (uint32_t)(ntile >> 3) << 17;
        idesc |= (uint32_t)(256 >> 4) << 24;
        const int kb0 = split * L.kb_split, kb1 = min(kb0 + L.kb_split, L.kb_total);
        mbar_wait(acc_empty, ((uint32_t)it & 1u) ^ 1u);   // the pair's epilogue warps have drained the accumulator
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = a_base + stage * WG2_A_BYTES, b_addr = b_base + stage * WG2_B_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 windows per MMA = 2 KB of each 64 x 64 box
            wg2_mma(tmem_base, smem_desc(a_addr + k * 2048u, 8192u, 1024u), smem_desc(b_addr + k * 2048u, 8192u, 1024u),
                    idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          wg2_commit(empty_bar(stage));
          if (++stage == WG2_STAGES) { stage = 0; phase ^= 1u; }
        }
        wg2_commit(acc_full);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the tile =====
    const int lq = warp & 3;               // TMEM lane quarter this warp may read
    const int colgroup = (warp - 4) >> 2;  // which half of the tile's columns
    unsigned char* stage = stage_smem + (warp - 4) * 32 * WG2_STG_PITCH;
    int it = 0;
    for (int unit = cluster_id; unit < units; unit += n_clusters, ++it) {
      const int tile = unit / L.splits;
      const Wg2Tile t = wg2_locate(L, tile);
      const Wg2Problem& P = L.p[t.problem];
      const int M = P.M, N = P.N;
      float* const C = P.C;
      const long long ldc = P.ldc;
      double* const sqp = P.sq;
      const int ntile = min(256, N - t.tn * 256);
      const int cpw = ntile >> 1;                               // columns per warp
      const int row0 = t.tm * 256 + (int)rank * 128 + lq * 32;  // first dW row of this warp
      const int col0 = t.tn * 256 + colgroup * cpw;
      const bool active = row0 < M;   // depends on the tile only: both halves of a split agree
      // Split contraction: the two clusters of a tile exchange halves.  The cluster of split s FINISHES the columns
      // of column group s and SENDS its partial sums of the other group; a sender stores them into the destination
      // and raises the flag of (tile, CTA rank, lane quarter, column group), the finisher of the other cluster
      // waits for it, adds its own sums and stores the result.  One writer and one reader per flag, no tickets.
      const int split = unit - tile * L.splits;
      unsigned* flag = L.flags + ((tile * 2 + (int)rank) * WG2_EPI_WARPS + (warp - 4));
      const bool sender = active && L.splits == 2 && colgroup != split;
      const bool combine = active && L.splits == 2 && colgroup == split;
      mbar_wait(acc_full, (uint32_t)it & 1u);
      tc_fence_after();
      if (warp == 4 && lane == 0) TL_CTA_MARK(1);   // this CTA's accumulator is complete
      float sq = 0.0f;
      if (active) {
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(colgroup * cpw);
        const int fr = lane >> 3, fch = lane & 7;   // flush: 4 rows x 128 B per instruction, whole lines
#pragma unroll 1
        for (int c = 0; c < cpw; c += 32) {
          uint32_t acc[32];
          tmem_ld32(taddr + (uint32_t)c, acc);
          float4* my = reinterpret_cast<float4*>(stage + lane * WG2_STG_PITCH);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            my[q] = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                                __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
          if (combine && c == 0) {   // the other cluster's partial sums of these columns must be in place
            const long long t0 = clock64();
            while (wg2_ld_acquire(flag) == 0u) {
              if (clock64() - t0 > 4000000000ll) {
                printf("msf_b200 wg2_gemm: split partner never arrived (block %d warp %d)\n", blockIdx.x, warp);
                __trap();
              }
            }
          }
          __syncwarp();
          float* const cbase = C + (long long)(row0 + fr) * ldc + col0 + c + fch * 4;
          float4 other[8];
          if (combine) {   // all eight loads in flight before the first store (they would be ordered behind it)
#pragma unroll
            for (int itr = 0; itr < 8; ++itr)
              other[itr] = (row0 + itr * 4 + fr < M) ? __ldcg(reinterpret_cast<const float4*>(cbase + (long long)itr * 4 * ldc))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int r = itr * 4 + fr;
            if (row0 + r < M) {
              float4 v = *reinterpret_cast<const float4*>(stage + r * WG2_STG_PITCH + fch * 16);
              if (combine) { v.x += other[itr].x; v.y += other[itr].y; v.z += other[itr].z; v.w += other[itr].w; }
              if (!sender) sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
              *reinterpret_cast<float4*>(cbase + (long long)itr * 4 * ldc) = v;
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) wg2_arrive_leader(acc_empty);
      if (active) {
        if (sender) {          // publish the partial sums: one fence, by the signalling lane, after the warp barrier
          __syncwarp();        // (fences are cumulative; a fence per lane is a MEMBAR per warp for nothing)
          if (lane == 0) {
            __threadfence();
            atomicExch(flag, 1u);
          }
        } else {
          if (combine && lane == 0) *flag = 0u;   // ready for the next launch
          if (sqp != nullptr) {
            double sd = (double)sq;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
            if (lane == 0 && sd != 0.0) atomicAdd(sqp, sd);
          }
        }
      }
    }
  }

  tc_fence_before();
  wg2_cluster_sync();   // neither CTA frees TMEM or exits while the pair's MMAs / the peer's signals are in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace

bool wg2_enabled() {
  const char* e = getenv("MSF_WG");
  return !(e && e[0] == 'v' && e[1] == '1');
}

bool wg2_shape_ok(int N, long long ldc, const void* C) {
  return N >= 128 && N % 128 == 0 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
}

Wg2Builder::Wg2Builder(long long rows_, cudaStream_t st, const char* lbl) : rows(rows_), stream(st), status(MSF_OK), label(lbl) {
  memset(&L, 0, sizeof(L));
}

int Wg2Builder::add_map(const void* base, long long r, long long cols, long long ld, long long depth, long long slice) {
  if (L.nmaps >= WG2_MAX_MAPS) {
    set_error("wg2_gemm: too many tensor maps in one launch");
    status = MSF_E_INVALID;
    return -1;
  }
  const int rc = tc_encode_map(&L.maps[L.nmaps], base, r, cols, ld, depth, slice, 64, 64);
  if (rc) {
    status = rc;
    return -1;
  }
  return L.nmaps++;
}

int Wg2Builder::add_problem(short a_map, int a_z, short b_map, int b_z, int M, int N, float* C, long long ldc, double* sq) {
  if (status != MSF_OK) return status;
  if (M <= 0 || N <= 0) return MSF_OK;
  if (a_map < 0 || b_map < 0) return status = MSF_E_INVALID;
  if (!wg2_shape_ok(N, ldc, C)) {
    set_error("wg2_gemm: dW with %d columns (ld %lld) not supported", N, ldc);
    return status = MSF_E_INVALID;
  }
  if (L.count >= WG2_MAX_PROBLEMS) {
    const int rc = flush();
    if (rc) return rc;
  }
  Wg2Problem p;
  memset(&p, 0, sizeof(p));
  p.a_map = a_map; p.a_z = a_z; p.b_map = b_map; p.b_z = b_z;
  p.M = M; p.N = N; p.C = C; p.ldc = ldc; p.sq = sq;
  p.tile_begin = L.total_tiles;
  L.total_tiles += (int)(ceil_div(M, 256) * ceil_div(N, 256));
  L.p[L.count++] = p;
  return MSF_OK;
}

int Wg2Builder::flush() {
  if (status != MSF_OK) return status;
  if (L.total_tiles == 0) return MSF_OK;
  MSF_REQUIRE(rows >= 1 && rows < (1ll << 31) - 64, "wg2_gemm: %lld windows out of range", rows);
  for (int i = L.nmaps; i < WG2_MAX_MAPS; ++i) L.maps[i] = L.maps[0];   // prefetched: keep every slot valid
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int max_clusters = sms / 2;
  L.kb_total = (int)ceil_div(rows, 64);
  // two halves per tile when both fit in one wave of CTA pairs (they wait for one another: they must be co-resident)
  L.splits = (L.kb_total >= 2 && 2 * L.total_tiles <= max_clusters && L.total_tiles <= WG2_MAX_FLAG_TILES &&
              !getenv("MSF_WG_NOSPLIT")) ? 2 : 1;
  L.kb_split = (int)ceil_div(L.kb_total, L.splits);
  static void* fl = nullptr;   // one ticket table per process and device context (one pass in flight at a time)
  if (fl == nullptr) MSF_CHECK_CUDA(cudaGetSymbolAddress(&fl, g_wg2_flags));
  L.flags = reinterpret_cast<unsigned*>(fl);
  const int units = L.total_tiles * L.splits;
  const int clusters = units < max_clusters ? units : max_clusters;
  if (prof_enabled()) {
    double flops = 0.0;
    for (int i = 0; i < L.count; ++i) flops += 2.0 * L.p[i].M * L.p[i].N * (double)rows;
    prof_begin(label, flops, stream);
  }
  MSF_CHECK_CUDA(cudaFuncSetAttribute(wg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG2_SMEM));
  MSF_CHECK_CUDA(launch_pdl(wg2_kernel, dim3(2 * clusters), dim3(WG2_THREADS), WG2_SMEM, stream, L));
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  L.count = 0;
  L.total_tiles = 0;
  return MSF_OK;
}

}  // namespace msf
