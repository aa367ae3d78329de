// proj_kernel: see proj_gemm.cuh.  12 warps: warp 0 TMA producer (all k-blocks of Wp_m at once), warp 1 MMA issuer,
// warp 2 TMEM allocator, warp 3 TMA store, warps 4..11 workers (X phase: 16-byte bf16 chunks of the A block from
// fp32 rows; epilogue: thread = accumulator row).  The weight blocks' shared memory is reused as the staging
// area of the output tile once the MMA has finished.  All mbarrier waits are bounded.
#include "proj_gemm.cuh"

#include <stdlib.h>

#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int PJ_THREADS = 384;
constexpr int PJ_WORKERS = 8;
constexpr uint32_t PJ_A_BYTES = 128 * 64 * 2;
constexpr int PJ_MAX_KB = 4;                 // in_dim <= 256
constexpr size_t PJ_SMEM_LIMIT = 232448;
constexpr int PJ_NBAR = 8;

__device__ __forceinline__ uint32_t pj_swz(int r, int c) {
  return (uint32_t)(c >> 6) * PJ_A_BYTES + (uint32_t)r * 128u + (uint32_t)((((c & 63) >> 3) ^ (r & 7)) << 4);
}
__device__ __forceinline__ uint4 pj_pack8(const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return pk;
}

// 8 consecutive input elements of a row as fp32, from fp32 or bf16 storage (shared or global memory)
__device__ __forceinline__ void pj_load8(const void* base, long long elem, bool is_bf16, bool global, float4& a, float4& b) {
  if (is_bf16) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem);
    const uint4 w = global ? __ldg(p) : *p;
    a = make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16),
                    __uint_as_float(w.y & 0xffff0000u));
    b = make_float4(__uint_as_float(w.z << 16), __uint_as_float(w.z & 0xffff0000u), __uint_as_float(w.w << 16),
                    __uint_as_float(w.w & 0xffff0000u));
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
    a = global ? __ldg(p) : p[0];
    b = global ? __ldg(p + 1) : p[1];
  }
}

__device__ long long g_proj_stamps[16];
#define PJ_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 128) g_proj_stamps[i] = clock64(); } while (0)

__global__ void __launch_bounds__(PJ_THREADS, 1) proj_kernel(const __grid_constant__ ProjLaunch L) {
  TL_KERNEL(0);
  // The last L.zero_ctas CTAs of the grid only clear the gradient slots this train pass accumulates into: walking
  // the range table costs microseconds of dependent constant-bank misses, which these CTAs spend on otherwise
  // idle SMs while the others work.  Chunks of 4096 floats are dealt round-robin.
  const int nwork = (int)gridDim.x - L.zero_ctas;
  if ((int)blockIdx.x >= nwork) {
    pdl_wait();     // the previous step's optimizer still reads these slots
    pdl_launch();
    const int me = (int)blockIdx.x - nwork;
    int chunk = 0;
    for (int i = 0; i < L.nzero; ++i) {
      const ProjZeroRange R = L.zero[i];
      const int per = (int)((R.count + 4095) >> 12);
      for (int b = 0; b < R.batch; ++b)
        for (int j = 0; j < per; ++j, ++chunk) {
          if (chunk % L.zero_ctas != me) continue;
          float* p = L.zero_base + R.begin + (long long)b * R.stride + ((long long)j << 12);
          const int n = (int)min((long long)4096, R.count - ((long long)j << 12));
          for (int e = threadIdx.x; e < n; e += PJ_THREADS) p[e] = 0.0f;
        }
    }
    return;
  }
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t smem_base = (off0 + 1023u) & ~1023u;
  const int H = L.H, M = L.M, KBO = L.H >> 6;
  const uint32_t WB = (uint32_t)H * 128u;                   // one k-block of Wp: [H rows][64]
  const uint32_t w_base = smem_base;                        // weight k-blocks, later the output staging tile
  const uint32_t a_base = w_base + L.w_area;                // A block: max_m D_m / 64 k-blocks
  const uint32_t x_base = a_base + L.a_area;                // fp32 input tile staged by one bulk copy (x_area may be 0)
  const uint32_t bar_base = x_base + L.x_area;
  const uint32_t w_full = bar_base + 0;      // weights landed (TMA -> MMA)
  const uint32_t a_ready = bar_base + 8;     // A block written (workers -> MMA, store)
  const uint32_t acc_full = bar_base + 16;   // GEMM finished (MMA -> workers, producer)
  const uint32_t out_ready = bar_base + 24;  // output tile staged (workers -> store)
  const uint32_t tile_done = bar_base + 32;  // stores have read A block and staging (store -> workers, producer)
  const uint32_t x_full = bar_base + 40;     // input tile landed (bulk copy -> workers)
  const uint32_t tmem_slot = bar_base + 48;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * PJ_NBAR - off0));
  float* mask_s = bias_s + H;                // mask column of the tile's 128 windows
  float* stat_s = mask_s + 128;              // LayerNorm: mean[128], rstd[128] of the tile's rows
  float* lnw_s = stat_s + 256;               // LayerNorm weight / bias of the item's modality (PJ_MAX_KB * 64 each)
  float* lnb_s = lnw_s + 64 * PJ_MAX_KB;
  const float* x_smem = reinterpret_cast<const float*>(smem_raw + (x_base - off0));
  const bool staged = L.x_area > 0;
  unsigned char* a_smem = smem_raw + (a_base - off0);
  unsigned char* o_smem = smem_raw + (w_base - off0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  PJ_STAMP(0);
  const uint32_t tmem_cols = H <= 32 ? 32u : H <= 64 ? 64u : H <= 128 ? 128u : 256u;

  if (warp == 0 && lane == 0) {
    for (int m = 0; m < M; ++m) {
      tma_prefetch_desc(&L.map_w[m]);
      tma_prefetch_desc(&L.map_xt[m]);
    }
    tma_prefetch_desc(&L.map_p);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_ready, PJ_WORKERS);
    mbar_init(acc_full, 1);
    mbar_init(out_ready, PJ_WORKERS);
    mbar_init(tile_done, 1);
    mbar_init(x_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  TL_WAITED(0);
  pdl_launch();
  PJ_STAMP(1);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < L.items; item += nwork) {
        const int m = L.active[item % L.n_active], KB = L.D[m] >> 6, m0 = (item / L.n_active) * 128;
        const uint32_t it_cur = it++;
        if (it_cur > 0) mbar_wait(tile_done, (it_cur - 1) & 1u);   // the staging tile in the weight area has been stored
        if (staged) {   // the tile's fp32 rows are contiguous in global memory: one bulk copy
          const int nrows = (L.rows - m0) < 128 ? (L.rows - m0) : 128;
          const uint32_t esz = L.x_bf16 ? 2u : 4u;
          const uint32_t bytes = (uint32_t)nrows * (uint32_t)L.D[m] * esz;
          mbar_expect_tx(x_full, bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(x_base), "l"(reinterpret_cast<const unsigned char*>(L.x[m]) + (long long)m0 * L.D[m] * esz),
                         "r"(bytes), "r"(x_full)
                       : "memory");
        }
        mbar_expect_tx(w_full, (uint32_t)KB * WB);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(w_base + kb * WB, &L.map_w[m], kb * 64, 0, 0, w_full);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc = instr_desc(H, false, false);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < L.items; item += nwork) {
        const int KB = L.D[L.active[item % L.n_active]] >> 6;
        const uint32_t it_cur = it++;
        mbar_wait(w_full, it_cur & 1u);
        mbar_wait(a_ready, it_cur & 1u);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base, smem_desc(a_base + kb * PJ_A_BYTES + k * 32, 16, 1024),
                        smem_desc(w_base + kb * WB + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(acc_full);
      }
    }
  } else if (warp == 3) {
    // =========================== TMA store ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < L.items; item += nwork) {
        const int m = L.active[item % L.n_active], m0 = (item / L.n_active) * 128, KB = L.D[m] >> 6;
        const uint32_t it_cur = it++;
        mbar_wait(a_ready, it_cur & 1u);
        for (int kb = 0; kb < KB; ++kb) tma_store_3d(&L.map_xt[m], a_base + kb * PJ_A_BYTES, kb * 64, m0, 0);
        tma_store_commit();
        mbar_wait(out_ready, it_cur & 1u);
        for (int kb = 0; kb < KBO; ++kb) tma_store_3d(&L.map_p, w_base + kb * PJ_A_BYTES, kb * 64, m0, m);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(tile_done);
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =========================== workers ================================
    const DropCfg drop = resolve_drop(L.drop);
    const int et = threadIdx.x - 128;
    const int lq = warp & 3, cg = (warp - 4) >> 2;
    const int trow = lq * 32 + lane;
    const int half = H >> 1, c_begin = cg * half;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;

    uint32_t it_next = 0;
    for (int item = blockIdx.x; item < L.items; item += nwork) {
      const int m = L.active[item % L.n_active], m0 = (item / L.n_active) * 128, D = L.D[m], c8n = D >> 3;
      const uint32_t it = it_next++;
      if (it > 0) mbar_wait(tile_done, (it - 1) & 1u);   // previous tile's A block / staging / bias row are free
      for (int e = et; e < H; e += 32 * PJ_WORKERS) bias_s[e] = L.bias[m] ? __ldg(L.bias[m] + e) : 0.0f;
      if (et < 128) mask_s[et] = (L.mask && (long long)m0 + et < L.rows) ? __ldg(L.mask + ((long long)m0 + et) * M + m) : 1.0f;
      const bool ln = L.ln_w[m] != nullptr || L.ln_b[m] != nullptr;
      if (ln)
        for (int e = et; e < D; e += 32 * PJ_WORKERS) {
          lnw_s[e] = L.ln_w[m] ? __ldg(L.ln_w[m] + e) : 1.0f;
          lnb_s[e] = L.ln_b[m] ? __ldg(L.ln_b[m] + e) : 0.0f;
        }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PJ_WORKERS) : "memory");   // bias row / mask column visible
      PJ_STAMP(2);
      if (staged) mbar_wait(x_full, it & 1u);
      PJ_STAMP(3);
      const float* xm = L.x[m];
      if (ln) {
        // row statistics, one warp per 16 rows, the row in registers (float4 per lane and 128 columns): mean, then
        // the centred sum of squares (biased variance, like nn.LayerNorm)
        const int c8s = D >> 3;   // 8-column groups per row: lane l takes group l (D <= 256)
        const float inv_d = 1.0f / (float)D;
        const bool xb = L.x_bf16 != 0;
#pragma unroll 1
        for (int rr = 0; rr < 16; ++rr) {
          const int r = (warp - 4) * 16 + rr;
          const long long row = (long long)m0 + r;
          float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
          const bool have = lane < c8s && row < L.rows;
          if (have) pj_load8(staged ? (const void*)x_smem : (const void*)xm, (staged ? (long long)r : row) * D + lane * 8, xb, !staged, va, vb);
          float s = va.x + va.y + va.z + va.w + vb.x + vb.y + vb.z + vb.w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * inv_d;
          float q = 0.0f;
          if (lane < c8s) {
            const float a0 = va.x - mean, a1 = va.y - mean, a2 = va.z - mean, a3 = va.w - mean;
            const float b0 = vb.x - mean, b1 = vb.y - mean, b2 = vb.z - mean, b3 = vb.w - mean;
            q = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
          if (lane == 0) {
            stat_s[r] = mean;
            stat_s[128 + r] = rsqrtf(q * inv_d + L.ln_eps);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * PJ_WORKERS) : "memory");   // every row's statistics visible
      }

      // ---- X phase: xt = bf16(drop0(LN(x) * mask)), 8 columns per thread and step ----
#pragma unroll 1
      for (int id = et; id < 128 * c8n; id += 32 * PJ_WORKERS) {
        const int r = id / c8n, c8 = id - r * c8n;
        const long long row = (long long)m0 + r;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row < L.rows) {
          float4 a, b;
          pj_load8(staged ? (const void*)x_smem : (const void*)xm, (staged ? (long long)r : row) * D + c8 * 8, L.x_bf16 != 0,
                   !staged, a, b);
          if (ln) {
            const float mu = stat_s[r], rs = stat_s[128 + r];
            const float4 g0 = *reinterpret_cast<const float4*>(lnw_s + c8 * 8), g1 = *reinterpret_cast<const float4*>(lnw_s + c8 * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(lnb_s + c8 * 8), b1 = *reinterpret_cast<const float4*>(lnb_s + c8 * 8 + 4);
            a.x = (a.x - mu) * rs * g0.x + b0.x; a.y = (a.y - mu) * rs * g0.y + b0.y;
            a.z = (a.z - mu) * rs * g0.z + b0.z; a.w = (a.w - mu) * rs * g0.w + b0.w;
            b.x = (b.x - mu) * rs * g1.x + b1.x; b.y = (b.y - mu) * rs * g1.y + b1.y;
            b.z = (b.z - mu) * rs * g1.z + b1.z; b.w = (b.w - mu) * rs * g1.w + b1.w;
          }
          const float mk = mask_s[r];
          float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
          if (drop.active) drop8(drop, SITE_INPUT, m, row, c8, dm);
          v[0] = a.x * mk * dm[0]; v[1] = a.y * mk * dm[1]; v[2] = a.z * mk * dm[2]; v[3] = a.w * mk * dm[3];
          v[4] = b.x * mk * dm[4]; v[5] = b.y * mk * dm[5]; v[6] = b.z * mk * dm[6]; v[7] = b.w * mk * dm[7];
        }
        *reinterpret_cast<uint4*>(a_smem + pj_swz(r, c8 * 8)) = pj_pack8(v);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      PJ_STAMP(4);

      // ---- epilogue: P = drop1(relu(acc + bias)); the Philox draws are made while the GEMM runs ----
      const long long row = (long long)m0 + trow;
      unsigned long long keep_lo = ~0ull, keep_hi = ~0ull;
      if (drop.active) {
        keep_lo = keep_hi = 0ull;
#pragma unroll 1
        for (int g8 = 0; g8 < (half >> 3); ++g8) {
          float d8[8];
          drop8(drop, SITE_PROJ, m, row, (c_begin >> 3) + g8, d8);
          unsigned long long m8 = 0ull;
#pragma unroll
          for (int j = 0; j < 8; ++j) m8 |= (d8[j] != 0.0f ? 1ull : 0ull) << j;
          if (g8 < 8) keep_lo |= m8 << (8 * g8);
          else keep_hi |= m8 << (8 * (g8 - 8));
        }
      }
      PJ_STAMP(5);
      mbar_wait(acc_full, it & 1u);
      PJ_STAMP(6);
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < (half >> 4); ++q) {
        const int c = c_begin + q * 16;
        uint32_t acc[16];
        tmem_ld16_issue(tmem_base + lane_base + (uint32_t)c, acc);
        tmem_wait16(acc);
        const uint32_t k16 = (uint32_t)((q < 4 ? keep_lo >> (16 * q) : keep_hi >> (16 * (q - 4))) & 0xffffull);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = ((k16 >> j) & 1u) ? fmaxf(__uint_as_float(acc[j]) + bias_s[c + j], 0.0f) * drop.scale : 0.0f;
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          float w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = v[h8 * 8 + e];
          *reinterpret_cast<uint4*>(o_smem + pj_swz(trow, c + 8 * h8)) = pj_pack8(w8);
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_ready);
      PJ_STAMP(7);
    }
  }

  tc_fence_before();
  __syncthreads();
  PJ_STAMP(8);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace

int proj_debug_stamps(long long* out16) {
  MSF_CHECK_CUDA(cudaDeviceSynchronize());
  MSF_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_proj_stamps, sizeof(long long) * 16));
  return MSF_OK;
}

bool proj_eligible(int H, int M, const int* D) {
  if (H % 64 != 0 || H < 64 || H > 256 || M < 1 || M > MSF_MAX_MODALITIES) return false;
  for (int m = 0; m < M; ++m)
    if (D[m] % 64 != 0 || D[m] < 64 || D[m] > 64 * PJ_MAX_KB) return false;
  return true;
}

int proj_launch(ProjLaunch& L, cudaStream_t stream, const char* label) {
  MSF_REQUIRE(proj_eligible(L.H, L.M, L.D), "proj_gemm: shape not supported");
  MSF_REQUIRE(L.rows >= 1, "proj_gemm: empty batch");
  L.row_tiles = (int)ceil_div(L.rows, 128);
  if (L.n_active <= 0) {
    L.n_active = L.M;
    for (int m = 0; m < L.M; ++m) L.active[m] = (short)m;
  }
  L.items = L.row_tiles * L.n_active;
  int max_d = 64;
  for (int m = 0; m < L.M; ++m) max_d = L.D[m] > max_d ? L.D[m] : max_d;
  const int max_kb = max_d / 64;
  const size_t out_tile = (size_t)(L.H / 64) * PJ_A_BYTES;
  L.w_area = (int)((size_t)max_kb * L.H * 128 > out_tile ? (size_t)max_kb * L.H * 128 : out_tile);
  L.a_area = max_kb * (int)PJ_A_BYTES;
  const size_t fixed = 1024 + 8 * PJ_NBAR + (size_t)(L.H + 128 + 256 + 2 * 64 * PJ_MAX_KB) * 4;
  L.x_area = 128 * max_d * 4;     // fp32 input tile staged by a bulk copy when it fits
  if (fixed + L.w_area + L.a_area + L.x_area > PJ_SMEM_LIMIT || getenv("MSF_PROJ_NO_STAGE")) L.x_area = 0;
  const size_t smem = fixed + L.w_area + L.a_area + L.x_area;
  MSF_REQUIRE(smem <= PJ_SMEM_LIMIT, "proj_gemm: not enough shared memory for hidden %d", L.H);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int nwork = L.items < sms ? L.items : sms;
  L.zero_ctas = L.nzero > 0 ? 16 : 0;
  const int grid = nwork + L.zero_ctas;
  if (prof_enabled()) {
    double fl = 0.0;
    for (int m = 0; m < L.M; ++m) fl += 2.0 * (double)L.rows * L.H * L.D[m];
    prof_begin(label, fl, stream);
  }
  MSF_CHECK_CUDA(cudaFuncSetAttribute(proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSF_CHECK_CUDA(launch_pdl(proj_kernel, dim3(grid), dim3(PJ_THREADS), smem, stream, L));
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  return MSF_OK;
}

}  // namespace msf
