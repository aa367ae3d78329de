// head_kernel: see head_gemm.cuh.  One persistent CTA per 128-window tile, 12 warps:
//
//   warp 0 (one lane)  TMA producer : W1, W2, W2^T, W1^T k-blocks through a shared-memory ring
//   warp 1 (one lane)  MMA issuer   : G1 -> TMEM [0,H), G2 -> [256,288), G3 -> [0,H), G4 -> [256,256+H)
//   warp 2             TMEM allocator
//   warp 3 (one lane)  TMA store    : fused / Hr / dH1 tiles from the A block to global memory
//   warps 4..11        workers      : P0 / P5 with a warp per window (lanes own 8-column chunks, scores by
//                                     warp shuffles), E1..E4 with a thread per accumulator row (tcgen05.ld)
//
// The "A block" (H/64 k-blocks of 128 x 64 bf16, 128B-swizzled K-major) is written by the workers and is at
// once the A operand of the next tcgen05.mma and the source box of the TMA store.  Its contents per tile:
// fused -> Hr -> dlog -> dH1 -> dfused.  All mbarrier waits are bounded (trap instead of hanging the GPU).
#include "head_gemm.cuh"

#include <math.h>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int HD_THREADS = 384;
constexpr int HD_WORKERS = 8;          // worker warps
constexpr int HD_MAX_STAGES = 4;
constexpr uint32_t HD_A_BYTES = 128 * 64 * 2;  // one k-block of the A block
constexpr size_t HD_SMEM_LIMIT = 232448;
constexpr int HD_RS = MSF_MAX_MODALITIES;      // stride of the per-row scalar arrays
constexpr int HD_NBAR = 2 * HD_MAX_STAGES + 4;

__host__ __device__ constexpr uint32_t hd_ring_bytes(int H) { return (uint32_t)H * 64 * 2; }
// the cross-entropy's logit staging tile [128][33] fp32 lives in k-blocks 1.. of the A block when the A block has at
// least three k-blocks, else in its own region behind the per-row scalars
__host__ __device__ constexpr int hd_zs_floats(int H) { return H >= 192 ? 0 : 128 * 33; }

// CTAs that have finished; the last one reduces row_loss in a fixed order and resets the ticket
__device__ unsigned int g_head_ticket = 0;
// phase time stamps (clock64) of CTA 0, worker warp 0: msf_debug_head_stamps()
__device__ long long g_head_stamps[16];
#define HD_STAMP(i)                                                          \
  do {                                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 128) g_head_stamps[i] = clock64(); \
  } while (0)

__device__ __forceinline__ float hd_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&out)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    out[2 * e] = __uint_as_float(w[e] << 16);
    out[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 pk;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return pk;
}
// byte offset of the 16-byte chunk holding columns [c, c+8) of tile row r inside the A block
__device__ __forceinline__ uint32_t swz_off(int r, int c) {
  return (uint32_t)(c >> 6) * HD_A_BYTES + (uint32_t)r * 128u + (uint32_t)((((c & 63) >> 3) ^ (r & 7)) << 4);
}
__device__ __forceinline__ void st_swz16(unsigned char* blk, int r, int c, const float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = v[q * 8 + e];
    *reinterpret_cast<uint4*>(blk + swz_off(r, c + 8 * q)) = pack8(h);
  }
}
__device__ __forceinline__ void ld8_smem_f32(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Sum NV per-lane values over the warp with about NV shuffles instead of 5 * NV: every butterfly step
// halves the number of values a lane keeps.  NV = 16: lanes 2c and 2c+1 end up with the total of value c;
// NV = 32: lane c ends up with the total of value c.
template <int HALF, int BIT, int NV>
__device__ __forceinline__ void tr_step(float (&p)[NV], int lane) {
  const bool hi = (lane & BIT) != 0;
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const float send = hi ? p[k] : p[k + HALF];
    const float keep = hi ? p[k + HALF] : p[k];
    p[k] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
  }
}
// 4 values: afterwards every lane holds the total of value (lane >> 3) & 3
__device__ __forceinline__ float transpose_reduce4(float (&p)[4], int lane) {
  tr_step<2, 16>(p, lane);
  tr_step<1, 8>(p, lane);
  float r = p[0];
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
// 8 values: afterwards every lane holds the total of value (lane >> 2) & 7
__device__ __forceinline__ float transpose_reduce8(float (&p)[8], int lane) {
  tr_step<4, 16>(p, lane);
  tr_step<2, 8>(p, lane);
  tr_step<1, 4>(p, lane);
  float r = p[0];
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

struct HeadSmem {
  unsigned char* ublk;
  float *b1s, *b2s, *db2s, *gws, *gbs, *rw, *rsoft, *rmk;
};

// Sum / max over the (up to 4) modalities of a window: lane l holds modality (l >> 3) & 3.
__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float group_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
  return v;
}

// Aggregated-token tiles reach P0 / P5 through the TMA ring: one sub-block = 32 windows x M modalities x H
// columns, staged as ceil(M/2) ring stages of [2 modalities][H/64 k-blocks][32 rows x 128 B, 128B-swizzled].
struct AggStage {
  const unsigned char* base[2];   // stage holding modalities {0,1} and {2,3}
  int KB;
};
// this lane's 16-byte chunk (columns lane*8 .. +8) of modality q, sub-block row rl
__device__ __forceinline__ uint4 agg_chunk(const AggStage& st, int q, int rl, int lane) {
  const unsigned char* p = st.base[q >> 1] + (((q & 1) * st.KB + (lane >> 3)) << 12) + rl * 128 +
                           (((lane & 7) ^ (rl & 7)) << 4);
  return *reinterpret_cast<const uint4*>(p);
}

// The row-wise phases are written for MINIMUM UNIQUE CODE: at small batches every CTA runs each phase once,
// so its time is set by cold instruction fetch (ncu: stall_no_inst), not by arithmetic.  One window per
// iteration of a rolled loop, everything else lane-parallel.
//
// ---- P0: gating + weighted sum for the 4 windows this warp owns in the staged 32-window sub-block.  Lanes own
// 8-column chunks for the dot products / weighted sum; after the transposed reduction lane l holds the score of
// modality (l >> 3) & 3, so the softmax arithmetic of src/fusion.py:464-478 is one pass of lane-parallel code. ----
__device__ __forceinline__ void head_p0(const HeadLaunch& L, const HeadSmem& S, const AggStage& st, int m0, int sb,
                                        int wq, int lane) {
  const int H = L.H, M = L.M, c = lane * 8;
  const bool c_ok = c < H;
  const int qi = (lane >> 3) & 3;
  const bool lane_ok = qi < M;
  const float gb = S.gbs[qi];
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int rl = wq * 4 + i, r = sb * 32 + rl;
    const long long row = (long long)m0 + r;
    const bool row_ok = row < L.rows;
    uint4 raw[4];
    float part[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      raw[q] = make_uint4(0u, 0u, 0u, 0u);
      part[q] = 0.0f;
      if (q < M && c_ok) {
        raw[q] = agg_chunk(st, q, rl, lane);
        float v[8], g[8];
        unpack8(raw[q], v);
        ld8_smem_f32(S.gws + q * H + c, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) part[q] = fmaf(v[j], g[j], part[q]);
      }
    }
    const float sc = transpose_reduce4(part, lane) + gb;              // fusion.py:459
    float mk = 0.0f;
    if (lane_ok) mk = (L.mask != nullptr && row_ok) ? __ldg(L.mask + row * M + qi) : 1.0f;
    // masked softmax with the reference's fallbacks (fusion.py:464-478)
    const float mx = group_max((lane_ok && mk > 0.0f) ? sc : -INFINITY);
    const float e = (lane_ok && mk > 0.0f && mx > -INFINITY) ? expf(sc - mx) : 0.0f;
    const float den = group_sum(e);
    const float soft = den > 0.0f ? e / den : 0.0f;
    float w = soft * mk;
    const float sum_w = group_sum(w), mask_sum = group_sum(mk);
    if (sum_w > 0.0f) w = w / (sum_w + 1e-8f);
    else w = lane_ok ? (mask_sum > 0.0f ? mk / (mask_sum + 1e-8f) : 1.0f / (float)M) : 0.0f;
    if (lane_ok && (lane & 7) == 0) {
      S.rw[r * HD_RS + qi] = w;
      S.rsoft[r * HD_RS + qi] = soft;
      S.rmk[r * HD_RS + qi] = mk;
      if (row_ok) {
        L.soft[row * M + qi] = soft;
        L.w[row * M + qi] = w;
        if (L.w_out != nullptr) L.w_out[row * M + qi] = w;
      }
    }
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float wb = __shfl_sync(0xffffffffu, w, 8 * q);
      float v[8];
      unpack8(raw[q], v);
      if (wb != 0.0f) {   // weight 0 <=> masked token: exactly 0 in the reference, possibly never written here
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(v[j], wb, f[j]);  // fusion.py:416-418
      }
    }
    // padding rows of the last tile stay zero
    if (c_ok) *reinterpret_cast<uint4*>(S.ublk + swz_off(r, c)) = row_ok ? pack8(f) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---- P5: backward of P0 with the same lane layout: d scores lane-parallel, then
// dS_q = (w_q dfused + ds_q gw_q) * mask_q / cnt_q. ----
__device__ __forceinline__ void head_p5(const HeadLaunch& L, const HeadSmem& S, const AggStage& st, int m0, int sb,
                                        int wq, int lane) {
  const int H = L.H, M = L.M, c = lane * 8;
  const bool c_ok = c < H;
  const int qi = (lane >> 3) & 3;
  const bool lane_ok = qi < M;
  const float icnt = L.inv_cnt[qi];
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int rl = wq * 4 + i, r = sb * 32 + rl;
    const long long row = (long long)m0 + r;
    const bool row_ok = row < L.rows;
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c_ok && row_ok) unpack8(*reinterpret_cast<const uint4*>(S.ublk + swz_off(r, c)), g);
    float part[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      part[q] = 0.0f;
      if (q < M && c_ok) {
        float v[8];
        unpack8(agg_chunk(st, q, rl, lane), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) part[q] = fmaf(v[j], g[j], part[q]);
      }
    }
    const float dw = transpose_reduce4(part, lane);   // d loss / d w_q for this lane's modality
    const float p = lane_ok ? S.rsoft[r * HD_RS + qi] : 0.0f;
    const float w = lane_ok ? S.rw[r * HD_RS + qi] : 0.0f;
    const float mk = lane_ok ? S.rmk[r * HD_RS + qi] : 0.0f;
    // w = n / (S + 1e-8), n = p * mask (only when S > 0; the fallbacks are constants)
    const float Ssum = group_sum(p * mk);
    const float inv = 1.0f / (Ssum + 1e-8f);
    const float dot = group_sum(dw * p * mk);
    const float dp = (dw * inv - dot * inv * inv) * mk;
    const float pdot = group_sum(dp * p);
    const float ds = (Ssum > 0.0f && row_ok && mk > 0.0f) ? p * (dp - pdot) : 0.0f;
    if (L.ds != nullptr && lane_ok && row_ok && (lane & 7) == 0) L.ds[row * M + qi] = ds;
    const float scl = mk * icnt;
#pragma unroll 1
    for (int q = 0; q < M; ++q) {
      const float wb = __shfl_sync(0xffffffffu, w, 8 * q), dsb = __shfl_sync(0xffffffffu, ds, 8 * q);
      const float sb2 = __shfl_sync(0xffffffffu, scl, 8 * q);
      if (c_ok && row_ok) {
        float gwq[8], o[8];
        ld8_smem_f32(S.gws + q * H + c, gwq);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(wb, g[j], dsb * gwq[j]) * sb2;
        *reinterpret_cast<uint4*>(L.dS + ((long long)q * L.rows + row) * H + c) = pack8(o);
      }
    }
  }
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * HD_WORKERS) : "memory"); }

__global__ void __launch_bounds__(HD_THREADS, 1) head_kernel(const __grid_constant__ HeadLaunch L) {
  TL_KERNEL(0);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t smem_base = (off0 + 1023u) & ~1023u;
  const int H = L.H, KB = L.H >> 6, STAGES = L.stages, M = L.M, C = L.C;
  const uint32_t RING = hd_ring_bytes(H);
  const uint32_t ring_base = smem_base;
  const uint32_t u_base = ring_base + STAGES * RING;
  const uint32_t bar_base = u_base + KB * HD_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (HD_MAX_STAGES + s); };
  const uint32_t a_ready = bar_base + 8u * (2 * HD_MAX_STAGES + 0);   // A block written (workers -> MMA, store)
  const uint32_t a_free = bar_base + 8u * (2 * HD_MAX_STAGES + 1);    // A block consumed (MMA + store -> workers)
  const uint32_t acc_full = bar_base + 8u * (2 * HD_MAX_STAGES + 2);  // a GEMM finished (MMA -> workers)
  const uint32_t tmem_slot = bar_base + 8u * (2 * HD_MAX_STAGES + 3);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* fbase = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * HD_NBAR - off0));
  HeadSmem S;
  S.ublk = smem_raw + (u_base - off0);
  S.b1s = fbase;
  S.b2s = S.b1s + H;
  S.db2s = S.b2s + 32;   // this CTA's share of the classifier.3 bias gradient
  S.gws = S.db2s + 32;
  S.gbs = S.gws + M * H;
  S.rw = S.gbs + HD_RS;
  S.rsoft = S.rw + 128 * HD_RS;
  S.rmk = S.rsoft + 128 * HD_RS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  HD_STAMP(11);
  const int nsteps = L.train ? 4 : 2;   // async-consumed A-block contents per tile
  const int TR = L.tile_rows, n_sb = TR >> 5;         // windows per tile (128, or 32 to spread a small batch over the SMs)
  const int n_aq = (M + 1) >> 1, n_agg = n_sb * n_aq; // ring stages per 32-window sub-block / per P0 or P5 pass

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.map_agg);
    tma_prefetch_desc(&L.map_w1);
    tma_prefetch_desc(&L.map_w2);
    if (L.train) {
      tma_prefetch_desc(&L.map_w2t);
      tma_prefetch_desc(&L.map_w1t);
      tma_prefetch_desc(&L.map_dh1);
    }
    tma_prefetch_desc(&L.map_fused);
    tma_prefetch_desc(&L.map_hr);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(a_ready, HD_WORKERS);
    mbar_init(a_free, 2);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 4) {  // constants of the whole launch -> shared memory
    const int et = threadIdx.x - 128;
    for (int e = et; e < H; e += 32 * HD_WORKERS) S.b1s[e] = __ldg(L.b1 + e);
    if (et < 32) S.b2s[et] = et < C ? __ldg(L.b2 + et) : 0.0f;
    if (et < 32) S.db2s[et] = 0.0f;
    for (int e = et; e < M * H; e += 32 * HD_WORKERS) S.gws[e] = __ldg(L.gate_w[e / H] + e % H);
    if (et < HD_RS) S.gbs[et] = et < M ? __ldg(L.gate_b[et]) : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();     TL_WAITED(0);  // everything above (incl. the parameter rows, written >= 2 kernels ago) overlapped the previous kernel
  pdl_launch();
  if (L.grad_sq != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *L.grad_sq = 0.0;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* map, int kcol, uint32_t bytes) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), bytes);
        tma_load_3d(ring_base + stage * RING, map, kcol, 0, 0, full_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      };
      auto load_agg = [&](int m0) {   // sub-blocks of 32 windows, ceil(M/2) stages each
        for (int sb = 0; sb < n_sb; ++sb)
          for (int j = 0; j < n_aq; ++j) {
            const int nq = (M - 2 * j) < 2 ? (M - 2 * j) : 2;
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), (uint32_t)(nq * KB) << 12);
            for (int qq = 0; qq < nq; ++qq)
              for (int kb = 0; kb < KB; ++kb)
                tma_load_3d(ring_base + stage * RING + (uint32_t)((qq * KB + kb) << 12), &L.map_agg, kb * 64,
                            m0 + sb * 32, 2 * j + qq, full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
      };
      for (int tile = blockIdx.x; tile < L.row_tiles; tile += gridDim.x) {
        load_agg(tile * TR);                                                     // P0
        for (int kb = 0; kb < KB; ++kb) load(&L.map_w1, kb * 64, RING);          // G1: W1 [H][64]
        for (int kb = 0; kb < KB; ++kb) load(&L.map_w2, kb * 64, 32u * 128u);    // G2: W2 [32][64]
        if (L.train) {
          load(&L.map_w2t, 0, RING);                                             // G3: W2^T [H][64] (cols >= Cp zero)
          for (int kb = 0; kb < KB; ++kb) load(&L.map_w1t, kb * 64, RING);       // G4: W1^T [H][64]
          load_agg(tile * TR);                                                   // P5
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc_h = instr_desc(H, false, false), idesc_c = instr_desc(32, false, false);
      int stage = 0;
      uint32_t phase = 0, n_ready = 0;
      auto gemm = [&](uint32_t d_tmem, uint32_t idesc, int kblocks, int ksteps) {
        mbar_wait(a_ready, n_ready & 1u);
        ++n_ready;
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = u_base + kb * HD_A_BYTES, b_addr = ring_base + stage * RING;
          for (int k = 0; k < ksteps; ++k)
            tc_mma_bf16(d_tmem, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(acc_full);
        tc_commit(a_free);   // 1 of 2: the GEMM no longer reads the A block
      };
      auto skip = [&](int n) {   // ring stages consumed by the workers (aggregated-token sub-blocks)
        for (int i = 0; i < n; ++i)
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int tile = blockIdx.x; tile < L.row_tiles; tile += gridDim.x) {
        skip(n_agg);                                     // P0
        gemm(tmem_base, idesc_h, KB, 4);                 // G1: fused . W1^T
        gemm(tmem_base + 256u, idesc_c, KB, 4);          // G2: Hr . W2^T
        if (L.train) {
          gemm(tmem_base, idesc_h, 1, 2);                // G3: dlog . W2   (K = 32)
          gemm(tmem_base + 256u, idesc_h, KB, 4);        // G4: dH1 . W1
          skip(n_agg);                                   // P5
        }
      }
    }
  } else if (warp == 3) {
    // =========================== TMA store ==============================
    if (lane == 0) {
      uint32_t n_ready = 0;
      for (int tile = blockIdx.x; tile < L.row_tiles; tile += gridDim.x) {
        const int m0 = tile * TR;
        for (int step = 0; step < nsteps; ++step) {
          mbar_wait(a_ready, n_ready & 1u);
          ++n_ready;
          const CUtensorMap* map = nullptr;
          if (step == 0 && (L.train || L.store_acts)) map = &L.map_fused;
          if (step == 1 && (L.train || L.store_acts)) map = &L.map_hr;
          if (step == 3) map = &L.map_dh1;
          if (map != nullptr) {
            for (int kb = 0; kb < KB; ++kb) tma_store_3d(map, u_base + kb * HD_A_BYTES, kb * 64, m0, 0);
            tma_store_commit();
            tma_store_wait_read();
          }
          mbar_arrive(a_free);  // 2 of 2
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =========================== workers ================================
    const DropCfg drop = resolve_drop(L.drop);
    const int wq = warp - 4;
    const int lq = warp & 3, cg = wq >> 2;
    const int trow = lq * 32 + lane;                 // accumulator row = TMEM lane
    const int half = H >> 1, c_begin = cg * half;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const bool warp_live = lq * 32 < L.tile_rows;    // 32-window tiles: TMEM lanes 32..127 are MMA padding
    uint32_t n_full = 0, n_free = 0;
    int wstage = 0;          // the workers' view of the ring position
    uint32_t wphase = 0;
    auto ring_skip = [&](int n) {   // stages consumed by the MMA warp (weights)
      for (int i = 0; i < n; ++i)
        if (++wstage == STAGES) { wstage = 0; wphase ^= 1u; }
    };
    // One pass over the tile's aggregated tokens: staged sub-blocks of 32 windows, each warp 4 windows.
    auto agg_pass = [&](int m0, bool backward) {
#pragma unroll 1
      for (int sb = 0; sb < n_sb; ++sb) {
        AggStage st;
        st.KB = KB;
        st.base[0] = st.base[1] = nullptr;
        int rel0 = 0, rel1 = 0;
        for (int j = 0; j < n_aq; ++j) {
          mbar_wait(full_bar(wstage), wphase);
          st.base[j] = smem_raw + (ring_base + wstage * RING - off0);
          if (j == 0) rel0 = wstage;
          else rel1 = wstage;
          if (++wstage == STAGES) { wstage = 0; wphase ^= 1u; }
        }
        if (backward) head_p5(L, S, st, m0, sb, wq, lane);
        else head_p0(L, S, st, m0, sb, wq, lane);
        workers_sync();   // every warp is done with the sub-block: its stages may be refilled
        if (threadIdx.x == 128) {
          mbar_arrive(empty_bar(rel0));
          if (n_aq > 1) mbar_arrive(empty_bar(rel1));
        }
      }
    };
    auto publish = [&]() {  // A block written: visible to the async proxy, then signal
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    };
    auto acquire = [&]() {  // accumulator complete and the A block reusable
      mbar_wait(acc_full, n_full & 1u);
      ++n_full;
      tc_fence_after();
      mbar_wait(a_free, n_free & 1u);
      ++n_free;
    };
    for (int tile = blockIdx.x; tile < L.row_tiles; tile += gridDim.x) {
      const int m0 = tile * TR;
      const long long row = (long long)m0 + trow;
      const bool row_ok = trow < TR && row < L.rows;   // rows beyond the tile's windows are MMA padding

      // ---- P0 ----
      HD_STAMP(0);
      agg_pass(m0, false);
      publish();
      HD_STAMP(1);

      // ---- E1: Hr = drop(relu(acc + b1)); the relu/dropout mask of this thread's columns stays in registers.
      // The Philox draws do not depend on the accumulator: they are made while G1 runs. ----
      unsigned long long bits_lo = ~0ull, bits_hi = ~0ull;   // first the dropout keep mask, then keep & relu'
      if (drop.active && warp_live) {
        bits_lo = bits_hi = 0ull;
#pragma unroll 1
        for (int g8 = 0; g8 < (half >> 3); ++g8) {
          float d8[8];
          drop8(drop, SITE_CLS, 0, row, (c_begin >> 3) + g8, d8);
          unsigned long long m8 = 0ull;
#pragma unroll
          for (int j = 0; j < 8; ++j) m8 |= (d8[j] != 0.0f ? 1ull : 0ull) << j;
          if (g8 < 8) bits_lo |= m8 << (8 * g8);
          else bits_hi |= m8 << (8 * (g8 - 8));
        }
      }
      ring_skip(2 * KB);   // W1, W2 stages belong to the MMA warp
      acquire();
      HD_STAMP(2);
#pragma unroll 1
      for (int it = 0; it < (warp_live ? (half >> 4) : 0); ++it) {
        const int c = c_begin + it * 16;
        uint32_t acc[16];
        tmem_ld16_issue(tmem_base + lane_base + (uint32_t)c, acc);
        tmem_wait16(acc);
        const uint32_t k16 = (uint32_t)((it < 4 ? bits_lo >> (16 * it) : bits_hi >> (16 * (it - 4))) & 0xffffull);
        float v[16];
        uint32_t m16 = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = ((k16 >> j) & 1u) ? fmaxf(__uint_as_float(acc[j]) + S.b1s[c + j], 0.0f) * drop.scale : 0.0f;
          m16 |= (v[j] > 0.0f ? 1u : 0u) << j;
        }
        // keep & relu' replaces the keep bits of these 16 columns (m16 is a subset of k16)
        if (it < 4) bits_lo &= ~(0xffffull << (16 * it)) | ((unsigned long long)m16 << (16 * it));
        else bits_hi &= ~(0xffffull << (16 * (it - 4))) | ((unsigned long long)m16 << (16 * (it - 4)));
        st_swz16(S.ublk, trow, c, v);
      }
      publish();
      HD_STAMP(3);

      // ---- E2: logits, softmax / cross-entropy.  The warps that own the accumulator rows move the 32 logit
      // columns (+ bias) to shared memory; then ALL worker warps share the arithmetic, 8 threads per window
      // (4 columns each, reductions over the 8-lane group), 32 windows per pass. ----
      acquire();
      HD_STAMP(4);
      float* zs = hd_zs_floats(H) ? S.rmk + 128 * HD_RS : reinterpret_cast<float*>(S.ublk + HD_A_BYTES);   // [tile rows][33]
      if (cg == 0 && warp_live) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + lane_base + 256u, acc);
#pragma unroll
        for (int c = 0; c < 32; ++c) zs[trow * 33 + c] = __uint_as_float(acc[c]) + S.b2s[c];
      }
      tc_fence_before();
      workers_sync();
      {
        const int sub = lane & 7, c0 = sub * 4;           // this thread's 4 columns
        const float off = L.smoothing / (float)C * L.grad_scale, hit = (1.0f - L.smoothing) * L.grad_scale;
#pragma unroll 1
        for (int r0 = 0; r0 < L.tile_rows; r0 += 32) {
          const int r = r0 + wq * 4 + (lane >> 3);
          const long long grow = (long long)m0 + r;
          const bool ok = grow < L.rows;
          float z[4];
          float mx = -INFINITY, zsum = 0.0f;
          int arg = c0;
          bool has_nan = false;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            z[e] = zs[r * 33 + c0 + e];
            if (c0 + e < C) {
              if (ok) L.logits[grow * C + c0 + e] = z[e];
              has_nan |= (z[e] != z[e]);
              if (z[e] > mx) { mx = z[e]; arg = c0 + e; }   // first max inside the thread's columns
              zsum += z[e];
            }
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {                    // over the window's 8 threads; ties -> lowest column
            const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
            if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
            zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
            has_nan |= (__shfl_xor_sync(0xffffffffu, has_nan ? 1 : 0, o) != 0);
          }
          const int y = (L.train && ok) ? (int)L.labels[grow] : -1;
          float ex[4], se = 0.0f, zy = 0.0f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            ex[e] = (c0 + e < C) ? expf(z[e] - mx) : 0.0f;
            se += ex[e];
            if (c0 + e == y) zy = z[e];
          }
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            zy += __shfl_xor_sync(0xffffffffu, zy, o);
          }
          if (!L.train) {
            if (L.conf != nullptr && ok && sub == 0) {   // eval.py:89-90; a NaN logit gives (NaN, 0) like torch.max
              L.conf[grow] = has_nan ? nanf("") : 1.0f / se;
              L.pred[grow] = has_nan ? 0 : arg;
            }
          } else {
            if (ok && sub == 0) {
              const float lse = mx + logf(se);
              const float nll = lse - zy, smooth = lse - zsum / (float)C;   // mean_c(-log p_c)
              L.row_loss[grow] = (1.0f - L.smoothing) * nll + L.smoothing * smooth;
            }
            const float inv = ok ? L.grad_scale / se : 0.0f;
            float dl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              dl[e] = (c0 + e < C && ok) ? fmaf(ex[e], inv, -(off + (c0 + e == y ? hit : 0.0f))) : 0.0f;
            uint2 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
            h2[0] = __floats2bfloat162_rn(dl[0], dl[1]);
            h2[1] = __floats2bfloat162_rn(dl[2], dl[3]);
            // A operand of G3 (k < 32): 8-byte half of the 16-byte swizzle chunk
            *reinterpret_cast<uint2*>(S.ublk + swz_off(r, c0 & ~7) + ((c0 & 4) ? 8 : 0)) = pk;
            if (ok && c0 < L.Cp) *reinterpret_cast<uint2*>(L.dlog + grow * L.Cp + c0) = pk;
            // classifier.3 bias gradient: sum the warp's 4 windows into the CTA's shared-memory accumulators
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float cs = dl[e];
              cs += __shfl_xor_sync(0xffffffffu, cs, 8);
              cs += __shfl_xor_sync(0xffffffffu, cs, 16);
              if (lane < 8 && c0 + e < C) atomicAdd(S.db2s + c0 + e, cs);
            }
          }
        }
      }
      HD_STAMP(5);
      if (!L.train) {
        workers_sync();   // the next tile's P0 overwrites the per-row scalars and the A block
        continue;
      }
      publish();

      // ---- E3: dH1 = acc * relu'(Hr) * drop ----
      acquire();
      HD_STAMP(6);
#pragma unroll 1
      for (int it = 0; it < (warp_live ? (half >> 4) : 0); ++it) {
        const int c = c_begin + it * 16;
        uint32_t acc[16];
        tmem_ld16_issue(tmem_base + lane_base + (uint32_t)c, acc);
        tmem_wait16(acc);
        const uint32_t m16 = (uint32_t)((it < 4 ? bits_lo >> (16 * it) : bits_hi >> (16 * (it - 4))) & 0xffffull);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ((m16 >> j) & 1u) ? __uint_as_float(acc[j]) * drop.scale : 0.0f;
        st_swz16(S.ublk, trow, c, v);
      }
      publish();
      HD_STAMP(7);

      // ---- E4: dfused -> A block (read back by P5 through the generic proxy) ----
      acquire();
      HD_STAMP(8);
#pragma unroll 1
      for (int c = c_begin; c < (warp_live ? c_begin + half : c_begin); c += 16) {
        uint32_t acc[16];
        tmem_ld16_issue(tmem_base + lane_base + 256u + (uint32_t)c, acc);
        tmem_wait16(acc);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
        st_swz16(S.ublk, trow, c, v);
      }
      tc_fence_before();
      workers_sync();
      HD_STAMP(9);

      // ---- P5 ----
      ring_skip(1 + KB);   // W2^T, W1^T stages
      agg_pass(m0, true);
      workers_sync();
      HD_STAMP(10);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (L.train && (int)threadIdx.x < C) atomicAdd(L.db2 + threadIdx.x, S.db2s[threadIdx.x]);
  HD_STAMP(12);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }

  // ---- mean loss: the last CTA sums row_loss in a fixed order (deterministic) ----
  if (L.train && L.loss_out != nullptr) {
    __shared__ bool last;
    __shared__ double sh[HD_THREADS];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&g_head_ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    HD_STAMP(13);
    if (!last) return;
    __threadfence();
    double s = 0.0;
    for (long long i = threadIdx.x; i < L.rows; i += HD_THREADS) s += (double)__ldcg(L.row_loss + i);
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 128) sh[threadIdx.x] += sh[threadIdx.x + 256];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      L.loss_out[0] = (float)(sh[0] / (double)L.rows);
      g_head_ticket = 0;
    }
  }
}

size_t head_float_smem(int H, int M) {
  return (size_t)(H + 64 + M * H + HD_RS + 3 * 128 * HD_RS + hd_zs_floats(H)) * 4;
}

}  // namespace

bool head_eligible(int H, int M, int C) {
  return H % 64 == 0 && H >= 64 && H <= 256 && M >= 1 && M <= 4 && C >= 1 && C <= 32;
}

static int head_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      sms = 148;
  }
  return sms;
}

// small batches: 32-window tiles (the MMA tile is padded to 128 rows) so that the row-wise phases of more
// tiles run at the same time on otherwise idle SMs
int head_tile_rows(long long rows) {
  return (ceil_div(rows, 128) * 2 <= head_sms() && !getenv("MSF_HEAD_TILE128")) ? 32 : 128;
}

int head_launch(HeadLaunch& L, cudaStream_t stream, const char* label) {
  MSF_REQUIRE(head_eligible(L.H, L.M, L.C), "head_gemm: hidden %d / modalities %d / classes %d not supported", L.H,
              L.M, L.C);
  MSF_REQUIRE(L.rows >= 1, "head_gemm: empty batch");
  const int sms = head_sms();
  L.tile_rows = head_tile_rows(L.rows);
  L.row_tiles = (int)ceil_div(L.rows, L.tile_rows);
  const size_t fixed = 1024 + 8 * HD_NBAR + head_float_smem(L.H, L.M) + (size_t)(L.H / 64) * HD_A_BYTES;
  const size_t statics = 4096;   // the loss reduction's static shared memory
  int stages = (int)((HD_SMEM_LIMIT - statics - fixed) / hd_ring_bytes(L.H));
  if (stages > HD_MAX_STAGES) stages = HD_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "head_gemm: not enough shared memory for hidden %d", L.H);
  L.stages = stages;
  const size_t smem = fixed + (size_t)stages * hd_ring_bytes(L.H);
  const int grid = L.row_tiles < sms ? L.row_tiles : sms;
  if (prof_enabled()) {
    const double fwd = 2.0 * (double)L.rows * ((double)L.H * L.H + (double)L.H * L.C);
    prof_begin(label, L.train ? 2.0 * fwd : fwd, stream);
  }
  MSF_CHECK_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MSF_CHECK_CUDA(launch_pdl(head_kernel, dim3(grid), dim3(HD_THREADS), smem, stream, L));
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  return MSF_OK;
}

int head_debug_stamps(long long* out16) {
  MSF_CHECK_CUDA(cudaDeviceSynchronize());
  MSF_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_head_stamps, sizeof(long long) * 16));
  return MSF_OK;
}

}  // namespace msf
