// chain2_kernel: the chained pair GEMMs of chain_gemm.cuh, software-pipelined at HALF-pair granularity
// so that the tensor pipe and the epilogue warps work at the same time (hidden % 128 == 0).
//
// One pair (T = A1 W1^T -> epilogue1 -> ACC += T' W2^T) is split along the columns of the intermediate:
//
//     half-chain h = (pair i, half b):   T_b   = A1[inner] . W1[pair][b*HW:(b+1)*HW, :]^T      (N = HW = H/2)
//                                        T'_b  = epilogue1(T_b) -> bf16 -> staging block u[h&1]
//                                        ACC  += T'_b . W2[pair][:, b*HW:(b+1)*HW]^T            (K = HW)
//
// TMEM holds two intermediate buffers T[0], T[1] (HW columns each) next to ACC (H columns), shared memory
// two staging blocks u[0], u[1]; the MMA warp issues  G1(0) | G1(1) G2(0) | G1(2) G2(1) | ... | G2(n-1),
// so while the epilogue warps turn T[h&1] into u[h&1] the tensor pipe already runs G2(h-1) and G1(h+1).
// Roles as in chain_kernel: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 TMA
// store, warps 4..11 epilogue (thread = accumulator row).  All mbarrier waits are bounded.
#include "chain_gemm.cuh"

#include "tc_ptx.cuh"

namespace msf {

namespace {

constexpr int C2_THREADS = 384;
constexpr int C2_EPI_WARPS = 8;
constexpr int C2_MAX_STAGES = 5;
constexpr uint32_t C2_A_BYTES = 128 * 64 * 2;  // one k-block of an A operand: 128 rows x 64 bf16
constexpr size_t C2_SMEM_LIMIT = 232448;
constexpr int C2_NBAR = 2 * C2_MAX_STAGES + 12;

struct Raw16 {
  uint4 lo, hi;
};
__device__ __forceinline__ Raw16 ld_row16(const __nv_bfloat16* p, bool ok) {
  Raw16 r;
  r.lo = make_uint4(0u, 0u, 0u, 0u);
  r.hi = r.lo;
  if (ok) {
    r.lo = __ldg(reinterpret_cast<const uint4*>(p));
    r.hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  }
  return r;
}
__device__ __forceinline__ void unpack16(const Raw16& r, float (&out)[16]) {
  const uint32_t w[8] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, r.hi.z, r.hi.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    out[2 * e] = __uint_as_float(w[e] << 16);
    out[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}
// 16 results of row `trow`, block columns [c, c+16) -> K-major 128B-swizzled k-blocks starting at blk
__device__ __forceinline__ void st_swizzled16(unsigned char* blk, int trow, int c, const float (&v)[16]) {
  unsigned char* rowp = blk + (c >> 6) * C2_A_BYTES + trow * 128;
  const int j0 = (c & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(rowp + (((j0 + q) ^ (trow & 7)) << 4)) = pk;
  }
}

// wait-time accounting of CTA 0 (clock64 cycles): msf_debug_chain_stamps()
//  [0] kernel start  [1] MMA: sum of waits for operand stages  [2] MMA: waits for T drained  [3] MMA: waits for u staged
//  [4] MMA loop end  [5] epilogue warp 4: waits for G1  [6] epilogue: waits for u free  [7] epilogue: compute
//  [8] epilogue: final wait  [9] epilogue end  [10] producer: waits for free stages  [11] producer end
__device__ long long g_chain_stamps[16];
#define C2_T0() const long long _t0 = clock64()
#define C2_ACC(i) do { dbg_acc[(i) & 3] += clock64() - _t0; } while (0)      /* register accumulators */
#define C2_FLUSH(i) do { if (blockIdx.x == 0) g_chain_stamps[i] = dbg_acc[(i) & 3]; } while (0)
#define C2_SET(i) do { if (blockIdx.x == 0) g_chain_stamps[i] = clock64(); } while (0)

template <int MODE>
__global__ void __launch_bounds__(C2_THREADS, 1) chain2_kernel(const __grid_constant__ ChainLaunch L) {
  TL_KERNEL(MODE);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t off0 = smem_u32(smem_raw);
  const uint32_t smem_base = (off0 + 1023u) & ~1023u;
  const int H = L.H, HW = L.H >> 1, KB = L.H >> 6, KBH = L.H >> 7, STAGES = L.stages, M = L.M;
  // G1: A k-block (16 KB) + W1 half k-block (HW rows); G2: W2 k-block (H rows)
  const uint32_t STAGE = (C2_A_BYTES + (uint32_t)HW * 128u > (uint32_t)H * 128u) ? C2_A_BYTES + (uint32_t)HW * 128u
                                                                                   : (uint32_t)H * 128u;
  const uint32_t UB = (uint32_t)KBH * C2_A_BYTES;         // one staging block
  const uint32_t ring_base = smem_base;
  const uint32_t u_base = ring_base + STAGES * STAGE;     // u[0], u[1] contiguous: together the final output tile
  const uint32_t bar_base = u_base + 2 * UB;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C2_MAX_STAGES + s); };
  auto t_full = [&](int b) { return bar_base + 8u * (2 * C2_MAX_STAGES + 0 + b); };    // G1 finished (MMA -> epilogue)
  auto t_empty = [&](int b) { return bar_base + 8u * (2 * C2_MAX_STAGES + 2 + b); };   // T drained (epilogue -> MMA)
  auto us_ready = [&](int b) { return bar_base + 8u * (2 * C2_MAX_STAGES + 4 + b); };  // u written (epilogue -> MMA, store)
  auto us_free = [&](int b) { return bar_base + 8u * (2 * C2_MAX_STAGES + 6 + b); };   // u reusable (MMA + store -> epilogue)
  const uint32_t acc_full = bar_base + 8u * (2 * C2_MAX_STAGES + 8);
  const uint32_t acc_empty = bar_base + 8u * (2 * C2_MAX_STAGES + 9);
  const uint32_t out_ready = bar_base + 8u * (2 * C2_MAX_STAGES + 10);
  const uint32_t tmem_slot = bar_base + 8u * (2 * C2_MAX_STAGES + 11);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - off0));
  float* bias_smem = reinterpret_cast<float*>(smem_raw + (bar_base + 8u * C2_NBAR - off0));  // (inner + 1) x H
  unsigned char* u_smem = smem_raw + (u_base - off0);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * H <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&L.map_a1);
    tma_prefetch_desc(&L.map_w1);
    tma_prefetch_desc(&L.map_w2);
    tma_prefetch_desc(&L.map_out1);
    tma_prefetch_desc(&L.map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full(b), 1);
      mbar_init(t_empty(b), C2_EPI_WARPS);
      mbar_init(us_ready(b), C2_EPI_WARPS);
      mbar_init(us_free(b), 2);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, C2_EPI_WARPS);
    mbar_init(out_ready, C2_EPI_WARPS);
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
  pdl_wait();     TL_WAITED(MODE);  // the prologue above overlapped the previous kernel; its outputs are visible from here on
  // The forward chain lets its successor (the head kernel) become resident now.  The backward chain does not: its
  // successor is the weight-gradient GEMM, whose early blocks would sit in pdl_wait() on the 20 SMs this grid
  // leaves idle for the whole kernel, which is where the column sums beside it are meant to run; without the
  // trigger the successor starts when this grid has drained (a trigger from the MMA thread once its last MMA is
  // issued hides the successor's launch latency but measured the same: 154.4 vs 154.6 us per step).
  if (MODE != 1) pdl_launch();
  long long dbg_acc[4] = {0, 0, 0, 0};
  if (blockIdx.x == 0 && threadIdx.x == 0) g_chain_stamps[0] = clock64();

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next_stage = [&]() {
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int item = blockIdx.x; item < L.items; item += gridDim.x) {
        const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
        const int nh = 2 * L.outer[o].n;
        auto load_g1 = [&](int h) {
          const int i = h >> 1, b = h & 1;
          const int zin = L.outer[o].inner[i], zp = L.outer[o].pair[i];
          for (int kb = 0; kb < KB; ++kb) {
            { C2_T0(); mbar_wait(empty_bar(stage), phase ^ 1u); C2_ACC(10); }
            mbar_expect_tx(full_bar(stage), C2_A_BYTES + (uint32_t)HW * 128u);
            tma_load_3d(ring_base + stage * STAGE, &L.map_a1, kb * 64, m0, zin, full_bar(stage));
            tma_load_3d(ring_base + stage * STAGE + C2_A_BYTES, &L.map_w1, kb * 64, b * HW, zp, full_bar(stage));
            next_stage();
          }
        };
        auto load_g2 = [&](int h) {
          const int i = h >> 1, b = h & 1;
          const int zp = L.outer[o].pair[i];
          for (int kk = 0; kk < KBH; ++kk) {
            { C2_T0(); mbar_wait(empty_bar(stage), phase ^ 1u); C2_ACC(10); }
            mbar_expect_tx(full_bar(stage), (uint32_t)H * 128u);
            tma_load_3d(ring_base + stage * STAGE, &L.map_w2, (b * KBH + kk) * 64, 0, zp, full_bar(stage));
            next_stage();
          }
        };
        if (nh > 0) {
          load_g1(0);
          for (int h = 1; h < nh; ++h) {
            load_g1(h);
            load_g2(h - 1);
          }
          load_g2(nh - 1);
        }
      }
      C2_FLUSH(10);
      C2_SET(11);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      const uint32_t idesc_h = instr_desc(H, false, false), idesc_hw = instr_desc(HW, false, false);
      int stage = 0;
      uint32_t phase = 0, item_cnt = 0;
      uint32_t t_uses[2] = {0u, 0u}, u_uses[2] = {0u, 0u};
      auto next_stage = [&]() {
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      };
      for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
        const int nh = 2 * L.outer[L.active[item % L.n_active]].n;
        mbar_wait(acc_empty, (item_cnt & 1u) ^ 1u);
        tc_fence_after();
        auto g1 = [&](int h) {   // T[b] = A1 . W1half^T
          const int b = h & 1;
          { C2_T0(); mbar_wait(t_empty(b), (t_uses[b] & 1u) ^ 1u); C2_ACC(2); }
          ++t_uses[b];
          tc_fence_after();
          const uint32_t d = tmem_base + (uint32_t)(b * HW);
          for (int kb = 0; kb < KB; ++kb) {
            { C2_T0(); mbar_wait(full_bar(stage), phase); C2_ACC(1); }
            tc_fence_after();
            const uint32_t a_addr = ring_base + stage * STAGE, b_addr = a_addr + C2_A_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(d, smem_desc(a_addr + k * 32, 16, 1024), smem_desc(b_addr + k * 32, 16, 1024), idesc_hw,
                          (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            next_stage();
          }
          tc_commit(t_full(b));
        };
        auto g2 = [&](int h) {   // ACC += u[b] . W2[:, half]^T
          const int b = h & 1;
          { C2_T0(); mbar_wait(us_ready(b), u_uses[b] & 1u); C2_ACC(3); }
          ++u_uses[b];
          tc_fence_after();
          for (int kk = 0; kk < KBH; ++kk) {
            { C2_T0(); mbar_wait(full_bar(stage), phase); C2_ACC(1); }
            tc_fence_after();
            const uint32_t a_addr = u_base + b * UB + kk * C2_A_BYTES, b_addr = ring_base + stage * STAGE;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_base + (uint32_t)H, smem_desc(a_addr + k * 32, 16, 1024),
                          smem_desc(b_addr + k * 32, 16, 1024), idesc_h, (h > 0 || kk > 0 || k > 0) ? 1u : 0u);
            tc_commit(empty_bar(stage));
            next_stage();
          }
          tc_commit(us_free(b));  // 1 of 2: GEMM2 no longer reads the staging block
        };
        if (nh > 0) {
          g1(0);
          for (int h = 1; h < nh; ++h) {
            g1(h);
            g2(h - 1);
          }
          g2(nh - 1);
        }
        tc_commit(acc_full);
      }
      C2_FLUSH(1); C2_FLUSH(2); C2_FLUSH(3);
      C2_SET(4);
    }
  } else if (warp == 3) {
    // =========================== TMA store ==============================
    if (lane == 0) {
      uint32_t u_uses[2] = {0u, 0u}, item_cnt = 0;
      for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
        const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
        const int nh = 2 * L.outer[o].n;
        for (int h = 0; h < nh; ++h) {
          const int b = h & 1;
          mbar_wait(us_ready(b), u_uses[b] & 1u);
          ++u_uses[b];
          if (L.store1) {
            for (int kk = 0; kk < KBH; ++kk)
              tma_store_3d(&L.map_out1, u_base + b * UB + kk * C2_A_BYTES, (b * KBH + kk) * 64, m0,
                           L.outer[o].pair[h >> 1]);
            tma_store_commit();
            tma_store_wait_read();
          }
          mbar_arrive(us_free(b));  // 2 of 2
        }
        mbar_wait(out_ready, item_cnt & 1u);
        for (int kb = 0; kb < KB; ++kb) tma_store_3d(&L.map_out, u_base + kb * C2_A_BYTES, kb * 64, m0, o);
        tma_store_commit();
        tma_store_wait_read();
        for (int b = 0; b < 2; ++b) {  // the final tile has no GEMM2 reader: both arrivals come from here
          mbar_arrive(us_free(b));
          mbar_arrive(us_free(b));
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =========================== epilogue ===============================
    const DropCfg drop = resolve_drop(L.drop);
    const int lq = warp & 3, cg = (warp - 4) >> 2;
    const int trow = lq * 32 + lane;                 // row inside the tile = TMEM lane
    const int et = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
    const int hc_begin = cg * (HW >> 1), hc_end = hc_begin + (HW >> 1);   // this thread's columns inside a half
    uint32_t t_uses[2] = {0u, 0u}, f_uses[2] = {0u, 0u}, item_cnt = 0;
    for (int item = blockIdx.x; item < L.items; item += gridDim.x, ++item_cnt) {
      const int o = L.active[item % L.n_active], m0 = (item / L.n_active) * 128;
      const int n = L.outer[o].n, nh = 2 * n;
      const long long row = (long long)m0 + trow;
      const bool row_ok = row < L.rows;

      // bias rows of this item -> shared memory: value_proj bias per pair, then the summed out_proj biases
      if (MODE == 0) {
        if (item_cnt > 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * C2_EPI_WARPS) : "memory");  // previous readers done
        for (int e = et; e < (n + 1) * H; e += 32 * C2_EPI_WARPS) {
          const int i = e / H, cc = e % H;
          float val = 0.0f;
          if (i < n) {
            const float* bsrc = L.bias1[L.outer[o].pair[i]];
            val = bsrc ? __ldg(bsrc + cc) : 0.0f;
          } else {
            for (int j = 0; j < n; ++j) {
              const float* bsrc = L.bias2[L.outer[o].pair[j]];
              if (bsrc) val += __ldg(bsrc + cc);
            }
            for (int j = 0; j < L.outer[o].nb; ++j) {
              const float* bsrc = L.bias2[L.outer[o].bias_only[j]];
              if (bsrc) val += __ldg(bsrc + cc);
            }
          }
          bias_smem[e] = val;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * C2_EPI_WARPS) : "memory");
      }

      for (int h = 0; h < nh; ++h) {
        const int i = h >> 1, b = h & 1;
        const int zp = L.outer[o].pair[i];
        const float* bias_s = bias_smem + i * H + b * HW;
        float gate_mask = 1.0f;
        if (MODE == 0 && L.mask != nullptr && row_ok) gate_mask = __ldg(L.mask + row * M + L.outer[o].mask_col[i]);
        const int sub = L.outer[o].sub[i];
        float* gate_out = (MODE == 0 && L.gate_out) ? L.gate_out + ((long long)zp * L.rows + row) * L.heads : nullptr;
        const float* gate_in = (MODE == 1) ? L.gate_in + ((long long)zp * L.rows + row) * L.heads : nullptr;
        unsigned char* ub = u_smem + b * UB;

        { C2_T0(); mbar_wait(t_full(b), t_uses[b] & 1u); C2_ACC(5); }
        ++t_uses[b];
        tc_fence_after();
        { C2_T0(); mbar_wait(us_free(b), (f_uses[b] & 1u) ^ 1u); C2_ACC(6); }
        ++f_uses[b];
        const long long _tc = clock64();
        int cur_head = -1;
        float cur_gate = 0.0f;
        if ((L.head_dim & 15) == 0) {
          // 32 accumulator columns per TMEM load / wait (half as many round trips as x16); every 16-column group
          // lies inside one head
#pragma unroll 1
          for (int c = hc_begin; c < hc_end; c += 32) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(b * HW + c), acc);
#pragma unroll
            for (int g16 = 0; g16 < 2; ++g16) {
              const int cc = c + 16 * g16, ca = b * HW + cc;   // column of the full intermediate
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
              float v[16];
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == 0) b4 = *reinterpret_cast<const float4*>(bias_s + cc + 4 * q4);
                v[4 * q4 + 0] = (__uint_as_float(acc[16 * g16 + 4 * q4 + 0]) + b4.x) * cur_gate;
                v[4 * q4 + 1] = (__uint_as_float(acc[16 * g16 + 4 * q4 + 1]) + b4.y) * cur_gate;
                v[4 * q4 + 2] = (__uint_as_float(acc[16 * g16 + 4 * q4 + 2]) + b4.z) * cur_gate;
                v[4 * q4 + 3] = (__uint_as_float(acc[16 * g16 + 4 * q4 + 3]) + b4.w) * cur_gate;
              }
              st_swizzled16(ub, trow, cc, v);
            }
          }
        } else {
#pragma unroll 1
          for (int c = hc_begin; c < hc_end; c += 16) {   // narrow heads (head_dim not a multiple of 16): per-column gate
            const int ca = b * HW + c;
            uint32_t acc[16];
            tmem_ld16_issue(tmem_base + lane_base + (uint32_t)(b * HW + c), acc);
            tmem_wait16(acc);
            float v[16];
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
              const int head = (ca + j) / L.head_dim;
              float g;
              if (MODE == 0) {
                g = (gate_mask != 0.0f) ? 1.0f : 0.0f;
                if (drop.active) g *= drop1(drop, SITE_ATTN, sub, row, head);
                if (gate_out != nullptr && row_ok && ca + j == head * L.head_dim) gate_out[head] = g;
              } else {
                g = row_ok ? __ldg(gate_in + head) : 0.0f;
              }
              const float bb = (MODE == 0) ? bias_s[c + j] : 0.0f;
              float a = 0.0f;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                if (jj == j) a = __uint_as_float(acc[jj]);
              const float r = (a + bb) * g;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj)
                if (jj == j) v[jj] = r;
            }
            st_swizzled16(ub, trow, c, v);
          }
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(t_empty(b));
          mbar_arrive(us_ready(b));
        }
        dbg_acc[3] += clock64() - _tc;
      }

      // ---- final epilogue over ACC (columns [H, 2H) of TMEM) into both staging blocks ----
      const float* bias_s = bias_smem + n * H;
      const int c_begin = cg * HW, c_end = c_begin + HW;
      const __nv_bfloat16* aux_row = L.aux + ((long long)o * L.rows + row) * H;
      const __nv_bfloat16* aux2_row = (MODE == 1) ? L.aux2 + ((long long)o * L.rows + row) * H : nullptr;
      float rscale = 1.0f;
      if (MODE == 0) rscale = L.inv_cnt[o] * ((L.mask != nullptr && row_ok) ? __ldg(L.mask + row * M + o) : 1.0f);
      Raw16 nxt = ld_row16(aux_row + c_begin, row_ok), nxt2;
      if (MODE == 1) nxt2 = ld_row16(aux2_row + c_begin, row_ok);

      { C2_T0(); mbar_wait(acc_full, item_cnt & 1u); C2_ACC(8); }
      tc_fence_after();
      for (int b = 0; b < 2; ++b) {
        mbar_wait(us_free(b), (f_uses[b] & 1u) ^ 1u);
        ++f_uses[b];
      }
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 16) {
        uint32_t acc[16];
        if (n > 0) {
          tmem_ld16_issue(tmem_base + lane_base + (uint32_t)(H + c), acc);
          tmem_wait16(acc);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0u;  // no pair module: nothing was accumulated
        }
        float aux[16], aux2[16], v[16];
        unpack16(nxt, aux);
        const bool more = c + 16 < c_end;
        nxt = ld_row16(aux_row + c + 16, row_ok && more);
        if (MODE == 1) {
          unpack16(nxt2, aux2);
          nxt2 = ld_row16(aux2_row + c + 16, row_ok && more);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = __uint_as_float(acc[j]);
          if (MODE == 0) v[j] = rscale != 0.0f ? (a + bias_s[c + j] + aux[j]) * rscale : 0.0f;  // masked row: exact 0
          else v[j] = (a + aux[j]) * (aux2[j] > 0.0f ? L.scale : 0.0f);
        }
        st_swizzled16(u_smem, trow, c, v);
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_empty);
        mbar_arrive(out_ready);
      }
      if (threadIdx.x == 128) { C2_FLUSH(5); C2_FLUSH(6); C2_FLUSH(7); C2_FLUSH(8); C2_SET(9); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

size_t chain2_fixed_smem(int H) { return 1024 + 8 * C2_NBAR + (size_t)(CHAIN_MAX_INNER + 1) * H * 4; }

}  // namespace

bool chain2_eligible(int H, int M) { return H % 128 == 0 && H >= 128 && H <= 256 && M >= 1 && M <= MSF_MAX_MODALITIES; }

int chain2_launch(ChainLaunch& L, cudaStream_t stream, const char* label) {
  MSF_REQUIRE(chain2_eligible(L.H, L.M), "chain2_gemm: hidden %d / modalities %d not supported", L.H, L.M);
  MSF_REQUIRE(L.rows >= 1, "chain2_gemm: empty batch");
  L.row_tiles = (int)ceil_div(L.rows, 128);
  if (L.n_active <= 0) {   // default: every modality is an outer modality
    L.n_active = L.M;
    for (int m = 0; m < L.M; ++m) L.active[m] = (short)m;
  }
  L.items = L.row_tiles * L.n_active;
  const size_t g1_stage = C2_A_BYTES + (size_t)(L.H / 2) * 128, g2_stage = (size_t)L.H * 128;
  const size_t stage = g1_stage > g2_stage ? g1_stage : g2_stage;
  const size_t ublocks = (size_t)(L.H / 64) * C2_A_BYTES;
  int stages = (int)((C2_SMEM_LIMIT - chain2_fixed_smem(L.H) - ublocks) / stage);
  if (stages > C2_MAX_STAGES) stages = C2_MAX_STAGES;
  MSF_REQUIRE(stages >= 2, "chain2_gemm: not enough shared memory for hidden %d", L.H);
  L.stages = stages;
  const size_t smem = chain2_fixed_smem(L.H) + ublocks + (size_t)stages * stage;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    MSF_CHECK_CUDA(cudaGetDevice(&dev));
    MSF_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = L.items < sms ? L.items : sms;
  if (prof_enabled()) {
    double pairs = 0.0;
    for (int o = 0; o < L.M; ++o) pairs += L.outer[o].n;
    prof_begin(label, 2.0 * 2.0 * (double)L.rows * L.H * L.H * pairs, stream, prof_repeat());
  }
  // the kernel only reads its operands and overwrites its outputs: repeating it (profiling) changes nothing
  for (int rep = prof_repeat(); rep > 0; --rep) {
    if (L.mode == 0) {
      MSF_CHECK_CUDA(cudaFuncSetAttribute(chain2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSF_CHECK_CUDA(launch_pdl(chain2_kernel<0>, dim3(grid), dim3(C2_THREADS), smem, stream, L));
    } else {
      MSF_CHECK_CUDA(cudaFuncSetAttribute(chain2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MSF_CHECK_CUDA(launch_pdl(chain2_kernel<1>, dim3(grid), dim3(C2_THREADS), smem, stream, L));
    }
  }
  MSF_LAUNCH_CHECK();
  prof_end(stream);
  return MSF_OK;
}

int chain_debug_stamps(long long* out16) {
  MSF_CHECK_CUDA(cudaDeviceSynchronize());
  MSF_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_chain_stamps, sizeof(long long) * 16));
  return MSF_OK;
}

}  // namespace msf
