// fp32 FFMA grouped GEMM (64x64x16 tiles, 256 threads, 4x4 register blocking)
// with the HybridFusion epilogues fused in.  See simt_gemm.cuh.
#include "simt_gemm.cuh"

namespace msf {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;  // LDS: padded smem row (floats)

__device__ __forceinline__ float ldg_guard(const float* p, bool ok) { return ok ? __ldg(p) : 0.0f; }

// Stage one (rows x BK) operand tile into smem as T[k][row].
//  kcontig: element (row, k) at base[row*rs + k]  (cs == 1)
//  else   : element (row, k) at base[k*cs + row]  (rs == 1)
__device__ __forceinline__ void load_tile(float (*T)[LDS], const float* __restrict__ base, long long rs,
                                          long long cs, int row0, int k0, int rows, int K, int tid) {
  if (cs == 1) {
    const int k = tid & 15, r = tid >> 4;
    const bool kok = (k0 + k) < K;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = r + 16 * j;
      const bool ok = kok && (row0 + row) < rows;
      T[k][row] = ldg_guard(base + (long long)(row0 + row) * rs + (k0 + k), ok);
    }
  } else {
    const int row = tid & 63, kq = tid >> 6;
    const bool rok = (row0 + row) < rows;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kq + 4 * j;
      const bool ok = rok && (k0 + k) < K;
      T[k][row] = ldg_guard(base + (long long)(k0 + k) * cs + (long long)(row0 + row) * rs, ok);
    }
  }
}

__global__ void __launch_bounds__(256) simt_gemm_kernel(const __grid_constant__ SimtProblemList list) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];

  // locate this block's problem (<= 28 entries, block-uniform scan)
  int pi = 0;
  const int tile = blockIdx.x;
  while (pi + 1 < list.count && tile >= list.p[pi + 1].tile_begin) ++pi;
  const SimtProblem& P = list.p[pi];
  const int local = tile - P.tile_begin;
  const int tiles_n = (P.N + BN - 1) / BN;
  const int m0 = (local / tiles_n) * BM, n0 = (local % tiles_n) * BN;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int s = 0; s < P.nseg; ++s) {
    const float* __restrict__ A = P.A[s];
    const float* __restrict__ B = P.B[s];
    for (int k0 = 0; k0 < P.K; k0 += BK) {
      load_tile(As, A, P.a_rs, P.a_cs, m0, k0, P.M, P.K, tid);
      load_tile(Bs, B, P.b_rs, P.b_cs, n0, k0, P.N, P.K, tid);
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue -------------------------------------------------------------
  const DropCfg drop = resolve_drop(list.drop);
  const int col0 = n0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = col0 + j;
    if (col < P.N)
      for (int s = 0; s < P.nseg; ++s)
        if (P.bias[s]) bias[j] += __ldg(P.bias[s] + col);
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= P.M) continue;
    float v[4];
    float dm[4] = {1.f, 1.f, 1.f, 1.f};
    const bool need_drop = drop.active && (P.epi == EPI_BIAS_RELU_DROP || P.epi == EPI_DX);
    if (need_drop && col0 < P.N) drop4(drop, P.site, P.sub, row, col0 >> 2, dm);
    const float mrow = (P.mask != nullptr) ? __ldg(P.mask + (long long)row * P.mask_ld + P.mask_col) : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + j;
      if (col >= P.N) { v[j] = 0.f; continue; }
      const float a = acc[i][j];
      float r;
      switch (P.epi) {
        default:
        case EPI_STORE:
          r = a * P.scale + bias[j];
          break;
        case EPI_BIAS_RELU_DROP:
          r = fmaxf(a + bias[j], 0.0f) * dm[j];
          break;
        case EPI_VALUE_GATE: {
          const int head = col / P.head_dim;
          float g = (mrow != 0.0f) ? 1.0f : 0.0f;  // softmax over one key: 1, or NaN->0 when masked
          if (drop.active) g *= drop1(drop, SITE_ATTN, P.sub, row, head);
          if (P.gate_out != nullptr && (col % P.head_dim) == 0)
            P.gate_out[(long long)row * P.heads + head] = g;
          r = (a + bias[j]) * g;
          break;
        }
        case EPI_OUT_MEAN:
          r = (a + bias[j] + __ldg(P.aux + (long long)row * P.ld_aux + col)) / P.scale * mrow;
          break;
        case EPI_RELU_GRAD:
          r = a * ((__ldg(P.aux + (long long)row * P.ld_aux + col) > 0.0f) ? P.scale : 0.0f);
          break;
        case EPI_GATE_MUL:
          r = a * __ldg(P.gate_in + (long long)row * P.heads + col / P.head_dim);
          break;
        case EPI_ADD_RELU_GRAD:
          r = (a + __ldg(P.aux + (long long)row * P.ld_aux + col)) *
              ((__ldg(P.aux2 + (long long)row * P.ld_aux2 + col) > 0.0f) ? P.scale : 0.0f);
          break;
        case EPI_DX:
          r = a * mrow * dm[j];
          break;
      }
      v[j] = r;
    }
    float* dst = P.C + (long long)row * P.ldc + col0;
    if (col0 + 3 < P.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col0 + j < P.N) dst[j] = v[j];
    }
  }
}

// column sums: dst[c] = sum_r src[r*ld + c].  grid = (ceil(cols/32), problem)
struct ColsumList {
  ColsumProblem p[SIMT_MAX_PROBLEMS];
  int count;
};

__global__ void __launch_bounds__(256) colsum_kernel(const __grid_constant__ ColsumList list) {
  const ColsumProblem& P = list.p[blockIdx.y];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  if (blockIdx.x * 32 >= P.cols) return;
  const int w = threadIdx.x >> 5;
  float s = 0.0f;
  if (col < P.cols)
    for (int r = w; r < P.rows; r += 8) s += __ldg(P.src + (long long)r * P.ld + col);
  __shared__ float red[8][33];
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && col < P.cols) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    P.dst[col] = t;
  }
}

}  // namespace

int simt_gemm_launch(const SimtProblem* problems, int count, const DropCfg& drop, cudaStream_t stream) {
  int done = 0;
  while (done < count) {
    SimtProblemList list;
    list.drop = drop;
    const int n = (count - done) < SIMT_MAX_PROBLEMS ? (count - done) : SIMT_MAX_PROBLEMS;
    int tiles = 0, kept = 0;
    for (int i = 0; i < n; ++i) {
      SimtProblem p = problems[done + i];
      if (p.M <= 0 || p.N <= 0) continue;
      if (p.nseg < 1 || p.nseg > SIMT_MAX_SEG) {
        set_error("simt_gemm: nseg %d out of range", p.nseg);
        return MSF_E_INVALID;
      }
      if (!((p.a_cs == 1) || (p.a_rs == 1)) || !((p.b_cs == 1) || (p.b_rs == 1))) {
        set_error("simt_gemm: operands need one unit stride");
        return MSF_E_INVALID;
      }
      p.tile_begin = tiles;
      tiles += (int)(ceil_div(p.M, BM) * ceil_div(p.N, BN));
      list.p[kept++] = p;
    }
    list.count = kept;
    list.total_tiles = tiles;
    if (tiles > 0) {
      simt_gemm_kernel<<<tiles, 256, 0, stream>>>(list);
      MSF_LAUNCH_CHECK();
    }
    done += n;
  }
  return MSF_OK;
}

int colsum_launch(const ColsumProblem* problems, int count, cudaStream_t stream) {
  int done = 0;
  while (done < count) {
    ColsumList list;
    const int n = (count - done) < SIMT_MAX_PROBLEMS ? (count - done) : SIMT_MAX_PROBLEMS;
    int maxc = 0;
    for (int i = 0; i < n; ++i) {
      list.p[i] = problems[done + i];
      if (list.p[i].cols > maxc) maxc = list.p[i].cols;
    }
    list.count = n;
    if (maxc > 0) {
      dim3 grid((unsigned)ceil_div(maxc, 32), (unsigned)n);
      colsum_kernel<<<grid, 256, 0, stream>>>(list);
      MSF_LAUNCH_CHECK();
    }
    done += n;
  }
  return MSF_OK;
}

}  // namespace msf
