// Grouped fp32 FFMA GEMM with fused epilogues — the MSF_PREC_F32 (parity) path.
// One launch runs a list of independent problems; each problem may sum several
// K-segments (different A and B per segment) into one accumulator, which is how
// aggregated[q] = P_q + sum_k out_proj(q,k)(...) is produced in a single pass.
#pragma once

#include "msf_common.cuh"

namespace msf {

constexpr int SIMT_MAX_SEG = MSF_MAX_MODALITIES;
constexpr int SIMT_MAX_PROBLEMS = 28;

enum SimtEpilogue {
  EPI_STORE = 0,          // v = acc*scale + bias
  EPI_BIAS_RELU_DROP,     // v = relu(acc + bias) * drop(site, sub, row, col)
  EPI_VALUE_GATE,         // v = (acc + bias) * gate(row, head); records gate
  EPI_OUT_MEAN,           // v = (acc + sum bias + aux) / scale * mask[row, mask_col]
  EPI_RELU_GRAD,          // v = acc * (aux > 0 ? scale : 0)
  EPI_GATE_MUL,           // v = acc * gate_in[row, head]
  EPI_ADD_RELU_GRAD,      // v = (acc + aux) * (aux2 > 0 ? scale : 0)
  EPI_DX                  // v = acc * mask[row, mask_col] * drop(site, sub, row, col)
};

struct SimtProblem {
  const float* A[SIMT_MAX_SEG];
  const float* B[SIMT_MAX_SEG];
  const float* bias[SIMT_MAX_SEG];
  int nseg;
  int M, N, K;
  long long a_rs, a_cs;   // A(m,k) = A[m*a_rs + k*a_cs]
  long long b_rs, b_cs;   // B(n,k) = B[n*b_rs + k*b_cs]
  float* C;
  long long ldc;
  int epi;
  float scale;
  const float* aux;
  long long ld_aux;
  const float* aux2;
  long long ld_aux2;
  const float* mask;      // (rows, mask_ld) or nullptr (= ones)
  int mask_ld, mask_col;
  float* gate_out;        // (rows, heads)
  const float* gate_in;
  int head_dim, heads;
  int site, sub;
  int tile_begin;         // filled by the launcher
};

struct SimtProblemList {
  SimtProblem p[SIMT_MAX_PROBLEMS];
  int count;
  int total_tiles;
  DropCfg drop;
};

// Enqueue all problems (splits into several launches if > SIMT_MAX_PROBLEMS).
int simt_gemm_launch(const SimtProblem* problems, int count, const DropCfg& drop, cudaStream_t stream);

struct ColsumProblem {
  const float* src;
  long long ld;
  int rows, cols;
  float* dst;
};
int colsum_launch(const ColsumProblem* problems, int count, cudaStream_t stream);

}  // namespace msf
