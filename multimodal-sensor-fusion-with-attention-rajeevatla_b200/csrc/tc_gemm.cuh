// Grouped bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory) with the
// HybridFusion epilogues fused in.  sm_100a only.
//
// One persistent launch processes a list of problems; each problem may sum up
// to 7 K-segments (different A / B per segment) into one TMEM accumulator.
// Operands are addressed through a small table of TMA descriptors over
// *stacked* tensors ([z][rows][cols]); a segment names (descriptor, z).
//
//   MAJOR_K  (forward / dgrad):  A[M,K] and B[N,K] row-major, K contiguous
//   MAJOR_MN (wgrad):            A = dY[Kb,M], B = X[Kb,N] row-major, contraction
//                                 over the rows (windows), M / N contiguous
#pragma once

#include <cuda.h>

#include "msf_common.cuh"

namespace msf {

constexpr int TC_MAX_SEG = 7;
constexpr int TC_MAX_PROBLEMS = 40;
constexpr int TC_MAX_MAPS = 24;
constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int TC_MAX_STAGES = 6;   // operand pipeline depth is chosen per launch from the shared memory left
constexpr int TC_EPI_COLGROUPS = 2;                 // epilogue warps per TMEM lane quarter
constexpr int TC_EPI_WARPS = 4 * TC_EPI_COLGROUPS;
constexpr int TC_THREADS = 128 + 32 * TC_EPI_WARPS; // TMA, MMA, TMEM-alloc, spare + epilogue warps

enum TcEpilogue {
  TC_EPI_STORE = 0,        // v = acc*scale + bias
  TC_EPI_BIAS_RELU_DROP,   // v = relu(acc + bias) * drop(site, sub, row, col)
  TC_EPI_VALUE_GATE,       // v = (acc + bias) * gate(row, head); records gate
  TC_EPI_OUT_MEAN,         // v = (acc + sum bias + aux) / scale * mask[row, mask_col]
  TC_EPI_RELU_GRAD,        // v = acc * (aux > 0 ? scale : 0)
  TC_EPI_GATE_MUL,         // v = acc * gate_in[row, head]
  TC_EPI_ADD_RELU_GRAD,    // v = (acc + aux) * (aux2 > 0 ? scale : 0)
  TC_EPI_DX,               // v = acc * mask[row, mask_col] * drop(site, sub, row, col)
  TC_EPI_LSTM              // LSTM cell: columns are gate-interleaved (4*unit + {i,f,g,o}); updates `cell`, writes h
};

struct TcSegment {
  short a_map, b_map;   // indices into TcLaunch::maps
  int a_z, b_z;         // slice of the stacked tensor
};

struct TcProblem {
  TcSegment seg[TC_MAX_SEG];
  const float* bias[TC_MAX_SEG];  // fp32, summed
  int nseg;
  int M, N, K;            // output rows, output cols, contraction length per segment
  void* C;                // row-major, ld elements
  long long ldc;
  int c_bf16;             // 1: bf16 output, 0: fp32 output
  int epi;
  float scale;
  const __nv_bfloat16* aux;
  long long ld_aux;
  const __nv_bfloat16* aux2;
  long long ld_aux2;
  const float* mask;
  int mask_ld, mask_col;
  float* gate_out;
  const float* gate_in;
  int head_dim, heads;
  int site, sub;
  int tile_begin;
  // TC_EPI_LSTM (N = 4 * hidden): c_t = f*c + i*g in place in `cell` (fp32, rows x hidden); h_t = o*tanh(c_t) goes
  // to C as bf16 in k-block-major layout [hidden/64][rows][64] (slice pitch h_slice elements: the next step's A
  // operand) and, if h32 != nullptr, to h32 (fp32, rows x hidden)
  float* cell;
  float* h32;
  long long h_slice;
  // TC_EPI_STORE: if set, the sum of squares of everything this problem stores is added here (the weight-gradient
  // GEMM hands the optimizer its share of the gradient norm)
  double* sq;
};

struct TcLaunch {
  CUtensorMap maps[TC_MAX_MAPS];
  TcProblem p[TC_MAX_PROBLEMS];
  int count;
  int nmaps;
  int stages;             // operand pipeline depth of this launch
  int stage_pitch;        // bytes between rows of an epilogue warp's staging block
  int dbg;                // debug switches (MSF_TC_DEBUG): 1 = skip stores, 2 = skip TMEM loads
  int total_tiles;
  int block_n;            // 32..256, multiple of 16; uniform per launch
  DropCfg drop;
};

// Host-side builder: collects problems over a shared table of TMA descriptors and
// launches them; more than TC_MAX_PROBLEMS problems are split over several
// launches (the descriptor table is kept).
struct TcBuilder {
  TcLaunch L;
  int nmaps;
  bool mn_major;
  cudaStream_t stream;
  int status;             // first error seen (MSF_OK otherwise)
  const char* label;      // shown by msf_prof_report
  TcBuilder(bool mn, int block_n, const DropCfg& drop, cudaStream_t st, const char* label = "tc_gemm");
  // Stacked bf16 tensor [depth][rows][cols] (row pitch = ld elements, slice pitch = slice elements).
  // role_rows: box height along `rows` for a K-major operand (TC_BLOCK_M for A, block_n for B);
  // ignored for MN-major operands (box = 64 x 64).
  int add_map(const void* base, long long rows, long long cols, long long ld, long long depth,
              long long slice, int role_rows);
  int add_problem(const TcProblem& p);
  int flush(bool keep = false);   // launch what has been collected so far (keep: leave the problems in place
                                  // so that the same launch can be patched and issued again)
};

TcProblem tc_blank_problem();
DropCfg no_dropout();

int tc_init();  // resolves cuTensorMapEncodeTiled; MSF_OK or error
// 3-D bf16 tensor map over [depth][rows][cols] (row pitch ld, slice pitch slice elements), 128B swizzle,
// box = box_cols x box_rows x 1 (box_cols * 2 bytes must be 128)
int tc_encode_map(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, long long depth,
                  long long slice, int box_cols, int box_rows);

}  // namespace msf
