// Weight-gradient GEMMs on CTA pairs: dW[M, N] = dY[rows, M]^T . X[rows, N] over the windows (MN-major operands),
// one 256 x 256 output tile per cluster of two CTAs (tcgen05.mma.cta_group::2), the contraction split in two
// halves that meet in the destination.  See wg2_gemm.cu.
#pragma once

#include <cuda.h>

#include "msf_common.cuh"

namespace msf {

constexpr int WG2_MAX_PROBLEMS = 40;
constexpr int WG2_MAX_MAPS = 24;
constexpr int WG2_MAX_FLAG_TILES = 96;   // split launches keep one ticket word per (tile, CTA rank, epilogue warp)

struct Wg2Problem {
  short a_map, b_map;   // indices into Wg2Launch::maps: dY [z][rows][M], X [z][rows][N]
  int a_z, b_z;
  int M, N;             // dW is M x N (row-major, ldc elements, fp32)
  float* C;
  long long ldc;
  double* sq;           // optional: the sum of squares of the finished dW is added here
  int tile_begin;
};

struct Wg2Launch {
  CUtensorMap maps[WG2_MAX_MAPS];
  Wg2Problem p[WG2_MAX_PROBLEMS];
  int count, nmaps;
  int kb_total;      // ceil(rows / 64)
  int splits;        // 1 or 2 halves of the contraction per tile
  int kb_split;      // k-blocks per half
  int total_tiles;
  unsigned* flags;   // splits == 2: WG2_MAX_FLAG_TILES * 16 words, zero between launches
};

// One launch: a list of weight-gradient problems over a shared table of TMA descriptors (64 x 64 boxes,
// 128B swizzle), all contracting over the same `rows` windows.
struct Wg2Builder {
  Wg2Launch L;
  long long rows;
  cudaStream_t stream;
  int status;
  const char* label;
  Wg2Builder(long long rows, cudaStream_t st, const char* label);
  int add_map(const void* base, long long rows, long long cols, long long ld, long long depth, long long slice);
  int add_problem(short a_map, int a_z, short b_map, int b_z, int M, int N, float* C, long long ldc, double* sq);
  int flush();
};

// dW of N input features with this leading dimension can be produced by wg2_kernel
bool wg2_shape_ok(int N, long long ldc, const void* C);
bool wg2_enabled();   // MSF_WG=v1 selects the single-CTA tc_gemm_kernel<1> path instead

}  // namespace msf
