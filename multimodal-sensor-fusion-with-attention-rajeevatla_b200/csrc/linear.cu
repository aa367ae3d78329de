// nn.Linear forward/backward on the fp32 FFMA grouped GEMM (simt_gemm.cu).
#include "simt_gemm.cuh"

namespace msf {
static SimtProblem base_problem() {
  SimtProblem p;
  memset(&p, 0, sizeof(p));
  p.nseg = 1; p.scale = 1.0f; p.head_dim = 1; p.heads = 1;
  return p;
}
static DropCfg no_drop() {
  DropCfg d;
  d.seed = 0; d.offset = 0; d.p = 0.f; d.scale = 1.f; d.active = 0; d.state = nullptr;
  return d;
}
}  // namespace msf

extern "C" {

int msf_linear_forward(const float* x, const float* w, const float* b, float* y, int64_t rows, int32_t in_dim,
                       int32_t out_dim, int32_t relu, void* stream) {
  MSF_REQUIRE(x && w && y && rows >= 0 && in_dim >= 1 && out_dim >= 1, "msf_linear_forward: bad arguments");
  MSF_REQUIRE(rows < (1ll << 31), "msf_linear_forward: too many rows");
  if (rows == 0) return MSF_OK;
  msf::SimtProblem p = msf::base_problem();
  p.A[0] = x; p.a_rs = in_dim; p.a_cs = 1;
  p.B[0] = w; p.b_rs = in_dim; p.b_cs = 1;
  p.bias[0] = b;
  p.M = (int)rows; p.N = out_dim; p.K = in_dim;
  p.C = y; p.ldc = out_dim;
  p.epi = relu ? msf::EPI_BIAS_RELU_DROP : msf::EPI_STORE;
  return msf::simt_gemm_launch(&p, 1, msf::no_drop(), (cudaStream_t)stream);
}

int msf_linear_backward(const float* x, const float* w, const float* y, const float* dy, float* dy_scratch,
                        float* dx, float* dw, float* db, int64_t rows, int32_t in_dim, int32_t out_dim,
                        int32_t relu, void* stream) {
  MSF_REQUIRE(x && w && dy && dw && rows >= 0 && in_dim >= 1 && out_dim >= 1, "msf_linear_backward: bad arguments");
  MSF_REQUIRE(!relu || (y && dy_scratch), "msf_linear_backward: relu needs y and dy_scratch");
  MSF_REQUIRE(rows < (1ll << 31), "msf_linear_backward: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  const msf::DropCfg nd = msf::no_drop();
  int rc;
  if (rows == 0) {
    MSF_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)in_dim * out_dim, st));
    if (db) MSF_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)out_dim, st));
    return MSF_OK;
  }
  const float* g = dy;
  if (relu) {  // g = dy * (y > 0): an identity-matrix-free elementwise pass via the K=0 epilogue
    msf::SimtProblem p = msf::base_problem();
    p.A[0] = dy; p.B[0] = w; p.a_rs = out_dim; p.a_cs = 1; p.b_rs = in_dim; p.b_cs = 1;
    p.M = (int)rows; p.N = out_dim; p.K = 0;
    p.C = dy_scratch; p.ldc = out_dim;
    p.epi = msf::EPI_ADD_RELU_GRAD; p.scale = 1.0f;
    p.aux = dy; p.ld_aux = out_dim; p.aux2 = y; p.ld_aux2 = out_dim;
    if ((rc = msf::simt_gemm_launch(&p, 1, nd, st))) return rc;
    g = dy_scratch;
  }
  msf::SimtProblem probs[2];
  int n = 0;
  {
    msf::SimtProblem p = msf::base_problem();  // dw[o,i] = sum_r g[r,o] x[r,i]
    p.A[0] = g; p.a_rs = 1; p.a_cs = out_dim;
    p.B[0] = x; p.b_rs = 1; p.b_cs = in_dim;
    p.M = out_dim; p.N = in_dim; p.K = (int)rows;
    p.C = dw; p.ldc = in_dim; p.epi = msf::EPI_STORE;
    probs[n++] = p;
  }
  if (dx) {
    msf::SimtProblem p = msf::base_problem();  // dx[r,i] = sum_o g[r,o] w[o,i]
    p.A[0] = g; p.a_rs = out_dim; p.a_cs = 1;
    p.B[0] = w; p.b_rs = 1; p.b_cs = in_dim;
    p.M = (int)rows; p.N = in_dim; p.K = out_dim;
    p.C = dx; p.ldc = in_dim; p.epi = msf::EPI_STORE;
    probs[n++] = p;
  }
  if ((rc = msf::simt_gemm_launch(probs, n, nd, st))) return rc;
  if (db) {
    msf::ColsumProblem c{g, out_dim, (int)rows, out_dim, db};
    if ((rc = msf::colsum_launch(&c, 1, st))) return rc;
  }
  return MSF_OK;
}

}  // extern "C"
