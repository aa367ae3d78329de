// Shared helpers for the msf_b200 kernels: error plumbing, parameter-arena
// layout, Philox4x32-10 dropout draws.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/msf_b200.h"

namespace msf {

// ---------------------------------------------------------------------------
// error plumbing (never throw across the C ABI)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define MSF_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::msf::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,             \
                       cudaGetErrorString(_e));                                       \
      return MSF_E_CUDA;                                                              \
    }                                                                                 \
  } while (0)

#define MSF_REQUIRE(cond, ...)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::msf::set_error(__VA_ARGS__);                                                  \
      return MSF_E_INVALID;                                                           \
    }                                                                                 \
  } while (0)

// every kernel launch of the library is counted (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
#define MSF_LAUNCH_CHECK()                  \
  do {                                      \
    ++::msf::g_launch_count;                \
    MSF_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

// Optional per-launch timing with CUDA events on the launching stream (msf_prof_enable / msf_prof_report);
// bench.py uses it to time the dominant kernel live.  Not usable while a stream is being captured.
// msf_prof_enable(R > 1): launchers whose kernel is idempotent issue it R times back to back inside one event
// bracket (prof_repeat()), so the average launch duration includes what a programmatic dependent launch hides
// in the real step (the prologue under the predecessor's tail) and not the cost of the event records.
bool prof_enabled();
int prof_repeat();
void prof_begin(const char* label, double flops, cudaStream_t stream, int reps = 1);
void prof_end(cudaStream_t stream);

// ---------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor on the
// stream is still running; it runs its prologue (barrier init, TMEM allocation, descriptor prefetch, constants
// written at least two kernels earlier), then pdl_wait() blocks until the predecessor has completed and its
// writes are visible.  Every such kernel calls pdl_launch() only AFTER pdl_wait(), so at most two kernels
// overlap and whatever a prologue reads was written by a kernel that has already completed.
// Works eagerly and under stream capture (programmatic graph edges).  MSF_PDL=0 turns it off.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

// ---------------------------------------------------------------------------
// Step timeline (debug builds only, -DMSF_TIMELINE): per kernel, the earliest CTA entry, the earliest / latest
// return from pdl_wait() and the earliest / latest CTA end, in %globaltimer ns.  One table per translation
// unit, registered with api.cu at load time; msf_debug_timeline() prints and clears them.
// ---------------------------------------------------------------------------
#ifdef MSF_TIMELINE
static __device__ unsigned long long g_tl[4][8];
static __device__ unsigned long long g_tl_cta[4][256];   // per CTA (grids of <= 256): end time << 8 | SM id
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct TlScope {
  int id;
  __device__ __forceinline__ explicit TlScope(int i) : id(i) {
    if (threadIdx.x == blockDim.x - 1 && threadIdx.y == 0) atomicMin(&g_tl[id][0], tl_now());
  }
  __device__ __forceinline__ ~TlScope() {
    if (threadIdx.x == blockDim.x - 1 && threadIdx.y == 0) {
      const unsigned long long t = tl_now();
      atomicMin(&g_tl[id][3], t);
      atomicMax(&g_tl[id][4], t);
      atomicAdd(&g_tl[id][5], 1ull);
      if (gridDim.x <= 256 && blockIdx.y == 0) {
        unsigned int sm;
        asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        g_tl_cta[id][blockIdx.x] = (t << 8) | (sm & 255u);
      }
    }
  }
};
__device__ __forceinline__ void tl_waited(int id) {
  if (threadIdx.x == blockDim.x - 1 && threadIdx.y == 0) {
    const unsigned long long t = tl_now();
    atomicMin(&g_tl[id][1], t);
    atomicMax(&g_tl[id][2], t);
  }
}
__device__ __forceinline__ void tl_mark(int id, int slot) {   // slot 6 or 7: latest time any CTA passed the mark
  if (threadIdx.x == blockDim.x - 1 && threadIdx.y == 0) atomicMax(&g_tl[id][slot], tl_now());
}
// per-CTA mark under a kernel id of its own (one thread per CTA calls it once): shows up as "<tu>#<id>" with the
// earliest / latest mark in the "CTA end" columns and every CTA's time in the .cta line
__device__ __forceinline__ void tl_cta_mark(int id) {
  const unsigned long long t = tl_now();
  atomicMin(&g_tl[id][0], t);
  atomicMin(&g_tl[id][3], t);
  atomicMax(&g_tl[id][4], t);
  atomicAdd(&g_tl[id][5], 1ull);
  if (gridDim.x <= 256) {
    unsigned int sm;
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    g_tl_cta[id][blockIdx.x] = (t << 8) | (sm & 255u);
  }
}
typedef int (*tl_dump_fn)(unsigned long long*, int);
void tl_register(const char* tu, tl_dump_fn fn);
static int tl_dump_local(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, g_tl, sizeof(unsigned long long) * 32) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(out + 32, g_tl_cta, sizeof(unsigned long long) * 1024) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long init[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 8; ++j) init[i][j] = (j == 0 || j == 1 || j == 3) ? ~0ull : 0ull;
    if (cudaMemcpyToSymbol(g_tl, init, sizeof(init)) != cudaSuccess) return 1;
    static unsigned long long zeros[1024];
    if (cudaMemcpyToSymbol(g_tl_cta, zeros, sizeof(zeros)) != cudaSuccess) return 1;
  }
  return 0;
}
namespace {
struct TlRegistrar {
  TlRegistrar() { tl_register(__BASE_FILE__, tl_dump_local); }
};
static TlRegistrar tl_registrar_instance;
}  // namespace
#define TL_KERNEL(id) TlScope tl_scope_(id)
#define TL_WAITED(id) tl_waited(id)
#define TL_MARK(id, slot) tl_mark(id, slot)
#define TL_CTA_MARK(id) tl_cta_mark(id)
#else
#define TL_KERNEL(id)
#define TL_WAITED(id)
#define TL_MARK(id, slot)
#define TL_CTA_MARK(id)
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// master-arena layout (see include/msf_b200.h; src/fusion.py:291-328 order)
// ---------------------------------------------------------------------------
struct Layout {
  int M, H, heads, C;
  int D[MSF_MAX_MODALITIES];
  uint64_t present;
  int64_t proj_w[MSF_MAX_MODALITIES], proj_b[MSF_MAX_MODALITIES];
  int64_t pair_base;   // first attention module
  int64_t pair_stride; // 4*(H*H+H)
  int64_t gate_w[MSF_MAX_MODALITIES], gate_b[MSF_MAX_MODALITIES];
  int64_t cls_w1, cls_b1, cls_w2, cls_b2;
  int64_t total;

  __host__ __device__ int num_pairs() const { return M * (M - 1); }
  // q-major index among ordered pairs q != k
  __host__ __device__ int pair_index(int q, int k) const { return q * (M - 1) + (k < q ? k : k - 1); }
  __host__ __device__ bool has_pair(int q, int k) const { return (present >> (q * M + k)) & 1ull; }
  // which: 0 query, 1 key, 2 value, 3 out
  __host__ __device__ int64_t pair_w(int p, int which) const {
    return pair_base + (int64_t)p * pair_stride + (int64_t)which * ((int64_t)H * H + H);
  }
  __host__ __device__ int64_t pair_b(int p, int which) const { return pair_w(p, which) + (int64_t)H * H; }
  // number of tensors averaged into aggregated[q]: itself + present pair modules (fusion.py:403-407)
  __host__ __device__ int mean_count(int q) const {
    int c = 1;
    for (int k = 0; k < M; ++k) c += (k != q && has_pair(q, k)) ? 1 : 0;
    return c;
  }
};

int make_layout(const msf_fusion_shape* s, Layout* out);

// ---------------------------------------------------------------------------
// Philox4x32-10 (counter-based; forward and backward regenerate the same draw)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct DropCfg {
  uint64_t seed, offset;
  float p;         // drop probability
  float scale;     // 1/(1-p)
  int active;      // training && p > 0
  const unsigned long long* state;  // optional device {seed, offset}: overrides the two fields above
};

// Resolve a device-resident RNG state once per kernel (block-uniform).
__device__ __forceinline__ DropCfg resolve_drop(DropCfg d) {
  if (d.state != nullptr) {
    d.seed = d.state[0];
    d.offset = d.state[1];
  }
  return d;
}

// Eight multipliers for elements (row, col8*8 .. col8*8+7) of a dropout site: one
// Philox4x32-10 call yields 8 x 16-bit uniforms; an element is kept when its
// uniform u16 >= round(p * 65536) (keep probability 1 - p to within 2^-16).
__host__ __device__ __forceinline__ void drop8(const DropCfg& d, int site, int sub, int64_t row,
                                               int col8, float (&out)[8]) {
  uint32_t c[4] = {(uint32_t)row, (uint32_t)col8, ((uint32_t)site << 24) | (uint32_t)sub,
                   (uint32_t)d.offset};  // rows < 2^32
  philox4x32_10(c, (uint32_t)d.seed, (uint32_t)(d.seed >> 32) ^ (uint32_t)(d.offset >> 32));
  const uint32_t thr = (uint32_t)(d.p * 65536.0f + 0.5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = ((c[i] & 0xffffu) >= thr) ? d.scale : 0.0f;
    out[2 * i + 1] = ((c[i] >> 16) >= thr) ? d.scale : 0.0f;
  }
}

// Four multipliers for elements (row, col4*4 .. col4*4+3): one half of a drop8 group.
__host__ __device__ __forceinline__ void drop4(const DropCfg& d, int site, int sub, int64_t row,
                                               int col4, float (&out)[4]) {
  float v[8];
  drop8(d, site, sub, row, col4 >> 1, v);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = (col4 & 1) ? v[4 + i] : v[i];
}

// Single multiplier for element (row, col).
__host__ __device__ __forceinline__ float drop1(const DropCfg& d, int site, int sub, int64_t row, int col) {
  if (!d.active) return 1.0f;
  float v[8];
  drop8(d, site, sub, row, col >> 3, v);
  float r = v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = ((col & 7) == i) ? v[i] : r;
  return r;
}

enum { SITE_INPUT = 0, SITE_PROJ = 1, SITE_ATTN = 2, SITE_CLS = 3 };

// Learning rate of an optimizer launch: the host value, or (host value < 0) the float whose bits sit in the low
// word of train_state[3] — a captured CUDA graph then follows a schedule without being re-captured
// (msf_train_state_set_lr; src/train.py:395-404 steps a cosine schedule per epoch).
__device__ __forceinline__ float resolve_lr(float lr, const unsigned long long* train_state) {
  return (lr < 0.0f && train_state != nullptr) ? __uint_as_float((uint32_t)train_state[3]) : lr;
}

}  // namespace msf
