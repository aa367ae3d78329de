// Library-level entry points: version, error string, device check, arena layout.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "msf_common.cuh"

namespace msf {

namespace {
struct ProfRecord {
  const char* label;
  double flops;
  int reps;
  cudaEvent_t start, stop;
};
bool g_prof_on = false;
int g_prof_reps = 1;
std::vector<ProfRecord> g_prof;
}  // namespace

bool prof_enabled() { return g_prof_on; }
int prof_repeat() { return g_prof_on ? g_prof_reps : 1; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MSF_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

void prof_begin(const char* label, double flops, cudaStream_t stream, int reps) {
  if (!g_prof_on) return;
  ProfRecord r;
  r.label = label;
  r.flops = flops;
  r.reps = reps < 1 ? 1 : reps;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, stream);
  g_prof.push_back(r);
}

void prof_end(cudaStream_t stream) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().stop, stream);
}

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_layout(const msf_fusion_shape* s, Layout* L) {
  MSF_REQUIRE(s != nullptr && L != nullptr, "null shape");
  MSF_REQUIRE(s->num_modalities >= 1 && s->num_modalities <= MSF_MAX_MODALITIES,
              "num_modalities %d out of range [1, %d]", s->num_modalities, MSF_MAX_MODALITIES);
  MSF_REQUIRE(s->hidden >= 1 && s->num_heads >= 1 && s->hidden % s->num_heads == 0,
              "hidden_dim (%d) must be divisible by num_heads (%d)", s->hidden, s->num_heads);
  MSF_REQUIRE(s->num_classes >= 1, "num_classes must be >= 1");
  L->M = s->num_modalities;
  L->H = s->hidden;
  L->heads = s->num_heads;
  L->C = s->num_classes;
  L->present = s->pair_present;
  int64_t off = 0;
  const int64_t H = L->H;
  for (int m = 0; m < L->M; ++m) {
    MSF_REQUIRE(s->in_dims[m] >= 1, "in_dims[%d] must be >= 1", m);
    L->D[m] = s->in_dims[m];
    L->proj_w[m] = off;
    off += H * L->D[m];
    L->proj_b[m] = off;
    off += H;
  }
  L->pair_base = off;
  L->pair_stride = 4 * (H * H + H);
  off += (int64_t)L->num_pairs() * L->pair_stride;
  for (int m = 0; m < L->M; ++m) {
    L->gate_w[m] = off;
    off += H;
    L->gate_b[m] = off;
    off += 1;
  }
  L->cls_w1 = off;
  off += H * H;
  L->cls_b1 = off;
  off += H;
  L->cls_w2 = off;
  off += (int64_t)L->C * H;
  L->cls_b2 = off;
  off += L->C;
  L->total = off;
  return MSF_OK;
}

}  // namespace msf

#ifdef MSF_TIMELINE
namespace msf {
struct TlEntry {
  const char* tu;
  tl_dump_fn fn;
};
static TlEntry* tl_entries() {
  static TlEntry e[64];
  return e;
}
static int& tl_count() {
  static int n = 0;
  return n;
}
void tl_register(const char* tu, tl_dump_fn fn) {
  if (tl_count() < 64) tl_entries()[tl_count()++] = {tu, fn};
}
}  // namespace msf
#endif

extern "C" {

int msf_abi_version(void) { return MSF_ABI_VERSION; }

int msf_struct_sizes(int32_t* shape_bytes, int32_t* call_bytes) {
  if (shape_bytes) *shape_bytes = (int32_t)sizeof(msf_fusion_shape);
  if (call_bytes) *call_bytes = (int32_t)sizeof(msf_fusion_call);
  return MSF_OK;
}

int msf_lstm_seq_bytes(void) { return (int)sizeof(msf_lstm_seq); }

uint64_t msf_launch_count(void) { return msf::g_launch_count; }

const char* msf_last_error(void) { return msf::g_err; }

int msf_prof_enable(int32_t on) {
  for (auto& r : msf::g_prof) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  msf::g_prof.clear();
  msf::g_prof_on = on != 0;
  msf::g_prof_reps = on > 1 ? on : 1;
  return MSF_OK;
}

int msf_prof_report(char* buf, size_t cap) {
  MSF_REQUIRE(buf != nullptr && cap > 0, "msf_prof_report: no buffer");
  MSF_CHECK_CUDA(cudaDeviceSynchronize());
  struct Agg { std::string label; long long n; double ms, flops; };
  std::vector<Agg> agg;
  for (auto& r : msf::g_prof) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) continue;
    Agg* a = nullptr;
    for (auto& x : agg)
      if (x.label == r.label) a = &x;
    if (!a) {
      agg.push_back(Agg{r.label, 0, 0.0, 0.0});
      a = &agg.back();
    }
    a->n += r.reps;
    a->ms += ms;
    a->flops += r.flops * r.reps;
  }
  std::string out;
  char line[256];
  for (auto& x : agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.0f\n", x.label.c_str(), x.n, x.ms, x.flops);
    out += line;
  }
  if (out.size() + 1 > cap) {
    msf::set_error("msf_prof_report: buffer too small (%zu needed)", out.size() + 1);
    return MSF_E_INVALID;
  }
  memcpy(buf, out.c_str(), out.size() + 1);
  return MSF_OK;
}

int msf_device_check(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  MSF_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MSF_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (prop.major < 10) {
    msf::set_error("device '%s' is sm_%d%d; this library is built for sm_100a only", prop.name,
                   prop.major, prop.minor);
    return MSF_E_UNSUPPORTED;
  }
  return MSF_OK;
}

int msf_fusion_param_count(const msf_fusion_shape* shape, int64_t* count) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(count != nullptr, "null count");
  *count = L.total;
  return MSF_OK;
}

int msf_fusion_param_offset(const msf_fusion_shape* shape, int32_t kind, int32_t idx, int64_t* offset) {
  msf::Layout L;
  int rc = msf::make_layout(shape, &L);
  if (rc) return rc;
  MSF_REQUIRE(offset != nullptr, "null offset");
  const int M = L.M;
  if (kind == 0 || kind == 1 || kind == 10 || kind == 11) {
    MSF_REQUIRE(idx >= 0 && idx < M, "modality index %d out of range", idx);
    *offset = kind == 0 ? L.proj_w[idx] : kind == 1 ? L.proj_b[idx] : kind == 10 ? L.gate_w[idx] : L.gate_b[idx];
    return MSF_OK;
  }
  if (kind >= 2 && kind <= 9) {
    const int q = idx / M, k = idx % M;
    MSF_REQUIRE(idx >= 0 && q < M && q != k, "pair index %d invalid", idx);
    const int p = L.pair_index(q, k);
    const int which = (kind - 2) / 2;
    *offset = ((kind - 2) % 2 == 0) ? L.pair_w(p, which) : L.pair_b(p, which);
    return MSF_OK;
  }
  if (kind >= 12 && kind <= 15) {
    *offset = kind == 12 ? L.cls_w1 : kind == 13 ? L.cls_b1 : kind == 14 ? L.cls_w2 : L.cls_b2;
    return MSF_OK;
  }
  msf::set_error("unknown tensor kind %d", kind);
  return MSF_E_INVALID;
}

int msf_memcpy_batch(void* const* dst, const void* const* src, const size_t* bytes, int32_t n, void* stream) {
  MSF_REQUIRE(dst && src && bytes && n >= 0, "msf_memcpy_batch: bad arguments");
  for (int i = 0; i < n; ++i)
    if (bytes[i] > 0)
      MSF_CHECK_CUDA(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDefault, (cudaStream_t)stream));
  return MSF_OK;
}


#ifdef MSF_TIMELINE
// debug builds only (msf_common.cuh): prints every translation unit's step-timeline table and clears it
int msf_debug_timeline(char* buf, size_t cap, int reset) {
  size_t used = 0;
  if (cap) buf[0] = 0;
  for (int i = 0; i < msf::tl_count(); ++i) {
    static unsigned long long t[32 + 1024];
    if (msf::tl_entries()[i].fn(t, reset) != 0) return MSF_E_CUDA;
    for (int k = 0; k < 4; ++k) {
      if (t[k * 8 + 5] == 0) continue;
      const char* tu = strrchr(msf::tl_entries()[i].tu, '/');
      int n = snprintf(buf + used, cap - used, "%s#%d\t%llu\t%llu\t%llu\t%llu\t%llu\t%llu\t%llu\t%llu\n", tu ? tu + 1 : msf::tl_entries()[i].tu,
                       k, t[k * 8 + 0], t[k * 8 + 1], t[k * 8 + 2], t[k * 8 + 3], t[k * 8 + 4], t[k * 8 + 5], t[k * 8 + 6], t[k * 8 + 7]);
      if (n < 0 || (size_t)n >= cap - used) return MSF_E_INVALID;
      used += (size_t)n;
      if (t[32 + k * 256] == 0) continue;
      n = snprintf(buf + used, cap - used, "%s#%d.cta", tu ? tu + 1 : msf::tl_entries()[i].tu, k);
      used += (size_t)n;
      for (int b = 0; b < 256 && t[32 + k * 256 + b] != 0; ++b) {
        if (cap - used < 64) return MSF_E_INVALID;
        used += (size_t)snprintf(buf + used, cap - used, "\t%llu:%llu", t[32 + k * 256 + b] >> 8, t[32 + k * 256 + b] & 255ull);
      }
      used += (size_t)snprintf(buf + used, cap - used, "\n");
    }
  }
  return MSF_OK;
}
#endif

}  // extern "C"
