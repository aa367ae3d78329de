// Signal-block layout (uint64 words) and system-scope access helpers shared by the data-parallel
// kernels (dp_optim.cu) and the fused optimizer (opt_pack.cu).
#pragma once

#include "msf_common.cuh"

namespace msf {

constexpr int SIG_DONE_B = 0;    // [p]: rank p has finished the backward pass of step s
constexpr int SIG_DONE_R = 8;    // [p]: rank p has finished reducing its slice of step s
constexpr int SIG_NORM = 16;     // [p]: square-norm of rank p's slice (double bits)
constexpr int SIG_DONE_W = 24;   // [p]: rank p has pushed its share of the updated bf16 weights everywhere (sharded step)
constexpr int SIG_ACC = 32;      // local accumulator of the slice norm (double)
constexpr int SIG_TICKET = 33;   // local last-block ticket (reduce kernel)
constexpr int SIG_TICKET2 = 34;  // local last-block ticket (reduce kernel, phase 2)
constexpr int SIG_TICKET3 = 35;  // local last-block ticket (update kernel)
constexpr int SIG_EPOCH = 40;    // local count of completed data-parallel steps: the barrier epoch.  Kept apart from
                                 // the Adam step counter, which callers may roll back (graph warm-up).
constexpr int SIG_TIME = 48;     // [0..5]: %globaltimer stamps of the last step (reduce start / barrier passed / end,
                                 // update start / barrier passed / end) — cheap always-on instrumentation

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// data a peer pushed into this GPU's memory: system-scope load, never served from a stale L1 line
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// NVLink multicast (NVLS): one address that stands for the same offset in every rank's buffer.  A load-reduce is
// answered by the switch with the sum over the ranks' copies; a store is delivered to every rank's copy.
__device__ __forceinline__ float4 mm_ld_reduce4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ float mm_ld_reduce1(const float* mc) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f32(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void mm_st_bf16x4(void* mc, uint2 v) {   // 4 bf16 = 8 bytes
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(mc), "r"(v.x), "r"(v.y) : "memory");
}

// every thread of the block returns once all ranks have published `epoch` in words [base, base + world)
__device__ __forceinline__ void wait_all(const unsigned long long* sig, int base, int world, unsigned long long epoch) {
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(sig + base + threadIdx.x) < epoch) {
      if (clock64() - t0 > 8000000000ll) {
        printf("msf_b200 dp_optim: rank wait timed out (word %d, peer %d, epoch %llu)\n", base, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

}  // namespace msf
