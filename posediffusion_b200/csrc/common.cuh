// Shared device/host helpers for the posediff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdb {

constexpr int kTargetDim = 9;
constexpr int kMaxFrames = 128;

// ---------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce 16 per-lane values across the 32 lanes of a warp with 16 shuffles (instead of 80):
// at every butterfly step each lane hands half of its slots to its partner.  On return lane L
// (both lanes of a pair L, L^1 hold the same value) owns slot  ((L>>4)&1)*8 + ((L>>3)&1)*4 +
// ((L>>2)&1)*2 + ((L>>1)&1)  in v[0].  Summation order is fixed -> deterministic.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? v[i] : v[i + 8];
      float keep = hi ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? v[i] : v[i + 4];
      float keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? v[i] : v[i + 2];
      float keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    float send = hi ? v[0] : v[1];
    float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
__device__ __forceinline__ int warp_reduce16_slot(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ---------------------------------------------------------------------------------------------
// Memory helpers, mbarrier + bulk asynchronous copy
// ---------------------------------------------------------------------------------------------
#if defined(PDB_EMU) && defined(__CUDACC__)
#error "PDB_EMU is the CPU test harness of tests/host/cuda_emu.h; it must never be defined in an nvcc (product) build"
#endif
#ifdef PDB_EMU
// TEST HARNESS ONLY (tests/host/cuda_emu.h): the inline-PTX helpers below restated for the CPU emulation that runs these
// kernels unmodified on the host.  Never defined in a product build (nvcc does not see PDB_EMU).
inline float4 ld_stream_f4(const float4* p) { return *p; }
inline unsigned ld_acquire_u32(const unsigned* p) {
  sched_yield();  // the poller spins on another OS thread's progress
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
inline unsigned ld_relaxed_u32(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void red_release_add_u32(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
inline uint32_t smem_u32(const void* p) { return emu::shared_addr(p); }
inline void mbar_init(uint32_t bar, uint32_t count) { emu::mbar_init(bar, count); }
inline void mbar_fence_init() {}
inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { emu::mbar_arrive_expect_tx(bar, bytes); }
inline bool mbar_try_wait(uint32_t bar, uint32_t parity) { return emu::mbar_try_wait(bar, parity); }
inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
inline void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) { emu::bulk_copy_g2s(dst_smem, src, bytes, bar); }
#else
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// mbarrier + bulk asynchronous copy (global -> shared, completion counted in bytes on an mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy executed by the TMA unit: `bytes` (multiple of 16) from global to this CTA's shared memory
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
#endif

// The same barrier executed by ONE warp per CTA (the rest of the CTA waits at a later __syncthreads).
// All 32 lanes may have issued global atomics before; __syncwarp orders them before lane 0's release.
__device__ __forceinline__ void warp_group_barrier(unsigned* counter, unsigned target) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    red_release_add_u32(counter, 1u);
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncwarp();
}

}  // namespace pdb
