// Shared device helpers for the lwse kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "../../include/lwse.h"

namespace lwse {

constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs

// L2 eviction policy for the small tables every tick reads again (node table and its index,
// placement requests, occupancy): evict_last, so that the tens of MB a sweep streams through L2
// go first.  (createpolicy; the asm is not volatile, so identical requests in one kernel fold
// into one instruction.)  On C3 with 126 MB of L2 this measures the same as plain loads (tick
// 22.7-22.9 us either way); it is a guard for bigger sweeps, not a speed-up.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// 128-bit streaming load: read-only path, do not allocate in L1 (each row is consumed once per
// sweep).  (An L2 evict_first hint on top was measured slower — the group pass re-reads the state
// words of event pods, which then come from DRAM again: fused kernel 9.6 -> 10.8 us, tick 22.7 -> 27.2 us.)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Loads of small tables that every tick reads again: last to go in L2.
__device__ __forceinline__ uint4 ldg_keep(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy_evict_last()));
  return r;
}
__device__ __forceinline__ uint32_t ldg_keep_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy_evict_last()));
  return r;
}

// 128-bit load through L1 (rows shared by neighbouring tiles: owner rows, node rows).
__device__ __forceinline__ uint4 ldg_cached(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) {
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

// Lane mask of the W-wide tile this lane belongs to.
template <int W>
__device__ __forceinline__ uint32_t tile_mask() {
  if constexpr (W == 32) {
    return 0xFFFFFFFFu;
  } else {
    const uint32_t lane = threadIdx.x & 31u;
    return ((1u << W) - 1u) << (lane & ~(uint32_t)(W - 1));
  }
}

template <int W>
__device__ __forceinline__ uint32_t tile_or(uint32_t v) {
  if constexpr (W == 1) return v;
  return __reduce_or_sync(tile_mask<W>(), v);
}
template <int W>
__device__ __forceinline__ uint32_t tile_min(uint32_t v) {
  if constexpr (W == 1) return v;
  return __reduce_min_sync(tile_mask<W>(), v);
}
template <int W>
__device__ __forceinline__ int32_t tile_max(int32_t v) {
  if constexpr (W == 1) return v;
  return __reduce_max_sync(tile_mask<W>(), v);
}
template <int W>
__device__ __forceinline__ int32_t tile_add(int32_t v) {
  if constexpr (W == 1) return v;
  return __reduce_add_sync(tile_mask<W>(), v);
}

// intstr.GetScaledValueFromIntOrPercent in integers.  The reference evaluates
// ceil/floor(float64(v)*float64(total)/100); for |v*total| < 2^53 the float
// quotient is on the same side of every integer as the exact one, so integer
// floor/ceil division gives the identical result ("float vs int").
__device__ __forceinline__ int32_t scaled_value(int32_t val, bool is_percent, int32_t total,
                                                bool round_up) {
  if (!is_percent) return val;
  const int64_t num = (int64_t)val * (int64_t)total;
  int64_t q = num / 100;
  const int64_t r = num - q * 100;
  if (r != 0) {
    if (round_up && num > 0) q += 1;
    if (!round_up && num < 0) q -= 1;
  }
  return (int32_t)q;
}

// Programmatic dependent launch (PDL): the kernels of one sweep are launched
// with programmatic stream serialization, call pdl_launch_dependents() first
// thing (the next kernel may start its own prologue) and pdl_wait_prior()
// right before the first access to data the previous kernel produces.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace lwse
