// Placement round (sm_100a) — build-defined spec, see oracle/lwse_oracle_place.c
// and DESIGN.md "Placement (parity unpinned)".
//
// One cooperative, persistent kernel (one CTA per SM):
//   phase 0  every CTA stages the node table through shared memory with TMA
//            (cp.async.bulk + mbarrier, 64 KB chunks) and condenses it, with the
//            occupancy vector, into a 4-byte/node word (free slots | domain) and
//            the per-domain free capacity — both kept in shared memory for the
//            whole kernel;
//   phase 1  pinned requests (leader already scheduled) claim their domain with
//            atomicMin on the 64-bit holder key;
//   phase 2  deferred-acceptance rounds: every unplaced / displaced request is
//            taken by one warp, whose lanes score every (request, node) pair
//            against the shared-memory node words, arg-max across the warp, then
//            atomicMin the holder of the winning domain.  Holder keys only ever
//            decrease, so the fixed point is unique and equals the sequential
//            "ascending key takes its best free domain" statement of the oracle.
#include <cooperative_groups.h>

#include "lwse_device.cuh"

namespace cg = cooperative_groups;

namespace lwse {

struct PlaceArgs {
  const lwse_node_rec* nodes;
  const lwse_place_req* reqs;
  const uint32_t* occupancy;  // nullable
  lwse_place_out* out;
  unsigned long long* holder;  // n_namespaces x n_domains
  uint32_t* choice;            // per request: proposed node (or NONE)
  uint32_t* state;             // per request: 1 = unschedulable
  uint32_t* counters;          // [0..2] proposals per round (rotating), [3] rounds
  uint32_t* g_compact;         // fallback when the node words do not fit in shared memory
  uint32_t* g_dom_free;
  uint32_t n_nodes, n_domains, n_reqs, n_namespaces;
  uint32_t smem_nodes;  // 1: node words + domain capacities live in shared memory
};

constexpr uint32_t kStageRows = 4096;  // 64 KB TMA stage
constexpr uint32_t kPlaceThreads = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global → shared, completion reported to the mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ unsigned long long place_key(const lwse_place_req& r, uint32_t index, bool pinned) {
  return ((unsigned long long)(pinned ? 0 : 1) << 63) | (((r.priority >> 25) & 0x7FFFFFFFFFull) << 24) |
         (unsigned long long)(index & 0xFFFFFFu);
}

constexpr uint32_t kUnusable = 0xFFFFFFFFu;
// node word: min(free,15) << 28 | domain (28 bits); the exact free count only feeds dom_free

__global__ void __launch_bounds__(kPlaceThreads, 1) place_kernel(const PlaceArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint4* stage = reinterpret_cast<uint4*>(smem + 128);
  uint32_t* s_words = reinterpret_cast<uint32_t*>(smem + 128 + kStageRows * sizeof(lwse_node_rec));
  uint32_t* compact = a.smem_nodes ? s_words : a.g_compact;
  uint32_t* dom_free = a.smem_nodes ? s_words + ((a.n_nodes + 31u) & ~31u) : a.g_dom_free;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const bool builder = a.smem_nodes || blockIdx.x == 0;  // who condenses the node table

  // ---------------- phase 0: holders, node words, domain capacities ----------------
  const uint64_t n_hold = (uint64_t)a.n_namespaces * a.n_domains;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + tid; i < n_hold; i += (uint64_t)gridDim.x * blockDim.x)
    a.holder[i] = ~0ull;
  if (blockIdx.x == 0 && tid < 4) a.counters[tid] = 0;
  if (builder) {
    for (uint32_t d = tid; d < a.n_domains; d += blockDim.x) dom_free[d] = 0;
    if (tid == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (uint32_t base = 0; base < a.n_nodes; base += kStageRows) {
      const uint32_t rows = min(kStageRows, a.n_nodes - base);
      if (tid == 0) {
        mbar_expect_tx(bar, rows * (uint32_t)sizeof(lwse_node_rec));
        tma_bulk_g2s(stage, a.nodes + base, rows * (uint32_t)sizeof(lwse_node_rec), bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1u;
      for (uint32_t i = tid; i < rows; i += blockDim.x) {
        const uint4 nr = stage[i];
        const uint32_t n = base + i, d = nr.z, cap = nr.w & 0xFFFFu, nflags = nr.w >> 16;
        const uint32_t occ = a.occupancy ? __ldg(a.occupancy + n) : 0u;
        const bool usable = (nflags & LWSE_NODE_SCHEDULABLE) && (nflags & LWSE_NODE_HAS_TOPOLOGY) &&
                            d < a.n_domains;
        const uint32_t fr = usable && cap > occ ? cap - occ : 0u;
        compact[n] = usable ? ((min(fr, 15u) << 28) | d) : kUnusable;
        if (usable && fr) atomicAdd(dom_free + d, fr);
      }
      __syncthreads();  // everyone is done with the stage before the next TMA overwrites it
    }
  }
  __threadfence();
  grid.sync();

  // ---------------- phase 1: pinned claims ----------------
  for (uint32_t r = blockIdx.x * blockDim.x + tid; r < a.n_reqs; r += gridDim.x * blockDim.x) {
    const lwse_place_req rq = a.reqs[r];
    a.choice[r] = LWSE_NONE;
    a.state[r] = 0;
    if (rq.leader_node != LWSE_NONE && rq.ns < a.n_namespaces && rq.leader_node < a.n_nodes) {
      const uint4 nr = __ldg(reinterpret_cast<const uint4*>(a.nodes + rq.leader_node));
      const uint32_t d = nr.z;
      if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && d < a.n_domains)
        atomicMin(a.holder + (uint64_t)rq.ns * a.n_domains + d, place_key(rq, r, true));
    }
  }
  __threadfence();
  grid.sync();

  // ---------------- phase 2: deferred-acceptance rounds for the unpinned ----------------
  const uint32_t warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const uint32_t warp_id = blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
  uint32_t round = 0;
  for (;; round++) {
    // three rotating counters: the one for round k+1 is cleared during round k,
    // when no CTA can still be reading it (it was last read after round k-2)
    uint32_t* counter = a.counters + (round % 3u);
    if (blockIdx.x == 0 && tid == 0) a.counters[(round + 1u) % 3u] = 0;
    for (uint32_t r = warp_id; r < a.n_reqs; r += warps_per_grid) {
      const lwse_place_req rq = a.reqs[r];
      if (rq.leader_node != LWSE_NONE) continue;  // pinned
      if (rq.ns >= a.n_namespaces || rq.size < 1) {
        if (lane == 0) a.state[r] = 1;
        continue;
      }
      if (a.state[r]) continue;
      const unsigned long long key = place_key(rq, r, false);
      unsigned long long* hold = a.holder + (uint64_t)rq.ns * a.n_domains;
      const uint32_t cur = a.choice[r];
      if (cur != LWSE_NONE) {
        const uint32_t cd = compact[cur] & 0x0FFFFFFFu;
        if (__ldcg(hold + cd) == key) continue;  // still holding its domain
      }
      // score every (request, node) pair
      const uint32_t key_lo = (uint32_t)rq.group_key, key_hi = (uint32_t)(rq.group_key >> 32);
      const uint32_t size = (uint32_t)rq.size;
      unsigned long long best = 0;
      uint32_t best_n = LWSE_NONE;
      for (uint32_t n = lane; n < a.n_nodes; n += 32u) {
        const uint32_t w = compact[n];
        if (w == kUnusable || (w >> 28) == 0u) continue;
        const uint32_t d = w & 0x0FFFFFFFu;
        const uint32_t df = dom_free[d];
        if (df < size) continue;
        if (__ldcg(hold + d) < key) continue;  // held by a higher-priority group (monotone: never frees)
        const uint32_t slack = (df - size) / size;
        const uint32_t bucket = slack > 7u ? 7u : slack;
        const uint32_t hi = ((7u - bucket) << 29) | (mix32(key_lo ^ (d * 0x9E3779B1u)) >> 3);
        const uint32_t lo = ((w >> 28) << 28) | (mix32(key_hi ^ (n * 0x85EBCA77u)) >> 4);
        const unsigned long long s = ((unsigned long long)hi << 32) | lo;
        if (best_n == LWSE_NONE || s > best) {  // n ascends per lane: ties keep the lower index
          best = s;
          best_n = n;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long os = __shfl_xor_sync(0xFFFFFFFFu, best, off);
        const uint32_t on = __shfl_xor_sync(0xFFFFFFFFu, best_n, off);
        const bool take = on != LWSE_NONE && (best_n == LWSE_NONE || os > best || (os == best && on < best_n));
        if (take) {
          best = os;
          best_n = on;
        }
      }
      if (lane == 0) {
        if (best_n == LWSE_NONE) {
          a.state[r] = 1;  // nothing feasible now, and the feasible set only shrinks
          a.choice[r] = LWSE_NONE;
        } else {
          const uint32_t d = compact[best_n] & 0x0FFFFFFFu;
          atomicMin(hold + d, key);
          a.choice[r] = best_n;
          a.out[r].score = (uint32_t)(best >> 32);
          atomicAdd(counter, 1u);
        }
      }
    }
    __threadfence();
    grid.sync();
    const uint32_t proposals = __ldcg(counter);
    if (proposals == 0u || round > a.n_reqs + 2u) break;
  }
  if (blockIdx.x == 0 && tid == 0) a.counters[3] = round + 1u;

  // ---------------- results ----------------
  for (uint32_t r = blockIdx.x * blockDim.x + tid; r < a.n_reqs; r += gridDim.x * blockDim.x) {
    const lwse_place_req rq = a.reqs[r];
    lwse_place_out o;
    o.domain_id = LWSE_NONE;
    o.leader_node = LWSE_NONE;
    o.flags = 0;
    o.score = 0;
    if (rq.leader_node != LWSE_NONE) {
      o.flags = LWSE_PLACE_PINNED;
      o.leader_node = rq.leader_node;
      if (rq.ns >= a.n_namespaces) {
        o.flags |= LWSE_PLACE_UNSCHEDULABLE;
      } else if (rq.leader_node < a.n_nodes) {
        const uint4 nr = __ldg(reinterpret_cast<const uint4*>(a.nodes + rq.leader_node));
        const uint32_t d = nr.z;
        if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && d < a.n_domains) {
          o.domain_id = d;
          const unsigned long long h = __ldcg(a.holder + (uint64_t)rq.ns * a.n_domains + d);
          o.flags |= h == place_key(rq, r, true) ? LWSE_PLACE_PLACED : LWSE_PLACE_CONFLICT;
        }
      }
    } else {
      const uint32_t cur = a.choice[r];
      if (a.state[r] || cur == LWSE_NONE) {
        o.flags = LWSE_PLACE_UNSCHEDULABLE;
      } else {
        o.domain_id = compact[cur] & 0x0FFFFFFFu;
        o.leader_node = cur;
        o.flags = LWSE_PLACE_PLACED;
        o.score = a.out[r].score;
      }
    }
    a.out[r] = o;
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t place_scratch_bytes(uint32_t n_nodes, uint32_t n_domains, uint32_t n_reqs, uint32_t n_namespaces) {
  return align_up((size_t)n_namespaces * n_domains * 8, 256) + align_up((size_t)n_reqs * 4, 256) * 2 + 256 +
         align_up((size_t)n_nodes * 4, 256) + align_up((size_t)n_domains * 4, 256) + 1024;
}

int launch_place(const lwse_node_rec* d_nodes, uint32_t n_nodes, uint32_t n_domains,
                 const lwse_place_req* d_reqs, uint32_t n_reqs, const uint32_t* d_occupancy,
                 uint32_t n_namespaces, lwse_place_out* d_out, void* d_scratch, size_t scratch_bytes,
                 uint32_t* h_rounds, int sm_count, cudaStream_t s, int* cuda_err) {
  *cuda_err = 0;
  if (n_reqs > 0xFFFFFFu) {
    *cuda_err = (int)cudaErrorInvalidValue;
    return -1;
  }
  if (scratch_bytes < place_scratch_bytes(n_nodes, n_domains, n_reqs, n_namespaces)) {
    *cuda_err = (int)cudaErrorInvalidValue;
    return -1;
  }
  uint8_t* p = static_cast<uint8_t*>(d_scratch);
  PlaceArgs a{};
  a.nodes = d_nodes;
  a.reqs = d_reqs;
  a.occupancy = d_occupancy;
  a.out = d_out;
  a.holder = reinterpret_cast<unsigned long long*>(p);
  p += align_up((size_t)n_namespaces * n_domains * 8, 256);
  a.choice = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_reqs * 4, 256);
  a.state = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_reqs * 4, 256);
  a.counters = reinterpret_cast<uint32_t*>(p);
  p += 256;
  a.g_compact = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_nodes * 4, 256);
  a.g_dom_free = reinterpret_cast<uint32_t*>(p);
  a.n_nodes = n_nodes;
  a.n_domains = n_domains;
  a.n_reqs = n_reqs;
  a.n_namespaces = n_namespaces;

  const size_t stage_bytes = 128 + (size_t)kStageRows * sizeof(lwse_node_rec);
  const size_t words_bytes = ((size_t)((n_nodes + 31u) & ~31u) + n_domains) * 4;
  size_t smem = stage_bytes + words_bytes;
  a.smem_nodes = smem <= 227u * 1024u ? 1u : 0u;
  if (!a.smem_nodes) smem = stage_bytes;
  cudaError_t e = cudaFuncSetAttribute(place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    *cuda_err = (int)e;
    return -1;
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_kernel, (int)kPlaceThreads, smem);
  if (e != cudaSuccess || per_sm < 1) {
    *cuda_err = (int)(e != cudaSuccess ? e : cudaErrorLaunchOutOfResources);
    return -1;
  }
  void* params[] = {&a};
  e = cudaLaunchCooperativeKernel((const void*)place_kernel, dim3((unsigned)sm_count), dim3(kPlaceThreads), params,
                                  smem, s);
  if (e != cudaSuccess) {
    *cuda_err = (int)e;
    return -1;
  }
  if (h_rounds) {
    e = cudaMemcpyAsync(h_rounds, a.counters + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      *cuda_err = (int)e;
      return -1;
    }
  }
  return 1;
}

}  // namespace lwse
