// Placement round (sm_100a) — build-defined spec, see oracle/lwse_oracle_place.c
// and DESIGN.md "Placement (parity unpinned)".
//
// One cooperative, persistent kernel (one CTA per SM):
//   phase 0  the grid condenses the node table (16 B/node) and the occupancy
//            vector into one word per node (free slots | domain) plus the
//            per-domain free capacity, and pinned requests (leader already
//            scheduled) claim their domain with atomicMin on the 64-bit holder
//            key; one grid.sync;
//   stage    every CTA pulls the condensed node table and the domain capacities
//            into shared memory with TMA (cp.async.bulk + mbarrier);
//   phase 2  deferred-acceptance rounds: every unplaced / displaced request is
//            taken by one CTA, whose threads score every (request, node) pair
//            from shared memory, arg-max across the CTA, then atomicMin the
//            holder of the winning domain; one grid.sync per round.  Holder keys only ever
//            decrease, so the fixed point is unique and equals the sequential
//            "ascending key takes its best free domain" statement of the oracle.
#include <cooperative_groups.h>

#include <cstdlib>

#include "lwse_device.cuh"

namespace cg = cooperative_groups;

extern const uint32_t* g_last_place_counters;

namespace lwse {

struct PlaceArgs {
  const lwse_node_rec* nodes;
  const lwse_place_req* reqs;
  const uint32_t* occupancy;  // nullable
  lwse_place_out* out;
  unsigned long long* holder;  // n_namespaces x n_domains
  uint32_t* choice;            // per request: proposed node (or NONE)
  uint32_t* state;             // per request: 1 = unschedulable
  uint32_t* counters;          // [0..2] unsettled claims per round (rotating), [3] rounds, [4] unpinned requests
  uint32_t* g_compact;         // condensed node words (global copy, TMA source)
  uint32_t* g_dom_free;
  uint32_t* unpinned;          // indices of the unpinned requests (built in phase 0)
  unsigned long long* next_holder;  // the scratch half the NEXT call will use: reset here
  uint32_t* next_zero;         // its counters + domain capacities
  uint32_t next_zero_words;
  uint32_t n_nodes, n_domains, n_reqs, n_namespaces;
  uint32_t smem_nodes;  // 1: node words + domain capacities live in shared memory
  // gathered form (multi-GPU): `reqs` / `occupancy` point into part 0 of n_parts equally laid out
  // parts (one per rank, as all-gathered); request r lives in part r / reqs_per_part, and the
  // occupancy of a node is the sum over the parts
  uint32_t n_parts, reqs_per_part;
  uint64_t part_stride_bytes;
  uint32_t* soft_bar;  // non-null: [0] arrivals, [1] generation of the software grid barrier
};


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

__device__ __forceinline__ lwse_place_req load_req(const PlaceArgs& a, uint32_t r) {
  const lwse_place_req* p = a.reqs;
  if (a.n_parts > 1u) {
    const uint32_t part = r / a.reqs_per_part, j = r - part * a.reqs_per_part;
    p = reinterpret_cast<const lwse_place_req*>(reinterpret_cast<const uint8_t*>(a.reqs) +
                                                (uint64_t)part * a.part_stride_bytes) + j;
  } else {
    p += r;
  }
  const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  lwse_place_req q;
  q.priority = u64_of(lo.x, lo.y);
  q.group_key = u64_of(lo.z, lo.w);
  q.group = hi.x;
  q.ns = hi.y;
  q.size = (int32_t)hi.z;
  q.leader_node = hi.w;
  return q;
}

__device__ __forceinline__ uint32_t load_occupancy(const PlaceArgs& a, uint32_t n) {
  if (!a.occupancy) return 0u;
  uint32_t occ = 0;
  for (uint32_t p = 0; p < a.n_parts; p++)
    occ += __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(a.occupancy) +
                                                    (uint64_t)p * a.part_stride_bytes) + n);
  return occ;
}

// phase timestamps (ns, %globaltimer) of CTA 0 for tuning: counters[16 + 2k, 16 + 2k + 1]
__device__ __forceinline__ void stamp(const PlaceArgs& a, uint32_t k) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && k < 16u) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.counters[16 + 2 * k] = (uint32_t)t;
    a.counters[17 + 2 * k] = (uint32_t)(t >> 32);
  }
}

// Grid barrier.  Cooperative launches do not share the GPU with other grids, so the placement
// round could never overlap the sweep kernels of the same step; launched as an ordinary grid of
// at most one CTA per SM (every CTA becomes resident as soon as the other, finite, kernels
// drain) the round uses this arrive-and-spin barrier instead.
__device__ __forceinline__ void grid_barrier(const PlaceArgs& a, cg::grid_group& grid) {
  if (a.soft_bar == nullptr) {
    __threadfence();
    grid.sync();
    return;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(a.soft_bar + 1) : "memory");
    __threadfence();
    const uint32_t arrived = atomicAdd(a.soft_bar, 1u);
    if (arrived == gridDim.x - 1u) {
      a.soft_bar[0] = 0u;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.soft_bar + 1), "r"(gen + 1u) : "memory");
    } else {
      uint32_t now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(a.soft_bar + 1) : "memory");
      } while (now == gen);
    }
  }
  __syncthreads();
}

constexpr uint32_t kUnusable = 0xFFFFFFFFu;
// node word: min(free,15) << 28 | domain (28 bits); the exact free count only feeds dom_free

// A CTA's view of one of its unpinned requests, kept in shared memory across the rounds
// (write-through: choice / state are mirrored in global memory for the CTAs' requests that
// do not fit in the cache).
struct ReqCache {
  unsigned long long key;
  uint32_t r, ns, size, key_lo, key_hi, cur, dead, pad;
};
constexpr uint32_t kCacheQ = 64;  // requests per CTA held in shared memory
constexpr uint32_t kSmemHeader = 256 + kCacheQ * sizeof(ReqCache);

__device__ __forceinline__ lwse_place_out pinned_result(const PlaceArgs& a, const lwse_place_req& rq, uint32_t r) {
  lwse_place_out o;
  o.domain_id = LWSE_NONE;
  o.leader_node = rq.leader_node;
  o.flags = LWSE_PLACE_PINNED;
  o.score = 0;
  if (rq.ns >= a.n_namespaces) {
    o.flags |= LWSE_PLACE_UNSCHEDULABLE;
  } else if (rq.leader_node < a.n_nodes) {
    const uint4 nr = __ldg(reinterpret_cast<const uint4*>(a.nodes + rq.leader_node));
    const uint32_t d = nr.z;
    if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && d < a.n_domains) {
      o.domain_id = d;
      // pinned keys (bit 63 clear) are below every unpinned key: once all pinned claims are in,
      // the holder of a pinned domain never changes again
      const unsigned long long h = __ldcg(a.holder + (uint64_t)rq.ns * a.n_domains + d);
      o.flags |= h == place_key(rq, r, true) ? LWSE_PLACE_PLACED : LWSE_PLACE_CONFLICT;
    }
  }
  return o;
}

// kPlaceThreads threads per CTA, at least MINB CTAs' worth of registers per SM left to ptxas:
// the round shares each SM with the sweep kernels of the same tick, so what it leaves free
// (registers above all) decides how much of the sweep overlaps it.
template <uint32_t kPlaceThreads, int MINB>
__global__ void __launch_bounds__(kPlaceThreads, MINB) place_kernel(const PlaceArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem + 16);  // 16 warps
  uint32_t* s_best_n = reinterpret_cast<uint32_t*>(smem + 16 + 16 * 8);
  uint32_t* s_flag = reinterpret_cast<uint32_t*>(smem + 16 + 16 * 8 + 16 * 4);
  ReqCache* s_req = reinterpret_cast<ReqCache*>(smem + 256);
  uint32_t* s_words = reinterpret_cast<uint32_t*>(smem + kSmemHeader);
  const uint32_t n_pad = (a.n_nodes + 31u) & ~31u;
  const uint32_t nd4 = (a.n_domains + 3u) & ~3u;
  uint32_t* s_hi = s_words + n_pad + nd4;  // per-domain score of the current request, 0 = may not claim
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t gtid = blockIdx.x * blockDim.x + tid, gsize = gridDim.x * blockDim.x;
  stamp(a, 0);

  // ---------------- phase 0: node words, domain capacities, pinned claims ----------------
  // (this scratch half was left clean — holders ~0, counters and capacities 0 — by the previous call)
  // Every CTA condenses its slice of the node table: 16-byte node row + occupancy →
  // one word (free slots | domain); domain capacities accumulate with global atomics.
  for (uint32_t n = gtid; n < n_pad; n += gsize) {
    uint32_t word = kUnusable;
    if (n < a.n_nodes) {
      const uint4 nr = ldg_stream(reinterpret_cast<const uint4*>(a.nodes + n));
      const uint32_t d = nr.z, cap = nr.w & 0xFFFFu, nflags = nr.w >> 16;
      const uint32_t occ = load_occupancy(a, n);
      const bool usable = (nflags & LWSE_NODE_SCHEDULABLE) && (nflags & LWSE_NODE_HAS_TOPOLOGY) && d < a.n_domains;
      const uint32_t fr = usable && cap > occ ? cap - occ : 0u;
      if (usable) word = (min(fr, 15u) << 28) | d;
      if (usable && fr) atomicAdd(a.g_dom_free + d, fr);
    }
    a.g_compact[n] = word;
  }
  // leave the other scratch half clean for the next call (it is idle: calls are stream-ordered)
  {
    const uint64_t n_hold = (uint64_t)a.n_namespaces * a.n_domains;
    for (uint64_t i = gtid; i < n_hold; i += gsize) a.next_holder[i] = ~0ull;
    for (uint32_t i = gtid; i < a.next_zero_words; i += gsize) a.next_zero[i] = 0u;
  }
  for (uint32_t r = gtid; r < a.n_reqs; r += gsize) {
    const lwse_place_req rq = load_req(a, r);
    if (rq.leader_node == LWSE_NONE) {
      a.choice[r] = LWSE_NONE;
      a.state[r] = 0;
      a.unpinned[atomicAdd(a.counters + 4, 1u)] = r;  // order is irrelevant: the fixed point is unique
    } else if (rq.ns < a.n_namespaces && rq.leader_node < a.n_nodes) {
      const uint4 nr = __ldg(reinterpret_cast<const uint4*>(a.nodes + rq.leader_node));
      const uint32_t d = nr.z;
      if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && d < a.n_domains)
        atomicMin(a.holder + (uint64_t)rq.ns * a.n_domains + d, place_key(rq, r, true));
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stamp(a, 1);
  grid_barrier(a, grid);
  stamp(a, 2);

  const uint32_t n_unpinned = __ldcg(a.counters + 4);
  uint32_t* compact = a.g_compact;
  uint32_t* dom_free = a.g_dom_free;
  if (n_unpinned && a.smem_nodes && tid == 0) {
    // TMA-stage the condensed node table and the domain capacities into shared memory (every
    // (request, node) pair below is scored from on-chip memory); the copy flies while the pinned
    // results are written and the request cache is filled
    const uint32_t bytes_w = n_pad * 4u, bytes_d = nd4 * 4u;
    mbar_expect_tx(bar, bytes_w + bytes_d);
    tma_bulk_g2s(s_words, a.g_compact, bytes_w, bar);
    tma_bulk_g2s(s_words + n_pad, a.g_dom_free, bytes_d, bar);
  }
  // results of the pinned requests: final as of now
  for (uint32_t r = gtid; r < a.n_reqs; r += gsize) {
    const lwse_place_req rq = load_req(a, r);
    if (rq.leader_node != LWSE_NONE) a.out[r] = pinned_result(a, rq, r);
  }

  uint32_t round = 0;
  if (n_unpinned) {
    // this CTA's requests: k = blockIdx.x, blockIdx.x + gridDim.x, …; the first kCacheQ live in shared memory
    const uint32_t my_count = blockIdx.x < n_unpinned ? (n_unpinned - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
    if (tid < min(my_count, kCacheQ)) {
      const uint32_t r = __ldcg(a.unpinned + blockIdx.x + tid * gridDim.x);
      const lwse_place_req rq = load_req(a, r);
      ReqCache q;
      q.key = place_key(rq, r, false);
      q.r = r;
      q.ns = rq.ns;
      q.size = (uint32_t)rq.size;
      q.key_lo = (uint32_t)rq.group_key;
      q.key_hi = (uint32_t)(rq.group_key >> 32);
      q.cur = LWSE_NONE;
      q.dead = (rq.ns >= a.n_namespaces || rq.size < 1) ? 1u : 0u;
      q.pad = 0;
      if (q.dead) a.state[r] = 1;
      s_req[tid] = q;
    }
    if (a.smem_nodes) {
      mbar_wait(bar, 0);
      compact = s_words;
      dom_free = s_words + n_pad;
    }
    __syncthreads();
    stamp(a, 3);

    // ---------------- phase 2: deferred-acceptance rounds, one CTA per request ----------------
    for (;; round++) {
      // three rotating counters: the one for round k+1 is cleared during round k,
      // when no CTA can still be reading it (it was last read after round k-2)
      uint32_t* counter = a.counters + (round % 3u);
      if (blockIdx.x == 0 && tid == 0) a.counters[(round + 1u) % 3u] = 0;
      for (uint32_t j = 0; j < my_count; j++) {
        ReqCache q;
        if (j < kCacheQ) {
          q = s_req[j];
        } else {  // overflow: the request lives in global memory (state / choice are written by this CTA only)
          const uint32_t r = __ldcg(a.unpinned + blockIdx.x + j * gridDim.x);
          const lwse_place_req rq = load_req(a, r);
          q.key = place_key(rq, r, false);
          q.r = r;
          q.ns = rq.ns;
          q.size = (uint32_t)rq.size;
          q.key_lo = (uint32_t)rq.group_key;
          q.key_hi = (uint32_t)(rq.group_key >> 32);
          q.cur = __ldcg(a.choice + r);
          q.dead = (rq.ns >= a.n_namespaces || rq.size < 1) ? 1u : __ldcg(a.state + r);
          if (q.dead && tid == 0) a.state[r] = 1;
        }
        if (q.dead) continue;  // uniform
        const unsigned long long key = q.key;
        unsigned long long* hold = a.holder + (uint64_t)q.ns * a.n_domains;
        const uint32_t size = q.size;
        unsigned long long best = 0;
        uint32_t best_n = LWSE_NONE;
        if (a.smem_nodes) {
          // Pass 1 — domains.  Capacity, "not held by a higher-priority group" and the domain part
          // of the score are per-domain facts: compute them once per domain into shared memory
          // and find the winning domain score H.  The same pass notices whether this request
          // still holds the domain it proposed to earlier (then nothing is left to do).  Holders
          // change under our feet during a round (other CTAs' atomicMin), so every decision the
          // CTA branches on is reduced through shared memory — threads never branch on their own
          // reading of a holder.  (Per-node holder loads cost 8.8 us per request, per-node
          // hashing 5 us, a serial chain of request/state loads 2 us — measured with %globaltimer.)
          if (round == 0 && j == 0) stamp(a, 10);
          const uint32_t cur_dom = q.cur != LWSE_NONE ? (compact[q.cur] & 0x0FFFFFFFu) : LWSE_NONE;
          uint32_t my_hi = 0, my_holding = 0;
          for (uint32_t d = tid; d < a.n_domains; d += kPlaceThreads) {
            const unsigned long long h = __ldcg(hold + d);
            uint32_t hi = 0;
            if (dom_free[d] >= size && h >= key) hi = mix32(q.key_lo ^ (d * 0x9E3779B1u)) | 1u;
            if (d == cur_dom && h == key) my_holding = 1u;
            s_hi[d] = hi;  // 0 = this request may not claim d
            my_hi = max(my_hi, hi);
          }
          my_hi = __reduce_max_sync(0xFFFFFFFFu, my_hi);
          my_holding = __reduce_or_sync(0xFFFFFFFFu, my_holding);
          __syncthreads();  // the previous request's arg-max is done with s_best / s_best_n
          if (lane == 0) {
            s_best_n[warp] = my_hi;
            s_best[warp] = my_holding;
          }
          __syncthreads();
          uint32_t H = 0, holding = 0;
#pragma unroll
          for (int w = 0; w < (int)(kPlaceThreads / 32); w++) {
            H = max(H, s_best_n[w]);
            holding |= (uint32_t)s_best[w];
          }
          __syncthreads();  // s_best / s_best_n are reused by the arg-max below
          if (round == 0 && j == 0) stamp(a, 11);
          if (holding) continue;  // uniform: derived from shared memory
          // Pass 2 — nodes of the winning domain(s): every (request, node) pair is looked at, but
          // only nodes whose domain carries the winning score are hashed and ranked.
          if (H != 0u) {
            // four node words per 128-bit shared-memory load; the four domain-score lookups are
            // independent (padding words are kUnusable)
            const uint4* words4 = reinterpret_cast<const uint4*>(compact);
            for (uint32_t n4 = tid; n4 < (n_pad >> 2); n4 += kPlaceThreads) {
              const uint4 w4 = words4[n4];
              const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
              uint32_t hit[4];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const bool ok = w[k] != kUnusable && (w[k] >> 28) != 0u;
                hit[k] = ok ? s_hi[w[k] & 0x0FFFFFFFu] : 0u;
              }
#pragma unroll
              for (int k = 0; k < 4; k++) {
                if (hit[k] != H) continue;
                const uint32_t n = n4 * 4u + (uint32_t)k;
                const uint32_t lo = ((w[k] >> 28) << 28) | (mix32(q.key_hi ^ (n * 0x85EBCA77u)) >> 4);
                const unsigned long long sc = ((unsigned long long)H << 32) | lo;
                if (best_n == LWSE_NONE || sc > best) {  // n ascends per thread: ties keep the lower index
                  best = sc;
                  best_n = n;
                }
              }
            }
          }
        } else {
          // node words in global memory (table too large for shared memory): one thread decides
          // whether the request still holds its domain and broadcasts, then one fused pass
          __syncthreads();
          if (tid == 0)
            *s_flag = (q.cur != LWSE_NONE && __ldcg(hold + (compact[q.cur] & 0x0FFFFFFFu)) == key) ? 0u : 1u;
          __syncthreads();
          if (!*s_flag) continue;
          for (uint32_t n = tid; n < a.n_nodes; n += kPlaceThreads) {
            const uint32_t w = compact[n];
            if (w == kUnusable || (w >> 28) == 0u) continue;
            const uint32_t d = w & 0x0FFFFFFFu;
            if (dom_free[d] < size) continue;
            if (__ldcg(hold + d) < key) continue;  // held by a higher-priority group (monotone: never frees)
            const uint32_t hi = mix32(q.key_lo ^ (d * 0x9E3779B1u)) | 1u;
            const uint32_t lo = ((w >> 28) << 28) | (mix32(q.key_hi ^ (n * 0x85EBCA77u)) >> 4);
            const unsigned long long sc = ((unsigned long long)hi << 32) | lo;
            if (best_n == LWSE_NONE || sc > best) {
              best = sc;
              best_n = n;
            }
          }
        }
        if (round == 0 && j == 0) stamp(a, 12);
        auto better = [](unsigned long long os, uint32_t on, unsigned long long s, uint32_t n) {
          return on != LWSE_NONE && (n == LWSE_NONE || os > s || (os == s && on < n));
        };
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const unsigned long long os = __shfl_xor_sync(0xFFFFFFFFu, best, off);
          const uint32_t on = __shfl_xor_sync(0xFFFFFFFFu, best_n, off);
          if (better(os, on, best, best_n)) {
            best = os;
            best_n = on;
          }
        }
        if (lane == 0) {
          s_best[warp] = best;
          s_best_n[warp] = best_n;
        }
        __syncthreads();
        if (warp == 0) {
          best = lane < kPlaceThreads / 32 ? s_best[lane] : 0ull;
          best_n = lane < kPlaceThreads / 32 ? s_best_n[lane] : LWSE_NONE;
#pragma unroll
          for (int off = 8; off > 0; off >>= 1) {
            const unsigned long long os = __shfl_xor_sync(0xFFFFFFFFu, best, off);
            const uint32_t on = __shfl_xor_sync(0xFFFFFFFFu, best_n, off);
            if (better(os, on, best, best_n)) {
              best = os;
              best_n = on;
            }
          }
          if (lane == 0) {
            if (best_n == LWSE_NONE) {
              a.state[q.r] = 1;  // nothing feasible now, and the feasible set only shrinks
              a.choice[q.r] = LWSE_NONE;
              if (j < kCacheQ) {
                s_req[j].dead = 1;
                s_req[j].cur = LWSE_NONE;
              }
            } else {
              const uint32_t d = compact[best_n] & 0x0FFFFFFFu;
              // A claim on an empty domain settles at once.  Any other outcome leaves somebody
              // without a domain — the previous holder (old > key) or this request (old < key) —
              // who proposes again next round: count it, so that the round after the last
              // displacement is never run just to find nothing to do.
              const unsigned long long old = atomicMin(hold + d, key);
              a.choice[q.r] = best_n;
              if (j < kCacheQ) s_req[j].cur = best_n;
              a.out[q.r].score = (uint32_t)(best >> 32);
              if (old != ~0ull) atomicAdd(counter, 1u);
            }
          }
        }
        __syncthreads();  // s_best and the cache entry are settled before the next request
        if (round == 0 && j == 0) stamp(a, 13);
      }
      stamp(a, 4 + 2 * round);
      grid_barrier(a, grid);
      stamp(a, 5 + 2 * round);
      const uint32_t unsettled = __ldcg(counter);
      if (unsettled == 0u || round > a.n_reqs + 2u) break;
    }

    // results of this CTA's unpinned requests (fixed point: every live request holds its choice)
    for (uint32_t j = tid; j < my_count; j += kPlaceThreads) {
      uint32_t r, cur, dead;
      if (j < kCacheQ) {
        r = s_req[j].r;
        cur = s_req[j].cur;
        dead = s_req[j].dead;
      } else {
        r = __ldcg(a.unpinned + blockIdx.x + j * gridDim.x);
        cur = __ldcg(a.choice + r);
        dead = __ldcg(a.state + r);
      }
      lwse_place_out o;
      if (dead || cur == LWSE_NONE) {
        o.domain_id = LWSE_NONE;
        o.leader_node = LWSE_NONE;
        o.flags = LWSE_PLACE_UNSCHEDULABLE;
        o.score = 0;
      } else {
        o.domain_id = compact[cur] & 0x0FFFFFFFu;
        o.leader_node = cur;
        o.flags = LWSE_PLACE_PLACED;
        o.score = a.out[r].score;
      }
      a.out[r] = o;
    }
  }
  if (blockIdx.x == 0 && tid == 0) a.counters[3] = n_unpinned ? round + 1u : 0u;
  stamp(a, 15);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// scratch: two identical halves used alternately by successive calls; each half is
// [holder | counters(256 B) | g_dom_free | choice | state | unpinned | g_compact].  A call resets the
// holders / counters / capacities of the *other* half, so no memset sits on the critical path
// (the engine zero-fills the whole scratch once, when it allocates it).
static size_t place_half_bytes(uint32_t n_nodes, uint32_t n_domains, uint32_t n_reqs, uint32_t n_namespaces) {
  return align_up((size_t)n_namespaces * n_domains * 8, 256) + 256 + align_up((size_t)n_domains * 4 + 16, 256) +
         align_up((size_t)n_reqs * 4, 256) * 3 + align_up((size_t)n_nodes * 4 + 128, 256);
}
size_t place_scratch_bytes(uint32_t n_nodes, uint32_t n_domains, uint32_t n_reqs, uint32_t n_namespaces) {
  return 2 * place_half_bytes(n_nodes, n_domains, n_reqs, n_namespaces) + 1024;
}

// `fresh`: the scratch was (re)allocated or its geometry changed → initialise both halves first.
int launch_place(const lwse_node_rec* d_nodes, uint32_t n_nodes, uint32_t n_domains,
                 const lwse_place_req* d_reqs, uint32_t n_reqs, const uint32_t* d_occupancy,
                 uint32_t n_namespaces, lwse_place_out* d_out, void* d_scratch, size_t scratch_bytes,
                 uint32_t* h_rounds, int sm_count, cudaStream_t s, int* cuda_err, uint32_t call_index,
                 bool fresh, uint32_t n_parts, uint32_t reqs_per_part, uint64_t part_stride_bytes) {
  *cuda_err = 0;
  if (n_reqs > 0xFFFFFFu || n_domains >= (1u << 28) ||
      scratch_bytes < place_scratch_bytes(n_nodes, n_domains, n_reqs, n_namespaces)) {
    *cuda_err = (int)cudaErrorInvalidValue;
    return -1;
  }
  const size_t half = place_half_bytes(n_nodes, n_domains, n_reqs, n_namespaces);
  const size_t holder_bytes = align_up((size_t)n_namespaces * n_domains * 8, 256);
  const size_t zero_bytes = 256 + align_up((size_t)n_domains * 4 + 16, 256);
  uint8_t* base = static_cast<uint8_t*>(d_scratch);
  uint8_t* p = base + (call_index & 1u) * half;
  uint8_t* q = base + ((call_index + 1u) & 1u) * half;
  cudaError_t e = cudaSuccess;
  if (fresh) {
    for (int h = 0; h < 2 && e == cudaSuccess; h++) {
      e = cudaMemsetAsync(base + h * half, 0xFF, holder_bytes, s);
      if (e == cudaSuccess) e = cudaMemsetAsync(base + h * half + holder_bytes, 0, zero_bytes, s);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(base + 2 * half, 0, 1024, s);  // software-barrier words
    if (e != cudaSuccess) {
      *cuda_err = (int)e;
      return -1;
    }
  }
  PlaceArgs a{};
  a.nodes = d_nodes;
  a.reqs = d_reqs;
  a.occupancy = d_occupancy;
  a.out = d_out;
  a.holder = reinterpret_cast<unsigned long long*>(p);
  a.next_holder = reinterpret_cast<unsigned long long*>(q);
  p += holder_bytes;
  q += holder_bytes;
  a.counters = reinterpret_cast<uint32_t*>(p);
  a.g_dom_free = reinterpret_cast<uint32_t*>(p + 256);
  a.next_zero = reinterpret_cast<uint32_t*>(q);
  a.next_zero_words = (uint32_t)(zero_bytes / 4);
  p += zero_bytes;
  a.choice = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_reqs * 4, 256);
  a.state = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_reqs * 4, 256);
  a.unpinned = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n_reqs * 4, 256);
  a.g_compact = reinterpret_cast<uint32_t*>(p);
  a.n_nodes = n_nodes;
  a.n_domains = n_domains;
  a.n_reqs = n_reqs;
  a.n_namespaces = n_namespaces;
  a.n_parts = n_parts ? n_parts : 1u;
  a.reqs_per_part = reqs_per_part ? reqs_per_part : n_reqs;
  a.part_stride_bytes = part_stride_bytes;

  const size_t words_bytes = ((size_t)((n_nodes + 31u) & ~31u) + 2 * (size_t)((n_domains + 3u) & ~3u)) * 4 + 16;
  size_t smem = kSmemHeader + words_bytes;
  a.smem_nodes = smem <= 227u * 1024u ? 1u : 0u;
  if (!a.smem_nodes) smem = kSmemHeader;
  static const int variant = [] {
    const char* v = getenv("LWSE_PLACE_VARIANT");
    return v ? atoi(v) : 0;
  }();
  // 256 threads x 64 registers: a quarter of the register file (the 512 x 92 first version held
  // three quarters and the sweep's group pass of the same tick could not start next to it:
  // 40.9 us per tick against 29.5 us)
  void (*kernel)(PlaceArgs) = place_kernel<256, 2>;
  unsigned threads = 256;
  if (variant == 1) kernel = place_kernel<512, 1>, threads = 512;
  if (variant == 2) kernel = place_kernel<512, 2>, threads = 512;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      *cuda_err = (int)e;
      return -1;
    }
    smem_set = smem;
  }
  ::g_last_place_counters = a.counters;
  void* params[] = {&a};
  // one CTA per SM at most; a grid barrier costs more the more CTAs take part, and a round
  // needs no more CTAs than there are requests
  static const int env_ctas = [] {
    const char* v = getenv("LWSE_PLACE_CTAS");
    return v ? atoi(v) : 0;
  }();
  unsigned ctas = (unsigned)sm_count;
  if (n_reqs < ctas) ctas = n_reqs < 16u ? 16u : n_reqs;
  if (env_ctas > 0 && (unsigned)env_ctas < ctas) ctas = (unsigned)env_ctas;
  static const bool soft = [] {
    const char* v = getenv("LWSE_PLACE_SOFT_BARRIER");
    return v && atoi(v) != 0;
  }();
  if (soft) {
    a.soft_bar = reinterpret_cast<uint32_t*>(base + 2 * half);  // zero-filled when the scratch is allocated
    kernel<<<dim3(ctas), dim3(threads), smem, s>>>(a);
    e = cudaGetLastError();
  } else {
    e = cudaLaunchCooperativeKernel((const void*)kernel, dim3(ctas), dim3(threads), params, smem, s);
  }
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
