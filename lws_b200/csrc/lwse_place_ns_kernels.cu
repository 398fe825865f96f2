// Placement round, namespace-parallel form (sm_100a) — same build-defined spec as
// lwse_place_kernels.cu (oracle/lwse_oracle_place.c), for request tables GROUPED BY NAMESPACE.
//
// Exclusivity is per namespace and capacity is a snapshot that no claim consumes, so the
// namespaces are independent sub-problems.  That turns the round's hot state — the holder table
// of one namespace (8 B x domains), the per-domain free capacity and the condensed node-topology
// table (one word per usable node: free slots | domain, in domain order) — into a few tens of KB
// that fit the shared memory of ONE CTA:
//
//   place_condense_kernel   node rows + occupancy -> sorted node words + per-domain capacity
//                           (a warp per domain over its run of the static domain-sorted index:
//                           no atomics), and the namespace index of the request table
//                           (ns_first[v] = first request of namespace v; order violations counted).
//   place_ns_kernel         one CTA per namespace.  The capacities and the domain index are staged
//                           with cp.async.bulk + mbarrier (issued by one thread while the others
//                           clear the holder table and take the pinned claims), the holder table
//                           lives in shared memory, claims are 64-bit shared-memory atomicMin,
//                           rounds are separated by __syncthreads() — no grid or cluster barrier,
//                           no L2 round trip per domain.  Deferred acceptance runs to ITS OWN
//                           fixed point per namespace.
//     two-level form        512 threads.  After the pinned claims the CTA compacts the CANDIDATE
//                           domains — no pinned holder (a pinned key beats every unpinned key) and
//                           capacity left — into a list {domain, free}; the unpinned holder table
//                           is indexed by list position, so a (request, round) search is two
//                           coalesced shared-memory streams over the candidates (a fourth of the
//                           domains on C3) and one read of the chosen domain's (free | node) run.
//                           16 KB + 36 B per domain of shared memory: no staging of the node
//                           table, and the sweep kernels keep the SMs' shared memory.
//     scan form             256 threads; the node-word table and the node order are staged whole
//                           with TMA (cp.async.bulk.tensor.2d through a CUtensorMap + cp.async.bulk)
//                           and every (request x node) pair is scored from shared memory:
//                           feasibility + both rendezvous hashes per pair, lexicographic arg-max
//                           (domain score, lower domain, node score, lower node) — identical rows.
//   Inside a tick both forms also append the result rows that differ from the previous tick's to
//   the device change list (PlaceNsArgs::prev …) — no separate diff kernel.
#include <cuda.h>

#include <cstdlib>

#include "lwse_device.cuh"

namespace lwse {

struct PlaceNsArgs {
  const lwse_node_rec* nodes;
  const lwse_place_req* reqs;  // grouped by namespace (ns non-decreasing)
  const uint32_t* occupancy;   // nullable; n_parts blocks of n_nodes counters, part_stride_bytes apart
  lwse_place_out* out;
  const uint32_t* dom_first;   // static index (lwse_upload_nodes)
  const uint32_t* node_order;
  uint32_t* g_words;     // condensed node words in domain-sorted position, padded with zeros to whole rows of 256
  uint32_t* g_nodes2;    // the same positions: free slots (4 bits) | node index — the two-level form's level 2
  uint32_t* g_dom_free;  // n_domains (+ padding)
  uint32_t* ns_first;    // n_namespaces + 1
  uint32_t* unpinned;    // n_reqs: namespace v's live unpinned requests at [ns_first[v], …)
  uint32_t* counters;    // this call's block: [0] max rounds [1] live unpinned [2] order violations [3] (request, round) scans
  uint32_t* next_counters;  // the other block, cleared for the next call
  uint32_t n_nodes, n_usable, n_domains, n_reqs, n_namespaces;
  uint32_t n_parts;
  uint64_t part_stride_bytes;
  uint32_t word_rows;   // rows of 256 words of g_words
  uint32_t scan;        // 1 = brute-force (request x node) form
  uint32_t stage_words; // the word table fits shared memory and is staged
  uint32_t use_tensor_map;
  // a tick's change list (all null / 0 = off): rows that differ from `prev` are appended, `prev` is updated
  lwse_place_out* prev;
  uint32_t* chg_rows;
  lwse_place_out* chg_outs;
  uint32_t* chg_count;
  uint32_t chg_capacity;
  // multi-rank ticks: `occupancy` is the base of the exchange buffers; the blocks to sum are those of
  // step (lagged ? step - 1 : step) where step = *xch_step (the push kernel right before stored it)
  const unsigned long long* xch_step;
  uint64_t xch_half_bytes;
  uint32_t xch_lagged;
};

__device__ __forceinline__ uint32_t mix32n(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ unsigned long long ns_place_key(unsigned long long priority, uint32_t index, bool pinned) {
  return ((unsigned long long)(pinned ? 0 : 1) << 63) | (((priority >> 25) & 0x7FFFFFFFFFull) << 24) |
         (unsigned long long)(index & 0xFFFFFFu);
}

__device__ __forceinline__ uint32_t occupancy_sum(const PlaceNsArgs& a, const uint32_t* base, uint32_t n) {
  if (!base) return 0u;
  uint32_t occ = 0;
  for (uint32_t p = 0; p < a.n_parts; p++)
    occ += ldg_keep_u32(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(base) +
                                                           (uint64_t)p * a.part_stride_bytes) + n);
  return occ;
}

__device__ __forceinline__ void store_row(lwse_place_out* p, uint32_t d, uint32_t n, uint32_t flags, uint32_t score) {
  *reinterpret_cast<uint4*>(p) = make_uint4(d, n, flags, score);
}

// --------------------------------------------------------------------------
// condense: node words, domain capacities, namespace index
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) place_condense_kernel(const PlaceNsArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  pdl_launch_dependents();  // the namespace kernel may load its requests meanwhile; it waits before the tables below
  pdl_wait_prior();         // the previous call's namespace kernel may still read the tables written here
  const uint32_t* occ_base = a.occupancy;
  if (a.xch_step != nullptr && occ_base != nullptr) {
    const unsigned long long step = *reinterpret_cast<const volatile unsigned long long*>(a.xch_step);
    const unsigned long long read_step = (a.xch_lagged && step > 1ull) ? step - 1ull : step;
    occ_base = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(occ_base) + (read_step % 3ull) * a.xch_half_bytes);
  }
  // A. a warp per domain: its nodes are a run of the sorted index
  for (uint32_t d = gwarp; d < a.n_domains; d += n_warps) {
    const uint32_t first = ldg_keep_u32(a.dom_first + d), last = ldg_keep_u32(a.dom_first + d + 1u);
    uint32_t sum = 0;
    for (uint32_t i = first + lane; i < last; i += 32u) {
      const uint32_t n = ldg_keep_u32(a.node_order + i);
      const uint4 nr = ldg_keep(reinterpret_cast<const uint4*>(a.nodes + n));
      const uint32_t cap = nr.w & 0xFFFFu, occ = occupancy_sum(a, occ_base, n);
      const uint32_t fr = cap > occ ? cap - occ : 0u;
      a.g_words[i] = (min(fr, 15u) << 28) | d;
      a.g_nodes2[i] = (min(fr, 15u) << 28) | n;
      sum += fr;
    }
    sum = __reduce_add_sync(0xFFFFFFFFu, sum);
    if (lane == 0) a.g_dom_free[d] = sum;
  }
  // B. the namespace index of the (promised) grouped request table
  if (a.n_reqs == 0u) {
    for (uint32_t v = gtid; v <= a.n_namespaces; v += gsize) a.ns_first[v] = 0u;
  }
  for (uint32_t i = gtid; i < a.n_reqs; i += gsize) {
    const uint4 hi = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + i) + 1);
    const uint32_t b = hi.y;
    uint32_t lo_v;  // namespaces (lo_v, b] start at request i; for i == 0 also namespace lo_v itself
    if (i == 0u) {
      for (uint32_t v = 0; v <= min(b, a.n_namespaces); v++) a.ns_first[v] = 0u;
    } else {
      const uint32_t prev = ldg_keep_u32(reinterpret_cast<const uint32_t*>(a.reqs + i - 1) + 5);  // .ns
      if (prev > b) atomicAdd(a.counters + 2, 1u);  // not grouped: the caller broke its promise
      lo_v = min(prev, a.n_namespaces);
      for (uint32_t v = lo_v + 1u; v <= min(b, a.n_namespaces); v++) a.ns_first[v] = i;
    }
    if (i == a.n_reqs - 1u)
      for (uint32_t v = min(b, a.n_namespaces) + 1u; v <= a.n_namespaces; v++) a.ns_first[v] = a.n_reqs;
    if (b >= a.n_namespaces) {  // no such namespace: never takes part in a round
      const bool pinned = hi.w != LWSE_NONE;
      const uint4 row = make_uint4(LWSE_NONE, pinned ? hi.w : LWSE_NONE, (pinned ? LWSE_PLACE_PINNED : 0u) | LWSE_PLACE_UNSCHEDULABLE, 0u);
      *reinterpret_cast<uint4*>(a.out + i) = row;
      if (a.prev != nullptr) {
        const uint4 o = __ldcg(reinterpret_cast<const uint4*>(a.prev + i));
        if (row.x != o.x || row.y != o.y || row.z != o.z || row.w != o.w) {
          const uint32_t k = atomicAdd(a.chg_count, 1u);
          if (k < a.chg_capacity) {
            a.chg_rows[k] = i;
            *reinterpret_cast<uint4*>(a.chg_outs + k) = row;
          }
          *reinterpret_cast<uint4*>(a.prev + i) = row;
        }
      }
    }
  }
  if (gtid < 8u && a.next_counters != nullptr) a.next_counters[gtid] = 0u;
}

// --------------------------------------------------------------------------
// one CTA per namespace
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr uint32_t kNsScanThreads = 256;  // scan form: two CTAs (2 x ~100 KB of staged tables) share an SM
constexpr uint32_t kNsThreads = 512;      // two-level form: 16 warps, a warp per unpinned request and round
constexpr uint32_t kNsKeep = 2;           // requests per thread kept in registers between the claim and the result pass
constexpr uint32_t kNsCache = 256;        // unpinned requests of a namespace whose state lives in shared memory

// An unpinned request across the rounds (shared memory; written by the one warp that owns it).
// cur: what it proposed to last — a domain (scan form) or a candidate slot (two-level form).
struct NsReq {
  unsigned long long key;
  uint32_t r, key_lo, key_hi, size, cur, dead;
};

__device__ __forceinline__ void mbar_wait(unsigned long long* mbar) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}

// A tick's change list: a result row that differs from the previous tick's is appended (and the
// previous-tick copy brought up to date) by the thread that produced it.
__device__ __forceinline__ void diff_row(const PlaceNsArgs& a, uint32_t r, const uint4 v) {
  if (a.prev == nullptr) return;
  const uint4 o = __ldcg(reinterpret_cast<const uint4*>(a.prev + r));
  if (v.x != o.x || v.y != o.y || v.z != o.z || v.w != o.w) {
    const uint32_t i = atomicAdd(a.chg_count, 1u);
    if (i < a.chg_capacity) {
      a.chg_rows[i] = r;
      *reinterpret_cast<uint4*>(a.chg_outs + i) = v;
    }
    *reinterpret_cast<uint4*>(a.prev + r) = v;
  }
}

// Register budget: LWSE_NS_MAXNREG (build-time experiment switch) caps the kernel with __maxnreg__ instead of
// __launch_bounds__ (the two cannot be combined) so that a CTA of this kernel and CTAs of the sweep fit one SM's
// register file side by side in a tick.
#ifdef LWSE_NS_MAXNREG
#define LWSE_NS_BOUNDS __maxnreg__(LWSE_NS_MAXNREG)
#else
#define LWSE_NS_BOUNDS __launch_bounds__(kScan ? kNsScanThreads : kNsThreads, 2)
#endif
template <bool kScan>
__global__ void LWSE_NS_BOUNDS
    place_ns_kernel(const PlaceNsArgs a, const __grid_constant__ CUtensorMap words_map) {
  extern __shared__ uint8_t s_dyn[];
  // layout (from a 128-byte aligned base: the TMA destination):
  //   scan form, when they fit:  [words: word_rows x 256 u32] [node order: word_rows x 256 u32]
  //   [dom_free: n_domains u32] [dom_first: n_domains + 1 u32] [holder by domain: n_domains u64]
  //   two-level form:            [candidates: n_domains x {domain, free}] [holder by candidate: n_domains u64]
  //                              [slot of domain: n_domains u32]
  //   [request cache: kNsCache x 32 B] [mbarrier]
  uint8_t* s_raw = s_dyn + ((128u - (smem_u32(s_dyn) & 127u)) & 127u);
  const uint32_t table_bytes = (kScan && a.stage_words) ? a.word_rows * 1024u : 0u;
  const uint32_t free_bytes = ((a.n_domains * 4u) + 15u) & ~15u;
  const uint32_t first_bytes = (((a.n_domains + 1u) * 4u) + 15u) & ~15u;
  uint32_t* s_words = reinterpret_cast<uint32_t*>(s_raw);
  uint32_t* s_order = reinterpret_cast<uint32_t*>(s_raw + table_bytes);
  uint32_t* s_free = reinterpret_cast<uint32_t*>(s_raw + 2u * table_bytes);
  uint32_t* s_first = reinterpret_cast<uint32_t*>(s_raw + 2u * table_bytes + free_bytes);
  unsigned long long* s_hold = reinterpret_cast<unsigned long long*>(s_raw + 2u * table_bytes + free_bytes + first_bytes);
  uint2* s_cand = reinterpret_cast<uint2*>(s_hold + a.n_domains);
  unsigned long long* s_holdc = reinterpret_cast<unsigned long long*>(s_cand + (kScan ? 0u : a.n_domains));
  uint32_t* s_slot = reinterpret_cast<uint32_t*>(s_holdc + (kScan ? 0u : a.n_domains));
  NsReq* s_req = reinterpret_cast<NsReq*>(reinterpret_cast<uint8_t*>(s_slot) + (kScan ? 0u : free_bytes));
  unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_req + kNsCache);
  __shared__ uint32_t s_n_unp, s_n_cand, s_unsettled[3];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const uint32_t* words = table_bytes ? s_words : a.g_words;
  const uint32_t* order = table_bytes ? s_order : a.node_order;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // (no early griddepcontrol.launch_dependents: a dependent would only park its CTAs on SMs the
  // concurrent sweep wants; it still launches ahead and starts the moment this grid ends)
  pdl_wait_prior();  // the condense kernel's tables (and the request table a tick's scatter patched)
  if (tid == 0) {
    // TMA: the capacity vector and the domain index — and, scan form, the node-word table (a 2D
    // tensor of 256-word rows) and the node order — land in shared memory while the other threads
    // clear the holder table and take the pinned claims; one mbarrier collects the bytes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_mbar)),
                 "r"(2u * table_bytes + free_bytes + first_bytes)
                 : "memory");
    if (table_bytes) {
      if (a.use_tensor_map) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                         smem_u32(s_words)),
                     "l"(&words_map), "r"(smem_u32(s_mbar)), "r"(0), "r"(0)
                     : "memory");
      } else {
        bulk_load(s_words, a.g_words, table_bytes, s_mbar);
      }
      bulk_load(s_order, a.node_order, table_bytes, s_mbar);
    }
    bulk_load(s_free, a.g_dom_free, free_bytes, s_mbar);
    bulk_load(s_first, a.dom_first, first_bytes, s_mbar);
  }
  bool staged = false;

  for (uint32_t ns = blockIdx.x; ns < a.n_namespaces; ns += gridDim.x) {
    const uint32_t first = __ldcg(a.ns_first + ns), last = __ldcg(a.ns_first + ns + 1u);
    if (tid == 0) {
      s_n_unp = 0u;
      s_n_cand = 0u;
      s_unsettled[0] = s_unsettled[1] = s_unsettled[2] = 0u;
    }
    for (uint32_t d = tid; d < a.n_domains; d += blockDim.x) s_hold[d] = ~0ull;
    __syncthreads();
    if (first >= last) continue;  // CTA-uniform

    // ---- phase 1: pinned claims, the list of live unpinned requests ----
    // the first kNsKeep requests of a thread stay in registers until their result is written
    uint4 k_lo[kNsKeep], k_hi[kNsKeep];
    uint32_t k_dom[kNsKeep];
#pragma unroll
    for (uint32_t j = 0; j < kNsKeep; j++) {  // all request loads of the thread in flight together
      const uint32_t r = first + tid + j * blockDim.x;
      if (r < last) {
        k_lo[j] = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r));
        k_hi[j] = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r) + 1);
      }
    }
    auto claim = [&](uint32_t r, const uint4 lo, const uint4 hi) -> uint32_t {
      const uint32_t leader = hi.w;
      uint32_t d = LWSE_NONE;
      if (leader == LWSE_NONE) {
        const bool dead = (int32_t)hi.z < 1;
        const uint4 row = make_uint4(LWSE_NONE, LWSE_NONE, dead ? LWSE_PLACE_UNSCHEDULABLE : 0u, 0u);
        *reinterpret_cast<uint4*>(a.out + r) = row;
        if (dead) {
          diff_row(a, r, row);
        } else {
          const uint32_t k = atomicAdd(&s_n_unp, 1u);
          a.unpinned[first + k] = r;
          if (k < kNsCache) {
            NsReq q;
            q.key = ns_place_key(u64_of(lo.x, lo.y), r, false);
            q.r = r;
            q.key_lo = lo.z;
            q.key_hi = lo.w;
            q.size = hi.z;
            q.cur = LWSE_NONE;
            q.dead = 0u;
            s_req[k] = q;
          }
        }
      } else {
        if (leader < a.n_nodes) {
          const uint4 nr = ldg_keep(reinterpret_cast<const uint4*>(a.nodes + leader));
          if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && nr.z < a.n_domains) d = nr.z;
        }
        if (d != LWSE_NONE) atomicMin(s_hold + d, ns_place_key(u64_of(lo.x, lo.y), r, true));
      }
      return d;
    };
    auto result = [&](uint32_t r, const uint4 lo, const uint4 hi, uint32_t d) {
      if (hi.w == LWSE_NONE) return;
      uint32_t flags = LWSE_PLACE_PINNED;
      if (d != LWSE_NONE) flags |= s_hold[d] == ns_place_key(u64_of(lo.x, lo.y), r, true) ? LWSE_PLACE_PLACED : LWSE_PLACE_CONFLICT;
      const uint4 row = make_uint4(d, hi.w, flags, 0u);
      *reinterpret_cast<uint4*>(a.out + r) = row;
      diff_row(a, r, row);
    };
#pragma unroll
    for (uint32_t j = 0; j < kNsKeep; j++) {
      const uint32_t r = first + tid + j * blockDim.x;
      if (r < last) k_dom[j] = claim(r, k_lo[j], k_hi[j]);
    }
    for (uint32_t r = first + tid + kNsKeep * blockDim.x; r < last; r += blockDim.x)  // big namespaces: the rest, unkept
      claim(r, ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r)), ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r) + 1));
    __syncthreads();
    // results of the pinned requests: final from here on (pinned keys are below every unpinned key)
#pragma unroll
    for (uint32_t j = 0; j < kNsKeep; j++) {
      const uint32_t r = first + tid + j * blockDim.x;
      if (r < last) result(r, k_lo[j], k_hi[j], k_dom[j]);
    }
    for (uint32_t r = first + tid + kNsKeep * blockDim.x; r < last; r += blockDim.x) {
      const uint4 lo = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r)), hi = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + r) + 1);
      uint32_t d = LWSE_NONE;
      if (hi.w != LWSE_NONE && hi.w < a.n_nodes) {
        const uint4 nr = ldg_keep(reinterpret_cast<const uint4*>(a.nodes + hi.w));
        if (((nr.w >> 16) & LWSE_NODE_HAS_TOPOLOGY) && nr.z < a.n_domains) d = nr.z;
      }
      result(r, lo, hi, d);
    }
    const uint32_t n_unp = s_n_unp;
    if (n_unp == 0u) {
      __syncthreads();
      continue;
    }
    if (tid == 0) atomicAdd(a.counters + 1, n_unp);
    if (!staged) {  // the first namespace with work waits for the staged tables
      mbar_wait(s_mbar);
      staged = true;
    }
    uint32_t n_cand = 0;
    if constexpr (!kScan) {
      // The domains an unpinned request can ever win: no pinned holder (a pinned key is below every
      // unpinned key) and capacity left.  The rounds look at this list only; the holder table of
      // the unpinned claims is indexed by the list position.
      for (uint32_t d = tid; d < a.n_domains; d += blockDim.x) {
        const uint32_t fr = s_free[d];
        if (s_hold[d] == ~0ull && fr != 0u) {
          const uint32_t c = atomicAdd(&s_n_cand, 1u);
          s_cand[c] = make_uint2(d, fr);
          s_holdc[c] = ~0ull;
          s_slot[d] = c;
        }
      }
      __syncthreads();
      n_cand = s_n_cand;
    }

    // ---- deferred-acceptance rounds of this namespace, one warp per request ----
    uint32_t round = 0;
    for (;; round++) {
      // three rotating counters: the one for round k+1 is cleared during round k, when nobody can
      // still be reading it (it was last read after round k-2)
      uint32_t* unsettled_ctr = &s_unsettled[round % 3u];
      if (tid == 0) s_unsettled[(round + 1u) % 3u] = 0u;
      uint32_t scans = 0;
      for (uint32_t k = warp; k < n_unp; k += n_warps) {
        NsReq q;
        const bool cached = k < kNsCache;
        if (cached) {
          q = s_req[k];
        } else {  // overflow: the state lives in the result row (written by this warp only)
          q.r = __ldcg(a.unpinned + first + k);
          const uint4 o = __ldcg(reinterpret_cast<const uint4*>(a.out + q.r));
          const uint4 lo = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + q.r)), hi = ldg_keep(reinterpret_cast<const uint4*>(a.reqs + q.r) + 1);
          q.key = ns_place_key(u64_of(lo.x, lo.y), q.r, false);
          q.key_lo = lo.z;
          q.key_hi = lo.w;
          q.size = hi.z;
          q.cur = o.x;
          if (!kScan && o.x != LWSE_NONE) q.cur = s_slot[o.x];
          q.dead = (o.z & LWSE_PLACE_UNSCHEDULABLE) ? 1u : 0u;
        }
        __syncwarp();  // every lane has its copy before lane 0 updates the cache entry below
        if (q.dead) continue;  // warp-uniform
        const unsigned long long key = q.key;
        const uint32_t key_lo = q.key_lo, key_hi = q.key_hi, size = q.size, r = q.r;
        if (q.cur != LWSE_NONE && (kScan ? s_hold[q.cur] : s_holdc[q.cur]) == key) continue;  // still holds what it proposed to
        scans++;
        uint32_t H = 0, best_d = LWSE_NONE, best_n = LWSE_NONE, best_c = LWSE_NONE;
        if constexpr (!kScan) {
          // Level 1 — the domain, over the candidate list in shared memory (holders change under
          // our feet during a round; a stale reading can only make a proposal fail: holder keys
          // only decrease).
          uint32_t my_hi = 0, my_d = LWSE_NONE, my_c = LWSE_NONE;
#pragma unroll 4
          for (uint32_t c = lane; c < n_cand; c += 32u) {
            const uint2 cd = s_cand[c];
            const unsigned long long hk = s_holdc[c];
            const uint32_t h = (cd.y >= size && hk >= key) ? (mix32n(key_lo ^ (cd.x * 0x9E3779B1u)) | 1u) : 0u;
            if (h > my_hi || (h == my_hi && h != 0u && cd.x < my_d)) {  // ties keep the lower domain
              my_hi = h;
              my_d = cd.x;
              my_c = c;
            }
          }
          H = __reduce_max_sync(0xFFFFFFFFu, my_hi);
          if (H != 0u) {
            best_d = __reduce_min_sync(0xFFFFFFFFu, my_hi == H ? my_d : LWSE_NONE);
            const uint32_t owner = __ffs(__ballot_sync(0xFFFFFFFFu, my_hi == H && my_d == best_d)) - 1u;
            best_c = __shfl_sync(0xFFFFFFFFu, my_c, (int)owner);
            // Level 2 — the node: the domain's run of the condensed (free | node) words.
            const uint32_t f2 = s_first[best_d], l2 = s_first[best_d + 1u];
            uint32_t my_lo = 0, my_n = LWSE_NONE;
            for (uint32_t i = f2 + lane; i < l2; i += 32u) {
              const uint32_t w = __ldcg(a.g_nodes2 + i);
              if ((w >> 28) == 0u) continue;
              const uint32_t n = w & 0x0FFFFFFFu;
              const uint32_t l = ((w >> 28) << 28) | (mix32n(key_hi ^ (n * 0x85EBCA77u)) >> 4);
              if (l > my_lo || (l == my_lo && n < my_n)) {
                my_lo = l;
                my_n = n;
              }
            }
            const uint32_t L = __reduce_max_sync(0xFFFFFFFFu, my_lo);
            best_n = __reduce_min_sync(0xFFFFFFFFu, (my_lo == L && L != 0u) ? my_n : LWSE_NONE);
          }
        } else {
          // Brute force: every (request, node) pair of the staged table.  Lexicographic arg-max of
          // (domain score, lower domain, node score, lower node) — the two-level result.
          uint32_t my_hi = 0, my_d = LWSE_NONE, my_lo = 0, my_n = LWSE_NONE;
          for (uint32_t i = lane; i < a.n_usable; i += 32u) {
            const uint32_t w = words[i], d = w & 0x0FFFFFFFu;
            if ((w >> 28) == 0u || s_free[d] < size || s_hold[d] < key) continue;
            const uint32_t h = mix32n(key_lo ^ (d * 0x9E3779B1u)) | 1u;
            if (h < my_hi || (h == my_hi && d > my_d)) continue;
            const uint32_t n = order[i];
            const uint32_t l = ((w >> 28) << 28) | (mix32n(key_hi ^ (n * 0x85EBCA77u)) >> 4);
            const bool better = h > my_hi || d < my_d || l > my_lo || (l == my_lo && n < my_n);
            if (better) {
              my_hi = h;
              my_d = d;
              my_lo = l;
              my_n = n;
            }
          }
          H = __reduce_max_sync(0xFFFFFFFFu, my_hi);
          if (H != 0u) {
            best_d = __reduce_min_sync(0xFFFFFFFFu, my_hi == H ? my_d : LWSE_NONE);
            const bool in = my_hi == H && my_d == best_d;
            const uint32_t L = __reduce_max_sync(0xFFFFFFFFu, in ? my_lo : 0u);
            best_n = __reduce_min_sync(0xFFFFFFFFu, (in && my_lo == L && L != 0u) ? my_n : LWSE_NONE);
          }
          best_c = best_d;
        }
        if (lane == 0) {
          if (H == 0u || best_n == LWSE_NONE) {  // nothing feasible now, and the feasible set only shrinks
            store_row(a.out + r, LWSE_NONE, LWSE_NONE, LWSE_PLACE_UNSCHEDULABLE, 0u);
            if (cached) s_req[k].dead = 1u;
          } else {
            store_row(a.out + r, best_d, best_n, LWSE_PLACE_PLACED, H);
            if (cached) s_req[k].cur = best_c;
            // A claim on an empty domain settles at once; any other outcome leaves somebody without
            // a domain who proposes again next round: count it.
            const unsigned long long old = atomicMin((kScan ? s_hold : s_holdc) + best_c, key);
            if (old != ~0ull) atomicAdd(unsettled_ctr, 1u);
          }
        }
        __syncwarp();
      }
      if (lane == 0 && scans) atomicAdd(a.counters + 3, scans);
      __syncthreads();
      const uint32_t unsettled = *unsettled_ctr;
      if (unsettled == 0u || round > n_unp + 2u) break;
    }
    if (tid == 0) atomicMax(a.counters + 0, round + 1u);
    // a tick's change list: the rows of the unpinned requests are final now
    if (a.prev != nullptr) {
      for (uint32_t k = tid; k < n_unp; k += blockDim.x) {
        const uint32_t r = k < kNsCache ? s_req[k].r : __ldcg(a.unpinned + first + k);
        diff_row(a, r, __ldcg(reinterpret_cast<const uint4*>(a.out + r)));
      }
    }
    __syncthreads();
  }
  if (!staged) mbar_wait(s_mbar);  // never leave a bulk copy in flight behind a CTA that exits
}

// --------------------------------------------------------------------------
// launcher
// --------------------------------------------------------------------------
struct PlaceNsChanges {  // a tick's change list (device memory); prev == null: off
  lwse_place_out* prev;
  uint32_t* rows;
  lwse_place_out* outs;
  uint32_t* count;
  uint32_t capacity;
  int tick_slot;  // 0 / 1: the round's counters live in the tick slot's own block, which the tick's publish
                  // kernel clears after reading (two ticks may be in flight); -1: the two alternating blocks
  void* mid_event;  // cudaEvent_t or null: recorded between the condense and the namespace kernel (a tick graph
                    // holds the sweep back until then, so that both branches reach the SMs together)
};

struct PlaceNsExchange {  // multi-rank: d_occupancy is the base of the exchange buffers (see PlaceNsArgs::xch_step)
  const unsigned long long* step_ctr;
  uint64_t half_bytes;
  bool lagged;
};

static size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// scratch: [words (rows x 1 KB)] [nodes2 (same)] [dom_free] [ns_first] [unpinned] [counters: 2 blocks of 8 words]
struct NsLayout {
  size_t words, nodes2, dom_free, ns_first, unpinned, counters, total;
  uint32_t word_rows;
};
static NsLayout ns_layout(uint32_t n_nodes, uint32_t n_domains, uint32_t n_reqs, uint32_t n_namespaces) {
  NsLayout l{};
  l.word_rows = (n_nodes + 255u) / 256u;
  if (l.word_rows == 0) l.word_rows = 1;
  size_t off = 0;
  l.words = off;
  off += up256((size_t)l.word_rows * 1024);
  l.nodes2 = off;
  off += up256((size_t)l.word_rows * 1024);
  l.dom_free = off;
  off += up256((size_t)n_domains * 4 + 16);
  l.ns_first = off;
  off += up256(((size_t)n_namespaces + 2) * 4);
  l.unpinned = off;
  off += up256((size_t)n_reqs * 4 + 16);
  l.counters = off;
  off += 256;
  l.total = off;
  return l;
}
size_t place_ns_scratch_bytes(uint32_t n_nodes, uint32_t n_domains, uint32_t n_reqs, uint32_t n_namespaces) {
  return ns_layout(n_nodes, n_domains, n_reqs, n_namespaces).total;
}

// shared memory of the namespace kernel: the scan form with / without the staged word table, the
// two-level form (candidate list, holder table by candidate, slot of domain)
static size_t ns_smem_bytes(uint32_t word_rows, uint32_t n_domains, bool scan, bool stage) {
  const size_t free_bytes = ((size_t)n_domains * 4 + 15) & ~(size_t)15;
  const size_t base = free_bytes + ((((size_t)n_domains + 1) * 4 + 15) & ~(size_t)15) + (size_t)n_domains * 8 + 256 * 32 + 16 + 128;
  if (scan) return base + (stage ? (size_t)word_rows * 2048 : 0);
  return base + (size_t)n_domains * 16 + free_bytes;
}
bool place_ns_supported(uint32_t n_nodes, uint32_t n_domains) {
  return ns_smem_bytes((n_nodes + 255u) / 256u, n_domains, false, false) <= 200u * 1024u && n_domains < (1u << 28) &&
         n_nodes < (1u << 28);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode_words_map(CUtensorMap* map, void* g_words, uint32_t word_rows) {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  if (!fn || word_rows == 0 || word_rows > 256) return false;
  const cuuint64_t dims[2] = {256, word_rows};
  const cuuint64_t strides[1] = {1024};
  const cuuint32_t box[2] = {256, word_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, g_words, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns kernels launched or -1.  `fresh`: the scratch was (re)allocated: zero it first.
int launch_place_ns(const lwse_node_rec* d_nodes, const uint32_t* d_dom_first, const uint32_t* d_node_order, uint32_t n_nodes,
                    uint32_t n_usable, uint32_t n_domains, const lwse_place_req* d_reqs, uint32_t n_reqs,
                    const uint32_t* d_occupancy, uint32_t n_parts, uint64_t part_stride_bytes, uint32_t n_namespaces,
                    lwse_place_out* d_out, void* d_scratch, size_t scratch_bytes, bool fresh, uint32_t call_index, bool scan,
                    int sm_count, cudaStream_t s, int* cuda_err, const uint32_t** d_counters_out, bool first_pdl,
                    const PlaceNsChanges* changes, const PlaceNsExchange* xch) {
  *cuda_err = 0;
  const NsLayout l = ns_layout(n_nodes, n_domains, n_reqs, n_namespaces);
  if (scratch_bytes < l.total || n_reqs > 0xFFFFFFu) {
    *cuda_err = (int)cudaErrorInvalidValue;
    return -1;
  }
  uint8_t* base = static_cast<uint8_t*>(d_scratch);
  cudaError_t e = cudaSuccess;
  if (fresh) {
    e = cudaMemsetAsync(base, 0, l.total, s);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
  }
  PlaceNsArgs a{};
  a.nodes = d_nodes;
  a.reqs = d_reqs;
  a.occupancy = d_occupancy;
  a.out = d_out;
  a.dom_first = d_dom_first;
  a.node_order = d_node_order;
  a.g_words = reinterpret_cast<uint32_t*>(base + l.words);
  a.g_nodes2 = reinterpret_cast<uint32_t*>(base + l.nodes2);
  a.g_dom_free = reinterpret_cast<uint32_t*>(base + l.dom_free);
  a.ns_first = reinterpret_cast<uint32_t*>(base + l.ns_first);
  a.unpinned = reinterpret_cast<uint32_t*>(base + l.unpinned);
  // counters: 256 bytes = 8 blocks of 8 words.  Blocks 0 / 1 alternate between calls (a call clears the
  // block the next one uses); blocks 2 / 3 belong to the tick slots.
  const int tick_slot = (changes && changes->prev) ? changes->tick_slot : -1;
  if (tick_slot >= 0) {
    a.counters = reinterpret_cast<uint32_t*>(base + l.counters) + (2u + (uint32_t)(tick_slot & 1)) * 8u;
    a.next_counters = nullptr;
  } else {
    a.counters = reinterpret_cast<uint32_t*>(base + l.counters) + (call_index & 1u) * 8u;
    a.next_counters = reinterpret_cast<uint32_t*>(base + l.counters) + ((call_index + 1u) & 1u) * 8u;
  }
  a.n_nodes = n_nodes;
  a.n_usable = n_usable;
  a.n_domains = n_domains;
  a.n_reqs = n_reqs;
  a.n_namespaces = n_namespaces;
  a.n_parts = d_occupancy ? (n_parts ? n_parts : 1u) : 0u;
  a.part_stride_bytes = part_stride_bytes;
  a.word_rows = l.word_rows;
  a.scan = scan ? 1u : 0u;
  if (xch && xch->step_ctr) {
    a.xch_step = xch->step_ctr;
    a.xch_half_bytes = xch->half_bytes;
    a.xch_lagged = xch->lagged ? 1u : 0u;
  }
  if (d_counters_out) *d_counters_out = a.counters;
  if (changes && changes->prev) {
    a.prev = changes->prev;
    a.chg_rows = changes->rows;
    a.chg_outs = changes->outs;
    a.chg_count = changes->count;
    a.chg_capacity = changes->capacity;
  }

  // does the word table fit next to the holder table?  (227 KB of shared memory per CTA)
  const size_t with_words = ns_smem_bytes(l.word_rows, n_domains, true, true);
  a.stage_words = (scan && with_words <= 100u * 1024u) ? 1u : 0u;  // two CTAs per SM keep their tables side by side
  const size_t smem = ns_smem_bytes(l.word_rows, n_domains, scan, a.stage_words != 0);
  // the tensor map of the word table: re-encoded when the scratch moved
  static thread_local CUtensorMap map;
  static thread_local void* map_for = nullptr;
  static thread_local uint32_t map_rows = 0;
  static thread_local bool map_ok = false;
  if (a.stage_words && (map_for != a.g_words || map_rows != l.word_rows)) {
    map_ok = encode_words_map(&map, a.g_words, l.word_rows);
    map_for = a.g_words;
    map_rows = l.word_rows;
  }
  static const bool no_tensor_map = [] {
    const char* v = getenv("LWSE_PLACE_NO_TENSOR_MAP");
    return v && atoi(v) != 0;
  }();
  a.use_tensor_map = (a.stage_words && map_ok && !no_tensor_map) ? 1u : 0u;

  {  // condense: a warp per domain, a thread per request
    uint64_t threads = (uint64_t)n_domains * 32u;
    if (threads < n_reqs) threads = n_reqs;
    uint64_t grid = (threads + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > (uint64_t)sm_count * 8u) grid = (uint64_t)sm_count * 8u;
    // first_pdl: the predecessor on the stream is a kernel — programmatically dependent then (the
    // kernel waits at its top, before its first read, so this is safe even behind a tick's scatter
    // kernel, whose tail the launch latency overlaps); behind a copy / memset / event an ordinary launch
    e = launch_pdl(place_condense_kernel, dim3((unsigned)grid), dim3(256), 0, s, first_pdl, a);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    if (changes && changes->mid_event) {
      e = cudaEventRecord(static_cast<cudaEvent_t>(changes->mid_event), s);
      if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    }
  }
  {
    unsigned grid = n_namespaces < (unsigned)sm_count * 2u ? n_namespaces : (unsigned)sm_count * 2u;
    if (grid < 1u) grid = 1u;
    auto kern = scan ? place_ns_kernel<true> : place_ns_kernel<false>;
    static thread_local size_t attr_set[2] = {0, 0};
    if (smem > attr_set[scan ? 1 : 0]) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
      attr_set[scan ? 1 : 0] = smem;
    }
    e = launch_pdl(kern, dim3(grid), dim3(scan ? kNsScanThreads : kNsThreads), smem, s, true, a, map);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
  }
  return 2;
}

cudaError_t set_place_ns_carveout(int pct) {  // see set_sweep_carveout (lwse_lws_kernels.cu)
  cudaError_t e = cudaSuccess;
  auto set = [&](const void* k) {
    const cudaError_t r = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (r != cudaSuccess) e = r;
  };
  set((const void*)place_condense_kernel);
  set((const void*)place_ns_kernel<false>);
  set((const void*)place_ns_kernel<true>);
  return e;
}

}  // namespace lwse
