// LWS sweep kernels (sm_100a).  One sweep = two launches on one stream when groups are small
// and no per-node pod count is wanted — group_fused_kernel (the pod scan and the group pass of
// 256 consecutive groups in one CTA, bitmaps in shared memory) and lws_sweep_kernel — and
// three otherwise:
//
//   pod_scan_kernel<U>   pod-centric streaming pass over the 4-byte pod state
//                        column: every lane takes U pods with coalesced loads
//                        (all U loads in flight before the first use), derives
//                        "Pending" (pod_controller.go:356) and "has a restart or
//                        deletion event" (pod_utils.go:29-50) per pod and packs
//                        them with __ballot_sync into two bitmaps (1 bit / pod).
//                        No dependence on group rows → pure HBM streaming.
//   group_sweep_kernel   one lane per pod group.  Reads its 64-byte group row,
//                        the first 16 bytes of its owner row and the bitmap words
//                        of its pod range; only pods whose event bit is set are
//                        visited individually (state word + 12-byte identity row:
//                        workerPodBelongsToLeader, pod_controller.go:268-295).
//                        Emits the per-replica state bits
//                        (leaderworkerset_controller.go:608-638, :433-476), the
//                        restart verdict (pod_controller.go:204-266) and the
//                        leader pod's worker-sts gating (:100-198).  16 B out.
//   lws_sweep_kernel<W>  one W-lane tile per LeaderWorkerSet over its groups'
//                        flag words: counters via redux.sync, the five cases of
//                        rollingUpdateParameters, the partition walk (:643-673)
//                        as three reductions.  32 B out per object.
//
// All three are integer/compare kernels bound by HBM traffic.
#include <cstdlib>

#include "lwse_device.cuh"

namespace lwse {

// --------------------------------------------------------------------------
// pod scan
// --------------------------------------------------------------------------
struct PodScanArgs {
  const uint32_t* state;
  uint32_t* pending_bits;  // ceil(n_pods / 32) words
  uint32_t* event_bits;
  uint32_t* occupancy;  // nullable: scheduled pods per node
  uint64_t n_pods;
  uint32_t n_nodes;
  uint32_t* event_count;  // nullable: += pods with an event bit (the host entry point sizes its
                          // identity-column transfer with it)
};

__device__ __forceinline__ bool pod_has_event(uint32_t bits) {
  const uint32_t phase = bits & LWSE_POD_PHASE_MASK;
  // ContainerRestarted (pod_utils.go:29-45) || PodDeleted (:48)
  return ((phase - 1u) < 2u && (bits & LWSE_POD_ANY_RESTART)) || (bits & LWSE_POD_DELETING);
}

// 4 predicate bits of the 4 pods a lane holds, moved to the lane's nibble of its
// 8-lane segment; a 3-step xor butterfly ORs the segment into one bitmap word
// (32 consecutive pods) that every lane of the segment ends up holding.
// (__reduce_or_sync with a sub-warp mask compiles to a serialised per-segment
// loop on sm_100a — measured slower — so the butterfly uses plain shuffles.)
__device__ __forceinline__ uint32_t seg8_or(uint32_t nibble, uint32_t lane) {
  uint32_t v = nibble << ((lane & 7u) * 4u);
  v |= __shfl_xor_sync(0xFFFFFFFFu, v, 1);
  v |= __shfl_xor_sync(0xFFFFFFFFu, v, 2);
  v |= __shfl_xor_sync(0xFFFFFFFFu, v, 4);
  return v;
}

// SWAR over the 4 pods of a lane: byte k of X = low byte of pod k's state word (phase bits 0-1,
// any-restart bit 2, deleting bit 3), both predicates evaluated on the four bytes at once, then
// the four bit-0s are gathered into a nibble with one multiply.
__device__ __forceinline__ void pod4_predicates(const uint4 v, uint32_t& pend, uint32_t& ev) {
  const uint32_t X = __byte_perm(__byte_perm(v.x, v.y, 0x0040), __byte_perm(v.z, v.w, 0x0040), 0x5410);
  const uint32_t y = X >> 1, z = X >> 2, w3 = X >> 3;
  const uint32_t P = X & ~y & 0x01010101u;                // phase == Pending (bit0 ∧ ¬bit1)
  const uint32_t E = (((X ^ y) & z) | w3) & 0x01010101u;  // (phase ∈ {Pending,Running} ∧ restart) ∨ deleting
  pend = (P * 0x10204080u) >> 28;
  ev = (E * 0x10204080u) >> 28;
}

template <int U, bool OCC>
__global__ void __launch_bounds__(256) pod_scan_kernel(const PodScanArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = (uint64_t)blockIdx.x * 8u + (threadIdx.x >> 5);
  const uint64_t n_warps = (uint64_t)gridDim.x * 8u;
  const uint64_t n_words = (a.n_pods + 31u) >> 5;
  constexpr uint64_t kPodsPerChunk = 128ull * U;  // 32 lanes x 4 pods x U
  const uint4* vec = reinterpret_cast<const uint4*>(a.state);
  pdl_launch_dependents();
  bool waited = false;
  uint32_t events = 0;
  for (uint64_t base = warp * kPodsPerChunk; base < a.n_pods; base += n_warps * kPodsPerChunk) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; j++) {  // all U 128-bit loads in flight before the first use
      const uint64_t idx = base + (uint64_t)j * 128u + lane * 4u;
      if (idx + 3u < a.n_pods) {
        v[j] = ldg_stream(vec + (idx >> 2));
      } else {  // ragged tail of the column
        v[j].x = idx + 0u < a.n_pods ? __ldg(a.state + idx + 0u) : 0u;
        v[j].y = idx + 1u < a.n_pods ? __ldg(a.state + idx + 1u) : 0u;
        v[j].z = idx + 2u < a.n_pods ? __ldg(a.state + idx + 2u) : 0u;
        v[j].w = 0u;
      }
    }
    if (!waited) {  // the bitmaps (and counters) may still be read by the previous sweep's group pass
      pdl_wait_prior();
      waited = true;
    }
#pragma unroll
    for (int j = 0; j < U; j++) {
      uint32_t pend, ev;
      pod4_predicates(v[j], pend, ev);
      if (OCC) {
        const uint32_t b[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (b[k] & LWSE_POD_SCHEDULED) {
            const uint32_t node = b[k] >> LWSE_POD_NODE_SHIFT;
            if (node < a.n_nodes) atomicAdd(a.occupancy + node, 1u);
          }
        }
      }
      const uint32_t wp = seg8_or(pend, lane), we = seg8_or(ev, lane);
      const uint64_t w = ((base + (uint64_t)j * 128u) >> 5) + (lane >> 3);
      if ((lane & 7u) == 0u && w < n_words) {
        a.pending_bits[w] = wp;
        a.event_bits[w] = we;
        events += __popc(we);
      }
    }
  }
  if (a.event_count != nullptr) {
    events = __reduce_add_sync(0xFFFFFFFFu, events);
    if (lane == 0 && events) atomicAdd(a.event_count, events);
  }
}

// --------------------------------------------------------------------------
// group pass
// --------------------------------------------------------------------------
// Optional change list: result rows that differ from what the output table held before.
struct ChangeList {
  uint32_t* rows;   // nullptr = off
  void* outs;       // packed result rows
  uint32_t* count;  // device counter
  uint32_t capacity;
};

template <int N>  // N = uint4 per result row
__device__ __forceinline__ void emit_if_changed(const ChangeList& c, uint4* slot, uint32_t row, const uint4 (&v)[N]) {
  if (c.rows != nullptr) {
    bool diff = false;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const uint4 old = slot[k];
      diff |= old.x != v[k].x || old.y != v[k].y || old.z != v[k].z || old.w != v[k].w;
    }
    if (diff) {
      const uint32_t i = atomicAdd(c.count, 1u);
      if (i < c.capacity) {
        c.rows[i] = row;
#pragma unroll
        for (int k = 0; k < N; k++) reinterpret_cast<uint4*>(c.outs)[(size_t)i * N + k] = v[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N; k++) stg_stream(slot + k, v[k]);
}

struct GroupSweepArgs {
  const lwse_lws_rec* lws;
  const lwse_group_rec* groups;
  const uint32_t* pod_state;
  const lwse_pod_ident* pod_ident;
  const uint32_t* pending_bits;
  const uint32_t* event_bits;
  const lwse_node_rec* nodes;
  lwse_group_out* out;
  uint64_t n_pods;
  uint32_t n_lws;
  uint32_t n_groups;
  uint32_t n_nodes;
  uint32_t sweep_flags;
  ChangeList changes;
};

// bits [lo, hi) of a 32-bit word, 0 <= lo <= hi <= 32
__device__ __forceinline__ uint32_t bit_range(uint32_t lo, uint32_t hi) {
  const uint32_t upto_hi = hi >= 32u ? 0xFFFFFFFFu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u) & (lo >= 32u ? 0u : 0xFFFFFFFFu);
}

// Where a group's bitmap words come from: the scan kernel's global bitmaps, or the
// shared-memory window of the fused kernel.
struct GlobalBits {
  const uint32_t* pending_bits;
  const uint32_t* event_bits;
  __device__ __forceinline__ uint32_t pending(uint32_t w) const { return __ldg(pending_bits + w); }
  __device__ __forceinline__ uint32_t event(uint32_t w) const { return __ldg(event_bits + w); }
};

// The group pass proper for group g (row ca..cd, first 16 bytes of its owner L): one W-lane
// tile per group, the lanes share the group's bitmap words (lane j takes words j, j+W, …) so
// that the rare per-pod visits of one group run in parallel; W = 1 for small groups.
// Returns the 16-byte result row.
template <int W, class Bits>
__device__ __forceinline__ uint4 group_body(const GroupSweepArgs& a, const Bits& bits, uint32_t lane, const uint4 ca,
                                            const uint4 cb, const uint4 cc, const uint4 cd, const uint4 L,
                                            const bool bad) {
  const uint32_t pod_base = cc.z, pod_count = cc.w, gflags = cd.y;
  uint32_t oflags = 0, first_out = LWSE_NONE, domain = LWSE_NONE;
  int32_t worker_replicas = 0;
  if (bad) {
    oflags = LWSE_GOUT_BAD_TABLE;
  } else {
    const int32_t size = (int32_t)L.z;
    const uint32_t lflags = L.w;
    const uint32_t policy = (lflags & LWSE_LWS_RESTART_MASK) >> LWSE_LWS_RESTART_SHIFT;
    const bool policy_on = policy == LWSE_RESTART_ON_POD_RESTART || policy == LWSE_RESTART_AFTER_START;

    // ---- pendingPodsInGroup :338-362 from the pending bitmap ----
    const uint32_t pod_end = pod_base + pod_count;  // <= n_pods < 2^32 (checked by the entry points)
    const uint32_t w_first = pod_base >> 5, w_last = pod_count ? ((pod_end - 1u) >> 5) : w_first;
    uint32_t any_bits = 0;  // bit0 pending, bit1 event
    if (pod_count) {
      for (uint32_t w = w_first + lane; w <= w_last; w += W) {
        const uint32_t lo = w == w_first ? (pod_base & 31u) : 0u;
        const uint32_t hi = w == w_last ? (((pod_end - 1u) & 31u) + 1u) : 32u;
        const uint32_t m = bit_range(lo, hi);
        if (bits.pending(w) & m) any_bits |= 1u;
        if (bits.event(w) & m) any_bits |= 2u;
      }
    }
    any_bits = tile_or<W>(any_bits);
    const bool pending = (uint32_t)size != pod_count || (any_bits & 1u);
    if (pending) oflags |= LWSE_GOUT_PENDING;
    // :222 skip when pending ∧ (AfterStart ∨ annotation)
    const bool suppressed =
        pending && (policy == LWSE_RESTART_AFTER_START || (lflags & LWSE_LWS_RECREATE_AFTER_START_ANNOT));

    // ---- handleRestartPolicy :204-266 for the pods that have an event ----
    bool leader_deleted = false;
    if ((any_bits & 2u) && policy_on && !suppressed) {  // tile-uniform
      const bool leader_found = (gflags & (LWSE_GRP_POD_PRESENT | LWSE_GRP_POD_NAME_MATCH)) ==
                                (LWSE_GRP_POD_PRESENT | LWSE_GRP_POD_NAME_MATCH);  // :233
      constexpr uint32_t kChain =
          LWSE_GRP_WSTS_FOUND | LWSE_GRP_WSTS_OWNER_IS_POD | LWSE_GRP_WSTS_OWNER_NAME_MATCH;
      const bool wsts_chain_ok = (gflags & kChain) == kChain && cc.x == cb.z;  // sts owner uid == leader uid
      uint32_t acc = 0, first = LWSE_NONE;
      // The pods to visit are collected four at a time so that their state and
      // identity loads are all in flight together (one DRAM round trip per batch
      // instead of one per pod).
      auto visit = [&](const uint32_t* ev, int cnt) {
        uint32_t bits[4], id_lo[4], id_hi[4], id_owner[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (k < cnt) {
            const uint32_t p = ev[k];
            bits[k] = __ldg(a.pod_state + p);
            const uint32_t* idp = reinterpret_cast<const uint32_t*>(a.pod_ident + p);
            id_lo[k] = __ldg(idp);
            id_hi[k] = __ldg(idp + 1);
            id_owner[k] = __ldg(idp + 2);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (k < cnt) {
            const uint32_t b = bits[k];
            bool cand, deleting;
            if (b & LWSE_POD_IS_LEADER) {
              cand = true;  // leader = pod (:251)
              deleting = b & LWSE_POD_DELETING;
            } else if (!(b & LWSE_POD_NAME_OK)) {
              acc |= LWSE_GOUT_RESTART_ERROR;  // :230
              cand = false;
              deleting = false;
            } else {
              const uint32_t kind = (b & LWSE_POD_OWNER_MASK) >> LWSE_POD_OWNER_SHIFT;
              // workerPodBelongsToLeader :268-295
              const bool belongs = (b & LWSE_POD_OWNER_NAME_MATCH) &&
                                   ((kind == 1u && id_owner[k] == cb.z) ||
                                    (kind == 2u && id_owner[k] == cb.w && wsts_chain_ok));
              cand = leader_found && id_lo[k] == ca.x && id_hi[k] == ca.y && belongs;  // :239
              deleting = gflags & LWSE_GRP_POD_DELETING;
            }
            if (cand) {
              acc |= deleting ? LWSE_GOUT_LEADER_DELETING : LWSE_GOUT_DELETE_LEADER;  // :255 / :259
              if (b & LWSE_POD_IS_LEADER) acc |= 0x80000000u;
              first = min(first, ev[k] - pod_base);
            }
          }
        }
      };
      uint32_t ev[4];
      int cnt = 0;
      for (uint32_t w = w_first + lane; w <= w_last; w += W) {
        const uint32_t lo = w == w_first ? (pod_base & 31u) : 0u;
        const uint32_t hi = w == w_last ? (((pod_end - 1u) & 31u) + 1u) : 32u;
        uint32_t m = bits.event(w) & bit_range(lo, hi);
        while (m) {
          const uint32_t p = (w << 5) + (__ffs(m) - 1u);
          m &= m - 1u;
          // fixed-slot insert keeps ev[] in registers
          if (cnt == 0) ev[0] = p;
          else if (cnt == 1) ev[1] = p;
          else if (cnt == 2) ev[2] = p;
          else ev[3] = p;
          if (++cnt == 4) {
            visit(ev, 4);
            cnt = 0;
          }
        }
      }
      if (cnt) visit(ev, cnt);
      acc = tile_or<W>(acc);
      first_out = tile_min<W>(first);
      leader_deleted = acc & 0x80000000u;
      oflags |= acc & 0x7FFFFFFFu;
    }

    // ---- per-replica state bits (consumed by lws_sweep_kernel) ----
    const bool no_wsts = size == 1;
    const uint64_t rev = u64_of(L.x, L.y);
    const bool leader_updated = u64_of(ca.x, ca.y) == rev;
    const bool wsts_updated = u64_of(ca.z, ca.w) == rev;
    const bool leader_ready = (gflags & (LWSE_GRP_POD_RUNNING | LWSE_GRP_POD_READY)) ==
                              (LWSE_GRP_POD_RUNNING | LWSE_GRP_POD_READY);  // PodRunningAndReady
    const bool wsts_ready = cb.x == cb.y && (gflags & LWSE_GRP_WSTS_REV_SETTLED);  // StatefulsetReady
    const bool ready = leader_ready && (no_wsts || wsts_ready);
    const bool updated = leader_updated && (no_wsts || wsts_updated);
    // getReplicaStates :609-617 — names decide whether the slot is live
    const bool named = (gflags & LWSE_GRP_POD_NAME_MATCH) &&
                       (no_wsts || (gflags & LWSE_GRP_WSTS_LABEL_NAME_MATCH));
    if (named && ready) oflags |= LWSE_GOUT_STATE_READY;
    if (named && updated) oflags |= LWSE_GOUT_STATE_UPDATED;
    // updateConditions :433-476 — existing leader pods whose worker sts is found
    const bool counted = (gflags & LWSE_GRP_POD_PRESENT) && (no_wsts || (gflags & LWSE_GRP_WSTS_FOUND));
    if (counted) {
      oflags |= LWSE_GOUT_COUNTED;
      if (ready) oflags |= LWSE_GOUT_COND_READY;
      if (updated) oflags |= LWSE_GOUT_COND_UPDATED;
    }

    // ---- the leader pod's own Reconcile tail, pod_controller.go:95-198 ----
    bool go_on = (gflags & LWSE_GRP_POD_PRESENT) && !leader_deleted &&
                 !(gflags & (LWSE_GRP_MISTAKEN_ANNOTATION | LWSE_GRP_POD_DELETING));
    if (go_on) {
      if (a.sweep_flags & LWSE_SWEEP_GANG) oflags |= LWSE_GOUT_CREATE_PODGROUP;  // :130
      go_on = !no_wsts &&                                                          // :138
              !((lflags & LWSE_LWS_STARTUP_LEADER_READY) && !(gflags & LWSE_GRP_POD_READY));  // :143
    }
    if (go_on && !(gflags & LWSE_GRP_REVISION_EXISTS)) {  // :152
      oflags |= LWSE_GOUT_REQUEUE_REVISION;
      go_on = false;
    }
    if (go_on && (lflags & LWSE_LWS_EXCLUSIVE_TOPOLOGY)) {  // :162
      const uint32_t node = cc.y;
      if (node == LWSE_NONE) {  // :164
        oflags |= LWSE_GOUT_WAIT_SCHEDULE;
        go_on = false;
      } else if (node != LWSE_NODE_NOT_FOUND && node < a.n_nodes) {
        const uint4 nr = ldg_cached(reinterpret_cast<const uint4*>(a.nodes + node));
        const uint32_t nflags = nr.w >> 16;
        if (!(nflags & LWSE_NODE_HAS_TOPOLOGY)) {  // :330
          oflags |= LWSE_GOUT_TOPOLOGY_ERROR;
          go_on = false;
        } else {
          domain = nr.z;
        }
      }  // Node NotFound → empty value, nil error (:327)
    }
    if (go_on && !(gflags & LWSE_GRP_WSTS_FOUND)) {  // :188-192
      oflags |= LWSE_GOUT_CREATE_WSTS;
      worker_replicas = size - 1;  // :437; ordinals start at 1 (:440)
    }
  }
  return make_uint4(oflags, first_out, (uint32_t)worker_replicas, domain);
}

// 128-thread CTAs capped at 64 registers: a tick's placement round holds part of the register
// file of some SMs while this kernel runs, and the groups of a 100k-group table should still
// fit in one wave next to it.
constexpr uint32_t kGroupThreads = 128;
template <int W>
__global__ void __launch_bounds__(kGroupThreads, 8) group_sweep_kernel(const GroupSweepArgs a) {
  constexpr uint32_t kTilesPerBlock = kGroupThreads / W;
  const uint32_t lane = threadIdx.x & (W - 1);
  const uint32_t stride = gridDim.x * kTilesPerBlock;
  const GlobalBits bits{a.pending_bits, a.event_bits};
  pdl_launch_dependents();
  bool waited = false;
  for (uint32_t g = blockIdx.x * kTilesPerBlock + threadIdx.x / W; g < a.n_groups; g += stride) {
    const uint4* row = reinterpret_cast<const uint4*>(a.groups + g);
    const uint4 ca = ldg_cached(row + 0), cb = ldg_cached(row + 1), cc = ldg_cached(row + 2),
                cd = ldg_cached(row + 3);
    const uint32_t pod_base = cc.z, pod_count = cc.w, lws_index = cd.x;
    const bool bad = lws_index >= a.n_lws || (uint64_t)pod_base + pod_count > a.n_pods;
    // owner row: only its first 16 bytes (rev_hash, size, flags)
    uint4 L = make_uint4(0, 0, 0, 0);
    if (!bad) L = ldg_cached(reinterpret_cast<const uint4*>(a.lws + lws_index));
    if (!waited) {  // everything above is input; the scan's bitmaps and group_out come next
      pdl_wait_prior();
      waited = true;
    }
    const uint4 v[1] = {group_body<W>(a, bits, lane, ca, cb, cc, cd, L, bad)};
    if (lane == 0) emit_if_changed<1>(a.changes, reinterpret_cast<uint4*>(a.out + g), g, v);
  }
}

// --------------------------------------------------------------------------
// fused pod scan + group pass
// --------------------------------------------------------------------------
// One CTA per 256 consecutive groups, one thread per group.  The CTA first streams the pod
// state words of its groups' pod window (the union of their ranges: contiguous for tables the
// encoder lays out, pods of group g right after those of g-1) with the scan's coalesced 128-bit
// loads and packs the two predicate bitmaps into SHARED memory; the group pass then reads its
// bitmap words from there.  Against the two-kernel form this drops one kernel boundary and the
// bitmaps' round trip through L2, and the state words of the event pods the group pass visits
// were just loaded by the same SM.  Used when no per-node occupancy is wanted (that count needs
// every pod exactly once, whatever the group table says) and groups are small (W = 1).
// A window that does not fit (irregular tables: ranges far apart) falls back to deriving each
// bitmap word from the state column directly — slow, still exact.
constexpr uint32_t kFusedThreads = 256;
constexpr uint32_t kWinWords = 2048;  // 65 536 pods per CTA: 2 x 8 KB of shared memory

struct WindowBits {
  const uint32_t* pend;  // word (w - w0)
  const uint32_t* ev;
  uint32_t w0;
  __device__ __forceinline__ uint32_t pending(uint32_t w) const { return pend[w - w0]; }
  __device__ __forceinline__ uint32_t event(uint32_t w) const { return ev[w - w0]; }
};

struct DirectBits {
  const uint32_t* state;
  uint64_t n_pods;
  __device__ __forceinline__ void word(uint32_t w, uint32_t& pend, uint32_t& ev) const {
    pend = ev = 0;
    for (uint32_t k = 0; k < 32u; k++) {
      const uint64_t p = ((uint64_t)w << 5) + k;
      if (p >= n_pods) break;
      const uint32_t b = __ldg(state + p);
      if ((b & LWSE_POD_PHASE_MASK) == LWSE_POD_PHASE_PENDING) pend |= 1u << k;
      if (pod_has_event(b)) ev |= 1u << k;
    }
  }
  __device__ __forceinline__ uint32_t pending(uint32_t w) const {
    uint32_t p, e;
    word(w, p, e);
    return p;
  }
  __device__ __forceinline__ uint32_t event(uint32_t w) const {
    uint32_t p, e;
    word(w, p, e);
    return e;
  }
};

__global__ void __launch_bounds__(kFusedThreads, 3) group_fused_kernel(const GroupSweepArgs a) {
  __shared__ uint32_t s_pend[kWinWords], s_ev[kWinWords];
  __shared__ uint32_t s_lo[kFusedThreads / 32], s_hi[kFusedThreads / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t g = blockIdx.x * kFusedThreads + tid;
  const bool valid = g < a.n_groups;
  pdl_launch_dependents();

  uint4 ca = make_uint4(0, 0, 0, 0), cb = ca, cc = ca, cd = ca, L = ca;
  bool bad = true;
  if (valid) {
    const uint4* row = reinterpret_cast<const uint4*>(a.groups + g);
    ca = ldg_cached(row + 0), cb = ldg_cached(row + 1), cc = ldg_cached(row + 2), cd = ldg_cached(row + 3);
    bad = cd.x >= a.n_lws || (uint64_t)cc.z + cc.w > a.n_pods;
    if (!bad) L = ldg_cached(reinterpret_cast<const uint4*>(a.lws + cd.x));  // first 16 bytes of the owner row
  }
  // the CTA's window of bitmap words
  const bool has_pods = valid && !bad && cc.w != 0u;
  uint32_t lo = has_pods ? (cc.z >> 5) : 0xFFFFFFFFu;
  uint32_t hi = has_pods ? (((cc.z + cc.w - 1u) >> 5) + 1u) : 0u;
  lo = __reduce_min_sync(0xFFFFFFFFu, lo);
  hi = __reduce_max_sync(0xFFFFFFFFu, hi);
  if (lane == 0) {
    s_lo[warp] = lo;
    s_hi[warp] = hi;
  }
  __syncthreads();
  lo = s_lo[0], hi = s_hi[0];
#pragma unroll
  for (int k = 1; k < (int)(kFusedThreads / 32); k++) {
    lo = min(lo, s_lo[k]);
    hi = max(hi, s_hi[k]);
  }
  const uint32_t n_words = hi > lo ? hi - lo : 0u;
  const bool windowed = n_words <= kWinWords;  // CTA-uniform

  if (windowed && n_words) {
    // the scan: 128 pods (4 bitmap words) per warp and chunk, U chunks in flight per warp
    // (U = 8 is 0.7 us faster alone but takes 80 registers: three resident CTAs would then leave
    // no SM with room for a placement CTA and the tick serialises — 27.0 us against 22.7 us)
    constexpr int U = 4;
    const uint32_t n_chunks = (n_words + 3u) >> 2;
    const uint4* vec = reinterpret_cast<const uint4*>(a.pod_state);
    for (uint32_t c0 = warp * U; c0 < n_chunks; c0 += (kFusedThreads / 32) * U) {
      uint4 v[U];
#pragma unroll
      for (int j = 0; j < U; j++) {
        const uint32_t c = c0 + (uint32_t)j;
        const uint64_t idx = ((uint64_t)lo << 5) + (uint64_t)c * 128u + lane * 4u;
        if (c >= n_chunks) {
          v[j] = make_uint4(0, 0, 0, 0);
        } else if (idx + 3u < a.n_pods) {
          v[j] = ldg_stream(vec + (idx >> 2));
        } else {  // ragged tail of the column
          v[j].x = idx + 0u < a.n_pods ? __ldg(a.pod_state + idx + 0u) : 0u;
          v[j].y = idx + 1u < a.n_pods ? __ldg(a.pod_state + idx + 1u) : 0u;
          v[j].z = idx + 2u < a.n_pods ? __ldg(a.pod_state + idx + 2u) : 0u;
          v[j].w = 0u;
        }
      }
#pragma unroll
      for (int j = 0; j < U; j++) {
        uint32_t pend, ev;
        pod4_predicates(v[j], pend, ev);
        const uint32_t wp = seg8_or(pend, lane), we = seg8_or(ev, lane);
        const uint32_t w = (c0 + (uint32_t)j) * 4u + (lane >> 3);
        if ((lane & 7u) == 0u && w < n_words) {
          s_pend[w] = wp;
          s_ev[w] = we;
        }
      }
    }
  }
  __syncthreads();

  uint4 v[1];
  if (windowed) {
    const WindowBits bits{s_pend, s_ev, lo};
    v[0] = group_body<1>(a, bits, 0u, ca, cb, cc, cd, L, bad);
  } else {
    const DirectBits bits{a.pod_state, a.n_pods};
    v[0] = group_body<1>(a, bits, 0u, ca, cb, cc, cd, L, bad);
  }
  pdl_wait_prior();  // group_out may still be read by the previous sweep's LWS pass
  if (valid) emit_if_changed<1>(a.changes, reinterpret_cast<uint4*>(a.out + g), g, v);
}

// --------------------------------------------------------------------------
// LWS-level pass
// --------------------------------------------------------------------------
struct LwsSweepArgs {
  const lwse_lws_rec* lws;
  const lwse_group_out* gout;
  lwse_lws_out* out;
  uint32_t n_lws;
  uint32_t n_groups;
  uint32_t sweep_flags;
  ChangeList changes;
};

__device__ __forceinline__ int32_t want_replicas(int32_t lws_replicas, int32_t surge, int32_t mu,
                                                 int32_t unready, int32_t sts_replicas,
                                                 uint32_t& event) {
  // calculateRollingUpdateReplicas :685-696
  int32_t fin;
  if (unready <= surge) {
    fin = lws_replicas + max(0, unready - mu);
  } else {
    fin = lws_replicas + surge;
  }
  if (fin == sts_replicas - 1)  // :313
    event = LWSE_EVENT_DELETE_ONE;
  else if (fin < sts_replicas)  // :315
    event = LWSE_EVENT_DELETE_RANGE;
  return fin;
}

template <int W>
__global__ void __launch_bounds__(256) lws_sweep_kernel(const LwsSweepArgs a) {
  constexpr uint32_t kTilesPerBlock = 256 / W;
  const uint32_t lane = threadIdx.x & (W - 1);
  const uint32_t n_tiles = gridDim.x * kTilesPerBlock;
  pdl_launch_dependents();
  bool waited = false;
  for (uint32_t i = blockIdx.x * kTilesPerBlock + threadIdx.x / W; i < a.n_lws; i += n_tiles) {
    const uint4* row = reinterpret_cast<const uint4*>(a.lws + i);
    const uint4 r0 = ldg_stream(row + 0), r1 = ldg_stream(row + 1), r2 = ldg_stream(row + 2),
                r3 = ldg_stream(row + 3);
    if (!waited) {  // the object rows are input; the group pass's flag words come next
      pdl_wait_prior();
      waited = true;
    }
    const int32_t size = (int32_t)r0.z;
    const uint32_t lflags = r0.w;
    const int32_t R = (int32_t)r1.x, P = (int32_t)r1.y;
    const int32_t n = (int32_t)r2.x;  // sts replicas
    const int32_t cur_partition = (int32_t)r2.y, annot = (int32_t)r2.z;
    const uint32_t gbase = r3.z, gc = r3.w;

    uint32_t oflags = 0, event = LWSE_EVENT_NONE;
    int32_t o_partition = 0, o_replicas = 0, o_mu = 0, o_ready = 0, o_updated = 0, o_unready = 0,
            o_min_member = 0;

    if ((uint64_t)gbase + gc > a.n_groups) {
      oflags = LWSE_LOUT_BAD_TABLE;
    } else {
      const lwse_group_out* go = a.gout + gbase;
      const bool sts_exists = lflags & LWSE_LWS_STS_EXISTS;
      const bool intstr_bad = lflags & LWSE_LWS_INTSTR_INVALID;
      int32_t mu = scaled_value((int32_t)r1.w, lflags & LWSE_LWS_UNAVAIL_IS_PERCENT, R, false);
      int32_t surge = scaled_value((int32_t)r1.z, lflags & LWSE_LWS_SURGE_IS_PERCENT, R, true);
      if (surge > R) surge = R;  // :307
      const int32_t burst = R + surge;

      // ---- pass A: counters over every group slot of the object ----
      // updateConditions counters (:450-475) and, for slots below the leader
      // sts's replica count, the getReplicaStates view.
      int32_t c_ready = 0, c_updated = 0, c_cur_nb = 0, c_upd_nb = 0, c_ready_nb = 0, c_upd_rdy = 0;
      int32_t c_ok_below_R = 0;  // slots < min(R, n) that are ready ∧ updated
      int32_t max_bad = -1;      // highest slot < n that is not (ready ∧ updated)
      const int32_t n_live = min(n, (int32_t)min(gc, 0x7FFFFFFFu));
      for (uint32_t idx = lane; idx < gc; idx += W) {
        const uint32_t f = __ldg(&go[idx].flags);
        const int32_t ix = (int32_t)idx;
        const bool in_nb = ix < R && ix >= P;
        if (f & LWSE_GOUT_COUNTED) {
          const bool rd = f & LWSE_GOUT_COND_READY, up = f & LWSE_GOUT_COND_UPDATED;
          c_ready += rd;
          c_updated += up;
          c_cur_nb += in_nb;
          c_upd_nb += in_nb && up;
          c_ready_nb += (ix < R) && rd;
          c_upd_rdy += in_nb && rd && up;
        }
        if (ix < n_live) {
          const bool ok = (f & (LWSE_GOUT_STATE_READY | LWSE_GOUT_STATE_UPDATED)) ==
                          (LWSE_GOUT_STATE_READY | LWSE_GOUT_STATE_UPDATED);
          if (ok && ix < R) c_ok_below_R++;
          if (!ok) max_bad = ix;  // idx ascends per lane
        }
      }
      // pack the six small counters pairwise to halve the reductions
      c_ready = tile_add<W>(c_ready);
      c_updated = tile_add<W>(c_updated);
      c_cur_nb = tile_add<W>(c_cur_nb);
      c_upd_nb = tile_add<W>(c_upd_nb);
      c_ready_nb = tile_add<W>(c_ready_nb);
      c_upd_rdy = tile_add<W>(c_upd_rdy);
      c_ok_below_R = tile_add<W>(c_ok_below_R);
      max_bad = tile_max<W>(max_bad);
      if (n > n_live) max_bad = n - 1;  // slots without any object are zero-valued states

      // ---- status / conditions (:478-501) ----
      if (lflags & LWSE_LWS_GROUP_LABEL_INVALID) {
        oflags |= LWSE_LOUT_STATUS_ERROR;  // :434-437
      } else {
        o_ready = c_ready;
        o_updated = c_updated;
        uint32_t cond;
        if (c_upd_nb < c_cur_nb)
          cond = LWSE_COND_UPDATE_IN_PROGRESS;
        else if (c_ready_nb == R && c_upd_rdy == c_cur_nb)
          cond = LWSE_COND_AVAILABLE;
        else
          cond = LWSE_COND_PROGRESSING;
        oflags |= cond << LWSE_LOUT_COND_SHIFT;
        if (P == 0 && c_upd_rdy == R) oflags |= LWSE_LOUT_UPDATE_DONE;
      }

      // ---- rollingUpdateParameters (:280-373) ----
      bool err = false;
      if (!sts_exists) {  // Case 1
        o_partition = 0;
        o_replicas = R;
      } else if (intstr_bad) {
        err = true;
      } else if (lflags & LWSE_LWS_UPDATED) {  // Case 2
        o_partition = min(R, n);
        o_replicas = n < R ? R : want_replicas(R, surge, mu, R, n, event);
      } else if (cur_partition == 0 && n == R) {  // Case 3
        o_partition = 0;
        o_replicas = R;
      } else if (n < R) {
        o_partition = cur_partition;
        o_replicas = R;
      } else {
        // calculateLWSUnreadyReplicas :675-683 (R >= 0 here because n >= R is not implied; clamp)
        const int32_t unready = max(R, 0) - c_ok_below_R;
        o_unready = unready;
        if (!(lflags & LWSE_LWS_ANNOT_VALID)) {
          err = true;  // :351-354
        } else if (annot != R) {  // Case 4
          o_partition = min(cur_partition, burst);
          o_replicas = want_replicas(R, surge, mu, unready, n, event);
        } else {  // Case 5
          const int32_t step = mu + surge - (burst - n);  // :366-369
          // rollingUpdatePartition :643-673
          const int32_t cont_ready = n - 1 - max_bad;  // calculateContinuousReadyReplicas
          const int32_t rsp = max(0, n - cont_ready - step);
          // pass B: unavailable = #{idx < rsp : !ready}
          int32_t unavail = 0;
          const int32_t rsp_live = min(rsp, n_live);
          for (int32_t idx = (int32_t)lane; idx < rsp_live; idx += W)
            unavail += !(__ldg(&go[idx].flags) & LWSE_GOUT_STATE_READY);
          unavail = tile_add<W>(unavail) + (rsp - rsp_live);
          int32_t part = rsp + unavail;
          const int32_t hi = min(part, n - 1);
          if (hi >= rsp) {
            // pass C: the walk stops at the highest idx in [rsp, hi] that is ready ∧ ¬updated
            int32_t blocker = -1;
            const int32_t hi_live = min(hi, n_live - 1);
            for (int32_t idx = rsp + (int32_t)lane; idx <= hi_live; idx += W) {
              const uint32_t f = __ldg(&go[idx].flags);
              if ((f & LWSE_GOUT_STATE_READY) && !(f & LWSE_GOUT_STATE_UPDATED)) blocker = idx;
            }
            blocker = tile_max<W>(blocker);
            if (blocker < 0)
              part = rsp;
            else if (blocker < hi)
              part = blocker + 1;
            // blocker == hi: the first probe breaks, partition keeps rsp + unavail
          }
          o_partition = min(part, cur_partition);
          o_replicas = want_replicas(R, surge, mu, unready, n, event);
        }
      }
      if (err) {
        oflags |= LWSE_LOUT_RUP_ERROR;
        o_partition = 0;
        o_replicas = 0;
        event = LWSE_EVENT_NONE;
      }
      o_partition = max(o_partition, P);  // deferred clamp :285-288
      oflags |= event << LWSE_LOUT_EVENT_SHIFT;

      // stsMaxUnavailable :811-830
      if (!intstr_bad) o_mu = max(1, mu + surge);
      // PodGroup MinMember, volcano_provider.go:72,81-83
      if (a.sweep_flags & LWSE_SWEEP_GANG)
        o_min_member = (lflags & LWSE_LWS_STARTUP_LEADER_READY) ? 1 : size;
      if (lflags & LWSE_LWS_IRREGULAR) oflags |= LWSE_LOUT_IRREGULAR;
    }
    if (lane == 0) {
      const uint4 v[2] = {make_uint4((uint32_t)o_partition, (uint32_t)o_replicas, (uint32_t)o_mu, (uint32_t)o_ready),
                          make_uint4((uint32_t)o_updated, (uint32_t)o_min_member, oflags, (uint32_t)o_unready)};
      emit_if_changed<2>(a.changes, reinterpret_cast<uint4*>(a.out + i), i, v);
    }
  }
}

// --------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------
// LWSE_NO_PDL=1 turns programmatic dependent launch off (A/B measurements).
static const bool g_pdl = [] {
  const char* v = getenv("LWSE_NO_PDL");
  return !(v && v[0] == '1');
}();

static int pick_tile(uint64_t items, uint64_t owners) {
  // Lanes per object: every lane of a tile runs the object-level logic (the five cases, the walk)
  // itself, so lanes only pay off when there are several group words per lane to read — next power
  // of two >= (average groups per object) / kGroupsPerLane, in [1, 32].  One lane per group
  // (divisor 1) made the C5 shape (16 groups per object) instruction-bound: 42 us for 100 k objects.
  static const uint64_t div = [] {
    const char* v = getenv("LWSE_LWS_TILE_DIV");
    const int d = v ? atoi(v) : 8;
    return (uint64_t)(d < 1 ? 1 : d);
  }();
  if (owners == 0) return 1;
  const uint64_t avg = ((items + owners - 1) / owners + div - 1) / div;
  int w = 1;
  while (w < 32 && (uint64_t)w < avg) w <<= 1;
  return w;
}

// Persistent-style grids: as many CTAs as are resident at once (SM count x
// occupancy), each thread/tile striding over its share of the rows.
template <typename K>
static uint32_t resident_ctas(K kernel, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1)
    per_sm = 4;
  return (uint32_t)sm_count * (uint32_t)per_sm;
}

template <int W>
static cudaError_t launch_group(const GroupSweepArgs& a, int sm_count, cudaStream_t s) {
  constexpr uint32_t kTilesPerBlock = kGroupThreads / W;
  (void)sm_count;
  const uint32_t want = (a.n_groups + kTilesPerBlock - 1) / kTilesPerBlock;
  return launch_pdl(group_sweep_kernel<W>, dim3(want < (1u << 20) ? want : (1u << 20)), dim3(kGroupThreads), 0, s,
                    g_pdl, a);
}

template <int W>
static cudaError_t launch_lws(const LwsSweepArgs& a, int sm_count, cudaStream_t s) {
  constexpr uint32_t kTilesPerBlock = 256 / W;
  static uint32_t resident = 0;
  if (resident == 0) resident = resident_ctas(lws_sweep_kernel<W>, sm_count);
  const uint32_t want = (a.n_lws + kTilesPerBlock - 1) / kTilesPerBlock;
  return launch_pdl(lws_sweep_kernel<W>, dim3(want < resident ? want : resident), dim3(256), 0, s, g_pdl, a);
}

constexpr int kScanUnroll = 4;  // 128-bit loads per lane per chunk: 4 x 512 B = 2 KB in flight per warp

size_t lws_sweep_scratch_bytes(uint64_t n_pods) {
  const uint64_t words = (n_pods + 31u) / 32u;
  return (size_t)(2 * words * sizeof(uint32_t) + 256);
}

// Returns the number of kernels launched (>=0) or -1 with *cuda_err set.
// scratch: lws_sweep_scratch_bytes(n_pods) bytes of device memory.
struct SweepChangeLists {  // device pointers; all null = off
  uint32_t* lws_rows = nullptr;
  lwse_lws_out* lws_out = nullptr;
  uint32_t lws_capacity = 0;
  uint32_t* group_rows = nullptr;
  lwse_group_out* group_out = nullptr;
  uint32_t group_capacity = 0;
  uint32_t* counts = nullptr;  // [0] lws, [1] groups; zeroed by the caller
};

int launch_lws_sweep(const lwse_lws_tables* t, const lwse_node_rec* d_nodes, uint32_t n_nodes,
                     void* scratch, int sm_count, cudaStream_t s, int* cuda_err, const SweepChangeLists* cl,
                     uint32_t* d_event_count) {
  *cuda_err = 0;
  int launches = 0;
  cudaError_t e = cudaSuccess;
  const uint64_t words = (t->n_pods + 31u) / 32u;
  uint32_t* pending_bits = static_cast<uint32_t*>(scratch);
  uint32_t* event_bits = pending_bits + ((words + 31u) & ~(uint64_t)31u);
  const bool group_pass = t->n_groups && !(t->flags & LWSE_SWEEP_SKIP_GROUP_PASS);

  // lanes per group: one per bitmap word of an average group, in {1, 2, 4, 8}
  const uint64_t avg_pods = t->n_groups ? (t->n_pods + t->n_groups - 1) / t->n_groups : 0;
  const int w = avg_pods > 2048 ? 8 : avg_pods > 1024 ? 4 : avg_pods > 512 ? 2 : 1;
  static const bool no_fuse = [] {
    const char* v = getenv("LWSE_NO_FUSE");
    return v && atoi(v) != 0;
  }();
  // scan + group pass in one kernel: small groups (the pod window of 256 consecutive groups has to
  // fit the CTA's bitmap window of 65 536 pods), no occupancy count, both passes wanted.
  // (A variant that staged the window's state words in shared memory with one TMA bulk copy per CTA
  // — 128 groups, 32 KB — and evaluated the predicates from there was slower: 11.7 us against 9.6 us;
  // the register scan already keeps ~48 KB in flight per SM, and a single bulk copy per CTA does not.)
  const bool fused = group_pass && avg_pods <= 256 && !t->node_occupancy && !no_fuse && !d_event_count &&
                     !(t->flags & LWSE_SWEEP_SKIP_POD_SCAN);
  if (fused) {
    GroupSweepArgs a{t->lws,   t->groups, t->pod_state, t->pod_ident, nullptr, nullptr, d_nodes,
                     t->group_out, t->n_pods, t->n_lws, t->n_groups, n_nodes, t->flags, ChangeList{}};
    if (cl && cl->group_rows) a.changes = ChangeList{cl->group_rows, cl->group_out, cl->counts + 1, cl->group_capacity};
    const uint32_t grid = (t->n_groups + kFusedThreads - 1) / kFusedThreads;
    e = launch_pdl(group_fused_kernel, dim3(grid), dim3(kFusedThreads), 0, s, g_pdl, a);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    launches++;
  }
  if (!fused && t->node_occupancy && n_nodes && !(t->flags & LWSE_SWEEP_SKIP_POD_SCAN)) {
    e = cudaMemsetAsync(t->node_occupancy, 0, (size_t)n_nodes * sizeof(uint32_t), s);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
  }
  if (!fused && t->n_pods && (t->n_groups || t->node_occupancy) && !(t->flags & LWSE_SWEEP_SKIP_POD_SCAN)) {
    PodScanArgs a{t->pod_state, pending_bits, event_bits, t->node_occupancy, t->n_pods, n_nodes, d_event_count};
    // one chunk per warp: the kernel is a few microseconds long, so let the
    // hardware CTA scheduler balance it instead of a persistent grid-stride loop
    const uint64_t chunks = (t->n_pods + 128ull * kScanUnroll - 1) / (128ull * kScanUnroll);
    const uint64_t want = (chunks + 7) / 8;
    const uint32_t grid = (uint32_t)(want < (1u << 20) ? want : (1u << 20));
    if (t->node_occupancy)
      e = launch_pdl(pod_scan_kernel<kScanUnroll, true>, dim3(grid), dim3(256), 0, s, false, a);  // follows a memset
    else
      e = launch_pdl(pod_scan_kernel<kScanUnroll, false>, dim3(grid), dim3(256), 0, s, g_pdl, a);
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    launches++;
  }
  if (group_pass && !fused) {
    GroupSweepArgs a{t->lws,   t->groups, t->pod_state, t->pod_ident, pending_bits, event_bits, d_nodes,
                     t->group_out, t->n_pods, t->n_lws, t->n_groups, n_nodes, t->flags, ChangeList{}};
    if (cl && cl->group_rows) a.changes = ChangeList{cl->group_rows, cl->group_out, cl->counts + 1, cl->group_capacity};
    switch (w) {
      case 8: e = launch_group<8>(a, sm_count, s); break;
      case 4: e = launch_group<4>(a, sm_count, s); break;
      case 2: e = launch_group<2>(a, sm_count, s); break;
      default: e = launch_group<1>(a, sm_count, s); break;
    }
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    launches++;
  }
  if (t->n_lws && !(t->flags & LWSE_SWEEP_SKIP_LWS_PASS)) {
    LwsSweepArgs a{t->lws, t->group_out, t->lws_out, t->n_lws, t->n_groups, t->flags, ChangeList{}};
    if (cl && cl->lws_rows) a.changes = ChangeList{cl->lws_rows, cl->lws_out, cl->counts + 0, cl->lws_capacity};
    switch (pick_tile(t->n_groups, t->n_lws)) {
      case 1: e = launch_lws<1>(a, sm_count, s); break;
      case 2: e = launch_lws<2>(a, sm_count, s); break;
      case 4: e = launch_lws<4>(a, sm_count, s); break;
      case 8: e = launch_lws<8>(a, sm_count, s); break;
      case 16: e = launch_lws<16>(a, sm_count, s); break;
      default: e = launch_lws<32>(a, sm_count, s); break;
    }
    if (e != cudaSuccess) { *cuda_err = (int)e; return -1; }
    launches++;
  }
  return launches;
}

// --------------------------------------------------------------------------
// row patches for the resident tables
// --------------------------------------------------------------------------
template <int WORDS>  // 32-bit words per row
__global__ void __launch_bounds__(256) scatter_rows_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ rows,
                                                           const uint32_t* __restrict__ values, uint32_t n,
                                                           uint64_t table_rows) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n * WORDS) return;
  const uint32_t k = (uint32_t)(i / WORDS), w = (uint32_t)(i % WORDS);
  const uint32_t r = __ldg(rows + k);
  if (r < table_rows) table[(uint64_t)r * WORDS + w] = __ldg(values + i);
}

int launch_scatter(int row_words, void* table, uint64_t table_rows, const uint32_t* rows, const void* values,
                   uint32_t n, cudaStream_t s, int* cuda_err) {
  *cuda_err = 0;
  if (n == 0) return 0;
  const uint64_t threads = (uint64_t)n * (uint64_t)row_words;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  uint32_t* tb = static_cast<uint32_t*>(table);
  const uint32_t* v = static_cast<const uint32_t*>(values);
  switch (row_words) {
    case 16: scatter_rows_kernel<16><<<grid, 256, 0, s>>>(tb, rows, v, n, table_rows); break;
    case 3: scatter_rows_kernel<3><<<grid, 256, 0, s>>>(tb, rows, v, n, table_rows); break;
    case 1: scatter_rows_kernel<1><<<grid, 256, 0, s>>>(tb, rows, v, n, table_rows); break;
    default: *cuda_err = (int)cudaErrorInvalidValue; return -1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *cuda_err = (int)e;
    return -1;
  }
  return 1;
}

}  // namespace lwse
