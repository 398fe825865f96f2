/*
 * lwse_oracle_place.c — sequential statement of the placement SPEC.
 *
 * TEST INFRASTRUCTURE ONLY (see lwse_oracle.c).
 *
 * PARITY UNPINNED: the reference has no node scoring.  It only emits the
 * constraints — exclusive pod affinity / anti-affinity per topology domain
 * (pkg/webhooks/pod_webhook.go:185-227, namespace-scoped: no Namespaces field
 * on the terms), the workers' nodeSelector copied from the leader's node
 * (pkg/controllers/pod_controller.go:297-336) and the PodGroup
 * (pkg/schedulerprovider/volcano_provider.go:58-87); kube-scheduler / Volcano,
 * whose sources are not part of the reference, do the actual filtering and
 * scoring.  What follows is therefore this build's own specification of a
 * placement round that honours those constraints, written sequentially; the
 * CUDA kernels must reproduce it bit for bit (DESIGN.md "Placement").
 *
 * Spec
 *   free[n]      = schedulable ∧ has-topology ? max(0, capacity[n] − occupancy[n]) : 0
 *   dom_free[d]  = Σ free[n] over the nodes of domain d
 *   key(r)       = unpinned(r) << 63 | (priority(r) >> 25) << 24 | index(r)     (smaller wins)
 *   pinned r     (leader already scheduled on a node with a topology label) claims
 *                its node's domain in its namespace; the smallest key among the
 *                pinned claimants of a (namespace, domain) holds it (PLACED|PINNED),
 *                the others get PINNED|CONFLICT.  Pinned claims are facts: they are
 *                never displaced by unpinned requests (bit 63 of the key).
 *   unpinned r   in ascending key order picks, in two levels,
 *                  1. the domain d with   dom_free[d] ≥ size(r) ∧ (ns(r), d) not held   that has
 *                     the largest      hi = mix(key_lo ^ d·φ) | 1        (ties → lower d)
 *                     — a per-group rendezvous hash: any capacity-derived term would give all
 *                     groups the same preference order and serialise the claim rounds, so
 *                     capacity is a feasibility filter only;
 *                  2. in it the node n with free[n] ≥ 1 that has the largest
 *                                      lo = min(free[n], 15) << 28 | mix(key_hi ^ n·ψ) >> 4
 *                     — emptiest node first                              (ties → lower n)
 *                and then holds (ns, d); score = hi.  No feasible domain → UNSCHEDULABLE.
 *                (size ≥ 1, so a feasible domain always has a node with a free slot.)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lwse.h"

#define LWSO_API __attribute__((visibility("default")))

static uint32_t mix32(uint32_t x) { /* murmur3 finalizer */
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

uint64_t lwso_place_key(const lwse_place_req* r, uint32_t index, int pinned) {
  return ((uint64_t)(pinned ? 0 : 1) << 63) | (((r->priority >> 25) & 0x7FFFFFFFFFull) << 24) |
         (uint64_t)(index & 0xFFFFFFu);
}

typedef struct {
  uint64_t key;
  uint32_t idx;
} order_ent;

static int cmp_order(const void* a, const void* b) {
  uint64_t x = ((const order_ent*)a)->key, y = ((const order_ent*)b)->key;
  return x < y ? -1 : x > y ? 1 : 0;
}

/* Exclusivity is per namespace and capacity is a snapshot (no claim consumes it), so the
 * namespaces are independent sub-problems: solving them one after the other in ascending key
 * order (threads = 1) or side by side on several host threads gives the same rows. */
typedef struct {
  const lwse_node_rec* nodes;
  const uint32_t* free_;
  const uint32_t* dom_free;
  uint32_t n_nodes, n_domains;
  const lwse_place_req* reqs;
  lwse_place_out* out;
  const uint32_t* dom_first; /* usable nodes grouped by domain: dom_nodes[dom_first[d] .. dom_first[d+1]) */
  const uint32_t* dom_nodes; /* ascending node index inside a domain */
  const order_ent* items;    /* requests bucketed by namespace, ascending key inside a bucket */
  const uint32_t* ns_first;  /* n_namespaces + 1 */
  uint32_t n_namespaces;
  int thread, threads;
} place_job;

static void solve_namespace(const place_job* j, uint32_t ns, uint64_t* hold) {
  const uint32_t n_domains = j->n_domains, n_nodes = j->n_nodes;
  const lwse_node_rec* nodes = j->nodes;
  lwse_place_out* out = j->out;
  memset(hold, 0xFF, sizeof(uint64_t) * ((size_t)n_domains + 1));
  for (uint32_t k = j->ns_first[ns]; k < j->ns_first[ns + 1]; k++) {
    const uint32_t i = j->items[k].idx;
    const lwse_place_req* r = &j->reqs[i];
    if (r->leader_node != LWSE_NONE) {
      /* pinned: the leader's node decides */
      if (r->leader_node >= n_nodes) continue; /* node object missing: no claim */
      const lwse_node_rec* nd = &nodes[r->leader_node];
      if (!(nd->flags & LWSE_NODE_HAS_TOPOLOGY) || nd->domain_id >= n_domains) continue;
      out[i].domain_id = nd->domain_id;
      if (hold[nd->domain_id] == ~0ull) {
        hold[nd->domain_id] = j->items[k].key;
        out[i].flags |= LWSE_PLACE_PLACED;
      } else {
        out[i].flags |= LWSE_PLACE_CONFLICT;
      }
      continue;
    }
    if (r->size < 1) {
      out[i].flags |= LWSE_PLACE_UNSCHEDULABLE;
      continue;
    }
    const uint32_t key_lo = (uint32_t)r->group_key, key_hi = (uint32_t)(r->group_key >> 32);
    /* level 1: the domain */
    uint32_t best_hi = 0, d = LWSE_NONE;
    for (uint32_t c = 0; c < n_domains; c++) {
      if (hold[c] != ~0ull || j->dom_free[c] < (uint32_t)r->size) continue;
      uint32_t hi = mix32(key_lo ^ (c * 0x9E3779B1u)) | 1u;
      if (hi > best_hi) { /* ties keep the lower domain index */
        best_hi = hi;
        d = c;
      }
    }
    if (d == LWSE_NONE) {
      out[i].flags |= LWSE_PLACE_UNSCHEDULABLE;
      continue;
    }
    /* level 2: the node in it */
    uint32_t best_lo = 0, best_n = LWSE_NONE;
    for (uint32_t q = j->dom_first[d]; q < j->dom_first[d + 1]; q++) { /* the nodes of domain d, ascending */
      const uint32_t n = j->dom_nodes[q];
      if (j->free_[n] < 1) continue;
      uint32_t lo = ((j->free_[n] > 15 ? 15u : j->free_[n]) << 28) | (mix32(key_hi ^ (n * 0x85EBCA77u)) >> 4);
      if (lo > best_lo) { /* ties keep the lower node index */
        best_lo = lo;
        best_n = n;
      }
    }
    if (best_n == LWSE_NONE) { /* cannot happen: dom_free[d] >= size >= 1 */
      out[i].flags |= LWSE_PLACE_UNSCHEDULABLE;
      continue;
    }
    hold[d] = j->items[k].key;
    out[i].domain_id = d;
    out[i].leader_node = best_n;
    out[i].flags |= LWSE_PLACE_PLACED;
    out[i].score = best_hi;
  }
}

static void* place_worker(void* arg) {
  const place_job* j = (const place_job*)arg;
  uint64_t* hold = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)j->n_domains + 1));
  for (uint32_t ns = (uint32_t)j->thread; ns < j->n_namespaces; ns += (uint32_t)j->threads)
    if (j->ns_first[ns + 1] > j->ns_first[ns]) solve_namespace(j, ns, hold);
  free(hold);
  return NULL;
}

LWSO_API int lwso_place_mt(const lwse_node_rec* nodes, uint32_t n_nodes, const uint32_t* occupancy,
                           uint32_t n_domains, uint32_t n_namespaces, const lwse_place_req* reqs,
                           uint32_t n_reqs, lwse_place_out* out, int threads) {
  if (n_reqs > 0xFFFFFFu) return LWSE_ERR_UNSUPPORTED;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  uint32_t* free_ = (uint32_t*)calloc(n_nodes ? n_nodes : 1, sizeof(uint32_t));
  uint32_t* dom_free = (uint32_t*)calloc(n_domains ? n_domains : 1, sizeof(uint32_t));
  order_ent* order = (order_ent*)malloc(sizeof(order_ent) * (n_reqs ? n_reqs : 1));
  order_ent* items = (order_ent*)malloc(sizeof(order_ent) * (n_reqs ? n_reqs : 1));
  uint32_t* ns_first = (uint32_t*)calloc((size_t)n_namespaces + 2, sizeof(uint32_t));
  uint32_t* dom_first = (uint32_t*)calloc((size_t)n_domains + 2, sizeof(uint32_t));
  uint32_t* dom_nodes = (uint32_t*)malloc(sizeof(uint32_t) * (n_nodes ? n_nodes : 1));
  for (uint32_t n = 0; n < n_nodes; n++) {
    const lwse_node_rec* nd = &nodes[n];
    uint32_t occ = occupancy ? occupancy[n] : 0;
    int usable = (nd->flags & LWSE_NODE_SCHEDULABLE) && (nd->flags & LWSE_NODE_HAS_TOPOLOGY) &&
                 nd->domain_id < n_domains;
    free_[n] = usable && nd->capacity > occ ? nd->capacity - occ : 0;
    if (usable) {
      dom_free[nd->domain_id] += free_[n];
      dom_first[nd->domain_id + 1]++;
    }
  }
  for (uint32_t d = 0; d < n_domains; d++) dom_first[d + 1] += dom_first[d];
  {
    uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_domains + 1));
    memcpy(cur, dom_first, sizeof(uint32_t) * ((size_t)n_domains + 1));
    for (uint32_t n = 0; n < n_nodes; n++) {
      const lwse_node_rec* nd = &nodes[n];
      if ((nd->flags & LWSE_NODE_SCHEDULABLE) && (nd->flags & LWSE_NODE_HAS_TOPOLOGY) && nd->domain_id < n_domains)
        dom_nodes[cur[nd->domain_id]++] = n;
    }
    free(cur);
  }
  /* classify, default outputs */
  for (uint32_t i = 0; i < n_reqs; i++) {
    const lwse_place_req* r = &reqs[i];
    int pinned = r->leader_node != LWSE_NONE;
    order[i].key = lwso_place_key(r, i, pinned);
    order[i].idx = i;
    out[i].domain_id = LWSE_NONE;
    out[i].leader_node = pinned ? r->leader_node : LWSE_NONE;
    out[i].flags = pinned ? LWSE_PLACE_PINNED : 0;
    out[i].score = 0;
    if (r->ns >= n_namespaces)
      out[i].flags |= LWSE_PLACE_UNSCHEDULABLE;
    else
      ns_first[r->ns + 1]++;
  }
  qsort(order, n_reqs, sizeof(order_ent), cmp_order);
  for (uint32_t ns = 0; ns < n_namespaces; ns++) ns_first[ns + 1] += ns_first[ns];
  {
    uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)n_namespaces + 1));
    memcpy(cursor, ns_first, sizeof(uint32_t) * ((size_t)n_namespaces + 1));
    for (uint32_t k = 0; k < n_reqs; k++) { /* ascending key order is kept inside every bucket */
      const uint32_t ns = reqs[order[k].idx].ns;
      if (ns < n_namespaces) items[cursor[ns]++] = order[k];
    }
    free(cursor);
  }
  place_job jobs[256];
  pthread_t tid[256];
  for (int k = 0; k < threads; k++)
    jobs[k] = (place_job){nodes, free_, dom_free, n_nodes, n_domains, reqs, out, dom_first, dom_nodes, items, ns_first,
                          n_namespaces, k, threads};
  if (threads == 1) {
    place_worker(&jobs[0]);
  } else {
    for (int k = 0; k < threads; k++) pthread_create(&tid[k], NULL, place_worker, &jobs[k]);
    for (int k = 0; k < threads; k++) pthread_join(tid[k], NULL);
  }
  free(free_);
  free(dom_free);
  free(dom_first);
  free(dom_nodes);
  free(order);
  free(items);
  free(ns_first);
  return LWSE_OK;
}

LWSO_API int lwso_place(const lwse_node_rec* nodes, uint32_t n_nodes, const uint32_t* occupancy,
                        uint32_t n_domains, uint32_t n_namespaces, const lwse_place_req* reqs,
                        uint32_t n_reqs, lwse_place_out* out) {
  return lwso_place_mt(nodes, n_nodes, occupancy, n_domains, n_namespaces, reqs, n_reqs, out, 1);
}
