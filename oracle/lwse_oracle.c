/*
 * lwse_oracle.c — CPU restatement of the reference's reconcile arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA engine: it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product
 * (lws_b200/, liblwse.so) never links or calls it; there is no CPU fallback.
 *
 * It restates, function by function, the Go code of kubernetes-sigs/lws
 * @ 1d9204a2 over the fixed-width record tables of include/lwse.h (the
 * reference cannot be compiled here: no Go toolchain).  Each
 * function cites the reference lines it follows.  It is written the way the
 * reference computes — one object at a time, materialising the per-replica
 * state slice — not the way the GPU kernels do, so that agreement between the
 * two is evidence and not tautology.
 *
 * Parity pinning: tests/test_oracle_golden.py drives this file with every
 * transcribable vector of the reference's own tests.  The
 * placement spec oracle (lwso_place) is build-defined — the reference has no
 * node scoring — and is labelled "parity unpinned".
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lwse.h"

#define LWSO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* third-party arithmetic                                                    */
/* ------------------------------------------------------------------------- */

/* k8s.io/apimachinery v0.36.1 util/intstr.GetScaledValueFromIntOrPercent:
 * Int → IntVal; "N%" → int(math.Ceil|Floor(float64(N) * float64(total) / 100)).
 * Kept in float64 exactly like the Go source. */
static int scaled_value(int32_t val, int is_percent, int total, int round_up) {
  if (!is_percent) return val;
  double v = (double)val * (double)total / 100.0;
  return round_up ? (int)ceil(v) : (int)floor(v);
}

/* pkg/utils/utils.go:45-50 NonZeroValue */
static int32_t non_zero_value(int32_t v) { return v < 0 ? 0 : v; }

static int32_t min32(int32_t a, int32_t b) { return a < b ? a : b; }
static int32_t max32(int32_t a, int32_t b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------- */
/* leaderworkerset_controller.go                                             */
/* ------------------------------------------------------------------------- */

typedef struct {
  int ready;
  int updated;
} replica_state; /* :569-574 */

/* pkg/utils/pod/pod_utils.go:58-60 PodRunningAndReady */
static int pod_running_and_ready(const lwse_group_rec* g) {
  return (g->flags & LWSE_GRP_POD_RUNNING) && (g->flags & LWSE_GRP_POD_READY);
}

/* pkg/utils/statefulset/statefulset_utils.go:48-51 StatefulsetReady */
static int statefulset_ready(const lwse_group_rec* g) {
  return g->wsts_spec_replicas == g->wsts_avail_replicas && (g->flags & LWSE_GRP_WSTS_REV_SETTLED);
}

/* :576-641 getReplicaStates.  The two List+SortByIndex passes are the
 * encoder's job (row r = group index r; a missing / mis-named object clears
 * the NAME_MATCH flags, which is how a zero-valued slot shows up at :609). */
static void get_replica_states(const lwse_lws_rec* l, const lwse_group_rec* groups,
                               int32_t sts_replicas, replica_state* states) {
  int no_worker_sts = l->size == 1;
  for (int32_t idx = 0; idx < sts_replicas; idx++) {
    states[idx].ready = 0;
    states[idx].updated = 0;
    if ((uint32_t)idx >= l->group_count) continue; /* nothing with this index exists */
    const lwse_group_rec* g = &groups[l->group_base + idx];
    /* :609 nominatedName != sortedPods[idx].Name || (!noWorkerSts && nominatedName != sortedSts[idx].Name) */
    if (!(g->flags & LWSE_GRP_POD_NAME_MATCH) ||
        (!no_worker_sts && !(g->flags & LWSE_GRP_WSTS_LABEL_NAME_MATCH))) {
      continue;
    }
    int leader_updated = g->leader_rev_hash == l->rev_hash; /* :618 */
    int leader_ready = pod_running_and_ready(g);             /* :619 */
    if (no_worker_sts) {
      states[idx].ready = leader_ready;
      states[idx].updated = leader_updated;
      continue;
    }
    int workers_updated = g->wsts_rev_hash == l->rev_hash; /* :629 */
    int workers_ready = statefulset_ready(g);               /* :630 */
    states[idx].ready = leader_ready && workers_ready;
    states[idx].updated = leader_updated && workers_updated;
  }
}

/* :698-708 */
static int32_t calculate_continuous_ready_replicas(const replica_state* states, int32_t n) {
  int32_t count = 0;
  for (int32_t idx = n - 1; idx >= 0; idx--) {
    if (!states[idx].ready || !states[idx].updated) break;
    count++;
  }
  return count;
}

/* :643-673 */
static int32_t rolling_update_partition(const replica_state* states, int32_t sts_replicas,
                                        int32_t rolling_step, int32_t current_partition) {
  int32_t continuous_ready = calculate_continuous_ready_replicas(states, sts_replicas);
  int32_t rolling_step_partition = non_zero_value(sts_replicas - continuous_ready - rolling_step);
  int32_t unavailable = 0;
  for (int32_t idx = 0; idx < rolling_step_partition; idx++) {
    if (!states[idx].ready) unavailable++;
  }
  int32_t partition = rolling_step_partition + unavailable;
  for (int32_t idx = min32(partition, sts_replicas - 1); idx >= rolling_step_partition; idx--) {
    if (!states[idx].ready || states[idx].updated) {
      partition = idx;
    } else {
      break;
    }
  }
  return min32(partition, current_partition);
}

/* :675-683 */
static int32_t calculate_lws_unready_replicas(const replica_state* states, int32_t len,
                                              int32_t lws_replicas) {
  int32_t unready = 0;
  for (int32_t idx = 0; idx < lws_replicas; idx++) {
    if (idx >= len || !states[idx].ready || !states[idx].updated) unready++;
  }
  return unready;
}

/* :685-696 */
LWSO_API int32_t lwso_calculate_rolling_update_replicas(int32_t lws_replicas, int32_t max_surge,
                                                        int32_t max_unavailable,
                                                        int32_t unready_replicas) {
  int32_t burst = lws_replicas + max_surge;
  if (unready_replicas <= max_surge) {
    int32_t required = non_zero_value(unready_replicas - max_unavailable);
    return lws_replicas + required;
  }
  return burst;
}

/* :280-373 rollingUpdateParameters.  Returns 0 or -1 (err). */
static int rolling_update_parameters(const lwse_lws_rec* l, const lwse_group_rec* groups,
                                     int32_t* out_partition, int32_t* out_replicas,
                                     uint32_t* out_event, int32_t* out_unready) {
  int32_t lws_replicas = l->replicas;
  int32_t sts_partition = 0, replicas = 0;
  int err = 0;
  *out_event = LWSE_EVENT_NONE;
  *out_unready = 0;

  do {
    /* Case 1 :293 */
    if (!(l->flags & LWSE_LWS_STS_EXISTS)) {
      sts_partition = 0;
      replicas = lws_replicas;
      break;
    }
    int32_t sts_replicas = l->sts_replicas;
    /* :298-305 */
    if (l->flags & LWSE_LWS_INTSTR_INVALID) {
      err = -1;
      break;
    }
    int max_surge =
        scaled_value(l->max_surge, l->flags & LWSE_LWS_SURGE_IS_PERCENT, (int)lws_replicas, 1);
    int max_unavailable = scaled_value(l->max_unavailable, l->flags & LWSE_LWS_UNAVAIL_IS_PERCENT,
                                       (int)lws_replicas, 0);
    if (max_surge > (int)lws_replicas) max_surge = (int)lws_replicas; /* :307 */
    int32_t burst_replicas = lws_replicas + (int32_t)max_surge;

#define WANT_REPLICAS(unready, dst)                                                     \
  do {                                                                                  \
    int32_t final_ = lwso_calculate_rolling_update_replicas(                            \
        lws_replicas, (int32_t)max_surge, (int32_t)max_unavailable, (unready));         \
    if (final_ == sts_replicas - 1) /* :313 */                                          \
      *out_event = LWSE_EVENT_DELETE_ONE;                                               \
    else if (final_ < sts_replicas) /* :315 */                                          \
      *out_event = LWSE_EVENT_DELETE_RANGE;                                             \
    (dst) = final_;                                                                     \
  } while (0)

    /* Case 2 :325 */
    if (l->flags & LWSE_LWS_UPDATED) {
      int32_t partition = min32(lws_replicas, sts_replicas);
      sts_partition = partition;
      if (sts_replicas < lws_replicas) {
        replicas = lws_replicas;
        break;
      }
      WANT_REPLICAS(lws_replicas, replicas);
      break;
    }

    int32_t partition = l->sts_partition;
    /* Case 3 :335-343 */
    if (partition == 0 && sts_replicas == lws_replicas) {
      sts_partition = 0;
      replicas = lws_replicas;
      break;
    }
    if (sts_replicas < lws_replicas) {
      sts_partition = partition;
      replicas = lws_replicas;
      break;
    }

    /* :345-349 */
    replica_state* states =
        (replica_state*)malloc(sizeof(replica_state) * (size_t)(sts_replicas > 0 ? sts_replicas : 1));
    get_replica_states(l, groups, sts_replicas, states);
    int32_t lws_unready = calculate_lws_unready_replicas(states, sts_replicas, lws_replicas);
    *out_unready = lws_unready;

    /* :351-354 */
    if (!(l->flags & LWSE_LWS_ANNOT_VALID)) {
      free(states);
      err = -1;
      break;
    }
    int replicas_updated = (int)l->sts_replicas_annotation != (int)lws_replicas;
    /* Case 4 :358-361 */
    if (replicas_updated) {
      sts_partition = min32(partition, burst_replicas);
      WANT_REPLICAS(lws_unready, replicas);
      free(states);
      break;
    }
    /* Case 5 :366-372 */
    int rolling_step = max_unavailable;
    rolling_step += max_surge - ((int)burst_replicas - (int)sts_replicas);
    sts_partition = rolling_update_partition(states, sts_replicas, (int32_t)rolling_step, partition);
    WANT_REPLICAS(lws_unready, replicas);
    free(states);
  } while (0);
#undef WANT_REPLICAS

  if (err) {
    sts_partition = 0;
    replicas = 0;
    *out_event = LWSE_EVENT_NONE;
  }
  /* deferred clamp :285-288 — runs on every return, error paths included */
  sts_partition = max32(sts_partition, l->partition);
  *out_partition = sts_partition;
  *out_replicas = replicas;
  return err;
}

/* :811-830 */
static int32_t sts_max_unavailable(const lwse_lws_rec* l) {
  int lws_replicas = (int)l->replicas;
  int mu = scaled_value(l->max_unavailable, l->flags & LWSE_LWS_UNAVAIL_IS_PERCENT, lws_replicas, 0);
  int ms = scaled_value(l->max_surge, l->flags & LWSE_LWS_SURGE_IS_PERCENT, lws_replicas, 1);
  if (ms > lws_replicas) ms = lws_replicas;
  int32_t v = (int32_t)(mu + ms);
  if (v < 1) v = 1;
  return v;
}

/* :414-509 updateConditions counters + condition choice. */
static int update_conditions(const lwse_lws_rec* l, const lwse_group_rec* groups,
                             lwse_lws_out* o) {
  if (l->flags & LWSE_LWS_GROUP_LABEL_INVALID) return -1; /* :434-437 */
  int ready_count = 0, updated_count = 0, ready_non_burst = 0;
  int part_updated_nb = 0, part_current_nb = 0, part_updated_ready = 0;
  int no_worker_sts = l->size == 1;
  int lws_partition = (int)l->partition;
  int lws_replicas = (int)l->replicas;
  for (uint32_t r = 0; r < l->group_count; r++) {
    const lwse_group_rec* g = &groups[l->group_base + r];
    if (!(g->flags & LWSE_GRP_POD_PRESENT)) continue; /* only existing leader pods are listed */
    int index = (int)r;
    if (!no_worker_sts && !(g->flags & LWSE_GRP_WSTS_FOUND)) continue; /* :441-447 */
    if (index < lws_replicas && index >= lws_partition) part_current_nb++;
    int ready = 0, updated = 0;
    if ((no_worker_sts || statefulset_ready(g)) && pod_running_and_ready(g)) {
      ready = 1;
      ready_count++;
    }
    if ((no_worker_sts || g->wsts_rev_hash == l->rev_hash) && g->leader_rev_hash == l->rev_hash) {
      updated = 1;
      updated_count++;
      if (index < lws_replicas && index >= lws_partition) part_updated_nb++;
    }
    if (index < lws_replicas) {
      if (ready) ready_non_burst++;
      if (index >= lws_partition && ready && updated) part_updated_ready++;
    }
  }
  o->ready_replicas = ready_count;
  o->updated_replicas = updated_count;
  uint32_t cond;
  if (part_updated_nb < part_current_nb) {
    cond = LWSE_COND_UPDATE_IN_PROGRESS;
  } else if (ready_non_burst == lws_replicas && part_updated_ready == part_current_nb) {
    cond = LWSE_COND_AVAILABLE;
  } else {
    cond = LWSE_COND_PROGRESSING;
  }
  o->flags |= cond << LWSE_LOUT_COND_SHIFT;
  if (lws_partition == 0 && part_updated_ready == lws_replicas) o->flags |= LWSE_LOUT_UPDATE_DONE;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* pod_controller.go                                                         */
/* ------------------------------------------------------------------------- */

/* one pod, assembled from the state and identity columns */
typedef struct {
  uint64_t rev_hash;
  uint32_t owner_uid_hash;
  uint32_t bits;  /* the state byte */
  uint32_t place; /* the identity row's cold word: name check, node binding */
} pod_view;

typedef struct {
  const lwse_pod_state* state;
  const lwse_pod_ident* ident;
} pod_cols;

static pod_view get_pod(const pod_cols* c, uint64_t i) {
  pod_view p;
  p.rev_hash = c->ident[i].rev_hash;
  p.owner_uid_hash = c->ident[i].owner_uid_hash;
  p.bits = c->state[i];
  p.place = c->ident[i].place;
  return p;
}

/* pkg/utils/pod/pod_utils.go:29-45 ContainerRestarted */
static int container_restarted(const pod_view* p) {
  uint32_t phase = p->bits & LWSE_POD_PHASE_MASK;
  if (phase == LWSE_POD_PHASE_RUNNING || phase == LWSE_POD_PHASE_PENDING)
    return (p->bits & LWSE_POD_ANY_RESTART) != 0;
  return 0;
}

/* :338-362 pendingPodsInGroup */
static int pending_pods_in_group(const lwse_group_rec* g, const pod_cols* pods, int group_size) {
  if ((uint32_t)group_size != g->pod_count) return 1;
  for (uint32_t i = 0; i < g->pod_count; i++) {
    if ((pods->state[g->pod_base + i] & LWSE_POD_PHASE_MASK) == LWSE_POD_PHASE_PENDING) return 1;
  }
  return 0;
}

/* :268-295 workerPodBelongsToLeader */
static int worker_pod_belongs_to_leader(const pod_view* p, const lwse_group_rec* g) {
  uint32_t kind = (p->bits & LWSE_POD_OWNER_MASK) >> LWSE_POD_OWNER_SHIFT;
  if (kind == 0) return 0; /* owner == nil */
  if (kind == 1) /* Pod */
    return (p->bits & LWSE_POD_OWNER_NAME_MATCH) && p->owner_uid_hash == g->leader_uid_hash;
  if (kind != 2) return 0;
  /* Get(sts named owner.Name): the encoder only resolves the group's own sts */
  if (!(p->bits & LWSE_POD_OWNER_NAME_MATCH) || !(g->flags & LWSE_GRP_WSTS_FOUND)) return 0;
  if (g->wsts_uid_hash != p->owner_uid_hash) return 0;
  if (!(g->flags & LWSE_GRP_WSTS_OWNER_IS_POD)) return 0; /* nil or non-Pod owner */
  return (g->flags & LWSE_GRP_WSTS_OWNER_NAME_MATCH) && g->wsts_owner_uid_hash == g->leader_uid_hash;
}

/* :204-266 handleRestartPolicy for one pod event.
 * returns 1 = leaderDeleted (true,nil); 0 = (false,nil); -1 = (false,err).
 * *issued_delete = 1 when r.Delete(leader) would be called. */
static int handle_restart_policy(const lwse_lws_rec* l, const lwse_group_rec* g,
                                 const pod_cols* pods, const pod_view* p,
                                 int* issued_delete) {
  *issued_delete = 0;
  uint32_t policy = (l->flags & LWSE_LWS_RESTART_MASK) >> LWSE_LWS_RESTART_SHIFT;
  if (policy != LWSE_RESTART_ON_POD_RESTART && policy != LWSE_RESTART_AFTER_START) return 0;
  if (!container_restarted(p) && !(p->bits & LWSE_POD_DELETING)) return 0;
  int pending = pending_pods_in_group(g, pods, (int)l->size);
  int has_annot = (l->flags & LWSE_LWS_RECREATE_AFTER_START_ANNOT) != 0;
  if (pending && (policy == LWSE_RESTART_AFTER_START || has_annot)) return 0;
  int leader_deleting;
  if (!(p->bits & LWSE_POD_IS_LEADER)) {
    if (!(p->place & LWSE_PODID_NAME_OK)) return -1; /* :230-232 */
    /* :233-237 Get(leader by parsed name) */
    if (!((g->flags & LWSE_GRP_POD_PRESENT) && (g->flags & LWSE_GRP_POD_NAME_MATCH))) return 0;
    if (g->leader_rev_hash != p->rev_hash) return 0; /* :239 */
    if (!worker_pod_belongs_to_leader(p, g)) return 0; /* :244-250 */
    leader_deleting = (g->flags & LWSE_GRP_POD_DELETING) != 0;
  } else {
    leader_deleting = (p->bits & LWSE_POD_DELETING) != 0; /* leader = pod */
  }
  if (leader_deleting) return 1; /* :255 */
  *issued_delete = 1;            /* :259 */
  return 1;
}

/* One pod group: restart sweep over its pods (B1/B2), then the leader pod's
 * own Reconcile tail (:100-198) for worker-sts gating and topology (B3/B4). */
static void reconcile_group(const lwse_lws_rec* l, const lwse_group_rec* g,
                            const pod_cols* pods, const lwse_node_rec* nodes,
                            uint32_t n_nodes, uint32_t sweep_flags, lwse_group_out* o) {
  uint32_t f = 0;
  o->first_trigger = LWSE_NONE;
  o->worker_replicas = 0;
  o->domain_id = LWSE_NONE;
  int no_worker_sts = l->size == 1;

  /* the per-replica state bits that the LWS-level pass consumes */
  {
    replica_state s = {0, 0};
    if ((g->flags & LWSE_GRP_POD_NAME_MATCH) &&
        (no_worker_sts || (g->flags & LWSE_GRP_WSTS_LABEL_NAME_MATCH))) {
      int lu = g->leader_rev_hash == l->rev_hash, lr = pod_running_and_ready(g);
      s.ready = lr && (no_worker_sts || statefulset_ready(g));
      s.updated = lu && (no_worker_sts || g->wsts_rev_hash == l->rev_hash);
    }
    if (s.ready) f |= LWSE_GOUT_STATE_READY;
    if (s.updated) f |= LWSE_GOUT_STATE_UPDATED;
    if ((g->flags & LWSE_GRP_POD_PRESENT) && (no_worker_sts || (g->flags & LWSE_GRP_WSTS_FOUND))) {
      f |= LWSE_GOUT_COUNTED;
      if ((no_worker_sts || statefulset_ready(g)) && pod_running_and_ready(g)) f |= LWSE_GOUT_COND_READY;
      if ((no_worker_sts || g->wsts_rev_hash == l->rev_hash) && g->leader_rev_hash == l->rev_hash)
        f |= LWSE_GOUT_COND_UPDATED;
    }
  }

  if (pending_pods_in_group(g, pods, (int)l->size)) f |= LWSE_GOUT_PENDING;

  int leader_deleted = 0; /* result of handleRestartPolicy for the leader pod's own event */
  for (uint32_t i = 0; i < g->pod_count; i++) {
    const pod_view pv = get_pod(pods, (uint64_t)g->pod_base + i);
    const pod_view* p = &pv;
    int issued = 0;
    int r = handle_restart_policy(l, g, pods, p, &issued);
    if (r < 0) {
      f |= LWSE_GOUT_RESTART_ERROR;
      continue;
    }
    if (r > 0) {
      if (o->first_trigger == LWSE_NONE) o->first_trigger = i;
      f |= issued ? LWSE_GOUT_DELETE_LEADER : LWSE_GOUT_LEADER_DELETING;
      if (p->bits & LWSE_POD_IS_LEADER) leader_deleted = 1;
    }
  }

  /* pod_controller.go:95-198 for the leader pod of this group */
  do {
    if (!(g->flags & LWSE_GRP_POD_PRESENT)) break;
    if (leader_deleted) break;                                  /* :95 */
    if (g->flags & LWSE_GRP_MISTAKEN_ANNOTATION) break;         /* :106 */
    if (g->flags & LWSE_GRP_POD_DELETING) break;                /* :125 */
    if (sweep_flags & LWSE_SWEEP_GANG) f |= LWSE_GOUT_CREATE_PODGROUP; /* :130 */
    if (l->size == 1) break;                                    /* :138 */
    if ((l->flags & LWSE_LWS_STARTUP_LEADER_READY) && !(g->flags & LWSE_GRP_POD_READY)) break; /* :143 */
    if (!(g->flags & LWSE_GRP_REVISION_EXISTS)) {               /* :152 */
      f |= LWSE_GOUT_REQUEUE_REVISION;
      break;
    }
    if (l->flags & LWSE_LWS_EXCLUSIVE_TOPOLOGY) {               /* :162 */
      if (g->leader_node == LWSE_NONE) {                        /* :164 */
        f |= LWSE_GOUT_WAIT_SCHEDULE;
        break;
      }
      /* :315-336 topologyValueFromPod */
      if (g->leader_node != LWSE_NODE_NOT_FOUND && g->leader_node < n_nodes) {
        const lwse_node_rec* n = &nodes[g->leader_node];
        if (!(n->flags & LWSE_NODE_HAS_TOPOLOGY)) {             /* :330 */
          f |= LWSE_GOUT_TOPOLOGY_ERROR;
          break;
        }
        o->domain_id = n->domain_id;
      } /* NotFound node: empty value, nil error (:327) */
    }
    if (!(g->flags & LWSE_GRP_WSTS_FOUND)) {                    /* :188-192 */
      f |= LWSE_GOUT_CREATE_WSTS;
      o->worker_replicas = l->size - 1;                         /* :437, ordinals start 1 :440 */
    }
  } while (0);
  o->flags = f;
}

/* ------------------------------------------------------------------------- */
/* whole-table sweep                                                         */
/* ------------------------------------------------------------------------- */

static void sweep_one_lws(const lwse_lws_tables* t, uint32_t i) {
  const lwse_lws_rec* l = &t->lws[i];
  lwse_lws_out* o = &t->lws_out[i];
  memset(o, 0, sizeof(*o));
  if ((uint64_t)l->group_base + l->group_count > t->n_groups) {
    o->flags = LWSE_LOUT_BAD_TABLE;
    return;
  }
  uint32_t event = 0;
  int32_t unready = 0;
  if (rolling_update_parameters(l, t->groups, &o->sts_partition, &o->sts_replicas, &event, &unready))
    o->flags |= LWSE_LOUT_RUP_ERROR;
  o->flags |= event << LWSE_LOUT_EVENT_SHIFT;
  o->unready_replicas = unready;
  o->sts_max_unavailable = (l->flags & LWSE_LWS_INTSTR_INVALID) ? 0 : sts_max_unavailable(l);
  if (update_conditions(l, t->groups, o)) o->flags |= LWSE_LOUT_STATUS_ERROR;
  /* volcano_provider.go:72,81-83 */
  if (t->flags & LWSE_SWEEP_GANG)
    o->min_member = (l->flags & LWSE_LWS_STARTUP_LEADER_READY) ? 1 : l->size;
  if (l->flags & LWSE_LWS_IRREGULAR) o->flags |= LWSE_LOUT_IRREGULAR;
}

static void sweep_one_group(const lwse_lws_tables* t, const lwse_node_rec* nodes, uint32_t n_nodes,
                            uint32_t r) {
  const lwse_group_rec* g = &t->groups[r];
  lwse_group_out* o = &t->group_out[r];
  if (g->lws_index >= t->n_lws || (uint64_t)g->pod_base + g->pod_count > t->n_pods) {
    o->flags = LWSE_GOUT_BAD_TABLE;
    o->first_trigger = LWSE_NONE;
    o->worker_replicas = 0;
    o->domain_id = LWSE_NONE;
    return;
  }
  pod_cols cols = {t->pod_state, t->pod_ident};
  reconcile_group(&t->lws[g->lws_index], g, &cols, nodes, n_nodes, t->flags, o);
}

/* Sweep every object, one at a time (the reference runs one reconcile worker
 * per controller).  threads > 1 splits the row ranges over pthreads — used
 * only by bench.py's all-cores baseline. */
typedef struct {
  const lwse_lws_tables* t;
  const lwse_node_rec* nodes;
  uint32_t n_nodes;
  uint32_t lws_begin, lws_end, grp_begin, grp_end;
} sweep_job;

static void* sweep_range(void* arg) {
  sweep_job* j = (sweep_job*)arg;
  for (uint32_t i = j->lws_begin; i < j->lws_end; i++) sweep_one_lws(j->t, i);
  for (uint32_t r = j->grp_begin; r < j->grp_end; r++) sweep_one_group(j->t, j->nodes, j->n_nodes, r);
  return NULL;
}

LWSO_API int lwso_sweep_lws(const lwse_lws_tables* t, const lwse_node_rec* nodes, uint32_t n_nodes,
                            int threads) {
  if (!t) return LWSE_ERR_INVALID_ARG;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads == 1 || t->n_groups < (uint32_t)threads * 4u) {
    sweep_job j = {t, nodes, n_nodes, 0, t->n_lws, 0, t->n_groups};
    sweep_range(&j);
  } else {
    pthread_t tid[256];
    sweep_job jobs[256];
    for (int k = 0; k < threads; k++) {
      jobs[k] = (sweep_job){t, nodes, n_nodes,
                            (uint32_t)((uint64_t)t->n_lws * (uint64_t)k / (uint64_t)threads),
                            (uint32_t)((uint64_t)t->n_lws * (uint64_t)(k + 1) / (uint64_t)threads),
                            (uint32_t)((uint64_t)t->n_groups * (uint64_t)k / (uint64_t)threads),
                            (uint32_t)((uint64_t)t->n_groups * (uint64_t)(k + 1) / (uint64_t)threads)};
      pthread_create(&tid[k], NULL, sweep_range, &jobs[k]);
    }
    for (int k = 0; k < threads; k++) pthread_join(tid[k], NULL);
  }
  if (t->node_occupancy) {
    /* scheduled pods per node, over the whole pod table */
    memset(t->node_occupancy, 0, sizeof(uint32_t) * (size_t)n_nodes);
    for (uint64_t p = 0; p < t->n_pods; p++) {
      uint32_t b = t->pod_ident[p].place;
      if (b & LWSE_PODID_SCHEDULED) {
        uint32_t node = b >> LWSE_PODID_NODE_SHIFT;
        if (node < n_nodes) t->node_occupancy[node]++;
      }
    }
  }
  return LWSE_OK;
}

/* ------------------------------------------------------------------------- */
/* event-driven form (bench.py's churn workloads)                            */
/* ------------------------------------------------------------------------- */
/* A controller does not sweep: watch events enqueue the objects they touch and Reconcile()
 * runs for those (leaderworkerset_controller.go:106, pod_controller.go:69).  The CPU arm of a
 * churn step therefore (1) applies the step's row patches to its host tables and (2) reconciles
 * exactly the dirty groups / objects, whose row numbers the event source hands it. */
LWSO_API void lwso_apply_patch(void* table, uint32_t row_bytes, uint64_t table_rows, const uint32_t* rows,
                               const void* values, uint32_t n) {
  uint8_t* t = (uint8_t*)table;
  const uint8_t* v = (const uint8_t*)values;
  for (uint32_t i = 0; i < n; i++)
    if (rows[i] < table_rows) memcpy(t + (size_t)rows[i] * row_bytes, v + (size_t)i * row_bytes, row_bytes);
}

typedef struct {
  const lwse_lws_tables* t;
  const lwse_node_rec* nodes;
  uint32_t n_nodes;
  const uint32_t *dg, *dl;
  uint32_t dg_begin, dg_end, dl_begin, dl_end;
} dirty_job;

static void* dirty_range(void* arg) {
  dirty_job* j = (dirty_job*)arg;
  for (uint32_t k = j->dl_begin; k < j->dl_end; k++)
    if (j->dl[k] < j->t->n_lws) sweep_one_lws(j->t, j->dl[k]);
  for (uint32_t k = j->dg_begin; k < j->dg_end; k++)
    if (j->dg[k] < j->t->n_groups) sweep_one_group(j->t, j->nodes, j->n_nodes, j->dg[k]);
  return NULL;
}

LWSO_API int lwso_sweep_dirty(const lwse_lws_tables* t, const lwse_node_rec* nodes, uint32_t n_nodes,
                              const uint32_t* dirty_groups, uint32_t n_dg, const uint32_t* dirty_lws,
                              uint32_t n_dl, int threads) {
  if (!t) return LWSE_ERR_INVALID_ARG;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads == 1 || (uint64_t)n_dg + n_dl < (uint64_t)threads * 4u) {
    dirty_job j = {t, nodes, n_nodes, dirty_groups, dirty_lws, 0, n_dg, 0, n_dl};
    dirty_range(&j);
    return LWSE_OK;
  }
  pthread_t tid[256];
  dirty_job jobs[256];
  for (int k = 0; k < threads; k++) {
    jobs[k] = (dirty_job){t, nodes, n_nodes, dirty_groups, dirty_lws,
                          (uint32_t)((uint64_t)n_dg * (uint64_t)k / (uint64_t)threads),
                          (uint32_t)((uint64_t)n_dg * (uint64_t)(k + 1) / (uint64_t)threads),
                          (uint32_t)((uint64_t)n_dl * (uint64_t)k / (uint64_t)threads),
                          (uint32_t)((uint64_t)n_dl * (uint64_t)(k + 1) / (uint64_t)threads)};
    pthread_create(&tid[k], NULL, dirty_range, &jobs[k]);
  }
  for (int k = 0; k < threads; k++) pthread_join(tid[k], NULL);
  return LWSE_OK;
}

/* leaf helpers exported for the known-answer tests */
LWSO_API int lwso_scaled_value(int32_t val, int is_percent, int total, int round_up) {
  return scaled_value(val, is_percent, total, round_up);
}

/* pkg/webhooks/pod_webhook.go:249-255 getSubGroupIndex */
LWSO_API int lwso_sub_group_index(int pod_count, int sub_group_size, int worker_index) {
  if ((pod_count - 1) % sub_group_size == 0) return (worker_index - 1) / sub_group_size;
  return worker_index / sub_group_size;
}

LWSO_API uint32_t lwso_abi_version(void) { return LWSE_ABI_VERSION; }
