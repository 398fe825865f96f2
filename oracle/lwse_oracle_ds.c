/*
 * lwse_oracle_ds.c — CPU restatement of the DisaggregatedSet rollout arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY (see lwse_oracle.c).  Follows
 *   pkg/controllers/disaggregatedset/planner.go:61-352   (planner, float64 kept)
 *   pkg/controllers/disaggregatedset/executor.go:199-302 (planner state, config,
 *                                                         stability, newest-first)
 *   pkg/controllers/disaggregatedset/executor.go:306-398 (scaleUpNew / scaleDownOld)
 *   pkg/controllers/disaggregatedset/disaggregatedset_controller.go:95-112,137-186,203-236
 *   pkg/controllers/disaggregatedset/service_manager.go:57-89,174-189
 *   pkg/utils/disaggregatedset/utils.go:160-190
 * of kubernetes-sigs/lws @ 1d9204a2, over the DS tables of include/lwse.h.
 * Pinned by tests/test_oracle_ds_golden.py against planner_test.go and
 * executor_test.go vectors.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/lwse.h"

#define LWSO_API __attribute__((visibility("default")))
#define MAXR LWSE_DS_MAX_ROLES

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
  int max_surge, max_unavailable;
} ru_config; /* planner.go:47-50 */

/* planner.go:61-66 */
static int batch_size(int max_surge, int max_unavailable) {
  if (max_surge > 0) return max_surge;
  return imax(1, max_unavailable);
}

/* planner.go:68-78 */
static int compute_total_steps(int n, const int* initial_old, const int* target, const ru_config* cfg) {
  int total = 0;
  for (int i = 0; i < n; i++) {
    int max_replicas = imax(imax(initial_old[i], target[i]), 0);
    int b = batch_size(cfg[i].max_surge, cfg[i].max_unavailable);
    int role_steps = (max_replicas + b - 1) / b;
    total = imax(total, role_steps);
  }
  return total;
}

/* planner.go:80-113 */
static void compute_next_new_replicas(int n, const int* target, const int* current_new, int total_steps,
                                      int* result) {
  if (total_steps == 0) {
    memcpy(result, target, sizeof(int) * (size_t)n);
    return;
  }
  int min_step_idx = total_steps;
  for (int i = 0; i < n; i++) {
    int step_idx;
    if (target[i] == 0)
      step_idx = total_steps;
    else
      step_idx = (int)((double)current_new[i] * (double)total_steps / (double)target[i]);
    min_step_idx = imin(min_step_idx, step_idx);
  }
  int next_step_idx = min_step_idx + 1;
  for (int i = 0; i < n; i++) {
    double progress = (double)next_step_idx * (double)target[i] / (double)total_steps;
    int computed = imin((int)ceil(progress), target[i]);
    result[i] = imax(computed, current_new[i]);
  }
}

/* planner.go:115-149 */
static void compute_next_old_replicas(int n, const int* initial_old, const int* current_old,
                                      int total_steps, int* result) {
  if (total_steps == 0) {
    memset(result, 0, sizeof(int) * (size_t)n);
    return;
  }
  int max_step_idx = 0;
  for (int i = 0; i < n; i++) {
    if (initial_old[i] == 0) continue;
    int removed = initial_old[i] - current_old[i];
    int step_idx = (int)((double)removed * (double)total_steps / (double)initial_old[i]);
    max_step_idx = imax(max_step_idx, step_idx);
  }
  int next_step_idx = max_step_idx + 1;
  for (int i = 0; i < n; i++) {
    double progress = (double)next_step_idx * (double)initial_old[i] / (double)total_steps;
    int computed = imax(0, initial_old[i] - (int)floor(progress));
    result[i] = imin(computed, current_old[i]);
  }
}

/* Exported for the tests that pin the helpers on the reference's own helper-level vectors
 * (planner_test.go:674-720, :800-902). */
LWSO_API int lwso_ds_batch_size(int max_surge, int max_unavailable) { return batch_size(max_surge, max_unavailable); }
LWSO_API int lwso_ds_total_steps(int n, const int* initial_old, const int* target, const int* max_surge,
                                 const int* max_unavailable) {
  ru_config cfg[16];
  if (n < 0 || n > 16) return -1;
  for (int i = 0; i < n; i++) {
    cfg[i].max_surge = max_surge[i];
    cfg[i].max_unavailable = max_unavailable[i];
  }
  return compute_total_steps(n, initial_old, target, cfg);
}
LWSO_API void lwso_ds_next_new(int n, const int* target, const int* current_new, int total_steps, int* result) {
  compute_next_new_replicas(n, target, current_new, total_steps, result);
}
LWSO_API void lwso_ds_next_old(int n, const int* initial_old, const int* current_old, int total_steps, int* result) {
  compute_next_old_replicas(n, initial_old, current_old, total_steps, result);
}

/* planner.go:252-262 */
static int can_drain_all_to_zero(int n, const int* next_new, const int* initial_old, const int* target,
                                 const ru_config* cfg) {
  for (int i = 0; i < n; i++) {
    if (initial_old[i] >= target[i]) {
      int min_required = target[i] - cfg[i].max_unavailable;
      if (next_new[i] < min_required) return 0;
    }
  }
  return 1;
}

/* planner.go:264-294 */
static void apply_orphan_prevention(int n, int* next_old, const int* current_new, const int* initial_old,
                                    const int* target, const ru_config* cfg) {
  int any_drains_to_zero = 0, all_drain_to_zero = 1;
  for (int i = 0; i < n; i++) {
    if (initial_old[i] == 0) continue;
    if (next_old[i] == 0)
      any_drains_to_zero = 1;
    else
      all_drain_to_zero = 0;
  }
  if (!any_drains_to_zero || all_drain_to_zero) return;
  if (can_drain_all_to_zero(n, current_new, initial_old, target, cfg)) {
    for (int i = 0; i < n; i++) next_old[i] = 0;
    return;
  }
  for (int i = 0; i < n; i++)
    if (next_old[i] == 0 && initial_old[i] > 0) next_old[i] = 1;
}

/* planner.go:320-352 ComputeNextStep.  Returns 1 and fills past/new_ when a
 * step exists, 0 for nil. */
LWSO_API int lwso_ds_compute_next_step(int n, const int* initial_old, const int* current_old,
                                       const int* current_new, const int* target_new,
                                       const int* max_surge, const int* max_unavailable, int* past,
                                       int* new_) {
  ru_config cfg[MAXR];
  if (n < 0 || n > (int)MAXR) return -1;
  for (int i = 0; i < n; i++) {
    cfg[i].max_surge = max_surge[i];
    cfg[i].max_unavailable = max_unavailable[i];
  }
  /* isComplete :173-180 */
  int complete = 1;
  for (int i = 0; i < n; i++)
    if (current_old[i] != 0 || current_new[i] < target_new[i]) complete = 0;
  if (complete) return 0;

  int total_steps = compute_total_steps(n, initial_old, target_new, cfg);
  if (total_steps == 0) return 0;

  /* correctAbnormalState :151-171 */
  {
    int needs = 0;
    int expected[MAXR];
    for (int i = 0; i < n; i++) {
      expected[i] = imin(initial_old[i], current_old[i]);
      if (current_old[i] > expected[i]) needs = 1;
    }
    if (needs) {
      memcpy(past, expected, sizeof(int) * (size_t)n);
      memcpy(new_, current_new, sizeof(int) * (size_t)n);
      return 1;
    }
  }
  /* isNewAtTarget :182-189 */
  {
    int at = 1;
    for (int i = 0; i < n; i++)
      if (current_new[i] < target_new[i]) at = 0;
    if (at) {
      memset(past, 0, sizeof(int) * (size_t)n);
      memcpy(new_, current_new, sizeof(int) * (size_t)n);
      return 1;
    }
  }
  int next_new[MAXR], min_old[MAXR];
  compute_next_new_replicas(n, target_new, current_new, total_steps, next_new);
  /* computeMinOld :203-211 */
  for (int i = 0; i < n; i++) {
    min_old[i] = 0;
    if (initial_old[i] >= target_new[i])
      min_old[i] = imax(0, target_new[i] - cfg[i].max_unavailable - current_new[i]);
  }
  /* tryScaleUp :213-228 (+ canScaleUp :191-201) */
  {
    int needs = 0;
    for (int i = 0; i < n; i++)
      if (next_new[i] > current_new[i]) {
        needs = 1;
        break;
      }
    if (needs) {
      int can = 1;
      for (int i = 0; i < n; i++) {
        if (target_new[i] == 0) continue;
        if (current_old[i] + next_new[i] > target_new[i] + cfg[i].max_surge) can = 0;
      }
      if (can) {
        memcpy(past, current_old, sizeof(int) * (size_t)n);
        memcpy(new_, next_new, sizeof(int) * (size_t)n);
        return 1;
      }
    }
  }
  /* tryProportionalDrain :230-250 */
  {
    int next_old[MAXR];
    compute_next_old_replicas(n, initial_old, current_old, total_steps, next_old);
    for (int i = 0; i < n; i++) next_old[i] = imax(next_old[i], min_old[i]);
    apply_orphan_prevention(n, next_old, current_new, initial_old, target_new, cfg);
    int needs = 0;
    for (int i = 0; i < n; i++)
      if (next_old[i] < current_old[i]) {
        needs = 1;
        break;
      }
    if (needs) {
      memcpy(past, next_old, sizeof(int) * (size_t)n);
      memcpy(new_, current_new, sizeof(int) * (size_t)n);
      return 1;
    }
  }
  /* tryForceDrain :296-318 */
  {
    int drained[MAXR];
    int needs = 0;
    for (int i = 0; i < n; i++) {
      int max_old = target_new[i] + cfg[i].max_surge - next_new[i];
      drained[i] = imax(0, imin(current_old[i], max_old));
      if (initial_old[i] >= target_new[i]) {
        int min_old_for_role = imax(0, target_new[i] - cfg[i].max_unavailable - next_new[i]);
        drained[i] = imax(drained[i], min_old_for_role);
      }
      if (drained[i] < current_old[i]) needs = 1;
    }
    if (needs) {
      apply_orphan_prevention(n, drained, next_new, initial_old, target_new, cfg);
      memcpy(past, drained, sizeof(int) * (size_t)n);
      memcpy(new_, next_new, sizeof(int) * (size_t)n);
      return 1;
    }
  }
  return 0;
}

/* k8s.io/apimachinery intstr.GetScaledValueFromIntOrPercent (float64 kept) */
static int scaled(int32_t val, int is_percent, int total, int round_up) {
  if (!is_percent) return val;
  double v = (double)val * (double)total / 100.0;
  return round_up ? (int)ceil(v) : (int)floor(v);
}

/* executor.go:330-398 scaleDownOld over revisions given newest-first.
 * replicas[r*n+i] (or -1 when the role has no LWS in revision r) is updated in place. */
LWSO_API int lwso_ds_scale_down_old(int n, int n_revs, int* replicas, const int* order,
                                    const int* current, const int* target, int* scaled_mask) {
  int budget[MAXR];
  if (n < 0 || n > (int)MAXR) return -1;
  for (int i = 0; i < n; i++) budget[i] = current[i] - target[i];
  for (int k = 0; k < n_revs; k++) {
    int all_zero = 1; /* allZero :400-407: no budget left */
    for (int i = 0; i < n; i++)
      if (budget[i] > 0) all_zero = 0;
    if (all_zero) break;
    int r = order ? order[k] : k;
    int new_rep[MAXR], planned[MAXR], trig[MAXR];
    int any = 0;
    for (int i = 0; i < n; i++) {
      trig[i] = 0;
      planned[i] = 0;
      new_rep[i] = 0;
      int rep = replicas[r * n + i];
      if (rep < 0) continue;
      int drain = imin(budget[i], rep);
      planned[i] = drain;
      new_rep[i] = rep - drain;
      if (new_rep[i] == 0) {
        trig[i] = 1;
        any = 1;
      }
    }
    if (any)
      for (int i = 0; i < n; i++)
        if (replicas[r * n + i] >= 0) new_rep[i] = 0;
    for (int i = 0; i < n; i++) {
      int rep = replicas[r * n + i];
      if (rep < 0) continue;
      if (rep <= new_rep[i]) continue;
      replicas[r * n + i] = new_rep[i]; /* LWSManager.Scale */
      if (scaled_mask) scaled_mask[r * n + i] = 1;
      if (trig[i] || !any) budget[i] -= planned[i];
    }
  }
  return 0;
}

static void sweep_one_ds(const lwse_ds_tables* t, uint32_t d) {
  const lwse_ds_rec* ds = &t->ds[d];
  lwse_ds_out* o = &t->ds_out[d];
  memset(o, 0, sizeof(*o));
  const int n = (int)ds->n_roles, S = (int)ds->n_spec_roles, V = (int)ds->n_old_revs;
  if (ds->n_roles > MAXR || ds->n_spec_roles > ds->n_roles || ds->n_old_revs > LWSE_DS_MAX_OLD_REVS ||
      (uint64_t)ds->role_base + ds->n_roles > t->n_roles ||
      (uint64_t)ds->rev_base + (uint64_t)(V + 1) * (uint64_t)n > t->n_revroles) {
    o->flags = LWSE_DOUT_BAD_TABLE;
    return;
  }
  const lwse_ds_role_rec* roles = t->roles + ds->role_base;
  const lwse_ds_revrole_rec* rr = t->revroles + ds->rev_base;
  const lwse_ds_revrole_rec* nw = rr + (size_t)V * (size_t)n;
  lwse_ds_role_out* ro = t->role_out + ds->role_base;
  lwse_ds_revrole_out* rro = t->revrole_out + ds->rev_base;

  for (int i = 0; i < n; i++) ro[i].next_old = ro[i].next_new = -1;
  for (int k = 0; k < (V + 1) * n; k++) rro[k] = (rr[k].flags & LWSE_RR_EXISTS) ? rr[k].replicas : -1;

  /* cleanupDrainedLWS, disaggregatedset_controller.go:203-236 */
  for (int r = 0; r < V; r++) {
    int any = 0, all_drained = 1;
    for (int i = 0; i < n; i++) {
      const lwse_ds_revrole_rec* x = &rr[r * n + i];
      if (!(x->flags & LWSE_RR_EXISTS)) continue;
      any = 1;
      int rep = (x->flags & LWSE_RR_REPLICAS_NIL) ? 0 : x->replicas;
      if (rep != 0) all_drained = 0;
    }
    if (any && all_drained) o->drained_revs |= 1u << r;
  }
  /* service readiness, service_manager.go:57-89,174-189 */
  for (int r = 0; r <= V; r++) {
    int ready = 1;
    for (int i = 0; i < S; i++) {
      const lwse_ds_revrole_rec* x = &rr[r * n + i];
      if (!(x->flags & LWSE_RR_EXISTS) || x->ready_replicas < 1) ready = 0;
    }
    if (r < V) {
      if (ready) o->ready_revs |= 1u << r;
    } else if (ready) {
      o->flags |= LWSE_DOUT_NEW_READY;
    }
  }

  /* Reconcile :95-112 */
  int total_old = 0;
  for (int i = 0; i < S; i++)
    for (int r = 0; r < V; r++)
      if (rr[r * n + i].flags & LWSE_RR_EXISTS) total_old += rr[r * n + i].replicas;
  if (!(V > 0 && total_old > 0)) {
    /* reconcileSimple :137-186 */
    for (int i = 0; i < S; i++) rro[V * n + i] = roles[i].target_replicas;
    return;
  }
  o->flags |= LWSE_DOUT_ROLLING;
  if (!(ds->flags & LWSE_DS_HAS_NEW_REVISION)) {
    /* initRollingUpdate executor.go:85-124: new LWS per spec role at 0 replicas */
    o->flags |= LWSE_DOUT_INIT;
    for (int i = 0; i < S; i++) rro[V * n + i] = 0;
    return;
  }
  /* isRevisionStable executor.go:270-281 */
  for (int i = 0; i < S; i++) {
    if (!(nw[i].flags & LWSE_RR_EXISTS)) return;
    if (nw[i].replicas != nw[i].ready_replicas) return;
  }
  o->flags |= LWSE_DOUT_STABLE;

  /* buildPlannerState executor.go:199-221 (+ utils.go:160-190) */
  int initial_old[MAXR], current_old[MAXR], current_new[MAXR], target_new[MAXR];
  int ms[MAXR], mu[MAXR];
  for (int i = 0; i < n; i++) {
    initial_old[i] = current_old[i] = current_new[i] = target_new[i] = 0;
    for (int r = 0; r < V; r++) {
      const lwse_ds_revrole_rec* x = &rr[r * n + i];
      if (!(x->flags & LWSE_RR_EXISTS)) continue;
      current_old[i] += x->replicas;
      initial_old[i] += x->initial_replicas >= 0 ? x->initial_replicas : x->replicas;
    }
    if (roles[i].flags & LWSE_ROLE_IN_SPEC) {
      if (nw[i].flags & LWSE_RR_EXISTS) current_new[i] = nw[i].replicas;
      target_new[i] = roles[i].target_replicas;
    }
    /* extractRollingUpdateConfig executor.go:235-260 */
    ms[i] = 1;
    mu[i] = 0;
    if ((roles[i].flags & LWSE_ROLE_IN_SPEC) && (roles[i].flags & LWSE_ROLE_HAS_ROLLING_CONFIG)) {
      int replicas = roles[i].target_replicas;
      int surge = (roles[i].flags & LWSE_ROLE_SURGE_INVALID)
                      ? 0
                      : scaled(roles[i].max_surge, roles[i].flags & LWSE_ROLE_SURGE_IS_PERCENT, replicas, 1);
      int unavail = (roles[i].flags & LWSE_ROLE_UNAVAIL_INVALID)
                        ? 0
                        : scaled(roles[i].max_unavailable, roles[i].flags & LWSE_ROLE_UNAVAIL_IS_PERCENT,
                                 replicas, 0);
      if (unavail > 0) {
        mu[i] = unavail;
        ms[i] = surge;
      } else if (surge > 0) {
        ms[i] = surge;
      }
    }
  }
  int past[MAXR], new_[MAXR];
  int has = lwso_ds_compute_next_step(n, initial_old, current_old, current_new, target_new, ms, mu, past, new_);
  if (has <= 0) {
    o->flags |= LWSE_DOUT_COMPLETE;
    return;
  }
  o->flags |= LWSE_DOUT_STEP;
  for (int i = 0; i < n; i++) {
    ro[i].next_old = past[i];
    ro[i].next_new = new_[i];
  }
  /* scaleUpNew executor.go:306-328 */
  for (int i = 0; i < n; i++) {
    if (!(roles[i].flags & LWSE_ROLE_IN_SPEC) || current_new[i] >= new_[i]) continue;
    rro[V * n + i] = new_[i];
  }
  /* sortByNewestTimestamp executor.go:283-302 (stable insertion sort, newest first) */
  int order[LWSE_DS_MAX_OLD_REVS];
  uint32_t ts[LWSE_DS_MAX_OLD_REVS];
  for (int r = 0; r < V; r++) {
    uint32_t m = 0;
    for (int i = 0; i < n; i++)
      if (rr[r * n + i].flags & LWSE_RR_EXISTS) {
        uint32_t v = rr[r * n + i].flags >> LWSE_RR_TS_SHIFT;
        if (v > m) m = v;
      }
    ts[r] = m;
    order[r] = r;
  }
  for (int a = 1; a < V; a++) {
    int x = order[a], b = a - 1;
    while (b >= 0 && ts[order[b]] < ts[x]) {
      order[b + 1] = order[b];
      b--;
    }
    order[b + 1] = x;
  }
  /* scaleDownOld executor.go:330-398 */
  int reps[LWSE_DS_MAX_OLD_REVS * MAXR];
  for (int k = 0; k < V * n; k++) reps[k] = (rr[k].flags & LWSE_RR_EXISTS) ? rr[k].replicas : -1;
  lwso_ds_scale_down_old(n, V, reps, order, current_old, past, NULL);
  for (int k = 0; k < V * n; k++) rro[k] = reps[k];
}

LWSO_API int lwso_sweep_ds(const lwse_ds_tables* t) {
  if (!t) return LWSE_ERR_INVALID_ARG;
  for (uint32_t d = 0; d < t->n_ds; d++) sweep_one_ds(t, d);
  return LWSE_OK;
}
