"""DisaggregatedSetReconciler.Reconcile in miniature (disaggregatedset_controller.go:55-126):
one sweep decides, this module applies the decisions to the child LWS objects the way the
reference's LWSManager would, and plays `simulateAllReady` (executor_test.go:190-201)."""
from __future__ import annotations

from lws_b200 import api, encoder
from lws_b200 import records as R


class DsSim:
    def __init__(self, ds: api.DisaggregatedSet, revision: str, children: list, sweep_ds):
        self.ds, self.revision, self.children, self.sweep_ds = ds, revision, list(children), sweep_ds
        self.clock = 100.0
        self.history = []

    def child(self, role, rev):
        for c in self.children:
            if c.role == role and c.revision == rev:
                return c
        return None

    def reconcile(self):
        t = encoder.encode_ds([encoder.DsItem(self.ds, self.revision, self.children)])
        ds_out, role_out, rr_out = self.sweep_ds(t)
        flags = int(ds_out[0]["flags"])
        names, old_revs = t.role_names[0], t.old_revisions[0]
        n = len(names)
        # cleanupDrainedLWS (:203-236)
        for r, rev in enumerate(old_revs):
            if int(ds_out[0]["drained_revs"]) >> r & 1:
                self.children = [c for c in self.children if c.revision != rev]
        if flags & R.DOUT_INIT:
            # initRollingUpdate (executor.go:85-124): snapshot initial-replicas, create new LWS at 0
            for c in self.children:
                if c.revision != self.revision:
                    c.annotations[api.DSInitialReplicasAnnotationKey] = str(1 if c.replicas is None else c.replicas)
        # replicas decided by the sweep: scaleUpNew / scaleDownOld / reconcileSimple / init
        for r, rev in enumerate(old_revs + [self.revision]):
            for i, role in enumerate(names):
                want = int(rr_out[r * n + i])
                c = self.child(role, rev)
                if c is None:
                    if want >= 0 and rev == self.revision and not (int(ds_out[0]["drained_revs"]) and False):
                        self.clock += 1
                        self.children.append(api.ChildLWS(role, rev, want, 0, self.clock))
                elif want >= 0 and c in self.children:
                    c.replicas = want
        self.history.append({(c.revision, c.role): c.replicas for c in self.children})
        return flags

    def simulate_all_ready(self):
        for c in self.children:
            if c.replicas is not None:
                c.readyReplicas = c.replicas

    def replicas(self, role, rev):
        c = self.child(role, rev)
        return -1 if c is None else c.replicas
