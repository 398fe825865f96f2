"""A miniature of the reference's integration-test environment.

test/integration/controllers/suite_test.go:69-122 runs the real LWS and Pod
reconcilers against envtest (apiserver + etcd, *no* StatefulSet controller,
kubelet or scheduler); the tests play those missing parts by hand with the
helpers of test/testutils/util.go.  ``LwsSim`` reproduces exactly that set-up
over the host object model so the integration traces of
test/integration/controllers/leaderworkerset_test.go can be replayed against
any ``sweep`` implementation (CPU oracle or the CUDA engine through the C ABI).

sweep(tables: encoder.LwsTables, flags) -> (lws_out, group_out)
"""
from __future__ import annotations

import copy

from lws_b200 import api, encoder
from lws_b200 import records as R


class LwsSim:
    def __init__(self, lws: api.LeaderWorkerSet, sweep, gang=False, nodes=None, topology_key=None):
        self.lws = lws
        self.sweep = sweep
        self.gang = gang
        self.template_rev = 1  # bumped by update_template(): a new ControllerRevision
        self.leader_sts = None
        self.pods: dict[str, api.Pod] = {}  # leader pods
        self.workers: dict[str, api.Pod] = {}  # worker pods (only the entries that need them create any)
        self.stss: dict[str, api.StatefulSet] = {}
        self.nodes = nodes or []
        self.topology_key = topology_key
        self.status = dict(readyReplicas=0, updatedReplicas=0, condition=None, updateDone=False)
        self.last_lws_out = None
        self.last_group_out = None
        self.last_tables = None
        self.deleted_leaders = []

    # ------------------------------------------------------------------ utils
    @property
    def rev_key(self) -> str:
        return f"rev-{self.template_rev}"

    def _cluster(self):
        return encoder.Cluster(
            pods=list(self.pods.values()) + list(self.workers.values()), statefulsets=list(self.stss.values()),
            nodes=self.nodes
        )

    def _sweep(self):
        lws_updated = (
            self.leader_sts is not None and self.leader_sts.labels.get(api.RevisionKey) != self.rev_key
        )
        item = encoder.LwsItem(
            lws=self.lws, revision_key=self.rev_key, lws_updated=lws_updated, leader_sts=self.leader_sts
        )
        tables = encoder.encode_lws([item], self._cluster(), self.topology_key)
        lws_out, group_out = self.sweep(tables, R.SWEEP_GANG if self.gang else 0)
        self.last_tables, self.last_lws_out, self.last_group_out = tables, lws_out, group_out
        return tables, lws_out, group_out

    # ------------------------------------------------- the two reconcilers
    def reconcile_lws(self) -> bool:
        """LeaderWorkerSetReconciler.Reconcile (leaderworkerset_controller.go:111-211)."""
        _, lws_out, _ = self._sweep()
        o = lws_out[0]
        assert not (o["flags"] & R.LOUT_RUP_ERROR), "rollingUpdateParameters error"
        before = copy.deepcopy((self.leader_sts, self.status))
        if self.leader_sts is None:
            self.leader_sts = api.StatefulSet(name=self.lws.name, namespace=self.lws.namespace)
        sts = self.leader_sts
        # SSAWithStatefulset (:375-411, :831-853)
        sts.replicas = int(o["sts_replicas"])
        sts.partition = int(o["sts_partition"])
        sts.labels = {api.SetNameLabelKey: self.lws.name, api.RevisionKey: self.rev_key}
        sts.annotations = {api.ReplicasAnnotationKey: str(self.lws.replicas)}
        # updateStatus happens in the same Reconcile, after the SSA (:196): the
        # conditions do not depend on the leader sts, so one sweep serves both.
        if not (o["flags"] & R.LOUT_STATUS_ERROR):
            self.status = dict(
                readyReplicas=int(o["ready_replicas"]),
                updatedReplicas=int(o["updated_replicas"]),
                condition=int((o["flags"] & R.LOUT_COND_MASK) >> R.LOUT_COND_SHIFT),
                updateDone=bool(o["flags"] & R.LOUT_UPDATE_DONE),
            )
        return before != (self.leader_sts, self.status)

    def reconcile_pods(self) -> bool:
        """PodReconciler.Reconcile for every pod (pod_controller.go:69-202)."""
        tables, _, group_out = self._sweep()
        changed = False
        for gi in range(len(group_out)):
            go = group_out[gi]
            leader_name = f"{self.lws.name}-{gi}"
            if go["flags"] & R.GOUT_DELETE_LEADER and leader_name in self.pods:
                # Foreground delete of the leader (:259); the test environment has no GC
                self.pods[leader_name].deletionTimestamp = True
                self.deleted_leaders.append(leader_name)
                changed = True
            if go["flags"] & R.GOUT_CREATE_WSTS:
                pod = self.pods[leader_name]
                # constructWorkerStatefulSetApplyConfiguration (:386-458)
                self.stss[leader_name] = api.StatefulSet(
                    name=leader_name,
                    namespace=self.lws.namespace,
                    uid=f"uid-sts-{leader_name}-{len(self.stss)}-{self.template_rev}",
                    labels={
                        api.SetNameLabelKey: self.lws.name,
                        api.GroupIndexLabelKey: pod.labels[api.GroupIndexLabelKey],
                        api.RevisionKey: pod.labels.get(api.RevisionKey, ""),
                    },
                    replicas=int(go["worker_replicas"]),
                    ownerReferences=[api.OwnerReference("Pod", pod.name, pod.uid)],
                )
                changed = True
        return changed

    def settle(self, limit=50):
        """gomega.Eventually: run both controllers to a fixed point."""
        for _ in range(limit):
            a = self.reconcile_lws()
            b = self.reconcile_pods()
            if not a and not b:
                return
        raise AssertionError("controllers did not settle")

    # ----------------------------------------- test/testutils/util.go helpers
    def create_leader_pods(self, start, end, rev_key=None):
        """CreateLeaderPods (:140-147,:230-243): pods of the *current* template revision."""
        for i in range(start, end):
            name = f"{self.lws.name}-{i}"
            self.pods[name] = api.Pod(
                name=name,
                namespace=self.lws.namespace,
                uid=f"uid-pod-{name}-{self.template_rev}-{len(self.pods)}",
                labels={
                    api.SetNameLabelKey: self.lws.name,
                    api.WorkerIndexLabelKey: "0",
                    api.GroupIndexLabelKey: str(i),
                    api.RevisionKey: rev_key or self.rev_key,
                },
                annotations={api.SizeAnnotationKey: str(self.lws.size)},
                ownerReferences=[api.OwnerReference("StatefulSet", self.lws.name, self.leader_sts.uid)],
            )
        self.settle()

    def set_pod_group_ready(self, idx):
        """SetPodGroupToReady (:386-401) = SetLeaderPodToReady (:331-361) + sts status."""
        name = f"{self.lws.name}-{idx}"
        pod = self.pods[name]
        pod.labels[api.RevisionKey] = self.leader_sts.labels[api.RevisionKey]
        pod.phase = "Running"
        pod.readyCondition = True
        if self.lws.size > 1:
            self.stss.pop(name, None)  # deleteWorkerStatefulSetIfExists → pod controller recreates it
            self.reconcile_pods()
            sts = self.stss[name]
            sts.availableReplicas = sts.replicas
            sts.statusReplicas = sts.replicas
            sts.currentRevision = sts.updateRevision = ""
        self.settle()

    def set_all_ready(self):
        """SetSuperPodToReady (:304-329)."""
        for name in sorted(self.pods):
            self.set_pod_group_ready(int(name.rsplit("-", 1)[1]))

    def set_sts_unready(self, idx):
        """SetStatefulsetToUnReady (:404-411)."""
        sts = self.stss[f"{self.lws.name}-{idx}"]
        sts.currentRevision, sts.updateRevision = "fuz", "bar"
        self.settle()

    def delete_leader_pod(self, start, end):
        """DeleteLeaderPod (:126-138)."""
        for i in range(start, end):
            name = f"{self.lws.name}-{i}"
            del self.pods[name]
            self.stss.pop(name, None)
        self.settle()

    def delete_leader_pods_above_replicas(self):
        """DeleteLeaderPods (:101-124)."""
        for name in list(self.pods):
            if int(name.rsplit("-", 1)[1]) >= self.lws.replicas:
                del self.pods[name]
                self.stss.pop(name, None)
        self.settle()

    def create_worker_pods(self, group):
        """CreateWorkerPodsForLeaderPod (test/testutils/util.go:58-88): workers 1..size-1 of the group,
        owned by the leader Pod, carrying the leader pod's current revision key."""
        leader = self.pods[f"{self.lws.name}-{group}"]
        for w in range(1, self.lws.size):
            name = f"{leader.name}-{w}"
            self.workers[name] = api.Pod(
                name=name,
                namespace=self.lws.namespace,
                labels={
                    api.SetNameLabelKey: self.lws.name,
                    api.WorkerIndexLabelKey: str(w),
                    api.GroupIndexLabelKey: leader.labels[api.GroupIndexLabelKey],
                    api.RevisionKey: leader.labels[api.RevisionKey],
                },
                annotations={api.SizeAnnotationKey: str(self.lws.size)},
                phase="Running",
                ownerReferences=[api.OwnerReference("Pod", leader.name, leader.uid)],
            )
        self.settle()

    def delete_worker_pods(self, names=None):
        """k8sClient.Delete on worker pods: the pod controller sees them with a deletionTimestamp
        (PodDeleted, pod_utils.go:48); afterwards they are gone."""
        names = list(self.workers) if names is None else list(names)
        for n in names:
            self.workers[n].deletionTimestamp = True
        self.settle()
        for n in names:
            del self.workers[n]
        self.settle()

    # ------------------------------------------------------------ lws edits
    def update_template(self):
        self.template_rev += 1
        self.settle()

    def set_replicas(self, n, settle=True):
        self.lws.replicas = n
        if settle:
            self.settle()

    def set_partition(self, p):
        self.lws.rollingUpdate.partition = p
        self.settle()

    # ------------------------------------------------------------ observers
    def state(self):
        return (
            self.leader_sts.partition,
            self.leader_sts.replicas,
            self.status["readyReplicas"],
            self.status["updatedReplicas"],
        )

    def bootstrap(self):
        """Create the LWS, let the controller create the leader sts, create and
        ready every group — the common preamble of the integration entries."""
        self.settle()
        self.create_leader_pods(0, self.lws.replicas)
        self.set_all_ready()
        return self
