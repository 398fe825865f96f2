"""Restart-policy vectors of the reference, run through encoder + oracle.

  pkg/controllers/pod_controller_test.go:427-532   current vs stale worker-sts owner
  test/integration/controllers/leaderworkerset_test.go:408-563   None / OnPodRestart /
      annotation+pending / AfterStart pending / AfterStart running (workers owned by the
      leader Pod, test/testutils/util.go:58-88)
  pkg/utils/pod/pod_utils_test.go:28-101            ContainerRestarted
  pkg/utils/statefulset/statefulset_utils_test.go:26-153  GetParentNameAndOrdinal, StatefulsetReady
"""
import pytest

from lws_b200 import api, encoder
from lws_b200 import records as R


def labels(lws, group, worker, rev="revision-1"):
    return {api.SetNameLabelKey: lws.name, api.WorkerIndexLabelKey: str(worker), api.GroupIndexLabelKey: str(group),
            api.RevisionKey: rev}


def group_flags(oracle_sweep, lws, pods, stss):
    item = encoder.LwsItem(lws=lws, revision_key="revision-1",
                           leader_sts=api.StatefulSet(name=lws.name, replicas=lws.replicas,
                                                      annotations={api.ReplicasAnnotationKey: str(lws.replicas)}))
    t = encoder.encode_lws([item], encoder.Cluster(pods=pods, statefulsets=stss))
    assert not t.lws[0]["flags"] & R.LWS_IRREGULAR
    _, go = oracle_sweep(t)
    return int(go[0]["flags"]), int(go[0]["first_trigger"]), t


# pod_controller_test.go:427-532
@pytest.mark.parametrize("owner_uid,want_deleted", [("sts-current", True), ("sts-stale", False)])
def test_handle_restart_policy_uses_current_worker_ownership(oracle_sweep, owner_uid, want_deleted):
    lws = api.LeaderWorkerSet("test-sample", replicas=1, size=2, restartPolicy=api.RecreateGroupOnPodRestart)
    leader = api.Pod("test-sample-0", uid="leader-current", labels=labels(lws, 0, 0))
    sts = api.StatefulSet("test-sample-0", uid="sts-current", labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: "0"},
                          ownerReferences=[api.OwnerReference("Pod", leader.name, leader.uid)])
    worker = api.Pod("test-sample-0-1", labels=labels(lws, 0, 1), deletionTimestamp=True,
                     ownerReferences=[api.OwnerReference("StatefulSet", "test-sample-0", owner_uid)])
    flags, first, t = group_flags(oracle_sweep, lws, [leader, worker], [sts])
    assert bool(flags & R.GOUT_DELETE_LEADER) == want_deleted
    assert not flags & (R.GOUT_LEADER_DELETING | R.GOUT_RESTART_ERROR)
    if want_deleted:
        assert t.group_pod_names[0][first] == "test-sample-0-1"


def _group_of_four(lws, pending=None, running=False, deleting="test-sample-0-1"):
    leader = api.Pod("test-sample-0", labels=labels(lws, 0, 0), phase="Running" if running else "")
    pods = [leader]
    for i in range(1, 4):
        name = f"test-sample-0-{i}"
        pods.append(api.Pod(name, labels=labels(lws, 0, i), phase="Pending" if name == pending else ("Running" if running else ""),
                            deletionTimestamp=(name == deleting),
                            ownerReferences=[api.OwnerReference("Pod", leader.name, leader.uid)]))
    return pods


# test/integration/controllers/leaderworkerset_test.go:408-563
@pytest.mark.parametrize(
    "policy,annot,pending,running,want",
    [(api.NoneRestartPolicy, False, None, False, False),
     (api.RecreateGroupOnPodRestart, False, None, False, True),
     (api.RecreateGroupOnPodRestart, True, "test-sample-0-2", False, False),
     (api.RecreateGroupAfterStart, False, "test-sample-0-2", False, False),
     (api.RecreateGroupAfterStart, False, None, True, True)],
)
def test_integration_restart_policies(oracle_sweep, policy, annot, pending, running, want):
    lws = api.LeaderWorkerSet("test-sample", replicas=1, size=4, restartPolicy=policy,
                              annotations={api.RecreateGroupAfterStartAnnotationKey: "true"} if annot else {})
    flags, first, t = group_flags(oracle_sweep, lws, _group_of_four(lws, pending, running), [])
    assert bool(flags & R.GOUT_DELETE_LEADER) == want
    if want:  # "Worker pod test-sample-0-1 failed, deleted leader pod test-sample-0 to recreate group 0"
        assert t.group_pod_names[0][first] == "test-sample-0-1"
    assert bool(flags & R.GOUT_PENDING) == (pending is not None)


def test_leader_already_deleting_returns_true_without_delete(oracle_sweep):
    """pod_controller.go:255-257."""
    lws = api.LeaderWorkerSet("test-sample", replicas=1, size=4)
    pods = _group_of_four(lws)
    pods[0].deletionTimestamp = True
    flags, _, _ = group_flags(oracle_sweep, lws, pods, [])
    assert flags & R.GOUT_LEADER_DELETING and not flags & R.GOUT_DELETE_LEADER
    assert not flags & R.GOUT_CREATE_WSTS  # :125 leader deleting → no worker sts


# pkg/utils/pod/pod_utils_test.go:28-101 ContainerRestarted
@pytest.mark.parametrize(
    "phase,init,main,want",
    [("Running", [], [1], True), ("Pending", [2], [0], True), ("Running", [], [0], False),
     ("Failed", [], [3], False), ("Succeeded", [1], [1], False)],
)
def test_container_restarted(oracle_sweep, phase, init, main, want):
    lws = api.LeaderWorkerSet("test-sample", replicas=1, size=2)
    leader = api.Pod("test-sample-0", labels=labels(lws, 0, 0), phase="Running")
    worker = api.Pod("test-sample-0-1", labels=labels(lws, 0, 1), phase=phase, initContainerRestartCounts=init,
                     containerRestartCounts=main, ownerReferences=[api.OwnerReference("Pod", leader.name, leader.uid)])
    flags, _, _ = group_flags(oracle_sweep, lws, [leader, worker], [])
    assert bool(flags & R.GOUT_DELETE_LEADER) == want


# pkg/utils/statefulset/statefulset_utils_test.go:26-153
@pytest.mark.parametrize(
    "name,parent,ordinal",
    [  # the reference's 7 cases (statefulset_utils_test.go:26-75) …
     ("lws-samples-132", "lws-samples", 132), ("lws-samples-132-u", "", -1), ("lws-samples-", "", -1),
     ("lws-samples-0", "lws-samples", 0), ("lws-samples--1", "lws-samples-", 1), ("lws-samples1", "", -1),
     ("lws-samples-1-0", "lws-samples-1", 0),
     # … plus the int32 overflow rule of strconv.ParseInt(…, 10, 32)
     ("x-99999999999999", "x", -1)],
)
def test_get_parent_name_and_ordinal(name, parent, ordinal):
    assert encoder.get_parent_name_and_ordinal(name) == (parent, ordinal)


# pkg/schedulerprovider/volcano_provider_test.go:49-139: PodGroup MinMember = size (LeaderCreated) or 1 (LeaderReady)
@pytest.mark.parametrize("startup,want", [(api.LeaderCreatedStartupPolicy, 3), (api.LeaderReadyStartupPolicy, 1)])
def test_podgroup_min_member(oracle_sweep, startup, want):
    lws = api.LeaderWorkerSet("test-lws", replicas=1, size=3, startupPolicy=startup)
    leader = api.Pod("test-lws-0", labels=labels(lws, 0, 0), phase="Running")
    item = encoder.LwsItem(lws=lws, revision_key="revision-1",
                           leader_sts=api.StatefulSet(name=lws.name, replicas=1, annotations={api.ReplicasAnnotationKey: "1"}))
    t = encoder.encode_lws([item], encoder.Cluster(pods=[leader]))
    lo, go = oracle_sweep(t, R.SWEEP_GANG)
    assert int(lo[0]["min_member"]) == want
    assert go[0]["flags"] & R.GOUT_CREATE_PODGROUP  # pod_controller.go:130 reached for the leader pod
    lo, go = oracle_sweep(t, 0)  # no SchedulerProvider configured
    assert int(lo[0]["min_member"]) == 0 and not go[0]["flags"] & R.GOUT_CREATE_PODGROUP


# pkg/utils/utils_test.go:25-64 SortByIndex: out-of-range and unparsable indices are dropped,
# the last writer wins — observable through the encoder's row order
def test_sort_by_index_semantics_in_encoder():
    lws = api.LeaderWorkerSet("s", replicas=3, size=1)
    mk = lambda name, idx: api.Pod(name, labels={api.SetNameLabelKey: "s", api.WorkerIndexLabelKey: "0",
                                                  api.GroupIndexLabelKey: idx}, phase="Running", readyCondition=True)
    pods = [mk("s-2", "2"), mk("s-0", "0"), mk("s-x", "abc")]
    item = encoder.LwsItem(lws=lws, revision_key="r", leader_sts=api.StatefulSet(name="s", replicas=3, annotations={api.ReplicasAnnotationKey: "3"}))
    t = encoder.encode_lws([item], encoder.Cluster(pods=pods))
    assert len(t.groups) == 3  # slots 0..2; slot 1 is a zero-valued entry
    present = [(int(g["flags"]) & R.GRP_POD_PRESENT) != 0 for g in t.groups]
    assert present == [True, False, True]
    assert t.lws[0]["flags"] & R.LWS_GROUP_LABEL_INVALID  # updateConditions would fail on "abc" (:434)


# pkg/utils/statefulset/statefulset_utils_test.go:81-153 StatefulsetReady: availableReplicas == *spec.replicas
# and currentRevision == updateRevision — observable as the group's ready state (a ready leader pod
# with a ready worker sts, leaderworkerset_controller.go:620-627)
@pytest.mark.parametrize("available,current,update,want",
                         [(3, "rev-1", "rev-1", True), (2, "rev-1", "rev-1", False),
                          (3, "rev-1", "rev-2", False), (2, "rev-1", "rev-2", False)])
def test_statefulset_ready(oracle_sweep, available, current, update, want):
    lws = api.LeaderWorkerSet("test-lws", replicas=1, size=4)
    leader = api.Pod("test-lws-0", labels=labels(lws, 0, 0), phase="Running", readyCondition=True)
    wsts = api.StatefulSet("test-lws-0", replicas=3, availableReplicas=available, currentRevision=current,
                           updateRevision=update,
                           labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: "0", api.RevisionKey: "revision-1"})
    flags, _, _ = group_flags(oracle_sweep, lws, [leader], [wsts])
    assert bool(flags & R.GOUT_STATE_READY) == want
    assert bool(flags & R.GOUT_COND_READY) == want
    assert flags & R.GOUT_STATE_UPDATED  # revisions of the leader pod and of the sts template both match


def test_podgroup_min_resources():
    """volcano_provider_test.go:65-139: three pods requesting 100m CPU each → MinResources cpu 300m, for
    both startup policies (test/testutils/util.go:846-865: the same for LeaderCreated and LeaderReady)."""
    worker = {"cpu": 100}
    assert encoder.pod_group_min_resources(None, worker, 3) == {"cpu": 300}
    assert encoder.pod_group_min_resources({"cpu": 100}, worker, 3) == {"cpu": 300}
    # a leader template of its own, a resource only the workers ask for, size 1 (no workers)
    assert encoder.pod_group_min_resources({"cpu": 500, "memory": 64}, {"cpu": 100, "nvidia.com/gpu": 8000}, 4) == {
        "cpu": 800, "memory": 64, "nvidia.com/gpu": 24000}
    assert encoder.pod_group_min_resources({"cpu": 500}, worker, 1) == {"cpu": 500}


# pkg/controllers/pod_controller_test.go:42-425 TestConstructWorkerStatefulSetApplyConfiguration: the worker
# StatefulSet has size − 1 replicas and ordinals starting at 1 (:434-443); Reconcile only builds it for
# size > 1 (:138).  (Labels, templates and volume claims of that test are object construction.)
@pytest.mark.parametrize("size,want_create,want_replicas", [(1, False, 0), (2, True, 1), (5, True, 4)])
def test_worker_statefulset_replicas_and_gate(oracle_sweep, size, want_create, want_replicas):
    lws = api.LeaderWorkerSet("test-sample", replicas=2, size=size)
    leader = api.Pod("test-sample-1", labels=labels(lws, 1, 0), phase="Running", readyCondition=True)
    item = encoder.LwsItem(lws=lws, revision_key="revision-1",
                           leader_sts=api.StatefulSet(name=lws.name, replicas=2, annotations={api.ReplicasAnnotationKey: "2"}))
    t = encoder.encode_lws([item], encoder.Cluster(pods=[leader]))
    _, go = oracle_sweep(t)
    assert bool(go[1]["flags"] & R.GOUT_CREATE_WSTS) == want_create
    assert int(go[1]["worker_replicas"]) == want_replicas  # pods <lws>-1-1 … <lws>-1-(size-1)
    assert not go[0]["flags"] & R.GOUT_CREATE_WSTS  # group 0 has no leader pod yet
