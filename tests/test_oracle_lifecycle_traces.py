"""More entries of test/integration/controllers/leaderworkerset_test.go replayed through tests/sim.py —
the scale / create / startup-policy / condition entries (the rolling-update family is tests/traces.py):

  :90   scale up number of groups            :109  scale down number of groups
  :128  scale down to 0                      :147  scale up from 0
  :166  group size is 1                      :187  zero replicas
  :200  2 groups, size 2                     :215  deleted worker StatefulSet is recreated
  :346  available state                      :359  progressing → available → progressing
  :2091 startupPolicy LeaderReady            :2120 startupPolicy LeaderCreated
  :1878 a not-yet-updated group that restarts during an update gets its worker sts back with the OLD spec
  :2575 PodGroup per leader pod, MinMember = size; gang rolling update 2 → 1 → 0
  :395  exclusive placement: worker StatefulSets wait for the leader pod to be scheduled; topology value
  :2277 resize (size 3 → 4) rolls every group and resizes the worker StatefulSets
  :1964 RecreateGroupOnPodRestart during a rolling update: deleting the workers of the OLD revision does not
        delete the already updated leader (handleRestartPolicy compares revisions, pod_controller.go:239);
        after the update a worker deletion recreates the group

What the reference asserts there and what is checked here: the leader StatefulSet's replica count
(ExpectValidLeaderStatefulSet, test/testutils/validators.go), one worker StatefulSet of size − 1 replicas
per existing leader pod and none for size 1 (ExpectValidWorkerStatefulSets / …NotCreated), and the
LWS condition (ExpectLeaderWorkerSetAvailable / Progressing).
"""
import pytest

from lws_b200 import api
from lws_b200 import records as R
from sim import LwsSim

AV, PR = R.COND_AVAILABLE, R.COND_PROGRESSING


def build(replicas=2, size=2, startup=api.LeaderCreatedStartupPolicy):  # wrappers.BuildLeaderWorkerSet defaults
    return api.LeaderWorkerSet(name="test-sample", replicas=replicas, size=size, startupPolicy=startup,
                               restartPolicy=api.RecreateGroupOnPodRestart,
                               rollingUpdate=api.RollingUpdateConfiguration(partition=0, maxUnavailable=1, maxSurge=0))


def start(sweep, **kw):
    """The table's preamble: create the LWS, wait for the leader sts, create the leader pods."""
    sim = LwsSim(build(**kw), sweep)
    sim.settle()
    sim.create_leader_pods(0, sim.lws.replicas)
    return sim


def expect_valid(sim, replicas):
    assert sim.leader_sts.replicas == replicas
    want = {name for name in sim.pods} if sim.lws.size > 1 else set()
    assert set(sim.stss) == want, (sorted(sim.stss), sorted(want))
    for sts in sim.stss.values():
        assert sts.replicas == sim.lws.size - 1


def set_leader_pods_ready(sim, a, b):
    """SetLeaderPodsToReady (test/testutils/util.go:693-719): leader pods only, no worker sts status."""
    for i in range(a, b):
        pod = sim.pods[f"{sim.lws.name}-{i}"]
        pod.phase, pod.readyCondition = "Running", True
    sim.settle()


def run_lifecycle_entries(sweep):
    # :90 scale up number of groups
    sim = start(sweep, replicas=2)
    sim.set_replicas(3)
    sim.create_leader_pods(2, 3)
    expect_valid(sim, 3)
    # :109 scale down number of groups
    sim = start(sweep, replicas=4)
    sim.set_replicas(3)
    sim.delete_leader_pods_above_replicas()
    expect_valid(sim, 3)
    # :128 scale down to 0
    sim = start(sweep, replicas=2)
    sim.set_replicas(0)
    sim.delete_leader_pods_above_replicas()
    expect_valid(sim, 0)
    assert not sim.pods and not sim.stss
    # :147 scale up from 0
    sim = start(sweep, replicas=0)
    expect_valid(sim, 0)
    sim.set_replicas(3)
    sim.create_leader_pods(0, 3)
    expect_valid(sim, 3)
    # :166 group size is 1: no worker StatefulSets; ready leaders alone make it available
    sim = start(sweep, size=1)
    expect_valid(sim, 2)
    set_leader_pods_ready(sim, 0, 2)
    assert sim.status["condition"] == AV and sim.status["readyReplicas"] == 2
    # :187 zero replicas
    sim = start(sweep, replicas=0)
    expect_valid(sim, 0)
    # :200 two groups of size 2: a worker sts per leader pod
    sim = start(sweep)
    expect_valid(sim, 2)
    assert sorted(sim.stss) == ["test-sample-0", "test-sample-1"]
    # :215 a deleted worker StatefulSet is recreated
    del sim.stss["test-sample-0"]
    sim.settle()
    expect_valid(sim, 2)
    # :346 available state
    sim = start(sweep)
    sim.set_all_ready()
    assert sim.status["condition"] == AV
    # :359 progressing → available → progressing (0 of 2 ready, 2 of 2, then 2 of 3 after a scale-up)
    sim = start(sweep)
    assert sim.status["condition"] == PR and sim.status["readyReplicas"] == 0
    sim.set_all_ready()
    assert sim.status["condition"] == AV and sim.status["readyReplicas"] == 2
    sim.set_replicas(3)
    assert sim.status["condition"] == PR and sim.status["readyReplicas"] == 2
    # :2091 startupPolicy LeaderReady: worker StatefulSets only for leaders that are ready
    sim = start(sweep, replicas=4, startup=api.LeaderReadyStartupPolicy)
    assert not sim.stss
    set_leader_pods_ready(sim, 0, 2)
    assert sorted(sim.stss) == ["test-sample-0", "test-sample-1"]
    set_leader_pods_ready(sim, 2, 4)
    assert sorted(sim.stss) == [f"test-sample-{i}" for i in range(4)]
    # :2120 startupPolicy LeaderCreated: all of them at once
    sim = start(sweep, replicas=4, startup=api.LeaderCreatedStartupPolicy)
    assert sorted(sim.stss) == [f"test-sample-{i}" for i in range(4)]


    # :1878-1962 restart of a not-yet-updated group during a rolling update
    UP = R.COND_UPDATE_IN_PROGRESS
    rev = lambda sim, i: sim.stss[f"test-sample-{i}"].labels[api.RevisionKey]
    sim = start(sweep, replicas=4)
    sim.set_all_ready()
    assert sim.status["condition"] == AV and (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 4)
    expect_valid(sim, 4)
    sim.update_template()  # UpdateWorkerTemplate
    assert sim.status["condition"] == UP and (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 0)
    sim.set_pod_group_ready(3)
    assert (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 1) and sim.status["condition"] == UP
    assert rev(sim, 3) == "rev-2" and [rev(sim, i) for i in (2, 1, 0)] == ["rev-1"] * 3
    sim.delete_leader_pod(0, 1)
    sim.create_leader_pods(0, 1, rev_key="rev-1")  # CreateLeaderPodsFromRevisionNumber(…, 1): the old template
    expect_valid(sim, 4)
    assert rev(sim, 3) == "rev-2" and [rev(sim, i) for i in (2, 1, 0)] == ["rev-1"] * 3  # old spec again
    assert sim.status["condition"] == UP
    sim.set_all_ready()
    expect_valid(sim, 4)
    assert sim.status["condition"] == AV and all(rev(sim, i) == "rev-2" for i in range(4))


    # :1964-2054 the leader is only restarted once during a rolling update
    sim = start(sweep, replicas=4)
    sim.set_all_ready()
    sim.create_worker_pods(3)
    assert sim.status["condition"] == AV and (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 4)
    sim.update_template()
    assert sim.status["condition"] == UP and (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 0)
    sim.set_pod_group_ready(3)  # the leader pod of group 3 now carries the new revision
    assert (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (4, 1)
    sim.delete_worker_pods()  # old-revision workers go away: this must not delete the leader
    assert not sim.pods["test-sample-3"].deletionTimestamp and not sim.deleted_leaders
    sim.set_all_ready()
    sim.create_worker_pods(3)
    expect_valid(sim, 4)
    assert sim.status["condition"] == AV
    sim.delete_worker_pods(["test-sample-3-1"])  # same revision now: the group is recreated
    assert sim.pods["test-sample-3"].deletionTimestamp and sim.deleted_leaders == ["test-sample-3"]


def run_more_entries(sweep):
    """Entries added after the GPU budget of the round was spent: oracle only for now."""
    # :2277-2310 resize should update the size of the replicas (a size change is a template change)
    sim = start(sweep, replicas=2, size=3)
    sim.set_all_ready()
    assert sim.status["condition"] == AV and sim.leader_sts.partition == 0
    expect_valid(sim, 2)
    assert (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (2, 2)
    sim.lws.size = 4  # UpdateSize
    sim.update_template()
    sim.set_all_ready()
    assert sim.status["condition"] == AV and sim.leader_sts.partition == 0
    expect_valid(sim, 2)  # worker StatefulSets now have size - 1 = 3 replicas
    assert all(sts.replicas == 3 for sts in sim.stss.values())
    assert (sim.status["readyReplicas"], sim.status["updatedReplicas"]) == (2, 2)


    # :2575-2660 gang scheduling (a SchedulerProvider is configured): a PodGroup per leader pod with
    # MinMember = size, and the rolling update walks the partition 2 → 1 → 0 with 3 replicas
    UP = R.COND_UPDATE_IN_PROGRESS

    def pod_groups_ok(sim, n):
        sim.reconcile_pods()
        go = sim.last_group_out
        assert int(sim.last_lws_out[0]["min_member"]) == sim.lws.size
        assert [bool(go[i]["flags"] & R.GOUT_CREATE_PODGROUP) for i in range(n)] == [True] * n

    sim = LwsSim(build(replicas=2), sweep, gang=True)
    sim.settle()
    sim.create_leader_pods(0, 2)
    pod_groups_ok(sim, 2)
    sim = LwsSim(build(replicas=3), sweep, gang=True)
    sim.settle()
    sim.create_leader_pods(0, 3)
    sim.set_all_ready()
    assert sim.status["condition"] == AV and sim.leader_sts.partition == 0
    pod_groups_ok(sim, 3)
    sim.update_template()
    assert sim.status["condition"] == UP and sim.leader_sts.partition == 2
    sim.set_pod_group_ready(2)
    assert sim.leader_sts.partition == 1
    sim.set_pod_group_ready(1)
    assert sim.leader_sts.partition == 0
    sim.set_pod_group_ready(0)
    assert sim.status["condition"] == AV
    pod_groups_ok(sim, 3)


    # :395-406 exclusive placement: no worker StatefulSet until the leader pod is scheduled
    # (validators.go:224-226: "only expect sts count to be 1"), then one per leader, carrying the
    # topology value of the leader's node as nodeSelector (pod_controller.go:162-172, :297-336)
    key = "cloud.google.com/gke-nodepool"
    nodes = [api.Node("node-a", labels={key: "pool-1"}), api.Node("node-b", labels={key: "pool-2"}),
             api.Node("node-nolabel")]
    lws = build()
    lws.annotations[api.ExclusiveKeyAnnotationKey] = key
    sim = LwsSim(lws, sweep, nodes=nodes, topology_key=key)
    sim.settle()
    sim.create_leader_pods(0, 2)
    assert sim.leader_sts.replicas == 2 and not sim.stss  # leaders are not scheduled yet
    assert all(sim.last_group_out[i]["flags"] & R.GOUT_WAIT_SCHEDULE for i in range(2))
    sim.pods["test-sample-0"].nodeName = "node-a"
    sim.pods["test-sample-1"].nodeName = "node-b"
    sim.settle()
    assert sorted(sim.stss) == ["test-sample-0", "test-sample-1"]
    node_dom = {n: int(sim.last_tables.nodes["domain_id"][k]) for k, n in enumerate(["node-a", "node-b", "node-nolabel"])}
    # (the domain is reported together with CREATE_WSTS, i.e. while the worker sts does not exist yet)
    # pod_controller.go:330: a node without the topology label is an error — no worker sts for that group
    sim.stss.clear()
    sim.pods["test-sample-1"].nodeName = "node-nolabel"
    sim._sweep()
    go = sim.last_group_out
    assert go[0]["flags"] & R.GOUT_CREATE_WSTS and int(go[0]["domain_id"]) == node_dom["node-a"]
    assert go[1]["flags"] & R.GOUT_TOPOLOGY_ERROR and not go[1]["flags"] & R.GOUT_CREATE_WSTS
    # :327 a node that does not exist yields an empty value and no error: the worker sts is created
    sim.pods["test-sample-1"].nodeName = "node-gone"
    sim._sweep()
    go = sim.last_group_out
    assert go[1]["flags"] & R.GOUT_CREATE_WSTS and not go[1]["flags"] & R.GOUT_TOPOLOGY_ERROR
    assert int(go[1]["domain_id"]) == R.NONE


def test_lifecycle_entries_on_the_oracle(oracle_sweep):
    run_lifecycle_entries(oracle_sweep)
    run_more_entries(oracle_sweep)
