"""Golden traces of the LWS rolling-update path, transcribed from the reference's
integration table test/integration/controllers/leaderworkerset_test.go (line
ranges per trace).  Each step = (action, expected) where expected is
(sts partition, sts replicas, status.readyReplicas, status.updatedReplicas,
condition) with None = not asserted by the reference at that step.

The actions are the reference's own test helpers (tests/sim.py mirrors them).
"""
from __future__ import annotations

from lws_b200 import api
from lws_b200 import records as R

AV, PR, UP = R.COND_AVAILABLE, R.COND_PROGRESSING, R.COND_UPDATE_IN_PROGRESS


def _lws(replicas, mu=1, ms=0, partition=0, size=2):
    return dict(replicas=replicas, maxUnavailable=mu, maxSurge=ms, partition=partition, size=size)


def rdy(*idx):
    return lambda s: [s.set_pod_group_ready(i) for i in idx]


def all_rdy(s):
    s.set_all_ready()


def update(s):
    s.update_template()


def unready(i):
    return lambda s: s.set_sts_unready(i)


def create(a, b):
    return lambda s: s.create_leader_pods(a, b)


def create_stale(a, b):
    """CreateLeaderPods called in the same lwsUpdateFn as a template edit: the helper is
    handed the *pre-edit* lws object (leaderworkerset_test.go:76-80 fetches it before the
    fn runs), so NewRevision() yields the previous template's revision key."""
    return lambda s: s.create_leader_pods(a, b, rev_key=f"rev-{s.template_rev - 1}")


def delete(a, b):
    return lambda s: s.delete_leader_pod(a, b)


def replicas(n):
    return lambda s: s.set_replicas(n)


def partition(p):
    return lambda s: s.set_partition(p)


def seq(*fns):
    def run(s):
        for f in fns:
            f(s)

    return run


def update_and_replicas(n):
    """UpdateLeaderTemplate + UpdateReplicaCount in one lwsUpdateFn: both edits land
    before the controller has settled on either (one object update each, but the
    checks only look at the end state)."""

    def run(s):
        s.template_rev += 1
        s.lws.replicas = n
        s.settle()

    return run


def del_surge(s):
    s.delete_leader_pods_above_replicas()


def replicas_then(n, *fns):
    """UpdateReplicaCount followed by more helper calls in the same lwsUpdateFn."""

    def run(s):
        s.set_replicas(n)
        for f in fns:
            f(s)

    return run


TRACES = {
    # :631-727 leaderTemplate changed with default strategy
    "T1": (_lws(4), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (update, (3, 4, 4, 0, UP)),
        (rdy(3), (2, 4, 4, 1, UP)),
        (rdy(1), (2, 4, 4, 2, UP)),
        (unready(3), (2, 4, 3, 2, UP)),
        (all_rdy, (0, 4, 4, 4, AV)),
    ]),
    # :729-805 workerTemplate changed with maxUnavailable=2
    "T2": (_lws(4, mu=2), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (update, (2, 4, 4, 0, UP)),
        (rdy(3), (1, 4, 4, 1, UP)),
        (rdy(2), (0, 4, 4, 2, UP)),
        (rdy(1, 0), (0, 4, 4, 4, AV)),
    ]),
    # :807-854 maxUnavailable greater than replicas
    "T3": (_lws(4, mu=10), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (update, (0, 4, 4, 0, UP)),
        (all_rdy, (0, 4, 4, 4, AV)),
    ]),
    # :1088-1205 rolling update with maxSurge set, maxUnavailable=0
    "T7": (_lws(4, mu=0, ms=1), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update, create_stale(4, 5)), (4, 5, 4, 0, UP)),
        (rdy(4), (3, 5, 5, 1, UP)),
        (rdy(3), (2, 5, 5, 2, UP)),
        (rdy(2), (1, 5, 5, 3, UP)),
        (rdy(1), (0, 5, 5, 4, UP)),
        (seq(rdy(0), del_surge), (0, 4, 4, 4, AV)),
    ]),
    # :1207-1324 rolling update with maxSurge set (maxUnavailable=1)
    "T8": (_lws(4, mu=1, ms=1), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update, create_stale(4, 5)), (3, 5, 4, 0, UP)),
        (rdy(4), (2, 5, 5, 1, UP)),
        (rdy(3), (1, 5, 5, 2, UP)),
        (rdy(2), (0, 5, 5, 3, UP)),
        (seq(rdy(1), del_surge), (0, 4, 4, 3, UP)),
        (rdy(0), (0, 4, 4, 4, AV)),
    ]),
    # :856-914 rolling update with both worker template and number of replicas changed
    "T4": (_lws(4), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update_and_replicas(6), create_stale(4, 6)), (4, 6, 4, 0, UP)),
        (all_rdy, (0, 6, 6, 6, AV)),
    ]),
    # :916-1006 replicas increases during rolling update
    "T5": (_lws(4), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (update, (3, 4, 4, 0, UP)),
        (rdy(3), (2, 4, 4, 1, UP)),
        (replicas_then(6, create(4, 6), rdy(4), rdy(5)), (2, 6, 6, 3, UP)),
        (all_rdy, (0, 6, 6, 6, AV)),
    ]),
    # :1008-1086 replicas decreases during rolling update
    "T6": (_lws(6), [
        (all_rdy, (0, 6, 6, 6, AV)),
        (update, (5, 6, 6, 0, UP)),
        (replicas_then(3, del_surge), (2, 3, 3, 0, UP)),
        (all_rdy, (0, 3, 3, 3, AV)),
    ]),
    # :1326-1402 rolling update with maxUnavailable and maxSurge set
    "T9": (_lws(4, mu=2, ms=2), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update, create_stale(4, 6)), (2, 6, 4, 0, UP)),
        (seq(rdy(3), rdy(2), delete(4, 6)), (0, 4, 4, 2, None)),
        (rdy(1, 0), (0, 4, 4, 4, AV)),
    ]),
    # :1404-1471 rolling update with replicas scaled up and maxSurge set
    "T10": (_lws(2, ms=2), [
        (all_rdy, (0, 2, 2, 2, AV)),
        (seq(update_and_replicas(4), create_stale(2, 6)), (2, 6, None, None, UP)),
        (all_rdy, (None, None, None, None, AV)),
        (del_surge, (0, 4, 4, 4, AV)),
    ]),
    # :1473-1537 rolling update with replicas scaled down and maxSurge set
    "T11": (_lws(6, ms=2), [
        (all_rdy, (0, 6, 6, 6, AV)),
        (update_and_replicas(3), (2, 5, None, None, UP)),
        (seq(delete(5, 6), all_rdy), (None, None, None, None, AV)),
        (del_surge, (0, 3, 3, 3, AV)),
    ]),
    # :1539-1607 rolling update with maxSurge greater than replicas
    "T12": (_lws(2, ms=4), [
        (all_rdy, (0, 2, 2, 2, AV)),
        (seq(update, create_stale(2, 3)), (1, 3, None, None, UP)),
        (all_rdy, (None, None, None, None, AV)),
        (del_surge, (0, 2, 2, 2, AV)),
    ]),
    # :1609-1764 scale up and down during rolling update with maxSurge set
    "T13": (_lws(4, ms=2), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update, create_stale(4, 6)), (3, 6, 4, 0, UP)),
        (replicas_then(6, create(6, 8)), (3, 8, 4, 2, UP)),
        (rdy(7, 6), (3, 8, 6, 2, UP)),
        (replicas_then(2, delete(4, 8), delete(3, 4)), (1, 3, 3, 0, UP)),
        (rdy(2), (0, 3, 3, 1, UP)),
        (seq(rdy(1), delete(2, 3)), (0, 2, 2, 1, UP)),
        (rdy(0), (0, 2, 2, 2, AV)),
    ]),
    # :1766-1876 multiple rolling update with maxSurge set
    "T14": (_lws(4, ms=2), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (seq(update, create_stale(4, 6)), (3, 6, 4, 0, UP)),
        (rdy(5, 4), (1, 6, 6, 2, UP)),
        (update, (3, 6, 6, 0, UP)),
        (all_rdy, (None, None, None, None, AV)),
        (del_surge, (0, 4, 4, 4, AV)),
    ]),
    # :2199-2275 rolling update with no ready replicas
    "T16": (_lws(2, mu=1), [
        (all_rdy, (0, 2, 2, 2, AV)),
        (unready(0), (0, None, None, None, PR)),
        (unready(1), (0, None, None, None, PR)),
        (update, (1, 2, 0, 0, UP)),
        (rdy(1), (0, None, None, None, None)),
        (rdy(0), (None, None, None, None, AV)),
    ]),
    # :2408-2496 rolling update with the partition and maxSurge setting
    "T18": (_lws(3, ms=1, partition=2), [
        (all_rdy, (2, 3, 3, 3, AV)),
        (update, (2, 4, 3, 0, UP)),
        (seq(create(3, 4), rdy(3), rdy(2)), (2, 4, 4, 2, AV)),
        (partition(0), (0, 4, 4, 2, UP)),
        (rdy(1, 0), (0, 3, 4, 4, AV)),
        (delete(3, 4), (0, 3, 3, 3, AV)),
    ]),
    # :2132-2197 unready replica below the partition counts as unavailable
    "T15": (_lws(4, mu=2), [
        (all_rdy, (0, 4, 4, 4, AV)),
        (unready(1), (None, None, 3, 4, None)),
        (update, (3, 4, 3, 0, UP)),
        (all_rdy, (0, 4, 4, 4, AV)),
    ]),
    # :2312-2406 rolling update with lws partition
    "T17": (_lws(6, partition=4), [
        (all_rdy, (4, 6, 6, 6, AV)),
        (update, (5, 6, 6, 0, UP)),
        (rdy(5), (4, 6, 6, 1, UP)),
        (rdy(4), (4, 6, 6, 2, None)),
        (partition(2), (3, 6, 6, 2, UP)),
        (rdy(3), (2, 6, 6, 3, UP)),
        (rdy(2), (2, 6, 6, 4, None)),
    ]),
}


def make_lws(cfg, name="test-sample"):
    return api.LeaderWorkerSet(
        name=name,
        replicas=cfg["replicas"],
        size=cfg["size"],
        rollingUpdate=api.RollingUpdateConfiguration(
            partition=cfg["partition"], maxUnavailable=cfg["maxUnavailable"], maxSurge=cfg["maxSurge"]
        ),
    )
