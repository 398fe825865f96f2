"""More of the reference's DisaggregatedSet vectors, replayed on the oracle (and, through
tests/test_gpu_other_paths.py, on the GPU engine):

  executor_test.go:486-556   TestSortByNewestTimestamp — old revisions drain newest first; a revision's
                             age is the MAX creation timestamp over its roles; the caller's slice order
                             does not matter
  executor_test.go:558-588   TestIsRevisionStable — the new revision is stable iff every role has
                             replicas == readyReplicas (0/0 counts as stable)
  service_manager_test.go    revision readiness (service_manager.go:57-89,174-189): every spec role
                             has readyReplicas >= 1
"""
import pytest

import oracle
from ds_sim import DsSim
from lws_b200 import api, encoder
from lws_b200 import records as R


def oracle_sweep_ds(t):
    return oracle.sweep_ds(t.ds, t.roles, t.revroles)


def _ds(target):
    cfg = api.RollingUpdateConfiguration(maxSurge=1, maxUnavailable=0)
    return api.DisaggregatedSet("test", roles=[api.DisaggregatedRoleSpec("prefill", target, cfg),
                                               api.DisaggregatedRoleSpec("decode", target, cfg)])


def drain_order(sweep_ds, revisions):
    """revisions: [(hash, ts_prefill, ts_decode)] in the caller's (arbitrary) order, one replica per
    role each.  Roll out a new revision with maxSurge 1 and record the order in which the old
    revisions reach zero."""
    children = []
    for h, tp, td in revisions:
        children += [api.ChildLWS("prefill", h, 1, 1, float(tp)), api.ChildLWS("decode", h, 1, 1, float(td))]
    sim = DsSim(_ds(len(revisions)), "new", children, sweep_ds)
    order = []
    for _ in range(6 * len(revisions) + 6):
        sim.reconcile()
        sim.simulate_all_ready()
        for h, _, _ in revisions:
            gone = all(sim.replicas(role, h) in (0, -1) for role in ("prefill", "decode"))
            if gone and h not in order:
                order.append(h)
    assert sim.replicas("prefill", "new") == len(revisions) and sim.replicas("decode", "new") == len(revisions)
    return order


SORT_CASES = [  # executor_test.go:506-511 (offsets in minutes) and :536-555 (max over the roles)
    ([("hash1", 0, 0)], ["hash1"]),
    ([("newest", 120, 120), ("oldest", 0, 0), ("middle", 60, 60)], ["newest", "middle", "oldest"]),
    ([("d", 30, 30), ("a", 0, 0), ("c", 20, 20), ("b", 10, 10)], ["d", "c", "b", "a"]),
    ([("A", 0, 20), ("B", 10, 10)], ["A", "B"]),
    ([("B", 10, 10), ("A", 0, 20)], ["A", "B"]),  # "does not modify original slice": input order is irrelevant
]


@pytest.mark.parametrize("revisions,want", SORT_CASES)
def test_old_revisions_drain_newest_first(revisions, want):
    assert drain_order(oracle_sweep_ds, revisions) == want


STABLE_CASES = [  # executor_test.go:566-572: (prefill replicas, ready, decode replicas, ready) → stable
    ((3, 3, 2, 2), True), ((3, 2, 2, 2), False), ((3, 3, 2, 1), False), ((3, 1, 2, 0), False), ((0, 0, 0, 0), True),
]


def stability_tables(case):
    pr, prd, dr, drd = case
    ds = api.DisaggregatedSet("test", roles=[api.DisaggregatedRoleSpec("prefill", 4), api.DisaggregatedRoleSpec("decode", 4)])
    children = [api.ChildLWS("prefill", "old", 2, 2, 1.0), api.ChildLWS("decode", "old", 2, 2, 1.0),
                api.ChildLWS("prefill", "hash1", pr, prd, 2.0), api.ChildLWS("decode", "hash1", dr, drd, 2.0)]
    return encoder.encode_ds([encoder.DsItem(ds, "hash1", children)])


@pytest.mark.parametrize("case,want", STABLE_CASES)
def test_is_revision_stable(case, want):
    ds_out, _, rr = oracle_sweep_ds(stability_tables(case))
    assert bool(ds_out[0]["flags"] & R.DOUT_STABLE) == want
    if not want:  # an unstable new revision freezes the old one (executor.go:160-164)
        assert list(rr[:2]) == [2, 2]


READY_CASES = [  # service_manager.go:174-189: a revision is ready iff every spec role has readyReplicas >= 1
    ((1, 1), True), ((3, 1), True), ((0, 1), False), ((1, 0), False), ((0, 0), False),
]


def readiness_tables(ready):
    ds = api.DisaggregatedSet("test", roles=[api.DisaggregatedRoleSpec("prefill", 2), api.DisaggregatedRoleSpec("decode", 2)])
    children = [api.ChildLWS("prefill", "old", 2, 2, 1.0), api.ChildLWS("decode", "old", 2, 2, 1.0),
                api.ChildLWS("prefill", "new", 3, ready[0], 2.0), api.ChildLWS("decode", "new", 3, ready[1], 2.0)]
    return encoder.encode_ds([encoder.DsItem(ds, "new", children)])


@pytest.mark.parametrize("ready,want", READY_CASES)
def test_revision_readiness_for_services(ready, want):
    ds_out, _, _ = oracle_sweep_ds(readiness_tables(ready))
    assert bool(ds_out[0]["flags"] & R.DOUT_NEW_READY) == want
    assert int(ds_out[0]["ready_revs"]) & 1  # the old revision (bit 0) has both roles ready
