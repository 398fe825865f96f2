"""Multi-cycle DisaggregatedSet flows of the reference, replayed on the oracle:
  executor_test.go:1244-1257  TestMidRolloutABC             A(1,2)+B(1,2) → C(2,4)
  executor_test.go:1261-1298  TestAsymmetricSizesCoordinatedDrain  A(1,2), B(3,1) → C(4,3), surge (1,2):
                              no revision may ever have exactly one role at 0
"""
import pytest

import oracle
from ds_sim import DsSim
from lws_b200 import api


def oracle_sweep_ds(t):
    return oracle.sweep_ds(t.ds, t.roles, t.revroles)


def abc(target, a, b, surge, order):
    cfg = [api.RollingUpdateConfiguration(maxSurge=surge[i], maxUnavailable=0) for i in range(2)]
    ds = api.DisaggregatedSet("test", roles=[api.DisaggregatedRoleSpec("prefill", target[0], cfg[0]),
                                             api.DisaggregatedRoleSpec("decode", target[1], cfg[1])])
    kids = {"A": [api.ChildLWS("prefill", "A", a[0], a[0], 0.0), api.ChildLWS("decode", "A", a[1], a[1], 0.0)],
            "B": [api.ChildLWS("prefill", "B", b[0], b[0], 0.0), api.ChildLWS("decode", "B", b[1], b[1], 0.0)]}
    children = [c for rev in order for c in kids[rev]]  # GroupByRevision ranges a Go map: both orders are legal
    return ds, children


def run_flow(sweep_ds, target, a, b, surge, order, check_orphans=False):
    ds, children = abc(target, a, b, surge, order)
    sim = DsSim(ds, "C", children, sweep_ds)
    for i in range(20):
        sim.reconcile()
        sim.simulate_all_ready()
        if check_orphans:
            norm = lambda v: 0 if v == -1 else v
            for rev in ("A", "B"):
                p, d = norm(sim.replicas("prefill", rev)), norm(sim.replicas("decode", rev))
                assert (p == 0) == (d == 0), f"step {i}: {rev} orphaned prefill={p} decode={d}"
    for rev in ("A", "B"):
        for role in ("prefill", "decode"):
            assert sim.replicas(role, rev) in (0, -1), f"{rev}/{role} not drained"
    assert (sim.replicas("prefill", "C"), sim.replicas("decode", "C")) == tuple(target)
    return sim


@pytest.mark.parametrize("order", [("A", "B"), ("B", "A")])
def test_mid_rollout_abc(order):
    run_flow(oracle_sweep_ds, (2, 4), (1, 2), (1, 2), (1, 1), order)


@pytest.mark.parametrize("order", [("A", "B"), ("B", "A")])
def test_asymmetric_sizes_coordinated_drain(order):
    run_flow(oracle_sweep_ds, (4, 3), (1, 2), (3, 1), (1, 2), order, check_orphans=True)
