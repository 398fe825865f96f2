"""Pins the DisaggregatedSet oracle (oracle/lwse_oracle_ds.c) to the reference's
own vectors:
  pkg/controllers/disaggregatedset/planner_test.go:106-505  22 exact sequences
      (tests/golden/planner_sequences.json, extracted by extract_planner_vectors.py)
  planner_test.go:507-672, :904-1055   completion + surge / unavailable invariants
  planner_test.go:722-798              nil-when-done / abnormal-state vectors
  executor_test.go:806-1014            scaleDownOld budget split (7 + 3 cases)
  executor_test.go:685-804             extractRollingUpdateConfig percent scaling
  executor_test.go:1151-1238           one ReconcileRollingUpdate call (3 rows)
  executor_test.go:334-481             single-reconcile behaviours
"""
import json
import os

import pytest

import oracle
from lws_b200 import api, encoder
from lws_b200 import records as R

HERE = os.path.dirname(os.path.abspath(__file__))
SEQ = json.load(open(os.path.join(HERE, "golden", "planner_sequences.json")))["cases"]


def all_steps(initial_old, target, surge, unavail):
    return oracle.ds_compute_all_steps(list(initial_old), list(target), list(surge), list(unavail))


def completes(steps, target):
    past, new = steps[-1]
    return all(p == 0 for p in past) and new == list(target)


@pytest.mark.parametrize("case", SEQ, ids=[c["name"] for c in SEQ])
def test_exact_sequences(case):
    surge = [c[0] for c in case["config"]]
    unavail = [c[1] for c in case["config"]]
    got = all_steps(case["source"], case["target"], surge, unavail)
    assert [[p, n] for p, n in got] == case["steps"]


N_ROLE = [  # planner_test.go:904-968
    ([3, 3, 3], [3, 3, 3], [1, 1, 1], [0, 0, 0]), ([6, 3, 2], [6, 3, 2], [2, 1, 1], [0, 0, 0]),
    ([4, 4, 4], [4, 4, 4], [2, 1, 3], [0, 0, 0]), ([2, 2, 2], [4, 4, 4], [1, 1, 1], [0, 0, 0]),
    ([4, 4, 4], [2, 2, 2], [1, 1, 1], [0, 0, 0]), ([0, 0, 0], [3, 3, 3], [1, 1, 1], [0, 0, 0]),
    ([4, 4, 4], [4, 4, 4], [0, 0, 0], [1, 1, 1]), ([4, 4, 4], [4, 4, 4], [1, 0, 2], [0, 1, 0]),
    ([4, 4, 4, 4], [4, 4, 4, 4], [1, 1, 1, 1], [0, 0, 0, 0]), ([8, 4, 2, 1], [8, 4, 2, 1], [2, 2, 1, 1], [0, 0, 0, 0]),
    ([1, 1, 1, 1], [3, 3, 3, 3], [1, 1, 1, 1], [0, 0, 0, 0]), ([5, 5, 5, 5], [2, 2, 2, 2], [1, 1, 1, 1], [0, 0, 0, 0]),
    ([0, 0, 0, 0], [4, 4, 4, 4], [1, 1, 1, 1], [0, 0, 0, 0]),
    ([5] * 5, [5] * 5, [1] * 5, [0] * 5), ([10, 5, 3, 2, 1], [10, 5, 3, 2, 1], [2, 2, 1, 1, 1], [0] * 5),
    ([1] * 5, [2] * 5, [1] * 5, [0] * 5), ([6] * 5, [3] * 5, [1] * 5, [0] * 5), ([0] * 5, [5] * 5, [1] * 5, [0] * 5),
    ([4, 4, 0], [4, 4, 4], [1, 1, 1], [0, 0, 0]), ([3, 3, 3, 0], [3, 3, 3, 3], [1] * 4, [0] * 4),
    ([2, 2, 2, 2, 2, 0], [2] * 6, [1] * 6, [0] * 6),
    ([4, 4, 4], [4, 4, 0], [1, 1, 1], [0, 0, 0]), ([3, 3, 3, 3], [3, 3, 3, 0], [1] * 4, [0] * 4),
    ([10, 10, 10, 10, 0, 10], [10, 10, 10, 10, 10, 0], [1] * 6, [0] * 6),
    ([10, 2], [6, 8], [2, 2], [0, 0]), ([10, 10], [4, 4], [2, 2], [0, 0]), ([8, 4], [3, 2], [1, 1], [0, 0]),
    ([5, 2], [5, 2], [0, 0], [1, 1]),
]


@pytest.mark.parametrize("io,tg,su,un", N_ROLE)
def test_n_role_rollout_completes(io, tg, su, un):
    assert completes(all_steps(io, tg, su, un), tg)


@pytest.mark.parametrize(
    "io,tg,su",
    [([3, 3, 3], [3, 3, 3], [1, 1, 1]), ([4] * 4, [4] * 4, [1] * 4), ([5] * 5, [5] * 5, [1] * 5),
     ([10, 2], [6, 8], [2, 2]), ([10, 10], [4, 4], [2, 2]), ([8, 4], [3, 2], [1, 1])],
)
def test_n_role_surge_constraint(io, tg, su):  # planner_test.go:970-1012
    for past, new in all_steps(io, tg, su, [0] * len(io)):
        for i in range(len(tg)):
            if new[i] == 0:
                continue
            assert past[i] + new[i] <= tg[i] + su[i]


@pytest.mark.parametrize(
    "io,tg,un", [([4, 4], [4, 4], [1, 1]), ([5, 2], [5, 2], [1, 1]), ([2, 5], [2, 5], [1, 1]), ([3, 3, 3], [3, 3, 3], [1, 1, 1])]
)
def test_n_role_unavailable_constraint(io, tg, un):  # planner_test.go:1014-1055
    steps = all_steps(io, tg, [0] * len(io), un)
    assert completes(steps, tg)
    for past, new in steps:
        for i in range(len(tg)):
            if io[i] < tg[i]:
                continue
            assert past[i] + new[i] >= tg[i] - un[i]


def test_unavailable_and_surge_priority():  # planner_test.go:507-672
    def totals(steps):
        return [sum(p) + sum(n) for p, n in steps]

    s = all_steps([4, 4], [4, 4], [0, 0], [1, 1])
    assert completes(s, [4, 4]) and min(totals(s)) < 8
    s = all_steps([4, 4], [4, 4], [1, 1], [1, 1])
    assert completes(s, [4, 4]) and min(totals(s)) >= 8
    s = all_steps([6, 6], [6, 6], [2, 2], [0, 0])
    assert completes(s, [6, 6]) and min(totals(s)) >= 12
    s = all_steps([6, 6], [6, 6], [0, 0], [2, 2])
    assert completes(s, [6, 6]) and min(totals(s)) < 12
    s = all_steps([6, 6], [6, 6], [2, 2], [2, 2])
    assert completes(s, [6, 6]) and min(totals(s)) >= 12
    for cfg in (([1, 0], [0, 1]), ([0, 1], [1, 0])):
        assert completes(all_steps([4, 4], [4, 4], *cfg), [4, 4])
    for sp, sd in ((6, 2), (2, 6), (8, 4)):
        assert completes(all_steps([sp, sd], [sp, sd], [1, 0], [0, 1]), [sp, sd])
    assert completes(all_steps([2, 2], [4, 4], [0, 0], [1, 1]), [4, 4])
    assert completes(all_steps([4, 4], [2, 2], [0, 0], [1, 1]), [2, 2])
    assert completes(all_steps([0, 0], [4, 4], [0, 0], [1, 1]), [4, 4])
    for size, un in ((4, 1), (6, 2), (10, 5)):
        s = all_steps([size, size], [size, size], [0, 0], [un, un])
        assert len(s) <= size * 4 and completes(s, [size, size])


def test_next_step_nil_when_done_and_abnormal_state():  # planner_test.go:722-798
    step = oracle.ds_compute_next_step
    assert step([3, 3], [0, 0], [3, 3], [3, 3], [1, 1], [0, 0]) is None  # complete
    assert step([0, 0], [0, 0], [0, 0], [0, 0], [1, 1], [0, 0]) is None  # totalSteps == 0
    # correctAbnormalState: currentOld above initialOld is clamped first, new untouched
    assert step([2, 2], [3, 2], [1, 1], [2, 2], [1, 1], [0, 0]) == ([2, 2], [1, 1])
    # new at target → drain everything
    assert step([3, 3], [2, 1], [3, 3], [3, 3], [1, 1], [0, 0]) == ([0, 0], [3, 3])


SCALE_DOWN = [  # executor_test.go:806-918: (revisions oldest→newest, budget, expected)
    ([(4, 4)], (2, 2), [(2, 2)]),
    ([(2, 2), (2, 2)], (2, 2), [(2, 2), (0, 0)]),
    ([(3, 2)], (1, 2), [(0, 0)]),
    ([(6, 6)], (2, 2), [(4, 4)]),
    ([(2, 2), (2, 2), (2, 2)], (4, 4), [(2, 2), (0, 0), (0, 0)]),
    ([(1, 2), (3, 3)], (1, 1), [(1, 2), (2, 2)]),
    ([(1, 1), (3, 3)], (2, 2), [(1, 1), (1, 1)]),
]


@pytest.mark.parametrize("revs,budget,want", SCALE_DOWN)
def test_scale_down_old(revs, budget, want):
    n = 2
    current = [sum(r[i] for r in revs) for i in range(n)]
    target = [current[i] - budget[i] for i in range(n)]
    order = list(range(len(revs)))[::-1]  # newest first
    got = oracle.ds_scale_down_old([list(r) for r in revs], order, current, target)
    assert [tuple(r) for r in got] == want


@pytest.mark.parametrize("budget,want", [((0, 1, 0), (4, 3)), ((1, 1, 0), (3, 3)), ((4, 0, 0), (0, 0))])
def test_scale_down_old_with_missing_role(budget, want):  # executor_test.go:920-1014
    current = [4, 4, 0]
    target = [current[i] - budget[i] for i in range(3)]
    got = oracle.ds_scale_down_old([[4, 4, -1]], [0], current, target)
    assert tuple(got[0][:2]) == want


def _ds(target=(4, 4), cfg=None):
    roles = [api.DisaggregatedRoleSpec("prefill", target[0], cfg), api.DisaggregatedRoleSpec("decode", target[1], cfg)]
    return api.DisaggregatedSet("test", roles=roles)


def _child(role, rev, replicas, ts, initial=None, ready=None):
    ann = {} if initial is None else {api.DSInitialReplicasAnnotationKey: str(initial)}
    return api.ChildLWS(role, rev, replicas, replicas if ready is None else ready, ts, ann)


@pytest.mark.parametrize(
    "a,b,c,want",  # executor_test.go:1151-1238: target (4,4), default config, initial-replicas=2
    [((2, 2), (2, 2), (0, 0), {"A": (2, 2), "B": (2, 2), "C": (1, 1)}),
     (None, (2, 2), (2, 2), {"B": (2, 2), "C": (3, 3)}),
     ((2, 2), (2, 2), (2, 2), {"A": (2, 2), "B": (1, 1), "C": (2, 2)})],
)
def test_reconcile_rolling_update_one_call(a, b, c, want):
    children = []
    for name, reps, ts in (("A", a, 1.0), ("B", b, 2.0), ("C", c, 3.0)):
        if reps is None:
            continue
        initial = 2 if name != "C" else None
        children += [_child("prefill", name, reps[0], ts, initial), _child("decode", name, reps[1], ts, initial)]
    t = encoder.encode_ds([encoder.DsItem(_ds(), "C", children)])
    ds_out, role_out, rr_out = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert ds_out[0]["flags"] & R.DOUT_ROLLING and ds_out[0]["flags"] & R.DOUT_STABLE
    revs = t.old_revisions[0] + ["C"]
    got = {rev: (int(rr_out[k * 2]), int(rr_out[k * 2 + 1])) for k, rev in enumerate(revs)}
    assert got == want


@pytest.mark.parametrize(
    "surge,unavail,replicas,want",  # executor_test.go:685-804 (percent scaling, priority rule)
    [("50%", 0, 4, (2, 0)), ("25%", 0, 4, (1, 0)), ("25%", "25%", 10, (3, 2)), ("100%", 0, 5, (5, 0)),
     (0, 0, 4, (1, 0)), (0, 2, 4, (0, 2)), (3, 0, 4, (3, 0))],
)
def test_extract_rolling_update_config(surge, unavail, replicas, want):
    """The config is only observable through the planner: compare with a run that is
    handed the expected (surge, unavailable) directly."""
    cfg = api.RollingUpdateConfiguration(maxSurge=surge, maxUnavailable=unavail)
    ds = _ds((replicas, replicas), cfg)
    children = [_child("prefill", "old", replicas, 1.0), _child("decode", "old", replicas, 1.0),
                _child("prefill", "new", 0, 2.0), _child("decode", "new", 0, 2.0)]
    t = encoder.encode_ds([encoder.DsItem(ds, "new", children)])
    ds_out, role_out, _ = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    direct = oracle.ds_compute_next_step([replicas] * 2, [replicas] * 2, [0, 0], [replicas] * 2,
                                         [want[0]] * 2, [want[1]] * 2)
    assert ds_out[0]["flags"] & R.DOUT_STEP
    assert [int(role_out[i]["next_old"]) for i in range(2)] == direct[0]
    assert [int(role_out[i]["next_new"]) for i in range(2)] == direct[1]


def test_reconciler_single_reconcile_behaviours():  # executor_test.go:334-481
    # fully drained old revision is deleted
    t = encoder.encode_ds([encoder.DsItem(_ds((2, 2)), "new", [
        _child("prefill", "old", 0, 1.0), _child("decode", "old", 0, 1.0),
        _child("prefill", "new", 2, 2.0), _child("decode", "new", 2, 2.0)])])
    ds_out, _, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert ds_out[0]["drained_revs"] == 1 and not ds_out[0]["flags"] & R.DOUT_ROLLING
    # first reconcile of an update: no new revision yet → init, new LWS at 0
    t = encoder.encode_ds([encoder.DsItem(_ds((2, 2)), "new", [
        _child("prefill", "old", 2, 1.0), _child("decode", "old", 2, 1.0)])])
    ds_out, _, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert ds_out[0]["flags"] & R.DOUT_INIT and list(rr[2:4]) == [0, 0]
    # no old scale-down while the new revision is not ready
    t = encoder.encode_ds([encoder.DsItem(_ds((2, 2)), "new", [
        _child("prefill", "old", 2, 1.0), _child("decode", "old", 2, 1.0),
        _child("prefill", "new", 1, 2.0, ready=0), _child("decode", "new", 1, 2.0, ready=0)])])
    ds_out, _, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert not ds_out[0]["flags"] & R.DOUT_STABLE and list(rr) == [2, 2, 1, 1]
    # old scales down once new is ready
    t = encoder.encode_ds([encoder.DsItem(_ds((2, 2)), "new", [
        _child("prefill", "old", 2, 1.0), _child("decode", "old", 2, 1.0),
        _child("prefill", "new", 1, 2.0), _child("decode", "new", 1, 2.0)])])
    ds_out, _, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert ds_out[0]["flags"] & R.DOUT_STEP and list(rr) == [1, 1, 1, 1]


def test_initial_replicas_annotation_rules():  # pkg/utils/disaggregatedset/utils_test.go:20-120
    mk = lambda v: api.ChildLWS("r", "x", 3, annotations={} if v is None else {api.DSInitialReplicasAnnotationKey: v})
    assert encoder.get_initial_replicas(mk(None)) == -1
    assert encoder.get_initial_replicas(mk("")) == -1
    assert encoder.get_initial_replicas(mk("abc")) == -1
    assert encoder.get_initial_replicas(mk("5")) == 5
    assert encoder.get_initial_replicas(mk("99999999999")) == -1


# ---- helper-level vectors of the reference (the oracle exports its helpers for exactly this) ----
@pytest.mark.parametrize("surge,unavail,want", [(1, 0, 1), (2, 0, 2), (0, 1, 1), (0, 2, 2), (0, 0, 1), (3, 2, 3)])
def test_batch_size(surge, unavail, want):  # planner_test.go:674-690
    assert oracle.ds_batch_size(surge, unavail) == want


@pytest.mark.parametrize(
    "initial_old,target,surge,want",  # planner_test.go:692-720 (DefaultRollingUpdateConfig = surge 1, unavailable 0)
    [((4, 4), (4, 4), (1, 1), 4), ((6, 2), (6, 2), (1, 1), 6), ((4, 4), (4, 4), (2, 2), 2), ((0, 0), (3, 3), (1, 1), 3)],
)
def test_compute_total_steps(initial_old, target, surge, want):
    assert oracle.ds_total_steps(list(initial_old), list(target), list(surge), [0, 0]) == want


def test_compute_next_new_replicas_edge_cases():  # planner_test.go:800-846
    r = oracle.ds_next_new([0, 4], [0, 2], 4)
    assert r[0] == 0 and r[1] > 2
    r = oracle.ds_next_new([4, 0], [2, 0], 4)
    assert r[0] > 2 and r[1] == 0
    assert oracle.ds_next_new([4, 4], [2, 2], 0) == [4, 4]


def test_compute_next_old_replicas_edge_cases():  # planner_test.go:852-898
    r = oracle.ds_next_old([0, 4], [0, 3], 4)
    assert r[0] == 0 and r[1] <= 3
    r = oracle.ds_next_old([4, 0], [3, 0], 4)
    assert r[0] <= 3 and r[1] == 0
    assert oracle.ds_next_old([4, 4], [2, 2], 0) == [0, 0]


def test_n_role_default_config():  # planner_test.go:1057-1069: surge 1, unavailable 0 when no config is given
    for n in (3, 4, 5):
        roles = [api.DisaggregatedRoleSpec(f"r{i}", 2) for i in range(n)]
        kids = [api.ChildLWS(f"r{i}", "old", 2, 2, 1.0) for i in range(n)] + \
               [api.ChildLWS(f"r{i}", "new", 0, 0, 2.0) for i in range(n)]
        t = encoder.encode_ds([encoder.DsItem(api.DisaggregatedSet("t", roles=roles), "new", kids)])
        _, role_out, _ = oracle.sweep_ds(t.ds, t.roles, t.revroles)
        direct = oracle.ds_compute_next_step([2] * n, [2] * n, [0] * n, [2] * n, [1] * n, [0] * n)
        assert [int(role_out[i]["next_new"]) for i in range(n)] == direct[1] == [1] * n


def test_fresh_deployment_and_scaling_without_rolling_update():
    """pkg/controllers/disaggregatedset/disaggregatedset_controller_test.go:70-102 (TestFreshDeploymentNoRollingUpdate)
    and :104-140 (TestScalingWithoutRollingUpdate): no old revision → the simple path of Reconcile
    (disaggregatedset_controller.go:95-112): the target revision's LWS go straight to spec.replicas."""
    roles = [api.DisaggregatedRoleSpec("prefill", 3, None), api.DisaggregatedRoleSpec("decode", 2, None)]
    t = encoder.encode_ds([encoder.DsItem(api.DisaggregatedSet("fresh-deploy", roles=roles), "rev", [])])
    ds_out, role_out, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert not ds_out[0]["flags"] & R.DOUT_ROLLING and not ds_out[0]["flags"] & R.DOUT_BAD_TABLE
    assert [int(x) for x in rr[-2:]] == [3, 2]  # prefill LWS created with 3 replicas, decode with 2
    roles = [api.DisaggregatedRoleSpec("prefill", 5, None), api.DisaggregatedRoleSpec("decode", 4, None)]
    t = encoder.encode_ds([encoder.DsItem(api.DisaggregatedSet("scale-test", roles=roles), "rev",
                                          [_child("prefill", "rev", 3, 1.0), _child("decode", "rev", 2, 1.0)])])
    ds_out, role_out, rr = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    assert not ds_out[0]["flags"] & R.DOUT_ROLLING
    assert [int(x) for x in rr[-2:]] == [5, 4]  # "prefill replicas should be scaled to 5", "decode … to 4"


@pytest.mark.parametrize("annotations,want", [  # lws_manager_test.go:51-101 TestParseInitialReplicasAnnotation
    ({}, -1), ({"other-key": "value"}, -1), ({api.DSInitialReplicasAnnotationKey: "not-a-number"}, -1),
    ({api.DSInitialReplicasAnnotationKey: "5"}, 5), ({api.DSInitialReplicasAnnotationKey: "0"}, 0)])
def test_parse_initial_replicas_annotation(annotations, want):
    assert encoder.get_initial_replicas(api.ChildLWS("r", "x", 3, annotations=dict(annotations))) == want


@pytest.mark.parametrize("replicas,want", [(None, 1), (5, 5), (0, 0)])  # lws_manager_test.go:103-138 TestGetLWSReplicas
def test_lws_replicas_nil_is_one(replicas, want):
    """getLWSReplicas (utils.go:160-165): nil spec.replicas counts as 1 — in the record row and in the sums."""
    t = encoder.encode_ds([encoder.DsItem(_ds((4, 4)), "new", [
        api.ChildLWS("prefill", "old", replicas, 0, 1.0, {}), _child("decode", "old", 2, 1.0),
        _child("prefill", "new", 0, 2.0), _child("decode", "new", 0, 2.0)])])
    assert int(t.revroles[0]["replicas"]) == want
    assert bool(int(t.revroles[0]["flags"]) & R.RR_REPLICAS_NIL) == (replicas is None)


INITIAL_STATE = json.load(open(os.path.join(HERE, "golden", "ds_initial_state_vectors.json")))["cases"]


@pytest.mark.parametrize("case", INITIAL_STATE, ids=[c["name"] for c in INITIAL_STATE])
def test_compute_initial_replica_state(case):
    """pkg/utils/disaggregatedset/utils_test.go:177-412 TestComputeInitialReplicaState (extracted by
    tests/golden/extract_ds_initial_state_vectors.py): ComputeInitialReplicaState (utils.go:55-81) sums, per
    role, the initial-replicas annotation of the old LeaderWorkerSets — spec.replicas when the annotation is
    absent or unparsable, 1 when that is nil too.  Here the rule is split between the encoder (the record row:
    replicas with nil → 1, initial_replicas or -1) and the planner-state sum every implementation applies to the
    rows (oracle/lwse_oracle_ds.c:394, lwse_ds_kernels.cu): each list element becomes the child of its own old
    revision, and the per-role sum over the encoded rows must be the reference's total."""
    assert api.DSRoleLabelKey == "disaggregatedset.x-k8s.io/role"
    children = []
    for i, lws in enumerate(case["lwsList"]):
        md = lws.get("metadata") or {}
        role = (md.get("labels") or {}).get(api.DSRoleLabelKey, "")
        if role == "":
            continue  # utils.go:61-63
        children.append(api.ChildLWS(role, f"old-{i}", (lws.get("spec") or {}).get("replicas"), 0, float(i + 1),
                                     dict(md.get("annotations") or {})))
    t = encoder.encode_ds([encoder.DsItem(_ds((4, 4)), "new", children)])
    names = t.role_names[0]
    n = len(names)
    got = {name: 0 for name in names}
    rows = t.revroles
    assert len(rows) % n == 0
    for k, row in enumerate(rows):
        if not int(row["flags"]) & R.RR_EXISTS:
            continue
        init = int(row["initial_replicas"])
        got[names[k % n]] += init if init >= 0 else int(row["replicas"])
    for role, want in case["want"].items():
        assert got.get(role, 0) == want
