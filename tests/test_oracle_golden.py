"""Pins the CPU oracle (oracle/lwse_oracle.c) to the reference's own test vectors.

Sources (paths relative to the reference root):
  pkg/controllers/leaderworkerset_controller_test.go:818-885, :887-1011
  test/integration/controllers/leaderworkerset_test.go (traces, see traces.py)
"""
import pytest

import oracle
from lws_b200 import api, encoder
from lws_b200 import records as R
from sim import LwsSim
from traces import TRACES, make_lws


# pkg/controllers/leaderworkerset_controller_test.go:818-885
@pytest.mark.parametrize(
    "lws_replicas,max_surge,max_unavailable,unready,want",
    [(1, 1, 0, 1, 2), (4, 2, 1, 2, 5), (2, 2, 1, 2, 3), (1, 1, 0, 0, 1), (1, 1, 1, 1, 1), (3, 0, 0, 1, 3)],
)
def test_calculate_rolling_update_replicas(lws_replicas, max_surge, max_unavailable, unready, want):
    got = oracle.lib().lwso_calculate_rolling_update_replicas(lws_replicas, max_surge, max_unavailable, unready)
    assert got == want


# pkg/controllers/leaderworkerset_controller_test.go:887-1011
@pytest.mark.parametrize(
    "replicas,mu,ms,updated,want",
    [(3, 0, 1, False, (0, 3)), (3, 0, 1, True, (2, 3)), (2, 1, 2, True, (2, 3))],
)
def test_rolling_update_parameters_kats(oracle_sweep, replicas, mu, ms, updated, want):
    lws = api.LeaderWorkerSet(
        name="test-sample", replicas=replicas, size=1,
        rollingUpdate=api.RollingUpdateConfiguration(partition=0, maxUnavailable=mu, maxSurge=ms),
    )
    sts = api.StatefulSet(name=lws.name, replicas=2, partition=0, annotations={api.ReplicasAnnotationKey: "2"})
    item = encoder.LwsItem(lws=lws, revision_key="rev-new", lws_updated=updated, leader_sts=sts)
    t = encoder.encode_lws([item], encoder.Cluster())
    lws_out, _ = oracle_sweep(t)
    assert not lws_out[0]["flags"] & R.LOUT_RUP_ERROR
    assert (lws_out[0]["sts_partition"], lws_out[0]["sts_replicas"]) == want


# k8s.io/apimachinery intstr.GetScaledValueFromIntOrPercent, pinned by
# pkg/controllers/disaggregatedset/executor_test.go:685-760 and
# pkg/controllers/leaderworkerset_controller_test.go:672-740
@pytest.mark.parametrize(
    "val,pct,total,up,want",
    [(50, 1, 4, 1, 2), (25, 1, 4, 1, 1), (25, 1, 10, 1, 3), (25, 1, 10, 0, 2), (100, 1, 5, 1, 5),
     (50, 1, 2, 0, 1), (3, 0, 10, 1, 3), (10, 1, 16, 1, 2), (10, 1, 16, 0, 1)],
)
def test_scaled_value(val, pct, total, up, want):
    assert oracle.lib().lwso_scaled_value(val, pct, total, up) == want


@pytest.mark.parametrize("name", sorted(TRACES))
def test_integration_trace(oracle_sweep, name):
    cfg, steps = TRACES[name]
    sim = LwsSim(make_lws(cfg), oracle_sweep)
    sim.settle()
    sim.create_leader_pods(0, cfg["replicas"])
    for i, (action, want) in enumerate(steps):
        action(sim)
        got = sim.state() + (sim.status["condition"],)
        for k, (g, w) in enumerate(zip(got, want)):
            if w is not None:
                assert g == w, f"{name} step {i}: field {k} got {got} want {want}"


# pkg/webhooks/pod_webhook_test.go:29-53 (genGroupUniqueKey) + hashlib cross-check
@pytest.mark.parametrize(
    "ns,pod,want",
    [("default", "test-sample", "95e88034e460983f51a9952fe128729fbc0663b5"),
     ("default", "podName", "390b34ab671d29e9997d7d4252b8bbf8da02f5b7"),
     ("leaderworkerset", "test-sample", "39f5d7e9122b9d94d3932e3720b43fd3b56347e8")],
)
def test_group_unique_key_kats(ns, pod, want):
    assert oracle.sha1([f"{ns}/{pod}"])[0].tobytes().hex() == want


def test_sha1_against_hashlib_all_padding_lengths():
    import hashlib

    msgs = [("x" * n).encode() + bytes([n % 251]) for n in range(0, 200)] + [b"", b"a" * 1000]
    got = oracle.sha1(msgs)
    for m, d in zip(msgs, got):
        assert d.tobytes() == hashlib.sha1(m).digest()


# pkg/webhooks/pod_webhook_test.go:272-305 (getSubGroupIndex)
@pytest.mark.parametrize("pod_count,sg,widx,want", [(4, 2, 2, 1), (5, 2, 2, 0), (9, 4, 8, 1), (8, 4, 7, 1), (8, 4, 3, 0)])
def test_sub_group_index(pod_count, sg, widx, want):
    assert oracle.lib().lwso_sub_group_index(pod_count, sg, widx) == want


# pkg/controllers/leaderworkerset_controller_test.go:50-758 TestLeaderStatefulSetApplyConfig: the numeric fields
# of the leader StatefulSet apply configuration — spec.replicas, the replicas annotation and
# rollingUpdate.maxUnavailable = max(1, maxUnavailable + min(maxSurge, lws replicas)) (:811-830).  No
# StatefulSet exists in these cases except the last one (3 replicas), so replicas = lws replicas (:290-298).
@pytest.mark.parametrize(
    "replicas,size,mu,ms,sts_replicas,want_mu,want_replicas",
    [(1, 1, 1, 0, None, 1, 1),       # :63   defaults
     (1, 2, 1, 0, None, 1, 1),       # :140  exclusive placement
     (2, 2, 1, 0, None, 1, 2),       # :209  leader template
     (1, 1, 2, 1, None, 3, 1),       # :279  2 maxUnavailable + 1 maxSurge
     (1, 1, 0, 2, None, 1, 1),       # :346  maxSurge capped at the replica count: 0 + min(2, 1)
     (1, 2, 1, 0, None, 1, 1),       # :413  subgroup size
     (1, 1, 1, 0, None, 1, 1),       # :483  volumeClaimTemplates
     (0, 1, 0, 0, None, 1, 0),       # :601  0 replicas, 0 / 0: at least 1
     (2, 1, 1, "50%", 3, 2, None)],  # :670  50 % of the LWS replicas (2), not of the sts replicas (3): 1 + 1
)
def test_leader_statefulset_apply_config_numbers(oracle_sweep, replicas, size, mu, ms, sts_replicas, want_mu, want_replicas):
    lws = api.LeaderWorkerSet(name="test-sample", replicas=replicas, size=size,
                              rollingUpdate=api.RollingUpdateConfiguration(partition=0, maxUnavailable=mu, maxSurge=ms))
    sts = None
    if sts_replicas is not None:
        sts = api.StatefulSet(name=lws.name, replicas=sts_replicas, partition=0,
                              annotations={api.ReplicasAnnotationKey: str(replicas)})
    item = encoder.LwsItem(lws=lws, revision_key="rev", lws_updated=False, leader_sts=sts)
    t = encoder.encode_lws([item], encoder.Cluster())
    lws_out, _ = oracle_sweep(t)
    assert not lws_out[0]["flags"] & R.LOUT_RUP_ERROR
    assert int(lws_out[0]["sts_max_unavailable"]) == want_mu
    if want_replicas is not None:
        assert int(lws_out[0]["sts_replicas"]) == want_replicas and int(lws_out[0]["sts_partition"]) == 0


# pkg/webhooks/leaderworkerset_webhook_test.go:27-83 getPercentValue: the input domain of the percent fields —
# an int is not a percentage, a string without '%' is invalid, "1%" and "101%" parse (the > 100 check is a
# separate validation, :126-170).  The encoder's parser is what feeds the kernels' percent flags.
@pytest.mark.parametrize("value,want_val,want_pct,want_valid",
                         [(1, 1, False, True), ("1", 0, False, False), ("1%", 1, True, True), ("101%", 101, True, True)])
def test_percent_value_parsing(value, want_val, want_pct, want_valid):
    assert encoder.parse_int_or_percent(value) == (want_val, want_pct, want_valid)
