"""Patch emission against the reference's own table tests:
pkg/controllers/leaderworkerset_controller_test.go:50-758 (TestLeaderStatefulSetApplyConfig, 9 entries)
and pkg/controllers/pod_controller_test.go:42-425 (TestConstructWorkerStatefulSetApplyConfiguration,
4 entries), extracted into tests/golden/apply_configs.json by tests/golden/extract_apply_config_vectors.py.

Each entry's LeaderWorkerSet is a wrappers.* builder chain; it is replayed on the host object model,
the numeric fields come from the oracle's sweep of the encoded object (the same numbers the GPU
sweep emits), and the emitted patch must equal the expected apply configuration field for field
(ownerReferences aside: the table tests call the construct* function directly, the reconciler adds
the controller reference afterwards, leaderworkerset_controller.go:384)."""
import json
import os

import pytest

from lws_b200 import api, encoder, patches
from lws_b200 import records as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "apply_configs.json")))


def replay(chain):
    """wrappers.BuildBasicLeaderWorkerSet(name, ns).<steps>.Obj() → (api.LeaderWorkerSet, patches.Templates).
    BuildBasicLeaderWorkerSet (test/wrappers/wrappers.go:212-224): 1 replica, size 1, rolling update
    maxUnavailable 1 / maxSurge 0, restart policy RecreateGroupOnPodRestart, SubdomainShared."""
    steps = chain["$chain"]
    assert steps[0][0] == "BuildBasicLeaderWorkerSet"
    lws = api.LeaderWorkerSet(name=steps[0][1], namespace=steps[0][2], replicas=1, size=1,
                              rollingUpdate=api.RollingUpdateConfiguration(partition=0, maxUnavailable=1, maxSurge=0))
    tm = patches.Templates(workerTemplate={"spec": {}}, subdomainPolicy="Shared")

    def pod_spec(v):
        return json.loads(json.dumps(GOLD["pod_specs"][v["$chain"][0][0]]))

    for step in steps[1:]:
        name, args = step[0], step[1:]
        if name == "Replica":
            lws.replicas = args[0]
        elif name == "Size":
            lws.size = args[0]
        elif name == "RolloutStrategy":
            cfg = args[0].get("rollingUpdateConfiguration", {})
            lws.rollingUpdate = api.RollingUpdateConfiguration(partition=cfg.get("partition", 0) or 0,
                                                               maxUnavailable=cfg.get("maxUnavailable", 0),
                                                               maxSurge=cfg.get("maxSurge", 0))
            tm.rolloutStrategyType = args[0]["type"]
        elif name == "WorkerTemplateSpec":
            tm.workerTemplate = {"spec": pod_spec(args[0])}
        elif name == "LeaderTemplateSpec":
            tm.leaderTemplate = {"spec": pod_spec(args[0])}
        elif name == "Annotation":
            lws.annotations.update(args[0])
        elif name == "SubGroupSize":
            lws.subGroupSize = args[0]
            lws.subGroupPolicyType = lws.subGroupPolicyType or api.SubGroupPolicyTypeLeaderWorker
        elif name == "SubGroupType":
            lws.subGroupPolicyType = args[0]
        elif name == "RestartPolicy":
            lws.restartPolicy = args[0]
        elif name == "VolumeClaimTemplates":
            tm.volumeClaimTemplates = args[0]
        elif name == "PersistentVolumeClaimRetentionPolicy":
            tm.pvcRetentionPolicy = args[0]
        elif name == "Obj":
            pass
        else:
            raise AssertionError(f"unhandled builder step {name}")
    return lws, tm


def strip_owner(obj):
    obj = json.loads(json.dumps(obj))
    obj["metadata"].pop("ownerReferences", None)
    return obj


@pytest.mark.parametrize("case", GOLD["leader_statefulset"]["cases"], ids=lambda c: c["name"])
def test_leader_statefulset_patch(case, oracle_sweep):
    lws, tm = replay(case["lws"])
    rev = case["revisionKey"]
    sts = None
    if case.get("stsReplicas") is not None:  # the entry's existing leader StatefulSet (:672: 3 replicas)
        sts = api.StatefulSet(name=lws.name, namespace=lws.namespace, replicas=case["stsReplicas"], partition=0,
                              annotations={api.ReplicasAnnotationKey: str(lws.replicas)})
    t = encoder.encode_lws([encoder.LwsItem(lws=lws, revision_key=rev, leader_sts=sts)], encoder.Cluster())
    lws_out, _ = oracle_sweep(t)
    o = lws_out[0]
    assert not o["flags"] & R.LOUT_RUP_ERROR
    tmpl = patches.LeaderStatefulSetTemplate(lws, tm, rev)
    # the table test calls constructLeaderStatefulSetApplyConfiguration(lws, 0, replicas, key) with the
    # entry's replicas (the existing sts's when the entry has one)
    replicas = case["stsReplicas"] if case.get("stsReplicas") is not None else int(o["sts_replicas"])
    body = tmpl.emit(partition=0, replicas=replicas, max_unavailable=int(o["sts_max_unavailable"]))
    got = json.loads(body)
    assert strip_owner(got) == case["wantApplyConfig"]
    assert got["metadata"]["ownerReferences"][0]["kind"] == "LeaderWorkerSet"
    # the compiled byte template and a from-scratch serialisation agree byte for byte
    want_obj = json.loads(json.dumps(tmpl.build(0, replicas, int(o["sts_max_unavailable"]))))
    assert body == json.dumps(want_obj, sort_keys=True, separators=(",", ":")).encode()


@pytest.mark.parametrize("case", GOLD["worker_statefulset"]["cases"], ids=lambda c: c["name"])
def test_worker_statefulset_patch(case):
    lws, tm = replay(case["lws"])
    pod = case["pod"]["metadata"]
    labels = pod["labels"]
    tmpl = patches.WorkerStatefulSetTemplate(lws, tm)
    got = tmpl.build(leader_name=pod["name"], leader_uid="uid", group_index=labels[api.GroupIndexLabelKey],
                     group_key=labels[api.GroupUniqueHashLabelKey], revision_key=labels[api.RevisionKey])
    assert strip_owner(got) == case["wantStatefulSetConfig"]
    assert got["spec"]["replicas"] == lws.size - 1 and got["spec"]["ordinals"] == {"start": 1}


def test_worker_patch_node_selector_and_batch():
    """setNodeSelectorForWorkerPods (pod_controller.go:297-313) + batch emission from sweep outputs."""
    import oracle
    from lws_b200 import synth

    lws = api.LeaderWorkerSet(name="s", size=4, annotations={api.ExclusiveKeyAnnotationKey: "zone"})
    tm = patches.Templates(workerTemplate={"spec": {"containers": [{"name": "w", "image": "i", "resources": {}}]}})
    tmpl = patches.WorkerStatefulSetTemplate(lws, tm, exclusive_topology_key="zone")
    got = tmpl.build(leader_name="s-3", leader_uid="u", group_index=3, group_key="k", revision_key="r", topology_value="z1")
    assert got["spec"]["template"]["spec"]["nodeSelector"] == {"zone": "z1"}
    assert got["metadata"]["name"] == "s-3" and got["spec"]["selector"]["matchLabels"][api.GroupIndexLabelKey] == "3"
    # batch: one patch per group whose leader reconcile reaches Create(worker sts)
    t = synth.make("fuzz", 0.05, seed=3)
    lo, go, _ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)
    leaders = [(f"s-{g}", f"u{g}", g, f"k{g}", "r") for g in range(len(go))]
    out = patches.emit_worker_patches(tmpl, go, leaders, [f"dom-{d}" for d in range(t.n_domains)])
    n_create = int(((go["flags"] & R.GOUT_CREATE_WSTS) != 0).sum())
    assert sum(p is not None for p in out) == n_create > 0
    for p, o in zip(out, go):
        if p is not None and o["domain_id"] != R.NONE:
            assert json.loads(p)["spec"]["template"]["spec"]["nodeSelector"]["zone"] == f"dom-{int(o['domain_id'])}"
