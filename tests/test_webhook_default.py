"""The rest of PodWebhook.Default against the reference's own tables, extracted by
tests/golden/extract_webhook_vectors.py into tests/golden/webhook_vectors.json:
  pkg/webhooks/pod_webhook_test.go:66-169 (SetExclusiveAffinities), :171-270 (exclusiveAffinityApplied),
  pkg/utils/pod/pod_utils_test.go:103-186 (AddLWSVariables),
  pkg/utils/accelerators/tpu_test.go:34-293, :295-346, :348-588 (AddTPUVariables, …Skip, …SubGroup),
  :590-656, :658-704, :706-756 (getContainersRequestingTPUs, getContainerRequestingTPUs, PodRequestsTPUs),
  pkg/utils/pod/pod_utils_test.go:188-256 (getEnvVarValueIfInContainer)."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from lws_b200 import api
from lws_b200 import webhook as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "webhook_vectors.json")))
FUNCS = GOLD["wrappers"]["functions"]


def make_pod_with_labels(set_name, group_index, worker_index, namespace, size):
    """test/wrappers/wrappers.go:257-277 MakePodWithLabels"""
    name = f"{set_name}-{group_index}" if worker_index == "0" else f"{set_name}-{group_index}-{worker_index}"
    return {"spec": resolve({"$chain": [["MakePodSpecWithInitContainer"]]}),
            "metadata": {"name": name, "namespace": namespace,
                         "labels": {api.GroupIndexLabelKey: group_index, api.SetNameLabelKey: set_name,
                                    api.WorkerIndexLabelKey: worker_index},
                         "annotations": {api.SizeAnnotationKey: str(size)}}}


def call(name, args):
    if name == "MakePodWithLabels":
        return make_pod_with_labels(*args)
    if name == "MakeContainerWithTPUAndEnvVars":  # wrappers.go:398-402
        c = call("MakeContainerWithTPU", args[:1])
        c["env"] = [resolve(e) for e in args[1:]]
        return c
    if name == "MakeLeaderPodSpecWithTPUAndEnvVars":  # :380-384
        s = call("MakeLeaderPodSpecWithTPUResource", [])
        s["containers"][0]["env"] = [resolve(e) for e in args]
        return s
    f = FUNCS[name]
    v = copy.deepcopy(f["value"])
    v = json.loads(json.dumps(v).replace('"$name"', json.dumps(args[0]))) if f["params"] == ["name"] else v
    return resolve(v)


def resolve(v):
    """Replace the symbolic builder calls of the extracted literals by what the builders return."""
    if isinstance(v, dict):
        if "$chain" in v:
            out = None
            for step in v["$chain"]:
                out = call(step[0], step[1:])
            return out
        if "$call" in v:
            return call(v["$call"], v["args"])
        return {k: resolve(x) for k, x in v.items()}
    if isinstance(v, list):
        return [resolve(x) for x in v]
    return v


def env_of(c, name):
    return W.get_env_var_value_if_in_container(c, name)


@pytest.mark.parametrize("case", GOLD["set_exclusive_affinities"]["cases"], ids=lambda c: c["name"])
def test_set_exclusive_affinities(case):
    pod = resolve(case["pod"])
    W.set_exclusive_affinities(pod, case["groupUniqueKey"], case["topologyKey"], case["podAffinityKey"])
    assert pod == resolve(case["expectedPod"])
    again = copy.deepcopy(pod)
    W.set_exclusive_affinities(again, "other-key", case["topologyKey"], case["podAffinityKey"])
    assert again == pod  # idempotent on the topology key (:186)


@pytest.mark.parametrize("case", GOLD["exclusive_affinity_applied"]["cases"], ids=lambda c: c["name"])
def test_exclusive_affinity_applied(case):
    assert W.exclusive_affinity_applied(resolve(case["pod"]), case["topologyKey"]) is bool(case.get("expectedAppliedExclusivePlacement", False))


@pytest.mark.parametrize("case", GOLD["add_lws_variables"]["cases"], ids=lambda c: c["name"])
def test_add_lws_variables(case):
    pod = resolve(case["pod"])
    assert W.add_lws_variables(pod) is None
    cs = pod["spec"]["containers"] + pod["spec"].get("initContainers", [])
    assert cs
    for c in cs:
        assert [e["value"] for e in c["env"][:3]] == [case["expectedLwsLeaderAddress"], str(case["expectedGroupSize"]),
                                                      case["expectedWorkerIndex"]]
        assert [e["name"] for e in c["env"]] == ["LWS_LEADER_ADDRESS", "LWS_GROUP_SIZE", "LWS_WORKER_INDEX", "key1", "key2"]


def check_tpu_env(pod, case, sub_group):
    cs = W.containers_requesting_tpus(pod["spec"])
    if not cs:
        for c in pod["spec"].get("containers", []) + pod["spec"].get("initContainers", []):
            assert not c.get("env")
        return
    labels, ann = pod["metadata"].get("labels", {}), pod["metadata"].get("annotations", {})
    for i, c in enumerate(cs if not sub_group else cs[:1]):
        assert env_of(c, "TPU_NAME") == (True, case["expectedTpuName"])
        assert env_of(c, "TPU_WORKER_HOSTNAMES") == (True, case["expectedTpuWorkerHostNames"])
        if "expectedTpuProcessAddresses" in case:
            assert env_of(c, "TPU_PROCESS_ADDRESSES") == (True, case["expectedTpuProcessAddresses"])
        if sub_group:
            assert env_of(c, "TPU_WORKER_ID") == (True, str(case["expectedTpuWorkerId"]))
            if "expectedTpuProcessPort" in case:
                assert env_of(c, "TPU_PROCESS_PORT") == (True, case["expectedTpuProcessPort"])
        else:  # the reference's own check: id = i + containers x pod worker index, port = expected + i (tpu_test.go:255-287)
            if labels.get(api.WorkerIndexLabelKey) == "0":
                pwi = 0
            else:
                pwi = W.get_parent_name_and_ordinal(pod["metadata"]["name"])[1]
                if ann.get(W.LeaderRequestsTPUsAnnotationKey) != "true":
                    pwi -= 1
            assert env_of(c, "TPU_WORKER_ID") == (True, str(i + len(cs) * pwi))
            assert env_of(c, "TPU_PROCESS_PORT") == (True, str(int(case["expectedTpuProcessPort"]) + i))


@pytest.mark.parametrize("case", GOLD["add_tpu_variables"]["cases"], ids=lambda c: c["name"])
def test_add_tpu_variables(case):
    pod = resolve(case["pod"])
    pod.setdefault("spec", {}).setdefault("subdomain", "default")  # tpu_test.go:229-231
    if not pod["spec"]["subdomain"]:
        pod["spec"]["subdomain"] = "default"
    err = W.add_tpu_variables(pod, case["size"])
    if case.get("expectedError"):
        assert err
        return
    assert err is None
    check_tpu_env(pod, case, sub_group=False)


@pytest.mark.parametrize("case", GOLD["add_tpu_variables_subgroup"]["cases"], ids=lambda c: c["name"])
def test_add_tpu_variables_subgroup(case):
    pod = resolve(case["pod"])
    pod.setdefault("spec", {})
    if not pod["spec"].get("subdomain"):
        pod["spec"]["subdomain"] = "default"
    err = W.add_tpu_variables_subgroup(pod)
    if case.get("expectedError"):
        assert err
        return
    assert err is None
    check_tpu_env(pod, case, sub_group=True)


@pytest.mark.parametrize("case", GOLD["add_tpu_variables_skip"]["cases"], ids=lambda c: c["name"])
def test_add_tpu_variables_skip(case):
    pod = resolve(case["pod"])
    before = copy.deepcopy(pod)
    assert W.add_tpu_variables(pod, 2) is None and pod["spec"] == before["spec"]  # nothing is injected twice


@pytest.mark.parametrize("case", GOLD["get_containers_requesting_tpus"]["cases"], ids=lambda c: c["name"])
def test_get_containers_requesting_tpus(case):
    """tpu_test.go:590-656: only the count is checked there; the order (containers, then init containers) by :658-704."""
    got = W.containers_requesting_tpus(resolve(case["podSpec"]) or {})
    assert len(got) == case["expectedNumContainer"]


def _comparable_container(c):
    """cmp.Diff over corev1.Container: quantities compare by value ("4" == 4), absent == empty."""
    if c is None:
        return None
    out = {k: v for k, v in c.items() if v not in (None, [], {})}
    res = {part: {k: str(q) for k, q in qs.items()} for part, qs in (out.get("resources") or {}).items() if qs}
    if res:
        out["resources"] = res
    return out


@pytest.mark.parametrize("case", GOLD["get_container_requesting_tpus"]["cases"], ids=lambda c: c["name"])
def test_get_container_requesting_tpus(case):
    """tpu_test.go:658-704; getContainerRequestingTPUs (tpu.go:91-97) is the first of containers_requesting_tpus."""
    got = W.containers_requesting_tpus(resolve(case["podSpec"]))
    first = got[0] if got else None
    assert _comparable_container(first) == _comparable_container(resolve(case["expectedContainer"]))


@pytest.mark.parametrize("case", GOLD["pod_requests_tpus"]["cases"], ids=lambda c: c["name"])
def test_pod_requests_tpus(case):
    """tpu_test.go:706-756"""
    assert W.pod_requests_tpus(resolve(case["podSpec"]) or {}) is case["expected"]


@pytest.mark.parametrize("case", GOLD["get_env_var_if_in_container"]["cases"], ids=lambda c: c["name"])
def test_get_env_var_value_if_in_container(case):
    """pod_utils_test.go:188-256"""
    found, value = W.get_env_var_value_if_in_container(case["container"], case["envVarName"])
    assert found is case["expectEnvVar"]
    assert value == case["expectedEnvValue"]


def test_default_batch_end_to_end():
    """Leader and worker of a subgroup-exclusive TPU group through the whole Default: labels, SHA-1
    keys (one batched hash call), affinity terms on both label keys, TPU and LWS variables."""
    def sha1_batch(strings):
        return np.array([list(hashlib.sha1(s.encode()).digest()) for s in strings], dtype=np.uint8)

    tpu = {"name": "c", "image": "i", "resources": {"limits": {"google.com/tpu": "4"}}}
    common_ann = {api.SizeAnnotationKey: "5", api.ExclusiveKeyAnnotationKey: "zone", api.SubGroupSizeAnnotationKey: "2",
                  api.SubGroupExclusiveKeyAnnotationKey: "rack", W.LeaderRequestsTPUsAnnotationKey: "true"}
    leader = {"metadata": {"name": "lws-3", "namespace": "ns", "labels": {api.SetNameLabelKey: "lws", api.WorkerIndexLabelKey: "0"},
                           "annotations": dict(common_ann)},
              "spec": {"containers": [copy.deepcopy(tpu)], "subdomain": "lws"}}
    worker = {"metadata": {"name": "lws-3-4", "namespace": "ns",
                           "labels": {api.SetNameLabelKey: "lws", api.GroupIndexLabelKey: "3"},
                           "annotations": {**common_ann, api.LeaderPodNameAnnotationKey: "lws-3"}},
              "spec": {"containers": [copy.deepcopy(tpu)], "subdomain": "lws"}}
    stranger = {"metadata": {"name": "x", "labels": {"app": "y"}}, "spec": {}}
    bad = {"metadata": {"name": "lws-zz", "labels": {api.SetNameLabelKey: "lws"}, "annotations": {api.SizeAnnotationKey: "2"}}, "spec": {}}
    errs = W.default_batch([leader, worker, stranger, bad], sha1_batch)
    assert errs[:3] == [None, None, None] and errs[3] == "parsing pod ordinal for pod lws-zz"
    gk = hashlib.sha1(b"ns/lws-3").hexdigest()
    sk0, sk1 = hashlib.sha1(b"lws-3/0").hexdigest(), hashlib.sha1(b"lws-3/1").hexdigest()
    ll, wl = leader["metadata"]["labels"], worker["metadata"]["labels"]
    assert ll[api.GroupIndexLabelKey] == "3" and ll[api.GroupUniqueHashLabelKey] == gk
    assert ll[api.SubGroupIndexLabelKey] == "0" and ll[api.SubGroupUniqueHashLabelKey] == sk0
    # worker 4 of 5 pods, subgroup size 2: (5-1) % 2 == 0 → (4-1)/2 = 1
    assert wl[api.WorkerIndexLabelKey] == "4" and wl[api.SubGroupIndexLabelKey] == "1" and wl[api.SubGroupUniqueHashLabelKey] == sk1
    req = "requiredDuringSchedulingIgnoredDuringExecution"
    la = leader["spec"]["affinity"]
    assert [(t["topologyKey"], t["labelSelector"]["matchExpressions"][0]["values"]) for t in la["podAffinity"][req]] == \
        [("zone", [gk]), ("rack", [sk0])]
    assert [t["labelSelector"]["matchExpressions"][1]["operator"] for t in la["podAntiAffinity"][req]] == ["NotIn", "NotIn"]
    wa = worker["spec"]["affinity"]  # workers get only the subgroup terms (:136-156)
    assert [(t["topologyKey"], t["labelSelector"]["matchExpressions"][0]["key"]) for t in wa["podAffinity"][req]] == \
        [("rack", api.SubGroupUniqueHashLabelKey)]
    assert env_of(worker["spec"]["containers"][0], "LWS_LEADER_ADDRESS") == (True, "lws-3.lws.ns")
    assert env_of(worker["spec"]["containers"][0], "TPU_WORKER_ID") == (True, "0")  # 4 % 2
    assert "affinity" not in stranger["spec"]
