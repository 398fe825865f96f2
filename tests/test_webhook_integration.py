"""PodWebhook.Default on the reference's integration table: the 20 entries of
test/integration/webhooks/pod_test.go:265-868 — TPU / LWS env vars, the injected subdomain, exclusive and
subgroup-exclusive affinity terms, idempotence, "does not override other terms" — extracted by
tests/golden/extract_webhook_integration_vectors.py, run through ``lws_b200.webhook.default_batch`` (the pod
as the admission request delivers it: JSON), and judged by restatements of the validators the entries call
(test/testutils/util.go:440-600).  The label entries of the same table (:68-263) are in test_webhook_batch.py."""
import copy
import hashlib
import json
import os

import numpy as np
import pytest

from lws_b200 import api
from lws_b200 import webhook as W
from test_webhook_default import resolve

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "webhook_integration_vectors.json")))
NAMESPACE = "ns-7w4kq"  # the integration suite generates one per entry; any name will do

TPU_WORKER_HOSTNAMES, TPU_WORKER_ID, TPU_NAME, TPU_PROCESS_PORT = "TPU_WORKER_HOSTNAMES", "TPU_WORKER_ID", "TPU_NAME", "TPU_PROCESS_PORT"
LWS_ENV = ["LWS_LEADER_ADDRESS", "LWS_GROUP_SIZE", "LWS_WORKER_INDEX"]
LEADER_REQUESTS_TPUS = "leaderworkerset.sigs.k8s.io/leader-requests-tpus"
REQ = "requiredDuringSchedulingIgnoredDuringExecution"


def sha1_batch(strings):
    return np.array([np.frombuffer(hashlib.sha1(s.encode()).digest(), np.uint8) for s in strings])


def meta(pod):
    return pod.get("metadata", {})


# ---- test/testutils/util.go ------------------------------------------------------------------------------
def has_all_env_vars_populated(pod, names):  # :449-473
    tpu_check = TPU_WORKER_ID in names or TPU_WORKER_HOSTNAMES in names
    containers = list(pod["spec"].get("containers") or []) + list(pod["spec"].get("initContainers") or [])
    checked = False
    for c in containers:
        if tpu_check and W.num_tpus_requested(c) == 0:
            continue
        checked = True
        have = {e["name"] for e in c.get("env") or []}
        if not all(n in have for n in names):
            return False
    return checked


def check_tpu_container_env(pod, hostnames):  # CheckTPUContainerHasCorrectEnvVars :515-562
    labels, ann = meta(pod).get("labels") or {}, meta(pod).get("annotations") or {}
    tpu = [c for c in pod["spec"].get("containers") or [] if W.num_tpus_requested(c) > 0]
    for i, c in enumerate(tpu):
        for e in c.get("env") or []:
            if e["name"] == TPU_WORKER_HOSTNAMES:
                assert e["value"] == hostnames
            if e["name"] == TPU_WORKER_ID:
                wi = int(labels.get(api.WorkerIndexLabelKey) or 0)
                if api.SubGroupSizeAnnotationKey in ann:
                    sg = int(ann[api.SubGroupSizeAnnotationKey])
                    pwi = wi % sg if ann.get(LEADER_REQUESTS_TPUS) == "true" else int((wi - 1) - sg * int((wi - 1) / sg))  # Go's %
                    want = pwi * len(tpu) + i
                elif labels.get(api.WorkerIndexLabelKey) == "0" or ann.get(LEADER_REQUESTS_TPUS) == "true":
                    want = wi * len(tpu) + i
                else:
                    want = (wi - 1) * len(tpu) + i
                assert e["value"] == str(want), (e, want)
            if e["name"] == TPU_PROCESS_PORT:
                assert e["value"] == str(8476 + i)


def validate_exclusive_terms(pod, exclusive_annotation, hash_label):  # ValidatePodExclusivePlacementTerms :564-597
    aff = pod["spec"].get("affinity") or {}
    if not aff.get("podAffinity") or not aff.get("podAntiAffinity"):
        return False
    topo = (meta(pod).get("annotations") or {}).get(exclusive_annotation, "")
    terms, valid_aff, valid_anti = 0, False, False
    for t in aff["podAffinity"].get(REQ) or []:
        if t.get("topologyKey", "") == topo:
            r = t["labelSelector"]["matchExpressions"][0]
            if r["key"] == hash_label and r["operator"] == "In" and r["values"][0] != "":
                valid_aff = True
                terms += 1
    for t in aff["podAntiAffinity"].get(REQ) or []:
        if t.get("topologyKey", "") == topo:
            rs = t["labelSelector"]["matchExpressions"]
            has_exist = any(r["key"] == hash_label and r["operator"] == "Exists" for r in rs)
            has_not_in = any(r["key"] == hash_label and r["operator"] == "NotIn" and r["values"][0] != "" for r in rs)
            valid_anti = has_exist and has_not_in
    return valid_aff and valid_anti and terms == 1


def build_pod(case):
    lit = json.loads(json.dumps(case["pod"]).replace("$NAMESPACE", NAMESPACE))
    pod = {"metadata": lit.get("ObjectMeta") or lit.get("metadata") or {}, "spec": resolve(lit.get("Spec") or lit.get("spec"))}
    for step in case["pre"]:
        if step["op"] == "SetExclusiveAffinities":  # the pod already went through the webhook once
            W.set_exclusive_affinities(pod, *step["args"])
        elif step["op"] == "AppendRequiredTerm":
            kind, term = step["args"]
            key = kind[0].lower() + kind[1:]
            pod["spec"].setdefault("affinity", {}).setdefault(key, {}).setdefault(REQ, []).append(copy.deepcopy(term))
    return pod


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"][:70] for c in GOLD["cases"]])
def test_defaulting_entry(case):
    pod = build_pod(case)
    expected = copy.deepcopy(pod)
    errs = W.default_batch([pod], sha1_batch)
    assert errs == [None]
    assert case["checks"], "an entry without checks would pin nothing"
    for chk in case["checks"]:
        fn, args, want = chk["fn"], chk["args"], chk["want"]
        if fn == "HasTPUEnvVarsPopulated":
            assert has_all_env_vars_populated(pod, [TPU_WORKER_HOSTNAMES, TPU_WORKER_ID, TPU_NAME]) is want
        elif fn == "HasLWSEnvVarsPopulated":
            assert has_all_env_vars_populated(pod, LWS_ENV) is want
        elif fn == "CheckTPUContainerHasCorrectEnvVars":
            check_tpu_container_env(pod, args[0])
        elif fn == "ValidatePodExclusivePlacementTerms":
            assert validate_exclusive_terms(pod, *args) is want
        elif fn == "IsContainerFirstEnvVarLWSLeaderAddress":  # :501-509
            for c in pod["spec"]["containers"]:
                assert c["env"][0]["name"] == "LWS_LEADER_ADDRESS"
        elif fn == "CheckContainerHasCorrectEnvVar":  # :490-499
            env = {k.lower(): v for k, v in args[0].items()}
            value = env["value"].replace("$NAMESPACE", meta(expected)["namespace"])
            for c in pod["spec"]["containers"]:
                for e in c.get("env") or []:
                    if e["name"] == env["name"]:
                        assert e["value"] == value
        elif fn == "FirstAffinityTermsKeepKey":  # pod_test.go:795-797
            aff = pod["spec"]["affinity"]
            assert (aff["podAffinity"][REQ][0]["labelSelector"]["matchExpressions"][0]["key"] == args[0]
                    or aff["podAntiAffinity"][REQ][0]["labelSelector"]["matchExpressions"][0]["key"] == args[0])
        else:
            raise AssertionError(f"unknown check {fn}")


def test_the_fixture_is_the_whole_range():
    assert GOLD["source"] == "test/integration/webhooks/pod_test.go:265-868" and len(GOLD["cases"]) == 20
    assert sum(len(c["checks"]) for c in GOLD["cases"]) == 36


def test_gang_scheduling_entry_adds_the_pod_group_annotation():
    """test/integration/webhooks/pod_test.go:913-935 ("should add pod group annotation when creating a lws pod"),
    with the Volcano provider configured (pod_webhook.go:159-164 → volcano_provider.go:103-109): the pod's
    annotation scheduling.k8s.io/group-name is "<lws>-<group index>-<revision>" = "test-0-1"."""
    pod = {"metadata": {"name": "test-pod-0", "namespace": NAMESPACE,
                        "labels": {api.SetNameLabelKey: "test", api.GroupIndexLabelKey: "0", api.RevisionKey: "1"},
                        "annotations": {api.SizeAnnotationKey: "2"}},
           "spec": resolve({"$chain": [["MakeLeaderPodSpec"]]})}
    assert W.default_batch([pod], sha1_batch, inject_pod_group_metadata=W.volcano_inject_pod_group_metadata) == [None]
    assert pod["metadata"]["annotations"][W.KubeGroupNameAnnotationKey] == "test-0-1"
    # no provider configured: no annotation (schedulerProvider == nil, pod_webhook.go:159)
    pod2 = {"metadata": {"name": "test-pod-0", "namespace": NAMESPACE,
                         "labels": {api.SetNameLabelKey: "test", api.GroupIndexLabelKey: "0", api.RevisionKey: "1"},
                         "annotations": {api.SizeAnnotationKey: "2"}},
            "spec": resolve({"$chain": [["MakeLeaderPodSpec"]]})}
    assert W.default_batch([pod2], sha1_batch) == [None]
    assert W.KubeGroupNameAnnotationKey not in pod2["metadata"]["annotations"]
