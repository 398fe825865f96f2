"""The label pass of PodWebhook.Default, batched (lws_b200/webhook.py), on the reference's vectors:

  test/integration/webhooks/pod_test.go:68-263   labels of non-LWS / worker / leader pods, with and
                                                 without subgroups (the tests mask the SHA-1 values;
                                                 here they are checked against hashlib as well)
  pkg/webhooks/pod_webhook_test.go:29-53         genGroupUniqueKey KATs
  pkg/webhooks/pod_webhook_test.go:272-305       getSubGroupIndex
  pod_webhook.go:92-99,104-107,139-142           the error paths
The CPU run hashes with the oracle's SHA-1; tests/test_gpu_other_paths.py runs the same cases over the
CUDA SHA-1 kernel."""
import hashlib

import pytest

import oracle
from lws_b200 import api, webhook

NS = "default"
L, A = api, api


def sha(s):
    return hashlib.sha1(s.encode()).hexdigest()


def cases():
    """(pod, expected labels, expected subdomain)"""
    out = []
    # :68-90 not a leaderworkerset pod: untouched
    out.append((api.Pod("randompod", NS, labels={"foo": "bar"}), {"foo": "bar"}, ""))
    # :91-118 worker pod: worker index from the name
    out.append((api.Pod("test-1-1", NS, labels={L.SetNameLabelKey: "test", L.GroupIndexLabelKey: "1"},
                        annotations={A.SizeAnnotationKey: "2"}),
                {L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "1", L.GroupIndexLabelKey: "1"}, ""))
    # :119-151 leader pod: group index from the name, group key
    out.append((api.Pod("test-1", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0"},
                        annotations={A.SizeAnnotationKey: "2"}),
                {L.GroupIndexLabelKey: "1", L.SetNameLabelKey: "test", L.GroupUniqueHashLabelKey: sha(f"{NS}/test-1"),
                 L.WorkerIndexLabelKey: "0"}, ""))
    # :152-191 leader with subgroups: subgroup 0 and its key
    out.append((api.Pod("test-1", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0"},
                        annotations={A.SizeAnnotationKey: "5", A.SubGroupSizeAnnotationKey: "4"}),
                {L.GroupIndexLabelKey: "1", L.SetNameLabelKey: "test", L.GroupUniqueHashLabelKey: sha(f"{NS}/test-1"),
                 L.SubGroupUniqueHashLabelKey: sha("test-1/0"), L.WorkerIndexLabelKey: "0", L.SubGroupIndexLabelKey: "0"}, ""))
    # :192-228 worker 3 of size 4, subgroup size 2 → subgroup 1 ((4-1) % 2 != 0 → 3 / 2)
    out.append((api.Pod("test-1-3", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "3", L.GroupIndexLabelKey: "1"},
                        annotations={A.SizeAnnotationKey: "4", A.SubGroupSizeAnnotationKey: "2", A.LeaderPodNameAnnotationKey: "test-1"}),
                {L.SetNameLabelKey: "test", L.SubGroupUniqueHashLabelKey: sha("test-1/1"), L.WorkerIndexLabelKey: "3",
                 L.SubGroupIndexLabelKey: "1", L.GroupIndexLabelKey: "1"}, ""))
    # :229-264 worker 4 of size 5, subgroup size 2 → (5-1) % 2 == 0 → (4-1) / 2 = 1
    out.append((api.Pod("test-1-4", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "4", L.GroupIndexLabelKey: "1"},
                        annotations={A.SizeAnnotationKey: "5", A.SubGroupSizeAnnotationKey: "2", A.LeaderPodNameAnnotationKey: "test-1"}),
                {L.SetNameLabelKey: "test", L.SubGroupUniqueHashLabelKey: sha("test-1/1"), L.WorkerIndexLabelKey: "4",
                 L.SubGroupIndexLabelKey: "1", L.GroupIndexLabelKey: "1"}, ""))
    # :357-371 subdomainPolicy UniquePerReplica: the leader's subdomain is its own name
    out.append((api.Pod("test-sample-1", NS, labels={L.SetNameLabelKey: "test-sample", L.WorkerIndexLabelKey: "0"},
                        annotations={A.SizeAnnotationKey: "5", webhook.SubdomainPolicyAnnotationKey: webhook.SubdomainUniquePerReplica}),
                {L.SetNameLabelKey: "test-sample", L.WorkerIndexLabelKey: "0", L.GroupIndexLabelKey: "1",
                 L.GroupUniqueHashLabelKey: sha(f"{NS}/test-sample-1")}, "test-sample-1"))
    # leader with LeaderExcluded subgroup policy gets no subgroup labels (:125-127)
    out.append((api.Pod("test-2", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0"},
                        annotations={A.SizeAnnotationKey: "5", A.SubGroupSizeAnnotationKey: "2",
                                     A.SubGroupPolicyTypeAnnotationKey: webhook.SubGroupPolicyTypeLeaderExcluded}),
                {L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0", L.GroupIndexLabelKey: "2",
                 L.GroupUniqueHashLabelKey: sha(f"{NS}/test-2")}, ""))
    # labels that are already there are kept (:103, :115, :126, :143)
    out.append((api.Pod("test-3", NS, labels={L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0", L.GroupIndexLabelKey: "7",
                                              L.GroupUniqueHashLabelKey: "given"},
                        annotations={A.SizeAnnotationKey: "2"}),
                {L.SetNameLabelKey: "test", L.WorkerIndexLabelKey: "0", L.GroupIndexLabelKey: "7", L.GroupUniqueHashLabelKey: "given"}, ""))
    return out


def run_cases(sha1_batch):
    cs = cases()
    pods = [c[0] for c in cs]
    errors = webhook.default_labels_batch(pods, sha1_batch)
    assert errors == [None] * len(pods)
    for pod, (_, want_labels, want_subdomain) in zip(pods, cs):
        assert pod.labels == want_labels, pod.name
        assert pod.subdomain == want_subdomain


def test_default_labels_batch_on_reference_vectors():
    run_cases(oracle.sha1)


def test_gen_group_unique_key_kats():  # pod_webhook_test.go:29-53
    pods = [api.Pod(name, ns, labels={L.SetNameLabelKey: "x", L.WorkerIndexLabelKey: "0", L.GroupIndexLabelKey: "0"},
                    annotations={A.SizeAnnotationKey: "1"})
            for name, ns in (("test-sample", "default"), ("podName", "default"), ("test-sample", "leaderworkerset"))]
    assert webhook.default_labels_batch(pods, oracle.sha1) == [None] * 3
    assert [p.labels[L.GroupUniqueHashLabelKey] for p in pods] == [
        "95e88034e460983f51a9952fe128729fbc0663b5", "390b34ab671d29e9997d7d4252b8bbf8da02f5b7",
        "39f5d7e9122b9d94d3932e3720b43fd3b56347e8"]


@pytest.mark.parametrize("pod_count,sub_size,worker,want", [(4, 2, 2, "1"), (5, 2, 2, "0")])
def test_get_sub_group_index(pod_count, sub_size, worker, want):  # pod_webhook_test.go:272-305
    assert webhook.get_sub_group_index(pod_count, sub_size, worker) == want
    assert str(oracle.lib().lwso_sub_group_index(pod_count, sub_size, worker)) == want


def test_error_paths():
    pods = [api.Pod("a-0", NS, labels={L.SetNameLabelKey: "a", L.WorkerIndexLabelKey: "0"}),  # no size annotation (:92-95)
            api.Pod("a-0", NS, labels={L.SetNameLabelKey: "a", L.WorkerIndexLabelKey: "0"}, annotations={A.SizeAnnotationKey: "x"}),
            api.Pod("noordinal", NS, labels={L.SetNameLabelKey: "a", L.WorkerIndexLabelKey: "0"}, annotations={A.SizeAnnotationKey: "2"}),
            api.Pod("noordinal", NS, labels={L.SetNameLabelKey: "a"}, annotations={A.SizeAnnotationKey: "2"}),
            api.Pod("a-0-1", NS, labels={L.SetNameLabelKey: "a"}, annotations={A.SizeAnnotationKey: "2", A.SubGroupSizeAnnotationKey: "z"})]
    errs = webhook.default_labels_batch(pods, oracle.sha1)
    assert errs[0] == "size annotation is unexpectedly missing for pod a-0"
    assert "invalid syntax" in errs[1]
    assert errs[2] == errs[3] == "parsing pod ordinal for pod noordinal"
    assert "invalid syntax" in errs[4] and pods[4].labels[L.WorkerIndexLabelKey] == "1"
