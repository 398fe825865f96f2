"""Revision detection against pkg/utils/revision/revision_utils_test.go:
TestApplyRevision (:33-80), the eight TestEqualRevision entries (:82-180), TestSetMatchesRevision
(:182-223), TestGetHighestRevision (:225-270) — replayed on the objects' JSON — plus the pieces the
reference takes from its dependencies: Go's encoding/json rules, FNV-1 and rand.SafeEncodeString."""
import copy
import json

import pytest

from lws_b200 import revision as rv

IMG = "docker.io/nginxinc/nginx-unprivileged:1.27"
# test/wrappers/wrappers.go:279-294, :360-369, :423-468 as the API serves them (omitempty applied;
# a corev1.Container always carries "resources": {})
WORKER = {"containers": [{"name": "worker", "image": IMG, "ports": [{"containerPort": 8080, "protocol": "TCP"}], "resources": {}}]}
LEADER = {"containers": [{"name": "leader", "image": IMG, "resources": {}}]}
WITH_VOLUME = {"containers": [{"name": "leader", "image": IMG, "ports": [{"containerPort": 8080, "protocol": "TCP"}], "resources": {}}],
               "volumes": [{"name": "dshm"}]}
WITH_VOLUME_NIL_IMAGE = copy.deepcopy(WITH_VOLUME)  # VolumeSource{Image: nil}: omitempty drops the field


def build_lws(worker=WORKER, leader=LEADER, subdomain="Shared", uid="", generation=0):
    """wrappers.BuildLeaderWorkerSet("default") (:226-254) as a JSON object."""
    spec = {
        "replicas": 2,
        "leaderWorkerTemplate": {"restartPolicy": "RecreateGroupOnPodRestart", "size": 2,
                                 "leaderTemplate": {"metadata": {"creationTimestamp": None}, "spec": copy.deepcopy(leader)},
                                 "workerTemplate": {"metadata": {"creationTimestamp": None}, "spec": copy.deepcopy(worker)}},
        "rolloutStrategy": {"type": "RollingUpdate",
                            "rollingUpdateConfiguration": {"partition": 0, "maxUnavailable": 1, "maxSurge": 0}},
        "startupPolicy": "LeaderCreated",
    }
    if subdomain is not None:
        spec["networkConfig"] = {"subdomainPolicy": subdomain}
    md = {"name": "test-sample", "namespace": "default", "creationTimestamp": None}
    if uid:
        md["uid"] = uid
    if generation:
        md["generation"] = generation
    return {"metadata": md, "spec": spec, "status": {}}


def test_apply_revision():
    lws = build_lws()
    revision = rv.new_revision(lws)
    current = copy.deepcopy(lws)
    lws["spec"]["leaderWorkerTemplate"]["leaderTemplate"]["spec"]["containers"][0]["name"] = "update-name"
    lws["spec"]["networkConfig"] = {"subdomainPolicy": "UniquePerReplica"}
    lws["spec"]["rolloutStrategy"] = {"type": "RollingUpdate", "rollingUpdateConfiguration": {"maxUnavailable": 2, "maxSurge": 1}}
    restored = rv.apply_revision(lws, revision)
    assert rv.equal_revision(revision, rv.new_revision(restored))
    assert restored["spec"]["leaderWorkerTemplate"] == current["spec"]["leaderWorkerTemplate"]
    assert restored["spec"]["networkConfig"] == current["spec"]["networkConfig"]
    assert restored["spec"]["rolloutStrategy"] == lws["spec"]["rolloutStrategy"]  # not part of a revision


@pytest.mark.parametrize(
    "left,right,lkey,rkey,equal",
    [(build_lws(), build_lws(), "", "", True),
     (build_lws(), build_lws(), "", "templateHash", True),
     (build_lws(subdomain="Shared"), build_lws(subdomain=None), "", "", True),  # nil network config defaults to Shared
     (None, None, "", "", True),
     (build_lws(worker=WITH_VOLUME_NIL_IMAGE), build_lws(worker=WITH_VOLUME), "", "", True),
     (None, build_lws(), "", "", False),
     (build_lws(subdomain="UniquePerReplica"), build_lws(), "", "", False),
     (build_lws(), build_lws(worker=LEADER), "", "", False)],
    ids=["same", "different revision key", "shared vs nil subdomain policy", "nil nil", "semantically same",
         "nil vs set", "different network config", "different template"])
def test_equal_revision(left, right, lkey, rkey, equal):
    l = rv.new_revision(left, lkey) if left is not None else None
    r = rv.new_revision(right, rkey) if right is not None else None
    assert rv.equal_revision(l, r) is equal
    if l is not None and r is not None and equal:
        assert l.name == r.name  # same bytes → same hash → same name
        assert (l.key == r.key) == (lkey == rkey)


def test_set_matches_revision():
    lws = build_lws(uid="test-uid", generation=1)
    revision = rv.new_revision(lws)
    revision.resourceVersion = "100"
    proposed = rv.new_revision(lws)
    cache = rv.RevisionEqualityCache(10)
    assert rv.set_matches_revision(lws, proposed, revision, cache) and len(cache) == 1
    proposed.raw = b'{"spec":{"leaderWorkerTemplate":{"$patch":"replace"}}}'
    assert rv.set_matches_revision(lws, proposed, revision, cache)  # cache hit
    lws["metadata"]["generation"] = 2
    assert not rv.set_matches_revision(lws, proposed, revision, cache)


def test_old_serialisation_is_a_semantic_match():
    """leaderworkerset_controller_test.go:1097: a revision stored by an older client ("creationTimestamp":
    null inside the templates) must not look like a template update."""
    lws = build_lws(uid="u", generation=3)
    today = copy.deepcopy(lws)
    for t in ("leaderTemplate", "workerTemplate"):
        today["spec"]["leaderWorkerTemplate"][t]["metadata"] = {}
    stored = rv.new_revision(lws)      # old bytes (with creationTimestamp: null)
    stored.resourceVersion = "7"
    proposed = rv.new_revision(today)  # today's bytes
    assert not rv.equal_revision(proposed, stored)
    cache = rv.RevisionEqualityCache()
    # ApplyRevision(today, stored) re-serialised gives the stored bytes, not today's: still an update …
    assert rv.get_updated_revision(today, True, stored, cache) is not None
    # … unless today's object is what the stored patch restores to
    assert rv.get_updated_revision(lws, True, stored, cache) is None
    assert rv.updated_bits([(lws, True, stored), (today, True, stored), (today, False, stored)]) == [False, True, False]


def test_get_highest_revision():
    mk = lambda n: rv.ControllerRevision(f"r{n}", "default", {}, n, b"{}")
    assert rv.get_highest_revision([]) is None
    assert rv.get_highest_revision([mk(1)]).revision == 1
    assert rv.get_highest_revision([mk(1), mk(3), mk(2)]).revision == 3
    assert rv.new_revision(build_lws(), existing=[mk(1), mk(3), mk(2)]).revision == 4


def test_get_patch_bytes_follow_gos_json_encoder():
    lws = build_lws(subdomain=None)
    raw = rv.get_patch(lws)
    obj = json.loads(raw)
    assert list(obj) == ["spec"] and list(obj["spec"]) == ["leaderWorkerTemplate", "networkConfig"]  # sorted keys
    assert obj["spec"]["networkConfig"] == {"$patch": "replace", "subdomainPolicy": "Shared"}
    assert raw.startswith(b'{"spec":{"leaderWorkerTemplate":{"$patch":"replace","leaderTemplate":{"metadata":{"creationTimestamp":null}')
    assert b" " not in raw.replace(b"nginx-unprivileged", b"")
    # Go escapes <, >, & and U+2028 in strings; numbers go through float64
    assert rv.go_json_marshal({"b": "<a&b> ", "a": [1, 1.5, 1e21, 1e-7, 100.0, -0.0]}) == \
        '{"a":[1,1.5,1e+21,1e-7,100,-0],"b":"\\u003ca\\u0026b\\u003e\\u2028"}'
    assert rv.go_json_marshal({"x": 123456789012, "y": 0.000001, "z": 12345.678}) == '{"x":123456789012,"y":0.000001,"z":12345.678}'


def test_fnv1_and_safe_encode():
    # FNV-1 32-bit published test values (draft-eastlake-fnv): "" and "a" and "foobar"
    assert rv.fnv1_32(b"") == 0x811C9DC5
    assert rv.fnv1_32(b"a") == 0x050C5D7E
    assert rv.fnv1_32(b"foobar") == 0x31F0B262
    # rand.SafeEncodeString: c -> alphabet[c % 27]; digits '0'..'9' = 48..57 → 21..26, 0..3
    assert rv.safe_encode_string("0123456789") == "".join(rv.SAFE_ALPHANUMS[(48 + i) % 27] for i in range(10))
    assert set(rv.hash_revision(b"abc")) <= set(rv.SAFE_ALPHANUMS)
    assert rv.revision_name("x" * 300, "h", 2) == "x" * 220 + "-h-2"
