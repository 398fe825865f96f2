"""The constraint semantics the placement spec claims to restate, as table-driven cases on the spec
oracle (the reference has no node scoring to pin the assignment itself — see DESIGN.md "Placement"):
what SetExclusiveAffinities (pkg/webhooks/pod_webhook.go:185-227) makes a scheduler honour.

  affinity      key In [group key], topologyKey           → a group's pods share ONE domain
  anti-affinity key Exists ∧ key NotIn [group key], same  → no OTHER exclusive group in that domain
  no `namespaces` on the terms                            → both hold per namespace only
  subgroup terms select on the SUBGROUP key (:132-134)    → subgroups exclude subgroups, not groups
The same cases run through the CUDA kernels in tests/test_gpu_other_paths.py (every form)."""
import numpy as np
import pytest

import oracle
from lws_b200 import encoder
from lws_b200 import records as R


def nodes(n_domains, per_domain=2, capacity=8):
    n = R.aligned_empty(n_domains * per_domain, R.NODE_REC)
    n["domain_id"] = np.arange(len(n)) // per_domain
    n["topo_value_hash"] = n["domain_id"] + 1000
    n["capacity"] = capacity
    n["flags"] = R.NODE_HAS_TOPOLOGY | R.NODE_SCHEDULABLE
    return n


def reqs(rows):
    """rows: (priority, ns, size, leader_node or None)"""
    r = R.aligned_empty(len(rows), R.PLACE_REQ)
    for i, (prio, ns, size, node) in enumerate(rows):
        r[i] = (prio << 25, 0x1234567 * (i + 1), i, ns, size, R.NONE if node is None else node)
    return r


CASES = [
    # name, n_domains, n_namespaces, requests, expected (per request: "placed" | "unschedulable" | "conflict" | ("same", i) | ("differs", i))
    ("two groups of one namespace never share a domain", 2, 1, [(1, 0, 2, None), (2, 0, 2, None)], ["placed", ("differs", 0)]),
    ("one free domain, two groups of one namespace: the second is unschedulable", 1, 1, [(1, 0, 2, None), (2, 0, 2, None)],
     ["placed", "unschedulable"]),
    ("the terms carry no namespaces: groups of DIFFERENT namespaces may share the domain", 1, 2, [(1, 0, 2, None), (1, 1, 2, None)],
     ["placed", ("same", 0)]),
    ("a scheduled leader pins its domain; an unscheduled group of the namespace goes elsewhere", 2, 1, [(9, 0, 2, 0), (1, 0, 2, None)],
     ["placed", ("differs", 0)]),
    ("… but not a group of another namespace", 1, 2, [(9, 0, 2, 0), (1, 1, 2, None)], ["placed", ("same", 0)]),
    ("two leaders already in one domain: the affinity is violated for the later one (reported, not moved)", 2, 1,
     [(1, 0, 2, 0), (2, 0, 2, 1)], ["placed", "conflict"]),
    ("a group needs its whole size in the domain (capacity is a filter)", 1, 1, [(1, 0, 64, None)], ["unschedulable"]),
]


@pytest.mark.parametrize("name,n_dom,n_ns,rows,want", CASES, ids=[c[0] for c in CASES])
def test_constraint_cases(name, n_dom, n_ns, rows, want):
    n = nodes(n_dom)
    rq = reqs(rows)
    out = oracle.place(n, None, n_dom, n_ns, rq)
    check(out, want)


def check(out, want):
    for i, w in enumerate(want):
        f = int(out["flags"][i])
        if w == "placed":
            assert f & R.PLACE_PLACED and out["domain_id"][i] != R.NONE, i
        elif w == "unschedulable":
            assert f & R.PLACE_UNSCHEDULABLE and not f & R.PLACE_PLACED, i
        elif w == "conflict":
            assert f & R.PLACE_CONFLICT and not f & R.PLACE_PLACED, i
        elif w[0] == "same":
            assert f & R.PLACE_PLACED and out["domain_id"][i] == out["domain_id"][w[1]], i
        elif w[0] == "differs":
            assert f & R.PLACE_PLACED and out["domain_id"][i] != out["domain_id"][w[1]], i


def test_pods_without_the_label_do_not_exclude():
    """`key Exists`: only pods carrying the group-key label are repelled — groups that are not exclusive make
    no request and hold nothing; they only consume capacity (through the occupancy counters)."""
    n = nodes(1, per_domain=2, capacity=4)
    occ = np.array([3, 3], dtype=np.uint32)  # non-exclusive pods fill most of the only domain
    out = oracle.place(n, occ, 1, 1, reqs([(1, 0, 2, None)]))
    assert out["flags"][0] & R.PLACE_PLACED  # 2 free slots: fits, nobody holds the domain
    out = oracle.place(n, occ, 1, 1, reqs([(1, 0, 3, None)]))
    assert out["flags"][0] & R.PLACE_UNSCHEDULABLE


@pytest.mark.parametrize("size,sg,excluded,want", [
    (5, 2, False, [(0, 3), (3, 2)]),   # (5-1) % 2 == 0: the leader is an extra pod of subgroup 0 (workers 1,2 + leader)
    (4, 2, False, [(0, 2), (2, 2)]),   # 4 % 2 == 0: workers 0,1 | 2,3
    (5, 2, True, [(1, 2), (3, 2)]),    # LeaderExcluded: the leader is in no subgroup
    (9, 4, False, [(0, 5), (5, 4)]),
    (8, 4, False, [(0, 4), (4, 4)]),
])
def test_sub_group_layout_follows_get_sub_group_index(size, sg, excluded, want):
    got = encoder.sub_group_layout(size, sg, excluded)
    assert got == want
    for w in range(1, size):  # every worker sits where the reference's index puts it
        idx = oracle.lib().lwso_sub_group_index(size, sg, w)
        first, pods = got[idx]
        assert first <= w < first + pods + (1 if (first == 0 and not excluded and (size - 1) % sg == 0) else 0)


def test_subgroup_exclusive_requests_are_their_own_class():
    """Subgroup requests live in namespace ids n_namespaces + ns: they exclude each other, never the
    group-level requests of the same namespace (different label keys in the anti-affinity selectors)."""
    from lws_b200 import synth

    p = synth.profile("fuzz", 0.05)
    p.size_choices, p.replicas_choices, p.p_exclusive, p.fuzz, p.p_short_group, p.n_nodes, p.nodes_per_domain = (5,), (2,), 1.0, 0.0, 0.0, 400, 4
    p.node_capacity, p.p_leader_unscheduled = 40, 0.5
    t = synth.make(p, seed=3)
    t.lws["subgroup_size"] = 2
    group_reqs = t.place_requests()
    sub = encoder.encode_subgroup_place_requests(t.lws, t.groups, t.pod_ident, np.ones(len(t.lws), bool), t.ns_of_lws,
                                                 t.n_namespaces)
    assert len(sub) == 2 * int(((t.groups["flags"] & R.GRP_POD_PRESENT != 0) & (t.groups["pod_count"] == 5)).sum())
    assert set(sub["size"].tolist()) == {3, 2}
    both = R.aligned_empty(len(group_reqs) + len(sub), R.PLACE_REQ)
    both[: len(group_reqs)], both[len(group_reqs):] = group_reqs, sub
    assert np.all(np.diff(both["ns"].astype(np.int64)) >= 0)
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    out = oracle.place(t.nodes, occ, t.n_domains, 2 * t.n_namespaces, both)
    alone = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, group_reqs)
    # request indices shift the tie-break field only for rows after the insert point: group rows keep theirs
    assert out[: len(group_reqs)].tobytes() == alone.tobytes()  # the subgroup class changes nothing for the groups
    placed = (out["flags"] & R.PLACE_PLACED) != 0
    sub_rows = np.arange(len(both)) >= len(group_reqs)
    for ns in np.unique(both["ns"][sub_rows]):
        sel = sub_rows & placed & (both["ns"] == ns)
        doms = out["domain_id"][sel]
        assert len(np.unique(doms)) == len(doms)  # one subgroup per domain within the class
