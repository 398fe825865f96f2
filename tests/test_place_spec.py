"""The placement spec oracle (build-defined, parity unpinned): sanity properties
that the sequential statement must have, on CPU."""
import numpy as np

import oracle
from lws_b200 import encoder, synth
from lws_b200 import records as R


def _setup(seed=5, scale=0.004, p_exclusive=0.5, unsched=0.5):
    p = synth.profile("C3", scale)
    p.p_exclusive, p.p_leader_unscheduled, p.n_nodes, p.size_choices = p_exclusive, unsched, 3200, (16,)
    t = synth.make(p, seed=seed)
    reqs = encoder.encode_place_requests(t.lws, t.groups)
    _, _, occ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, want_occupancy=True)
    return t, reqs, occ


def test_exclusive_one_group_per_domain_and_capacity():
    t, reqs, occ = _setup()
    out = oracle.place(t.nodes, occ * 0, t.n_domains, 1, reqs)
    placed = (out["flags"] & R.PLACE_PLACED) != 0
    doms = out["domain_id"][placed]
    assert len(doms) == len(set(doms.tolist())), "two groups hold one domain"
    # every unpinned placement has room for the whole group and a free node for the leader
    free = np.where((t.nodes["flags"] & 3) == 3, t.nodes["capacity"].astype(np.int64), 0)
    dom_free = np.bincount(t.nodes["domain_id"][free > 0], weights=free[free > 0], minlength=t.n_domains)
    unp = placed & ((out["flags"] & R.PLACE_PINNED) == 0)
    assert unp.sum() > 0
    assert (dom_free[out["domain_id"][unp]] >= reqs["size"][unp]).all()
    assert (free[out["leader_node"][unp]] >= 1).all()
    assert (t.nodes["domain_id"][out["leader_node"][unp]] == out["domain_id"][unp]).all()
    # a pinned group either holds its leader's domain or is flagged as a conflict
    pin = (out["flags"] & R.PLACE_PINNED) != 0
    ok = ((out["flags"] & (R.PLACE_PLACED | R.PLACE_CONFLICT)) != 0) | (out["domain_id"] == R.NONE)
    assert ok[pin].all()


def test_priority_order_is_respected():
    """Removing the lowest-priority request never changes anyone else's result."""
    t, reqs, occ = _setup(seed=9)
    full = oracle.place(t.nodes, None, t.n_domains, 1, reqs)
    unp = np.flatnonzero(reqs["leader_node"] == R.NONE)
    worst = unp[np.argmax((reqs["priority"][unp] >> np.uint64(25)))]
    keep = np.ones(len(reqs), bool)
    keep[worst] = False
    reqs2 = R.aligned_empty(int(keep.sum()), R.PLACE_REQ)
    reqs2[:] = reqs[keep]
    sub = oracle.place(t.nodes, None, t.n_domains, 1, reqs2)
    # index tie-breaks may shift by one, so compare the placements of everyone but the removed one
    assert np.array_equal(sub["domain_id"], full["domain_id"][keep])


def test_unschedulable_when_capacity_is_short():
    t, reqs, occ = _setup(seed=3)
    full_occ = np.full(len(t.nodes), 1000, dtype=np.uint32)
    out = oracle.place(t.nodes, full_occ, t.n_domains, 1, reqs)
    unp = reqs["leader_node"] == R.NONE
    assert ((out["flags"][unp] & R.PLACE_UNSCHEDULABLE) != 0).all()


# (group_key low word, d1, d2, shared score): mix(key ^ d1·φ)|1 == mix(key ^ d2·φ)|1, found by
# inverting the mixer (tests/golden/find_place_ties.py)
PLACE_TIES = [(0x5DD5FF0E, 24, 56, 0x3337DF), (0x0ECD4C3C, 11, 23, 0x43259F), (0xA7E9EB8E, 26, 35, 0x7036E3),
              (0x5D7C00AC, 31, 36, 0x812E1B), (0xD88EC2CA, 6, 41, 0x932713)]


def tie_case(key_lo, d1, d2, n_domains=64, nodes_per_domain=4):
    """Only d1 and d2 have room; the request scores both the same."""
    nodes = R.aligned_empty(n_domains * nodes_per_domain, R.NODE_REC)
    nodes["topo_value_hash"] = 0
    nodes["domain_id"] = np.arange(len(nodes)) % n_domains  # interleaved: runs are not contiguous
    nodes["capacity"] = np.where(np.isin(nodes["domain_id"], [d1, d2]), 8, 0)
    nodes["flags"] = R.NODE_HAS_TOPOLOGY | R.NODE_SCHEDULABLE
    reqs = R.aligned_empty(3, R.PLACE_REQ)
    reqs["priority"] = [5 << 25, 1 << 25, 9 << 25]
    reqs["group_key"] = [key_lo | (0xABCDEF01 << 32), 0x1111, 0x2222]  # only row 0 carries the tie
    reqs["group"] = [0, 1, 2]
    reqs["ns"] = 0
    reqs["size"] = [4, 100, 4]  # row 1 fits nowhere
    reqs["leader_node"] = R.NONE
    return nodes, reqs


def test_equal_domain_scores_keep_the_lower_domain():
    for key_lo, d1, d2, score in PLACE_TIES:
        nodes, reqs = tie_case(key_lo, d1, d2)
        out = oracle.place(nodes, None, 64, 1, reqs)
        assert out["flags"][0] == R.PLACE_PLACED and out["domain_id"][0] == d1 and out["score"][0] == score
        assert nodes["domain_id"][out["leader_node"][0]] == d1
        assert out["flags"][1] == R.PLACE_UNSCHEDULABLE
        assert out["flags"][2] == R.PLACE_PLACED and out["domain_id"][2] == d2  # what is left
