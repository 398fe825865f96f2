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
