"""The churn plan both bench arms digest, on the CPU: applying a tick's patches and reconciling
only the dirty rows the plan names gives the same tables as a full sweep (so the CPU arm of
bench.py does all the work a tick requires, and no more)."""
import numpy as np

import oracle
from lws_b200 import churn
from lws_b200 import records as R
from lws_b200 import synth


def test_dirty_sets_cover_every_changed_result():
    p = synth.profile("fuzz", 0.4)
    p.n_namespaces = 4
    t = synth.make(p, seed=21)
    reqs = t.place_requests()
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    base_place = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, reqs)
    plan = churn.make_plan(t, reqs, base_place, 0.02, 0.04, n_sets=4, seed=2)
    pst, grp, rq = t.pod_state.copy(), t.groups.copy(), reqs.copy()
    lo, go, _ = oracle.sweep_lws(t.lws, grp, pst, t.pod_ident, t.nodes, flags=t.flags)
    n_unpinned0 = int((rq["leader_node"] == R.NONE).sum())
    for k in range(8):
        ps = plan[k % len(plan)]
        assert len(ps.pod_rows) > 0 and len(ps.req_rows) > 0
        oracle.apply_patch(pst, ps.pod_rows, ps.pod_vals)
        oracle.apply_patch(grp, ps.grp_rows, ps.grp_vals)
        oracle.apply_patch(rq, ps.req_rows, ps.req_vals)
        oracle.sweep_dirty(t.lws, grp, pst, t.pod_ident, t.nodes, lo, go, ps.dirty_groups, ps.dirty_lws, flags=t.flags, threads=3)
        f_lo, f_go, _ = oracle.sweep_lws(t.lws, grp, pst, t.pod_ident, t.nodes, flags=t.flags)
        assert lo.tobytes() == f_lo.tobytes() and go.tobytes() == f_go.tobytes(), f"set {k}"
        assert np.array_equal(grp["leader_node"][rq["group"]], rq["leader_node"])  # group rows follow the requests
    assert int((rq["leader_node"] == R.NONE).sum()) == n_unpinned0  # odd sets undo even sets


def test_threaded_placement_equals_sequential():
    p = synth.profile("fuzz", 0.5)
    p.n_namespaces = 7
    t = synth.make(p, seed=22)
    reqs = t.place_requests()
    reqs["ns"][::11] = 9  # out-of-range namespaces are unschedulable in either form
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    a = oracle.place(t.nodes, occ, t.n_domains, 7, reqs, threads=1)
    b = oracle.place(t.nodes, occ, t.n_domains, 7, reqs, threads=5)
    assert a.tobytes() == b.tobytes() and (a["flags"] & R.PLACE_UNSCHEDULABLE).any()


def test_c3_profile_is_feasible_and_conflict_free():
    t = synth.make("C3", 0.05)
    reqs = t.place_requests()
    assert len(reqs) == (t.groups["flags"] & R.GRP_POD_PRESENT != 0).sum()  # every group is exclusive
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    out = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, reqs)
    pinned = reqs["leader_node"] != R.NONE
    assert not (out["flags"][pinned] & R.PLACE_CONFLICT).any()
    assert 0.02 < (~pinned).mean() < 0.09
    assert ((out["flags"][~pinned] & R.PLACE_PLACED) != 0).mean() > 0.95


def test_plan_for_a_request_table_sharded_by_namespace_owner():
    """Strong scaling: the rank's request table (namespaces it owns) refers to groups of OTHER ranks' sweep
    shards — the plan then patches request rows only (patch_groups=False) and stays self-consistent."""
    import oracle
    from lws_b200 import distributed as D

    t = synth.make("C3", 0.03, seed=synth.SEED)
    world, rank = 4, 1
    s_lws, s_grp, s_pst, s_pid, lrows, _ = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)[rank]
    reqs, n_ns, _ = D.requests_of_rank(t.place_requests(), world, rank)
    shard = synth.Tables(profile=t.profile, lws=s_lws, groups=s_grp, pod_state=s_pst, pod_ident=s_pid, nodes=t.nodes,
                         n_domains=t.n_domains, flags=t.flags, ns_of_lws=t.ns_of_lws[lrows], n_namespaces=n_ns)
    assert int(reqs["group"].max()) >= len(s_grp)  # global group rows: not indices into the shard
    po = oracle.place(t.nodes, R.occupancy_of(t.pod_ident, len(t.nodes)), t.n_domains, n_ns, reqs)
    plan = churn.make_plan(shard, reqs, po, 0.01, 0.02, n_sets=4, seed=11, patch_groups=False)
    m_pst, m_grp, m_req = s_pst.copy(), s_grp.copy(), reqs.copy()
    for ps in plan:
        assert len(ps.grp_rows) == 0 and len(ps.req_rows) > 0 and len(ps.pod_rows) > 0
        churn.apply_to_mirror(ps, m_pst, m_grp, m_req)
    assert m_grp.tobytes() == s_grp.tobytes()
    # sets 0/1 and 2/3 are scheduling events and their undo: the request table is back where it started
    assert m_req.tobytes() == reqs.tobytes()
