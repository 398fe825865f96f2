"""The incremental encoder: after any stream of watch events the slotted resident
tables — maintained by row patches only — give the oracle exactly the results a fresh full encode of the
same objects gives, object by object and group by group; the patches are the only thing that changed."""
import numpy as np
import pytest

import oracle
from informer_world import World, make_world, outputs_by_name
from lws_b200 import encoder, informer
from lws_b200 import records as R


def sweep(lws, groups, pst, pid, nodes):
    lo, go, _ = oracle.sweep_lws(lws, groups, pst, pid, nodes, flags=R.SWEEP_GANG)
    return lo, go


def fresh_results(world, topology_key):
    items = list(world.items.values())
    t = encoder.encode_lws(items, world.cluster(), topology_key)
    lo, go = sweep(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes)
    names = {i: (it.lws.namespace, it.lws.name) for i, it in enumerate(items)}
    return outputs_by_name(t.lws, t.groups, lo, go, names), t


def resident_results(enc, mirror):
    lws, groups, pst, pid, _ = mirror
    lo, go = sweep(lws, groups, pst, pid, enc.node_rec)
    names = {sl.lws_row: key for key, sl in enc.slots.items()}
    return outputs_by_name(lws, groups, lo, go, names)


def strip_layout(group_bytes_by_key):
    return group_bytes_by_key


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_event_stream_equals_fresh_encode(seed):
    items, cluster = make_world(seed)
    world = World(items, cluster, seed + 10)
    enc = informer.IncrementalEncoder(list(world.items.values()), world.cluster(), "zone")
    # the engine's side: copies of the tables, changed by the emitted patches only
    mirror = [a.copy() for a in enc.full_tables()]
    by_table = {R.TABLE_LWS: 0, R.TABLE_GROUPS: 1, R.TABLE_POD_STATE: 2, R.TABLE_POD_IDENT: 3, R.TABLE_PLACE_REQS: 4}
    (want_l, want_g), _ = fresh_results(world, "zone")
    got_l, got_g = resident_results(enc, mirror)
    assert got_l == want_l and got_g == want_g
    n_patched = 0
    for tick in range(40):
        for _ in range(int(world.rng.integers(1, 6))):
            world.step(enc)
        patches = enc.flush()
        assert not enc.needs_reload
        for table, rows, vals in patches.segments:
            assert len(np.unique(rows)) == len(rows)
            mirror[by_table[table]][rows] = vals
            n_patched += len(rows)
        for a, b in zip(mirror, enc.full_tables()):
            assert a.tobytes() == b.tobytes()  # the patches carry every change
        (want_l, want_g), fresh = fresh_results(world, "zone")
        got_l, got_g = resident_results(enc, mirror)
        assert got_l == want_l, f"tick {tick}"
        # node occupancy counted over EVERY identity row of the slotted table (what the engine does)
        # = the fresh encode's: rows an object vacated were patched back to zero
        assert np.array_equal(R.occupancy_of(mirror[3], len(enc.node_rec)), R.occupancy_of(fresh.pod_ident, len(fresh.nodes))), f"tick {tick}"
        # group rows: compare what both have; a fresh encode has exactly the referenced rows
        assert got_g.keys() == want_g.keys(), f"tick {tick}"
        for k in want_g:
            a, b = np.frombuffer(got_g[k], R.GROUP_OUT)[0], np.frombuffer(want_g[k], R.GROUP_OUT)[0]
            assert a.tobytes() == b.tobytes(), f"tick {tick} {k}"
    assert n_patched > 0


def test_requests_follow_the_groups_and_stay_grouped():
    items, cluster = make_world(5)
    world = World(items, cluster, 6)
    enc = informer.IncrementalEncoder(list(world.items.values()), world.cluster(), "zone")
    for _ in range(60):
        world.step(enc)
    enc.flush()
    lws, groups, pst, pid, reqs = enc.full_tables()
    live = reqs[reqs["size"] >= 1]
    # exactly the groups of exclusive objects whose leader pod exists have a live request, pointing at their row
    want = set()
    for key, sl in enc.slots.items():
        row = sl.lws_row
        if not (int(lws["flags"][row]) & R.LWS_EXCLUSIVE_TOPOLOGY):
            continue
        for gi in range(int(lws["group_count"][row])):
            if int(groups["flags"][sl.group_base + gi]) & R.GRP_POD_PRESENT:
                want.add(sl.group_base + gi)
    assert set(live["group"].tolist()) == want
    assert np.array_equal(groups["leader_node"][live["group"]], live["leader_node"])
    ns = reqs["ns"].astype(np.int64)
    assert np.all(np.diff(ns[: enc.r_end]) >= 0) or enc.r_end > sum(s.group_cap for s in enc.slots.values()) - 1


def test_overflow_relocates_then_asks_for_a_reload():
    from lws_b200 import api

    items, cluster = make_world(7, n_lws=3)
    world = World(items, cluster, 8)
    enc = informer.IncrementalEncoder(list(world.items.values()), world.cluster(), "zone", spare=0.5)
    it = list(world.items.values())[0]
    sl0 = enc.slots[(it.lws.namespace, it.lws.name)]
    # many new groups appear for one object: first it moves to the spare rows …
    for g in range(sl0.group_cap + 1):
        p = api.Pod(name=f"{it.lws.name}-{g}", namespace=it.lws.namespace, phase="Running",
                    labels={api.SetNameLabelKey: it.lws.name, api.GroupIndexLabelKey: str(g), api.WorkerIndexLabelKey: "0"})
        world.pods[(p.namespace, p.name)] = p
        enc.pod_event("ADDED", p)
    enc.flush()
    sl1 = enc.slots[(it.lws.namespace, it.lws.name)]
    assert not enc.needs_reload and sl1.group_base != sl0.group_base
    (want_l, want_g), _ = fresh_results(world, "zone")
    got_l, got_g = resident_results(enc, list(enc.full_tables()))
    assert got_l == want_l and got_g == want_g
    assert (enc.reqs[sl0.req_base: sl0.req_base + sl0.group_cap]["size"] == 0).all()  # the vacated requests are inert
    # … and when even the spare rows are used up the encoder says so instead of writing out of range
    for g in range(400):
        p = api.Pod(name=f"{it.lws.name}-{g}", namespace=it.lws.namespace, phase="Running",
                    labels={api.SetNameLabelKey: it.lws.name, api.GroupIndexLabelKey: str(g), api.WorkerIndexLabelKey: "0"})
        enc.pod_event("ADDED", p)
    enc.flush()
    assert enc.needs_reload
