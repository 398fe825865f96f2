"""The resident tick (lwse_resident_tick) on the GPU: watch-event patches in, changed result
rows out — every tick's full outputs and change lists against the oracle run on host mirrors."""
import numpy as np
import pytest

from lws_b200 import churn
from lws_b200 import records as R
from lws_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from lws_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def _rows_differ(a, b):
    return np.flatnonzero(np.any(a.view(np.uint8).reshape(len(a), -1) != b.view(np.uint8).reshape(len(b), -1), axis=1))


def _check_changes(res, prefix, before, after):
    want = _rows_differ(before, after)
    rows, outs = res[f"{prefix}_rows"], res[f"{prefix}_out"]
    assert res[f"n_{'groups' if prefix == 'group' else prefix}"] == len(want), prefix
    o = np.argsort(rows)
    assert np.array_equal(rows[o], want), prefix
    assert outs[o].tobytes() == after[want].tobytes(), prefix


@pytest.mark.parametrize("in_arena", [True, False])
def test_ticks_match_the_oracle(engine, in_arena):
    _ticks_match_the_oracle(engine, in_arena)


def test_graph_replayed_ticks_match_the_oracle(monkeypatch):
    """LWSE_TICK_GRAPH=2: every eligible tick — also a lone one — is replayed as a CUDA graph (scatter
    descriptors through the pinned slot, device-side sequence counter).  Same checks, every tick."""
    from lws_b200.engine import Engine

    monkeypatch.setenv("LWSE_TICK_GRAPH", "2")
    e = Engine(0)
    try:
        _ticks_match_the_oracle(e, True, n_ticks=14)
    finally:
        e.close()


def _ticks_match_the_oracle(engine, in_arena, n_ticks=9):
    import oracle

    p = synth.profile("fuzz", 0.5)
    p.n_namespaces = 3
    t = synth.make(p, seed=41)
    reqs = t.place_requests()
    assert len(reqs) > 100
    engine.upload_nodes(t.nodes, t.n_domains)
    engine.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
    engine.resident_place_load(reqs, t.n_namespaces)
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    assert np.array_equal(engine.resident_occupancy(), occ)
    flags = t.flags | R.TICK_PLACE
    m_pst, m_grp, m_req = t.pod_state.copy(), t.groups.copy(), reqs.copy()

    def oracle_now():
        lo, go, _ = oracle.sweep_lws(t.lws, m_grp, m_pst, t.pod_ident, t.nodes, flags=t.flags)
        return lo, go, oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, m_req)

    prev = oracle_now()
    first = engine.resident_tick(engine.make_tick((), flags))
    assert first["n_lws"] == len(t.lws) and first["n_groups"] == len(t.groups) and first["n_place"] == len(reqs)
    assert engine.resident_place_outputs().tobytes() == prev[2].tobytes()
    again = engine.resident_tick(engine.make_tick((), flags))
    assert (again["n_lws"], again["n_groups"], again["n_place"]) == (0, 0, 0)

    plan = churn.make_plan(t, reqs, prev[2], 0.03, 0.05, n_sets=6, seed=5)
    if in_arena:
        ticks = churn.ArenaPlan(engine, plan, flags).ticks
    else:  # caller-owned pageable buffers: the call stages them
        ticks = []
        for ps in plan:
            segs = [(R.TABLE_POD_STATE, ps.pod_rows, ps.pod_vals)]
            if len(ps.grp_rows):
                segs.append((R.TABLE_GROUPS, ps.grp_rows, ps.grp_vals))
            if len(ps.req_rows):
                segs.append((R.TABLE_PLACE_REQS, ps.req_rows, ps.req_vals))
            ticks.append(engine.make_tick(segs, flags))
    for k in range(n_ticks):
        ps = plan[k % len(plan)]
        churn.apply_to_mirror(ps, m_pst, m_grp, m_req)
        now = oracle_now()
        res = engine.resident_tick(ticks[k % len(plan)])
        g_lo, g_go = engine.resident_outputs()
        assert g_lo.tobytes() == now[0].tobytes() and g_go.tobytes() == now[1].tobytes(), f"tick {k}"
        assert engine.resident_place_outputs().tobytes() == now[2].tobytes(), f"tick {k} placement"
        _check_changes(res, "lws", prev[0], now[0])
        _check_changes(res, "group", prev[1], now[1])
        _check_changes(res, "place", prev[2], now[2])
        assert res["rounds"] >= 1
        prev = now


def test_pipelined_ticks_report_what_serial_ticks_report(engine):
    """lwse_resident_tick_submit / _wait with two ticks in flight: every tick's change lists (copied
    at wait time, while the NEXT tick is already running) equal the oracle's diff of consecutive
    states, and a third submit without a wait is refused."""
    import oracle
    from lws_b200.engine import LwseError

    p = synth.profile("fuzz", 0.5)
    p.n_namespaces = 3
    t = synth.make(p, seed=47)
    reqs = t.place_requests()
    engine.upload_nodes(t.nodes, t.n_domains)
    engine.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
    engine.resident_place_load(reqs, t.n_namespaces)
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    flags = t.flags | R.TICK_PLACE
    m_pst, m_grp, m_req = t.pod_state.copy(), t.groups.copy(), reqs.copy()

    def oracle_now():
        lo, go, _ = oracle.sweep_lws(t.lws, m_grp, m_pst, t.pod_ident, t.nodes, flags=t.flags)
        return lo, go, oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, m_req)

    states = [oracle_now()]
    engine.resident_tick(engine.make_tick((), flags))
    plan = churn.make_plan(t, reqs, states[0][2], 0.03, 0.05, n_sets=6, seed=9)
    ticks = churn.ArenaPlan(engine, plan, flags).ticks  # six disjoint arena regions
    n = 12
    for k in range(n):
        churn.apply_to_mirror(plan[k % len(plan)], m_pst, m_grp, m_req)
        states.append(oracle_now())

    def check(k, res):
        before, after = states[k], states[k + 1]
        res = {key: (v.copy() if isinstance(v, np.ndarray) else v) for key, v in res.items()}
        _check_changes(res, "lws", before[0], after[0])
        _check_changes(res, "group", before[1], after[1])
        _check_changes(res, "place", before[2], after[2])

    engine.resident_tick_submit(ticks[0])
    for k in range(1, n):
        engine.resident_tick_submit(ticks[k % len(ticks)])
        if k == 3:  # two in flight: the third is refused, nothing is lost
            with pytest.raises(LwseError):
                engine.resident_tick_submit(ticks[(k + 1) % len(ticks)])
        check(k - 1, engine.resident_tick_wait())
    check(n - 1, engine.resident_tick_wait())
    with pytest.raises(LwseError):
        engine.resident_tick_wait()
    g_lo, g_go = engine.resident_outputs()
    assert g_lo.tobytes() == states[-1][0].tobytes() and g_go.tobytes() == states[-1][1].tobytes()
    assert engine.resident_place_outputs().tobytes() == states[-1][2].tobytes()
    # a synchronous call with a tick in flight drains it first
    engine.resident_tick_submit(ticks[0])
    again = engine.resident_tick(engine.make_tick((), flags))
    assert (again["n_lws"], again["n_groups"], again["n_place"]) == (0, 0, 0)


def test_range_patch_and_identity_patches_keep_the_occupancy(engine):
    import oracle

    t = synth.make("fuzz", 0.3, seed=43)
    engine.upload_nodes(t.nodes, t.n_domains)
    engine.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
    engine.resident_tick(engine.make_tick((), t.flags))
    rng = np.random.default_rng(3)
    m_pst, m_pid = t.pod_state.copy(), t.pod_ident.copy()
    # (1) a whole-column range patch of the state bytes
    new = R.aligned_empty(len(m_pst), R.POD_STATE)
    new[:] = m_pst ^ np.where(rng.random(len(m_pst)) < 0.3, R.POD_ANY_RESTART, 0).astype(np.uint8)
    m_pst[:] = new
    engine.resident_tick(engine.make_tick([(R.TABLE_POD_STATE, 0, new, True)], t.flags))
    # (2) pods move: scattered identity-row patches move their occupancy count with them
    rows = np.unique(rng.integers(0, len(m_pid), size=len(m_pid) // 20)).astype(np.uint32)
    vals = R.aligned_empty(len(rows), R.POD_IDENT)
    vals[:] = m_pid[rows]
    node = rng.integers(0, len(t.nodes), size=len(rows)).astype(np.uint32)
    vals["place"] = np.where(rng.random(len(rows)) < 0.8, R.PODID_SCHEDULED | R.PODID_NAME_OK | (node << R.PODID_NODE_SHIFT),
                             R.PODID_NAME_OK).astype(np.uint32)
    vals["owner_uid_hash"] ^= (rng.random(len(rows)) < 0.1).astype(np.uint32)
    m_pid[rows] = vals
    engine.resident_tick(engine.make_tick([(R.TABLE_POD_IDENT, rows, vals)], t.flags))
    lo, go, occ = oracle.sweep_lws(t.lws, t.groups, m_pst, m_pid, t.nodes, flags=t.flags, want_occupancy=True)
    g_lo, g_go = engine.resident_outputs()
    assert g_lo.tobytes() == lo.tobytes() and g_go.tobytes() == go.tobytes()
    assert np.array_equal(engine.resident_occupancy(), occ)
    # (3) a range patch of the identity column: the engine recounts
    first = len(m_pid) // 3
    seg = R.aligned_empty(len(m_pid) - first, R.POD_IDENT)
    seg[:] = m_pid[first:]
    seg["place"] &= ~np.uint32(R.PODID_SCHEDULED)
    m_pid[first:] = seg
    engine.resident_tick(engine.make_tick([(R.TABLE_POD_IDENT, first, seg, True)], t.flags))
    assert np.array_equal(engine.resident_occupancy(), R.occupancy_of(m_pid, len(t.nodes)))


def test_watch_events_pipelined_through_the_arena(engine):
    """The informer's patch segments written into alternating halves of the engine's arena and submitted
    with two ticks in flight (the later ones replayed as graphs): the final resident state equals the
    oracle on a fresh encode of the same objects, by name — sweep rows, occupancy, placement rows."""
    import oracle
    from informer_world import World, make_world, outputs_by_name
    from lws_b200 import encoder, informer

    items, cluster = make_world(5, n_lws=24, n_nodes=40)
    world = World(items, cluster, 6)
    enc = informer.IncrementalEncoder(list(world.items.values()), world.cluster(), "zone")
    lws, groups, pst, pid, reqs = enc.full_tables()
    engine.upload_nodes(enc.node_rec, len(enc.domain_values))
    engine.resident_load(lws, groups, pst, pid)
    engine.resident_place_load(reqs, enc.n_namespaces)
    flags = R.SWEEP_GANG | R.TICK_PLACE
    engine.resident_tick(engine.make_tick((), flags))
    half = 1 << 20
    arena = engine.resident_arena(2 * half)

    def in_arena(segments, which):
        cursor, out = which * half, []
        for table, rows, vals in segments:
            views = []
            for a in (np.ascontiguousarray(rows, dtype=np.uint32), np.ascontiguousarray(vals)):
                cursor = (cursor + 255) // 256 * 256
                v = arena[cursor: cursor + a.nbytes].view(a.dtype)
                v[...] = a
                cursor += a.nbytes
                views.append(v)
            out.append((table, views[0], views[1]))
        assert cursor <= (which + 1) * half
        return out

    n_ticks, in_flight = 24, 0
    for tick in range(n_ticks):
        for _ in range(int(world.rng.integers(1, 6))):
            world.step(enc)
        patches = enc.flush()
        assert not enc.needs_reload
        t = engine.make_tick(in_arena(patches.segments, tick & 1), flags)
        if in_flight == 2:  # the tick that used this half of the arena two ticks ago
            engine.resident_tick_wait()
            in_flight -= 1
        engine.resident_tick_submit(t)
        in_flight += 1
    while in_flight:
        engine.resident_tick_wait()
        in_flight -= 1
    names = {sl.lws_row: key for key, sl in enc.slots.items()}
    g_lo, g_go = engine.resident_outputs()
    got_l, got_g = outputs_by_name(enc.lws[: enc.n_lws], enc.groups, g_lo, g_go, names)
    fresh_items = list(world.items.values())
    ft = encoder.encode_lws(fresh_items, world.cluster(), "zone")
    w_lo, w_go, w_occ = oracle.sweep_lws(ft.lws, ft.groups, ft.pod_state, ft.pod_ident, ft.nodes, flags=R.SWEEP_GANG,
                                         want_occupancy=True)
    want_l, want_g = outputs_by_name(ft.lws, ft.groups, w_lo, w_go,
                                     {i: (it.lws.namespace, it.lws.name) for i, it in enumerate(fresh_items)})
    assert got_l == want_l and got_g == want_g
    assert np.array_equal(engine.resident_occupancy(), w_occ)
    want_p = oracle.place(enc.node_rec, w_occ, len(enc.domain_values), enc.n_namespaces, enc.reqs)
    assert engine.resident_place_outputs().tobytes() == want_p.tobytes()


def test_ds_sweep_c4_full_size(engine):
    """BASELINE.json configs[3]: 50k two-role DisaggregatedSets, every output against the oracle."""
    import oracle

    d = synth.make_ds(50_000, (2,))
    got = engine.sweep_ds_host(d.ds, d.roles, d.revroles)
    want = oracle.sweep_ds(d.ds, d.roles, d.revroles)
    for a, b, name in zip(got, want, ("ds_out", "role_out", "revrole_out")):
        assert a.tobytes() == b.tobytes(), name


def test_watch_events_through_the_incremental_encoder(engine):
    """Watch events → lws_b200.informer.IncrementalEncoder → patch segments → lwse_resident_tick (sweep +
    placement): after every tick the engine's resident results equal the oracle's on a FRESH full encode of
    the same objects, object by object (the two layouts differ; results are matched by name)."""
    import oracle
    from informer_world import World, make_world, outputs_by_name
    from lws_b200 import encoder, informer

    items, cluster = make_world(3, n_lws=20, n_nodes=40)
    world = World(items, cluster, 4)
    enc = informer.IncrementalEncoder(list(world.items.values()), world.cluster(), "zone")
    lws, groups, pst, pid, reqs = enc.full_tables()
    engine.upload_nodes(enc.node_rec, len(enc.domain_values))
    engine.resident_load(lws, groups, pst, pid)
    engine.resident_place_load(reqs, enc.n_namespaces)
    flags = R.SWEEP_GANG | R.TICK_PLACE
    engine.resident_tick(engine.make_tick((), flags))
    names = {sl.lws_row: key for key, sl in enc.slots.items()}
    for tick in range(25):
        for _ in range(int(world.rng.integers(1, 8))):
            world.step(enc)
        patches = enc.flush()
        assert not enc.needs_reload
        engine.resident_tick(engine.make_tick(patches.segments, flags))
        g_lo, g_go = engine.resident_outputs()
        got_l, got_g = outputs_by_name(enc.lws[: enc.n_lws], enc.groups, g_lo, g_go, names)
        fresh_items = list(world.items.values())
        ft = encoder.encode_lws(fresh_items, world.cluster(), "zone")
        w_lo, w_go, w_occ = oracle.sweep_lws(ft.lws, ft.groups, ft.pod_state, ft.pod_ident, ft.nodes, flags=R.SWEEP_GANG,
                                             want_occupancy=True)
        want_l, want_g = outputs_by_name(ft.lws, ft.groups, w_lo, w_go,
                                         {i: (it.lws.namespace, it.lws.name) for i, it in enumerate(fresh_items)})
        assert got_l == want_l and got_g == want_g, f"tick {tick}"
        # the engine's occupancy counters followed the identity-row patches
        assert np.array_equal(engine.resident_occupancy(), w_occ), f"tick {tick}"
        # placement over the resident request table = the spec oracle on the same rows
        want_p = oracle.place(enc.node_rec, w_occ, len(enc.domain_values), enc.n_namespaces, enc.reqs)
        assert engine.resident_place_outputs().tobytes() == want_p.tobytes(), f"tick {tick} placement"
