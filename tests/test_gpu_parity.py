"""GPU parity: the CUDA engine, called through the C ABI, against the CPU oracle
on the same seeded tables — bit-exact on every output field."""
import numpy as np
import pytest

from lws_b200 import records as R
from lws_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from lws_b200.engine import Engine

    e = Engine(0)  # raises if liblwse.so or the GPU is missing — no fallback
    yield e
    e.close()


def assert_same(got, want, what):
    assert got.dtype == want.dtype and got.shape == want.shape
    for name in got.dtype.names or [None]:
        a = got[name] if name else got
        b = want[name] if name else want
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a != b)
            i = int(bad[0])
            raise AssertionError(
                f"{what}.{name}: {len(bad)} rows differ; first row {i}: got {a[i]!r} want {b[i]!r}"
            )


def run_both(engine, t, occupancy=True):
    import oracle

    engine.upload_nodes(t.nodes, t.n_domains)
    want = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags, want_occupancy=occupancy)
    got = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags, want_occupancy=occupancy)
    assert_same(got[0], want[0], "lws_out")
    assert_same(got[1], want[1], "group_out")
    if occupancy:
        assert np.array_equal(got[2], want[2]), "node occupancy"
        # without the occupancy count the scan and the group pass run as ONE fused kernel
        # (small groups); same results
        fused = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags, want_occupancy=False)
        assert_same(fused[0], want[0], "fused.lws_out")
        assert_same(fused[1], want[1], "fused.group_out")
    return got


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz(engine, seed):
    run_both(engine, synth.make("fuzz", 1.0, seed=seed))


@pytest.mark.parametrize("size", [1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 100, 129, 513, 1025, 2049])
def test_every_tile_width(engine, size):
    """pods-per-group selects the tile width W of the group kernel (1, 2, 4, 8)."""
    p = synth.profile("fuzz", 0.2 if size <= 100 else 0.01)
    p.size_choices = (size,)
    run_both(engine, synth.make(p, seed=size))


@pytest.mark.parametrize("replicas", [(0,), (1,), (3,), (12,), (31, 32, 33), (60,), (100,), (200,), (1, 1000)])
def test_every_lws_tile_width(engine, replicas):
    p = synth.profile("fuzz", 0.1)
    p.size_choices = (2,)
    p.replicas_choices = replicas
    run_both(engine, synth.make(p, seed=7))


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C2", 1.0), ("C3", 0.1), ("C5", 0.05), ("C3-steady", 0.1)])
def test_baseline_configs(engine, name, scale):
    run_both(engine, synth.make(name, scale))


@pytest.mark.parametrize("name", ["C3", "C5"])
def test_baseline_configs_full_size(engine, name):
    """BASELINE.json's configs at scale 1.0 (100k objects; 6.3M / 12.8M pod rows): every output
    row of the fused sweep — and, for C3, of the placement round over its 98k requests in 200
    namespaces — against the oracle (which sweeps them on the host threads in well under a second)."""
    import os

    import oracle

    t = synth.make(name, 1.0)
    threads = min(len(os.sched_getaffinity(0)), 32)
    engine.upload_nodes(t.nodes, t.n_domains)
    want = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags, threads=threads)
    got = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags)
    assert_same(got[0], want[0], "lws_out")
    assert_same(got[1], want[1], "group_out")
    reqs = t.place_requests()
    if len(reqs):
        occ = R.occupancy_of(t.pod_ident, len(t.nodes))
        want_p = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, reqs, threads=threads)
        got_p, rounds = engine.place_host(reqs, occ, t.n_namespaces)
        assert_same(got_p, want_p, "place_out")
        assert rounds >= 1


def test_empty_tables(engine):
    t = synth.make("C1")
    engine.upload_nodes(t.nodes, t.n_domains)
    e_lws = R.aligned_empty(0, R.LWS_REC)
    e_grp = R.aligned_empty(0, R.GROUP_REC)
    e_pst = R.aligned_empty(0, R.POD_STATE)
    e_pid = R.aligned_empty(0, R.POD_IDENT)
    lo, go, _ = engine.sweep_lws_host(e_lws, e_grp, e_pst, e_pid)
    assert len(lo) == 0 and len(go) == 0
    # objects without any group rows
    t.lws["group_count"] = 0
    run_both(engine, synth.Tables(t.profile, t.lws, e_grp, e_pst, e_pid, t.nodes, t.n_domains, t.flags))


def test_bad_tables_are_flagged_not_fatal(engine):
    t = synth.make("fuzz", 0.05, seed=11)
    t.lws["group_base"][3] = len(t.groups) + 5
    t.groups["lws_index"][7] = len(t.lws) + 1
    t.groups["pod_base"][9] = len(t.pod_state)
    t.groups["pod_count"][9] = 3
    lo, go, _ = run_both(engine, t)
    assert lo["flags"][3] & R.LOUT_BAD_TABLE
    assert go["flags"][7] & R.GOUT_BAD_TABLE and go["flags"][9] & R.GOUT_BAD_TABLE


def test_device_pointer_entry(engine):
    """lwse_sweep_lws_device with torch-owned device memory on torch's stream."""
    import torch
    import oracle

    t = synth.make("C5", 0.02)
    engine.upload_nodes(t.nodes, t.n_domains)
    dev = torch.device("cuda:0")

    def up(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    d_lws, d_grp, d_pst, d_pid = up(t.lws), up(t.groups), up(t.pod_state), up(t.pod_ident)
    d_lo = torch.zeros(len(t.lws) * R.LWS_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_go = torch.zeros(len(t.groups) * R.GROUP_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_occ = torch.zeros(len(t.nodes), dtype=torch.int32, device=dev)
    side = torch.cuda.Stream()  # a non-default torch stream (0 / NULL would mean "the engine's own stream")
    side.wait_stream(torch.cuda.current_stream())
    stream = side.cuda_stream
    before = engine.launch_count
    engine.sweep_lws_device(d_lws, len(t.lws), d_grp, len(t.groups), d_pst, d_pid, len(t.pod_state), d_lo, d_go,
                            d_occ, flags=t.flags, stream=stream)
    torch.cuda.synchronize()
    # occupancy count + pod scan + group pass + LWS pass (C5's 8-pod groups are below the fused kernel's threshold)
    assert engine.launch_count - before == 4
    want = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags, want_occupancy=True)
    assert_same(d_lo.cpu().numpy().view(R.LWS_OUT), want[0], "lws_out")
    assert_same(d_go.cpu().numpy().view(R.GROUP_OUT), want[1], "group_out")
    assert np.array_equal(d_occ.cpu().numpy().astype(np.uint32), want[2])


def test_reference_integration_traces_on_gpu(engine):
    """The reference's own integration traces, reconciled by the CUDA engine."""
    from sim import LwsSim
    from traces import TRACES, make_lws

    def sweep(tables, flags=0):
        engine.upload_nodes(tables.nodes, tables.n_domains)
        lo, go, _ = engine.sweep_lws_host(tables.lws, tables.groups, tables.pod_state, tables.pod_ident, flags=flags)
        return lo, go

    for name, (cfg, steps) in sorted(TRACES.items()):
        sim = LwsSim(make_lws(cfg), sweep)
        sim.settle()
        sim.create_leader_pods(0, cfg["replicas"])
        for i, (action, want) in enumerate(steps):
            action(sim)
            got = sim.state() + (sim.status["condition"],)
            for g, w in zip(got, want):
                if w is not None:
                    assert g == w, f"{name} step {i}: got {got} want {want}"


def test_large_idempotent_and_order_independent(engine):
    """Full-size property checks (no oracle needed): sweeping twice gives the same
    bytes, and permuting the objects permutes the outputs."""
    t = synth.make("C3", 0.25)
    engine.upload_nodes(t.nodes, t.n_domains)
    a = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags)
    b = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    # reverse the LWS table (group rows keep their place; lws_index is remapped)
    n = len(t.lws)
    perm = np.arange(n)[::-1].copy()
    lws2 = R.aligned_empty(n, R.LWS_REC)
    lws2[:] = t.lws[perm]
    grp2 = R.aligned_empty(len(t.groups), R.GROUP_REC)
    grp2[:] = t.groups
    grp2["lws_index"] = (n - 1 - t.groups["lws_index"].astype(np.int64)).astype(np.uint32)
    c = engine.sweep_lws_host(lws2, grp2, t.pod_state, t.pod_ident, flags=t.flags)
    assert c[0].tobytes() == a[0][perm].tobytes()
    assert c[1].tobytes() == a[1].tobytes()


def test_reuse_pod_ident_flag(engine):
    """LWSE_SWEEP_REUSE_POD_IDENT keeps the identity column of the previous host sweep."""
    import oracle

    t = synth.make("fuzz", 0.3, seed=17)
    engine.upload_nodes(t.nodes, t.n_domains)
    run_both(engine, t, occupancy=False)  # uploads both pod columns
    # pod status changes (restarts appear), identities stay
    rng = np.random.default_rng(5)
    flip = rng.random(len(t.pod_state)) < 0.1
    t.pod_state[flip] |= np.uint8(R.POD_ANY_RESTART)
    garbage = R.aligned_empty(len(t.pod_ident), R.POD_IDENT)  # must NOT be read
    want = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)
    got = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, garbage, flags=t.flags | R.SWEEP_REUSE_POD_IDENT)
    assert_same(got[0], want[0], "lws_out")
    assert_same(got[1], want[1], "group_out")
    # a different pod count invalidates the resident column: the flag is ignored
    t2 = synth.make("fuzz", 0.2, seed=18)
    engine.upload_nodes(t2.nodes, t2.n_domains)
    want2 = oracle.sweep_lws(t2.lws, t2.groups, t2.pod_state, t2.pod_ident, t2.nodes, flags=t2.flags)
    got2 = engine.sweep_lws_host(t2.lws, t2.groups, t2.pod_state, t2.pod_ident, flags=t2.flags | R.SWEEP_REUSE_POD_IDENT)
    assert_same(got2[1], want2[1], "group_out")


def test_resident_tables_patches_and_change_lists(engine):
    """lwse_resident_*: tables stay on the device, the host sends row patches and gets back
    exactly the result rows that changed."""
    import oracle

    t = synth.make("fuzz", 0.5, seed=31)
    engine.upload_nodes(t.nodes, t.n_domains)
    engine.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
    before = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)

    lr, lo, gr, go, nl, ng = engine.resident_sweep(t.flags)
    assert nl == len(t.lws) and ng == len(t.groups), "the first sweep after a load reports every row"
    o = np.argsort(lr)
    assert np.array_equal(lr[o], np.arange(len(t.lws))) and lo[o].tobytes() == before[0].tobytes()
    o = np.argsort(gr)
    assert np.array_equal(gr[o], np.arange(len(t.groups))) and go[o].tobytes() == before[1].tobytes()

    _, _, _, _, nl, ng = engine.resident_sweep(t.flags)
    assert (nl, ng) == (0, 0), "nothing changed → nothing reported"

    # watch events: some pods restart / go pending, some groups lose readiness, some objects move their partition
    rng = np.random.default_rng(9)
    prow = np.unique(rng.integers(0, len(t.pod_state), size=len(t.pod_state) // 50)).astype(np.uint32)
    t.pod_state[prow] ^= np.where(rng.random(len(prow)) < 0.5, R.POD_ANY_RESTART, R.POD_PHASE_PENDING | R.POD_PHASE_RUNNING).astype(np.uint8)
    grow = np.unique(rng.integers(0, len(t.groups), size=len(t.groups) // 100)).astype(np.uint32)
    t.groups["flags"][grow] ^= R.GRP_POD_READY
    lrow = np.unique(rng.integers(0, len(t.lws), size=len(t.lws) // 100)).astype(np.uint32)
    t.lws["sts_partition"][lrow] += 1
    irow = prow[:7]
    t.pod_ident["owner_uid_hash"][irow] ^= 1
    engine.resident_patch(R.TABLE_POD_STATE, prow, t.pod_state[prow])
    engine.resident_patch(R.TABLE_GROUPS, grow, t.groups[grow])
    engine.resident_patch(R.TABLE_LWS, lrow, t.lws[lrow])
    engine.resident_patch(R.TABLE_POD_IDENT, irow, t.pod_ident[irow])
    after = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)

    lr, lo, gr, go, nl, ng = engine.resident_sweep(t.flags)
    want_l = np.flatnonzero([a.tobytes() != b.tobytes() for a, b in zip(before[0], after[0])])
    want_g = np.flatnonzero([a.tobytes() != b.tobytes() for a, b in zip(before[1], after[1])])
    assert len(want_l) > 0 and len(want_g) > 0
    o = np.argsort(lr)
    assert np.array_equal(lr[o], want_l) and lo[o].tobytes() == after[0][want_l].tobytes()
    o = np.argsort(gr)
    assert np.array_equal(gr[o], want_g) and go[o].tobytes() == after[1][want_g].tobytes()
    full = engine.resident_outputs()
    assert full[0].tobytes() == after[0].tobytes() and full[1].tobytes() == after[1].tobytes()

    # a change list that is too small: the count is still exact, the full outputs are the fallback
    t.lws["replicas"][:] += 1
    engine.resident_patch(R.TABLE_LWS, np.arange(len(t.lws), dtype=np.uint32), t.lws)
    after2 = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)
    lr, lo, gr, go, nl, ng = engine.resident_sweep(t.flags, lws_capacity=3, group_capacity=3)
    want_l2 = sum(a.tobytes() != b.tobytes() for a, b in zip(after[0], after2[0]))
    assert nl == want_l2 and len(lr) == 3
    full = engine.resident_outputs()
    assert full[0].tobytes() == after2[0].tobytes() and full[1].tobytes() == after2[1].tobytes()


def test_fused_pass_with_scattered_pod_ranges(engine):
    """Pod ranges of neighbouring groups far apart (even groups' pods first, odd groups' after):
    the fused kernel's shared-memory window does not fit and it derives the bitmap words from
    the state column directly — same results."""
    p = synth.profile("fuzz", 1.0)
    p.n_lws, p.size_choices = 6000, (32, 64)
    t = synth.make(p, seed=77)
    assert len(t.pod_state) > 4 * 65536
    order = np.concatenate([np.arange(0, len(t.groups), 2), np.arange(1, len(t.groups), 2)])
    base, count = t.groups["pod_base"].astype(np.int64), t.groups["pod_count"].astype(np.int64)
    new_base = np.zeros(len(t.groups), np.int64)
    new_base[order] = np.concatenate([[0], np.cumsum(count[order])[:-1]])
    src = np.concatenate([np.arange(base[g], base[g] + count[g]) for g in order]) if len(order) else np.zeros(0, np.int64)
    pst, pid = R.aligned_empty(len(src), R.POD_STATE), R.aligned_empty(len(src), R.POD_IDENT)
    pst[:], pid[:] = t.pod_state[src], t.pod_ident[src]
    grp = R.aligned_empty(len(t.groups), R.GROUP_REC)
    grp[:] = t.groups
    grp["pod_base"] = new_base.astype(np.uint32)
    import oracle

    engine.upload_nodes(t.nodes, t.n_domains)
    want = oracle.sweep_lws(t.lws, grp, pst, pid, t.nodes, flags=t.flags, want_occupancy=False)
    got = engine.sweep_lws_host(t.lws, grp, pst, pid, flags=t.flags, want_occupancy=False)
    assert_same(got[0], want[0], "lws_out")
    assert_same(got[1], want[1], "group_out")


@pytest.mark.parametrize("p_event", [0.002, 0.6])
def test_host_entry_reads_pinned_identity_rows_in_place(engine, p_event):
    """Pinned, mapped host tables: the identity column is not uploaded when few pods have an
    event (the group pass reads those rows over PCIe); with many events it is uploaded after the
    scan has counted them.  Same results either way."""
    import torch
    import oracle

    p = synth.profile("C3", 0.05)
    p.p_restarted = p_event
    t = synth.make(p, seed=31)
    keep = []

    def pinned(a):
        ten = torch.empty(max(a.nbytes, 16), dtype=torch.uint8).pin_memory()
        view = ten.numpy()[: a.nbytes].view(a.dtype)
        view[...] = a
        keep.append(ten)
        return view

    lws, grp, pst, pid = pinned(t.lws), pinned(t.groups), pinned(t.pod_state), pinned(t.pod_ident)
    engine.upload_nodes(t.nodes, t.n_domains)
    want = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags, want_occupancy=True)
    # (True: the occupancy count reads the whole identity column → uploaded; False: first call
    # optimistic zero-copy, second call decided by the event count the first one measured)
    for occupancy in (True, False, False):
        got = engine.sweep_lws_host(lws, grp, pst, pid, flags=t.flags, want_occupancy=occupancy)
        assert_same(got[0], want[0], "lws_out")
        assert_same(got[1], want[1], "group_out")
        if occupancy:
            assert np.array_equal(got[2], want[2])


@pytest.mark.parametrize("seed", range(100, 124))
def test_fuzz_campaign(engine, seed):
    """More seeds at a small scale: both sweep forms and, from the same tables, a placement round
    fed by the swept occupancy — every output row against the oracle."""
    import oracle
    from lws_b200 import encoder

    rng = np.random.Generator(np.random.PCG64(seed))
    p = synth.profile("fuzz", 0.15)
    p.n_nodes = int(rng.choice([8, 64, 640]))
    p.nodes_per_domain = int(rng.choice([1, 4, 16]))
    p.node_capacity = int(rng.choice([1, 4, 40]))
    p.p_exclusive = float(rng.choice([0.1, 0.5, 1.0]))
    p.p_leader_unscheduled = float(rng.choice([0.1, 0.6, 1.0]))
    t = synth.make(p, seed=seed)
    got = run_both(engine, t)
    reqs = encoder.encode_place_requests(t.lws, t.groups)
    n_ns = 3
    reqs["ns"] = (np.arange(len(reqs)) * 7 + seed) % (n_ns + 1)  # includes one out-of-range namespace
    want = oracle.place(t.nodes, got[2], t.n_domains, n_ns, reqs)
    out, _ = engine.place_host(reqs, got[2], n_ns)
    assert_same(out, want, "place_out")


def test_reference_lifecycle_entries_on_gpu(engine):
    """The scale / create / startup-policy / condition / restart-during-update entries of the reference's
    integration table (tests/test_oracle_lifecycle_traces.py), reconciled by the CUDA engine."""
    from test_oracle_lifecycle_traces import run_lifecycle_entries

    def sweep(tables, flags=0):
        engine.upload_nodes(tables.nodes, tables.n_domains)
        lo, go, _ = engine.sweep_lws_host(tables.lws, tables.groups, tables.pod_state, tables.pod_ident, flags=flags)
        return lo, go

    run_lifecycle_entries(sweep)
