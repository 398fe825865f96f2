"""GPU parity for the placement round, the DisaggregatedSet sweep and the SHA-1
group keys — CUDA engine through the C ABI vs the CPU oracle, bit-exact."""
import hashlib

import numpy as np
import pytest

from lws_b200 import encoder, synth
from lws_b200 import records as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from lws_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def same(got, want, what):
    for name in got.dtype.names or [None]:
        a, b = (got[name], want[name]) if name else (got, want)
        if not np.array_equal(a, b):
            i = int(np.flatnonzero(a != b)[0])
            raise AssertionError(f"{what}.{name}: {int((a != b).sum())} rows differ; row {i}: got {a[i]} want {b[i]}")


# ------------------------------------------------------------------ placement
def place_case(n_lws, n_nodes, size, p_excl, p_unsched, seed, fuzz=False, capacity=4, nodes_per_domain=16):
    p = synth.profile("fuzz" if fuzz else "C3", 1.0)
    p.n_lws, p.n_nodes, p.size_choices, p.replicas_choices = n_lws, n_nodes, (size,), (1, 2)
    p.p_exclusive, p.p_leader_unscheduled, p.node_capacity, p.nodes_per_domain = p_excl, p_unsched, capacity, nodes_per_domain
    t = synth.make(p, seed=seed)
    return t, encoder.encode_place_requests(t.lws, t.groups)


@pytest.mark.parametrize(
    "kw",
    [dict(n_lws=300, n_nodes=3200, size=16, p_excl=0.6, p_unsched=0.5, seed=1),
     dict(n_lws=2000, n_nodes=10000, size=64, p_excl=0.5, p_unsched=0.4, seed=2),
     dict(n_lws=500, n_nodes=640, size=8, p_excl=0.9, p_unsched=0.7, seed=3),  # heavy contention
     dict(n_lws=400, n_nodes=4096, size=4, p_excl=0.7, p_unsched=0.5, seed=4, fuzz=True),
     dict(n_lws=50, n_nodes=60000, size=8, p_excl=1.0, p_unsched=0.5, seed=5, nodes_per_domain=100),  # node words in global
     dict(n_lws=64, n_nodes=48, size=2, p_excl=1.0, p_unsched=1.0, seed=6, nodes_per_domain=2, capacity=1)],
)
def test_placement_matches_spec_oracle(engine, kw):
    import oracle

    t, reqs = place_case(**kw)
    engine.upload_nodes(t.nodes, t.n_domains)
    _, _, occ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, want_occupancy=True)
    for occupancy in (None, occ // 8):
        want = oracle.place(t.nodes, occupancy, t.n_domains, 1, reqs)
        got, rounds = engine.place_host(reqs, occupancy, 1)
        same(got, want, "place_out")
        assert rounds >= 1


@pytest.mark.parametrize(
    "kw,n_ns",
    [(dict(n_lws=300, n_nodes=3200, size=16, p_excl=0.6, p_unsched=0.5, seed=21), 1),
     (dict(n_lws=3000, n_nodes=10000, size=64, p_excl=1.0, p_unsched=0.3, seed=22), 7),
     (dict(n_lws=900, n_nodes=640, size=8, p_excl=0.9, p_unsched=0.7, seed=23), 3),  # heavy contention, many rounds
     (dict(n_lws=400, n_nodes=4096, size=4, p_excl=0.7, p_unsched=0.5, seed=24, fuzz=True), 5),
     (dict(n_lws=50, n_nodes=60000, size=8, p_excl=1.0, p_unsched=0.5, seed=25, nodes_per_domain=100), 2),  # table not staged
     (dict(n_lws=700, n_nodes=3000, size=2, p_excl=1.0, p_unsched=0.6, seed=26, nodes_per_domain=1, capacity=2), 4),  # hostname topology
     (dict(n_lws=64, n_nodes=48, size=2, p_excl=1.0, p_unsched=1.0, seed=27, nodes_per_domain=2, capacity=1), 300)],  # mostly empty namespaces
)
def test_placement_forms_agree(engine, kw, n_ns):
    """The three forms of the round — general (one warp per request over the whole grid, holders in
    L2), namespace-parallel (one CTA per namespace, holders + TMA-staged node table in shared memory)
    and its brute-force (request x node) scan — give the spec oracle's rows on tables grouped by
    namespace, including an out-of-range namespace at the end."""
    import torch

    import oracle

    t, reqs = place_case(**kw)
    rng = np.random.default_rng(kw["seed"])
    ns = np.sort(rng.integers(0, n_ns, size=len(reqs))).astype(np.uint32)
    ns[-3:] = n_ns + 2  # no such namespace
    reqs["ns"] = ns
    engine.upload_nodes(t.nodes, t.n_domains)
    occ = R.occupancy_of(t.pod_ident, len(t.nodes)) // 4
    want = oracle.place(t.nodes, occ, t.n_domains, n_ns, reqs)
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
    d_reqs, d_occ = up(reqs), up(occ)
    d_out = torch.zeros(len(reqs) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)

    def rows():
        torch.cuda.synchronize()
        return d_out.cpu().numpy().view(R.PLACE_OUT).copy()

    rounds = engine.place_device(d_reqs, len(reqs), d_occ, n_ns, d_out, want_rounds=True)
    same(rows(), want, "general")
    for flags, name in ((0, "grouped"), (R.SWEEP_PLACE_SCAN, "scan")):
        d_out.zero_()
        r2, scans = engine.place_grouped_device(d_reqs, len(reqs), d_occ, n_ns, d_out, flags=flags, want_rounds=True)
        same(rows(), want, name)
        live = int(((reqs["leader_node"] == R.NONE) & (reqs["size"] >= 1) & (reqs["ns"] < n_ns)).sum())
        assert (r2 >= 1 and scans >= live) if live else (r2 == 0 and scans == 0)
        assert rounds >= 1 or not live
    # twice in a row (the scratch is reused, counters alternate)
    d_out.zero_()
    engine.place_grouped_device(d_reqs, len(reqs), d_occ, n_ns, d_out)
    same(rows(), want, "grouped, second call")


def test_placement_namespaces_and_empty(engine):
    import oracle

    t, reqs = place_case(n_lws=400, n_nodes=2048, size=8, p_excl=0.8, p_unsched=0.6, seed=8)
    reqs["ns"] = np.arange(len(reqs)) % 3
    reqs["ns"][5] = 7  # out of range → unschedulable
    engine.upload_nodes(t.nodes, t.n_domains)
    same(engine.place_host(reqs, None, 3)[0], oracle.place(t.nodes, None, t.n_domains, 3, reqs), "place_out")
    empty = R.aligned_empty(0, R.PLACE_REQ)
    assert len(engine.place_host(empty, None, 1)[0]) == 0


def test_placement_occupancy_from_the_sweep(engine):
    """The occupancy vector the LWS sweep counts feeds the placement round."""
    import oracle

    t, reqs = place_case(n_lws=1000, n_nodes=4096, size=16, p_excl=0.5, p_unsched=0.5, seed=11, capacity=40)
    engine.upload_nodes(t.nodes, t.n_domains)
    _, _, occ_gpu = engine.sweep_lws_host(t.lws, t.groups, t.pod_state, t.pod_ident, flags=t.flags, want_occupancy=True)
    want = oracle.place(t.nodes, occ_gpu, t.n_domains, 1, reqs)
    same(engine.place_host(reqs, occ_gpu, 1)[0], want, "place_out")


# ------------------------------------------------------------------------- DS
@pytest.mark.parametrize(
    "kw",
    [dict(n_ds=5000, n_roles_choices=(2,), seed=1), dict(n_ds=3000, n_roles_choices=(2, 3, 4, 5, 6, 9, 10), seed=2),
     dict(n_ds=4000, n_roles_choices=(2, 3, 10), seed=3, fuzz=0.1), dict(n_ds=1, n_roles_choices=(2,), seed=4)],
)
def test_ds_sweep_matches_oracle(engine, kw):
    import oracle

    t = synth.make_ds(**kw)
    want = oracle.sweep_ds(t.ds, t.roles, t.revroles)
    got = engine.sweep_ds_host(t.ds, t.roles, t.revroles)
    same(got[0], want[0], "ds_out")
    same(got[1], want[1], "role_out")
    same(got[2], want[2], "revrole_out")
    # the caller's own result arrays (bench.py hands page-locked ones in)
    out = (np.full(len(t.ds), 0x5A, np.uint8).repeat(R.DS_OUT.itemsize).view(R.DS_OUT),
           np.full(len(t.roles), 0x5A, np.uint8).repeat(R.DS_ROLE_OUT.itemsize).view(R.DS_ROLE_OUT),
           np.full(len(t.revroles), 0x5A5A5A5A, R.DS_REVROLE_OUT))
    back = engine.sweep_ds_host(t.ds, t.roles, t.revroles, out=out)
    assert all(a is b for a, b in zip(back, out))
    for a, b, what in zip(out, want, ("ds_out", "role_out", "revrole_out")):
        same(a, b, what + " (caller's arrays)")
    flags = want[0]["flags"]
    if kw["n_ds"] >= 1000:  # the generator reaches every branch
        for bit in (R.DOUT_ROLLING, R.DOUT_INIT, R.DOUT_STABLE, R.DOUT_STEP, R.DOUT_NEW_READY):
            assert (flags & bit).any(), f"no DS exercises flag {bit}"


def test_ds_reference_sequences_on_gpu(engine):
    """planner_test.go's 22 exact sequences, advanced step by step by the CUDA planner."""
    import json
    import os

    from lws_b200 import api

    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "planner_sequences.json")))["cases"]
    for c in cases:
        if sum(c["source"]) == 0:
            continue  # nothing old to roll: Reconcile takes reconcileSimple, the planner is not consulted
        cfgs = [api.RollingUpdateConfiguration(maxSurge=ms, maxUnavailable=mu) for ms, mu in c["config"]]
        ds = api.DisaggregatedSet("t", roles=[api.DisaggregatedRoleSpec(f"r{i}", c["target"][i], cfgs[i]) for i in range(2)])
        old, new = list(c["source"]), [0, 0]
        seq = [[old[:], new[:]]]
        for _ in range(100):
            ch = []
            for i in range(2):
                ch.append(api.ChildLWS(f"r{i}", "old", old[i], old[i], 1.0,
                                       {api.DSInitialReplicasAnnotationKey: str(c["source"][i])}))
                ch.append(api.ChildLWS(f"r{i}", "new", new[i], new[i], 2.0))
            t = encoder.encode_ds([encoder.DsItem(ds, "new", ch)])
            ds_out, role_out, _ = engine.sweep_ds_host(t.ds, t.roles, t.revroles)
            if not ds_out[0]["flags"] & R.DOUT_STEP:
                break
            old = [int(role_out[i]["next_old"]) for i in range(2)]
            new = [int(role_out[i]["next_new"]) for i in range(2)]
            seq.append([old[:], new[:]])
        # an all-zero old revision is not "rolling" any more: the reference sequence's tail
        # (drain to zero) is the last step the rolling path can emit
        assert seq == c["steps"][: len(seq)] and len(seq) >= len(c["steps"]) - 1, c["name"]


# ---------------------------------------------------------------------- SHA-1
def test_group_keys_kats_and_hashlib(engine):
    kats = {"default/test-sample": "95e88034e460983f51a9952fe128729fbc0663b5",
            "default/podName": "390b34ab671d29e9997d7d4252b8bbf8da02f5b7",
            "leaderworkerset/test-sample": "39f5d7e9122b9d94d3932e3720b43fd3b56347e8"}
    got = engine.group_keys_host(list(kats))
    assert [d.tobytes().hex() for d in got] == list(kats.values())
    msgs = [("x" * n).encode() + bytes([n % 251]) for n in range(0, 300)] + [b"", b"a" * 5000]
    msgs += [f"ns-{i % 7}/lws-{i}-{i % 13}".encode() for i in range(20000)]
    got = engine.group_keys_host(msgs)
    for m, d in zip(msgs, got):
        assert d.tobytes() == hashlib.sha1(m).digest()


def test_placement_gathered_parts_matches_oracle(engine):
    """lwse_place_gathered_device over the payloads of three (simulated) ranks."""
    import torch
    import oracle
    from lws_b200 import distributed as D

    p = synth.profile("C3", 0.03)
    p.p_exclusive, p.p_leader_unscheduled, p.node_capacity, p.size_choices, p.n_nodes = 0.4, 0.5, 40, (8,), 3000
    t = synth.make(p, seed=21)
    world = 3
    shards = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)
    engine.upload_nodes(t.nodes, t.n_domains)
    parts, caps = [], []
    for lws, grp, pst, pid, _, _ in shards:
        _, _, occ = oracle.sweep_lws(lws, grp, pst, pid, t.nodes, want_occupancy=True)
        parts.append((occ, encoder.encode_place_requests(lws, grp)))
    cap = max(len(r) for _, r in parts)
    blob = np.concatenate([D.pack_part(o, r, cap) for o, r in parts])
    stride, off = D.part_layout(len(t.nodes), cap)
    occ_all, reqs_all = D.unpack_parts(blob, world, len(t.nodes), cap)
    want = oracle.place(t.nodes, occ_all, t.n_domains, 1, reqs_all)
    d_blob = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(world * cap * R.PLACE_OUT.itemsize, dtype=torch.uint8, device="cuda")
    rounds = engine.place_gathered_device(d_blob, world, stride, off, cap, 1, d_out, want_rounds=True)
    got = d_out.cpu().numpy().view(R.PLACE_OUT)
    same(got, want, "place_out(gathered)")
    assert rounds >= 1 and ((want["flags"] & R.PLACE_PLACED) != 0).sum() > 0


@pytest.mark.parametrize("order", [("A", "B"), ("B", "A")])
def test_ds_multi_cycle_flows_on_gpu(engine, order):
    """executor_test.go:1244-1298 — 20 reconcile cycles each, decided by the CUDA planner."""
    from test_oracle_ds_flows import run_flow

    sweep = lambda t: engine.sweep_ds_host(t.ds, t.roles, t.revroles)
    run_flow(sweep, (2, 4), (1, 2), (1, 2), (1, 1), order)
    run_flow(sweep, (4, 3), (1, 2), (3, 1), (1, 2), order, check_orphans=True)


def test_reconcile_tick_sweep_and_placement_concurrently(engine):
    """lwse_reconcile_device: sweep on the caller's stream, placement round on the engine's side
    stream, repeated ticks (the scratch halves alternate) — both results equal the oracle's."""
    import torch
    import oracle

    t, reqs = place_case(n_lws=3000, n_nodes=4096, size=16, p_excl=0.5, p_unsched=0.6, seed=21, capacity=40)
    engine.upload_nodes(t.nodes, t.n_domains)
    dev = torch.device("cuda:0")

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    want_lo, want_go, occ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags,
                                             want_occupancy=True)
    want_place = oracle.place(t.nodes, occ, t.n_domains, 1, reqs)
    d_lo = torch.zeros(len(t.lws) * R.LWS_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_go = torch.zeros(len(t.groups) * R.GROUP_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_po = torch.zeros(len(reqs) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_occ = torch.from_numpy(occ.view(np.int32)).to(dev)
    d_in = [up(t.lws), up(t.groups), up(t.pod_state), up(t.pod_ident)]  # the descriptor holds raw pointers only
    tables = engine.device_tables(d_in[0], len(t.lws), d_in[1], len(t.groups), d_in[2], d_in[3], len(t.pod_state),
                                  d_lo, d_go, None, flags=t.flags)
    d_reqs = up(reqs)
    torch.cuda.synchronize()
    for tick in range(5):
        d_lo.zero_(), d_go.zero_(), d_po.zero_()
        torch.cuda.synchronize()
        before = engine.launch_count
        engine.reconcile_device(tables, d_reqs, len(reqs), d_occ, 1, d_po)
        torch.cuda.synchronize()
        assert engine.launch_count - before == 3  # placement + fused scan/group pass + LWS pass
        same(d_lo.cpu().numpy().view(R.LWS_OUT), want_lo, f"tick{tick}.lws_out")
        same(d_go.cpu().numpy().view(R.GROUP_OUT), want_go, f"tick{tick}.group_out")
        same(d_po.cpu().numpy().view(R.PLACE_OUT), want_place, f"tick{tick}.place_out")


def test_placement_equal_domain_scores_keep_the_lower_domain(engine):
    """Two domains with the same rendezvous score (hash collisions found by inverting the mixer):
    the two-level spec takes the lower domain; interleaved node rows exercise the sorted index."""
    import oracle
    from test_place_spec import PLACE_TIES, tie_case

    for key_lo, d1, d2, score in PLACE_TIES:
        nodes, reqs = tie_case(key_lo, d1, d2)
        engine.upload_nodes(nodes, 64)
        got, _ = engine.place_host(reqs, None, 1)
        same(got, oracle.place(nodes, None, 64, 1, reqs), "place_out")
        assert got["domain_id"][0] == d1 and got["score"][0] == score


def test_peer_exchange_placement_step_on_two_gpus():
    """lwse_exchange_* / lwse_reconcile_exchanged_device: needs two GPUs of one node (skipped on a
    one-GPU box); the check itself is tests/multi_gpu/exchange_check.py under torchrun."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multi_gpu", "exchange_check.py")],
        capture_output=True, text=True, timeout=300)
    assert "EXCHANGE_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_shared_occupancy_tick_on_every_gpu_of_the_box():
    """lwse_reconcile_shared_device / LWSE_TICK_SHARED_OCCUPANCY on all GPUs of the box (2, 4 or 8;
    skipped on a one-GPU box): tests/multi_gpu/shared_occupancy_check.py under torchrun."""
    import os
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
         "127.0.0.1", "--master-port", "29534", os.path.join(root, "tests", "multi_gpu", "shared_occupancy_check.py")],
        capture_output=True, text=True, timeout=600)
    assert "SHARED_OCCUPANCY_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_reconcile_host_one_call(engine):
    """lwse_reconcile_host = lwse_sweep_lws_host + lwse_place_host in one call (placement on the
    side stream while the tables upload); pageable and pinned tables."""
    import torch
    import oracle

    t, reqs = place_case(n_lws=2500, n_nodes=4096, size=16, p_excl=0.5, p_unsched=0.6, seed=23, capacity=40)
    engine.upload_nodes(t.nodes, t.n_domains)
    want_lo, want_go, occ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags,
                                             want_occupancy=True)
    want_place = oracle.place(t.nodes, occ, t.n_domains, 1, reqs)
    keep = []

    def pinned(a):
        ten = torch.empty(max(a.nbytes, 16), dtype=torch.uint8).pin_memory()
        view = ten.numpy()[: a.nbytes].view(a.dtype)
        view[...] = a
        keep.append(ten)
        return view

    for conv in (lambda a: a, pinned):
        lo, go, po = engine.reconcile_host(conv(t.lws), conv(t.groups), conv(t.pod_state), conv(t.pod_ident),
                                           conv(reqs), occ, 1, flags=t.flags)
        same(lo, want_lo, "lws_out")
        same(go, want_go, "group_out")
        same(po, want_place, "place_out")
    # no requests: sweep only
    lo, go, po = engine.reconcile_host(t.lws, t.groups, t.pod_state, t.pod_ident, reqs[:0], None, 1, flags=t.flags)
    same(lo, want_lo, "lws_out")
    assert len(po) == 0
    # LWSE_SWEEP_PLACE_GROUPED: the caller promises the grouping, the call skips its own pass over the table …
    from lws_b200.engine import LwseError

    t3 = synth.make("C3", 0.02)
    reqs3 = t3.place_requests()
    assert bool(np.all(np.diff(reqs3["ns"].astype(np.int64)) >= 0)) and t3.n_namespaces > 1
    engine.upload_nodes(t3.nodes, t3.n_domains)
    w_lo, w_go, occ3 = oracle.sweep_lws(t3.lws, t3.groups, t3.pod_state, t3.pod_ident, t3.nodes, flags=t3.flags, want_occupancy=True)
    lo, go, po = engine.reconcile_host(t3.lws, t3.groups, t3.pod_state, t3.pod_ident, reqs3, occ3, t3.n_namespaces,
                                       flags=t3.flags | R.SWEEP_PLACE_GROUPED)
    same(lo, w_lo, "lws_out")
    same(po, oracle.place(t3.nodes, occ3, t3.n_domains, t3.n_namespaces, reqs3), "place_out (promised grouping)")
    # … and a broken promise is found on the device
    bad = reqs3[::-1].copy()
    assert not bool(np.all(np.diff(bad["ns"].astype(np.int64)) >= 0))
    with pytest.raises(LwseError):
        engine.reconcile_host(t3.lws, t3.groups, t3.pod_state, t3.pod_ident, bad, occ3, t3.n_namespaces,
                              flags=t3.flags | R.SWEEP_PLACE_GROUPED)
    # without the flag the same table takes the general kernel: fine
    lo, go, po = engine.reconcile_host(t3.lws, t3.groups, t3.pod_state, t3.pod_ident, bad, occ3, t3.n_namespaces, flags=t3.flags)
    same(po, oracle.place(t3.nodes, occ3, t3.n_domains, t3.n_namespaces, bad), "place_out (ungrouped, general kernel)")


def test_peer_exchange_degenerate_world_of_one():
    """The exchange path on one GPU (world = 1: the part is pushed to the local buffer, the flag wait
    sees its own flag): push kernel → programmatically dependent placement round → results equal the
    oracle's, over several ticks with changing parts (both buffer halves)."""
    import torch
    import oracle
    from lws_b200 import distributed as D
    from lws_b200.engine import Engine

    eng = Engine(0)  # its own engine: one exchange per engine
    try:
        t, reqs = place_case(n_lws=1500, n_nodes=2048, size=16, p_excl=0.5, p_unsched=0.6, seed=41, capacity=40)
        eng.upload_nodes(t.nodes, t.n_domains)
        cap = len(reqs) + 5
        handle = eng.exchange_create(cap, 1, 0)
        eng.exchange_connect(handle)
        stride, _ = D.part_layout(len(t.nodes), cap)
        assert eng.exchange_part_bytes == stride
        dev = torch.device("cuda:0")
        d_po = torch.zeros(cap * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
        for tick in range(4):
            rng = np.random.Generator(np.random.PCG64(tick))
            occ = rng.integers(0, 3, size=len(t.nodes)).astype(np.uint32)
            rq = reqs.copy()
            rq["leader_node"][rng.random(len(rq)) < 0.2] = R.NONE
            part = D.pack_part(occ, rq, cap)
            d_part = torch.from_numpy(part).to(dev)
            torch.cuda.synchronize()
            eng.reconcile_exchanged_device(None, d_part, 1, d_po)
            torch.cuda.synchronize()
            occ_sum, reqs_cat = D.unpack_parts(part, 1, len(t.nodes), cap)
            same(d_po.cpu().numpy().view(R.PLACE_OUT), oracle.place(t.nodes, occ_sum, t.n_domains, 1, reqs_cat),
                 f"tick{tick}.place_out")
        assert eng.exchange_status() == 0
    finally:
        eng.close()


def test_ds_order_stability_readiness_vectors_on_gpu(engine):
    """executor_test.go:486-588 and the service readiness rule, through the CUDA DS sweep
    (tests/test_oracle_ds_order_stability.py holds the vectors and checks the oracle)."""
    import test_oracle_ds_order_stability as V

    def gpu_sweep_ds(t):
        return engine.sweep_ds_host(t.ds, t.roles, t.revroles)

    for revisions, want in V.SORT_CASES:
        assert V.drain_order(gpu_sweep_ds, revisions) == want
    for case, want in V.STABLE_CASES:
        ds_out, _, rr = gpu_sweep_ds(V.stability_tables(case))
        assert bool(ds_out[0]["flags"] & R.DOUT_STABLE) == want
    for ready, want in V.READY_CASES:
        ds_out, _, _ = gpu_sweep_ds(V.readiness_tables(ready))
        assert bool(ds_out[0]["flags"] & R.DOUT_NEW_READY) == want


def test_webhook_label_batch_over_the_cuda_sha1_kernel(engine):
    """lws_b200/webhook.py with Engine.group_keys_host as its hasher: the reference's webhook label
    vectors and the three genGroupUniqueKey KATs (tests/test_webhook_batch.py holds them)."""
    import test_webhook_batch as V
    from lws_b200 import api, webhook

    V.run_cases(engine.group_keys_host)
    pods = [api.Pod(name, ns, labels={api.SetNameLabelKey: "x", api.WorkerIndexLabelKey: "0", api.GroupIndexLabelKey: "0"},
                    annotations={api.SizeAnnotationKey: "1"})
            for name, ns in (("test-sample", "default"), ("podName", "default"), ("test-sample", "leaderworkerset"))]
    assert webhook.default_labels_batch(pods, engine.group_keys_host) == [None] * 3
    assert [p.labels[api.GroupUniqueHashLabelKey] for p in pods] == [
        "95e88034e460983f51a9952fe128729fbc0663b5", "390b34ab671d29e9997d7d4252b8bbf8da02f5b7",
        "39f5d7e9122b9d94d3932e3720b43fd3b56347e8"]


def test_subgroup_index_and_keys_on_gpu(engine):
    """lwse_subgroup_keys_host: getSubGroupIndex (pod_webhook.go:249-255, Go's truncating division) and
    SHA-1("<leaderName>/<index>") formed on the device — against the host restatement + hashlib, including
    the reference's five getSubGroupIndex vectors (pod_webhook_test.go:272-305), worker 0 with subgroup size 1
    (index -1) and a zero subgroup size."""
    from lws_b200 import webhook as W

    rng = np.random.default_rng(12)
    kat = [(4, 2, 2, 1), (5, 2, 2, 0), (9, 4, 8, 1), (8, 4, 7, 1), (8, 4, 3, 0)]
    n = 5000
    pc = rng.integers(1, 70, size=n).astype(np.int32)
    sg = rng.integers(1, 9, size=n).astype(np.int32)
    wi = rng.integers(0, 70, size=n).astype(np.int32)
    for k, (a, b, c, _) in enumerate(kat):
        pc[k], sg[k], wi[k] = a, b, c
    pc[10], sg[10], wi[10] = 3, 1, 0  # (0 - 1) / 1 = -1 in Go
    sg[11] = 0
    names = [f"lws-{'x' * int(rng.integers(0, 90))}-{i}" for i in range(n)]
    index, digests = engine.subgroup_keys_host(names, pc, sg, wi)
    for k, (_, _, _, want) in enumerate(kat):
        assert index[k] == want
    assert index[10] == -1 and index[11] == np.iinfo(np.int32).min and not digests[11].any()
    for i in range(n):
        if sg[i] == 0:
            continue
        want = int(W.get_sub_group_index(int(pc[i]), int(sg[i]), int(wi[i])))
        assert index[i] == want, i
        assert bytes(digests[i]) == hashlib.sha1(f"{names[i]}/{want}".encode()).digest(), i


def test_sha1_large_batch_and_odd_layouts(engine):
    """Keys longer than one block, empty keys, a warp whose strings do not fit the staging buffer, n % 32 != 0."""
    rng = np.random.default_rng(13)
    strings = [("k%d/" % i) + "y" * int(rng.integers(0, 200)) for i in range(1000)] + ["", "a" * 55, "b" * 56, "c" * 64, "d" * 119, "e" * 3000]
    got = engine.group_keys_host(strings)
    for s, d in zip(strings, got):
        assert bytes(d) == hashlib.sha1(s.encode()).digest()


def test_constraint_cases_on_gpu(engine):
    """tests/test_place_constraints.py's table (namespace scoping, Exists ∧ NotIn, pinned leaders, capacity
    filter) and the subgroup-exclusive class through the CUDA kernels: every row equals the spec oracle's."""
    import oracle
    from test_place_constraints import CASES, check, nodes, reqs

    for name, n_dom, n_ns, rows, want in CASES:
        n, rq = nodes(n_dom), reqs(rows)
        engine.upload_nodes(n, n_dom)
        got, _ = engine.place_host(rq, None, n_ns)
        same(got, oracle.place(n, None, n_dom, n_ns, rq), name)
        check(got, want)
    p = synth.profile("fuzz", 0.05)
    p.size_choices, p.replicas_choices, p.p_exclusive, p.fuzz, p.p_short_group, p.n_nodes, p.nodes_per_domain = (5,), (2,), 1.0, 0.0, 0.0, 400, 4
    p.node_capacity, p.p_leader_unscheduled = 40, 0.5
    t = synth.make(p, seed=3)
    t.lws["subgroup_size"] = 2
    sub = encoder.encode_subgroup_place_requests(t.lws, t.groups, t.pod_ident, np.ones(len(t.lws), bool), t.ns_of_lws, t.n_namespaces)
    g = t.place_requests()
    both = R.aligned_empty(len(g) + len(sub), R.PLACE_REQ)
    both[: len(g)], both[len(g):] = g, sub
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    engine.upload_nodes(t.nodes, t.n_domains)
    got, _ = engine.place_host(both, occ, 2 * t.n_namespaces)
    same(got, oracle.place(t.nodes, occ, t.n_domains, 2 * t.n_namespaces, both), "group + subgroup classes")
