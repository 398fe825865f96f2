"""The N>1 path on CPU: world_size-2 gloo processes exercise the host-side logic —
sharding by UID hash, the single all-gather payload of a placement step, and the
invariant that sharding changes no result.  The sweep/placement backend here is the
CPU oracle (there is no GPU in this container); the GPU run uses the same plumbing
with NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import oracle
    from lws_b200 import distributed as D
    from lws_b200 import encoder, synth
    from lws_b200 import records as R

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank builds the same global cluster, keeps its own shard
        p = synth.profile("C3", 0.02)
        p.p_exclusive, p.p_leader_unscheduled, p.node_capacity, p.size_choices, p.n_nodes = 0.3, 0.5, 40, (8,), 2000
        t = synth.make(p, seed=77)
        shards = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)
        lws, grp, pst, pid, lrows, grows = shards[rank]
        # sweep of the shard (no collective)
        lo, go, occ = oracle.sweep_lws(lws, grp, pst, pid, t.nodes, flags=t.flags, want_occupancy=True)
        # placement: ONE all-gather of [occupancy | requests]
        reqs = encoder.encode_place_requests(lws, grp)
        cap_t = torch.tensor([len(reqs)])
        dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
        cap = int(cap_t.item())
        part = torch.from_numpy(D.pack_part(occ, reqs, cap))
        gathered = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        occ_all, reqs_all = D.unpack_parts(torch.cat(gathered).numpy(), world, len(t.nodes), cap)
        pout = oracle.place(t.nodes, occ_all, t.n_domains, 1, reqs_all)
        mine = pout[rank * cap: rank * cap + len(reqs)]
        q.put((rank, lrows, grows, lo.tobytes(), go.tobytes(), occ, reqs_all.tobytes(), pout.tobytes(),
               mine.tobytes(), reqs["group"].copy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_step_matches_unsharded():
    import torch.multiprocessing as mp

    import oracle
    from lws_b200 import synth
    from lws_b200 import records as R

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0

    # reference: the same cluster, unsharded
    p = synth.profile("C3", 0.02)
    p.p_exclusive, p.p_leader_unscheduled, p.node_capacity, p.size_choices, p.n_nodes = 0.3, 0.5, 40, (8,), 2000
    t = synth.make(p, seed=77)
    lo, go, occ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags, want_occupancy=True)
    seen_l, seen_g = np.zeros(len(t.lws), bool), np.zeros(len(t.groups), bool)
    occ_sum = np.zeros(len(t.nodes), np.uint64)
    for rank, lrows, grows, lo_b, go_b, occ_r, reqs_all_b, pout_b, mine_b, req_groups in res:
        s_lo = np.frombuffer(lo_b, dtype=R.LWS_OUT)
        s_go = np.frombuffer(go_b, dtype=R.GROUP_OUT)
        assert s_lo.tobytes() == lo[lrows].tobytes(), "sharding changed an LWS result"
        assert s_go.tobytes() == go[grows].tobytes(), "sharding changed a group result"
        seen_l[lrows] = True
        seen_g[grows] = True
        occ_sum += occ_r
    assert seen_l.all() and seen_g.all(), "a row fell between the shards"
    assert np.array_equal(occ_sum.astype(np.uint32), occ), "summed shard occupancy != global occupancy"
    # every rank solved the identical placement problem and got the identical answer
    assert res[0][6] == res[1][6] and res[0][7] == res[1][7]
    pout = np.frombuffer(res[0][7], dtype=R.PLACE_OUT)
    placed = (pout["flags"] & R.PLACE_PLACED) != 0
    doms = pout["domain_id"][placed]
    assert len(doms) == len(set(doms.tolist())), "two groups hold one domain across shards"
    assert placed.sum() > 0


def test_shard_of_matches_c_abi_and_balances():
    from lws_b200 import distributed as D
    from lws_b200 import engine

    rng = np.random.default_rng(1)
    uid = rng.integers(0, 1 << 63, size=4000, dtype=np.uint64)
    for world in (1, 2, 4, 8):
        got = D.shard_index(uid, world)
        want = np.array([engine.shard_of(int(u), world) for u in uid[:500]], dtype=np.uint32)
        assert np.array_equal(got[:500], want)
        counts = np.bincount(got, minlength=world)
        assert counts.min() > 0.8 * len(uid) / world


def test_pack_unpack_roundtrip():
    from lws_b200 import distributed as D
    from lws_b200 import records as R

    rng = np.random.default_rng(2)
    world, n_nodes, cap = 3, 37, 5
    parts, occs, allreq = [], [], []
    for r in range(world):
        occ = rng.integers(0, 9, size=n_nodes).astype(np.uint32)
        reqs = R.aligned_empty(r + 2, R.PLACE_REQ)
        reqs["priority"] = rng.integers(0, 1 << 62, size=len(reqs), dtype=np.uint64)
        reqs["size"] = 4
        reqs["leader_node"] = R.NONE
        parts.append(D.pack_part(occ, reqs, cap))
        occs.append(occ)
        allreq.append(D.pad_requests(reqs, cap))
    occ_sum, reqs_all = D.unpack_parts(np.concatenate(parts), world, n_nodes, cap)
    assert np.array_equal(occ_sum, np.sum(occs, axis=0).astype(np.uint32))
    assert reqs_all.tobytes() == np.concatenate(allreq).tobytes()
    stride, off = D.part_layout(n_nodes, cap)
    assert stride % 16 == 0 and off % 16 == 0
    # padding rows are inert: unpinned and size 0
    pad = reqs_all[cap - 1]
    assert pad["leader_node"] == R.NONE and pad["size"] == 0


def _shared_worker(rank, world, port, q):
    """The shared-occupancy layout on CPU (gloo): sweep sharded by UID hash, placement by namespace owner,
    ONE all-gather of the occupancy vector — what lwse_reconcile_shared_device does with peer stores."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import oracle
    from lws_b200 import distributed as D
    from lws_b200 import synth
    from lws_b200 import records as R

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = synth.profile("C3", 0.03)
        p.n_namespaces, p.n_nodes, p.p_leader_unscheduled = 10, 1024, 0.4
        t = synth.make(p, seed=78)
        shards = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)
        lws, grp, pst, pid, _, _ = shards[rank]
        _, _, occ = oracle.sweep_lws(lws, grp, pst, pid, t.nodes, flags=t.flags, want_occupancy=True)
        mine = torch.from_numpy(D.padded_occupancy(occ).view(np.int32).copy())
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)  # the single collective of the step
        occ_sum = sum(x.numpy().view(np.uint32)[: len(t.nodes)].astype(np.uint64) for x in parts).astype(np.uint32)
        reqs_global = t.place_requests()
        reqs, n_ns, rows = D.requests_of_rank(reqs_global, world, rank)
        assert np.all(np.diff(reqs["ns"].astype(np.int64)) >= 0)  # still grouped by namespace
        out = oracle.place(t.nodes, occ_sum, t.n_domains, n_ns, reqs)
        q.put((rank, rows, out.tobytes(), occ_sum.tobytes()))
    finally:
        dist.destroy_process_group()


def test_shared_occupancy_layout_gloo_world2():
    import torch.multiprocessing as mp

    import oracle
    from lws_b200 import records as R
    from lws_b200 import synth

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = synth.profile("C3", 0.03)
    p.n_namespaces, p.n_nodes, p.p_leader_unscheduled = 10, 1024, 0.4
    t = synth.make(p, seed=78)
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    assert got[0][3] == occ.tobytes() == got[1][3]  # summed shard occupancies = the cluster's
    reqs = t.place_requests()
    whole = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, reqs)
    seen = np.zeros(len(reqs), bool)
    for rank, rows, out, _ in got:
        assert np.frombuffer(out, R.PLACE_OUT).tobytes() == whole[rows].tobytes()  # every rank's rows = the unsharded answer
        seen[rows] = True
    assert seen.all()
