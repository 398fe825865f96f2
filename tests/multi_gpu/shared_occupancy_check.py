"""Run under torchrun on N >= 2 GPUs of one node (any N up to 8):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29534 tests/multi_gpu/shared_occupancy_check.py
ONE cluster, sharded two ways at once: the sweep by LWS UID hash, the placement by namespace
(lws_b200.distributed.requests_of_rank).  Per tick every rank sweeps its shard and pushes its
occupancy counters to the peers over NVLink (lwse_reconcile_shared_device: peer stores + flags, no
collective library); its local requests are solved against the SUM.  Checked on every rank, every
tick: sweep outputs == oracle on the shard, placement rows == spec oracle on (local requests, summed
occupancy of the right step) — in step and lagged — and the union over ranks == the oracle on the
whole, unsharded request table.  Also the resident tick with LWSE_TICK_SHARED_OCCUPANCY."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lws_b200 import distributed as D  # noqa: E402
from lws_b200 import synth  # noqa: E402
from lws_b200 import records as R  # noqa: E402
from lws_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import oracle

    p = synth.profile("C3", 0.05)
    p.n_namespaces, p.n_nodes, p.p_leader_unscheduled = 24, 2048, 0.3
    t = synth.make(p, seed=9)
    shards = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)
    lws, grp, pst, pid, _, _ = shards[rank]
    reqs_global = t.place_requests()
    reqs, n_ns, rows = D.requests_of_rank(reqs_global, world, rank)
    n_nodes = len(t.nodes)
    eng = Engine(local)
    eng.upload_nodes(t.nodes, t.n_domains)
    D.connect_exchange(eng, 0, world, rank, device=dev)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    d_in = [up(lws), up(grp), up(pst), up(pid)]
    d_lo = torch.zeros(len(lws) * R.LWS_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_go = torch.zeros(len(grp) * R.GROUP_OUT.itemsize, dtype=torch.uint8, device=dev)
    tables = eng.device_tables(d_in[0], len(lws), d_in[1], len(grp), d_in[2], d_in[3], len(pst), d_lo, d_go, None, flags=t.flags)
    want_lo, want_go, _ = oracle.sweep_lws(lws, grp, pst, pid, t.nodes, flags=t.flags)
    d_reqs = up(reqs)
    d_po = torch.zeros(max(len(reqs), 1) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
    occ_of = lambda r, tick: (R.occupancy_of(shards[r][3], n_nodes) + np.uint32(tick * (r + 1) % 3)).astype(np.uint32)
    ok = True
    prev_sum = None
    for tick in range(1, 8):
        lagged = tick >= 4
        occ_local = occ_of(rank, tick)
        occ_sum = sum(occ_of(r, tick).astype(np.uint64) for r in range(world)).astype(np.uint32)
        d_occ = up(D.padded_occupancy(occ_local))
        torch.cuda.synchronize()
        eng.reconcile_shared_device(tables, d_reqs, len(reqs), d_occ, n_ns, d_po, flags=R.EXCHANGE_LAGGED if lagged else 0)
        torch.cuda.synchronize()
        seen = prev_sum if lagged else occ_sum  # lagged: the snapshot every rank pushed one call earlier
        want = oracle.place(t.nodes, seen, t.n_domains, n_ns, reqs)
        got = d_po.cpu().numpy()[: len(reqs) * R.PLACE_OUT.itemsize].view(R.PLACE_OUT)
        same = got.tobytes() == want.tobytes()
        same = same and d_lo.cpu().numpy().tobytes() == want_lo.tobytes() and d_go.cpu().numpy().tobytes() == want_go.tobytes()
        # the union over ranks is the unsharded problem's answer
        whole = oracle.place(t.nodes, seen, t.n_domains, t.n_namespaces, reqs_global)
        same = same and got.tobytes() == whole[rows].tobytes()
        if not same:
            print(f"rank {rank} tick {tick} lagged={lagged}: MISMATCH", flush=True)
            ok = False
        prev_sum = occ_sum
        dist.barrier()
    # the resident tick on the same shard
    eng2 = Engine(local)
    ok2 = True
    try:
        eng2.upload_nodes(t.nodes, t.n_domains)
        D.connect_exchange(eng2, 0, world, rank, device=dev)
        eng2.resident_load(lws, grp, pst, pid)
        eng2.resident_place_load(reqs, n_ns)
        flags = t.flags | R.TICK_PLACE | R.TICK_SHARED_OCCUPANCY
        for k in range(3):
            eng2.resident_tick(eng2.make_tick((), flags))
            dist.barrier()
        occ_sum = sum(R.occupancy_of(shards[r][3], n_nodes).astype(np.uint64) for r in range(world)).astype(np.uint32)
        want = oracle.place(t.nodes, occ_sum, t.n_domains, n_ns, reqs)
        ok2 = eng2.resident_place_outputs().tobytes() == want.tobytes()
        g_lo, g_go = eng2.resident_outputs()
        ok2 = ok2 and g_lo.tobytes() == want_lo.tobytes() and g_go.tobytes() == want_go.tobytes()
        if not ok2:
            print(f"rank {rank}: resident tick with shared occupancy MISMATCH", flush=True)
    finally:
        err2 = eng2.exchange_status()
    err = eng.exchange_status()
    flag = torch.tensor([1 if (ok and ok2 and err == 0 and err2 == 0) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0:
        print("SHARED_OCCUPANCY_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", "world", world, flush=True)
    eng.close()
    eng2.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
