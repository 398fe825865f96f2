"""Run under torchrun on N >= 2 GPUs of one node:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu/exchange_check.py
Every rank shards one cluster by UID hash, sweeps its shard, pushes its [occupancy | requests]
part to the peers over NVLink (lwse_reconcile_exchanged_device) and solves the placement round
over all parts; the result must equal (a) the oracle's on the unpacked parts and (b) what the
NCCL all-gather form (lwse_place_gathered_device) computes.  Several ticks with changing
requests exercise the two buffer halves and the step flags."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lws_b200 import distributed as D  # noqa: E402
from lws_b200 import encoder, synth  # noqa: E402
from lws_b200 import records as R  # noqa: E402
from lws_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import oracle

    p = synth.profile("C3", 0.02)
    p.p_exclusive, p.p_leader_unscheduled, p.n_nodes = 0.3, 0.5, 2048
    t = synth.make(p, seed=5)
    shards = D.shard_lws_tables(t.lws, t.groups, t.pod_state, t.pod_ident, world)
    lws, grp, pst, pid, _, _ = shards[rank]
    eng = Engine(local)
    eng.upload_nodes(t.nodes, t.n_domains)
    n_nodes = len(t.nodes)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)

    reqs_all = [encoder.encode_place_requests(s[0], s[1]) for s in shards]
    cap = max(len(r) for r in reqs_all) + 3
    D.connect_exchange(eng, cap, world, rank, device=dev)
    stride, off = D.part_layout(n_nodes, cap)
    assert eng.exchange_part_bytes == stride
    d_in = [up(lws), up(grp), up(pst), up(pid)]
    d_lo = torch.zeros(len(lws) * R.LWS_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_go = torch.zeros(len(grp) * R.GROUP_OUT.itemsize, dtype=torch.uint8, device=dev)
    tables = eng.device_tables(d_in[0], len(lws), d_in[1], len(grp), d_in[2], d_in[3], len(pst), d_lo, d_go, None,
                               flags=t.flags)
    want_lo, want_go, _ = oracle.sweep_lws(lws, grp, pst, pid, t.nodes, flags=t.flags, want_occupancy=False)
    d_po = torch.zeros(world * cap * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
    d_po2 = torch.zeros_like(d_po)
    ok = True
    for tick in range(6):
        # per-rank occupancy and requests change every tick (deterministic per (rank, tick) on every rank)
        parts = []
        for r in range(world):
            rng = np.random.Generator(np.random.PCG64(1000 * tick + r))
            occ = rng.integers(0, 3, size=n_nodes).astype(np.uint32)
            rq = reqs_all[r].copy()
            rq["priority"] ^= np.uint64(tick) << np.uint64(40)
            drop = rng.random(len(rq)) < 0.2
            rq["leader_node"][drop] = R.NONE
            parts.append(D.pack_part(occ, rq, cap))
        d_part = up(parts[rank])
        torch.cuda.synchronize()  # the engine's streams are non-blocking: no implicit order with torch's copies
        eng.reconcile_exchanged_device(tables, d_part, 1, d_po)
        torch.cuda.synchronize()
        occ_sum, reqs_cat = D.unpack_parts(np.concatenate(parts), world, n_nodes, cap)
        want = oracle.place(t.nodes, occ_sum, t.n_domains, 1, reqs_cat)
        got = d_po.cpu().numpy().view(R.PLACE_OUT)
        # the NCCL form on the same parts
        gathered = torch.empty(world * stride, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, d_part)
        torch.cuda.synchronize()
        eng.place_gathered_device(gathered, world, stride, off, cap, 1, d_po2)
        torch.cuda.synchronize()
        got2 = d_po2.cpu().numpy().view(R.PLACE_OUT)
        same = got.tobytes() == want.tobytes() and got2.tobytes() == want.tobytes()
        same = same and d_lo.cpu().numpy().tobytes() == want_lo.tobytes() and d_go.cpu().numpy().tobytes() == want_go.tobytes()
        if not same:
            what = [n for n, a, b in (("peer", got, want), ("nccl", got2, want)) if a.tobytes() != b.tobytes()]
            print(f"rank {rank} tick {tick}: MISMATCH in {what} (+ sweep outputs if empty)", flush=True)
            ok = False
    err = eng.exchange_status()
    flag = torch.tensor([1 if (ok and err == 0) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0:
        print("EXCHANGE_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", "world", world, flush=True)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
