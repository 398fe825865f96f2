"""Host-side plumbing for the multi-GPU layout: one process per GPU, objects
sharded by LWS UID hash, no data-path collective for the sweep, and ONE
all-gather per step for the placement round (per-node occupancy + the shard's
placement requests), after which every rank solves the same small placement
problem (``lwse_place_gathered_device``).

Everything here is numpy / byte layout; the collective itself is issued by the
caller with ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import records as R
from .engine import shard_of


def part_layout(n_nodes: int, reqs_per_part: int) -> tuple[int, int]:
    """→ (part_stride_bytes, reqs_offset_bytes) of one rank's all-gather payload."""
    reqs_offset = (n_nodes * 4 + 15) // 16 * 16
    stride = reqs_offset + reqs_per_part * R.PLACE_REQ.itemsize
    return stride, reqs_offset


def pad_requests(reqs: np.ndarray, cap: int) -> np.ndarray:
    """Pad to `cap` rows with inert requests (unpinned, size 0 → never claim anything)."""
    out = R.aligned_empty(cap, R.PLACE_REQ)
    out["leader_node"] = R.NONE
    out["size"] = 0
    out[: len(reqs)] = reqs
    return out


def pack_part(occupancy: np.ndarray, reqs: np.ndarray, cap: int) -> np.ndarray:
    """This rank's payload: [occupancy u32 x n_nodes | pad | cap request rows] as bytes."""
    n_nodes = len(occupancy)
    stride, off = part_layout(n_nodes, cap)
    buf = np.zeros(stride, dtype=np.uint8)
    buf[: n_nodes * 4] = np.ascontiguousarray(occupancy, dtype=np.uint32).view(np.uint8)
    buf[off:] = pad_requests(reqs, cap).view(np.uint8)
    return buf


def unpack_parts(gathered: np.ndarray, world: int, n_nodes: int, cap: int):
    """→ (summed occupancy, all request rows in gathered order) — what the kernel sees."""
    stride, off = part_layout(n_nodes, cap)
    g = np.ascontiguousarray(gathered, dtype=np.uint8).reshape(world, stride)
    occ = g[:, : n_nodes * 4].copy().view(np.uint32).reshape(world, n_nodes).sum(axis=0, dtype=np.uint64).astype(np.uint32)
    reqs = R.aligned_empty(world * cap, R.PLACE_REQ)
    reqs.view(np.uint8)[:] = g[:, off:].reshape(-1)
    return occ, reqs


def connect_exchange(engine, reqs_per_part: int, world: int, rank: int, device=None) -> None:
    """Set up the engine's peer exchange (``lwse_exchange_*``): create the local buffer, all-gather the
    64-byte IPC handles with ``torch.distributed`` (setup only — the data path uses peer stores, not a
    collective), map every peer's buffer.  Call on every rank after ``upload_nodes``."""
    import torch
    import torch.distributed as dist

    handle = engine.exchange_create(reqs_per_part, world, rank)
    mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    engine.exchange_connect(b"".join(bytes(t.cpu().numpy().tobytes()) for t in every))
    dist.barrier()  # nobody pushes before every buffer is mapped everywhere


def shard_index(uid_hash: np.ndarray, world: int) -> np.ndarray:
    """lwse_shard_of for a whole column."""
    if world <= 1:
        return np.zeros(len(uid_hash), dtype=np.uint32)
    x = uid_hash.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x % np.uint64(world)).astype(np.uint32)


def shard_lws_tables(lws, groups, pod_state, pod_ident, world: int):
    """Split one cluster's tables into `world` self-contained shards by LWS UID hash.
    Group and pod rows travel with their object; bases are re-indexed.  Returns, per shard,
    (lws, groups, pod_state, pod_ident, lws_rows, group_rows) where *_rows map shard rows back
    to the rows of the unsharded tables."""
    shard = shard_index(lws["uid_hash"], world)
    g_owner = groups["lws_index"].astype(np.int64)
    out = []
    for rank in range(world):
        lrows = np.flatnonzero(shard == rank)
        new_lws_index = np.full(len(lws), -1, dtype=np.int64)
        new_lws_index[lrows] = np.arange(len(lrows))
        grows = np.flatnonzero(shard[g_owner] == rank) if len(groups) else np.zeros(0, np.int64)
        s_lws = R.aligned_empty(len(lrows), R.LWS_REC)
        s_lws[:] = lws[lrows]
        s_grp = R.aligned_empty(len(grows), R.GROUP_REC)
        s_grp[:] = groups[grows]
        s_grp["lws_index"] = new_lws_index[g_owner[grows]].astype(np.uint32)
        # group rows of an object stay contiguous and in order: new base = rank of its first row
        gcount = s_lws["group_count"].astype(np.int64)
        s_lws["group_base"] = (np.concatenate([[0], np.cumsum(gcount)[:-1]]) if len(lrows) else np.zeros(0)).astype(np.uint32)
        pcount = s_grp["pod_count"].astype(np.int64)
        new_pod_base = np.concatenate([[0], np.cumsum(pcount)[:-1]]).astype(np.int64) if len(grows) else np.zeros(0, np.int64)
        idx = np.repeat(s_grp["pod_base"].astype(np.int64) - new_pod_base, pcount) + np.arange(int(pcount.sum()))
        s_pst = R.aligned_empty(len(idx), R.POD_STATE)
        s_pst[:] = pod_state[idx]
        s_pid = R.aligned_empty(len(idx), R.POD_IDENT)
        s_pid[:] = pod_ident[idx]
        s_grp["pod_base"] = new_pod_base.astype(np.uint32)
        out.append((s_lws, s_grp, s_pst, s_pid, lrows, grows))
    return out


__all__ = ["connect_exchange", "part_layout", "pad_requests", "pack_part", "unpack_parts", "shard_index", "shard_lws_tables", "shard_of"]


# --------------------------------------------------------------------------- #
# shared occupancy: ranks own namespaces, nodes are common
# --------------------------------------------------------------------------- #
def namespace_owner(ns: np.ndarray, world: int) -> np.ndarray:
    """Which rank solves the placement rounds of a namespace (exclusivity is per namespace, so whole
    namespaces — never single requests — are the unit of placement work)."""
    return (np.asarray(ns, dtype=np.uint64) % np.uint64(max(world, 1))).astype(np.uint32)


def requests_of_rank(reqs: np.ndarray, world: int, rank: int):
    """The rows of a (namespace-grouped) global request table that `rank` solves, with dense local
    namespace ids → (requests, n_namespaces, global row numbers).  Strong scaling: the sweep shards by
    LWS UID hash, the placement by namespace; the only per-tick exchange is the occupancy vector."""
    if world <= 1:
        n_ns = int(reqs["ns"].max()) + 1 if len(reqs) else 1
        return reqs, n_ns, np.arange(len(reqs))
    rows = np.flatnonzero(namespace_owner(reqs["ns"], world) == rank)
    mine = R.aligned_empty(len(rows), R.PLACE_REQ)
    mine[:] = reqs[rows]
    mine["ns"] = mine["ns"] // np.uint32(world)  # dense and still non-decreasing
    n_ns = int(mine["ns"].max()) + 1 if len(mine) else 1
    return mine, n_ns, rows


def padded_occupancy(occupancy: np.ndarray) -> np.ndarray:
    """The rank's occupancy counters padded to a multiple of 16 bytes (what lwse_reconcile_shared_device pushes)."""
    n = (len(occupancy) * 4 + 15) // 16 * 4
    out = np.zeros(n, dtype=np.uint32)
    out[: len(occupancy)] = occupancy
    return out
