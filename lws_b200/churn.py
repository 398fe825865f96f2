"""Synthetic watch-event churn for the resident tick (``lwse_resident_tick``).

A controller sees events, not tables.  A *churn plan* is a list of patch sets; one set is what
one reconcile pass of the work queue has to digest:

* pod status updates — ``frac_pods`` of the pod state bytes get a new value (phase flips
  Pending <-> Running, restart counts move, a few deletions start);
* scheduling events — ``frac_reqs`` of the placement requests change sides: unpinned leaders
  get bound to the node the previous placement round chose for them, as many bound ones go
  back to unscheduled (pod recreated); the group rows carry the same ``leader_node``.

Set 2j+1 undoes the scheduling events of set 2j, so a long run stays at the profile's 5 %
unscheduled leaders.  Both bench arms consume the same plan: the GPU engine as row patches,
the CPU arm as (apply patches, reconcile the dirty objects) — ``dirty_*`` are the row numbers
the event source hands a controller for free.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import records as R


@dataclass
class PatchSet:
    pod_rows: np.ndarray
    pod_vals: np.ndarray
    req_rows: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    req_vals: np.ndarray = field(default_factory=lambda: R.aligned_empty(0, R.PLACE_REQ))
    grp_rows: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    grp_vals: np.ndarray = field(default_factory=lambda: R.aligned_empty(0, R.GROUP_REC))
    dirty_groups: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    dirty_lws: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))

    def h2d_bytes(self) -> int:
        return int(self.pod_rows.nbytes + self.pod_vals.nbytes + self.req_rows.nbytes + self.req_vals.nbytes
                   + self.grp_rows.nbytes + self.grp_vals.nbytes)


def _toggle_states(vals: np.ndarray, rng) -> np.ndarray:
    v = vals.astype(np.uint8).copy()
    r = rng.random(len(v))
    phase = v & R.POD_PHASE_MASK
    flip = (r < 0.60) & ((phase == R.POD_PHASE_PENDING) | (phase == R.POD_PHASE_RUNNING))
    v[flip] ^= np.uint8(3)  # Pending <-> Running
    v[(r >= 0.60) & (r < 0.80)] ^= np.uint8(R.POD_ANY_RESTART)
    v[(r >= 0.80) & (r < 0.85)] ^= np.uint8(R.POD_DELETING)
    return v


def make_plan(t, reqs: np.ndarray | None, place_out: np.ndarray | None, frac_pods: float, frac_reqs: float = 0.0,
              n_sets: int = 4, seed: int = 7, patch_groups: bool = True) -> list[PatchSet]:
    """Patch sets for tables ``t`` (a ``synth.Tables``).  Applied in order, cyclically, they keep
    mirror tables on the host consistent: every set's values are computed from the state the
    previous sets leave behind (the plan simulates the run once).
    ``patch_groups=False``: the request table is sharded differently from the group table (strong
    scaling: placement by namespace owner, sweep by UID hash) — a scheduling event then patches the
    request row here and the group row on the rank that sweeps the group (not generated)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_pods = len(t.pod_state)
    state = t.pod_state.copy()
    groups = t.groups.copy()
    cur_reqs = None if reqs is None else reqs.copy()
    pod_group = np.repeat(np.arange(len(groups), dtype=np.uint32), groups["pod_count"].astype(np.int64))
    # only valid when the pod table is laid out group after group (synth tables are)
    laid_out = len(pod_group) == n_pods and bool(
        np.all(groups["pod_base"].astype(np.int64) == np.concatenate([[0], np.cumsum(groups["pod_count"].astype(np.int64))[:-1]])))
    sets: list[PatchSet] = []
    n_p = int(round(n_pods * frac_pods))
    swap = None
    for k in range(n_sets):
        rows = np.sort(rng.choice(n_pods, size=n_p, replace=False)).astype(np.uint32) if 0 < n_p < n_pods else (
            np.arange(n_pods, dtype=np.uint32) if n_p >= n_pods else np.zeros(0, np.uint32))
        vals = R.aligned_empty(len(rows), R.POD_STATE)
        vals[:] = _toggle_states(state[rows], rng)
        state[rows] = vals
        ps = PatchSet(pod_rows=rows, pod_vals=vals)
        dg = pod_group[rows] if laid_out else np.zeros(0, np.uint32)
        dl = np.zeros(0, np.uint32)
        if laid_out and len(rows):
            leader = rows.astype(np.int64) == groups["pod_base"][dg].astype(np.int64)
            dl = groups["lws_index"][dg[leader]]
        if cur_reqs is not None and frac_reqs > 0 and len(cur_reqs):
            if k % 2 == 0:
                n_s = max(1, int(round(len(cur_reqs) * frac_reqs / 2)))
                unp = np.flatnonzero((cur_reqs["leader_node"] == R.NONE) & (cur_reqs["size"] >= 1))
                if place_out is not None:  # only those the round placed: bind them where it put them
                    unp = unp[(place_out["flags"][unp] & R.PLACE_PLACED) != 0]
                pin = np.flatnonzero(cur_reqs["leader_node"] < R.NODE_NOT_FOUND)
                a = rng.choice(unp, size=min(n_s, len(unp)), replace=False) if len(unp) else np.zeros(0, np.int64)
                b = rng.choice(pin, size=min(n_s, len(pin)), replace=False) if len(pin) else np.zeros(0, np.int64)
                new_nodes = np.concatenate([place_out["leader_node"][a] if place_out is not None
                                            else np.zeros(len(a), np.uint32), np.full(len(b), R.NONE, np.uint32)])
                rr = np.concatenate([a, b]).astype(np.uint32)
                old_nodes = cur_reqs["leader_node"][rr].copy()
                swap = (rr, old_nodes)
            else:
                rr, new_nodes = swap  # undo the previous set's scheduling events
            order = np.argsort(rr, kind="stable")
            rr, new_nodes = rr[order], new_nodes[order]
            cur_reqs["leader_node"][rr] = new_nodes
            rv = R.aligned_empty(len(rr), R.PLACE_REQ)
            rv[:] = cur_reqs[rr]
            ps.req_rows, ps.req_vals = rr, rv
            if patch_groups:
                gr = cur_reqs["group"][rr].astype(np.uint32)
                groups["leader_node"][gr] = new_nodes
                gv = R.aligned_empty(len(gr), R.GROUP_REC)
                gv[:] = groups[gr]
                ps.grp_rows, ps.grp_vals = gr, gv
                dg = np.concatenate([dg, gr])
                dl = np.concatenate([dl, groups["lws_index"][gr]])
        ps.dirty_groups = np.unique(dg).astype(np.uint32)
        ps.dirty_lws = np.unique(dl).astype(np.uint32)
        sets.append(ps)
    return sets


def apply_to_mirror(ps: PatchSet, pod_state: np.ndarray, groups: np.ndarray, reqs: np.ndarray | None) -> None:
    """The host-side mirror of one tick's patches (plain numpy; the CPU arm uses the oracle's C loop)."""
    pod_state[ps.pod_rows] = ps.pod_vals
    if len(ps.grp_rows):
        groups[ps.grp_rows] = ps.grp_vals
    if reqs is not None and len(ps.req_rows):
        reqs[ps.req_rows] = ps.req_vals


class ArenaPlan:
    """A churn plan laid out in an engine's patch arena: the rows / values the tick descriptors
    point at are views of pinned, mapped memory the GPU reads in place."""

    def __init__(self, engine, sets: list[PatchSet], flags: int, range_threshold: float = 0.2):
        need = 4096
        for ps in sets:
            need += sum(((a.nbytes + 255) // 256 + 1) * 256 for a in
                        (ps.pod_rows, ps.pod_vals, ps.req_rows, ps.req_vals, ps.grp_rows, ps.grp_vals))
        self.arena = engine.resident_arena(need)
        self._cursor = 0
        self.ticks = []
        self.h2d_bytes = []
        n_pods = engine._resident_pods
        for ps in sets:
            segs = []
            if len(ps.pod_rows) >= range_threshold * n_pods and len(ps.pod_rows) == n_pods:
                # the whole column changed: one contiguous DMA copy instead of a scatter
                segs.append((R.TABLE_POD_STATE, 0, self._put(ps.pod_vals), True))
                nbytes = ps.pod_vals.nbytes
            else:
                if len(ps.pod_rows):
                    segs.append((R.TABLE_POD_STATE, self._put(ps.pod_rows), self._put(ps.pod_vals)))
                nbytes = ps.pod_rows.nbytes + ps.pod_vals.nbytes
            if len(ps.grp_rows):
                segs.append((R.TABLE_GROUPS, self._put(ps.grp_rows), self._put(ps.grp_vals)))
                nbytes += ps.grp_rows.nbytes + ps.grp_vals.nbytes
            if len(ps.req_rows):
                segs.append((R.TABLE_PLACE_REQS, self._put(ps.req_rows), self._put(ps.req_vals)))
                nbytes += ps.req_rows.nbytes + ps.req_vals.nbytes
            self.ticks.append(engine.make_tick(segs, flags))
            self.h2d_bytes.append(int(nbytes))

    def _put(self, a: np.ndarray) -> np.ndarray:
        off = (self._cursor + 255) // 256 * 256
        view = self.arena[off: off + a.nbytes].view(a.dtype)
        view[...] = a.reshape(-1) if a.dtype.fields is None else a
        self._cursor = off + a.nbytes
        return view
