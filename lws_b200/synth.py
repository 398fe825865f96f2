"""Seeded synthetic record tables: the BASELINE.json configurations and a
branch-covering fuzzer for the parity tests.

BASELINE.json fixes only counts (objects x size x nodes); the value
distributions below are this build's choice and are stated in
``describe()`` so that every benchmark line can name its workload.  Everything
is vectorised numpy on a PCG64 stream seeded with 0x4C5753 ("LWS").
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import records as R

SEED = 0x4C5753


@dataclass
class Profile:
    """Value distributions of one synthetic cluster."""

    name: str
    n_lws: int
    size_choices: tuple = (8,)
    replicas_choices: tuple = (1,)
    n_nodes: int = 1000
    nodes_per_domain: int = 16
    node_capacity: int = 4
    max_surge: tuple = ((0, False),)  # (value, is_percent) choices
    max_unavailable: tuple = ((1, False),)
    # fractions of objects in each rollingUpdateParameters case
    f_no_sts: float = 0.02
    f_just_updated: float = 0.0
    f_scaling: float = 0.0  # sts replicas < lws replicas
    f_replicas_changed: float = 0.0
    f_mid_update: float = 0.1
    # per-group state
    p_group_updated: float = 0.5  # within mid-update objects
    p_group_ready: float = 0.9
    p_exclusive: float = 0.0
    p_leader_unscheduled: float = 0.05
    # per-pod state
    p_restarted: float = 0.01
    p_deleting: float = 0.005
    p_pending: float = 0.02
    p_short_group: float = 0.01  # group with a missing pod
    restart_policy_p: tuple = (0.1, 0.7, 0.2)  # None / OnPodRestart / AfterStart
    p_leader_ready_policy: float = 0.1
    fuzz: float = 0.0  # probability of flipping "can't happen" bits (parity fuzzing)
    n_namespaces: int = 1  # objects are dealt to namespaces round-robin (exclusivity is per namespace)
    conflict_free: bool = False  # scheduled leaders of one namespace sit in distinct domains
    gang: bool = True
    extra: dict = field(default_factory=dict)


# BASELINE.json configs (counts) with the SURVEY.md §8(d) distributions.
def profile(name: str, scale: float = 1.0) -> Profile:
    s = lambda n: max(1, int(round(n * scale)))
    if name == "C1":  # 10 LWS x size 4 on 16 nodes (plumbing)
        return Profile("C1", 10, size_choices=(4,), replicas_choices=(2,), n_nodes=16, nodes_per_domain=4,
                       p_exclusive=0.3)
    if name == "C2":  # 10k LWS x size 8, 1k nodes, placement off
        return Profile("C2", s(10_000), size_choices=(8,), replicas_choices=(1, 2, 4, 8), n_nodes=1000,
                       f_mid_update=0.10)
    if name == "C3":
        # 100k LWS x size 64, 10k nodes, topology-aware gang placement ON for every group: every object
        # carries the exclusive-topology annotation.  One domain holds one group per namespace, so the
        # objects live in enough namespaces to fit (200 x 625 domains >= 100k groups); scheduled leaders
        # of a namespace sit in distinct domains (what a scheduler honouring the constraint leaves
        # behind), 5 % of the leaders are not scheduled yet and compete for the free domains.
        # BASELINE's counts put ~600 pods on a node: capacity 650.
        return Profile("C3", s(100_000), size_choices=(64,), replicas_choices=(1,), n_nodes=10_000,
                       nodes_per_domain=16, node_capacity=650, p_exclusive=1.0, f_mid_update=0.10,
                       n_namespaces=max(1, int(round(200 * min(scale, 1.0)))) if scale < 1.0 else 200,
                       conflict_free=True)
    if name == "C3-steady":  # round 1's placement load: 1 % of the objects exclusive, one namespace
        return Profile("C3-steady", s(100_000), size_choices=(64,), replicas_choices=(1,), n_nodes=10_000,
                       nodes_per_domain=16, node_capacity=650, p_exclusive=0.01, f_mid_update=0.10)
    if name == "C5":  # 100k LWS rolling update maxSurge=10% + restart sweep
        return Profile("C5", s(100_000), size_choices=(8,), replicas_choices=(16,), n_nodes=10_000,
                       max_surge=((10, True),), max_unavailable=((1, False),), f_no_sts=0.02,
                       f_just_updated=0.10, f_scaling=0.08, f_replicas_changed=0.05, f_mid_update=0.60,
                       restart_policy_p=(0.1, 0.7, 0.2))
    if name == "fuzz":
        return Profile("fuzz", s(2000), size_choices=(1, 2, 3, 4, 8, 33, 64, 70),
                       replicas_choices=(0, 1, 2, 3, 4, 7, 16, 40),
                       n_nodes=64, nodes_per_domain=4,
                       max_surge=((0, False), (1, False), (2, False), (5, False), (10, True), (50, True), (100, True)),
                       max_unavailable=((0, False), (1, False), (2, False), (10, False), (25, True), (100, True)),
                       f_no_sts=0.05, f_just_updated=0.1, f_scaling=0.1, f_replicas_changed=0.1,
                       f_mid_update=0.5, p_group_ready=0.7, p_exclusive=0.4, p_leader_unscheduled=0.2,
                       p_restarted=0.05, p_deleting=0.03, p_pending=0.05, p_short_group=0.1,
                       restart_policy_p=(0.2, 0.5, 0.3), p_leader_ready_policy=0.3, fuzz=0.08)
    raise KeyError(name)


@dataclass
class Tables:
    profile: Profile
    lws: np.ndarray
    groups: np.ndarray
    pod_state: np.ndarray
    pod_ident: np.ndarray
    nodes: np.ndarray
    n_domains: int
    flags: int
    ns_of_lws: np.ndarray = None  # dense namespace id per object
    n_namespaces: int = 1

    def place_requests(self) -> np.ndarray:
        from . import encoder

        return encoder.encode_place_requests(self.lws, self.groups, self.ns_of_lws)

    def algorithmic_bytes(self) -> int:
        """Compulsory HBM traffic of one sweep: every input row read once, every
        output row written once."""
        return (
            len(self.lws) * (R.LWS_REC.itemsize + R.LWS_OUT.itemsize)
            + len(self.groups) * (R.GROUP_REC.itemsize + R.GROUP_OUT.itemsize)
            + len(self.pod_state) * R.POD_STATE.itemsize
            + self.event_pods() * R.POD_IDENT.itemsize
        )

    def event_pods(self) -> int:
        """Pods with a restart / deletion event: the only identity rows a sweep reads."""
        b = self.pod_state
        phase = b & R.POD_PHASE_MASK
        ev = (((phase == R.POD_PHASE_PENDING) | (phase == R.POD_PHASE_RUNNING)) & ((b & R.POD_ANY_RESTART) != 0)) | (
            (b & R.POD_DELETING) != 0
        )
        return int(ev.sum())

    def table_bytes(self) -> int:
        """Bytes of all input tables (what a full host→device upload moves)."""
        return int(self.lws.nbytes + self.groups.nbytes + self.pod_state.nbytes + self.pod_ident.nbytes)

    def describe(self) -> dict:
        p = self.profile
        return {
            "workload": p.name,
            "lws": int(len(self.lws)),
            "groups": int(len(self.groups)),
            "pods": int(len(self.pod_state)),
            "nodes": int(len(self.nodes)),
            "domains": int(self.n_domains),
            "namespaces": int(self.n_namespaces),
            "size": list(p.size_choices),
            "replicas": list(p.replicas_choices),
            "seed": hex(SEED),
        }


def _u64(rng, n):
    return rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(
        0, 2, size=n, dtype=np.uint64
    )


def _u32(rng, n):
    return rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)


def make_nodes(p: Profile, rng) -> tuple[np.ndarray, int]:
    n = p.n_nodes
    nodes = R.aligned_empty(n, R.NODE_REC)
    dom = (np.arange(n) // p.nodes_per_domain).astype(np.uint32)
    n_domains = int(dom.max()) + 1 if n else 0
    dom_hash = _u64(rng, max(n_domains, 1))
    nodes["topo_value_hash"] = dom_hash[dom]
    nodes["domain_id"] = dom
    nodes["capacity"] = p.node_capacity
    flags = np.full(n, R.NODE_HAS_TOPOLOGY | R.NODE_SCHEDULABLE, dtype=np.uint16)
    if p.fuzz:
        drop = rng.random(n) < p.fuzz
        flags[drop] &= ~np.uint16(R.NODE_HAS_TOPOLOGY)
        nodes["domain_id"][drop] = R.NONE
        cord = rng.random(n) < p.fuzz
        flags[cord] &= ~np.uint16(R.NODE_SCHEDULABLE)
    nodes["flags"] = flags
    return nodes, n_domains


def make(name_or_profile, scale: float = 1.0, seed: int = SEED) -> Tables:
    p = name_or_profile if isinstance(name_or_profile, Profile) else profile(name_or_profile, scale)
    rng = np.random.Generator(np.random.PCG64(seed))
    n = p.n_lws
    nodes, n_domains = make_nodes(p, rng)

    # ------------------------------------------------------------- LWS rows
    lws = R.aligned_empty(n, R.LWS_REC)
    size = rng.choice(np.array(p.size_choices, dtype=np.int32), size=n)
    replicas = rng.choice(np.array(p.replicas_choices, dtype=np.int32), size=n)
    si = rng.integers(0, len(p.max_surge), size=n)
    ui = rng.integers(0, len(p.max_unavailable), size=n)
    surge_v = np.array([v for v, _ in p.max_surge], dtype=np.int32)[si]
    surge_p = np.array([b for _, b in p.max_surge], dtype=bool)[si]
    unav_v = np.array([v for v, _ in p.max_unavailable], dtype=np.int32)[ui]
    unav_p = np.array([b for _, b in p.max_unavailable], dtype=bool)[ui]
    surge_abs = np.where(surge_p, -(-(surge_v.astype(np.int64) * replicas) // 100), surge_v).astype(np.int32)
    surge_abs = np.minimum(surge_abs, replicas)

    case = rng.choice(
        6, size=n,
        p=_norm([p.f_no_sts, p.f_just_updated, p.f_scaling, p.f_replicas_changed, p.f_mid_update,
                 max(0.0, 1 - p.f_no_sts - p.f_just_updated - p.f_scaling - p.f_replicas_changed - p.f_mid_update)]),
    )  # 0 no sts, 1 just updated, 2 scaling up, 3 replicas changed, 4 mid update, 5 steady
    flags = np.zeros(n, dtype=np.uint32)
    flags |= np.where(surge_p, R.LWS_SURGE_IS_PERCENT, 0).astype(np.uint32)
    flags |= np.where(unav_p, R.LWS_UNAVAIL_IS_PERCENT, 0).astype(np.uint32)
    flags |= np.where(case != 0, R.LWS_STS_EXISTS, 0).astype(np.uint32)
    flags |= np.where(case == 1, R.LWS_UPDATED, 0).astype(np.uint32)
    flags |= np.uint32(R.LWS_ANNOT_VALID)
    policy = rng.choice(3, size=n, p=_norm(p.restart_policy_p)).astype(np.uint32)
    flags |= policy << np.uint32(R.LWS_RESTART_SHIFT)
    flags |= np.where(rng.random(n) < 0.05, R.LWS_RECREATE_AFTER_START_ANNOT, 0).astype(np.uint32)
    flags |= np.where(rng.random(n) < p.p_leader_ready_policy, R.LWS_STARTUP_LEADER_READY, 0).astype(np.uint32)
    exclusive = rng.random(n) < p.p_exclusive
    flags |= np.where(exclusive, R.LWS_EXCLUSIVE_TOPOLOGY, 0).astype(np.uint32)
    if p.fuzz:
        for bit in (R.LWS_GROUP_LABEL_INVALID, R.LWS_INTSTR_INVALID, R.LWS_IRREGULAR):
            flags |= np.where(rng.random(n) < p.fuzz / 4, bit, 0).astype(np.uint32)
        flags &= ~np.where(rng.random(n) < p.fuzz / 2, R.LWS_ANNOT_VALID, 0).astype(np.uint32)

    sts_replicas = replicas.copy()
    mid = (case == 4) | (case == 3)
    sts_replicas[mid] = (replicas + np.where(rng.random(n) < 0.7, surge_abs, 0))[mid]
    sts_replicas[case == 2] = np.maximum(replicas[case == 2] - rng.integers(1, 3, size=(case == 2).sum()), 0)
    sts_partition = np.zeros(n, dtype=np.int32)
    sts_partition[mid] = (rng.random(mid.sum()) * (sts_replicas[mid] + 1)).astype(np.int32)
    annot = replicas.copy()
    annot[case == 3] += rng.choice(np.array([-2, -1, 1, 3]), size=(case == 3).sum()).astype(np.int32)
    lws_partition = np.where(rng.random(n) < (0.05 + p.fuzz), (rng.random(n) * (replicas + 2)).astype(np.int32), 0)
    if p.fuzz:
        odd = rng.random(n) < p.fuzz
        sts_replicas[odd] = np.maximum(0, sts_replicas[odd] + rng.integers(-2, 4, size=odd.sum())).astype(np.int32)
        sts_partition[odd] = rng.integers(0, 6, size=odd.sum())

    # group rows per object: every slot of the leader sts, sometimes fewer / more
    group_count = np.where(case == 0, 0, sts_replicas).astype(np.int64)
    jitter = rng.random(n)
    lo = 0.03 + p.fuzz
    group_count = np.where(jitter < lo, np.maximum(group_count - 1, 0), group_count)
    group_count = np.where(jitter > 1 - lo, group_count + 1, group_count)
    group_base = np.concatenate([[0], np.cumsum(group_count)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
    G = int(group_count.sum())

    lws["rev_hash"] = _u64(rng, n)
    lws["size"] = size
    lws["flags"] = flags
    lws["replicas"] = replicas
    lws["partition"] = lws_partition
    lws["max_surge"] = surge_v
    lws["max_unavailable"] = unav_v
    lws["sts_replicas"] = sts_replicas
    lws["sts_partition"] = sts_partition
    lws["sts_replicas_annotation"] = annot
    lws["subgroup_size"] = 0
    lws["uid_hash"] = _u64(rng, n)
    lws["group_base"] = group_base.astype(np.uint32)
    lws["group_count"] = group_count.astype(np.uint32)

    # ----------------------------------------------------------- group rows
    owner = np.repeat(np.arange(n, dtype=np.int64), group_count)
    gidx = np.arange(G, dtype=np.int64) - group_base[owner]
    g = R.aligned_empty(G, R.GROUP_REC)
    o_rev = lws["rev_hash"][owner]
    o_size = size[owner]
    o_case = case[owner]
    old_rev = o_rev ^ np.uint64(0x9E3779B97F4A7C15)
    # mid-update objects: slots at/above a moving frontier carry the new revision
    upd = np.ones(G, dtype=bool)
    in_update = (o_case == 4) | (o_case == 3) | (o_case == 1)
    upd[in_update] = rng.random(in_update.sum()) < p.p_group_updated
    upd[o_case == 1] = False
    g["leader_rev_hash"] = np.where(upd, o_rev, old_rev)
    wsts_upd = upd & (rng.random(G) > (0.02 + p.fuzz))
    g["wsts_rev_hash"] = np.where(wsts_upd, o_rev, old_rev)
    g["wsts_spec_replicas"] = o_size - 1
    ready = rng.random(G) < np.where(in_update, p.p_group_ready, 0.98)
    g["wsts_avail_replicas"] = np.where(ready, o_size - 1, np.maximum(o_size - 2, 0))
    g["leader_uid_hash"] = _u32(rng, G)
    g["wsts_uid_hash"] = _u32(rng, G)
    stale_owner = rng.random(G) < (0.02 + p.fuzz)
    g["wsts_owner_uid_hash"] = np.where(stale_owner, _u32(rng, G), g["leader_uid_hash"])
    sched = rng.random(G) >= p.p_leader_unscheduled
    node = rng.integers(0, max(p.n_nodes, 1), size=G).astype(np.uint32)
    ns_of_lws = (np.arange(n, dtype=np.int64) % max(p.n_namespaces, 1)).astype(np.uint32)
    if p.conflict_free and n_domains:
        # the k-th group of a namespace takes the k-th domain of that namespace's own permutation
        # (groups beyond the number of domains stay unscheduled), a random node inside it
        g_ns = ns_of_lws[owner].astype(np.int64)
        order = np.argsort(g_ns, kind="stable")
        first = np.searchsorted(g_ns[order], np.arange(max(p.n_namespaces, 1)))
        rank = np.empty(G, dtype=np.int64)
        rank[order] = np.arange(G) - first[g_ns[order]]
        perm = np.argsort(rng.random((max(p.n_namespaces, 1), n_domains)), axis=1)
        dom = perm[g_ns, np.minimum(rank, n_domains - 1)]
        sched &= rank < n_domains
        node = np.minimum(dom * p.nodes_per_domain + rng.integers(0, max(p.nodes_per_domain, 1), size=G),
                          max(p.n_nodes - 1, 0)).astype(np.uint32)
    leader_node = np.where(sched, node, np.uint32(R.NONE)).astype(np.uint32)
    if p.fuzz:
        leader_node = np.where(rng.random(G) < p.fuzz / 2, np.uint32(R.NODE_NOT_FOUND), leader_node)
    g["leader_node"] = leader_node
    g["lws_index"] = owner.astype(np.uint32)
    gf = np.full(
        G,
        R.GRP_POD_PRESENT | R.GRP_POD_NAME_MATCH | R.GRP_POD_RUNNING | R.GRP_WSTS_LABEL_NAME_MATCH
        | R.GRP_WSTS_FOUND | R.GRP_WSTS_REV_SETTLED | R.GRP_WSTS_OWNER_IS_POD | R.GRP_WSTS_OWNER_NAME_MATCH
        | R.GRP_REVISION_EXISTS,
        dtype=np.uint32,
    )
    gf |= np.where(ready, R.GRP_POD_READY, 0).astype(np.uint32)
    leader_deleting = rng.random(G) < p.p_deleting
    gf |= np.where(leader_deleting, R.GRP_POD_DELETING, 0).astype(np.uint32)
    # a group whose worker sts does not exist yet (→ CREATE_WSTS path)
    no_wsts = rng.random(G) < (0.03 + p.fuzz)
    gf &= ~np.where(no_wsts, R.GRP_WSTS_FOUND | R.GRP_WSTS_LABEL_NAME_MATCH, 0).astype(np.uint32)
    if p.fuzz:
        for bit in (R.GRP_POD_PRESENT, R.GRP_POD_NAME_MATCH, R.GRP_POD_RUNNING, R.GRP_WSTS_LABEL_NAME_MATCH,
                    R.GRP_WSTS_FOUND, R.GRP_WSTS_REV_SETTLED, R.GRP_WSTS_OWNER_IS_POD,
                    R.GRP_WSTS_OWNER_NAME_MATCH, R.GRP_REVISION_EXISTS):
            gf &= ~np.where(rng.random(G) < p.fuzz / 2, bit, 0).astype(np.uint32)
        gf |= np.where(rng.random(G) < p.fuzz / 2, R.GRP_MISTAKEN_ANNOTATION, 0).astype(np.uint32)
    g["flags"] = gf

    # ------------------------------------------------------------- pod rows
    pod_count = o_size.astype(np.int64).copy()
    short = rng.random(G) < p.p_short_group
    pod_count[short] = np.maximum(pod_count[short] - 1, 0)
    if p.fuzz:
        extra = rng.random(G) < p.fuzz / 2
        pod_count[extra] += 1
        pod_count[~(gf & R.GRP_POD_PRESENT).astype(bool) & (rng.random(G) < 0.5)] = 0
    pod_base = np.concatenate([[0], np.cumsum(pod_count)[:-1]]).astype(np.int64) if G else np.zeros(0, np.int64)
    Pn = int(pod_count.sum())
    g["pod_base"] = pod_base.astype(np.uint32)
    g["pod_count"] = pod_count.astype(np.uint32)

    pg = np.repeat(np.arange(G, dtype=np.int64), pod_count)  # pod → group
    within = np.arange(Pn, dtype=np.int64) - pod_base[pg]
    same_rev = rng.random(Pn) > (0.02 + p.fuzz)
    pod_rev = np.where(same_rev, g["leader_rev_hash"][pg], g["leader_rev_hash"][pg] ^ np.uint64(1))
    is_leader = within == 0
    if p.fuzz:
        is_leader &= rng.random(Pn) > p.fuzz / 2
    kind = np.full(Pn, R.POD_OWNER_STS, dtype=np.uint32)
    owner_uid = g["wsts_uid_hash"][pg].copy()
    if p.fuzz:
        r = rng.random(Pn)
        k_pod = r < p.fuzz
        k_none = (r >= p.fuzz) & (r < 1.5 * p.fuzz)
        k_other = (r >= 1.5 * p.fuzz) & (r < 2 * p.fuzz)
        kind[k_pod] = R.POD_OWNER_POD
        owner_uid[k_pod] = g["leader_uid_hash"][pg][k_pod]
        kind[k_none] = R.POD_OWNER_NONE
        kind[k_other] = R.POD_OWNER_OTHER
    stale = rng.random(Pn) < (0.01 + p.fuzz)
    owner_uid = np.where(stale, _u32(rng, Pn), owner_uid)
    phase = np.full(Pn, R.POD_PHASE_RUNNING, dtype=np.uint32)
    phase[rng.random(Pn) < p.p_pending] = R.POD_PHASE_PENDING
    if p.fuzz:
        phase[rng.random(Pn) < p.fuzz] = 0
        phase[rng.random(Pn) < p.fuzz / 4] = 3
    bits = phase
    bits |= np.where(rng.random(Pn) < p.p_restarted, R.POD_ANY_RESTART, 0).astype(np.uint32)
    bits |= np.where(rng.random(Pn) < p.p_deleting, R.POD_DELETING, 0).astype(np.uint32)
    bits |= np.where(is_leader & leader_deleting[pg], R.POD_DELETING, 0).astype(np.uint32)
    bits |= kind << np.uint32(R.POD_OWNER_SHIFT)
    name_match = rng.random(Pn) > p.fuzz
    bits |= np.where(name_match, R.POD_OWNER_NAME_MATCH, 0).astype(np.uint32)
    bits |= np.where(is_leader, R.POD_IS_LEADER, 0).astype(np.uint32)
    place = np.where(rng.random(Pn) > p.fuzz / 2, R.PODID_NAME_OK, 0).astype(np.uint32)
    # placement: pods of a group sit in the leader's domain, spread over its nodes
    gl = g["leader_node"][pg]
    g_sched = (gl != R.NONE) & (gl != R.NODE_NOT_FOUND)
    dom_first = (gl // np.uint32(max(p.nodes_per_domain, 1))) * np.uint32(max(p.nodes_per_domain, 1))
    pnode = np.minimum(dom_first + (within % max(p.nodes_per_domain, 1)).astype(np.uint32),
                       np.uint32(max(p.n_nodes - 1, 0)))
    p_sched = g_sched & (rng.random(Pn) > p.p_pending)
    place |= np.where(p_sched, np.uint32(R.PODID_SCHEDULED) | (pnode << np.uint32(R.PODID_NODE_SHIFT)), 0).astype(np.uint32)
    pod_ident = R.pod_ident_table(pod_rev, owner_uid, place)
    pod_state = R.aligned_empty(Pn, R.POD_STATE)
    pod_state[:] = bits.astype(np.uint8)

    return Tables(profile=p, lws=lws, groups=g, pod_state=pod_state, pod_ident=pod_ident, nodes=nodes, n_domains=n_domains,
                  flags=R.SWEEP_GANG if p.gang else 0, ns_of_lws=ns_of_lws, n_namespaces=max(p.n_namespaces, 1))


def _norm(ps):
    a = np.asarray(ps, dtype=np.float64)
    return a / a.sum()


# --------------------------------------------------------------------------- #
# DisaggregatedSet tables (BASELINE.json configs[3]: 2-role prefill/decode)
# --------------------------------------------------------------------------- #
@dataclass
class DsTables:
    ds: np.ndarray
    roles: np.ndarray
    revroles: np.ndarray

    def algorithmic_bytes(self) -> int:
        return (len(self.ds) * (R.DS_REC.itemsize + R.DS_OUT.itemsize)
                + len(self.roles) * (R.DS_ROLE_REC.itemsize + R.DS_ROLE_OUT.itemsize)
                + len(self.revroles) * (R.DS_REVROLE_REC.itemsize + 4))


def make_ds(n_ds: int, n_roles_choices=(2,), max_old_revs=3, seed: int = SEED, fuzz: float = 0.0) -> DsTables:
    """Random mid-rollout DisaggregatedSets.  C4: 2 roles (prefill:decode ratios from
    {1:1, 2:1, 5:2, 10:1}, replicas <= 32), maxSurge in {1, 2, 25%}, 1-3 old revisions
    with distinct creation stamps, some roles missing / drained / not yet ready."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_roles = rng.choice(np.array(n_roles_choices), size=n_ds).astype(np.int64)
    n_removed = np.where(rng.random(n_ds) < (0.1 + fuzz), 1, 0) * (n_roles < R.DS_MAX_ROLES)
    n_all = n_roles + n_removed
    n_old = rng.integers(0 if fuzz else 1, max_old_revs + 1, size=n_ds)
    role_base = np.concatenate([[0], np.cumsum(n_all)[:-1]])
    rr_count = (n_old + 1) * n_all
    rev_base = np.concatenate([[0], np.cumsum(rr_count)[:-1]])
    ds = R.aligned_empty(n_ds, R.DS_REC)
    ds["uid_hash"] = _u64(rng, n_ds)
    ds["role_base"], ds["n_roles"], ds["n_spec_roles"] = role_base, n_all, n_roles
    ds["rev_base"], ds["n_old_revs"] = rev_base, n_old
    ds["flags"] = np.where(rng.random(n_ds) < 0.9, R.DS_HAS_NEW_REVISION, 0)

    NR = int(n_all.sum())
    owner = np.repeat(np.arange(n_ds), n_all)
    ridx = np.arange(NR) - role_base[owner]
    in_spec = ridx < n_roles[owner]
    roles = R.aligned_empty(NR, R.DS_ROLE_REC)
    ratio = rng.choice(np.array([1, 2, 5, 10]), size=NR)
    base = rng.integers(0 if fuzz else 1, 5, size=NR)
    target = np.minimum(base * np.where(ridx == 0, ratio, rng.choice(np.array([1, 2]), size=NR)), 32)
    roles["target_replicas"] = np.where(in_spec, target, 0)
    kind = rng.integers(0, 6, size=NR)  # 0 none, 1 surge1, 2 surge2, 3 surge25%, 4 unavail1, 5 both
    rflags = np.where(in_spec, R.ROLE_IN_SPEC, 0).astype(np.uint32)
    rflags |= np.where(in_spec & (kind > 0), R.ROLE_HAS_ROLLING_CONFIG, 0).astype(np.uint32)
    surge = np.select([kind == 1, kind == 2, kind == 3, kind == 5], [1, 2, 25, 1], 0)
    rflags |= np.where(kind == 3, R.ROLE_SURGE_IS_PERCENT, 0).astype(np.uint32)
    unav = np.select([kind == 4, kind == 5], [1, 2], 0)
    if fuzz:
        rflags |= np.where(rng.random(NR) < fuzz, R.ROLE_SURGE_INVALID, 0).astype(np.uint32)
        rflags |= np.where(rng.random(NR) < fuzz, R.ROLE_UNAVAIL_IS_PERCENT, 0).astype(np.uint32)
        unav = np.where(rng.random(NR) < fuzz, 50, unav)
    roles["max_surge"], roles["max_unavailable"], roles["flags"] = surge, unav, rflags

    NRR = int(rr_count.sum())
    d_of = np.repeat(np.arange(n_ds), rr_count)
    k = np.arange(NRR) - rev_base[d_of]
    rev = k // n_all[d_of]
    role = k % n_all[d_of]
    is_new = rev == n_old[d_of]
    tgt = roles["target_replicas"][role_base[d_of] + role].astype(np.int64)
    rr = R.aligned_empty(NRR, R.DS_REVROLE_REC)
    spec_role = role < n_roles[d_of]
    exists = np.where(is_new, spec_role & (ds["flags"][d_of] & R.DS_HAS_NEW_REVISION != 0),
                      rng.random(NRR) > (0.05 + fuzz))
    # old revisions hold a random share of an initial count; the new one a random progress
    initial = np.maximum(tgt + rng.integers(-2, 3, size=NRR), 0)
    initial = np.where(spec_role, initial, rng.integers(1, 5, size=NRR))
    progress = rng.random(NRR)
    replicas = np.where(is_new, np.floor(progress * (tgt + 1)), np.ceil(progress * initial)).astype(np.int64)
    replicas = np.where(rng.random(NRR) < 0.15, 0, replicas)
    rr["replicas"] = np.where(exists, replicas, 0)
    has_annot = rng.random(NRR) < 0.8
    rr["initial_replicas"] = np.where(exists & ~is_new & has_annot, initial, -1)
    not_ready = rng.random(NRR) < 0.1
    rr["ready_replicas"] = np.where(exists, np.where(not_ready, np.maximum(replicas - 1, 0), replicas), 0)
    stamp = (rev + 1) * 7 + (d_of % 5)  # distinct per revision within a DS
    if fuzz:
        stamp = np.where(rng.random(NRR) < fuzz, 3, stamp)  # equal stamps exercise the stable order
    f = np.where(exists, R.RR_EXISTS, 0).astype(np.uint32) | (stamp.astype(np.uint32) << np.uint32(R.RR_TS_SHIFT))
    f |= np.where(exists & (rng.random(NRR) < 0.02), R.RR_REPLICAS_NIL, 0).astype(np.uint32)
    rr["replicas"] = np.where((f & R.RR_REPLICAS_NIL) != 0, 1, rr["replicas"])
    rr["flags"] = np.where(exists, f, 0)
    return DsTables(ds, roles, rr)
