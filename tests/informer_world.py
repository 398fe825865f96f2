"""A small random cluster and a random watch-event stream for the incremental-encoder tests
(tests/test_informer.py, tests/test_gpu_tick.py)."""
import copy

import numpy as np

from lws_b200 import api, encoder


def make_world(seed=0, n_lws=12, n_nodes=24, topology_key="zone"):
    rng = np.random.default_rng(seed)
    nodes = [api.Node(name=f"node-{i}", labels={topology_key: f"z{i // 4}"} if i % 11 else {}, capacity=16) for i in range(n_nodes)]
    items, pods, stss = [], [], []
    for li in range(n_lws):
        ns = f"ns-{li % 3}"
        size = int(rng.choice([1, 2, 3, 5]))
        replicas = int(rng.choice([1, 2, 3, 4]))
        ann = {api.ExclusiveKeyAnnotationKey: topology_key} if li % 2 == 0 else {}
        lws = api.LeaderWorkerSet(name=f"lws-{li}", namespace=ns, replicas=replicas, size=size, annotations=ann,
                                  rollingUpdate=api.RollingUpdateConfiguration(0, 1, int(rng.choice([0, 1]))),
                                  restartPolicy=str(rng.choice([api.RecreateGroupOnPodRestart, api.RecreateGroupAfterStart, api.NoneRestartPolicy])))
        rev = "rev-1"
        sts = api.StatefulSet(name=lws.name, namespace=ns, replicas=replicas, partition=0,
                              labels={api.SetNameLabelKey: lws.name, api.RevisionKey: rev},
                              annotations={api.ReplicasAnnotationKey: str(replicas)})
        items.append(encoder.LwsItem(lws=lws, revision_key=rev, leader_sts=sts))
        for g in range(replicas):
            leader = api.Pod(name=f"{lws.name}-{g}", namespace=ns, phase="Running", readyCondition=True,
                             labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: str(g),
                                     api.WorkerIndexLabelKey: "0", api.RevisionKey: rev},
                             nodeName=f"node-{int(rng.integers(0, n_nodes))}" if rng.random() < 0.7 else "")
            pods.append(leader)
            if size > 1:
                w = api.StatefulSet(name=leader.name, namespace=ns, replicas=size - 1, availableReplicas=size - 1,
                                    currentRevision="a", updateRevision="a",
                                    labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: str(g), api.RevisionKey: rev},
                                    ownerReferences=[api.OwnerReference("Pod", leader.name, leader.uid)])
                stss.append(w)
                for wi in range(1, size):
                    pods.append(api.Pod(name=f"{leader.name}-{wi}", namespace=ns, phase="Running", readyCondition=True,
                                        labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: str(g),
                                                api.WorkerIndexLabelKey: str(wi), api.RevisionKey: rev},
                                        ownerReferences=[api.OwnerReference("StatefulSet", w.name, w.uid)],
                                        nodeName=leader.nodeName))
    return items, encoder.Cluster(pods=pods, statefulsets=stss, nodes=nodes)


class World:
    """The 'API server': holds the objects, mutates them at random, reports each mutation as a watch event."""

    def __init__(self, items, cluster, seed=1):
        self.items = {(it.lws.namespace, it.lws.name): it for it in items}
        self.pods = {(p.namespace, p.name): p for p in cluster.pods}
        self.stss = {(s.namespace, s.name): s for s in cluster.statefulsets}
        self.nodes = cluster.nodes
        self.rng = np.random.default_rng(seed)
        self.serial = 0

    def cluster(self):
        return encoder.Cluster(pods=list(self.pods.values()), statefulsets=list(self.stss.values()), nodes=self.nodes)

    def step(self, enc):
        """One random mutation, delivered to `enc` as the informer would."""
        rng = self.rng
        r = rng.random()
        self.serial += 1
        if r < 0.45 and self.pods:  # pod status update
            key = list(self.pods)[int(rng.integers(0, len(self.pods)))]
            p = copy.deepcopy(self.pods[key])
            c = rng.random()
            if c < 0.3:
                p.phase = "Pending" if p.phase == "Running" else "Running"
            elif c < 0.5:
                p.readyCondition = not p.readyCondition
            elif c < 0.7:
                p.containerRestartCounts = [int(rng.integers(0, 3))]
            elif c < 0.8:
                p.deletionTimestamp = not p.deletionTimestamp
            else:
                p.nodeName = f"node-{int(rng.integers(0, len(self.nodes)))}" if rng.random() < 0.8 else ""
            self.pods[key] = p
            enc.pod_event("MODIFIED", p)
        elif r < 0.6 and self.pods:  # pod deleted
            key = list(self.pods)[int(rng.integers(0, len(self.pods)))]
            p = self.pods.pop(key)
            enc.pod_event("DELETED", p)
        elif r < 0.75:  # a pod (re)created, possibly in a new group slot
            it = list(self.items.values())[int(rng.integers(0, len(self.items)))]
            lws = it.lws
            g = int(rng.integers(0, lws.replicas + 2))
            wi = int(rng.integers(0, max(lws.size, 1)))
            name = f"{lws.name}-{g}" if wi == 0 else f"{lws.name}-{g}-{wi}"
            owner = [] if wi == 0 else [api.OwnerReference("StatefulSet", f"{lws.name}-{g}", f"uid-sts-{lws.namespace}-{lws.name}-{g}")]
            p = api.Pod(name=name, namespace=lws.namespace, uid=f"uid-{self.serial}", phase=str(rng.choice(["Pending", "Running"])),
                        readyCondition=bool(rng.random() < 0.5),
                        labels={api.SetNameLabelKey: lws.name, api.GroupIndexLabelKey: str(g), api.WorkerIndexLabelKey: str(wi),
                                api.RevisionKey: it.revision_key}, ownerReferences=owner)
            kind = "MODIFIED" if (lws.namespace, name) in self.pods else "ADDED"
            self.pods[(lws.namespace, name)] = p
            enc.pod_event(kind, p)
        elif r < 0.85 and self.stss:  # worker sts status
            key = list(self.stss)[int(rng.integers(0, len(self.stss)))]
            s = copy.deepcopy(self.stss[key])
            s.availableReplicas = int(rng.integers(0, s.replicas + 1))
            s.updateRevision = str(rng.choice(["a", "b"]))
            self.stss[key] = s
            enc.statefulset_event("MODIFIED", s)
        elif r < 0.93:  # leader sts moves (partition / replicas), as the LWS controller's SSA would
            it = list(self.items.values())[int(rng.integers(0, len(self.items)))]
            s = copy.deepcopy(it.leader_sts)
            s.partition = int(rng.integers(0, s.replicas + 1))
            s.replicas = max(0, s.replicas + int(rng.integers(-1, 3)))
            it.leader_sts = s
            enc.statefulset_event("MODIFIED", s)
        else:  # the object itself: scale / template update
            key = list(self.items)[int(rng.integers(0, len(self.items)))]
            it = copy.deepcopy(self.items[key])
            if rng.random() < 0.5:
                it.lws.replicas = max(0, it.lws.replicas + int(rng.integers(-1, 3)))
            else:
                it.revision_key = f"rev-{self.serial}"
                it.lws_updated = True
            self.items[key] = it
            enc.lws_event("MODIFIED", it)


def outputs_by_name(tables_lws, tables_groups, lws_out, group_out, names):
    """{(object, group index): group_out row bytes}, {object: lws_out row bytes} — layout independent."""
    g, l = {}, {}
    for row, name in names.items():
        l[name] = lws_out[row].tobytes()
        base, count = int(tables_lws["group_base"][row]), int(tables_lws["group_count"][row])
        for gi in range(count):
            g[(name, gi)] = group_out[base + gi].tobytes()
    return l, g
