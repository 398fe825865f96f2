#!/usr/bin/env python
"""Extract the reference's expected StatefulSet apply configurations into JSON fixtures.

    python tests/golden/extract_apply_config_vectors.py   # needs /root/reference (this container only)

Sources (kubernetes-sigs/lws @ 1d9204a2):
  pkg/controllers/leaderworkerset_controller_test.go:50-758  TestLeaderStatefulSetApplyConfig
  pkg/controllers/pod_controller_test.go:42-425              TestConstructWorkerStatefulSetApplyConfiguration
The table entries are Go composite literals (apply-configuration structs, builder chains,
`ptr.To`, `intstr.From*`, label-key constants).  A small recursive-descent parser turns every
entry into plain data:
  * `T{Field: v}`            -> {"field": v}; TypeMeta / ObjectMeta embeddings fold into
                                 kind / apiVersion / "metadata" as the JSON encoding of the
                                 apply configuration does;
  * `X().WithA(a).WithB(b)`  -> {"a": a, "b": b}   (apply-configuration builders);
  * `ptr.To[T](v)`, `intstr.FromInt32(v)`, `resource.MustParse(v)` -> v;
  * known constants          -> their values; test-local variables (revision keys) -> "$name".
The LeaderWorkerSet of an entry is a wrappers.* builder chain; it is kept as a list of
[method, args...] steps for the test to replay on the host object model.
Writes tests/golden/apply_configs.json.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CONSTS = {
    "leaderworkerset.SetNameLabelKey": "leaderworkerset.sigs.k8s.io/name",
    "leaderworkerset.GroupIndexLabelKey": "leaderworkerset.sigs.k8s.io/group-index",
    "leaderworkerset.WorkerIndexLabelKey": "leaderworkerset.sigs.k8s.io/worker-index",
    "leaderworkerset.GroupUniqueHashLabelKey": "leaderworkerset.sigs.k8s.io/group-key",
    "leaderworkerset.RevisionKey": "leaderworkerset.sigs.k8s.io/template-revision-hash",
    "leaderworkerset.SubGroupIndexLabelKey": "leaderworkerset.sigs.k8s.io/subgroup-index",
    "leaderworkerset.SubGroupSizeAnnotationKey": "leaderworkerset.sigs.k8s.io/subgroup-size",
    "leaderworkerset.SubGroupUniqueHashLabelKey": "leaderworkerset.sigs.k8s.io/subgroup-key",
    "leaderworkerset.ExclusiveKeyAnnotationKey": "leaderworkerset.sigs.k8s.io/exclusive-topology",
    "leaderworkerset.SubGroupExclusiveKeyAnnotationKey": "leaderworkerset.sigs.k8s.io/subgroup-exclusive-topology",
    "leaderworkerset.SizeAnnotationKey": "leaderworkerset.sigs.k8s.io/size",
    "leaderworkerset.LeaderPodNameAnnotationKey": "leaderworkerset.sigs.k8s.io/leader-name",
    "leaderworkerset.ReplicasAnnotationKey": "leaderworkerset.sigs.k8s.io/replicas",
    "leaderworkerset.SubGroupPolicyTypeAnnotationKey": "leaderworkerset.sigs.k8s.io/subgroup-policy-type",
    "corev1.PersistentVolumeFilesystem": "Filesystem",
    "leaderworkerset.RollingUpdateStrategyType": "RollingUpdate",
    "leaderworkerset.RecreateGroupOnPodRestart": "RecreateGroupOnPodRestart",
    "leaderworkerset.SubGroupPolicyTypeLeaderWorker": "LeaderWorker",
    "leaderworkerset.SubGroupPolicyTypeLeaderExcluded": "LeaderExcluded",
    "leaderworkerset.SubdomainShared": "Shared",
    "leaderworkerset.SubdomainUniquePerReplica": "UniquePerReplica",
    "appsv1.ParallelPodManagement": "Parallel",
    "appsv1.RollingUpdateStatefulSetStrategyType": "RollingUpdate",
    "appsv1.RetainPersistentVolumeClaimRetentionPolicyType": "Retain",
    "appsv1.DeletePersistentVolumeClaimRetentionPolicyType": "Delete",
    "corev1.ProtocolTCP": "TCP",
    "corev1.ReadWriteOnce": "ReadWriteOnce",
    "corev1.ResourceStorage": "storage",
    "true": True, "false": False, "nil": None,
}

TOKEN = re.compile(r'\s*(?:(//[^\n]*)|("(?:[^"\\]|\\.)*")|(`[^`]*`)|([A-Za-z_][A-Za-z0-9_]*)|(\d+)|(.))', re.S)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            out.append(("str", json.loads(m.group(2))))
        elif m.group(3) is not None:
            out.append(("str", m.group(3)[1:-1]))
        elif m.group(4) is not None:
            out.append(("id", m.group(4)))
        elif m.group(5) is not None:
            out.append(("num", int(m.group(5))))
        elif m.group(6).strip():
            out.append(("p", m.group(6)))
    return out


def lower(name):
    return name[0].lower() + name[1:] if name else name


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def eat(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise SyntaxError(f"expected {kind} {val}, got {tok} near token {self.i}: {self.t[max(0, self.i - 6): self.i + 6]}")
        self.i += 1
        return tok

    def at(self, val):
        return self.peek() == ("p", val)

    # a (possibly qualified, possibly generic) name: a.b.c[T]
    def name(self):
        parts = [self.eat("id")[1]]
        while self.at(".") and self.peek(1)[0] == "id":
            self.eat()
            parts.append(self.eat("id")[1])
        return ".".join(parts)

    def type_suffix(self):
        """[T] / [K]V after a name, or leading []T / map[K]V / *T — consumed and ignored."""
        while self.at("["):
            depth = 0
            while True:
                tok = self.eat()
                if tok == ("p", "["):
                    depth += 1
                elif tok == ("p", "]"):
                    depth -= 1
                    if depth == 0:
                        break

    def value(self):
        tok = self.peek()
        if tok == ("p", "&") or tok == ("p", "*"):
            self.eat()
            return self.value()
        if tok[0] == "str":
            self.eat()
            return tok[1]
        if tok[0] == "num":
            self.eat()
            return tok[1]
        if tok == ("p", "-"):
            self.eat()
            return -self.eat("num")[1]
        if tok == ("p", "["):  # []T{...}
            self.type_suffix()
            if self.peek()[0] == "id":
                self.name()
                self.type_suffix()
            return self.composite(listy=True)
        if tok == ("p", "{"):  # element of a []T literal with the type elided
            return self.composite()
        if tok == ("id", "map"):
            self.eat()
            self.type_suffix()
            self.name()
            return self.composite(mapy=True)
        name = self.name()
        self.type_suffix()
        v = None
        if self.at("{"):
            v = self.composite(type_name=name)
        elif self.at("("):
            v = self.call(name)
        else:
            v = CONSTS.get(name, "$" + name)
        # method chain
        while self.at(".") and self.peek(1)[0] == "id":
            self.eat()
            m = self.eat("id")[1]
            self.type_suffix()
            args = self.args()
            v = self.method(v, m, args)
        return v

    def args(self):
        self.eat("p", "(")
        out = []
        while not self.at(")"):
            out.append(self.value())
            if self.at(","):
                self.eat()
            if self.at("."):  # variadic spread "..."
                while self.at("."):
                    self.eat()
        self.eat("p", ")")
        return out

    def call(self, name):
        args = self.args()
        if name.startswith("ptr.To") or name in ("intstr.FromInt32", "intstr.FromInt", "intstr.FromString", "resource.MustParse",
                                                 "corev1.Protocol", "appsv1.StatefulSetUpdateStrategyType", "string",
                                                 "int32", "metav1.NewTime", "appsv1.PodManagementPolicyType", "corev1.ResourceName",
                                                 "types.UID", "corev1.PodPhase"):
            return args[0]
        if name.startswith("wrappers.") or name.startswith("testutils."):
            return {"$chain": [[name.split(".", 1)[1]] + args]}
        # an apply-configuration constructor: StatefulSet(name, ns) / StatefulSetSpec() / ...
        if name.endswith(".StatefulSet") and len(args) == 2:
            return {"kind": "StatefulSet", "apiVersion": "apps/v1", "metadata": {"name": args[0], "namespace": args[1]}}
        return {} if not args else {"$call": name, "args": args}

    def method(self, v, m, args):
        if isinstance(v, dict) and "$chain" in v:
            v["$chain"].append([m] + args)
            return v
        if m.startswith("With") and isinstance(v, dict):
            key = lower(m[4:])
            val = args[0] if len(args) == 1 else args
            if key in ("labels", "annotations") and "kind" in v:
                v.setdefault("metadata", {})[key] = val
            else:
                v[key] = val
            return v
        if m in ("Obj", "DeepCopy"):
            return v
        return {"$method": m, "on": v, "args": args}

    def composite(self, type_name=None, listy=False, mapy=False):
        self.eat("p", "{")
        if listy:
            out = []
            while not self.at("}"):
                out.append(self.value())
                if self.at(","):
                    self.eat()
            self.eat("p", "}")
            return out
        out = {}
        while not self.at("}"):
            keyed = mapy or not (self.peek()[0] == "id" and self.peek(1) == ("p", ":"))  # a named map type: T{k: v}
            if keyed:
                k = self.value()
            else:
                k = self.eat("id")[1]
            self.eat("p", ":")
            v = self.value()
            if not keyed and k in CONSTS:  # a constant used as the key of a named map type
                k = CONSTS[k]
            elif not keyed:
                if k == "TypeMetaApplyConfiguration" or k == "TypeMeta":
                    out.update(v)
                    k = None
                elif k in ("ObjectMetaApplyConfiguration", "ObjectMeta"):
                    k = "metadata"
                else:
                    k = "apiVersion" if k == "APIVersion" else lower(k)
            if k is not None:
                out[k] = v
            if self.at(","):
                self.eat()
        self.eat("p", "}")
        return out


def table_entries(src, start_marker):
    """The `tests := []struct{...}{ {entry}, {entry}, ... }` table that follows start_marker."""
    i = src.index(start_marker)
    i = src.index("tests := []struct", i)
    i = src.index("}{", i) + 1  # the opening brace of the table literal
    toks = tokenize(src[i:])
    p = Parser(toks)
    p.eat("p", "{")
    entries = []
    while not p.at("}"):
        entries.append(p.composite())
        if p.at(","):
            p.eat()
    return entries


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    out = {}
    src = open(os.path.join(REF, "pkg/controllers/leaderworkerset_controller_test.go")).read()
    out["leader_statefulset"] = {
        "source": "pkg/controllers/leaderworkerset_controller_test.go:50-758 TestLeaderStatefulSetApplyConfig",
        "cases": table_entries(src, "func TestLeaderStatefulSetApplyConfig("),
    }
    src = open(os.path.join(REF, "pkg/controllers/pod_controller_test.go")).read()
    out["worker_statefulset"] = {
        "source": "pkg/controllers/pod_controller_test.go:42-425 TestConstructWorkerStatefulSetApplyConfiguration",
        "cases": table_entries(src, "func TestConstructWorkerStatefulSetApplyConfiguration("),
    }
    # the pod specs the builders refer to (test/wrappers/wrappers.go:279-294, :360-369)
    out["pod_specs"] = {
        "source": "test/wrappers/wrappers.go:279-294 MakeWorkerPodSpec, :360-369 MakeLeaderPodSpec",
        "MakeWorkerPodSpec": {"containers": [{"name": "worker", "image": "docker.io/nginxinc/nginx-unprivileged:1.27",
                                              "ports": [{"containerPort": 8080, "protocol": "TCP"}], "resources": {}}]},
        "MakeLeaderPodSpec": {"containers": [{"name": "leader", "image": "docker.io/nginxinc/nginx-unprivileged:1.27",
                                              "resources": {}}]},
    }
    path = os.path.join(HERE, "apply_configs.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, {k: len(v.get("cases", [])) for k, v in out.items() if isinstance(v, dict) and "cases" in v})


if __name__ == "__main__":
    main()
