"""The C++ host mirror (lws_b200/csrc/host/lws_host.hpp): its encoder must produce the
same bytes as the Python encoder; on a GPU its reconciler facades reproduce the
reference's unit KATs through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

from lws_b200 import api, build, encoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cpp", "host_check.cc")
EXE = os.path.join(ROOT, "tests", "host_cpp", "host_check")


@pytest.fixture(scope="module")
def exe():
    lib = build.build()
    deps = [SRC, os.path.join(ROOT, "lws_b200", "csrc", "host", "lws_host.hpp"), os.path.join(ROOT, "include", "lwse.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps + [lib]):
        libdir = os.path.dirname(lib)
        subprocess.run(["g++", "-std=c++17", "-O1", "-o", EXE, SRC, f"-L{libdir}", "-llwse", f"-Wl,-rpath,{libdir}"],
                       check=True, capture_output=True)
    return EXE


def scenario():
    nodes = [api.Node(f"node-{i}", {"rack": "r0" if i < 2 else "r1"}, 8) for i in range(3)] + [api.Node("node-3", {}, 8)]
    a = api.LeaderWorkerSet("alpha", uid="uid-alpha", replicas=3, size=2,
                            rollingUpdate=api.RollingUpdateConfiguration(maxSurge="50%"),
                            annotations={api.ExclusiveKeyAnnotationKey: "rack"})
    asts = api.StatefulSet("alpha", replicas=3, partition=2, annotations={api.ReplicasAnnotationKey: "3"})
    items = [encoder.LwsItem(a, "rev-2", False, asts)]
    pods, stss = [], []

    def labels(lws, g, w, rev):
        return {api.SetNameLabelKey: lws, api.WorkerIndexLabelKey: str(w), api.GroupIndexLabelKey: str(g), api.RevisionKey: rev}

    for g in range(3):
        ln, rev = f"alpha-{g}", "rev-2" if g == 2 else "rev-1"
        pods.append(api.Pod(ln, uid=f"uid-{ln}", labels=labels("alpha", g, 0, rev), phase="Running", readyCondition=g != 1,
                            nodeName="node-0" if g == 0 else ("node-3" if g == 1 else "")))
        stss.append(api.StatefulSet(ln, uid=f"uid-sts-{ln}",
                                    labels={api.SetNameLabelKey: "alpha", api.GroupIndexLabelKey: str(g), api.RevisionKey: rev},
                                    replicas=1, availableReplicas=1, ownerReferences=[api.OwnerReference("Pod", ln, f"uid-{ln}")]))
        pods.append(api.Pod(f"{ln}-1", uid=f"uid-{ln}-1", labels=labels("alpha", g, 1, rev),
                            phase="Pending" if g == 2 else "Running", containerRestartCounts=[2 if g == 0 else 0],
                            deletionTimestamp=g == 1,
                            ownerReferences=[api.OwnerReference("StatefulSet", ln, "uid-stale" if g == 1 else f"uid-sts-{ln}")],
                            nodeName="node-1" if g == 0 else ""))
    b = api.LeaderWorkerSet("beta", uid="uid-beta", replicas=2, size=1, restartPolicy="None",
                            startupPolicy=api.LeaderReadyStartupPolicy,
                            rollingUpdate=api.RollingUpdateConfiguration(maxUnavailable="oops"))
    items.append(encoder.LwsItem(b, "rev-9", True, None))
    pods.append(api.Pod("beta-0", uid="uid-beta-0", labels=labels("beta", 0, 0, "rev-9"), phase="Running", readyCondition=True))
    return items, encoder.Cluster(pods=pods, statefulsets=stss, nodes=nodes)


def test_cpp_encoder_matches_python_encoder(exe):
    out = subprocess.run([exe, "encode"], check=True, capture_output=True, text=True).stdout
    got = dict(line.split(" ", 1) for line in out.strip().splitlines())
    items, cluster = scenario()
    t = encoder.encode_lws(items, cluster, "rack")
    for name, arr in (("lws", t.lws), ("groups", t.groups), ("pod_state", t.pod_state), ("pod_ident", t.pod_ident),
                      ("nodes", t.nodes)):
        assert got[name].strip() == np.ascontiguousarray(arr).tobytes().hex(), f"{name} table differs"
    assert int(got["n_domains"]) == t.n_domains


@pytest.mark.gpu
def test_cpp_facades_reproduce_reference_kats(exe):
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ok" in r.stdout
