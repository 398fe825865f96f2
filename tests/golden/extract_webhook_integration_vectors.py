#!/usr/bin/env python
"""Extract the pod-webhook integration table into a JSON fixture (tests/golden/webhook_integration_vectors.json).

    python tests/golden/extract_webhook_integration_vectors.py    # needs /root/reference (this container only)

Source (kubernetes-sigs/lws @ 1d9204a2): test/integration/webhooks/pod_test.go:265-868 — the entries of the
"Defaulting" table that check TPU / LWS env vars, the injected subdomain and the exclusive-placement terms of a
pod that went through PodWebhook.Default.  (:68-263, the label entries, are in tests/test_webhook_batch.py.)

Per entry:  name; the pod literal of makePod (builder calls such as wrappers.MakeLeaderPodSpecWithTPUResource()
stay symbolic, as in webhook_vectors.json); the statements makePod runs on the pod before returning it
(SetExclusiveAffinities / appended affinity terms, parsed as literals); and the checks of checkExpectedPod as a
list of {fn, args, want}: the testutils validators it calls (test/testutils/util.go:475-600), with the polarity
the entry expects.  The validators themselves are restated in tests/test_webhook_integration.py.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import extract_apply_config_vectors as G  # noqa: E402
import extract_webhook_vectors as W  # noqa: E402,F401  (extends G.CONSTS)

G.CONSTS.update({
    "leaderworkerset.SubdomainPolicyAnnotationKey": "leaderworkerset.sigs.k8s.io/subdomainPolicy",
    "leaderworkerset.SubGroupExclusiveKeyAnnotationKey": "leaderworkerset.sigs.k8s.io/subgroup-exclusive-topology",
    "leaderworkerset.SubGroupUniqueHashLabelKey": "leaderworkerset.sigs.k8s.io/subgroup-key",
    "leaderworkerset.GroupUniqueHashLabelKey": "leaderworkerset.sigs.k8s.io/group-key",
    "leaderworkerset.ExclusiveKeyAnnotationKey": "leaderworkerset.sigs.k8s.io/exclusive-topology",
    "acceleratorutils.LeaderRequestsTPUsAnnotationKey": "leaderworkerset.sigs.k8s.io/leader-requests-tpus",
    "metav1.LabelSelectorOpIn": "In", "metav1.LabelSelectorOpNotIn": "NotIn", "metav1.LabelSelectorOpExists": "Exists",
})

SRC = "test/integration/webhooks/pod_test.go"
FIRST, LAST = 265, 868


def literal_after(text, pos):
    p = G.Parser(G.tokenize(text[pos:]))
    return p.value()


def const(name):
    name = name.strip()
    if name.startswith('"'):
        return json.loads(name)
    for k in (name, name.split(".")[-1]):
        if k in G.CONSTS:
            return G.CONSTS[k]
    raise KeyError(name)


def checks_of(body):
    out = []
    for m in re.finditer(r"if (!?)testutils\.(HasTPUEnvVarsPopulated|HasLWSEnvVarsPopulated)\(got\)", body):
        # `if X(got) { return error }` expects false; `if !X(got) { return error }` expects true
        out.append({"fn": m.group(2), "args": [], "want": m.group(1) == "!"})
    for m in re.finditer(r'testutils\.CheckTPUContainerHasCorrectEnvVars\(got, ("(?:[^"\\]|\\.)*")\)', body):
        out.append({"fn": "CheckTPUContainerHasCorrectEnvVars", "args": [json.loads(m.group(1))], "want": None})
    for m in re.finditer(r"testutils\.ValidatePodExclusivePlacementTerms\(got, ([\w.]+), ([\w.]+)\)\)\.To\(gomega\.Be(True|False)\(\)\)", body):
        out.append({"fn": "ValidatePodExclusivePlacementTerms", "args": [const(m.group(1)), const(m.group(2))],
                    "want": m.group(3) == "True"})
    if "testutils.IsContainerFirstEnvVarLWSLeaderAddress(got)" in body:
        out.append({"fn": "IsContainerFirstEnvVarLWSLeaderAddress", "args": [], "want": None})
    # env vars built in the closure and checked with CheckContainerHasCorrectEnvVar
    envs = {}
    for m in re.finditer(r"(\w+) := corev1\.EnvVar\{", body):
        lit = literal_after(body, m.start() + len(m.group(1)) + 4)
        envs[m.group(1)] = lit
    for m in re.finditer(r"testutils\.CheckContainerHasCorrectEnvVar\(got, (\w+)\)", body):
        out.append({"fn": "CheckContainerHasCorrectEnvVar", "args": [envs[m.group(1)]], "want": None})
    if "existing pod affinity terms are unexpectedly overridden" in body:
        out.append({"fn": "FirstAffinityTermsKeepKey", "args": ["key"], "want": None})
    return out


def main():
    if not os.path.isdir(G.REF):
        sys.exit("needs /root/reference")
    lines = open(os.path.join(G.REF, SRC)).read().split("\n")
    text = "\n".join(lines[FIRST - 1:LAST])
    # fmt.Sprintf("…%s", expected.ObjectMeta.Namespace) → the namespace placeholder the test fills in
    text = re.sub(r'fmt\.Sprintf\(("(?:[^"\\]|\\.)*"), expected\.ObjectMeta\.Namespace\)',
                  lambda m: json.dumps(json.loads(m.group(1)).replace("%s", "$NAMESPACE")), text)
    text = text.replace("ns.Name", '"$NAMESPACE"')
    text = text.replace("string(leaderworkerset.SubdomainUniquePerReplica)", '"UniquePerReplica"')
    cases = []
    starts = [m.start() for m in re.finditer(r"ginkgo\.Entry\(", text)]
    for a, b in zip(starts, starts[1:] + [len(text)]):
        entry = text[a:b]
        name = json.loads(re.match(r'ginkgo\.Entry\(("(?:[^"\\]|\\.)*")', entry).group(1))
        mk = entry.index("makePod:")
        ck = entry.index("checkExpectedPod:")
        make, check = entry[mk:ck], entry[ck:]
        pm = re.search(r"(return |pod := &)corev1\.Pod\{", make)
        pod = literal_after(make, pm.end() - len("corev1.Pod{"))
        pre = []
        m = re.search(r'webhooks\.SetExclusiveAffinities\(pod, ("[^"]*"), ("[^"]*"), ([\w.]+)\)', make)
        if m:
            pre.append({"op": "SetExclusiveAffinities", "args": [json.loads(m.group(1)), json.loads(m.group(2)), const(m.group(3))]})
        for kind in ("PodAffinity", "PodAntiAffinity"):
            m = re.search(r"pod\.Spec\.Affinity\.%s\.RequiredDuringSchedulingIgnoredDuringExecution = append\([^,]+,\s*corev1\.PodAffinityTerm\{" % kind, make)
            if m:
                pre.append({"op": "AppendRequiredTerm", "args": [kind, literal_after(make, m.end() - len("corev1.PodAffinityTerm{"))]})
        cases.append({"name": name, "pod": pod, "pre": pre, "checks": checks_of(check)})
    out = {"source": f"{SRC}:{FIRST}-{LAST}", "cases": cases}
    path = os.path.join(HERE, "webhook_integration_vectors.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(f"{path}: {len(cases)} entries; checks per entry: {[len(c['checks']) for c in cases]}")


if __name__ == "__main__":
    main()
