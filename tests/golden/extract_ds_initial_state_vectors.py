#!/usr/bin/env python
"""Extract TestComputeInitialReplicaState into tests/golden/ds_initial_state_vectors.json.

    python tests/golden/extract_ds_initial_state_vectors.py    # needs /root/reference (this container only)

Source (kubernetes-sigs/lws @ 1d9204a2): pkg/utils/disaggregatedset/utils_test.go:177-412 — eight sub-tests,
each a `lwsList := []leaderworkersetv1.LeaderWorkerSet{…}` literal followed by
`assert.Equal(t, N, state[role], …)` lines.  Per sub-test: name, the list as plain data (labels, annotations,
spec.replicas; nil stays absent) and the expected per-role totals.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import extract_apply_config_vectors as G  # noqa: E402

G.CONSTS.update({
    "disaggregatedsetv1.RoleLabelKey": "disaggregatedset.x-k8s.io/role",
    "disaggregatedsetv1.InitialReplicasAnnotationKey": "disaggregatedset.x-k8s.io/initial-replicas",
    "testUtilsRolePrefill": "prefill", "testUtilsRoleDecode": "decode",
})
SRC = "pkg/utils/disaggregatedset/utils_test.go"
ROLE = {"testUtilsRolePrefill": "prefill", "testUtilsRoleDecode": "decode"}


def main():
    if not os.path.isdir(G.REF):
        sys.exit("needs /root/reference")
    text = open(os.path.join(G.REF, SRC)).read()
    types = open(os.path.join(G.REF, "api/disaggregatedset/v1/disaggregatedset_types.go")).read()  # :31, :38
    for k in ("disaggregatedsetv1.RoleLabelKey", "disaggregatedsetv1.InitialReplicasAnnotationKey"):
        assert '"%s"' % G.CONSTS[k] in types, k
    body = text[text.index("func TestComputeInitialReplicaState("):]
    starts = [m.start() for m in re.finditer(r"\tt\.Run\(", body)]
    cases = []
    for a, b in zip(starts, starts[1:] + [len(body)]):
        sub = body[a:b]
        name = json.loads(re.match(r'\tt\.Run\(("(?:[^"\\]|\\.)*")', sub).group(1))
        m = re.search(r"lwsList := (\[\]leaderworkersetv1\.LeaderWorkerSet\{)", sub)
        lst = G.Parser(G.tokenize(sub[m.start(1):])).value()
        want = {}
        for am in re.finditer(r"assert\.Equal\(t, (\d+), state\[(\w+)\]", sub):
            want[ROLE[am.group(2)]] = int(am.group(1))
        cases.append({"name": name, "lwsList": lst, "want": want})
    out = {"source": f"{SRC}:177-412", "cases": cases}
    path = os.path.join(HERE, "ds_initial_state_vectors.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, len(cases), [c["want"] for c in cases])


if __name__ == "__main__":
    main()
