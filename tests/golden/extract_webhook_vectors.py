#!/usr/bin/env python
"""Extract the reference's pod-webhook table tests into JSON fixtures (tests/golden/webhook_vectors.json).

    python tests/golden/extract_webhook_vectors.py    # needs /root/reference (this container only)

Sources (kubernetes-sigs/lws @ 1d9204a2), parsed with the Go composite-literal parser of
extract_apply_config_vectors.py:
  pkg/webhooks/pod_webhook_test.go:66-169    TestSetExclusiveAffinities
  pkg/webhooks/pod_webhook_test.go:171-270   TestExclusiveAffinityApplied
  pkg/utils/pod/pod_utils_test.go:103-186    TestAddLWSVariables
  pkg/utils/accelerators/tpu_test.go:34-293  TestAddTPUVariables, :295-346 TestAddTPUVariablesSkip,
                                             :348-588 TestAddTPUVariablesSubGroup
  test/wrappers/wrappers.go                  the Make*PodSpec / MakeContainerWithTPU literals the entries use
Builder calls (wrappers.MakeX(args)) stay symbolic ({"$chain": [[name, args...]]}); the test resolves
them against the "wrappers" section.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import extract_apply_config_vectors as G  # noqa: E402

REF = G.REF
G.CONSTS.update({
    "LeaderRequestsTPUsAnnotationKey": "leaderworkerset.sigs.k8s.io/leader-requests-tpus",
    "TpuWorkerHostNames": "TPU_WORKER_HOSTNAMES", "TpuProcessAddresses": "TPU_PROCESS_ADDRESSES",
    "TpuProcessPortName": "TPU_PROCESS_PORT", "TpuWorkerId": "TPU_WORKER_ID", "TpuName": "TPU_NAME",
    "TpuResourceName": "google.com/tpu", "tpuResourceName": "google.com/tpu",
    "leaderworkerset.SubGroupPolicyTypeAnnotationKey": "leaderworkerset.sigs.k8s.io/subgroup-policy-type",
})


def wrapper_literals(src):
    """func MakeX(...) T { return T{...} } → {name: {"params": [...], "value": parsed literal}}"""
    out = {}
    for m in re.finditer(r"func (Make\w+)\(([^)]*)\) [\w.*]+ \{\n\treturn ", src):
        name, params = m.group(1), m.group(2)
        p = G.Parser(G.tokenize(src[m.end():]))
        try:
            val = p.value()
        except SyntaxError:
            continue
        out[name] = {"params": [q.strip().split(" ")[0] for q in params.split(",") if q.strip()], "value": val}
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    rd = lambda p: open(os.path.join(REF, p)).read()
    out = {}
    wh = rd("pkg/webhooks/pod_webhook_test.go")
    out["set_exclusive_affinities"] = {"source": "pkg/webhooks/pod_webhook_test.go:66-169",
                                       "cases": G.table_entries(wh, "func TestSetExclusiveAffinities(")}
    out["exclusive_affinity_applied"] = {"source": "pkg/webhooks/pod_webhook_test.go:171-270",
                                         "cases": G.table_entries(wh, "func TestExclusiveAffinityApplied(")}
    out["add_lws_variables"] = {"source": "pkg/utils/pod/pod_utils_test.go:103-186",
                                "cases": G.table_entries(rd("pkg/utils/pod/pod_utils_test.go"), "func TestAddLWSVariables(")}
    tpu = rd("pkg/utils/accelerators/tpu_test.go")
    out["add_tpu_variables"] = {"source": "pkg/utils/accelerators/tpu_test.go:34-293",
                                "cases": G.table_entries(tpu, "func TestAddTPUVariables(")}
    out["add_tpu_variables_skip"] = {"source": "pkg/utils/accelerators/tpu_test.go:295-346",
                                     "cases": G.table_entries(tpu, "func TestAddTPUVariablesSkip(")}
    out["add_tpu_variables_subgroup"] = {"source": "pkg/utils/accelerators/tpu_test.go:348-588",
                                         "cases": G.table_entries(tpu, "func TestAddTPUVariablesSubGroup(")}
    out["get_containers_requesting_tpus"] = {"source": "pkg/utils/accelerators/tpu_test.go:590-656",
                                             "cases": G.table_entries(tpu, "func TestGetContainersRequestingTPUs(")}
    out["get_container_requesting_tpus"] = {"source": "pkg/utils/accelerators/tpu_test.go:658-704",
                                            "cases": G.table_entries(tpu, "func TestGetContainerRequestingTPUs(")}
    out["pod_requests_tpus"] = {"source": "pkg/utils/accelerators/tpu_test.go:706-756",
                                "cases": G.table_entries(tpu, "func TestPodRequestsTPUs(")}
    out["get_env_var_if_in_container"] = {"source": "pkg/utils/pod/pod_utils_test.go:188-256",
                                          "cases": G.table_entries(rd("pkg/utils/pod/pod_utils_test.go"), "func TestGetEnvVarIfInContainer(")}
    out["wrappers"] = {"source": "test/wrappers/wrappers.go", "functions": wrapper_literals(rd("test/wrappers/wrappers.go"))}
    path = os.path.join(HERE, "webhook_vectors.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, {k: len(v["cases"]) for k, v in out.items() if "cases" in v}, sorted(out["wrappers"]["functions"]))


if __name__ == "__main__":
    main()
