#!/usr/bin/env python
"""Extracts the DisaggregatedSet planner golden vectors from the reference's own
test file into tests/golden/planner_sequences.json.

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/extract_planner_vectors.py /root/reference

Source: pkg/controllers/disaggregatedset/planner_test.go:106-505
(TestComputeAllSteps_ExactSequence — 22 exact step sequences of
ComputeAllSteps, planner.go:355-385).
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
path = os.path.join(ref, "pkg/controllers/disaggregatedset/planner_test.go")
text = open(path).read()
start = text.index("func TestComputeAllSteps_ExactSequence")
end = text.index("for _, tc := range testCases", start)
body = text[start:end]
cases = []
for chunk in body.split("name:")[1:]:
    name = re.search(r'"([^"]+)"', chunk).group(1)
    src = [int(re.search(r"sourceRole0:\s*(\d+)", chunk).group(1)), int(re.search(r"sourceRole1:\s*(\d+)", chunk).group(1))]
    tgt = [int(re.search(r"targetRole0:\s*(\d+)", chunk).group(1)), int(re.search(r"targetRole1:\s*(\d+)", chunk).group(1))]
    cfg_line = re.search(r"config:\s*(.*)", chunk).group(1)
    if "DefaultRollingUpdateConfig" in cfg_line:
        cfg = [[1, 0], [1, 0]]
    else:
        cfg = []
        for m in re.finditer(r"\{([^{}]*)\}", cfg_line[cfg_line.index("{") + 1:]):
            s = m.group(1)
            ms = re.search(r"MaxSurge:\s*(\d+)", s)
            mu = re.search(r"MaxUnavailable:\s*(\d+)", s)
            cfg.append([int(ms.group(1)) if ms else 0, int(mu.group(1)) if mu else 0])
    steps = [
        [[int(a), int(b)], [int(c), int(d)]]
        for a, b, c, d in re.findall(r"step\(\[\]int\{(\d+), (\d+)\}, \[\]int\{(\d+), (\d+)\}\)", chunk)
    ]
    cases.append({"name": name, "source": src, "target": tgt, "config": cfg, "steps": steps})
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "planner_sequences.json")
json.dump({"source": "pkg/controllers/disaggregatedset/planner_test.go:106-505", "cases": cases}, open(out, "w"), indent=1)
print(len(cases), "cases;", [len(c["steps"]) for c in cases])
