#!/bin/bash
# Register-budget experiment (profiles/r2_summary.md): runs bench.py on variant libraries built here beforehand with
#   nvcc <build.NVCC_FLAGS> [-DLWSE_NS_MAXNREG=48] [-DLWSE_FUSED_MAXNREG=72|64] -Ilws_b200/csrc -Iinclude \
#        -o lws_b200/variants/liblwse_<name>.so lws_b200/csrc/*.cu
# usage (on the GPU box, from the repo root): profiles/microbench/variants_run.sh v0 v1 ...
cp lws_b200/liblwse.so /tmp/liblwse_keep.so
for v in "$@"; do
  cp lws_b200/variants/liblwse_$v.so lws_b200/liblwse.so
  timeout 200 python bench.py --steps 100 --warmup 5 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/var_{v}.json"))
    e = d["e2e"]
    print(v, "tick %.2f us" % (d["ms_per_step"] * 1e3), "sweep %.2f" % (d["ms_sweep_only"] * 1e3), "place %.2f" % (d["ms_placement_only"] * 1e3),
          "fused %.2f" % (d["roofline"]["ms_per_launch"] * 1e3), "e2e %.2f us" % (e["ms_per_step"] * 1e3), "lone %.2f" % (e["latency_ms_per_step"] * 1e3),
          d.get("oracle_check"))
except Exception as ex:
    print(v, "FAILED", ex)
PY
done
cp /tmp/liblwse_keep.so lws_b200/liblwse.so
