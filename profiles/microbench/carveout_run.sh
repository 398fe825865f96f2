#!/bin/bash
# Shared-memory carveout experiment: bench.py under LWSE_SMEM_CARVEOUT = each argument ("d" = unset)
for c in "$@"; do
  if [ "$c" = d ]; then unset LWSE_SMEM_CARVEOUT; else export LWSE_SMEM_CARVEOUT=$c; fi
  timeout 200 python bench.py --steps 100 --warmup 5 > gpurun_out/carve_$c.json 2> gpurun_out/carve_$c.err
  python - "$c" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/carve_{v}.json"))
    e = d["e2e"]
    print("carveout", v, "tick %.2f us" % (d["ms_per_step"] * 1e3), "sweep %.2f" % (d["ms_sweep_only"] * 1e3), "place %.2f" % (d["ms_placement_only"] * 1e3),
          "fused %.2f" % (d["roofline"]["ms_per_launch"] * 1e3), "e2e %.2f us" % (e["ms_per_step"] * 1e3), "lone %.2f" % (e["latency_ms_per_step"] * 1e3),
          d.get("oracle_check"))
except Exception as ex:
    print(v, "FAILED", ex)
PY
done
