"""The bench.py JSON line: the committed N=1 line of this round (profiles/r2_bench_line_n1.json) has
every key the contract names, and the reference arm — which runs on CPU — still prints its line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["config"]["workload"] == "C3" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["gpu_launches"] >= 3 * d["steps"]  # a device tick is four launches (fused, LWS pass, condense, namespace kernel)
    # the headline e2e is the pipelined resident tick; its latency (one tick at a time) and the churn variants are stated
    e = d["e2e"]
    assert e["ms_per_step"] > 0 and e["latency_ms_per_step"] >= e["ms_per_step"]
    assert abs(e["value"] - d["config"]["groups"] / (e["ms_per_step"] * 1e-3)) / e["value"] < 1e-6
    assert {"churn_1pct", "churn_10pct", "churn_100pct", "no_churn"} <= set(e["variants"])
    assert e["oracle_check"]["sweep_equals_oracle"] is True and e["oracle_check"]["placement_equals_spec_oracle"] is True
    assert d["oracle_check"]["sweep_equals_oracle"] is True
    assert d["roofline"]["traffic"] and d["roofline"]["traffic"] >= d["roofline"]["bytes_per_launch"]
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line_reference.json")))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["config"]["groups"] == d["config"]["groups"] and ref["e2e"]["h2d_bytes_per_step"] == 0
    assert abs(d["value"] - d["config"]["groups"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_reference_arm_prints_its_line_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C2",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "groups/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
