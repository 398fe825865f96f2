#!/usr/bin/env python
"""Turn bench.py JSON lines (committed under profiles/) into the markdown rows BASELINE.md §4 quotes.
usage: python profiles/make_r2_tables.py profiles/r2_bench_line_n1.json [more lines ...]"""
import json
import sys


def us(ms):
    return f"{ms * 1e3:.1f} µs" if ms < 1 else f"{ms:.2f} ms"


def g(v):
    return f"{v / 1e9:.2f}·10⁹" if v >= 1e9 else f"{v / 1e6:.1f}·10⁶"


for path in sys.argv[1:]:
    d = json.loads([ln for ln in open(path).read().splitlines() if ln.startswith("{")][-1])
    cfg = d.get("config", {})
    print(f"## {path}: {cfg.get('workload')} n_gpus={d.get('n_gpus')} scaling={d.get('scaling')} impl={d.get('impl', 'ours')}")
    print(f"| device tick | {g(d['value'])} {d['unit']} | {us(d['ms_per_step'])} |")
    if "ms_sweep_only" in d:
        print(f"| sweep alone / placement alone | | {us(d['ms_sweep_only'])} / {us(d['ms_placement_only'])} |")
    e = d.get("e2e", {})
    if "ms_per_step" in e:
        print(f"| e2e (headline) | {g(e['value'])} | {us(e['ms_per_step'])} | h2d {e['h2d_bytes_per_step']} B, d2h {e['d2h_bytes_per_step']} B |")
    if "latency_ms_per_step" in e:
        print(f"| e2e one tick at a time | | {us(e['latency_ms_per_step'])} |")
    for k, v in (e.get("variants") or {}).items():
        print(f"| e2e {k} | | serial {us(v['ms_per_step'])}" + (f", two in flight {us(v['pipelined_ms_per_step'])}" if "pipelined_ms_per_step" in v else "")
              + f" | h2d {v['h2d_bytes_per_step']} B, changed {v['changed_rows_last_step']} |")
    if e.get("full_handover"):
        f = e["full_handover"]
        print(f"| e2e full handover | {g(f['value'])} | {us(f['ms_per_step'])} | h2d {f['h2d_bytes_per_step']} B, d2h {f['d2h_bytes_per_step']} B |")
    r = d.get("roofline")
    if r:
        print(f"| roofline {r.get('kernel')} | {r['achieved']:.0f} / {r['peak']:.0f} {r['unit']} = {r['frac']:.3f} | {us(r.get('ms_per_launch', 0))} | {r.get('bytes_per_launch')} B algorithmic, traffic {r.get('traffic')} |")
        for k, v in (r.get("passes") or {}).items():
            print(f"|   pass {k} | {v['gbs']:.0f} GB/s = {v['frac']:.3f} | {us(v['ms'])} | {v['bytes']} B |")
    pl = cfg.get("placement") or {}
    for k, v in (pl.get("forms") or {}).items():
        print(f"| placement form {k} | rounds {v.get('rounds')} scans {v.get('request_scans')} | {us(v['ms'])} | oracle-equal {v.get('equals_spec_oracle')} |")
    if pl.get("pair_evals"):
        print(f"| pair evals | {pl['pair_evals']['pair_evals_per_s']:.3g} /s | {us(pl['pair_evals']['ms_per_call'])} | {pl['pair_evals']['pairs_per_call']} pairs |")
    c = d.get("cpu_baseline")
    if c:
        print(f"| cpu_baseline ({c['kind']}, {c['cores']} cores) | {g(c['value'])} | | {c['sample'][:120]} |")
    print(f"| oracle_check | {d.get('oracle_check')} {e.get('oracle_check')} |")
    print(f"| clocks | {d.get('clocks')} |")
    print()
