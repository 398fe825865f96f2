"""Kernel / copy timelines (CUPTI via torch.profiler) of (1) the grouped placement round alone, (2) resident
ticks with 1 % churn, (3) one full-handover call.  Prints stream, start offset and duration per activity."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from lws_b200 import synth, churn, records as R
from lws_b200.engine import Engine
from torch.profiler import profile, ProfilerActivity

t = synth.make(os.environ.get("WORKLOAD", "C3"), float(os.environ.get("SCALE", "1.0")))
reqs = t.place_requests()
occ = R.occupancy_of(t.pod_ident, len(t.nodes))
e = Engine(0)
e.upload_nodes(t.nodes, t.n_domains)
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d_reqs, d_occ = up(reqs), up(occ)
po = torch.empty(len(reqs) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)


def show(prof, title, limit=60):
    path = os.path.join(tempfile.gettempdir(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [x for x in json.load(open(path))["traceEvents"] if x.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda x: x["ts"])
    print(f"--- {title}: {len(ev)} activities")
    if not ev:
        return
    t0 = ev[0]["ts"]
    for x in ev[:limit]:
        print(f"{x['ts'] - t0:9.2f} us  +{x['dur']:7.2f}  s{x['args'].get('stream')}  {x['name'][:60]}")


for _ in range(10):
    e.place_grouped_device(d_reqs, len(reqs), d_occ, t.n_namespaces, po)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        e.place_grouped_device(d_reqs, len(reqs), d_occ, t.n_namespaces, po)
    torch.cuda.synchronize()
show(prof, "grouped placement x3")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        e.place_grouped_device(d_reqs, len(reqs), d_occ, t.n_namespaces, po, flags=R.SWEEP_PLACE_SCAN)
    torch.cuda.synchronize()
show(prof, "scan placement x2")

e.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
e.resident_place_load(reqs, t.n_namespaces)
flags = t.flags | R.TICK_PLACE
e.resident_tick(e.make_tick((), flags))
base = e.resident_place_outputs()
plan = churn.make_plan(t, reqs, base, 0.01, 0.01, n_sets=4, seed=11)
ap = churn.ArenaPlan(e, plan, flags)
for k in range(8):
    e.resident_tick(ap.ticks[k % 4])
import time
t0 = time.perf_counter()
for k in range(200):
    e.resident_tick(ap.ticks[k % 4])
print("tick wall us", (time.perf_counter() - t0) / 200 * 1e6)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for k in range(3):
        e.resident_tick(ap.ticks[k % 4])
    torch.cuda.synchronize()
show(prof, "resident ticks x3 (1 % churn)")

# two ticks in flight (submit / wait): graph replay after the first few
e.resident_tick_submit(ap.ticks[0])
for k in range(1, 12):
    e.resident_tick_submit(ap.ticks[k % 4])
    e.resident_tick_wait()
e.resident_tick_wait()
t0 = time.perf_counter()
e.resident_tick_submit(ap.ticks[0])
for k in range(1, 200):
    e.resident_tick_submit(ap.ticks[k % 4])
    e.resident_tick_wait()
e.resident_tick_wait()
print("pipelined tick wall us", (time.perf_counter() - t0) / 200 * 1e6)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e.resident_tick_submit(ap.ticks[0])
    for k in range(1, 4):
        e.resident_tick_submit(ap.ticks[k % 4])
        e.resident_tick_wait()
    e.resident_tick_wait()
    torch.cuda.synchronize()
show(prof, "pipelined resident ticks x4 (1 % churn, two in flight)", limit=80)

keep = []
def pinned(a):
    ten = torch.empty(max(a.nbytes, 16), dtype=torch.uint8).pin_memory()
    v = ten.numpy()[: a.nbytes].view(a.dtype); v[...] = a; keep.append(ten); return v
h = [pinned(x) for x in (t.lws, t.groups, t.pod_state, t.pod_ident, reqs, occ)]
lo, go, pout = pinned(R.aligned_empty(len(t.lws), R.LWS_OUT)), pinned(R.aligned_empty(len(t.groups), R.GROUP_OUT)), pinned(R.aligned_empty(len(reqs), R.PLACE_OUT))
for _ in range(3):
    e.reconcile_host(h[0], h[1], h[2], h[3], h[4], h[5], t.n_namespaces, flags=t.flags, out=(lo, go), place_out=pout)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e.reconcile_host(h[0], h[1], h[2], h[3], h[4], h[5], t.n_namespaces, flags=t.flags, out=(lo, go), place_out=pout)
    torch.cuda.synchronize()
show(prof, "full handover x1")
