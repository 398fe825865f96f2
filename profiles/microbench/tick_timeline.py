"""Kernel timeline of a few reconcile ticks (CUPTI via torch.profiler): do the placement round
and the sweep kernels really overlap?  Prints stream, start offset and duration per kernel."""
import json, os, sys, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from lws_b200 import synth, encoder, records as R
from lws_b200.engine import Engine

t = synth.make("C3", float(os.environ.get("SCALE", "1.0")))
reqs = encoder.encode_place_requests(t.lws, t.groups)
occ = R.occupancy_of(t.pod_ident, len(t.nodes))
e = Engine(0)
e.upload_nodes(t.nodes, t.n_domains)
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d = [up(t.lws), up(t.groups), up(t.pod_state), up(t.pod_ident)]
lo = torch.empty(len(t.lws) * R.LWS_OUT.itemsize, dtype=torch.uint8, device=dev)
go = torch.empty(len(t.groups) * R.GROUP_OUT.itemsize, dtype=torch.uint8, device=dev)
po = torch.empty(len(reqs) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
d_occ, d_reqs = torch.from_numpy(occ.view(np.int32)).to(dev), up(reqs)
tab = e.device_tables(d[0], len(t.lws), d[1], len(t.groups), d[2], d[3], len(t.pod_state), lo, go, None, flags=t.flags)
for _ in range(20):
    e.reconcile_device(tab, d_reqs, len(reqs), d_occ, 1, po)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        e.reconcile_device(tab, d_reqs, len(reqs), d_occ, 1, po)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "tick_trace.json")
prof.export_chrome_trace(path)
ev = [x for x in json.load(open(path))["traceEvents"] if x.get("cat") == "kernel"]
ev.sort(key=lambda x: x["ts"])
t0 = ev[0]["ts"]
for x in ev:
    print(f"{x['ts'] - t0:9.2f} us  +{x['dur']:7.2f}  stream {x['args'].get('stream')}  {x['name'][:40]}")
