"""A few grouped placement rounds (C3 burst) and resident ticks — the target of `ncu` captures."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from lws_b200 import synth, churn, records as R
from lws_b200.engine import Engine

t = synth.make(os.environ.get("WORKLOAD", "C3"), float(os.environ.get("SCALE", "1.0")))
reqs = t.place_requests()
occ = R.occupancy_of(t.pod_ident, len(t.nodes))
e = Engine(0)
e.upload_nodes(t.nodes, t.n_domains)
dev = torch.device("cuda:0")
up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
d_reqs, d_occ = up(reqs), up(occ)
po = torch.empty(len(reqs) * R.PLACE_OUT.itemsize, dtype=torch.uint8, device=dev)
for _ in range(int(os.environ.get("REPS", "4"))):
    e.place_grouped_device(d_reqs, len(reqs), d_occ, t.n_namespaces, po)
torch.cuda.synchronize()
if os.environ.get("TICKS"):
    e.resident_load(t.lws, t.groups, t.pod_state, t.pod_ident)
    e.resident_place_load(reqs, t.n_namespaces)
    flags = t.flags | R.TICK_PLACE
    e.resident_tick(e.make_tick((), flags))
    plan = churn.make_plan(t, reqs, e.resident_place_outputs(), 0.01, 0.01, n_sets=2, seed=11)
    ap = churn.ArenaPlan(e, plan, flags)
    for k in range(int(os.environ["TICKS"])):
        e.resident_tick(ap.ticks[k % 2])
    torch.cuda.synchronize()
print("done")
