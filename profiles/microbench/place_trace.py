import ctypes as C, numpy as np, sys
sys.path.insert(0,'/root/repo')
from lws_b200 import synth, encoder, records as R
from lws_b200.engine import Engine, lib
t = synth.make("C3", 1.0)
reqs = encoder.encode_place_requests(t.lws, t.groups)
occ = R.occupancy_of(t.pod_ident, len(t.nodes))
e = Engine(0); e.upload_nodes(t.nodes, t.n_domains)
for i in range(5):
    out, rounds = e.place_host(reqs, occ, 1)
fn = lib().lwse_debug_place_trace; fn.argtypes=[C.c_void_p, C.c_void_p]; fn.restype=C.c_int
buf = np.zeros(16, dtype=np.uint64)
print("rc", fn(e._h, buf.ctypes.data), "rounds", rounds, "reqs", len(reqs), "unpinned", int((reqs["leader_node"]==R.NONE).sum()))
t0 = int(buf[0])
print([ (k, (int(v)-t0)/1e3) for k,v in enumerate(buf) if v])
