"""SHA-1 group-key throughput (W1): every pod admission of C3 hashes "<ns>/<pod name>" — 6.4 M keys per
full re-admission.  Device-resident blob, CUDA events on the engine's stream; prints keys/s and bytes/s."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from lws_b200.engine import Engine  # noqa: E402

n = int(os.environ.get("N", "6400000"))
e = Engine(0)
dev = torch.device("cuda:0")
# "ns-XXX/lws-XXXXXX-G" / "...-G-W": 24-34 bytes, like real pod names
rng = np.random.default_rng(1)
lens = rng.integers(24, 35, size=n).astype(np.int64)
offsets = np.zeros(n + 1, dtype=np.uint32)
offsets[1:] = np.cumsum(lens)
blob = rng.integers(97, 123, size=int(offsets[-1]) + 4, dtype=np.uint8)
d_blob, d_off = torch.from_numpy(blob).to(dev), torch.from_numpy(offsets.view(np.int32)).to(dev)
d_dig = torch.empty(n * 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.ExternalStream(e.stream, device=dev)
for _ in range(3):
    e.group_keys_device(d_blob, d_off, n, d_dig, stream=e.stream)
torch.cuda.synchronize()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(reps):
        e.group_keys_device(d_blob, d_off, n, d_dig, stream=e.stream)
    e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# spot check against hashlib
import hashlib
dig = d_dig.cpu().numpy().reshape(n, 20)
for i in (0, 1, n // 2, n - 1):
    assert bytes(dig[i]) == hashlib.sha1(blob[offsets[i]:offsets[i + 1]].tobytes()).digest()
t0 = time.perf_counter()
for i in range(20000):
    hashlib.sha1(blob[offsets[i]:offsets[i + 1]].tobytes()).digest()
cpu = 20000 / (time.perf_counter() - t0)
print(json.dumps({"keys": n, "ms": ms, "keys_per_s": n / (ms * 1e-3), "bytes_in_per_s": float(offsets[-1]) / (ms * 1e-3),
                  "bytes_out_per_s": n * 20 / (ms * 1e-3), "mean_key_bytes": float(lens.mean()),
                  "hashlib_one_core_keys_per_s": cpu}))
