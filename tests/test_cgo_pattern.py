"""The Go shim's calling pattern, exercised from C (no Go toolchain in this image): C-malloc'd pageable
tables, FOUR threads sharing ONE engine handle — the stateless sweep, placement rounds, SHA-1 group
keys and the resident tick all at once (tests/host_c/cgo_pattern_check.c).  Every result file is
compared with the oracle.  This is also the regression test for the staging-buffer race the
round-1 advisor found in lwse_place_host / lwse_group_keys_host."""
import hashlib
import os
import subprocess
import tempfile

import numpy as np
import pytest

from lws_b200 import records as R
from lws_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_c", "cgo_pattern_check.c")


def build(tmp):
    exe = os.path.join(tmp, "cgo_pattern_check")
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-o", exe, SRC, "-L" + os.path.join(ROOT, "lws_b200"), "-llwse",
                    "-lpthread", "-Wl,-rpath," + os.path.join(ROOT, "lws_b200")], check=True)
    return exe


def test_c_driver_compiles_and_links_against_the_abi():
    with tempfile.TemporaryDirectory() as tmp:
        assert os.path.exists(build(tmp))


@pytest.mark.gpu
def test_four_threads_share_one_handle():
    import oracle

    p = synth.profile("fuzz", 1.0)
    p.n_namespaces = 3
    t = synth.make(p, seed=61)
    reqs = t.place_requests()
    occ = R.occupancy_of(t.pod_ident, len(t.nodes))
    keys = [f"ns-{i % 5}/lws-{i}-{i % 9}".encode() for i in range(20000)]
    off = np.zeros(len(keys) + 1, np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        for name, arr in (("lws", t.lws), ("groups", t.groups), ("pod_state", t.pod_state), ("pod_ident", t.pod_ident),
                          ("nodes", t.nodes), ("reqs", reqs), ("occ", occ), ("key_offsets", off)):
            np.ascontiguousarray(arr).tofile(os.path.join(tmp, name + ".bin"))
        open(os.path.join(tmp, "keys.bin"), "wb").write(b"".join(keys) + b"\0\0\0\0")
        open(os.path.join(tmp, "meta.txt"), "w").write(f"{t.n_domains} {t.n_namespaces} {t.flags}\n")
        r = subprocess.run([exe, tmp, "6"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        want_lo, want_go, _ = oracle.sweep_lws(t.lws, t.groups, t.pod_state, t.pod_ident, t.nodes, flags=t.flags)
        rd = lambda n, dt: np.fromfile(os.path.join(tmp, n), dtype=dt)
        assert rd("lws_out.bin", R.LWS_OUT).tobytes() == want_lo.tobytes()
        assert rd("group_out.bin", R.GROUP_OUT).tobytes() == want_go.tobytes()
        assert rd("tick_group_out.bin", R.GROUP_OUT).tobytes() == want_go.tobytes()  # the patches cancel out
        want_po = oracle.place(t.nodes, occ, t.n_domains, t.n_namespaces, reqs)
        assert rd("place_out.bin", R.PLACE_OUT).tobytes() == want_po.tobytes()
        dig = rd("digests.bin", np.uint8).reshape(-1, 20)
        assert all(bytes(dig[i]) == hashlib.sha1(keys[i]).digest() for i in range(0, len(keys), 97))
