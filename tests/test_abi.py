"""CPU-side checks of the C-ABI library: it loads, exports every symbol that
include/lwse.h declares, and its record sizes match the numpy mirrors.  No
compute call is made (there is no GPU here and no CPU fallback to call)."""
import ctypes as C
import os
import re

import pytest

from lws_b200 import build, engine
from lws_b200 import records as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return engine.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lwse.h")).read()
    return sorted(set(re.findall(r"LWSE_API\s+[\w\s\*]+?\b(lwse_\w+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"liblwse.so does not export {n}"
    assert sorted(engine.SYMBOLS) == names, "engine.py binding table drifted from lwse.h"


def test_abi_version_and_helpers(lib):
    assert lib.lwse_abi_version() == R.ABI_VERSION
    assert b"no CPU fallback" in lib.lwse_strerror(-2)
    for s in ["", "rev-1", "default/test-sample-0", "x" * 300]:
        b = s.encode()
        assert lib.lwse_hash64(b, len(b)) == R.hash64(s)
    shards = [lib.lwse_shard_of(R.hash64(f"uid-{i}"), 8) for i in range(4000)]
    assert set(shards) == set(range(8))
    assert max(shards.count(k) for k in range(8)) < 4000 / 8 * 1.25
    assert lib.lwse_shard_of(12345, 1) == 0


def test_record_sizes_match_header():
    """sizeof() of every struct in lwse.h, via a gcc-compiled probe."""
    import subprocess
    import tempfile

    structs = {
        "lwse_lws_rec": R.LWS_REC, "lwse_group_rec": R.GROUP_REC, "lwse_pod_ident": R.POD_IDENT,
        "lwse_node_rec": R.NODE_REC, "lwse_lws_out": R.LWS_OUT, "lwse_group_out": R.GROUP_OUT,
        "lwse_place_req": R.PLACE_REQ, "lwse_place_out": R.PLACE_OUT, "lwse_ds_rec": R.DS_REC,
        "lwse_ds_role_rec": R.DS_ROLE_REC, "lwse_ds_revrole_rec": R.DS_REVROLE_REC,
        "lwse_ds_out": R.DS_OUT, "lwse_ds_role_out": R.DS_ROLE_OUT,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/lwse.h"', "int main(){"]
    for s, dt in structs.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in dt.names:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines += ['printf("lwse_lws_tables %zu\\n", sizeof(lwse_lws_tables));',
              'printf("lwse_ds_tables %zu\\n", sizeof(lwse_ds_tables));',
              'printf("lwse_patch_seg %zu\\n", sizeof(lwse_patch_seg));',
              'printf("lwse_tick %zu\\n", sizeof(lwse_tick));',
              'printf("lwse_tick.place_rounds %zu\\n", offsetof(lwse_tick, place_rounds));',
              'printf("lwse_pod_state %zu\\n", sizeof(lwse_pod_state));', "return 0;}"]
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for s, dt in structs.items():
        assert int(got[s]) == dt.itemsize, s
        for f in dt.names:
            assert int(got[f"{s}.{f}"]) == dt.fields[f][1], f"{s}.{f}"
    assert int(got["lwse_lws_tables"]) == C.sizeof(R.LwsTables)
    assert int(got["lwse_ds_tables"]) == C.sizeof(R.DsTables)
    assert int(got["lwse_patch_seg"]) == C.sizeof(R.PatchSeg)
    assert int(got["lwse_tick"]) == C.sizeof(R.Tick)
    assert int(got["lwse_tick.place_rounds"]) == R.Tick.place_rounds.offset
    assert int(got["lwse_pod_state"]) == R.POD_STATE.itemsize == 1


def test_create_without_gpu_fails_loudly(lib):
    """No device → LWSE_ERR_NO_DEVICE, never a silent CPU path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    cfg = R.Config(R.ABI_VERSION, 0, 0, 0)
    assert lib.lwse_create(C.byref(cfg), C.byref(h)) == -2
    with pytest.raises(engine.LwseError):
        engine.Engine(0)
