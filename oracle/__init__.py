"""ctypes loader for the CPU oracle (``oracle/lwse_oracle*.c`` → ``liblwso.so``).

TEST INFRASTRUCTURE ONLY — see the header of ``lwse_oracle.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` import this package.  Nothing under ``lws_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from lws_b200 import records as R

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liblwso.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith(".c")]
    deps = srcs + [os.path.join(_DIR, "..", "include", "lwse.h"), os.path.join(_DIR, "Makefile")]
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(d) for d in deps)
    ):
        subprocess.run(["make", "-C", _DIR, "-B", "all"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()  # no-op when liblwso.so is newer than its sources
        _lib = C.CDLL(_LIB_PATH)
        _lib.lwso_sweep_lws.argtypes = [C.POINTER(R.LwsTables), C.c_void_p, C.c_uint32, C.c_int]
        _lib.lwso_sweep_lws.restype = C.c_int
        _lib.lwso_calculate_rolling_update_replicas.argtypes = [C.c_int32] * 4
        _lib.lwso_calculate_rolling_update_replicas.restype = C.c_int32
        _lib.lwso_scaled_value.argtypes = [C.c_int32, C.c_int, C.c_int, C.c_int]
        _lib.lwso_scaled_value.restype = C.c_int
        _lib.lwso_sub_group_index.argtypes = [C.c_int] * 3
        _lib.lwso_sub_group_index.restype = C.c_int
        for name, args, res in _OPTIONAL:
            if hasattr(_lib, name):
                fn = getattr(_lib, name)
                fn.argtypes, fn.restype = args, res
    return _lib


_OPTIONAL = [
    (
        "lwso_place",
        [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p],
        C.c_int,
    ),
    (
        "lwso_place_mt",
        [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int],
        C.c_int,
    ),
    ("lwso_apply_patch", [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32], None),
    (
        "lwso_sweep_dirty",
        [C.POINTER(R.LwsTables), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int],
        C.c_int,
    ),
    ("lwso_sweep_ds", [C.POINTER(R.DsTables)], C.c_int),
    ("lwso_sha1", [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p], C.c_int),
    (
        "lwso_ds_compute_next_step",
        [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        C.c_int,
    ),
    ("lwso_ds_batch_size", [C.c_int, C.c_int], C.c_int),
    ("lwso_ds_total_steps", [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    ("lwso_ds_next_new", [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], None),
    ("lwso_ds_next_old", [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], None),
    (
        "lwso_ds_scale_down_old",
        [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        C.c_int,
    ),
]


def sweep_lws(lws, groups, pod_state, pod_ident, nodes=None, flags=0, want_occupancy=False, threads=1):
    """Run the restatement over record tables → (lws_out, group_out, occupancy|None)."""
    n_nodes = 0 if nodes is None else len(nodes)
    lws_out = R.aligned_empty(len(lws), R.LWS_OUT)
    group_out = R.aligned_empty(len(groups), R.GROUP_OUT)
    occ = np.zeros(max(n_nodes, 1), dtype=np.uint32) if want_occupancy else None
    t = R.LwsTables(
        R.ptr(lws), len(lws), R.ptr(groups), len(groups), R.ptr(pod_state), R.ptr(pod_ident), len(pod_state),
        R.ptr(lws_out), R.ptr(group_out), R.ptr(occ), flags,
    )
    rc = lib().lwso_sweep_lws(C.byref(t), R.ptr(nodes) if n_nodes else None, n_nodes, threads)
    if rc != 0:
        raise RuntimeError(f"lwso_sweep_lws failed: {rc}")
    return lws_out, group_out, (occ[:n_nodes] if occ is not None else None)


def apply_patch(table: np.ndarray, rows: np.ndarray, values: np.ndarray) -> None:
    """table[rows] = values, in C (the CPU arm of a churn step applies its events the same way)."""
    assert table.dtype == values.dtype and rows.dtype == np.uint32 and len(rows) == len(values)
    lib().lwso_apply_patch(R.ptr(table), table.dtype.itemsize, len(table), R.ptr(rows), R.ptr(values), len(rows))


def sweep_dirty(lws, groups, pod_state, pod_ident, nodes, lws_out, group_out, dirty_groups, dirty_lws, flags=0, threads=1):
    """Reconcile only the listed group / object rows (results written in place)."""
    n_nodes = 0 if nodes is None else len(nodes)
    dg = np.ascontiguousarray(dirty_groups, dtype=np.uint32)
    dl = np.ascontiguousarray(dirty_lws, dtype=np.uint32)
    t = R.LwsTables(R.ptr(lws), len(lws), R.ptr(groups), len(groups), R.ptr(pod_state), R.ptr(pod_ident), len(pod_state),
                    R.ptr(lws_out), R.ptr(group_out), None, flags)
    rc = lib().lwso_sweep_dirty(C.byref(t), R.ptr(nodes) if n_nodes else None, n_nodes, R.ptr(dg), len(dg), R.ptr(dl), len(dl),
                                threads)
    if rc != 0:
        raise RuntimeError(f"lwso_sweep_dirty failed: {rc}")


def sweep_ds(ds, roles, revroles):
    ds_out = R.aligned_empty(len(ds), R.DS_OUT)
    role_out = R.aligned_empty(len(roles), R.DS_ROLE_OUT)
    revrole_out = R.aligned_empty(len(revroles), R.DS_REVROLE_OUT)
    t = R.DsTables(R.ptr(ds), len(ds), R.ptr(roles), len(roles), R.ptr(revroles), len(revroles),
                   R.ptr(ds_out), R.ptr(role_out), R.ptr(revrole_out))
    rc = lib().lwso_sweep_ds(C.byref(t))
    if rc != 0:
        raise RuntimeError(f"lwso_sweep_ds failed: {rc}")
    return ds_out, role_out, revrole_out


def ds_compute_next_step(initial_old, current_old, current_new, target_new, max_surge, max_unavailable):
    """planner.go:320 ComputeNextStep → (past, new) or None."""
    n = len(initial_old)
    arr = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    a = [arr(x) for x in (initial_old, current_old, current_new, target_new, max_surge, max_unavailable)]
    past, new = np.zeros(n, np.int32), np.zeros(n, np.int32)
    rc = lib().lwso_ds_compute_next_step(n, *[R.ptr(x) for x in a], R.ptr(past), R.ptr(new))
    if rc < 0:
        raise RuntimeError("lwso_ds_compute_next_step failed")
    return (past.tolist(), new.tolist()) if rc else None


def ds_batch_size(max_surge, max_unavailable):
    """planner.go:61 batchSize."""
    return int(lib().lwso_ds_batch_size(int(max_surge), int(max_unavailable)))


def ds_total_steps(initial_old, target, max_surge, max_unavailable):
    """planner.go:68 computeTotalSteps."""
    arr = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    a = [arr(x) for x in (initial_old, target, max_surge, max_unavailable)]
    return int(lib().lwso_ds_total_steps(len(initial_old), *[R.ptr(x) for x in a]))


def ds_next_new(target, current_new, total_steps):
    """planner.go:80 computeNextNewReplicas."""
    arr = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    t, c, out = arr(target), arr(current_new), np.zeros(len(target), np.int32)
    lib().lwso_ds_next_new(len(t), R.ptr(t), R.ptr(c), int(total_steps), R.ptr(out))
    return out.tolist()


def ds_next_old(initial_old, current_old, total_steps):
    """planner.go:115 computeNextOldReplicas."""
    arr = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    i, c, out = arr(initial_old), arr(current_old), np.zeros(len(initial_old), np.int32)
    lib().lwso_ds_next_old(len(i), R.ptr(i), R.ptr(c), int(total_steps), R.ptr(out))
    return out.tolist()


def ds_compute_all_steps(initial_old, target, max_surge, max_unavailable):
    """planner.go:355-385 ComputeAllSteps."""
    n = len(initial_old)
    cur_old, cur_new = list(initial_old), [0] * n
    steps = [(list(initial_old), [0] * n)]
    for _ in range(max(max(initial_old + target + [0]) * 2 + 10, 0)):
        s = ds_compute_next_step(initial_old, cur_old, cur_new, target, max_surge, max_unavailable)
        if s is None:
            break
        steps.append(s)
        cur_old, cur_new = s
    return steps


def ds_scale_down_old(replicas, order, current, target):
    """executor.go:330-398; replicas[r][i] (-1 = role missing) newest-first via order."""
    n = len(current)
    reps = np.ascontiguousarray(replicas, dtype=np.int32).reshape(-1)
    o = np.ascontiguousarray(order, dtype=np.int32)
    cur = np.ascontiguousarray(current, dtype=np.int32)  # keep alive across the call
    tgt = np.ascontiguousarray(target, dtype=np.int32)
    rc = lib().lwso_ds_scale_down_old(n, len(order), R.ptr(reps), R.ptr(o), R.ptr(cur), R.ptr(tgt), None)
    if rc != 0:
        raise RuntimeError("lwso_ds_scale_down_old failed")
    return reps.reshape(len(order), n).tolist()


def place(nodes, occupancy, n_domains, n_namespaces, reqs, threads=1):
    """Sequential statement of the placement spec (parity unpinned, build-defined); ``threads`` > 1
    solves the (independent) namespaces side by side — same rows."""
    out = R.aligned_empty(len(reqs), R.PLACE_OUT)
    occ = None if occupancy is None else np.ascontiguousarray(occupancy, dtype=np.uint32)
    rc = lib().lwso_place_mt(R.ptr(nodes), len(nodes), R.ptr(occ), n_domains, n_namespaces, R.ptr(reqs), len(reqs),
                             R.ptr(out), threads)
    if rc != 0:
        raise RuntimeError(f"lwso_place failed: {rc}")
    return out


def sha1(strings) -> np.ndarray:
    """SHA-1 of each string → (n, 20) uint8 (pkg/utils/utils.go:39-43 before hex encoding)."""
    enc = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    offsets = np.zeros(len(enc) + 1, dtype=np.uint32)
    if enc:
        offsets[1:] = np.cumsum([len(b) for b in enc])
    blob = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8)
    digests = np.zeros((len(enc), 20), dtype=np.uint8)
    rc = lib().lwso_sha1(R.ptr(blob), R.ptr(offsets), len(enc), R.ptr(digests))
    if rc != 0:
        raise RuntimeError("lwso_sha1 failed")
    return digests
