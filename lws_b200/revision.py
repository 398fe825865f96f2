"""Revision detection — SURVEY.md §8(f) rank 4: what produces the ``LWSE_LWS_UPDATED`` bit.

The reconciler decides whether a LeaderWorkerSet's template changed by building a
ControllerRevision from the current spec and comparing it with the stored one
(pkg/controllers/leaderworkerset_controller.go:138-156, :722-767;
pkg/utils/revision/revision_utils.go:52-93 NewRevision, :188-196 EqualRevision, :207-235
SetMatchesRevision, :265-300 getPatch, :334-346 hashRevision / revisionName).  All of it is byte
work on the object's JSON; this module restates it on the JSON the watch stream delivers, so the
encoder can compute the bit for a batch of objects without a Go round trip:

* ``get_patch``       the strategic-merge patch that restores spec.leaderWorkerTemplate and
                      spec.networkConfig (``$patch: replace`` on both), serialised the way Go's
                      ``encoding/json`` serialises a ``map[string]interface{}``: keys sorted,
                      compact, ``<``, ``>``, ``&``, U+2028/9 escaped, numbers as float64;
* ``hash_revision``   FNV-1 (32 bit) of the patch bytes, printed in decimal and passed through
                      ``rand.SafeEncodeString`` (k8s.io/apimachinery v0.36.1 pkg/util/rand: every
                      character c -> "bcdfghjklmnpqrstvwxz2456789"[c % 27]);
* ``revision_name``   ``<lws name, at most 220 bytes>-<hash>-<revision number>``;
* ``equal_revision``, ``apply_revision``, ``set_matches_revision`` (with its
  (uid, generation, resourceVersion) cache), ``get_updated_revision`` → the update bit.

Pinned on the reference's own table (pkg/utils/revision/revision_utils_test.go:33-223:
TestApplyRevision, the eight TestEqualRevision entries, TestSetMatchesRevision,
TestGetHighestRevision) in tests/test_revision.py.  The SafeEncodeString / FNV pair has no literal
vector in the reference; it is pinned on FNV-1's published test values and the alphabet rule.
"""
from __future__ import annotations

import copy
import json
import math
from dataclasses import dataclass, field
from typing import Optional

SAFE_ALPHANUMS = "bcdfghjklmnpqrstvwxz2456789"
SubdomainShared = "Shared"


# --------------------------------------------------------------------------- #
# Go's encoding/json for map[string]interface{}
# --------------------------------------------------------------------------- #
def _go_float(x: float) -> str:
    """encoding/json's floatEncoder: strconv.AppendFloat(x, fmt, -1, 64) with fmt 'e' when
    |x| < 1e-6 or >= 1e21 and 'f' otherwise; "e-09" is cleaned up to "e-9"."""
    if math.isnan(x) or math.isinf(x):
        raise ValueError("json: unsupported value")
    if x == 0:
        return "-0" if math.copysign(1.0, x) < 0 else "0"
    from decimal import Decimal

    sign, digits, exp = Decimal(repr(float(x))).as_tuple()  # shortest round-trip digits
    ds = "".join(map(str, digits)).rstrip("0") or "0"
    exp += len(digits) - len(ds)  # value = 0.ds... : ds x 10^exp
    neg = "-" if sign else ""
    a = abs(x)
    if a < 1e-6 or a >= 1e21:
        e10 = exp + len(ds) - 1
        mant = ds[0] + ("." + ds[1:] if len(ds) > 1 else "")
        es = f"{abs(e10):02d}"
        out = f"{neg}{mant}e{'-' if e10 < 0 else '+'}{es}"
        if e10 < 0 and es[0] == "0" and len(es) == 2:  # e-09 -> e-9
            out = f"{neg}{mant}e-{es[1]}"
        return out
    if exp >= 0:
        return neg + ds + "0" * exp
    point = len(ds) + exp
    if point > 0:
        return neg + ds[:point] + "." + ds[point:]
    return neg + "0." + "0" * (-point) + ds


def _go_string(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif o < 0x20 or ch in "<>&":
            out.append(f"\\u{o:04x}")
        elif o == 0x2028 or o == 0x2029:
            out.append(f"\\u{o:04x}")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def go_json_marshal(v) -> str:
    """json.Marshal of a value that was json.Unmarshal'ed into interface{} (maps sorted by key)."""
    if v is None:
        return "null"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, str):
        return _go_string(v)
    if isinstance(v, (int, float)):
        return _go_float(float(v))
    if isinstance(v, list):
        return "[" + ",".join(go_json_marshal(x) for x in v) + "]"
    if isinstance(v, dict):
        # Go sorts map keys as strings (byte order)
        items = sorted(v.items(), key=lambda kv: kv[0].encode())
        return "{" + ",".join(_go_string(k) + ":" + go_json_marshal(x) for k, x in items) + "}"
    raise TypeError(type(v))


# --------------------------------------------------------------------------- #
# revision_utils.go
# --------------------------------------------------------------------------- #
def get_patch(lws: dict) -> bytes:
    """:265-300 — `lws` is the LeaderWorkerSet as the API serves it (a JSON object)."""
    spec = lws.get("spec", {})
    network = copy.deepcopy(spec.get("networkConfig"))
    if network is None:  # :272-278 an object written before NetworkConfig existed
        network = {"subdomainPolicy": SubdomainShared}
    template = copy.deepcopy(spec.get("leaderWorkerTemplate", {}))
    network["$patch"] = "replace"
    template["$patch"] = "replace"
    return go_json_marshal({"spec": {"networkConfig": network, "leaderWorkerTemplate": template}}).encode()


def fnv1_32(data: bytes) -> int:
    """hash/fnv New32: FNV-1 (multiply, then xor)."""
    h = 0x811C9DC5
    for b in data:
        h = (h * 0x01000193) & 0xFFFFFFFF
        h ^= b
    return h


def safe_encode_string(s: str) -> str:
    return "".join(SAFE_ALPHANUMS[ord(c) % len(SAFE_ALPHANUMS)] for c in s)


def hash_revision(raw: bytes) -> str:
    """:334-343 with Data.Object == nil (NewRevision only sets Data.Raw)."""
    return safe_encode_string(str(fnv1_32(raw)))


def revision_name(prefix: str, hash_: str, revision_number: int) -> str:
    """:325-331"""
    if len(prefix.encode()) > 220:
        prefix = prefix.encode()[:220].decode(errors="ignore")
    return f"{prefix}-{hash_}-{revision_number}"


@dataclass
class ControllerRevision:
    name: str
    namespace: str
    labels: dict
    revision: int
    raw: bytes  # Data.Raw
    resourceVersion: str = ""
    ownerUID: str = ""

    @property
    def key(self) -> str:  # GetRevisionKey :103-108
        return self.labels.get("leaderworkerset.sigs.k8s.io/template-revision-hash", "")


def get_highest_revision(revisions: list) -> Optional[ControllerRevision]:
    """:304-320 — `max <= revision.Revision`: of equal numbers the LAST one wins."""
    best, mx = None, 0
    for r in revisions:
        if mx <= r.revision:
            mx, best = r.revision, r
    return best


def new_revision(lws: dict, revision_key: str = "", existing: Optional[list] = None) -> ControllerRevision:
    """:52-93"""
    highest = get_highest_revision(existing or [])
    number = highest.revision + 1 if highest is not None else 1
    patch = get_patch(lws)
    h = hash_revision(patch)
    md = lws.get("metadata", {})
    name = md.get("name", "")
    return ControllerRevision(
        name=revision_name(name, h, number), namespace=md.get("namespace", ""),
        labels={"leaderworkerset.sigs.k8s.io/name": name,
                "leaderworkerset.sigs.k8s.io/template-revision-hash": revision_key or h},
        revision=number, raw=patch, ownerUID=md.get("uid", ""))


def equal_revision(lhs: Optional[ControllerRevision], rhs: Optional[ControllerRevision]) -> bool:
    """:188-196 (Data.Object is nil on both sides: the semantic DeepEqual of nil, nil is true)."""
    if lhs is None or rhs is None:
        return lhs is rhs
    return lhs.raw == rhs.raw


def apply_revision(lws: dict, revision: ControllerRevision) -> dict:
    """:160-186 ApplyRevision: a strategic merge patch whose two members say `$patch: replace` —
    spec.leaderWorkerTemplate and spec.networkConfig are replaced wholesale, nothing else moves."""
    restored = copy.deepcopy(lws)
    patch = json.loads(revision.raw)
    for member, value in patch.get("spec", {}).items():
        value = dict(value)
        value.pop("$patch", None)
        restored.setdefault("spec", {})[member] = value
    return restored


@dataclass
class RevisionEqualityCache:
    """utils/lru keyed by (LWS UID, generation, revision resourceVersion) (:198-205)."""

    size: int = 1024
    _keys: dict = field(default_factory=dict)

    def get(self, key) -> bool:
        return key in self._keys

    def add(self, key) -> None:
        if len(self._keys) >= self.size:
            self._keys.pop(next(iter(self._keys)))
        self._keys[key] = True

    def __len__(self):
        return len(self._keys)


def set_matches_revision(lws: dict, proposed: ControllerRevision, existing: ControllerRevision,
                         cache: RevisionEqualityCache) -> bool:
    """:207-235 — raw bytes differ but the stored revision, re-serialised by today's encoder, gives
    today's bytes (e.g. `"creationTimestamp": null` written by an older client)."""
    md = lws.get("metadata", {})
    key = (md.get("uid", ""), md.get("generation", 0), existing.resourceVersion)
    if cache.get(key):
        return True
    try:
        latest = apply_revision(lws, existing)
        reconstructed = get_patch(latest)
    except Exception:
        return False
    if proposed.raw == reconstructed:
        cache.add(key)
        return True
    return False


def get_updated_revision(lws: dict, leader_sts_exists: bool, revision: Optional[ControllerRevision],
                         cache: RevisionEqualityCache, existing: Optional[list] = None) -> Optional[ControllerRevision]:
    """leaderworkerset_controller.go:747-767 getUpdatedRevision.  A non-None result is the new
    revision to create and means `leaderWorkerSetUpdated` (the LWSE_LWS_UPDATED bit)."""
    if not leader_sts_exists:
        return None
    current = new_revision(lws, "", existing)
    if not equal_revision(current, revision):
        if revision is not None and set_matches_revision(lws, current, revision, cache):
            return None
        return current
    return None


def updated_bits(objects: list, cache: Optional[RevisionEqualityCache] = None) -> list:
    """Batch form for the encoder: objects = [(lws json, leader sts exists, stored revision or None)]
    → [bool]: the LWSE_LWS_UPDATED bit of every object."""
    cache = cache or RevisionEqualityCache()
    return [get_updated_revision(lws, sts, rev, cache) is not None for lws, sts, rev in objects]
