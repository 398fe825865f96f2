"""Batched label pass of ``PodWebhook.Default`` (pkg/webhooks/pod_webhook.go:83-157) — SURVEY §8(f)
rank 3, the part of the admission path that is arithmetic: group / worker / subgroup indices and the
SHA-1 group and subgroup keys of a whole batch of pods.  Names are parsed on the host
(``GetParentNameAndOrdinal``, ``strconv.Atoi``); every key of the batch is hashed in ONE call of the
hasher the caller passes — ``Engine.group_keys_host`` (the CUDA SHA-1 kernel, ``lwse_group_keys_host``)
in production; the tests also run it over the CPU oracle's SHA-1.  Affinity terms, the PodGroup metadata
and the ``LWS_*`` / ``TPU_*`` environment strings (``:122-155, :159-176``) are object and string
construction and stay in the Go webhook.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

from . import api
from .encoder import atoi, get_parent_name_and_ordinal

SubdomainPolicyAnnotationKey = "leaderworkerset.sigs.k8s.io/subdomainPolicy"  # leaderworkerset_types.go:94
SubdomainUniquePerReplica = "UniquePerReplica"  # :263
SubGroupPolicyTypeLeaderExcluded = "LeaderExcluded"  # :241


def _go_div(a: int, b: int) -> int:
    """Go's integer division: truncates toward zero (Python's // floors)."""
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def _go_mod(a: int, b: int) -> int:
    return a - b * _go_div(a, b)


def get_sub_group_index(pod_count: int, sub_group_size: int, worker_index: int) -> str:
    """pod_webhook.go:249-255 with Go's truncating division: worker 0 under the "leader is extra"
    rule gives (0 - 1) / sg = 0 for sg > 1 but -1 for sg == 1, exactly as the Go expression does.
    sub_group_size == 0 is a division by zero in Go (a panic the caller turns into a per-pod
    error): ZeroDivisionError here."""
    if sub_group_size == 0:
        raise ZeroDivisionError("subGroupSize is 0")
    if _go_mod(pod_count - 1, sub_group_size) == 0:
        return str(_go_div(worker_index - 1, sub_group_size))
    return str(_go_div(worker_index, sub_group_size))


def default_labels_batch(pods: Sequence[api.Pod], sha1_batch: Callable[[list], "np.ndarray"]) -> list[Optional[str]]:
    """Apply the label part of ``Default`` to every pod in place; → one error string (or None) per pod,
    worded like the reference's.  ``pod.subdomain`` is set where the reference sets ``pod.Spec.Subdomain``."""
    errors: list[Optional[str]] = [None] * len(pods)
    want: list[tuple[int, str, str]] = []  # (pod row, label key, string to hash)
    for i, pod in enumerate(pods):
        if api.SetNameLabelKey not in pod.labels:  # :88-91 not part of a leaderworkerset
            continue
        size = pod.annotations.get(api.SizeAnnotationKey)
        if size is None:
            errors[i] = f"size annotation is unexpectedly missing for pod {pod.name}"
            continue
        pod_count = atoi(size)
        if pod_count is None:
            errors[i] = f'strconv.Atoi: parsing "{size}": invalid syntax'
            continue
        if pod.labels.get(api.WorkerIndexLabelKey) == "0":  # podutils.LeaderPod
            if api.GroupIndexLabelKey not in pod.labels:
                _, group_index = get_parent_name_and_ordinal(pod.name)
                if group_index == -1:
                    errors[i] = f"parsing pod ordinal for pod {pod.name}"
                    continue
                pod.labels[api.GroupIndexLabelKey] = str(group_index)
            if pod.annotations.get(SubdomainPolicyAnnotationKey) == SubdomainUniquePerReplica:
                pod.subdomain = pod.name
            if api.GroupUniqueHashLabelKey not in pod.labels:
                want.append((i, api.GroupUniqueHashLabelKey, f"{pod.namespace}/{pod.name}"))  # genGroupUniqueKey :180
            if (api.SubGroupSizeAnnotationKey in pod.annotations and pod.labels.get(api.SubGroupIndexLabelKey, "") == ""
                    and pod.annotations.get(api.SubGroupPolicyTypeAnnotationKey, "") != SubGroupPolicyTypeLeaderExcluded):
                pod.labels[api.SubGroupIndexLabelKey] = "0"  # the leader always lands on subgroup 0
                want.append((i, api.SubGroupUniqueHashLabelKey, f"{pod.name}/0"))
        else:
            _, worker_index = get_parent_name_and_ordinal(pod.name)
            if worker_index == -1:
                errors[i] = f"parsing pod ordinal for pod {pod.name}"
                continue
            pod.labels[api.WorkerIndexLabelKey] = str(worker_index)
            sub_size = pod.annotations.get(api.SubGroupSizeAnnotationKey)
            if sub_size is not None and pod.labels.get(api.SubGroupIndexLabelKey, "") == "":
                sub_size_int = atoi(sub_size)
                if sub_size_int is None:
                    errors[i] = f'strconv.Atoi: parsing "{sub_size}": invalid syntax'
                    continue
                leader_name = pod.annotations.get(api.LeaderPodNameAnnotationKey, "")
                if sub_size_int == 0:  # Go panics (integer divide by zero): one pod's admission fails, not the batch
                    errors[i] = "runtime error: integer divide by zero"
                    continue
                idx = get_sub_group_index(pod_count, sub_size_int, worker_index)
                pod.labels[api.SubGroupIndexLabelKey] = idx
                want.append((i, api.SubGroupUniqueHashLabelKey, f"{leader_name}/{idx}"))
    if want:
        digests = sha1_batch([s for _, _, s in want])  # (n, 20) uint8 — one batch for all keys
        for (i, key, _), d in zip(want, digests):
            pods[i].labels[key] = bytes(d).hex()  # Sha1Hash: hex of the digest (utils.go:39-43)
    return errors


# --------------------------------------------------------------------------- #
# The rest of PodWebhook.Default on the admission request's JSON (pods as dicts:
# {"metadata": {...}, "spec": {...}}) — SURVEY §8(f) rank 3.
# --------------------------------------------------------------------------- #
LeaderRequestsTPUsAnnotationKey = "leaderworkerset.sigs.k8s.io/leader-requests-tpus"
TpuResourceName = "google.com/tpu"
TpuWorkerHostNames, TpuProcessAddresses = "TPU_WORKER_HOSTNAMES", "TPU_PROCESS_ADDRESSES"
TpuProcessPortName, TpuWorkerId, TpuName = "TPU_PROCESS_PORT", "TPU_WORKER_ID", "TPU_NAME"
TpuProcessDefaultPort = 8476
LwsLeaderAddress, LwsGroupSize, LwsWorkerIndex = "LWS_LEADER_ADDRESS", "LWS_GROUP_SIZE", "LWS_WORKER_INDEX"


def _md(pod: dict) -> dict:
    return pod.setdefault("metadata", {})


def _labels(pod: dict) -> dict:
    return _md(pod).get("labels") or {}


def _annotations(pod: dict) -> dict:
    return _md(pod).get("annotations") or {}


def exclusive_affinity_applied(pod: dict, topology_key: str) -> bool:
    """pod_webhook.go:230-247: idempotence is keyed on the topology key only."""
    aff = (pod.get("spec") or {}).get("affinity")
    if not aff or aff.get("podAffinity") is None or aff.get("podAntiAffinity") is None:
        return False
    req = "requiredDuringSchedulingIgnoredDuringExecution"
    has = any(t.get("topologyKey") == topology_key for t in aff["podAffinity"].get(req) or [])
    has_anti = any(t.get("topologyKey") == topology_key for t in aff["podAntiAffinity"].get(req) or [])
    return has and has_anti


def set_exclusive_affinities(pod: dict, group_unique_key: str, topology_key: str, pod_affinity_key: str) -> None:
    """pod_webhook.go:185-227: affinity `key In [group key]` + anti-affinity `key Exists ∧ NotIn [group key]`,
    both on the topology key, no namespaces set (→ the pod's own namespace)."""
    if exclusive_affinity_applied(pod, topology_key):
        return
    aff = pod.setdefault("spec", {}).setdefault("affinity", {})
    req = "requiredDuringSchedulingIgnoredDuringExecution"
    if aff.get("podAffinity") is None:
        aff["podAffinity"] = {}
    if aff.get("podAntiAffinity") is None:
        aff["podAntiAffinity"] = {}
    pa, paa = aff["podAffinity"], aff["podAntiAffinity"]
    pa[req] = list(pa.get(req) or []) + [{
        "labelSelector": {"matchExpressions": [{"key": pod_affinity_key, "operator": "In", "values": [group_unique_key]}]},
        "topologyKey": topology_key}]
    paa[req] = list(paa.get(req) or []) + [{
        "labelSelector": {"matchExpressions": [{"key": pod_affinity_key, "operator": "Exists"},
                                               {"key": pod_affinity_key, "operator": "NotIn", "values": [group_unique_key]}]},
        "topologyKey": topology_key}]


def get_env_var_value_if_in_container(c: dict, name: str):
    """pkg/utils/pod/pod_utils.go:95-102"""
    for env in c.get("env") or []:
        if env.get("name") == name:
            return True, env.get("value", "")
    return False, ""


def add_env_vars_if_not_exists(c: dict, first: dict, *rest: dict) -> None:
    """pod_utils.go:104-129: the given variables first, in order, then the container's own that are not among them."""
    names, new = set(), []
    for env in (first,) + rest:
        new.append(dict(env))
        names.add(env["name"])
    for env in c.get("env") or []:
        if env.get("name") not in names:
            new.append(env)
            names.add(env.get("name"))
    c["env"] = new


def add_lws_variables(pod: dict) -> Optional[str]:
    """pod_utils.go:131-180 AddLWSVariables → error string or None."""
    labels, ann = _labels(pod), _annotations(pod)
    ref = f'{_md(pod).get("namespace", "")}/{_md(pod).get("name", "")}'
    if api.SetNameLabelKey not in labels:
        return f"Failure constructing environment variables, no name label found for pod {ref}"
    if api.GroupIndexLabelKey not in labels:
        return f"Failure constructing environment variables, no group index label found for pod {ref}"
    spec = pod.setdefault("spec", {})
    leader = {"name": LwsLeaderAddress,
              "value": f'{labels[api.SetNameLabelKey]}-{labels[api.GroupIndexLabelKey]}.{spec.get("subdomain", "")}.{_md(pod).get("namespace", "")}'}
    if api.SizeAnnotationKey not in ann:
        return f"Failure constructing environment variables, no size annotation found for pod {ref}"
    size = {"name": LwsGroupSize, "value": ann[api.SizeAnnotationKey]}
    if api.WorkerIndexLabelKey not in labels:
        return f"Failure constructing environment variables, no worker index label found for pod {ref}"
    widx = {"name": LwsWorkerIndex, "value": labels[api.WorkerIndexLabelKey]}
    for c in spec.get("containers") or []:
        add_env_vars_if_not_exists(c, leader, size, widx)
    for c in spec.get("initContainers") or []:
        add_env_vars_if_not_exists(c, leader, size, widx)
    return None


def _quantity_is_zero(q) -> bool:
    """resource.Quantity.IsZero for the plain forms a TPU count takes ("4", 4, "0")."""
    try:
        return float(str(q)) == 0.0
    except ValueError:
        return False


def num_tpus_requested(c: dict) -> int:
    """pkg/utils/accelerators/tpu.go:46-58: limits first, then requests."""
    res = c.get("resources") or {}
    for part in ("limits", "requests"):
        q = (res.get(part) or {}).get(TpuResourceName)
        if q is not None and not _quantity_is_zero(q):
            return int(float(str(q)))
    return 0


def containers_requesting_tpus(spec: dict) -> list:
    """tpu.go:72-86: containers first, then init containers."""
    return [c for c in (spec.get("containers") or []) if num_tpus_requested(c)] + \
           [c for c in (spec.get("initContainers") or []) if num_tpus_requested(c)]


def pod_requests_tpus(spec: dict) -> bool:
    return bool(containers_requesting_tpus(spec))


def _append_tpu_env(container: dict, hostnames, worker_id, tpu_name, addresses, port, port_in_container) -> None:
    env = container.setdefault("env", [])
    if env is None:
        env = container["env"] = []
    env += [{"name": TpuWorkerHostNames, "value": ",".join(hostnames)}, {"name": TpuWorkerId, "value": str(worker_id)},
            {"name": TpuName, "value": tpu_name}, {"name": TpuProcessAddresses, "value": ",".join(addresses)}]
    if not port_in_container:
        env.append({"name": TpuProcessPortName, "value": port})


def add_tpu_variables_subgroup(pod: dict) -> Optional[str]:
    """tpu.go:99-200 addTPUVariablesSubGroup → error string or None."""
    spec = pod.get("spec") or {}
    cs = containers_requesting_tpus(spec)
    if not cs:
        return None
    container = cs[0]
    for env in container.get("env") or []:
        if env.get("name") in (TpuWorkerHostNames, TpuWorkerId):
            return None
    labels, ann = _labels(pod), _annotations(pod)
    name = _md(pod).get("name", "")
    sub = spec.get("subdomain", "")
    vals = []
    for s in (ann.get(api.SubGroupSizeAnnotationKey), labels.get(api.SubGroupIndexLabelKey), labels.get(api.WorkerIndexLabelKey)):
        v = atoi(s)
        if v is None:
            return f'strconv.Atoi: parsing "{s if s is not None else ""}": invalid syntax'
        vals.append(v)
    sg_size, sg_index, worker_index = vals
    if sg_size == 0:
        return "runtime error: integer divide by zero"
    leader_tpus = ann.get(LeaderRequestsTPUsAnnotationKey) == "true"
    tpu_worker_id = _go_mod(worker_index, sg_size)
    if not leader_tpus:
        tpu_worker_id = _go_mod(worker_index - 1, sg_size)
    in_c, port = get_env_var_value_if_in_container(container, TpuProcessPortName)
    if not in_c:
        port = str(TpuProcessDefaultPort)
    start, end = sg_size * sg_index + 1, sg_size * (sg_index + 1)
    hostnames, addresses = [], []
    leader_name = name
    if labels.get(api.WorkerIndexLabelKey) == "0":
        hostnames.append(f"{leader_name}.{sub}")
        addresses.append(f"{leader_name}.{sub}:{port}")
        end -= 1
    else:
        leader_name, _ = get_parent_name_and_ordinal(name)
        if leader_name == "":
            return f"parsing parent name from pod {name}"
        if leader_tpus and sg_index == 0:
            end -= 1
            hostnames.append(f"{leader_name}.{sub}")
            addresses.append(f"{leader_name}.{sub}:{port}")
        elif leader_tpus:
            start -= 1
            end -= 1
    for i in range(start, end + 1):
        hostnames.append(f"{leader_name}-{i}.{sub}")
        addresses.append(f"{leader_name}-{i}.{sub}:{port}")
    _append_tpu_env(container, hostnames, tpu_worker_id, leader_name, addresses, port, in_c)
    return None


def add_tpu_variables(pod: dict, size: int) -> Optional[str]:
    """tpu.go:203-300 AddTPUVariables → error string or None."""
    spec = pod.get("spec") or {}
    labels, ann = _labels(pod), _annotations(pod)
    if api.SubGroupSizeAnnotationKey in ann:
        return add_tpu_variables_subgroup(pod)
    cs = containers_requesting_tpus(spec)
    n = len(cs)
    if n == 0:
        return None
    for env in cs[0].get("env") or []:
        if env.get("name") in (TpuWorkerHostNames, TpuWorkerId):
            return None
    name, sub = _md(pod).get("name", ""), spec.get("subdomain", "")
    leader_tpus = ann.get(LeaderRequestsTPUsAnnotationKey) == "true"
    if labels.get(api.WorkerIndexLabelKey) == "0":
        leader_name, pod_worker_index = name, 0
    else:
        leader_name, pod_worker_index = get_parent_name_and_ordinal(name)
        if leader_name == "":
            return f"parsing parent name from pod {name}"
        if not leader_tpus:
            pod_worker_index -= 1
    ports = []
    for i, c in enumerate(cs):
        found, val = get_env_var_value_if_in_container(c, TpuProcessPortName)
        ports.append(val if found else str(TpuProcessDefaultPort + i))
    hostnames, addresses = [], []
    if leader_tpus or labels.get(api.WorkerIndexLabelKey) == "0":
        host = f"{leader_name}.{sub}"
        for i in range(n):
            hostnames.append(host)
            addresses.append(f"{host}:{ports[i]}")
    for i in range(1, size):
        host = f"{leader_name}-{i}.{sub}"
        for j in range(n):
            hostnames.append(host)
            addresses.append(f"{host}:{ports[j]}")
    for i, c in enumerate(cs):
        in_c, _ = get_env_var_value_if_in_container(c, TpuProcessPortName)
        _append_tpu_env(c, hostnames, pod_worker_index * n + i, leader_name, addresses, ports[i], in_c)
    return None


KubeGroupNameAnnotationKey = "scheduling.k8s.io/group-name"  # volcano.sh/apis v1.12.1 scheduling/v1beta1


def volcano_inject_pod_group_metadata(pod: dict) -> Optional[str]:
    """VolcanoProvider.InjectPodGroupMetadata (pkg/schedulerprovider/volcano_provider.go:103-109): the pod's
    PodGroup is ``GetPodGroupName(lws, groupIndex, revision)`` = "<lws>-<group index>-<revision key>"
    (interface.go:35, :49-51).  Go writes into ``pod.Annotations`` without a nil check: a pod without
    annotations panics there — reported as an error string here."""
    md = _md(pod)
    labels = md.get("labels") or {}
    if md.get("annotations") is None:
        return "assignment to entry in nil map"
    md["annotations"][KubeGroupNameAnnotationKey] = "%s-%s-%s" % (
        labels.get(api.SetNameLabelKey, ""), labels.get(api.GroupIndexLabelKey, ""), labels.get(api.RevisionKey, ""))
    return None


def default_batch(pods: Sequence[dict], sha1_batch: Callable[[list], "np.ndarray"],
                  inject_pod_group_metadata: Optional[Callable[[dict], Optional[str]]] = None) -> list[Optional[str]]:
    """The whole of ``PodWebhook.Default`` (pod_webhook.go:83-178) over a batch of admission requests
    (pods as JSON objects, mutated in place) → one error string (or None) per pod.  Every SHA-1 of the
    batch — group keys and subgroup keys — is hashed in ONE call of ``sha1_batch``
    (``Engine.group_keys_host``: the CUDA kernel); the affinity terms need the digests, so the pass is
    split in two around that call exactly where the reference calls ``genGroupUniqueKey``."""
    errors: list[Optional[str]] = [None] * len(pods)
    want: list[tuple[int, str, str, Optional[str], Optional[str]]] = []  # pod, label, string, exclusive key annotation, affinity label
    stage2: list[int] = []
    counts: dict[int, int] = {}
    for i, pod in enumerate(pods):
        md = _md(pod)
        labels = md.get("labels")
        if not labels or api.SetNameLabelKey not in labels:  # :88-91
            continue
        ann = md.get("annotations") or {}
        size = ann.get(api.SizeAnnotationKey)
        if size is None:
            errors[i] = f'size annotation is unexpectedly missing for pod {md.get("name", "")}'
            continue
        pod_count = atoi(size)
        if pod_count is None:
            errors[i] = f'strconv.Atoi: parsing "{size}": invalid syntax'
            continue
        counts[i] = pod_count
        name, ns = md.get("name", ""), md.get("namespace", "")
        if labels.get(api.WorkerIndexLabelKey) == "0":
            if api.GroupIndexLabelKey not in labels:
                _, gi = get_parent_name_and_ordinal(name)
                if gi == -1:
                    errors[i] = f"parsing pod ordinal for pod {name}"
                    continue
                labels[api.GroupIndexLabelKey] = str(gi)
            if ann.get(SubdomainPolicyAnnotationKey) == SubdomainUniquePerReplica:
                pod.setdefault("spec", {})["subdomain"] = name
            ep = ann.get(api.ExclusiveKeyAnnotationKey)
            if api.GroupUniqueHashLabelKey not in labels:
                want.append((i, api.GroupUniqueHashLabelKey, f"{ns}/{name}", ep, api.GroupUniqueHashLabelKey))
            elif ep is not None:
                set_exclusive_affinities(pod, labels[api.GroupUniqueHashLabelKey], ep, api.GroupUniqueHashLabelKey)
            if (api.SubGroupSizeAnnotationKey in ann and labels.get(api.SubGroupIndexLabelKey, "") == ""
                    and ann.get(api.SubGroupPolicyTypeAnnotationKey, "") != SubGroupPolicyTypeLeaderExcluded):
                labels[api.SubGroupIndexLabelKey] = "0"
                want.append((i, api.SubGroupUniqueHashLabelKey, f"{name}/0", ann.get(api.SubGroupExclusiveKeyAnnotationKey),
                             api.SubGroupUniqueHashLabelKey))
        else:
            _, wi = get_parent_name_and_ordinal(name)
            if wi == -1:
                errors[i] = f"parsing pod ordinal for pod {name}"
                continue
            labels[api.WorkerIndexLabelKey] = str(wi)
            sub_size = ann.get(api.SubGroupSizeAnnotationKey)
            if sub_size is not None and labels.get(api.SubGroupIndexLabelKey, "") == "":
                ssi = atoi(sub_size)
                if ssi is None:
                    errors[i] = f'strconv.Atoi: parsing "{sub_size}": invalid syntax'
                    continue
                if ssi == 0:
                    errors[i] = "runtime error: integer divide by zero"
                    continue
                idx = get_sub_group_index(pod_count, ssi, wi)
                labels[api.SubGroupIndexLabelKey] = idx
                want.append((i, api.SubGroupUniqueHashLabelKey, f"{ann.get(api.LeaderPodNameAnnotationKey, '')}/{idx}",
                             ann.get(api.SubGroupExclusiveKeyAnnotationKey), api.SubGroupUniqueHashLabelKey))
        stage2.append(i)
    if want:
        digests = sha1_batch([s for _, _, s, _, _ in want])
        for (i, key, _, ep, aff_key), d in zip(want, digests):
            if errors[i] is not None:
                continue
            hexd = bytes(d).hex()
            _md(pods[i])["labels"][key] = hexd
            if ep is not None:
                set_exclusive_affinities(pods[i], hexd, ep, aff_key)
    for i in stage2:
        if errors[i] is not None:
            continue
        pod = pods[i]
        if inject_pod_group_metadata is not None:  # SchedulerProvider.InjectPodGroupMetadata :159-164
            err = inject_pod_group_metadata(pod)
            if err:
                errors[i] = err
                continue
        if pod_requests_tpus(pod.get("spec") or {}):  # :167-171
            err = add_tpu_variables(pod, counts[i])
            if err:
                errors[i] = err
                continue
        err = add_lws_variables(pod)  # :173
        if err:
            errors[i] = err
    return errors
