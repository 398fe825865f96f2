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
