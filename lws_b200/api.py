"""Object model of the host side: the fields of the Kubernetes objects that the
hot path reads, named as in the reference's Go types.

This is *not* a re-implementation of the CRDs (the CRD surface stays in Go,
``api/leaderworkerset/v1`` and ``api/disaggregatedset/v1``); it is the minimal
in-memory form the record encoder consumes, so that the parity tests can be
written the way the reference's own tests are (build objects, reconcile, check
partition / replicas / conditions).

Reference: api/leaderworkerset/v1/leaderworkerset_types.go:26-99 (label and
annotation keys), :111-358 (spec), :362-395 (status).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Union

# api/leaderworkerset/v1/leaderworkerset_types.go:26-99
ExclusiveKeyAnnotationKey = "leaderworkerset.sigs.k8s.io/exclusive-topology"
SubGroupExclusiveKeyAnnotationKey = "leaderworkerset.sigs.k8s.io/subgroup-exclusive-topology"
SetNameLabelKey = "leaderworkerset.sigs.k8s.io/name"
GroupIndexLabelKey = "leaderworkerset.sigs.k8s.io/group-index"
WorkerIndexLabelKey = "leaderworkerset.sigs.k8s.io/worker-index"
SizeAnnotationKey = "leaderworkerset.sigs.k8s.io/size"
ReplicasAnnotationKey = "leaderworkerset.sigs.k8s.io/replicas"
GroupUniqueHashLabelKey = "leaderworkerset.sigs.k8s.io/group-key"
LeaderPodNameAnnotationKey = "leaderworkerset.sigs.k8s.io/leader-name"
RevisionKey = "leaderworkerset.sigs.k8s.io/template-revision-hash"
SubGroupIndexLabelKey = "leaderworkerset.sigs.k8s.io/subgroup-index"
SubGroupSizeAnnotationKey = "leaderworkerset.sigs.k8s.io/subgroup-size"
SubGroupUniqueHashLabelKey = "leaderworkerset.sigs.k8s.io/subgroup-key"
SubGroupPolicyTypeAnnotationKey = "leaderworkerset.sigs.k8s.io/subgroup-policy-type"
RecreateGroupAfterStartAnnotationKey = (
    "leaderworkerset.sigs.k8s.io/experimental-recreate-group-after-start"
)

# :323-358
RecreateGroupOnPodRestart = "RecreateGroupOnPodRestart"
RecreateGroupAfterStart = "RecreateGroupAfterStart"
DeprecatedDefaultRestartPolicy = "Default"
NoneRestartPolicy = "None"
LeaderReadyStartupPolicy = "LeaderReady"
LeaderCreatedStartupPolicy = "LeaderCreated"
SubGroupPolicyTypeLeaderWorker = "LeaderWorker"
SubGroupPolicyTypeLeaderExcluded = "LeaderExcluded"

# condition types (:397-409)
LeaderWorkerSetAvailable = "Available"
LeaderWorkerSetProgressing = "Progressing"
LeaderWorkerSetUpdateInProgress = "UpdateInProgress"

IntOrString = Union[int, str]  # k8s.io/apimachinery intstr.IntOrString


@dataclass
class OwnerReference:
    kind: str
    name: str
    uid: str
    controller: bool = True


@dataclass
class RollingUpdateConfiguration:
    partition: int = 0
    maxUnavailable: IntOrString = 1
    maxSurge: IntOrString = 0


@dataclass
class LeaderWorkerSet:
    """Defaults are those of test/wrappers/wrappers.go:226-254 (BuildLeaderWorkerSet)."""

    name: str
    namespace: str = "default"
    uid: str = ""
    replicas: int = 2
    size: int = 2
    rollingUpdate: RollingUpdateConfiguration = field(default_factory=RollingUpdateConfiguration)
    restartPolicy: str = RecreateGroupOnPodRestart
    startupPolicy: str = LeaderCreatedStartupPolicy
    subGroupSize: Optional[int] = None
    subGroupPolicyType: Optional[str] = None
    annotations: dict = field(default_factory=dict)

    def __post_init__(self):
        if not self.uid:
            self.uid = f"uid-lws-{self.namespace}-{self.name}"


@dataclass
class Pod:
    name: str
    namespace: str = "default"
    uid: str = ""
    labels: dict = field(default_factory=dict)
    annotations: dict = field(default_factory=dict)
    phase: str = ""  # "", Pending, Running, Succeeded, Failed
    readyCondition: bool = False  # status.conditions[Ready] == True
    deletionTimestamp: bool = False
    initContainerRestartCounts: list = field(default_factory=list)
    containerRestartCounts: list = field(default_factory=list)
    ownerReferences: list = field(default_factory=list)
    nodeName: str = ""
    subdomain: str = ""  # spec.subdomain (set by the pod webhook for SubdomainUniquePerReplica)

    def __post_init__(self):
        if not self.uid:
            self.uid = f"uid-pod-{self.namespace}-{self.name}"


@dataclass
class StatefulSet:
    name: str
    namespace: str = "default"
    uid: str = ""
    labels: dict = field(default_factory=dict)
    annotations: dict = field(default_factory=dict)
    replicas: int = 1  # *spec.replicas
    partition: int = 0  # spec.updateStrategy.rollingUpdate.partition
    statusReplicas: int = 0
    availableReplicas: int = 0
    currentRevision: str = ""
    updateRevision: str = ""
    ownerReferences: list = field(default_factory=list)

    def __post_init__(self):
        if not self.uid:
            self.uid = f"uid-sts-{self.namespace}-{self.name}"


@dataclass
class Node:
    name: str
    labels: dict = field(default_factory=dict)
    capacity: int = 0  # pod slots available to LWS pods
    schedulable: bool = True


def controller_of(obj) -> Optional[OwnerReference]:
    """metav1.GetControllerOf."""
    for ref in obj.ownerReferences:
        if ref.controller:
            return ref
    return None


# --------------------------------------------------------------------------- #
# DisaggregatedSet (api/disaggregatedset/v1/disaggregatedset_types.go:24-72)
# --------------------------------------------------------------------------- #
DSSetNameLabelKey = "disaggregatedset.x-k8s.io/name"
DSRoleLabelKey = "disaggregatedset.x-k8s.io/role"
DSRevisionLabelKey = "disaggregatedset.x-k8s.io/revision"
DSInitialReplicasAnnotationKey = "disaggregatedset.x-k8s.io/initial-replicas"


@dataclass
class DisaggregatedRoleSpec:
    name: str
    replicas: Optional[int] = 1  # spec.replicas (nil → 1, executor.go:223-233)
    rollingUpdate: Optional[RollingUpdateConfiguration] = None  # rolloutStrategy.rollingUpdateConfiguration


@dataclass
class DisaggregatedSet:
    name: str
    namespace: str = "default"
    uid: str = ""
    roles: list = field(default_factory=list)

    def __post_init__(self):
        if not self.uid:
            self.uid = f"uid-ds-{self.namespace}-{self.name}"


@dataclass
class ChildLWS:
    """A LeaderWorkerSet owned by a DisaggregatedSet (lws_manager.go:60-130)."""

    role: str
    revision: str
    replicas: Optional[int] = 1  # spec.replicas, None = nil
    readyReplicas: int = 0
    creationTimestamp: float = 0.0
    annotations: dict = field(default_factory=dict)
