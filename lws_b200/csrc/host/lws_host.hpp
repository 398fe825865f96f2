// lws_host.hpp — C++ host side above the C ABI: the object model the reference's
// reconcilers read, the objects → records encoder, and reconciler facades whose
// methods carry the reference's names and return what the reference returns.
//
// The reference is compiled Go; with no Go toolchain in this image the host mirror
// is C++ (header-only, C++17).  It contains no arithmetic of the path: every
// decision comes out of liblwse.so (CUDA).  Python's lws_b200/encoder.py is the
// same encoder used by the test-suite; tests/test_host_cpp.py checks that both
// produce byte-identical tables.
//
// Reference: api/leaderworkerset/v1/leaderworkerset_types.go:26-99 (keys),
// pkg/controllers/leaderworkerset_controller.go:280,414,576 and
// pkg/controllers/pod_controller.go:204 (facade method names).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <regex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/lwse.h"

namespace lws {

// api/leaderworkerset/v1/leaderworkerset_types.go:26-99
inline constexpr const char* ExclusiveKeyAnnotationKey = "leaderworkerset.sigs.k8s.io/exclusive-topology";
inline constexpr const char* SetNameLabelKey = "leaderworkerset.sigs.k8s.io/name";
inline constexpr const char* GroupIndexLabelKey = "leaderworkerset.sigs.k8s.io/group-index";
inline constexpr const char* WorkerIndexLabelKey = "leaderworkerset.sigs.k8s.io/worker-index";
inline constexpr const char* ReplicasAnnotationKey = "leaderworkerset.sigs.k8s.io/replicas";
inline constexpr const char* LeaderPodNameAnnotationKey = "leaderworkerset.sigs.k8s.io/leader-name";
inline constexpr const char* RevisionKey = "leaderworkerset.sigs.k8s.io/template-revision-hash";
inline constexpr const char* RecreateGroupAfterStartAnnotationKey =
    "leaderworkerset.sigs.k8s.io/experimental-recreate-group-after-start";
inline constexpr const char* RecreateGroupOnPodRestart = "RecreateGroupOnPodRestart";
inline constexpr const char* RecreateGroupAfterStart = "RecreateGroupAfterStart";
inline constexpr const char* LeaderReadyStartupPolicy = "LeaderReady";

using Labels = std::map<std::string, std::string>;

struct IntOrString {  // k8s.io/apimachinery intstr.IntOrString
  bool is_string = false;
  int32_t int_val = 0;
  std::string str_val;
  static IntOrString FromInt(int32_t v) { return {false, v, {}}; }
  static IntOrString FromString(std::string s) { return {true, 0, std::move(s)}; }
};

struct OwnerReference {
  std::string kind, name, uid;
  bool controller = true;
};

struct RollingUpdateConfiguration {
  int32_t partition = 0;
  IntOrString maxUnavailable = IntOrString::FromInt(1);
  IntOrString maxSurge = IntOrString::FromInt(0);
};

struct LeaderWorkerSet {  // defaults: test/wrappers/wrappers.go:226-254
  std::string name, ns = "default", uid;
  int32_t replicas = 2, size = 2;
  RollingUpdateConfiguration rollingUpdate;
  std::string restartPolicy = RecreateGroupOnPodRestart, startupPolicy = "LeaderCreated";
  std::optional<int32_t> subGroupSize;
  std::string subGroupPolicyType;
  Labels annotations;
};

struct Pod {
  std::string name, ns = "default", uid;
  Labels labels, annotations;
  std::string phase;  // "", Pending, Running, …
  bool readyCondition = false, deletionTimestamp = false;
  std::vector<int32_t> initContainerRestartCounts, containerRestartCounts;
  std::vector<OwnerReference> ownerReferences;
  std::string nodeName;
};

struct StatefulSet {
  std::string name, ns = "default", uid;
  Labels labels, annotations;
  int32_t replicas = 1, partition = 0, availableReplicas = 0;
  std::string currentRevision, updateRevision;
  std::vector<OwnerReference> ownerReferences;
};

struct Node {
  std::string name;
  Labels labels;
  uint16_t capacity = 0;
  bool schedulable = true;
};

// ---------------------------------------------------------------------------
inline uint64_t hash64(const std::string& s) { return lwse_hash64(s.data(), s.size()); }
inline uint32_t hash32(const std::string& s) {
  const uint64_t h = hash64(s);
  return (uint32_t)(h ^ (h >> 32));
}
inline const std::string& get(const Labels& m, const char* k) {
  static const std::string empty;
  auto it = m.find(k);
  return it == m.end() ? empty : it->second;
}
inline bool has(const Labels& m, const char* k) { return m.find(k) != m.end(); }

// strconv.Atoi
inline std::optional<long long> atoi_go(const std::string& s) {
  static const std::regex re("[+-]?[0-9]+");
  if (!std::regex_match(s, re)) return std::nullopt;
  try {
    return std::stoll(s);
  } catch (...) {
    return std::nullopt;
  }
}

// pkg/utils/statefulset/statefulset_utils.go:33-45
inline std::pair<std::string, int> GetParentNameAndOrdinal(const std::string& name) {
  static const std::regex re("([\\s\\S]*)-([0-9]+)$");
  std::smatch m;
  if (!std::regex_match(name, m, re)) return {"", -1};
  long long v = 0;
  try {
    v = std::stoll(m[2].str());
  } catch (...) {
    return {m[1].str(), -1};
  }
  if (v > 0x7FFFFFFFLL) return {m[1].str(), -1};
  return {m[1].str(), (int)v};
}

inline const OwnerReference* controllerOf(const std::vector<OwnerReference>& refs) {
  for (const auto& r : refs)
    if (r.controller) return &r;
  return nullptr;
}

struct LwsItem {  // what Reconcile looks at for one LeaderWorkerSet
  LeaderWorkerSet lws;
  std::string revisionKey;
  bool lwsUpdated = false;
  std::optional<StatefulSet> leaderSts;
};

struct Cluster {
  std::vector<Pod> pods;
  std::vector<StatefulSet> statefulsets;
  std::vector<Node> nodes;
};

struct Tables {
  std::vector<lwse_lws_rec> lws;
  std::vector<lwse_group_rec> groups;
  std::vector<lwse_pod_state> pod_state;
  std::vector<lwse_pod_ident> pod_ident;
  std::vector<lwse_node_rec> nodes;
  uint32_t n_domains = 0;
  std::vector<std::string> domain_values;
  std::vector<std::vector<std::string>> group_pod_names;
};

// Objects → records.  Same rules as lws_b200/encoder.py:encode_lws (see its docstring and
// DESIGN.md "Encoder invariants").
inline Tables EncodeLws(const std::vector<LwsItem>& items, const Cluster& c, const std::optional<std::string>& topologyKey) {
  Tables t;
  std::map<std::string, uint32_t> domains, node_index;
  for (size_t i = 0; i < c.nodes.size(); i++) {
    const Node& n = c.nodes[i];
    lwse_node_rec r{};
    r.flags = n.schedulable ? LWSE_NODE_SCHEDULABLE : 0;
    r.domain_id = LWSE_NONE;
    r.capacity = n.capacity;
    if (topologyKey && has(n.labels, topologyKey->c_str())) {
      const std::string& val = get(n.labels, topologyKey->c_str());
      r.flags |= LWSE_NODE_HAS_TOPOLOGY;
      auto it = domains.find(val);
      if (it == domains.end()) {
        it = domains.emplace(val, (uint32_t)domains.size()).first;
        t.domain_values.push_back(val);
      }
      r.domain_id = it->second;
      r.topo_value_hash = hash64(val);
    }
    t.nodes.push_back(r);
    node_index[n.name] = (uint32_t)i;
  }
  t.n_domains = (uint32_t)domains.size();

  for (size_t li = 0; li < items.size(); li++) {
    const LwsItem& it = items[li];
    const LeaderWorkerSet& lws = it.lws;
    bool irregular = false;
    uint32_t flags = 0;
    auto parse = [&](const IntOrString& v, int32_t& out, uint32_t pct_flag) {
      if (!v.is_string) {
        out = v.int_val;
        return true;
      }
      if (!v.str_val.empty() && v.str_val.back() == '%') {
        auto n = atoi_go(v.str_val.substr(0, v.str_val.size() - 1));
        if (n) {
          out = (int32_t)*n;
          flags |= pct_flag;
          return true;
        }
      }
      out = 0;
      return false;
    };
    int32_t surge = 0, unav = 0;
    const bool ok_s = parse(lws.rollingUpdate.maxSurge, surge, LWSE_LWS_SURGE_IS_PERCENT);
    const bool ok_u = parse(lws.rollingUpdate.maxUnavailable, unav, LWSE_LWS_UNAVAIL_IS_PERCENT);
    if (!(ok_s && ok_u)) flags |= LWSE_LWS_INTSTR_INVALID;
    int32_t sts_replicas = 0, sts_partition = 0, annot = 0;
    if (it.leaderSts) {
      flags |= LWSE_LWS_STS_EXISTS;
      sts_replicas = it.leaderSts->replicas;
      sts_partition = it.leaderSts->partition;
      if (auto a = atoi_go(get(it.leaderSts->annotations, ReplicasAnnotationKey))) {
        flags |= LWSE_LWS_ANNOT_VALID;
        annot = (int32_t)*a;
      }
    }
    if (it.lwsUpdated) flags |= LWSE_LWS_UPDATED;
    uint32_t policy = lws.restartPolicy == RecreateGroupOnPodRestart ? LWSE_RESTART_ON_POD_RESTART
                      : lws.restartPolicy == RecreateGroupAfterStart ? LWSE_RESTART_AFTER_START
                                                                     : LWSE_RESTART_NONE;
    flags |= policy << LWSE_LWS_RESTART_SHIFT;
    if (has(lws.annotations, RecreateGroupAfterStartAnnotationKey)) flags |= LWSE_LWS_RECREATE_AFTER_START_ANNOT;
    if (lws.startupPolicy == LeaderReadyStartupPolicy) flags |= LWSE_LWS_STARTUP_LEADER_READY;
    if (lws.subGroupPolicyType == "LeaderExcluded") flags |= LWSE_LWS_SUBGROUP_LEADER_EXCLUDED;
    if (has(lws.annotations, ExclusiveKeyAnnotationKey)) {
      flags |= LWSE_LWS_EXCLUSIVE_TOPOLOGY;
      if (!topologyKey || get(lws.annotations, ExclusiveKeyAnnotationKey) != *topologyKey) irregular = true;
    }

    std::map<int, const Pod*> leader_by_idx;
    std::map<int, const StatefulSet*> sts_by_idx;
    std::map<int, std::vector<const Pod*>> pods_by_idx;
    std::map<std::string, const StatefulSet*> sts_by_name;
    for (const auto& s : c.statefulsets)
      if (s.ns == lws.ns) sts_by_name[s.name] = &s;
    for (const auto& p : c.pods) {
      if (p.ns != lws.ns || get(p.labels, SetNameLabelKey) != lws.name) continue;
      if (get(p.labels, WorkerIndexLabelKey) == "0") {
        auto idx = atoi_go(get(p.labels, GroupIndexLabelKey));
        if (!idx) {
          flags |= LWSE_LWS_GROUP_LABEL_INVALID;
        } else if (*idx < 0) {
          irregular = true;
        } else {
          if (leader_by_idx.count((int)*idx)) irregular = true;
          leader_by_idx[(int)*idx] = &p;
        }
      }
      if (has(p.labels, GroupIndexLabelKey)) {
        const std::string& gi = get(p.labels, GroupIndexLabelKey);
        auto idx = atoi_go(gi);
        if (!idx || *idx < 0 || std::to_string(*idx) != gi)
          irregular = true;
        else
          pods_by_idx[(int)*idx].push_back(&p);
      }
    }
    for (const auto& s : c.statefulsets) {
      if (s.ns != lws.ns || get(s.labels, SetNameLabelKey) != lws.name) continue;
      auto idx = atoi_go(get(s.labels, GroupIndexLabelKey));
      if (idx && *idx >= 0) sts_by_idx[(int)*idx] = &s;
    }
    int n_groups = 0;
    if (!leader_by_idx.empty()) n_groups = std::max(n_groups, leader_by_idx.rbegin()->first + 1);
    if (!sts_by_idx.empty()) n_groups = std::max(n_groups, sts_by_idx.rbegin()->first + 1);
    if (!pods_by_idx.empty()) n_groups = std::max(n_groups, pods_by_idx.rbegin()->first + 1);
    const uint32_t group_base = (uint32_t)t.groups.size();
    for (int idx = 0; idx < n_groups; idx++) {
      const std::string nominated = lws.name + "-" + std::to_string(idx);
      lwse_group_rec g{};
      g.leader_node = LWSE_NONE;
      g.lws_index = (uint32_t)li;
      const Pod* pod = leader_by_idx.count(idx) ? leader_by_idx[idx] : nullptr;
      const StatefulSet* wsts = nullptr;
      if (pod) {
        g.flags |= LWSE_GRP_POD_PRESENT;
        if (pod->name == nominated)
          g.flags |= LWSE_GRP_POD_NAME_MATCH;
        else
          irregular = true;
        if (pod->phase == "Running") g.flags |= LWSE_GRP_POD_RUNNING;
        if (pod->readyCondition) g.flags |= LWSE_GRP_POD_READY;
        if (pod->deletionTimestamp) g.flags |= LWSE_GRP_POD_DELETING;
        if (!get(pod->annotations, LeaderPodNameAnnotationKey).empty()) g.flags |= LWSE_GRP_MISTAKEN_ANNOTATION;
        g.leader_rev_hash = hash64(get(pod->labels, RevisionKey));
        g.leader_uid_hash = hash32(pod->uid);
        g.flags |= LWSE_GRP_REVISION_EXISTS;
        if (!pod->nodeName.empty()) {
          auto ni = node_index.find(pod->nodeName);
          g.leader_node = ni == node_index.end() ? LWSE_NODE_NOT_FOUND : ni->second;
        }
        auto si = sts_by_name.find(pod->name);
        if (si != sts_by_name.end()) wsts = si->second;
      }
      if (sts_by_idx.count(idx) && sts_by_idx[idx]->name == nominated) g.flags |= LWSE_GRP_WSTS_LABEL_NAME_MATCH;
      if (wsts) {
        g.flags |= LWSE_GRP_WSTS_FOUND;
        g.wsts_rev_hash = hash64(get(wsts->labels, RevisionKey));
        g.wsts_uid_hash = hash32(wsts->uid);
        g.wsts_spec_replicas = wsts->replicas;
        g.wsts_avail_replicas = wsts->availableReplicas;
        if (wsts->currentRevision == wsts->updateRevision) g.flags |= LWSE_GRP_WSTS_REV_SETTLED;
        if (const OwnerReference* o = controllerOf(wsts->ownerReferences); o && o->kind == "Pod") {
          g.flags |= LWSE_GRP_WSTS_OWNER_IS_POD;
          g.wsts_owner_uid_hash = hash32(o->uid);
          if (pod && o->name == pod->name) g.flags |= LWSE_GRP_WSTS_OWNER_NAME_MATCH;
        }
      }
      g.pod_base = (uint32_t)t.pod_state.size();
      std::vector<std::string> names;
      if (pods_by_idx.count(idx)) {
        for (const Pod* p : pods_by_idx[idx]) {
          uint32_t bits = p->phase == "Pending" ? LWSE_POD_PHASE_PENDING : p->phase == "Running" ? LWSE_POD_PHASE_RUNNING : 0u;
          auto any_pos = [](const std::vector<int32_t>& v) { return std::any_of(v.begin(), v.end(), [](int32_t x) { return x > 0; }); };
          if (any_pos(p->initContainerRestartCounts) || any_pos(p->containerRestartCounts)) bits |= LWSE_POD_ANY_RESTART;
          if (p->deletionTimestamp) bits |= LWSE_POD_DELETING;
          uint32_t owner_uid = 0, place = 0;  // place: the identity row's cold word (name check, node binding)
          if (const OwnerReference* o = controllerOf(p->ownerReferences)) {
            const uint32_t kind = o->kind == "Pod" ? 1u : o->kind == "StatefulSet" ? 2u : 3u;
            bits |= kind << LWSE_POD_OWNER_SHIFT;
            owner_uid = hash32(o->uid);
            if (o->name == nominated)
              bits |= LWSE_POD_OWNER_NAME_MATCH;
            else if (kind == 2u)
              irregular = true;
          }
          if (get(p->labels, WorkerIndexLabelKey) == "0") {
            bits |= LWSE_POD_IS_LEADER;
            place |= LWSE_PODID_NAME_OK;
          } else {
            auto po = GetParentNameAndOrdinal(p->name);
            if (po.second != -1) {
              place |= LWSE_PODID_NAME_OK;
              if (po.first != nominated) irregular = true;
            }
          }
          if (!p->nodeName.empty()) {
            auto ni = node_index.find(p->nodeName);
            if (ni != node_index.end() && ni->second <= LWSE_POD_NODE_MAX)
              place |= LWSE_PODID_SCHEDULED | (ni->second << LWSE_PODID_NODE_SHIFT);
          }
          const uint64_t rev = hash64(get(p->labels, RevisionKey));
          t.pod_state.push_back((lwse_pod_state)bits);
          t.pod_ident.push_back({rev, owner_uid, place});
          names.push_back(p->name);
        }
      }
      g.pod_count = (uint32_t)names.size();
      t.group_pod_names.push_back(std::move(names));
      t.groups.push_back(g);
    }
    if (irregular) flags |= LWSE_LWS_IRREGULAR;
    lwse_lws_rec r{};
    r.rev_hash = hash64(it.revisionKey);
    r.size = lws.size;
    r.flags = flags;
    r.replicas = lws.replicas;
    r.partition = lws.rollingUpdate.partition;
    r.max_surge = surge;
    r.max_unavailable = unav;
    r.sts_replicas = sts_replicas;
    r.sts_partition = sts_partition;
    r.sts_replicas_annotation = annot;
    r.subgroup_size = lws.subGroupSize.value_or(0);
    r.uid_hash = hash64(lws.uid);
    r.group_base = group_base;
    r.group_count = (uint32_t)n_groups;
    t.lws.push_back(r);
  }
  return t;
}

// ---------------------------------------------------------------------------
// Engine + reconciler facades
// ---------------------------------------------------------------------------
class Engine {
 public:
  explicit Engine(int device = 0) {
    lwse_config cfg{LWSE_ABI_VERSION, device, 0, 0};
    const int rc = lwse_create(&cfg, &h_);
    if (rc != LWSE_OK) throw std::runtime_error(std::string("lwse_create: ") + lwse_strerror(rc));
  }
  ~Engine() { lwse_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  lwse_engine* handle() const { return h_; }

 private:
  lwse_engine* h_ = nullptr;
};

struct Sweep {  // one sweep's inputs and results
  Tables tables;
  std::vector<lwse_lws_out> lws_out;
  std::vector<lwse_group_out> group_out;
};

inline Sweep RunSweep(Engine& e, Tables tables, bool gang = false) {
  Sweep s;
  s.tables = std::move(tables);
  s.lws_out.resize(s.tables.lws.size());
  s.group_out.resize(s.tables.groups.size());
  int rc = lwse_upload_nodes(e.handle(), s.tables.nodes.data(), (uint32_t)s.tables.nodes.size(), s.tables.n_domains);
  if (rc != LWSE_OK) throw std::runtime_error(std::string("lwse_upload_nodes: ") + lwse_strerror(rc));
  lwse_lws_tables t{};
  t.lws = s.tables.lws.data();
  t.n_lws = (uint32_t)s.tables.lws.size();
  t.groups = s.tables.groups.data();
  t.n_groups = (uint32_t)s.tables.groups.size();
  t.pod_state = s.tables.pod_state.data();
  t.pod_ident = s.tables.pod_ident.data();
  t.n_pods = s.tables.pod_state.size();
  t.lws_out = s.lws_out.data();
  t.group_out = s.group_out.data();
  t.flags = gang ? LWSE_SWEEP_GANG : 0;
  rc = lwse_sweep_lws_host(e.handle(), &t);
  if (rc != LWSE_OK) throw std::runtime_error(std::string("lwse_sweep_lws_host: ") + lwse_strerror(rc));
  return s;
}

// pkg/controllers/leaderworkerset_controller.go — same method names, same results.
class LeaderWorkerSetReconciler {
 public:
  explicit LeaderWorkerSetReconciler(const Sweep& s) : s_(s) {}
  struct Params {
    int32_t stsPartition, replicas;
    bool err;
  };
  // :280 rollingUpdateParameters(ctx, lws, sts, revisionKey, leaderWorkerSetUpdated)
  Params rollingUpdateParameters(size_t lws_row) const {
    const lwse_lws_out& o = s_.lws_out.at(lws_row);
    return {o.sts_partition, o.sts_replicas, (o.flags & LWSE_LOUT_RUP_ERROR) != 0};
  }
  struct Conditions {
    int32_t readyReplicas, updatedReplicas;
    lwse_condition condition;
    bool updateDone, err;
  };
  // :414 updateConditions
  Conditions updateConditions(size_t lws_row) const {
    const lwse_lws_out& o = s_.lws_out.at(lws_row);
    return {o.ready_replicas, o.updated_replicas, (lwse_condition)((o.flags & LWSE_LOUT_COND_MASK) >> LWSE_LOUT_COND_SHIFT),
            (o.flags & LWSE_LOUT_UPDATE_DONE) != 0, (o.flags & LWSE_LOUT_STATUS_ERROR) != 0};
  }
  int32_t stsMaxUnavailable(size_t lws_row) const { return s_.lws_out.at(lws_row).sts_max_unavailable; }  // :811-830

 private:
  const Sweep& s_;
};

// pkg/controllers/pod_controller.go
class PodReconciler {
 public:
  explicit PodReconciler(const Sweep& s) : s_(s) {}
  struct Restart {
    bool leaderDeleted, issuedDelete, err;
    std::string triggerPod;
  };
  // :204 handleRestartPolicy, for the group of row `group_row`
  Restart handleRestartPolicy(size_t group_row) const {
    const lwse_group_out& g = s_.group_out.at(group_row);
    Restart r{};
    r.issuedDelete = g.flags & LWSE_GOUT_DELETE_LEADER;
    r.leaderDeleted = g.flags & (LWSE_GOUT_DELETE_LEADER | LWSE_GOUT_LEADER_DELETING);
    r.err = g.flags & LWSE_GOUT_RESTART_ERROR;
    if (g.first_trigger != LWSE_NONE) r.triggerPod = s_.tables.group_pod_names.at(group_row).at(g.first_trigger);
    return r;
  }
  // :186-198 — whether the leader pod's reconcile creates the worker StatefulSet (replicas size-1, ordinals from 1)
  std::optional<int32_t> createWorkerStatefulSet(size_t group_row) const {
    const lwse_group_out& g = s_.group_out.at(group_row);
    if (g.flags & LWSE_GOUT_CREATE_WSTS) return g.worker_replicas;
    return std::nullopt;
  }
  // :315 topologyValueFromPod → the nodeSelector value of the workers
  std::optional<std::string> topologyValueFromPod(size_t group_row) const {
    const lwse_group_out& g = s_.group_out.at(group_row);
    if (g.domain_id == LWSE_NONE) return std::nullopt;
    return s_.tables.domain_values.at(g.domain_id);
  }

 private:
  const Sweep& s_;
};

}  // namespace lws
