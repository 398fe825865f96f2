// Self-check of the C++ host mirror (lws_b200/csrc/host/lws_host.hpp).
//   host_check encode   prints the record tables of a fixed scenario as hex (CPU only);
//                       tests/test_host_cpp.py compares them with the Python encoder's.
//   host_check gpu      runs the reference's unit KATs through Engine + the reconciler
//                       facades (needs a B200): leaderworkerset_controller_test.go:887-1011,
//                       pod_controller_test.go:427-532.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "../../lws_b200/csrc/host/lws_host.hpp"

using namespace lws;

static Labels podLabels(const std::string& lws, int group, int worker, const std::string& rev) {
  return {{SetNameLabelKey, lws}, {WorkerIndexLabelKey, std::to_string(worker)}, {GroupIndexLabelKey, std::to_string(group)},
          {RevisionKey, rev}};
}

static void scenario(std::vector<LwsItem>& items, Cluster& c) {
  for (int i = 0; i < 3; i++) c.nodes.push_back({"node-" + std::to_string(i), {{"rack", i < 2 ? "r0" : "r1"}}, 8, true});
  c.nodes.push_back({"node-3", {}, 8, true});
  LeaderWorkerSet a;
  a.name = "alpha";
  a.uid = "uid-alpha";
  a.replicas = 3;
  a.size = 2;
  a.rollingUpdate.maxSurge = IntOrString::FromString("50%");
  a.annotations[ExclusiveKeyAnnotationKey] = "rack";
  StatefulSet asts;
  asts.name = "alpha";
  asts.replicas = 3;
  asts.partition = 2;
  asts.annotations[ReplicasAnnotationKey] = "3";
  items.push_back({a, "rev-2", false, asts});
  for (int g = 0; g < 3; g++) {
    const std::string ln = "alpha-" + std::to_string(g);
    Pod l;
    l.name = ln;
    l.uid = "uid-" + ln;
    l.labels = podLabels("alpha", g, 0, g == 2 ? "rev-2" : "rev-1");
    l.phase = "Running";
    l.readyCondition = g != 1;
    l.nodeName = g == 0 ? "node-0" : (g == 1 ? "node-3" : "");
    c.pods.push_back(l);
    StatefulSet w;
    w.name = ln;
    w.uid = "uid-sts-" + ln;
    w.labels = {{SetNameLabelKey, "alpha"}, {GroupIndexLabelKey, std::to_string(g)}, {RevisionKey, g == 2 ? "rev-2" : "rev-1"}};
    w.replicas = 1;
    w.availableReplicas = 1;
    w.ownerReferences = {{"Pod", ln, "uid-" + ln, true}};
    c.statefulsets.push_back(w);
    Pod wk;
    wk.name = ln + "-1";
    wk.uid = "uid-" + ln + "-1";
    wk.labels = podLabels("alpha", g, 1, g == 2 ? "rev-2" : "rev-1");
    wk.phase = g == 2 ? "Pending" : "Running";
    wk.containerRestartCounts = {g == 0 ? 2 : 0};
    wk.deletionTimestamp = g == 1;
    wk.ownerReferences = {{"StatefulSet", ln, g == 1 ? "uid-stale" : "uid-sts-" + ln, true}};
    wk.nodeName = g == 0 ? "node-1" : "";
    c.pods.push_back(wk);
  }
  LeaderWorkerSet b;
  b.name = "beta";
  b.uid = "uid-beta";
  b.replicas = 2;
  b.size = 1;
  b.restartPolicy = "None";
  b.startupPolicy = LeaderReadyStartupPolicy;
  b.rollingUpdate.maxUnavailable = IntOrString::FromString("oops");
  items.push_back({b, "rev-9", true, std::nullopt});
  Pod bl;
  bl.name = "beta-0";
  bl.uid = "uid-beta-0";
  bl.labels = podLabels("beta", 0, 0, "rev-9");
  bl.phase = "Running";
  bl.readyCondition = true;
  c.pods.push_back(bl);
}

template <typename T>
static void dump(const char* name, const std::vector<T>& v) {
  std::printf("%s ", name);
  const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
  for (size_t i = 0; i < v.size() * sizeof(T); i++) std::printf("%02x", p[i]);
  std::printf("\n");
}

#define CHECK(cond)                                                    \
  do {                                                                 \
    if (!(cond)) {                                                     \
      std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                    \
    }                                                                  \
  } while (0)

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "encode";
  if (mode == "encode") {
    std::vector<LwsItem> items;
    Cluster c;
    scenario(items, c);
    Tables t = EncodeLws(items, c, std::string("rack"));
    dump("lws", t.lws);
    dump("groups", t.groups);
    dump("pod_state", t.pod_state);
    dump("pod_ident", t.pod_ident);
    dump("nodes", t.nodes);
    std::printf("n_domains %u\n", t.n_domains);
    return 0;
  }
  // ---- gpu: the reference's unit KATs through the facades ----
  Engine eng(0);
  struct Kat {
    int32_t replicas, mu, ms;
    bool updated;
    int32_t want_partition, want_replicas;
  } kats[] = {{3, 0, 1, false, 0, 3}, {3, 0, 1, true, 2, 3}, {2, 1, 2, true, 2, 3}};  // :887-1011
  for (const Kat& k : kats) {
    LeaderWorkerSet l;
    l.name = "test-sample";
    l.uid = "u";
    l.replicas = k.replicas;
    l.size = 1;
    l.rollingUpdate.maxUnavailable = IntOrString::FromInt(k.mu);
    l.rollingUpdate.maxSurge = IntOrString::FromInt(k.ms);
    StatefulSet sts;
    sts.name = l.name;
    sts.replicas = 2;
    sts.annotations[ReplicasAnnotationKey] = "2";
    Sweep s = RunSweep(eng, EncodeLws({{l, "rev-new", k.updated, sts}}, Cluster{}, std::nullopt));
    auto p = LeaderWorkerSetReconciler(s).rollingUpdateParameters(0);
    CHECK(!p.err && p.stsPartition == k.want_partition && p.replicas == k.want_replicas);
  }
  for (int stale = 0; stale < 2; stale++) {  // pod_controller_test.go:427-532
    LeaderWorkerSet l;
    l.name = "test-sample";
    l.uid = "u";
    l.replicas = 1;
    l.size = 2;
    Cluster c;
    Pod leader;
    leader.name = "test-sample-0";
    leader.uid = "leader-current";
    leader.labels = podLabels(l.name, 0, 0, "revision-1");
    StatefulSet sts;
    sts.name = leader.name;
    sts.uid = "sts-current";
    sts.labels = {{SetNameLabelKey, l.name}, {GroupIndexLabelKey, "0"}};
    sts.ownerReferences = {{"Pod", leader.name, leader.uid, true}};
    Pod worker;
    worker.name = "test-sample-0-1";
    worker.uid = "w";
    worker.labels = podLabels(l.name, 0, 1, "revision-1");
    worker.deletionTimestamp = true;
    worker.ownerReferences = {{"StatefulSet", sts.name, stale ? "sts-stale" : "sts-current", true}};
    c.pods = {leader, worker};
    c.statefulsets = {sts};
    StatefulSet lsts;
    lsts.name = l.name;
    lsts.replicas = 1;
    lsts.annotations[ReplicasAnnotationKey] = "1";
    Sweep s = RunSweep(eng, EncodeLws({{l, "revision-1", false, lsts}}, c, std::nullopt));
    auto r = PodReconciler(s).handleRestartPolicy(0);
    CHECK(!r.err && r.leaderDeleted == !stale);
    if (!stale) CHECK(r.issuedDelete && r.triggerPod == "test-sample-0-1");
  }
  {  // the scenario's topology lookups
    std::vector<LwsItem> items;
    Cluster c;
    scenario(items, c);
    Sweep s = RunSweep(eng, EncodeLws(items, c, std::string("rack")));
    PodReconciler pr(s);
    CHECK(pr.topologyValueFromPod(0).value_or("") == "r0");
    CHECK(s.group_out[1].flags & LWSE_GOUT_TOPOLOGY_ERROR);  // node-3 has no rack label
    CHECK(s.group_out[2].flags & LWSE_GOUT_WAIT_SCHEDULE);   // leader unscheduled
    CHECK(s.lws_out[1].flags & LWSE_LOUT_RUP_ERROR || !(s.tables.lws[1].flags & LWSE_LWS_STS_EXISTS));
  }
  std::puts("host_check gpu: ok");
  return 0;
}
