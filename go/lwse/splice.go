package lwse

// How the reconcilers consume a sweep (sketch of the splice points; the encoder that
// fills the tables from the informer cache is the next row of SURVEY.md §8(f)).
//
//   cmd/main.go:158            after ctrl.NewManager: eng, err := lwse.New(dev); defer eng.Close()
//                              a Sweeper goroutine encodes the cache into tables, calls
//                              eng.SweepLws / eng.Place / eng.SweepDs and publishes the result
//                              tables behind an atomic pointer keyed by object UID.
//
//   leaderworkerset_controller.go:159
//       partition, replicas, err := r.rollingUpdateParameters(ctx, lws, leaderSts, key, lwsUpdated)
//     becomes
//       o, ok := r.Sweep.LwsOut(lws.UID)            // lwse_lws_out row of this object
//       if !ok || o.Irregular() { /* stock path above */ }
//       if o.RupError() { return ctrl.Result{}, errRollingUpdate }
//       partition, replicas := o.StsPartition, o.StsReplicas
//       // o.Event(): 1 → "deleting surge replica %s-%d", 2 → "deleting surge replicas from … to …"
//
//   leaderworkerset_controller.go:196/:553 (updateConditions)
//       readyCount, updatedCount := o.ReadyReplicas, o.UpdatedReplicas
//       cond := o.Condition()  // 0 Progressing, 1 Available, 2 UpdateInProgress+Progressing
//       updateDone := o.UpdateDone()
//
//   leaderworkerset_controller.go:811-830: stsMaxUnavailable := o.StsMaxUnavailable
//
//   pod_controller.go:91 (handleRestartPolicy), for the pod's group row g:
//       if g.DeleteLeader() { r.Delete(ctx, &leader, foreground) ; return true }   // :259
//       if g.LeaderDeleting() { return true }                                       // :255
//       if g.RestartError() { return false, fmt.Errorf("parsing pod name …") }      // :231
//
//   pod_controller.go:100-198 (leader pod): g.CreatePodGroup() → SchedulerProvider.CreatePodGroupIfNotExists
//       (MinMember = o.MinMember); g.RequeueRevision() → RequeueAfter 1s; g.WaitSchedule() → return;
//       g.TopologyError() → error; g.CreateWsts() → create worker sts with replicas g.WorkerReplicas,
//       ordinals.start = 1, nodeSelector{topologyKey: domainValue[g.DomainID]}
//
//   disaggregatedset/executor.go:153,166: nextStep/scale targets from lwse_ds_role_out /
//       lwse_ds_revrole_out; disaggregatedset_controller.go:203-236 from lwse_ds_out.drained_revs.
