/*
 * cgo_pattern_check.c — the calling pattern of the Go shim (go/lwse), in C, because this image has no
 * Go toolchain: plain C, linked against liblwse.so exactly as cgo links it.
 *
 *   - every table lives in C-malloc'd memory (cgo forbids the C side to keep Go-heap pointers, so the
 *     shim allocates with C.malloc) — pageable, 64-byte aligned by posix_memalign;
 *   - ONE engine handle is shared by several threads, the way the LeaderWorkerSet, Pod and
 *     DisaggregatedSet reconcilers and the webhook handlers (goroutines on different OS threads) share
 *     it: each thread loops over its own entry point with its own buffers, all at once;
 *   - a "sweeper" thread drives the resident tick from the engine's pinned arena.
 *
 * usage: cgo_pattern_check <dir> <iterations>
 *   <dir> holds the raw little-endian tables written by tests/test_cgo_pattern.py:
 *   lws.bin groups.bin pod_state.bin pod_ident.bin nodes.bin reqs.bin occ.bin keys.bin key_offsets.bin meta.txt
 *   The program writes lws_out.bin group_out.bin place_out.bin digests.bin tick_group_out.bin next to them;
 *   the Python test compares them with the oracle.  Exit code 0 = every call returned LWSE_OK and repeated
 *   calls of a thread gave identical bytes.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/lwse.h"

typedef struct {
  void* p;
  size_t bytes;
} buf;

static buf load(const char* dir, const char* name) {
  char path[1024];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "rb");
  buf b = {NULL, 0};
  if (!f) {
    fprintf(stderr, "missing %s\n", path);
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  b.bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  if (posix_memalign(&b.p, 64, b.bytes + 64)) exit(2);  /* C.malloc'd, aligned for the ABI */
  if (b.bytes && fread(b.p, 1, b.bytes, f) != b.bytes) exit(2);
  fclose(f);
  return b;
}

static void save(const char* dir, const char* name, const void* p, size_t bytes) {
  char path[1024];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "wb");
  if (!f || (bytes && fwrite(p, 1, bytes, f) != bytes)) exit(2);
  fclose(f);
}

static void* xalloc(size_t bytes) {
  void* p = NULL;
  if (posix_memalign(&p, 64, bytes + 64)) exit(2);
  memset(p, 0, bytes + 64);
  return p;
}

typedef struct {
  lwse_engine* e;
  const char* dir;
  int iterations;
  buf lws, groups, pst, pid, nodes, reqs, occ, keys, koff;
  uint32_t n_domains, n_namespaces, sweep_flags;
  int failed;
} ctx;

/* LeaderWorkerSet / Pod reconcilers: the stateless sweep */
static void* sweep_thread(void* arg) {
  ctx* c = (ctx*)arg;
  lwse_lws_tables t;
  memset(&t, 0, sizeof t);
  t.lws = (const lwse_lws_rec*)c->lws.p;
  t.n_lws = (uint32_t)(c->lws.bytes / sizeof(lwse_lws_rec));
  t.groups = (const lwse_group_rec*)c->groups.p;
  t.n_groups = (uint32_t)(c->groups.bytes / sizeof(lwse_group_rec));
  t.pod_state = (const lwse_pod_state*)c->pst.p;
  t.pod_ident = (const lwse_pod_ident*)c->pid.p;
  t.n_pods = c->pst.bytes;
  t.flags = c->sweep_flags;
  size_t b_lo = (size_t)t.n_lws * sizeof(lwse_lws_out), b_go = (size_t)t.n_groups * sizeof(lwse_group_out);
  void* first_lo = NULL;
  void* first_go = NULL;
  for (int it = 0; it < c->iterations; it++) {
    t.lws_out = (lwse_lws_out*)xalloc(b_lo);
    t.group_out = (lwse_group_out*)xalloc(b_go);
    int rc = lwse_sweep_lws_host(c->e, &t);
    if (rc != LWSE_OK) {
      fprintf(stderr, "sweep: %s\n", lwse_strerror(rc));
      c->failed = 1;
      return NULL;
    }
    if (it == 0) {
      first_lo = t.lws_out;
      first_go = t.group_out;
    } else {
      if (memcmp(first_lo, t.lws_out, b_lo) || memcmp(first_go, t.group_out, b_go)) {
        fprintf(stderr, "sweep: iteration %d differs from iteration 0\n", it);
        c->failed = 1;
      }
      free(t.lws_out);
      free(t.group_out);
    }
  }
  save(c->dir, "lws_out.bin", first_lo, b_lo);
  save(c->dir, "group_out.bin", first_go, b_go);
  return NULL;
}

/* the scheduler side: placement rounds */
static void* place_thread(void* arg) {
  ctx* c = (ctx*)arg;
  uint32_t n = (uint32_t)(c->reqs.bytes / sizeof(lwse_place_req));
  size_t b = (size_t)n * sizeof(lwse_place_out);
  void* first = NULL;
  for (int it = 0; it < c->iterations; it++) {
    lwse_place_out* out = (lwse_place_out*)xalloc(b);
    uint32_t rounds = 0;
    int rc = lwse_place_host(c->e, (const lwse_place_req*)c->reqs.p, n, (const uint32_t*)c->occ.p, c->n_namespaces, out, &rounds);
    if (rc != LWSE_OK) {
      fprintf(stderr, "place: %s\n", lwse_strerror(rc));
      c->failed = 1;
      return NULL;
    }
    if (it == 0) {
      first = out;
    } else {
      if (memcmp(first, out, b)) {
        fprintf(stderr, "place: iteration %d differs from iteration 0\n", it);
        c->failed = 1;
      }
      free(out);
    }
  }
  save(c->dir, "place_out.bin", first, b);
  return NULL;
}

/* webhook admissions: SHA-1 group keys */
static void* keys_thread(void* arg) {
  ctx* c = (ctx*)arg;
  uint32_t n = (uint32_t)(c->koff.bytes / 4) - 1;
  size_t b = (size_t)n * 20;
  void* first = NULL;
  for (int it = 0; it < c->iterations; it++) {
    uint8_t* dig = (uint8_t*)xalloc(b);
    int rc = lwse_group_keys_host(c->e, (const uint8_t*)c->keys.p, (const uint32_t*)c->koff.p, n, dig);
    if (rc != LWSE_OK) {
      fprintf(stderr, "keys: %s\n", lwse_strerror(rc));
      c->failed = 1;
      return NULL;
    }
    if (it == 0) {
      first = dig;
    } else {
      if (memcmp(first, dig, b)) {
        fprintf(stderr, "keys: iteration %d differs from iteration 0\n", it);
        c->failed = 1;
      }
      free(dig);
    }
  }
  save(c->dir, "digests.bin", first, b);
  return NULL;
}

/* the sweeper goroutine: resident tables, ticks fed from the pinned arena */
static void* tick_thread(void* arg) {
  ctx* c = (ctx*)arg;
  lwse_lws_tables t;
  memset(&t, 0, sizeof t);
  t.lws = (const lwse_lws_rec*)c->lws.p;
  t.n_lws = (uint32_t)(c->lws.bytes / sizeof(lwse_lws_rec));
  t.groups = (const lwse_group_rec*)c->groups.p;
  t.n_groups = (uint32_t)(c->groups.bytes / sizeof(lwse_group_rec));
  t.pod_state = (const lwse_pod_state*)c->pst.p;
  t.pod_ident = (const lwse_pod_ident*)c->pid.p;
  t.n_pods = c->pst.bytes;
  int rc = lwse_resident_load(c->e, &t);
  void* arena = NULL;
  uint64_t arena_bytes = 0;
  if (rc == LWSE_OK) rc = lwse_resident_arena(c->e, 1 << 20, &arena, &arena_bytes);
  if (rc != LWSE_OK) {
    fprintf(stderr, "resident: %s\n", lwse_strerror(rc));
    c->failed = 1;
    return NULL;
  }
  /* watch events: flip the restart bit of every 7th pod forth and back; the table ends as it began */
  uint32_t n_patch = (uint32_t)(t.n_pods / 7);
  uint32_t* rows = (uint32_t*)arena;
  uint8_t* vals = (uint8_t*)arena + (((size_t)n_patch * 4 + 255) & ~(size_t)255);
  lwse_tick tick;
  for (int it = 0; it < 2 * c->iterations; it++) {
    for (uint32_t k = 0; k < n_patch; k++) {
      rows[k] = k * 7;
      vals[k] = ((const uint8_t*)c->pst.p)[k * 7] ^ ((it & 1) ? 0 : LWSE_POD_ANY_RESTART);
    }
    lwse_patch_seg seg;
    memset(&seg, 0, sizeof seg);
    seg.table = LWSE_TABLE_POD_STATE;
    seg.n = n_patch;
    seg.rows = rows;
    seg.values = vals;
    memset(&tick, 0, sizeof tick);
    tick.segs = &seg;
    tick.n_segs = 1;
    tick.flags = c->sweep_flags;
    rc = lwse_resident_tick(c->e, &tick);
    if (rc != LWSE_OK) {
      fprintf(stderr, "tick: %s\n", lwse_strerror(rc));
      c->failed = 1;
      return NULL;
    }
  }
  size_t b_go = (size_t)t.n_groups * sizeof(lwse_group_out);
  lwse_group_out* go = (lwse_group_out*)xalloc(b_go);
  lwse_lws_out* lo = (lwse_lws_out*)xalloc((size_t)t.n_lws * sizeof(lwse_lws_out));
  rc = lwse_resident_outputs(c->e, lo, go);
  if (rc != LWSE_OK) c->failed = 1;
  save(c->dir, "tick_group_out.bin", go, b_go);
  return NULL;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  ctx c;
  memset(&c, 0, sizeof c);
  c.dir = argv[1];
  c.iterations = atoi(argv[2]);
  c.lws = load(c.dir, "lws.bin");
  c.groups = load(c.dir, "groups.bin");
  c.pst = load(c.dir, "pod_state.bin");
  c.pid = load(c.dir, "pod_ident.bin");
  c.nodes = load(c.dir, "nodes.bin");
  c.reqs = load(c.dir, "reqs.bin");
  c.occ = load(c.dir, "occ.bin");
  c.keys = load(c.dir, "keys.bin");
  c.koff = load(c.dir, "key_offsets.bin");
  {
    char path[1024];
    snprintf(path, sizeof path, "%s/meta.txt", c.dir);
    FILE* f = fopen(path, "r");
    if (!f || fscanf(f, "%u %u %u", &c.n_domains, &c.n_namespaces, &c.sweep_flags) != 3) return 2;
    fclose(f);
  }
  lwse_config cfg = {LWSE_ABI_VERSION, 0, 0, 0};
  int rc = lwse_create(&cfg, &c.e);
  if (rc != LWSE_OK) {
    fprintf(stderr, "lwse_create: %s\n", lwse_strerror(rc));
    return 3;
  }
  rc = lwse_upload_nodes(c.e, (const lwse_node_rec*)c.nodes.p, (uint32_t)(c.nodes.bytes / sizeof(lwse_node_rec)), c.n_domains);
  if (rc != LWSE_OK) return 4;
  pthread_t th[4];
  pthread_create(&th[0], NULL, sweep_thread, &c);
  pthread_create(&th[1], NULL, place_thread, &c);
  pthread_create(&th[2], NULL, keys_thread, &c);
  pthread_create(&th[3], NULL, tick_thread, &c);
  for (int k = 0; k < 4; k++) pthread_join(th[k], NULL);
  lwse_destroy(c.e);
  if (c.failed) return 1;
  printf("cgo pattern ok: 4 threads x %d iterations on one handle\n", c.iterations);
  return 0;
}
