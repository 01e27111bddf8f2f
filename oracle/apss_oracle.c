/*
 * apss_oracle.c -- CPU ORACLE for inverted-index all-pairs similarity scoring.
 *
 * >>> TEST INFRASTRUCTURE ONLY.  Nothing under all-pairs-similarity_b200/ may link,
 * >>> import or call this file.  Only tests/, __graft_entry__.smoke() and the
 * >>> cpu_baseline / --impl reference legs of bench.py use it, as the checker / baseline.
 *
 * >>> PARITY UNPINNED: the reference (Scala 2.10 / Akka 2.3.4) ships no tests, no golden
 * >>> vectors and cannot be compiled or run in this image (no JVM).  This file is a plain-C
 * >>> restatement written from the cited reference lines; it is pinned only by the
 * >>> hand-derived known-answer tests of SURVEY.md section 8(c) and a brute-force check.
 *
 * Reference files restated (paths relative to /root/reference/core/src/main/scala/cpslab):
 *   IWA = deploy/server/IndexingWorkerActor.scala
 *   WWA = deploy/server/WriteWorkerActor.scala
 *   EPA = deploy/server/EntryProxyActor.scala
 *   CU  = deploy/CommonUtils.scala
 *
 * Two algorithms live here:
 *   ALGO_FAITHFUL  the as-built actor pipeline: admission filter (EPA:81-93), value prune
 *                  (WWA:185-202), dimension routing to emulated index workers (WWA:164-183,
 *                  CU:28-40, EPA:37-49), id-only posting sets (IWA:61-71), per-candidate
 *                  hash-join dot product (CU:98-117), de-dup of passing candidates only and
 *                  the first-posting-list skip (IWA:74-111).  Semantics R0 (as built) or R1
 *                  (the guard bug fixed); see SURVEY.md section 8(a).
 *   ALGO_FAST      weighted postings + dense fp64 accumulator (R1 only).  Used where the
 *                  faithful algorithm would take hours; cross-checked against it in tests.
 *                  oracle_set_pruning() adds the EXACT INDEX REDUCTION the reference planned around
 *                  its max-weight stub (EPA:51-57,81-93) -- not reference behaviour,
 *                  a restatement of the GPU library's rule so that its counters can be predicted:
 *                  the pair set must equal the un-pruned one (tests check exactly that).
 *
 * Similarities are accumulated in ascending-dimension order with separate multiply and add
 * (compile with -ffp-contract=off), which is the order the CUDA fp64 verify kernel uses, so
 * parity on similarity values is bit-exact.  The reference's own order is the iteration order
 * of a scala.collection.mutable.HashMap (CU:110), which is unspecified (last-ulp effects only).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_R1 0
#define ORC_R0 1
#define ORC_ALGO_FAITHFUL 0
#define ORC_ALGO_FAST 1
#define ORC_FLAG_QUERY_ONLY 1
#define ORC_FLAG_INDEX_ONLY 2   /* bulk load (IWA:61-71 only); used to build the CPU baseline's index */
#define ORC_FLAG_SKIP_ADMIT 4   /* IndexData carries vectors that already passed EPA:81-93 upstream     */

#define ORC_ST_REJECTED 0   /* failed EPA:81-93 admission                         */
#define ORC_ST_EMPTY    1   /* admitted, pruned to zero components (WWA:192-199)  */
#define ORC_ST_ACTIVE   2   /* admitted, >=1 component: indexed and queried       */

/* ------------------------------------------------------------------ small containers */

typedef struct { int32_t *v; int32_t n, cap; } ivec_t;
static void ivec_push(ivec_t *a, int32_t x) {
  if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4; a->v = (int32_t*)realloc(a->v, sizeof(int32_t) * (size_t)a->cap); }
  a->v[a->n++] = x;
}

/* open-addressing map int64 -> int64 (keys >= 0 or any; EMPTY marks by separate flag) */
typedef struct { int64_t *k; int64_t *val; uint8_t *used; int64_t cap, n; } map64_t;
static uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static void map64_init(map64_t *m, int64_t cap) {
  int64_t c = 8; while (c < cap * 2) c <<= 1;
  m->cap = c; m->n = 0;
  m->k = (int64_t*)malloc(sizeof(int64_t) * (size_t)c);
  m->val = (int64_t*)malloc(sizeof(int64_t) * (size_t)c);
  m->used = (uint8_t*)calloc((size_t)c, 1);
}
static void map64_free(map64_t *m) { free(m->k); free(m->val); free(m->used); memset(m, 0, sizeof(*m)); }
static int64_t map64_find(const map64_t *m, int64_t key) {
  if (!m->cap) return -1;
  uint64_t i = mix64((uint64_t)key) & (uint64_t)(m->cap - 1);
  while (m->used[i]) { if (m->k[i] == key) return (int64_t)i; i = (i + 1) & (uint64_t)(m->cap - 1); }
  return -1;
}
static void map64_put(map64_t *m, int64_t key, int64_t val);
static void map64_grow(map64_t *m) {
  map64_t o = *m; map64_init(m, o.cap);
  for (int64_t i = 0; i < o.cap; i++) if (o.used[i]) map64_put(m, o.k[i], o.val[i]);
  free(o.k); free(o.val); free(o.used);
}
static void map64_put(map64_t *m, int64_t key, int64_t val) {
  if (!m->cap) map64_init(m, 8);
  if ((m->n + 1) * 2 > m->cap) map64_grow(m);
  uint64_t i = mix64((uint64_t)key) & (uint64_t)(m->cap - 1);
  while (m->used[i]) { if (m->k[i] == key) { m->val[i] = val; return; } i = (i + 1) & (uint64_t)(m->cap - 1); }
  m->used[i] = 1; m->k[i] = key; m->val[i] = val; m->n++;
}
static void map64_clear(map64_t *m) { if (m->cap) memset(m->used, 0, (size_t)m->cap); m->n = 0; }

/* ------------------------------------------------------------------ oracle state */

typedef struct {            /* a stored (pruned) vector: what WWA:193-194 puts in vectorsStore */
  int64_t key;              /* stand-in for the caller's String id                            */
  int32_t nnz;
  int32_t *idx;             /* ascending (SparseVector.scala:96-108 sorts)                    */
  double *val;
  uint8_t *skip;            /* index reduction: 1 = component kept out of the index (else NULL) */
} ovec_t;

typedef struct {            /* one emulated IndexingWorkerActor (IWA:21-25)                   */
  ivec_t store_vec;         /* vectorsStore: position -> global ordinal of the vector         */
  map64_t dim2list;         /* invertedIndex keys: dim -> index into lists[]                  */
  ivec_t *lists; int32_t n_lists, cap_lists;   /* posting "sets": positions in store_vec      */
} oworker_t;

typedef struct { int64_t qkey, ckey; int32_t q, c; double sim; } opair_t;

typedef struct {
  int32_t dim; double sim_thr, idx_thr;
  int32_t semantics, algo;
  int32_t max_shard, max_entry, max_index_actor;   /* WWA:48, CU:24, EPA:21 */
  int32_t set_order;        /* 0: Scala 2.10.4 Set1-4 / HashSet trie order (UNVERIFIED recall), 1: ascending */
  int32_t frozen;           /* IWA:143-144 */
  double *maxw;             /* EPA:51-57: NULL -> 1.0 for every dim < vectorDim */
  /* global vector table, ordinal = order of submission in indexing calls */
  ovec_t *vecs; int64_t n_vecs, cap_vecs;
  /* faithful */
  oworker_t *workers; int32_t n_workers;
  /* fast */
  ivec_t *fp_ids; double **fp_w; int32_t *fp_wcap;  /* per-dim weighted postings */
  /* last batch results */
  opair_t *pairs; int64_t n_pairs, cap_pairs;
  uint8_t *status; int32_t n_status;
  int64_t id_base;
  int64_t postings_visited, candidates_unique, dot_calls_ref;
  int64_t tot_postings_visited, tot_candidates_unique, tot_dot_calls_ref, tot_pairs;
  int32_t threads;
  /* exact index reduction (ALGO_FAST): off unless oracle_set_pruning() was called */
  int32_t prune; double prune_lim, max_qnorm; int32_t *df; int64_t n_unindexed;
  char err[256];
} oracle_t;

/* ------------------------------------------------------------------ arithmetic */

/* CU:98-117 restated literally: two int->double hash maps are built, then every entry of
 * vector1's map probes vector2's map.  Iteration here is ascending index (see file header). */
typedef struct { int32_t *k; double *v; int32_t cap; } smap_t;
static void smap_build(smap_t *m, const int32_t *idx, const double *val, int32_t n, int32_t *kbuf, double *vbuf, int32_t cap) {
  m->k = kbuf; m->v = vbuf; m->cap = cap;
  for (int32_t i = 0; i < cap; i++) kbuf[i] = -1;
  for (int32_t i = 0; i < n; i++) {
    uint32_t h = ((uint32_t)idx[i] * 2654435761u) & (uint32_t)(cap - 1);
    while (kbuf[h] >= 0) h = (h + 1) & (uint32_t)(cap - 1);
    kbuf[h] = idx[i]; vbuf[h] = val[i];
  }
}
static int smap_get(const smap_t *m, int32_t key, double *out) {
  uint32_t h = ((uint32_t)key * 2654435761u) & (uint32_t)(m->cap - 1);
  while (m->k[h] >= 0) { if (m->k[h] == key) { *out = m->v[h]; return 1; } h = (h + 1) & (uint32_t)(m->cap - 1); }
  return 0;
}
static int32_t pow2_at_least(int32_t n) { int32_t c = 8; while (c < 2 * n) c <<= 1; return c; }

/* CU:98-117 calculateSimilarity(vector1 = candidate, vector2 = query) */
static double calc_similarity_hashjoin(const ovec_t *v1, const ovec_t *v2) {
  int32_t c1 = pow2_at_least(v1->nnz), c2 = pow2_at_least(v2->nnz);
  int32_t *kb = (int32_t*)malloc(sizeof(int32_t) * (size_t)(c1 + c2));
  double *vb = (double*)malloc(sizeof(double) * (size_t)(c1 + c2));
  smap_t m1, m2;
  smap_build(&m1, v1->idx, v1->val, v1->nnz, kb, vb, c1);           /* CU:102-105 */
  smap_build(&m2, v2->idx, v2->val, v2->nnz, kb + c1, vb + c1, c2); /* CU:106-109 */
  double similarity = 0.0;                                          /* CU:101 */
  for (int32_t i = 0; i < v1->nnz; i++) {                           /* CU:110 (ascending order here) */
    double a = 0.0, b = 0.0;
    smap_get(&m1, v1->idx[i], &a);
    if (smap_get(&m2, v1->idx[i], &b)) { double p = a * b; similarity = similarity + p; }  /* CU:111-114 */
  }
  free(kb); free(vb);
  return similarity;
}

/* same value by sorted merge; used by the brute-force helper */
static double dot_merge(const int32_t *ia, const double *va, int32_t na, const int32_t *ib, const double *vb, int32_t nb, int32_t *n_shared) {
  double s = 0.0; int32_t i = 0, j = 0, sh = 0;
  while (i < na && j < nb) {
    if (ia[i] < ib[j]) i++; else if (ia[i] > ib[j]) j++;
    else { double p = va[i] * vb[j]; s = s + p; sh++; i++; j++; }
  }
  if (n_shared) *n_shared = sh;
  return s;
}

/* Scala 2.10.4 immutable.HashSet.improve (hashing of an Int key; UNVERIFIED recall) */
static uint32_t scala_improve(uint32_t hcode) {
  uint32_t h = hcode + ~(hcode << 9);
  h = h ^ (h >> 14);
  h = h + (h << 4);
  return h ^ (h >> 10);
}
/* trie iteration key: 5-bit digits from the least-significant end, compared lexicographically */
static uint64_t scala_trie_key(int32_t x) {
  uint32_t h = scala_improve((uint32_t)x);
  uint64_t k = 0;
  for (int lvl = 0; lvl < 7; lvl++) { k = (k << 5) | ((h >> (5 * lvl)) & 31u); }
  return k;
}
typedef struct { uint64_t k; int32_t d; } kd_t;
static int kd_cmp(const void *a, const void *b) { uint64_t x = ((const kd_t*)a)->k, y = ((const kd_t*)b)->k; return x < y ? -1 : x > y; }

/* Iteration order of `sparseVector.indices.toSet.filter(...)` (WWA:172, EPA:43, IWA:102):
 * a Set built from <= 4 ascending ints is a Set1..Set4 (insertion order); from >= 5 it is a
 * HashSet whose filter stays a HashSet (hash-trie order).  n_total = nnz of the whole vector. */
void oracle_set_iteration_order(int32_t mode, int32_t n_total, const int32_t *dims, int32_t n, int32_t *out) {
  if (mode == 1 || n_total <= 4) { memcpy(out, dims, sizeof(int32_t) * (size_t)n); return; }
  kd_t *t = (kd_t*)malloc(sizeof(kd_t) * (size_t)(n ? n : 1));
  for (int32_t i = 0; i < n; i++) { t[i].k = scala_trie_key(dims[i]); t[i].d = dims[i]; }
  qsort(t, (size_t)n, sizeof(kd_t), kd_cmp);
  for (int32_t i = 0; i < n; i++) out[i] = t[i].d;
  free(t);
}

/* ------------------------------------------------------------------ lifecycle */

oracle_t *oracle_create(int32_t dim, double sim_thr, double idx_thr, int32_t semantics, int32_t algo,
                        int32_t max_shard, int32_t max_entry, int32_t max_index_actor, int32_t set_order,
                        const double *maxw) {
  oracle_t *o = (oracle_t*)calloc(1, sizeof(oracle_t));
  o->dim = dim; o->sim_thr = sim_thr; o->idx_thr = idx_thr; o->semantics = semantics; o->algo = algo;
  o->max_shard = max_shard < 1 ? 1 : max_shard; o->max_entry = max_entry < 1 ? 1 : max_entry;
  o->max_index_actor = max_index_actor < 1 ? 1 : max_index_actor; o->set_order = set_order;
  if (maxw) { o->maxw = (double*)malloc(sizeof(double) * (size_t)dim); memcpy(o->maxw, maxw, sizeof(double) * (size_t)dim); }
  if (algo == ORC_ALGO_FAITHFUL) {
    /* a worker = (shardId, child) : shardId = d % maxShardNum (WWA:172,197); the EntryProxyActor
     * instance is addressed by (shard = shardId, entry = shardId % maxEntryNum) (CU:28-40), i.e.
     * one per shardId; child = d % maxIndexEntryActorNum (EPA:43). */
    o->n_workers = o->max_shard * o->max_index_actor;
    o->workers = (oworker_t*)calloc((size_t)o->n_workers, sizeof(oworker_t));
  } else {
    o->fp_ids = (ivec_t*)calloc((size_t)dim, sizeof(ivec_t));
    o->fp_w = (double**)calloc((size_t)dim, sizeof(double*));
    o->fp_wcap = (int32_t*)calloc((size_t)dim, sizeof(int32_t));
  }
  o->threads = 1;
  return o;
}

void oracle_set_threads(oracle_t *o, int32_t t) { o->threads = t < 1 ? 1 : t; }

/* Exact index reduction, same rule as the GPU library (include/apss.h `pruning`): a vector keeps out of
 * the index the longest prefix, in (document frequency descending, dim ascending) order, whose squared
 * weights sum to <= alpha * (t / max_query_norm)^2 * (1 - 2^-20).  Cauchy-Schwarz: the un-indexed part can
 * contribute < t to any dot product with a query of norm <= max_query_norm, so every pair with dot >= t
 * still shares an indexed component.  Returns 0, or -1 for a bad argument / wrong algorithm. */
int32_t oracle_set_pruning(oracle_t *o, double alpha, double max_qnorm) {
  if (o->algo != ORC_ALGO_FAST || o->n_vecs) return -1;
  if (alpha == 0.0) alpha = 0.8;
  if (max_qnorm == 0.0) max_qnorm = 1.0;
  if (!(alpha > 0.0 && alpha < 1.0) || !(max_qnorm > 0.0)) return -1;
  double t = o->sim_thr;
  o->prune = 1; o->max_qnorm = max_qnorm;
  o->prune_lim = t > 0.0 ? alpha * t * t / (max_qnorm * max_qnorm) * (1.0 - ldexp(1.0, -20)) : 0.0;
  o->df = (int32_t*)calloc((size_t)o->dim, sizeof(int32_t));
  return 0;
}
int64_t oracle_n_unindexed(const oracle_t *o) { return o->n_unindexed; }
int32_t oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_destroy(oracle_t *o) {
  if (!o) return;
  for (int64_t i = 0; i < o->n_vecs; i++) { free(o->vecs[i].idx); free(o->vecs[i].val); free(o->vecs[i].skip); }
  free(o->vecs); free(o->df);
  for (int32_t w = 0; w < o->n_workers; w++) {
    oworker_t *wk = &o->workers[w];
    free(wk->store_vec.v); map64_free(&wk->dim2list);
    for (int32_t l = 0; l < wk->n_lists; l++) free(wk->lists[l].v);
    free(wk->lists);
  }
  free(o->workers);
  if (o->fp_ids) { for (int32_t d = 0; d < o->dim; d++) { free(o->fp_ids[d].v); free(o->fp_w[d]); } }
  free(o->fp_ids); free(o->fp_w); free(o->fp_wcap);
  free(o->pairs); free(o->status); free(o->maxw);
  free(o);
}

void oracle_freeze(oracle_t *o) { o->frozen = 1; }                  /* IWA:143-144 */
const char *oracle_last_error(oracle_t *o) { return o->err; }

/* ------------------------------------------------------------------ per-vector steps */

/* EPA:81-93 with EPA:51-57: admit iff sum over dims d < vectorDim of maxw(d)*v(d) >= t.
 * maxWeightMap only holds keys [0, vectorDim) so other dims contribute nothing.  Runs on the
 * UN-pruned vector (EPA:97 precedes WWA).  Ascending-index fp64 sum. */
static int admit_vector(const oracle_t *o, const int32_t *idx, const double *val, int32_t n) {
  double s = 0.0;
  for (int32_t i = 0; i < n; i++) {
    if (idx[i] < 0 || idx[i] >= o->dim) continue;
    double mw = o->maxw ? o->maxw[idx[i]] : 1.0;
    double p = mw * val[i];
    s = s + p;
  }
  return s >= o->sim_thr;
}

/* WWA:185-202: keep value > indexThreshold (strict, WWA:192), sorted by index (WWA:193). */
static void prune_vector(const oracle_t *o, const int32_t *idx, const double *val, int32_t n, ovec_t *out) {
  int32_t m = 0;
  for (int32_t i = 0; i < n; i++) if (val[i] > o->idx_thr) m++;
  out->nnz = m;
  out->idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(m ? m : 1));
  out->val = (double*)malloc(sizeof(double) * (size_t)(m ? m : 1));
  m = 0;
  for (int32_t i = 0; i < n; i++) if (val[i] > o->idx_thr) { out->idx[m] = idx[i]; out->val[m] = val[i]; m++; }
}

static void push_pair(opair_t **buf, int64_t *n, int64_t *cap, opair_t p) {
  if (*n == *cap) { *cap = *cap ? *cap * 2 : 64; *buf = (opair_t*)realloc(*buf, sizeof(opair_t) * (size_t)*cap); }
  (*buf)[(*n)++] = p;
}

/* ------------------------------------------------------------------ faithful worker */

/* a wrapper as delivered to one worker: SparseVectorWrapper(indices routed here, (id, vector)) */
typedef struct { int32_t ord; int32_t *dims; int32_t n_dims; } owrap_t;

/* IWA:61-71 */
static void worker_build_index(oworker_t *wk, const owrap_t *ws, int32_t n) {
  for (int32_t i = 0; i < n; i++) {
    ivec_push(&wk->store_vec, ws[i].ord);                         /* IWA:64 */
    int32_t current_idx = wk->store_vec.n - 1;                    /* IWA:65 */
    for (int32_t j = 0; j < ws[i].n_dims; j++) {                  /* IWA:66 */
      int64_t slot = map64_find(&wk->dim2list, ws[i].dims[j]);
      int32_t li;
      if (slot < 0) {                                             /* getOrElseUpdate, IWA:67 */
        if (wk->n_lists == wk->cap_lists) { wk->cap_lists = wk->cap_lists ? wk->cap_lists * 2 : 16; wk->lists = (ivec_t*)realloc(wk->lists, sizeof(ivec_t) * (size_t)wk->cap_lists); }
        li = wk->n_lists++; memset(&wk->lists[li], 0, sizeof(ivec_t));
        map64_put(&wk->dim2list, ws[i].dims[j], li);
      } else li = (int32_t)wk->dim2list.val[slot];
      ivec_push(&wk->lists[li], current_idx);                     /* IWA:67-68 */
    }
  }
}

/* IWA:74-111 for one query wrapper.  `sims` is this query's outputSimSet entry (ckey -> pair
 * slot) and *has_entry says whether outputSimSet.contains(q.id). */
/* `stripe` of `nstripes`: the threads of the parallel driver may share ONE query by candidate key (key % nstripes); a
 * candidate's visits all fall into the same stripe in the same order, so pairs, dot calls and the first-list skip are
 * what the serial loop gives (the posting count is taken by stripe 0).  Pure parallelisation of the restatement. */
static void worker_query_one(const oracle_t *o, const oworker_t *wk, const owrap_t *q, map64_t *sims, int *has_entry,
                             opair_t **out, int64_t *n_out, int64_t *cap_out, int64_t *postings, int64_t *dot_calls,
                             int32_t stripe, int32_t nstripes) {
  const ovec_t *qv = &o->vecs[q->ord];
  for (int32_t j = 0; j < q->n_dims; j++) {                       /* IWA:102 (Set iteration order) */
    int64_t slot = map64_find(&wk->dim2list, q->dims[j]);         /* IWA:104; Q15: missing dim -> empty list */
    int64_t list_start = *n_out;
    if (slot >= 0) {
      const ivec_t *list = &wk->lists[wk->dim2list.val[slot]];
      if (stripe == 0) *postings += list->n;
      for (int32_t p = 0; p < list->n; p++) {                     /* IWA:86 */
        int32_t c_ord = wk->store_vec.v[list->v[p]];              /* IWA:87 */
        const ovec_t *cv = &o->vecs[c_ord];
        if (nstripes > 1 && (int32_t)((uint64_t)cv->key % (uint64_t)nstripes) != stripe) continue;
        if (*has_entry && map64_find(sims, cv->key) < 0 && qv->key != cv->key) {   /* IWA:89-91 */
          double sim = calc_similarity_hashjoin(cv, qv);          /* IWA:92 */
          (*dot_calls)++;
          if (sim >= o->sim_thr) {                                /* IWA:93 */
            opair_t pr; pr.qkey = qv->key; pr.ckey = cv->key; pr.q = q->ord; pr.c = c_ord; pr.sim = sim;
            push_pair(out, n_out, cap_out, pr);                   /* IWA:94 (similarityHashMap) */
          }
        }
      }
    }
    /* IWA:106-107: getOrElseUpdate(q.id, new) ++= similarVectors */
    *has_entry = 1;
    for (int64_t k = list_start; k < *n_out; k++) map64_put(sims, (*out)[k].ckey, k);
  }
}

/* ------------------------------------------------------------------ fast (weighted) index */

typedef struct { int32_t df, dim, pos; } rank_t;
static int rank_cmp(const void *a, const void *b) {
  const rank_t *x = (const rank_t*)a, *y = (const rank_t*)b;
  if (x->df != y->df) return x->df > y->df ? -1 : 1;
  return x->dim < y->dim ? -1 : x->dim > y->dim;
}
static void prune_select_vector(oracle_t *o, ovec_t *v) {
  rank_t *r = (rank_t*)malloc(sizeof(rank_t) * (size_t)(v->nnz ? v->nnz : 1));
  for (int32_t i = 0; i < v->nnz; i++) { r[i].df = o->df[v->idx[i]]; r[i].dim = v->idx[i]; r[i].pos = i; }
  qsort(r, (size_t)v->nnz, sizeof(rank_t), rank_cmp);
  v->skip = (uint8_t*)calloc((size_t)(v->nnz ? v->nnz : 1), 1);
  double s = 0.0;
  for (int32_t k = 0; k < v->nnz; k++) {
    double x = v->val[r[k].pos];
    double s2 = s + x * x;
    if (!(s2 <= o->prune_lim)) break;
    s = s2; v->skip[r[k].pos] = 1; o->n_unindexed++;
  }
  free(r);
}

static void fast_index_vector(oracle_t *o, int32_t ord) {
  const ovec_t *v = &o->vecs[ord];
  for (int32_t i = 0; i < v->nnz; i++) {
    int32_t d = v->idx[i];
    if (v->skip && v->skip[i]) continue;
    ivec_t *l = &o->fp_ids[d];
    if (l->n == o->fp_wcap[d]) {
      int32_t nc = o->fp_wcap[d] ? o->fp_wcap[d] * 2 : 4;
      o->fp_w[d] = (double*)realloc(o->fp_w[d], sizeof(double) * (size_t)nc); o->fp_wcap[d] = nc;
    }
    o->fp_w[d][l->n] = v->val[i];
    ivec_push(l, ord);
  }
}

/* ------------------------------------------------------------------ batch entry point */

/* One insertNewVector batch = one IndexData per worker (parity configuration P0 of SURVEY 8(a)
 * makes batch boundaries explicit).  Steps: EPA:81-93 admit -> WWA:185-202 prune -> routing
 * -> IWA:125-127 index (unless frozen / query-only) -> IWA:128-132 query.
 * keys == NULL: key = ordinal.  Returns 0 or a negative error (Q9: all-or-nothing validation). */
int32_t oracle_insert_batch(oracle_t *o, int32_t n, const int64_t *indptr, const int32_t *indices, const double *values,
                            const int64_t *keys, int32_t flags) {
  /* validation mirrors SparseVector.scala:96-108 (strictly increasing, < size) */
  for (int32_t v = 0; v < n; v++) {
    if (indptr[v + 1] < indptr[v]) { snprintf(o->err, sizeof o->err, "indptr not monotone at %d", v); return -2; }
    int32_t prev = -1;
    for (int64_t p = indptr[v]; p < indptr[v + 1]; p++) {
      if (indices[p] <= prev || indices[p] >= o->dim) { snprintf(o->err, sizeof o->err, "bad index in vector %d", v); return -3; }
      prev = indices[p];
    }
  }
  int query_only = (flags & ORC_FLAG_QUERY_ONLY) || o->frozen;
  int index_only = (flags & ORC_FLAG_INDEX_ONLY) && !query_only;
  int64_t base = o->n_vecs;
  o->id_base = base;
  o->n_pairs = 0; o->postings_visited = 0; o->candidates_unique = 0; o->dot_calls_ref = 0;
  o->status = (uint8_t*)realloc(o->status, (size_t)(n ? n : 1)); o->n_status = n;

  /* vectors are appended to the table even in query-only mode (then removed again below) */
  if (o->n_vecs + n > o->cap_vecs) { o->cap_vecs = (o->n_vecs + n) * 2; o->vecs = (ovec_t*)realloc(o->vecs, sizeof(ovec_t) * (size_t)o->cap_vecs); }
  for (int32_t v = 0; v < n; v++) {
    const int32_t *idx = indices + indptr[v]; const double *val = values + indptr[v];
    int32_t nn = (int32_t)(indptr[v + 1] - indptr[v]);
    ovec_t *ov = &o->vecs[base + v];
    ov->key = keys ? keys[v] : base + v;
    ov->skip = NULL;
    if (!(flags & ORC_FLAG_SKIP_ADMIT) && !admit_vector(o, idx, val, nn)) { o->status[v] = ORC_ST_REJECTED; ov->nnz = 0; ov->idx = NULL; ov->val = NULL; continue; }
    prune_vector(o, idx, val, nn, ov);
    o->status[v] = ov->nnz ? ORC_ST_ACTIVE : ORC_ST_EMPTY;
  }
  if (o->prune) {   /* the bound needs |q| <= max_query_norm for every query: refuse the batch otherwise */
    int bad = 0;
    for (int32_t v = 0; v < n && !bad; v++) {
      const ovec_t *ov = &o->vecs[base + v]; double sq = 0.0;
      for (int32_t i = 0; i < ov->nnz; i++) sq += ov->val[i] * ov->val[i];
      if (sq > o->max_qnorm * o->max_qnorm * (1.0 + 1e-9)) bad = 1;
    }
    if (bad) {
      for (int32_t v = 0; v < n; v++) { free(o->vecs[base + v].idx); free(o->vecs[base + v].val); }
      snprintf(o->err, sizeof o->err, "pruning: vector norm exceeds max_query_norm"); return -4;
    }
  }
  o->n_vecs += n;

  if (o->algo == ORC_ALGO_FAITHFUL) {
    /* routing: WWA:164-183 then EPA:37-49.  For every shardId holding >= 1 dim of v the whole
     * vector goes to every child 0..maxIndexEntryActorNum-1 of that shard's entry actor, with
     * the (possibly empty) subset of dims d % maxIndexEntryActorNum == child. */
    int32_t nw = o->n_workers;
    owrap_t **wr = (owrap_t**)calloc((size_t)nw, sizeof(owrap_t*));
    int32_t *wn = (int32_t*)calloc((size_t)nw, sizeof(int32_t)), *wc = (int32_t*)calloc((size_t)nw, sizeof(int32_t));
    for (int32_t v = 0; v < n; v++) {
      if (o->status[v] != ORC_ST_ACTIVE) continue;
      const ovec_t *ov = &o->vecs[base + v];
      int32_t *ordered = (int32_t*)malloc(sizeof(int32_t) * (size_t)ov->nnz);
      int32_t *sub = (int32_t*)malloc(sizeof(int32_t) * (size_t)ov->nnz);
      for (int32_t s = 0; s < o->max_shard; s++) {
        int32_t ns = 0;
        for (int32_t i = 0; i < ov->nnz; i++) if (ov->idx[i] % o->max_shard == s) sub[ns++] = ov->idx[i];
        if (!ns) continue;                                         /* WWA:173 */
        for (int32_t ch = 0; ch < o->max_index_actor; ch++) {      /* EPA:41 */
          int32_t w = s * o->max_index_actor + ch;
          int32_t nd = 0;
          for (int32_t i = 0; i < ns; i++) if (sub[i] % o->max_index_actor == ch) ordered[nd++] = sub[i];
          if (wn[w] == wc[w]) { wc[w] = wc[w] ? wc[w] * 2 : 16; wr[w] = (owrap_t*)realloc(wr[w], sizeof(owrap_t) * (size_t)wc[w]); }
          owrap_t *x = &wr[w][wn[w]++];
          x->ord = (int32_t)(base + v); x->n_dims = nd;
          x->dims = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nd ? nd : 1));
          oracle_set_iteration_order(o->set_order, ov->nnz, ordered, nd, x->dims);
        }
      }
      free(ordered); free(sub);
    }
    /* IWA:125-127 : the whole batch is indexed before any of it is queried */
    if (!query_only) for (int32_t w = 0; w < nw; w++) worker_build_index(&o->workers[w], wr[w], wn[w]);
    /* IWA:128-132 */
    int64_t postings = 0, dots = 0;
    int nthreads = o->threads;
    /* duplicate keys inside a batch share an outputSimSet entry (IWA:89,106): run serially then */
    int dup_keys = 0;
    if (keys) { map64_t seen; map64_init(&seen, n); for (int32_t v = 0; v < n && !dup_keys; v++) { if (map64_find(&seen, keys[v]) >= 0) dup_keys = 1; map64_put(&seen, keys[v], 1); } map64_free(&seen); }
    if (dup_keys) nthreads = 1;
    for (int32_t w = 0; w < nw && !index_only; w++) {
      const oworker_t *wk = &o->workers[w];
      int32_t nq = wn[w];
      if (!nq) continue;
      if (nthreads == 1) {
        /* literal: one outputSimSet keyed by q.id for the whole IndexData */
        map64_t key2slot; map64_init(&key2slot, nq);
        map64_t *sims = (map64_t*)calloc((size_t)nq, sizeof(map64_t)); int *has = (int*)calloc((size_t)nq, sizeof(int)); int32_t n_ent = 0;
        for (int32_t i = 0; i < nq; i++) {
          int64_t qkey = o->vecs[wr[w][i].ord].key;
          int64_t s = map64_find(&key2slot, qkey); int32_t e;
          if (s < 0) { e = n_ent++; map64_put(&key2slot, qkey, e); has[e] = (o->semantics == ORC_R1); } else e = (int32_t)key2slot.val[s];
          /* pair slots index into o->pairs: keep per-entry map consistent by rebuilding offsets */
          int64_t before = o->n_pairs;
          worker_query_one(o, wk, &wr[w][i], &sims[e], &has[e], &o->pairs, &o->n_pairs, &o->cap_pairs, &postings, &dots, 0, 1);
          (void)before;
        }
        for (int32_t e = 0; e < n_ent; e++) map64_free(&sims[e]);
        free(sims); free(has); map64_free(&key2slot);
      } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
        {
          opair_t *lp = NULL; int64_t ln = 0, lc = 0, lpost = 0, ldots = 0;
          map64_t sims; memset(&sims, 0, sizeof sims);
          /* few queries, many threads (bench samples against a large index): share each query by candidate key */
          const int32_t nstripes = nq >= 2 * nthreads ? 1 : (2 * nthreads + nq - 1) / nq;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1) nowait
#endif
          for (int64_t task = 0; task < (int64_t)nq * nstripes; task++) {
            const int32_t i = (int32_t)(task / nstripes), stripe = (int32_t)(task % nstripes);
            int has = (o->semantics == ORC_R1);
            map64_clear(&sims);
            /* sims maps ckey -> slot in lp; slots are relative to this query only for lookups */
            worker_query_one(o, wk, &wr[w][i], &sims, &has, &lp, &ln, &lc, &lpost, &ldots, stripe, nstripes);
          }
          map64_free(&sims);
#ifdef _OPENMP
#pragma omp critical
#endif
          {
            for (int64_t k = 0; k < ln; k++) push_pair(&o->pairs, &o->n_pairs, &o->cap_pairs, lp[k]);
            postings += lpost; dots += ldots;
          }
          free(lp);
        }
      }
    }
    o->postings_visited = postings; o->dot_calls_ref = dots;
    o->candidates_unique = -1;       /* defined on R1 only; use ALGO_FAST or oracle_bruteforce */
    for (int32_t w = 0; w < nw; w++) { for (int32_t i = 0; i < wn[w]; i++) free(wr[w][i].dims); free(wr[w]); }
    free(wr); free(wn); free(wc);
  } else {
    /* ---- ALGO_FAST: R1 by accumulation over weighted postings (same arithmetic order as
     * the ascending-index dot: acc starts at 0.0, products added in ascending dim). */
    if (!query_only && o->prune) {     /* document frequencies include the whole batch, then rank and mark */
      for (int32_t v = 0; v < n; v++) { const ovec_t *ov = &o->vecs[base + v]; if (o->status[v] == ORC_ST_ACTIVE) for (int32_t i = 0; i < ov->nnz; i++) o->df[ov->idx[i]]++; }
      for (int32_t v = 0; v < n; v++) if (o->status[v] == ORC_ST_ACTIVE) prune_select_vector(o, &o->vecs[base + v]);
    }
    if (!query_only) for (int32_t v = 0; v < n; v++) if (o->status[v] == ORC_ST_ACTIVE) fast_index_vector(o, (int32_t)(base + v));
    int64_t nvis = query_only ? base : o->n_vecs;      /* candidates are ordinals < nvis */
    int64_t postings = 0, cands = 0;
    int nthreads = o->threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) reduction(+:postings, cands) if(!index_only)
#endif
    if (!index_only) {
      double *acc = (double*)calloc((size_t)(nvis ? nvis : 1), sizeof(double));
      uint8_t *touched = (uint8_t*)calloc((size_t)(nvis ? nvis : 1), 1);
      int32_t *tl = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nvis ? nvis : 1));
      opair_t *lp = NULL; int64_t ln = 0, lc = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4) nowait
#endif
      for (int32_t v = 0; v < n; v++) {
        if (o->status[v] != ORC_ST_ACTIVE) continue;
        const ovec_t *qv = &o->vecs[base + v];
        int64_t nt = 0;
        for (int32_t i = 0; i < qv->nnz; i++) {
          int32_t d = qv->idx[i]; double wq = qv->val[i];
          const ivec_t *l = &o->fp_ids[d]; const double *lw = o->fp_w[d];
          postings += l->n;
          for (int32_t p = 0; p < l->n; p++) {
            int32_t c = l->v[p];
            double pr = wq * lw[p];
            if (!touched[c]) { touched[c] = 1; tl[nt++] = c; acc[c] = 0.0; }
            acc[c] = acc[c] + pr;
          }
        }
        for (int64_t k = 0; k < nt; k++) {
          int32_t c = tl[k]; touched[c] = 0;
          const ovec_t *cv = &o->vecs[c];
          if (cv->key == qv->key) continue;                        /* IWA:91 */
          cands++;
          double sim = acc[c];
          if (o->prune) { int32_t ns; sim = dot_merge(cv->idx, cv->val, cv->nnz, qv->idx, qv->val, qv->nnz, &ns); }   /* acc holds the indexed part only */
          if (sim >= o->sim_thr) {                              /* IWA:93 */
            opair_t pr; pr.qkey = qv->key; pr.ckey = cv->key; pr.q = (int32_t)(base + v); pr.c = c; pr.sim = sim;
            push_pair(&lp, &ln, &lc, pr);
          }
        }
      }
#ifdef _OPENMP
#pragma omp critical
#endif
      { for (int64_t k = 0; k < ln; k++) push_pair(&o->pairs, &o->n_pairs, &o->cap_pairs, lp[k]); }
      free(lp); free(acc); free(touched); free(tl);
    }
    o->postings_visited = postings; o->candidates_unique = cands; o->dot_calls_ref = -1;
  }

  if (query_only) {   /* the batch was never stored (IWA:125): drop it from the table again */
    for (int64_t i = base; i < o->n_vecs; i++) { free(o->vecs[i].idx); free(o->vecs[i].val); free(o->vecs[i].skip); }
    o->n_vecs = base;
  }
  o->tot_postings_visited += o->postings_visited;
  if (o->candidates_unique >= 0) o->tot_candidates_unique += o->candidates_unique;
  if (o->dot_calls_ref >= 0) o->tot_dot_calls_ref += o->dot_calls_ref;
  o->tot_pairs += o->n_pairs;
  return 0;
}

/* ------------------------------------------------------------------ result access */

int64_t oracle_n_pairs(const oracle_t *o) { return o->n_pairs; }
int64_t oracle_id_base(const oracle_t *o) { return o->id_base; }
int64_t oracle_n_vectors(const oracle_t *o) { return o->n_vecs; }
/* q is returned relative to the batch (q - id_base); c is the global ordinal. */
void oracle_fetch_pairs(const oracle_t *o, int32_t *q, int32_t *c, int64_t *qkey, int64_t *ckey, double *sim) {
  for (int64_t i = 0; i < o->n_pairs; i++) {
    if (q) q[i] = (int32_t)(o->pairs[i].q - o->id_base);
    if (c) c[i] = o->pairs[i].c;
    if (qkey) qkey[i] = o->pairs[i].qkey;
    if (ckey) ckey[i] = o->pairs[i].ckey;
    if (sim) sim[i] = o->pairs[i].sim;
  }
}
void oracle_fetch_status(const oracle_t *o, uint8_t *st) { memcpy(st, o->status, (size_t)o->n_status); }
/* out[0..6] = postings_visited, candidates_unique, dot_calls_ref (last batch), then totals, tot_pairs */
void oracle_counters(const oracle_t *o, int64_t *out) {
  out[0] = o->postings_visited; out[1] = o->candidates_unique; out[2] = o->dot_calls_ref;
  out[3] = o->tot_postings_visited; out[4] = o->tot_candidates_unique; out[5] = o->tot_dot_calls_ref; out[6] = o->tot_pairs;
}

/* ------------------------------------------------------------------ brute force helper */

/* O(nq * nc * nnz) check used by the tests: for every query row and candidate row with at least
 * one shared dim (and differing key), count it and report it if dot >= thr.  first_dim (nullable,
 * per query) applies the R0 rule for the single-worker configuration P0: drop the pair when the
 * shared dims are a subset of {first_dim[q]}.  Returns number of pairs written (<= cap). */
int64_t oracle_bruteforce(int32_t nq, const int64_t *qptr, const int32_t *qidx, const double *qval, const int64_t *qkey,
                          int32_t nc, const int64_t *cptr, const int32_t *cidx, const double *cval, const int64_t *ckey,
                          double thr, const int32_t *first_dim,
                          int32_t *out_q, int32_t *out_c, double *out_sim, int64_t cap, int64_t *n_candidates) {
  int64_t np = 0, ncand = 0;
  for (int32_t q = 0; q < nq; q++) for (int32_t c = 0; c < nc; c++) {
    if (qkey[q] == ckey[c]) continue;
    int32_t sh = 0;
    double s = dot_merge(cidx + cptr[c], cval + cptr[c], (int32_t)(cptr[c + 1] - cptr[c]),
                         qidx + qptr[q], qval + qptr[q], (int32_t)(qptr[q + 1] - qptr[q]), &sh);
    if (!sh) continue;
    ncand++;
    if (first_dim) {
      int32_t nonfirst = 0, i = 0, j = 0;
      const int32_t *ia = cidx + cptr[c], *ib = qidx + qptr[q];
      int32_t na = (int32_t)(cptr[c + 1] - cptr[c]), nb = (int32_t)(qptr[q + 1] - qptr[q]);
      while (i < na && j < nb) { if (ia[i] < ib[j]) i++; else if (ia[i] > ib[j]) j++; else { if (ia[i] != first_dim[q]) nonfirst++; i++; j++; } }
      if (!nonfirst) continue;
    }
    if (s >= thr) { if (np < cap) { out_q[np] = q; out_c[np] = c; out_sim[np] = s; } np++; }
  }
  if (n_candidates) *n_candidates = ncand;
  return np;
}
