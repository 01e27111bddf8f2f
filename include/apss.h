/*
 * apss.h -- C ABI of the B200-native inverted-index all-pairs similarity scorer.
 *
 * This is the drop-in boundary for ONE path of mcgill-cpslab/all-pairs-similarity: what an
 * IndexingWorkerActor does with an IndexData message (index the batch, score it against every
 * indexed vector, threshold, emit similar pairs).  The reference has no FFI of its own; the seam
 * is its actor message protocol, so each entry point below names the reference code it replaces
 * (paths relative to core/src/main/scala/cpslab/ in the reference):
 *
 *   IWA = deploy/server/IndexingWorkerActor.scala      WWA = deploy/server/WriteWorkerActor.scala
 *   EPA = deploy/server/EntryProxyActor.scala          CU  = deploy/CommonUtils.scala
 *
 * A JVM host binds these through jni/apss_jni.c; this repo's own host side
 * binds them through ctypes.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Threading: a handle is single-caller but not thread-affine (an Akka actor handles one message at
 * a time on whatever dispatcher thread; IWA:122).  Every entry point selects the handle's device
 * itself.  Distinct handles may be used concurrently.
 *
 * Errors: every function returns APSS_OK (0) or a negative apss_status; nothing aborts or throws
 * across the boundary (the reference swallows exceptions per batch, IWA:124-137).  A batch is
 * all-or-nothing: one that fails validation never touches the index, and one that fails later (out of
 * memory, CUDA error) is rolled back before the call returns -- the index, the id counter and the
 * document frequencies are what they were.  Where a roll-back is impossible (the tile index rewrites
 * its open tile in place; a sticky CUDA error) the handle is retired: every later call returns
 * APSS_E_STATE and the only valid operation is apss_destroy.
 */
#ifndef APSS_H_
#define APSS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APSS_ABI_VERSION 3
#define APSS_MAX_DEVICES 16

typedef enum apss_status {
  APSS_OK = 0,
  APSS_E_INVALID = -1,   /* bad argument / config                                  */
  APSS_E_CUDA = -2,      /* CUDA runtime error (text in apss_last_error)           */
  APSS_E_NOMEM = -3,     /* device or host allocation failed                       */
  APSS_E_INPUT = -4,     /* batch failed validation (SparseVector.scala:96-108)    */
  APSS_E_STATE = -5,     /* call not valid in this state (e.g. fetch before batch) */
  APSS_E_NO_DEVICE = -6  /* no usable CUDA device: there is NO CPU fallback        */
} apss_status;

/* which result set is reported */
#define APSS_SEM_R1 0 /* specified semantics: every pair sharing >= 1 dim with dot >= t            */
#define APSS_SEM_R0 1 /* as-built: additionally drop pairs whose shared dims are all first(q)       */
                      /* (IWA:89 + IWA:106-107); needs first_dim[] from the host                    */

/* apss_insert_batch flags */
#define APSS_BATCH_QUERY_ONLY 1u  /* score but do not index (frozen index, IWA:125,143-144)         */
#define APSS_BATCH_DEVICE_PTRS 2u /* indptr/indices/values/ext_keys/first_dim are DEVICE pointers   */
#define APSS_BATCH_SKIP_ADMIT 4u  /* vectors already passed EPA:81-93 upstream (IndexData carries   */
                                  /* admitted, pruned vectors): do not re-apply the admission filter */
#define APSS_BATCH_INDEX_ONLY 8u  /* bulk load: IWA:61-71 only, no scoring (HBase LoadData-style     */
                                  /* warm start; not a path the reference's IndexData offers)        */

/* per-input-vector status (apss_fetch_status) */
#define APSS_ST_REJECTED 0 /* failed the admission filter (EPA:81-93)                               */
#define APSS_ST_EMPTY 1    /* admitted but value-pruned to nothing (WWA:192-199): stored, never sent */
#define APSS_ST_ACTIVE 2   /* indexed and queried                                                    */

typedef struct apss_handle apss_handle;

typedef struct apss_config {
  int32_t struct_size;         /* = sizeof(apss_config)                                              */
  int32_t dim;                 /* cpslab.allpair.vectorDim            (EPA:25, WWA:31)               */
  double similarity_threshold; /* cpslab.allpair.similarityThreshold  (IWA:23, EPA:24)               */
  double index_threshold;      /* cpslab.allpair.indexThreshold       (WWA:35); must be >= 0         */
  const double *max_weight;    /* dim entries, or NULL = 1.0 for every dim (the stub at EPA:51-57)   */
  int32_t device;              /* CUDA device ordinal                                                */
  int32_t semantics;           /* APSS_SEM_R1 / APSS_SEM_R0                                          */
  int32_t tile_vectors;        /* candidates per index tile; 0 = default (6144 for the default kernel)  */
  int32_t kernel_variant;      /* 0 = default.  Otherwise: bits 16-23 scoring kernel (3 dense-head [default],   */
                               /* 1 warp-per-(query,tile) row kernel, 2 query-block kernel), bits 8-15 warps per */
                               /* CTA, bits 24-31 queries per block, bits 0-7 unroll / candidates per thread.    */
                               /* All variants give identical results; they exist for measurement.               */
  int64_t reserve_vectors;     /* capacity hints; 0 = grow on demand                                 */
  int64_t reserve_nnz;
  int64_t reserve_pairs;
  /* Exact index reduction (the pruning the reference planned around its max-weight stub,  */
  /* EPA:51-57,81-93).  Off by default: with it on, the pair set and similarities are unchanged but        */
  /* postings_visited / candidates_unique count only what the reduced index makes the kernel touch.         */
  int32_t pruning;             /* 0 = off (parity counters); 1 = on, tile kernel on the reduced index;     */
                               /* 2 = on, candidate-major kernel (the batch is inverted instead and the    */
                               /* stored vectors are streamed); 3 = on, query-major posting-list traversal */
                               /* of dimension-sorted CSR posting segments with shared-memory hash         */
                               /* accumulators (the fastest; the reference's own loop order, IWA:101-104). */
                               /* 1, 2 and 3 give the same pairs, similarities and counters.               */
  int32_t reserved0;
  double prune_alpha;          /* share of (t / max_query_norm)^2 a vector may keep out of the index;      */
                               /* 0 = default 0.8; must be < 1                                             */
  double max_query_norm;       /* promise: every vector of every batch has L2 norm <= this after the value */
                               /* prune; 0 = default 1.0 (LoadGenerator.scala:35-37 normalises).  A batch  */
                               /* that breaks the promise is refused (APSS_E_INPUT), never mis-scored.      */
  /* Shard dispatch below the ABI (replaces the remote router EPA:37-49,113-122): with       */
  /* n_devices > 1 the ONE handle an actor holds owns all the listed GPUs.  The index is partitioned by          */
  /* internal-id range, block-cyclic (block = one insert batch, owner = batch number mod n_devices); every      */
  /* batch is copied to device_ids[0] once and fanned out over NVLink peer copies, each GPU scores it against   */
  /* its own shard (the owner also indexes it), and the per-shard pair lists are gathered into one result.      */
  int32_t n_devices;           /* 0 or 1: one GPU, `device`.  N > 1: device_ids[0..N-1] (then `device` is ignored) */
  int32_t device_ids[APSS_MAX_DEVICES];
  int32_t reserved1;
} apss_config;

typedef struct apss_batch_result {
  int64_t id_base;           /* internal id of the batch's first vector (IWA:64-65 `currentIdx`)    */
  int32_t n_vectors;         /* vectors in the call                                                  */
  int32_t n_rejected;        /* EPA:81-93                                                            */
  int32_t n_empty;           /* WWA:192-199                                                          */
  int32_t n_active;
  int64_t n_pairs;           /* pairs reported (after the R0 post-filter when semantics = R0)        */
  int64_t n_pairs_r1;        /* pairs with dot >= t before the R0 post-filter                        */
  int64_t n_prefilter;       /* fp32 guard-band survivors handed to the fp64 verify kernel           */
  int64_t postings_visited;  /* sum over query terms of visible posting-list lengths (IWA:86)        */
  int64_t candidates_unique; /* (q, c) with >= 1 shared dim, c.id != q.id: "candidate dot-products"  */
  int64_t work_items;        /* (query, index tile) items the scoring kernel processed; pruning = 2: stored   */
                             /* vectors that took the heavy pass                                             */
  double score_ms;           /* CUDA-event time of the scoring kernel(s), on the handle's stream     */
  double device_ms;          /* CUDA-event time first kernel -> last kernel of the call              */
  int64_t dense_postings;    /* default kernel: postings scored through the dense FFMA rows; the other    */
                             /* postings_visited - dense_postings went through shared-memory atomics      */
  int64_t dense_fma;         /* default kernel: FMAs the dense phase executed (zeros and padding included) */
} apss_batch_result;

typedef struct apss_stats {
  int64_t n_vectors;         /* vectors stored in this shard (all statuses)                          */
  int64_t n_postings;        /* postings resident in HBM                                             */
  int64_t n_tiles;           /* index tiles; pruning = 3: posting segments                           */
  int64_t bytes_postings;    /* 8 B each                                                             */
  int64_t bytes_directory;
  int64_t bytes_forward;     /* fp64 forward store used by the verify kernel                         */
  int64_t tot_postings_visited;
  int64_t tot_candidates_unique;
  int64_t tot_pairs;
  int64_t tot_prefilter;
  int64_t score_launches;    /* launches of the scoring kernel so far                                */
  int64_t kernel_launches;   /* all kernels launched by this handle so far                           */
  double tot_score_ms;
  int64_t phase_cycles[8];   /* profiling build (libapss_b200_prof.so) only, last batch, dense-head kernel:  */
                             /* thread-0 cycles summed over CTAs in [0] item setup, [1] look-ups + short     */
                             /* segments, [2] query-weight fill + dense FFMA, [3] queued segments, [4]        */
                             /* epilogue loop, [5] barrier + item fetch, [7] self-clear; [6] rare-path count  */
  int32_t frozen;
  int32_t tile_vectors;
  int32_t warps_per_cta;
  int32_t sm_count;
  int64_t n_unindexed;       /* stored components kept out of the index by exact index reduction     */
  int64_t segment_merges;    /* pruning = 3: posting-segment merges so far and the postings they copied     */
  int64_t merged_postings;
  int32_t n_devices;         /* GPUs behind this handle; the other fields are sums over the shards           */
  int32_t reserved;
} apss_stats;

int32_t apss_abi_version(void);

/* Construct one index worker on one GPU.  Replaces `new IndexingWorkerActor(conf)` (IWA:21-39,
 * swap point EPA:116) plus the per-vector filters configured on WWA/EPA.  Fails with
 * APSS_E_NO_DEVICE when CUDA is unusable. */
int32_t apss_create(const apss_config *cfg, apss_handle **out);
void apss_destroy(apss_handle *h);

/* One IndexData message = one call.  Replaces, in order: the admission filter EPA:81-93 (on the
 * un-pruned vector), the value prune WWA:185-202, buildInvertedIndex IWA:61-71 (skipped when
 * frozen / APSS_BATCH_QUERY_ONLY), querySimilarItems IWA:74-111 with calculateSimilarity
 * CU:98-117, and the threshold IWA:93.  Synchronous: results are ready on return.
 *   indptr[n+1], indices[indptr[n]] strictly increasing per vector and < dim, values fp64:
 *     the CSR form of Set[(String, SparkSparseVector)] (Message.scala:13); strings stay on the host.
 *   ext_keys[n] or NULL: one integer per distinct caller id string; vectors with equal keys are
 *     never paired (the string compare at IWA:91).  NULL = every vector distinct (its key is its
 *     internal id; callers that mix both must keep their keys disjoint from the internal ids).
 *   first_dim[n] or NULL: first(q) in Scala Set iteration order, only read when semantics = R0. */
int32_t apss_insert_batch(apss_handle *h, int32_t n, const int64_t *indptr, const int32_t *indices,
                          const double *values, const int64_t *ext_keys, const int32_t *first_dim,
                          uint32_t flags, apss_batch_result *out);

/* Copy the last batch's pairs out: q = index of the query within the batch, c = internal id of the
 * similar vector, sim = fp64 dot product.  This is SimilarityOutput.output (Message.scala:20-21,
 * sent at IWA:130) before the host maps ids back to strings.  Writes min(n_pairs, capacity). */
int32_t apss_fetch_pairs(apss_handle *h, int32_t *q, int32_t *c, double *sim, int64_t capacity, int64_t *n_out);

/* Device-resident view of the same arrays (valid until the next call on the handle). */
int32_t apss_pairs_device(apss_handle *h, const int32_t **q, const int32_t **c, const double **sim, int64_t *n);

/* Per-vector APSS_ST_* of the last batch (which queries get an output entry, IWA:106). */
int32_t apss_fetch_status(apss_handle *h, uint8_t *status, int32_t capacity);

/* ReceiveTimeout => stopUpdateIndex (IWA:143-144): every later batch is query-only. */
int32_t apss_freeze(apss_handle *h);

/* Internal id the next indexed vector will get (default: consecutive from 0).  Used by the shard
 * dispatcher, which owns the global id space when the index is sharded by id range. */
int32_t apss_set_next_id(apss_handle *h, int64_t next_id);

int32_t apss_get_stats(apss_handle *h, apss_stats *out);
const char *apss_last_error(apss_handle *h);

/* The cudaStream_t all work of this handle is enqueued on (for CUDA-event timing by the caller). */
void *apss_stream(apss_handle *h);

/* Micro-benchmark of shared-memory accumulator update primitives (plain RMW, fixed-point atomics,
 * float CAS atomics; mode 6: register FP32 FMA), used to choose the accumulator design and as the measured on-chip
 * peaks of bench.py's roofline; reports updates (mode 6: FMAs) per second. */
int32_t apss_microbench_accumulators(int32_t device, int32_t mode, int32_t warps, int32_t iters, double *updates_per_sec);

#ifdef __cplusplus
}
#endif
#endif /* APSS_H_ */
