// apss_kernels.cuh -- device code of the B200 (sm_100a) all-pairs similarity scorer.
//
// Data layout in HBM (one shard = one GPU):
//   forward store   fwd_ptr[n+1] (int64), fwd_idx[nnz] (int32), fwd_val[nnz] (fp64): the pruned vectors
//                   exactly as WWA:193-194 stores them; read only by the fp64 verify kernel.
//   index tiles     the shard's vectors are cut into tiles of CR consecutive local ids.  Each tile
//                   owns (a) its postings, sorted by dimension then id, 8 B each:
//                   (id within tile : int32, weight : fp32)  and (b) a dense directory
//                   dir[tile][0..D] of int32 offsets into the tile's postings.  Tiles are
//                   append-only: an insert rebuilds at most the last, partially filled tile.
//   query batch     q_ptr[nq+1] (int32), q_dim[] (int32), q_w[] (fp32) + q_val[] (fp64 for verify).
//
// Reference code replaced (core/src/main/scala/cpslab/deploy/...):
//   k_prefilter_*   EntryProxyActor.scala:81-93 (admission) + WriteWorkerActor.scala:185-202 (prune)
//   k_emit_postings / k_build_dir    IndexingWorkerActor.scala:61-71 (buildInvertedIndex)
//   k_score         IndexingWorkerActor.scala:74-111 + CommonUtils.scala:98-117 (fp32 pre-filter)
//   k_verify        CommonUtils.scala:98-117 in fp64 + the threshold at IndexingWorkerActor.scala:93
#pragma once
#include <cassert>
#include <cstdint>
#include <cuda_runtime.h>

// APSS_DEBUG builds (libapss_b200_dbg.so) bounds-check every indexed access with device asserts.
#ifdef APSS_DEBUG
#define DBG_ASSERT(c) assert(c)
#else
#define DBG_ASSERT(c) ((void)0)
#endif

namespace apss {

enum Counter : int {
  C_PF = 0,        // fp32 guard-band survivors (prefilter records)
  C_POSTINGS = 1,  // postings visited
  C_CANDS = 2,     // touched accumulators (candidates_unique)
  C_FINAL = 3,     // pairs reported
  C_R1 = 4,        // pairs with dot >= t
  C_ERR = 5,       // validation error code (0 = ok)
  C_WORK = 6,      // persistent-kernel work cursor
  C_NREJ = 7, C_NEMPTY = 8, C_NACTIVE = 9, C_MAXNNZ = 10,
  C_MAXSQ = 11,    // max squared L2 norm of a pruned vector (bits of a non-negative double)
  C_SKIPPED = 12,  // components of this batch left out of the index by exact index reduction
  C_HEAVY = 13,    // candidate-major kernel: stored vectors deferred to the heavy pass (this launch)
  C_HEAVY_TOT = 14, // same, summed over the query slices of the batch
  C_TOTNNZ = 15,   // components kept by the value prune, whole batch (64-bit: the per-vector counts are int32)
  C_PHASE = 16,    // 8 per-phase cycle totals of the dense kernel (thread 0 of every CTA)
  C_DENSE_POST = 24, // dense-head kernel: postings scored through the dense FFMA rows (the rest went through shared-memory atomics)
  C_DENSE_FMA = 25,  // dense-head kernel: FMAs the dense phase executed (zeros and padding included)
  C_HOTN = 26,       // query-major pipelined kernel: entries reserved in the hot-candidate buffer (chunks of QP_CHUNK)
  C_COUNT = 27,      // device counters; the pinned host mirror has C_COUNT + 2 words:
  C_ITEMS = 27,      //   host only: (pieces << 36 | postings) of the batch, query-major kernel
  C_SCRATCH = 28     //   host only: one-word read-backs
};

static constexpr unsigned FULL = 0xffffffffu;
static constexpr unsigned NEG0 = 0x80000000u;   // accumulator "never touched" marker (-0.0f)
// fp32 weights are clamped to >= 2^-60 so that the product of any two is a positive normal float: a
// shared dimension always leaves a non-zero trace in the accumulator (exact candidate counts).
#define W_MIN 8.6736173798840355e-19f

__device__ __forceinline__ uint2 ld_stream(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// ------------------------------------------------------------------ K5: admission + value prune

// One WARP per input vector.  Validates (SparseVector.scala:96-108: strictly increasing indices,
// all < size), evaluates the admission predicate of EPA:81-93 on the UN-pruned vector in ascending
// index order (fp64, separate multiply and add: the lanes load 32 components at a time, the sum is
// then chained through them in lane order, the same chain on every lane) and counts the components
// that survive `value > indexThreshold` (strict; WWA:192).
__global__ void k_prefilter_count(int n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                  const double* __restrict__ val, int D, const double* __restrict__ maxw,
                                  double sim_thr, double idx_thr, int32_t* __restrict__ cnt,
                                  uint8_t* __restrict__ status, float* __restrict__ q_nrm, unsigned long long* counters) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v > n) return;
  if (v == n) { if (lane == 0) cnt[n] = 0; return; }
  const int64_t a = ptr[v], b = ptr[v + 1];
  if (b < a || (v == 0 && a != 0)) { if (lane == 0) { atomicMax(&counters[C_ERR], 2ULL); cnt[v] = 0; status[v] = 0; } return; }
  double s = 0.0, sq = 0.0; int kept = 0; int prev = -1; bool bad = false;
  for (int64_t p0 = a; p0 < b; p0 += 32) {
    const int64_t p = p0 + lane;
    const bool on = p < b;
    const int d = on ? idx[p] : 0x7fffffff;
    const double x = on ? val[p] : 0.0;
    int before = __shfl_up_sync(FULL, d, 1);
    if (lane == 0) before = prev;
    bad |= __any_sync(FULL, on && (d <= before || d >= D));
    if (bad) break;
    prev = __shfl_sync(FULL, d, 31);
    const bool keep = on && x > idx_thr;
    kept += __popc(__ballot_sync(FULL, keep));
    const double t = on ? __dmul_rn((maxw ? maxw[d] : 1.0), x) : 0.0;
    const int cntl = (int)min((int64_t)32, b - p0);
    for (int k = 0; k < cntl; ++k) {
      s = __dadd_rn(s, __shfl_sync(FULL, t, k));
      const double xk = __shfl_sync(FULL, x, k);
      if (xk > idx_thr) sq = fma(xk, xk, sq);
    }
  }
  if (lane) return;
  if (bad) { atomicMax(&counters[C_ERR], 3ULL); cnt[v] = 0; status[v] = 0; return; }
  uint8_t st;
  if (!(s >= sim_thr)) { st = 0; kept = 0; atomicAdd(&counters[C_NREJ], 1ULL); }
  else if (kept == 0) { st = 1; atomicAdd(&counters[C_NEMPTY], 1ULL); }
  else {
    st = 2; atomicAdd(&counters[C_NACTIVE], 1ULL); atomicMax(&counters[C_MAXNNZ], (unsigned long long)kept);
    atomicAdd(&counters[C_TOTNNZ], (unsigned long long)kept);
    atomicMax(&counters[C_MAXSQ], (unsigned long long)__double_as_longlong(sq));
  }
  status[v] = st; cnt[v] = kept;
  // upper bound of the L2 norm of the pruned vector (used only by the index-reduction bound)
  if (q_nrm) q_nrm[v] = st == 2 ? __double2float_ru(sqrt(sq) * (1.0 + 1e-9)) : 0.f;
}

// warp per vector: the surviving components, compacted in order
__global__ void k_prefilter_write(int n, const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                  const double* __restrict__ val, double idx_thr, const uint8_t* __restrict__ status,
                                  const int32_t* __restrict__ q_ptr, int32_t* __restrict__ q_dim,
                                  double* __restrict__ q_val, float* __restrict__ q_w) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n || status[v] != 2) return;
  int o = q_ptr[v];
  const int64_t a = ptr[v], b = ptr[v + 1];
  for (int64_t p0 = a; p0 < b; p0 += 32) {
    const int64_t p = p0 + lane;
    const double x = p < b ? val[p] : 0.0;
    const bool keep = p < b && x > idx_thr;
    const unsigned bal = __ballot_sync(FULL, keep);
    if (keep) { const int k = o + __popc(bal & ((1u << lane) - 1u)); q_dim[k] = idx[p]; q_val[k] = x; q_w[k] = fmaxf((float)x, W_MIN); }
    o += __popc(bal);
  }
}

__global__ void k_default_keys(int n, int64_t id_base, int64_t* __restrict__ keys) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) keys[v] = id_base + v;     // the default key of a vector is its internal id
}


// ------------------------------------------------------------------ exact index reduction

// Which components of a vector c stay OUT of the index: rank its components by document frequency
// (descending; ties by ascending dim) and take the longest prefix U whose squared weights sum to at most
// `lim` = alpha * (t / max query norm)^2.  For any query q with |q| <= max query norm,
// dot(q, c restricted to U) <= |q| |c_U| <= sqrt(alpha) t < t, so a pair with dot >= t always shares an
// indexed component, and the candidate test becomes  indexed part + |q| |c_U| >= t  (checked in the
// scoring kernel's epilogue; the fp64 verify kernel then computes the full dot product as before).
__global__ void k_df_update(int nnz, const int32_t* __restrict__ q_dim, int32_t* __restrict__ df, int delta) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) atomicAdd(df + q_dim[p], delta);
}

// sort key (row, max_df - df): a stable sort keeps ascending dims among equal frequencies
__global__ void k_rank_keys(int n, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim,
                            const int32_t* __restrict__ df, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= n) return;
  for (int p = q_ptr[v] + (threadIdx.x & 31); p < q_ptr[v + 1]; p += 32) {      // warp per vector
    keys[p] = ((unsigned long long)(unsigned)v << 31) | (unsigned long long)(0x7fffffff - df[q_dim[p]]);
    vals[p] = (unsigned long long)(unsigned)p;
  }
}

// one warp per vector: walk its components in rank order, fp64 running sum of squares (separate multiply and
// add, like the oracle; 32 components loaded at a time, the sum chained through them in lane order), mark the
// prefix, record |c_U| rounded up and the number of indexed components
// dfmin[v]: the smallest document frequency among the components left out (the last one of the prefix, the ranking
// being by descending frequency); INT_MAX when nothing was left out.  Used by the query-major kernels' candidate test.
__global__ void k_prune_mark(int n, const int32_t* __restrict__ q_ptr, const double* __restrict__ q_val,
                             const unsigned long long* __restrict__ ranked, const unsigned long long* __restrict__ ranked_keys, double lim,
                             uint8_t* __restrict__ skip, float* __restrict__ cu, int32_t* __restrict__ icnt, int32_t* __restrict__ dfmin,
                             unsigned long long* counters) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v > n) return;
  if (v == n) { if (lane == 0) icnt[n] = 0; return; }
  const int a = q_ptr[v], b = q_ptr[v + 1];
  double s = 0.0; int k = a; bool open = true;          // k: first component (in rank order) that stays indexed
  for (int p0 = a; p0 < b; p0 += 32) {
    const int p = p0 + lane;
    const int pos = p < b ? (int)ranked[p] : -1;
    const double x = p < b ? q_val[pos] : 0.0;
    const double xx = __dmul_rn(x, x);
    int nskip = 0;
    if (open) {
      const int cntl = min(32, b - p0);
      for (int j = 0; j < cntl; ++j) {
        const double s2 = __dadd_rn(s, __shfl_sync(FULL, xx, j));
        if (!(s2 <= lim)) { open = false; break; }
        s = s2; ++nskip;
      }
      k = p0 + nskip;
    }
    if (p < b) skip[pos] = lane < nskip ? 1 : 0;
  }
  if (lane) return;
  if (k > a) atomicAdd(&counters[C_SKIPPED], (unsigned long long)(k - a));
  cu[v] = s > 0.0 ? __double2float_ru(sqrt(s) * (1.0 + 1e-9)) : 0.f;
  icnt[v] = b - k;
  if (dfmin) dfmin[v] = k > a ? (int32_t)(0x7fffffffLL - (long long)(ranked_keys[k - 1] & 0x7fffffffULL)) : 0x7fffffff;
}

// k_rank_keys + the radix sort + k_prune_mark in one launch, for batches whose longest vector has <= PRL_MAX components
// (the usual case; longer ones take the global sort above).  One warp per vector: the ranking keys
// ((0x7fffffff - df) << 32 | position) go to shared memory, every component's rank is the number of smaller keys
// (positions are distinct, so the keys are; ascending position = ascending dim, the stable sort's tie order), the
// positions are scattered into rank order and the walk of k_prune_mark follows, arithmetic and order unchanged.
static constexpr int PRL_MAX = 512;       // components per vector the in-warp ranking takes
static constexpr int PRL_WARPS = 8;
__global__ void __launch_bounds__(PRL_WARPS * 32) k_prune_rank_mark(int n, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim,
                                  const double* __restrict__ q_val, const int32_t* __restrict__ df, double lim,
                                  uint8_t* __restrict__ skip, float* __restrict__ cu, int32_t* __restrict__ icnt,
                                  int32_t* __restrict__ dfmin, unsigned long long* counters) {
  __shared__ unsigned long long s_key[PRL_WARPS][PRL_MAX];
  __shared__ unsigned short s_ord[PRL_WARPS][PRL_MAX];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x * PRL_WARPS + w;
  if (v > n) return;
  if (v == n) { if (lane == 0) icnt[n] = 0; return; }
  const int a = q_ptr[v], b = q_ptr[v + 1], m = b - a;
  DBG_ASSERT(m <= PRL_MAX);
  unsigned long long* key = s_key[w]; unsigned short* ord = s_ord[w];
  for (int i = lane; i < m; i += 32)
    key[i] = ((unsigned long long)(unsigned)(0x7fffffff - __ldg(df + __ldg(q_dim + a + i))) << 32) | (unsigned)i;
  __syncwarp();
  for (int i = lane; i < m; i += 32) {
    const unsigned long long ki = key[i];
    int r = 0;
    for (int j = 0; j < m; ++j) r += key[j] < ki;          // same j in every lane: shared-memory broadcast
    ord[r] = (unsigned short)i;
  }
  __syncwarp();
  double s = 0.0; int k = 0; bool open = true;             // k: first rank that stays indexed
  for (int p0 = 0; p0 < m; p0 += 32) {
    const int p = p0 + lane;
    const int pos = p < m ? a + (int)ord[p] : -1;
    const double x = p < m ? q_val[pos] : 0.0;
    const double xx = __dmul_rn(x, x);
    int nskip = 0;
    if (open) {
      const int cntl = min(32, m - p0);
      for (int j = 0; j < cntl; ++j) {
        const double s2 = __dadd_rn(s, __shfl_sync(FULL, xx, j));
        if (!(s2 <= lim)) { open = false; break; }
        s = s2; ++nskip;
      }
      k = p0 + nskip;
    }
    if (p < m) skip[pos] = lane < nskip ? 1 : 0;
  }
  if (lane) return;
  if (k > 0) atomicAdd(&counters[C_SKIPPED], (unsigned long long)k);
  cu[v] = s > 0.0 ? __double2float_ru(sqrt(s) * (1.0 + 1e-9)) : 0.f;
  icnt[v] = m - k;
  if (dfmin) dfmin[v] = k > 0 ? (int32_t)(0x7fffffffLL - (long long)(key[ord[k - 1]] >> 32)) : 0x7fffffff;
}

// compact forward store of the INDEXED components only, (dim, fp32 weight), for the candidate-major kernel:
// appended behind the current end, which lives on the device in ifw_ptr[n_old]
__global__ void k_ifw_append(int n, int64_t n_old, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim,
                             const float* __restrict__ q_w, const uint8_t* __restrict__ skip, const int32_t* __restrict__ iptr,
                             int64_t* __restrict__ ifw_ptr, uint2* __restrict__ ifw) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n) return;
  const int64_t base = n_old ? ifw_ptr[n_old] : 0;
  if (v == 0 && n_old == 0 && lane == 0) ifw_ptr[0] = 0;
  int64_t o = base + iptr[v];
  for (int p0 = q_ptr[v]; p0 < q_ptr[v + 1]; p0 += 32) {       // warp per vector, compaction in order
    const int p = p0 + lane;
    const bool keep = p < q_ptr[v + 1] && !skip[p];
    const unsigned bal = __ballot_sync(FULL, keep);
    if (keep) ifw[o + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned)q_dim[p], __float_as_uint(q_w[p]));
    o += __popc(bal);
  }
  if (lane == 0) ifw_ptr[n_old + v + 1] = base + iptr[v + 1];
}

// ------------------------------------------------------------------ K1: index append

__global__ void k_append_rows(int n, const int32_t* __restrict__ q_ptr, int64_t nnz_base, int64_t n_local,
                              int64_t id_base, const int64_t* __restrict__ ext_keys,
                              int64_t* __restrict__ fwd_ptr, int32_t* __restrict__ gid, int64_t* __restrict__ key) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  if (v == 0 && n_local == 0) fwd_ptr[0] = 0;
  fwd_ptr[n_local + v + 1] = nnz_base + q_ptr[v + 1];
  gid[n_local + v] = (int32_t)(id_base + v);
  key[n_local + v] = ext_keys ? ext_keys[v] : (id_base + v);
}

// One thread per stored component in [nnz_lo, nnz_hi): find its row, emit (sort key, posting).
__global__ void k_emit_postings(int64_t nnz_lo, int64_t nnz_hi, int64_t row_lo, int64_t row_hi,
                                const int64_t* __restrict__ fwd_ptr, const int32_t* __restrict__ fwd_idx,
                                const double* __restrict__ fwd_val, const uint8_t* __restrict__ fwd_skip, int ntiles_aff,
                                int CR, int64_t tile0, int dimbits,
                                unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  int64_t p = nnz_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz_hi) return;
  if (fwd_skip && fwd_skip[p]) {             // not indexed: sorts behind the last affected tile and is dropped
    keys[p - nnz_lo] = (unsigned long long)ntiles_aff << dimbits; vals[p - nnz_lo] = 0ULL;
    return;
  }
  int64_t lo = row_lo, hi = row_hi;          // largest row with fwd_ptr[row] <= p
  while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (fwd_ptr[mid] <= p) lo = mid; else hi = mid; }
  int64_t row = lo;
  unsigned long long tile_rel = (unsigned long long)(row / CR - tile0);
  unsigned in_tile = (unsigned)(row % CR);
  keys[p - nnz_lo] = (tile_rel << dimbits) | (unsigned)fwd_idx[p];
  vals[p - nnz_lo] = ((unsigned long long)__float_as_uint(fmaxf((float)fwd_val[p], W_MIN)) << 32) | in_tile;   // uint2{x = id, y = w}
}

__device__ __forceinline__ int64_t lower_bound_u64(const unsigned long long* __restrict__ a, int64_t n, unsigned long long key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

__global__ void k_tile_starts(const unsigned long long* __restrict__ keys, int64_t m, int ntiles_aff, int dimbits,
                              int64_t post_base, int64_t tile0, int64_t* __restrict__ tile_start, int64_t* __restrict__ tile_base) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles_aff) return;
  int64_t s = lower_bound_u64(keys, m, (unsigned long long)t << dimbits);
  tile_start[t] = s;
  if (t < ntiles_aff) tile_base[tile0 + t] = post_base + s;
}

__global__ void k_build_dir(const unsigned long long* __restrict__ keys, int64_t m, int ntiles_aff, int D, int dimbits,
                            const int64_t* __restrict__ tile_start, int64_t tile0, int32_t* __restrict__ dir) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per = (int64_t)D + 1;
  if (i >= per * ntiles_aff) return;
  int t = (int)(i / per); int d = (int)(i - (int64_t)t * per);
  int64_t lb = lower_bound_u64(keys, m, ((unsigned long long)t << dimbits) + (unsigned long long)d);
  dir[(tile0 + t) * per + d] = (int32_t)(lb - tile_start[t]);
}

// ------------------------------------------------------------------ K2 + K3: scoring + threshold/compaction

struct ScoreArgs {
  const int32_t* q_ptr; const int32_t* q_dim; const float* q_w; const int64_t* q_key;
  const uint2* post; const int32_t* dir; const int64_t* tile_base; const int64_t* c_key;
  int32_t nq, ntiles, D, CR;
  int64_t q_local_base;      // shard-local id of query 0 when the batch was indexed in this call, else -1
  float thr_emit;            // t * (1 - guard band), rounded down
  int32_t* out_q; int32_t* out_c; float* out_est; unsigned long long out_cap;
  unsigned long long* counters;
  unsigned long long total_items;
  const int32_t* tile_cnt;   // debug builds: sparse postings per tile (bounds checks)
  int32_t seg_cap;           // dense-head kernel: capacity of the per-item segment queue
  long long post_cap;        // debug builds: capacity of post[]
  const float* row_ub;       // index reduction: per stored vector, upper bound of the L2 norm of its un-indexed part (else NULL)
  const float* q_nrm;        // index reduction: per query, upper bound of its L2 norm
};

// Persistent kernel.  One CTA per SM, WARPS warps per CTA; each warp owns one row of CR fp32
// accumulators in shared memory and loops over work items (index tile, query) drawn from a global
// cursor (tile-major, so the warps of all SMs sweep the same tile's postings out of L2 together).
// For one item the warp walks the query's terms: 32 directory look-ups at a time (one per lane),
// then for every non-empty posting segment the lanes stream postings with coalesced 8 B loads and
// do a plain LDS/FFMA/STS update -- no atomics are needed because (a) the row belongs to this warp
// alone and (b) ids inside one posting list are distinct, so lanes never collide.  The epilogue
// scans the row once: counts touched accumulators (candidates_unique), emits those >= thr_emit by
// warp-aggregated atomic compaction, and resets the row to the "untouched" marker -0.0f.
template <int WARPS, int UNROLL, bool DUPKEYS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_score(const ScoreArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* row = smem + (size_t)warp * a.CR;
  const float4 neg0 = make_float4(__uint_as_float(NEG0), __uint_as_float(NEG0), __uint_as_float(NEG0), __uint_as_float(NEG0));
  for (int i = lane * 4; i < a.CR; i += 128) *reinterpret_cast<float4*>(row + i) = neg0;
  __syncwarp();

  unsigned long long n_post = 0, n_cand = 0;
  unsigned long long item = 0;
  if (lane == 0) item = atomicAdd(&a.counters[C_WORK], 1ULL);
  item = __shfl_sync(FULL, item, 0);

  while (item < a.total_items) {
    unsigned long long next = 0;
    if (lane == 0) next = atomicAdd(&a.counters[C_WORK], 1ULL);   // latency hidden behind this item
    const int tile = (int)(item / (unsigned)a.nq);
    const int q = (int)(item - (unsigned long long)tile * (unsigned)a.nq);
    const int ts = __ldg(a.q_ptr + q), te = __ldg(a.q_ptr + q + 1);
    if (ts < te) {
      const int32_t* __restrict__ dirt = a.dir + (size_t)tile * ((size_t)a.D + 1);
      const uint2* __restrict__ pt = a.post + __ldg(a.tile_base + tile);
      for (int t0 = ts; t0 < te; t0 += 32) {
        const int t = t0 + lane;
        int s = 0, e = 0; float wq = 0.f;
        if (t < te) {
          const int d = __ldg(a.q_dim + t);
          wq = __ldg(a.q_w + t);
          s = __ldg(dirt + d); e = __ldg(dirt + d + 1);
        }
        n_post += (unsigned)(e - s);
        unsigned m = __ballot_sync(FULL, e > s);
        while (m) {
          const int j = __ffs(m) - 1; m &= m - 1;
          const int sj = __shfl_sync(FULL, s, j), ej = __shfl_sync(FULL, e, j);
          const float wj = __shfl_sync(FULL, wq, j);
          int p = sj + lane;
          for (; p + 32 * (UNROLL - 1) < ej; p += 32 * UNROLL) {
            uint2 pp[UNROLL]; float av[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) pp[u] = ld_stream(pt + p + 32 * u);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) av[u] = row[pp[u].x];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) av[u] = fmaf(wj, __uint_as_float(pp[u].y), av[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) row[pp[u].x] = av[u];
          }
          for (; p < ej; p += 32) {
            const uint2 pp = ld_stream(pt + p);
            row[pp.x] = fmaf(wj, __uint_as_float(pp.y), row[pp.x]);
          }
        }
      }
      __syncwarp();
      // ---- epilogue: count, threshold, compact, reset
      const long long c0 = (long long)tile * a.CR;
      const long long self = (a.q_local_base >= 0) ? (a.q_local_base + q - c0) : -1;   // q's own slot in this tile, if any
      long long qkey = 0;
      if (DUPKEYS) qkey = __ldg(a.q_key + q);
      for (int i = lane * 4; i < a.CR; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + i);
        *reinterpret_cast<float4*>(row + i) = neg0;
        const float vv[4] = {v.x, v.y, v.z, v.w};
        unsigned tm = 0, pm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool touched = __float_as_uint(vv[k]) != NEG0;
          tm |= (unsigned)touched << k;
          pm |= (unsigned)(touched && vv[k] >= a.thr_emit) << k;
        }
        if (self >= i && self < i + 4) { const unsigned bit = 1u << (int)(self - i); tm &= ~bit; pm &= ~bit; }
        if (DUPKEYS && tm) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((tm >> k) & 1u) { if (__ldg(a.c_key + c0 + i + k) == qkey) { tm &= ~(1u << k); pm &= ~(1u << k); } }
        }
        n_cand += __popc(tm);
        if (__ballot_sync(FULL, pm != 0)) {
          const int cnt = __popc(pm);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
          const int tot = __shfl_sync(FULL, incl, 31);
          unsigned long long base = 0;
          if (lane == 31) base = atomicAdd(&a.counters[C_PF], (unsigned long long)tot);
          base = __shfl_sync(FULL, base, 31);
          unsigned long long slot = base + (unsigned)(incl - cnt);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((pm >> k) & 1u) {
              if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)(c0 + i + k); a.out_est[slot] = vv[k]; }
              ++slot;
            }
        }
      }
      __syncwarp();
    }
    item = __shfl_sync(FULL, next, 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_post += __shfl_down_sync(FULL, n_post, o);
    n_cand += __shfl_down_sync(FULL, n_cand, o);
  }
  if (lane == 0) { atomicAdd(&a.counters[C_POSTINGS], n_post); atomicAdd(&a.counters[C_CANDS], n_cand); }
}

// ------------------------------------------------------------------ K2b: query-block scoring (fixed-point atomics)

// Per-batch transposition of the query batch into blocks of QB consecutive queries: for every block
// the distinct dimensions it uses and, per dimension, the (row, weight) list of the queries having it.
struct BlockArgs {
  const int32_t* ud_dim;    // distinct (block, dim) entries, block-major, dim ascending
  const int32_t* ud_start;  // [n_ud + 1] offsets into bt
  const int32_t* bd_ptr;    // [n_qblocks + 1] ranges of a block in ud_*
  const uint2* bt;          // (row * CR, weight * 2^F as fp32 bits), grouped by (block, dim)
  int32_t QB, n_qblocks;
  unsigned thr_int;         // emit iff acc >= thr_int  (t * 2^F with guard band, rounded down)
  float inv_scale;          // 2^-F
  float scale;              // 2^F
};

__global__ void k_bt_emit(int nq, int nnz, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim,
                          const float* __restrict__ q_w, int QB, int CR, int dimbits, float scale,
                          unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  int lo = 0, hi = nq;                      // largest q with q_ptr[q] <= t
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (q_ptr[mid] <= t) lo = mid; else hi = mid; }
  const int q = lo;
  keys[t] = ((unsigned long long)(q / QB) << dimbits) | (unsigned)q_dim[t];
  vals[t] = ((unsigned long long)__float_as_uint(q_w[t] * scale) << 32) | (unsigned)((q % QB) * CR);
}

__global__ void k_bt_heads(int nnz, const unsigned long long* __restrict__ keys, int32_t* __restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nnz) return;
  flags[t] = (t < nnz) && (t == 0 || keys[t] != keys[t - 1]);
}

__global__ void k_bt_scatter(int nnz, const unsigned long long* __restrict__ keys, const int32_t* __restrict__ flags,
                             const int32_t* __restrict__ pos, int dimbits, unsigned long long* __restrict__ ud_key,
                             int32_t* __restrict__ ud_dim, int32_t* __restrict__ ud_start) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nnz) return;
  if (t == nnz) { ud_start[pos[nnz]] = nnz; return; }     // pos[nnz] = number of distinct entries
  if (flags[t]) { const int j = pos[t]; ud_key[j] = keys[t]; ud_dim[j] = (int32_t)(keys[t] & ((1ULL << dimbits) - 1)); ud_start[j] = t; }
}

__global__ void k_bt_blocks(int n_qblocks, const int32_t* __restrict__ pos, int nnz, const unsigned long long* __restrict__ ud_key,
                            int dimbits, int32_t* __restrict__ bd_ptr) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_qblocks) return;
  bd_ptr[b] = (int32_t)lower_bound_u64(ud_key, pos[nnz], (unsigned long long)b << dimbits);
}

static constexpr int LONG_CAP = 512;     // posting segments longer than one warp chunk, per work item

// Persistent kernel, one CTA per SM.  A work item is (index tile, block of QB queries); the CTA holds
// QB x CR u32 fixed-point accumulators in shared memory.  Every posting of the tile that belongs to a
// dimension used by the block is loaded ONCE (coalesced 8 B loads) and applied to every query row that
// has the dimension: acc[row][id] += ceil(wq * 2^F * wc + 0.5) with a native shared-memory atomic
// (ATOMS.ADD, ~2x the rate of an LDS/FFMA/STS round trip, order-independent => bit-reproducible).
// Every contribution is >= 1, so "touched" (a candidate, IWA:86-92) is exactly acc != 0.
//   phase 1  warps look up 32 directory entries at a time; segments of <= 32 postings are applied
//            at once, longer ones are queued in shared memory
//   phase 2  the queued segments are cut into 32-posting chunks, dealt round-robin to the warps
//   phase 3  scan: count candidates, emit acc >= thr_int by warp-aggregated compaction, clear
template <int WARPS, bool DUPKEYS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_score_blk(const ScoreArgs a, const BlockArgs b) {
  extern __shared__ __align__(16) unsigned smem_u[];
  const int QB = b.QB, CR = a.CR;
  unsigned* acc = smem_u;
  int4* longlist = reinterpret_cast<int4*>(acc + (size_t)QB * CR);
  int* lprefix = reinterpret_cast<int*>(longlist + LONG_CAP);     // [LONG_CAP + 1] chunk prefix
  __shared__ unsigned long long s_item;
  __shared__ int s_nlong;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  constexpr int NT = WARPS * 32;
  const int nacc = QB * CR;
  for (int i = tid * 4; i < nacc; i += NT * 4) *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
  unsigned long long n_post = 0, n_cand = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_item = atomicAdd(&a.counters[C_WORK], 1ULL); s_nlong = 0; }
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= a.total_items) break;
    const int tile = (int)(item / (unsigned)b.n_qblocks);
    const int qb = (int)(item - (unsigned long long)tile * (unsigned)b.n_qblocks);
    const int32_t* __restrict__ dirt = a.dir + (size_t)tile * ((size_t)a.D + 1);
    const uint2* __restrict__ pt = a.post + __ldg(a.tile_base + tile);
    const int u0 = __ldg(b.bd_ptr + qb), u1 = __ldg(b.bd_ptr + qb + 1);

    // ---- phase 1
    for (int ub = u0 + warp * 32; ub < u1; ub += NT) {
      const int u = ub + lane;
      int s = 0, e = 0, rs = 0, re = 0;
      if (u < u1) {
        const int d = __ldg(b.ud_dim + u);
        rs = __ldg(b.ud_start + u); re = __ldg(b.ud_start + u + 1);
        s = __ldg(dirt + d); e = __ldg(dirt + d + 1);
      }
      const int len = e - s;
      n_post += (unsigned long long)(unsigned)len * (unsigned)(re - rs);
      bool queued = false;
      if (len > 32) {
        const int slot = atomicAdd(&s_nlong, 1);
        if (slot < LONG_CAP) { longlist[slot] = make_int4(s, e, rs, re); queued = true; }
      }
      unsigned m = __ballot_sync(FULL, len > 0 && !queued);
      while (m) {
        const int j = __ffs(m) - 1; m &= m - 1;
        const int sj = __shfl_sync(FULL, s, j), ej = __shfl_sync(FULL, e, j);
        const int rsj = __shfl_sync(FULL, rs, j), nr = __shfl_sync(FULL, re, j) - rsj;    // nr <= QB <= 32 rows
        uint2 rw = make_uint2(0, 0);
        if (lane < nr) rw = __ldg(b.bt + rsj + lane);        // the dimension's (row, weight) list, one per lane
        for (int p0 = sj; p0 < ej; p0 += 32) {               // <= 32 postings unless the queue overflowed
          const int p = p0 + lane;
          uint2 pp = make_uint2(0, 0);
          if (p < ej) pp = ld_stream(pt + p);
          const float wc = __uint_as_float(pp.y);
          for (int r = 0; r < nr; ++r) {
            const unsigned ro = __shfl_sync(FULL, rw.x, r);
            const float ws = __uint_as_float(__shfl_sync(FULL, rw.y, r));
            if (p < ej) atomicAdd(acc + ro + pp.x, __float2uint_ru(fmaf(ws, wc, 0.5f)));
          }
        }
      }
    }
    __syncthreads();
    // ---- chunk prefix of the queued segments (warp 0)
    const int nlong = min(s_nlong, LONG_CAP);
    if (warp == 0) {
      int run = 0;
      for (int k0 = 0; k0 < nlong; k0 += 32) {
        const int k = k0 + lane;
        int nc = 0;
        if (k < nlong) { const int4 L = longlist[k]; nc = (L.y - L.x + 31) >> 5; }
        int incl = nc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
        if (k < nlong) lprefix[k] = run + incl - nc;
        run += __shfl_sync(FULL, incl, 31);
      }
      if (lane == 0) lprefix[nlong] = run;
    }
    __syncthreads();
    // ---- phase 2: 32-posting chunks, round-robin over the warps, next chunk's loads in flight
    {
      const int nchunks = lprefix[nlong];
      int k = 0;
      int ch = warp;
      uint2 pp_n = make_uint2(0, 0), rw_n = make_uint2(0, 0); int nr_n = 0; bool ok_n = false;
      if (ch < nchunks) {
        while (lprefix[k + 1] <= ch) ++k;
        const int4 L = longlist[k];
        const int p = L.x + ((ch - lprefix[k]) << 5) + lane;
        nr_n = L.w - L.z; ok_n = p < L.y;
        if (ok_n) pp_n = ld_stream(pt + p);
        if (lane < nr_n) rw_n = __ldg(b.bt + L.z + lane);
      }
      while (ch < nchunks) {
        const uint2 pp = pp_n, rw = rw_n; const int nr = nr_n; const bool ok = ok_n;
        ch += WARPS;
        if (ch < nchunks) {
          while (lprefix[k + 1] <= ch) ++k;
          const int4 L = longlist[k];
          const int p = L.x + ((ch - lprefix[k]) << 5) + lane;
          nr_n = L.w - L.z; ok_n = p < L.y;
          if (ok_n) pp_n = ld_stream(pt + p);
          if (lane < nr_n) rw_n = __ldg(b.bt + L.z + lane);
        }
        const float wc = __uint_as_float(pp.y);
        unsigned* col = acc + pp.x;
#pragma unroll 4
        for (int r = 0; r < nr; ++r) {
          const unsigned ro = __shfl_sync(FULL, rw.x, r);
          const float ws = __uint_as_float(__shfl_sync(FULL, rw.y, r));
          if (ok) atomicAdd(col + ro, __float2uint_ru(fmaf(ws, wc, 0.5f)));
        }
      }
    }
    __syncthreads();
    // ---- phase 3: epilogue
    {
      const long long c0 = (long long)tile * CR;
      const int q0 = qb * QB;
      const long long self0 = (a.q_local_base >= 0) ? (a.q_local_base + q0 - c0) : (long long)-(1LL << 40);   // self column of row 0
      for (int i = tid * 4; i < nacc; i += NT * 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(acc + i);
        const unsigned any = v.x | v.y | v.z | v.w;
        const unsigned wany = __ballot_sync(FULL, any != 0);
        if (!wany) continue;
        if (any) *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
        const int row = i / CR, col = i - row * CR;
        const int q = q0 + row;
        const unsigned vv[4] = {v.x, v.y, v.z, v.w};
        unsigned tm = 0, pm = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { tm |= (unsigned)(vv[k] != 0) << k; pm |= (unsigned)(vv[k] != 0 && vv[k] >= b.thr_int) << k; }
        const long long selfc = self0 + row;
        if (selfc >= col && selfc < col + 4) { const unsigned bit = 1u << (int)(selfc - col); tm &= ~bit; pm &= ~bit; }
        if (DUPKEYS && tm) {
          const long long qkey = __ldg(a.q_key + q);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((tm >> k) & 1u) { if (__ldg(a.c_key + c0 + col + k) == qkey) { tm &= ~(1u << k); pm &= ~(1u << k); } }
        }
        n_cand += __popc(tm);
        if (__ballot_sync(FULL, pm != 0)) {
          const int cnt = __popc(pm);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
          const int tot = __shfl_sync(FULL, incl, 31);
          unsigned long long base = 0;
          if (lane == 31) base = atomicAdd(&a.counters[C_PF], (unsigned long long)tot);
          base = __shfl_sync(FULL, base, 31);
          unsigned long long slot = base + (unsigned)(incl - cnt);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if ((pm >> k) & 1u) {
              if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)(c0 + col + k); a.out_est[slot] = (float)vv[k] * b.inv_scale; }
              ++slot;
            }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_post += __shfl_down_sync(FULL, n_post, o);
    n_cand += __shfl_down_sync(FULL, n_cand, o);
  }
  if (lane == 0) { atomicAdd(&a.counters[C_POSTINGS], n_post); atomicAdd(&a.counters[C_CANDS], n_cand); }
}

// ------------------------------------------------------------------ K1b/K2c: dense-head tiles + FFMA scoring

// On power-law data almost all posting visits fall on a few dozen near-ubiquitous dimensions.  A tile
// therefore stores the dimensions present in >= 1/2^dense_shift of its vectors ("dense dims", at most
// KD per tile) as dense fp32 rows dense_w[tile][slot][id] (0 = absent) instead of postings; the
// scoring kernel handles them with register-tiled FFMAs (1 instruction per 32 updates instead of ~12
// on the sparse path) and everything else with the sparse path of k_score_blk.
static constexpr int KD = 64;      // dense slots per tile
static constexpr int HS = 128;     // open-addressing hash (dim -> slot) per tile

struct DenseTiles {
  const int32_t* cnt;      // [ntiles]
  const int32_t* dim;      // [ntiles][KD] ascending
  const int32_t* len;      // [ntiles][KD] true posting-list length of the dim in the tile
  const int2* hash;        // [ntiles][HS] (dim, slot); dim = -1 empty
  const float* w;          // [ntiles][KD][CR]
};

__device__ __forceinline__ unsigned dense_hash_fn(int d) { return ((unsigned)d * 2654435761u) >> 25; }   // 7 bits

// one CTA (1024 threads) per affected tile: pick the dense dims (ascending dim order, first KD that qualify).
// The directory is read in super-chunks of 64 x 1024 dims: pass 1 counts the qualifying dims per (step, warp),
// a block scan turns the counts into slot offsets, pass 2 assigns the slots -- three barriers per 65 536 dims.
__global__ void __launch_bounds__(1024) k_dense_select(int D, int CR, int64_t tile0, int64_t n_local, int dense_shift, const int32_t* __restrict__ dir,
                               int32_t* __restrict__ d_cnt, int32_t* __restrict__ d_dim, int32_t* __restrict__ d_len,
                               int2* __restrict__ d_hash, int32_t* __restrict__ tile_cnt) {
  constexpr int NT = 1024, NW = 32, IT = 64;
  const int64_t tile = tile0 + blockIdx.x;
  const int32_t* dirt = dir + (size_t)tile * ((size_t)D + 1);
  __shared__ int s_cnt[IT * NW];
  __shared__ int s_wtot[NW];
  __shared__ int s_run, s_removed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nt = min((int64_t)CR, n_local - tile * CR);
  // dense_shift < 16: threshold nt / 2^shift; otherwise (dense_shift - 16) sixteenths of the tile
  const int thr = dense_shift < 16 ? max(8, (int)((nt + (1 << dense_shift) - 1) >> dense_shift))
                                   : max(8, (int)((nt * (dense_shift - 16) + 15) / 16));
  if (tid == 0) { s_run = 0; s_removed = 0; }
  for (int h = tid; h < HS; h += NT) d_hash[tile * HS + h] = make_int2(-1, -1);
  __syncthreads();
  for (int64_t base = 0; base < D; base += (int64_t)NT * IT) {
    const int run0 = s_run;
    if (run0 >= KD) break;
    for (int it = 0; it < IT; ++it) {
      const int64_t d = base + (int64_t)it * NT + tid;
      int len = 0;
      if (d < D) len = dirt[d + 1] - dirt[d];
      const unsigned bal = __ballot_sync(FULL, len >= thr);
      if (lane == 0) s_cnt[it * NW + warp] = __popc(bal);
    }
    __syncthreads();
    // exclusive scan of the IT * NW = 2 * NT counts: two per thread
    const int a0 = s_cnt[2 * tid], a1 = s_cnt[2 * tid + 1];
    int inc = a0 + a1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_wtot[w];
    const int excl = before + inc - (a0 + a1);
    s_cnt[2 * tid] = excl; s_cnt[2 * tid + 1] = excl + a0;
    __syncthreads();
    int total = 0;
    for (int w = 0; w < NW; ++w) total += s_wtot[w];
    for (int it = 0; it < IT; ++it) {
      const int64_t d = base + (int64_t)it * NT + tid;
      int len = 0;
      if (d < D) len = dirt[d + 1] - dirt[d];
      const bool q = len >= thr;
      const unsigned bal = __ballot_sync(FULL, q);
      const int slot = run0 + s_cnt[it * NW + warp] + __popc(bal & ((1u << lane) - 1));
      if (q && slot < KD) {
        d_dim[tile * KD + slot] = (int)d; d_len[tile * KD + slot] = len;
        atomicAdd(&s_removed, len);
        unsigned h = dense_hash_fn((int)d);
        while (atomicCAS(&d_hash[tile * HS + h].x, -1, (int)d) != -1) h = (h + 1) & (HS - 1);
        d_hash[tile * HS + h].y = slot;
      }
    }
    __syncthreads();
    if (tid == 0) s_run = run0 + total;
    __syncthreads();
  }
  if (tid == 0) { d_cnt[tile] = min(s_run, KD); tile_cnt[tile] = dirt[D] - s_removed; }
}

__global__ void k_tile_bases(int64_t tile0, int ntiles_aff, const int32_t* __restrict__ tile_cnt, int64_t* __restrict__ tile_base) {
  if (blockIdx.x || threadIdx.x) return;
  int64_t b = tile0 == 0 ? 0 : tile_base[tile0 - 1] + tile_cnt[tile0 - 1];
  for (int t = 0; t < ntiles_aff; ++t) { tile_base[tile0 + t] = b; b += tile_cnt[tile0 + t]; }
}

__device__ __forceinline__ int dense_lookup(const int2* __restrict__ hash, int d) {
  unsigned h = dense_hash_fn(d);
  for (;;) { const int2 e = hash[h]; if (e.x == d) return e.y; if (e.x < 0) return -1; h = (h + 1) & (HS - 1); }
}

// thread per sorted posting: dense dims go to dense_w, the rest are compacted into the tile's postings
__global__ void k_post_scatter(int64_t m, const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ vals,
                               int dimbits, int64_t tile0, int ntiles_aff, int CR, const int64_t* __restrict__ tile_start,
                               const int64_t* __restrict__ tile_base, const int32_t* __restrict__ d_cnt, const int32_t* __restrict__ d_dim,
                               const int32_t* __restrict__ d_len, const int2* __restrict__ d_hash, float* __restrict__ d_w,
                               unsigned long long* __restrict__ post, long long post_cap) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long key = keys[i], val = vals[i];
  const int trel = (int)(key >> dimbits); const int d = (int)(key & ((1ULL << dimbits) - 1));
  if (trel >= ntiles_aff) return;           // components left out of the index (exact index reduction)
  const int64_t tile = tile0 + trel;
  const int slot = dense_lookup(d_hash + tile * HS, d);
  DBG_ASSERT(trel >= 0 && (unsigned)(val & 0xffffffffu) < (unsigned)CR && slot < KD);
  if (slot >= 0) { d_w[((size_t)tile * KD + slot) * CR + (unsigned)(val & 0xffffffffu)] = __uint_as_float((unsigned)(val >> 32)); return; }
  int removed = 0;
  const int n = d_cnt[tile];
  for (int k = 0; k < n; ++k) { if (d_dim[tile * KD + k] < d) removed += d_len[tile * KD + k]; else break; }
  DBG_ASSERT((i - tile_start[trel]) - removed >= 0 && tile_base[tile] + (i - tile_start[trel]) - removed < post_cap);
  post[tile_base[tile] + (i - tile_start[trel]) - removed] = val;
}

__global__ void k_dir_fix(int ntiles_aff, int D, int64_t tile0, const int32_t* __restrict__ d_cnt, const int32_t* __restrict__ d_dim,
                          const int32_t* __restrict__ d_len, int32_t* __restrict__ dir) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)D + 1;
  if (i >= per * ntiles_aff) return;
  const int64_t tile = tile0 + i / per; const int d = (int)(i % per);
  int removed = 0;
  const int n = d_cnt[tile];
  for (int k = 0; k < n; ++k) { if (d_dim[tile * KD + k] < d) removed += d_len[tile * KD + k]; else break; }
  dir[tile * per + d] -= removed;
}

// Scoring kernel for dense-head tiles.  Same work items, fixed-point accumulators, sparse phases and
// epilogue as k_score_blk; in addition the dimensions of the query block that are dense in the tile
// are applied by FFMA: thread = COLS adjacent candidates x all QB query rows in registers.
static constexpr int SEG_CAP = 1280;    // max queued segment pieces per work item (shared memory); ScoreArgs.seg_cap <= this
static constexpr int SPLIT = 8;         // segments up to this length are walked by the lane that looked them up
static constexpr int SEG_PIECE = 256;   // queued segments are cut into pieces of at most this many postings
static constexpr int SHORT_PIECE = 64;  // pieces up to this length use the 2-slot piece loop

// Accumulators of the dense-head kernel are u16 fixed point, two per 32-bit word: word (row*CR + c)/2,
// half c & 1 (CR is even).  An update is one native shared atomic add of (value << 16*(c&1)); halves
// cannot carry into each other because every sum is < 2^16 by the choice of the scale.
// float -> fixed point without the (slow, XU-pipe) F2I: for 0 <= p < 2^23, fma(ws, wc, 2^23) rounded UP has
// the integer ceil(p) in its low mantissa bits.  ceil(p) >= 1 for p > 0 and never below the exact product:
// every contribution is at least one quantum and over-shoots by less than one.
__device__ __forceinline__ unsigned fx_contrib(float ws, float wc) {
  return __float_as_uint(__fmaf_ru(ws, wc, 8388608.0f)) & 0x7fffffu;
}
__device__ __forceinline__ unsigned fx_ceil(float v) {          // exact ceil(v) for 0 <= v < 2^23; 0 stays 0
  return __float_as_uint(__fadd_ru(v, 8388608.0f)) & 0x7fffffu;
}
// predicated shared-memory reduction on a 32-bit shared-space address: one instruction, no branch and no
// convergence barrier around it (the compiler wraps `if (p) atomicAdd(..)` in BSSY / BRA / BSYNC)
__device__ __forceinline__ void red_shared_if(unsigned saddr, unsigned val, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q red.shared.add.u32 [%0], %1;\n\t}" :: "r"(saddr), "r"(val), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void acc_add16(unsigned* acc, unsigned rowoff_words, unsigned c, float ws, float wc) {
  atomicAdd(acc + rowoff_words + (c >> 1), fx_contrib(ws, wc) << ((c & 1u) << 4));
}

// apply postings [p, pe) (<= SPLIT of them) of one sparse segment to the rows [rs, rs+nr) of the block:
// one LANE per segment; collisions between lanes are safe (atomics)
__device__ __forceinline__ void lane_walk(unsigned* acc, const uint2* __restrict__ pt, const uint2* __restrict__ bt,
                                          int p, int pe, int rs, int nr, uint2 rw0, int nwords, int CR) {
  DBG_ASSERT(p >= 0 && pe >= p && pe - p <= SPLIT);
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc);
  while (__any_sync(FULL, p < pe)) {
    uint2 pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { pp[u] = make_uint2(0u, 0u); if (p + u < pe) pp[u] = ld_stream(pt + p + u); }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = p + u < pe;
      const float wc = __uint_as_float(pp[u].y);
      DBG_ASSERT(!ok || (pp[u].x < (unsigned)CR && rw0.x + (pp[u].x >> 1) < (unsigned)nwords));
      const unsigned w4 = (pp[u].x >> 1) << 2, sh = (pp[u].x & 1u) << 4;
      red_shared_if(acc_s + (rw0.x << 2) + w4, fx_contrib(__uint_as_float(rw0.y), wc) << sh, ok);
      if (ok)
        for (int r = 1; r < nr; ++r) {
          const uint2 rw = __ldg(bt + rs + r);
          acc_add16(acc, rw.x, pp[u].x, __uint_as_float(rw.y), wc);
        }
    }
    p += 4;
  }
}

// Piece loop of phase L: one warp per piece (<= 32 * NCH postings), handed out dynamically from `next`.
// A lane keeps the piece's NCH postings (one per 32-posting chunk) in registers -- NCH independent
// load -> FFMA -> shift -> red.shared chains -- and each register slot is refilled with the NEXT piece's
// chunk as soon as it has been consumed.  Chunks past the end of the piece are skipped warp-uniformly;
// only the last chunk is predicated per lane.  dir = +1 / -1: the queue grows up / down from `segs`.
// predicated streaming load: ONE instruction under a predicate, no branch / convergence barrier around it
__device__ __forceinline__ uint2 ld_stream_if(const uint2* p, bool on) {
  uint2 r = make_uint2(0u, 0u);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];\n\t}"
      : "+r"(r.x), "+r"(r.y) : "l"(p), "r"((unsigned)on));
  return r;
}

__device__ __forceinline__ void red_shared(unsigned saddr, unsigned val) {
  asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(saddr), "r"(val) : "memory");
}

// One piece through its NCH register slots.  WHOLE: the piece fills every slot of every lane (most queued pieces are
// cuts of exactly 32 * NCH postings) -- the reductions carry no predicate, so there is no branch / convergence barrier
// around them (ptxas turns a predicated ATOMS into BSSY / BRA / BSYNC); otherwise a chunk is skipped warp-uniformly when
// the piece ends before it and only its last chunk is predicated per lane.  TWO: the dimension is in two rows of the
// block.  The refill loads are single predicated instructions and the compiler hoists them to the top.
template <int NCH, bool WHOLE, bool TWO>
__device__ __forceinline__ void piece_slots(uint2 (&pn)[NCH], const uint2* __restrict__ pt, const int4 S, const int4 Sn, int lane,
                                            unsigned base0, unsigned base1, float ws0, float ws1, int CR) {
  const int left = S.y - S.x - lane;                          // chunk u holds one of my postings iff left > 32 u
  const int len = S.y - S.x;
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    const uint2 pp = pn[u];
    const int pnx = Sn.x + lane + 32 * u;
    pn[u] = ld_stream_if(pt + pnx, pnx < Sn.y);              // refill the slot with the next piece's chunk
    const float wc = __uint_as_float(pp.y);
    const unsigned w4 = (pp.x >> 1) << 2, sh = (pp.x & 1u) << 4;
    if (WHOLE) {
      DBG_ASSERT(pp.x < (unsigned)CR);
      red_shared(base0 + w4, fx_contrib(ws0, wc) << sh);
      if (TWO) red_shared(base1 + w4, fx_contrib(ws1, wc) << sh);
    } else if (len > 32 * u) {                                // warp-uniform
      const bool ok = left > 32 * u;
      DBG_ASSERT(!ok || pp.x < (unsigned)CR);
      red_shared_if(base0 + w4, fx_contrib(ws0, wc) << sh, ok);
      if (TWO) red_shared_if(base1 + w4, fx_contrib(ws1, wc) << sh, ok);
    }
  }
}

// Rows beyond the second -- rare for a sparse dimension in a 16-query block -- are applied in a separate loop BEFORE the
// slots are refilled.
template <int NCH>
__device__ __forceinline__ void process_pieces(unsigned* acc, const int4* segs, int dir, int nseg, int* next,
                                               const uint2* __restrict__ pt, const uint2* __restrict__ bt, int lane, int CR) {
  const unsigned acc_s = (unsigned)__cvta_generic_to_shared(acc);
  // the queue cursor is read one piece ahead of its use: lane 0 holds the raw ticket, the broadcast (which waits for
  // the atomic) happens an iteration later, so the atomic's latency is off the piece -> descriptor -> load chain
  int k = 0, kraw = 0;
  if (lane == 0) { k = atomicAdd(next, 1); kraw = atomicAdd(next, 1); }
  k = __shfl_sync(FULL, k, 0);
  int4 S = make_int4(0, 0, 0, 0); uint2 rw = make_uint2(0, 0); uint2 pn[NCH];
  if (k < nseg) {
    S = segs[dir * k];
    if (lane < S.w) rw = __ldg(bt + S.z + lane);
  }
#pragma unroll
  for (int u = 0; u < NCH; ++u) {
    const int p = S.x + lane + 32 * u;
    pn[u] = ld_stream_if(pt + p, p < S.y);
  }
  while (k < nseg) {
    k = __shfl_sync(FULL, kraw, 0);
    if (lane == 0) kraw = atomicAdd(next, 1);
    int4 Sn = make_int4(0, 0, 0, 0); uint2 rwn = make_uint2(0, 0);
    if (k < nseg) {
      Sn = segs[dir * k];
      if (lane < Sn.w) rwn = __ldg(bt + Sn.z + lane);
    }
    const unsigned ro0 = __shfl_sync(FULL, rw.x, 0), ro1 = __shfl_sync(FULL, rw.x, 1);
    const float ws0 = __uint_as_float(__shfl_sync(FULL, rw.y, 0)), ws1 = __uint_as_float(__shfl_sync(FULL, rw.y, 1));
    const unsigned base0 = acc_s + (ro0 << 2), base1 = acc_s + (ro1 << 2);
    if (S.w > 2) {                                            // warp-uniform, rare
      const int left = S.y - S.x - lane;
      for (int r = 2; r < S.w; ++r) {
        const unsigned bs = acc_s + (__shfl_sync(FULL, rw.x, r) << 2);
        const float ws = __uint_as_float(__shfl_sync(FULL, rw.y, r));
#pragma unroll
        for (int u = 0; u < NCH; ++u)
          red_shared_if(bs + ((pn[u].x >> 1) << 2), fx_contrib(ws, __uint_as_float(pn[u].y)) << ((pn[u].x & 1u) << 4), left > 32 * u);
      }
    }
    const bool whole = S.y - S.x == 32 * NCH, two = S.w > 1;   // warp-uniform
    if (whole) {
      if (two) piece_slots<NCH, true, true>(pn, pt, S, Sn, lane, base0, base1, ws0, ws1, CR);
      else piece_slots<NCH, true, false>(pn, pt, S, Sn, lane, base0, base1, ws0, ws1, CR);
    } else {
      if (two) piece_slots<NCH, false, true>(pn, pt, S, Sn, lane, base0, base1, ws0, ws1, CR);
      else piece_slots<NCH, false, false>(pn, pt, S, Sn, lane, base0, base1, ws0, ws1, CR);
    }
    S = Sn; rw = rwn;
  }
}

// Scoring kernel for dense-head tiles.  Persistent, one CTA per SM; a work item is (index tile, block
// of QB queries); QB x CR u16 accumulators live in shared memory.  Per item:
//   phase 1  32 directory look-ups per warp step.  Dims that are dense in the tile go to the dense list;
//            sparse segments of <= SPLIT postings are walked at once, one lane per segment; longer ones
//            are queued
//   phase D  dense dims by FFMA: thread = COLS adjacent candidates x all QB query rows in registers,
//            result added to the accumulators (exclusive phase, plain read-modify-write)
//   phase L  queued segments, one warp per segment: coalesced 8 B posting loads (next chunk in flight),
//            the dimension's (row, weight) list held one entry per lane and broadcast by shuffle
//   phase 3  epilogue: count candidates (non-zero halves), emit those >= thr by warp-aggregated
//            compaction, clear
// bt entries for this kernel hold (row * CR / 2, weight * 2^F).
template <int QB, int WARPS, int COLS, bool DUPKEYS, bool PRUNED = false>
__global__ void __launch_bounds__(WARPS * 32, (WARPS <= 8 ? 2 : 1)) k_score_dense(const ScoreArgs a, const BlockArgs b, const DenseTiles dt) {
  extern __shared__ __align__(16) unsigned smem_u[];
  const int CR = a.CR;
  const int RW = CR >> 1;                                               // words per accumulator row
  unsigned* acc = smem_u;                                               // [QB][RW]
  float* Wq = reinterpret_cast<float*>(acc + (size_t)QB * RW);          // [KD][QB] scaled query weights, by dense entry
  int4* dl = reinterpret_cast<int4*>(Wq + KD * QB);                     // [KD] (slot, rs, nr, -)
  int2* hsh = reinterpret_cast<int2*>(dl + KD);                         // [HS]
  int4* segs = reinterpret_cast<int4*>(hsh + HS);                       // [SEG_CAP] (s, e, rs, nr)
  __shared__ unsigned long long s_item;
  __shared__ int s_nseg, s_segvalid, s_nshort, s_shortvalid, s_ndense, s_next, s_next2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  constexpr int NT = WARPS * 32;
  const int nwords = QB * RW;
  for (int i = tid * 4; i < nwords; i += NT * 4) *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
  // tallies: 32-bit per thread and item, summed per warp into shared 64-bit totals at the top of every item
  // (keeps four 64-bit counters out of the register file for the whole kernel)
  __shared__ unsigned long long s_tal[4];
  if (tid < 4) s_tal[tid] = 0ULL;
  unsigned n_post = 0, n_cand = 0, n_dpost = 0;
  __syncthreads();
#ifdef APSS_PHASE_TIMERS
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tc = clock64();
#define PHASE_MARK(k) do { if (tid == 0) { const long long now_ = clock64(); ph[k] += now_ - tc; tc = now_; } } while (0)
#else
#define PHASE_MARK(k) ((void)0)
#endif
  const unsigned thr_lo = b.thr_int;
  const unsigned thr_hi = thr_lo << 16;      // high half >= thr  <=>  word >= thr << 16

  for (;;) {
    {
      unsigned long long tp = n_post, tc_ = n_cand, td = n_dpost;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_down_sync(FULL, tp, o);
        tc_ += __shfl_down_sync(FULL, tc_, o);
        td += __shfl_down_sync(FULL, td, o);
      }
      if (lane == 0) { if (tp) atomicAdd(&s_tal[0], tp); if (tc_) atomicAdd(&s_tal[1], tc_); if (td) atomicAdd(&s_tal[2], td); }
      n_post = 0; n_cand = 0; n_dpost = 0;
    }
    __syncthreads();
    PHASE_MARK(5);
    if (tid == 0) { s_item = atomicAdd(&a.counters[C_WORK], 1ULL); s_nseg = 0; s_segvalid = a.seg_cap - (a.seg_cap * 3 >> 3); s_nshort = 0; s_shortvalid = a.seg_cap * 3 >> 3; s_ndense = 0; s_next = 0; s_next2 = 0; }
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= a.total_items) break;
    const int tile = (int)(item / (unsigned)b.n_qblocks);
    const int qb = (int)(item - (unsigned long long)tile * (unsigned)b.n_qblocks);
    const int32_t* __restrict__ dirt = a.dir + (size_t)tile * ((size_t)a.D + 1);
    const uint2* __restrict__ pt = a.post + __ldg(a.tile_base + tile);
    const int u0 = __ldg(b.bd_ptr + qb), u1 = __ldg(b.bd_ptr + qb + 1);
    const int ndt = __ldg(dt.cnt + tile);
    const int tcnt = a.tile_cnt ? __ldg(a.tile_cnt + tile) : 0x7fffffff;
    DBG_ASSERT(tile >= 0 && tile < a.ntiles && qb < b.n_qblocks && ndt >= 0 && ndt <= KD && u0 <= u1);
    DBG_ASSERT(!a.tile_cnt || __ldg(a.tile_base + tile) + tcnt <= a.post_cap);
    for (int h = tid; h < HS; h += NT) hsh[h] = __ldg(dt.hash + (size_t)tile * HS + h);
    for (int i = tid; i < KD * QB; i += NT) Wq[i] = 0.f;
    __syncthreads();
    PHASE_MARK(0);

    // ---- phase 1
    for (int ub = u0 + warp * 32; ub < u1; ub += NT) {
      const int u = ub + lane;
      int s = 0, e = 0, rs = 0, nr = 0;
      uint2 rw0 = make_uint2(0, 0);
      if (u < u1) {
        const int d = __ldg(b.ud_dim + u);
        rs = __ldg(b.ud_start + u); nr = __ldg(b.ud_start + u + 1) - rs;
        const int slot = ndt ? dense_lookup(hsh, d) : -1;
        if (slot >= 0) {
          const int k = atomicAdd(&s_ndense, 1);
          DBG_ASSERT(k < KD && slot < ndt && nr >= 1 && nr <= QB);
          dl[k] = make_int4(slot, rs, nr, 0);
          const unsigned dp = (unsigned)__ldg(dt.len + (size_t)tile * KD + slot) * (unsigned)nr;
          n_post += dp; n_dpost += dp;
        } else {
          s = __ldg(dirt + d); e = __ldg(dirt + d + 1);
          DBG_ASSERT(d >= 0 && d < a.D && s >= 0 && e >= s && e <= tcnt && nr >= 1 && nr <= QB);
        }
      }
      const int len = e - s;
      n_post += (unsigned)len * (unsigned)nr;
      if (len > 0 && len <= SPLIT) rw0 = __ldg(b.bt + rs);
      bool coop = false;                      // segment queue full: walk it here, warp-cooperatively
      if (len > SPLIT) {
        // queue in pieces of <= SEG_PIECE postings (balance across warps); a (last) piece of <= SHORT_PIECE
        // postings goes to the short queue (2 chunk slots per piece instead of 8)
        const int rem = len % SEG_PIECE;
        const int ns = (rem > 0 && rem <= SHORT_PIECE) ? 1 : 0;
        const int nl = (len + SEG_PIECE - 1) / SEG_PIECE - ns;
        // two independent queues: long pieces in segs[0, cap_l), short ones in segs[cap_l, seg_cap).  Each
        // cursor only grows, so once a reservation on a cursor fails every later one on it fails too and
        // the valid prefix of that queue is [0, first failed base).
        const int cap_l = a.seg_cap - (a.seg_cap * 3 >> 3), cap_s = a.seg_cap - cap_l;
        const int kl = nl ? atomicAdd(&s_nseg, nl) : 0;
        const int ks = ns ? atomicAdd(&s_nshort, ns) : 0;
        const bool okl = kl + nl <= cap_l, oks = ks + ns <= cap_s;
        if (!okl && nl) atomicMin(&s_segvalid, kl);
        if (!oks && ns) atomicMin(&s_shortvalid, ks);
        if (okl && oks) {
          for (int j = 0; j < nl; ++j) segs[kl + j] = make_int4(s + j * SEG_PIECE, min(e, s + (j + 1) * SEG_PIECE), rs, nr);
          if (ns) segs[cap_l + ks] = make_int4(e - rem, e, rs, nr);
        } else {                                  // all or nothing: neutralise the half that was reserved
          if (okl) for (int j = 0; j < nl; ++j) segs[kl + j] = make_int4(0, 0, rs, 0);
          if (oks && ns) segs[cap_l + ks] = make_int4(0, 0, rs, 0);
          coop = true;
        }
      }
      lane_walk(acc, pt, b.bt, s, len <= SPLIT ? e : s, rs, nr, rw0, nwords, CR);
      unsigned m = __ballot_sync(FULL, coop);
      while (m) {
        const int j = __ffs(m) - 1; m &= m - 1;
        const int sj = __shfl_sync(FULL, s, j), ej = __shfl_sync(FULL, e, j);
        const int rsj = __shfl_sync(FULL, rs, j), nrj = __shfl_sync(FULL, nr, j);
        uint2 rw = make_uint2(0, 0);
        if (lane < nrj) rw = __ldg(b.bt + rsj + lane);
        for (int p0 = sj; p0 < ej; p0 += 32) {
          const int p = p0 + lane;
          uint2 pp = make_uint2(0, 0);
          if (p < ej) pp = ld_stream(pt + p);
          const float wc = __uint_as_float(pp.y);
          for (int r = 0; r < nrj; ++r) {
            const unsigned ro = __shfl_sync(FULL, rw.x, r);
            const float ws = __uint_as_float(__shfl_sync(FULL, rw.y, r));
            if (p < ej) acc_add16(acc, ro, pp.x, ws, wc);
          }
        }
      }
    }
    __syncthreads();
    PHASE_MARK(1);
    const int nd = s_ndense;
    const int nlong = min(s_nseg, s_segvalid), nshort = min(s_nshort, s_shortvalid);
    // pad the dense list to a multiple of 4 with entries whose query weights stay zero (Wq is cleared per item)
    if (tid < 4 && nd + tid < ((nd + 3) & ~3)) dl[nd + tid] = make_int4(0, 0, 0, 0);
    // ---- Wq[entry][row] from the block's row lists
    for (int k = warp; k < nd; k += WARPS) {
      const int4 e = dl[k];
      if (lane < e.z) {
        const uint2 rw = __ldg(b.bt + e.y + lane);
        DBG_ASSERT(rw.x / (unsigned)RW < (unsigned)QB && k < KD);
        Wq[k * QB + rw.x / (unsigned)RW] = __uint_as_float(rw.y);
      }
    }
    __syncthreads();
    // ---- phase D: dense dims by FFMA, thread = COLS adjacent candidates x QB rows
    if (nd) {
      if (tid == 0) s_tal[3] += (unsigned long long)((nd + 3) & ~3) * (unsigned long long)QB * (unsigned long long)CR;
      const float* __restrict__ wbase = dt.w + (size_t)tile * KD * CR;
      for (int cb = tid * COLS; cb < CR; cb += NT * COLS) {
        float av[COLS][QB];
#pragma unroll
        for (int c = 0; c < COLS; ++c)
#pragma unroll
          for (int r = 0; r < QB; ++r) av[c][r] = 0.f;
        float wa[4][COLS], wb[4][COLS];                   // ping-pong: no register copies between groups
        auto load4 = [&](float (&dst)[4][COLS], int e0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* src = wbase + (size_t)dl[e0 + j].x * CR + cb;
            if (COLS == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(src)); dst[j][0] = t.x; dst[j][COLS - 1] = t.y; }
            else if (COLS == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(src)); dst[j][0] = t.x; dst[j][1 % COLS] = t.y; dst[j][2 % COLS] = t.z; dst[j][3 % COLS] = t.w; }
            else { for (int c = 0; c < COLS; ++c) dst[j][c] = __ldg(src + c); }
          }
        };
        auto fma4 = [&](const float (&w)[4][COLS], int e0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4* q4 = reinterpret_cast<const float4*>(Wq + (e0 + j) * QB);
#pragma unroll
            for (int r4 = 0; r4 < QB / 4; ++r4) {
              const float4 q = q4[r4];
#pragma unroll
              for (int c = 0; c < COLS; ++c) {
                av[c][r4 * 4 + 0] = fmaf(q.x, w[j][c], av[c][r4 * 4 + 0]);
                av[c][r4 * 4 + 1] = fmaf(q.y, w[j][c], av[c][r4 * 4 + 1]);
                av[c][r4 * 4 + 2] = fmaf(q.z, w[j][c], av[c][r4 * 4 + 2]);
                av[c][r4 * 4 + 3] = fmaf(q.w, w[j][c], av[c][r4 * 4 + 3]);
              }
            }
          }
        };
        const int nd4 = (nd + 3) & ~3;
        load4(wa, 0);
        for (int e0 = 0; e0 < nd4; e0 += 8) {
          if (e0 + 4 < nd4) load4(wb, e0 + 4);
          fma4(wa, e0);
          if (e0 + 4 < nd4) {
            if (e0 + 8 < nd4) load4(wa, e0 + 8);
            fma4(wb, e0 + 4);
          }
        }
        // add to the accumulators: exclusive phase (between barriers), plain read-modify-write of
        // whole words (COLS is even, so a thread owns both halves of each word it touches)
#pragma unroll
        for (int r = 0; r < QB; ++r)
#pragma unroll
          for (int c = 0; c < COLS; c += 2) {
            const unsigned add = fx_ceil(av[c][r]) | (fx_ceil(av[c + 1][r]) << 16);
            if (add) acc[r * RW + ((cb + c) >> 1)] += add;
          }
      }
    }
    __syncthreads();
    PHASE_MARK(2);
    // ---- phase L: queued pieces.  Long pieces (65..256 postings) with 8 chunk slots, short ones (<= 64)
    // with 2: most pieces are short and the per-piece cost grows with the number of slots.
    process_pieces<SEG_PIECE / 32>(acc, segs, +1, nlong, &s_next, pt, b.bt, lane, CR);
    process_pieces<SHORT_PIECE / 32>(acc, segs + (a.seg_cap - (a.seg_cap * 3 >> 3)), +1, nshort, &s_next2, pt, b.bt, lane, CR);
    __syncthreads();
    PHASE_MARK(3);
    // ---- phase 3: epilogue
    {
      const long long c0 = (long long)tile * CR;
      const int q0 = qb * QB;
      // a query indexed in this call sees its own postings: clear its own accumulator first (IWA:91)
      if (a.q_local_base >= 0 && tid < QB) {
        const long long selfc = a.q_local_base + q0 + tid - c0;
        if (selfc >= 0 && selfc < CR) acc[tid * RW + (int)(selfc >> 1)] &= (selfc & 1) ? 0x0000ffffu : 0xffff0000u;
      }
      if (PRUNED && tid == 0) s_nseg = 0;                       // reused as the cursor of the touched-word queue
      __syncthreads();
      PHASE_MARK(7);
      if (PRUNED) {
        // Reduced index: every touched candidate has its own threshold (thr - bound of its un-indexed part), so
        // all touched words need global loads.  Pass 1 scans and clears the accumulators and queues the non-zero
        // words (the segment queue is free now); pass 2 tests them one per thread, latencies in parallel.
        uint2* tq = reinterpret_cast<uint2*>(segs);
        const int tcap = a.seg_cap * 2;
        auto test_word = [&](unsigned widx, unsigned word) {
          const int row = (int)(widx / (unsigned)RW), wcol = (int)(widx - (unsigned)row * (unsigned)RW);
          const int q = q0 + row;
          const long long c = c0 + (long long)wcol * 2;
          const float2 ub2 = __ldg(reinterpret_cast<const float2*>(a.row_ub + c));
          const float qn = __fmul_ru(__ldg(a.q_nrm + q), b.scale);
          long long qkey = 0;
          if (DUPKEYS) qkey = __ldg(a.q_key + q);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const unsigned val = (word >> (k << 4)) & 0xffffu;
            if (!val) continue;
            if (DUPKEYS) { if (__ldg(a.c_key + c + k) == qkey) continue; ++n_cand; }
            // dot(q, c) <= indexed part + |q| * |un-indexed part of c| (Cauchy-Schwarz), everything rounded up
            const unsigned ub = fx_ceil(fminf(__fmul_ru(k ? ub2.y : ub2.x, qn), 8388607.f));
            if (val + ub >= b.thr_int) {
              const unsigned long long slot = atomicAdd(&a.counters[C_PF], 1ULL);
              if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)(c + k); a.out_est[slot] = (float)val * b.inv_scale; }
            }
          }
        };
        for (int i = tid * 4; i < nwords; i += NT * 4) {
          const uint4 v = *reinterpret_cast<const uint4*>(acc + i);
          if ((v.x | v.y | v.z | v.w) == 0) continue;
          *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
          const unsigned vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (!vv[k]) continue;
            if (!DUPKEYS) n_cand += (unsigned)((vv[k] & 0xffffu) != 0) + (unsigned)(vv[k] > 0xffffu);
            const int slot = atomicAdd(&s_nseg, 1);
            if (slot < tcap) tq[slot] = make_uint2((unsigned)(i + k), vv[k]);
            else test_word((unsigned)(i + k), vv[k]);           // queue full: test in place
          }
        }
        __syncthreads();
        const int nt = min(s_nseg, tcap);
        for (int e = tid; e < nt; e += NT) { const uint2 w = tq[e]; test_word(w.x, w.y); }
      } else
      for (int i = tid * 4; i < nwords; i += NT * 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(acc + i);
        if ((v.x | v.y | v.z | v.w) == 0) continue;
        *reinterpret_cast<uint4*>(acc + i) = make_uint4(0, 0, 0, 0);
        const unsigned vv[4] = {v.x, v.y, v.z, v.w};
        unsigned hot = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                            // plain compares (the SIMD-in-word intrinsics are emulated)
          const unsigned lo = vv[k] & 0xffffu;
          if (!DUPKEYS) n_cand += (unsigned)(lo != 0) + (unsigned)(vv[k] > 0xffffu);
          hot |= (unsigned)(lo != 0 && lo >= thr_lo) | (unsigned)(vv[k] > 0xffffu && vv[k] >= thr_hi);
        }
        if (DUPKEYS || hot) {                                   // rare: something to emit (or key checks)
#ifdef APSS_PHASE_TIMERS
          if (tid == 0) ph[6] += 1;
#endif
          const int row = i / RW, wcol = i - row * RW;
          const int q = q0 + row;
          long long qkey = 0;
          if (DUPKEYS) qkey = __ldg(a.q_key + q);
          unsigned pm = 0;                                      // halves to emit
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const unsigned val = (vv[k >> 1] >> ((k & 1) << 4)) & 0xffffu;
            if (!val) continue;
            if (DUPKEYS) { if (__ldg(a.c_key + c0 + (long long)(wcol + (k >> 1)) * 2 + (k & 1)) == qkey) continue; ++n_cand; }
            if (val >= b.thr_int) pm |= 1u << k;
          }
          if (pm) {                                             // one slot claim per thread (<= 8 pairs)
            unsigned long long slot = atomicAdd(&a.counters[C_PF], (unsigned long long)__popc(pm));
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if ((pm >> k) & 1u) {
                const unsigned val = (vv[k >> 1] >> ((k & 1) << 4)) & 0xffffu;
                if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)(c0 + (long long)(wcol + (k >> 1)) * 2 + (k & 1)); a.out_est[slot] = (float)val * b.inv_scale; }
                ++slot;
              }
          }
        }
      }
      PHASE_MARK(4);
    }
  }
  // the loop flushed every thread's tallies before the break; the barrier at its top ordered them before this read
  if (tid == 0) {
    if (s_tal[0]) atomicAdd(&a.counters[C_POSTINGS], s_tal[0]);
    if (s_tal[1]) atomicAdd(&a.counters[C_CANDS], s_tal[1]);
    if (s_tal[2]) atomicAdd(&a.counters[C_DENSE_POST], s_tal[2]);
    if (s_tal[3]) atomicAdd(&a.counters[C_DENSE_FMA], s_tal[3]);
  }
#ifdef APSS_PHASE_TIMERS
  if (tid == 0) { for (int k = 0; k < 8; ++k) atomicAdd(&a.counters[C_PHASE + k], (unsigned long long)ph[k]); }
#endif
#undef PHASE_MARK
}


// ------------------------------------------------------------------ K2c: candidate-major scoring on the reduced index

// With exact index reduction on, a batch of 16 K queries touches ~10^8 postings instead of 2 * 10^11, and the
// tile sweep of the kernels above (every query dimension looked up in every tile's directory, every
// accumulator tile scanned) costs far more than the updates themselves.  This kernel turns the join around:
// the BATCH is inverted (dim -> (query, weight) lists, rebuilt per call: k_qi_emit + sort + k_qdir) and the
// stored vectors are streamed once; a warp takes one stored vector c, walks the query lists of c's INDEXED
// components and accumulates dot(q, c_indexed) for the queries it meets in a small per-warp hash table in
// shared memory.  The same counters result: postings visited = sum over (c, indexed d) of |queries with d|,
// candidates = touched (q, c) pairs.  A touched pair survives to the fp64 verify kernel iff
//   estimate * (1 + guard band) + |q| * |c_unindexed|  >=  t.
// Stored vectors whose lists are too long for the table are deferred to k_score_cand_heavy.
struct CandArgs {
  const int64_t* ifw_ptr; const uint2* ifw;     // indexed components of the stored vectors: (dim, fp32 weight)
  const float* row_ub; const int64_t* c_key;
  const int32_t* qdir;       // [D + 1] offsets into qi
  const uint2* qi;           // (query, weight as fp32 bits), by dim, query ascending
  const float* q_nrm; const int64_t* q_key;
  int64_t n_rows;            // stored vectors visible to this batch
  int64_t q_local_base;      // shard-local id of query 0 when the batch was indexed in this call, else -1
  int32_t nq;
  int32_t q_lo, q_hi;        // the query slice this launch scores (its lists are what qdir points at)
  float thr, band1;          // t and 1 + guard band of the fp32 estimate
  float scale, inv_scale;    // 2^F, 2^-F: fixed-point accumulators of k_score_cand
  int32_t* out_q; int32_t* out_c; float* out_est; unsigned long long out_cap;
  unsigned long long* counters;
  int32_t* heavy; int64_t heavy_cap;
};

__global__ void k_qi_emit(int n, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim, const float* __restrict__ q_w,
                          int qsub, int dimbits, unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= n) return;
  const unsigned long long slice = (unsigned long long)(v / qsub) << dimbits;
  for (int p = q_ptr[v] + (threadIdx.x & 31); p < q_ptr[v + 1]; p += 32) {      // warp per vector
    keys[p] = slice | (unsigned long long)(unsigned)q_dim[p];
    vals[p] = ((unsigned long long)__float_as_uint(q_w[p]) << 32) | (unsigned)v;      // uint2{x = query, y = weight}
  }
}

// one directory of D + 1 offsets per query slice
__global__ void k_qdir(const unsigned long long* __restrict__ keys, int nnz, int D, int dimbits, int slices, int32_t* __restrict__ qdir) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)D + 1;
  if (i >= per * slices) return;
  const long long sl = i / per, d = i - sl * per;
  qdir[i] = (int32_t)lower_bound_u64(keys, nnz, ((unsigned long long)sl << dimbits) + (unsigned long long)d);
}

static constexpr int CAND_CHUNK = 16;      // stored vectors per work-cursor fetch
static constexpr int CAND_SHORT = 8;       // lists up to this length are walked by the lane that looked them up

// fixed-point accumulation (native shared-memory atomic add; every contribution rounded up, so it is >= 1)
__device__ __forceinline__ void cand_insert(unsigned* keys, unsigned* vals, unsigned mask, unsigned q, float p, float scale) {
  unsigned h = ((q * 2654435761u) >> 12) & mask;
  for (;;) {
    const unsigned old = atomicCAS(keys + h, 0u, q + 1u);
    if (old == 0u || old == q + 1u) break;
    h = (h + 1u) & mask;
  }
  atomicAdd(vals + h, __float2uint_ru(p * scale));
}

__device__ __forceinline__ void cand_test_emit(const CandArgs& a, int q, long long c, float est, float cu, float qn, bool same_key,
                                               unsigned long long& n_cand) {
  if (a.q_local_base >= 0 && a.q_local_base + q == c) return;           // a query never meets itself (IWA:91)
  if (same_key) return;                                                  // same external id (IWA:91)
  ++n_cand;
  const float ub = __fmul_ru(cu, qn);
  if (__fmaf_ru(est, a.band1, ub) >= a.thr) {
    const unsigned long long slot = atomicAdd(&a.counters[C_PF], 1ULL);
    if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)c; a.out_est[slot] = est; }
  }
}

static constexpr int CAND_FEAT = 64;       // non-empty query lists per stored vector handled by the flat walk

template <int WARPS, int TBL>
__global__ void __launch_bounds__(WARPS * 32, 1) k_score_cand(const CandArgs a) {
  constexpr int LIMIT = TBL / 2;              // longest total list length handled in the table (load factor <= 1/2)
  extern __shared__ __align__(16) unsigned smem_u[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned* keys = smem_u + (size_t)warp * (2 * TBL);
  unsigned* vals = keys + TBL;
  // per-warp list of the vector's non-empty query lists: start offset in the flattened walk, list start, weight
  int* pre = reinterpret_cast<int*>(smem_u + (size_t)WARPS * (2 * TBL)) + warp * (3 * CAND_FEAT + 4);
  int* fs = pre + CAND_FEAT + 4;
  float* fw = reinterpret_cast<float*>(fs + CAND_FEAT);
  for (int i = lane; i < TBL; i += 32) { keys[i] = 0u; vals[i] = 0u; }
  __syncwarp();
  unsigned long long n_post = 0, n_cand = 0;
  constexpr int RC = 4;                       // components per lane kept in registers between the two passes
  // fallback walk of one component's query list: short lists by the owning lane, long ones by the whole warp
  auto walk = [&](int s, int len, float w, unsigned mask) {
    if (len <= CAND_SHORT)
      for (int p = s; p < s + len; ++p) { const uint2 x = __ldg(a.qi + p); cand_insert(keys, vals, mask, x.x, w * __uint_as_float(x.y), a.scale); }
    unsigned m = __ballot_sync(FULL, len > CAND_SHORT);
    while (m) {
      const int src = __ffs(m) - 1; m &= m - 1;
      const int sj = __shfl_sync(FULL, s, src), lj = __shfl_sync(FULL, len, src);
      const float wj = __shfl_sync(FULL, w, src);
      for (int p = sj + lane; p < sj + lj; p += 32) { const uint2 x = __ldg(a.qi + p); cand_insert(keys, vals, mask, x.x, wj * __uint_as_float(x.y), a.scale); }
    }
  };
  auto lookup = [&](long long j, long long fe, int& s, int& len, float& w) {
    s = 0; len = 0; w = 0.f;
    if (j < fe) {
      const uint2 f = __ldg(a.ifw + j);
      s = __ldg(a.qdir + f.x); len = __ldg(a.qdir + f.x + 1) - s;
      w = __uint_as_float(f.y);
    }
  };
  for (;;) {
    long long base = 0;
    if (lane == 0) base = (long long)atomicAdd(&a.counters[C_WORK], (unsigned long long)CAND_CHUNK);
    base = __shfl_sync(FULL, base, 0);
    if (base >= a.n_rows) break;
    const int nc = (int)min((long long)CAND_CHUNK, (long long)a.n_rows - base);
    long long myptr = 0;
    if (lane <= nc) myptr = __ldg(a.ifw_ptr + base + lane);        // the chunk's row pointers, one load
    for (int ci = 0; ci < nc; ++ci) {
      const long long c = base + ci;
      const long long fa = __shfl_sync(FULL, myptr, ci), fe = __shfl_sync(FULL, myptr, ci + 1);
      int s_[RC], len_[RC]; float w_[RC];
      unsigned total = 0;
#pragma unroll
      for (int it = 0; it < RC; ++it) {
        s_[it] = 0; len_[it] = 0; w_[it] = 0.f;
        if (fa + it * 32 < fe) { lookup(fa + it * 32 + lane, fe, s_[it], len_[it], w_[it]); total += (unsigned)len_[it]; }
      }
      const bool tail = fe - fa > RC * 32;
      if (tail) for (long long j = fa + RC * 32 + lane; j < fe; j += 32) { int s, len; float w; lookup(j, fe, s, len, w); total += (unsigned)len; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
      if (!total) continue;
      if (total > (unsigned)LIMIT) {
        if (lane == 0) {
          const unsigned long long k = atomicAdd(&a.counters[C_HEAVY], 1ULL); atomicAdd(&a.counters[C_HEAVY_TOT], 1ULL);
          if ((long long)k < a.heavy_cap) a.heavy[k] = (int32_t)c;
        }
        continue;
      }
      if (lane == 0) n_post += total;
      unsigned size = 64; while (size < 2u * total) size <<= 1;
      const unsigned mask = size - 1u;
      // compact the non-empty lists: slot and exclusive prefix of the lengths
      int nf = 0, run = 0;
#pragma unroll
      for (int it = 0; it < RC; ++it) {
        if (fa + it * 32 >= fe) break;                           // warp-uniform
        const unsigned bal = __ballot_sync(FULL, len_[it] > 0);
        int inc = len_[it];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        const int slot = nf + __popc(bal & ((1u << lane) - 1u));
        if (len_[it] > 0 && slot < CAND_FEAT) { pre[slot] = run + inc - len_[it]; fs[slot] = s_[it]; fw[slot] = w_[it]; }
        nf += __popc(bal); run += __shfl_sync(FULL, inc, 31);
      }
      if (!tail && nf <= CAND_FEAT) {
        for (int i = nf + lane; i < CAND_FEAT + 1; i += 32) pre[i] = 0x7fffffff;      // padding for the fixed-depth search
        __syncwarp();
        // flat walk: item t of the concatenated lists -> (list, position); iterations are independent
        for (int t0 = 0; t0 < (int)total; t0 += 64) {
          int f_[2], t_[2]; uint2 x_[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            t_[u] = t0 + u * 32 + lane; f_[u] = 0;
            if (t_[u] < (int)total) {
              int lo = 0;                                        // largest f with pre[f] <= t (pre[0] = 0, padded with INT_MAX)
#pragma unroll
              for (int st = CAND_FEAT / 2; st > 0; st >>= 1) lo += (pre[lo + st] <= t_[u]) ? st : 0;
              f_[u] = lo;
              x_[u] = __ldg(a.qi + fs[lo] + (t_[u] - pre[lo]));
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (t_[u] < (int)total) cand_insert(keys, vals, mask, x_[u].x, fw[f_[u]] * __uint_as_float(x_[u].y), a.scale);
        }
      } else {
#pragma unroll
        for (int it = 0; it < RC; ++it) if (fa + it * 32 < fe) walk(s_[it], len_[it], w_[it], mask);
        for (long long j0 = fa + RC * 32; j0 < fe; j0 += 32) { int s, len; float w; lookup(j0 + lane, fe, s, len, w); walk(s, len, w, mask); }
      }
      __syncwarp();
      const float cu = __ldg(a.row_ub + c);
      const long long ckey = a.q_key ? __ldg(a.c_key + c) : 0;
      for (unsigned i0 = 0; i0 < size; i0 += 128) {
        unsigned k_[4], v_[4]; float qn_[4]; bool same_[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned i = i0 + u * 32 + lane;
          k_[u] = i < size ? keys[i] : 0u; v_[u] = 0u; qn_[u] = 0.f; same_[u] = false;
          if (k_[u]) {
            v_[u] = vals[i]; keys[i] = 0u; vals[i] = 0u;
            qn_[u] = __ldg(a.q_nrm + (k_[u] - 1u));
            if (a.q_key) same_[u] = __ldg(a.q_key + (k_[u] - 1u)) == ckey;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k_[u]) cand_test_emit(a, (int)(k_[u] - 1u), c, __uint2float_ru(v_[u]) * a.inv_scale, cu, qn_[u], same_[u], n_cand);
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);
  if (lane == 0) { atomicAdd(&a.counters[C_POSTINGS], n_post); atomicAdd(&a.counters[C_CANDS], n_cand); }
}

// Heavy pass: one CTA per deferred stored vector, dense fp32 accumulators over a chunk of `qc` queries in
// shared memory (positive products: touched <=> non-zero), repeated for every chunk of the batch.
__global__ void __launch_bounds__(512, 1) k_score_cand_heavy(const CandArgs a, int qc) {
  extern __shared__ __align__(16) unsigned smem_u[];
  float* acc = reinterpret_cast<float*>(smem_u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const long long nh = min((long long)a.counters[C_HEAVY], (long long)a.heavy_cap);
  unsigned long long n_post = 0, n_cand = 0;
  for (long long hidx = blockIdx.x; hidx < nh; hidx += gridDim.x) {
    const long long c = a.heavy[hidx];
    const long long fa = __ldg(a.ifw_ptr + c), fe = __ldg(a.ifw_ptr + c + 1);
    const float cu = __ldg(a.row_ub + c);
    const long long ckey = a.q_key ? __ldg(a.c_key + c) : 0;
    for (int q_lo = a.q_lo; q_lo < a.q_hi; q_lo += qc) {
      const int q_hi = min(a.q_hi, q_lo + qc);
      for (int i = tid; i < q_hi - q_lo; i += blockDim.x) acc[i] = 0.f;
      __syncthreads();
      for (long long j0 = fa + (long long)warp * 32; j0 < fe; j0 += (long long)nw * 32) {      // 32 components per warp step
        const long long j = j0 + lane;
        int s = 0, e = 0; float w = 0.f;
        if (j < fe) {
          const uint2 f = __ldg(a.ifw + j);
          s = __ldg(a.qdir + f.x); e = __ldg(a.qdir + f.x + 1);
          w = __uint_as_float(f.y);
        }
        if (q_lo == a.q_lo) n_post += (unsigned)(e - s);
        const bool longl = e - s > 32;
        if (!longl)                                            // short list: the lane that looked it up
          for (int p = s; p < e; ++p) {
            const uint2 x = __ldg(a.qi + p);
            if ((int)x.x >= q_lo && (int)x.x < q_hi) atomicAdd(acc + ((int)x.x - q_lo), w * __uint_as_float(x.y));
          }
        unsigned m = __ballot_sync(FULL, longl);
        while (m) {                                            // long list: the whole warp
          const int src = __ffs(m) - 1; m &= m - 1;
          const int sj = __shfl_sync(FULL, s, src), ej = __shfl_sync(FULL, e, src);
          const float wj = __shfl_sync(FULL, w, src);
          for (int p = sj + lane; p < ej; p += 32) {
            const uint2 x = __ldg(a.qi + p);
            if ((int)x.x >= q_lo && (int)x.x < q_hi) atomicAdd(acc + ((int)x.x - q_lo), wj * __uint_as_float(x.y));
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < q_hi - q_lo; i += blockDim.x) {
        const float est = acc[i];
        if (est != 0.f) cand_test_emit(a, q_lo + i, c, est, cu, __ldg(a.q_nrm + q_lo + i), a.q_key && __ldg(a.q_key + q_lo + i) == ckey, n_cand);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { n_post += __shfl_down_sync(FULL, n_post, o); n_cand += __shfl_down_sync(FULL, n_cand, o); }
  if (lane == 0 && (n_post | n_cand)) { atomicAdd(&a.counters[C_POSTINGS], n_post); atomicAdd(&a.counters[C_CANDS], n_cand); }
}

// ------------------------------------------------------------------ K4: fp64 verify

// One WARP per pre-filter record: exact sparse dot of the query and the stored candidate in ascending
// dimension order, fp64, multiply and add rounded separately (CU:98-117; bit-identical to the CPU oracle).
// The lanes look the query's dimensions up in the candidate's (binary search, 32 at a time); the matched
// products are then added one by one in lane order = ascending dimension, the same chain on every lane.
// Applies `sim >= similarityThreshold` (IWA:93) and, for the as-built semantics R0, drops pairs whose
// shared dims all equal first(q) (IWA:89 + IWA:106-107).
static constexpr int VERIFY_WIN = 256;       // candidate dimensions staged per warp in shared memory

__global__ void __launch_bounds__(256) k_verify(const unsigned long long* counters, unsigned long long pf_cap,
                         const int32_t* __restrict__ pf_q, const int32_t* __restrict__ pf_c,
                         const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim, const double* __restrict__ q_val,
                         const int64_t* __restrict__ fwd_ptr, const int32_t* __restrict__ fwd_idx, const double* __restrict__ fwd_val,
                         const int32_t* __restrict__ gid, double thr, int sem_r0, const int32_t* __restrict__ first_dim,
                         int32_t* __restrict__ out_q, int32_t* __restrict__ out_c, double* __restrict__ out_sim,
                         unsigned long long* wcounters) {
  __shared__ int32_t s_win[8][VERIFY_WIN];
  unsigned long long n = counters[C_PF];
  if (n > pf_cap) n = pf_cap;
  const int lane = threadIdx.x & 31;
  int32_t* win = s_win[threadIdx.x >> 5];
  const unsigned long long wid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long r = wid; r < n; r += nw) {
    const int q = pf_q[r], c = pf_c[r];
    const int i0 = q_ptr[q], ie = q_ptr[q + 1];
    const int64_t j0 = fwd_ptr[c], je = fwd_ptr[c + 1];
    const int fd = (sem_r0 && first_dim) ? first_dim[q] : -1;
    // the candidate's dimensions go to shared memory once (coalesced) when they fit: the look-ups below then cost
    // shared-memory latency instead of a chain of dependent global loads
    const int nc = (int)min((int64_t)VERIFY_WIN + 1, je - j0);
    const bool staged = nc <= VERIFY_WIN;
    __syncwarp();
    if (staged) for (int k = lane; k < nc; k += 32) win[k] = fwd_idx[j0 + k];
    __syncwarp();
    double s = 0.0; int nonfirst = 0;
    for (int ib = i0; ib < ie; ib += 32) {
      const int i = ib + lane;
      double p = 0.0; bool hit = false;
      if (i < ie) {
        const int di = q_dim[i];
        int64_t pos;
        if (staged) {
          int lo = 0, hi = nc;                       // first k with win[k] >= di
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (win[mid] < di) lo = mid + 1; else hi = mid; }
          hit = lo < nc && win[lo] == di; pos = j0 + lo;
        } else {
          int64_t lo = j0, hi = je;                  // first j with fwd_idx[j] >= di
          while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (fwd_idx[mid] < di) lo = mid + 1; else hi = mid; }
          hit = lo < je && fwd_idx[lo] == di; pos = lo;
        }
        if (hit) { p = __dmul_rn(fwd_val[pos], q_val[i]); nonfirst += (di != fd); }
      }
      unsigned m = __ballot_sync(FULL, hit);
      while (m) {                                    // ascending lane = ascending dimension
        const int src = __ffs(m) - 1; m &= m - 1;
        s = __dadd_rn(s, __shfl_sync(FULL, p, src));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nonfirst += __shfl_xor_sync(FULL, nonfirst, o);
    if (lane == 0 && s >= thr) {
      atomicAdd(&wcounters[C_R1], 1ULL);
      if (!sem_r0 || nonfirst > 0) {
        const unsigned long long slot = atomicAdd(&wcounters[C_FINAL], 1ULL);
        out_q[slot] = q; out_c[slot] = gid[c]; out_sim[slot] = s;
      }
    }
  }
}

// ------------------------------------------------------------------ accumulator micro-benchmark

// mode 6: register FFMA, 8 independent chains per thread (FP32 FMA peak)
// mode 0: LDS/FFMA/STS random   1: same, consecutive addresses   2: ATOMS.ADD u32 random
// 3: ATOMS.ADD u32 consecutive  4: float atomicAdd (CAS loop) random   5: FFMA+F2I+ATOMS.ADD random
template <int MODE>
__global__ void k_microbench(int CR, int iters, unsigned* sink) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* row = smem + (size_t)warp * CR;
  unsigned* urow = reinterpret_cast<unsigned*>(row);
  for (int i = lane; i < CR; i += 32) row[i] = 0.f;
  __syncwarp();
  unsigned x = (blockIdx.x * 1315423911u) ^ (threadIdx.x * 2654435761u) ^ 12345u;
  const float w = 1.0f + lane * 1e-3f;
  if (MODE == 6) {
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = w + k;
    const float m = 1.0f + 1e-7f * lane, c = 1e-9f * warp;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], m, c);          // 32 FMAs per iteration (reported as 4 "updates" x 8)
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += f[k];
    if (sum == 0.12345f) sink[0] = 1u;
    return;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x = x * 1664525u + 1013904223u;
      int idx;   // CR is a power of two here
      if (MODE == 1 || MODE == 3) idx = (int)((__shfl_sync(FULL, x, 0) >> 8) & (unsigned)(CR - 32)) + lane;   // 32 consecutive slots
      else idx = (int)((x >> 8) & (unsigned)(CR - 1));                                                         // independent random slots
      if (MODE == 0 || MODE == 1) row[idx] = fmaf(w, 0.5f, row[idx]);
      else if (MODE == 2 || MODE == 3) atomicAdd(urow + idx, 3u);
      else if (MODE == 4) atomicAdd(row + idx, w);
      else atomicAdd(urow + idx, __float2uint_ru(fmaf(w, 1000.f, 0.5f)));
    }
  }
  __syncwarp();
  unsigned acc = 0;
  for (int i = lane; i < CR; i += 32) acc += urow[i];
  if (acc == 0x12345678u) sink[0] = acc;
}

}  // namespace apss
