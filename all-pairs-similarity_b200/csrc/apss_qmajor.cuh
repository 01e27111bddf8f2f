// apss_qmajor.cuh -- query-major scoring on the REDUCED index (include/apss.h `pruning` = 3).
//
// What the reference does per query (IWA:74-111): for every dimension of q walk that dimension's posting list
// and score the candidates met.  This file is that loop on the GPU, over the exactly reduced index of
// DESIGN.md 4b (only the components a vector could not keep out by the Cauchy-Schwarz bound are posted):
//
//   index      dimension-sorted CSR posting SEGMENTS (LSM style).  A segment covers a contiguous range of
//              shard-local vector ids: dir[D + 1] int32 offsets + 8-byte postings (local id : int32,
//              weight : fp32), ids ascending inside a list.  An insert builds one small segment from the
//              batch (k_seg_emit + radix sort by dimension + k_seg_dir) and segments of similar size are
//              merged by per-dimension concatenation (k_merge_dir / k_merge_copy: a streaming copy, no sort),
//              so a query term meets O(log(#batches)) lists.  "Appended in place on insert": nothing older
//              than the merged suffix is ever rewritten.
//   per batch  k_qm_count / scan / k_qm_emit cut every (query term, segment) list into PIECES of <= 64
//              postings (one 128-bit load per lane): (pointer, length, query weight * 2^F), contiguous per query.
//   scoring    k_score_qm: persistent, ONE CTA PER QUERY at a time.  The warps take the query's pieces
//              round-robin and stream them with 128-bit loads (two postings per lane, 512 B..1 KB per warp
//              instruction, coalesced); every posting is accumulated into a shared-memory open-addressing
//              hash table keyed by candidate id (ATOMS.CAS on the key, native u32 fixed-point ATOMS.ADD on
//              the value, every contribution rounded up).  The add returns the old sum, so the one update
//              that lifts a candidate over the query's smallest possible emission threshold appends the slot
//              to a short "hot list": after the walk only those slots are tested against
//                    estimate * (1 + guard band) + |q| * |c_unindexed|  >=  t
//              and handed to the fp64 verify kernel; the table is then cleared with plain vector stores (no
//              scan).  Queries whose lists exceed the table are scored in passes over candidate-id ranges
//              (ranges sized by an exact counting walk; pieces outside the range are skipped by their first
//              and last id).
//
// Counters are the ones the oracle's restatement (oracle_set_pruning, ALGO_FAST) defines: postings visited =
// sum of the list lengths walked, candidates = distinct (q, c) touched with c.key != q.key.
#pragma once
#include "apss_kernels.cuh"

namespace apss {

static constexpr int QM_MAXSEG = 40;      // segments a handle may hold (LSM: ~log2(#batches) in practice)
static constexpr int QM_HOT = 2048;       // hot-list capacity per pass (slots); more => full scan of the table
static constexpr int QM_TBL = 26624;      // table slots, one 1024-thread CTA per SM (keys + values = 208 KB of shared memory)
static constexpr int QM_CAP = 14336;      // list entries scored per pass (load factor <= 0.54)
static constexpr int QM_TBL2 = 13056;     // two 512-thread CTAs per SM: half the table each
static constexpr int QM_CAP2 = 7040;

struct QmItem { unsigned long long post; int32_t len; float wqs; };      // 16 B
static_assert(sizeof(QmItem) == 16, "QmItem is read with one 128-bit load");

struct SegList { const uint2* post[QM_MAXSEG]; const int32_t* dir[QM_MAXSEG]; int32_t n; };

__device__ __forceinline__ uint4 ld_stream4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ------------------------------------------------------------------ segment build (IWA:61-71 on the reduced index)

// warp per vector of the batch: (sort key = dimension, or D for a component that stays out of the index;
// value = posting).  The batch CSR is in ascending (row, dim) order, so a stable sort by dimension leaves
// every list in ascending id order.
__global__ void k_seg_emit(int n, int64_t n_old, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim,
                           const float* __restrict__ q_w, const uint8_t* __restrict__ skip, int D,
                           unsigned* __restrict__ keys, unsigned long long* __restrict__ vals) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (v >= n) return;
  const unsigned id = (unsigned)(n_old + v);
  for (int p = q_ptr[v] + (threadIdx.x & 31); p < q_ptr[v + 1]; p += 32) {
    keys[p] = (skip && skip[p]) ? (unsigned)D : (unsigned)q_dim[p];
    vals[p] = ((unsigned long long)__float_as_uint(q_w[p]) << 32) | id;        // uint2{x = id, y = weight}
  }
}

__global__ void k_seg_dir(const unsigned* __restrict__ keys, int nnz, int D, int32_t* __restrict__ dir) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > D) return;
  int lo = 0, hi = nnz;                      // first position with key >= d; dir[D] = number of postings
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < (unsigned)d) lo = mid + 1; else hi = mid; }
  dir[d] = lo;
}

// ------------------------------------------------------------------ segment merge (per-dimension concatenation)

struct MergeSrc { const uint2* post[QM_MAXSEG]; const int32_t* dir[QM_MAXSEG]; int32_t n; };

__global__ void k_merge_dir(int D, const MergeSrc m, int32_t* __restrict__ out_dir) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > D) return;
  int s = 0;
  for (int k = 0; k < m.n; ++k) s += m.dir[k][d];
  out_dir[d] = s;
}

// thread per posting of source `src` (sources are in age order = ascending id ranges)
__global__ void k_merge_copy(int src, int n_post, int D, const MergeSrc m, const int32_t* __restrict__ out_dir, uint2* __restrict__ out_post) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_post) return;
  const int32_t* __restrict__ dir = m.dir[src];
  int lo = 0, hi = D;                        // the dimension d with dir[d] <= p < dir[d + 1]: first d with dir[d + 1] > p
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(dir + mid + 1) > p) hi = mid; else lo = mid + 1; }
  const int d = lo;
  int dst = __ldg(out_dir + d) + (p - __ldg(dir + d));
  for (int k = 0; k < src; ++k) dst += __ldg(m.dir[k] + d + 1) - __ldg(m.dir[k] + d);
  out_post[dst] = m.post[src][p];
}

// ------------------------------------------------------------------ per batch: the pieces of every query

// A piece is a run of <= QM_PMAX postings of one list that starts on a 16-byte pair boundary (except the first piece
// of a list that starts on the upper half of a pair; segment buffers are 256-byte aligned): ONE bulk copy moves it.
// The bulk-copy engine handles a small request in about as many cycles as a large one (measured: ~80 cycles per
// request per SM), so pieces are as long as the stage allows, not one warp-load long.
static constexpr int QM_PMAX = 1024;
__device__ __forceinline__ int qm_pieces(int a, int b) {
  const int len = b - a;
  if (len <= 0) return 0;
  const int first = min(len, QM_PMAX - (a & 1));
  return 1 + (len - first + QM_PMAX - 1) / QM_PMAX;
}

// thread per query term: (pieces << 36 | postings) of its lists over all segments; ONE 64-bit scan then yields both
// the piece offsets and the per-query posting totals
__global__ void k_qm_count(int nnz, const int32_t* __restrict__ q_dim, const SegList sl, unsigned long long* __restrict__ cnt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nnz) return;
  unsigned long long c = 0;
  if (t < nnz) {
    const int d = q_dim[t];
    for (int s = 0; s < sl.n; ++s) {
      const int a = __ldg(sl.dir[s] + d), b = __ldg(sl.dir[s] + d + 1);
      c += ((unsigned long long)qm_pieces(a, b) << 36) + (unsigned long long)(b - a);
    }
  }
  cnt[t] = c;
}

__global__ void k_qm_emit(int nnz, const int32_t* __restrict__ q_dim, const float* __restrict__ q_w, float scale, const SegList sl,
                          const unsigned long long* __restrict__ off, QmItem* __restrict__ items, long long cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  long long o = (long long)(off[t] >> 36);
  if (o == (long long)(off[t + 1] >> 36)) return;
  const int d = q_dim[t];
  const float wqs = q_w[t] * scale;
  for (int s = 0; s < sl.n; ++s) {
    const int a = __ldg(sl.dir[s] + d), b = __ldg(sl.dir[s] + d + 1);
    for (int p = a; p < b; ++o) {
      const int len = min(b - p, QM_PMAX - (p & 1));
      if (o < cap) { QmItem it; it.post = (unsigned long long)(sl.post[s] + p); it.len = len; it.wqs = wqs; items[o] = it; }
      p += len;
    }
  }
}

// ------------------------------------------------------------------ scoring

struct QmArgs {
  const int32_t* q_ptr;          // pruned batch CSR
  const unsigned long long* item_off;   // [batch_nnz + 1] per query term: (piece offset << 36 | posting offset)
  const QmItem* items; long long item_cap;
  const float* q_nrm; const int64_t* q_key;
  const float* row_ub; const int64_t* c_key;
  const int32_t* row_dfmin;      // per stored vector: smallest document frequency among its un-indexed components
  const float* q_bkt;            // per query: 32 suffix sums of squared weights by document-frequency bucket (see k_qm_qnorms)
  int64_t n_rows;                // stored vectors visible to this batch
  int64_t q_local_base;          // shard-local id of query 0 when the batch was indexed in this call, else -1
  int32_t nq;
  float thr, band1;              // t (rounded down) and 1 + guard band of the fp32 estimate
  float scale, inv_scale;        // 2^F, 2^-F
  float cu_max;                  // upper bound of every row_ub[]
  int32_t cap;                   // list entries scored per pass (QM_CAP; tests lower it to reach the ranged passes)
  int32_t* out_q; int32_t* out_c; float* out_est; unsigned long long out_cap;
  unsigned long long* counters;
  int32_t* deferred; int32_t deferred_cap;   // queries the pipelined kernel left to the ranged kernel (count in C_HEAVY)
  unsigned* hot_used; unsigned n_chunks;    // entries written per QP_CHUNK-sized chunk of the hot-candidate buffer (zeroed by the host)
  int32_t* hot_q; int32_t* hot_c; float* hot_est; unsigned hot_cap;     // pipelined kernel: candidates that crossed the query's
                                 // coarse threshold, written in per-CTA chunks reserved on C_HOTN (q = -1: unused entry)
  int32_t from_list;             // k_score_qm: 0 = take every query from the cursor, 1 = take the deferred list
  int32_t sizef;                 // pipelined kernel: table slots per list entry, in halves (6 = 3.0: load factor <= 1/3 where the table allows)
  int32_t dry;                   // measurement only (APSS_QM_DRY, bits): 1 = no table updates, 2 = no copies, 4 = L2 prefetch of the
                                 // next stage's pieces while the producer waits (measured: slower -- it doubles the bulk requests)
};

// ---- the candidate test.  dot(q, c) = dot(q, c_indexed) + dot(q, c_unindexed), and by Cauchy-Schwarz on the un-indexed
// dimensions U_c alone  dot(q, c_unindexed) <= |q restricted to U_c| * |c_unindexed|.  U_c holds c's most frequent
// dimensions: every d in U_c had df(d) >= dfmin(c) when c was indexed, and frequencies only grow, so U_c is inside
// {d : df_now(d) >= dfmin(c)} and the query's norm over THAT set bounds the first factor.  It is tabulated per query over
// buckets b(df) = floor(log2(df + 1)) (suffix sums, rounded up): frequent dimensions carry small IDF weights, so this is
// far below |q| -- ~40x fewer records reach the fp64 verify kernel than with the plain |q| |c_U| bound, same pairs.
__device__ __forceinline__ int qm_df_bucket(int df) { return 31 - __clz((int)((unsigned)max(df, 0) + 1u)); }

__device__ __forceinline__ bool qm_candidate_passes(const QmArgs& a, int q, long long c, float est) {
  const float cu = __ldg(a.row_ub + c);
  float ub = 0.f;
  if (cu > 0.f) {
    const float s2 = __ldg(a.q_bkt + (size_t)q * 32 + qm_df_bucket(__ldg(a.row_dfmin + c)));
    ub = __fmul_ru(cu, __fmul_ru(__fsqrt_ru(s2), 1.000001f));
  }
  return __fmaf_ru(est, a.band1, ub) >= a.thr;
}

// warp per query: suffix sums of the squared fp32 weights over the document-frequency buckets of its dimensions
__global__ void k_qm_qnorms(int n, const int32_t* __restrict__ q_ptr, const int32_t* __restrict__ q_dim, const float* __restrict__ q_w,
                            const int32_t* __restrict__ df, float* __restrict__ q_bkt) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n) return;
  float mine = 0.f;                                   // lane b accumulates bucket b
  for (int p0 = q_ptr[v]; p0 < q_ptr[v + 1]; p0 += 32) {
    const int p = p0 + lane;
    int b = -1; float w2 = 0.f;
    if (p < q_ptr[v + 1]) { b = qm_df_bucket(__ldg(df + q_dim[p])); const float w = q_w[p]; w2 = __fmul_ru(w, w); }
    for (int k = 0; k < 32; ++k) {                    // (batch-sized work: ~100 components per query)
      const int bk = __shfl_sync(FULL, b, k); const float wk = __shfl_sync(FULL, w2, k);
      if (bk == lane) mine = __fadd_ru(mine, wk);
    }
  }
  float suf = mine;                                   // suffix sum over lanes >= me, every add rounded up
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_down_sync(FULL, suf, o); if (lane + o < 32) suf = __fadd_ru(suf, t); }
  q_bkt[(size_t)v * 32 + lane] = __fmul_ru(suf, 1.000001f);
}

template <int NT>
__device__ __forceinline__ long long qm_block_sum(long long v, long long* red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  __syncthreads();                            // red[] may still be read from the previous call
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  long long s = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) s += red[w];
  return s;
}

template <int NT, int TBL, bool DUPKEYS>
__global__ void __launch_bounds__(NT, 2048 / NT / 2 * 1) k_score_qm(const QmArgs a) {
  extern __shared__ __align__(16) unsigned qm_smem[];
  unsigned* keys = qm_smem;
  unsigned* vals = qm_smem + TBL;
  int* hot = reinterpret_cast<int*>(qm_smem + 2 * TBL);
  __shared__ long long red[NT / 32];
  __shared__ int s_q;
  __shared__ unsigned s_hot_n;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid * 4; i < 2 * TBL; i += NT * 4) *reinterpret_cast<uint4*>(qm_smem + i) = make_uint4(0, 0, 0, 0);
  unsigned long long n_post = 0; unsigned n_cand = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      if (a.from_list) {
        const unsigned long long k = atomicAdd(&a.counters[C_HEAVY_TOT], 1ULL);
        s_q = k < min(a.counters[C_HEAVY], (unsigned long long)a.deferred_cap) ? a.deferred[k] : a.nq;
      } else s_q = (int)atomicAdd(&a.counters[C_WORK], 1ULL);
    }
    __syncthreads();
    const int q = s_q;
    if (q >= a.nq) break;
    const int t0 = __ldg(a.q_ptr + q), t1 = __ldg(a.q_ptr + q + 1);
    if (t0 == t1) continue;
    const unsigned long long o0 = __ldg(a.item_off + t0), o1 = __ldg(a.item_off + t1);
    const long long i0 = (long long)(o0 >> 36), i1 = min((long long)(o1 >> 36), a.item_cap);
    if (i0 >= i1) continue;
    const long long total = (long long)((o1 & 0xfffffffffULL) - (o0 & 0xfffffffffULL));
    if (tid == 0) n_post += (unsigned long long)total;
    const float qn = __ldg(a.q_nrm + q);
    const unsigned self = a.q_local_base >= 0 ? (unsigned)(a.q_local_base + q) : 0xffffffffu;
    long long qkey = 0;
    if (DUPKEYS) qkey = __ldg(a.q_key + q);
    // smallest fixed-point sum a candidate of this query needs before the exact test can pass (0: test them all)
    unsigned thr_fix = 0;
    {
      const float em = (a.thr - a.cu_max * qn * 1.000001f) / a.band1;
      if (em > 0.f) thr_fix = (unsigned)fminf(floorf(em * a.scale * 0.99999f), 4294967040.f);
    }
    const bool scan_all = DUPKEYS || thr_fix == 0;
    // the query's pieces, a contiguous share per warp
    const int per = (int)((i1 - i0 + NW - 1) / NW);
    const long long w_lo = i0 + (long long)warp * per, w_hi = min(i1, w_lo + per);

    // one pass: candidates with lo <= id < hi (ranged) or all of them; count_only: how many entries the range holds.
    // A warp takes its pieces one after the other and streams each in 1 KB chunks (two postings per lane, 128-bit loads).
    auto walk = [&](const bool ranged, const bool count_only, const unsigned lo, const unsigned hi, const unsigned size) -> long long {
      long long cnt = 0;
      for (long long i = w_lo; i < w_hi; ++i) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(a.items + i));          // same address on every lane
        const unsigned long long pp = ((unsigned long long)raw.y << 32) | raw.x;
        const int len = (int)raw.z; const float wqs = __uint_as_float(raw.w);
        const int odd = (int)((pp >> 3) & 1ULL);                                       // the piece starts on the upper half of a 16-byte pair
        const uint4* p4 = reinterpret_cast<const uint4*>(pp - 8ULL * odd);
        for (int j0 = -odd; j0 < len; j0 += 64) {
          const int j = j0 + 2 * lane;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (j + 1 >= 0 && j < len) v = ld_stream4(p4 + ((j + odd) >> 1));
          unsigned cc[2], slot[2], contrib[2]; bool on[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            cc[u] = u ? v.z : v.x;
            const float w = __uint_as_float(u ? v.w : v.y);
            on[u] = j + u >= 0 && j + u < len && cc[u] != self;
            if (ranged) on[u] = on[u] && cc[u] >= lo && cc[u] < hi;
            contrib[u] = __float2uint_ru(__fmul_ru(w, wqs));
            slot[u] = __umulhi(cc[u] * 0x9E3779B1u, size);
          }
          if (count_only) { cnt += (int)on[0] + (int)on[1]; continue; }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            const unsigned k1 = cc[u] + 1u;
            unsigned sl_ = slot[u];
            for (;;) {
              const unsigned o = atomicCAS(keys + sl_, 0u, k1);
              if (o == 0u) { ++n_cand; break; }
              if (o == k1) break;
              if (++sl_ == size) sl_ = 0;
            }
            const unsigned prev = atomicAdd(vals + sl_, contrib[u]);
            if (!scan_all && prev < thr_fix && prev + contrib[u] >= thr_fix) {
              const unsigned h = atomicAdd(&s_hot_n, 1u);
              if (h < (unsigned)QM_HOT) hot[h] = (int)sl_;
            }
          }
        }
      }
      return cnt;
    };
    auto test_emit = [&](const unsigned k, const unsigned v) {
      const long long c = (long long)k - 1;
      if (DUPKEYS) { if (__ldg(a.c_key + c) == qkey) return; ++n_cand; }
      const float est = __uint2float_ru(v) * a.inv_scale;
      if (qm_candidate_passes(a, q, c, est)) {
        const unsigned long long slot = atomicAdd(&a.counters[C_PF], 1ULL);
        if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)c; a.out_est[slot] = est; }
      }
    };
    auto pass = [&](const bool ranged, const unsigned lo, const unsigned hi, const long long entries) {
      const unsigned size = (unsigned)min((long long)TBL, max(256LL, (3 * entries + 31) & ~31LL));
      if (tid == 0) s_hot_n = 0u;
      __syncthreads();
      const unsigned cand0 = n_cand;
      walk(ranged, false, lo, hi, size);
      if (DUPKEYS) n_cand = cand0;             // counted in the scan instead (same-key candidates do not count)
      __syncthreads();
      const unsigned nh = s_hot_n;
      if (scan_all || nh > (unsigned)QM_HOT) {
        for (unsigned i = tid; i < size; i += NT) { const unsigned k = keys[i]; if (k) test_emit(k, vals[i]); }
      } else {
        for (unsigned e = tid; e < nh; e += NT) { const int s = hot[e]; test_emit(keys[s], vals[s]); }
      }
      __syncthreads();
      for (unsigned i = tid * 4; i < size; i += NT * 4) {
        *reinterpret_cast<uint4*>(keys + i) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(vals + i) = make_uint4(0, 0, 0, 0);
      }
    };

    if (total <= a.cap) pass(false, 0u, 0xffffffffu, total);
    else {
      long long lo = 0;
      while (lo < a.n_rows) {
        long long width = (long long)((double)a.cap * 0.7 * (double)a.n_rows / (double)total);
        long long hi = min((long long)a.n_rows, lo + max(1LL, width));
        long long cnt = qm_block_sum<NT>(walk(true, true, (unsigned)lo, (unsigned)hi, 1u), red, tid);
        while (cnt > a.cap && hi - lo > a.cap) {
          hi = lo + (hi - lo) / 2;
          cnt = qm_block_sum<NT>(walk(true, true, (unsigned)lo, (unsigned)hi, 1u), red, tid);
        }
        if (cnt) pass(true, (unsigned)lo, (unsigned)hi, min(cnt, hi - lo));
        lo = hi;
      }
    }
  }
  unsigned long long nc = n_cand;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nc += __shfl_down_sync(FULL, nc, o);
  if (lane == 0 && nc) atomicAdd(&a.counters[C_CANDS], nc);
  if (tid == 0 && n_post) atomicAdd(&a.counters[C_POSTINGS], n_post);
}

// ------------------------------------------------------------------ the pipelined kernel (bulk-async producer / consumers)

// k_score_qm above pays a chain of dependent global round trips per query (cursor -> q_ptr -> offsets -> piece
// descriptors -> postings) with the whole CTA in lock step.  In k_score_qm_flat (below) one PRODUCER warp runs ahead of
// 31 CONSUMER warps through a two-stage shared-memory ring guarded by mbarrier full / empty pairs; the postings arrive by
// cp.async.bulk (TMA bulk copy), the consumers only ever touch shared memory.
static constexpr int QP_NCW = 31;                 // consumer warps
static constexpr int QP_STAGES = 2;
static constexpr int QP_QG = 8;                   // queries whose set-up chain the producer runs at once (one per lane)
static constexpr int QP_TBL = 19712;              // table slots: 232448 - ring - hot list - metadata, in 8-byte slots
static constexpr int QP_CAP = 13568;              // entries per query handled here (load factor <= 0.69: two-choice hashing keeps probes short)
static constexpr int QP_F_FIRST = 1, QP_F_LAST = 2, QP_F_END = 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "QP_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra QP_DONE_%=;\n\t"
      "bra QP_WAIT_%=;\n\t"
      "QP_DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// predicated shared-memory atomics: no branch around the instruction (the consumer loop is issue bound)
__device__ __forceinline__ unsigned atoms_cas_if(unsigned addr, unsigned val, bool on) {
  unsigned old = 0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.cas.b32 %0, [%1], 0, %2;\n\t}" : "+r"(old) : "r"(addr), "r"(val), "r"((unsigned)on) : "memory");
  return old;
}
__device__ __forceinline__ unsigned atoms_add_if(unsigned addr, unsigned val, bool on) {
  unsigned old = 0;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}" : "+r"(old) : "r"(addr), "r"(val), "r"((unsigned)on) : "memory");
  return old;
}
static constexpr unsigned QP_CHUNK = 32768;        // entries of the hot-candidate buffer a CTA reserves at a time (>= QP_TBL)

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(QP_NCW * 32) : "memory"); }

// ------------------------------------------------------------------ the flat-stage variant (default)

// A first version gave every piece a fixed 512-byte slot and every consumer warp two slots per stage; with the short
// lists of the young posting segments the slots were ~40 % full and the consumer loop spent most of its issue slots on
// nothing.  Here the producer PACKS the pieces of a stage back to back (16-byte units, a
// running offset from a warp prefix sum) and publishes a piece table (start unit, length, parity, weight) plus, per
// 512-byte window, the first piece that reaches into it.  A consumer warp takes whole windows: lane = one 16-byte unit =
// two postings, finds its piece with a short forward walk from the window's first piece, and every lane has work.
static constexpr int QF_PPL = 3;                  // pieces per producer lane and stage
static constexpr int QF_NP = 32 * QF_PPL;         // pieces per stage, at most
static constexpr int QF_UNITS = 2 * QP_NCW * 32;  // 16-byte units per stage (two windows per consumer warp) = 31 744 B
static constexpr int QF_STAGE_BYTES = QF_UNITS * 16;
struct QfHdr { int q, np, flags, total; float qn; int units; long long qkey; };      // 32 B per stage
static constexpr size_t QF_META = (size_t)(QF_NP + 1) * 8 + 64 + sizeof(QfHdr);     // piece table + window table + header
static constexpr size_t QF_SMEM = (size_t)2 * QP_TBL * 4 + (size_t)QM_HOT * 4 + (size_t)QP_STAGES * QF_STAGE_BYTES +
                                  (size_t)QP_STAGES * QF_META + (size_t)QP_STAGES * 16;

template <bool DUPKEYS>
__global__ void __launch_bounds__(1024, 1) k_score_qm_flat(const QmArgs a) {
  extern __shared__ __align__(128) unsigned char qp_smem[];
  unsigned* keys = reinterpret_cast<unsigned*>(qp_smem);
  unsigned* vals = keys + QP_TBL;
  int* hot = reinterpret_cast<int*>(vals + QP_TBL);
  unsigned char* ring = reinterpret_cast<unsigned char*>(hot + QM_HOT);                  // QP_STAGES x QF_STAGE_BYTES, 128-byte aligned
  unsigned char* metab = ring + (size_t)QP_STAGES * QF_STAGE_BYTES;                      // per stage: tab[QF_NP + 1] | first[64] | header
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(metab + (size_t)QP_STAGES * QF_META);
  auto tab_of = [&](int st) { return reinterpret_cast<uint2*>(metab + (size_t)st * QF_META); };
  auto first_of = [&](int st) { return metab + (size_t)st * QF_META + (size_t)(QF_NP + 1) * 8; };
  auto hdr_of = [&](int st) { return reinterpret_cast<QfHdr*>(metab + (size_t)st * QF_META + (size_t)(QF_NP + 1) * 8 + 64); };
  __shared__ unsigned s_hot_n, s_out_n, s_chunk_pos, s_chunk_end, s_chunk_next;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid * 4; i < 2 * QP_TBL; i += 1024 * 4) *reinterpret_cast<uint4*>(keys + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    s_hot_n = 0u;
    for (int s = 0; s < QP_STAGES; ++s) { mbar_init(smem_u32(bars + s), 1u); mbar_init(smem_u32(bars + QP_STAGES + s), (unsigned)QP_NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------ producer
    int stage = 0; unsigned phase = 0; unsigned long long n_post = 0;
    auto load_desc = [&](const long long base, const long long end, uint4 (&d)[QF_PPL]) {      // lane l: pieces 3l, 3l + 1, 3l + 2
#pragma unroll
      for (int r = 0; r < QF_PPL; ++r) {
        const long long k = base + lane * QF_PPL + r;
        d[r] = make_uint4(0, 0, 0, 0);
        if (k < end) d[r] = __ldg(reinterpret_cast<const uint4*>(a.items + k));
      }
    };
    for (;;) {
      int qbase = 0;
      if (lane == 0) qbase = (int)atomicAdd(&a.counters[C_WORK], (unsigned long long)QP_QG);
      qbase = __shfl_sync(FULL, qbase, 0);
      if (qbase >= a.nq) break;
      // ---- one query per lane: the whole set-up chain in parallel
      const int myq = qbase + lane;
      long long my_i0 = 0, my_i1 = 0, my_total = 0, my_qkey = 0; float my_qn = 0.f; bool my_ok = false;
      if (lane < QP_QG && myq < a.nq) {
        const int t0 = __ldg(a.q_ptr + myq), t1 = __ldg(a.q_ptr + myq + 1);
        if (t0 != t1) {
          const unsigned long long o0 = __ldg(a.item_off + t0), o1 = __ldg(a.item_off + t1);
          my_i0 = (long long)(o0 >> 36); my_i1 = min((long long)(o1 >> 36), a.item_cap);
          my_total = (long long)((o1 & 0xfffffffffULL) - (o0 & 0xfffffffffULL));
          my_ok = my_i0 < my_i1;
          if (my_ok && my_total > a.cap) {                     // too long for one table pass: the ranged kernel takes it
            const unsigned long long k = atomicAdd(&a.counters[C_HEAVY], 1ULL);
            if (k < (unsigned long long)a.deferred_cap) a.deferred[k] = myq;
            my_ok = false;
          }
          if (my_ok) {
            n_post += (unsigned long long)my_total;
            my_qn = __ldg(a.q_nrm + myq);
            if (DUPKEYS) my_qkey = __ldg(a.q_key + myq);
          }
        }
      }
      unsigned okmask = __ballot_sync(FULL, my_ok);
      if (!okmask) continue;
      int g = __ffs(okmask) - 1;
      long long base = __shfl_sync(FULL, my_i0, g), gi1 = __shfl_sync(FULL, my_i1, g);
      uint4 d[QF_PPL];
      load_desc(base, gi1, d);
      while (g >= 0) {
        // ---- pack: units per piece, running offsets, how many pieces fit the stage
        unsigned un[QF_PPL], st[QF_PPL], lsum = 0;
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r) {
          un[r] = (base + lane * QF_PPL + r < gi1) ? ((((d[r].x >> 3) & 1u) + d[r].z + 1u) >> 1) : 0u;
          lsum += un[r];
        }
        unsigned incl = lsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        st[0] = incl - lsum;
#pragma unroll
        for (int r = 1; r < QF_PPL; ++r) st[r] = st[r - 1] + un[r - 1];
        int fit = 0; unsigned endu = 0;
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r) if (un[r] && st[r] + un[r] <= (unsigned)QF_UNITS) { ++fit; endu = st[r] + un[r]; }
        int np = fit;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { np += __shfl_xor_sync(FULL, np, o); endu = max(endu, __shfl_xor_sync(FULL, endu, o)); }
        const bool first = base == __shfl_sync(FULL, my_i0, g), last = base + np >= gi1;
        // what comes next: the query's next stage, or the first stage of the group's next query
        int ng = g; long long nbase = base + np, ngi1 = gi1;
        if (last) {
          okmask &= ~(1u << g);
          ng = okmask ? __ffs(okmask) - 1 : -1;
          const int src = ng >= 0 ? ng : 0;
          nbase = __shfl_sync(FULL, my_i0, src); ngi1 = __shfl_sync(FULL, my_i1, src);
        }
        uint4 dn[QF_PPL];
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r) dn[r] = make_uint4(0, 0, 0, 0);
        if (ng >= 0) load_desc(nbase, ngi1, dn);
        const int qv = qbase + g; const long long tv = __shfl_sync(FULL, my_total, g);
        const float qnv = __shfl_sync(FULL, my_qn, g); const long long qkv = __shfl_sync(FULL, my_qkey, g);
        // the stage's postings start their way from HBM to L2 now, while the producer waits for a ring slot
        if ((a.dry & 6) == 4)
#pragma unroll
          for (int r = 0; r < QF_PPL; ++r)
            if (un[r] && st[r] + un[r] <= (unsigned)QF_UNITS)
              bulk_prefetch_l2(reinterpret_cast<const void*>((((unsigned long long)d[r].y << 32) | d[r].x) & ~15ULL), un[r] * 16u);
        // ---- fill the stage
        mbar_wait(smem_u32(bars + QP_STAGES + stage), phase ^ 1u);          // the consumers have released this stage
        uint2* tab = tab_of(stage); unsigned char* fw = first_of(stage);
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r) {
          const int k = lane * QF_PPL + r;
          if (un[r] && st[r] + un[r] <= (unsigned)QF_UNITS) {
            tab[k] = make_uint2(st[r] | (d[r].z << 12) | (((d[r].x >> 3) & 1u) << 24), d[r].w);
            for (unsigned w = (st[r] + 31u) >> 5; (w << 5) < st[r] + un[r]; ++w) fw[w] = (unsigned char)k;      // windows that start inside this piece
          }
        }
        if (lane == 0) {
          tab[np] = make_uint2(0xfffu, 0u);                  // sentinel: starts beyond every unit
          QfHdr hh; hh.q = qv; hh.np = np; hh.total = (int)tv; hh.qn = qnv; hh.units = (int)endu; hh.qkey = qkv;
          hh.flags = (first ? QP_F_FIRST : 0) | (last ? QP_F_LAST : 0);
          *hdr_of(stage) = hh;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_expect_tx(smem_u32(bars + stage), (a.dry & 2) ? 0u : endu * 16u);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r)
          if (un[r] && st[r] + un[r] <= (unsigned)QF_UNITS && !(a.dry & 2)) {
            const unsigned long long pp = ((unsigned long long)d[r].y << 32) | d[r].x;
            bulk_g2s(smem_u32(ring + (size_t)stage * QF_STAGE_BYTES + (size_t)st[r] * 16), reinterpret_cast<const void*>(pp & ~15ULL), un[r] * 16u, smem_u32(bars + stage));
          }
        if (++stage == QP_STAGES) { stage = 0; phase ^= 1u; }
        g = ng; base = nbase; gi1 = ngi1;
#pragma unroll
        for (int r = 0; r < QF_PPL; ++r) d[r] = dn[r];
      }
    }
    // no more queries: one empty END stage
    mbar_wait(smem_u32(bars + QP_STAGES + stage), phase ^ 1u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_post += __shfl_down_sync(FULL, n_post, o);
    if (lane == 0) {
      QfHdr hh; hh.q = a.nq; hh.np = 0; hh.total = 0; hh.qn = 0.f; hh.units = 0; hh.qkey = 0; hh.flags = QP_F_END;
      *hdr_of(stage) = hh;
      mbar_arrive(smem_u32(bars + stage));
      if (n_post) atomicAdd(&a.counters[C_POSTINGS], n_post);
    }
    return;
  }

  // -------------------------------------------------- consumers
  const int cw = warp - 1, ctid = tid - 32;
  constexpr int CT = QP_NCW * 32;
  const unsigned keys_s = smem_u32(keys), vals_s = smem_u32(vals);
  if (ctid == 0) {               // two chunks of the hot-candidate buffer up front; a new one is reserved whenever one is taken
    s_chunk_pos = (unsigned)atomicAdd(&a.counters[C_HOTN], (unsigned long long)QP_CHUNK);
    s_chunk_end = s_chunk_pos + QP_CHUNK;
    s_chunk_next = (unsigned)atomicAdd(&a.counters[C_HOTN], (unsigned long long)QP_CHUNK);
    s_out_n = 0u;
  }
  // (consumer thread 0 only) close the current chunk -- its fill goes to the chunk directory k_qm_filter reads -- and
  // move to the one reserved ahead
  auto close_chunk = [&]() {
    const unsigned c0 = s_chunk_end - QP_CHUNK, ci = c0 / QP_CHUNK;
    if (ci < a.n_chunks) a.hot_used[ci] = s_chunk_pos - c0;
  };
  auto next_chunk = [&]() {
    close_chunk();
    s_chunk_pos = s_chunk_next; s_chunk_end = s_chunk_next + QP_CHUNK;
    s_chunk_next = (unsigned)atomicAdd(&a.counters[C_HOTN], (unsigned long long)QP_CHUNK);
  };
  int stage = 0; unsigned phase = 0; unsigned n_cand = 0;
  unsigned size = 256u, thr_fix = 0u, self = 0xffffffffu; bool scan_all = true; float qn = 0.f; long long qkey = 0; int q = 0;
  for (;;) {
    mbar_wait(smem_u32(bars + stage), phase);
    const QfHdr hh = *hdr_of(stage);
    if (hh.flags & QP_F_END) break;
    if (hh.flags & QP_F_FIRST) {
      q = hh.q; qn = hh.qn; qkey = hh.qkey;
      size = (unsigned)min((long long)QP_TBL, max(256LL, ((long long)a.sizef * hh.total / 2 + 31) & ~31LL));
      self = a.q_local_base >= 0 ? (unsigned)(a.q_local_base + q) : 0xffffffffu;
      const float em = (a.thr - a.cu_max * qn * 1.000001f) / a.band1;
      thr_fix = em > 0.f ? (unsigned)fminf(floorf(em * a.scale * 0.99999f), 4294967040.f) : 0u;
      scan_all = DUPKEYS || thr_fix == 0u;
    }
    if (cw * 32 >= hh.units) {                                  // a short stage: no window for this warp, just release it --
      __syncwarp();                                             // once EVERY lane has read the header (the producer rewrites it)
      if (lane == 0) mbar_arrive(smem_u32(bars + QP_STAGES + stage));
    } else {
      const uint2* tab = tab_of(stage); const unsigned char* fw = first_of(stage);
      uint4 v[2]; int jj[2], ln[2]; float ws[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const unsigned u = (unsigned)(cw + QP_NCW * r) * 32u + (unsigned)lane;       // my 16-byte unit of the stage
        v[r] = make_uint4(0, 0, 0, 0); jj[r] = 0; ln[r] = 0; ws[r] = 0.f;
        if (u < (unsigned)hh.units) {
          int p = fw[cw + QP_NCW * r];
          uint2 cur = tab[p], nxt = tab[p + 1];
          while ((nxt.x & 0xfffu) <= u) { cur = nxt; ++p; nxt = tab[p + 1]; }        // forward walk: few pieces reach into one window
          ln[r] = (int)((cur.x >> 12) & 0xfffu); ws[r] = __uint_as_float(cur.y);
          jj[r] = 2 * (int)(u - (cur.x & 0xfffu)) - (int)(cur.x >> 24);
          v[r] = *reinterpret_cast<const uint4*>(ring + (size_t)stage * QF_STAGE_BYTES + (size_t)u * 16);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bars + QP_STAGES + stage));      // the postings are in registers: release the stage
      unsigned cc[4], slot[4], contrib[4], old[4]; bool on[4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int e = 2 * r + u;
          cc[e] = u ? v[r].z : v[r].x;
          const float w = __uint_as_float(u ? v[r].w : v[r].y);
          on[e] = jj[r] + u >= 0 && jj[r] + u < ln[r] && cc[e] != self && !(a.dry & 3);
          contrib[e] = __float2uint_ru(__fmul_ru(w, ws[r]));
          slot[e] = __umulhi(cc[e] * 0x9E3779B1u, size);
        }
#pragma unroll
      for (int e = 0; e < 4; ++e) old[e] = atoms_cas_if(keys_s + slot[e] * 4u, cc[e] + 1u, on[e]);
      // second choice, still straight-line: a slot held by another candidate sends the key to an independent second hash
      // before any probing loop -- the divergent loops below (the warp waits for its slowest lane) become rare and short
      bool c2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        c2[e] = on[e] && old[e] != 0u && old[e] != cc[e] + 1u && !(a.dry & 8);      // (dry bit 8: measurement, no second choice)
        const unsigned s2 = __umulhi((cc[e] ^ 0x5bd1e995u) * 0x85EBCA6Bu, size);
        if (c2[e]) slot[e] = s2;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { const unsigned o2 = atoms_cas_if(keys_s + slot[e] * 4u, cc[e] + 1u, c2[e]); if (c2[e]) old[e] = o2; }
      // both taken: linear probing from the second slot.  ONE loop for the (rarely more than one) keys a lane still has
      // to place -- four loops, one per register slot, each kept the whole warp for a lane or two
      unsigned pend = 0u;
#pragma unroll
      for (int e = 0; e < 4; ++e) pend |= (on[e] && old[e] != 0u && old[e] != cc[e] + 1u) ? (1u << e) : 0u;
      if (pend) {
        auto pick = [&](const unsigned (&x)[4], const int e) { return e == 0 ? x[0] : e == 1 ? x[1] : e == 2 ? x[2] : x[3]; };
        int e = __ffs((int)pend) - 1;
        unsigned k1 = pick(cc, e) + 1u, sl_ = pick(slot, e);
        for (;;) {
          if (++sl_ == size) sl_ = 0;
          const unsigned o = atomicCAS(keys + sl_, 0u, k1);
          if (o == 0u || o == k1) {
#pragma unroll
            for (int r = 0; r < 4; ++r) if (r == e) { old[r] = o; slot[r] = sl_; }
            pend &= pend - 1u;
            if (!pend) break;
            e = __ffs((int)pend) - 1;
            k1 = pick(cc, e) + 1u; sl_ = pick(slot, e);
          }
        }
      }
      unsigned prev[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!DUPKEYS) n_cand += (unsigned)(on[e] && old[e] == 0u);      // (caller keys: counted in the scan; same-key candidates do not count)
        prev[e] = atoms_add_if(vals_s + slot[e] * 4u, contrib[e], on[e]);
      }
      bool anyhot = false;
#pragma unroll
      for (int e = 0; e < 4; ++e) anyhot |= on[e] && prev[e] < thr_fix && prev[e] + contrib[e] >= thr_fix;
      if (anyhot && !scan_all) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (on[e] && prev[e] < thr_fix && prev[e] + contrib[e] >= thr_fix) {
            const unsigned h = atomicAdd(&s_hot_n, 1u);
            if (h < (unsigned)QM_HOT) hot[h] = (int)slot[e];
          }
      }
    }
    if (hh.flags & QP_F_LAST) {
      // ---- consumer-only epilogue.  No global round trip here: the candidates that crossed the coarse threshold go,
      // with their final sums, to this CTA's chunk of the hot-candidate buffer; k_qm_filter applies the exact test.
      consumer_bar();
      const unsigned nh = s_hot_n;
      const bool full = scan_all || nh > (unsigned)QM_HOT;
      if (full) {
        if (ctid == 0 && s_chunk_pos + size > s_chunk_end) next_chunk();
        consumer_bar();
      }
      const unsigned cpos = s_chunk_pos;
      auto put = [&](const unsigned at, const unsigned k, const unsigned vv) {
        const unsigned o = cpos + at;
        if (o < a.hot_cap) { a.hot_q[o] = q; a.hot_c[o] = (int32_t)(k - 1u); a.hot_est[o] = __uint2float_ru(vv) * a.inv_scale; }
      };
      if (full) {
        for (unsigned i = ctid; i < size; i += CT) {
          const unsigned k = keys[i];
          if (!k) continue;
          if (DUPKEYS) { if (__ldg(a.c_key + (k - 1u)) == qkey) continue; ++n_cand; }
          put(atomicAdd(&s_out_n, 1u), k, vals[i]);
        }
      } else {
        for (unsigned e = ctid; e < nh; e += CT) { const int sl_ = hot[e]; put(e, keys[sl_], vals[sl_]); }
      }
      consumer_bar();
      for (unsigned i = ctid * 4; i < size; i += CT * 4) {
        *reinterpret_cast<uint4*>(keys + i) = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(vals + i) = make_uint4(0, 0, 0, 0);
      }
      if (ctid == 0) {
        s_chunk_pos += full ? s_out_n : nh; s_out_n = 0u; s_hot_n = 0u;
        if (s_chunk_pos + (unsigned)QM_HOT > s_chunk_end) next_chunk();    // the next query's hot list always fits
      }
      consumer_bar();
    }
    if (++stage == QP_STAGES) { stage = 0; phase ^= 1u; }
  }
  if (ctid == 0) close_chunk();          // (every epilogue ended with a consumer barrier: s_chunk_pos is final)
  unsigned long long nc = n_cand;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nc += __shfl_down_sync(FULL, nc, o);
  if (lane == 0 && nc) atomicAdd(&a.counters[C_CANDS], nc);
}

// The exact candidate test of the pipelined kernel, moved out of its epilogue: one thread per entry of the hot-candidate
// buffer (q = -1: never written),  estimate * (1 + guard band) + |q| * |c_unindexed| >= t,  survivors compacted into
// the record list of the fp64 verify kernel (warp-aggregated).
static constexpr int QF_PARTS = 8;          // blocks per chunk
__global__ void k_qm_filter(const QmArgs a) {
  // blockIdx.x = chunk * QF_PARTS + part; the chunk directory says how many entries the scoring kernel wrote there
  const unsigned ci = blockIdx.x / QF_PARTS, part = blockIdx.x % QF_PARTS;
  if (ci >= a.n_chunks) return;
  const unsigned used = min(a.hot_used[ci], (unsigned)QP_CHUNK);
  constexpr unsigned PER = QP_CHUNK / QF_PARTS;
  const unsigned lo = part * PER, hi = min(used, lo + PER);
  const int lane = threadIdx.x & 31;
  const unsigned long long base0 = (unsigned long long)ci * QP_CHUNK;
  for (unsigned e0 = lo + (threadIdx.x & ~31u); e0 < hi; e0 += blockDim.x) {
    const unsigned e = e0 + lane;
    bool pass = false; int q = -1, c = 0; float est = 0.f;
    if (e < hi) {
      const unsigned long long i = base0 + e;
      q = a.hot_q[i]; c = a.hot_c[i]; est = a.hot_est[i];
      pass = qm_candidate_passes(a, q, c, est);
    }
    const unsigned bal = __ballot_sync(FULL, pass);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&a.counters[C_PF], (unsigned long long)__popc(bal));
    base = __shfl_sync(FULL, base, 0);
    if (pass) {
      const unsigned long long o = base + __popc(bal & ((1u << lane) - 1u));
      if (o < a.out_cap) { a.out_q[o] = q; a.out_c[o] = c; a.out_est[o] = est; }
    }
  }
}

}  // namespace apss
