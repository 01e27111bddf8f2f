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
//   per batch  k_qm_count / scan / k_qm_emit cut every (query term, segment) list into PIECES of <= QM_PIECE
//              postings: (pointer, length, query weight * 2^F), contiguous per query.
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

static constexpr int QM_PIECE = 128;      // postings per piece
static constexpr int QM_MAXSEG = 40;      // segments a handle may hold (LSM: ~log2(#batches) in practice)
static constexpr int QM_HOT = 2048;       // hot-list capacity per pass (slots); more => full scan of the table
static constexpr int QM_TBL = 26624;      // table slots (keys + values = 208 KB of shared memory)
static constexpr int QM_CAP = 14336;      // list entries scored per pass (load factor <= 0.54)

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

// thread per query term: pieces of its lists over all segments
__global__ void k_qm_count(int nnz, const int32_t* __restrict__ q_dim, const SegList sl, int32_t* __restrict__ cnt) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nnz) return;
  int c = 0;
  if (t < nnz) {
    const int d = q_dim[t];
    for (int s = 0; s < sl.n; ++s) { const int len = __ldg(sl.dir[s] + d + 1) - __ldg(sl.dir[s] + d); c += (len + QM_PIECE - 1) / QM_PIECE; }
  }
  cnt[t] = c;
}

__global__ void k_qm_emit(int nnz, const int32_t* __restrict__ q_dim, const float* __restrict__ q_w, float scale, const SegList sl,
                          const int32_t* __restrict__ off, QmItem* __restrict__ items, long long cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz) return;
  long long o = off[t];
  if (o == off[t + 1]) return;
  const int d = q_dim[t];
  const float wqs = q_w[t] * scale;
  for (int s = 0; s < sl.n; ++s) {
    const int a = __ldg(sl.dir[s] + d), b = __ldg(sl.dir[s] + d + 1);
    for (int p = a; p < b; p += QM_PIECE, ++o)
      if (o < cap) { QmItem it; it.post = (unsigned long long)(sl.post[s] + p); it.len = min(QM_PIECE, b - p); it.wqs = wqs; items[o] = it; }
  }
}

// ------------------------------------------------------------------ scoring

struct QmArgs {
  const int32_t* q_ptr;          // pruned batch CSR
  const int32_t* item_off;       // [batch_nnz + 1] piece offsets per query term
  const QmItem* items; long long item_cap;
  const float* q_nrm; const int64_t* q_key;
  const float* row_ub; const int64_t* c_key;
  int64_t n_rows;                // stored vectors visible to this batch
  int64_t q_local_base;          // shard-local id of query 0 when the batch was indexed in this call, else -1
  int32_t nq;
  float thr, band1;              // t (rounded down) and 1 + guard band of the fp32 estimate
  float scale, inv_scale;        // 2^F, 2^-F
  float cu_max;                  // upper bound of every row_ub[]
  int32_t cap;                   // list entries scored per pass (QM_CAP; tests lower it to reach the ranged passes)
  int32_t* out_q; int32_t* out_c; float* out_est; unsigned long long out_cap;
  unsigned long long* counters;
};

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

template <int NT, bool DUPKEYS>
__global__ void __launch_bounds__(NT, 1) k_score_qm(const QmArgs a) {
  extern __shared__ __align__(16) unsigned qm_smem[];
  unsigned* keys = qm_smem;
  unsigned* vals = qm_smem + QM_TBL;
  int* hot = reinterpret_cast<int*>(qm_smem + 2 * QM_TBL);
  __shared__ long long red[NT / 32];
  __shared__ int s_q;
  __shared__ unsigned s_hot_n;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid * 4; i < 2 * QM_TBL; i += NT * 4) *reinterpret_cast<uint4*>(qm_smem + i) = make_uint4(0, 0, 0, 0);
  unsigned long long n_post = 0; unsigned n_cand = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = (int)atomicAdd(&a.counters[C_WORK], 1ULL);
    __syncthreads();
    const int q = s_q;
    if (q >= a.nq) break;
    const int t0 = __ldg(a.q_ptr + q), t1 = __ldg(a.q_ptr + q + 1);
    if (t0 == t1) continue;
    const long long i0 = __ldg(a.item_off + t0), i1 = min((long long)__ldg(a.item_off + t1), a.item_cap);
    if (i0 >= i1) continue;
    long long mine = 0;
    for (long long i = i0 + tid; i < i1; i += NT) mine += __ldg(&a.items[i].len);
    const long long total = qm_block_sum<NT>(mine, red, tid);
    if (tid == 0) n_post += (unsigned long long)total;
    const float qn = __ldg(a.q_nrm + q);
    const unsigned self = a.q_local_base >= 0 ? (unsigned)(a.q_local_base + q) : 0xffffffffu;
    long long qkey = 0;
    if (DUPKEYS) qkey = __ldg(a.q_key + q);
    // smallest fixed-point sum a candidate of this query needs before the exact test can pass (0: test them all)
    unsigned thr_fix = 0;
    {
      const double em = ((double)a.thr - (double)a.cu_max * (double)qn * (1.0 + 1e-6)) / (double)a.band1;
      if (em > 0.0) thr_fix = (unsigned)fmin(floor(em * (double)a.scale * (1.0 - 1e-6)), 4294967295.0);
    }
    const bool scan_all = DUPKEYS || thr_fix == 0;

    // one pass: candidates with lo <= id < hi (ranged) or all of them
    auto walk = [&](const bool ranged, const bool count_only, const unsigned lo, const unsigned hi, const unsigned size) -> long long {
      long long cnt = 0;
      for (long long i = i0 + warp; i < i1; i += NW) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(a.items + i));          // same address on every lane
        const unsigned long long pp = ((unsigned long long)raw.y << 32) | raw.x;
        const int len = (int)raw.z; const float wqs = __uint_as_float(raw.w);
        if (ranged) {       // ids ascend inside a list: skip a piece that lies outside the range
          const uint2* p2 = reinterpret_cast<const uint2*>(pp);
          if (__ldg(&p2[len - 1].x) < lo || __ldg(&p2[0].x) >= hi) continue;
        }
        const int odd = (int)((pp >> 3) & 1ULL);                                       // list starts on the upper half of a 16-byte pair
        const uint4* p4 = reinterpret_cast<const uint4*>(pp - 8ULL * odd);
        for (int j0 = -odd; j0 < len; j0 += 64) {
          const int j = j0 + 2 * lane;
          if (j + 1 < 0 || j >= len) continue;
          const uint4 v = ld_stream4(p4 + ((j + odd) >> 1));
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const unsigned c = u ? v.z : v.x; const float w = __uint_as_float(u ? v.w : v.y);
            if (j + u < 0 || j + u >= len || c == self) continue;
            if (ranged && (c < lo || c >= hi)) continue;
            if (count_only) { ++cnt; continue; }
            const unsigned contrib = __float2uint_ru(__fmul_ru(w, wqs));
            const unsigned k = c + 1u;
            unsigned slot = __umulhi(c * 0x9E3779B1u, size);
            for (;;) {
              const unsigned old = atomicCAS(keys + slot, 0u, k);
              if (old == 0u) { ++n_cand; break; }
              if (old == k) break;
              if (++slot == size) slot = 0;
            }
            const unsigned prev = atomicAdd(vals + slot, contrib);
            if (!scan_all && prev < thr_fix && prev + contrib >= thr_fix) {
              const unsigned e = atomicAdd(&s_hot_n, 1u);
              if (e < (unsigned)QM_HOT) hot[e] = (int)slot;
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
      const float ub = __fmul_ru(__ldg(a.row_ub + c), qn);
      if (__fmaf_ru(est, a.band1, ub) >= a.thr) {
        const unsigned long long slot = atomicAdd(&a.counters[C_PF], 1ULL);
        if (slot < a.out_cap) { a.out_q[slot] = q; a.out_c[slot] = (int32_t)c; a.out_est[slot] = est; }
      }
    };
    auto pass = [&](const bool ranged, const unsigned lo, const unsigned hi, const long long entries) {
      unsigned size = (unsigned)min((long long)QM_TBL, max(256LL, (2 * entries + 31) & ~31LL));
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
        long long cnt = qm_block_sum<NT>(walk(true, true, (unsigned)lo, (unsigned)hi, 0u), red, tid);
        while (cnt > a.cap && hi - lo > a.cap) {
          hi = lo + (hi - lo) / 2;
          cnt = qm_block_sum<NT>(walk(true, true, (unsigned)lo, (unsigned)hi, 0u), red, tid);
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

}  // namespace apss
