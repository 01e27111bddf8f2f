// apss_api.cu -- host engine + C ABI (include/apss.h) of the B200 all-pairs similarity scorer.
// One handle = one index worker on one GPU (replaces IndexingWorkerActor, IWA:21-149).
#include "apss.h"
#include "apss_kernels.cuh"
#include "apss_qmajor.cuh"

#include <cub/cub.cuh>
#include <cuda.h>   // driver types only; entry points are resolved at run time (no libcuda link)

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace apss;

namespace {

// APSS_ALLOC_TRACE=1: every growth of a device buffer with its size and host time (growth stalls the caller)
struct AllocTrace {
  const char* what; size_t bytes; std::chrono::steady_clock::time_point t0; bool on;
  AllocTrace(const char* w, size_t b) : what(w), bytes(b), t0(std::chrono::steady_clock::now()), on(getenv("APSS_ALLOC_TRACE") != nullptr) {}
  ~AllocTrace() {
    if (on) std::fprintf(stderr, "apss alloc: %s to %.1f MB took %.2f ms\n", what, bytes / 1e6,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  // grow to >= n elements, preserving the first `keep` elements
  cudaError_t reserve(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return cudaSuccess;
    AllocTrace tr_("cudaMalloc", n * sizeof(T));
    // 25 % head-room on every (re)allocation: per-batch scratch is sized by the batch, batches differ by a few percent, and a
    // cudaMalloc + cudaFree pair costs tens of milliseconds here (large VMM reservations mapped) -- a dozen buffers growing in
    // the same call was a 0.5 s stall in the middle of a live stream
    size_t ncap = std::max(n + n / 4, cap + cap / 2);
    ncap = (ncap + 255) & ~size_t(255);
    T* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) { e = cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s); if (e != cudaSuccess) { cudaFree(np); return e; } }
    if (p) { cudaStreamSynchronize(s); cudaFree(p); }
    p = np; cap = ncap;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  size_t bytes() const { return cap * sizeof(T); }
};

// ---- growable device array backed by CUDA virtual memory management: one virtual range is reserved up
// front, physical chunks are mapped behind it as the index grows.  The pointer never moves, growth copies
// nothing and frees nothing (a cudaMalloc + copy + cudaFree of a few hundred MB stalls a live index for
// ~1 s), and the arrays can grow to the whole 180 GB of HBM without a 2x peak.
struct VmApi {
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

static VmApi& vm_api() {
  static VmApi api = [] {
    VmApi a;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult st;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn;
    };
    a.ok = get("cuMemAddressReserve", (void**)&a.AddressReserve) && get("cuMemAddressFree", (void**)&a.AddressFree) &&
           get("cuMemCreate", (void**)&a.Create) && get("cuMemRelease", (void**)&a.Release) && get("cuMemMap", (void**)&a.Map) &&
           get("cuMemUnmap", (void**)&a.Unmap) && get("cuMemSetAccess", (void**)&a.SetAccess) &&
           get("cuMemGetAllocationGranularity", (void**)&a.GetGranularity);
    if (getenv("APSS_NO_VMM")) a.ok = false;
    cudaGetLastError();
    return a;
  }();
  return api;
}

template <typename T>
struct VmBuf {
  T* p = nullptr;
  size_t cap = 0;                       // elements backed by physical memory
  int device = 0;
  size_t va_bytes = 0, mapped = 0, gran = 0;
  std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;
  bool vmm = false;
  DevBuf<T> fallback;                   // used when the VMM entry points are unavailable

  cudaError_t reserve(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return cudaSuccess;
    AllocTrace tr_("vmm-map", n * sizeof(T));
    VmApi& api = vm_api();
    if (!api.ok || (p && !vmm)) {
      cudaError_t e = fallback.reserve(n, keep, s);
      p = fallback.p; cap = fallback.cap;
      return e;
    }
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = device;
    if (!p) {
      if (api.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !gran) gran = 2u << 20;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      va_bytes = ((std::max<size_t>(total_b, (size_t)1 << 30) + gran - 1) / gran) * gran;    // never more than the whole HBM
      CUdeviceptr va = 0;
      if (api.AddressReserve(&va, va_bytes, 0, 0, 0) != CUDA_SUCCESS) {                      // fall back for good
        cudaError_t e = fallback.reserve(n, keep, s);
        p = fallback.p; cap = fallback.cap;
        return e;
      }
      p = reinterpret_cast<T*>(va); vmm = true;
    }
    size_t want = std::max(n * sizeof(T), mapped + mapped / 4);      // geometric, 25 %
    want = ((want + gran - 1) / gran) * gran;
    if (want > va_bytes) return cudaErrorMemoryAllocation;
    const size_t add = want - mapped;
    CUmemGenericAllocationHandle hnd;
    if (api.Create(&hnd, add, &prop, 0) != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
    if (api.Map((CUdeviceptr)p + mapped, add, 0, hnd, 0) != CUDA_SUCCESS) { api.Release(hnd); return cudaErrorMemoryAllocation; }
    CUmemAccessDesc acc{};
    acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (api.SetAccess((CUdeviceptr)p + mapped, add, &acc, 1) != CUDA_SUCCESS) { api.Unmap((CUdeviceptr)p + mapped, add); api.Release(hnd); return cudaErrorMemoryAllocation; }
    chunks.emplace_back(hnd, add);
    mapped = want; cap = mapped / sizeof(T);
    return cudaSuccess;
  }
  void release() {
    if (vmm) {
      VmApi& api = vm_api();
      size_t off = 0;
      for (auto& c : chunks) { api.Unmap((CUdeviceptr)p + off, c.second); api.Release(c.first); off += c.second; }
      chunks.clear();
      if (p) api.AddressFree((CUdeviceptr)p, va_bytes);
    } else fallback.release();
    p = nullptr; cap = 0; mapped = 0; vmm = false;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

}  // namespace

struct MultiState;

struct apss_handle {
  MultiState* multi = nullptr;   // n_devices > 1: this handle is the shard dispatcher over one engine per GPU (see the end of the file)
  apss_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int CR = 0, WARPS = 0, variant = 0, algo = 2, QB = 16;
  double max_sq = 0.0;   // largest squared norm of any pruned vector seen (stored or query)
  size_t smem_bytes = 0;
  size_t smem_optin = 0;   // device limit; the kernels' dynamic-smem attribute is set to (just under) it, not to this
                           // handle's need: the attribute is process-wide and handles with different tiles coexist
  std::string err;
  bool frozen = false, custom_keys = false;
  int64_t next_id = 0;
  int max_nnz_seen = 0;
  int64_t phase_cycles[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // last batch, profiling build only (see apss_stats)
  double host_us[6] = {0, 0, 0, 0, 0, 0};               // APSS_HOST_TIMING: host wall clock of the last batch by phase

  // shard
  int64_t n_local = 0, nnz = 0, n_post = 0;
  int64_t ntiles = 0;
  VmBuf<int64_t> fwd_ptr; VmBuf<int32_t> fwd_idx; VmBuf<double> fwd_val; VmBuf<int32_t> gid; VmBuf<int64_t> key;
  VmBuf<uint2> post; VmBuf<int32_t> dir; VmBuf<int64_t> tile_base;
  DevBuf<double> maxw;
  // batch staging
  DevBuf<int64_t> b_ptr; DevBuf<int32_t> b_idx; DevBuf<double> b_val; DevBuf<int64_t> b_key; DevBuf<int32_t> b_first;
  DevBuf<int32_t> q_cnt, q_ptr, q_dim; DevBuf<double> q_val; DevBuf<float> q_w; DevBuf<uint8_t> q_status;
  // build scratch
  DevBuf<unsigned long long> s_keys_in, s_keys_out, s_vals_in; DevBuf<int64_t> s_tile_start; DevBuf<char> cub_tmp;
  // dense-head tiles (algo 3)
  VmBuf<int32_t> dn_cnt, dn_dim, dn_len, tile_cnt; VmBuf<int2> dn_hash; VmBuf<float> dn_w; DevBuf<unsigned long long> s_vals_out;
  int dense_shift = 2, COLS = 4, ctas_per_sm = 1, seg_cap = SEG_CAP;
  // exact index reduction (cfg.pruning): document frequencies, per stored component "not indexed" flag, per
  // stored vector norm bound of its un-indexed part; per batch the same + rank-sort scratch
  bool prune = false; double prune_lim = 0.0, max_qnorm = 1.0;
  int prune_mode = 0;        // 1: tile kernels on the reduced index, 2: candidate-major kernel (no tiles are built)
  int cand_warps = 24;
  double cand_rate = -1.0;   // query-list entries met per stored vector per query, from the previous batch (-1: unknown)
  int cand_slices_env = 0;
  DevBuf<int32_t> qdir; VmBuf<int32_t> heavy;
  VmBuf<int64_t> ifw_ptr; VmBuf<uint2> ifw; DevBuf<int32_t> q_icnt, q_iptr;   // compact store of the indexed components
  DevBuf<int32_t> df; VmBuf<uint8_t> fwd_skip; VmBuf<float> row_ub;
  DevBuf<uint8_t> q_skip; DevBuf<float> q_cu, q_nrm, q_bkt; DevBuf<int32_t> q_dfmin; VmBuf<int32_t> row_dfmin;
  DevBuf<unsigned long long> pr_keys_in, pr_keys_out, pr_vals_in, pr_vals_out;
  int64_t tot_skipped = 0;
  // query-major scoring on the reduced index (prune_mode 3): LSM posting segments, oldest first
  // Segments are stacked by age in ONE growable arena (VMM: the pointer never moves, growth maps pages behind it and
  // copies nothing); a merge writes to the scratch arena and is copied back over its sources -- or, when it took
  // every segment, the two arenas simply swap.  Directories come from a fixed pool of (D + 1)-entry slots.
  struct Seg { int64_t off = 0; int32_t dir_slot = -1; int64_t n_post = -1; int64_t cap_post = 0; int64_t row_lo = 0, row_hi = 0; };
  std::vector<Seg> segs;
  VmBuf<uint2> seg_arena[2];
  DevBuf<int32_t> dir_pool; std::vector<int32_t> dir_free;
  uint2* seg_post(const Seg& g) const { return seg_arena[0].p + g.off; }
  int32_t* seg_dir(const Seg& g) const { return dir_pool.p + (size_t)g.dir_slot * ((size_t)cfg.dim + 1); }
  DevBuf<unsigned> sg_keys_in, sg_keys_out; DevBuf<unsigned long long> sg_vals_in;
  DevBuf<unsigned long long> qm_cnt, qm_off; DevBuf<QmItem> qm_items;
  int64_t merges = 0, merged_postings = 0;
  int64_t merge_ratio = 8;   // an older segment more than this many times the newer ones together is left alone (APSS_QM_MERGE_RATIO)
  int qm_cap = QM_CAP, qm_nt = 1024; size_t qm_items_cap0 = 0; bool qm_pipe = true; DevBuf<int32_t> qm_deferred, hot_q, hot_c; DevBuf<float> hot_est; DevBuf<unsigned> hot_used;     // test hooks: APSS_QM_CAP, APSS_QM_ITEMS_CAP
  bool broken = false;       // a failure after the index was touched that could not be rolled back: every later call fails
  // query-block transposition (v2 kernel)
  DevBuf<unsigned long long> bt_keys_in, bt_keys_out, bt_vals_in, bt_vals_out, ud_key; DevBuf<int32_t> bt_flags, bt_pos, ud_dim, ud_start, bd_ptr;
  // outputs
  DevBuf<int32_t> pf_q, pf_c; DevBuf<float> pf_est;
  DevBuf<int32_t> out_q, out_c; DevBuf<double> out_sim;
  unsigned long long* d_counters = nullptr;
  unsigned long long* h_counters = nullptr;   // pinned
  int32_t* h_total = nullptr;                 // pinned
  cudaEvent_t ev_b0 = nullptr, ev_b1 = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;

  // last batch
  int32_t last_n = -1; int64_t last_pairs = 0;
  std::vector<uint8_t> last_status;
  // totals
  int64_t tot_postings = 0, tot_cands = 0, tot_pairs = 0, tot_pf = 0, score_launches = 0, kernel_launches = 0;
  double tot_score_ms = 0;

  int32_t fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return h->fail(e_ == cudaErrorMemoryAllocation ? APSS_E_NOMEM : APSS_E_CUDA, \
                                          "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static constexpr int U16_MAX_NNZ = 60000;   // longest vector the packed-u16 tile kernel accepts (see transpose_query_blocks)
static inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

template <int WARPS, int UNROLL>
static cudaError_t launch_score_t(apss_handle* h, const ScoreArgs& a, bool dup) {
  auto kern = dup ? k_score<WARPS, UNROLL, true> : k_score<WARPS, UNROLL, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(h->smem_bytes, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  kern<<<h->sm_count, WARPS * 32, h->smem_bytes, h->stream>>>(a);
  return cudaGetLastError();
}

static cudaError_t launch_score(apss_handle* h, const ScoreArgs& a, bool dup) {
  const int unroll = (h->variant & 0xff) ? (h->variant & 0xff) : 8;
  switch (h->WARPS) {
    case 8: return unroll == 8 ? launch_score_t<8, 8>(h, a, dup) : launch_score_t<8, 4>(h, a, dup);
    case 16: return unroll == 8 ? launch_score_t<16, 8>(h, a, dup) : (unroll == 2 ? launch_score_t<16, 2>(h, a, dup) : launch_score_t<16, 4>(h, a, dup));
    case 32: return unroll == 8 ? launch_score_t<32, 8>(h, a, dup) : launch_score_t<32, 4>(h, a, dup);
    default: return cudaErrorInvalidValue;
  }
}

template <int WARPS>
static cudaError_t launch_blk_t(apss_handle* h, const ScoreArgs& a, const BlockArgs& b, bool dup) {
  auto kern = dup ? k_score_blk<WARPS, true> : k_score_blk<WARPS, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(h->smem_bytes, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  kern<<<h->sm_count, WARPS * 32, h->smem_bytes, h->stream>>>(a, b);
  return cudaGetLastError();
}

static cudaError_t launch_blk(apss_handle* h, const ScoreArgs& a, const BlockArgs& b, bool dup) {
  switch (h->WARPS) {
    case 8: return launch_blk_t<8>(h, a, b, dup);
    case 16: return launch_blk_t<16>(h, a, b, dup);
    case 32: return launch_blk_t<32>(h, a, b, dup);
    default: return cudaErrorInvalidValue;
  }
}

template <int QB, int WARPS, int COLS>
static cudaError_t launch_dense_t(apss_handle* h, const ScoreArgs& a, const BlockArgs& b, const DenseTiles& d, bool dup) {
  auto kern = dup ? k_score_dense<QB, WARPS, COLS, true> : k_score_dense<QB, WARPS, COLS, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(h->smem_bytes, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  kern<<<h->sm_count * h->ctas_per_sm, WARPS * 32, h->smem_bytes, h->stream>>>(a, b, d);
  return cudaGetLastError();
}

static cudaError_t launch_dense_pruned(apss_handle* h, const ScoreArgs& a, const BlockArgs& b, const DenseTiles& d, bool dup) {
  auto kern = dup ? k_score_dense<16, 16, 4, true, true> : k_score_dense<16, 16, 4, false, true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(h->smem_bytes, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  kern<<<h->sm_count * h->ctas_per_sm, 16 * 32, h->smem_bytes, h->stream>>>(a, b, d);
  return cudaGetLastError();
}

static cudaError_t launch_dense(apss_handle* h, const ScoreArgs& a, const BlockArgs& b, const DenseTiles& d, bool dup) {
  const int key = h->QB * 10000 + h->WARPS * 100 + h->COLS;
  if (h->prune_mode == 1) return key == 161604 ? launch_dense_pruned(h, a, b, d, dup) : cudaErrorInvalidValue;
  switch (key) {
    case 321602: return launch_dense_t<32, 16, 2>(h, a, b, d, dup);
    case 321202: return launch_dense_t<32, 12, 2>(h, a, b, d, dup);
    case 320802: return launch_dense_t<32, 8, 2>(h, a, b, d, dup);
    case 161602: return launch_dense_t<16, 16, 2>(h, a, b, d, dup);
    case 161604: return launch_dense_t<16, 16, 4>(h, a, b, d, dup);
    case 163202: return launch_dense_t<16, 32, 2>(h, a, b, d, dup);
    case 160804: return launch_dense_t<16, 8, 4>(h, a, b, d, dup);
    case 81604: return launch_dense_t<8, 16, 4>(h, a, b, d, dup);
    default: return cudaErrorInvalidValue;
  }
}

// Candidate-major scoring on the reduced index: invert the batch (dim -> (query, weight) lists), then stream the
// stored vectors (k_score_cand) and finish the deferred ones (k_score_cand_heavy).
static int32_t build_query_index(apss_handle* h, int32_t n, int32_t batch_nnz, int slices, int qsub) {
  cudaStream_t s = h->stream;
  const int D = h->cfg.dim;
  int dimbits = 1; while ((1LL << dimbits) < (int64_t)D + 1) ++dimbits;
  int sbits = 0; while ((1 << sbits) < slices) ++sbits;
  CK(h->bt_keys_in.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->bt_keys_out.reserve(std::max(batch_nnz, 1), 0, s));
  CK(h->bt_vals_in.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->bt_vals_out.reserve(std::max(batch_nnz, 1), 0, s));
  CK(h->qdir.reserve(((size_t)D + 1) * slices, 0, s));
  if (batch_nnz) {
    k_qi_emit<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, h->q_ptr.p, h->q_dim.p, h->q_w.p, qsub, dimbits, h->bt_keys_in.p, h->bt_vals_in.p);
    CK(cudaGetLastError());
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, h->bt_keys_in.p, h->bt_keys_out.p, h->bt_vals_in.p, h->bt_vals_out.p, batch_nnz, 0, dimbits + sbits, s));
    CK(h->cub_tmp.reserve(tb, 0, s));
    CK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tb, h->bt_keys_in.p, h->bt_keys_out.p, h->bt_vals_in.p, h->bt_vals_out.p, batch_nnz, 0, dimbits + sbits, s));
    h->kernel_launches += 3;
  }
  k_qdir<<<cdiv(((int64_t)D + 1) * slices, 256), 256, 0, s>>>(h->bt_keys_out.p, batch_nnz, D, dimbits, slices, h->qdir.p);
  CK(cudaGetLastError()); h->kernel_launches++;
  return APSS_OK;
}

static cudaError_t launch_cand(apss_handle* h, const CandArgs& a) {
  // 24 warps with 1024-slot tables (default) or 32 warps with 512-slot tables
  const int tbl = h->cand_warps == 32 ? 512 : 1024;
  const size_t smem = (size_t)h->cand_warps * (2 * tbl + 3 * CAND_FEAT + 4) * sizeof(unsigned);
  auto kern = h->cand_warps == 16 ? k_score_cand<16, 1024> : h->cand_warps == 32 ? k_score_cand<32, 512> : k_score_cand<24, 1024>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  kern<<<h->sm_count, h->cand_warps * 32, smem, h->stream>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int qc = 32768;
  e = cudaFuncSetAttribute(k_score_cand_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max((size_t)qc * 4, h->smem_optin - 1024));
  if (e != cudaSuccess) return e;
  k_score_cand_heavy<<<h->sm_count, 512, (size_t)qc * 4, h->stream>>>(a, qc);
  return cudaGetLastError();
}

extern "C" int32_t apss_abi_version(void) { return APSS_ABI_VERSION; }

static int32_t multi_create(const apss_config* cfg, apss_handle** out);
static void multi_destroy(apss_handle* h);
static int32_t multi_insert_batch(apss_handle* h, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                  const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags, apss_batch_result* out);
static int32_t multi_fetch_pairs(apss_handle* h, int32_t* q, int32_t* c, double* sim, int64_t capacity, int64_t* n_out);
static int32_t multi_get_stats(apss_handle* h, apss_stats* out);
static int32_t multi_fetch_status(apss_handle* h, uint8_t* status, int32_t capacity);
static int32_t multi_freeze(apss_handle* h);
static int32_t multi_set_next_id(apss_handle* h, int64_t next_id);

extern "C" int32_t apss_create(const apss_config* cfg, apss_handle** out) {
  if (!cfg || !out || cfg->struct_size != (int32_t)sizeof(apss_config)) return APSS_E_INVALID;
  *out = nullptr;
  if (cfg->n_devices < 0 || cfg->n_devices > APSS_MAX_DEVICES) return APSS_E_INVALID;
  if (cfg->n_devices > 1) return multi_create(cfg, out);
  if (cfg->dim <= 0 || cfg->dim > (1 << 30) || !(cfg->index_threshold >= 0.0) || !(cfg->similarity_threshold == cfg->similarity_threshold) || (cfg->semantics != APSS_SEM_R1 && cfg->semantics != APSS_SEM_R0)) return APSS_E_INVALID;
  int ndev = 0;
  const int dev0 = cfg->n_devices == 1 ? cfg->device_ids[0] : cfg->device;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || dev0 < 0 || dev0 >= ndev) { cudaGetLastError(); return APSS_E_NO_DEVICE; }
  apss_handle* h = new apss_handle();
  h->cfg = *cfg; h->cfg.max_weight = nullptr; h->device = cfg->n_devices == 1 ? cfg->device_ids[0] : cfg->device;
  h->cfg.device = h->device;
  h->row_dfmin.device = cfg->n_devices == 1 ? cfg->device_ids[0] : cfg->device;
  h->fwd_skip.device = h->row_ub.device = h->heavy.device = h->ifw_ptr.device = h->ifw.device = cfg->device;
  h->fwd_ptr.device = h->fwd_idx.device = h->fwd_val.device = h->gid.device = h->key.device = h->post.device = h->dir.device =
      h->tile_base.device = h->dn_cnt.device = h->dn_dim.device = h->dn_len.device = h->tile_cnt.device = h->dn_hash.device =
      h->dn_w.device = cfg->device;
  auto bail = [&](int32_t rc) { apss_destroy(h); return rc; };
  if (cudaSetDevice(h->device) != cudaSuccess) return bail(APSS_E_NO_DEVICE);
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return bail(APSS_E_NO_DEVICE);
  h->sm_count = prop.multiProcessorCount;
  const size_t max_smem1 = prop.sharedMemPerBlockOptin;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  // kernel_variant (see include/apss.h): bits 0-7 unroll of the row kernel / candidates per thread of the dense
  // phase, 8-15 warps per CTA, 16-23 scoring kernel (0/3 = dense-head, 1 = row, 2 = query-block), 24-31 queries
  // per block.  The index layout (tile size) follows from the kernel's shared-memory budget.
  int algo = (cfg->kernel_variant >> 16) & 0xff;   // 0 = default
  if (algo != 1 && algo != 2 && algo != 3) algo = 3;   // default: dense-head kernel
  int QB = (cfg->kernel_variant >> 24) & 0xff; if (QB <= 0) QB = 16;
  const bool qb_given = ((cfg->kernel_variant >> 24) & 0xff) != 0;
  if (QB > 32) QB = 32;   // a dimension's row list is staged one entry per lane
  if (algo == 3 && QB != 8 && QB != 16 && QB != 32) QB = qb_given ? 32 : 16;
  size_t blk_extra = (size_t)LONG_CAP * 16 + (size_t)(LONG_CAP + 1) * 4 + 64;
  const int want_w = (cfg->kernel_variant >> 8) & 0xff;
  const int ctas = (algo == 3 && want_w == 8) ? 2 : 1;          // 8-warp CTAs run two per SM
  const int seg_cap = ctas == 2 ? 640 : SEG_CAP;
  if (algo == 3) blk_extra = (size_t)seg_cap * 16 + 64 + (size_t)KD * QB * 4 + (size_t)KD * 16 + (size_t)HS * 8;
  const size_t max_smem = ctas == 2 ? (prop.sharedMemPerMultiprocessor / 2 - 1024) : max_smem1;
  const size_t acc_bytes = algo == 3 ? 2 : 4;   // u16 packed (dense-head kernel) or u32 / fp32
  int CR = cfg->tile_vectors;
  if (CR <= 0) {
    CR = algo == 1 ? 3584 : (int)(((max_smem - blk_extra) / ((size_t)QB * acc_bytes)) / 128 * 128);
    if (algo == 3 && CR >= 1024) CR = CR / 1024 * 1024;   // whole passes of the dense phase
  }
  CR = (CR + 127) / 128 * 128;
  int warps = 0;
  const int want = (cfg->kernel_variant >> 8) & 0xff;
  if (algo == 1) {
    // one accumulator row per warp: as many rows as fit in shared memory, from {32, 16, 8}
    for (int w : {32, 16, 8}) if ((size_t)w * CR * sizeof(float) + 1024 <= max_smem) { warps = w; break; }
    if ((want == 8 || want == 16 || want == 32) && (size_t)want * CR * sizeof(float) + 1024 <= max_smem) warps = want;
    h->smem_bytes = (size_t)warps * CR * sizeof(float);
  } else {
    while (QB > (algo == 3 ? 8 : 1) && (size_t)QB * CR * acc_bytes + blk_extra > max_smem) {   // explicit tile size: shrink the block
      if (algo == 3) blk_extra -= (size_t)KD * (QB / 2) * 4;
      QB >>= 1;
    }
    if ((size_t)QB * CR * acc_bytes + blk_extra > max_smem) return bail(APSS_E_INVALID);
    warps = (want == 8 || want == 12 || want == 16 || want == 32) ? want : 16;
    h->smem_bytes = (size_t)QB * CR * acc_bytes + blk_extra;
    if (algo == 3) {
      h->COLS = (cfg->kernel_variant & 0xff) == 2 ? 2 : 4;
      if (QB == 32) { h->COLS = 2; if (warps != 8 && warps != 12 && warps != 16) warps = 16; }
      else if (QB == 16) { if (h->COLS == 2) { if (warps != 32) warps = 16; } else if (warps != 8) warps = 16; }
      else { h->COLS = 4; warps = 16; }
      const char* ds = getenv("APSS_DENSE_SHIFT");
      if (ds && atoi(ds) >= 0 && atoi(ds) <= 32) h->dense_shift = atoi(ds);   // experiments: see k_dense_select
    }
  }
  if (!warps) return bail(APSS_E_INVALID);
  h->CR = CR; h->WARPS = warps; h->variant = cfg->kernel_variant; h->algo = algo; h->QB = QB;
  h->ctas_per_sm = ctas; h->seg_cap = seg_cap;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(APSS_E_CUDA);
  if (cudaMalloc(&h->d_counters, C_COUNT * sizeof(unsigned long long)) != cudaSuccess) return bail(APSS_E_NOMEM);
  if (cudaMallocHost(&h->h_counters, (C_COUNT + 2) * sizeof(unsigned long long)) != cudaSuccess) return bail(APSS_E_NOMEM);
  if (cudaMallocHost(&h->h_total, sizeof(int32_t) * 4) != cudaSuccess) return bail(APSS_E_NOMEM);
  cudaEventCreate(&h->ev_b0); cudaEventCreate(&h->ev_b1); cudaEventCreate(&h->ev_s0); cudaEventCreate(&h->ev_s1);
  if (cfg->max_weight) {
    if (h->maxw.reserve(cfg->dim, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (cudaMemcpyAsync(h->maxw.p, cfg->max_weight, sizeof(double) * cfg->dim, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return bail(APSS_E_CUDA);
  }
  if (cfg->pruning) {
    if (cfg->pruning < 1 || cfg->pruning > 3 || algo != 3) return bail(APSS_E_INVALID);   // only the default scoring kernel applies the bound
    h->prune_mode = cfg->pruning;
    if (cfg->pruning == 3) {
      h->seg_arena[0].device = h->seg_arena[1].device = cfg->device;
      if (h->dir_pool.reserve((size_t)(QM_MAXSEG + 2) * ((size_t)cfg->dim + 1), 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
      for (int k = QM_MAXSEG + 1; k >= 0; --k) h->dir_free.push_back(k);
      { const char* e = getenv("APSS_QM_CAP"); if (e && atoi(e) >= 8 && atoi(e) <= QM_CAP) h->qm_cap = atoi(e); }
      { const char* e = getenv("APSS_QM_MERGE_RATIO"); if (e && atoi(e) >= 1 && atoi(e) <= 1024) h->merge_ratio = atoi(e); }
      { const char* e = getenv("APSS_QM_NT"); if (e && atoi(e) == 512) h->qm_nt = 512; }
      { const char* e = getenv("APSS_QM_PIPE"); if (e && atoi(e) == 0) h->qm_pipe = false; }      // measurement / tests: ranged kernel only
      { const char* e = getenv("APSS_QM_ITEMS_CAP"); if (e && atoll(e) >= 1) h->qm_items_cap0 = (size_t)atoll(e); }
    }
    if (cfg->pruning == 1 && (QB != 16 || warps != 16 || h->COLS != 4)) return bail(APSS_E_INVALID);   // tile kernel: default shape only
    { const char* cw = getenv("APSS_CAND_WARPS"); if (cw && (atoi(cw) == 16 || atoi(cw) == 24 || atoi(cw) == 32)) h->cand_warps = atoi(cw); }
    { const char* cs = getenv("APSS_CAND_SLICES"); if (cs && atoi(cs) >= 1 && atoi(cs) <= 256) h->cand_slices_env = atoi(cs); }
    const double alpha = cfg->prune_alpha == 0.0 ? 0.8 : cfg->prune_alpha;
    const double qn = cfg->max_query_norm == 0.0 ? 1.0 : cfg->max_query_norm;
    if (!(alpha > 0.0 && alpha < 1.0) || !(qn > 0.0) || !std::isfinite(qn)) return bail(APSS_E_INVALID);
    const double t = cfg->similarity_threshold;
    h->prune = true; h->max_qnorm = qn;
    h->prune_lim = t > 0.0 ? alpha * t * t / (qn * qn) * (1.0 - std::ldexp(1.0, -20)) : 0.0;
    if (h->df.reserve(cfg->dim, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (cudaMemsetAsync(h->df.p, 0, sizeof(int32_t) * (size_t)cfg->dim, h->stream) != cudaSuccess) return bail(APSS_E_CUDA);
  }
  if (cfg->reserve_vectors > 0) {
    int64_t nv = cfg->reserve_vectors, nt = (nv + CR - 1) / CR;
    if (h->fwd_ptr.reserve(nv + 1, 0, h->stream) != cudaSuccess || h->gid.reserve(nv, 0, h->stream) != cudaSuccess ||
        h->key.reserve(nv, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (h->prune_mode < 2 && (h->dir.reserve((size_t)nt * ((size_t)cfg->dim + 1), 0, h->stream) != cudaSuccess ||      // (no tiles in modes 2 / 3)
                              h->tile_base.reserve(nt + 1, 0, h->stream) != cudaSuccess)) return bail(APSS_E_NOMEM);
    if (algo == 3 && h->prune_mode < 2 && (h->dn_cnt.reserve(nt, 0, h->stream) != cudaSuccess || h->tile_cnt.reserve(nt, 0, h->stream) != cudaSuccess ||
                      h->dn_dim.reserve((size_t)nt * KD, 0, h->stream) != cudaSuccess || h->dn_len.reserve((size_t)nt * KD, 0, h->stream) != cudaSuccess ||
                      h->dn_hash.reserve((size_t)nt * HS, 0, h->stream) != cudaSuccess || h->dn_w.reserve((size_t)nt * KD * CR, 0, h->stream) != cudaSuccess))
      return bail(APSS_E_NOMEM);
  }
  if (cfg->reserve_nnz > 0) {
    if (h->fwd_idx.reserve(cfg->reserve_nnz, 0, h->stream) != cudaSuccess || h->fwd_val.reserve(cfg->reserve_nnz, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (h->prune_mode < 2 && h->post.reserve(cfg->reserve_nnz, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (h->prune && h->fwd_skip.reserve(cfg->reserve_nnz, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
    if (h->prune_mode == 3 && (h->seg_arena[0].reserve((size_t)cfg->reserve_nnz / 2, 0, h->stream) != cudaSuccess ||
                               h->seg_arena[1].reserve((size_t)cfg->reserve_nnz / 2, 0, h->stream) != cudaSuccess)) return bail(APSS_E_NOMEM);
  }
  if (cfg->reserve_vectors > 0 && h->prune && h->row_ub.reserve(cfg->reserve_vectors, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
  if (cfg->reserve_vectors > 0 && h->prune_mode == 3 && h->row_dfmin.reserve(cfg->reserve_vectors, 0, h->stream) != cudaSuccess) return bail(APSS_E_NOMEM);
  {
    size_t np = cfg->reserve_pairs > 0 ? (size_t)cfg->reserve_pairs : (size_t)1 << 20;
    if (h->pf_q.reserve(np, 0, h->stream) != cudaSuccess || h->pf_c.reserve(np, 0, h->stream) != cudaSuccess || h->pf_est.reserve(np, 0, h->stream) != cudaSuccess ||
        h->out_q.reserve(np, 0, h->stream) != cudaSuccess || h->out_c.reserve(np, 0, h->stream) != cudaSuccess || h->out_sim.reserve(np, 0, h->stream) != cudaSuccess)
      return bail(APSS_E_NOMEM);
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(APSS_E_CUDA);
  *out = h;
  return APSS_OK;
}

extern "C" void apss_destroy(apss_handle* h) {
  if (!h) return;
  if (h->multi) { multi_destroy(h); return; }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->fwd_ptr.release(); h->fwd_idx.release(); h->fwd_val.release(); h->gid.release(); h->key.release();
  h->post.release(); h->dir.release(); h->tile_base.release(); h->maxw.release();
  h->b_ptr.release(); h->b_idx.release(); h->b_val.release(); h->b_key.release(); h->b_first.release();
  h->q_cnt.release(); h->q_ptr.release(); h->q_dim.release(); h->q_val.release(); h->q_w.release(); h->q_status.release();
  h->bt_keys_in.release(); h->bt_keys_out.release(); h->bt_vals_in.release(); h->bt_vals_out.release(); h->ud_key.release();
  h->bt_flags.release(); h->bt_pos.release(); h->ud_dim.release(); h->ud_start.release(); h->bd_ptr.release();
  h->dn_cnt.release(); h->dn_dim.release(); h->dn_len.release(); h->tile_cnt.release(); h->dn_hash.release(); h->dn_w.release(); h->s_vals_out.release();
  h->s_keys_in.release(); h->s_keys_out.release(); h->s_vals_in.release(); h->s_tile_start.release(); h->cub_tmp.release();
  h->pf_q.release(); h->pf_c.release(); h->pf_est.release(); h->out_q.release(); h->out_c.release(); h->out_sim.release();
  h->qdir.release(); h->heavy.release(); h->ifw_ptr.release(); h->ifw.release(); h->q_icnt.release(); h->q_iptr.release();
  h->df.release(); h->fwd_skip.release(); h->row_ub.release(); h->q_skip.release(); h->q_cu.release(); h->q_nrm.release();
  h->q_bkt.release(); h->q_dfmin.release(); h->row_dfmin.release();
  h->pr_keys_in.release(); h->pr_keys_out.release(); h->pr_vals_in.release(); h->pr_vals_out.release();
  h->segs.clear(); h->seg_arena[0].release(); h->seg_arena[1].release(); h->dir_pool.release();
  h->sg_keys_in.release(); h->sg_keys_out.release(); h->sg_vals_in.release(); h->qm_cnt.release(); h->qm_off.release(); h->qm_items.release(); h->qm_deferred.release(); h->hot_q.release(); h->hot_c.release(); h->hot_est.release(); h->hot_used.release();
  if (h->d_counters) cudaFree(h->d_counters);
  if (h->h_counters) cudaFreeHost(h->h_counters);
  if (h->h_total) cudaFreeHost(h->h_total);
  if (h->ev_b0) cudaEventDestroy(h->ev_b0);
  if (h->ev_b1) cudaEventDestroy(h->ev_b1);
  if (h->ev_s0) cudaEventDestroy(h->ev_s0);
  if (h->ev_s1) cudaEventDestroy(h->ev_s1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// IWA:61-71 on the GPU: append the batch's pruned vectors to the forward store and (re)build the
// index tiles they fall into: at most the last, partially filled tile plus the new ones.
// Exact index reduction: choose, for every vector of the batch, the components that stay out of the index
// (see k_prune_mark).  Fills q_skip[batch_nnz] and q_cu[n]; the document frequencies include this batch.
static int32_t prune_select(apss_handle* h, int32_t n, int32_t batch_nnz) {
  cudaStream_t s = h->stream;
  CK(h->q_skip.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->q_cu.reserve(n, 0, s)); CK(h->q_dfmin.reserve(n, 0, s));
  CK(h->q_icnt.reserve(n + 1, 0, s)); CK(h->q_iptr.reserve(n + 1, 0, s));
  // ranking inside a warp when every vector of the batch is short enough (C_MAXNNZ: read back after the prefilter)
  const bool local_rank = (int64_t)h->h_counters[C_MAXNNZ] <= PRL_MAX && !getenv("APSS_PRUNE_GLOBAL_SORT");
  if (batch_nnz) {
    k_df_update<<<cdiv(batch_nnz, 256), 256, 0, s>>>(batch_nnz, h->q_dim.p, h->df.p, 1);
    CK(cudaGetLastError()); h->kernel_launches++;
  }
  if (local_rank) {
    k_prune_rank_mark<<<cdiv((int64_t)n + 1, PRL_WARPS), PRL_WARPS * 32, 0, s>>>(n, h->q_ptr.p, h->q_dim.p, h->q_val.p, h->df.p, h->prune_lim, h->q_skip.p, h->q_cu.p,
                                                                               h->q_icnt.p, h->q_dfmin.p, h->d_counters);
    CK(cudaGetLastError()); h->kernel_launches++;
  } else {
    CK(h->pr_keys_in.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->pr_keys_out.reserve(std::max(batch_nnz, 1), 0, s));
    CK(h->pr_vals_in.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->pr_vals_out.reserve(std::max(batch_nnz, 1), 0, s));
    if (batch_nnz) {
      k_rank_keys<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, h->q_ptr.p, h->q_dim.p, h->df.p, h->pr_keys_in.p, h->pr_vals_in.p);
      CK(cudaGetLastError());
      int rowbits = 1; while ((1LL << rowbits) < n) ++rowbits;
      size_t tb = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, h->pr_keys_in.p, h->pr_keys_out.p, h->pr_vals_in.p, h->pr_vals_out.p, batch_nnz, 0, 31 + rowbits, s));
      CK(h->cub_tmp.reserve(tb, 0, s));
      CK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tb, h->pr_keys_in.p, h->pr_keys_out.p, h->pr_vals_in.p, h->pr_vals_out.p, batch_nnz, 0, 31 + rowbits, s));
      h->kernel_launches += 3;
    }
    k_prune_mark<<<cdiv(((int64_t)n + 1) * 32, 256), 256, 0, s>>>(n, h->q_ptr.p, h->q_val.p, h->pr_vals_out.p, h->pr_keys_out.p, h->prune_lim, h->q_skip.p, h->q_cu.p, h->q_icnt.p,
                                                                  h->q_dfmin.p, h->d_counters);
    CK(cudaGetLastError()); h->kernel_launches++;
  }
  if (h->prune_mode == 2) {
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, h->q_icnt.p, h->q_iptr.p, n + 1, s));
    CK(h->cub_tmp.reserve(tb, 0, s));
    CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, h->q_icnt.p, h->q_iptr.p, n + 1, s));
    h->kernel_launches++;
  }
  return APSS_OK;
}

static int32_t index_append(apss_handle* h, int32_t n, int32_t batch_nnz, const int64_t* d_ext_keys) {
  cudaStream_t s = h->stream;
  const int CR = h->CR; const int D = h->cfg.dim;
  const int64_t n_old = h->n_local, n_new = n_old + n, nnz_old = h->nnz, nnz_new = nnz_old + batch_nnz;
  CK(h->fwd_ptr.reserve(n_new + 1, n_old ? n_old + 1 : 0, s));
  CK(h->gid.reserve(n_new, n_old, s));
  CK(h->key.reserve(n_new, n_old, s));
  CK(h->fwd_idx.reserve(std::max<int64_t>(nnz_new, 1), nnz_old, s));
  CK(h->fwd_val.reserve(std::max<int64_t>(nnz_new, 1), nnz_old, s));
  k_append_rows<<<cdiv(n, 256), 256, 0, s>>>(n, h->q_ptr.p, nnz_old, n_old, h->next_id, d_ext_keys, h->fwd_ptr.p, h->gid.p, h->key.p);
  CK(cudaGetLastError()); h->kernel_launches++;
  if (batch_nnz) {
    CK(cudaMemcpyAsync(h->fwd_idx.p + nnz_old, h->q_dim.p, sizeof(int32_t) * batch_nnz, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(h->fwd_val.p + nnz_old, h->q_val.p, sizeof(double) * batch_nnz, cudaMemcpyDeviceToDevice, s));
  }
  if (h->prune) {
    CK(h->fwd_skip.reserve(std::max<int64_t>(nnz_new, 1), nnz_old, s)); CK(h->row_ub.reserve(n_new, n_old, s));
    if (batch_nnz) CK(cudaMemcpyAsync(h->fwd_skip.p + nnz_old, h->q_skip.p, (size_t)batch_nnz, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(h->row_ub.p + n_old, h->q_cu.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    if (h->prune_mode == 3) {
      CK(h->row_dfmin.reserve(n_new, n_old, s));
      CK(cudaMemcpyAsync(h->row_dfmin.p + n_old, h->q_dfmin.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    }
  }
  if (h->prune_mode == 3) {     // query-major scoring: the batch becomes one new posting segment (merged after the call)
    if (batch_nnz) {
      if ((int)h->segs.size() >= QM_MAXSEG || h->dir_free.empty()) return h->fail(APSS_E_STATE, "too many posting segments");
      apss_handle::Seg sg; sg.row_lo = n_old; sg.row_hi = n_new;
      sg.cap_post = (int64_t)batch_nnz + 64;      // the sort writes the un-indexed components behind the postings; + read slack
      if (!h->segs.empty()) { const apss_handle::Seg& b = h->segs.back(); sg.off = (b.off + b.n_post + 64 + 31) & ~(int64_t)31; }
      CK(h->seg_arena[0].reserve((size_t)(sg.off + sg.cap_post), 0, s));
      CK(h->sg_keys_in.reserve(batch_nnz, 0, s)); CK(h->sg_keys_out.reserve(batch_nnz, 0, s)); CK(h->sg_vals_in.reserve(batch_nnz, 0, s));
      sg.dir_slot = h->dir_free.back(); h->dir_free.pop_back();
      h->segs.push_back(sg);       // from here on a failure is undone by the roll-back
      int dimbits = 1; while ((1LL << dimbits) < (int64_t)D + 1) ++dimbits;
      k_seg_emit<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, n_old, h->q_ptr.p, h->q_dim.p, h->q_w.p, h->q_skip.p, D, h->sg_keys_in.p, h->sg_vals_in.p);
      CK(cudaGetLastError());
      size_t tb = 0;
      unsigned long long* vout = reinterpret_cast<unsigned long long*>(h->seg_post(sg));
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, h->sg_keys_in.p, h->sg_keys_out.p, h->sg_vals_in.p, vout, batch_nnz, 0, dimbits, s));
      CK(h->cub_tmp.reserve(tb, 0, s));
      CK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tb, h->sg_keys_in.p, h->sg_keys_out.p, h->sg_vals_in.p, vout, batch_nnz, 0, dimbits, s));
      k_seg_dir<<<cdiv((int64_t)D + 1, 256), 256, 0, s>>>(h->sg_keys_out.p, batch_nnz, D, h->seg_dir(sg));
      CK(cudaGetLastError()); h->kernel_launches += 5;
    }
    h->n_local = n_new; h->nnz = nnz_new; h->ntiles = 0;
    return APSS_OK;
  }
  if (h->prune_mode == 2) {     // candidate-major scoring streams the forward store: there is no tile index to maintain
    CK(h->heavy.reserve(n_new, 0, s));
    // indexed components so far: known on the host up to the previous batch; this batch adds at most batch_nnz
    const int64_t ifw_old = nnz_old - h->tot_skipped;
    CK(h->ifw_ptr.reserve(n_new + 1, n_old ? n_old + 1 : 0, s)); CK(h->ifw.reserve(std::max<int64_t>(ifw_old + batch_nnz, 1), ifw_old, s));
    k_ifw_append<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, n_old, h->q_ptr.p, h->q_dim.p, h->q_w.p, h->q_skip.p, h->q_iptr.p, h->ifw_ptr.p, h->ifw.p);
    CK(cudaGetLastError()); h->kernel_launches++;
    h->n_local = n_new; h->nnz = nnz_new; h->ntiles = 0;
    return APSS_OK;
  }
  // tiles to (re)build: [tile0, tile1)
  const int64_t tile0 = n_old / CR, tile1 = (n_new + CR - 1) / CR;
  const int ntiles_aff = (int)(tile1 - tile0);
  const int64_t row_lo = tile0 * CR;
  // first stored component of row_lo: the open tile's postings are rewritten in place
  int64_t nnz_lo = nnz_old;
  if (row_lo < n_old) {
    CK(cudaMemcpyAsync(h->h_counters + C_SCRATCH, h->fwd_ptr.p + row_lo, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    nnz_lo = (int64_t)h->h_counters[C_SCRATCH];
  }
  const int64_t m = nnz_new - nnz_lo;
  const int64_t post_base = nnz_lo;   // postings are stored in the same order of tiles as the forward store
  CK(h->post.reserve(std::max<int64_t>(nnz_new, 1), h->algo == 3 ? nnz_old : post_base, s));
  CK(h->dir.reserve((size_t)tile1 * ((size_t)D + 1), (size_t)tile0 * ((size_t)D + 1), s));
  CK(h->tile_base.reserve(tile1 + 1, tile0, s));
  CK(h->s_tile_start.reserve(ntiles_aff + 1, 0, s));
  int dimbits = 1; while ((1LL << dimbits) < (int64_t)D + 1) ++dimbits;
  int tilebits = 1; while ((1LL << tilebits) < ntiles_aff + 1) ++tilebits;   // + the "not indexed" sentinel tile
  if (m > 0) {
    // scratch sized for the most an append can touch (open tile + batch, longest vector seen): growing it later costs a
    // cudaMalloc + cudaFree (tens of ms each with the shard's large VMM reservations mapped) in the middle of a live stream
    const int64_t m_cap = std::max<int64_t>(m, (int64_t)(CR + n) * (int64_t)(h->max_nnz_seen + 8));
    CK(h->s_keys_in.reserve(m_cap, 0, s)); CK(h->s_keys_out.reserve(m_cap, 0, s)); CK(h->s_vals_in.reserve(m_cap, 0, s));
    k_emit_postings<<<cdiv(m, 256), 256, 0, s>>>(nnz_lo, nnz_new, row_lo, n_new, h->fwd_ptr.p, h->fwd_idx.p, h->fwd_val.p,
                                                  h->prune ? h->fwd_skip.p : nullptr, ntiles_aff, CR, tile0, dimbits, h->s_keys_in.p, h->s_vals_in.p);
    CK(cudaGetLastError()); h->kernel_launches++;
    size_t tmp = 0;
    unsigned long long* vals_out = reinterpret_cast<unsigned long long*>(h->post.p + post_base);
    if (h->algo == 3) { CK(h->s_vals_out.reserve(m_cap, 0, s)); vals_out = h->s_vals_out.p; }
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->s_keys_in.p, h->s_keys_out.p, h->s_vals_in.p, vals_out, m, 0, dimbits + tilebits, s));
    CK(h->cub_tmp.reserve(tmp, 0, s));
    CK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tmp, h->s_keys_in.p, h->s_keys_out.p, h->s_vals_in.p, vals_out, m, 0, dimbits + tilebits, s));
    h->kernel_launches += 4;
  }
  k_tile_starts<<<cdiv(ntiles_aff + 1, 128), 128, 0, s>>>(h->s_keys_out.p, m, ntiles_aff, dimbits, post_base, tile0, h->s_tile_start.p, h->tile_base.p);
  CK(cudaGetLastError());
  k_build_dir<<<cdiv(((int64_t)D + 1) * ntiles_aff, 256), 256, 0, s>>>(h->s_keys_out.p, m, ntiles_aff, D, dimbits, h->s_tile_start.p, tile0, h->dir.p);
  CK(cudaGetLastError()); h->kernel_launches += 2;
  if (h->algo == 3) {
    // dense-head extraction: pick each tile's dense dims, move their postings into dense fp32 rows,
    // compact the remaining postings and fix the directory
    CK(h->dn_cnt.reserve(tile1, tile0, s)); CK(h->tile_cnt.reserve(tile1, tile0, s));
    CK(h->dn_dim.reserve((size_t)tile1 * KD, (size_t)tile0 * KD, s)); CK(h->dn_len.reserve((size_t)tile1 * KD, (size_t)tile0 * KD, s));
    CK(h->dn_hash.reserve((size_t)tile1 * HS, (size_t)tile0 * HS, s));
    CK(h->dn_w.reserve((size_t)tile1 * KD * CR, (size_t)tile0 * KD * CR, s));
    k_dense_select<<<ntiles_aff, 1024, 0, s>>>(D, CR, tile0, n_new, h->dense_shift, h->dir.p, h->dn_cnt.p, h->dn_dim.p, h->dn_len.p, h->dn_hash.p, h->tile_cnt.p);
    CK(cudaGetLastError());
    k_tile_bases<<<1, 32, 0, s>>>(tile0, ntiles_aff, h->tile_cnt.p, h->tile_base.p);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(h->dn_w.p + (size_t)tile0 * KD * CR, 0, (size_t)ntiles_aff * KD * CR * sizeof(float), s));
    if (m > 0) {
      k_post_scatter<<<cdiv(m, 256), 256, 0, s>>>(m, h->s_keys_out.p, h->s_vals_out.p, dimbits, tile0, ntiles_aff, CR, h->s_tile_start.p, h->tile_base.p,
                                                   h->dn_cnt.p, h->dn_dim.p, h->dn_len.p, h->dn_hash.p, h->dn_w.p,
                                                   reinterpret_cast<unsigned long long*>(h->post.p), (long long)h->post.cap);
      CK(cudaGetLastError());
    }
    k_dir_fix<<<cdiv(((int64_t)D + 1) * ntiles_aff, 256), 256, 0, s>>>(ntiles_aff, D, tile0, h->dn_cnt.p, h->dn_dim.p, h->dn_len.p, h->dir.p);
    CK(cudaGetLastError()); h->kernel_launches += 4;
  }
  h->n_local = n_new; h->nnz = nnz_new; h->n_post = nnz_new; h->ntiles = tile1;
  return APSS_OK;
}

// Per-batch transposition of the (pruned) query batch into blocks of QB consecutive queries: for every block
// the distinct dimensions it uses and, per dimension, the (row, weight * 2^F) list of the queries having it.
// Also fixes the fixed-point scale F from the largest squared norm seen (Cauchy-Schwarz bound on any dot
// product) and the integer emission threshold with its guard band.
static int32_t transpose_query_blocks(apss_handle* h, int32_t n, int32_t batch_nnz, BlockArgs* blk_out, int* F_out, unsigned* thr_out) {
  cudaStream_t s = h->stream;
  const int D = h->cfg.dim;
  const double t = h->cfg.similarity_threshold;
  BlockArgs blk{};
  int F = 0; unsigned thr_int = 0;
  // fixed-point scale: every dot product is <= max squared norm (Cauchy-Schwarz); keep 2x headroom
  const double bound = std::max(h->max_sq, 1e-300) * (1.0 + 1e-6);
  // (dense-head kernel: u16 accumulators, sums stay below 2^15 + one quantum per shared dim)
  // (every shared dimension adds an over-shoot of up to one quantum: max_sq * 2^F + max_nnz must stay below 2^16;
  //  up to ~32 K components per vector this is the 2^15 headroom rule, beyond that the scale drops -- estimates only
  //  ever err upwards, so a coarser scale costs verify work, never a pair)
  F = h->algo == 3 ? (int)std::floor(std::log2(std::min(32768.0, 65535.0 - (double)h->max_nnz_seen - 1.0) / bound))
                   : (int)std::floor(std::log2(2147483648.0 / bound));
  F = std::max(-100, std::min(100, F));
  const double ts = t * std::ldexp(1.0, F) * (1.0 - std::ldexp(1.0, h->algo == 3 ? -16 : -20));
  const double tmax = h->algo == 3 ? 65535.0 : 4294967295.0;
  thr_int = ts <= 0 ? 0u : (ts >= tmax ? (unsigned)tmax : (unsigned)std::floor(ts));
  const int QB = h->QB; const int nqb = (n + QB - 1) / QB;
  int dimbits = 1; while ((1LL << dimbits) < (int64_t)D) ++dimbits;
  int qbbits = 1; while ((1LL << qbbits) < nqb) ++qbbits;
  CK(h->bt_keys_in.reserve(batch_nnz, 0, s)); CK(h->bt_keys_out.reserve(batch_nnz, 0, s)); CK(h->bt_vals_in.reserve(batch_nnz, 0, s));
  CK(h->bt_vals_out.reserve(batch_nnz, 0, s)); CK(h->ud_key.reserve(batch_nnz + 1, 0, s));
  CK(h->bt_flags.reserve(batch_nnz + 1, 0, s)); CK(h->bt_pos.reserve(batch_nnz + 1, 0, s));
  CK(h->ud_dim.reserve(batch_nnz + 1, 0, s)); CK(h->ud_start.reserve(batch_nnz + 2, 0, s)); CK(h->bd_ptr.reserve(nqb + 1, 0, s));
  k_bt_emit<<<cdiv(batch_nnz, 256), 256, 0, s>>>(n, batch_nnz, h->q_ptr.p, h->q_dim.p, h->q_w.p, QB, h->algo == 3 ? h->CR / 2 : h->CR, dimbits, (float)std::ldexp(1.0, F),
                                                 h->bt_keys_in.p, h->bt_vals_in.p);
  CK(cudaGetLastError());
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, h->bt_keys_in.p, h->bt_keys_out.p, h->bt_vals_in.p, h->bt_vals_out.p, batch_nnz, 0, dimbits + qbbits, s));
  CK(h->cub_tmp.reserve(tb, 0, s));
  CK(cub::DeviceRadixSort::SortPairs(h->cub_tmp.p, tb, h->bt_keys_in.p, h->bt_keys_out.p, h->bt_vals_in.p, h->bt_vals_out.p, batch_nnz, 0, dimbits + qbbits, s));
  k_bt_heads<<<cdiv(batch_nnz + 1, 256), 256, 0, s>>>(batch_nnz, h->bt_keys_out.p, h->bt_flags.p);
  CK(cudaGetLastError());
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, h->bt_flags.p, h->bt_pos.p, batch_nnz + 1, s));
  CK(h->cub_tmp.reserve(tb, 0, s));
  CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, h->bt_flags.p, h->bt_pos.p, batch_nnz + 1, s));
  k_bt_scatter<<<cdiv(batch_nnz + 1, 256), 256, 0, s>>>(batch_nnz, h->bt_keys_out.p, h->bt_flags.p, h->bt_pos.p, dimbits, h->ud_key.p, h->ud_dim.p, h->ud_start.p);
  CK(cudaGetLastError());
  k_bt_blocks<<<cdiv(nqb + 1, 128), 128, 0, s>>>(nqb, h->bt_pos.p, batch_nnz, h->ud_key.p, dimbits, h->bd_ptr.p);
  CK(cudaGetLastError());
  h->kernel_launches += 9;
  blk.ud_dim = h->ud_dim.p; blk.ud_start = h->ud_start.p; blk.bd_ptr = h->bd_ptr.p;
  blk.bt = reinterpret_cast<const uint2*>(h->bt_vals_out.p); blk.QB = QB; blk.n_qblocks = nqb;
    *blk_out = blk; *F_out = F; *thr_out = thr_int;
  return APSS_OK;
}

// How many query slices the candidate-major kernel scores this batch in: sized so that a stored vector meets
// ~160 query-list entries per slice on average (the per-warp table takes 512; the rate is the previous batch's).
static void plan_query_slices(const apss_handle* h, int32_t n, int* slices_out, int* qsub_out) {
  const int D = h->cfg.dim;
  int slices = 1;
  if (h->cand_slices_env) slices = h->cand_slices_env;
  else if (h->cand_rate > 0) slices = (int)std::ceil(h->cand_rate * n / (h->cand_warps == 32 ? 80.0 : 160.0));
  const int64_t max_by_mem = std::max<int64_t>(1, ((int64_t)256 << 20) / (((int64_t)D + 1) * 4));     // directories: <= 256 MB
  slices = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(slices, 256), std::min<int64_t>(max_by_mem, (n + 31) / 32)));
  const int qsub = (n + slices - 1) / slices;
  *slices_out = (n + qsub - 1) / qsub; *qsub_out = qsub;
}

// One scoring attempt of the candidate-major path: k_score_cand + k_score_cand_heavy per query slice.
static int32_t score_candidate_major(apss_handle* h, int32_t n, int32_t batch_nnz, int slices, int qsub, int64_t q_local_base,
                                     const int64_t* d_qkey) {
  cudaStream_t s = h->stream;
  const int D = h->cfg.dim;
  const double t = h->cfg.similarity_threshold;
  const int64_t n_rows = h->n_local;      // includes this batch when it was indexed (IWA:125-132)
  if (!n_rows || !batch_nnz) return APSS_OK;
  CandArgs ca{};
  ca.ifw_ptr = h->ifw_ptr.p; ca.ifw = h->ifw.p;
  ca.row_ub = h->row_ub.p; ca.c_key = h->key.p; ca.qi = reinterpret_cast<const uint2*>(h->bt_vals_out.p);
  ca.q_nrm = h->q_nrm.p; ca.q_key = h->custom_keys ? d_qkey : nullptr;
  ca.n_rows = n_rows; ca.q_local_base = q_local_base; ca.nq = n;
  ca.thr = (float)t; if ((double)ca.thr > t) ca.thr = std::nextafterf(ca.thr, -INFINITY);
  ca.band1 = (float)(1.0 + (double)(h->max_nnz_seen + 8) * std::ldexp(1.0, -22));
  {   // u32 fixed point: every dot product is <= the largest squared norm (Cauchy-Schwarz)
    const int Fc = std::max(-100, std::min(100, (int)std::floor(std::log2(2147483648.0 / (std::max(h->max_sq, 1e-300) * (1.0 + 1e-6))))));
    ca.scale = (float)std::ldexp(1.0, Fc); ca.inv_scale = (float)std::ldexp(1.0, -Fc);
  }
  ca.out_q = h->pf_q.p; ca.out_c = h->pf_c.p; ca.out_est = h->pf_est.p; ca.out_cap = h->pf_q.cap;
  ca.counters = h->d_counters; ca.heavy = h->heavy.p; ca.heavy_cap = (int64_t)h->heavy.cap;
  CK(cudaMemsetAsync(h->d_counters + C_HEAVY, 0, 2 * sizeof(unsigned long long), s));      // C_HEAVY, C_HEAVY_TOT
  for (int sl = 0; sl < slices; ++sl) {
    if (sl) { CK(cudaMemsetAsync(h->d_counters + C_WORK, 0, sizeof(unsigned long long), s)); CK(cudaMemsetAsync(h->d_counters + C_HEAVY, 0, sizeof(unsigned long long), s)); }
    ca.qdir = h->qdir.p + (size_t)sl * ((size_t)D + 1);
    ca.q_lo = sl * qsub; ca.q_hi = std::min(n, (sl + 1) * qsub);
    CK(launch_cand(h, ca));
    h->kernel_launches += 2;
  }
  h->score_launches++;
  return APSS_OK;
}

// One scoring attempt of the query-major path on the reduced index: cut the batch's lists into pieces
// (k_qm_count + scan + k_qm_emit), then k_score_qm.  h_total[1] receives the number of pieces needed.
static int32_t score_query_major(apss_handle* h, int32_t n, int32_t batch_nnz, int64_t q_local_base, const int64_t* d_qkey) {
  cudaStream_t s = h->stream;
  const double t = h->cfg.similarity_threshold;
  h->h_counters[C_ITEMS] = 0;
  if (!h->n_local || !batch_nnz || h->segs.empty()) return APSS_OK;
  SegList sl{}; sl.n = (int32_t)h->segs.size();
  for (int k = 0; k < sl.n; ++k) { sl.post[k] = h->seg_post(h->segs[k]); sl.dir[k] = h->seg_dir(h->segs[k]); }
  CK(h->qm_cnt.reserve((size_t)batch_nnz + 1, 0, s)); CK(h->qm_off.reserve((size_t)batch_nnz + 1, 0, s));
  if (!h->qm_items.cap) CK(h->qm_items.reserve(h->qm_items_cap0 ? h->qm_items_cap0 : std::max<size_t>((size_t)batch_nnz * 4, (size_t)1 << 16), 0, s));
  k_qm_count<<<cdiv((int64_t)batch_nnz + 1, 256), 256, 0, s>>>(batch_nnz, h->q_dim.p, sl, h->qm_cnt.p);
  CK(cudaGetLastError());
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, h->qm_cnt.p, h->qm_off.p, batch_nnz + 1, s));
  CK(h->cub_tmp.reserve(tb, 0, s));
  CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tb, h->qm_cnt.p, h->qm_off.p, batch_nnz + 1, s));
  CK(h->q_bkt.reserve((size_t)n * 32, 0, s));
  k_qm_qnorms<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, h->q_ptr.p, h->q_dim.p, h->q_w.p, h->df.p, h->q_bkt.p);
  CK(cudaGetLastError());
  QmArgs a{};
  {   // u32 fixed point: every dot product is <= the largest squared norm (Cauchy-Schwarz)
    const int Fc = std::max(-100, std::min(100, (int)std::floor(std::log2(2147483648.0 / (std::max(h->max_sq, 1e-300) * (1.0 + 1e-6))))));
    a.scale = (float)std::ldexp(1.0, Fc); a.inv_scale = (float)std::ldexp(1.0, -Fc);
  }
  k_qm_emit<<<cdiv(batch_nnz, 256), 256, 0, s>>>(batch_nnz, h->q_dim.p, h->q_w.p, a.scale, sl, h->qm_off.p, h->qm_items.p, (long long)h->qm_items.cap);
  CK(cudaGetLastError());
  a.q_ptr = h->q_ptr.p; a.item_off = h->qm_off.p; a.items = h->qm_items.p; a.item_cap = (long long)h->qm_items.cap;
  a.q_nrm = h->q_nrm.p; a.q_key = h->custom_keys ? d_qkey : nullptr; a.row_ub = h->row_ub.p; a.c_key = h->key.p;
  a.row_dfmin = h->row_dfmin.p; a.q_bkt = h->q_bkt.p;
  a.n_rows = h->n_local; a.q_local_base = q_local_base; a.nq = n;
  a.thr = (float)t; if ((double)a.thr > t) a.thr = std::nextafterf(a.thr, -INFINITY);
  a.band1 = (float)(1.0 + (double)(h->max_nnz_seen + 8) * std::ldexp(1.0, -22));
  {
    const double cm = std::sqrt(std::max(h->prune_lim, 0.0)) * (1.0 + 1e-9);
    a.cu_max = (float)cm; if ((double)a.cu_max < cm) a.cu_max = std::nextafterf(a.cu_max, INFINITY);
  }
  a.cap = h->qm_cap;
  a.out_q = h->pf_q.p; a.out_c = h->pf_c.p; a.out_est = h->pf_est.p; a.out_cap = h->pf_q.cap;
  a.counters = h->d_counters;
  CK(h->qm_deferred.reserve((size_t)n, 0, s));
  a.deferred = h->qm_deferred.p; a.deferred_cap = n;
  CK(cudaMemsetAsync(h->d_counters + C_HEAVY, 0, 2 * sizeof(unsigned long long), s));      // deferred count, ranged-kernel cursor
  if (h->qm_pipe) {
    // producer / consumer kernel for the queries that fit one table pass; the others land on the deferred list.  Its hot
    // candidates go to a chunked buffer (q = -1 marks unused entries), k_qm_filter applies the exact test afterwards.
    if (!h->hot_q.cap) {
      size_t c0 = std::max<size_t>((size_t)h->sm_count * 4 * QP_CHUNK, (size_t)1 << 22);
      { const char* e = getenv("APSS_QM_HOT_CAP"); if (e && atoll(e) >= 1) c0 = (size_t)atoll(e); }      // test hook: forces the grow-and-replay path
      CK(h->hot_q.reserve(c0, 0, s)); CK(h->hot_c.reserve(c0, 0, s)); CK(h->hot_est.reserve(c0, 0, s));
    }
    const size_t n_chunks = std::min<size_t>(h->hot_q.cap, 0xffff0000u) / QP_CHUNK;       // whole chunks only
    CK(h->hot_used.reserve(n_chunks, 0, s));
    CK(cudaMemsetAsync(h->hot_used.p, 0, sizeof(unsigned) * n_chunks, s));     // the chunk directory, not the buffer, is cleared
    CK(cudaMemsetAsync(h->d_counters + C_HOTN, 0, sizeof(unsigned long long), s));
    QmArgs ap = a; ap.cap = std::min(a.cap, QP_CAP);
    { const char* e = getenv("APSS_QM_DRY"); ap.dry = e ? atoi(e) : 0; }
    { const char* e = getenv("APSS_QM_SIZEF"); ap.sizef = e && atoi(e) >= 3 ? atoi(e) : 6; }
    ap.hot_q = h->hot_q.p; ap.hot_c = h->hot_c.p; ap.hot_est = h->hot_est.p; ap.hot_cap = (unsigned)(n_chunks * QP_CHUNK);
    ap.hot_used = h->hot_used.p; ap.n_chunks = (unsigned)n_chunks;
    {
      auto kern = h->custom_keys ? k_score_qm_flat<true> : k_score_qm_flat<false>;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QF_SMEM));
      kern<<<h->sm_count, 1024, QF_SMEM, s>>>(ap);
    }
    CK(cudaGetLastError());
    k_qm_filter<<<(unsigned)(n_chunks * QF_PARTS), 256, 0, s>>>(ap);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_counters + C_HOTN, h->d_counters + C_HOTN, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    a.from_list = 1; h->kernel_launches += 2;
  } else h->h_counters[C_HOTN] = 0;
  if (h->qm_nt == 512) {          // two CTAs per SM, half the table each
    a.cap = std::min(a.cap, QM_CAP2);
    const size_t smem = (size_t)(2 * QM_TBL2 + QM_HOT) * sizeof(unsigned);
    auto kern = h->custom_keys ? k_score_qm<512, QM_TBL2, true> : k_score_qm<512, QM_TBL2, false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<h->sm_count * 2, 512, smem, s>>>(a);
  } else {
    const size_t smem = (size_t)(2 * QM_TBL + QM_HOT) * sizeof(unsigned);
    auto kern = h->custom_keys ? k_score_qm<1024, QM_TBL, true> : k_score_qm<1024, QM_TBL, false>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, h->smem_optin - 1024)));
    kern<<<h->sm_count, 1024, smem, s>>>(a);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_counters + C_ITEMS, h->qm_off.p + batch_nnz, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  h->kernel_launches += 4; h->score_launches++;
  return APSS_OK;
}

// LSM policy: after a committed append, the youngest segments are merged while the next older one is at most
// merge_ratio (8) times their sum: the sizes of neighbouring segments differ by more than 8x, so a query term meets
// log_8(#batches) lists (3 at a million vectors); every posting is copied about (8/7) log_8(#batches) times by merges
// that are streaming copies.  Stream-ordered: the caller does not wait for it.
static int32_t merge_segments(apss_handle* h) {
  cudaStream_t s = h->stream;
  const int D = h->cfg.dim;
  const int size = (int)h->segs.size();
  if (size < 2) return APSS_OK;
  int j = 1; int64_t sum = h->segs.back().n_post;
  while (j < size) {
    const int64_t prev = h->segs[size - 1 - j].n_post;
    if (sum + prev > 0x7fff0000LL) break;                                  // directories are int32 per segment
    if (prev > h->merge_ratio * sum && size - j <= QM_MAXSEG - 8) break;
    sum += prev; ++j;
  }
  if (j < 2 || h->dir_free.empty()) return APSS_OK;
  apss_handle::Seg out; out.n_post = sum; out.cap_post = sum + 64;
  out.off = h->segs[size - j].off; out.row_lo = h->segs[size - j].row_lo; out.row_hi = h->segs.back().row_hi;
  CK(h->seg_arena[1].reserve((size_t)(sum + 64), 0, s));
  out.dir_slot = h->dir_free.back();
  MergeSrc m{}; m.n = j;
  for (int k = 0; k < j; ++k) { m.post[k] = h->seg_post(h->segs[size - j + k]); m.dir[k] = h->seg_dir(h->segs[size - j + k]); }
  k_merge_dir<<<cdiv((int64_t)D + 1, 256), 256, 0, s>>>(D, m, h->seg_dir(out));
  for (int k = 0; k < j; ++k) {
    const int64_t np = h->segs[size - j + k].n_post;
    if (np > 0) k_merge_copy<<<cdiv(np, 256), 256, 0, s>>>(k, (int)np, D, m, h->seg_dir(out), h->seg_arena[1].p);
  }
  { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return h->fail(APSS_E_CUDA, "segment merge: %s", cudaGetErrorString(e_)); }
  if (j == size) std::swap(h->seg_arena[0], h->seg_arena[1]);        // everything was merged: the scratch arena IS the index now
  else {
    CK(h->seg_arena[0].reserve((size_t)(out.off + out.cap_post), 0, s));
    CK(cudaMemcpyAsync(h->seg_arena[0].p + out.off, h->seg_arena[1].p, sizeof(uint2) * (size_t)sum, cudaMemcpyDeviceToDevice, s));
  }
  h->dir_free.pop_back();
  for (int k = 0; k < j; ++k) h->dir_free.push_back(h->segs[size - j + k].dir_slot);
  h->segs.resize(size - j);
  h->segs.push_back(out);
  h->kernel_launches += 1 + j; h->merges++; h->merged_postings += sum;
  return APSS_OK;
}

// One scoring attempt of the tile kernels (K2/K3): the row kernel, the query-block kernel or the dense-head kernel.
static int32_t score_tiles(apss_handle* h, int32_t n, int32_t batch_nnz, BlockArgs blk, int F, unsigned thr_int, int64_t q_local_base,
                           const int64_t* d_qkey) {
  const int D = h->cfg.dim;
  const double t = h->cfg.similarity_threshold;
  ScoreArgs a{};
  a.q_ptr = h->q_ptr.p; a.q_dim = h->q_dim.p; a.q_w = h->q_w.p; a.q_key = d_qkey;
  a.post = h->post.p; a.dir = h->dir.p; a.tile_base = h->tile_base.p; a.c_key = h->key.p;
  a.nq = n; a.ntiles = (int32_t)h->ntiles; a.D = D; a.CR = h->CR; a.q_local_base = q_local_base;
  if (t > 0) {     // fp32 estimate of the row kernel: guard band of one rounding per component
    const double band = (double)(h->max_nnz_seen + 8) * std::ldexp(1.0, -23);
    a.thr_emit = std::nextafterf((float)(t * (1.0 - band)), -INFINITY);
  } else a.thr_emit = -INFINITY;
  a.out_q = h->pf_q.p; a.out_c = h->pf_c.p; a.out_est = h->pf_est.p; a.out_cap = h->pf_q.cap;
  a.counters = h->d_counters;
  a.tile_cnt = h->algo == 3 ? h->tile_cnt.p : nullptr; a.post_cap = (long long)h->post.cap; a.seg_cap = h->seg_cap;
  a.total_items = (unsigned long long)h->ntiles * (unsigned long long)n;
  a.row_ub = h->prune ? h->row_ub.p : nullptr; a.q_nrm = h->prune ? h->q_nrm.p : nullptr;
  if (h->algo == 1) {
    if (a.total_items && batch_nnz) { CK(launch_score(h, a, h->custom_keys)); h->score_launches++; h->kernel_launches++; }
  } else if (h->ntiles && batch_nnz) {
    blk.thr_int = thr_int; blk.inv_scale = (float)std::ldexp(1.0, -F); blk.scale = (float)std::ldexp(1.0, F);
    a.total_items = (unsigned long long)h->ntiles * (unsigned long long)blk.n_qblocks;
    if (h->algo == 3) {
      DenseTiles dtl{h->dn_cnt.p, h->dn_dim.p, h->dn_len.p, h->dn_hash.p, h->dn_w.p};
      CK(launch_dense(h, a, blk, dtl, h->custom_keys));
    } else CK(launch_blk(h, a, blk, h->custom_keys));
    h->score_launches++; h->kernel_launches++;
  }
  return APSS_OK;
}

// What a batch may change on the host side of the handle; restored when the call fails after the index was touched.
struct BatchTxn {
  int64_t n_local, nnz, n_post, ntiles, next_id; size_t n_segs; int max_nnz_seen; double max_sq; bool custom_keys;
  bool touched = false;      // index_append ran (forward store / tiles / segments may hold the batch)
  bool df_updated = false;   // document frequencies include the batch
  int32_t batch_nnz = 0;
};

static int32_t insert_batch_impl(apss_handle* h, BatchTxn& txn, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                 const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags, apss_batch_result* out);

// Undo a failed batch: the append-only stores (forward store, compact store, posting segments) are simply cut
// back; the tile index rewrites its open tile in place, so there the handle is retired instead.
static void rollback_batch(apss_handle* h, const BatchTxn& txn) {
  const std::string why = h->err;
  bool ok = cudaSetDevice(h->device) == cudaSuccess;
  if (h->prune_mode != 2 && h->prune_mode != 3) ok = false;
  if (ok && txn.df_updated && txn.batch_nnz) {
    k_df_update<<<cdiv(txn.batch_nnz, 256), 256, 0, h->stream>>>(txn.batch_nnz, h->q_dim.p, h->df.p, -1);
    ok = cudaGetLastError() == cudaSuccess;
  }
  while (h->segs.size() > txn.n_segs) { h->dir_free.push_back(h->segs.back().dir_slot); h->segs.pop_back(); }
  if (ok) ok = cudaStreamSynchronize(h->stream) == cudaSuccess;
  h->n_local = txn.n_local; h->nnz = txn.nnz; h->n_post = txn.n_post; h->ntiles = txn.ntiles; h->next_id = txn.next_id;
  h->custom_keys = txn.custom_keys;
  h->last_n = -1; h->last_pairs = 0;
  if (!ok) { h->broken = true; cudaGetLastError(); }
  h->err = why + (ok ? " (batch rolled back: the index is unchanged)" : " (the index could not be rolled back: this handle is retired, destroy it)");
}

extern "C" int32_t apss_insert_batch(apss_handle* h, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                     const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags, apss_batch_result* out) {
  if (!h) return APSS_E_INVALID;
  if (h->multi) return multi_insert_batch(h, n, indptr, indices, values, ext_keys, first_dim, flags, out);
  if (h->broken) return h->fail(APSS_E_STATE, "handle retired after a failed batch that could not be rolled back; destroy it");
  BatchTxn txn{h->n_local, h->nnz, h->n_post, h->ntiles, h->next_id, h->segs.size(), h->max_nnz_seen, h->max_sq, h->custom_keys};
  const int32_t rc = insert_batch_impl(h, txn, n, indptr, indices, values, ext_keys, first_dim, flags, out);
  if (rc != APSS_OK && txn.touched) rollback_batch(h, txn);
  return rc;
}

static int32_t insert_batch_impl(apss_handle* h, BatchTxn& txn, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                 const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags, apss_batch_result* out) {
  if (n < 0 || (n > 0 && (!indptr || (!indices && !values)))) return h->fail(APSS_E_INVALID, "null batch arrays");
  if (n > (1 << 24)) return h->fail(APSS_E_INVALID, "batch too large: at most 2^24 vectors per call");
  if (h->n_local + n > 0x7fffff00LL) return h->fail(APSS_E_INVALID, "shard full: internal ids are int32");
  if (!((flags & APSS_BATCH_QUERY_ONLY) || h->frozen) && h->next_id + (int64_t)n > 0x7fffffffLL)
    return h->fail(APSS_E_INVALID, "id space exhausted: internal ids are int32 (next_id %lld + %d vectors)", (long long)h->next_id, n);
  CK(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const bool dev_ptrs = flags & APSS_BATCH_DEVICE_PTRS;
  const bool query_only = (flags & APSS_BATCH_QUERY_ONLY) || h->frozen;
  const int D = h->cfg.dim;
  h->last_n = -1; h->last_pairs = 0;
  apss_batch_result res{};
  res.n_vectors = n; res.id_base = h->next_id;
  if (n == 0) { h->last_n = 0; h->last_status.clear(); if (out) *out = res; return APSS_OK; }

  const auto hts0 = std::chrono::steady_clock::now();
  auto hmark = [&](int k) { h->host_us[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - hts0).count(); };
  CK(cudaEventRecord(h->ev_b0, s));
  // ---- stage the batch on the device
  const int64_t* d_ptr; const int32_t* d_idx; const double* d_val; const int64_t* d_keys = nullptr; const int32_t* d_first = nullptr;
  int64_t in_nnz = 0;
  if (dev_ptrs) {
    d_ptr = indptr; d_idx = indices; d_val = values; d_keys = ext_keys; d_first = first_dim;
  } else {
    in_nnz = indptr[n];
    if (indptr[0] != 0 || in_nnz < 0) return h->fail(APSS_E_INPUT, "indptr must start at 0 and be non-negative");
    if (in_nnz > 0x7fffff00LL) return h->fail(APSS_E_INVALID, "batch too large: at most 2^31 components per call");
    CK(h->b_ptr.reserve(n + 1, 0, s)); CK(h->b_idx.reserve(std::max<int64_t>(in_nnz, 1), 0, s)); CK(h->b_val.reserve(std::max<int64_t>(in_nnz, 1), 0, s));
    CK(cudaMemcpyAsync(h->b_ptr.p, indptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    if (in_nnz) {
      CK(cudaMemcpyAsync(h->b_idx.p, indices, sizeof(int32_t) * in_nnz, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(h->b_val.p, values, sizeof(double) * in_nnz, cudaMemcpyHostToDevice, s));
    }
    d_ptr = h->b_ptr.p; d_idx = h->b_idx.p; d_val = h->b_val.p;
    if (ext_keys) { CK(h->b_key.reserve(n, 0, s)); CK(cudaMemcpyAsync(h->b_key.p, ext_keys, sizeof(int64_t) * n, cudaMemcpyHostToDevice, s)); d_keys = h->b_key.p; }
    if (first_dim) { CK(h->b_first.reserve(n, 0, s)); CK(cudaMemcpyAsync(h->b_first.p, first_dim, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s)); d_first = h->b_first.p; }
  }
  if (h->cfg.semantics == APSS_SEM_R0 && !d_first) return h->fail(APSS_E_INVALID, "semantics R0 needs first_dim[]");

  // ---- K5: validate + admit + prune (count, scan, write)
  CK(h->q_cnt.reserve(n + 1, 0, s)); CK(h->q_ptr.reserve(n + 1, 0, s)); CK(h->q_status.reserve(n, 0, s));
  if (h->prune) CK(h->q_nrm.reserve(n, 0, s));
  CK(cudaMemsetAsync(h->d_counters, 0, C_COUNT * sizeof(unsigned long long), s));
  const double admit_thr = (flags & APSS_BATCH_SKIP_ADMIT) ? -INFINITY : h->cfg.similarity_threshold;
  k_prefilter_count<<<cdiv(((int64_t)n + 1) * 32, 256), 256, 0, s>>>(n, d_ptr, d_idx, d_val, D, h->maxw.p, admit_thr, h->cfg.index_threshold,
                                                     h->q_cnt.p, h->q_status.p, h->prune ? h->q_nrm.p : nullptr, h->d_counters);
  CK(cudaGetLastError());
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, h->q_cnt.p, h->q_ptr.p, n + 1, s));
  CK(h->cub_tmp.reserve(tmp, 0, s));
  CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp.p, tmp, h->q_cnt.p, h->q_ptr.p, n + 1, s));
  h->kernel_launches += 3;
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_total, h->q_ptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  h->last_status.resize(n);
  CK(cudaMemcpyAsync(h->last_status.data(), h->q_status.p, n, cudaMemcpyDeviceToHost, s));
  hmark(0);      // prefilter enqueued
  CK(cudaStreamSynchronize(s));
  hmark(1);      // prefilter done
  if (h->h_counters[C_ERR]) {   // all-or-nothing (Q9): the index has not been touched yet
    return h->fail(APSS_E_INPUT, h->h_counters[C_ERR] == 2 ? "indptr is not monotone / does not start at 0"
                                                           : "indices must be strictly increasing and < dim (SparseVector.scala:96-108)");
  }
  if (h->h_counters[C_TOTNNZ] > 0x7fffff00ULL)     // q_cnt / q_ptr are int32 (the device-pointer entry has no host-side bound)
    return h->fail(APSS_E_INVALID, "batch too large: at most 2^31 components per call");
  const int32_t batch_nnz = h->h_total[0];
  // dense-head tile kernel: two u16 accumulators per word; a sum is < max_sq * 2^F + one quantum per shared
  // dimension, so the longest vector bounds what the scale can absorb (see transpose_query_blocks)
  if (h->algo == 3 && h->prune_mode < 2 && std::max(h->max_nnz_seen, (int)h->h_counters[C_MAXNNZ]) > U16_MAX_NNZ)
    return h->fail(APSS_E_INPUT, "a vector keeps %d components after the value prune; the u16 tile kernel takes at most %d "
                                 "(use kernel_variant 2 << 16, u32 accumulators, or pruning = 3)", (int)h->h_counters[C_MAXNNZ], U16_MAX_NNZ);
  txn.batch_nnz = batch_nnz;
  res.n_rejected = (int32_t)h->h_counters[C_NREJ]; res.n_empty = (int32_t)h->h_counters[C_NEMPTY]; res.n_active = (int32_t)h->h_counters[C_NACTIVE];
  h->max_nnz_seen = std::max(h->max_nnz_seen, (int)h->h_counters[C_MAXNNZ]);
  {
    double sq; std::memcpy(&sq, &h->h_counters[C_MAXSQ], sizeof sq);
    // index reduction relies on |q| <= max_query_norm for EVERY query: refuse the batch rather than mis-score it
    if (h->prune && sq > h->max_qnorm * h->max_qnorm * (1.0 + 1e-9))
      return h->fail(APSS_E_INPUT, "pruning: a vector of this batch has L2 norm %.9g > max_query_norm %.9g", std::sqrt(sq), h->max_qnorm);
    if (sq > h->max_sq) h->max_sq = sq;
  }
  CK(h->q_dim.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->q_val.reserve(std::max(batch_nnz, 1), 0, s)); CK(h->q_w.reserve(std::max(batch_nnz, 1), 0, s));
  k_prefilter_write<<<cdiv((int64_t)n * 32, 256), 256, 0, s>>>(n, d_ptr, d_idx, d_val, h->cfg.index_threshold, h->q_status.p, h->q_ptr.p, h->q_dim.p, h->q_val.p, h->q_w.p);
  CK(cudaGetLastError()); h->kernel_launches++;
  if (d_keys) h->custom_keys = true;

  // ---- K1: index first, then query (IWA:125-132)
  int64_t q_local_base = -1;
  if (!query_only) {
    q_local_base = h->n_local;
    if (h->prune) { txn.df_updated = batch_nnz > 0; txn.touched = true; const int32_t rcp = prune_select(h, n, batch_nnz); if (rcp != APSS_OK) return rcp; }
    txn.touched = true;
    int32_t rc = index_append(h, n, batch_nnz, d_keys);
    if (rc != APSS_OK) return rc;
    h->next_id += n;
    if (getenv("APSS_TEST_FAIL_AFTER_APPEND")) return h->fail(APSS_E_NOMEM, "injected failure after the index append (APSS_TEST_FAIL_AFTER_APPEND)");
  }
  // commit of a successful indexing call (after the final synchronisation): skipped-component tally, size of the new
  // posting segment, then the stream-ordered segment merges
  auto commit_index = [&]() {
    if (!h->prune || query_only) return;
    const int64_t skipped = (int64_t)h->h_counters[C_SKIPPED];
    h->tot_skipped += skipped; h->n_post = h->nnz - h->tot_skipped;
    if (h->prune_mode == 3 && h->segs.size() > txn.n_segs) {
      apss_handle::Seg& sg = h->segs.back();
      sg.n_post = (int64_t)batch_nnz - skipped;
      if (sg.n_post <= 0) { h->dir_free.push_back(sg.dir_slot); h->segs.pop_back(); }
      const std::string keep = h->err;
      if (merge_segments(h) != APSS_OK) { cudaGetLastError(); h->err = keep; }     // left unmerged: the index stays valid
    }
  };

  if ((flags & APSS_BATCH_INDEX_ONLY) && !query_only) {
    CK(cudaEventRecord(h->ev_b1, s));
    if (h->prune) CK(cudaMemcpyAsync(h->h_counters + C_SKIPPED, h->d_counters + C_SKIPPED, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    commit_index();
    float ms0 = 0.f; CK(cudaEventElapsedTime(&ms0, h->ev_b0, h->ev_b1)); res.device_ms = ms0;
    h->last_n = n; h->last_pairs = 0;
    if (out) *out = res;
    return APSS_OK;
  }

  // ---- K2/K3: scoring + threshold/compaction, K4: fp64 verify.  Re-run on output overflow.
  const double t = h->cfg.similarity_threshold;
  // when keys are supplied, q_key lives in b_key (host path) or the caller's buffer (device path)
  const int64_t* d_qkey = d_keys;
  if (h->custom_keys && !d_qkey) {     // keys were used before: this batch gets the default keys (its internal ids)
    CK(h->b_key.reserve(n, 0, s));
    k_default_keys<<<cdiv(n, 256), 256, 0, s>>>(n, res.id_base, h->b_key.p);
    CK(cudaGetLastError()); h->kernel_launches++;
    d_qkey = h->b_key.p;
  }
  // per-batch query structure: the batch inverted by dimension (candidate-major kernel) or transposed into
  // query blocks (block kernels); the row kernel reads the CSR batch as it is
  BlockArgs blk{};
  int F = 0; unsigned thr_int = 0;
  int slices = 1, qsub = n;
  if (h->prune_mode == 2) {
    plan_query_slices(h, n, &slices, &qsub);
    const int32_t rc = build_query_index(h, n, batch_nnz, slices, qsub);
    if (rc != APSS_OK) return rc;
  } else if (h->prune_mode == 3) {
    // the pieces are cut inside score_query_major (they depend on the segments only)
  } else if (h->algo != 1 && batch_nnz) {
    const int32_t rc = transpose_query_blocks(h, n, batch_nnz, &blk, &F, &thr_int);
    if (rc != APSS_OK) return rc;
  }
  for (int attempt = 0; attempt < 4; ++attempt) {
    CK(cudaMemsetAsync(h->d_counters, 0, 7 * sizeof(unsigned long long), s));   // keep the prefilter tallies
    CK(cudaMemsetAsync(h->d_counters + C_PHASE, 0, 10 * sizeof(unsigned long long), s));      // phase timers + dense-phase tallies
    hmark(2);    // index append + per-batch query structures enqueued
    CK(cudaEventRecord(h->ev_s0, s));
    const int32_t rcs = h->prune_mode == 3 ? score_query_major(h, n, batch_nnz, q_local_base, d_qkey)
                      : h->prune_mode == 2 ? score_candidate_major(h, n, batch_nnz, slices, qsub, q_local_base, d_qkey)
                                           : score_tiles(h, n, batch_nnz, blk, F, thr_int, q_local_base, d_qkey);
    if (rcs != APSS_OK) return rcs;
    CK(cudaEventRecord(h->ev_s1, s));
    hmark(3);    // scoring enqueued
    CK(h->out_q.reserve(h->pf_q.cap, 0, s)); CK(h->out_c.reserve(h->pf_q.cap, 0, s)); CK(h->out_sim.reserve(h->pf_q.cap, 0, s));
    if (h->n_local) {
      k_verify<<<h->sm_count * 8, 256, 0, s>>>(h->d_counters, h->pf_q.cap, h->pf_q.p, h->pf_c.p, h->q_ptr.p, h->q_dim.p, h->q_val.p,
                                               h->fwd_ptr.p, h->fwd_idx.p, h->fwd_val.p, h->gid.p, t, h->cfg.semantics == APSS_SEM_R0, d_first,
                                               h->out_q.p, h->out_c.p, h->out_sim.p, h->d_counters);
      CK(cudaGetLastError()); h->kernel_launches++;
    }
    CK(cudaEventRecord(h->ev_b1, s));
    CK(cudaMemcpyAsync(h->h_counters, h->d_counters, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->h_counters + C_PHASE, h->d_counters + C_PHASE, 10 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    if (h->prune) CK(cudaMemcpyAsync(h->h_counters + C_SKIPPED, h->d_counters + C_SKIPPED, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    hmark(4);    // verify enqueued
    CK(cudaStreamSynchronize(s));
    hmark(5);    // all done
    const size_t items_need = h->prune_mode == 3 ? (size_t)(h->h_counters[C_ITEMS] >> 36) : 0;
    const bool items_short = items_need > h->qm_items.cap;
    const bool hot_short = h->prune_mode == 3 && h->qm_pipe && (size_t)h->h_counters[C_HOTN] > std::min<size_t>(h->hot_q.cap, 0xffff0000u) / QP_CHUNK * QP_CHUNK;
    if (hot_short) { const size_t need = (size_t)h->h_counters[C_HOTN] * 2; CK(h->hot_q.reserve(need, 0, s)); CK(h->hot_c.reserve(need, 0, s)); CK(h->hot_est.reserve(need, 0, s)); }
    if (h->h_counters[C_PF] <= h->pf_q.cap && !items_short && !hot_short) break;
    if (attempt == 3) return h->fail(APSS_E_NOMEM, "pair / piece buffer overflow persisted");
    if (items_short) CK(h->qm_items.reserve(items_need + items_need / 2 + 1024, 0, s));     // grow and replay
    if (h->h_counters[C_PF] > h->pf_q.cap) {
      const size_t need = (size_t)h->h_counters[C_PF] + 1024;
      CK(h->pf_q.reserve(need, 0, s)); CK(h->pf_c.reserve(need, 0, s)); CK(h->pf_est.reserve(need, 0, s));
    }
  }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev_s0, h->ev_s1)); res.score_ms = ms;
  CK(cudaEventElapsedTime(&ms, h->ev_b0, h->ev_b1)); res.device_ms = ms;
  res.n_prefilter = (int64_t)h->h_counters[C_PF];
  res.postings_visited = (int64_t)h->h_counters[C_POSTINGS];
  res.candidates_unique = (int64_t)h->h_counters[C_CANDS];
  res.n_pairs = (int64_t)h->h_counters[C_FINAL];
  res.n_pairs_r1 = (int64_t)h->h_counters[C_R1];
  res.dense_postings = (int64_t)h->h_counters[C_DENSE_POST]; res.dense_fma = (int64_t)h->h_counters[C_DENSE_FMA];
  for (int k = 0; k < 8; ++k) h->phase_cycles[k] = (int64_t)h->h_counters[C_PHASE + k];
  commit_index();
  res.work_items = (int64_t)((unsigned long long)h->ntiles * (unsigned long long)(h->algo == 1 ? n : (n + h->QB - 1) / h->QB));
  if (h->prune_mode == 2) {
    res.work_items = (int64_t)h->h_counters[C_HEAVY_TOT];      // (stored vector, query slice) pairs that took the heavy pass
    if (h->n_local > 0 && res.n_active > 0) h->cand_rate = (double)res.postings_visited / ((double)h->n_local * (double)res.n_active);
  }
  if (getenv("APSS_HOST_TIMING")) std::fprintf(stderr, "apss host us: prefilter-enq %.0f prefilter-done %.0f index-enq %.0f score-enq %.0f verify-enq %.0f done %.0f | device %.0f score %.0f\n",
                                               h->host_us[0], h->host_us[1], h->host_us[2], h->host_us[3], h->host_us[4], h->host_us[5], res.device_ms * 1e3, res.score_ms * 1e3);
  h->last_n = n; h->last_pairs = res.n_pairs;
  h->tot_postings += res.postings_visited; h->tot_cands += res.candidates_unique; h->tot_pairs += res.n_pairs; h->tot_pf += res.n_prefilter;
  h->tot_score_ms += res.score_ms;
  if (out) *out = res;
  return APSS_OK;
}

extern "C" int32_t apss_fetch_pairs(apss_handle* h, int32_t* q, int32_t* c, double* sim, int64_t capacity, int64_t* n_out) {
  if (!h) return APSS_E_INVALID;
  if (h->multi) return multi_fetch_pairs(h, q, c, sim, capacity, n_out);
  if (h->last_n < 0) return h->fail(APSS_E_STATE, "no completed batch to fetch from");
  CK(cudaSetDevice(h->device));
  const int64_t m = std::min<int64_t>(h->last_pairs, capacity < 0 ? 0 : capacity);
  if (m > 0) {
    if (q) CK(cudaMemcpyAsync(q, h->out_q.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, h->stream));
    if (c) CK(cudaMemcpyAsync(c, h->out_c.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, h->stream));
    if (sim) CK(cudaMemcpyAsync(sim, h->out_sim.p, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  if (n_out) *n_out = h->last_pairs;
  return APSS_OK;
}

extern "C" int32_t apss_pairs_device(apss_handle* h, const int32_t** q, const int32_t** c, const double** sim, int64_t* n) {
  if (!h) return APSS_E_INVALID;
  if (h->multi) return h->fail(APSS_E_STATE, "apss_pairs_device: the pairs of a multi-device handle live on several GPUs; use apss_fetch_pairs");
  if (h->last_n < 0) return h->fail(APSS_E_STATE, "no completed batch");
  if (q) *q = h->out_q.p;
  if (c) *c = h->out_c.p;
  if (sim) *sim = h->out_sim.p;
  if (n) *n = h->last_pairs;
  return APSS_OK;
}

extern "C" int32_t apss_fetch_status(apss_handle* h, uint8_t* status, int32_t capacity) {
  if (!h || !status) return APSS_E_INVALID;
  if (h->multi) return multi_fetch_status(h, status, capacity);
  if (h->last_n < 0) return h->fail(APSS_E_STATE, "no completed batch");
  const int32_t m = std::min<int32_t>(capacity, (int32_t)h->last_status.size());
  if (m > 0) std::memcpy(status, h->last_status.data(), m);
  return APSS_OK;
}

extern "C" int32_t apss_freeze(apss_handle* h) { if (!h) return APSS_E_INVALID; if (h->multi) return multi_freeze(h); h->frozen = true; return APSS_OK; }

extern "C" int32_t apss_set_next_id(apss_handle* h, int64_t next_id) {
  if (!h) return APSS_E_INVALID;
  if (h->multi) return multi_set_next_id(h, next_id);
  if (next_id < 0 || next_id > 0x7fffffffLL) return h->fail(APSS_E_INVALID, "next_id out of int32 range");
  h->next_id = next_id;
  return APSS_OK;
}

extern "C" int32_t apss_get_stats(apss_handle* h, apss_stats* out) {
  if (!h || !out) return APSS_E_INVALID;
  if (h->multi) return multi_get_stats(h, out);
  apss_stats s{};
  s.n_vectors = h->n_local; s.n_postings = h->n_post; s.n_tiles = h->prune_mode == 3 ? (int64_t)h->segs.size() : h->ntiles;   // pruning = 3: posting segments
  s.bytes_postings = h->n_post * 8; s.bytes_directory = s.n_tiles * ((int64_t)h->cfg.dim + 1) * 4;
  s.bytes_forward = h->nnz * 12 + (h->n_local + 1) * 8;
  s.tot_postings_visited = h->tot_postings; s.tot_candidates_unique = h->tot_cands; s.tot_pairs = h->tot_pairs; s.tot_prefilter = h->tot_pf;
  s.score_launches = h->score_launches; s.kernel_launches = h->kernel_launches; s.tot_score_ms = h->tot_score_ms;
  for (int k = 0; k < 8; ++k) s.phase_cycles[k] = h->phase_cycles[k];
  s.frozen = h->frozen; s.tile_vectors = h->CR; s.warps_per_cta = h->WARPS; s.sm_count = h->sm_count;
  s.n_unindexed = h->tot_skipped; s.segment_merges = h->merges; s.merged_postings = h->merged_postings; s.n_devices = 1;
  *out = s;
  return APSS_OK;
}

extern "C" const char* apss_last_error(apss_handle* h) { return h ? h->err.c_str() : "null handle"; }
// (multi-device handle: the stream of the first shard; every shard has its own)
extern "C" void* apss_stream(apss_handle* h);


// ================================================================================================ shard dispatch below the ABI
//
// apss_config.n_devices > 1 (replaces the remote router EPA:37-49,113-122): ONE handle owns several GPUs
// in one process.  The index is partitioned by internal-id range, block-cyclic: the batch with number b is indexed by
// shard b mod N (and scored there against itself: IWA:125-132), every other shard scores it query-only against its own
// vectors, so every pair is found exactly once.  One worker thread per GPU drives that GPU's engine (the single-device
// handle above); the calling thread only fans the batch out and merges the results:
//   host batch    every worker copies it to its own GPU (N host-to-device copies in parallel, one PCIe link each);
//   device batch  (APSS_BATCH_DEVICE_PTRS) staged once, then sent to the other GPUs with cudaMemcpyPeerAsync -- NVLink --
//                 and each engine reads its own copy.
// Pair lists stay on the shards until apss_fetch_pairs gathers them (candidate ids are global: the dispatcher owns the id
// space through apss_set_next_id).  A failure of the indexing shard is rolled back by that engine; a failure of a
// query-only shard after the owner has indexed retires the handle (the batch cannot be un-indexed from outside).
struct MultiState {
  std::vector<apss_handle*> sh;
  std::vector<std::thread> th;
  std::mutex mu; std::condition_variable cv_go, cv_done;
  std::vector<std::function<void()>> job; std::vector<uint64_t> want, done; bool stop = false;
  int64_t next_id = 0, batch_no = 0; bool frozen = false, broken = false;
  std::vector<apss_batch_result> res; std::vector<int32_t> rc;
  int32_t last_n = -1; int64_t last_pairs = 0; int owner_last = 0;
  // staging of device-pointer batches on every shard's GPU
  struct Stage { DevBuf<int64_t> ptr, key; DevBuf<int32_t> idx, first; DevBuf<double> val; };
  std::vector<Stage> stage;

  void worker(int w) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return stop || want[w] > seen; });
        if (stop) return;
        seen = want[w]; f = job[w];
      }
      f();
      { std::lock_guard<std::mutex> lk(mu); done[w] = seen; }
      cv_done.notify_all();
    }
  }
  // run f(w) on every shard's thread and wait for all of them
  void run_all(const std::function<void(int)>& f) {
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t w = 0; w < sh.size(); ++w) { job[w] = [f, w] { f((int)w); }; ++want[w]; }
    }
    cv_go.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { for (size_t w = 0; w < sh.size(); ++w) if (done[w] != want[w]) return false; return true; });
  }
};

static void multi_destroy(apss_handle* h) {
  MultiState* m = h->multi;
  { std::lock_guard<std::mutex> lk(m->mu); m->stop = true; }
  m->cv_go.notify_all();
  for (auto& t : m->th) if (t.joinable()) t.join();
  for (size_t w = 0; w < m->sh.size(); ++w) {
    if (!m->sh[w]) continue;
    cudaSetDevice(m->sh[w]->device);
    if (w < m->stage.size()) { auto& g = m->stage[w]; g.ptr.release(); g.key.release(); g.idx.release(); g.first.release(); g.val.release(); }
    apss_destroy(m->sh[w]);
  }
  delete m;
  delete h;
}

static int32_t multi_create(const apss_config* cfg, apss_handle** out) {
  const int N = cfg->n_devices;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return APSS_E_NO_DEVICE; }
  for (int a = 0; a < N; ++a) {
    if (cfg->device_ids[a] < 0 || cfg->device_ids[a] >= ndev) return APSS_E_NO_DEVICE;
    // (APSS_TEST_ALLOW_DUP_DEVICES: several shards on one GPU, so that the dispatch logic can be tested on a one-GPU box)
    for (int b = 0; b < a; ++b) if (cfg->device_ids[a] == cfg->device_ids[b] && !getenv("APSS_TEST_ALLOW_DUP_DEVICES")) return APSS_E_INVALID;
  }
  apss_handle* h = new apss_handle();
  h->cfg = *cfg; h->cfg.max_weight = nullptr; h->device = cfg->device_ids[0];
  MultiState* m = h->multi = new MultiState();
  m->sh.assign(N, nullptr); m->stage.resize(N); m->job.resize(N); m->want.assign(N, 0); m->done.assign(N, 0);
  m->res.resize(N); m->rc.assign(N, APSS_OK);
  for (int w = 0; w < N; ++w) {
    apss_config c = *cfg;
    c.n_devices = 0; c.device = cfg->device_ids[w];
    c.reserve_vectors = cfg->reserve_vectors > 0 ? cfg->reserve_vectors / N + (1 << 16) : 0;
    c.reserve_nnz = cfg->reserve_nnz > 0 ? cfg->reserve_nnz / N + (1 << 22) : 0;
    const int32_t rc = apss_create(&c, &m->sh[w]);
    if (rc != APSS_OK) { multi_destroy(h); return rc; }
  }
  // peer access for the NVLink fan-out of device-resident batches (best effort: without it the copies are staged by the driver)
  for (int a = 0; a < N; ++a) {
    cudaSetDevice(cfg->device_ids[a]);
    for (int b = 0; b < N; ++b) if (a != b) { int can = 0; cudaDeviceCanAccessPeer(&can, cfg->device_ids[a], cfg->device_ids[b]); if (can) cudaDeviceEnablePeerAccess(cfg->device_ids[b], 0); }
  }
  cudaGetLastError();
  for (int w = 0; w < N; ++w) m->th.emplace_back([m, w] { m->worker(w); });
  *out = h;
  return APSS_OK;
}

static int32_t multi_insert_batch(apss_handle* h, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                  const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags, apss_batch_result* out) {
  MultiState* m = h->multi;
  const int N = (int)m->sh.size();
  if (m->broken) return h->fail(APSS_E_STATE, "handle retired after a failed batch that could not be rolled back; destroy it");
  if (n < 0 || (n > 0 && (!indptr || (!indices && !values)))) return h->fail(APSS_E_INVALID, "null batch arrays");
  const bool query_only = (flags & APSS_BATCH_QUERY_ONLY) || m->frozen;
  const int owner = (int)(m->batch_no % N);
  m->last_n = -1; m->last_pairs = 0;
  // device-resident batch: one staged copy per GPU over NVLink, issued on the owning engine's stream by its worker
  const bool dev_ptrs = (flags & APSS_BATCH_DEVICE_PTRS) != 0;
  int src_dev = -1; int64_t nnz = 0;
  if (dev_ptrs && n > 0) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, indptr) != cudaSuccess || at.type != cudaMemoryTypeDevice) { cudaGetLastError(); return h->fail(APSS_E_INVALID, "APSS_BATCH_DEVICE_PTRS: indptr is not a device pointer"); }
    src_dev = at.device;
    cudaSetDevice(src_dev);
    if (cudaMemcpy(&nnz, indptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return h->fail(APSS_E_CUDA, "cannot read indptr[n]"); }
    if (nnz < 0 || nnz > 0x7fffff00LL) return h->fail(APSS_E_INVALID, "batch too large: at most 2^31 components per call");
  }
  m->run_all([&](int w) {
    apss_handle* e = m->sh[w];
    uint32_t f = flags;
    if (w != owner || query_only) f |= APSS_BATCH_QUERY_ONLY;
    if (w != owner) f &= ~(uint32_t)APSS_BATCH_INDEX_ONLY;
    if ((flags & APSS_BATCH_INDEX_ONLY) && w != owner && !query_only) {        // bulk load: the other shards have nothing to do
      m->res[w] = apss_batch_result{}; m->rc[w] = APSS_OK; return;
    }
    apss_set_next_id(e, m->next_id);                  // default keys and candidate ids are GLOBAL internal ids
    const int64_t* p = indptr; const int32_t* ix = indices; const double* v = values; const int64_t* k = ext_keys; const int32_t* fd = first_dim;
    if (dev_ptrs && n > 0 && e->device != src_dev) {
      MultiState::Stage& g = m->stage[w];
      cudaSetDevice(e->device);
      cudaStream_t s = e->stream;
      bool ok = g.ptr.reserve((size_t)n + 1, 0, s) == cudaSuccess && g.idx.reserve((size_t)std::max<int64_t>(nnz, 1), 0, s) == cudaSuccess &&
                g.val.reserve((size_t)std::max<int64_t>(nnz, 1), 0, s) == cudaSuccess;
      ok = ok && cudaMemcpyPeerAsync(g.ptr.p, e->device, indptr, src_dev, sizeof(int64_t) * ((size_t)n + 1), s) == cudaSuccess;
      if (ok && nnz) ok = cudaMemcpyPeerAsync(g.idx.p, e->device, indices, src_dev, sizeof(int32_t) * (size_t)nnz, s) == cudaSuccess &&
                          cudaMemcpyPeerAsync(g.val.p, e->device, values, src_dev, sizeof(double) * (size_t)nnz, s) == cudaSuccess;
      if (ok && ext_keys) { ok = g.key.reserve((size_t)n, 0, s) == cudaSuccess && cudaMemcpyPeerAsync(g.key.p, e->device, ext_keys, src_dev, sizeof(int64_t) * (size_t)n, s) == cudaSuccess; k = g.key.p; }
      if (ok && first_dim) { ok = g.first.reserve((size_t)n, 0, s) == cudaSuccess && cudaMemcpyPeerAsync(g.first.p, e->device, first_dim, src_dev, sizeof(int32_t) * (size_t)n, s) == cudaSuccess; fd = g.first.p; }
      if (!ok) { cudaGetLastError(); m->rc[w] = e->fail(APSS_E_CUDA, "peer copy of the batch to device %d failed", e->device); return; }
      p = g.ptr.p; ix = g.idx.p; v = g.val.p;
    }
    m->rc[w] = apss_insert_batch(e, n, p, ix, v, k, fd, f, &m->res[w]);
  });
  // ---- merge
  int32_t rc = m->rc[owner];
  int bad = rc != APSS_OK ? owner : -1;
  for (int w = 0; w < N && rc == APSS_OK; ++w) if (m->rc[w] != APSS_OK) { rc = m->rc[w]; bad = w; }
  if (rc != APSS_OK) {
    h->err = std::string("shard on device ") + std::to_string(m->sh[bad]->device) + ": " + m->sh[bad]->err;
    if (m->rc[owner] == APSS_OK && !query_only) {      // the owner has indexed a batch the caller is told has failed
      m->broken = true;
      h->err += " (the owning shard had already indexed the batch: this handle is retired, destroy it)";
    }
    return rc;
  }
  apss_batch_result r{};
  r.id_base = m->next_id; r.n_vectors = n;
  r.n_rejected = m->res[owner].n_rejected; r.n_empty = m->res[owner].n_empty; r.n_active = m->res[owner].n_active;
  for (int w = 0; w < N; ++w) {
    const apss_batch_result& x = m->res[w];
    r.n_pairs += x.n_pairs; r.n_pairs_r1 += x.n_pairs_r1; r.n_prefilter += x.n_prefilter; r.postings_visited += x.postings_visited;
    r.candidates_unique += x.candidates_unique; r.work_items += x.work_items; r.dense_postings += x.dense_postings; r.dense_fma += x.dense_fma;
    r.score_ms = std::max(r.score_ms, x.score_ms); r.device_ms = std::max(r.device_ms, x.device_ms);
  }
  if (!query_only) { m->next_id += n; m->batch_no += 1; }
  m->last_n = n; m->last_pairs = r.n_pairs; m->owner_last = owner;
  if (out) *out = r;
  return APSS_OK;
}

static int32_t multi_fetch_pairs(apss_handle* h, int32_t* q, int32_t* c, double* sim, int64_t capacity, int64_t* n_out) {
  MultiState* m = h->multi;
  if (m->last_n < 0) return h->fail(APSS_E_STATE, "no completed batch to fetch from");
  const int N = (int)m->sh.size();
  std::vector<int64_t> off(N + 1, 0);
  for (int w = 0; w < N; ++w) off[w + 1] = off[w] + m->res[w].n_pairs;
  const int64_t cap = capacity < 0 ? 0 : capacity;
  std::vector<int32_t> rc(N, APSS_OK);
  if (cap > 0 && (q || c || sim))
    m->run_all([&](int w) {                             // every shard copies its slice straight into the caller's arrays
      const int64_t lo = std::min(off[w], cap), hi = std::min(off[w + 1], cap);
      if (hi <= lo) return;
      int64_t got = 0;
      rc[w] = apss_fetch_pairs(m->sh[w], q ? q + lo : nullptr, c ? c + lo : nullptr, sim ? sim + lo : nullptr, hi - lo, &got);
    });
  for (int w = 0; w < N; ++w) if (rc[w] != APSS_OK) { h->err = m->sh[w]->err; return rc[w]; }
  if (n_out) *n_out = m->last_pairs;
  return APSS_OK;
}

static int32_t multi_fetch_status(apss_handle* h, uint8_t* status, int32_t capacity) {
  MultiState* m = h->multi;
  if (m->last_n < 0) return h->fail(APSS_E_STATE, "no completed batch");
  return apss_fetch_status(m->sh[m->owner_last], status, capacity);
}

static int32_t multi_freeze(apss_handle* h) { h->multi->frozen = true; for (apss_handle* e : h->multi->sh) apss_freeze(e); return APSS_OK; }

static int32_t multi_set_next_id(apss_handle* h, int64_t next_id) {
  if (next_id < 0 || next_id > 0x7fffffffLL) return h->fail(APSS_E_INVALID, "next_id out of int32 range");
  h->multi->next_id = next_id;
  return APSS_OK;
}

extern "C" void* apss_stream(apss_handle* h) { return !h ? nullptr : (h->multi ? (void*)h->multi->sh[0]->stream : (void*)h->stream); }

static int32_t multi_get_stats(apss_handle* h, apss_stats* out) {
  MultiState* m = h->multi;
  apss_stats s{};
  for (apss_handle* e : m->sh) {
    apss_stats x{};
    apss_get_stats(e, &x);
    s.n_vectors += x.n_vectors; s.n_postings += x.n_postings; s.n_tiles += x.n_tiles; s.bytes_postings += x.bytes_postings;
    s.bytes_directory += x.bytes_directory; s.bytes_forward += x.bytes_forward; s.tot_postings_visited += x.tot_postings_visited;
    s.tot_candidates_unique += x.tot_candidates_unique; s.tot_pairs += x.tot_pairs; s.tot_prefilter += x.tot_prefilter;
    s.score_launches += x.score_launches; s.kernel_launches += x.kernel_launches; s.tot_score_ms = std::max(s.tot_score_ms, x.tot_score_ms);
    s.n_unindexed += x.n_unindexed; s.segment_merges += x.segment_merges; s.merged_postings += x.merged_postings;
    s.tile_vectors = x.tile_vectors; s.warps_per_cta = x.warps_per_cta; s.sm_count = x.sm_count;
  }
  s.frozen = m->frozen; s.n_devices = (int32_t)m->sh.size();
  *out = s;
  return APSS_OK;
}

extern "C" int32_t apss_microbench_accumulators(int32_t device, int32_t mode, int32_t warps, int32_t iters, double* updates_per_sec) {
  if (!updates_per_sec || warps < 1 || warps > 32) return APSS_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return APSS_E_NO_DEVICE; }
  cudaSetDevice(device);
  cudaDeviceProp prop{}; cudaGetDeviceProperties(&prop, device);
  const int CR = warps > 16 ? 1024 : 2048;
  const size_t smem = (size_t)warps * CR * 4;
  unsigned* sink = nullptr; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int it) {
    switch (mode) {
#define MB(M) case M: cudaFuncSetAttribute(k_microbench<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                      k_microbench<M><<<prop.multiProcessorCount, warps * 32, smem>>>(CR, it, sink); break;
      MB(0) MB(1) MB(2) MB(3) MB(4) MB(5) MB(6)
#undef MB
      default: break;
    }
  };
  run(8); cudaDeviceSynchronize();
  cudaEventRecord(e0); run(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  cudaFree(sink); cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e != cudaSuccess) return APSS_E_CUDA;
  *updates_per_sec = (double)prop.multiProcessorCount * warps * 32.0 * (mode == 6 ? 32.0 : 4.0) * iters / (ms * 1e-3);
  return APSS_OK;
}
