// TEST DOUBLE ONLY: an engine with the interface apss_host::GpuIndexingWorkerActor expects, backed by the CPU oracle
// (oracle/apss_oracle.c).  Lives under tests/ on purpose -- the product never routes through the oracle; this lets the
// C++ host logic be exercised on a box without a GPU.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "apss.h"

extern "C" {
typedef struct oracle_s oracle_t;
oracle_t* oracle_create(int32_t dim, double sim_thr, double idx_thr, int32_t semantics, int32_t algo, int32_t max_shard, int32_t max_entry,
                        int32_t max_index_actor, int32_t set_order, const double* maxw);
void oracle_destroy(oracle_t* o);
void oracle_freeze(oracle_t* o);
const char* oracle_last_error(oracle_t* o);
int32_t oracle_insert_batch(oracle_t* o, int32_t n, const int64_t* indptr, const int32_t* indices, const double* values, const int64_t* keys,
                            int32_t flags);
int64_t oracle_n_pairs(const oracle_t* o);
int64_t oracle_id_base(const oracle_t* o);
void oracle_fetch_pairs(const oracle_t* o, int32_t* q, int32_t* c, int64_t* qkey, int64_t* ckey, double* sim);
void oracle_fetch_status(const oracle_t* o, uint8_t* st);
}

class OracleEngine {
 public:
  OracleEngine(int dim, double t, double idx_thr, bool as_built)
      : o_(oracle_create(dim, t, idx_thr, as_built ? 1 : 0, as_built ? 0 /* faithful */ : 1 /* fast */, 1, 1, 1, 0, nullptr)) {}
  ~OracleEngine() { oracle_destroy(o_); }
  apss_batch_result insert_batch(int32_t n, const int64_t* indptr, const int32_t* indices, const double* values, const int64_t* ext_keys,
                                 const int32_t* /*first_dim: the faithful oracle derives its own*/, uint32_t flags) {
    const int32_t oflags = ((flags & APSS_BATCH_QUERY_ONLY) ? 1 : 0) | ((flags & APSS_BATCH_SKIP_ADMIT) ? 4 : 0);
    const int32_t rc = oracle_insert_batch(o_, n, indptr, indices, values, ext_keys, oflags);
    if (rc != 0) throw std::runtime_error(std::string("oracle: ") + oracle_last_error(o_));
    apss_batch_result r{};
    r.id_base = oracle_id_base(o_); r.n_vectors = n; r.n_pairs = oracle_n_pairs(o_);
    n_status_ = n;
    return r;
  }
  void fetch_pairs(std::vector<int32_t>& q, std::vector<int32_t>& c, std::vector<double>& sim, int64_t n_pairs) {
    q.resize((size_t)n_pairs); c.resize((size_t)n_pairs); sim.resize((size_t)n_pairs);
    if (n_pairs) oracle_fetch_pairs(o_, q.data(), c.data(), nullptr, nullptr, sim.data());
  }
  void fetch_status(std::vector<uint8_t>& st, int32_t n) { st.resize((size_t)n); if (n) oracle_fetch_status(o_, st.data()); }
  void freeze() { oracle_freeze(o_); }

 private:
  oracle_t* o_;
  int32_t n_status_ = 0;
};
