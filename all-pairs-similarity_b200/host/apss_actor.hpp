// apss_actor.hpp -- C++17 host side above the C ABI (include/apss.h): the reference's index worker and client
// call with the reference's names, argument meaning and error behaviour, without Akka.
//
// The reference is JVM code (Scala 2.10 / Akka 2.3.4); no JVM exists in the build image, so the compiled-language
// host mirror is C++ (the JNI + Scala binding a maintainer adds is in integration/, the same protocol in Python is
// all-pairs-similarity_b200/worker.py).  Paths cited are relative to /root/reference/core/src/main/scala/cpslab/:
//   IWA = deploy/server/IndexingWorkerActor.scala      WWA = deploy/server/WriteWorkerActor.scala
//   EPA = deploy/server/EntryProxyActor.scala          MSG = message/Message.scala
//   CC  = deploy/client/ClientConnection.scala         SV  = vector/SparseVector.scala
//
// Header only.  `GpuIndexingWorkerActor<Engine>` is written against a small engine concept; `CApiEngine` is the
// product engine (libapss_b200.so through include/apss.h -- CUDA or nothing), tests may plug a test double.
#pragma once

#include <algorithm>
#include <array>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "apss.h"

namespace apss_host {

// ------------------------------------------------------------------------------------------------ vectors

// java.lang.Double.toString: shortest digits that round-trip, decimal for 1e-3 <= |x| < 1e7, else d.dddE[-]n
inline std::string java_double_to_string(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
  if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::scientific);   // shortest round trip
  std::string sci(buf, r.ptr);                                                                   // d[.ddd]e[+-]nn
  const size_t epos = sci.find('e');
  std::string digits = sci.substr(0, epos);
  const int e10 = std::stoi(sci.substr(epos + 1));
  digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
  const std::string sign = x < 0 ? "-" : "";
  if (e10 >= -3 && e10 < 7) {
    std::string ip, fp;
    if (e10 >= 0) {
      ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)e10 + 1));
      ip.append((size_t)e10 + 1 - ip.size(), '0');
      fp = digits.size() > (size_t)e10 + 1 ? digits.substr((size_t)e10 + 1) : "0";
    } else {
      ip = "0";
      fp = std::string((size_t)(-e10 - 1), '0') + digits;
    }
    return sign + ip + "." + fp;
  }
  return sign + digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(e10);
}

// org.apache.spark.mllib.linalg.SparseVector as the reference uses it: (size, ascending indices, values).
struct SparkSparseVector {
  int size = 0;
  std::vector<int32_t> indices;
  std::vector<double> values;

  SparkSparseVector() = default;
  // SV:96-108: strictly increasing indices, all < size
  SparkSparseVector(int size_, std::vector<int32_t> idx, std::vector<double> val) : size(size_), indices(std::move(idx)), values(std::move(val)) {
    if (indices.size() != values.size()) throw std::invalid_argument("indices and values must have the same length");
    int prev = -1;
    for (int i : indices) {
      if (!(prev < i)) throw std::invalid_argument("Found duplicate indices: " + std::to_string(i) + ".");
      prev = i;
    }
    if (!(prev < size)) throw std::invalid_argument("index out of range");
  }
  // Vectors.sparse(size, Seq[(Int, Double)]): elements sorted by index
  static SparkSparseVector sparse(int size, std::vector<std::pair<int32_t, double>> elems) {
    std::sort(elems.begin(), elems.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<int32_t> idx; std::vector<double> val;
    for (const auto& e : elems) { idx.push_back(e.first); val.push_back(e.second); }
    return SparkSparseVector(size, std::move(idx), std::move(val));
  }
  // SV:204-205  "(size,[i,..],[v,..])"
  std::string toString() const {
    std::string s = "(" + std::to_string(size) + ",[";
    for (size_t k = 0; k < indices.size(); ++k) s += (k ? "," : "") + std::to_string(indices[k]);
    s += "],[";
    for (size_t k = 0; k < values.size(); ++k) s += (k ? "," : "") + java_double_to_string(values[k]);
    return s + "])";
  }
};

using IdVector = std::pair<std::string, SparkSparseVector>;     // (String id, vector): MSG:13

// ------------------------------------------------------------------------------------------------ messages (MSG:10-43)

struct VectorIOMsg { std::vector<IdVector> vectors; };                                  // MSG:13
struct SparseVectorWrapper { std::set<int32_t> indices; IdVector sparseVector; };       // SparseVectorWrapper.scala:9
struct DataPacket { int32_t shardId; std::vector<SparseVectorWrapper> vectors; };      // MSG:16
struct IndexData { std::vector<SparseVectorWrapper> vectors; };                         // MSG:18
struct IOTicket {};                                                                     // MSG:39
struct IOTrigger {};
struct ReceiveTimeout {};                                                               // akka.actor.ReceiveTimeout
struct Test { std::string content; bool operator==(const Test& o) const { return content == o.content; } };   // MSG:37

// MSG:20-35.  The reference's mutable.HashMap iterates in an unspecified order; std::map gives a stable one.
struct SimilarityOutput {
  std::map<std::string, std::map<std::string, double>> output;
  int64_t outputMoment = 0;
  std::string toString() const {                                                        // MSG:23-34
    std::string sb;
    for (const auto& [q, sims] : output) {
      sb += "---------------------------------";
      sb += q + ":";
      for (const auto& [c, s] : sims) sb += c + "," + java_double_to_string(s) + ";";
      sb += "\n";
    }
    return sb;
  }
};

using OutMessage = std::variant<SimilarityOutput, Test>;
using Config = std::map<std::string, std::string>;       // flat typesafe-config keys, e.g. "cpslab.allpair.vectorDim"

inline const std::string& conf_required(const Config& c, const std::string& key) {
  auto it = c.find(key);
  if (it == c.end()) throw std::out_of_range("No configuration setting found for key '" + key + "'");   // ConfigException.Missing
  return it->second;
}
inline std::string conf_get(const Config& c, const std::string& key, const std::string& dflt) {
  auto it = c.find(key);
  return it == c.end() ? dflt : it->second;
}

// ------------------------------------------------------------------------------------------------ Scala Set order (R0 only)

inline uint32_t scala_improve(uint32_t h) {        // scala.collection.immutable.HashSet.improve, 2.10.4 (recalled, unverified)
  h = h + ~(h << 9);
  h ^= h >> 14;
  h = h + (h << 4);
  return h ^ (h >> 10);
}
// first element, in iteration order, of the immutable Set built from ascending dims (WWA:172, IWA:102)
inline int32_t scala_set_first(const std::vector<int32_t>& dims) {
  if (dims.empty()) return -1;
  if (dims.size() <= 4) return dims[0];
  auto key = [](int32_t x) {
    const uint32_t h = scala_improve((uint32_t)x);
    std::array<uint32_t, 7> k{};
    for (int lvl = 0; lvl < 7; ++lvl) k[lvl] = (h >> (5 * lvl)) & 31u;
    return k;
  };
  return *std::min_element(dims.begin(), dims.end(), [&](int32_t a, int32_t b) { return key(a) < key(b); });
}

// ------------------------------------------------------------------------------------------------ product engine: the C ABI

// One apss_handle.  Throws std::runtime_error with the library's text on any non-zero status: there is no fallback.
class CApiEngine {
 public:
  CApiEngine(int dim, double similarity_threshold, double index_threshold, int device, int semantics, int pruning = 0,
             const std::vector<int>& devices = {}) {
    apss_config cfg{};
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.dim = dim; cfg.similarity_threshold = similarity_threshold; cfg.index_threshold = index_threshold;
    cfg.device = device; cfg.semantics = semantics; cfg.pruning = pruning;
    if (!devices.empty()) {                       // the GPUs that share the index: id-range shards below the C ABI
      if (devices.size() > APSS_MAX_DEVICES) throw std::invalid_argument("too many devices");
      cfg.n_devices = (int32_t)devices.size(); cfg.device = devices[0];
      for (size_t i = 0; i < devices.size(); ++i) cfg.device_ids[i] = devices[i];
    }
    const int32_t rc = apss_create(&cfg, &h_);
    if (rc != APSS_OK) throw std::runtime_error("apss_create failed: status " + std::to_string(rc));
  }
  ~CApiEngine() { if (h_) apss_destroy(h_); }
  CApiEngine(const CApiEngine&) = delete;
  CApiEngine& operator=(const CApiEngine&) = delete;

  apss_batch_result insert_batch(int32_t n, const int64_t* indptr, const int32_t* indices, const double* values,
                                 const int64_t* ext_keys, const int32_t* first_dim, uint32_t flags) {
    apss_batch_result r{};
    check(apss_insert_batch(h_, n, indptr, indices, values, ext_keys, first_dim, flags, &r));
    return r;
  }
  void fetch_pairs(std::vector<int32_t>& q, std::vector<int32_t>& c, std::vector<double>& sim, int64_t n_pairs) {
    q.resize((size_t)n_pairs); c.resize((size_t)n_pairs); sim.resize((size_t)n_pairs);
    int64_t got = 0;
    check(apss_fetch_pairs(h_, q.data(), c.data(), sim.data(), n_pairs, &got));
    q.resize((size_t)got); c.resize((size_t)got); sim.resize((size_t)got);
  }
  void fetch_status(std::vector<uint8_t>& st, int32_t n) { st.resize((size_t)n); if (n) check(apss_fetch_status(h_, st.data(), n)); }
  void freeze() { check(apss_freeze(h_)); }

 private:
  void check(int32_t rc) { if (rc != APSS_OK) throw std::runtime_error(std::string("apss error ") + std::to_string(rc) + ": " + apss_last_error(h_)); }
  apss_handle* h_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ the index worker

// IndexingWorkerActor (IWA:21-149) with the inverted index and the scoring loop behind `Engine`.
// conf keys (same names as the reference; IWA:23,26,33,44 / WWA:31,35):
//   cpslab.allpair.similarityThreshold, cpslab.allpair.outputIODuration, cpslab.allpair.benchmark.expDuration
//   (must exist), cpslab.allpair.vectorDim, cpslab.allpair.indexThreshold (default 0),
//   cpslab.allpair.gpu.semantics ("R1" | "R0"), cpslab.allpair.gpu.device, cpslab.allpair.gpu.pruning.
template <class Engine>
class GpuIndexingWorkerActor {
 public:
  using Reply = std::function<void(const OutMessage&)>;

  GpuIndexingWorkerActor(const Config& conf, Engine& engine, Reply replyTo = nullptr)
      : similarityThreshold(std::stod(conf_required(conf, "cpslab.allpair.similarityThreshold"))),
        outputWritingDuration(std::stoll(conf_required(conf, "cpslab.allpair.outputIODuration"))),
        expDuration(std::stoll(conf_required(conf, "cpslab.allpair.benchmark.expDuration"))),
        vectorDim(std::stoi(conf_required(conf, "cpslab.allpair.vectorDim"))),
        indexThreshold(std::stod(conf_get(conf, "cpslab.allpair.indexThreshold", "0"))),
        as_built(conf_get(conf, "cpslab.allpair.gpu.semantics", "R1") == "R0"),
        engine_(engine), replyTo_(std::move(replyTo)) {}

  // ---- IWA:122-148
  void receive(const IndexData& m) {            // wrappers carry admitted, pruned vectors (EPA:97, WWA:192-194)
    std::vector<IdVector> vs; std::vector<int32_t> firsts;
    for (const auto& w : m.vectors) {
      vs.push_back(w.sparseVector);
      // as built, the skipped first posting list (IWA:89 + IWA:106-107) is the first element of the WRAPPER's Set (IWA:102)
      if (as_built) { std::vector<int32_t> d(w.indices.begin(), w.indices.end()); std::sort(d.begin(), d.end()); firsts.push_back(scala_set_first(d)); }
    }
    handle_batch(vs, /*skip_admit=*/true, as_built ? &firsts : nullptr);
  }
  void receive(const VectorIOMsg& m) { handle_batch(m.vectors, /*skip_admit=*/false); }
  void receive(IOTicket) {                                                   // IWA:138-142
    if (!writeBuffer.empty()) {
      reply(SimilarityOutput{writeBuffer, now_ms()});
      writeBuffer.clear();
    }
  }
  void receive(ReceiveTimeout) { stopUpdateIndex = true; engine_.freeze(); }  // IWA:143-144
  void receive(const Test& t) { reply(t); }                                  // IWA:145-147

  // buildInvertedIndex + querySimilarItems (IWA:61-111) for one batch; returns outputSimSet
  std::map<std::string, std::map<std::string, double>> query_and_index(const std::vector<IdVector>& vectors, bool skip_admit,
                                                                       const std::vector<int32_t>* firsts = nullptr) {
    std::map<std::string, std::map<std::string, double>> out;
    if (vectors.empty()) return out;
    const int32_t n = (int32_t)vectors.size();
    std::vector<int64_t> indptr(1, 0);
    std::vector<int32_t> indices; std::vector<double> values;
    for (const auto& v : vectors) {
      if (v.second.size != vectorDim)          // CommonUtils.scala:99 `require`
        throw std::invalid_argument("requirement failed: vector1 size: " + std::to_string(v.second.size) + ", vector2 size: " + std::to_string(vectorDim));
      indices.insert(indices.end(), v.second.indices.begin(), v.second.indices.end());
      values.insert(values.end(), v.second.values.begin(), v.second.values.end());
      indptr.push_back((int64_t)indices.size());
    }
    const int64_t base = (int64_t)ids_.size();
    std::vector<int64_t> keys((size_t)n);
    // String ids never cross the ABI: key = internal id of the first occurrence.  Nothing is recorded in first_of_ / ids_
    // until the engine has accepted the batch (a refused batch leaves no trace).
    std::map<std::string, int64_t> fresh; bool dups = dups_;
    for (int32_t i = 0; i < n; ++i) {
      auto it = first_of_.find(vectors[i].first);
      if (it != first_of_.end()) { dups = true; keys[i] = it->second; continue; }
      auto ins = fresh.emplace(vectors[i].first, base + i);
      if (!ins.second) dups = true;
      keys[i] = ins.first->second;
    }
    std::vector<int32_t> first_dim;
    if (as_built && firsts) first_dim = *firsts;
    else if (as_built) {
      first_dim.resize((size_t)n);
      for (int32_t i = 0; i < n; ++i) {
        std::vector<int32_t> kept;
        for (int64_t p = indptr[i]; p < indptr[i + 1]; ++p) if (values[p] > indexThreshold) kept.push_back(indices[p]);
        first_dim[i] = scala_set_first(kept);
      }
    }
    const uint32_t flags = (stopUpdateIndex ? APSS_BATCH_QUERY_ONLY : 0u) | (skip_admit ? APSS_BATCH_SKIP_ADMIT : 0u);
    const apss_batch_result res = engine_.insert_batch(n, indptr.data(), indices.data(), values.data(), dups ? keys.data() : nullptr,
                                                       as_built ? first_dim.data() : nullptr, flags);      // throws: batch dropped, nothing recorded
    dups_ = dups;
    if (!stopUpdateIndex) first_of_.insert(fresh.begin(), fresh.end());
    std::vector<uint8_t> status; engine_.fetch_status(status, n);
    std::vector<int32_t> q, c; std::vector<double> sim; engine_.fetch_pairs(q, c, sim, res.n_pairs);
    if (!stopUpdateIndex) for (const auto& v : vectors) ids_.push_back(v.first);
    for (int32_t i = 0; i < n; ++i) if (status[i] == APSS_ST_ACTIVE) out[vectors[i].first];     // a key for every q with >= 1 dim (IWA:106)
    for (size_t k = 0; k < q.size(); ++k) out[vectors[(size_t)q[k]].first][ids_.at((size_t)c[k])] = sim[k];
    last_result = res;
    return out;
  }

  const double similarityThreshold;
  const long long outputWritingDuration, expDuration;
  const int vectorDim;
  const double indexThreshold;
  const bool as_built;
  bool stopUpdateIndex = false;                                              // IWA:35
  std::map<std::string, std::map<std::string, double>> writeBuffer;          // IWA:28
  apss_batch_result last_result{};

  // System.currentTimeMillis; protocol tests put a virtual clock here (apss_loadgen.hpp)
  std::function<int64_t()> now_ms = [] {
    return (int64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
  };
  void setReplyTo(Reply r) { replyTo_ = std::move(r); }          // outputActor (IWA:44), resolved after construction

 private:
  void reply(const OutMessage& m) { if (replyTo_) replyTo_(m); }
  void handle_batch(const std::vector<IdVector>& vectors, bool skip_admit, const std::vector<int32_t>* firsts = nullptr) {
    try {                                                                    // IWA:124
      auto out = query_and_index(vectors, skip_admit, firsts);
      if (replyTo_) {                                                        // IWA:128
        if (outputWritingDuration <= 0) reply(SimilarityOutput{std::move(out), now_ms()});     // IWA:129-130
        else for (const auto& [qid, sims] : out) for (const auto& [cid, s] : sims) writeBuffer[qid][cid] = s;   // IWA:113-120
      }
    } catch (const std::exception& e) {                                      // IWA:135-137: printStackTrace, batch dropped
      std::fprintf(stderr, "GpuIndexingWorkerActor: %s\n", e.what());
    }
  }
  Engine& engine_;
  Reply replyTo_;
  std::vector<std::string> ids_;                  // internal id -> caller's String id
  std::map<std::string, int64_t> first_of_;       // String id -> internal id of its first occurrence
  bool dups_ = false;
};

// ------------------------------------------------------------------------------------------------ router + client

// What regionRouter -> ShardRegion -> EntryProxyActor -> WriteWorkerActor amount to for one GPU worker
// (SimilaritySearchService.scala:28-32, EPA:95-111, WWA:164-202): vectors wait for the IOTrigger tick and form ONE
// batch; with ioTriggerPeriod <= 0 every VectorIOMsg is its own batch (parity configuration P0).
template <class Engine>
class RegionRouter {
 public:
  RegionRouter(const Config& conf, GpuIndexingWorkerActor<Engine>& worker)
      : ioTriggerPeriod(std::stoll(conf_get(conf, "cpslab.allpair.ioTriggerPeriod", "0"))), worker_(worker) {}
  void tell(const VectorIOMsg& m) {
    if (ioTriggerPeriod <= 0) worker_.receive(m);
    else buffer_.insert(buffer_.end(), m.vectors.begin(), m.vectors.end());
  }
  void tell(IOTrigger) {
    if (!buffer_.empty()) { VectorIOMsg m{std::move(buffer_)}; buffer_.clear(); worker_.receive(m); }
  }
  // EPA:113-122 handleDataPacket: the reference splits the packet by dimension over its index workers (EPA:37-49); an
  // id-range shard holds all dimensions of its vectors, so the whole packet becomes ONE IndexData
  void tell(const DataPacket& m) { worker_.receive(IndexData{m.vectors}); }
  template <class M> void tell(const M& m) { worker_.receive(m); }
  const long long ioTriggerPeriod;

 private:
  GpuIndexingWorkerActor<Engine>& worker_;
  std::vector<IdVector> buffer_;
};

// in-process stand-in for the ActorSystem argument of ClientConnection
template <class Engine>
class LocalActorSystem {
 public:
  void registerRouter(const std::string& address, RegionRouter<Engine>& r) { routes_["akka.tcp://ClusterSystem@" + address + "/user/regionRouter"] = &r; }
  RegionRouter<Engine>& actorSelection(const std::string& path) const { return *routes_.at(path); }

 private:
  std::map<std::string, RegionRouter<Engine>*> routes_;
};

// CC:10-34: fire-and-forget VectorIOMsg to a randomly chosen regionRouter
template <class Engine>
class ClientConnection {
 public:
  ClientConnection(const std::vector<std::string>& remoteAddresses, const LocalActorSystem<Engine>& localActorSystem) {
    for (const auto& a : remoteAddresses)                                                        // CC:12-21
      remoteRouters_.push_back(&localActorSystem.actorSelection("akka.tcp://ClusterSystem@" + a + "/user/regionRouter"));
  }
  void insertNewVector(const std::vector<IdVector>& vectors) {                                   // CC:31-33
    remoteRouters_[rng_() % remoteRouters_.size()]->tell(VectorIOMsg{vectors});                  // CC:24-25
  }
  // README.md:8-10 form: vectors without ids get generated ones
  void insertNewVector(const std::vector<SparkSparseVector>& vectors) {
    std::vector<IdVector> vs;
    for (const auto& v : vectors) vs.emplace_back("auto-" + std::to_string(auto_++), v);
    insertNewVector(vs);
  }

 private:
  std::vector<RegionRouter<Engine>*> remoteRouters_;
  std::mt19937 rng_{20260102u};
  long long auto_ = 0;
};

}  // namespace apss_host
