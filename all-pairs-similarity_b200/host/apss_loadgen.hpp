// apss_loadgen.hpp -- C++17 mirror of the reference's latency driver, benchmark/LoadGenerator.scala:15-173 (LG below),
// above the same worker the actor scenario drives (apss_actor.hpp).  Same protocol as the Python mirror
// (all-pairs-similarity_b200/loadgen.py), kept AS BUILT:
//   * `childrenNum` LoadRunners (LG:105-109) tick every `writeBatchingDuration` ms (LG:44-46); a tick sends ONE vector,
//     videos(msgCount % videos.size) L2-normalised (LG:30-41), under the id msgCount.toString.  Runner i starts counting at
//     i * totalMessageCount (LG:22); the warm-up ends when msgCount > videos.size (LG:63-66, the crossing tick still sends).
//   * The parent's first ReceiveTimeout (`expDuration`, LG:100-102,159-168) sends StartTest: EVERY runner restarts at
//     msgCount = 0 (LG:79), so the test phase sends the same ids from all runners; each test tick reports
//     StartTime(id, now) first (LG:68) and the runner stops after totalMessageCount (LG:69-72).
//   * Response times (LG:135-156): endTime(q) = outputMoment whenever q's set of found pairs grows; postStop (LG:112-131)
//     prints count, average (integer division), max and min of endTime - startTime.
// Akka's dispatcher and timers become EventLoop: a deterministic single-threaded scheduler, one legal interleaving of
// the actors' mailboxes; virtual clock for tests, wall clock for measurements.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "apss_actor.hpp"

namespace apss_host {

struct StartTest {};                                                     // MSG:42
struct StartTime { std::string vectorId; int64_t moment; };             // MSG:43

class EventLoop {
 public:
  struct Event { std::function<void()> fn; int64_t period = 0; bool cancelled = false; };
  using Handle = std::shared_ptr<Event>;

  explicit EventLoop(bool virtual_clock = true, int64_t start_ms = 0) : virtual_(virtual_clock), now_(virtual_clock ? start_ms : wall()) {}
  int64_t now() const { return virtual_ ? now_ : wall(); }               // System.currentTimeMillis
  Handle schedule(int64_t delay_ms, std::function<void()> fn) { return push(now() + std::max<int64_t>(0, delay_ms), std::move(fn), 0); }
  Handle schedule_every(int64_t initial_ms, int64_t period_ms, std::function<void()> fn) {
    return push(now() + std::max<int64_t>(0, initial_ms), std::move(fn), std::max<int64_t>(1, period_ms));
  }
  static void cancel(const Handle& h) { if (h) h->cancelled = true; }
  void stop() { stopped_ = true; }                                       // context.system.shutdown()
  size_t run(size_t max_events = 10000000) {
    size_t n = 0;
    while (!q_.empty() && !stopped_ && n < max_events) {
      Item it = q_.top(); q_.pop();
      if (it.ev->cancelled) continue;
      if (virtual_) now_ = std::max(now_, it.at);
      else { const int64_t w = it.at - wall(); if (w > 0) std::this_thread::sleep_for(std::chrono::milliseconds(w)); }
      it.ev->fn();
      ++n;
      if (it.ev->period && !it.ev->cancelled) q_.push(Item{it.at + it.ev->period, ++seq_, it.ev});     // fixed rate
    }
    return n;
  }

 private:
  struct Item { int64_t at; uint64_t seq; Handle ev; };
  struct Later { bool operator()(const Item& a, const Item& b) const { return a.at != b.at ? a.at > b.at : a.seq > b.seq; } };
  static int64_t wall() {
    return (int64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
  }
  Handle push(int64_t at, std::function<void()> fn, int64_t period) {
    auto ev = std::make_shared<Event>(); ev->fn = std::move(fn); ev->period = period;
    q_.push(Item{at, ++seq_, ev});
    return ev;
  }
  bool virtual_; int64_t now_; bool stopped_ = false; uint64_t seq_ = 0;
  std::priority_queue<Item, std::vector<Item>, Later> q_;
};

class LoadGenerator;

class LoadRunner {                                                       // LG:15-92
 public:
  using Remote = std::function<void(const VectorIOMsg&)>;
  LoadRunner(int id, const Config& conf, const std::vector<IdVector>& videos, Remote remote, LoadGenerator* parent, EventLoop& loop)
      : writeBatching(std::stoll(conf_required(conf, "cpslab.allpair.benchmark.writeBatchingDuration"))),
        totalMessageCount(std::stoi(conf_required(conf, "cpslab.allpair.benchmark.totalMessageCount"))),
        vectorDim(std::stoi(conf_required(conf, "cpslab.allpair.vectorDim"))),
        msgCount((long long)id * totalMessageCount), videos_(videos), remote_(std::move(remote)), parent_(parent), loop_(loop) {
    ioTask_ = loop_.schedule_every(0, writeBatching, [this] { tick(); });                      // preStart, LG:44-46
  }
  VectorIOMsg generateVector() const {                                   // LG:30-41
    const SparkSparseVector& v = videos_[(size_t)(msgCount % (long long)videos_.size())].second;
    double sq = 0.0;
    for (double x : v.values) sq += x * x;
    const double squareSum = std::sqrt(sq);
    std::vector<double> values;
    for (double x : v.values) values.push_back(x / squareSum);
    return VectorIOMsg{{{std::to_string(msgCount), SparkSparseVector(vectorDim, v.indices, values)}}};
  }
  void receive(IOTicket);                                                // LG:59-74 (defined below: needs LoadGenerator)
  void receive(StartTest) {                                              // LG:75-83
    msgCount = 0; testPhaseStarted = true;
    EventLoop::cancel(ioTask_);       // (the reference leaves a still-running warm-up timer alive: two timers would double the rate)
    ioTask_ = loop_.schedule_every(0, writeBatching, [this] { tick(); });
  }
  const long long writeBatching; const int totalMessageCount, vectorDim;
  long long msgCount; bool testPhaseStarted = false, stopped = false;

 private:
  void tick() { if (!stopped) receive(IOTicket{}); }
  const std::vector<IdVector>& videos_; Remote remote_; LoadGenerator* parent_; EventLoop& loop_; EventLoop::Handle ioTask_;
};

class LoadGenerator {                                                    // LG:94-172
 public:
  struct Report { long long messages = 0, average = 0, max = 0, min = 0, with_both = 0; std::string line; };
  LoadGenerator(const Config& conf, const std::vector<IdVector>& videos, LoadRunner::Remote remote, EventLoop& loop,
                std::function<void(const std::string&)> log = nullptr)
      : totalMessageCount(std::stoi(conf_required(conf, "cpslab.allpair.benchmark.totalMessageCount"))),
        childNum(std::stoi(conf_required(conf, "cpslab.allpair.benchmark.childrenNum"))),
        expDuration(std::stoll(conf_required(conf, "cpslab.allpair.benchmark.expDuration"))), loop_(loop), log_(std::move(log)) {
    for (int i = 0; i < childNum; ++i) children.push_back(std::make_unique<LoadRunner>(i, conf, videos, remote, this, loop));   // preStart
    arm();
  }
  void receive(const SimilarityOutput& so) {                             // LG:135-156
    arm();
    if (!testPhaseStarted) return;
    for (const auto& [q, sims] : so.output) {
      for (const auto& [c, s] : sims) {
        auto it = findPair.find(q);
        const long long old = it == findPair.end() ? -1 : (long long)it->second.size();
        auto& fp = findPair[q];
        fp.insert({c, s});
        if ((long long)fp.size() != old) {
          // (the reference throws NoSuchElementException for an id without a StartTime; such ids are skipped in the log)
          if (log_ && startTime.count(q)) log_(q + " -> " + std::to_string(fp.size()) + " lasting Time:" + std::to_string(so.outputMoment - startTime[q]));
          endTime[q] = so.outputMoment;
        }
        if ((long long)fp.size() >= (long long)totalMessageCount * childNum - 1) readyVectors.insert(q);
      }
      if ((long long)readyVectors.size() >= (long long)totalMessageCount * childNum) loop_.stop();
    }
  }
  void receive(const StartTime& m) { arm(); startTime[m.vectorId] = m.moment; }                // LG:157-158
  void receive(ReceiveTimeout) {                                         // LG:159-168
    if (!testPhaseStarted) {
      testPhaseStarted = true;
      for (auto& w : children) w->receive(StartTest{});
      arm();
    } else loop_.stop();
  }
  Report report() const {                                                // postStop, LG:112-131
    Report r; r.messages = (long long)endTime.size();
    long long total = 0; bool any = false;
    for (const auto& [vid, start] : startTime) {
      auto it = endTime.find(vid);
      if (it == endTime.end()) continue;
      const long long d = it->second - start;
      total += d; ++r.with_both;
      if (!any || d > r.max) r.max = d;
      if (!any || d < r.min) r.min = d;
      any = true;
    }
    if (r.messages > 0) {
      r.average = total / r.messages;                                    // Long division
      r.line = "LoadGenerator stopped with " + std::to_string(r.messages) + " messages, average response time " +
               std::to_string(r.average) + ", max:" + std::to_string(r.max) + " min:" + std::to_string(r.min);
    }
    return r;
  }
  const int totalMessageCount, childNum; const long long expDuration;
  bool testPhaseStarted = false;
  std::map<std::string, int64_t> startTime, endTime;
  std::map<std::string, std::set<std::pair<std::string, double>>> findPair;
  std::set<std::string> readyVectors;
  std::vector<std::unique_ptr<LoadRunner>> children;

 private:
  void arm() {          // context.setReceiveTimeout: fires after expDuration without a message; every message re-arms it
    EventLoop::cancel(timeout_);
    if (expDuration > 0) timeout_ = loop_.schedule(expDuration, [this] { receive(ReceiveTimeout{}); });
  }
  EventLoop& loop_; std::function<void(const std::string&)> log_; EventLoop::Handle timeout_;
};

inline void LoadRunner::receive(IOTicket) {                              // LG:59-74
  ++msgCount;
  if (!testPhaseStarted) {
    if (msgCount > (long long)videos_.size()) EventLoop::cancel(ioTask_);
  } else {
    parent_->receive(StartTime{std::to_string(msgCount), loop_.now()});
    if (msgCount > totalMessageCount) { EventLoop::cancel(ioTask_); stopped = true; }          // context.stop(self), after this send
  }
  remote_(generateVector());
}

// Wire LoadGenerator -> worker -> LoadGenerator the way conf/app.conf does (remoteTarget = the entry actor, outputActor =
// the LoadGenerator) and run to shutdown.  The worker's own ReceiveTimeout (IWA:37-39,143-144) and its output IOTicket
// (IWA:48-50) are timers of the same loop; on a virtual clock the worker stamps outputMoment with that clock.
template <class Worker>
inline LoadGenerator::Report run_experiment(const Config& conf, const std::vector<IdVector>& videos, Worker& worker, EventLoop& loop,
                                            bool virtual_clock, std::function<void(const std::string&)> log = nullptr,
                                            std::function<void(int64_t, const VectorIOMsg&)> spy = nullptr) {
  const long long exp = std::stoll(conf_required(conf, "cpslab.allpair.benchmark.expDuration"));
  const long long out_every = std::stoll(conf_get(conf, "cpslab.allpair.outputIODuration", "0"));
  EventLoop::Handle wt;
  auto arm_worker = [&] {
    EventLoop::cancel(wt);
    if (exp > 0) wt = loop.schedule(exp, [&] { worker.receive(ReceiveTimeout{}); });
  };
  LoadGenerator* genp = nullptr;
  worker.setReplyTo([&](const OutMessage& m) { if (genp && std::holds_alternative<SimilarityOutput>(m)) genp->receive(std::get<SimilarityOutput>(m)); });
  if (virtual_clock) worker.now_ms = [&loop] { return loop.now(); };
  LoadGenerator gen(conf, videos, [&](const VectorIOMsg& m) { arm_worker(); if (spy) spy(loop.now(), m); worker.receive(m); }, loop, std::move(log));
  genp = &gen;
  arm_worker();
  if (out_every > 0) loop.schedule_every(0, out_every, [&] { worker.receive(IOTicket{}); });
  loop.run();
  return gen.report();
}

}  // namespace apss_host
