// Drives apss_host::GpuIndexingWorkerActor / RegionRouter / ClientConnection like the reference's actors are
// driven (same scenario as tests/test_worker_mirror.py).  Built twice by tests/test_cpp_host.py:
//   -DUSE_ORACLE_ENGINE  against the CPU oracle (test double)          -- runs anywhere
//   (default)            against libapss_b200.so through include/apss.h -- needs a GPU
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "apss_actor.hpp"
#include "apss_loadgen.hpp"
#ifdef USE_ORACLE_ENGINE
#include "oracle_engine.hpp"
using Engine = OracleEngine;
#else
using Engine = apss_host::CApiEngine;
#endif

using namespace apss_host;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond); ++failures; } } while (0)

static SparkSparseVector V(std::vector<std::pair<int32_t, double>> e, int size = 64) { return SparkSparseVector::sparse(size, std::move(e)); }

static std::vector<int> g_devices;      // GPUs behind the one engine the actor holds (shards below the C ABI); empty = GPU 0

static Engine* make_engine(int dim, double t, double idx_thr, bool as_built, int pruning) {
#ifdef USE_ORACLE_ENGINE
  (void)pruning;
  return new Engine(dim, t, idx_thr, as_built);
#else
  return new Engine(dim, t, idx_thr, 0, as_built ? APSS_SEM_R0 : APSS_SEM_R1, pruning, g_devices);
#endif
}

static Config base_conf() {
  return Config{{"cpslab.allpair.similarityThreshold", "0.5"}, {"cpslab.allpair.outputIODuration", "0"},
                {"cpslab.allpair.benchmark.expDuration", "0"}, {"cpslab.allpair.vectorDim", "64"}, {"cpslab.allpair.indexThreshold", "0.0"}};
}

static void scenario(int pruning) {
  Config conf = base_conf();
  Engine* eng = make_engine(64, 0.5, 0.0, false, pruning);
  std::vector<OutMessage> out;
  GpuIndexingWorkerActor<Engine> w(conf, *eng, [&](const OutMessage& m) { out.push_back(m); });
  w.receive(VectorIOMsg{{{"a", V({{0, .6}, {1, .8}})}}});
  w.receive(VectorIOMsg{{{"b", V({{1, .8}, {2, .6}})}, {"c", V({{0, .6}, {1, .8}})}, {"tiny", V({{5, .1}})}}});
  CHECK(out.size() == 2);
  const auto& o0 = std::get<SimilarityOutput>(out[0]).output;
  CHECK(o0.size() == 1 && o0.count("a") && o0.at("a").empty());
  const auto& o1 = std::get<SimilarityOutput>(out[1]).output;
  CHECK(o1.size() == 2 && o1.count("b") && o1.count("c"));              // "tiny" fails the admission filter: no entry
  CHECK(o1.at("c").at("a") == .6 * .6 + .8 * .8 && o1.at("b").at("a") == .8 * .8);
  CHECK(o1.at("b").at("c") == .8 * .8 && o1.at("c").at("b") == .8 * .8);  // in-batch pairs, both orders (IWA:125-132)
  CHECK(std::get<SimilarityOutput>(out[0]).toString() == "---------------------------------a:\n");
  CHECK(std::get<SimilarityOutput>(out[1]).toString() ==
        "---------------------------------b:a,0.6400000000000001;c,0.6400000000000001;\n---------------------------------c:a,1.0;b,0.6400000000000001;\n");
  // same external id is never paired with itself (IWA:91)
  w.receive(VectorIOMsg{{{"a", V({{0, .6}, {1, .8}})}}});
  const auto& o2 = std::get<SimilarityOutput>(out[2]).output;
  CHECK(o2.at("a").count("a") == 0 && o2.at("a").size() == 2 && o2.at("a").count("b") && o2.at("a").count("c"));
  // Test echo (IWA:145-147) and ReceiveTimeout freeze (IWA:143-144)
  w.receive(Test{"ping"});
  CHECK(std::get<Test>(out[3]) == Test{"ping"});
  w.receive(ReceiveTimeout{});
  w.receive(VectorIOMsg{{{"z1", V({{0, .6}, {1, .8}})}, {"z2", V({{0, .6}, {1, .8}})}}});
  const auto& o4 = std::get<SimilarityOutput>(out[4]).output;
  CHECK(o4.at("z1").count("z2") == 0 && o4.at("z1").count("a") == 1);
  // a batch with a wrong-size vector is dropped whole, like the swallowed exception at IWA:135-137
  const size_t n_before = out.size();
  w.receive(VectorIOMsg{{{"bad", SparkSparseVector(32, {1}, {1.0})}}});
  CHECK(out.size() == n_before);
  delete eng;
}

static void buffered_output_and_router() {
  Config conf = base_conf();
  conf["cpslab.allpair.outputIODuration"] = "50";
  Engine* eng = make_engine(64, 0.5, 0.0, false, 0);
  std::vector<OutMessage> out;
  GpuIndexingWorkerActor<Engine> w(conf, *eng, [&](const OutMessage& m) { out.push_back(m); });
  w.receive(VectorIOMsg{{{"a", V({{0, 1.0}})}}});
  w.receive(VectorIOMsg{{{"b", V({{0, 1.0}})}}});
  CHECK(out.empty());                                   // buffered until the IOTicket (IWA:131-132)
  w.receive(IOTicket{});
  CHECK(out.size() == 1);
  const auto& o = std::get<SimilarityOutput>(out[0]).output;
  CHECK(o.size() == 1 && o.at("b").at("a") == 1.0);     // only non-empty results are buffered (IWA:115-119)
  w.receive(IOTicket{});
  CHECK(out.size() == 1);                               // cleared after the send (IWA:141)
  delete eng;

  // client -> router -> worker, timer-driven batching (WWA:164-183)
  Config c2 = base_conf();
  c2["cpslab.allpair.ioTriggerPeriod"] = "10";
  Engine* e2 = make_engine(64, 0.5, 0.0, false, 0);
  std::vector<OutMessage> out2;
  GpuIndexingWorkerActor<Engine> w2(c2, *e2, [&](const OutMessage& m) { out2.push_back(m); });
  RegionRouter<Engine> router(c2, w2);
  LocalActorSystem<Engine> sys;
  sys.registerRouter("127.0.0.1:2551", router);
  ClientConnection<Engine> client({"127.0.0.1:2551"}, sys);
  client.insertNewVector(std::vector<IdVector>{{"a", V({{0, 1.0}})}});
  client.insertNewVector(std::vector<IdVector>{{"b", V({{0, 1.0}})}});
  CHECK(out2.empty());
  router.tell(IOTrigger{});
  CHECK(out2.size() == 1);
  const auto& ob = std::get<SimilarityOutput>(out2[0]).output;
  CHECK(ob.at("a").at("b") == 1.0 && ob.at("b").at("a") == 1.0);     // one batch: both orders
  // README form: vectors without ids
  client.insertNewVector(std::vector<SparkSparseVector>{V({{0, 1.0}})});
  router.tell(IOTrigger{});
  CHECK(out2.size() == 2 && std::get<SimilarityOutput>(out2[1]).output.count("auto-0") == 1);
  delete e2;
}

static void as_built_first_list_skip() {
  // KAT-B: batch1 a={0:.6,1:.8}, batch2 b={1:.8,2:.6}: the only shared dim is b's first -> nothing as built
  Config conf = base_conf();
  conf["cpslab.allpair.gpu.semantics"] = "R0";
  Engine* eng = make_engine(64, 0.5, 0.0, true, 0);
  std::vector<OutMessage> out;
  GpuIndexingWorkerActor<Engine> w(conf, *eng, [&](const OutMessage& m) { out.push_back(m); });
  w.receive(VectorIOMsg{{{"a", V({{0, .6}, {1, .8}})}}});
  w.receive(VectorIOMsg{{{"b", V({{1, .8}, {2, .6}})}}});
  CHECK(std::get<SimilarityOutput>(out[1]).output.at("b").empty());
  delete eng;
  Engine* e2 = make_engine(64, 0.5, 0.0, true, 0);
  std::vector<OutMessage> o2;
  GpuIndexingWorkerActor<Engine> w2(conf, *e2, [&](const OutMessage& m) { o2.push_back(m); });
  w2.receive(VectorIOMsg{{{"b", V({{1, .8}, {2, .6}})}}});
  w2.receive(VectorIOMsg{{{"a", V({{0, .6}, {1, .8}})}}});      // reversed arrival: dim 1 is not a's first
  CHECK(std::get<SimilarityOutput>(o2[1]).output.at("a").at("b") == .8 * .8);
  delete e2;
}

// IndexData / DataPacket (Message.scala:16-18): wrappers carry admitted, pruned vectors -- no second admission filter --
// and, as built, the skipped first posting list is the first element of the WRAPPER's Set (IWA:102)
static void index_data_and_data_packet(int pruning) {
  Config conf = base_conf();
  Engine* eng = make_engine(64, 0.5, 0.0, false, pruning);
  std::vector<OutMessage> out;
  GpuIndexingWorkerActor<Engine> w(conf, *eng, [&](const OutMessage& m) { out.push_back(m); });
  RegionRouter<Engine> router(conf, w);
  auto wrap = [](const char* id, SparkSparseVector v) { std::set<int32_t> d(v.indices.begin(), v.indices.end()); return SparseVectorWrapper{d, IdVector{id, v}}; };
  w.receive(IndexData{{wrap("p", V({{3, .3}}))}});                          // sum 0.3 < t: a VectorIOMsg would reject it (EPA:81-93)
  CHECK(out.size() == 1 && std::get<SimilarityOutput>(out[0]).output.count("p") == 1);
  router.tell(DataPacket{0, {wrap("q", V({{3, .9}, {4, .1}})), wrap("r", V({{3, 1.0}}))}});   // EPA:113-122 -> ONE IndexData
  CHECK(out.size() == 2);
  const auto& o = std::get<SimilarityOutput>(out[1]).output;
  CHECK(o.size() == 2 && o.at("r").count("q") == 1 && o.at("q").at("r") == .9 && o.at("q").count("p") == 0);   // .9 * .3 < t
  // a refused batch leaves no trace in the id tables: "dup" is new again afterwards and pairs with nothing of its own name
  w.receive(VectorIOMsg{{{"dup", SparkSparseVector(32, {1}, {1.0})}}});     // wrong size: dropped whole (IWA:135-137)
  w.receive(VectorIOMsg{{{"dup", V({{3, 1.0}})}}});
  CHECK(out.size() == 3 && std::get<SimilarityOutput>(out[2]).output.at("dup").count("r") == 1);
  delete eng;
  // as built: first(q) comes from the wrapper's Set
  Config c0 = base_conf(); c0["cpslab.allpair.gpu.semantics"] = "R0";
  Engine* e0 = make_engine(64, 0.5, 0.0, true, pruning);
  std::vector<OutMessage> o0;
  GpuIndexingWorkerActor<Engine> w0(c0, *e0, [&](const OutMessage& m) { o0.push_back(m); });
  w0.receive(IndexData{{wrap("a", V({{0, .6}, {1, .8}}))}});
  w0.receive(IndexData{{wrap("b", V({{1, .8}, {2, .6}}))}});               // only shared dim is b's first: dropped as built
  CHECK(std::get<SimilarityOutput>(o0[1]).output.at("b").empty());
  delete e0;
}

static void formats() {
  CHECK(java_double_to_string(1.0) == "1.0" && java_double_to_string(0.64) == "0.64" && java_double_to_string(1.25e-5) == "1.25E-5");
  CHECK(java_double_to_string(1e7) == "1.0E7" && java_double_to_string(123456.789) == "123456.789" && java_double_to_string(0.001) == "0.001");
  CHECK(V({{17, 0.25}, {3, 0.5}}, 1 << 20).toString() == "(1048576,[3,17],[0.5,0.25])");
  bool threw = false;
  try { SparkSparseVector(8, {3, 3}, {1.0, 1.0}); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  CHECK(scala_set_first({3, 17, 100, 1000, 65537, 200000}) == 200000 && scala_set_first({5, 10, 15, 20, 25, 30, 35}) == 5);   // SURVEY 8(a) examples
  threw = false;
  try { Config c; conf_required(c, "cpslab.allpair.vectorDim"); } catch (const std::out_of_range&) { threw = true; }
  CHECK(threw);
}

// The reference's latency experiment (benchmark/LoadGenerator.scala:15-173) on a virtual clock: two runners, warm-up,
// StartTest after the parent's ReceiveTimeout, every runner restarting its ids at 1, response time = outputMoment - StartTime
// (immediate output: 0 ms; output buffered for 25 ms: the wait for the next flush).  Same checks as tests/test_loadgen.py.
static void loadgen_protocol(int pruning, long long output_io_ms) {
  const int D = 32, total = 4, kids = 2, nvid = 6;
  Config conf{{"cpslab.allpair.similarityThreshold", "0.3"}, {"cpslab.allpair.outputIODuration", std::to_string(output_io_ms)},
              {"cpslab.allpair.vectorDim", std::to_string(D)}, {"cpslab.allpair.indexThreshold", "0.0"},
              {"cpslab.allpair.benchmark.expDuration", "500"}, {"cpslab.allpair.benchmark.writeBatchingDuration", "10"},
              {"cpslab.allpair.benchmark.totalMessageCount", std::to_string(total)}, {"cpslab.allpair.benchmark.childrenNum", std::to_string(kids)}};
  std::vector<IdVector> videos;
  for (int i = 0; i < nvid; ++i)       // three shared dimensions: every pair of videos is similar; NOT normalised (the runner does it)
    videos.push_back({"v" + std::to_string(i), SparkSparseVector(D, {0, 1, 2, 3 + i % 3, 8 + i}, {1.0 + 0.1 * i, 1.5, 0.7 + 0.05 * i, 1.1, 0.9})});
  Engine* eng = make_engine(D, 0.3, 0.0, false, pruning);
  GpuIndexingWorkerActor<Engine> w(conf, *eng, nullptr);
  EventLoop loop(true, 1000000);
  std::vector<std::pair<int64_t, std::string>> seen; std::vector<std::string> lines;
  const auto rep = run_experiment(conf, videos, w, loop, true, [&](const std::string& ln) { lines.push_back(ln); },
                                  [&](int64_t at, const VectorIOMsg& m) { seen.push_back({at, m.vectors[0].first}); });
  // warm-up: runner i counts from i * total (LG:22) until msgCount > videos.size (LG:63-66), one vector per 10 ms tick
  std::vector<int> warm_ids, test_ids; std::vector<int64_t> test_at;
  for (const auto& [at, id] : seen) { if (at < 1000500) warm_ids.push_back(std::stoi(id)); else { test_ids.push_back(std::stoi(id)); test_at.push_back(at); } }
  std::vector<int> want_warm;
  for (int i = 1; i <= nvid + 1; ++i) want_warm.push_back(i);
  for (int i = total + 1; i <= nvid + 1; ++i) want_warm.push_back(i);
  std::sort(warm_ids.begin(), warm_ids.end()); std::sort(want_warm.begin(), want_warm.end());
  CHECK(warm_ids == want_warm);
  CHECK(seen.size() >= 4 && seen[0].first == 1000000 && seen[1].first == 1000000 && seen[2].first == 1000010 && seen[3].first == 1000010);
  // test phase: every runner restarts at 1 (LG:79) -- the same ids from all children, as built
  std::vector<int> want_test;
  for (int k = 0; k < kids; ++k) for (int i = 1; i <= total + 1; ++i) want_test.push_back(i);
  std::sort(test_ids.begin(), test_ids.end()); std::sort(want_test.begin(), want_test.end());
  CHECK(test_ids == want_test);
  CHECK(!test_at.empty() && test_at[0] >= 1000500 && test_at.size() >= 4 && test_at[1] == test_at[0] && test_at[2] == test_at[0] + 10);
  CHECK(w.stopUpdateIndex);                                               // the worker's own ReceiveTimeout froze the index (IWA:143-144)
  CHECK(rep.messages >= 1 && rep.with_both == rep.messages && rep.min >= 0 && rep.min <= rep.average && rep.average <= rep.max);
  if (output_io_ms <= 0) CHECK(rep.max == 0);                             // answered inside the tick
  else CHECK(rep.max > 0 && rep.max <= output_io_ms);                     // the wait for the worker's next IOTicket
  CHECK(rep.line.rfind("LoadGenerator stopped with " + std::to_string(rep.messages) + " messages, average response time", 0) == 0);
  CHECK(!lines.empty() && lines[0].find(" lasting Time:") != std::string::npos);
  delete eng;
}

int main(int argc, char** argv) {
  const int pruning = argc > 1 ? std::atoi(argv[1]) : 0;
  for (int a = 2; a < argc; ++a) g_devices.push_back(std::atoi(argv[a]));      // e.g. "0 1 2 3": device_ids of the one handle
  formats();
  scenario(pruning);
  buffered_output_and_router();
  as_built_first_list_skip();
  index_data_and_data_packet(pruning);
  loadgen_protocol(pruning, 0);
  loadgen_protocol(pruning, 25);
  if (failures) { std::fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
  std::printf("actor scenario ok (pruning=%d, devices=%zu)\n", pruning, g_devices.empty() ? (size_t)1 : g_devices.size());
  return 0;
}
