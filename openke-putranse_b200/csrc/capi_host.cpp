// Host half of the C-ABI: the Base.so-compatible process-global surface and the pk_* exports of
// graph state.  See include/putranse.h for the map to the reference's symbols.
#include <atomic>
#include <cmath>
#include <cstring>
#include <ctime>
#include <thread>

#include "common.hpp"
#include "graph_host.hpp"
#include "global_state.hpp"
#include <mutex>
#include <shared_mutex>

namespace pk {

std::string& last_error() {
    thread_local std::string e;
    return e;
}
int& launch_counter() {
    thread_local int n = 0;
    return n;
}

Global& G() {
    static Global g;
    return g;
}

const TripleIndex& current_index() {
    Global& g = G();
    return g.swapped ? g.universe.local : g.graph.train;
}

uint64_t index_epoch() { return G().index_epoch; }

}  // namespace pk

using pk::G;

// pk_universes_build runs on caller threads beside the thread that owns the loaders (the orchestrator samples the
// next chunk's subgraphs in the background); it holds this lock shared, and whatever replaces the training graph
// (importTrainFiles, evolveTrainList, the incremental reset) takes it exclusively.
static std::shared_mutex& graph_mutex() {
    static std::shared_mutex m;
    return m;
}

extern "C" {

int pk_last_error(char* buf, int n) {
    const std::string& e = pk::last_error();
    if (buf && n > 0) {
        const int m = (int)std::min<size_t>(e.size(), (size_t)n - 1);
        std::memcpy(buf, e.data(), (size_t)m);
        buf[m] = 0;
    }
    return (int)e.size();
}

const char* pk_version(void) { return "putranse-b200 0.1 (sm_100a)"; }

int pk_last_launch_count(void) { return pk::launch_counter(); }

// ---------------------------------------------------------------- settings
void setInPath(char* path) { G().graph.in_path = path ? path : ""; }
void setOutPath(char*) {}
void setWorkThreads(PK_INT threads) { G().graph.work_threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads); }
PK_INT getWorkThreads(void) { return G().graph.work_threads; }
void setBern(PK_INT con) { G().graph.bern = con; }

PK_INT getEntityTotal(void) { return G().swapped ? G().universe.local.n_ent : G().graph.n_ent; }
PK_INT getRelationTotal(void) { return G().swapped ? G().universe.local.n_rel : G().graph.n_rel; }
PK_INT getTrainTotal(void) { return pk::current_index().n_tri(); }
PK_INT getTestTotal(void) { return (PK_INT)G().graph.test.size(); }
PK_INT getValidTotal(void) { return (PK_INT)G().graph.valid.size(); }
PK_INT getTripleTotal(void) { return (PK_INT)G().graph.all_hrt.size(); }

void setRandomSeed(PK_INT seed) {
    if (seed == -1) seed = (PK_INT)time(nullptr);
    G().seed = seed;
    G().rng.reseed((uint32_t)seed);
}
PK_INT getRandomSeed(void) { return G().seed; }
void randReset(void) {
    for (int64_t i = 0; i < G().graph.work_threads; ++i) G().lcg[i] = (uint64_t)(int64_t)G().rng.next();
}

void importTrainFiles(void) {
    std::unique_lock<std::shared_mutex> lk(graph_mutex());
    std::string err;
    if (!G().graph.import_train(&err)) {
        pk::fail(PK_ERR_IO, err);
        fprintf(stderr, "putranse: importTrainFiles failed: %s\n", err.c_str());
        return;
    }
    G().index_epoch++;
}

void importTestFiles(void) {
    std::string err;
    if (!G().graph.import_test(&err)) {
        pk::fail(PK_ERR_IO, err);
        fprintf(stderr, "putranse: importTestFiles failed: %s\n", err.c_str());
    }
    G().eval_epoch++;
}

int pk_import_count(void) { return (int)G().graph.import_count; }

// ---------------------------------------------------------------- universes (global, reference order of calls)
void getParallelUniverse(PK_INT tc, PK_REAL balance) {
    std::string err;
    pk::Universe& u = G().universe;
    u = pk::Universe();
    if (G().swapped) {
        pk::fail(PK_ERR_STATE, "getParallelUniverse: helpers are swapped; call resetUniverse first");
        return;
    }
    G().have_universe = G().graph.walk_universe(G().rng, tc, balance, &u, &err);
    if (!G().have_universe) {
        pk::fail(PK_ERR_STATE, err);
        fprintf(stderr, "putranse: getParallelUniverse failed: %s\n", err.c_str());
    }
    std::memcpy(u.lcg, G().lcg, sizeof u.lcg);
}
PK_INT getEntityTotalUniverse(void) { return G().swapped ? G().graph.n_ent : (G().have_universe ? G().universe.local.n_ent : 0); }
PK_INT getRelationTotalUniverse(void) { return G().swapped ? G().graph.n_rel : (G().have_universe ? G().universe.local.n_rel : 0); }
PK_INT getTrainTotalUniverse(void) { return G().swapped ? G().graph.train.n_tri() : (G().have_universe ? G().universe.local.n_tri() : 0); }
void getEntityRemapping(PK_INT* out) {
    if (!G().have_universe) return;
    for (size_t i = 0; i < G().universe.ent_remap.size(); ++i) out[i] = G().universe.ent_remap[i];
}
void getRelationRemapping(PK_INT* out) {
    if (!G().have_universe) return;
    for (size_t i = 0; i < G().universe.rel_remap.size(); ++i) out[i] = G().universe.rel_remap[i];
}
void swapHelpers(void) {
    if (!G().have_universe) {
        pk::fail(PK_ERR_STATE, "swapHelpers: no universe has been built");
        return;
    }
    G().swapped = !G().swapped;
    G().index_epoch++;
}
void resetUniverse(void) {
    if (G().swapped) swapHelpers();
    G().universe = pk::Universe();
    G().have_universe = false;
}

int pk_universe_triples(int32_t* out) {
    if (!G().have_universe) return pk::fail(PK_ERR_STATE, "pk_universe_triples: no universe");
    std::memcpy(out, G().universe.collected.data(), G().universe.collected.size() * sizeof(pk::Tri));
    return PK_OK;
}

int pk_train_index(int32_t* by_head, int32_t* by_tail, float* left_mean, float* right_mean) {
    const pk::TripleIndex& ix = pk::current_index();
    if (ix.n_tri() == 0) return pk::fail(PK_ERR_STATE, "pk_train_index: nothing imported");
    if (by_head) std::memcpy(by_head, ix.by_head.data(), ix.by_head.size() * sizeof(pk::Tri));
    if (by_tail) std::memcpy(by_tail, ix.by_tail.data(), ix.by_tail.size() * sizeof(pk::Tri));
    if (left_mean) std::memcpy(left_mean, ix.left_mean.data(), ix.left_mean.size() * sizeof(float));
    if (right_mean) std::memcpy(right_mean, ix.right_mean.data(), ix.right_mean.size() * sizeof(float));
    return PK_OK;
}

// n = capacity of the caller's buffer in elements: work_threads is process-global (the LAST loader's
// setWorkThreads), so the caller says how much it holds
int pk_get_lcg(uint64_t* s, int n) {
    if (!s || n < 0) return pk::fail(PK_ERR_ARG, "pk_get_lcg: null argument");
    std::memcpy(s, G().lcg, (size_t)std::min<int64_t>(n, std::min<int64_t>(G().graph.work_threads, 64)) * 8);
    return PK_OK;
}
int pk_set_lcg(const uint64_t* s, int n) {
    if (!s || n < 0) return pk::fail(PK_ERR_ARG, "pk_set_lcg: null argument");
    std::memcpy(G().lcg, s, (size_t)std::min<int64_t>(n, std::min<int64_t>(G().graph.work_threads, 64)) * 8);
    return PK_OK;
}

// acc = (float)(acc + v) over v in order: the reference's `float += double` metric accumulators
// (openke/base/Test.h:14-16,213-223), for the Python-side metric table
float pk_f32_running_sum(const double* v, int64_t n) {
    float acc = 0.f;
    for (int64_t i = 0; i < n; ++i) acc = (float)((double)acc + v[i]);
    return acc;
}

// ---------------------------------------------------------------- evaluation lists
int pk_eval_triples(int which, int32_t* hrt) {
    const std::vector<pk::Tri>& q = which == 0 ? G().graph.test : G().graph.valid;
    if (q.empty()) return pk::fail(PK_ERR_STATE, "pk_eval_triples: importTestFiles has not run");
    std::memcpy(hrt, q.data(), q.size() * sizeof(pk::Tri));
    return PK_OK;
}

int pk_filter_csr(int which, int side, int64_t* offsets, int32_t* cand, int64_t* n_cand) {
    if (G().graph.all_hrt.empty()) return pk::fail(PK_ERR_STATE, "pk_filter_csr: importTestFiles has not run");
    if (which < 0 || which > 1 || side < 0 || side > 1) return pk::fail(PK_ERR_ARG, "pk_filter_csr: bad selector");
    std::vector<int64_t> off;
    std::vector<int32_t> c;
    G().graph.filter_candidates(which, side, off, c);
    if (n_cand) *n_cand = (int64_t)c.size();
    if (offsets) std::memcpy(offsets, off.data(), off.size() * 8);
    if (cand && !c.empty()) std::memcpy(cand, c.data(), c.size() * 4);
    return PK_OK;
}

// candidate batches in the reference's order: slot 0 = the true triple, then every other entity
// ascending (Test.h:61-68, Valid.h:58-67)
static void fill_batch(const pk::Tri& x, bool head, PK_INT* ph, PK_INT* pt, PK_INT* pr) {
    const int64_t E = G().graph.n_ent;
    ph[0] = x.h; pt[0] = x.t; pr[0] = x.r;
    const int64_t truth = head ? x.h : x.t;
    for (int64_t i = 1; i < E; ++i) {
        const int64_t e = i - 1 < truth ? i - 1 : i;
        ph[i] = head ? e : x.h;
        pt[i] = head ? x.t : e;
        pr[i] = x.r;
    }
}
void initTest(void) { G().last_head = G().last_tail = 0; pk::test_metrics_reset(); }
void validInit(void) { G().last_valid_head = G().last_valid_tail = 0; pk::valid_metrics_reset(); }
void getHeadBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr) { fill_batch(G().graph.test[(size_t)G().last_head++], true, ph, pt, pr); }
void getTailBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr) { fill_batch(G().graph.test[(size_t)G().last_tail++], false, ph, pt, pr); }
void getValidHeadBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr) { fill_batch(G().graph.valid[(size_t)G().last_valid_head++], true, ph, pt, pr); }
void getValidTailBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr) { fill_batch(G().graph.valid[(size_t)G().last_valid_tail++], false, ph, pt, pr); }

// ---------------------------------------------------------------- adjacency queries over the current id space
// Reference openke/base/Base.cpp:312-468: scans of one entity's run in the (h,r,t) resp. (t,r,h) order
// of whichever index the sampler currently reads (global graph, or the universe after swapHelpers).
// The reference's Python loaders bind all six unconditionally (openke/data/TrainDataLoader.py:60-103).
namespace {
struct EntRun {
    const pk::Tri* rec = nullptr;
    int64_t n = 0;
    bool tail = false;
    int32_t other(int64_t i) const { return tail ? rec[i].h : rec[i].t; }
};
EntRun ent_run(PK_INT entity, bool entity_is_tail) {
    EntRun run;
    const pk::TripleIndex& ix = pk::current_index();
    if (entity < 0 || entity >= ix.n_ent || ix.rig_head.empty()) return run;
    const int64_t lef = entity_is_tail ? ix.lef_tail[(size_t)entity] : ix.lef_head[(size_t)entity];
    const int64_t rig = entity_is_tail ? ix.rig_tail[(size_t)entity] : ix.rig_head[(size_t)entity];
    if (rig < lef) return run;
    run.rec = (entity_is_tail ? ix.by_tail.data() : ix.by_head.data()) + lef;
    run.n = rig - lef + 1;
    run.tail = entity_is_tail;
    return run;
}
}  // namespace

PK_INT getNumOfNegatives(PK_INT entity, PK_INT relation, bool entity_is_tail) {   // Base.cpp:313-335
    const EntRun run = ent_run(entity, entity_is_tail);
    PK_INT n = 0;
    for (int64_t i = 0; i < run.n; ++i) n += run.rec[i].r != relation;
    return n;
}
PK_INT getNumOfPositives(PK_INT entity, PK_INT relation, bool entity_is_tail) {   // Base.cpp:337-359
    const EntRun run = ent_run(entity, entity_is_tail);
    PK_INT n = 0;
    for (int64_t i = 0; i < run.n; ++i) n += run.rec[i].r == relation;
    return n;
}
void getNegativeEntities(PK_INT* out, PK_INT entity, PK_INT relation, bool entity_is_tail) {   // Base.cpp:361-385
    const EntRun run = ent_run(entity, entity_is_tail);
    for (int64_t i = 0, o = 0; i < run.n; ++i)
        if (run.rec[i].r != relation) out[o++] = run.other(i);
}
void getPositiveEntities(PK_INT* out, PK_INT entity, PK_INT relation, bool entity_is_tail) {   // Base.cpp:387-411
    const EntRun run = ent_run(entity, entity_is_tail);
    for (int64_t i = 0, o = 0; i < run.n; ++i)
        if (run.rec[i].r == relation) out[o++] = run.other(i);
}
PK_INT getNumOfEntityRelations(PK_INT entity, bool entity_is_tail) {   // Base.cpp:413-439
    const EntRun run = ent_run(entity, entity_is_tail);
    PK_INT n = 0;
    for (int64_t i = 0; i < run.n; ++i) n += (i == 0 || run.rec[i].r != run.rec[i - 1].r);
    return n;
}
// Base.cpp:441-468 never advances its output cursor, so it leaves only the LAST distinct relation in
// slot 0.  The distinct relations are written in order here (slot 0 of the reference is then slot n-1);
// nothing in the reference's Python calls it.
void getEntityRelations(PK_INT* out, PK_INT entity, bool entity_is_tail) {
    const EntRun run = ent_run(entity, entity_is_tail);
    for (int64_t i = 0, o = 0; i < run.n; ++i)
        if (i == 0 || run.rec[i].r != run.rec[i - 1].r) out[o++] = run.rec[i].r;
}

// Reference openke/base/Reader.h:240-244 (the argument is ignored there as well): the next
// importTestFiles() reads the filter set from triple2id.txt instead of test ∪ train ∪ valid.
void activateLoadOfAllTriples(bool) { G().graph.load_all_triples = true; }

// ---------------------------------------------------------------- incremental setting (openke/base/Incremental.h)
// Bound by the reference's IncrementalTrainDataLoader / IncrementalTestDataLoader
// (openke/data/IncrementalTrainDataLoader.py:40-60, IncrementalTestDataLoader.py:34-66).
static void report(bool ok, const char* who, const std::string& err) {
    if (ok) return;
    pk::fail(PK_ERR_IO, err);
    fprintf(stderr, "putranse: %s failed: %s\n", who, err.c_str());
}
void activateIncrementalSetting(void) { G().graph.incremental = true; }          // Incremental.h:45-48
void initializeIncrementalSetting(void) {                                         // :207-217
    std::string err;
    G().graph.incremental = true;
    report(G().graph.read_global_totals(&err), "initializeIncrementalSetting", err);
}
void setNumSnapshots(PK_INT n) { G().graph.num_snapshots = n; }
PK_INT getNumSnapshots(void) { return G().graph.num_snapshots; }
void setNumOperationsRate(PK_INT n) { G().graph.ops_rate = n; }
void readGlobalNumEntities(void) { std::unique_lock<std::shared_mutex> lk(graph_mutex()); std::string err; report(G().graph.read_global_totals(&err), "readGlobalNumEntities", err); }
void readGlobalNumRelations(void) { std::unique_lock<std::shared_mutex> lk(graph_mutex()); std::string err; report(G().graph.read_global_totals(&err), "readGlobalNumRelations", err); }
void initializeTrainingOperations(int snapshot) {
    std::string err;
    report(G().graph.load_train_ops(snapshot, &err), "initializeTrainingOperations", err);
}
void evolveTrainList(void) {
    std::unique_lock<std::shared_mutex> lk(graph_mutex());
    std::string err;
    const bool ok = G().graph.evolve_train(&err);
    report(ok, "evolveTrainList", err);
    if (ok) G().index_epoch++;
}
void loadSnapshotTriples(int snapshot) {
    std::string err;
    const bool ok = G().graph.load_snapshot_triples(snapshot, &err);
    report(ok, "loadSnapshotTriples", err);
    if (ok) G().eval_epoch++;
}
void loadTestData(int snapshot) {
    std::string err;
    const bool ok = G().graph.load_snapshot_eval(snapshot, 0, &err);
    report(ok, "loadTestData", err);
    if (ok) G().eval_epoch++;
}
void loadValidData(int snapshot) {
    std::string err;
    const bool ok = G().graph.load_snapshot_eval(snapshot, 1, &err);
    report(ok, "loadValidData", err);
    if (ok) G().eval_epoch++;
}
// The triple-operation replay of the FILTER list (Incremental.h:323-348,926-949) was superseded in the reference by
// loadSnapshotTriples (experiments/incremental_experiment_PuTransE_on_WikidataEvolve.py:45-46).  The reference's
// IncrementalTestDataLoader still binds the symbol at construction (IncrementalTestDataLoader.py:38), so both exist
// and refuse at call time.
void initializeTripleOperations(int) {
    pk::fail(PK_ERR_UNSUPPORTED, "initializeTripleOperations: use loadSnapshotTriples (evolveTripleList2)");
    fprintf(stderr, "putranse: %s\n", pk::last_error().c_str());
}
void evolveTripleList(void) {
    pk::fail(PK_ERR_UNSUPPORTED, "evolveTripleList: use loadSnapshotTriples (evolveTripleList2)");
    fprintf(stderr, "putranse: %s\n", pk::last_error().c_str());
}
PK_INT getNumCurrentlyContainedEntities(void) { return (PK_INT)G().graph.contained_entities.size(); }
// leave the incremental setting (the reference cannot: its flag is set once per process)
int pk_incremental_reset(void) {
    std::unique_lock<std::shared_mutex> lk(graph_mutex());
    pk::Graph& g = G().graph;
    g.incremental = false;
    g.ops.clear(); g.next_op = 0; g.ops_rate = 0;
    g.train_count.clear(); g.train_rel_count.clear();
    g.train_rel_contained.clear(); g.train_rel_all.clear(); g.train_rel_deleted.clear();
    g.contained_entities.clear(); g.contained_relations.clear();
    return PK_OK;
}
// which: 0 currently contained train relations (the universe focus is drawn by position from this array),
// 1 all, 2 deleted, 3 entities of the snapshot's triple list, 4 its relations.  out = NULL: count only.
int64_t pk_incremental_list(int which, int32_t* out) {
    const pk::Graph& g = G().graph;
    const std::vector<int32_t>* v = which == 0 ? &g.train_rel_contained : which == 1 ? &g.train_rel_all : which == 2 ? &g.train_rel_deleted
                                    : which == 3 ? &g.contained_entities : which == 4 ? &g.contained_relations : nullptr;
    if (!v) return pk::fail(PK_ERR_ARG, "pk_incremental_list: bad selector");
    if (out && !v->empty()) std::memcpy(out, v->data(), v->size() * 4);
    return (int64_t)v->size();
}

// ---------------------------------------------------------------- many universes, many threads
struct pk_universe_set {
    std::vector<pk::Universe> u;
    int work_threads = 1;
};

static pk_universe_set* universes_build(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, int nthreads, bool helpers);
pk_universe_set* pk_universes_build(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, int nthreads) {
    return universes_build(n, seeds, tcs, balances, nthreads, true);
}
// the same without the (t,r,h) order, per-entity ranges and Bernoulli means of every universe: what training with
// filter_flag = 0 and bern_flag = 0 (every PuTrans* script) reads is the (h,r,t) list and the remaps
pk_universe_set* pk_universes_build_lean(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, int nthreads) {
    return universes_build(n, seeds, tcs, balances, nthreads, false);
}
static pk_universe_set* universes_build(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, int nthreads, bool helpers) {
    if (n < 0 || !seeds || !tcs || !balances) {
        pk::fail(PK_ERR_ARG, "pk_universes_build: null argument");
        return nullptr;
    }
    if (G().graph.train.n_tri() == 0) {
        pk::fail(PK_ERR_STATE, "pk_universes_build: importTrainFiles has not run");
        return nullptr;
    }
    std::shared_lock<std::shared_mutex> graph_lk(graph_mutex());
    pk_universe_set* s = new pk_universe_set();
    s->u.resize((size_t)n);
    s->work_threads = (int)G().graph.work_threads;
    if (nthreads < 1) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min(nthreads, std::max(n, 1));
    std::atomic<int> next{0};
    std::atomic<bool> failed{false};
    std::string first_err;
    std::mutex* mu = new std::mutex();
    const pk::Graph& g = G().graph;
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            std::string err;
            if (!g.build_universe(seeds[i], tcs[i], balances[i], &s->u[(size_t)i], &err, helpers)) {
                std::lock_guard<std::mutex> lk(*mu);
                if (!failed.exchange(true)) first_err = "universe " + std::to_string(i) + ": " + err;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    delete mu;
    if (failed) {
        pk::fail(PK_ERR_STATE, first_err);
        delete s;
        return nullptr;
    }
    return s;
}

void pk_universes_free(pk_universe_set* s) { delete s; }
int pk_universes_count(const pk_universe_set* s) { return s ? (int)s->u.size() : 0; }

int pk_universes_sizes(const pk_universe_set* s, int64_t* n_tri, int64_t* n_ent, int64_t* n_rel, int64_t* focus) {
    if (!s) return pk::fail(PK_ERR_ARG, "pk_universes_sizes: null set");
    for (size_t i = 0; i < s->u.size(); ++i) {
        if (n_tri) n_tri[i] = s->u[i].local.n_tri();
        if (n_ent) n_ent[i] = s->u[i].local.n_ent;
        if (n_rel) n_rel[i] = s->u[i].local.n_rel;
        if (focus) focus[i] = s->u[i].focus;
    }
    return PK_OK;
}

int pk_universes_export(const pk_universe_set* s, int32_t* tri_by_head, int32_t* tri_by_tail, int32_t* tri_collected,
                        int32_t* ent_remap, int32_t* rel_remap, float* left_mean, float* right_mean, uint64_t* lcg) {
    if (!s) return pk::fail(PK_ERR_ARG, "pk_universes_export: null set");
    size_t to = 0, eo = 0, ro = 0;
    for (size_t i = 0; i < s->u.size(); ++i)
        if (!s->u[i].has_helpers && (tri_by_tail || left_mean || right_mean))
            return pk::fail(PK_ERR_STATE, "pk_universes_export: the set was built lean (no (t,r,h) order / Bernoulli means)");
    for (size_t i = 0; i < s->u.size(); ++i) {
        const pk::Universe& u = s->u[i];
        const size_t nt = (size_t)u.local.n_tri(), ne = (size_t)u.local.n_ent, nr = (size_t)u.local.n_rel;
        if (tri_by_head) std::memcpy(tri_by_head + to * 3, u.local.by_head.data(), nt * sizeof(pk::Tri));
        if (tri_by_tail) std::memcpy(tri_by_tail + to * 3, u.local.by_tail.data(), nt * sizeof(pk::Tri));
        if (tri_collected) std::memcpy(tri_collected + to * 3, u.collected.data(), nt * sizeof(pk::Tri));
        if (ent_remap) std::memcpy(ent_remap + eo, u.ent_remap.data(), ne * 4);
        if (rel_remap) std::memcpy(rel_remap + ro, u.rel_remap.data(), nr * 4);
        if (left_mean) std::memcpy(left_mean + ro, u.local.left_mean.data(), nr * 4);
        if (right_mean) std::memcpy(right_mean + ro, u.local.right_mean.data(), nr * 4);
        if (lcg) std::memcpy(lcg + i * (size_t)s->work_threads, u.lcg, (size_t)s->work_threads * 8);
        to += nt; eo += ne; ro += nr;
    }
    return PK_OK;
}

}  // extern "C"
