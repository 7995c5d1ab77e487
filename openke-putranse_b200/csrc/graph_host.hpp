// Host-side knowledge-graph state for the PuTransE hot path: triple files, sorted indexes,
// Bernoulli statistics, the two random generators the reference consumes, universe (subgraph)
// construction and the link-prediction work lists.
//
// This is the B200 build's counterpart of the reference's header-only native core
// (reference: openke/base/{Reader,Random,Corrupt,UniverseConstructor,UniverseSetting,Test}.h).
// It is written from scratch around different data structures (int32 ids, std::vector storage,
// a re-entrant generator object instead of libc's hidden rand() state, a Fenwick tree and a hash
// set where the reference scans), but every integer it produces is bit-identical to the
// reference's, because subgraph sampling and negative corruption are integer work.
#pragma once
#include <cstdint>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace pk {

struct Tri {
    int32_t h, r, t;
    bool operator==(const Tri& o) const { return h == o.h && r == o.r && t == o.t; }
};

// Orderings of reference openke/base/Triple.h:9-23.
inline bool less_hrt(const Tri& a, const Tri& b) {
    return a.h != b.h ? a.h < b.h : (a.r != b.r ? a.r < b.r : a.t < b.t);
}
inline bool less_trh(const Tri& a, const Tri& b) {
    return a.t != b.t ? a.t < b.t : (a.r != b.r ? a.r < b.r : a.h < b.h);
}
inline bool less_rht(const Tri& a, const Tri& b) {
    return a.r != b.r ? a.r < b.r : (a.h != b.h ? a.h < b.h : a.t < b.t);
}

// glibc's rand()/srand() (TYPE_3 additive feedback generator, r[i] = r[i-3] + r[i-31]) as a value
// type, so that every universe can own its stream and universes can be built on many host threads.
// The reference calls libc rand() directly (openke/base/Random.h:11-15,32-34); libc's generator is
// process-global, which is what makes the reference single-threaded here.
class GlibcRand {
  public:
    explicit GlibcRand(uint32_t seed = 1) { reseed(seed); }
    void reseed(uint32_t seed);
    int32_t next();  // == rand()
    // reference Random.h:32-34  rand(a,b) = rand() % (b-a) + a
    int64_t range(int64_t a, int64_t b) { return (int64_t)next() % (b - a) + a; }

  private:
    uint32_t r_[34];
    int f_, b_;
};

// The per-thread 64-bit LCG of reference openke/base/Random.h:18-29.
constexpr uint64_t kLcgMul = 25214903917ULL;
constexpr uint64_t kLcgInc = 11ULL;

// A triple list with the sorted copies the sampler needs, in one id space (global graph or one
// universe's local ids).  Reference: loadHelpers (openke/base/Reader.h:58-167) builds four sorted
// copies and eight range arrays; the hot path reads only the (h,r,t) and (t,r,h) orders and the
// per-entity ranges of those two, so that is what is kept.
struct TripleIndex {
    int64_t n_ent = 0, n_rel = 0;
    std::vector<Tri> by_head;  // sorted (h,r,t), de-duplicated; this IS trainList / trainHead
    std::vector<Tri> by_tail;  // sorted (t,r,h)                 ; trainTail
    std::vector<int64_t> lef_head, rig_head, lef_tail, rig_tail;  // [n_ent], rig = -1 if absent
    std::vector<float> left_mean, right_mean;                     // [n_rel]
    int64_t n_tri() const { return (int64_t)by_head.size(); }
    // distinct (entity, relation) pairs per relation on the head / tail side
    void count_distinct(std::vector<int64_t>& heads_per_rel, std::vector<int64_t>& tails_per_rel) const;
    void build_ranges();
};

struct Universe {
    int64_t seed = 0, tc = 0;
    float balance = 0;
    int64_t focus = -1;
    std::vector<Tri> collected;      // global ids, collection order (trainListUniverse)
    std::vector<int32_t> ent_remap;  // local -> global
    std::vector<int32_t> rel_remap;
    TripleIndex local;               // local ids (trainListUniverseEnum + helpers)
    bool has_helpers = true;         // by_tail / ranges / means were built
    uint64_t lcg[64];                // sampler streams as randReset() leaves them (first work_threads used)
    int64_t draws = 0;               // libc-equivalent rand() calls consumed (diagnostic)
};

struct Graph {
    std::string in_path = "./";
    int64_t work_threads = 1;
    int64_t bern = 0;
    int64_t n_ent = 0, n_rel = 0;
    int64_t train_lines = 0;   // raw line count of train2id.txt
    uint64_t version = 0;      // unique per rebuilt training index (device mirrors key on it)
    int64_t import_count = 0;  // how many times the training files were imported (bern drift)
    bool load_all_triples = false;  // activateLoadOfAllTriples (Reader.h:240-244): filter set = triple2id.txt
    TripleIndex train;
    std::vector<Tri> by_rel;   // train sorted (r,h,t) + ranges, for the universe focus set
    std::vector<int64_t> lef_rel, rig_rel;
    // per relation: the entities it occurs with, ascending and distinct (offsets [n_rel+1] into rel_ent);
    // this is what every universe of that focus relation starts from (UniverseConstructor.h:69-80)
    std::vector<int64_t> rel_ent_off;
    std::vector<int32_t> rel_ent;
    std::vector<int32_t> tail_to_head;  // position in by_head of train.by_tail[i] (one id per triple)
    struct EntRange { int32_t lef_head, rig_head, lef_tail, rig_tail; };   // rig = -1 if absent
    std::vector<EntRange> ent_range;    // the four per-entity ranges of `train` in one cache line (the walk reads all four)
    std::vector<Tri> test, valid;  // sorted (r,h,t)   (reference Reader.h:311-312)
    std::vector<Tri> all_hrt;      // test ∪ train ∪ valid, distinct, sorted (h,r,t) (the filter set)
    std::vector<Tri> all_trh;      // same set sorted (t,r,h)

    bool import_train(std::string* err);
    bool import_test(std::string* err);
    void finish_train_index(bool same_shape, bool drift);

    // ---- incremental setting (reference openke/base/Incremental.h): the training list evolves snapshot by
    //      snapshot by replaying "h t r +|-" operations; evaluation lists and the filter set are per snapshot
    struct TripleOp { Tri t; char op; };
    bool incremental = false;
    int64_t num_snapshots = 0, ops_rate = 0;
    std::vector<TripleOp> ops;                         // KnowledgeGraphOperations
    size_t next_op = 0;
    std::unordered_map<uint64_t, int32_t> train_count; // the current training list as a multiset (duplicates are kept)
    std::vector<int64_t> train_rel_count;              // its triples per relation
    // the reference's arrays, in ITS order (append on first or renewed appearance, erase-and-shift on removal): the
    // universe focus is drawn by position from train_rel_contained (UniverseConstructor.h:336-339)
    std::vector<int32_t> train_rel_contained, train_rel_all, train_rel_deleted;
    std::vector<int32_t> head_canon;                   // by_head position -> first position of the same triple (empty: no duplicates)
    std::vector<int32_t> contained_entities;           // entities of the snapshot's triple list, ascending (Incremental.h:850-870)
    std::vector<int32_t> contained_relations;
    uint64_t tri_key(const Tri& x) const { return ((uint64_t)x.h * (uint64_t)n_rel + (uint64_t)x.r) * (uint64_t)n_ent + (uint64_t)x.t; }
    bool read_global_totals(std::string* err);                       // readGlobalNumEntities / readGlobalNumRelations
    bool load_train_ops(int snapshot, std::string* err);             // initializeTrainingOperations
    bool evolve_train(std::string* err);                             // evolveTrainList
    bool load_snapshot_triples(int snapshot, std::string* err);      // loadSnapshotTriples
    bool load_snapshot_eval(int snapshot, int which, std::string* err);  // loadTestData / loadValidData
    // Known-true candidates for query i of `which` (0 test, 1 valid) on `side` (0 head, 1 tail),
    // excluding the true entity itself; ascending.  Equivalent to the reference's per-candidate
    // _find() (openke/base/Corrupt.h:188-199) evaluated for every other entity.
    void filter_candidates(int which, int side, std::vector<int64_t>& offsets, std::vector<int32_t>& cand) const;
    // reference getParallelUniverse (openke/base/UniverseConstructor.h:327-397) continuing the given
    // generator, as the reference continues libc's after setRandomSeed()+randReset().
    // helpers = false: only what unfiltered, non-Bernoulli training reads is built (the (h,r,t)-sorted local list and
    // the remaps); the (t,r,h) order, the per-entity ranges and the Bernoulli means are skipped
    bool walk_universe(GlibcRand& rng, int64_t tc, float balance, Universe* out, std::string* err, bool helpers = true) const;
    // srand(seed); randReset(); getParallelUniverse(tc, balance) in one re-entrant call: what
    // Parallel_Universe_Config.set_random_seed + compile_train_datset do per universe
    // (openke/config/Parallel_Universe_Config.py:157-161,209-226).
    bool build_universe(int64_t seed, int64_t tc, float balance, Universe* out, std::string* err, bool helpers = true) const;
};

}  // namespace pk
