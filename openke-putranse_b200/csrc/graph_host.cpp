// Host-side graph state for the PuTransE hot path.  See graph_host.hpp for the map to the
// reference files.  Nothing here is a numerical fallback for the CUDA path: this is the integer
// work the reference also does on the host (file parsing, index building, subgraph sampling).
#include "graph_host.hpp"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <unordered_set>

namespace pk {

// ------------------------------------------------------------------------------------------------
// glibc rand(): srandom_r seeds 31 words with the Park-Miller LCG (Schrage form), then discards
// 310 outputs of x[i] = x[i-3] + x[i-31]; rand() returns the sum shifted right by one.
void GlibcRand::reseed(uint32_t seed) {
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    r_[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        r_[i] = (uint32_t)word;
    }
    f_ = 3;
    b_ = 0;
    for (int i = 0; i < 310; ++i) (void)next();
}

int32_t GlibcRand::next() {
    uint32_t v = (r_[f_] += r_[b_]);
    if (++f_ >= 31) f_ = 0;
    if (++b_ >= 31) b_ = 0;
    return (int32_t)(v >> 1);
}

// ------------------------------------------------------------------------------------------------
static bool read_triples(const std::string& path, std::vector<Tri>* out, int64_t* lines, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string buf((size_t)sz, '\0');
    if (sz > 0 && fread(&buf[0], 1, (size_t)sz, f) != (size_t)sz) {
        fclose(f);
        *err = "short read on " + path;
        return false;
    }
    fclose(f);
    // The reference takes the record count from the number of '\n' (openke/base/Utilities.h:47-57)
    // and then reads that many "h t r" records with fscanf (openke/base/Reader.h:188-197).
    int64_t n = 0;
    for (char c : buf) n += (c == '\n');
    *lines = n;
    out->clear();
    out->reserve((size_t)n);
    const char* p = buf.c_str();
    auto next_int = [&](int64_t* v) -> bool {
        while (*p && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
        if (!*p) return false;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; }
        if (*p < '0' || *p > '9') return false;
        int64_t x = 0;
        while (*p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
        *v = neg ? -x : x;
        return true;
    };
    for (int64_t i = 0; i < n; ++i) {
        int64_t h, t, r;
        if (!next_int(&h) || !next_int(&t) || !next_int(&r)) {
            *err = "malformed triple record in " + path;
            return false;
        }
        out->push_back(Tri{(int32_t)h, (int32_t)r, (int32_t)t});
    }
    return true;
}

static bool count_lines(const std::string& path, int64_t* n, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    char buf[1 << 16];
    int64_t c = 0;
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0)
        for (size_t i = 0; i < got; ++i) c += (buf[i] == '\n');
    fclose(f);
    *n = c;
    return true;
}

static void sort_unique(std::vector<Tri>& v, bool (*less)(const Tri&, const Tri&)) {
    std::sort(v.begin(), v.end(), less);
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

void TripleIndex::build_ranges() {
    // reference Reader.h:106-146: lef defaults to 0 (calloc), rig to -1
    lef_head.assign((size_t)n_ent, 0);
    rig_head.assign((size_t)n_ent, -1);
    lef_tail.assign((size_t)n_ent, 0);
    rig_tail.assign((size_t)n_ent, -1);
    const int64_t n = n_tri();
    for (int64_t i = 0; i < n; ++i) {
        const int32_t h = by_head[(size_t)i].h, t = by_tail[(size_t)i].t;
        if (i == 0 || by_head[(size_t)i - 1].h != h) lef_head[(size_t)h] = i;
        rig_head[(size_t)h] = i;
        if (i == 0 || by_tail[(size_t)i - 1].t != t) lef_tail[(size_t)t] = i;
        rig_tail[(size_t)t] = i;
    }
}

void TripleIndex::count_distinct(std::vector<int64_t>& heads, std::vector<int64_t>& tails) const {
    heads.assign((size_t)n_rel, 0);
    tails.assign((size_t)n_rel, 0);
    const int64_t n = n_tri();
    for (int64_t i = 0; i < n; ++i) {
        const Tri& a = by_head[(size_t)i];
        if (i == 0 || by_head[(size_t)i - 1].h != a.h || by_head[(size_t)i - 1].r != a.r) heads[(size_t)a.r]++;
        const Tri& b = by_tail[(size_t)i];
        if (i == 0 || by_tail[(size_t)i - 1].t != b.t || by_tail[(size_t)i - 1].r != b.r) tails[(size_t)b.r]++;
    }
}

// The reference adds 1.0 once per distinct (entity, relation) pair to a float that — on a repeated
// import in the same process — still holds the previous import's mean, and divides an un-zeroed,
// therefore n-times-accumulated, frequency by it (Reader.h:148-166 with Utilities.h:60-96).  The
// chain of float additions is reproduced literally so the resulting probabilities are bit-equal.
static float bump(float start, int64_t times) {
    float v = start;
    for (int64_t i = 0; i < times; ++i) v = (float)((double)v + 1.0);
    return v;
}

bool Graph::import_train(std::string* err) {
    int64_t nr, ne;
    if (!count_lines(in_path + "relation2id.txt", &nr, err)) return false;
    if (!count_lines(in_path + "entity2id.txt", &ne, err)) return false;
    std::vector<Tri> raw;
    if (!read_triples(in_path + "train2id.txt", &raw, &train_lines, err)) return false;
    if (raw.empty()) {
        *err = "empty training set in " + in_path;
        return false;
    }
    for (const Tri& x : raw)
        if (x.h < 0 || x.t < 0 || x.r < 0 || x.h >= ne || x.t >= ne || x.r >= nr) {
            *err = "triple id out of range in " + in_path + "train2id.txt";
            return false;
        }
    const bool same_shape = (nr == n_rel && (int64_t)train.left_mean.size() == nr);
    n_rel = nr;
    n_ent = ne;
    train.n_ent = ne;
    train.n_rel = nr;
    train.by_head = raw;
    sort_unique(train.by_head, less_hrt);
    finish_train_index(same_shape, /*drift=*/true);
    return true;
}

// Everything that is derived from the (h,r,t)-sorted training list `train.by_head` (with or without duplicate
// records): the (t,r,h) order, per-entity and per-relation ranges, the per-relation entity lists the universes
// start from, and the Bernoulli statistics.
void Graph::finish_train_index(bool same_shape, bool drift) {
    static std::atomic<uint64_t> next_version{1};
    version = next_version.fetch_add(1);
    const int64_t nr = n_rel;
    train.by_tail = train.by_head;
    std::sort(train.by_tail.begin(), train.by_tail.end(), less_trh);
    train.build_ranges();
    by_rel = train.by_head;
    std::sort(by_rel.begin(), by_rel.end(), less_rht);
    lef_rel.assign((size_t)nr, 0);
    rig_rel.assign((size_t)nr, -1);
    for (int64_t i = 0; i < (int64_t)by_rel.size(); ++i) {
        const int32_t r = by_rel[(size_t)i].r;
        if (i == 0 || by_rel[(size_t)i - 1].r != r) lef_rel[(size_t)r] = i;
        rig_rel[(size_t)r] = i;
    }
    // entity sets per relation and the triple ids of the (t,r,h) order: built once, read by every universe
    rel_ent_off.assign((size_t)nr + 1, 0);
    rel_ent.clear();
    {
        std::vector<int32_t> tmp;
        for (int64_t r = 0; r < nr; ++r) {
            tmp.clear();
            if (rig_rel[(size_t)r] >= 0)
                for (int64_t k = lef_rel[(size_t)r]; k <= rig_rel[(size_t)r]; ++k) {
                    tmp.push_back(by_rel[(size_t)k].h);
                    tmp.push_back(by_rel[(size_t)k].t);
                }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            rel_ent.insert(rel_ent.end(), tmp.begin(), tmp.end());
            rel_ent_off[(size_t)r + 1] = (int64_t)rel_ent.size();
        }
    }
    ent_range.resize((size_t)n_ent);
    for (int64_t e = 0; e < n_ent; ++e)
        ent_range[(size_t)e] = EntRange{(int32_t)train.lef_head[(size_t)e], (int32_t)train.rig_head[(size_t)e],
                                         (int32_t)train.lef_tail[(size_t)e], (int32_t)train.rig_tail[(size_t)e]};
    tail_to_head.resize(train.by_tail.size());
    for (size_t k = 0; k < train.by_tail.size(); ++k)
        tail_to_head[k] = (int32_t)(std::lower_bound(train.by_head.begin(), train.by_head.end(), train.by_tail[k], less_hrt) - train.by_head.begin());

    // positions of by_head that hold the same triple share one id for the walk's collected-before test (the
    // reference compares triple VALUES, UniverseConstructor.h:82-90); only the incremental list has duplicates
    head_canon.clear();
    for (size_t k = 1; k < train.by_head.size(); ++k)
        if (train.by_head[k] == train.by_head[k - 1]) {
            if (head_canon.empty()) {
                head_canon.resize(train.by_head.size());
                for (size_t j = 0; j < head_canon.size(); ++j) head_canon[j] = (int32_t)j;
            }
            head_canon[k] = head_canon[k - 1];
        }

    // Bernoulli statistics with the reference's import-count drift (static reader), or computed afresh
    // (incremental: resetIncrementalHelpers frees the arrays, Incremental.h:679-697).
    import_count = (drift && same_shape) ? import_count + 1 : 1;
    if (!same_shape || !drift) {
        train.left_mean.assign((size_t)nr, 0.f);
        train.right_mean.assign((size_t)nr, 0.f);
    }
    std::vector<int64_t> freq((size_t)nr, 0), dh, dt;
    for (const Tri& x : train.by_head) freq[(size_t)x.r]++;
    train.count_distinct(dh, dt);
    for (int64_t r = 0; r < nr; ++r) {
        const int64_t f = freq[(size_t)r] * import_count;  // INT accumulated over imports
        train.left_mean[(size_t)r] = (float)f / bump(train.left_mean[(size_t)r], dh[(size_t)r]);
        train.right_mean[(size_t)r] = (float)f / bump(train.right_mean[(size_t)r], dt[(size_t)r]);
    }
}

bool Graph::import_test(std::string* err) {
    int64_t nr, ne, lines;
    if (!count_lines(in_path + "relation2id.txt", &nr, err)) return false;
    if (!count_lines(in_path + "entity2id.txt", &ne, err)) return false;
    n_rel = nr;
    n_ent = ne;
    std::vector<Tri> tr;
    if (!read_triples(in_path + "test2id.txt", &test, &lines, err)) return false;
    if (!read_triples(in_path + "train2id.txt", &tr, &lines, err)) return false;
    if (!read_triples(in_path + "valid2id.txt", &valid, &lines, err)) return false;
    all_hrt.clear();
    all_hrt.reserve(test.size() + tr.size() + valid.size());
    all_hrt.insert(all_hrt.end(), test.begin(), test.end());
    all_hrt.insert(all_hrt.end(), tr.begin(), tr.end());
    all_hrt.insert(all_hrt.end(), valid.begin(), valid.end());
    if (load_all_triples) {   // Reader.h:295-308: the filter set is replaced by the file's triples
        if (!read_triples(in_path + "triple2id.txt", &all_hrt, &lines, err)) return false;
    }
    for (const Tri& x : all_hrt)
        if (x.h < 0 || x.t < 0 || x.r < 0 || x.h >= ne || x.t >= ne || x.r >= nr) {
            *err = "triple id out of range under " + in_path;
            return false;
        }
    // The reference keeps duplicates in tripleList; membership (_find) is all that is ever asked.
    sort_unique(all_hrt, less_hrt);
    all_trh = all_hrt;
    std::sort(all_trh.begin(), all_trh.end(), less_trh);
    std::sort(test.begin(), test.end(), less_rht);    // Reader.h:311
    std::sort(valid.begin(), valid.end(), less_rht);  // Reader.h:312
    return true;
}

void Graph::filter_candidates(int which, int side, std::vector<int64_t>& off, std::vector<int32_t>& cand) const {
    const std::vector<Tri>& q = which == 0 ? test : valid;
    off.assign(q.size() + 1, 0);
    cand.clear();
    for (size_t i = 0; i < q.size(); ++i) {
        const Tri& x = q[i];
        if (side == 0) {  // head prediction: all j with (j, r, t) known, j != h
            Tri lo{INT32_MIN, x.r, x.t};
            auto it = std::lower_bound(all_trh.begin(), all_trh.end(), lo, less_trh);
            for (; it != all_trh.end() && it->t == x.t && it->r == x.r; ++it)
                if (it->h != x.h) cand.push_back(it->h);
        } else {  // tail prediction: all j with (h, r, j) known, j != t
            Tri lo{x.h, x.r, INT32_MIN};
            auto it = std::lower_bound(all_hrt.begin(), all_hrt.end(), lo, less_hrt);
            for (; it != all_hrt.end() && it->h == x.h && it->r == x.r; ++it)
                if (it->t != x.t) cand.push_back(it->t);
        }
        off[i + 1] = (int64_t)cand.size();
    }
}

// ------------------------------------------------------------------------------------------------
// Universe construction.  Draw order of the libc generator is the contract (SURVEY.md appendix A,
// reference UniverseConstructor.h:39-67,92-191,327-397); the containers are not.
namespace {

// k-th remaining element of a sorted array under deletions (replaces std::advance over std::set).
struct Fenwick {
    std::vector<int32_t> t;
    int n, top;
    explicit Fenwick(int n_) : t((size_t)n_ + 1, 0), n(n_) {
        for (int i = 1; i <= n; ++i) t[(size_t)i] = i & -i;   // all ones: node i covers lowbit(i) elements
        top = 1;
        while (top * 2 <= n) top *= 2;
    }
    int kth(int k) const {  // 0-based rank -> 0-based position
        int pos = 0, rem = k + 1;
        for (int s = top; s > 0; s >>= 1)
            if (pos + s <= n && t[(size_t)(pos + s)] < rem) {
                pos += s;
                rem -= t[(size_t)pos];
            }
        return pos;
    }
    void remove(int p) {
        for (int i = p + 1; i <= n; i += i & -i) t[(size_t)i] -= 1;
    }
};


}  // namespace

bool Graph::build_universe(int64_t seed, int64_t tc, float balance, Universe* u, std::string* err, bool helpers) const {
    GlibcRand rng((uint32_t)seed);                       // setRandomSeed -> srand   (Random.h:38-45)
    std::memset(u->lcg, 0, sizeof u->lcg);
    const int64_t wt = std::min<int64_t>(work_threads, 64);
    for (int64_t i = 0; i < wt; ++i) u->lcg[i] = (uint64_t)(int64_t)rng.next();  // randReset (Random.h:11-15)
    u->seed = seed;
    const bool ok = walk_universe(rng, tc, balance, u, err, helpers);
    u->draws += wt;
    return ok;
}

bool Graph::walk_universe(GlibcRand& rng, int64_t tc, float balance, Universe* u, std::string* err, bool helpers) const {
    if (train.n_tri() == 0) {
        *err = "build_universe: no training graph imported";
        return false;
    }
    if (tc <= 0) {
        *err = "build_universe: triple_constraint must be positive";
        return false;
    }
    u->tc = tc;
    u->balance = balance;
    int64_t draws = 0;

    int64_t focus;
    if (incremental) {   // UniverseConstructor.h:336-339: a relation the evolving training list currently holds
        if (train_rel_contained.empty()) {
            *err = "build_universe: the incremental training list holds no relation (evolveTrainList has not run)";
            return false;
        }
        focus = train_rel_contained[(size_t)rng.range(0, (int64_t)train_rel_contained.size())];
    } else {
        focus = rng.range(0, n_rel);                     // UniverseConstructor.h:341
    }
    ++draws;
    u->focus = focus;
    const int64_t threshold = (int64_t)(balance * (float)tc);  // :345-346 (float product, truncated)

    // entities that occur with the focus relation, ascending (:69-80): precomputed at import
    const int32_t* focus_ent = rel_ent.data() + rel_ent_off[(size_t)focus];
    const int64_t n_focus = rel_ent_off[(size_t)focus + 1] - rel_ent_off[(size_t)focus];
    std::vector<int32_t> frontier;
    if (threshold >= 0 && n_focus > threshold) {  // :352-354, :55-67
        Fenwick fw((int)n_focus);
        int64_t remaining = n_focus;
        frontier.reserve((size_t)threshold);
        while ((int64_t)frontier.size() < threshold) {
            const int k = (int)((int64_t)rng.next() % remaining);
            ++draws;
            const int pos = fw.kth(k);
            frontier.push_back(focus_ent[pos]);
            fw.remove(pos);
            --remaining;
        }
        std::sort(frontier.begin(), frontier.end());
    } else {
        frontier.assign(focus_ent, focus_ent + n_focus);
    }

    // bidirectional random walk (:92-191)
    const TripleIndex& g = train;
    std::vector<Tri>& got = u->collected;
    got.clear();
    got.reserve((size_t)tc);
    // collected-before test: one BIT per training triple (by_head position), per thread, instead of hashing
    // (17 KB for WN18: stays in L1/L2 beside the graph arrays); the bits set by a walk are cleared when it ends
    static thread_local std::vector<uint64_t> seen, queued;
    static thread_local std::vector<int32_t> got_tid;
    if (seen.size() != (g.by_head.size() + 63) / 64 || queued.size() != ((size_t)n_ent + 63) / 64) {
        seen.assign((g.by_head.size() + 63) / 64, 0);
        queued.assign(((size_t)n_ent + 63) / 64, 0);
    }
    got_tid.clear();
    auto test_bit = [](const std::vector<uint64_t>& v, int64_t i) { return (v[(size_t)i >> 6] >> (i & 63)) & 1ULL; };
    auto set_bit = [](std::vector<uint64_t>& v, int64_t i) { v[(size_t)i >> 6] |= 1ULL << (i & 63); };
    auto clr_bit = [](std::vector<uint64_t>& v, int64_t i) { v[(size_t)i >> 6] &= ~(1ULL << (i & 63)); };
    // the reference's std::set of next starting points: a vector de-duplicated by a bitmap on insertion and
    // sorted when the round ends; it survives rounds (skipped entities resurface one round later)
    std::vector<int32_t> next_points;
    bool zero_seen = false, neg_queued = false;
    int64_t target = tc;
    int32_t last_dup_entity = -1;
    int dup_tol = 5, stall_tol = 20;
    int64_t last_size = 0;
    std::vector<int32_t> leftover;
    while ((int64_t)got.size() < target) {
        leftover.clear();
        size_t i = 0;
        while (i < frontier.size() && (int64_t)got.size() < target) {
            const int32_t e = frontier[i];
            const bool head_first = (rng.next() % 1000) < 500;  // :122 (prob is the float 500)
            ++draws;
            if (i + 1 < frontier.size()) __builtin_prefetch(&ent_range[(size_t)frontier[i + 1]]);
            const EntRange er = ent_range[(size_t)e];
            const bool has_h = er.rig_head != -1, has_t = er.rig_tail != -1;
            int side;  // 0 head, 1 tail
            if (head_first) side = has_h ? 0 : (has_t ? 1 : -1);
            else            side = has_t ? 1 : (has_h ? 0 : -1);
            Tri x{0, 0, 0};
            int32_t nxt = -1;
            int64_t tid = -1;   // the triple's position in by_head
            if (side == 0) {
                const int64_t idx = rng.range(er.lef_head, (int64_t)er.rig_head + 1);  // :40
                ++draws;
                x = g.by_head[(size_t)idx];
                nxt = x.t;
                tid = head_canon.empty() ? idx : head_canon[(size_t)idx];
            } else if (side == 1) {
                const int64_t idx = rng.range(er.lef_tail, (int64_t)er.rig_tail + 1);  // :48
                ++draws;
                x = g.by_tail[(size_t)idx];
                nxt = x.h;
                tid = tail_to_head[(size_t)idx];
            }
            if (tid < 0) {   // isolated entity: the reference compares the zero triple (0,0,0) (:141)
                const auto it0 = std::lower_bound(g.by_head.begin(), g.by_head.end(), x, less_hrt);
                if (it0 != g.by_head.end() && *it0 == x) tid = it0 - g.by_head.begin();
            }
            if (tid >= 0 ? test_bit(seen, tid) != 0 : zero_seen) {  // :141-154
                if (last_dup_entity == e) --dup_tol;
                else last_dup_entity = e;
                if (dup_tol == 0) {
                    dup_tol = 5;
                    leftover.push_back(e);
                    ++i;
                }
                continue;
            }
            got.push_back(x);
            if (tid >= 0) { set_bit(seen, tid); got_tid.push_back((int32_t)tid); } else zero_seen = true;
            if (nxt >= 0 && !test_bit(queued, nxt)) { set_bit(queued, nxt); next_points.push_back(nxt); }
            else if (nxt < 0 && !neg_queued) { neg_queued = true; next_points.push_back(nxt); }
            ++i;  // erase(it++)
        }
        for (; i < frontier.size(); ++i) leftover.push_back(frontier[i]);
        // entity_set.swap(new_starting_points) (:170)
        std::sort(next_points.begin(), next_points.end());
        frontier.swap(next_points);
        next_points.clear();
        for (int32_t e2 : frontier)   // (the old next_points) a new round: nothing is queued yet
            if (e2 >= 0) clr_bit(queued, e2);
        neg_queued = false;
        for (int32_t e2 : leftover)   // ascending and distinct already
            if (e2 >= 0) { set_bit(queued, e2); next_points.push_back(e2); }
            else if (!neg_queued) { neg_queued = true; next_points.push_back(e2); }
        if ((int64_t)got.size() == last_size) --stall_tol;
        else { last_size = (int64_t)got.size(); stall_tol = 20; }
        if (stall_tol == 0) {  // :181-186
            target = (int64_t)got.size();
            break;
        }
    }
    for (int32_t e2 : next_points)
        if (e2 >= 0) clr_bit(queued, e2);
    for (int32_t t2 : got_tid) clr_bit(seen, t2);
    u->draws = draws;
    if (got.empty()) {
        *err = "build_universe: random walk collected no triples";
        return false;
    }

    // local ids by first appearance: h, then t, then r (:193-233)
    static thread_local std::vector<int32_t> emap, rmap;   // all -1 between calls (touched entries are reset below)
    if (emap.size() != (size_t)n_ent) emap.assign((size_t)n_ent, -1);
    if (rmap.size() != (size_t)n_rel) rmap.assign((size_t)n_rel, -1);
    u->ent_remap.clear();
    u->rel_remap.clear();
    TripleIndex& L = u->local;
    L.by_head.resize(got.size());
    for (size_t k = 0; k < got.size(); ++k) {
        const Tri& x = got[k];
        if (emap[(size_t)x.h] < 0) { emap[(size_t)x.h] = (int32_t)u->ent_remap.size(); u->ent_remap.push_back(x.h); }
        if (emap[(size_t)x.t] < 0) { emap[(size_t)x.t] = (int32_t)u->ent_remap.size(); u->ent_remap.push_back(x.t); }
        if (rmap[(size_t)x.r] < 0) { rmap[(size_t)x.r] = (int32_t)u->rel_remap.size(); u->rel_remap.push_back(x.r); }
        L.by_head[k] = Tri{emap[(size_t)x.h], rmap[(size_t)x.r], emap[(size_t)x.t]};
    }
    for (int32_t e : u->ent_remap) emap[(size_t)e] = -1;
    for (int32_t r : u->rel_remap) rmap[(size_t)r] = -1;
    L.n_ent = (int64_t)u->ent_remap.size();
    L.n_rel = (int64_t)u->rel_remap.size();
    u->has_helpers = helpers;
    if (!helpers) {   // training without filter / Bernoulli reads the (h,r,t) list only
        L.by_tail.clear();
        L.lef_head.clear(); L.rig_head.clear(); L.lef_tail.clear(); L.rig_tail.clear();
        L.left_mean.clear(); L.right_mean.clear();
        if (L.n_ent < (1 << 21) && L.n_rel < (1 << 21)) {
            static thread_local std::vector<uint64_t> keys1;
            keys1.resize(got.size());
            for (size_t k = 0; k < got.size(); ++k) {
                const Tri& x = L.by_head[k];
                keys1[k] = ((uint64_t)x.h << 42) | ((uint64_t)x.r << 21) | (uint64_t)x.t;
            }
            std::sort(keys1.begin(), keys1.end());
            for (size_t k = 0; k < got.size(); ++k) {
                const uint64_t v = keys1[k];
                L.by_head[k] = Tri{(int32_t)(v >> 42), (int32_t)((v >> 21) & 0x1fffff), (int32_t)(v & 0x1fffff)};
            }
        } else {
            std::sort(L.by_head.begin(), L.by_head.end(), less_hrt);
        }
        return true;
    }
    L.by_tail.resize(got.size());
    if (L.n_ent < (1 << 21) && L.n_rel < (1 << 21)) {
        // the two orders (:236) as sorts of packed 63-bit keys
        static thread_local std::vector<uint64_t> keys;
        keys.resize(got.size());
        for (size_t k = 0; k < got.size(); ++k) {
            const Tri& x = L.by_head[k];
            keys[k] = ((uint64_t)x.h << 42) | ((uint64_t)x.r << 21) | (uint64_t)x.t;
        }
        std::sort(keys.begin(), keys.end());
        for (size_t k = 0; k < got.size(); ++k) {
            const uint64_t v = keys[k];
            L.by_head[k] = Tri{(int32_t)(v >> 42), (int32_t)((v >> 21) & 0x1fffff), (int32_t)(v & 0x1fffff)};
            keys[k] = ((v & 0x1fffff) << 42) | (v & (0x1fffffULL << 21)) | (v >> 42);
        }
        std::sort(keys.begin(), keys.end());
        for (size_t k = 0; k < got.size(); ++k) {
            const uint64_t v = keys[k];
            L.by_tail[k] = Tri{(int32_t)(v & 0x1fffff), (int32_t)((v >> 21) & 0x1fffff), (int32_t)(v >> 42)};
        }
    } else {
        std::sort(L.by_head.begin(), L.by_head.end(), less_hrt);  // :236
        L.by_tail = L.by_head;
        std::sort(L.by_tail.begin(), L.by_tail.end(), less_trh);
    }
    L.build_ranges();
    // universe-local Bernoulli statistics (:294-324); freshly allocated per universe, no drift
    std::vector<int64_t> freq((size_t)L.n_rel, 0), dh, dt;
    for (const Tri& x : L.by_head) freq[(size_t)x.r]++;
    L.count_distinct(dh, dt);
    L.left_mean.resize((size_t)L.n_rel);
    L.right_mean.resize((size_t)L.n_rel);
    for (int64_t r = 0; r < L.n_rel; ++r) {
        L.left_mean[(size_t)r] = (float)freq[(size_t)r] / bump(0.f, dh[(size_t)r]);
        L.right_mean[(size_t)r] = (float)freq[(size_t)r] / bump(0.f, dt[(size_t)r]);
    }
    return true;
}

}  // namespace pk

// ------------------------------------------------------------------------------------------------
// Incremental setting (reference openke/base/Incremental.h).  The reference keeps the evolving training list as
// a C array that it appends to and shift-deletes from, and re-derives "is this relation still present" by linear
// scans; here the list is a multiset keyed by the packed triple with a per-relation counter, which gives the same
// list (the reference sorts it before use, Incremental.h:710) and the same ORDER of the three relation arrays the
// universe focus is drawn from.
namespace pk {

bool Graph::read_global_totals(std::string* err) {   // Incremental.h:182-205
    int64_t ne, nr;
    if (!count_lines(in_path + "incremental/entity2id.txt", &ne, err)) return false;
    if (!count_lines(in_path + "incremental/relation2id.txt", &nr, err)) return false;
    if (ne != n_ent || nr != n_rel) {   // a different id space: nothing of the previous one survives
        train = TripleIndex();
        train_count.clear();
        train_rel_contained.clear();
        train_rel_all.clear();
        train_rel_deleted.clear();
        ops.clear();
        next_op = 0;
    }
    n_ent = ne;
    n_rel = nr;
    train.n_ent = ne;
    train.n_rel = nr;
    train_rel_count.resize((size_t)nr, 0);
    return true;
}

static bool read_ops(const std::string& path, int64_t ne, int64_t nr, std::vector<Graph::TripleOp>* out, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    out->clear();
    char line[256];
    while (fgets(line, sizeof line, f)) {
        long long h, t, r;
        char op[8] = {0};
        if (sscanf(line, "%lld %lld %lld %7s", &h, &t, &r, op) != 4) continue;   // the reference reads with fscanf per line count
        if (h < 0 || t < 0 || r < 0 || h >= ne || t >= ne || r >= nr || (op[0] != '+' && op[0] != '-')) {
            fclose(f);
            *err = "malformed operation in " + path;
            return false;
        }
        out->push_back(Graph::TripleOp{Tri{(int32_t)h, (int32_t)r, (int32_t)t}, op[0]});
    }
    fclose(f);
    return true;
}

bool Graph::load_train_ops(int snapshot, std::string* err) {   // initializeTrainingOperations, Incremental.h:299-321
    if (n_ent <= 0 || n_rel <= 0) {
        *err = "initializeTrainingOperations: readGlobalNumEntities/Relations have not run";
        return false;
    }
    if (!read_ops(in_path + "incremental/" + std::to_string(snapshot) + "/train-op2id.txt", n_ent, n_rel, &ops, err)) return false;
    next_op = 0;
    return true;
}

bool Graph::evolve_train(std::string* err) {   // evolveTrainList, Incremental.h:798-846
    size_t todo = ops_rate > 0 ? (size_t)ops_rate : ops.size();   // numOperationsRate == 0: the whole snapshot
    if (todo > ops.size() - next_op) todo = ops.size() - next_op;
    train_rel_count.resize((size_t)n_rel, 0);
    auto erase_value = [](std::vector<int32_t>& v, int32_t x) {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == x) { v.erase(v.begin() + (long)i); return true; }
        return false;
    };
    auto contains = [](const std::vector<int32_t>& v, int32_t x) {
        for (int32_t y : v) if (y == x) return true;
        return false;
    };
    for (size_t k = 0; k < todo; ++k, ++next_op) {
        const TripleOp& o = ops[next_op];
        const int32_t r = o.t.r;
        if (o.op == '+') {   // insertTrainTriple :562-573 with adjustTrainRelationSet :540-552
            if (train_rel_count[(size_t)r] == 0) {
                if (!contains(train_rel_all, r)) train_rel_all.push_back(r);
                else erase_value(train_rel_deleted, r);
                train_rel_contained.push_back(r);
            }
            ++train_count[tri_key(o.t)];
            ++train_rel_count[(size_t)r];
        } else {             // deleteTrainTriple :651-676: the first equal record goes; unknown triples are reported and skipped
            auto it = train_count.find(tri_key(o.t));
            if (it == train_count.end() || it->second == 0) continue;
            if (--it->second == 0) train_count.erase(it);
            if (--train_rel_count[(size_t)r] == 0) {   // trainRelationRemovalCheck :618-626
                erase_value(train_rel_contained, r);
                train_rel_deleted.push_back(r);
            }
        }
    }
    // resetSnapShot :219-230 is called at the end of evolveTrainList: operations that were not replayed are dropped
    ops_rate = 0;
    ops.clear();
    next_op = 0;
    std::vector<Tri> list;
    size_t total = 0;
    for (const auto& kv : train_count) total += (size_t)kv.second;
    if (total == 0) {
        *err = "evolveTrainList: the training list is empty";
        return false;
    }
    list.reserve(total);
    for (const auto& kv : train_count) {
        uint64_t key = kv.first;
        Tri x;
        x.t = (int32_t)(key % (uint64_t)n_ent); key /= (uint64_t)n_ent;
        x.r = (int32_t)(key % (uint64_t)n_rel); key /= (uint64_t)n_rel;
        x.h = (int32_t)key;
        for (int32_t c = 0; c < kv.second; ++c) list.push_back(x);
    }
    std::sort(list.begin(), list.end(), less_hrt);   // :710 (duplicates stay: the sampler indexes this list)
    train.n_ent = n_ent;
    train.n_rel = n_rel;
    train.by_head.swap(list);
    train_lines = (int64_t)train.by_head.size();
    finish_train_index(/*same_shape=*/false, /*drift=*/false);
    return true;
}

// loadSnapshotTriples (Incremental.h:891-924): the snapshot's whole triple list becomes the filter set, its
// entities the candidate set.  (The reference then copies the RELATION ids over the first entries of
// currently_contained_entities, :882-887 — a slip; the entity array is kept intact here.)
bool Graph::load_snapshot_triples(int snapshot, std::string* err) {
    std::vector<Tri> all;
    int64_t lines = 0;
    if (!read_triples(in_path + "incremental/" + std::to_string(snapshot) + "/global_triple2id.txt", &all, &lines, err)) return false;
    std::vector<int32_t> ents, rels;
    ents.reserve(all.size() * 2);
    for (const Tri& x : all) {
        if (x.h < 0 || x.t < 0 || x.r < 0 || x.h >= n_ent || x.t >= n_ent || x.r >= n_rel) {
            *err = "triple id out of range in global_triple2id.txt of snapshot " + std::to_string(snapshot);
            return false;
        }
        ents.push_back(x.h);
        ents.push_back(x.t);
        rels.push_back(x.r);
    }
    std::sort(ents.begin(), ents.end());
    ents.erase(std::unique(ents.begin(), ents.end()), ents.end());
    std::sort(rels.begin(), rels.end());
    rels.erase(std::unique(rels.begin(), rels.end()), rels.end());
    contained_entities.swap(ents);
    contained_relations.swap(rels);
    all_hrt.swap(all);
    sort_unique(all_hrt, less_hrt);
    all_trh = all_hrt;
    std::sort(all_trh.begin(), all_trh.end(), less_trh);
    return true;
}

// loadTestData / loadValidData (Incremental.h:248-296): the snapshot's list in FILE order (not re-sorted).
bool Graph::load_snapshot_eval(int snapshot, int which, std::string* err) {
    std::vector<Tri> q;
    int64_t lines = 0;
    const std::string name = which == 0 ? "/test2id.txt" : "/valid2id.txt";
    if (!read_triples(in_path + "incremental/" + std::to_string(snapshot) + name, &q, &lines, err)) return false;
    for (const Tri& x : q)
        if (x.h < 0 || x.t < 0 || x.r < 0 || x.h >= n_ent || x.t >= n_ent || x.r >= n_rel) {
            *err = "triple id out of range in " + name + " of snapshot " + std::to_string(snapshot);
            return false;
        }
    (which == 0 ? test : valid).swap(q);
    return true;
}

}  // namespace pk
