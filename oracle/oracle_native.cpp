// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product may include, link or load this file.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
// and only as the checker.
//
// Plain single-threaded C++ restatement of the integer half of the PuTransE hot path of
// luofeisg/OpenKE-PuTransE, written to be read side by side with the reference:
//   * triple file import, de-duplication, sorted copies, lef/rig ranges, Bernoulli means with the
//     import-count drift                    openke/base/Reader.h:58-234, openke/base/Utilities.h:60-96
//   * the per-thread LCG and sampling()     openke/base/Random.h:11-29, openke/base/Base.cpp:185-310
//   * corrupt_head / corrupt_tail           openke/base/Corrupt.h:9-105
//   * getParallelUniverse and its helpers   openke/base/UniverseConstructor.h:39-397 (libc rand())
//   * testHead / testTail rank counting     openke/base/Test.h:118-238, openke/base/Corrupt.h:188-199
// Parity status: PINNED — tests/test_oracle.py checks every function here against vectors produced
// by the unmodified reference (tests/golden/*.npz, minted by tests/golden/make_golden.py) and, when
// oracle/_ref/Base.so is present, against the reference library itself on fresh inputs.
//
// Deliberately naive data structures (std::set, linear scans) like the reference; int64 ids.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

typedef long INT;
typedef float REAL;

struct T3 { INT h, r, t; };
static bool cmp_head(const T3& a, const T3& b) { return a.h < b.h || (a.h == b.h && a.r < b.r) || (a.h == b.h && a.r == b.r && a.t < b.t); }
static bool cmp_tail(const T3& a, const T3& b) { return a.t < b.t || (a.t == b.t && a.r < b.r) || (a.t == b.t && a.r == b.r && a.h < b.h); }
static bool cmp_rel2(const T3& a, const T3& b) { return a.r < b.r || (a.r == b.r && a.h < b.h) || (a.r == b.r && a.h == b.h && a.t < b.t); }

struct Index {  // one id space: global graph or a universe
    INT nE = 0, nR = 0, nT = 0;
    std::vector<T3> list, head, tail, rel2;
    std::vector<INT> lefHead, rigHead, lefTail, rigTail, lefRel2, rigRel2;
    std::vector<REAL> left_mean, right_mean;
    std::vector<INT> freqRel;
};

struct Oracle {
    Index g;            // global graph
    Index u;            // current universe (local ids)
    std::vector<T3> u_global;  // trainListUniverse
    std::vector<INT> ent_remap, rel_remap;
    bool swapped = false;
    INT threads = 1, bern = 0;
    unsigned long long lcg[64];
    std::vector<T3> testList, validList, tripleList;
    int imports = 0;
};

static Index& cur(Oracle* o) { return o->swapped ? o->u : o->g; }

// Reader.h:58-167.  `keep_means`: the arrays survive from the previous import (the drift).
static void load_helpers(Index& x, bool keep_means, int import_no) {
    x.nT = (INT)x.list.size();
    std::sort(x.list.begin(), x.list.end(), cmp_head);
    x.head = x.tail = x.rel2 = x.list;
    std::sort(x.tail.begin(), x.tail.end(), cmp_tail);
    std::sort(x.rel2.begin(), x.rel2.end(), cmp_rel2);
    if (!keep_means || (INT)x.freqRel.size() != x.nR) x.freqRel.assign(x.nR, 0);
    for (INT i = 0; i < x.nT; i++) x.freqRel[x.list[i].r]++;   // realloc without zeroing => accumulates
    (void)import_no;
    x.lefHead.assign(x.nE, 0); x.rigHead.assign(x.nE, -1);
    x.lefTail.assign(x.nE, 0); x.rigTail.assign(x.nE, -1);
    x.lefRel2.assign(x.nR, 0); x.rigRel2.assign(x.nR, -1);
    for (INT i = 1; i < x.nT; i++) {
        if (x.tail[i].t != x.tail[i - 1].t) { x.rigTail[x.tail[i - 1].t] = i - 1; x.lefTail[x.tail[i].t] = i; }
        if (x.head[i].h != x.head[i - 1].h) { x.rigHead[x.head[i - 1].h] = i - 1; x.lefHead[x.head[i].h] = i; }
        if (x.rel2[i].r != x.rel2[i - 1].r) { x.rigRel2[x.rel2[i - 1].r] = i - 1; x.lefRel2[x.rel2[i].r] = i; }
    }
    x.lefHead[x.head[0].h] = 0; x.rigHead[x.head[x.nT - 1].h] = x.nT - 1;
    x.lefTail[x.tail[0].t] = 0; x.rigTail[x.tail[x.nT - 1].t] = x.nT - 1;
    x.lefRel2[x.rel2[0].r] = 0; x.rigRel2[x.rel2[x.nT - 1].r] = x.nT - 1;
    if (!keep_means || (INT)x.left_mean.size() != x.nR) { x.left_mean.assign(x.nR, 0.f); x.right_mean.assign(x.nR, 0.f); }
    for (INT i = 0; i < x.nE; i++) {
        for (INT j = x.lefHead[i] + 1; j <= x.rigHead[i]; j++)
            if (x.head[j].r != x.head[j - 1].r) x.left_mean[x.head[j].r] += 1.0;
        if (x.lefHead[i] <= x.rigHead[i]) x.left_mean[x.head[x.lefHead[i]].r] += 1.0;
        for (INT j = x.lefTail[i] + 1; j <= x.rigTail[i]; j++)
            if (x.tail[j].r != x.tail[j - 1].r) x.right_mean[x.tail[j].r] += 1.0;
        if (x.lefTail[i] <= x.rigTail[i]) x.right_mean[x.tail[x.lefTail[i]].r] += 1.0;
    }
    for (INT i = 0; i < x.nR; i++) {
        x.left_mean[i] = x.freqRel[i] / x.left_mean[i];
        x.right_mean[i] = x.freqRel[i] / x.right_mean[i];
    }
}

// Random.h:18-29
static unsigned long long randd(Oracle* o, INT id) {
    o->lcg[id] = o->lcg[id] * (unsigned long long)(25214903917) + 11;
    return o->lcg[id];
}
static INT rand_max(Oracle* o, INT id, INT x) {
    INT res = randd(o, id) % x;
    while (res < 0) res += x;
    return res;
}
static INT rand_ab(INT a, INT b) { return (rand() % (b - a)) + a; }  // Random.h:32-34 (libc)

// Corrupt.h:9-57 (fix = h, varying column = t over trainHead) and :59-105 (fix = t over trainTail)
static INT corrupt(Oracle* o, INT id, INT fix, INT r, bool filter, bool head_index) {
    Index& x = cur(o);
    if (!filter) {
        INT tmp = rand_max(o, id, x.nE - 1);
        return tmp < fix ? tmp : tmp + 1;
    }
    const std::vector<T3>& a = head_index ? x.head : x.tail;
    const std::vector<INT>& L = head_index ? x.lefHead : x.lefTail;
    const std::vector<INT>& R = head_index ? x.rigHead : x.rigTail;
    auto var = [&](INT i) { return head_index ? a[i].t : a[i].h; };
    INT lef = L[fix] - 1, rig = R[fix], mid, ll, rr;
    while (lef + 1 < rig) { mid = (lef + rig) >> 1; if (a[mid].r >= r) rig = mid; else lef = mid; }
    ll = rig;
    lef = L[fix]; rig = R[fix] + 1;
    while (lef + 1 < rig) { mid = (lef + rig) >> 1; if (a[mid].r <= r) lef = mid; else rig = mid; }
    rr = lef;
    INT tmp = rand_max(o, id, x.nE - (rr - ll + 1));
    if (tmp < var(ll)) return tmp;
    if (tmp > var(rr) - rr + ll - 1) return tmp + rr - ll + 1;
    lef = ll; rig = rr + 1;
    while (lef + 1 < rig) { mid = (lef + rig) >> 1; if (var(mid) - mid + ll - 1 < tmp) lef = mid; else rig = mid; }
    return tmp + lef - ll + 1;
}

extern "C" {

Oracle* oracle_new() { Oracle* o = new Oracle(); std::memset(o->lcg, 0, sizeof o->lcg); return o; }
void oracle_free(Oracle* o) { delete o; }
void oracle_set(Oracle* o, INT threads, INT bern) { o->threads = threads; o->bern = bern; }

// importTrainFiles on arrays instead of files (h,t,r columns as in train2id.txt); Reader.h:169-234
void oracle_import_train(Oracle* o, const INT* htr, INT n, INT nE, INT nR) {
    Index& g = o->g;
    const bool keep = (g.nR == nR && !g.left_mean.empty());
    g.nE = nE; g.nR = nR;
    g.list.resize(n);
    for (INT i = 0; i < n; i++) g.list[i] = T3{htr[3 * i], htr[3 * i + 2], htr[3 * i + 1]};
    std::sort(g.list.begin(), g.list.end(), cmp_head);
    std::vector<T3> d;
    d.push_back(g.list[0]);
    for (INT i = 1; i < n; i++)
        if (g.list[i].h != g.list[i - 1].h || g.list[i].r != g.list[i - 1].r || g.list[i].t != g.list[i - 1].t) d.push_back(g.list[i]);
    g.list.swap(d);
    o->imports = keep ? o->imports + 1 : 1;
    load_helpers(g, keep, o->imports);
}
INT oracle_train_total(Oracle* o) { return cur(o).nT; }
INT oracle_ent_total(Oracle* o) { return cur(o).nE; }
INT oracle_rel_total(Oracle* o) { return cur(o).nR; }
void oracle_means(Oracle* o, REAL* l, REAL* r) {
    Index& x = cur(o);
    for (INT i = 0; i < x.nR; i++) { l[i] = x.left_mean[i]; r[i] = x.right_mean[i]; }
}
void oracle_train_list(Oracle* o, INT* hrt) {
    Index& x = cur(o);
    for (INT i = 0; i < x.nT; i++) { hrt[3 * i] = x.list[i].h; hrt[3 * i + 1] = x.list[i].r; hrt[3 * i + 2] = x.list[i].t; }
}

// setRandomSeed + randReset (Random.h:11-15,38-45): libc generator, process-global like the reference
void oracle_seed(Oracle* o, INT seed) {
    srand((unsigned)seed);
    for (INT i = 0; i < o->threads; i++) o->lcg[i] = rand();
}
void oracle_get_lcg(Oracle* o, unsigned long long* s) { for (INT i = 0; i < o->threads; i++) s[i] = o->lcg[i]; }

// sampling() with the thread bodies run one after the other (their streams are independent);
// Base.cpp:185-264, mode 0, negRelRate 0
void oracle_sampling(Oracle* o, INT* batch_h, INT* batch_t, INT* batch_r, REAL* batch_y, INT batchSize, INT negRate, bool filter) {
    Index& x = cur(o);
    for (INT id = 0; id < o->threads; id++) {
        INT lef, rig;
        if (batchSize % o->threads == 0) { lef = id * (batchSize / o->threads); rig = (id + 1) * (batchSize / o->threads); }
        else { lef = id * (batchSize / o->threads + 1); rig = (id + 1) * (batchSize / o->threads + 1); if (rig > batchSize) rig = batchSize; }
        REAL prob = 500;
        for (INT batch = lef; batch < rig; batch++) {
            INT i = rand_max(o, id, x.nT);
            batch_h[batch] = x.list[i].h; batch_t[batch] = x.list[i].t; batch_r[batch] = x.list[i].r;
            if (batch_y) batch_y[batch] = 1;
            INT last = batchSize;
            for (INT times = 0; times < negRate; times++) {
                if (o->bern) prob = 1000 * x.right_mean[x.list[i].r] / (x.right_mean[x.list[i].r] + x.left_mean[x.list[i].r]);
                if (randd(o, id) % 1000 < prob) {
                    batch_h[batch + last] = x.list[i].h;
                    batch_t[batch + last] = corrupt(o, id, x.list[i].h, x.list[i].r, filter, true);
                } else {
                    batch_h[batch + last] = corrupt(o, id, x.list[i].t, x.list[i].r, filter, false);
                    batch_t[batch + last] = x.list[i].t;
                }
                batch_r[batch + last] = x.list[i].r;
                if (batch_y) batch_y[batch + last] = -1;
                last += batchSize;
            }
        }
    }
}

// getParallelUniverse (UniverseConstructor.h:327-397), libc rand() continuing from oracle_seed
INT oracle_universe(Oracle* o, INT tc, REAL balance) {
    Index& g = o->g;
    INT nT = tc;
    INT focus = rand_ab(0, g.nR);
    INT threshold = balance * tc;
    std::set<INT> entity_set;
    for (INT i = g.lefRel2[focus]; i < g.rigRel2[focus] + 1; i++) { entity_set.insert(g.rel2[i].h); entity_set.insert(g.rel2[i].t); }
    if ((INT)entity_set.size() > threshold) {  // get_entity_subset :55-67
        std::set<INT> sub;
        while ((INT)sub.size() < threshold) {
            std::set<INT>::iterator it = entity_set.begin();
            std::advance(it, rand() % entity_set.size());
            sub.insert(*it);
            entity_set.erase(it);
        }
        entity_set = sub;
    }
    // BidirectionalRandomWalk :92-191
    o->u_global.assign(nT, T3{0, 0, 0});
    INT universe_index = 0, last_dup = -1, dup_tol = 5, stall_tol = 20, last_size = 0;
    std::set<INT> next_points, ents, rels;
    REAL prob = 500;
    while (universe_index < nT) {
        for (std::set<INT>::iterator it = entity_set.begin(); it != entity_set.end() && universe_index < nT;) {
            INT e = *it, nh = 0, nr = 0, nt = 0, nxt = -1;
            bool from_head;
            if (rand() % 1000 < prob) from_head = g.rigHead[e] != -1 ? true : false;
            else from_head = g.rigTail[e] != -1 ? false : true;
            if (from_head) {
                INT idx = rand_ab(g.lefHead[e], g.rigHead[e] + 1);
                nh = g.head[idx].h; nr = g.head[idx].r; nt = g.head[idx].t; nxt = nt;
            } else {
                INT idx = rand_ab(g.lefTail[e], g.rigTail[e] + 1);
                nh = g.tail[idx].h; nr = g.tail[idx].r; nt = g.tail[idx].t; nxt = nh;
            }
            bool dup = false;
            for (INT i = 0; i < universe_index; i++)
                if (nh == o->u_global[i].h && nr == o->u_global[i].r && nt == o->u_global[i].t) { dup = true; break; }
            if (dup) {
                if (last_dup == e) dup_tol--; else last_dup = e;
                if (dup_tol == 0) { dup_tol = 5; it++; }
                continue;
            }
            o->u_global[universe_index] = T3{nh, nr, nt};
            next_points.insert(nxt); ents.insert(nt); ents.insert(nh); rels.insert(nr);
            entity_set.erase(it++);
            universe_index++;
        }
        entity_set.swap(next_points);
        if (universe_index == last_size) stall_tol--; else { last_size = universe_index; stall_tol = 20; }
        if (stall_tol == 0) { nT = universe_index; break; }
    }
    o->u_global.resize(nT);
    // enumerateTrainUniverseTriples :193-233
    Index& u = o->u;
    u = Index();
    u.nE = (INT)ents.size(); u.nR = (INT)rels.size();
    std::vector<INT> em(g.nE, -1), rm(g.nR, -1);
    o->ent_remap.clear(); o->rel_remap.clear();
    u.list.resize(nT);
    for (INT i = 0; i < nT; i++) {
        const T3& x = o->u_global[i];
        if (em[x.h] == -1) { em[x.h] = (INT)o->ent_remap.size(); o->ent_remap.push_back(x.h); }
        if (em[x.t] == -1) { em[x.t] = (INT)o->ent_remap.size(); o->ent_remap.push_back(x.t); }
        if (rm[x.r] == -1) { rm[x.r] = (INT)o->rel_remap.size(); o->rel_remap.push_back(x.r); }
        u.list[i] = T3{em[x.h], rm[x.r], em[x.t]};
    }
    load_helpers(u, false, 1);  // loadUniverseHelpers :235-325 (fresh arrays)
    return nT;
}
void oracle_universe_export(Oracle* o, INT* global_hrt, INT* ent_remap, INT* rel_remap) {
    for (size_t i = 0; i < o->u_global.size(); i++) { global_hrt[3 * i] = o->u_global[i].h; global_hrt[3 * i + 1] = o->u_global[i].r; global_hrt[3 * i + 2] = o->u_global[i].t; }
    for (size_t i = 0; i < o->ent_remap.size(); i++) ent_remap[i] = o->ent_remap[i];
    for (size_t i = 0; i < o->rel_remap.size(); i++) rel_remap[i] = o->rel_remap[i];
}
INT oracle_universe_ent(Oracle* o) { return (INT)o->ent_remap.size(); }
INT oracle_universe_rel(Oracle* o) { return (INT)o->rel_remap.size(); }
void oracle_swap(Oracle* o) { o->swapped = !o->swapped; }

// importTestFiles on arrays (Reader.h:246-342): test/valid sorted (r,h,t), tripleList = all sorted (h,r,t)
void oracle_import_test(Oracle* o, const INT* test_htr, INT nt, const INT* train_htr, INT ntr, const INT* valid_htr, INT nv) {
    auto conv = [](const INT* a, INT n, std::vector<T3>& out) { for (INT i = 0; i < n; i++) out.push_back(T3{a[3 * i], a[3 * i + 2], a[3 * i + 1]}); };
    o->testList.clear(); o->validList.clear(); o->tripleList.clear();
    conv(test_htr, nt, o->testList); conv(valid_htr, nv, o->validList);
    conv(test_htr, nt, o->tripleList); conv(train_htr, ntr, o->tripleList); conv(valid_htr, nv, o->tripleList);
    std::sort(o->tripleList.begin(), o->tripleList.end(), cmp_head);
    std::sort(o->testList.begin(), o->testList.end(), cmp_rel2);
    std::sort(o->validList.begin(), o->validList.end(), cmp_rel2);
}
void oracle_eval_list(Oracle* o, int which, INT* hrt) {
    const std::vector<T3>& q = which == 0 ? o->testList : o->validList;
    for (size_t i = 0; i < q.size(); i++) { hrt[3 * i] = q[i].h; hrt[3 * i + 1] = q[i].r; hrt[3 * i + 2] = q[i].t; }
}

static bool find_triple(Oracle* o, INT h, INT t, INT r) {  // Corrupt.h:188-199
    const std::vector<T3>& L = o->tripleList;
    INT lef = 0, rig = (INT)L.size() - 1;
    while (lef + 1 < rig) {
        INT mid = (lef + rig) >> 1;
        if ((L[mid].h < h) || (L[mid].h == h && L[mid].r < r) || (L[mid].h == h && L[mid].r == r && L[mid].t < t)) lef = mid; else rig = mid;
    }
    if (L[lef].h == h && L[lef].r == r && L[lef].t == t) return true;
    if (L[rig].h == h && L[rig].r == r && L[rig].t == t) return true;
    return false;
}

// testHead (head=1) / testTail (head=0) on a candidate-ordered score row (Test.h:118-238,240-359):
// out[0] = raw count, out[1] = filtered count.
void oracle_rank_row(Oracle* o, int which, const REAL* con, INT index, int head, INT nE, INT* out) {
    const T3& x = (which == 0 ? o->testList : o->validList)[index];
    INT offset = -1, s = 0, fs = 0;
    REAL minimal = con[0];
    const INT truth = head ? x.h : x.t;
    if (minimal != INFINITY) {
        for (INT j = 1; j < nE; j++) {
            REAL value = con[j];
            if (j + offset == truth) offset++;
            if (value < minimal) {
                s += 1;
                if (!(head ? find_triple(o, j + offset, x.t, x.r) : find_triple(o, x.h, j + offset, x.r))) fs += 1;
            }
        }
    } else {
        s = nE; fs = nE;
        for (INT j = 1; j < nE; j++) {
            if (j + offset == truth) offset++;
            if (head ? find_triple(o, j + offset, x.t, x.r) : find_triple(o, x.h, j + offset, x.r)) fs -= 1;
        }
    }
    out[0] = s; out[1] = fs;
}

}  // extern "C"
