// Universe subgraphs built ON THE GPU: the reference's getParallelUniverse (openke/base/UniverseConstructor.h:39-67,
// 92-233,327-397) for a whole chunk of universes at once (k_walk_universes: phases A-C, k_number_universes: phase D), bit-identical to the host builder
// (graph_host.cpp Graph::build_universe, itself pinned to the reference) — subgraph sampling is integer work.
//
// Why: with the training kernel at ~5 ms per 100 universes, the bit-exact glibc-rand() walk on the host cores
// (0.35-0.55 ms per universe and core) bounds the end-to-end rate as soon as several ranks share one host
// (SURVEY.md 8(f) rank 1).  The walk of ONE universe is sequential (every draw depends on what the previous draws
// collected), but universes are independent: one warp per universe, a few hundred universes in flight.
//
//   phase A  lane 0 replays srand(seed), the randReset() draws and the focus draw               (Random.h:11-15,38-45)
//   phase B  subset pick (UniverseConstructor.h:55-67): the k-th REMAINING entity of the focus relation's entity
//            list, `threshold` times.  The list is a 32-ary tree of counts whose groups of 32 siblings hold
//            inclusive prefix sums: a level of the descent is one load + one ballot, the removal decrements the
//            lanes behind the chosen child.  The picked set in ascending order = the cleared bits, enumerated.
//   phase C  the bidirectional walk (:92-191), lane 0: coin, random incident triple, collected-before test in a
//            shared-memory hash set keyed by the triple's position in the (h,r,t) order.  The reference's two
//            std::sets of starting points are two bitmaps over the entities that swap roles every round: next
//            round's starting points and the skipped ("resurface a round later") ones are OR-ed in fire-and-forget,
//            and a round starts by enumerating + clearing its bitmap (ascending order, duplicates gone, no sort); one
//            summary bit per bitmap word keeps a round with three starting points from scanning the whole bitmap.
//   phase D  local ids by first appearance (:193-233) without a map over all entities: sort (id, position) keys,
//            flag the first position of every id, local id = number of flags before it.  Then the (h,r,t) order
//            of the local triples as one sort of packed keys (:236).
//
// Only what unfiltered, non-Bernoulli training reads is produced (the "lean" universe of graph_host.hpp): the local
// (h,r,t)-sorted list and the two remaps.  Universes the kernel does not handle (more than PK_WALK_CAP triples,
// entities without any triple, a training list with duplicates) are reported with a status and left to the host.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.hpp"
#include "global_state.hpp"

namespace {

constexpr int CAP = PK_WALK_CAP;          // triples per universe (2048)
constexpr int HASH_SLOTS = 2 * CAP;       // collected-before set, open addressing, load <= 0.5
constexpr int MAX_LEVELS = 7;
// Two kernels, one warp (= one block) per universe each.  The walk (phases A-C) is one lane chasing L2 latency for a
// millisecond, so its blocks are kept small (24 KB: nine per SM) — every SM they sit on is lost to the training kernel,
// whose blocks fill an SM's shared memory.  The numbering (phase D) is short and needs the sort buffer.
constexpr int SM_X_BYTES = CAP * 4;                   // walk: starting points of the current round
constexpr int SM_TREE_BYTES = HASH_SLOTS * 4;         // walk: pick tree, then the collected-before hash set
constexpr int SM_RNG = SM_X_BYTES + SM_TREE_BYTES;    // walk: uint32 [32] generator state
constexpr int SM_WALK_TOTAL = SM_RNG + 32 * 4;
constexpr int SM_SORT_BYTES = 2 * CAP * 8;            // numbering: 2*CAP 64-bit keys
constexpr int SM_LOC = SM_SORT_BYTES;                 // uint16 [2*CAP] local entity id per occurrence
constexpr int SM_RLOC = SM_LOC + 2 * CAP * 2;         // uint16 [CAP]   local relation id per triple
constexpr int SM_FLAGS = SM_RLOC + CAP * 2;           // uint32 [2*CAP/32] first-appearance flags
constexpr int SM_PRE = SM_FLAGS + (2 * CAP / 32) * 4; // uint32 [2*CAP/32] flags before each word
constexpr int SM_NUMBER_TOTAL = SM_PRE + (2 * CAP / 32) * 4;

struct UniIn {            // per universe, filled by the host
    uint32_t seed;
    int32_t tc, threshold, focus, skip;   // skip = draws before the subset pick (randReset + focus)
    int32_t n_focus;
    int64_t focus_off;    // into rel_ent
    int64_t tree_off;     // words into the tree scratch, -1: the tree fits shared memory
};

struct WalkArgs {
    const int4* ent_range; const int2* head_rt; const int4* tail_rht; const int32_t* rel_ent;
    const UniIn* in;
    uint32_t* q;          // [n][2][qwords + swords] starting-point bitmaps + one summary bit per word (zero on entry)
    int32_t qwords, swords;
    uint32_t* tree;       // scratch for pick trees that do not fit shared memory
    int32_t* got;         // [n][CAP][3] collected triples, global ids, collection order
    int32_t* tri;         // [n][CAP][3] local ids sorted (h,r,t)
    int32_t* ent_remap;   // [n][2*CAP]
    int32_t* rel_remap;   // [n][CAP]
    int32_t* sizes;       // [n][8] nT nE nR focus draws status rounds -
};

// ---- glibc TYPE_3 generator, state in shared memory, used by lane 0 only
struct Rng {
    uint32_t* r; int f, b;
    __device__ __forceinline__ int32_t next() {
        const uint32_t v = r[f] + r[b];
        r[f] = v;
        f = f == 30 ? 0 : f + 1;
        b = b == 30 ? 0 : b + 1;
        return (int32_t)(v >> 1);
    }
    __device__ void reseed(uint32_t seed) {
        if (seed == 0) seed = 1;
        int32_t word = (int32_t)seed;
        r[0] = (uint32_t)word;
        for (int i = 1; i < 31; ++i) {
            const long long hi = word / 127773, lo = word % 127773;
            long long w = 16807 * lo - 2836 * hi;
            if (w < 0) w += 2147483647;
            word = (int32_t)w;
            r[i] = (uint32_t)word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; ++i) (void)next();
    }
};

__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) { return __ldcg(p); }

// ---- one warp sorts M (power of two) 64-bit keys in shared memory, ascending
__device__ void bitonic_sort(unsigned long long* a, int M, int lane) {
    for (int k = 2; k <= M; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (M >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long x = a[i], y = a[p];
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            __syncwarp();
        }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v = max(v, o);
    }
    return v;
}

// Set bits of bitmap q[0..words) in ascending order -> out[], bitmap cleared; returns the count (<= cap, else -1).
// `value`: nullptr = the bit index itself, else value[bit index].
__device__ int enumerate_clear(uint32_t* q, int words, int32_t* out, int cap, const int32_t* value, bool global_mem, int lane) {
    int base = 0;
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int w = w0 + lane;
        uint32_t m = 0;
        if (w < words) m = global_mem ? ld_cg(q + w) : q[w];
        if (__ballot_sync(0xffffffffu, m != 0) == 0) continue;
        const int c = __popc(m);
        const int incl = warp_incl_scan(c, lane);
        int o = base + incl - c;
        base += __shfl_sync(0xffffffffu, incl, 31);
        if (base > cap) return -1;
        if (m) {
            q[w] = 0;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const int idx = w * 32 + j;
                out[o++] = value ? value[idx] : idx;
            }
        }
    }
    __syncwarp();
    return base;
}

// A starting-point set: bitmap over the entities [qwords] followed by a summary [swords] with one bit per bitmap word, so
// that a round with a handful of starting points (relation-rich graphs: a focus relation with three entities walks in
// hundreds of rounds) does not scan the whole bitmap.
__device__ __forceinline__ void set_insert(uint32_t* q, int qwords, int e) {
    const int w = e >> 5;
    atomicOr(&q[w], 1u << (e & 31));
    atomicOr(&q[qwords + (w >> 5)], 1u << (w & 31));
}
// members in ascending order -> out[], set emptied; returns the count (<= cap, else -1)
__device__ int set_enumerate_clear(uint32_t* q, int qwords, int swords, int32_t* out, int cap, int lane) {
    uint32_t* S = q + qwords;
    int base = 0;
    for (int s0 = 0; s0 < swords; s0 += 32) {
        const int si = s0 + lane;
        const uint32_t sv = si < swords ? ld_cg(S + si) : 0u;
        uint32_t nz = __ballot_sync(0xffffffffu, sv != 0);
        if (sv) S[si] = 0;
        while (nz) {
            const int sl = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t group = __shfl_sync(0xffffffffu, sv, sl);   // the non-empty words among the 32 this summary word covers
            const int w = (s0 + sl) * 32 + lane;
            uint32_t m = ((group >> lane) & 1u) ? ld_cg(q + w) : 0u;
            const int c = __popc(m);
            const int incl = warp_incl_scan(c, lane);
            int o = base + incl - c;
            base += __shfl_sync(0xffffffffu, incl, 31);
            if (base > cap) return -1;
            if (m) {
                q[w] = 0;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    out[o++] = w * 32 + j;
                }
            }
        }
    }
    __syncwarp();
    return base;
}

// First-appearance numbering (UniverseConstructor.h:193-233) of the ids in the sorted keys a[0..cnt) = (id << 32 | position):
// loc[position] = local id, remap[local id] = id; returns the number of distinct ids.
__device__ int number_by_first_appearance(const unsigned long long* a, int cnt, unsigned short* loc, uint32_t* flags, uint32_t* pre,
                                          int flag_words, int32_t* remap, int lane) {
    for (int w = lane; w < flag_words; w += 32) flags[w] = 0;
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
        const uint32_t id = (uint32_t)(a[i] >> 32);
        if (i == 0 || (uint32_t)(a[i - 1] >> 32) != id) {
            const uint32_t p = (uint32_t)a[i];
            atomicOr(&flags[p >> 5], 1u << (p & 31));
        }
    }
    __syncwarp();
    int carry = 0;
    for (int w0 = 0; w0 < flag_words; w0 += 32) {
        const int w = w0 + lane;
        const int c = w < flag_words ? __popc(flags[w]) : 0;
        const int incl = warp_incl_scan(c, lane);
        if (w < flag_words) pre[w] = (uint32_t)(carry + incl - c);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    int run_head = -1;
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < cnt;
        uint32_t id = 0;
        bool head = false;
        if (valid) {
            id = (uint32_t)(a[i] >> 32);
            head = i == 0 || (uint32_t)(a[i - 1] >> 32) != id;
        }
        int h = warp_incl_max(head ? i : -1, lane);
        if (h < 0) h = run_head;
        run_head = __shfl_sync(0xffffffffu, h, 31);
        if (valid) {
            const uint32_t fp = (uint32_t)a[h];
            const int local = (int)pre[fp >> 5] + __popc(flags[fp >> 5] & ((1u << (fp & 31)) - 1u));
            loc[(uint32_t)a[i]] = (unsigned short)local;
            if (head) remap[local] = (int32_t)id;
        }
    }
    __syncwarp();
    return carry;
}

__global__ void __launch_bounds__(32) k_walk_universes(WalkArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int u = blockIdx.x, lane = threadIdx.x;
    const UniIn in = A.in[u];
    int32_t* X = reinterpret_cast<int32_t*>(smem);                              // [CAP]
    int32_t* hash = reinterpret_cast<int32_t*>(smem + SM_X_BYTES);              // [HASH_SLOTS] (after the tree is done)
    Rng rng{reinterpret_cast<uint32_t*>(smem + SM_RNG), 3, 0};
    int32_t* sizes = A.sizes + (size_t)u * 8;
    int draws = 0, status = PK_WALK_OK;
    if (in.tc > CAP || in.threshold > CAP) {   // the host builder's case
        if (lane == 0) { sizes[0] = 0; sizes[1] = 0; sizes[2] = 0; sizes[3] = in.focus; sizes[4] = 0; sizes[5] = PK_WALK_TOO_LARGE; sizes[6] = 0; sizes[7] = 0; }
        return;
    }

    // ---- phase A
    if (lane == 0) {
        rng.reseed(in.seed);
        for (int i = 0; i < in.skip; ++i) (void)rng.next();
    }
    draws = in.skip;
    __syncwarp();

    // ---- phase B: starting points
    const int32_t* focus_ent = A.rel_ent + in.focus_off;
    const int n = in.n_focus;
    int fsz;
    if (in.threshold >= 0 && n > in.threshold) {
        const bool tree_global = in.tree_off >= 0;
        uint32_t* T = tree_global ? A.tree + in.tree_off : reinterpret_cast<uint32_t*>(smem + SM_X_BYTES);
        // level 0 = presence bits [W0]; level l >= 1 = N_l prefix counts, N_1 = W0, N_{l+1} = ceil(N_l / 32)
        int off[MAX_LEVELS], cntl[MAX_LEVELS], L = 1;
        const int W0 = (n + 31) >> 5;
        off[0] = 0; cntl[0] = W0;
        off[1] = W0; cntl[1] = W0;
        while (cntl[L] > 32) { off[L + 1] = off[L] + cntl[L]; cntl[L + 1] = (cntl[L] + 31) >> 5; ++L; }
        for (int w = lane; w < W0; w += 32) T[w] = (w == W0 - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
        {
            long long span = 32;
            for (int l = 1; l <= L; ++l, span *= 32)
                for (int e = lane; e < cntl[l]; e += 32) {
                    const long long hi = min((long long)n, (long long)(e + 1) * span), lo = (long long)(e & ~31) * span;
                    T[off[l] + e] = (uint32_t)(hi - lo);
                }
        }
        __syncwarp();
        int remaining = n;
        for (int pick = 0; pick < in.threshold; ++pick) {
            int k = 0;
            if (lane == 0) k = (int)((uint32_t)rng.next() % (uint32_t)remaining);
            k = __shfl_sync(0xffffffffu, k, 0);
            int g = 0;
            for (int l = L; l >= 1; --l) {
                const int idx = g * 32 + lane;
                uint32_t* slot = T + off[l] + idx;
                const bool valid = idx < cntl[l];
                const uint32_t v = valid ? *slot : 0xffffffffu;
                const uint32_t bal = __ballot_sync(0xffffffffu, v > (uint32_t)k);
                const int child = __ffs(bal) - 1;
                const uint32_t below = __shfl_sync(0xffffffffu, v, child > 0 ? child - 1 : 0);
                if (child > 0) k -= (int)below;
                if (valid && lane >= child) *slot = v - 1;
                g = g * 32 + child;
            }
            const uint32_t word = T[g];
            const bool mine = ((word >> lane) & 1u) && __popc(word & ((1u << lane) - 1u)) == k;
            const int j = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            if (lane == 0) T[g] = word & ~(1u << j);
            --remaining;
            __syncwarp();
        }
        draws += in.threshold;
        // the picked entities, ascending = the cleared bits
        for (int w = lane; w < W0; w += 32) {
            const uint32_t full = (w == W0 - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
            T[w] = full & ~T[w];
        }
        __syncwarp();
        fsz = enumerate_clear(T, W0, X, CAP, focus_ent, false, lane);   // plain loads: only this warp wrote T
    } else {
        fsz = n <= CAP ? n : -1;
        for (int i = lane; i < n && fsz >= 0; i += 32) X[i] = focus_ent[i];
    }
    if (fsz < 0) status = PK_WALK_TOO_LARGE;
    __syncwarp();

    // ---- phase C: the walk
    for (int i = lane; i < HASH_SLOTS; i += 32) hash[i] = -1;
    __syncwarp();
    const int qstride = A.qwords + A.swords;
    uint32_t* Q0 = A.q + (size_t)u * 2 * qstride;
    int32_t* got = A.got + (size_t)u * CAP * 3;
    int ngot = 0, target = in.tc, rounds = 0;
    int last_dup = -1, dup_tol = 5, stall_tol = 20, last_size = 0;
    while (status == PK_WALK_OK && ngot < target) {
        uint32_t* Qnext = Q0 + (size_t)((rounds + 1) & 1) * qstride;   // next round's starting points
        uint32_t* Qlate = Q0 + (size_t)(rounds & 1) * qstride;         // skipped ones: the round after next
        int i = 0;
        if (lane == 0) {
            int4 er = fsz > 0 ? A.ent_range[X[0]] : make_int4(0, -1, 0, -1);
            while (i < fsz && ngot < target) {
                const int e = X[i];
                int4 er_next = make_int4(0, -1, 0, -1);
                if (i + 1 < fsz) er_next = A.ent_range[X[i + 1]];        // in flight while this entity is worked on
                bool advance = false;
                while (!advance && ngot < target) {
                    const bool head_first = ((uint32_t)rng.next() % 1000u) < 500u;   // :122
                    ++draws;
                    const bool has_h = er.y != -1, has_t = er.w != -1;
                    const int side = head_first ? (has_h ? 0 : (has_t ? 1 : -1)) : (has_t ? 1 : (has_h ? 0 : -1));
                    if (side < 0) { status = PK_WALK_ISOLATED; break; }
                    int tid, h, r, t, nxt;
                    if (side == 0) {
                        const int idx = (int)((uint32_t)rng.next() % (uint32_t)(er.y + 1 - er.x)) + er.x;   // :40
                        const int2 rt = A.head_rt[idx];
                        tid = idx; h = e; r = rt.x; t = rt.y; nxt = t;
                    } else {
                        const int idx = (int)((uint32_t)rng.next() % (uint32_t)(er.w + 1 - er.z)) + er.z;   // :48
                        const int4 rh = A.tail_rht[idx];
                        tid = rh.z; h = rh.y; r = rh.x; t = e; nxt = h;
                    }
                    ++draws;
                    // collected before? (:141-154)
                    uint32_t s = ((uint32_t)tid * 2654435761u) >> (32 - 12);
                    static_assert(HASH_SLOTS == 4096, "hash shift assumes 4096 slots");
                    bool dup = false;
                    for (;;) {
                        const int v = hash[s];
                        if (v == tid) { dup = true; break; }
                        if (v < 0) break;
                        s = (s + 1) & (HASH_SLOTS - 1);
                    }
                    if (dup) {
                        if (last_dup == e) --dup_tol; else last_dup = e;
                        if (dup_tol == 0) {
                            dup_tol = 5;
                            set_insert(Qlate, A.qwords, e);
                            advance = true;
                        }
                        continue;
                    }
                    hash[s] = tid;
                    got[3 * ngot + 0] = h; got[3 * ngot + 1] = r; got[3 * ngot + 2] = t;
                    ++ngot;
                    set_insert(Qnext, A.qwords, nxt);
                    advance = true;
                }
                if (status != PK_WALK_OK) break;
                if (advance) { ++i; er = er_next; }
            }
            __threadfence();
        }
        __syncwarp();
        i = __shfl_sync(0xffffffffu, i, 0);
        ngot = __shfl_sync(0xffffffffu, ngot, 0);
        status = __shfl_sync(0xffffffffu, status, 0);
        if (status != PK_WALK_OK) break;
        // the starting points this round did not reach stay for the round after next (:112-113,170)
        for (int k2 = i + lane; k2 < fsz; k2 += 32) set_insert(Qlate, A.qwords, X[k2]);
        __threadfence();
        __syncwarp();
        fsz = set_enumerate_clear(Qnext, A.qwords, A.swords, X, CAP, lane);
        if (fsz < 0) { status = PK_WALK_TOO_LARGE; break; }
        ++rounds;
        if (ngot == last_size) --stall_tol;
        else { last_size = ngot; stall_tol = 20; }
        if (stall_tol == 0) { target = ngot; break; }   // :181-186
    }
    draws = __shfl_sync(0xffffffffu, draws, 0);
    if (status == PK_WALK_OK && ngot == 0) status = PK_WALK_EMPTY;
    if (status != PK_WALK_OK) ngot = 0;
    if (lane == 0) { sizes[0] = ngot; sizes[1] = 0; sizes[2] = 0; sizes[3] = in.focus; sizes[4] = draws; sizes[5] = status; sizes[6] = rounds; sizes[7] = 0; }
}

// ---- phase D: local ids by first appearance, entities (h then t of every triple in collection order), relations, and
//      the (h,r,t) order of the local list
__global__ void __launch_bounds__(32) k_number_universes(WalkArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int u = blockIdx.x, lane = threadIdx.x;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    unsigned short* LOC = reinterpret_cast<unsigned short*>(smem + SM_LOC);
    unsigned short* RLOC = reinterpret_cast<unsigned short*>(smem + SM_RLOC);
    uint32_t* flags = reinterpret_cast<uint32_t*>(smem + SM_FLAGS);
    uint32_t* pre = reinterpret_cast<uint32_t*>(smem + SM_PRE);
    int32_t* sizes = A.sizes + (size_t)u * 8;
    const int ngot = sizes[0];
    if (sizes[5] != PK_WALK_OK || ngot <= 0) return;
    const int32_t* got = A.got + (size_t)u * CAP * 3;
    const int nocc = 2 * ngot;
    int M = 2;
    while (M < nocc) M <<= 1;
    for (int p = lane; p < M; p += 32) {
        unsigned long long key = ~0ull;
        if (p < nocc) key = ((unsigned long long)(uint32_t)got[3 * (p >> 1) + ((p & 1) ? 2 : 0)] << 32) | (unsigned long long)p;
        keys[p] = key;
    }
    __syncwarp();
    bitonic_sort(keys, M, lane);
    const int nE = number_by_first_appearance(keys, nocc, LOC, flags, pre, 2 * CAP / 32, A.ent_remap + (size_t)u * 2 * CAP, lane);
    // relations
    M = 2;
    while (M < ngot) M <<= 1;
    for (int p = lane; p < M; p += 32)
        keys[p] = p < ngot ? (((unsigned long long)(uint32_t)got[3 * p + 1] << 32) | (unsigned long long)p) : ~0ull;
    __syncwarp();
    bitonic_sort(keys, M, lane);
    const int nR = number_by_first_appearance(keys, ngot, RLOC, flags, pre, CAP / 32, A.rel_remap + (size_t)u * CAP, lane);
    // the (h,r,t) order of the local list (:236)
    for (int p = lane; p < M; p += 32)
        keys[p] = p < ngot ? (((unsigned long long)LOC[2 * p] << 42) | ((unsigned long long)RLOC[p] << 21) | (unsigned long long)LOC[2 * p + 1]) : ~0ull;
    __syncwarp();
    bitonic_sort(keys, M, lane);
    int32_t* tri = A.tri + (size_t)u * CAP * 3;
    for (int p = lane; p < ngot; p += 32) {
        const unsigned long long v = keys[p];
        tri[3 * p + 0] = (int32_t)(v >> 42);
        tri[3 * p + 1] = (int32_t)((v >> 21) & 0x1fffff);
        tri[3 * p + 2] = (int32_t)(v & 0x1fffff);
    }
    if (lane == 0) { sizes[1] = nE; sizes[2] = nR; }
}

// ---- the two remaps of a chunk back to back (what evaluation indexes), universes in order; block u finds its offsets
//      by summing the sizes of the universes before it
__global__ void __launch_bounds__(128) k_pack_remaps(int n, const int32_t* sizes, const int32_t* ent_remap, const int32_t* rel_remap,
                                                      int32_t* ent_out, int32_t* rel_out) {
    __shared__ long long part[2][4];
    const int u = blockIdx.x, tid = threadIdx.x;
    long long eo = 0, ro = 0;
    for (int v = tid; v < u; v += 128) { eo += sizes[v * 8 + 1]; ro += sizes[v * 8 + 2]; }
    for (int d = 16; d > 0; d >>= 1) {
        eo += __shfl_down_sync(0xffffffffu, eo, d);
        ro += __shfl_down_sync(0xffffffffu, ro, d);
    }
    if ((tid & 31) == 0) { part[0][tid >> 5] = eo; part[1][tid >> 5] = ro; }
    __syncthreads();
    eo = part[0][0] + part[0][1] + part[0][2] + part[0][3];
    ro = part[1][0] + part[1][1] + part[1][2] + part[1][3];
    const int nE = sizes[u * 8 + 1], nR = sizes[u * 8 + 2];
    for (int i = tid; i < nE; i += 128) ent_out[eo + i] = ent_remap[(size_t)u * 2 * CAP + i];
    for (int i = tid; i < nR; i += 128) rel_out[ro + i] = rel_remap[(size_t)u * CAP + i];
}

// ------------------------------------------------------------------------------------------------ host side
struct DevGraph {
    uint64_t version = 0;
    int device = -1;
    int4* ent_range = nullptr; int2* head_rt = nullptr; int4* tail_rht = nullptr; int32_t* rel_ent = nullptr;
};
// per-universe kernel inputs travel through a small ring of pinned buffers (a copy from pageable memory would
// synchronise the stream first); a ring slot is reused only after the launch that read it has finished
struct InRing {
    static constexpr int N = 4;
    UniIn* h[N] = {nullptr, nullptr, nullptr, nullptr};
    UniIn* d[N] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[N] = {0, 0, 0, 0};
    cudaEvent_t ev[N] = {nullptr, nullptr, nullptr, nullptr};
    int next = 0;
};
InRing g_ring;
DevGraph g_graph;
std::mutex g_mu;

int upload_graph(cudaStream_t st) {
    const pk::Graph& g = pk::G().graph;
    int dev = 0;
    PK_CUDA(cudaGetDevice(&dev));
    if (g_graph.version == g.version && g_graph.device == dev && g_graph.ent_range) return PK_OK;
    cudaFree(g_graph.ent_range); cudaFree(g_graph.head_rt); cudaFree(g_graph.tail_rht); cudaFree(g_graph.rel_ent);
    g_graph.ent_range = nullptr; g_graph.head_rt = nullptr; g_graph.tail_rht = nullptr; g_graph.rel_ent = nullptr;
    const size_t E = (size_t)g.n_ent, T = g.train.by_head.size();
    std::vector<int2> head_rt(T);
    std::vector<int4> tail_rht(T);
    for (size_t i = 0; i < T; ++i) {
        head_rt[i] = make_int2(g.train.by_head[i].r, g.train.by_head[i].t);
        tail_rht[i] = make_int4(g.train.by_tail[i].r, g.train.by_tail[i].h, g.tail_to_head[i], 0);
    }
    static_assert(sizeof(pk::Graph::EntRange) == sizeof(int4), "EntRange is four int32");
    PK_CUDA(cudaMalloc(&g_graph.ent_range, std::max<size_t>(E, 1) * sizeof(int4)));
    PK_CUDA(cudaMalloc(&g_graph.head_rt, std::max<size_t>(T, 1) * sizeof(int2)));
    PK_CUDA(cudaMalloc(&g_graph.tail_rht, std::max<size_t>(T, 1) * sizeof(int4)));
    PK_CUDA(cudaMalloc(&g_graph.rel_ent, std::max<size_t>(g.rel_ent.size(), 1) * 4));
    PK_CUDA(cudaMemcpyAsync(g_graph.ent_range, g.ent_range.data(), E * sizeof(int4), cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaMemcpyAsync(g_graph.head_rt, head_rt.data(), T * sizeof(int2), cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaMemcpyAsync(g_graph.tail_rht, tail_rht.data(), T * sizeof(int4), cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaMemcpyAsync(g_graph.rel_ent, g.rel_ent.data(), g.rel_ent.size() * 4, cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaStreamSynchronize(st));   // the staging vectors die here
    g_graph.version = g.version;
    g_graph.device = dev;
    return PK_OK;
}

int64_t tree_words(int64_t n) {
    int64_t w = (n + 31) >> 5, total = w, c = w;
    total += c;
    while (c > 32) { c = (c + 31) >> 5; total += c; }
    return total;
}

}  // namespace

extern "C" {

int pk_walk_cap(void) { return CAP; }

// Why the current training graph cannot be walked on the device (0 = it can).
int pk_walk_device_check(void) {
    const pk::Graph& g = pk::G().graph;
    if (g.train.n_tri() == 0) return pk::fail(PK_ERR_STATE, "pk_walk_device: importTrainFiles has not run");
    if (!g.head_canon.empty()) return pk::fail(PK_ERR_UNSUPPORTED, "pk_walk_device: the training list holds duplicates (host builder)");
    if (g.n_ent >= (1LL << 31) - 64 || g.train.n_tri() >= (1LL << 31) - 64) return pk::fail(PK_ERR_UNSUPPORTED, "pk_walk_device: graph too large for 32-bit positions");
    return PK_OK;
}

// Scratch the caller provides for n universes: [0] bytes of the starting-point bitmap area, [1] bytes of the collected-triples area (also an output), [2] bytes for pick trees that do not fit
// shared memory (worst case for this graph).
int pk_walk_scratch_bytes(int n, int64_t* out3) {
    if (!out3 || n < 0) return pk::fail(PK_ERR_ARG, "pk_walk_scratch_bytes: bad argument");
    const pk::Graph& g = pk::G().graph;
    const int64_t qwords = (g.n_ent + 31) / 32, swords = (qwords + 31) / 32;
    int64_t worst = 0;
    for (int64_t r = 0; r < g.n_rel; ++r) {
        const int64_t nf = g.rel_ent_off[(size_t)r + 1] - g.rel_ent_off[(size_t)r];
        if (tree_words(nf) * 4 > SM_TREE_BYTES) worst = std::max(worst, tree_words(nf));
    }
    out3[0] = (int64_t)n * 2 * (qwords + swords) * 4;
    out3[1] = (int64_t)n * CAP * 3 * 4;
    out3[2] = (int64_t)n * worst * 4;
    return PK_OK;
}

// srand(seed + u); randReset(); getParallelUniverse(tc, balance) for n universes in ONE launch on `stream`.
// Host outputs (filled before the call returns): h_lcg [n * work_threads] sampler streams as randReset() leaves
// them, h_focus [n].  Device outputs (valid when the stream reaches the end of the launch): d_tri [n][CAP][3] local
// ids sorted (h,r,t), d_ent_remap [n][2*CAP], d_rel_remap [n][CAP], d_sizes [n][8] = nT nE nR focus draws status
// rounds 0, d_got [n][CAP][3] the collected triples in global ids.  d_bitmaps / d_trees are scratch (pk_walk_scratch_bytes).
int pk_universes_walk_device(int n, const int64_t* seeds, const int64_t* tcs, const float* balances, uint64_t* h_lcg, int64_t* h_focus,
                             uint32_t* d_bitmaps, int32_t* d_got, uint32_t* d_trees, int32_t* d_tri, int32_t* d_ent_remap,
                             int32_t* d_rel_remap, int32_t* d_sizes, void* stream) {
    pk::launch_counter() = 0;
    if (n <= 0 || !seeds || !tcs || !balances || !d_bitmaps || !d_got || !d_tri || !d_ent_remap || !d_rel_remap || !d_sizes)
        return pk::fail(PK_ERR_ARG, "pk_universes_walk_device: null argument");
    if (int rc = pk_walk_device_check()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(g_mu);
    const pk::Graph& g = pk::G().graph;
    if (int rc = upload_graph(st)) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        PK_CUDA(cudaFuncSetAttribute(k_walk_universes, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_WALK_TOTAL));
        PK_CUDA(cudaFuncSetAttribute(k_number_universes, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_NUMBER_TOTAL));
        attr_done = true;
    }
    const int slot = g_ring.next;
    g_ring.next = (g_ring.next + 1) % InRing::N;
    if (!g_ring.ev[slot]) PK_CUDA(cudaEventCreateWithFlags(&g_ring.ev[slot], cudaEventDisableTiming));
    else PK_CUDA(cudaEventSynchronize(g_ring.ev[slot]));
    if (g_ring.cap[slot] < (size_t)n) {
        if (g_ring.h[slot]) cudaFreeHost(g_ring.h[slot]);
        if (g_ring.d[slot]) cudaFree(g_ring.d[slot]);
        g_ring.h[slot] = nullptr; g_ring.d[slot] = nullptr; g_ring.cap[slot] = 0;
        const size_t want = (size_t)n + (size_t)n / 2 + 16;
        PK_CUDA(cudaHostAlloc(&g_ring.h[slot], want * sizeof(UniIn), cudaHostAllocDefault));
        PK_CUDA(cudaMalloc(&g_ring.d[slot], want * sizeof(UniIn)));
        g_ring.cap[slot] = want;
    }
    UniIn* in = g_ring.h[slot];
    const int64_t wt = std::min<int64_t>(g.work_threads, 64);
    int64_t tree_off = 0;
    for (int i = 0; i < n; ++i) {
        UniIn& x = in[(size_t)i];
        if (tcs[i] <= 0) return pk::fail(PK_ERR_ARG, "pk_universes_walk_device: triple_constraint must be positive");
        pk::GlibcRand rng((uint32_t)seeds[i]);
        for (int64_t k = 0; k < wt; ++k) {
            const uint64_t v = (uint64_t)(int64_t)rng.next();
            if (h_lcg) h_lcg[(size_t)i * (size_t)wt + (size_t)k] = v;
        }
        int64_t focus;
        if (g.incremental) {
            if (g.train_rel_contained.empty()) return pk::fail(PK_ERR_STATE, "pk_universes_walk_device: the incremental training list holds no relation");
            focus = g.train_rel_contained[(size_t)rng.range(0, (int64_t)g.train_rel_contained.size())];
        } else {
            focus = rng.range(0, g.n_rel);
        }
        if (h_focus) h_focus[i] = focus;
        const int64_t threshold = (int64_t)(balances[i] * (float)tcs[i]);
        const int64_t nf = g.rel_ent_off[(size_t)focus + 1] - g.rel_ent_off[(size_t)focus];
        x.seed = (uint32_t)seeds[i];
        x.tc = (int32_t)std::min<int64_t>(tcs[i], CAP + 1);          // > CAP: reported as too large by the kernel
        x.threshold = (int32_t)std::max<int64_t>(std::min<int64_t>(threshold, CAP + 1), -1);
        x.focus = (int32_t)focus;
        x.skip = (int32_t)wt + 1;
        x.n_focus = (int32_t)nf;
        x.focus_off = g.rel_ent_off[(size_t)focus];
        x.tree_off = -1;
        if (threshold >= 0 && nf > threshold && tree_words(nf) * 4 > SM_TREE_BYTES) {
            if (!d_trees) return pk::fail(PK_ERR_ARG, "pk_universes_walk_device: tree scratch needed (pk_walk_scratch_bytes)");
            x.tree_off = tree_off;
            tree_off += tree_words(nf);
        }
    }
    PK_CUDA(cudaMemcpyAsync(g_ring.d[slot], in, (size_t)n * sizeof(UniIn), cudaMemcpyHostToDevice, st));
    const int64_t qwords = (g.n_ent + 31) / 32, swords = (qwords + 31) / 32;
    PK_CUDA(cudaMemsetAsync(d_bitmaps, 0, (size_t)n * 2 * (size_t)(qwords + swords) * 4, st));
    WalkArgs A;
    A.ent_range = g_graph.ent_range; A.head_rt = g_graph.head_rt; A.tail_rht = g_graph.tail_rht; A.rel_ent = g_graph.rel_ent;
    A.in = g_ring.d[slot];
    A.q = d_bitmaps; A.qwords = (int32_t)qwords; A.swords = (int32_t)swords;
    A.tree = d_trees; A.got = d_got; A.tri = d_tri; A.ent_remap = d_ent_remap; A.rel_remap = d_rel_remap; A.sizes = d_sizes;
    k_walk_universes<<<n, 32, SM_WALK_TOTAL, st>>>(A);
    PK_LAUNCHED("k_walk_universes");
    k_number_universes<<<n, 32, SM_NUMBER_TOTAL, st>>>(A);
    PK_LAUNCHED("k_number_universes");
    PK_CUDA(cudaEventRecord(g_ring.ev[slot], st));
    return PK_OK;
}

// d_ent_packed [sum nE], d_rel_packed [sum nR]: the remaps of universes 0..n-1 back to back (sizes from d_sizes).
int pk_walk_pack_remaps(int n, const int32_t* d_sizes, const int32_t* d_ent_remap, const int32_t* d_rel_remap, int32_t* d_ent_packed,
                        int32_t* d_rel_packed, void* stream) {
    pk::launch_counter() = 0;
    if (n <= 0 || !d_sizes || !d_ent_remap || !d_rel_remap || !d_ent_packed || !d_rel_packed)
        return pk::fail(PK_ERR_ARG, "pk_walk_pack_remaps: null argument");
    k_pack_remaps<<<n, 128, 0, (cudaStream_t)stream>>>(n, d_sizes, d_ent_remap, d_rel_remap, d_ent_packed, d_rel_packed);
    PK_LAUNCHED("k_pack_remaps");
    return PK_OK;
}

}  // extern "C"
