// K3 — link-prediction ranking on the device, and the PuTransE cross-universe energy aggregation.
//
// Reference path being replaced (per test triple and side, E = #entities):
//   TestDataLoader yields an [E]-long candidate batch (openke/base/Test.h:37-107), the model
//   scores it (openke/module/model/TransE.py:88-94 and TransH/TransD twins), the scores go back to
//   the host and testHead/testTail count, in an O(E) loop with a binary search per better-scored
//   candidate, how many candidates beat the true triple (openke/base/Test.h:118-359, _find in
//   openke/base/Corrupt.h:188-199).  PuTransE adds a Python loop that takes an elementwise minimum
//   over universes with one .item() per entity (openke/config/Parallel_Universe_Config.py:446-465).
//
// Energy of candidate c for a query with relation r (norm_flag on):
//   head side:  || y_r(c) + (r^ - y_r(t)) ||_p        (mode 'head_batch': h + (r - t))
//   tail side:  || (y_r(h) + r^) - y_r(c) ||_p        (mode 'tail_batch': (h + r) - t)
//   y_r(e) = normalize(e)                                   TransE
//          = normalize(e - (e.w^_r) w^_r)                   TransH
//          = normalize(normalize(e + (e.e_p) r_p))          TransD
// Every energy — the true triple's, the tile kernel's, the filter pass's — is produced by the same
// device functions below, evaluated by one thread sequentially over the dimension with explicit
// round-to-nearest intrinsics, so equal inputs give bit-equal energies and `<` is self-consistent.
#include <algorithm>
#include <cmath>

#include "common.hpp"
#include "kge_device.cuh"

namespace {

using namespace pkd;

constexpr int R_THREADS = 256;
constexpr int R_TE = 256;      // entities per tile (one per thread)
constexpr int R_QC = 64;       // queries per block chunk
constexpr int R_MAXD = 256;

struct SpaceView {
    const float* ent[2];
    const float* rel[2];
    int d, p_norm, norm_flag, model;
};

// y_r(e) for one entity, sequential over d.  `w` = w^ (H) or r_p (D), already prepared.
__device__ __forceinline__ void ent_operand(const SpaceView& sp, const float* __restrict__ e, const float* __restrict__ ep,
                                            const float* w, float* y) {
    const int d = sp.d;
    if (sp.model == TRANSE) {
        for (int i = 0; i < d; ++i) y[i] = e[i];
    } else if (sp.model == TRANSH) {
        float a = 0.f;
        for (int i = 0; i < d; ++i) a = __fmaf_rn(e[i], w[i], a);
        for (int i = 0; i < d; ++i) y[i] = __fsub_rn(e[i], __fmul_rn(a, w[i]));
    } else {
        float a = 0.f;
        for (int i = 0; i < d; ++i) a = __fmaf_rn(e[i], ep[i], a);
        float ss = 0.f;
        for (int i = 0; i < d; ++i) {
            y[i] = __fadd_rn(e[i], __fmul_rn(a, w[i]));
            ss = __fmaf_rn(y[i], y[i], ss);
        }
        const float n = fmaxf(__fsqrt_rn(ss), kNormEps);
        for (int i = 0; i < d; ++i) y[i] = __fdiv_rn(y[i], n);
    }
    if (sp.norm_flag) {
        float ss = 0.f;
        for (int i = 0; i < d; ++i) ss = __fmaf_rn(y[i], y[i], ss);
        const float n = fmaxf(__fsqrt_rn(ss), kNormEps);
        for (int i = 0; i < d; ++i) y[i] = __fdiv_rn(y[i], n);
    }
}

// relation-side preparation: rh = r^ ; w = w^ (H) / r_p (D)
__device__ __forceinline__ void rel_operand(const SpaceView& sp, int r, float* rh, float* w) {
    const int d = sp.d;
    const float* rr = sp.rel[0] + (size_t)r * d;
    for (int i = 0; i < d; ++i) rh[i] = rr[i];
    if (sp.norm_flag) {
        float ss = 0.f;
        for (int i = 0; i < d; ++i) ss = __fmaf_rn(rh[i], rh[i], ss);
        const float n = fmaxf(__fsqrt_rn(ss), kNormEps);
        for (int i = 0; i < d; ++i) rh[i] = __fdiv_rn(rh[i], n);
    }
    if (sp.model == TRANSH) {
        const float* ww = sp.rel[1] + (size_t)r * d;
        float ss = 0.f;
        for (int i = 0; i < d; ++i) { w[i] = ww[i]; ss = __fmaf_rn(w[i], w[i], ss); }
        const float n = fmaxf(__fsqrt_rn(ss), kNormEps);
        for (int i = 0; i < d; ++i) w[i] = __fdiv_rn(w[i], n);
    } else if (sp.model == TRANSD) {
        const float* ww = sp.rel[1] + (size_t)r * d;
        for (int i = 0; i < d; ++i) w[i] = ww[i];
    }
}

// the query vector: head side q = r^ - y(t);  tail side q = y(h) + r^
__device__ __forceinline__ void query_vector(int d, int side, const float* rh, const float* yfix, float* q) {
    if (side == 0) for (int i = 0; i < d; ++i) q[i] = __fsub_rn(rh[i], yfix[i]);
    else           for (int i = 0; i < d; ++i) q[i] = __fadd_rn(yfix[i], rh[i]);
}

// energy of a candidate operand y (strided access so that tiles can be stored transposed)
__device__ __forceinline__ float energy(int d, int p_norm, int side, const float* q, const float* y, int ystride) {
    float acc = 0.f;
    if (p_norm == 1) {
        for (int i = 0; i < d; ++i) {
            const float s = side == 0 ? __fadd_rn(y[(size_t)i * ystride], q[i]) : __fsub_rn(q[i], y[(size_t)i * ystride]);
            acc = __fadd_rn(acc, fabsf(s));
        }
        return acc;
    }
    for (int i = 0; i < d; ++i) {
        const float s = side == 0 ? __fadd_rn(y[(size_t)i * ystride], q[i]) : __fsub_rn(q[i], y[(size_t)i * ystride]);
        acc = __fmaf_rn(s, s, acc);
    }
    return __fsqrt_rn(acc);
}

// ------------------------------------------------------------------------------------------------
// pass 1: per (query, side) the query vector and the true triple's energy
struct RankParams {
    SpaceView sp;
    const int32_t* triples;  // [n*3] (h,r,t)
    int64_t n;
    int64_t n_ent;
    float* qvec;             // [n*2*d]
    float* target;           // [n*2]
    int32_t* ranks;          // [n*4] head raw, head filt, tail raw, tail filt
    const int64_t* foff[2];
    const int32_t* fcand[2];
};

__global__ void __launch_bounds__(128) k3_targets(const __grid_constant__ RankParams P) {
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;  // query*2 + side
    if (idx >= P.n * 2) return;
    const int64_t qi = idx >> 1;
    const int side = (int)(idx & 1);
    const int d = P.sp.d;
    const int32_t h = P.triples[qi * 3], r = P.triples[qi * 3 + 1], t = P.triples[qi * 3 + 2];
    float rh[R_MAXD], w[R_MAXD], y[R_MAXD];
    rel_operand(P.sp, r, rh, w);
    const int32_t fix = side == 0 ? t : h, truth = side == 0 ? h : t;
    ent_operand(P.sp, P.sp.ent[0] + (size_t)fix * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)fix * d : nullptr, w, y);
    float* q = P.qvec + idx * d;
    query_vector(d, side, rh, y, q);
    ent_operand(P.sp, P.sp.ent[0] + (size_t)truth * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)truth * d : nullptr, w, y);
    P.target[idx] = energy(d, P.sp.p_norm, side, q, y, 1);
    P.ranks[qi * 4 + side * 2 + 0] = 0;
    P.ranks[qi * 4 + side * 2 + 1] = 0;
}

// pass 2: raw ranks.  Block = (tile of R_TE entities) x (chunk of R_QC queries).  Queries arrive
// sorted by relation (reference testList order, Reader.h:311), so a chunk is a few relation runs;
// for each run the tile's operands y_r(e) are built once in shared memory (transposed: [d][R_TE])
// and every query of the run is scored against them.
__global__ void __launch_bounds__(R_THREADS) k3_raw(const __grid_constant__ RankParams P) {
    extern __shared__ __align__(16) float sm[];
    const int d = P.sp.d;
    float* ytile = sm;                      // [d][R_TE]
    float* wbuf = ytile + (size_t)d * R_TE; // [d]  w^ / r_p of the current run
    float* qbuf = wbuf + d;                 // [2][d] query vectors of the current query
    __shared__ int cnt[2];
    const int tid = threadIdx.x;
    const int64_t e = (int64_t)blockIdx.x * R_TE + tid;
    const bool live = e < P.n_ent;
    const int64_t q0 = (int64_t)blockIdx.y * R_QC, q1 = min(q0 + (int64_t)R_QC, P.n);
    int cur_rel = -1;
    for (int64_t qi = q0; qi < q1; ++qi) {
        const int32_t r = P.triples[qi * 3 + 1];
        if (r != cur_rel) {  // new relation run: rebuild the tile's operands
            __syncthreads();
            if (P.sp.model != TRANSE) {
                // every thread needs w; thread 0 prepares it once
                if (tid == 0) {
                    float rh[R_MAXD];
                    rel_operand(P.sp, r, rh, wbuf);
                }
                __syncthreads();
            }
            if (live) {
                float y[R_MAXD], w[R_MAXD];
                if (P.sp.model != TRANSE) for (int i = 0; i < d; ++i) w[i] = wbuf[i];
                ent_operand(P.sp, P.sp.ent[0] + (size_t)e * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)e * d : nullptr, w, y);
                for (int i = 0; i < d; ++i) ytile[(size_t)i * R_TE + tid] = y[i];
            }
            cur_rel = r;
        }
        __syncthreads();
        for (int i = tid; i < 2 * d; i += R_THREADS) qbuf[i] = P.qvec[qi * 2 * d + i];
        if (tid < 2) cnt[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            bool better = false;
            if (live) {
                const float en = energy(d, P.sp.p_norm, side, qbuf + side * d, ytile + tid, R_TE);
                better = en < P.target[qi * 2 + side];
            }
            const unsigned m = __ballot_sync(0xffffffffu, better);
            if ((tid & 31) == 0 && m) atomicAdd(&cnt[side], __popc(m));
        }
        __syncthreads();
        if (tid < 2 && cnt[tid]) atomicAdd(&P.ranks[qi * 4 + tid * 2 + 0], cnt[tid]);
    }
}

// pass 3: filtered rank = raw rank - #(known-true candidates that also beat the truth)
__global__ void __launch_bounds__(128) k3_filter(const __grid_constant__ RankParams P) {
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;  // query*2 + side
    if (idx >= P.n * 2) return;
    const int64_t qi = idx >> 1;
    const int side = (int)(idx & 1);
    const int d = P.sp.d;
    const int32_t r = P.triples[qi * 3 + 1];
    const int64_t lo = P.foff[side][qi], hi = P.foff[side][qi + 1];
    int sub = 0;
    if (hi > lo) {
        float rh[R_MAXD], w[R_MAXD], y[R_MAXD];
        rel_operand(P.sp, r, rh, w);
        const float* q = P.qvec + idx * d;
        const float tgt = P.target[idx];
        for (int64_t c = lo; c < hi; ++c) {
            const int32_t ce = P.fcand[side][c];
            ent_operand(P.sp, P.sp.ent[0] + (size_t)ce * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)ce * d : nullptr, w, y);
            if (energy(d, P.sp.p_norm, side, q, y, 1) < tgt) ++sub;
        }
    }
    // raw count is final (previous kernel finished); derive the filtered one
    P.ranks[qi * 4 + side * 2 + 1] = P.ranks[qi * 4 + side * 2 + 0] - sub;
}

// ------------------------------------------------------------------------------------------------
// PuTransE: energies of one universe's entities for one (key, side), min-folded into the key's row
struct EnergyParams {
    SpaceView sp;              // packed tables (all universes)
    const int64_t* ent_off;
    const int64_t* rel_off;
    const int32_t* n_ent;
    const int32_t* ent_remap;  // packed local -> global
    const pk_energy_item* items;
    int64_t n_items;
    float* energy;             // [n_keys, E]
    int64_t E;
};

__global__ void __launch_bounds__(R_THREADS) k3u_energies(const __grid_constant__ EnergyParams P) {
    extern __shared__ __align__(16) float sm[];
    const int d = P.sp.d;
    float* q = sm;          // [d]
    float* w = q + d;       // [d]
    const pk_energy_item it = P.items[blockIdx.x];
    const int64_t eo = P.ent_off[it.universe], ro = P.rel_off[it.universe];
    const int nE = P.n_ent[it.universe];
    SpaceView sp = P.sp;  // this universe's slice of the packed tables
    for (int i = 0; i < 2; ++i) {
        if (sp.ent[i]) sp.ent[i] += (size_t)eo * d;
        if (sp.rel[i]) sp.rel[i] += (size_t)ro * d;
    }
    if (threadIdx.x == 0) {
        float rh[R_MAXD], y[R_MAXD];
        rel_operand(sp, it.rel_local, rh, w);
        ent_operand(sp, sp.ent[0] + (size_t)it.fixed_local * d, sp.ent[1] ? sp.ent[1] + (size_t)it.fixed_local * d : nullptr, w, y);
        query_vector(d, it.side, rh, y, q);
    }
    __syncthreads();
    unsigned int* row = reinterpret_cast<unsigned int*>(P.energy + (size_t)it.key_row * P.E);
    const int32_t* remap = P.ent_remap + eo;
    for (int e = threadIdx.x; e < nE; e += R_THREADS) {
        float y[R_MAXD], wl[R_MAXD];
        if (sp.model != TRANSE) for (int i = 0; i < d; ++i) wl[i] = w[i];
        ent_operand(sp, sp.ent[0] + (size_t)e * d, sp.ent[1] ? sp.ent[1] + (size_t)e * d : nullptr, wl, y);
        const float en = energy(d, sp.p_norm, it.side, q, y, 1);
        // energies are >= 0, so their bit patterns order like unsigned integers; +inf = 0x7f800000
        atomicMin(row + remap[e], __float_as_uint(en));
    }
}

// PuTransE 'null_vector' handling (reference Parallel_Universe_Config.py:378-388,494-514): the tuple
// score of a key in one universe is the model's _calc() with the missing side replaced by a zero
// vector and WITHOUT the TransH / TransD projection (the reference calls _calc on the raw embedding
// rows):  head batch ||0 + (r^ - e^)||_p ,  tail batch ||(e^ + r^) - 0||_p ,  e^ = normalize(e_fixed).
// One thread per work item; min over universes with atomicMin on the bit pattern (scores >= 0).
__global__ void __launch_bounds__(128) k3u_tuple_scores(const __grid_constant__ EnergyParams P, float* tuple) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= P.n_items) return;
    const pk_energy_item it = P.items[i];
    const int d = P.sp.d;
    const float* e = P.sp.ent[0] + ((size_t)P.ent_off[it.universe] + it.fixed_local) * d;
    const float* r = P.sp.rel[0] + ((size_t)P.rel_off[it.universe] + it.rel_local) * d;
    float ne = 1.f, nr = 1.f;
    if (P.sp.norm_flag) {
        float se = 0.f, sr = 0.f;
        for (int j = 0; j < d; ++j) { se = __fmaf_rn(e[j], e[j], se); sr = __fmaf_rn(r[j], r[j], sr); }
        ne = fmaxf(__fsqrt_rn(se), kNormEps);
        nr = fmaxf(__fsqrt_rn(sr), kNormEps);
    }
    float acc = 0.f;
    for (int j = 0; j < d; ++j) {
        const float eh = __fdiv_rn(e[j], ne), rh = __fdiv_rn(r[j], nr);
        const float s = it.side == 0 ? __fadd_rn(0.f, __fsub_rn(rh, eh)) : __fsub_rn(__fadd_rn(eh, rh), 0.f);
        acc = P.sp.p_norm == 1 ? __fadd_rn(acc, fabsf(s)) : __fmaf_rn(s, s, acc);
    }
    if (P.sp.p_norm != 1) acc = __fsqrt_rn(acc);
    atomicMin(reinterpret_cast<unsigned int*>(tuple) + it.key_row, __float_as_uint(acc));
}

// every still-unscored candidate of a key row gets the key's tuple score (if some universe has one)
__global__ void k3u_fill_missing(float* energy, int64_t n_rows, int64_t E, const float* tuple) {
    const int64_t row = blockIdx.y;
    const float t = tuple[row];
    if (isinf(t)) return;
    float* p = energy + (size_t)row * E;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < E; c += (int64_t)gridDim.x * blockDim.x)
        if (isinf(p[c])) p[c] = t;
}

struct FromEnergyParams {
    const float* energy;
    int64_t E, n;
    const int32_t* key_row;
    const int32_t* truth;
    const int64_t* foff;
    const int32_t* fcand;
    int32_t* ranks;  // [n*2]
    int layout;      // 0: row indexed by entity id; 1: reference candidate order (slot 0 = truth)
    // incremental setting (Test.h:181-206 incremental branch): only entities the snapshot currently contains are
    // candidates (mask[e] != 0) and an unscored truth ranks behind all n_candidates of them; nullptr: every entity
    const uint8_t* mask;
    int64_t n_candidates;
};

__device__ __forceinline__ float row_at(const float* row, int layout, int32_t truth, int64_t c) {
    if (layout == 0) return row[c];
    return c == truth ? row[0] : (c < truth ? row[c + 1] : row[c]);
}

// reference testHead/testTail on one energy row (openke/base/Test.h:118-238), incl. the +inf branch
__global__ void __launch_bounds__(R_THREADS) k3_rank_rows(const __grid_constant__ FromEnergyParams P) {
    const int64_t qi = blockIdx.x;
    const float* row = P.energy + (size_t)(P.key_row ? P.key_row[qi] : qi) * P.E;
    const int32_t truth = P.truth[qi];
    const float tgt = row_at(row, P.layout, truth, truth);
    __shared__ int s_raw, s_sub;
    if (threadIdx.x == 0) { s_raw = 0; s_sub = 0; }
    __syncthreads();
    const bool missing = isinf(tgt) && tgt > 0.f;
    int raw = 0, sub = 0;
    if (!missing) {
        if (P.layout == 0 && P.mask) {
            for (int64_t c = threadIdx.x; c < P.E; c += R_THREADS) raw += (P.mask[c] != 0 && row[c] < tgt);
        } else if (P.layout == 0) {
            for (int64_t c = threadIdx.x; c < P.E; c += R_THREADS) raw += (row[c] < tgt);
        } else {
            for (int64_t c = 1 + threadIdx.x; c < P.E; c += R_THREADS) raw += (row[c] < tgt);
        }
    }
    const int64_t lo = P.foff ? P.foff[qi] : 0, hi = P.foff ? P.foff[qi + 1] : 0;
    for (int64_t c = lo + threadIdx.x; c < hi; c += R_THREADS) {
        const int32_t ce = P.fcand[c];
        if (P.mask && P.mask[ce] == 0) continue;      // a known triple whose entity has left the graph is no candidate
        if (missing) sub += 1;
        else sub += (row_at(row, P.layout, truth, ce) < tgt);
    }
    raw = (int)gsum<32>((float)raw);  // counts < 2^24: exact in float
    sub = (int)gsum<32>((float)sub);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_raw, raw); atomicAdd(&s_sub, sub); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int r = missing ? (int)(P.mask ? P.n_candidates : P.E) : s_raw;
        P.ranks[qi * 2 + 0] = r;
        P.ranks[qi * 2 + 1] = r - s_sub;
    }
}

__global__ void k_fill(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}


// ------------------------------------------------------------------------------------------------
// Model.forward / Model.predict on index batches (reference TransE.py:62-74,88-94): one thread per
// output score; index arrays broadcast by modulo exactly like the reference's view(-1, r.shape[0], d).
struct ScoreParams {
    SpaceView sp;
    const int64_t* h; const int64_t* t; const int64_t* r;
    int64_t nh, nt, nr, n;
    int head_batch;  // 1: h + (r - t) ; 0: (h + r) - t
    float* out;
    int64_t n_ent, n_rel;
    int* bad;
};

__global__ void __launch_bounds__(128) k_score_batch(const __grid_constant__ ScoreParams P) {
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= P.n) return;
    const int d = P.sp.d;
    const int64_t h = P.h[i % P.nh], t = P.t[i % P.nt], r = P.r[i % P.nr];
    if (h < 0 || h >= P.n_ent || t < 0 || t >= P.n_ent || r < 0 || r >= P.n_rel) { *P.bad = 1; P.out[i] = nanf(""); return; }
    float rh[R_MAXD], w[R_MAXD], y[R_MAXD], q[R_MAXD];
    rel_operand(P.sp, (int)r, rh, w);
    const int side = P.head_batch ? 0 : 1;
    const int64_t fix = side == 0 ? t : h, var = side == 0 ? h : t;
    ent_operand(P.sp, P.sp.ent[0] + (size_t)fix * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)fix * d : nullptr, w, y);
    query_vector(d, side, rh, y, q);
    ent_operand(P.sp, P.sp.ent[0] + (size_t)var * d, P.sp.ent[1] ? P.sp.ent[1] + (size_t)var * d : nullptr, w, y);
    P.out[i] = energy(d, P.sp.p_norm, side, q, y, 1);
}

struct Scratch {  // query vectors + targets, grown on demand, per thread
    float* p = nullptr;
    size_t cap = 0;
};
thread_local Scratch g_scratch;

int fill_space(SpaceView& sp, const pk_model_cfg* cfg, const pk_tables* tab, const char* who) {
    if (!cfg || !tab) return pk::fail(PK_ERR_ARG, std::string(who) + ": null argument");
    if (cfg->model < 0 || cfg->model > 2) return pk::fail(PK_ERR_ARG, std::string(who) + ": unknown model");
    if (cfg->dim < 1 || cfg->dim > R_MAXD) return pk::fail(PK_ERR_UNSUPPORTED, std::string(who) + ": dim must be in [1,256]");
    if (cfg->p_norm != 1 && cfg->p_norm != 2) return pk::fail(PK_ERR_UNSUPPORTED, std::string(who) + ": p_norm must be 1 or 2");
    const int ntE = cfg->model == PK_TRANSD ? 2 : 1, ntR = cfg->model == PK_TRANSE ? 1 : 2;
    for (int i = 0; i < 2; ++i) {
        sp.ent[i] = i < ntE ? tab->ent[i] : nullptr;
        sp.rel[i] = i < ntR ? tab->rel[i] : nullptr;
        if ((i < ntE && !tab->ent[i]) || (i < ntR && !tab->rel[i])) return pk::fail(PK_ERR_ARG, std::string(who) + ": missing table");
    }
    sp.d = cfg->dim; sp.p_norm = cfg->p_norm; sp.norm_flag = cfg->norm_flag; sp.model = cfg->model;
    return PK_OK;
}

}  // namespace

extern "C" int pk_rank_space(const pk_model_cfg* cfg, const pk_tables* tab, int64_t n, const int32_t* d_triples,
                             const int64_t* d_foff_head, const int32_t* d_fcand_head, const int64_t* d_foff_tail,
                             const int32_t* d_fcand_tail, int32_t* d_ranks, void* stream) {
    pk::launch_counter() = 0;
    RankParams P;
    int rc = fill_space(P.sp, cfg, tab, "pk_rank_space");
    if (rc != PK_OK) return rc;
    if (n < 0 || !d_triples || !d_ranks || !d_foff_head || !d_foff_tail) return pk::fail(PK_ERR_ARG, "pk_rank_space: null argument");
    if (n == 0) return PK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int d = cfg->dim;
    const size_t need = ((size_t)n * 2 * d + (size_t)n * 2) * sizeof(float);
    if (g_scratch.cap < need) {
        if (g_scratch.p) cudaFree(g_scratch.p);
        g_scratch.p = nullptr;
        g_scratch.cap = 0;
        PK_CUDA(cudaMalloc(&g_scratch.p, need));
        g_scratch.cap = need;
    }
    P.triples = d_triples; P.n = n; P.n_ent = tab->n_ent;
    P.qvec = g_scratch.p; P.target = g_scratch.p + (size_t)n * 2 * d;
    P.ranks = d_ranks;
    P.foff[0] = d_foff_head; P.fcand[0] = d_fcand_head; P.foff[1] = d_foff_tail; P.fcand[1] = d_fcand_tail;
    const unsigned qb = (unsigned)((n * 2 + 127) / 128);
    k3_targets<<<qb, 128, 0, st>>>(P);
    PK_LAUNCHED("k3_targets");
    const size_t smem = ((size_t)d * R_TE + 3 * (size_t)d) * sizeof(float);
    if (smem > 227 * 1024) return pk::fail(PK_ERR_UNSUPPORTED, "pk_rank_space: dim too large for the ranking tile (d <= 220)");
    PK_CUDA(cudaFuncSetAttribute(k3_raw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((tab->n_ent + R_TE - 1) / R_TE), (unsigned)((n + R_QC - 1) / R_QC));
    k3_raw<<<grid, R_THREADS, smem, st>>>(P);
    PK_LAUNCHED("k3_raw");
    k3_filter<<<qb, 128, 0, st>>>(P);
    PK_LAUNCHED("k3_filter");
    return PK_OK;
}

extern "C" int pk_universe_energies(const pk_model_cfg* cfg, const pk_tables* packed, const int64_t* d_ent_off,
                                    const int64_t* d_rel_off, const int32_t* d_n_ent, const int32_t* d_ent_remap,
                                    const pk_energy_item* d_items, int64_t n_items, float* d_energy, int64_t n_ent_global,
                                    void* stream) {
    pk::launch_counter() = 0;
    EnergyParams P;
    int rc = fill_space(P.sp, cfg, packed, "pk_universe_energies");
    if (rc != PK_OK) return rc;
    if (!d_ent_off || !d_rel_off || !d_n_ent || !d_ent_remap || !d_energy || (n_items > 0 && !d_items))
        return pk::fail(PK_ERR_ARG, "pk_universe_energies: null argument");
    if (n_items == 0) return PK_OK;
    P.ent_off = d_ent_off; P.rel_off = d_rel_off; P.n_ent = d_n_ent; P.ent_remap = d_ent_remap;
    P.items = d_items; P.n_items = n_items; P.energy = d_energy; P.E = n_ent_global;
    const size_t smem = 2 * (size_t)cfg->dim * sizeof(float);
    k3u_energies<<<(unsigned)n_items, R_THREADS, smem, (cudaStream_t)stream>>>(P);
    PK_LAUNCHED("k3u_energies");
    return PK_OK;
}

extern "C" int pk_universe_tuple_scores(const pk_model_cfg* cfg, const pk_tables* packed, const int64_t* d_ent_off,
                                        const int64_t* d_rel_off, const pk_energy_item* d_items, int64_t n_items, float* d_tuple,
                                        void* stream) {
    pk::launch_counter() = 0;
    EnergyParams P;
    int rc = fill_space(P.sp, cfg, packed, "pk_universe_tuple_scores");
    if (rc != PK_OK) return rc;
    if (!d_ent_off || !d_rel_off || !d_tuple || (n_items > 0 && !d_items)) return pk::fail(PK_ERR_ARG, "pk_universe_tuple_scores: null argument");
    if (n_items == 0) return PK_OK;
    P.ent_off = d_ent_off; P.rel_off = d_rel_off; P.n_ent = nullptr; P.ent_remap = nullptr;
    P.items = d_items; P.n_items = n_items; P.energy = nullptr; P.E = 0;
    k3u_tuple_scores<<<(unsigned)((n_items + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P, d_tuple);
    PK_LAUNCHED("k3u_tuple_scores");
    return PK_OK;
}

extern "C" int pk_fill_missing_energies(float* d_energy, int64_t n_rows, int64_t n_ent_global, const float* d_tuple, void* stream) {
    pk::launch_counter() = 0;
    if (!d_energy || !d_tuple || n_rows < 0) return pk::fail(PK_ERR_ARG, "pk_fill_missing_energies: null argument");
    if (n_rows == 0) return PK_OK;
    if (n_rows > 65535) return pk::fail(PK_ERR_UNSUPPORTED, "pk_fill_missing_energies: at most 65535 rows per call");
    dim3 grid((unsigned)std::min<int64_t>((n_ent_global + 255) / 256, 64), (unsigned)n_rows);
    k3u_fill_missing<<<grid, 256, 0, (cudaStream_t)stream>>>(d_energy, n_rows, n_ent_global, d_tuple);
    PK_LAUNCHED("k3u_fill_missing");
    return PK_OK;
}

extern "C" int pk_rank_from_energy(const float* d_energy, int64_t n_ent_global, int64_t n, const int32_t* d_key_row,
                                   const int32_t* d_truth, const int64_t* d_foff, const int32_t* d_fcand, int32_t* d_ranks,
                                   void* stream) {
    pk::launch_counter() = 0;
    if (!d_energy || !d_truth || !d_ranks || n < 0) return pk::fail(PK_ERR_ARG, "pk_rank_from_energy: null argument");
    if (n == 0) return PK_OK;
    FromEnergyParams P;
    P.energy = d_energy; P.E = n_ent_global; P.n = n; P.key_row = d_key_row; P.truth = d_truth;
    P.foff = d_foff; P.fcand = d_fcand; P.ranks = d_ranks; P.layout = 0;
    P.mask = nullptr; P.n_candidates = n_ent_global;
    k3_rank_rows<<<(unsigned)n, R_THREADS, 0, (cudaStream_t)stream>>>(P);
    PK_LAUNCHED("k3_rank_rows");
    return PK_OK;
}

// the same with a candidate mask: the incremental setting ranks among the entities the snapshot currently holds
extern "C" int pk_rank_from_energy_masked(const float* d_energy, int64_t n_ent_global, int64_t n, const int32_t* d_key_row,
                                          const int32_t* d_truth, const int64_t* d_foff, const int32_t* d_fcand, int32_t* d_ranks,
                                          const uint8_t* d_candidate_mask, int64_t n_candidates, void* stream) {
    pk::launch_counter() = 0;
    if (!d_energy || !d_truth || !d_ranks || !d_candidate_mask || n < 0) return pk::fail(PK_ERR_ARG, "pk_rank_from_energy_masked: null argument");
    if (n == 0) return PK_OK;
    FromEnergyParams P;
    P.energy = d_energy; P.E = n_ent_global; P.n = n; P.key_row = d_key_row; P.truth = d_truth;
    P.foff = d_foff; P.fcand = d_fcand; P.ranks = d_ranks; P.layout = 0;
    P.mask = d_candidate_mask; P.n_candidates = n_candidates;
    k3_rank_rows<<<(unsigned)n, R_THREADS, 0, (cudaStream_t)stream>>>(P);
    PK_LAUNCHED("k3_rank_rows");
    return PK_OK;
}

// one row in the reference's candidate order (slot 0 = truth): the kernel behind testHead/testTail
extern "C" int pk_rank_candidate_row(const float* d_con, int64_t n_ent, const int32_t* d_truth1, const int64_t* d_foff2,
                                     const int32_t* d_fcand, int32_t* d_ranks2, void* stream) {
    pk::launch_counter() = 0;
    if (!d_con || !d_ranks2 || !d_truth1) return pk::fail(PK_ERR_ARG, "pk_rank_candidate_row: null argument");
    FromEnergyParams P;
    P.energy = d_con; P.E = n_ent; P.n = 1; P.key_row = nullptr; P.truth = d_truth1;
    P.foff = d_foff2; P.fcand = d_fcand; P.ranks = d_ranks2; P.layout = 1;
    P.mask = nullptr; P.n_candidates = n_ent;
    k3_rank_rows<<<1, R_THREADS, 0, (cudaStream_t)stream>>>(P);
    PK_LAUNCHED("k3_rank_rows");
    return PK_OK;
}


extern "C" int pk_score_batch(const pk_model_cfg* cfg, const pk_tables* tab, const int64_t* d_h, int64_t nh, const int64_t* d_t,
                              int64_t nt, const int64_t* d_r, int64_t nr, int head_batch, float* d_out, int* d_bad_flag,
                              void* stream) {
    pk::launch_counter() = 0;
    ScoreParams P;
    int rc = fill_space(P.sp, cfg, tab, "pk_score_batch");
    if (rc != PK_OK) return rc;
    if (!d_h || !d_t || !d_r || !d_out || !d_bad_flag || nh < 1 || nt < 1 || nr < 1) return pk::fail(PK_ERR_ARG, "pk_score_batch: null/empty argument");
    P.h = d_h; P.t = d_t; P.r = d_r; P.nh = nh; P.nt = nt; P.nr = nr;
    P.n = std::max(nh, std::max(nt, nr));
    P.head_batch = head_batch; P.out = d_out; P.n_ent = tab->n_ent; P.n_rel = tab->n_rel; P.bad = d_bad_flag;
    k_score_batch<<<(unsigned)((P.n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P);
    PK_LAUNCHED("k_score_batch");
    return PK_OK;
}

// x / max(||x||, eps) per row with the ranking kernels' own sequence of operations (ent_operand / rel_operand above), so
// that a space whose tables were normalised ONCE scores bit-identically with norm_flag = 0 (TransE: the operand of a
// candidate is its normalised row, whatever the key; the evaluation of an ensemble visits every universe once per key)
__global__ void __launch_bounds__(128) k_normalise_rows(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int d) {
    const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (r >= rows) return;
    const float* x = in + r * d;
    float* y = out + r * d;
    float ss = 0.f;
    for (int i = 0; i < d; ++i) ss = __fmaf_rn(x[i], x[i], ss);
    const float n = fmaxf(__fsqrt_rn(ss), kNormEps);
    for (int i = 0; i < d; ++i) y[i] = __fdiv_rn(x[i], n);
}

extern "C" int pk_normalise_rows(const float* d_in, float* d_out, int64_t rows, int d, void* stream) {
    pk::launch_counter() = 0;
    if (!d_in || !d_out || rows < 0 || d < 1) return pk::fail(PK_ERR_ARG, "pk_normalise_rows: bad argument");
    if (rows == 0) return PK_OK;
    k_normalise_rows<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_in, d_out, rows, d);
    PK_LAUNCHED("k_normalise_rows");
    return PK_OK;
}

extern "C" int pk_fill_inf(float* d, int64_t n, void* stream) {
    pk::launch_counter() = 0;
    if (!d || n < 0) return pk::fail(PK_ERR_ARG, "pk_fill_inf: null argument");
    if (n == 0) return PK_OK;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16);
    k_fill<<<blocks, 256, 0, (cudaStream_t)stream>>>(d, n, INFINITY);
    PK_LAUNCHED("k_fill");
    return PK_OK;
}
