// K0 + K1 — device sampler and the fused single-space train step.
//
// One reference training step (openke/config/Trainer.py:44-56) is: sampling() on the host
// (openke/base/Base.cpp:266-310), 4 H2D copies, ~100-190 ATen kernels (gather, normalise,
// translate, norm, margin loss, autograd with a DENSE [E,d] gradient, dense optimizer) and a
// loss.item() sync.  Here it is a handful of launches that never leave the device:
//
//   k1_prepare   draws the batch with the reference's LCG streams (bit-exact; jump tables +
//                multiply-high modulo) — or converts a caller-supplied batch — into the compact form
//                (h, t, r, replacement entity | side), counts how often every table row occurs in
//                the batch and queues the rows that occur more than once and the relations touched;
//   k1_grad      persistent, one lane group per positive sample, software-pipelined: while tile i
//                is computed, the table rows (and Adagrad state) of tiles i+1 and i+2 are in flight
//                from HBM through cp.async into per-lane shared-memory staging, so the kernel is
//                bound by memory bandwidth, not by load latency.  A row that occurs ONCE in the
//                batch is read once and written once, in place, by the group that owns the sample
//                (the algorithmic minimum); rows that occur several times accumulate with native
//                RED.ADD.F32 into a dense accumulator, relations into one of C privatised copies
//                (a hot relation is in 10 % of the samples: one copy would serialise in L2);
//   k1_apply     optimizer for the queued rows and the touched relations (copies summed,
//                normalisation backward once per relation, cache of normalised relation operands
//                refreshed), occurrence counters back to zero, loss reduced in a fixed order.  Inside
//                pk_train_steps it is launched FUSED with the next batch's k1_prepare
//                (k1_apply_prepare: two block roles in one grid — both are latency-bound and share
//                the SMs; batches alternate between two "batch sets").
//
// The sparse update is exactly the reference's dense one: SGD and Adagrad (lr_decay = 0,
// weight_decay = 0, reference Trainer.py:34-35,65-70,84-88) leave zero-gradient rows bit-unchanged.
// Bound on big tables (E*d*4 >> L2): HBM, ~(3+k) rows read + written per positive.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.hpp"
#include "kge_device.cuh"

using namespace pkd;

struct pk_workspace {
    pk_model_cfg cfg;
    int64_t n_ent, n_rel, max_batch;
    // two "batch sets" (counts, compact ids, queues, counters): while step s trains on set s % 2,
    // the batch of step s + 1 is prepared into the other one
    int32_t* cnt_ent[2] = {nullptr, nullptr};   // [n_ent] occurrences in the batch
    int32_t* cnt_rel[2] = {nullptr, nullptr};   // [n_rel]
    float* acc_ent[2] = {nullptr, nullptr};  // dense gradient accumulators for multiply-occurring rows
    float* acc_rel = nullptr;     // [C][ntR][n_rel][d] privatised relation gradient sums
    float* relc[2] = {nullptr, nullptr};     // cached r^ ; w^ (TransH)
    float* reln = nullptr;        // [2][n_rel] clamped norms
    int32_t* dup_ent[2] = {nullptr, nullptr};   // queue of multiply-occurring entity rows
    int32_t* touched_rel[2] = {nullptr, nullptr};
    int32_t* counters[2] = {nullptr, nullptr};  // [0] queued entities, [1] touched relations, [2] error flag, [3] block ticket
    int64_t* step_ctr = nullptr;  // index of the next loss slot
    int32_t* ids[2] = {nullptr, nullptr};       // h[B] | t[B] | r[B] | c[k][B]
    float* loss_part = nullptr;   // per-block partial sums of k1_grad
    uint64_t* jump = nullptr;     // A[per] | C[per] | Aadv[64] | Cadv[64]  (LCG jump tables)
    int64_t jump_B = -1;
    int jump_W = -1, jump_k = -1;
    int64_t jump_cap = 0;
    cudaStream_t own_stream = nullptr;  // blocking stream used when the caller hands us the legacy default stream
    bool memset_reset = true;           // counts cleared by zeroing the arrays (small tables) or by scattering over the batch
    int64_t dup_cap_ent = 0;
    int rel_copies = 1;
    int max_blocks = 0;
    // the captured chunk of steps, kept across pk_train_steps calls with identical launch parameters
    // (an epoch loop calls with the same tables, sampler, batch size and loss buffer every time)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_key;
    int graph_launches = 0;
};

namespace pkk1 {

constexpr int K1_THREADS = 256;
constexpr int K1_STAGES = 3;      // tiles in flight per warp: i (compute), i+1, i+2 (cp.async)

struct K1Params {
    float* ent[2];
    float* rel[2];
    float* ent_state[2];
    float* rel_state[2];
    float* acc_ent[2];
    float* acc_rel;
    float* relc[2];
    float* reln;
    int32_t* cnt_ent;
    int32_t* cnt_rel;
    int32_t* dup_ent;
    int32_t* touched_rel;
    int32_t* counters;
    int64_t* step_ctr;
    int32_t* ids;
    float* loss_part;
    float* loss;        // [steps] indexed by *step_ctr
    int64_t B;
    int64_t n_ent, n_rel;
    int d, k, p_norm, norm_flag, opt;
    float margin, lr;
    int grad_blocks, rel_copies;
    int apply_blocks;   // blocks that take the optimizer-tail role (the fused kernel appends prepare blocks)
    int clear_mode;     // occurrence counts of this batch set after the step: 0 keep, 1 zero the whole arrays, 2 scatter over the batch
};

// ---- shared by the plain and the fused tail kernels
// ---- multiply-high modulo and LCG jump coefficients (same as the universe kernel's producer)
struct FastMod { uint64_t n, m; };
__device__ __forceinline__ FastMod make_fastmod(uint64_t n) {
    FastMod f;
    f.n = n;
    f.m = ~0ULL / n;
    return f;
}
__device__ __forceinline__ uint64_t fastmod(uint64_t x, const FastMod& f) {
    const uint64_t q = __umul64hi(x, f.m);
    uint64_t r = x - q * f.n;
    if (r >= f.n) r -= f.n;
    if (r >= f.n) r -= f.n;
    return r;
}
__device__ __forceinline__ void lcg_affine(uint64_t n, uint64_t& A, uint64_t& C) {
    uint64_t a = kLcgMul, c = kLcgInc, ra = 1, rc = 0;
    while (n) {
        if (n & 1) { ra = ra * a; rc = rc * a + c; }
        c = (a + 1) * c;
        a = a * a;
        n >>= 1;
    }
    A = ra;
    C = rc;
}

// ---- k1_prepare: the batch in compact form + occurrence counts + queues
struct PrepParams {
    SamplerView sv;
    const uint64_t* lcg;       // W stream states at the start of this batch (sampling mode), else NULL
    const uint64_t* jump;
    int64_t per;
    const int32_t* gh;         // supplied batch in the reference layout [B pos | B neg#1 | ...] (convert mode)
    const int32_t* gt;
    const int32_t* gr;
    int32_t* ids;              // out: h[B] | t[B] | r[B] | c[k][B]
    int32_t* cnt_ent;
    int32_t* cnt_rel;
    int32_t* dup_ent;
    int32_t* touched_rel;
    int32_t* counters;
    int64_t B, n_ent, n_rel;
    int k, bern, filter, W;
    int ahead;                 // 1: draw the batch AFTER the one the stream states stand at
    uint64_t* lcg_out;         // non-NULL: the last prepare block moves the streams one batch on (ticket in counters[3])
    int nblocks;               // blocks of the prepare grid
};

constexpr int PREP_QCAP = K1_THREADS * 3;
constexpr int PREP_RELBITS = 32768;   // relations de-duplicated per block in a shared-memory bitmap up to this many

// the work of one block of the prepare grid (block index `bid`); K1_THREADS threads
__device__ __forceinline__ void prepare_block(const PrepParams& S, int bid) {
    __shared__ int32_t q_ent[PREP_QCAP], q_rel[K1_THREADS];
    __shared__ unsigned rel_seen[PREP_RELBITS / 32];
    __shared__ int n_qe, n_qr, base_e, base_r;
    const int64_t b = (int64_t)bid * K1_THREADS + threadIdx.x;
    const bool rel_bitmap = S.n_rel <= PREP_RELBITS;
    if (threadIdx.x == 0) { n_qe = 0; n_qr = 0; }
    if (rel_bitmap)
        for (int i = threadIdx.x; i < (int)((S.n_rel + 31) / 32); i += K1_THREADS) rel_seen[i] = 0u;
    __syncthreads();
    int32_t h = 0, t = 0, r = 0;
    bool ok = b < S.B;
    if (ok && S.lcg) {
        // the reference sampling() (Base.cpp:185-264), one thread per positive, bit-exact
        const FastMod fm_tri = make_fastmod((uint64_t)S.sv.n_tri), fm_coin = make_fastmod(1000ULL),
                      fm_ent = make_fastmod((uint64_t)(S.sv.n_ent - 1));
        const int id = (int)(b / S.per);
        const int64_t j = b - (int64_t)id * S.per;
        uint64_t s0 = S.lcg[id];
        if (S.ahead) s0 = S.jump[2 * S.per + id] * s0 + S.jump[2 * S.per + 64 + id];
        uint64_t s = S.jump[j] * s0 + S.jump[S.per + j];
        const int64_t i = (int64_t)fastmod(lcg_next(s), fm_tri);
        h = S.sv.by_head[i * 3 + 0]; r = S.sv.by_head[i * 3 + 1]; t = S.sv.by_head[i * 3 + 2];
        float prob = 500.f;
        if (S.bern) {
            const float rm = S.sv.right_mean[r], lm = S.sv.left_mean[r];
            prob = __fdiv_rn(__fmul_rn(1000.f, rm), __fadd_rn(rm, lm));  // Base.cpp:220-221
        }
        for (int n = 0; n < S.k; ++n) {
            const uint64_t coin = fastmod(lcg_next(s), fm_coin);
            const uint64_t x = lcg_next(s);
            int32_t c, side;
            if ((float)coin < prob) {   // keep head, replace tail
                if (!S.filter) { const int64_t tmp = (int64_t)fastmod(x, fm_ent); c = (int32_t)(tmp < h ? tmp : tmp + 1); }
                else c = corrupt_entity(x, S.sv.by_head, S.sv.n_tri, S.sv.n_ent, h, r, 0, 2, true, S.sv.head_off);
                side = 0;
            } else {                    // keep tail, replace head
                if (!S.filter) { const int64_t tmp = (int64_t)fastmod(x, fm_ent); c = (int32_t)(tmp < t ? tmp : tmp + 1); }
                else c = corrupt_entity(x, S.sv.by_tail, S.sv.n_tri, S.sv.n_ent, t, r, 2, 0, true, S.sv.tail_off);
                side = 1;
            }
            S.ids[(3 + (int64_t)n) * S.B + b] = (int32_t)((uint32_t)c | ((uint32_t)side << 31));
        }
    } else if (ok) {
        // a caller-supplied batch: every negative must keep its positive's relation and differ from
        // it in at most one entity (what the reference sampler produces); anything else is refused
        h = S.gh[b]; t = S.gt[b]; r = S.gr[b];
        bool good = r >= 0 && r < S.n_rel && h >= 0 && h < S.n_ent && t >= 0 && t < S.n_ent;
        for (int n = 0; n < S.k && good; ++n) {
            const int64_t o = b + (int64_t)(1 + n) * S.B;
            const int32_t nh = S.gh[o], nt = S.gt[o];
            good = S.gr[o] == r && nh >= 0 && nh < S.n_ent && nt >= 0 && nt < S.n_ent && (nh == h || nt == t);
        }
        if (!good) {
            atomicExch(&S.counters[2], 1);
            ok = false;
        } else {
            for (int n = 0; n < S.k; ++n) {
                const int64_t o = b + (int64_t)(1 + n) * S.B;
                const int32_t nh = S.gh[o], nt = S.gt[o];
                const int32_t side = nh == h ? 0 : 1, c = nh == h ? nt : nh;
                S.ids[(3 + (int64_t)n) * S.B + b] = (int32_t)((uint32_t)c | ((uint32_t)side << 31));
            }
        }
    }
    if (ok) {
        S.ids[b] = h; S.ids[S.B + b] = t; S.ids[2 * S.B + b] = r;
        // occurrence counts; the SECOND occurrence of an entity row queues it, the FIRST of a relation queues it
        auto push_ent = [&](int32_t e) {
            const int slot = atomicAdd(&n_qe, 1);
            if (slot < PREP_QCAP) q_ent[slot] = e;
            else S.dup_ent[atomicAdd(&S.counters[0], 1)] = e;   // block queue full (k > 1): straight to the global one
        };
        if (atomicAdd(&S.cnt_ent[h], 1) == 1) push_ent(h);
        if (atomicAdd(&S.cnt_ent[t], 1) == 1) push_ent(t);
        // a hot relation is in a tenth of the samples: one global atomic per sample would serialise in
        // L2, so samples of a block are de-duplicated in a shared-memory bitmap first
        bool first_in_block = true;
        if (rel_bitmap) first_in_block = !((atomicOr(&rel_seen[r >> 5], 1u << (r & 31)) >> (r & 31)) & 1u);
        if (first_in_block && atomicExch(&S.cnt_rel[r], 1) == 0) q_rel[atomicAdd(&n_qr, 1)] = r;
        for (int n = 0; n < S.k; ++n) {
            const int32_t c = S.ids[(3 + (int64_t)n) * S.B + b] & 0x7fffffff;
            if (atomicAdd(&S.cnt_ent[c], 1) == 1) push_ent(c);
        }
    }
    __syncthreads();
    const int nqe = min(n_qe, PREP_QCAP);
    if (threadIdx.x == 0) {
        base_e = nqe ? atomicAdd(&S.counters[0], nqe) : 0;
        base_r = n_qr ? atomicAdd(&S.counters[1], n_qr) : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nqe; i += K1_THREADS) S.dup_ent[base_e + i] = q_ent[i];
    for (int i = threadIdx.x; i < n_qr; i += K1_THREADS) S.touched_rel[base_r + i] = q_rel[i];
    if (S.lcg_out) {   // every block has read the stream states: the last one to finish advances them
        __shared__ int last_prep;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last_prep = atomicAdd(&S.counters[3], 1) == S.nblocks - 1;
        }
        __syncthreads();
        if (last_prep) {
            if ((int)threadIdx.x < S.W)
                S.lcg_out[threadIdx.x] = S.jump[2 * S.per + threadIdx.x] * S.lcg_out[threadIdx.x] + S.jump[2 * S.per + 64 + threadIdx.x];
            if (threadIdx.x == 0) S.counters[3] = 0;
        }
    }
}


#ifdef PK_MODEL_TU

// ---- cp.async helpers (per-lane private staging: a lane later reads only what it copied itself,
//      so cp.async.wait_group is the only synchronisation needed)
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// this lane's chunks of one table row -> the same offsets of a staged row
template <class L>
__device__ __forceinline__ void stage_row(float* dst, const float* src, int d, int lane) {
#pragma unroll
    for (int c = 0; c < L::CPL; ++c) {
        const int e = (lane + c * L::G) * L::V;
        if (e < d) cp_async<L::V * 4>(dst + e, src + e);
    }
}

template <class L>
__device__ __forceinline__ void k1_apply_update(float* x_row, float* s_row, float (&x)[L::NF], float (&s)[L::NF], const float (&g)[L::NF],
                                                int d, int lane, int opt, float lr) {
    if (opt == PK_ADAGRAD) {
#pragma unroll
        for (int i = 0; i < L::NF; ++i) {
            s[i] = fmaf(g[i], g[i], s[i]);
            x[i] = fmaf(-lr * g[i], rcp_nr(sqrt0(s[i]) + 1e-10f), x[i]);
        }
        st_row<L>(s_row, d, lane, s);
    } else {
#pragma unroll
        for (int i = 0; i < L::NF; ++i) x[i] = fmaf(-lr, g[i], x[i]);
    }
    st_row<L>(x_row, d, lane, x);
}

// Entity-row handle of K1: table row id, occurrence code (-1 single, else multiply-occurring) and the
// operand's slot in the staged tile (0 h, 1 t, 2 first negative; -1 = not staged, read from global).
struct K1Tgt {
    int32_t id, code, slot;
};

// Staged tile layout, per sample (rows of d floats):
//   [0]              r^ (cached)            [1] w^ / r_p (TransH / TransD only)
//   then per operand slot s in {0,1,2}: NTE table rows, and NTE optimizer-state rows when Adagrad
template <int MODEL, class L>
struct K1Ctx {
    static constexpr bool kMerge3 = false;
    static constexpr int NTE = MODEL == TRANSD ? 2 : 1, NTR = MODEL == TRANSE ? 1 : 2;
    using Tgt = K1Tgt;
    const K1Params* P;
    float* sample;          // this group's sample in the current stage
    const int32_t* cnts;    // this lane's private slot of the stage: occurrence counts of h, t, c | one id
    const int32_t* gids;    // slot of the group's lane 0: word 3 of lanes 0..3 holds h, t, r, c of the sample
    int rows_per_slot;      // NTE * (1 + adagrad)
    int copy;               // which privatised relation accumulator this block adds to
    __device__ __forceinline__ float* slot_row(int slot, int tbl, bool state) const {
        return sample + (size_t)(NTR + slot * rows_per_slot + (state ? NTE : 0) + tbl) * P->d;
    }
    __device__ __forceinline__ void load_pos(int64_t b, bool act, Tgt& th, Tgt& tt, int32_t& r) const {
        th.id = act ? gids[3] : 0;        th.code = (act && cnts[0] > 1) ? 1 : -1; th.slot = 0;
        tt.id = act ? gids[4 + 3] : 0;    tt.code = (act && cnts[1] > 1) ? 1 : -1; tt.slot = 1;
        r = act ? gids[8 + 3] : 0;
    }
    __device__ __forceinline__ bool load_neg(int j, int64_t b, bool act, Tgt& tc) const {
        const int32_t cj = act ? (j == 0 ? gids[12 + 3] : P->ids[(3 + (int64_t)j) * P->B + b]) : 0;
        tc.id = cj & 0x7fffffff;
        if (j == 0) {
            tc.code = (act && cnts[2] > 1) ? 1 : -1; tc.slot = 2;
        } else {   // further negatives (k > 1) are not staged: count and rows come straight from global
            tc.code = (act && P->cnt_ent[tc.id] > 1) ? 1 : -1; tc.slot = -1;
        }
        return cj < 0;
    }
    __device__ __forceinline__ const float* ent_row(int tbl, const Tgt& tg) const {
        return tg.slot >= 0 ? slot_row(tg.slot, tbl, false) : P->ent[tbl] + (size_t)tg.id * P->d;
    }
    __device__ __forceinline__ void prefetch(Tgt&, int, bool) const {}
    __device__ __forceinline__ const float* rel_y(int) const { return sample; }
    __device__ __forceinline__ const float* rel_w(int) const { return sample + P->d; }
    // row += g with native reductions: one 128-bit RED per chunk when the layout is float4
    __device__ __forceinline__ void red_row(float* p, const float (&g)[L::NF], int lane) const {
        const int d = P->d;
        if constexpr (L::V == 4) {
#pragma unroll
            for (int c = 0; c < L::CPL; ++c) {
                const int e = (lane + c * L::G) * 4;
                if (e < d && (g[c * 4] != 0.f || g[c * 4 + 1] != 0.f || g[c * 4 + 2] != 0.f || g[c * 4 + 3] != 0.f))
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + e), "f"(g[c * 4]), "f"(g[c * 4 + 1]),
                                 "f"(g[c * 4 + 2]), "f"(g[c * 4 + 3]) : "memory");
            }
        } else {
#pragma unroll
            for (int i = 0; i < L::NF; ++i) {
                const int e = elem_of<L>(lane, i);
                if (e < d && g[i] != 0.f) atomicAdd(p + e, g[i]);
            }
        }
    }
    __device__ __forceinline__ void rel_add(int tbl, int r, const float (&g)[L::NF], int lane) const {
        red_row(P->acc_rel + (((size_t)copy * NTR + tbl) * P->n_rel + r) * P->d, g, lane);
    }
    __device__ __forceinline__ void add_ent(int tbl, const Tgt& tg, const float (&g)[L::NF], int lane, bool pred) const {
        if (!pred) return;
        const int d = P->d;
        if (tg.code < 0) {   // the only occurrence in this batch: nobody else reads or writes the row
            float x[L::NF], s[L::NF];
            float* xg = P->ent[tbl] + (size_t)tg.id * d;
            float* sg = P->opt == PK_ADAGRAD ? P->ent_state[tbl] + (size_t)tg.id * d : nullptr;
            if (tg.slot >= 0) {
                ld_row<L>(slot_row(tg.slot, tbl, false), d, lane, x);
                ld_row<L>(slot_row(tg.slot, tbl, true), d, lane, s, P->opt == PK_ADAGRAD);
            } else {
                ld_row<L>(xg, d, lane, x);
                ld_row<L>(sg, d, lane, s, P->opt == PK_ADAGRAD);
            }
            k1_apply_update<L>(xg, sg, x, s, g, d, lane, P->opt, P->lr);
        } else {
            red_row(P->acc_ent[tbl] + (size_t)tg.id * d, g, lane);
        }
    }
};

// ---- K1 main
template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS, 2) k1_grad(const __grid_constant__ K1Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int GPW = 32 / L::G, NW = K1_THREADS / 32;
    constexpr int NTE = MODEL == TRANSD ? 2 : 1, NTR = MODEL == TRANSE ? 1 : 2;
    const int tid = threadIdx.x, lane = tid % L::G, grp = (tid % 32) / L::G, warp = tid / 32, wl = tid % 32;
    const int d = P.d;
    const bool adagrad = P.opt == PK_ADAGRAD;
    const int rows_per_slot = NTE * (adagrad ? 2 : 1);
    const int rows_per_sample = NTR + 3 * rows_per_slot;
    const size_t sample_floats = (size_t)rows_per_sample * d;
    const size_t rows_bytes = (GPW * sample_floats * 4 + 15) & ~(size_t)15;
    const size_t stage_bytes = rows_bytes + 32 * 16;   // + 4 ints per lane (occurrence counts)
    unsigned char* wbase = smem + (size_t)warp * K1_STAGES * stage_bytes;
    __shared__ float part[NW];

    Hyper hp;
    hp.d = d; hp.k = P.k; hp.p_norm = P.p_norm; hp.norm_flag = P.norm_flag;
    hp.margin = P.margin;
    hp.inv_bk = 1.f / (float)(P.B * P.k);
    K1Ctx<MODEL, L> cx;
    cx.P = &P;
    cx.rows_per_slot = rows_per_slot;
    cx.copy = blockIdx.x % P.rel_copies;

    const bool bad = P.counters[2] != 0;
    const int64_t ntiles = bad ? 0 : (P.B + GPW - 1) / GPW;
    const int64_t wglobal = (int64_t)blockIdx.x * NW + warp, wtotal = (int64_t)gridDim.x * NW;
    const int32_t* idh = P.ids;
    const int32_t* idt = P.ids + P.B;
    const int32_t* idr = P.ids + 2 * P.B;
    const int32_t* idc = P.ids + 3 * P.B;

    // ids of one sample of a tile -> registers (every lane of the group loads the same words)
    struct Meta { int32_t h, t, r, c; bool act; };
    auto load_meta = [&](int64_t tile) {
        Meta m;
        const int64_t b = tile * GPW + grp;
        m.act = tile < ntiles && b < P.B;
        m.h = m.act ? idh[b] : 0; m.t = m.act ? idt[b] : 0; m.r = m.act ? idr[b] : 0; m.c = m.act ? idc[b] : 0;
        return m;
    };
    // request everything the sample needs: cached relation operands, the three entity operands'
    // table rows (+ Adagrad state), and the occurrence counts of the three entity rows
    auto issue = [&](const Meta& m, int stage) {
        unsigned char* sb = wbase + (size_t)stage * stage_bytes;
        float* sp = reinterpret_cast<float*>(sb) + (size_t)grp * sample_floats;
        int32_t* mt = reinterpret_cast<int32_t*>(sb + rows_bytes) + wl * 4;
        if (m.act) {
            const int32_t mc = m.c & 0x7fffffff;
            if (lane < 4) mt[3] = lane == 0 ? m.h : (lane == 1 ? m.t : (lane == 2 ? m.r : m.c));
            cp_async<4>(mt + 0, P.cnt_ent + m.h);
            cp_async<4>(mt + 1, P.cnt_ent + m.t);
            cp_async<4>(mt + 2, P.cnt_ent + mc);
            stage_row<L>(sp, P.relc[0] + (size_t)m.r * d, d, lane);
            if constexpr (MODEL == TRANSH) stage_row<L>(sp + d, P.relc[1] + (size_t)m.r * d, d, lane);
            if constexpr (MODEL == TRANSD) stage_row<L>(sp + d, P.rel[1] + (size_t)m.r * d, d, lane);
            const int32_t e3[3] = {m.h, m.t, mc};
#pragma unroll
            for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int t = 0; t < NTE; ++t) {
                    float* dst = sp + (size_t)(NTR + s * rows_per_slot + t) * d;
                    stage_row<L>(dst, P.ent[t] + (size_t)e3[s] * d, d, lane);
                    if (adagrad) stage_row<L>(dst + (size_t)NTE * d, P.ent_state[t] + (size_t)e3[s] * d, d, lane);
                }
            }
        }
        cp_async_commit();
    };

    float acc = 0.f;
    Meta ma = load_meta(wglobal);                 // tile 0 of this warp
    issue(ma, 0);
    ma = load_meta(wglobal + wtotal);             // tile 1
    issue(ma, 1);
    ma = load_meta(wglobal + 2 * wtotal);         // tile 2: requested inside the loop
    int stage = 0;
    for (int64_t tile = wglobal; tile < ntiles; tile += wtotal) {
        const Meta mb = load_meta(tile + 3 * wtotal);
        issue(ma, stage >= 1 ? stage - 1 : K1_STAGES - 1);   // (stage + 2) % 3: the stage computed last iteration
        cp_async_wait<K1_STAGES - 1>();           // this tile's rows have landed
        __syncwarp();                             // ... and the ids its group's lanes stored two iterations ago are visible
        unsigned char* sb = wbase + (size_t)stage * stage_bytes;
        cx.sample = reinterpret_cast<float*>(sb) + (size_t)grp * sample_floats;
        cx.cnts = reinterpret_cast<const int32_t*>(sb + rows_bytes) + wl * 4;
        cx.gids = reinterpret_cast<const int32_t*>(sb + rows_bytes) + (wl - lane) * 4;
        const int64_t b = tile * GPW + grp;
        const bool act = b < P.B;
        const float l = train_sample<MODEL, L>(cx, hp, lane, b, act);
        if (act) acc += l;
        ma = mb;
        stage = stage + 1 == K1_STAGES ? 0 : stage + 1;
    }
    cp_async_wait<0>();
    // per-block loss partial in a fixed order: groups of a warp, then warps
    float wsum = 0.f;
#pragma unroll
    for (int g = 0; g < GPW; ++g) wsum += __shfl_sync(0xffffffffu, acc, g * L::G);
    if (wl == 0) part[warp] = wsum;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < NW; ++i) s += part[i];
        P.loss_part[blockIdx.x] = s;
    }
}

// ---- K1 tail: optimizer for multiply-occurring entity rows and for the touched relations, cleanup, loss
template <int MODEL, class L>
__device__ __forceinline__ void apply_block(const K1Params& P, int bid) {
    constexpr int NG = K1_THREADS / L::G;
    const int nblk = P.apply_blocks;
    constexpr int NTE = MODEL == TRANSD ? 2 : 1, NTR = MODEL == TRANSE ? 1 : 2;
    const int tid = threadIdx.x, lane = tid % L::G, grp = tid / L::G;
    const unsigned gmask = group_mask<L::G>(tid);
    const int d = P.d;
    const bool bad = P.counters[2] != 0;
    const int nqe = bad ? 0 : P.counters[0], nqr = bad ? 0 : P.counters[1];
    for (int64_t it = (int64_t)bid * NG + grp; it < (int64_t)nqe + nqr; it += (int64_t)nblk * NG) {
        if (it < nqr) {   // a relation of the batch: sum the privatised copies, then treat as one row
            const int r = P.touched_rel[it];
            float g0[L::NF], g1[NTR == 2 ? L::NF : 1];
#pragma unroll
            for (int t = 0; t < NTR; ++t) {
                float sum[L::NF];
#pragma unroll
                for (int i = 0; i < L::NF; ++i) sum[i] = 0.f;
                for (int c0 = 0; c0 < P.rel_copies; c0 += 4) {   // four copies' loads in flight at a time
                    float v[4][L::NF];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        ld_row<L>(P.acc_rel + (((size_t)(c0 + u) * NTR + t) * P.n_rel + r) * d, d, lane, v[u], c0 + u < P.rel_copies);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        bool nz = false;
#pragma unroll
                        for (int i = 0; i < L::NF; ++i) { sum[i] += v[u][i]; nz |= v[u][i] != 0.f; v[u][i] = 0.f; }
                        if (nz) st_row<L>(P.acc_rel + (((size_t)(c0 + u) * NTR + t) * P.n_rel + r) * d, d, lane, v[u]);
                    }
                }
#pragma unroll
                for (int i = 0; i < L::NF; ++i) {
                    if (t == 0) g0[i] = sum[i];
                    else if constexpr (NTR == 2) g1[i] = sum[i];
                }
            }
            float y[L::NF], x[L::NF], s[L::NF];
            bool fr;
            // table 0: r (the gradient arrived w.r.t. r^)
            if (P.norm_flag) {
                ld_row<L>(P.relc[0] + (size_t)r * d, d, lane, y);
                const float n = P.reln[r];
                normalize_bwd<L>(y, n, n > kNormEps, g0, gmask);
            }
            float* xr = P.rel[0] + (size_t)r * d;
            float* sr = P.opt == PK_ADAGRAD ? P.rel_state[0] + (size_t)r * d : nullptr;
            ld_row<L>(xr, d, lane, x);
            ld_row<L>(sr, d, lane, s, P.opt == PK_ADAGRAD);
            k1_apply_update<L>(xr, sr, x, s, g0, d, lane, P.opt, P.lr);
            float n0 = 1.f;
            if (P.norm_flag) n0 = normalize_row<L>(x, fr, gmask);
            st_row<L>(P.relc[0] + (size_t)r * d, d, lane, x);
            if (lane == 0) P.reln[r] = n0;
            if constexpr (NTR == 2) {
                if constexpr (MODEL == TRANSH) {   // the gradient arrived w.r.t. w^
                    ld_row<L>(P.relc[1] + (size_t)r * d, d, lane, y);
                    const float n = P.reln[P.n_rel + r];
                    normalize_bwd<L>(y, n, n > kNormEps, g1, gmask);
                }
                float* xw = P.rel[1] + (size_t)r * d;
                float* sw = P.opt == PK_ADAGRAD ? P.rel_state[1] + (size_t)r * d : nullptr;
                ld_row<L>(xw, d, lane, x);
                ld_row<L>(sw, d, lane, s, P.opt == PK_ADAGRAD);
                k1_apply_update<L>(xw, sw, x, s, g1, d, lane, P.opt, P.lr);
                if constexpr (MODEL == TRANSH) {
                    const float n1 = normalize_row<L>(x, fr, gmask);
                    st_row<L>(P.relc[1] + (size_t)r * d, d, lane, x);
                    if (lane == 0) P.reln[P.n_rel + r] = n1;
                }
            }
        } else {
            const int id = P.dup_ent[it - nqr];
#pragma unroll
            for (int t = 0; t < NTE; ++t) {
                float* arow = P.acc_ent[t] + (size_t)id * d;
                float g[L::NF], x[L::NF], s[L::NF];
                ld_row<L>(arow, d, lane, g);
                float* xr = P.ent[t] + (size_t)id * d;
                float* sr = P.opt == PK_ADAGRAD ? P.ent_state[t] + (size_t)id * d : nullptr;
                ld_row<L>(xr, d, lane, x);
                ld_row<L>(sr, d, lane, s, P.opt == PK_ADAGRAD);
                k1_apply_update<L>(xr, sr, x, s, g, d, lane, P.opt, P.lr);
#pragma unroll
                for (int i = 0; i < L::NF; ++i) g[i] = 0.f;
                st_row<L>(arow, d, lane, g);
            }
        }
    }
    // occurrence counts of this batch set back to zero (the gradient kernel was their only reader)
    if (P.clear_mode == 1) {
        for (int64_t i = (int64_t)bid * K1_THREADS + tid; i < P.n_ent; i += (int64_t)nblk * K1_THREADS) P.cnt_ent[i] = 0;
        for (int64_t i = (int64_t)bid * K1_THREADS + tid; i < P.n_rel; i += (int64_t)nblk * K1_THREADS) P.cnt_rel[i] = 0;
    } else if (P.clear_mode == 2) {
        for (int64_t i = (int64_t)bid * K1_THREADS + tid; i < P.B; i += (int64_t)nblk * K1_THREADS) {
            P.cnt_ent[P.ids[i]] = 0;
            P.cnt_ent[P.ids[P.B + i]] = 0;
            P.cnt_rel[P.ids[2 * P.B + i]] = 0;
            for (int j = 0; j < P.k; ++j) P.cnt_ent[P.ids[(3 + (int64_t)j) * P.B + i] & 0x7fffffff] = 0;
        }
    }
    // the last block to get here closes the step: loss, queues emptied
    __shared__ int last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last = atomicAdd(&P.counters[3], 1) == nblk - 1;
    }
    __syncthreads();
    if (last && tid < 64) {
        if (tid < 32) {
            float s = 0.f;
            for (int i = tid; i < P.grad_blocks; i += 32) s += P.loss_part[i];
            s = gsum<32>(s);
            if (tid == 0) {
                const int64_t step = *P.step_ctr;
                if (P.loss) P.loss[step] = bad ? nanf("") : s / (float)(P.B * P.k) + P.margin;
                *P.step_ctr = step + 1;
            }
        }
        if (tid < 2) P.counters[tid] = 0;
        if (tid == 3) P.counters[3] = 0;
    }
}

template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS) k1_apply(const __grid_constant__ K1Params P) {
    apply_block<MODEL, L>(P, (int)blockIdx.x);
}

// The optimizer tail of step s and the preparation of batch s + 1 in ONE launch: both are latency-bound
// (IPC ~0.4), so blocks of the two roles share the SMs well; the first P.apply_blocks blocks take the
// tail, the rest the batch (which lives in the other batch set and does not depend on the embeddings).
template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS) k1_apply_prepare(const __grid_constant__ K1Params P, const __grid_constant__ PrepParams S) {
    if ((int)blockIdx.x < P.apply_blocks) apply_block<MODEL, L>(P, (int)blockIdx.x);
    else prepare_block(S, (int)blockIdx.x - P.apply_blocks);
}

// cached relation operands from the tables (start of every pk_train_step / pk_train_steps call)
template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS) k1_relcache(const __grid_constant__ K1Params P) {
    constexpr int NG = K1_THREADS / L::G;
    const int tid = threadIdx.x, lane = tid % L::G, grp = tid / L::G;
    const unsigned gmask = group_mask<L::G>(tid);
    const int d = P.d;
    for (int64_t r = (int64_t)blockIdx.x * NG + grp; r < P.n_rel; r += (int64_t)gridDim.x * NG) {
        float x[L::NF];
        bool fr;
        ld_row<L>(P.rel[0] + (size_t)r * d, d, lane, x);
        float n = 1.f;
        if (P.norm_flag) n = normalize_row<L>(x, fr, gmask);
        st_row<L>(P.relc[0] + (size_t)r * d, d, lane, x);
        if (lane == 0) P.reln[r] = n;
        if constexpr (MODEL == TRANSH) {
            ld_row<L>(P.rel[1] + (size_t)r * d, d, lane, x);
            n = normalize_row<L>(x, fr, gmask);
            st_row<L>(P.relc[1] + (size_t)r * d, d, lane, x);
            if (lane == 0) P.reln[P.n_rel + r] = n;
        }
    }
}

#endif

#ifndef PK_MODEL_TU
// jump[0..per) = A_j, jump[per..2per) = C_j : j samples into a slice; then Aadv[64], Cadv[64]: one batch
__global__ void k1_build_jump(uint64_t* jump, int64_t per, int64_t B, int W, int k) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < per) lcg_affine((uint64_t)j * (uint64_t)(1 + 2 * k), jump[j], jump[per + j]);
    if (j < W) {
        int64_t lef, rig;
        slice_of(B, W, (int)j, lef, rig);
        lcg_affine((uint64_t)(rig - lef) * (uint64_t)(1 + 2 * k), jump[2 * per + j], jump[2 * per + 64 + j]);
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_prepare(const __grid_constant__ PrepParams S) { prepare_block(S, (int)blockIdx.x); }

// ---- K0 stand-alone: one reference sampling() call in the reference's output layout
struct SampleParams {
    SamplerView sv;
    const uint64_t* lcg;
    int32_t* bh;
    int32_t* bt;
    int32_t* br;
    int64_t B;
    int W, k, bern, filter;
};

__global__ void __launch_bounds__(K1_THREADS) k0_sample(const __grid_constant__ SampleParams S) {
    const int64_t b = (int64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    if (b >= S.B) return;
    int64_t j;
    const int id = stream_of(S.B, S.W, b, j);
    sample_one(S.sv, S.lcg[id], j, S.B, S.k, S.bern != 0, S.filter != 0, b, S.bh, S.bt, S.br);
}

// advance the W stream states by one batch
__global__ void k0_commit_lcg(uint64_t* lcg, int64_t B, int W, int k) {
    const int id = threadIdx.x;
    uint64_t s = 0;
    if (id < W) s = lcg_advance_batch(lcg[id], W, id, B, k);
    __syncthreads();
    if (id < W) lcg[id] = s;
}

#endif

struct LaySel { int V, G, CPL; };
inline LaySel pick_layout(int model, int d, int opt = PK_SGD) {
    const int V = d % 4 == 0 ? 4 : (d % 2 == 0 ? 2 : 1);
    const int chunks = d / V;
    const int nf_cap = model == TRANSD ? 4 : 8;
    // 9..16 float4 chunks (d = 36..64): 16 lanes with one chunk each halve the registers and the staged
    // bytes per warp against 8 lanes x 2 chunks, so twice as many warps are resident.  Measured on S1
    // (d = 64): Adagrad 119 vs 135 us/step, SGD 112 vs 101 — so only where the staged state rows make
    // shared memory the occupancy limit.
    if (V == 4 && chunks > 8 && chunks <= 16 && opt == PK_ADAGRAD) return LaySel{4, 16, 1};
    for (int G : {8, 32}) {
        int cpl = (chunks + G - 1) / G, c2 = 1;
        while (c2 < cpl) c2 *= 2;
        if (c2 * V <= nf_cap || G == 32) return LaySel{V, G, std::max(c2, 1)};
    }
    return LaySel{V, 32, 1};
}

// shared memory of one k1_grad block (must match the carve-up inside the kernel)
inline size_t grad_smem(int model, const LaySel& l, int d, int opt) {
    const int gpw = 32 / l.G, nte = model == PK_TRANSD ? 2 : 1, ntr = model == PK_TRANSE ? 1 : 2;
    const size_t rows = ntr + 3 * nte * (opt == PK_ADAGRAD ? 2 : 1);
    const size_t stage = (((size_t)gpw * rows * d * 4 + 15) & ~(size_t)15) + 32 * 16;
    return stage * K1_STAGES * (K1_THREADS / 32);
}

#ifdef PK_MODEL_TU
template <int MODEL, int V, int G, int CPL>
int launch_step(const K1Params& P, const PrepParams* S, int what, int grad_blocks, int apply_blocks, size_t smem, cudaStream_t st) {
    using L = Lay<V, G, CPL>;
    if (what == 5) {
        k1_apply_prepare<MODEL, L><<<apply_blocks + S->nblocks, K1_THREADS, 0, st>>>(P, *S);
        PK_LAUNCHED("k1_apply_prepare");
        return PK_OK;
    }
    if (what == 0) {
        k1_relcache<MODEL, L><<<apply_blocks, K1_THREADS, 0, st>>>(P);
        PK_LAUNCHED("k1_relcache");
        return PK_OK;
    }
    if (what == 2) {   // occupancy query: resident blocks per SM for this shared-memory size
        int nb = 0;
        PK_CUDA(cudaFuncSetAttribute(k1_grad<MODEL, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1_grad<MODEL, L>, K1_THREADS, smem));
        return nb;
    }
    if (what == 1 || what == 3) {
        k1_grad<MODEL, L><<<grad_blocks, K1_THREADS, smem, st>>>(P);
        PK_LAUNCHED("k1_grad");
    }
    if (what == 1 || what == 4) {
        k1_apply<MODEL, L><<<apply_blocks, K1_THREADS, 0, st>>>(P);
        PK_LAUNCHED("k1_apply");
    }
    return PK_OK;
}

template <int MODEL>
int dispatch_step(const LaySel& l, const K1Params& P, const PrepParams* S, int what, int gb, int ab, size_t smem, cudaStream_t st) {
#define PK_CASE(v, g, c) if (l.V == v && l.G == g && l.CPL == c) return launch_step<MODEL, v, g, c>(P, S, what, gb, ab, smem, st);
    PK_CASE(4, 16, 1)
    PK_CASE(4, 8, 1) PK_CASE(4, 8, 2) PK_CASE(4, 32, 1) PK_CASE(4, 32, 2)
    PK_CASE(2, 8, 1) PK_CASE(2, 8, 2) PK_CASE(2, 8, 4) PK_CASE(2, 32, 1) PK_CASE(2, 32, 2) PK_CASE(2, 32, 4)
    PK_CASE(1, 8, 1) PK_CASE(1, 8, 2) PK_CASE(1, 8, 4) PK_CASE(1, 8, 8) PK_CASE(1, 32, 1) PK_CASE(1, 32, 2) PK_CASE(1, 32, 4) PK_CASE(1, 32, 8)
#undef PK_CASE
    return pk::fail(PK_ERR_UNSUPPORTED, "embedding dimension not supported by the train step (d <= 256)");
}

#define PK_CAT2(a, b) a##b
#define PK_CAT(a, b) PK_CAT2(a, b)
int PK_CAT(step_model, PK_MODEL_TU)(const LaySel& l, const K1Params& P, const PrepParams* S, int what, int gb, int ab, size_t smem, cudaStream_t st) {
    return dispatch_step<PK_MODEL_TU>(l, P, S, what, gb, ab, smem, st);
}
}  // namespace pkk1
#else
int step_model0(const LaySel& l, const K1Params& P, const PrepParams* S, int what, int gb, int ab, size_t smem, cudaStream_t st);
int step_model1(const LaySel& l, const K1Params& P, const PrepParams* S, int what, int gb, int ab, size_t smem, cudaStream_t st);
int step_model2(const LaySel& l, const K1Params& P, const PrepParams* S, int what, int gb, int ab, size_t smem, cudaStream_t st);

int step_model(int model, const LaySel& l, const K1Params& P, int what, int gb, int ab, size_t smem, cudaStream_t st,
               const PrepParams* S = nullptr) {
    if (model == PK_TRANSE) return step_model0(l, P, S, what, gb, ab, smem, st);
    if (model == PK_TRANSH) return step_model1(l, P, S, what, gb, ab, smem, st);
    return step_model2(l, P, S, what, gb, ab, smem, st);
}

int check_cfg(const pk_model_cfg* cfg, const char* who) {
    if (!cfg) return pk::fail(PK_ERR_ARG, std::string(who) + ": null cfg");
    if (cfg->model < 0 || cfg->model > 2) return pk::fail(PK_ERR_ARG, std::string(who) + ": unknown model");
    if (cfg->dim <= 0 || cfg->dim > 256) return pk::fail(PK_ERR_UNSUPPORTED, std::string(who) + ": dim must be in [1,256]");
    if (cfg->p_norm != 1 && cfg->p_norm != 2) return pk::fail(PK_ERR_UNSUPPORTED, std::string(who) + ": p_norm must be 1 or 2");
    if (cfg->neg_ent < 1) return pk::fail(PK_ERR_ARG, std::string(who) + ": neg_ent must be >= 1");
    if (cfg->opt != PK_SGD && cfg->opt != PK_ADAGRAD) return pk::fail(PK_ERR_UNSUPPORTED, std::string(who) + ": optimizer must be SGD or Adagrad");
    return PK_OK;
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

void use_set(K1Params& P, pk_workspace* ws, int set) {
    P.cnt_ent = ws->cnt_ent[set]; P.cnt_rel = ws->cnt_rel[set];
    P.dup_ent = ws->dup_ent[set]; P.touched_rel = ws->touched_rel[set];
    P.counters = ws->counters[set];
    P.ids = ws->ids[set];
}

void fill_params(K1Params& P, const pk_model_cfg* cfg, const pk_tables* tab, pk_workspace* ws, int64_t B, float margin, float lr,
                 float* d_loss) {
    for (int i = 0; i < 2; ++i) {
        P.ent[i] = tab->ent[i]; P.rel[i] = tab->rel[i];
        P.ent_state[i] = cfg->opt == PK_ADAGRAD ? tab->ent_state[i] : nullptr;
        P.rel_state[i] = cfg->opt == PK_ADAGRAD ? tab->rel_state[i] : nullptr;
        P.acc_ent[i] = ws->acc_ent[i];
        P.relc[i] = ws->relc[i];
    }
    P.acc_rel = ws->acc_rel; P.reln = ws->reln;
    use_set(P, ws, 0);
    P.step_ctr = ws->step_ctr;
    P.loss_part = ws->loss_part;
    P.loss = d_loss;
    P.B = B; P.n_ent = ws->n_ent; P.n_rel = ws->n_rel;
    P.d = cfg->dim; P.k = cfg->neg_ent; P.p_norm = cfg->p_norm; P.norm_flag = cfg->norm_flag; P.opt = cfg->opt;
    P.margin = margin; P.lr = lr;
    P.rel_copies = ws->rel_copies;
    P.grad_blocks = 1;
    P.apply_blocks = 1;
    P.clear_mode = ws->memset_reset ? 1 : 2;
}

int check_tables(const pk_model_cfg* cfg, const pk_tables* tab, const pk_workspace* ws, const char* who) {
    if (!tab || !ws) return pk::fail(PK_ERR_ARG, std::string(who) + ": null tables/workspace");
    const int ntE = cfg->model == PK_TRANSD ? 2 : 1, ntR = cfg->model == PK_TRANSE ? 1 : 2;
    for (int i = 0; i < ntE; ++i)
        if (!tab->ent[i] || (cfg->opt == PK_ADAGRAD && !tab->ent_state[i])) return pk::fail(PK_ERR_ARG, std::string(who) + ": missing entity table/state");
    for (int i = 0; i < ntR; ++i)
        if (!tab->rel[i] || (cfg->opt == PK_ADAGRAD && !tab->rel_state[i])) return pk::fail(PK_ERR_ARG, std::string(who) + ": missing relation table/state");
    if (tab->n_ent != ws->n_ent || tab->n_rel != ws->n_rel) return pk::fail(PK_ERR_ARG, std::string(who) + ": workspace was created for another table shape");
    if (ws->cfg.model != cfg->model || ws->cfg.dim != cfg->dim || ws->cfg.neg_ent != cfg->neg_ent)
        return pk::fail(PK_ERR_ARG, std::string(who) + ": workspace was created for another model configuration");
    return PK_OK;
}

// launch geometry of one step; the shared memory (hence the resident blocks) depends on the optimizer
struct StepGeom {
    LaySel lay;
    size_t smem;
    int grad_blocks, apply_blocks;
};

int step_geometry(const pk_model_cfg* cfg, const K1Params& P, pk_workspace* ws, StepGeom& g) {
    g.lay = pick_layout(cfg->model, cfg->dim, cfg->opt);
    g.smem = grad_smem(cfg->model, g.lay, cfg->dim, cfg->opt);
    if (g.smem > 227 * 1024) return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_step: embedding dimension too large for the staged train step");
    const int per_sm = step_model(cfg->model, g.lay, P, 2, 0, 0, g.smem, nullptr);
    if (per_sm < 0) return per_sm;
    const int gpw = 32 / g.lay.G, nw = K1_THREADS / 32;
    const int64_t tiles = (P.B + gpw - 1) / gpw;
    g.grad_blocks = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + nw - 1) / nw, (int64_t)num_sms() * std::max(per_sm, 1)));
    g.grad_blocks = std::min(g.grad_blocks, ws->max_blocks);
    g.apply_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ws->max_blocks, (P.B * (2 + P.k) / 2 + ws->n_rel + 31) / 32));
    return PK_OK;
}

int ensure_jump(pk_workspace* ws, int64_t B, int W, int k, cudaStream_t st) {
    if (ws->jump_B == B && ws->jump_W == W && ws->jump_k == k) return PK_OK;
    const int64_t per = (B % W == 0) ? B / W : B / W + 1;
    const int64_t need = 2 * per + 128;
    if (ws->jump_cap < need) {
        if (ws->jump) cudaFree(ws->jump);
        ws->jump = nullptr;
        ws->jump_cap = 0;
        PK_CUDA(cudaMalloc((void**)&ws->jump, (size_t)need * 8));
        ws->jump_cap = need;
    }
    const int64_t n = std::max<int64_t>(per, W);
    k1_build_jump<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws->jump, per, B, W, k);
    PK_LAUNCHED("k1_build_jump");
    ws->jump_B = B; ws->jump_W = W; ws->jump_k = k;
    return PK_OK;
}

void fill_prep(PrepParams& S, const pk_model_cfg* cfg, pk_workspace* ws, int64_t B) {
    S.sv.by_head = nullptr; S.sv.by_tail = nullptr; S.sv.left_mean = nullptr; S.sv.right_mean = nullptr;
    S.sv.n_tri = 0; S.sv.n_ent = (int32_t)ws->n_ent; S.sv.n_rel = (int32_t)ws->n_rel;
    S.sv.head_off = nullptr; S.sv.tail_off = nullptr;
    S.lcg = nullptr; S.jump = ws->jump; S.per = 0;
    S.gh = S.gt = S.gr = nullptr;
    S.B = B; S.n_ent = ws->n_ent; S.n_rel = ws->n_rel;
    S.k = cfg->neg_ent; S.bern = cfg->bern; S.filter = cfg->filter;
    S.ahead = 0;
    S.W = cfg->work_threads;
    S.lcg_out = nullptr;
    S.nblocks = (int)((B + K1_THREADS - 1) / K1_THREADS);
}

void use_set(PrepParams& S, pk_workspace* ws, int set) {
    S.ids = ws->ids[set]; S.cnt_ent = ws->cnt_ent[set]; S.cnt_rel = ws->cnt_rel[set];
    S.dup_ent = ws->dup_ent[set]; S.touched_rel = ws->touched_rel[set]; S.counters = ws->counters[set];
}

}  // namespace pkk1

using namespace pkk1;

extern "C" pk_workspace* pk_workspace_create(const pk_model_cfg* cfg, int64_t n_ent, int64_t n_rel, int64_t max_batch) {
    if (check_cfg(cfg, "pk_workspace_create") != PK_OK) return nullptr;
    if (n_ent < 2 || n_rel < 1 || max_batch < 1) {
        pk::fail(PK_ERR_ARG, "pk_workspace_create: bad sizes");
        return nullptr;
    }
    pk_workspace* ws = new pk_workspace();
    ws->cfg = *cfg;
    ws->n_ent = n_ent; ws->n_rel = n_rel; ws->max_batch = max_batch;
    const int d = cfg->dim, k = cfg->neg_ent;
    const int ntE = cfg->model == PK_TRANSD ? 2 : 1, ntR = cfg->model == PK_TRANSE ? 1 : 2;
    ws->dup_cap_ent = std::min<int64_t>(n_ent, max_batch * (2 + k)) + 4;
    ws->max_blocks = num_sms() * 8;
    // privatised relation accumulators: up to 32 copies, at most ~64 MB in total
    const size_t one = (size_t)ntR * n_rel * d * 4;
    ws->rel_copies = (int)std::max<size_t>(1, std::min<size_t>(32, (64u << 20) / std::max<size_t>(one, 1)));
    auto alloc0 = [&](void** p, size_t bytes) -> bool {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        return cudaMemset(*p, 0, bytes) == cudaSuccess;
    };
    // clearing the counts with two memsets beats scattering over the batch unless the tables are huge
    ws->memset_reset = n_ent <= 16 * (int64_t)(2 + k) * max_batch;
    bool ok = alloc0((void**)&ws->step_ctr, 8) && alloc0((void**)&ws->loss_part, (size_t)ws->max_blocks * 4) &&
              alloc0((void**)&ws->acc_rel, one * ws->rel_copies) && alloc0((void**)&ws->reln, (size_t)2 * n_rel * 4);
    for (int s = 0; ok && s < 2; ++s)
        ok = alloc0((void**)&ws->cnt_ent[s], (size_t)n_ent * 4) && alloc0((void**)&ws->cnt_rel[s], (size_t)n_rel * 4) &&
             alloc0((void**)&ws->dup_ent[s], (size_t)ws->dup_cap_ent * 4) && alloc0((void**)&ws->touched_rel[s], (size_t)n_rel * 4) &&
             alloc0((void**)&ws->counters[s], 16) && alloc0((void**)&ws->ids[s], (size_t)(3 + k) * max_batch * 4);
    for (int i = 0; ok && i < ntE; ++i) ok = alloc0((void**)&ws->acc_ent[i], (size_t)n_ent * d * 4);
    for (int i = 0; ok && i < (cfg->model == PK_TRANSH ? 2 : 1); ++i) ok = alloc0((void**)&ws->relc[i], (size_t)n_rel * d * 4);
    ok = ok && cudaStreamCreate(&ws->own_stream) == cudaSuccess;
    if (!ok) {
        pk::cuda_fail(cudaGetLastError(), "pk_workspace_create: cudaMalloc");
        pk_workspace_free(ws);
        return nullptr;
    }
    return ws;
}

extern "C" void pk_workspace_free(pk_workspace* ws) {
    if (!ws) return;
    cudaFree(ws->step_ctr); cudaFree(ws->loss_part);
    cudaFree(ws->acc_rel); cudaFree(ws->reln); cudaFree(ws->jump);
    for (int i = 0; i < 2; ++i) {
        cudaFree(ws->acc_ent[i]); cudaFree(ws->relc[i]);
        cudaFree(ws->cnt_ent[i]); cudaFree(ws->cnt_rel[i]); cudaFree(ws->dup_ent[i]); cudaFree(ws->touched_rel[i]);
        cudaFree(ws->counters[i]); cudaFree(ws->ids[i]);
    }
    if (ws->graph_exec) {
        cudaDeviceSynchronize();
        cudaGraphExecDestroy(ws->graph_exec);
        cudaGraphDestroy(ws->graph);
    }
    if (ws->own_stream) cudaStreamDestroy(ws->own_stream);
    delete ws;
}

extern "C" int pk_sample_batch(const pk_model_cfg* cfg, const pk_sampler* smp, int64_t B, int32_t* d_h, int32_t* d_t,
                               int32_t* d_r, void* stream) {
    pk::launch_counter() = 0;
    if (!cfg || !smp || !d_h || !d_t || !d_r) return pk::fail(PK_ERR_ARG, "pk_sample_batch: null argument");
    if (B < 1 || cfg->neg_ent < 0) return pk::fail(PK_ERR_ARG, "pk_sample_batch: bad batch size / neg_ent");
    if (cfg->work_threads < 1 || cfg->work_threads > 64) return pk::fail(PK_ERR_UNSUPPORTED, "pk_sample_batch: work_threads must be in [1,64]");
    if (smp->n_ent < 2 || smp->n_tri < 1) return pk::fail(PK_ERR_ARG, "pk_sample_batch: empty index");
    if (cfg->filter && !smp->by_tail) return pk::fail(PK_ERR_ARG, "pk_sample_batch: filter needs the (t,r,h) index");
    if (cfg->bern && (!smp->left_mean || !smp->right_mean)) return pk::fail(PK_ERR_ARG, "pk_sample_batch: bern needs the relation means");
    cudaStream_t st = (cudaStream_t)stream;
    SampleParams S;
    S.sv.by_head = smp->by_head; S.sv.by_tail = smp->by_tail; S.sv.left_mean = smp->left_mean; S.sv.right_mean = smp->right_mean;
    S.sv.n_tri = smp->n_tri; S.sv.n_ent = (int32_t)smp->n_ent; S.sv.n_rel = (int32_t)smp->n_rel;
    S.sv.head_off = smp->head_off; S.sv.tail_off = smp->tail_off;
    S.lcg = smp->lcg;
    S.bh = d_h; S.bt = d_t; S.br = d_r;
    S.B = B; S.W = cfg->work_threads; S.k = cfg->neg_ent; S.bern = cfg->bern; S.filter = cfg->filter;
    k0_sample<<<(unsigned)((B + K1_THREADS - 1) / K1_THREADS), K1_THREADS, 0, st>>>(S);
    PK_LAUNCHED("k0_sample");
    k0_commit_lcg<<<1, 64, 0, st>>>(smp->lcg, B, cfg->work_threads, cfg->neg_ent);
    PK_LAUNCHED("k0_commit_lcg");
    return PK_OK;
}

extern "C" int pk_train_step(const pk_model_cfg* cfg, const pk_tables* tab, pk_workspace* ws, int64_t B, const int32_t* d_h,
                             const int32_t* d_t, const int32_t* d_r, float margin, float lr, float* d_loss, void* stream) {
    pk::launch_counter() = 0;
    int rc = check_cfg(cfg, "pk_train_step");
    if (rc != PK_OK) return rc;
    rc = check_tables(cfg, tab, ws, "pk_train_step");
    if (rc != PK_OK) return rc;
    if (!d_h || !d_t || !d_r) return pk::fail(PK_ERR_ARG, "pk_train_step: null ids");
    if (B < 1 || B > ws->max_batch) return pk::fail(PK_ERR_ARG, "pk_train_step: batch size exceeds the workspace");
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P;
    fill_params(P, cfg, tab, ws, B, margin, lr, d_loss);
    StepGeom g;
    rc = step_geometry(cfg, P, ws, g);
    if (rc != PK_OK) return rc;
    P.grad_blocks = g.grad_blocks;
    P.apply_blocks = g.apply_blocks;
    PrepParams S;
    fill_prep(S, cfg, ws, B);
    use_set(S, ws, 0);
    S.gh = d_h; S.gt = d_t; S.gr = d_r;
    PK_CUDA(cudaMemsetAsync(ws->step_ctr, 0, 8, st));   // single steps write loss[0]
    rc = step_model(cfg->model, g.lay, P, 0, 0, g.apply_blocks, 0, st);   // relation cache from the live tables
    if (rc != PK_OK) return rc;
    k1_prepare<<<(unsigned)((B + K1_THREADS - 1) / K1_THREADS), K1_THREADS, 0, st>>>(S);
    PK_LAUNCHED("k1_prepare");
    return step_model(cfg->model, g.lay, P, 1, g.grad_blocks, g.apply_blocks, g.smem, st);   // grad + apply (clears the counts)
}

extern "C" int pk_train_steps(const pk_model_cfg* cfg, const pk_tables* tab, const pk_sampler* smp, pk_workspace* ws, int64_t B,
                              int64_t steps, float margin, float lr, float* d_loss, void* stream) {
    pk::launch_counter() = 0;
    int rc = check_cfg(cfg, "pk_train_steps");
    if (rc != PK_OK) return rc;
    rc = check_tables(cfg, tab, ws, "pk_train_steps");
    if (rc != PK_OK) return rc;
    if (!smp || !smp->by_head || !smp->lcg) return pk::fail(PK_ERR_ARG, "pk_train_steps: null sampler");
    if (B < 1 || B > ws->max_batch) return pk::fail(PK_ERR_ARG, "pk_train_steps: batch size exceeds the workspace");
    if (steps < 0) return pk::fail(PK_ERR_ARG, "pk_train_steps: negative step count");
    if (cfg->work_threads < 1 || cfg->work_threads > 64) return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_steps: work_threads must be in [1,64]");
    if (smp->n_ent != ws->n_ent || smp->n_rel != ws->n_rel) return pk::fail(PK_ERR_ARG, "pk_train_steps: sampler and tables disagree on the id space");
    if (cfg->filter && !smp->by_tail) return pk::fail(PK_ERR_ARG, "pk_train_steps: filter needs the (t,r,h) index");
    if (cfg->bern && (!smp->left_mean || !smp->right_mean)) return pk::fail(PK_ERR_ARG, "pk_train_steps: bern needs the relation means");
    if (steps == 0) return PK_OK;
    // Stream capture is not allowed on the legacy default stream (which is what torch hands out by
    // default); a blocking stream of our own is implicitly ordered against it in both directions.
    cudaStream_t st = (cudaStream_t)stream;
    if (st == nullptr || st == cudaStreamLegacy) st = ws->own_stream;
    const int k = cfg->neg_ent, W = cfg->work_threads;
    rc = ensure_jump(ws, B, W, k, st);
    if (rc != PK_OK) return rc;
    K1Params P;
    memset(&P, 0, sizeof(P));
    fill_params(P, cfg, tab, ws, B, margin, lr, d_loss);
    StepGeom g;
    memset(&g, 0, sizeof(g));
    rc = step_geometry(cfg, P, ws, g);
    if (rc != PK_OK) return rc;
    P.grad_blocks = g.grad_blocks;
    const int64_t per = (B % W == 0) ? B / W : B / W + 1;
    PrepParams S;
    memset(&S, 0, sizeof(S));
    fill_prep(S, cfg, ws, B);
    S.sv.by_head = smp->by_head; S.sv.by_tail = smp->by_tail; S.sv.left_mean = smp->left_mean; S.sv.right_mean = smp->right_mean;
    S.sv.n_tri = smp->n_tri;
    S.sv.head_off = smp->head_off; S.sv.tail_off = smp->tail_off;
    S.lcg = smp->lcg; S.per = per;
    const unsigned sb = (unsigned)((B + K1_THREADS - 1) / K1_THREADS);
    PK_CUDA(cudaMemsetAsync(ws->step_ctr, 0, 8, st));
    rc = step_model(cfg->model, g.lay, P, 0, 0, g.apply_blocks, 0, st);
    if (rc != PK_OK) return rc;
    // Schedule: batch s lives in batch set s % 2.  prepare(0) runs alone; then every step is
    //   k1_grad(set s)  ->  k1_apply_prepare: tail of step s  ||  prepare(batch s + 1 into the other set)
    // and the last step ends with the plain tail.  Each prepare advances the sampler streams once all
    // its blocks have read them, so after N steps they stand exactly N batches further.
    P.apply_blocks = g.apply_blocks;
    auto prepare_alone = [&](int set) -> int {
        use_set(S, ws, set);
        k1_prepare<<<sb, K1_THREADS, 0, st>>>(S);
        PK_LAUNCHED("k1_prepare");
        return PK_OK;
    };
    auto step = [&](int64_t s, bool fused) -> int {
        const int set = (int)(s & 1);
        use_set(P, ws, set);
        int r2 = step_model(cfg->model, g.lay, P, 3, g.grad_blocks, g.apply_blocks, g.smem, st);   // grad
        if (r2 != PK_OK) return r2;
        if (!fused) return step_model(cfg->model, g.lay, P, 4, g.grad_blocks, g.apply_blocks, g.smem, st);   // plain tail
        use_set(S, ws, set ^ 1);
        return step_model(cfg->model, g.lay, P, 5, g.grad_blocks, g.apply_blocks, g.smem, st, &S);      // tail || next batch
    };
    S.lcg_out = smp->lcg;
    rc = prepare_alone(0);
    if (rc != PK_OK) return rc;
    // Every launch parameter is step-invariant (the step index and the sampler streams live on the
    // device), so a chunk of fused steps is captured once into a CUDA graph and replayed; the chunk is
    // even, so every replay starts on batch set 0.  The final step is never part of a replay.
    const int64_t fused_steps = steps - 1;
    const int64_t chunk = fused_steps >= 64 ? 64 : (fused_steps & ~(int64_t)1);
    int64_t done = 0;
    if (chunk >= 4) {
        // key = every byte that ends up in a launch: kernel parameters (set 0), geometry, chunk, stream
        use_set(P, ws, 0);
        use_set(S, ws, 1);
        std::vector<unsigned char> key(sizeof(P) + sizeof(S) + sizeof(g) + sizeof(chunk) + sizeof(st) + sizeof(int));
        unsigned char* kp = key.data();
        memcpy(kp, &P, sizeof(P)); kp += sizeof(P);
        memcpy(kp, &S, sizeof(S)); kp += sizeof(S);
        memcpy(kp, &g, sizeof(g)); kp += sizeof(g);
        memcpy(kp, &chunk, sizeof(chunk)); kp += sizeof(chunk);
        memcpy(kp, &st, sizeof(st)); kp += sizeof(st);
        memcpy(kp, &cfg->model, sizeof(int));
        if (!ws->graph_exec || ws->graph_key != key) {
            if (ws->graph_exec) {   // parameters changed: the old graph may still be queued on its stream
                cudaDeviceSynchronize();
                cudaGraphExecDestroy(ws->graph_exec);
                cudaGraphDestroy(ws->graph);
                ws->graph_exec = nullptr;
                ws->graph = nullptr;
            }
            const int before = pk::launch_counter();
            PK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int64_t i = 0; i < chunk && rc == PK_OK; ++i) rc = step(i, true);
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc != PK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return pk::cuda_fail(ce, "cudaStreamEndCapture");
            ws->graph_launches = pk::launch_counter() - before;
            pk::launch_counter() = before;
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            if (ce != cudaSuccess) { cudaGraphDestroy(graph); return pk::cuda_fail(ce, "cudaGraphInstantiate"); }
            ws->graph = graph;
            ws->graph_exec = exec;
            ws->graph_key = key;
        }
        for (; done + chunk <= fused_steps; done += chunk) {
            PK_CUDA(cudaGraphLaunch(ws->graph_exec, st));
            pk::launch_counter() += ws->graph_launches;
        }
    }
    for (; done < fused_steps && rc == PK_OK; ++done) rc = step(done, true);
    if (rc == PK_OK) rc = step(steps - 1, false);
    return rc;
}

extern "C" int pk_workspace_check(pk_workspace* ws, void* stream) {
    if (!ws) return pk::fail(PK_ERR_ARG, "pk_workspace_check: null workspace");
    int32_t c[4] = {0, 0, 0, 0};
    PK_CUDA(cudaMemcpyAsync(c, ws->counters[0], 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (c[2]) {
        cudaMemsetAsync(ws->counters[0] + 2, 0, 4, (cudaStream_t)stream);
        return pk::fail(PK_ERR_ARG, "train step refused a batch: an id is out of range, a negative does not share its positive's relation, or it replaces both entities");
    }
    return PK_OK;
}
#endif  // !PK_MODEL_TU
