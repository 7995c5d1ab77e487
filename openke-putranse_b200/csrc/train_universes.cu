// K2 — batched universe training: ONE launch trains many PuTransE universes; one thread block per
// universe runs all of that universe's epochs x nbatches steps back to back (sampling, forward,
// analytic backward, Adagrad/SGD update) without ever returning to the host.
//
// Replaces, per universe, the reference's Python loop
//   Parallel_Universe_Config.train_embedding_space -> Trainer.run -> train_one_step
//   (openke/config/Parallel_Universe_Config.py:228-258, openke/config/Trainer.py:58-104,44-56)
// and the per-step ctypes call into sampling() (openke/base/Base.cpp:266-310), ~100-190 ATen
// launches per step.
//
// A universe's steps are strictly sequential and one step touches < 100 KB, so the bound is the
// LATENCY of one step on one SM, not bandwidth.  The block is therefore split into two roles that
// run concurrently (warp specialisation, named barriers):
//
//   producer warps   draw batch s+1 with the reference's LCG streams (bit-exact).  A producer thread owns
//                    the same samples of every batch, so its stream position lives in registers and
//                    advances by one affine map per batch (no jump tables, no division, multiply-high
//                    modulo); the triple-index loads of the batch after next are issued a call early.
//                    Which table rows occur more than once is found WITHOUT read-modify-write chains:
//                    every occurrence stores its index into the row's word (last store = the row's
//                    owner), losers flag the word, a third pass reads the verdict.  None of this
//                    depends on the embeddings, so it is off the critical path.
//   consumer warps   step s: one lane group per positive sample gathers h, t, r and the corrupted
//                    entity (shared memory), forward, margin loss, analytic backward.  A row that
//                    occurs ONCE in the batch is updated in place by the group that read it, with
//                    its Adagrad state prefetched from L2 at the start of the sample; a row that
//                    occurs several times accumulates into the scratch row of its owner occurrence with
//                    fire-and-forget L2 reductions (RED.ADD.F32; a shared-memory float add is a
//                    compare-and-swap loop) and is updated after a consumer-only barrier.  Samples
//                    whose margin term is switched off skip the backward pass (all their gradients
//                    are exactly zero).
//
// Data layout.  All universes of a launch are packed: entity tables [sum nE, d], relation tables
// [sum nR, d], sorted triple lists [sum nT, 3]; a descriptor per universe holds the offsets.  A
// block stages its universe's tables in shared memory when they fit (local ids are dense, so the
// staged table IS the universe's whole embedding space).  The Adagrad accumulators of the entity
// tables and the scratch rows stay in global memory (L2-resident).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "common.hpp"
#include "kge_device.cuh"

namespace pkk2 {

using namespace pkd;

struct K2Params {
    const pk_universe_desc* desc;
    float* ent[2];
    float* rel[2];
    float* ent_state[2];
    float* rel_state[2];
    const int32_t* by_head;
    const int32_t* by_tail;
    const float* left_mean;
    const float* right_mean;
    float* loss;
    int d, k, p_norm, norm_flag, opt, bern, filter, W;
    int mE, mR, mB;         // launch-wide maxima: the shared-memory carve-up is uniform
    int np;                 // producer warps
    long long* timer;       // optional (pk_debug_universe_timer): per block (loss_off, ns spent)
    int timer_base;
    float* scratch;         // global (L2-resident) gradient sums of multiply-occurring entity rows: [blocks][slots][ntE][d]
    size_t scratch_stride;  // floats per block
};

__host__ __device__ inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory carve-up, identical on host (sizing) and device (pointers)
struct K2Smem {
    size_t ent[2], rel[2], rel_state[2], relc[2], reln, relacc, scratch, map, batch[2], lossv, lcg, bar, total;
    int slots;      // entity scratch rows: a multiply-occurring entity needs >= 2 of the (2+k)B occurrences
    int nrelcap;    // distinct relations a batch can hold
    int per;        // samples per LCG stream slice
    int batch_ints; // ints in one batch buffer
    // fast: the FAST kernel instance (k = 1, L1, normalised, Adagrad, four-lane layout) is the one launched; with a
    // single relation table (TransE) its relation gradient sums need ONE int32 limb per element instead of three,
    // which is what lets relation-rich universes (FB15K shape: 400-600 relations) keep their relation side in
    // shared memory
    __host__ __device__ K2Smem(int model, int d, int k, int W, int mE, int mR, int mB, int stage, int fast) {
        const int ntE = model == TRANSD ? 2 : 1, ntR = model == TRANSE ? 1 : 2;
        const int nrc = model == TRANSH ? 2 : 1;   // cached relation operands: r^ (all), w^ (TransH)
        const long long occ = (long long)(2 + k) * mB;
        slots = (int)min_((long long)mE, occ / 2 + 1);
        nrelcap = (int)min_(mR, mB);
        per = (mB + W - 1) / W + 1;
        size_t o = 0;
        for (int i = 0; i < 2; ++i) { ent[i] = o; if (stage && i < ntE) o = up16(o + (size_t)mE * d * 4); }
        for (int i = 0; i < 2; ++i) { rel[i] = o; if (i < ntR) o = up16(o + (size_t)mR * d * 4); }
        for (int i = 0; i < 2; ++i) { rel_state[i] = o; if (i < ntR) o = up16(o + (size_t)mR * d * 4); }
        for (int i = 0; i < 2; ++i) { relc[i] = o; if (i < nrc) o = up16(o + (size_t)mR * d * 4); }
        reln = o;    o = up16(o + (size_t)2 * mR * 4);
        relacc = o;  o = up16(o + (size_t)mR * ntR * d * ((fast && ntR == 1) ? 4 : 12));  // fixed-point relation gradient sums
        scratch = o;   // (the entity scratch rows live in global memory, see K2Params::scratch)
        map = o;     o = up16(o + ((size_t)mE + mR) * 4);
        // one batch buffer: h | t | r | c[k] | code_h | code_t | code_c[k] | dup[slots] | dupslot[slots] | rel_ids | ndup | nrel
        batch_ints = (int)((3 + k) * (long long)mB + occ) + 2 * slots + nrelcap + 4;
        for (int i = 0; i < 2; ++i) { batch[i] = o; o = up16(o + (size_t)batch_ints * 4); }
        lossv = o;   o = up16(o + (size_t)2 * mB * 4);   // per-sample loss terms, double-buffered
        // s0[8] | Aadv[8] | Cadv[8] | A[per] | C[per]
        lcg = o;     o = up16(o + (size_t)(24 + 2 * per) * 8);
        bar = o;     o = up16(o + 16);   // mbarrier of the bulk (TMA) copy that stages the entity tables
        total = o;
    }
    __host__ __device__ static long long min_(long long a, long long b) { return a < b ? a : b; }
};

#ifdef PK_MODEL_TU

// x mod n for a divisor fixed per universe: multiply-high by m = floor((2^64-1)/n), then at most two
// conditional subtractions (q underestimates floor(x/n) by < 3).  ~12 instructions instead of the
// ~80 of a 64-bit division.
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

// coefficients of n LCG steps: x -> A x + C
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

// ---- bulk asynchronous copies (TMA, 1-D): the staged entity tables of a universe travel global <-> shared memory as a
//      few cp.async.bulk transfers issued by ONE thread (SASS: UBLKCP) instead of a float4 loop through the registers
//      of all 512; completion of the load is counted in bytes on an mbarrier, of the store by the bulk async-group
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
constexpr uint32_t kBulkChunk = 32768;   // bytes per transfer (a multiple of 16)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    for (uint32_t o = 0; o < bytes; o += kBulkChunk) {
        const uint32_t n = min(kBulkChunk, bytes - o);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(static_cast<unsigned char*>(smem_dst) + o)), "l"(static_cast<const unsigned char*>(gmem_src) + o),
                       "r"(n), "r"(smem_u32(bar)) : "memory");
    }
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    for (uint32_t o = 0; o < bytes; o += kBulkChunk) {
        const uint32_t n = min(kBulkChunk, bytes - o);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(static_cast<unsigned char*>(gmem_dst) + o), "r"(smem_u32(static_cast<const unsigned char*>(smem_src) + o)), "r"(n) : "memory");
    }
}

__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one batch as the producers hand it to the consumers
struct BatchView {
    int32_t* h;
    int32_t* t;
    int32_t* r;
    int32_t* c;            // [k][B]  corrupted entity | (side << 31); side 0: tail replaced, 1: head replaced
    int32_t* code_h;       // -1: the entity row occurs once in this batch (update in place); else its scratch row
    int32_t* code_t;
    int32_t* code_c;
    int32_t* dup;          // multiply-occurring entity rows of the batch ...
    int32_t* dupslot;      // ... and the scratch row each accumulates in (the owner's occurrence index)
    int32_t* rel_ids;      // distinct relations of the batch
    int32_t* ndup;
    int32_t* nrel;
    __device__ __forceinline__ BatchView(unsigned char* base, int B, int k, int slots, int nrelcap) {
        int32_t* p = reinterpret_cast<int32_t*>(base);
        h = p; p += B;
        t = p; p += B;
        r = p; p += B;
        c = p; p += (size_t)B * k;
        code_h = p; p += B;
        code_t = p; p += B;
        code_c = p; p += (size_t)B * k;
        dup = p; p += slots;
        dupslot = p; p += slots;
        rel_ids = p; p += nrelcap;
        ndup = p;
        nrel = p + 1;
    }
};

// an entity row as the backward pass addresses it: id, duplicate-slot code, prefetched Adagrad state
template <class L, int NTE>
struct K2Tgt {
    int32_t id, code;
    float st[NTE][L::NF];
};

// a row read straight from L2 (scalar accesses: only the short scratch rows are read this way)
template <class L>
__device__ __forceinline__ void ld_row_cg(const float* p, int d, int lane, float (&x)[L::NF]) {
#pragma unroll
    for (int i = 0; i < L::NF; ++i) {
        const int e = elem_of<L>(lane, i);
        x[i] = in_row<L>(e, d) ? __ldcg(p + e) : 0.f;
    }
}

// x <- optimizer(x, g): SGD  x -= lr g ;  Adagrad  s += g^2, x -= lr g / (sqrt(s) + 1e-10)
// (torch.optim.SGD / Adagrad as configured by reference Trainer.py:65-70,84-88; lr_decay = weight_decay = 0)
template <class L>
__device__ __forceinline__ void apply_update(float* x_row, float* s_row, float (&s)[L::NF], const float (&g)[L::NF], int d, int lane,
                                             int opt, float lr) {
    float x[L::NF];
    ld_row<L>(x_row, d, lane, x);
    if (opt == PK_ADAGRAD) {
#pragma unroll
        for (int i = 0; i < L::NF; ++i) {
            s[i] = fmaf(g[i], g[i], s[i]);
            // 1 / (sqrt(s) + 1e-10) from the two approximate units (rsqrt: 2^-22.4, rcp: 1 ulp): the step
            // lr*g/(...) is then within ~4 ulp OF THE STEP, i.e. below half an ulp of x for lr <= 0.1
            float r;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s[i], 1.17549435e-38f)));
            float inv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(s[i], r, 1e-10f)));
            x[i] = fmaf(-lr * g[i], inv, x[i]);
        }
        st_row<L>(s_row, d, lane, s);
    } else {
#pragma unroll
        for (int i = 0; i < L::NF; ++i) x[i] = fmaf(-lr, g[i], x[i]);
    }
    st_row<L>(x_row, d, lane, x);
}

struct RelCache;
template <class L>
__device__ __forceinline__ void fix_add_row(int32_t* row, int d, int lane, const float (&g)[L::NF]);
template <class L>
__device__ __forceinline__ void fix1_add_row(int32_t* row, int d, int lane, const float (&g)[L::NF]);

template <class L, int NTE_, int NTR_, int FAST_>
struct K2Ctx {
    static constexpr int NTE = NTE_, NTR = NTR_;
    static constexpr int ACC = (FAST_ != 0 && NTR_ == 1) ? 1 : 3;   // int32 limbs per relation gradient element
    static constexpr bool kMerge3 = FAST_ != 0 && NTE_ == 1 && NTR_ == 1;   // FAST TransE: see train_sample
    using Tgt = K2Tgt<L, NTE_>;
    float* ent[2];        // working entity tables: shared memory when staged, else global
    float* ent_state[2];  // global; nullptr for SGD
    float* scratch;       // [slots][NTE][d]
    const float* rel_c[2];   // cached relation operands: r^ ; w^ (TransH) or the raw r_p table (TransD)
    int32_t* rel_acc;        // fixed-point relation gradient sums [mR][NTR][3][d]
    const int32_t* bh;       // the current batch (BatchView of this step)
    const int32_t* bt;
    const int32_t* br;
    const int32_t* bc;
    const int32_t* code_h;
    const int32_t* code_t;
    const int32_t* code_c;
    int B;
    int d, opt;
    float lr;
    __device__ __forceinline__ void load_pos(int64_t b, bool act, Tgt& th, Tgt& tt, int32_t& r) const {
        th.id = act ? bh[b] : 0; th.code = act ? code_h[b] : 0;
        tt.id = act ? bt[b] : 0; tt.code = act ? code_t[b] : 0;
        r = act ? br[b] : 0;
    }
    __device__ __forceinline__ bool load_neg(int j, int64_t b, bool act, Tgt& tc) const {
        const int32_t cj = act ? bc[(uint32_t)(j * B) + (uint32_t)b] : 0;
        tc.id = cj & 0x7fffffff;
        tc.code = act ? code_c[(uint32_t)(j * B) + (uint32_t)b] : 0;
        return cj < 0;
    }
    __device__ __forceinline__ const float* rel_y(int r) const { return rel_c[0] + (uint32_t)r * (uint32_t)d; }
    __device__ __forceinline__ const float* rel_w(int r) const { return rel_c[1] + (uint32_t)r * (uint32_t)d; }
    __device__ __forceinline__ void rel_add(int tbl, int r, const float (&g)[L::NF], int lane) const {
        if (FAST_ && tbl == 0) fix1_add_row<L>(rel_acc + (uint32_t)(r * NTR * ACC * d), d, lane, g);
        else fix_add_row<L>(rel_acc + (uint32_t)((r * NTR + tbl) * 3 * d), d, lane, g);
    }
    __device__ __forceinline__ const float* ent_row(int tbl, const Tgt& tg) const { return ent[tbl] + (uint32_t)tg.id * (uint32_t)d; }
    // issue the loads of a singly-occurring row's optimizer state early; they complete behind the math
    __device__ __forceinline__ void prefetch(K2Tgt<L, NTE>& tg, int lane, bool pred) const {
        if (opt != PK_ADAGRAD) return;
#pragma unroll
        for (int t = 0; t < NTE; ++t) ld_row<L>(ent_state[t] + (uint32_t)tg.id * (uint32_t)d, d, lane, tg.st[t], pred && tg.code < 0);
    }
    // Three rows at once (FAST: Adagrad): the in-place arithmetic runs for all of them unconditionally
    // and only the stores are predicated, so nothing diverges and the 3 x NF rsqrt/rcp chains overlap.
    __device__ __forceinline__ void add_ent3(const K2Tgt<L, NTE>& a, const float (&ga)[L::NF], const K2Tgt<L, NTE>& b,
                                             const float (&gb)[L::NF], const K2Tgt<L, NTE>& c, const float (&gc)[L::NF], int lane,
                                             bool pred) const {
        const K2Tgt<L, NTE>* tg[3] = {&a, &b, &c};
        const float* gg[3] = {ga, gb, gc};
        float x[3][L::NF], s[3][L::NF];
#pragma unroll
        for (int q = 0; q < 3; ++q) ld_row<L>(ent[0] + (uint32_t)tg[q]->id * (uint32_t)d, d, lane, x[q]);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int i = 0; i < L::NF; ++i) {
                const float g = gg[q][i];
                s[q][i] = fmaf(g, g, tg[q]->st[0][i]);
                float r, inv;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(s[q][i], 1.17549435e-38f)));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(s[q][i], r, 1e-10f)));
                x[q][i] = fmaf(-lr * g, inv, x[q][i]);
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (pred && tg[q]->code < 0) {
                st_row<L>(ent_state[0] + (uint32_t)tg[q]->id * (uint32_t)d, d, lane, s[q]);
                st_row<L>(ent[0] + (uint32_t)tg[q]->id * (uint32_t)d, d, lane, x[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (pred && tg[q]->code >= 0) {
                float* p = scratch + (uint32_t)(tg[q]->code * d);
#pragma unroll
                for (int i = 0; i < L::NF; ++i) {
                    const int e = elem_of<L>(lane, i);
                    if (in_row<L>(e, d) && gg[q][i] != 0.f) atomicAdd(p + e, gg[q][i]);
                }
            }
        }
    }
    __device__ __forceinline__ void add_ent(int tbl, const K2Tgt<L, NTE>& tgc, const float (&g)[L::NF], int lane, bool pred) const {
        if (!pred) return;
        K2Tgt<L, NTE>& tg = const_cast<K2Tgt<L, NTE>&>(tgc);
        if (tg.code < 0) {
            apply_update<L>(ent[tbl] + (uint32_t)tg.id * (uint32_t)d, ent_state[tbl] ? ent_state[tbl] + (uint32_t)tg.id * (uint32_t)d : nullptr, tg.st[tbl], g, d,
                            lane, opt, lr);
        } else {
            // fire-and-forget RED.ADD.F32 at L2 (a shared-memory float add is a compare-and-swap loop whose
            // round trips, five per row, sat on the critical path of every sample)
            float* p = scratch + (uint32_t)((tg.code * NTE + tbl) * d);
#pragma unroll
            for (int i = 0; i < L::NF; ++i) {
                const int e = elem_of<L>(lane, i);
                if (in_row<L>(e, d) && g[i] != 0.f) atomicAdd(p + e, g[i]);
            }
        }
    }
};

// Relation-side operands, cached per step: r^ = normalize(r) (or r when !norm_flag), for TransH also
// w^ = normalize(norm_vector[r]); their norms; and the accumulator of the raw relation gradients
// (w.r.t. r^ and w^ / r_p).  A relation occurs in many samples of a batch (the universe's focus
// relation in half of them), and shared memory has native atomic adds only for 32-bit integers
// (fp32 and 64-bit adds are compare-and-swap loops that collapse under that contention).  The sums
// are therefore kept in 2^-44 fixed point split into three int32 limbs of 20 bits (room for 2047
// addends): exact for every term >= 2^-21, order-independent, hence deterministic, and each add is
// three fire-and-forget ATOMS.ADD.  The normalisation backward and the optimizer run once per
// relation after the consumer barrier.
struct RelCache {
    float* rel[2];     // working relation tables (shared memory)
    float* state[2];   // their Adagrad state (shared memory)
    float* c[2];       // cached operands
    float* n;          // [2][mR] clamped norms
    int32_t* acc;      // [mR][ntR][3 limbs][d]
    int mR;
};
constexpr float kFixScale = 17592186044416.f;          // 2^44
constexpr float kFixInv = 1.f / 17592186044416.f;
constexpr float kFixClamp = 262144.f;                   // |g| <= 2^18 keeps the top limb far from overflow

template <class L>
__device__ __forceinline__ void fix_add_row(int32_t* row, int d, int lane, const float (&g)[L::NF]) {
#pragma unroll
    for (int i = 0; i < L::NF; ++i) {
        const int e = elem_of<L>(lane, i);
        if (in_row<L>(e, d) && g[i] != 0.f) {
            const long long v = __float2ll_rn(fminf(fmaxf(g[i], -kFixClamp), kFixClamp) * kFixScale);
            atomicAdd(row + e, (int32_t)(v & 0xfffff));
            atomicAdd(row + d + e, (int32_t)((v >> 20) & 0xfffff));
            atomicAdd(row + 2 * d + e, (int32_t)(v >> 40));
        }
    }
}
// One-limb variant for the gradient w.r.t. the relation operand r^ under the L1 energy (FAST): every
// element of a sample's term is a sum of (1 + k) directions in [-1, 1] weighted by g_j <= 1/(B k), so
// |term| <= 2/B and |sum over the batch| <= 2: 2^-29 fixed point fits one int32 with a bit to spare,
// each add is ONE fire-and-forget ATOMS.ADD and the conversion is one F2I (exact to 2^-30 per term).
constexpr float kFix1Scale = 536870912.f;   // 2^29
constexpr float kFix1Inv = 1.f / 536870912.f;
template <class L>
__device__ __forceinline__ void fix1_add_row(int32_t* row, int d, int lane, const float (&g)[L::NF]) {
#pragma unroll
    for (int i = 0; i < L::NF; ++i) {
        const int e = elem_of<L>(lane, i);
        if (in_row<L>(e, d) && g[i] != 0.f) atomicAdd(row + e, __float2int_rn(fminf(fmaxf(g[i], -2.f), 2.f) * kFix1Scale));
    }
}
template <class L>
__device__ __forceinline__ bool fix1_take_row(int32_t* row, int d, int lane, float (&g)[L::NF]) {
    bool nz = false;
#pragma unroll
    for (int i = 0; i < L::NF; ++i) {
        const int e = elem_of<L>(lane, i);
        int32_t v = 0;
        if (in_row<L>(e, d)) {
            v = row[e];
            if (v != 0) row[e] = 0;
        }
        nz |= v != 0;
        g[i] = (float)v * kFix1Inv;
    }
    return nz;
}
// read a fixed-point row as fp32 and reset it; returns whether this lane saw a non-zero sum
template <class L>
__device__ __forceinline__ bool fix_take_row(int32_t* row, int d, int lane, float (&g)[L::NF]) {
    bool nz = false;
#pragma unroll
    for (int i = 0; i < L::NF; ++i) {
        const int e = elem_of<L>(lane, i);
        long long v = 0;
        if (in_row<L>(e, d)) {
            const int32_t a0 = row[e], a1 = row[d + e], a2 = row[2 * d + e];
            if ((a0 | a1 | a2) != 0) { row[e] = 0; row[d + e] = 0; row[2 * d + e] = 0; }
            v = ((long long)a2 << 40) + ((long long)a1 << 20) + (long long)a0;
        }
        nz |= v != 0;
        g[i] = (float)v * kFixInv;
    }
    return nz;
}

// FAST = the configuration every PuTrans* experiment script uses (k = 1, L1 energy, normalised operands,
// Adagrad): those launch parameters become compile-time constants.  With an exact layout (L::EX) the
// row length is a constant too.
template <int MODEL, class L, int NT, int STAGE, int FAST>
__global__ void __launch_bounds__(NT) k2_train_universes(const __grid_constant__ K2Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int ntE = MODEL == TRANSD ? 2 : 1, ntR = MODEL == TRANSE ? 1 : 2;
    const pk_universe_desc& U = P.desc[blockIdx.x];
    const K2Smem S(MODEL, L::EX ? L::D : P.d, FAST ? 1 : P.k, P.W, P.mE, P.mR, P.mB, STAGE, FAST);
    constexpr int ACC = K2Ctx<L, MODEL == TRANSD ? 2 : 1, MODEL == TRANSE ? 1 : 2, FAST>::ACC;
    const int tid = threadIdx.x;
    long long t_begin = 0;
    if (P.timer && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
    const int d = L::EX ? L::D : P.d, k = FAST ? 1 : P.k, B = U.batch_size, nE = U.n_ent, nR = U.n_rel, W = P.W;
    const int opt = FAST ? (int)PK_ADAGRAD : P.opt, p_norm = FAST ? 1 : P.p_norm, norm_flag = FAST ? 1 : P.norm_flag;
    const int NC = NT / 32 - P.np;          // consumer warps
    const int n_cons = NC * 32, n_prod = P.np * 32;
    const bool producer = tid >= n_cons;
    constexpr int GPW = 32 / L::G;          // lane groups per warp
    const int lane = tid % L::G, grp = tid / L::G;
    const unsigned gmask = group_mask<L::G>(tid);

    float* g_ent[2];   // this universe's tables in global memory
    float* g_rel[2];
    float* g_rel_state[2];
    K2Ctx<L, ntE, ntR, FAST> cx;
    RelCache rc;
    cx.d = d; cx.opt = opt; cx.lr = U.lr; cx.B = B;
    rc.mR = P.mR;
    for (int i = 0; i < 2; ++i) {
        g_ent[i] = (i < ntE) ? P.ent[i] + (size_t)U.ent_off * d : nullptr;
        g_rel[i] = (i < ntR) ? P.rel[i] + (size_t)U.rel_off * d : nullptr;
        g_rel_state[i] = (i < ntR && opt == PK_ADAGRAD) ? P.rel_state[i] + (size_t)U.rel_off * d : nullptr;
        cx.ent_state[i] = (i < ntE && opt == PK_ADAGRAD) ? P.ent_state[i] + (size_t)U.ent_off * d : nullptr;
        cx.ent[i] = STAGE ? reinterpret_cast<float*>(smem + S.ent[i]) : g_ent[i];
        rc.rel[i] = reinterpret_cast<float*>(smem + S.rel[i]);
        rc.state[i] = reinterpret_cast<float*>(smem + S.rel_state[i]);
        rc.c[i] = reinterpret_cast<float*>(smem + S.relc[i]);
    }
    rc.n = reinterpret_cast<float*>(smem + S.reln);
    rc.acc = reinterpret_cast<int32_t*>(smem + S.relacc);
    cx.rel_acc = rc.acc;
    cx.rel_c[0] = rc.c[0];
    cx.rel_c[1] = MODEL == TRANSD ? rc.rel[1] : rc.c[1];
    cx.scratch = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
    int32_t* map = reinterpret_cast<int32_t*>(smem + S.map);   // per table row: owner occurrence | (occurs again) << 31
    float* lossv = reinterpret_cast<float*>(smem + S.lossv);
    uint64_t* s0 = reinterpret_cast<uint64_t*>(smem + S.lcg);  // stream states at the start of the next batch
    uint64_t* Aadv = s0 + 8;
    uint64_t* Cadv = Aadv + 8;
    uint64_t* Aj = Cadv + 8;
    uint64_t* Cj = Aj + S.per;
    const int per = (B % W == 0) ? B / W : B / W + 1;   // Base.cpp:199-207
    uint64_t* stage_bar = reinterpret_cast<uint64_t*>(smem + S.bar);
    const uint32_t table_bytes = (uint32_t)nE * (uint32_t)d * 4u;
    // bulk copies move multiples of 16 bytes between 16-byte aligned addresses: rows of d % 4 == 0 floats at a row offset
    // of the packed table (whose base the allocator aligns to 256 bytes)
    const bool bulk = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(g_ent[0]) | (ntE > 1 ? reinterpret_cast<uintptr_t>(g_ent[1]) : 0)) & 15) == 0;
    const size_t batch_stride = S.batch[1] - S.batch[0];

    // refresh the cached operands of relation r from the working tables (group-masked shuffles:
    // groups of one warp work on different relations)
    auto recache = [&](int r) {
        float x[L::NF];
        ld_row<L>(rc.rel[0] + (uint32_t)r * (uint32_t)d, d, lane, x);
        float n = 1.f;
        bool fr = false;
        if (norm_flag) n = normalize_row<L>(x, fr, gmask);
        st_row<L>(rc.c[0] + (uint32_t)r * (uint32_t)d, d, lane, x);
        if (lane == 0) rc.n[r] = n;
        if constexpr (MODEL == TRANSH) {
            ld_row<L>(rc.rel[1] + (uint32_t)r * (uint32_t)d, d, lane, x);
            n = normalize_row<L>(x, fr, gmask);
            st_row<L>(rc.c[1] + (uint32_t)r * (uint32_t)d, d, lane, x);
            if (lane == 0) rc.n[rc.mR + r] = n;
        }
    };

    // ---- stage tables, clear scratch, build the LCG jump tables
    {
        if (STAGE && bulk) {
            if (tid == 0) {
                mbar_init(stage_bar, 1);
                mbar_expect_tx(stage_bar, (uint32_t)ntE * table_bytes);
                for (int t = 0; t < ntE; ++t) bulk_load(cx.ent[t], g_ent[t], table_bytes, stage_bar);
            }
        } else {
            for (int t = 0; t < ntE && STAGE; ++t)
                for (int i = tid; i < nE * d; i += NT) cx.ent[t][i] = g_ent[t][i];
        }
        for (int t = 0; t < ntR; ++t)
            for (int i = tid; i < nR * d; i += NT) {
                rc.rel[t][i] = g_rel[t][i];
                rc.state[t][i] = g_rel_state[t] ? g_rel_state[t][i] : 0.f;
            }
        for (int i = tid; i < (2 + k) * P.mB * ntE * d; i += NT) __stcg(cx.scratch + i, 0.f);
        for (int i = tid; i < nR * ntR * ACC * d; i += NT) rc.acc[i] = 0;
        if (tid < 8) s0[tid] = U.lcg[tid];
        if (tid < W) {
            int64_t lef, rig;
            slice_of(B, W, tid, lef, rig);
            lcg_affine((uint64_t)(rig - lef) * (uint64_t)(1 + 2 * k), Aadv[tid], Cadv[tid]);
        }
        for (int j = tid; j < per; j += NT) lcg_affine((uint64_t)j * (uint64_t)(1 + 2 * k), Aj[j], Cj[j]);
    }
    __syncthreads();
    if (STAGE && bulk) mbar_wait(stage_bar, 0);   // the staged tables have landed (and are visible to whoever waited)
    for (int r = grp; r < nR; r += NT / L::G) recache(r);

    const long long steps = (long long)U.epochs * U.nbatches;

    // ------------------------------------------------------------------------------------ producer
    // The reference sampling() call (Base.cpp:185-310, Corrupt.h:9-105), bit-exact, one thread per
    // positive, plus the occurrence analysis of the batch.
    const int32_t* by_head = P.by_head + (size_t)U.tri_off * 3;
    const int32_t* by_tail = P.by_tail ? P.by_tail + (size_t)U.tri_off * 3 : nullptr;
    const FastMod fm_tri = make_fastmod((uint64_t)U.n_tri), fm_coin = make_fastmod(1000ULL),
                  fm_ent = make_fastmod((uint64_t)(nE - 1));
    // Occurrence analysis without read-modify-write chains (the producers' dependent shared-memory
    // round trips are what bounds a step once the consumers are busy).  Every table-row occurrence of a
    // batch has an index occ = role * B + b (role 0: head, 1: tail, 2 + j: negative j).
    //   pass 1  each occurrence stores occ into its row's word of `map` (plain stores: whichever lands
    //           last is the row's OWNER); relations likewise (owner = a sample index);
    //   pass 2  an occurrence that is not the owner sets the word's top bit (all of them write the same value);
    //   pass 3  top bit set: the row occurs several times and accumulates in scratch row `owner`; else -1.
    //           Owners of such rows and owners of relations append themselves to the step's work lists.
    // Pass 1 of the next batch overwrites every word before it is read again: nothing to clear.
    // The positives of the batch AFTER the one being produced are requested one call early, so the
    // triple-index loads (L2) are off the chain as well.
    // Fast path (k = 1, unfiltered; at most two samples per producer thread): a producer thread owns
    // the same samples b = ptid, ptid + n_prod of every batch, so everything that does not change between
    // batches stays in registers.  Its position in its LCG stream advances by the same affine map
    // every batch:  x' = Aadv x + K  with  K = Aj Cadv + Cj (1 - Aadv)  (x = Aj s0 + Cj), so no shared
    // jump tables are read per step; all three draws of the batch AFTER the one being handed over are
    // evaluated, and its triple-index loads issued, one call early.
    struct PState {
        uint64_t x, Aad, K;                 // stream position before the sample's first draw; its per-batch advance
        int32_t h, r, t, coin, tmp;         // the next batch's positive and raw draws
        bool on;
    };
    PState ps[2];
    const bool pfast = producer && k == 1 && !P.filter;
    auto draw_next = [&](PState& q) {
        uint64_t sx = q.x;
        const int64_t i = (int64_t)fastmod(lcg_next(sx), fm_tri);
        q.h = by_head[i * 3 + 0]; q.r = by_head[i * 3 + 1]; q.t = by_head[i * 3 + 2];
        q.coin = (int32_t)fastmod(lcg_next(sx), fm_coin);
        q.tmp = (int32_t)fastmod(lcg_next(sx), fm_ent);
        q.x = q.Aad * q.x + q.K;
    };
    if (producer) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int b = q * n_prod + (tid - n_cons);
            ps[q].on = pfast && b < B;
            if (ps[q].on) {
                const int id = b / per, j = b - id * per;
                ps[q].Aad = Aadv[id];
                ps[q].x = Aj[j] * s0[id] + Cj[j];
                ps[q].K = Aj[j] * Cadv[id] + Cj[j] * (1ULL - ps[q].Aad);
                draw_next(ps[q]);
            }
        }
    }
    // one sample drawn from the shared jump tables (any k, filtered corruption, samples beyond the fast path)
    auto sample_generic = [&](int b, const BatchView& bv) {
        const int id = b / per, j = b - id * per;
        uint64_t sx = Aj[j] * s0[id] + Cj[j];
        const int64_t i = (int64_t)fastmod(lcg_next(sx), fm_tri);
        const int32_t h = by_head[i * 3 + 0], r = by_head[i * 3 + 1], t = by_head[i * 3 + 2];
        bv.h[b] = h; bv.t[b] = t; bv.r[b] = r;
        map[h] = b; map[t] = B + b; map[nE + r] = b;
        float prob = 500.f;
        if (P.bern) {
            const float rm = P.right_mean[U.rel_off + r], lm = P.left_mean[U.rel_off + r];
            prob = __fdiv_rn(__fmul_rn(1000.f, rm), __fadd_rn(rm, lm));  // Base.cpp:220-221
        }
        for (int n = 0; n < k; ++n) {
            const uint64_t coin = fastmod(lcg_next(sx), fm_coin);
            const uint64_t x = lcg_next(sx);
            int32_t c, side;
            if ((float)coin < prob) {   // keep head, replace tail (corrupt_head, Corrupt.h:9-57)
                if (!P.filter) { const int64_t tmp = (int64_t)fastmod(x, fm_ent); c = (int32_t)(tmp < h ? tmp : tmp + 1); }
                else c = corrupt_entity(x, by_head, U.n_tri, nE, h, r, 0, 2, true);
                side = 0;
            } else {                    // keep tail, replace head (corrupt_tail, Corrupt.h:59-105)
                if (!P.filter) { const int64_t tmp = (int64_t)fastmod(x, fm_ent); c = (int32_t)(tmp < t ? tmp : tmp + 1); }
                else c = corrupt_entity(x, by_tail, U.n_tri, nE, t, r, 2, 0, true);
                side = 1;
            }
            bv.c[(uint32_t)(n * B) + (uint32_t)b] = (int32_t)((uint32_t)c | ((uint32_t)side << 31));
            map[c] = (2 + n) * B + b;
        }
    };
    auto produce = [&](int buf) {
        const int ptid = tid - n_cons;
        BatchView bv(smem + S.batch[0] + (size_t)buf * batch_stride, B, k, S.slots, S.nrelcap);
        if (ptid == 0) { *bv.ndup = 0; *bv.nrel = 0; }
        // ---- pass 1: the reference sampling() call (Base.cpp:185-310, Corrupt.h:9-105), bit-exact
        int32_t eh[2], et[2], ec[2];   // this thread's fast-path samples, kept for passes 2 and 3
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!ps[q].on) continue;
            const int b = q * n_prod + ptid;
            const int32_t h = ps[q].h, r = ps[q].r, t = ps[q].t, coin = ps[q].coin, tmp = ps[q].tmp;
            draw_next(ps[q]);   // the batch after this one: its loads have a whole step to arrive
            float prob = 500.f;
            if (P.bern) {
                const float rm = P.right_mean[U.rel_off + r], lm = P.left_mean[U.rel_off + r];
                prob = __fdiv_rn(__fmul_rn(1000.f, rm), __fadd_rn(rm, lm));  // Base.cpp:220-221
            }
            const bool keep_head = (float)coin < prob;
            const int32_t fix = keep_head ? h : t;
            const int32_t c = tmp < fix ? tmp : tmp + 1;
            bv.h[b] = h; bv.t[b] = t; bv.r[b] = r;
            bv.c[b] = (int32_t)((uint32_t)c | (keep_head ? 0u : 0x80000000u));
            map[h] = b; map[t] = B + b; map[nE + r] = b; map[c] = 2 * B + b;
            eh[q] = h; et[q] = t; ec[q] = c;
        }
        for (int b = (pfast ? 2 * n_prod : 0) + ptid; b < B; b += n_prod) sample_generic(b, bv);
        named_barrier(2, n_prod);
        // ---- pass 2: losers flag their row
        auto flag = [&](int32_t e, int32_t occ) {
            const int32_t m = map[e];
            if ((m & 0x7fffffff) != occ) map[e] = m | (int32_t)0x80000000;
        };
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!ps[q].on) continue;
            const int b = q * n_prod + ptid;
            flag(eh[q], b); flag(et[q], B + b); flag(ec[q], 2 * B + b);
        }
        for (int b = (pfast ? 2 * n_prod : 0) + ptid; b < B; b += n_prod) {
            flag(bv.h[b], b);
            flag(bv.t[b], B + b);
            for (int n = 0; n < k; ++n) flag(bv.c[(uint32_t)(n * B) + (uint32_t)b] & 0x7fffffff, (2 + n) * B + b);
        }
        named_barrier(2, n_prod);
        // ---- pass 3: codes and work lists
        auto code_of = [&](int32_t e, int32_t occ) -> int32_t {
            const int32_t m = map[e];
            if (m >= 0) return -1;
            const int32_t own = m & 0x7fffffff;
            if (own == occ) {
                const int32_t i = atomicAdd(bv.ndup, 1);
                bv.dup[i] = e;
                bv.dupslot[i] = own;
            }
            return own;
        };
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (!ps[q].on) continue;
            const int b = q * n_prod + ptid;
            bv.code_h[b] = code_of(eh[q], b);
            bv.code_t[b] = code_of(et[q], B + b);
            bv.code_c[b] = code_of(ec[q], 2 * B + b);
            const int32_t r = bv.r[b];
            if (map[nE + r] == b) bv.rel_ids[atomicAdd(bv.nrel, 1)] = r;   // distinct relations of the batch
        }
        for (int b = (pfast ? 2 * n_prod : 0) + ptid; b < B; b += n_prod) {
            bv.code_h[b] = code_of(bv.h[b], b);
            bv.code_t[b] = code_of(bv.t[b], B + b);
            for (int n = 0; n < k; ++n)
                bv.code_c[(uint32_t)(n * B) + (uint32_t)b] = code_of(bv.c[(uint32_t)(n * B) + (uint32_t)b] & 0x7fffffff, (2 + n) * B + b);
            const int32_t r = bv.r[b];
            if (map[nE + r] == b) bv.rel_ids[atomicAdd(bv.nrel, 1)] = r;
        }
        if (ptid < W) s0[ptid] = Aadv[ptid] * s0[ptid] + Cadv[ptid];
    };

    // deterministic loss reduction by the first producer warp: mean + margin (MarginLoss.py:28)
    auto reduce_loss = [&](long long st) {
        const int ptid = tid - n_cons;
        if (ptid < 32 && U.loss_off >= 0) {
            const float* lv = lossv + (int)(st & 1) * P.mB;
            float acc = 0.f;
            for (int b = ptid; b < B; b += 32) acc += lv[b];
            acc = gsum<32>(acc);
            if (ptid == 0) P.loss[U.loss_off + st] = acc / (float)((long long)B * k) + U.margin;
        }
    };
    if (producer && steps > 0) produce(0);
    __syncthreads();   // also orders the initial recache() before the first step

    // ------------------------------------------------------------------------------------ consumers
    const int NG = NC * GPW;
    Hyper hp;
    hp.d = d; hp.k = k; hp.p_norm = p_norm; hp.norm_flag = norm_flag;
    hp.margin = U.margin;
    hp.inv_bk = 1.f / (float)((long long)B * k);

    // The two roles run SEPARATE step loops (their register allocations stay independent) and meet once
    // per step at barrier 0 (bar.sync counts arriving threads, whichever instruction they arrive from).
    if (producer) {
        for (long long step = 0; step < steps; ++step) {
            const int buf = (int)(step & 1);
            if (step + 1 < steps) produce(buf ^ 1);
            // the loss of the step that just finished is reduced here, off the consumers' critical path
            if (step > 0) reduce_loss(step - 1);
            named_barrier(0, NT);
        }
    } else {
      for (long long step = 0; step < steps; ++step) {
        const int buf = (int)(step & 1);
        {
            const BatchView bv(smem + S.batch[0] + (size_t)buf * batch_stride, B, k, S.slots, S.nrelcap);
            cx.bh = bv.h; cx.bt = bv.t; cx.br = bv.r; cx.bc = bv.c;
            cx.code_h = bv.code_h; cx.code_t = bv.code_t; cx.code_c = bv.code_c;
            // ---- phase A: forward + analytic backward; singly-occurring entity rows updated in place
            for (int base = 0; base < B; base += NG) {
                const int b = base + grp;
                const bool act = b < B;
                const float l = train_sample<MODEL, L>(cx, hp, lane, b, act);
                if (act && lane == 0) lossv[buf * P.mB + b] = l;
            }
            named_barrier(1, n_cons);
            // ---- phase B: one item per distinct relation (take the fixed-point gradient sums, normalisation
            //      backward, update, refresh the cache) and per multiply-occurring entity row
            const int nrel = *bv.nrel, nd = *bv.ndup;
            for (int it = grp; it < nrel + nd; it += NG) {
                if (it < nrel) {
                    const int r = bv.rel_ids[it];
                    int32_t* pr = rc.acc + (uint32_t)(r * ntR * ACC * d);
                    float g0[L::NF], g1[ntR == 2 ? L::NF : 1];
                    bool nz = FAST ? fix1_take_row<L>(pr, d, lane, g0) : fix_take_row<L>(pr, d, lane, g0);
                    if constexpr (ntR == 2) nz |= fix_take_row<L>(pr + 3 * d, d, lane, g1);
                    // an all-zero sum updates nothing (SGD and Adagrad leave zero-gradient rows unchanged)
                    if (__ballot_sync(gmask, nz) != 0u) {
                        float y[L::NF], st[L::NF];
                        if (norm_flag) {
                            ld_row<L>(rc.c[0] + (uint32_t)r * (uint32_t)d, d, lane, y);
                            const float n = rc.n[r];
                            normalize_bwd<L>(y, n, n > kNormEps, g0, gmask);
                        }
                        ld_row<L>(rc.state[0] + (uint32_t)r * (uint32_t)d, d, lane, st);
                        apply_update<L>(rc.rel[0] + (uint32_t)r * (uint32_t)d, rc.state[0] + (uint32_t)r * (uint32_t)d, st, g0, d, lane, opt, U.lr);
                        if constexpr (MODEL == TRANSH) {
                            ld_row<L>(rc.c[1] + (uint32_t)r * (uint32_t)d, d, lane, y);
                            const float n = rc.n[rc.mR + r];
                            normalize_bwd<L>(y, n, n > kNormEps, g1, gmask);
                        }
                        if constexpr (ntR == 2) {
                            ld_row<L>(rc.state[1] + (uint32_t)r * (uint32_t)d, d, lane, st);
                            apply_update<L>(rc.rel[1] + (uint32_t)r * (uint32_t)d, rc.state[1] + (uint32_t)r * (uint32_t)d, st, g1, d, lane, opt, U.lr);
                        }
                        recache(r);
                    }
                } else {
                    const int id = bv.dup[it - nrel];
                    const int s = bv.dupslot[it - nrel];
#pragma unroll
                    for (int t = 0; t < ntE; ++t) {
                        float* grow = cx.scratch + (uint32_t)((s * ntE + t) * d);
                        float g[L::NF], st[L::NF];
                        ld_row_cg<L>(grow, d, lane, g);   // summed by L2 reductions: must not come from L1
                        float* xrow = cx.ent[t] + (uint32_t)id * (uint32_t)d;
                        float* srow = opt == PK_ADAGRAD ? cx.ent_state[t] + (uint32_t)id * (uint32_t)d : nullptr;
                        ld_row<L>(srow, d, lane, st, opt == PK_ADAGRAD);
                        apply_update<L>(xrow, srow, st, g, d, lane, opt, U.lr);
#pragma unroll
                        for (int i = 0; i < L::NF; ++i) {
                            const int e = elem_of<L>(lane, i);
                            if (in_row<L>(e, d) && g[i] != 0.f) __stcg(grow + e, 0.f);
                        }
                    }
                }
            }
        }
        named_barrier(0, NT);
      }
    }

    if (producer && steps > 0) reduce_loss(steps - 1);

    // ---- write staged tables back
    {
        if (STAGE && bulk) {
            // shared-memory writes of the generic proxy (the updates) must be ordered before the async proxy reads them
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                for (int t = 0; t < ntE; ++t) bulk_store(g_ent[t], cx.ent[t], table_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            for (int t = 0; t < ntE && STAGE; ++t)
                for (int i = tid; i < nE * d; i += NT) g_ent[t][i] = cx.ent[t][i];
        }
        for (int t = 0; t < ntR; ++t)
            for (int i = tid; i < nR * d; i += NT) {
                g_rel[t][i] = rc.rel[t][i];
                if (g_rel_state[t]) g_rel_state[t][i] = rc.state[t][i];
            }
    }
    if (STAGE && bulk && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the block must not retire before its stores
    if (P.timer && tid == 0) {
        long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        P.timer[2 * (size_t)(P.timer_base + blockIdx.x)] = U.loss_off;
        P.timer[2 * (size_t)(P.timer_base + blockIdx.x) + 1] = t_end - t_begin;
    }
}

#endif  // PK_MODEL_TU

// ---- dispatch over (model, layout)
struct LaySel { int V, G, CPL, NQ; };

inline LaySel pick_layout(int model, int d) {
    // PK_K2_LAYOUT=V,G,CPL forces one of the instantiated layouts (experiments)
    if (const char* e = getenv("PK_K2_LAYOUT")) {
        LaySel f{0, 0, 0, 0};
        if (sscanf(e, "%d,%d,%d,%d", &f.V, &f.G, &f.CPL, &f.NQ) >= 3 && f.G * (4 * f.NQ + f.V * f.CPL) >= d &&
            (f.G != 4 || f.G * (4 * f.NQ + f.V * f.CPL) == d))
            return f;
    }
    // Short rows (the PuTrans* scripts all use d = 20): four lanes per row with scalar chunks keep every
    // lane busy (d = 20: 4 lanes x 5 floats, against 5 of 8 lanes with 128-bit chunks) and put EIGHT
    // samples in a warp, so a batch of up to 104 positives is one pass of the 13 consumer warps.
    // d = 20 exactly: one 128-bit access plus one 32-bit access per lane instead of five 32-bit ones (fewer
    // memory instructions in flight, whole sectors of the L2-resident Adagrad state per request).
    if (d == 20) return LaySel{1, 4, 1, 1};
    if (d <= 20 && d % 4 == 0) return LaySel{1, 4, d / 4, 0};
    const int V = d % 4 == 0 ? 4 : (d % 2 == 0 ? 2 : 1);
    const int chunks = d / V;
    const int nf_cap = model == TRANSD ? 4 : 8;  // registers per row per lane
    for (int G : {8, 32}) {
        int cpl = (chunks + G - 1) / G;
        int c2 = 1;
        while (c2 < cpl) c2 *= 2;
        if (c2 * V <= nf_cap || G == 32) return LaySel{V, G, std::max(c2, 1), 0};
    }
    return LaySel{V, 32, 1, 0};
}

// whether dispatch_layout launches the FAST instance for this configuration (four-lane layouts only)
inline int fast_instance(const pk_model_cfg* cfg) {
    const LaySel l = pick_layout(cfg->model, cfg->dim);
    return (cfg->neg_ent == 1 && cfg->p_norm == 1 && cfg->norm_flag == 1 && cfg->opt == PK_ADAGRAD && l.V == 1 && l.G == 4) ? 1 : 0;
}

// threads per block: as many consumer warps as the register budget allows
constexpr int k2_threads(int model, int nf) { return nf > 5 ? 256 : (model == 2 ? 384 : 512); }

#ifdef PK_MODEL_TU
template <int MODEL, class L, int FAST>
int launch_k2(const K2Params& P, int stage, int n, size_t smem, cudaStream_t st) {
    constexpr int NT = k2_threads(MODEL, L::NF);
    if (stage) {
        auto kern = k2_train_universes<MODEL, L, NT, 1, FAST>;
        PK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n, NT, smem, st>>>(P);
    } else {
        auto kern = k2_train_universes<MODEL, L, NT, 0, FAST>;
        PK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<n, NT, smem, st>>>(P);
    }
    PK_LAUNCHED("k2_train_universes");
    return PK_OK;
}

template <int MODEL>
int dispatch_layout(const LaySel& l, const K2Params& P, int stage, int n, size_t smem, cudaStream_t st) {
    // four-lane layouts are only ever picked when they fit the row exactly (pick_layout)
    const bool fast = P.k == 1 && P.p_norm == 1 && P.norm_flag == 1 && P.opt == PK_ADAGRAD;
#define PK_CASE4(c, q)                                                                          \
    if (l.V == 1 && l.G == 4 && l.CPL == c && l.NQ == q && P.d == 4 * (4 * q + c))              \
        return fast ? launch_k2<MODEL, Lay<1, 4, c, 1, q>, 1>(P, stage, n, smem, st)            \
                    : launch_k2<MODEL, Lay<1, 4, c, 1, q>, 0>(P, stage, n, smem, st);
#ifdef PK_K2_DEV
    PK_CASE4(5, 0) PK_CASE4(1, 1)
#undef PK_CASE4
    return pk::fail(PK_ERR_UNSUPPORTED, "PK_K2_DEV build: only d = 20");
#else
    PK_CASE4(1, 0) PK_CASE4(2, 0) PK_CASE4(3, 0) PK_CASE4(4, 0) PK_CASE4(5, 0) PK_CASE4(1, 1)
#undef PK_CASE4
#define PK_CASE(v, g, c) if (l.V == v && l.G == g && l.CPL == c && l.NQ == 0) return launch_k2<MODEL, Lay<v, g, c>, 0>(P, stage, n, smem, st);
    PK_CASE(4, 8, 1) PK_CASE(4, 8, 2) PK_CASE(4, 32, 1) PK_CASE(4, 32, 2)
    PK_CASE(2, 8, 1) PK_CASE(2, 8, 2) PK_CASE(2, 8, 4) PK_CASE(2, 32, 1) PK_CASE(2, 32, 2) PK_CASE(2, 32, 4)
    PK_CASE(1, 8, 1) PK_CASE(1, 8, 2) PK_CASE(1, 8, 4) PK_CASE(1, 8, 8) PK_CASE(1, 32, 1) PK_CASE(1, 32, 2) PK_CASE(1, 32, 4) PK_CASE(1, 32, 8)
#undef PK_CASE
    return pk::fail(PK_ERR_UNSUPPORTED, "embedding dimension not supported by the universe kernel (d <= 256)");
#endif
}

// one translation unit per model keeps the build parallel: -DPK_MODEL_TU=0|1|2
#define PK_CAT2(a, b) a##b
#define PK_CAT(a, b) PK_CAT2(a, b)
int PK_CAT(launch_model, PK_MODEL_TU)(const LaySel& l, const K2Params& P, int stage, int n, size_t smem, cudaStream_t st) {
    return dispatch_layout<PK_MODEL_TU>(l, P, stage, n, smem, st);
}
}  // namespace pkk2
#else
int launch_model0(const LaySel& l, const K2Params& P, int stage, int n, size_t smem, cudaStream_t st);
int launch_model1(const LaySel& l, const K2Params& P, int stage, int n, size_t smem, cudaStream_t st);
int launch_model2(const LaySel& l, const K2Params& P, int stage, int n, size_t smem, cudaStream_t st);

// Device copies of the universe descriptors.  Launches on different streams (pieces of one chunk run
// concurrently) must not share a buffer, and stream-ordered cudaMallocAsync goes back to the driver on
// every call with the default pool settings, so a small per-thread pool hands out buffers whose last
// launch has completed (event query) and grows otherwise.
struct DescSlot {
    pk_universe_desc* d = nullptr;
    size_t cap = 0;
    // pinned staging for the descriptors: a copy from pageable memory beyond 64 KB (a few hundred universes) makes the
    // launching thread wait for the stream, i.e. for the launches ahead of this one
    pk_universe_desc* h = nullptr;
    size_t hcap = 0;
    pk_universe_desc* stage(const pk_universe_desc* src, size_t bytes) {
        if (hcap < bytes) {
            if (h) cudaFreeHost(h);
            h = nullptr;
            hcap = 0;
            const size_t want = bytes + bytes / 2 + 4096;
            if (cudaHostAlloc(&h, want, cudaHostAllocDefault) != cudaSuccess) return nullptr;
            hcap = want;
        }
        std::memcpy(h, src, bytes);
        return h;
    }
    cudaEvent_t done = nullptr;
    bool busy = false;
    uint64_t seq = 0;      // launch order (the oldest busy buffer is the one to wait for)
};
thread_local std::vector<DescSlot> g_desc_pool;
thread_local uint64_t g_desc_seq = 0;
constexpr size_t kMaxDescSlots = 16;
constexpr size_t kDescArea = 256;   // one descriptor (128 B), padded

// The unstaged class (universes whose tables do not fit in shared memory) runs beside the staged one
// on a side stream: fork from / join into the caller's stream with events.
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
// one per caller stream: launches of different caller streams (chunks in flight together) must not meet on one side stream
thread_local std::map<cudaStream_t, SideStream> g_sides;
long long* g_timer = nullptr;   // pk_debug_universe_timer

DescSlot* acquire_desc(size_t bytes) {
    // a free buffer that is large enough (the smallest such), else the largest free one regrown with
    // headroom (chunks differ in size from call to call; a cudaMalloc per call costs milliseconds), else a new one
    DescSlot* fit = nullptr;
    DescSlot* grow = nullptr;
    for (auto& s : g_desc_pool) {
        if (s.busy && cudaEventQuery(s.done) != cudaSuccess) continue;
        s.busy = false;
        if (s.cap >= bytes) { if (!fit || s.cap < fit->cap) fit = &s; }
        else if (!grow || s.cap > grow->cap) grow = &s;
    }
    if (fit) return fit;
    // A caller that queues launches faster than they finish (twenty steps issued back to back) would grow the pool by one
    // buffer per launch, and every allocation (cudaMalloc + cudaHostAlloc) stalls behind the launches in flight: beyond
    // kMaxDescSlots buffers the launching thread waits for the oldest launch instead.
    if (!grow && g_desc_pool.size() >= kMaxDescSlots) {
        DescSlot* oldest = &g_desc_pool[0];
        for (auto& s : g_desc_pool)
            if (s.seq < oldest->seq) oldest = &s;
        cudaEventSynchronize(oldest->done);
        oldest->busy = false;
        if (oldest->cap >= bytes) return oldest;
        grow = oldest;
    }
    const size_t want = bytes + bytes / 2 + (1u << 20);
    if (grow) {
        if (grow->d) cudaFree(grow->d);
        grow->d = nullptr;
        grow->cap = 0;
        if (cudaMalloc(&grow->d, want) != cudaSuccess) return nullptr;
        grow->cap = want;
        return grow;
    }
    DescSlot s;
    if (cudaMalloc(&s.d, want) != cudaSuccess) return nullptr;
    s.cap = want;
    if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) { cudaFree(s.d); return nullptr; }
    g_desc_pool.push_back(s);
    return &g_desc_pool.back();
}

}  // namespace pkk2

using namespace pkk2;

// Debug/profiling aid: when set, every universe block of later pk_train_universes calls writes
// (loss_off, nanoseconds it ran) to d_buf[2*i], d_buf[2*i+1] (i = launch order; n_universes pairs).
extern "C" int pk_debug_universe_timer(long long* d_buf) {
    g_timer = d_buf;
    return PK_OK;
}

extern "C" int pk_universe_kernel_class(const pk_model_cfg* cfg, int64_t n_ent, int64_t n_rel, int64_t batch_size) {
    if (!cfg || cfg->model < 0 || cfg->model > 2 || cfg->dim <= 0 || cfg->neg_ent < 1 || cfg->work_threads < 1 || n_ent < 2 || n_rel < 1 ||
        batch_size < 1)
        return pk::fail(PK_ERR_ARG, "pk_universe_kernel_class: bad argument");
    int dev = 0, max_smem = 0;
    PK_CUDA(cudaGetDevice(&dev));
    PK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int k = cfg->neg_ent;
    if ((long long)(3 + k) * batch_size >= 32768 || (long long)batch_size * k > 2047 || cfg->dim > 256 || cfg->work_threads > 8) return 2;
    const int fast = fast_instance(cfg);
    K2Smem a(cfg->model, cfg->dim, k, cfg->work_threads, (int)n_ent, (int)n_rel, (int)batch_size, 1, fast);
    if (a.total <= (size_t)max_smem) return 0;
    K2Smem b(cfg->model, cfg->dim, k, cfg->work_threads, (int)n_ent, (int)n_rel, (int)batch_size, 0, fast);
    return b.total <= (size_t)max_smem ? 1 : 2;
}

extern "C" int pk_train_universes(const pk_model_cfg* cfg, const pk_tables* packed, const int32_t* d_by_head,
                                  const int32_t* d_by_tail, const float* d_left_mean, const float* d_right_mean,
                                  const pk_universe_desc* h_desc, int n, float* d_loss, void* stream) {
    pk::launch_counter() = 0;
    if (!cfg || !packed || !h_desc || n < 0) return pk::fail(PK_ERR_ARG, "pk_train_universes: null argument");
    if (n == 0) return PK_OK;
    if (cfg->model < 0 || cfg->model > 2) return pk::fail(PK_ERR_ARG, "pk_train_universes: unknown model");
    if (cfg->dim <= 0 || cfg->dim > 256) return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: dim must be in [1,256]");
    if (cfg->p_norm != 1 && cfg->p_norm != 2) return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: p_norm must be 1 or 2");
    if (cfg->neg_ent < 1) return pk::fail(PK_ERR_ARG, "pk_train_universes: neg_ent must be >= 1");
    if (cfg->work_threads < 1 || cfg->work_threads > 8) return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: work_threads must be in [1,8]");
    if (cfg->filter && !d_by_tail) return pk::fail(PK_ERR_ARG, "pk_train_universes: filter needs the (t,r,h) index");
    if (cfg->bern && (!d_left_mean || !d_right_mean)) return pk::fail(PK_ERR_ARG, "pk_train_universes: bern needs the relation means");
    if (!d_by_head) return pk::fail(PK_ERR_ARG, "pk_train_universes: null triple index");
    cudaStream_t st = (cudaStream_t)stream;
    const int d = cfg->dim, k = cfg->neg_ent, W = cfg->work_threads;

    int dev = 0, max_smem = 0;
    PK_CUDA(cudaGetDevice(&dev));
    PK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    const int fast = fast_instance(cfg);
    // split into a staged launch (entity tables fit in shared memory) and an unstaged one
    std::vector<pk_universe_desc> cls[2];
    int mE[2] = {1, 1}, mR[2] = {1, 1}, mB[2] = {1, 1};
    for (int i = 0; i < n; ++i) {
        const pk_universe_desc& u = h_desc[i];
        if (u.n_ent < 2 || u.n_rel < 1 || u.n_tri < 1 || u.batch_size < 1 || u.nbatches < 0 || u.epochs < 0)
            return pk::fail(PK_ERR_ARG, "pk_train_universes: degenerate universe descriptor");
        if ((long long)(3 + k) * u.batch_size >= 32768 || (long long)u.batch_size * k > 2047)
            return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: batch too large for the universe kernel (B*k <= 2047); use pk_train_steps");
        K2Smem own(cfg->model, d, k, W, u.n_ent, u.n_rel, u.batch_size, 1, fast);
        const int c = own.total <= (size_t)max_smem ? 0 : 1;
        cls[c].push_back(u);
        mE[c] = std::max(mE[c], u.n_ent);
        mR[c] = std::max(mR[c], u.n_rel);
        mB[c] = std::max(mB[c], u.batch_size);
    }
    // the uniform carve-up uses the class maxima; if that overflows, demote the largest universes
    for (;;) {
        if (cls[0].empty()) break;
        K2Smem s(cfg->model, d, k, W, mE[0], mR[0], mB[0], 1, fast);
        if (s.total <= (size_t)max_smem) break;
        size_t worst = 0;
        for (size_t i = 1; i < cls[0].size(); ++i)
            if ((long long)cls[0][i].n_ent * 4 + cls[0][i].batch_size > (long long)cls[0][worst].n_ent * 4 + cls[0][worst].batch_size) worst = i;
        cls[1].push_back(cls[0][worst]);
        mE[1] = std::max(mE[1], cls[0][worst].n_ent);
        mR[1] = std::max(mR[1], cls[0][worst].n_rel);
        mB[1] = std::max(mB[1], cls[0][worst].batch_size);
        cls[0].erase(cls[0].begin() + (long)worst);
        mE[0] = mR[0] = mB[0] = 1;
        for (const auto& u : cls[0]) {
            mE[0] = std::max(mE[0], u.n_ent);
            mR[0] = std::max(mR[0], u.n_rel);
            mB[0] = std::max(mB[0], u.batch_size);
        }
    }

    const LaySel lay = pick_layout(cfg->model, d);
    const int threads = k2_threads(cfg->model, lay.V * lay.CPL + 4 * lay.NQ);
    int np = threads >= 512 ? 3 : 2;
    if (const char* e = getenv("PK_K2_PRODUCERS")) np = std::max(1, std::min(threads / 32 - 1, atoi(e)));
    const bool both = !cls[0].empty() && !cls[1].empty();
    cudaStream_t caller = st;
    SideStream& g_side = g_sides[caller];
    if (both) {
        if (!g_side.st) {
            PK_CUDA(cudaStreamCreateWithFlags(&g_side.st, cudaStreamNonBlocking));
            PK_CUDA(cudaEventCreateWithFlags(&g_side.fork, cudaEventDisableTiming));
            PK_CUDA(cudaEventCreateWithFlags(&g_side.join, cudaEventDisableTiming));
        }
        PK_CUDA(cudaEventRecord(g_side.fork, caller));
        PK_CUDA(cudaStreamWaitEvent(g_side.st, g_side.fork, 0));
    }
    for (int c = 1; c >= 0; --c) {   // the unstaged class first: its universes are the slowest
        if (cls[c].empty()) continue;
        st = (both && c == 1) ? g_side.st : caller;
        K2Smem s(cfg->model, d, k, W, mE[c], mR[c], mB[c], c == 0, fast);
        if (s.total > (size_t)max_smem) {
            // the class maxima do not fit together (c == 1 only): one launch per universe, each with its
            // own carve-up; a universe that does not fit alone belongs to pk_train_steps
            // (pk_universe_kernel_class tells the caller beforehand)
            for (const auto& u : cls[c]) {
                K2Smem own(cfg->model, d, k, W, u.n_ent, u.n_rel, u.batch_size, 0, fast);
                if (own.total > (size_t)max_smem)
                    return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: a universe's relation tables and batch scratch exceed shared memory; train it with pk_train_steps (see pk_universe_kernel_class)");
            }
            for (const auto& u : cls[c]) {
                K2Smem own(cfg->model, d, k, W, u.n_ent, u.n_rel, u.batch_size, 0, fast);
                const size_t stride1 = (size_t)(2 + k) * u.batch_size * (cfg->model == PK_TRANSD ? 2 : 1) * d;
                DescSlot* slot1 = acquire_desc(kDescArea + stride1 * sizeof(float));
                if (!slot1) return pk::cuda_fail(cudaGetLastError(), "pk_train_universes: descriptor buffer");
                const pk_universe_desc* staged1 = slot1->stage(&u, sizeof(pk_universe_desc));
                if (!staged1) return pk::cuda_fail(cudaGetLastError(), "pk_train_universes: pinned descriptor staging");
                PK_CUDA(cudaMemcpyAsync(slot1->d, staged1, sizeof(pk_universe_desc), cudaMemcpyHostToDevice, st));
                K2Params P1;
                P1.desc = slot1->d;
                for (int i = 0; i < 2; ++i) {
                    P1.ent[i] = packed->ent[i]; P1.rel[i] = packed->rel[i];
                    P1.ent_state[i] = packed->ent_state[i]; P1.rel_state[i] = packed->rel_state[i];
                }
                P1.by_head = d_by_head; P1.by_tail = d_by_tail; P1.left_mean = d_left_mean; P1.right_mean = d_right_mean;
                P1.loss = d_loss;
                P1.d = d; P1.k = k; P1.p_norm = cfg->p_norm; P1.norm_flag = cfg->norm_flag; P1.opt = cfg->opt;
                P1.bern = cfg->bern; P1.filter = cfg->filter; P1.W = W;
                P1.mE = u.n_ent; P1.mR = u.n_rel; P1.mB = u.batch_size;
                P1.np = np;
                P1.timer = nullptr; P1.timer_base = 0;
                P1.scratch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(slot1->d) + kDescArea);
                P1.scratch_stride = stride1;
                int rc1 = cfg->model == PK_TRANSE ? launch_model0(lay, P1, 0, 1, own.total, st)
                          : (cfg->model == PK_TRANSH ? launch_model1(lay, P1, 0, 1, own.total, st) : launch_model2(lay, P1, 0, 1, own.total, st));
                if (rc1 != PK_OK) return rc1;
                slot1->busy = true;
                slot1->seq = ++g_desc_seq;
                PK_CUDA(cudaEventRecord(slot1->done, st));
            }
            continue;
        }
        // the longest universes first: blocks are scheduled in index order
        std::stable_sort(cls[c].begin(), cls[c].end(), [](const pk_universe_desc& a, const pk_universe_desc& b) {
            return (long long)a.epochs * a.nbatches * a.batch_size > (long long)b.epochs * b.nbatches * b.batch_size;
        });
        const size_t bytes = cls[c].size() * sizeof(pk_universe_desc);
        // behind the descriptors: every block's scratch rows (zeroed by the block itself)
        const size_t desc_area = (bytes + 255) & ~(size_t)255;
        const size_t stride = (size_t)(2 + k) * mB[c] * (cfg->model == PK_TRANSD ? 2 : 1) * d;
        DescSlot* slot = acquire_desc(desc_area + cls[c].size() * stride * sizeof(float));
        if (!slot) return pk::cuda_fail(cudaGetLastError(), "pk_train_universes: descriptor buffer");
        pk_universe_desc* d_desc = slot->d;
        const pk_universe_desc* staged = slot->stage(cls[c].data(), bytes);
        if (!staged) return pk::cuda_fail(cudaGetLastError(), "pk_train_universes: pinned descriptor staging");
        PK_CUDA(cudaMemcpyAsync(d_desc, staged, bytes, cudaMemcpyHostToDevice, st));
        K2Params P;
        P.desc = d_desc;
        for (int i = 0; i < 2; ++i) {
            P.ent[i] = packed->ent[i]; P.rel[i] = packed->rel[i];
            P.ent_state[i] = packed->ent_state[i]; P.rel_state[i] = packed->rel_state[i];
        }
        P.by_head = d_by_head; P.by_tail = d_by_tail; P.left_mean = d_left_mean; P.right_mean = d_right_mean;
        P.loss = d_loss;
        P.d = d; P.k = k; P.p_norm = cfg->p_norm; P.norm_flag = cfg->norm_flag; P.opt = cfg->opt;
        P.bern = cfg->bern; P.filter = cfg->filter; P.W = W;
        P.mE = mE[c]; P.mR = mR[c]; P.mB = mB[c];
        P.np = np;
        P.timer = g_timer; P.timer_base = c == 1 ? (int)cls[0].size() : 0;
        P.scratch = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(slot->d) + desc_area);
        P.scratch_stride = stride;
        // the slot (device buffer + pinned staging) is handed out again only after this launch's `done` event
        int rc = PK_OK;
        const int nblk = (int)cls[c].size();
        if (cfg->model == PK_TRANSE) rc = launch_model0(lay, P, c == 0, nblk, s.total, st);
        else if (cfg->model == PK_TRANSH) rc = launch_model1(lay, P, c == 0, nblk, s.total, st);
        else rc = launch_model2(lay, P, c == 0, nblk, s.total, st);
        if (rc != PK_OK) return rc;
        slot->busy = true;
        slot->seq = ++g_desc_seq;
        PK_CUDA(cudaEventRecord(slot->done, st));
    }
    if (both) {
        PK_CUDA(cudaEventRecord(g_side.join, g_side.st));
        PK_CUDA(cudaStreamWaitEvent(caller, g_side.join, 0));
    }
    return PK_OK;
}
#endif  // !PK_MODEL_TU
