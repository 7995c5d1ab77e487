// K0 + K1 — device sampler and the fused single-space train step.
//
// One reference training step (openke/config/Trainer.py:44-56) is: sampling() on the host
// (openke/base/Base.cpp:266-310), 4 H2D copies, ~100-190 ATen kernels (gather, normalise,
// translate, norm, margin loss, autograd with a DENSE [E,d] gradient, dense optimizer) and a
// loss.item() sync.  Here it is three launches that never leave the device:
//
//   k1_sample_count  draws the batch with the reference's LCG streams (bit-exact, jump-ahead per
//                    sample) and counts how often each table row occurs in the batch;
//   k1_grad          one lane group per positive sample: gather h/r/t and the corrupted rows with
//                    vector loads, forward, margin loss, analytic backward.  A row that occurs
//                    ONCE in the batch is updated in place by the group that read it (read once,
//                    written once: the algorithmic minimum).  A row that occurs several times has
//                    its gradient summed into a dense accumulator with atomics and is queued;
//   k1_apply         optimizer step for the queued rows, accumulators and counters back to zero,
//                    per-block loss partials reduced in a fixed order.
//
// The sparse update is exactly the reference's dense one: SGD and Adagrad (lr_decay = 0,
// weight_decay = 0, reference Trainer.py:34-35,65-70,84-88) leave zero-gradient rows bit-unchanged.
// Bound on big tables (E*d*4 >> L2): HBM, ~(3+k) rows read + written per positive.
#include <algorithm>
#include <map>
#include <vector>

#include "common.hpp"
#include "kge_device.cuh"

using namespace pkd;

struct pk_workspace {
    pk_model_cfg cfg;
    int64_t n_ent, n_rel, max_batch;
    int32_t* cnt_ent = nullptr;   // [n_ent] occurrences in the current batch (bit 30: queued)
    int32_t* cnt_rel = nullptr;   // [n_rel]
    float* acc_ent[2] = {nullptr, nullptr};  // dense gradient accumulators for multiply-occurring rows
    float* acc_rel[2] = {nullptr, nullptr};
    int32_t* dup_ent = nullptr;   // queue of multiply-occurring rows
    int32_t* dup_rel = nullptr;
    int32_t* counters = nullptr;  // [0] queued entities, [1] queued relations, [2] error flag
    int64_t* step_ctr = nullptr;  // steps taken since the LCG base state was last committed
    int32_t* ids = nullptr;       // [3 * max_batch * (1+k)] sampled batch (h | t | r)
    float* loss_part = nullptr;   // per-block partial sums of k1_grad
    cudaStream_t own_stream = nullptr;  // blocking stream used when the caller hands us the legacy default stream
    int64_t dup_cap_ent = 0, dup_cap_rel = 0;
    int max_blocks = 0;
};

namespace pkk1 {

constexpr int K1_THREADS = 256;
constexpr int32_t QUEUED = 1 << 30;

struct K1Params {
    float* ent[2];
    float* rel[2];
    float* ent_state[2];
    float* rel_state[2];
    float* acc_ent[2];
    float* acc_rel[2];
    int32_t* cnt_ent;
    int32_t* cnt_rel;
    int32_t* dup_ent;
    int32_t* dup_rel;
    int32_t* counters;
    int64_t* step_ctr;
    const int32_t* bh;
    const int32_t* bt;
    const int32_t* br;
    float* loss_part;
    float* loss;        // [steps] indexed by *step_ctr
    int64_t B;
    int64_t n_ent, n_rel;
    int d, k, p_norm, norm_flag, opt;
    float margin, lr;
    int grad_blocks;
};

#ifdef PK_MODEL_TU
template <class L>
__device__ __forceinline__ void k1_apply_update(float* x_row, float* s_row, const float (&g)[L::NF], int d, int lane, int opt, float lr) {
    float x[L::NF];
    ld_row<L>(x_row, d, lane, x);
    if (opt == PK_ADAGRAD) {
        float s[L::NF];
        ld_row<L>(s_row, d, lane, s);
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

template <class L>
struct K1Ctx {
    const K1Params* P;
    __device__ __forceinline__ const float* ent_row(int tbl, int id) const { return P->ent[tbl] + (size_t)id * P->d; }
    __device__ __forceinline__ const float* rel_row(int tbl, int id) const { return P->rel[tbl] + (size_t)id * P->d; }
    __device__ __forceinline__ void emit(float* table, float* state, float* acc, int32_t* cnt, int32_t* queue, int32_t* qn,
                                         bool first_table, int id, const float (&g)[L::NF], int lane) const {
        const int d = P->d;
        const int32_t c = cnt[id];
        if ((c & ~QUEUED) == 1) {  // the only occurrence in this batch: nobody else reads or writes the row
            k1_apply_update<L>(table + (size_t)id * d, state ? state + (size_t)id * d : nullptr, g, d, lane, P->opt, P->lr);
        } else {
            float* p = acc + (size_t)id * d;
#pragma unroll
            for (int i = 0; i < L::NF; ++i) {
                const int e = elem_of<L>(lane, i);
                if (e < d && g[i] != 0.f) atomicAdd(p + e, g[i]);
            }
            if (first_table && lane == 0 && !(c & QUEUED)) {
                const int32_t old = atomicOr(&cnt[id], QUEUED);
                if (!(old & QUEUED)) queue[atomicAdd(qn, 1)] = id;
            }
        }
    }
    __device__ __forceinline__ void add_ent(int tbl, int id, const float (&g)[L::NF], int lane, bool pred) const {
        if (pred) emit(P->ent[tbl], P->ent_state[tbl], P->acc_ent[tbl], P->cnt_ent, P->dup_ent, P->counters + 0, tbl == 0, id, g, lane);
    }
    __device__ __forceinline__ void add_rel(int tbl, int id, const float (&g)[L::NF], int lane, bool pred) const {
        if (pred) emit(P->rel[tbl], P->rel_state[tbl], P->acc_rel[tbl], P->cnt_rel, P->dup_rel, P->counters + 1, tbl == 0, id, g, lane);
    }
};

#endif

#ifndef PK_MODEL_TU
// ---- K0: the reference sampling() on the device, optionally followed by the occurrence count
struct SampleParams {
    SamplerView sv;
    const uint64_t* lcg;       // W stream states at the last commit
    const int64_t* step_ctr;   // batches drawn since then (may be NULL: 0)
    int32_t* bh;
    int32_t* bt;
    int32_t* br;
    int32_t* cnt_ent;          // may be NULL: sample only
    int32_t* cnt_rel;
    int32_t* counters;
    int64_t B;
    int W, k, bern, filter;
};

__device__ __forceinline__ void count_sample(int32_t* cnt_ent, int32_t* cnt_rel, const int32_t* bh, const int32_t* bt,
                                             const int32_t* br, int64_t B, int k, int64_t b) {
    const int32_t h = bh[b], t = bt[b];
    atomicAdd(&cnt_ent[h], 1);
    atomicAdd(&cnt_ent[t], 1);
    atomicAdd(&cnt_rel[br[b]], 1);
    for (int j = 0; j < k; ++j) {
        const int64_t o = b + (int64_t)(1 + j) * B;
        const int32_t nh = bh[o], nt = bt[o];
        if (nh != h) atomicAdd(&cnt_ent[nh], 1);
        if (nt != t) atomicAdd(&cnt_ent[nt], 1);
    }
}

__global__ void __launch_bounds__(K1_THREADS) k1_sample_count(const __grid_constant__ SampleParams S) {
    const int64_t b = (int64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    if (b == 0 && S.counters) { S.counters[0] = 0; S.counters[1] = 0; }
    if (b >= S.B) return;
    // the stream's state at the start of THIS batch: `done` whole batches after the last commit
    int64_t j, lef, rig;
    const int id = stream_of(S.B, S.W, b, j);
    slice_of(S.B, S.W, id, lef, rig);
    const uint64_t done = S.step_ctr ? (uint64_t)*S.step_ctr : 0;
    const uint64_t s0 = lcg_skip(S.lcg[id], done * (uint64_t)(rig - lef) * (uint64_t)(1 + 2 * S.k));
    sample_one(S.sv, s0, j, S.B, S.k, S.bern != 0, S.filter != 0, b, S.bh, S.bt, S.br);
    if (S.cnt_ent) count_sample(S.cnt_ent, S.cnt_rel, S.bh, S.bt, S.br, S.B, S.k, b);
}

__global__ void __launch_bounds__(K1_THREADS) k1_count(int32_t* cnt_ent, int32_t* cnt_rel, int32_t* counters, const int32_t* bh,
                                                       const int32_t* bt, const int32_t* br, int64_t B, int k, int64_t n_ent,
                                                       int64_t n_rel) {
    const int64_t b = (int64_t)blockIdx.x * K1_THREADS + threadIdx.x;
    if (b == 0) { counters[0] = 0; counters[1] = 0; }
    if (b >= B) return;
    // ids come from the caller here: refuse out-of-range rows instead of corrupting memory
    bool ok = br[b] >= 0 && br[b] < n_rel;
    for (int j = 0; j <= k; ++j) {
        const int64_t o = b + (int64_t)j * B;
        ok = ok && bh[o] >= 0 && bh[o] < n_ent && bt[o] >= 0 && bt[o] < n_ent && br[o] == br[b];
    }
    if (!ok) { atomicExch(&counters[2], 1); return; }
    count_sample(cnt_ent, cnt_rel, bh, bt, br, B, k, b);
}

// advance the W stream states by `fixed` batches, or by *step_ctr batches (then zero it)
__global__ void k1_commit_lcg(uint64_t* lcg, int64_t* step_ctr, int64_t fixed, int64_t B, int W, int k) {
    const int id = threadIdx.x;
    const uint64_t done = step_ctr ? (uint64_t)*step_ctr : (uint64_t)fixed;
    uint64_t s = 0;
    if (id < W) {
        int64_t lef, rig;
        slice_of(B, W, id, lef, rig);
        s = lcg_skip(lcg[id], done * (uint64_t)(rig - lef) * (uint64_t)(1 + 2 * k));
    }
    __syncthreads();
    if (id < W) lcg[id] = s;
    if (id == 0 && step_ctr) *step_ctr = 0;
}

#endif

#ifdef PK_MODEL_TU
// ---- K1 main: forward + backward + in-place update of singly-occurring rows
template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS) k1_grad(const __grid_constant__ K1Params P) {
    constexpr int NG = K1_THREADS / L::G;
    const int tid = threadIdx.x, lane = tid % L::G, grp = tid / L::G;
    __shared__ float part[NG];
    Hyper hp;
    hp.d = P.d; hp.k = P.k; hp.p_norm = P.p_norm; hp.norm_flag = P.norm_flag;
    hp.margin = P.margin;
    hp.inv_bk = 1.f / (float)(P.B * P.k);
    K1Ctx<L> cx;
    cx.P = &P;
    float acc = 0.f;
    const bool bad = P.counters[2] != 0;
    for (int64_t base = (int64_t)blockIdx.x * NG; base < P.B; base += (int64_t)gridDim.x * NG) {
        const int64_t b = base + grp;
        const bool act = b < P.B && !bad;
        const float l = process_sample<MODEL, L>(cx, hp, lane, P.B, b, act, P.bh, P.bt, P.br);
        if (act) acc += l;
    }
    if (lane == 0) part[grp] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int i = 0; i < NG; ++i) s += part[i];
        P.loss_part[blockIdx.x] = s;
    }
}

// ---- K1 tail: optimizer for multiply-occurring rows, cleanup, loss
template <int MODEL, class L>
__global__ void __launch_bounds__(K1_THREADS) k1_apply(const __grid_constant__ K1Params P) {
    constexpr int NG = K1_THREADS / L::G;
    constexpr int ntE = MODEL == TRANSD ? 2 : 1, ntR = MODEL == TRANSE ? 1 : 2;
    const int tid = threadIdx.x, lane = tid % L::G, grp = tid / L::G;
    const int d = P.d;
    const int nqe = P.counters[0], nqr = P.counters[1];
    for (int64_t s = (int64_t)blockIdx.x * NG + grp; s < (int64_t)nqe + nqr; s += (int64_t)gridDim.x * NG) {
        const bool is_ent = s < nqe;
        const int id = is_ent ? P.dup_ent[s] : P.dup_rel[s - nqe];
        const int nt = is_ent ? ntE : ntR;
        for (int t = 0; t < nt; ++t) {
            float* arow = (is_ent ? P.acc_ent[t] : P.acc_rel[t]) + (size_t)id * d;
            float g[L::NF];
            ld_row<L>(arow, d, lane, g);
            float* table = is_ent ? P.ent[t] : P.rel[t];
            float* state = is_ent ? P.ent_state[t] : P.rel_state[t];
            k1_apply_update<L>(table + (size_t)id * d, state ? state + (size_t)id * d : nullptr, g, d, lane, P.opt, P.lr);
#pragma unroll
            for (int i = 0; i < L::NF; ++i) g[i] = 0.f;
            st_row<L>(arow, d, lane, g);
        }
    }
    // occurrence counters back to zero (every id of the batch; equal values race benignly)
    const int64_t nids = P.B * (1 + P.k);
    for (int64_t i = (int64_t)blockIdx.x * K1_THREADS + tid; i < nids; i += (int64_t)gridDim.x * K1_THREADS) {
        P.cnt_ent[P.bh[i]] = 0;
        P.cnt_ent[P.bt[i]] = 0;
        if (i < P.B) P.cnt_rel[P.br[i]] = 0;
    }
    if (blockIdx.x == 0 && tid < 32) {
        float s = 0.f;
        for (int i = tid; i < P.grad_blocks; i += 32) s += P.loss_part[i];
        s = gsum<32>(s);
        if (tid == 0) {
            const int64_t step = *P.step_ctr;
            if (P.loss) P.loss[step] = s / (float)(P.B * P.k) + P.margin;
            *P.step_ctr = step + 1;
        }
    }
}

#endif

struct LaySel { int V, G, CPL; };
inline LaySel pick_layout(int model, int d) {
    const int V = d % 4 == 0 ? 4 : (d % 2 == 0 ? 2 : 1);
    const int chunks = d / V;
    const int nf_cap = model == TRANSD ? 4 : 8;
    for (int G : {8, 32}) {
        int cpl = (chunks + G - 1) / G, c2 = 1;
        while (c2 < cpl) c2 *= 2;
        if (c2 * V <= nf_cap || G == 32) return LaySel{V, G, std::max(c2, 1)};
    }
    return LaySel{V, 32, 1};
}

#ifdef PK_MODEL_TU
template <int MODEL, int V, int G, int CPL>
int launch_step(const K1Params& P, int grad_blocks, int apply_blocks, cudaStream_t st) {
    k1_grad<MODEL, Lay<V, G, CPL>><<<grad_blocks, K1_THREADS, 0, st>>>(P);
    PK_LAUNCHED("k1_grad");
    k1_apply<MODEL, Lay<V, G, CPL>><<<apply_blocks, K1_THREADS, 0, st>>>(P);
    PK_LAUNCHED("k1_apply");
    return PK_OK;
}

template <int MODEL>
int dispatch_step(const LaySel& l, const K1Params& P, int gb, int ab, cudaStream_t st) {
#define PK_CASE(v, g, c) if (l.V == v && l.G == g && l.CPL == c) return launch_step<MODEL, v, g, c>(P, gb, ab, st);
    PK_CASE(4, 8, 1) PK_CASE(4, 8, 2) PK_CASE(4, 32, 1) PK_CASE(4, 32, 2)
    PK_CASE(2, 8, 1) PK_CASE(2, 8, 2) PK_CASE(2, 8, 4) PK_CASE(2, 32, 1) PK_CASE(2, 32, 2) PK_CASE(2, 32, 4)
    PK_CASE(1, 8, 1) PK_CASE(1, 8, 2) PK_CASE(1, 8, 4) PK_CASE(1, 8, 8) PK_CASE(1, 32, 1) PK_CASE(1, 32, 2) PK_CASE(1, 32, 4) PK_CASE(1, 32, 8)
#undef PK_CASE
    return pk::fail(PK_ERR_UNSUPPORTED, "embedding dimension not supported by the train step (d <= 256)");
}

#define PK_CAT2(a, b) a##b
#define PK_CAT(a, b) PK_CAT2(a, b)
int PK_CAT(step_model, PK_MODEL_TU)(const LaySel& l, const K1Params& P, int gb, int ab, cudaStream_t st) {
    return dispatch_step<PK_MODEL_TU>(l, P, gb, ab, st);
}
}  // namespace pkk1
#else
int step_model0(const LaySel& l, const K1Params& P, int gb, int ab, cudaStream_t st);
int step_model1(const LaySel& l, const K1Params& P, int gb, int ab, cudaStream_t st);
int step_model2(const LaySel& l, const K1Params& P, int gb, int ab, cudaStream_t st);

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

void fill_params(K1Params& P, const pk_model_cfg* cfg, const pk_tables* tab, pk_workspace* ws, int64_t B, const int32_t* h,
                 const int32_t* t, const int32_t* r, float margin, float lr, float* d_loss) {
    for (int i = 0; i < 2; ++i) {
        P.ent[i] = tab->ent[i]; P.rel[i] = tab->rel[i];
        P.ent_state[i] = cfg->opt == PK_ADAGRAD ? tab->ent_state[i] : nullptr;
        P.rel_state[i] = cfg->opt == PK_ADAGRAD ? tab->rel_state[i] : nullptr;
        P.acc_ent[i] = ws->acc_ent[i]; P.acc_rel[i] = ws->acc_rel[i];
    }
    P.cnt_ent = ws->cnt_ent; P.cnt_rel = ws->cnt_rel;
    P.dup_ent = ws->dup_ent; P.dup_rel = ws->dup_rel;
    P.counters = ws->counters; P.step_ctr = ws->step_ctr;
    P.bh = h; P.bt = t; P.br = r;
    P.loss_part = ws->loss_part;
    P.loss = d_loss;
    P.B = B; P.n_ent = ws->n_ent; P.n_rel = ws->n_rel;
    P.d = cfg->dim; P.k = cfg->neg_ent; P.p_norm = cfg->p_norm; P.norm_flag = cfg->norm_flag; P.opt = cfg->opt;
    P.margin = margin; P.lr = lr;
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

int run_step(const pk_model_cfg* cfg, const K1Params& P0, pk_workspace* ws, cudaStream_t st) {
    K1Params P = P0;
    const LaySel lay = pick_layout(cfg->model, cfg->dim);
    const int ng = K1_THREADS / lay.G;
    int gb = (int)std::min<int64_t>((P.B + ng - 1) / ng, (int64_t)ws->max_blocks);
    gb = std::max(gb, 1);
    P.grad_blocks = gb;
    const int ab = std::max(1, std::min(ws->max_blocks, (int)((P.B * (1 + P.k) + K1_THREADS - 1) / K1_THREADS)));
    if (cfg->model == PK_TRANSE) return step_model0(lay, P, gb, ab, st);
    if (cfg->model == PK_TRANSH) return step_model1(lay, P, gb, ab, st);
    return step_model2(lay, P, gb, ab, st);
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
    ws->dup_cap_ent = std::min<int64_t>(n_ent, max_batch * (2 + k));
    ws->dup_cap_rel = std::min<int64_t>(n_rel, max_batch);
    ws->max_blocks = num_sms() * 8;
    auto alloc0 = [&](void** p, size_t bytes) -> bool {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        return cudaMemset(*p, 0, bytes) == cudaSuccess;
    };
    bool ok = alloc0((void**)&ws->cnt_ent, (size_t)n_ent * 4) && alloc0((void**)&ws->cnt_rel, (size_t)n_rel * 4) &&
              alloc0((void**)&ws->dup_ent, (size_t)ws->dup_cap_ent * 4) && alloc0((void**)&ws->dup_rel, (size_t)ws->dup_cap_rel * 4) &&
              alloc0((void**)&ws->counters, 16) && alloc0((void**)&ws->step_ctr, 8) &&
              alloc0((void**)&ws->ids, (size_t)3 * max_batch * (1 + k) * 4) && alloc0((void**)&ws->loss_part, (size_t)ws->max_blocks * 4);
    for (int i = 0; ok && i < ntE; ++i) ok = alloc0((void**)&ws->acc_ent[i], (size_t)n_ent * d * 4);
    for (int i = 0; ok && i < ntR; ++i) ok = alloc0((void**)&ws->acc_rel[i], (size_t)n_rel * d * 4);
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
    cudaFree(ws->cnt_ent); cudaFree(ws->cnt_rel); cudaFree(ws->dup_ent); cudaFree(ws->dup_rel);
    cudaFree(ws->counters); cudaFree(ws->step_ctr); cudaFree(ws->ids); cudaFree(ws->loss_part);
    for (int i = 0; i < 2; ++i) { cudaFree(ws->acc_ent[i]); cudaFree(ws->acc_rel[i]); }
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
    S.lcg = smp->lcg; S.step_ctr = nullptr;
    S.bh = d_h; S.bt = d_t; S.br = d_r;
    S.cnt_ent = nullptr; S.cnt_rel = nullptr; S.counters = nullptr;
    S.B = B; S.W = cfg->work_threads; S.k = cfg->neg_ent; S.bern = cfg->bern; S.filter = cfg->filter;
    k1_sample_count<<<(unsigned)((B + K1_THREADS - 1) / K1_THREADS), K1_THREADS, 0, st>>>(S);
    PK_LAUNCHED("k1_sample_count");
    k1_commit_lcg<<<1, 64, 0, st>>>(smp->lcg, nullptr, 1, B, cfg->work_threads, cfg->neg_ent);
    PK_LAUNCHED("k1_commit_lcg");
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
    fill_params(P, cfg, tab, ws, B, d_h, d_t, d_r, margin, lr, d_loss);
    k1_count<<<(unsigned)((B + K1_THREADS - 1) / K1_THREADS), K1_THREADS, 0, st>>>(ws->cnt_ent, ws->cnt_rel, ws->counters, d_h, d_t,
                                                                                    d_r, B, cfg->neg_ent, ws->n_ent, ws->n_rel);
    PK_LAUNCHED("k1_count");
    rc = run_step(cfg, P, ws, st);
    if (rc != PK_OK) return rc;
    // single steps write loss[0]: rewind the step counter (stream-ordered)
    PK_CUDA(cudaMemsetAsync(ws->step_ctr, 0, 8, st));
    return PK_OK;
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
    const int k = cfg->neg_ent;
    int32_t* bh = ws->ids;
    int32_t* bt = bh + B * (1 + k);
    int32_t* br = bt + B * (1 + k);
    SampleParams S;
    S.sv.by_head = smp->by_head; S.sv.by_tail = smp->by_tail; S.sv.left_mean = smp->left_mean; S.sv.right_mean = smp->right_mean;
    S.sv.n_tri = smp->n_tri; S.sv.n_ent = (int32_t)smp->n_ent; S.sv.n_rel = (int32_t)smp->n_rel;
    S.lcg = smp->lcg; S.step_ctr = ws->step_ctr;
    S.bh = bh; S.bt = bt; S.br = br;
    S.cnt_ent = ws->cnt_ent; S.cnt_rel = ws->cnt_rel; S.counters = ws->counters;
    S.B = B; S.W = cfg->work_threads; S.k = k; S.bern = cfg->bern; S.filter = cfg->filter;
    K1Params P;
    fill_params(P, cfg, tab, ws, B, bh, bt, br, margin, lr, d_loss);
    const unsigned sb = (unsigned)((B + K1_THREADS - 1) / K1_THREADS);

    // Every launch parameter is step-invariant (the step index lives in *step_ctr on the device),
    // so a chunk of steps is captured once into a CUDA graph and replayed.
    const int64_t chunk = std::min<int64_t>(steps, 64);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int64_t done = 0;
    if (steps >= 4) {
        PK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int64_t i = 0; i < chunk; ++i) {
            k1_sample_count<<<sb, K1_THREADS, 0, st>>>(S);
            ++pk::launch_counter();
            rc = run_step(cfg, P, ws, st);
            if (rc != PK_OK) break;
        }
        cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (rc != PK_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return pk::cuda_fail(ce, "cudaStreamEndCapture");
        const int per_chunk = pk::launch_counter();
        PK_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        int launches = 0;
        for (; done + chunk <= steps; done += chunk) {
            PK_CUDA(cudaGraphLaunch(exec, st));
            launches += per_chunk;
        }
        pk::launch_counter() = launches;
    }
    for (; done < steps; ++done) {
        k1_sample_count<<<sb, K1_THREADS, 0, st>>>(S);
        PK_LAUNCHED("k1_sample_count");
        rc = run_step(cfg, P, ws, st);
        if (rc != PK_OK) break;
    }
    if (rc == PK_OK) {
        k1_commit_lcg<<<1, 64, 0, st>>>(smp->lcg, ws->step_ctr, 0, B, cfg->work_threads, k);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = pk::cuda_fail(e, "k1_commit_lcg");
        else ++pk::launch_counter();
    }
    if (exec) {
        // the graph must outlive its queued launches
        cudaStreamSynchronize(st);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
    }
    return rc;
}

extern "C" int pk_workspace_check(pk_workspace* ws, void* stream) {
    if (!ws) return pk::fail(PK_ERR_ARG, "pk_workspace_check: null workspace");
    int32_t c[4] = {0, 0, 0, 0};
    PK_CUDA(cudaMemcpyAsync(c, ws->counters, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (c[2]) {
        cudaMemsetAsync(ws->counters + 2, 0, 4, (cudaStream_t)stream);
        return pk::fail(PK_ERR_ARG, "train step refused a batch: an id is out of range, or a negative does not share its positive's relation");
    }
    return PK_OK;
}
#endif  // !PK_MODEL_TU
