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
// Data layout.  All universes of a launch are packed: entity tables [sum nE, d], relation tables
// [sum nR, d], sorted triple lists [sum nT, 3]; a descriptor per universe holds the offsets.  A
// block stages its universe's tables in shared memory when they fit (local ids are dense, so the
// staged table IS the universe's whole embedding space); the per-step gradient scratch (one row per
// touched table row), the slot maps and the batch ids always live in shared memory.  The Adagrad
// accumulators stay in global memory (L2-resident: touched once per step, off the critical path).
//
// Bound: this kernel is latency-bound by construction (a universe's steps are strictly sequential
// and a step touches < 100 KB); the roofline that matters is "steps per second per SM".
#include <algorithm>
#include <vector>

#include "common.hpp"
#include "kge_device.cuh"

namespace pkk2 {

using namespace pkd;

constexpr int K2_THREADS = 256;

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
    int stage;              // tables staged in shared memory for every universe of this launch
    int mE, mR, mB;         // launch-wide maxima: shared-memory carve-up is uniform
};

__host__ __device__ inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

// shared-memory carve-up, identical on host (sizing) and device (pointers)
struct K2Smem {
    size_t ent[2], rel[2], gent[2], grel[2], lossv, slot_ent, slot_rel, touched_ent, touched_rel, ids, counters, lcg, total;
    int slotsE, slotsR;
    __host__ __device__ K2Smem(int model, int d, int k, int mE, int mR, int mB, int stage) {
        const int ntE = model == TRANSD ? 2 : 1, ntR = model == TRANSE ? 1 : 2;
        slotsE = min_((long long)mE, (long long)(2 + k) * mB);
        slotsR = min_((long long)mR, (long long)mB);
        size_t o = 0;
        for (int i = 0; i < 2; ++i) { ent[i] = o; if (stage && i < ntE) o = up16(o + (size_t)mE * d * 4); }
        for (int i = 0; i < 2; ++i) { rel[i] = o; if (stage && i < ntR) o = up16(o + (size_t)mR * d * 4); }
        for (int i = 0; i < 2; ++i) { gent[i] = o; if (i < ntE) o = up16(o + (size_t)slotsE * d * 4); }
        for (int i = 0; i < 2; ++i) { grel[i] = o; if (i < ntR) o = up16(o + (size_t)slotsR * d * 4); }
        lossv = o;       o = up16(o + (size_t)mB * 4);
        slot_ent = o;    o = up16(o + (size_t)mE * 4);
        slot_rel = o;    o = up16(o + (size_t)mR * 4);
        touched_ent = o; o = up16(o + (size_t)slotsE * 4);
        touched_rel = o; o = up16(o + (size_t)slotsR * 4);
        ids = o;         o = up16(o + (size_t)3 * mB * (1 + k) * 4);
        counters = o;    o = up16(o + 16);
        lcg = o;         o = up16(o + 8 * 8);
        total = o;
    }
    __host__ __device__ static int min_(long long a, long long b) { return (int)(a < b ? a : b); }
};

#ifdef PK_MODEL_TU
template <class L>
struct K2Ctx {
    float* ent[2];
    float* rel[2];
    float* gent[2];
    float* grel[2];
    const int* slot_ent;
    const int* slot_rel;
    int d;
    __device__ __forceinline__ const float* ent_row(int tbl, int id) const { return ent[tbl] + (size_t)id * d; }
    __device__ __forceinline__ const float* rel_row(int tbl, int id) const { return rel[tbl] + (size_t)id * d; }
    __device__ __forceinline__ void add(float* base, int slot, const float (&g)[L::NF], int lane) const {
        float* p = base + (size_t)slot * d;
#pragma unroll
        for (int i = 0; i < L::NF; ++i) {
            const int e = elem_of<L>(lane, i);
            if (e < d && g[i] != 0.f) atomicAdd(p + e, g[i]);
        }
    }
    __device__ __forceinline__ void add_ent(int tbl, int id, const float (&g)[L::NF], int lane, bool pred) const {
        if (pred) add(gent[tbl], slot_ent[id], g, lane);
    }
    __device__ __forceinline__ void add_rel(int tbl, int id, const float (&g)[L::NF], int lane, bool pred) const {
        if (pred) add(grel[tbl], slot_rel[id], g, lane);
    }
};

// x <- optimizer(x, g): SGD  x -= lr g ;  Adagrad  s += g^2, x -= lr g / (sqrt(s) + 1e-10)
// (torch.optim.SGD / Adagrad as configured by reference Trainer.py:65-70,84-88; lr_decay = weight_decay = 0)
template <class L>
__device__ __forceinline__ void apply_update(float* x_row, float* s_row, const float (&g)[L::NF], int d, int lane, int opt, float lr) {
    float x[L::NF];
    ld_row<L>(x_row, d, lane, x);
    if (opt == PK_ADAGRAD) {
        float s[L::NF];
        ld_row<L>(s_row, d, lane, s);
#pragma unroll
        for (int i = 0; i < L::NF; ++i) {
            s[i] = fmaf(g[i], g[i], s[i]);
            x[i] = x[i] + div0(-lr * g[i], sqrt0(s[i]) + 1e-10f);
        }
        st_row<L>(s_row, d, lane, s);
    } else {
#pragma unroll
        for (int i = 0; i < L::NF; ++i) x[i] = fmaf(-lr, g[i], x[i]);
    }
    st_row<L>(x_row, d, lane, x);
}

template <int MODEL, class L>
__global__ void __launch_bounds__(K2_THREADS) k2_train_universes(const __grid_constant__ K2Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int ntE = MODEL == TRANSD ? 2 : 1, ntR = MODEL == TRANSE ? 1 : 2;
    constexpr int NG = K2_THREADS / L::G;
    const pk_universe_desc& U = P.desc[blockIdx.x];
    const K2Smem S(MODEL, P.d, P.k, P.mE, P.mR, P.mB, P.stage);
    const int tid = threadIdx.x, lane = tid % L::G, grp = tid / L::G;
    const int d = P.d, k = P.k, B = U.batch_size, nE = U.n_ent, nR = U.n_rel;

    float* g_ent[2];   // this universe's tables in global memory
    float* g_rel[2];
    float* st_ent[2];
    float* st_rel[2];
    for (int i = 0; i < 2; ++i) {
        g_ent[i] = (i < ntE) ? P.ent[i] + (size_t)U.ent_off * d : nullptr;
        g_rel[i] = (i < ntR) ? P.rel[i] + (size_t)U.rel_off * d : nullptr;
        st_ent[i] = (i < ntE && P.opt == PK_ADAGRAD) ? P.ent_state[i] + (size_t)U.ent_off * d : nullptr;
        st_rel[i] = (i < ntR && P.opt == PK_ADAGRAD) ? P.rel_state[i] + (size_t)U.rel_off * d : nullptr;
    }
    K2Ctx<L> cx;
    cx.d = d;
    for (int i = 0; i < 2; ++i) {
        cx.ent[i] = P.stage ? reinterpret_cast<float*>(smem + S.ent[i]) : g_ent[i];
        cx.rel[i] = P.stage ? reinterpret_cast<float*>(smem + S.rel[i]) : g_rel[i];
        cx.gent[i] = reinterpret_cast<float*>(smem + S.gent[i]);
        cx.grel[i] = reinterpret_cast<float*>(smem + S.grel[i]);
    }
    float* lossv = reinterpret_cast<float*>(smem + S.lossv);
    int* slot_ent = reinterpret_cast<int*>(smem + S.slot_ent);
    int* slot_rel = reinterpret_cast<int*>(smem + S.slot_rel);
    int* touched_ent = reinterpret_cast<int*>(smem + S.touched_ent);
    int* touched_rel = reinterpret_cast<int*>(smem + S.touched_rel);
    int32_t* bh = reinterpret_cast<int32_t*>(smem + S.ids);
    int32_t* bt = bh + (size_t)B * (1 + k);
    int32_t* br = bt + (size_t)B * (1 + k);
    int* counters = reinterpret_cast<int*>(smem + S.counters);  // [0] touched entities, [1] touched relations
    uint64_t* lcg = reinterpret_cast<uint64_t*>(smem + S.lcg);
    cx.slot_ent = slot_ent;
    cx.slot_rel = slot_rel;

    // ---- stage tables, clear scratch
    if (P.stage) {
        for (int t = 0; t < ntE; ++t)
            for (int i = tid; i < nE * d; i += K2_THREADS) cx.ent[t][i] = g_ent[t][i];
        for (int t = 0; t < ntR; ++t)
            for (int i = tid; i < nR * d; i += K2_THREADS) cx.rel[t][i] = g_rel[t][i];
    }
    for (int t = 0; t < ntE; ++t)
        for (int i = tid; i < S.slotsE * d; i += K2_THREADS) cx.gent[t][i] = 0.f;
    for (int t = 0; t < ntR; ++t)
        for (int i = tid; i < S.slotsR * d; i += K2_THREADS) cx.grel[t][i] = 0.f;
    for (int i = tid; i < nE; i += K2_THREADS) slot_ent[i] = -1;
    for (int i = tid; i < nR; i += K2_THREADS) slot_rel[i] = -1;
    if (tid < 8) lcg[tid] = U.lcg[tid];
    if (tid < 2) counters[tid] = 0;
    __syncthreads();

    SamplerView sv;
    sv.by_head = P.by_head + (size_t)U.tri_off * 3;
    sv.by_tail = P.by_tail ? P.by_tail + (size_t)U.tri_off * 3 : nullptr;
    sv.left_mean = P.left_mean ? P.left_mean + U.rel_off : nullptr;
    sv.right_mean = P.right_mean ? P.right_mean + U.rel_off : nullptr;
    sv.n_tri = U.n_tri;
    sv.n_ent = nE;
    sv.n_rel = nR;

    Hyper hp;
    hp.d = d; hp.k = k; hp.p_norm = P.p_norm; hp.norm_flag = P.norm_flag;
    hp.margin = U.margin;
    hp.inv_bk = 1.f / (float)((long long)B * k);
    const long long steps = (long long)U.epochs * U.nbatches;
    const int nids = B * (1 + k);

    for (long long step = 0; step < steps; ++step) {
        // ---- phase 0: the reference sampling() call, bit-exact, one thread per positive
        for (int b = tid; b < B; b += K2_THREADS) {
            int64_t j;
            const int id = stream_of(B, P.W, b, j);
            sample_one(sv, lcg[id], j, B, k, P.bern != 0, P.filter != 0, b, bh, bt, br);
        }
        __syncthreads();
        // ---- phase 1a: give every touched table row a scratch slot
        for (int i = tid; i < nids; i += K2_THREADS) {
            const int a = bh[i], c = bt[i];
            if (atomicCAS(&slot_ent[a], -1, -2) == -1) { const int s = atomicAdd(&counters[0], 1); touched_ent[s] = a; slot_ent[a] = s; }
            if (atomicCAS(&slot_ent[c], -1, -2) == -1) { const int s = atomicAdd(&counters[0], 1); touched_ent[s] = c; slot_ent[c] = s; }
            if (i < B) {
                const int r = br[i];
                if (atomicCAS(&slot_rel[r], -1, -2) == -1) { const int s = atomicAdd(&counters[1], 1); touched_rel[s] = r; slot_rel[r] = s; }
            }
        }
        if (tid < P.W) lcg[tid] = lcg_advance_batch(lcg[tid], P.W, tid, B, k);
        __syncthreads();
        // ---- phase 1b: forward + analytic backward, gradients accumulated per touched row
        for (int base = 0; base < B; base += NG) {
            const int b = base + grp;
            const bool act = b < B;
            const float l = process_sample<MODEL, L>(cx, hp, lane, B, b, act, bh, bt, br);
            if (act && lane == 0) lossv[b] = l;
        }
        __syncthreads();
        // ---- phase 2: optimizer on the touched rows, scratch back to zero
        const int nte = counters[0], ntr = counters[1];
        for (int s = grp; s < nte + ntr; s += NG) {
            const bool is_ent = s < nte;
            const int slot = is_ent ? s : s - nte;
            const int id = is_ent ? touched_ent[slot] : touched_rel[slot];
            const int nt = is_ent ? ntE : ntR;
            for (int t = 0; t < nt; ++t) {
                float* grow = (is_ent ? cx.gent[t] : cx.grel[t]) + (size_t)slot * d;
                float g[L::NF];
                ld_row<L>(grow, d, lane, g);
                float* xrow = (is_ent ? cx.ent[t] : cx.rel[t]) + (size_t)id * d;
                float* srow = P.opt == PK_ADAGRAD ? (is_ent ? st_ent[t] : st_rel[t]) + (size_t)id * d : nullptr;
                apply_update<L>(xrow, srow, g, d, lane, P.opt, U.lr);
#pragma unroll
                for (int i = 0; i < L::NF; ++i) g[i] = 0.f;
                st_row<L>(grow, d, lane, g);
            }
            if (lane == 0) { if (is_ent) slot_ent[id] = -1; else slot_rel[id] = -1; }
        }
        if (tid < 32 && U.loss_off >= 0) {  // deterministic loss reduction: mean + margin (MarginLoss.py:28)
            float acc = 0.f;
            for (int b = tid; b < B; b += 32) acc += lossv[b];
            acc = gsum<32>(acc);
            if (tid == 0) P.loss[U.loss_off + step] = acc / (float)((long long)B * k) + U.margin;
        }
        __syncthreads();
        if (tid < 2) counters[tid] = 0;
        // (next phase 0 does not touch counters; phase 1a runs after the next barrier)
    }

    // ---- write staged tables back
    if (P.stage) {
        __syncthreads();
        for (int t = 0; t < ntE; ++t)
            for (int i = tid; i < nE * d; i += K2_THREADS) g_ent[t][i] = cx.ent[t][i];
        for (int t = 0; t < ntR; ++t)
            for (int i = tid; i < nR * d; i += K2_THREADS) g_rel[t][i] = cx.rel[t][i];
    }
}

#endif  // PK_MODEL_TU

// ---- dispatch over (model, layout)
struct LaySel { int V, G, CPL; };

inline LaySel pick_layout(int model, int d) {
    const int V = d % 4 == 0 ? 4 : (d % 2 == 0 ? 2 : 1);
    const int chunks = d / V;
    const int nf_cap = model == TRANSD ? 4 : 8;  // registers per row per lane
    for (int G : {8, 32}) {
        int cpl = (chunks + G - 1) / G;
        int c2 = 1;
        while (c2 < cpl) c2 *= 2;
        if (c2 * V <= nf_cap || G == 32) return LaySel{V, G, std::max(c2, 1)};
    }
    return LaySel{V, 32, 1};
}

#ifdef PK_MODEL_TU
template <int MODEL, int V, int G, int CPL>
int launch_k2(const K2Params& P, int n, size_t smem, cudaStream_t st) {
    auto kern = k2_train_universes<MODEL, Lay<V, G, CPL>>;
    PK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n, K2_THREADS, smem, st>>>(P);
    PK_LAUNCHED("k2_train_universes");
    return PK_OK;
}

template <int MODEL>
int dispatch_layout(const LaySel& l, const K2Params& P, int n, size_t smem, cudaStream_t st) {
#define PK_CASE(v, g, c) if (l.V == v && l.G == g && l.CPL == c) return launch_k2<MODEL, v, g, c>(P, n, smem, st);
    PK_CASE(4, 8, 1) PK_CASE(4, 8, 2) PK_CASE(4, 32, 1) PK_CASE(4, 32, 2)
    PK_CASE(2, 8, 1) PK_CASE(2, 8, 2) PK_CASE(2, 8, 4) PK_CASE(2, 32, 1) PK_CASE(2, 32, 2) PK_CASE(2, 32, 4)
    PK_CASE(1, 8, 1) PK_CASE(1, 8, 2) PK_CASE(1, 8, 4) PK_CASE(1, 8, 8) PK_CASE(1, 32, 1) PK_CASE(1, 32, 2) PK_CASE(1, 32, 4) PK_CASE(1, 32, 8)
#undef PK_CASE
    return pk::fail(PK_ERR_UNSUPPORTED, "embedding dimension not supported by the universe kernel (d <= 256)");
}

// one translation unit per model keeps the build parallel: -DPK_MODEL_TU=0|1|2
#define PK_CAT2(a, b) a##b
#define PK_CAT(a, b) PK_CAT2(a, b)
int PK_CAT(launch_model, PK_MODEL_TU)(const LaySel& l, const K2Params& P, int n, size_t smem, cudaStream_t st) {
    return dispatch_layout<PK_MODEL_TU>(l, P, n, smem, st);
}
}  // namespace pkk2
#else
int launch_model0(const LaySel& l, const K2Params& P, int n, size_t smem, cudaStream_t st);
int launch_model1(const LaySel& l, const K2Params& P, int n, size_t smem, cudaStream_t st);
int launch_model2(const LaySel& l, const K2Params& P, int n, size_t smem, cudaStream_t st);

struct DescBuf {  // device copy of the descriptors, grown on demand, one per thread
    pk_universe_desc* d = nullptr;
    size_t cap = 0;
};
thread_local DescBuf g_desc[2];

}  // namespace pkk2

using namespace pkk2;

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
    cudaStream_t st = (cudaStream_t)stream;
    const int d = cfg->dim, k = cfg->neg_ent;

    int dev = 0, max_smem = 0;
    PK_CUDA(cudaGetDevice(&dev));
    PK_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    // split into a staged launch (tables fit in shared memory) and an unstaged one
    std::vector<pk_universe_desc> cls[2];
    int mE[2] = {1, 1}, mR[2] = {1, 1}, mB[2] = {1, 1};
    for (int i = 0; i < n; ++i) {
        const pk_universe_desc& u = h_desc[i];
        if (u.n_ent < 2 || u.n_rel < 1 || u.n_tri < 1 || u.batch_size < 1 || u.nbatches < 0 || u.epochs < 0)
            return pk::fail(PK_ERR_ARG, "pk_train_universes: degenerate universe descriptor");
        K2Smem own(cfg->model, d, k, u.n_ent, u.n_rel, u.batch_size, 1);
        const int c = own.total <= (size_t)max_smem ? 0 : 1;
        cls[c].push_back(u);
        mE[c] = std::max(mE[c], u.n_ent);
        mR[c] = std::max(mR[c], u.n_rel);
        mB[c] = std::max(mB[c], u.batch_size);
    }
    // the uniform carve-up uses the class maxima; if that overflows, demote the largest universes
    for (;;) {
        if (cls[0].empty()) break;
        K2Smem s(cfg->model, d, k, mE[0], mR[0], mB[0], 1);
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
    for (int c = 0; c < 2; ++c) {
        if (cls[c].empty()) continue;
        const size_t bytes = cls[c].size() * sizeof(pk_universe_desc);
        if (g_desc[c].cap < bytes) {
            if (g_desc[c].d) cudaFree(g_desc[c].d);
            PK_CUDA(cudaMalloc(&g_desc[c].d, bytes));
            g_desc[c].cap = bytes;
        }
        // the longest universes first: blocks are scheduled in index order
        std::stable_sort(cls[c].begin(), cls[c].end(), [](const pk_universe_desc& a, const pk_universe_desc& b) {
            return (long long)a.epochs * a.nbatches * a.batch_size > (long long)b.epochs * b.nbatches * b.batch_size;
        });
        PK_CUDA(cudaMemcpyAsync(g_desc[c].d, cls[c].data(), bytes, cudaMemcpyHostToDevice, st));
        K2Params P;
        P.desc = g_desc[c].d;
        for (int i = 0; i < 2; ++i) {
            P.ent[i] = packed->ent[i]; P.rel[i] = packed->rel[i];
            P.ent_state[i] = packed->ent_state[i]; P.rel_state[i] = packed->rel_state[i];
        }
        P.by_head = d_by_head; P.by_tail = d_by_tail; P.left_mean = d_left_mean; P.right_mean = d_right_mean;
        P.loss = d_loss;
        P.d = d; P.k = k; P.p_norm = cfg->p_norm; P.norm_flag = cfg->norm_flag; P.opt = cfg->opt;
        P.bern = cfg->bern; P.filter = cfg->filter; P.W = cfg->work_threads;
        P.stage = c == 0;
        P.mE = mE[c]; P.mR = mR[c]; P.mB = mB[c];
        K2Smem s(cfg->model, d, k, mE[c], mR[c], mB[c], P.stage);
        if (s.total > (size_t)max_smem)
            return pk::fail(PK_ERR_UNSUPPORTED, "pk_train_universes: a universe's batch scratch exceeds shared memory; use pk_train_steps");
        // descriptors were copied from pageable host memory owned by this call: the copy has
        // completed (or been staged) when cudaMemcpyAsync returns, so cls[c] may go out of scope
        int rc = PK_OK;
        const int nblk = (int)cls[c].size();
        if (cfg->model == PK_TRANSE) rc = launch_model0(lay, P, nblk, s.total, st);
        else if (cfg->model == PK_TRANSH) rc = launch_model1(lay, P, nblk, s.total, st);
        else rc = launch_model2(lay, P, nblk, s.total, st);
        if (rc != PK_OK) return rc;
    }
    return PK_OK;
}
#endif  // !PK_MODEL_TU
