// Device-backed members of the Base.so-compatible surface: sampling() and the per-row rank
// functions testHead/testTail/validHead/validTail.  They keep the reference's signatures (HOST
// buffers, process-global state) but do the work with the same CUDA kernels as the pk_* entry
// points: K0 for sampling (openke/base/Base.cpp:266-310) and the candidate-row ranking kernel for
// openke/base/Test.h:118-359 / openke/base/Valid.h:116-240.  There is no host implementation to
// fall back to; without a CUDA device these functions report the failure and leave buffers
// untouched.
#include <cstring>
#include <vector>

#include "common.hpp"
#include "global_state.hpp"

namespace {

struct DevIndex {  // device mirror of the sampler's current id space
    uint64_t epoch = 0;
    int32_t* by_head = nullptr;
    int32_t* by_tail = nullptr;
    float* left_mean = nullptr;
    float* right_mean = nullptr;
    uint64_t* lcg = nullptr;
    int32_t* ids = nullptr;
    size_t ids_cap = 0;
    std::vector<int32_t> host_ids;
};
DevIndex g_dev;

struct DevFilter {  // device mirror of the filter CSR of one (which, side)
    bool ready = false;
    uint64_t epoch = 0;  // pk::Global::eval_epoch it was built from
    std::vector<int64_t> off;
    int32_t* cand = nullptr;
};
DevFilter g_filter[2][2];
float* g_con = nullptr;
size_t g_con_cap = 0;
int32_t* g_small = nullptr;  // truth[1] | ranks[2] | pad ; then int64 foff[2]

// The reference's float accumulators (openke/base/Test.h:14-16, Valid.h:33-34).
struct Acc {
    float tot10 = 0, tot3 = 0, tot1 = 0, rank = 0, reci = 0;
    float f_tot10 = 0, f_tot3 = 0, f_tot1 = 0, f_rank = 0, f_reci = 0;
};
Acc g_l, g_r;
float g_valid_l10 = 0, g_valid_r10 = 0;
float g_mrr = 0, g_mr = 0, g_hit10 = 0, g_hit3 = 0, g_hit1 = 0;

int upload_index() {
    const pk::TripleIndex& ix = pk::current_index();
    if (ix.n_tri() == 0) return pk::fail(PK_ERR_STATE, "sampling: importTrainFiles has not run");
    if (g_dev.epoch == pk::index_epoch()) return PK_OK;
    cudaFree(g_dev.by_head); cudaFree(g_dev.by_tail); cudaFree(g_dev.left_mean); cudaFree(g_dev.right_mean);
    g_dev.by_head = g_dev.by_tail = nullptr;
    g_dev.left_mean = g_dev.right_mean = nullptr;
    const size_t tb = (size_t)ix.n_tri() * sizeof(pk::Tri), rb = (size_t)ix.n_rel * sizeof(float);
    PK_CUDA(cudaMalloc(&g_dev.by_head, tb));
    PK_CUDA(cudaMalloc(&g_dev.by_tail, tb));
    PK_CUDA(cudaMalloc(&g_dev.left_mean, rb));
    PK_CUDA(cudaMalloc(&g_dev.right_mean, rb));
    if (!g_dev.lcg) PK_CUDA(cudaMalloc(&g_dev.lcg, 64 * 8));
    PK_CUDA(cudaMemcpy(g_dev.by_head, ix.by_head.data(), tb, cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(g_dev.by_tail, ix.by_tail.data(), tb, cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(g_dev.left_mean, ix.left_mean.data(), rb, cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(g_dev.right_mean, ix.right_mean.data(), rb, cudaMemcpyHostToDevice));
    g_dev.epoch = pk::index_epoch();
    return PK_OK;
}

int do_sampling(PK_INT* bh, PK_INT* bt, PK_INT* br, PK_REAL* by, PK_INT B, PK_INT k, bool filter) {
    int rc = upload_index();
    if (rc != PK_OK) return rc;
    pk::Global& g = pk::G();
    const pk::TripleIndex& ix = pk::current_index();
    const size_t n = (size_t)B * (size_t)(1 + k);
    if (g_dev.ids_cap < 3 * n) {
        cudaFree(g_dev.ids);
        g_dev.ids = nullptr;
        PK_CUDA(cudaMalloc(&g_dev.ids, 3 * n * 4));
        g_dev.ids_cap = 3 * n;
    }
    g_dev.host_ids.resize(3 * n);
    const int W = (int)g.graph.work_threads;
    PK_CUDA(cudaMemcpy(g_dev.lcg, g.lcg, (size_t)W * 8, cudaMemcpyHostToDevice));
    pk_model_cfg cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.neg_ent = (int32_t)k; cfg.bern = g.graph.bern ? 1 : 0; cfg.filter = filter ? 1 : 0; cfg.work_threads = W;
    pk_sampler smp;
    smp.by_head = g_dev.by_head; smp.by_tail = g_dev.by_tail; smp.left_mean = g_dev.left_mean; smp.right_mean = g_dev.right_mean;
    smp.lcg = g_dev.lcg; smp.n_tri = ix.n_tri(); smp.n_ent = ix.n_ent; smp.n_rel = ix.n_rel;
    smp.head_off = nullptr; smp.tail_off = nullptr;
    rc = pk_sample_batch(&cfg, &smp, B, g_dev.ids, g_dev.ids + n, g_dev.ids + 2 * n, nullptr);
    if (rc != PK_OK) return rc;
    PK_CUDA(cudaMemcpy(g_dev.host_ids.data(), g_dev.ids, 3 * n * 4, cudaMemcpyDeviceToHost));
    PK_CUDA(cudaMemcpy(g.lcg, g_dev.lcg, (size_t)W * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        bh[i] = g_dev.host_ids[i];
        bt[i] = g_dev.host_ids[n + i];
        br[i] = g_dev.host_ids[2 * n + i];
        if (by) by[i] = i < (size_t)B ? 1.f : -1.f;
    }
    return PK_OK;
}

// rank one candidate-ordered host row on the device
int rank_row(const PK_REAL* con, int which, int side, PK_INT index, int* raw, int* filt) {
    pk::Global& g = pk::G();
    const std::vector<pk::Tri>& q = which == 0 ? g.graph.test : g.graph.valid;
    if (index < 0 || (size_t)index >= q.size()) return pk::fail(PK_ERR_ARG, "rank: test index out of range");
    DevFilter& f = g_filter[which][side];
    if (!f.ready || f.epoch != g.eval_epoch) {   // a later importTestFiles replaced the lists
        if (f.cand) cudaFree(f.cand);
        f.cand = nullptr;
        f.ready = false;
        std::vector<int32_t> cand;
        g.graph.filter_candidates(which, side, f.off, cand);
        PK_CUDA(cudaMalloc(&f.cand, std::max<size_t>(cand.size(), 1) * 4));
        PK_CUDA(cudaMemcpy(f.cand, cand.data(), cand.size() * 4, cudaMemcpyHostToDevice));
        f.ready = true;
        f.epoch = g.eval_epoch;
    }
    const int64_t E = g.graph.n_ent;
    if (g_con_cap < (size_t)E) {
        cudaFree(g_con);
        g_con = nullptr;
        PK_CUDA(cudaMalloc(&g_con, (size_t)E * 4));
        g_con_cap = (size_t)E;
    }
    if (!g_small) PK_CUDA(cudaMalloc(&g_small, 64));
    const pk::Tri& x = q[(size_t)index];
    struct { int32_t truth, r0, r1, pad; int64_t foff[2]; } h;
    h.truth = side == 0 ? x.h : x.t;
    h.r0 = h.r1 = h.pad = 0;
    h.foff[0] = 0;
    h.foff[1] = f.off[(size_t)index + 1] - f.off[(size_t)index];
    PK_CUDA(cudaMemcpy(g_small, &h, sizeof h, cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(g_con, con, (size_t)E * 4, cudaMemcpyHostToDevice));
    int rc = pk_rank_candidate_row(g_con, E, g_small, reinterpret_cast<const int64_t*>(g_small + 4), f.cand + f.off[(size_t)index],
                                   g_small + 1, nullptr);
    if (rc != PK_OK) return rc;
    int32_t out[2];
    PK_CUDA(cudaMemcpy(out, g_small + 1, 8, cudaMemcpyDeviceToHost));
    *raw = out[0];
    *filt = out[1];
    return PK_OK;
}

void accumulate(Acc& a, int raw, int filt) {  // Test.h:213-223, float accumulators like the reference
    if (filt < 10) a.f_tot10 += 1;
    if (raw < 10) a.tot10 += 1;
    if (filt < 3) a.f_tot3 += 1;
    if (raw < 3) a.tot3 += 1;
    if (filt < 1) a.f_tot1 += 1;
    if (raw < 1) a.tot1 += 1;
    a.f_rank += (filt + 1);
    a.rank += (1 + raw);
    a.f_reci += 1.0 / (filt + 1);
    a.reci += 1.0 / (raw + 1);
}

}  // namespace


// ---------------------------------------------------------------------------------------------
// Triple-classification negatives (openke/base/Test.h:573-599).  The reference walks the test list
// once on LCG stream 0: per triple one draw for the coin (randd(0) % 1000 < 500 keeps the head) and
// one inside the FILTERED corruptor (Corrupt.h:9-105, filter_flag defaults to true).  Two draws per
// triple, always, so thread i starts 2i draws into the stream.
//
// The corruptor here restates the reference's searches with their exact sentinels instead of reusing
// the sampler's corrupt_entity(): a test pair (entity, relation) need not occur in the training
// index, and then the reference's two boundary searches do NOT produce an empty run — they settle on
// a neighbouring record of the entity — so what is excluded depends on those sentinels.
// An entity with no record on that side makes the reference read trainHead[-1] (undefined); it is
// given an empty exclusion set here.
__device__ int64_t neg_test_corrupt(uint64_t x, const int32_t* __restrict__ idx, const int64_t* __restrict__ lefs,
                                    const int64_t* __restrict__ rigs, int64_t n_ent, int32_t fix, int32_t r, int col) {
    const int64_t L = lefs[fix], R = rigs[fix];
    if (R < 0) return (int64_t)(x % (uint64_t)n_ent);
    int64_t lo = L - 1, hi = R;
    while (lo + 1 < hi) {                       // Corrupt.h:27-33 / :77-83
        const int64_t mid = (lo + hi) >> 1;
        if (idx[mid * 3 + 1] >= r) hi = mid; else lo = mid;
    }
    const int64_t ll = hi;
    lo = L; hi = R + 1;
    while (lo + 1 < hi) {                       // Corrupt.h:35-42 / :85-92
        const int64_t mid = (lo + hi) >> 1;
        if (idx[mid * 3 + 1] <= r) lo = mid; else hi = mid;
    }
    const int64_t rr = lo;
    const int64_t tmp = (int64_t)(x % (uint64_t)(n_ent - (rr - ll + 1)));
    if (tmp < idx[ll * 3 + col]) return tmp;
    if (tmp > idx[rr * 3 + col] - rr + ll - 1) return tmp + rr - ll + 1;
    lo = ll; hi = rr + 1;
    while (lo + 1 < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (idx[mid * 3 + col] - mid + ll - 1 < tmp) lo = mid; else hi = mid;
    }
    return tmp + lo - ll + 1;
}

__global__ void k_neg_test(const int32_t* __restrict__ test, int64_t n, const int32_t* __restrict__ by_head,
                           const int32_t* __restrict__ by_tail, const int64_t* __restrict__ ranges /*[4][E]*/, int64_t n_ent,
                           const uint64_t* __restrict__ lcg, int32_t* __restrict__ neg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = lcg[0];
    {   // 2i draws ahead: x -> A x + C by squaring
        uint64_t a = 25214903917ULL, c = 11ULL, ra = 1, rc = 0, m = 2ULL * (uint64_t)i;
        while (m) {
            if (m & 1) { ra = ra * a; rc = rc * a + c; }
            c = (a + 1) * c;
            a = a * a;
            m >>= 1;
        }
        s = ra * s + rc;
    }
    const int32_t h = test[i * 3 + 0], r = test[i * 3 + 1], t = test[i * 3 + 2];
    s = s * 25214903917ULL + 11ULL;
    const bool keep_head = (s % 1000ULL) < 500ULL;
    s = s * 25214903917ULL + 11ULL;
    int32_t nh = h, nt = t;
    if (keep_head) nt = (int32_t)neg_test_corrupt(s, by_head, ranges, ranges + n_ent, n_ent, h, r, 2);
    else nh = (int32_t)neg_test_corrupt(s, by_tail, ranges + 2 * n_ent, ranges + 3 * n_ent, n_ent, t, r, 0);
    neg[i * 3 + 0] = nh; neg[i * 3 + 1] = r; neg[i * 3 + 2] = nt;
}

namespace {
std::vector<int32_t> g_neg_test;   // (h, r, t) per test triple

int do_neg_test() {
    pk::Global& g = pk::G();
    const size_t n = g.graph.test.size();
    if (n == 0) return pk::fail(PK_ERR_STATE, "getNegTest: importTestFiles has not run");
    int rc = upload_index();
    if (rc != PK_OK) return rc;
    const pk::TripleIndex& ix = pk::current_index();
    const int64_t E = ix.n_ent;
    std::vector<int64_t> ranges((size_t)(4 * E));
    std::memcpy(ranges.data(), ix.lef_head.data(), (size_t)E * 8);
    std::memcpy(ranges.data() + E, ix.rig_head.data(), (size_t)E * 8);
    std::memcpy(ranges.data() + 2 * E, ix.lef_tail.data(), (size_t)E * 8);
    std::memcpy(ranges.data() + 3 * E, ix.rig_tail.data(), (size_t)E * 8);
    int64_t* d_ranges = nullptr;
    int32_t* d_test = nullptr;
    int32_t* d_neg = nullptr;
    auto release = [&]() { cudaFree(d_ranges); cudaFree(d_test); cudaFree(d_neg); };
    cudaError_t e = cudaMalloc(&d_ranges, ranges.size() * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_test, n * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_neg, n * 12);
    if (e == cudaSuccess) e = cudaMemcpy(d_ranges, ranges.data(), ranges.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_test, g.graph.test.data(), n * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(g_dev.lcg, g.lcg, 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { release(); return pk::cuda_fail(e, "getNegTest: staging"); }
    k_neg_test<<<(unsigned)((n + 255) / 256), 256>>>(d_test, (int64_t)n, g_dev.by_head, g_dev.by_tail, d_ranges, E, g_dev.lcg, d_neg);
    pk::launch_counter() = 1;
    g_neg_test.resize(n * 3);
    e = cudaMemcpy(g_neg_test.data(), d_neg, n * 12, cudaMemcpyDeviceToHost);
    release();
    if (e != cudaSuccess) return pk::cuda_fail(e, "getNegTest: kernel");
    // stream 0 has consumed two draws per test triple
    uint64_t a = 25214903917ULL, c = 11ULL, ra = 1, rc2 = 0, m = 2ULL * n;
    while (m) {
        if (m & 1) { ra = ra * a; rc2 = rc2 * a + c; }
        c = (a + 1) * c;
        a = a * a;
        m >>= 1;
    }
    g.lcg[0] = ra * g.lcg[0] + rc2;
    return PK_OK;
}
}  // namespace

namespace pk {
void test_metrics_reset() { g_l = Acc(); g_r = Acc(); }
void valid_metrics_reset() { g_valid_l10 = g_valid_r10 = 0; }
}  // namespace pk

extern "C" {

int pk_cuda_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return pk::cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

void sampling(PK_INT* batch_h, PK_INT* batch_t, PK_INT* batch_r, PK_REAL* batch_y, PK_INT batchSize, PK_INT negRate,
              PK_INT negRelRate, PK_INT mode, bool filter_flag, bool p, bool val_loss) {
    (void)p;
    if (negRelRate != 0 || mode != 0 || val_loss) {
        pk::fail(PK_ERR_UNSUPPORTED, "sampling: only mode 0 with negRelRate 0 is on the PuTransE hot path");
        fprintf(stderr, "putranse: %s\n", pk::last_error().c_str());
        return;
    }
    if (do_sampling(batch_h, batch_t, batch_r, batch_y, batchSize, negRate, filter_flag) != PK_OK)
        fprintf(stderr, "putranse: sampling failed: %s\n", pk::last_error().c_str());
}

void testHead(PK_REAL* con, PK_INT index, bool) {
    int raw, filt;
    if (rank_row(con, 0, 0, index, &raw, &filt) == PK_OK) accumulate(g_l, raw, filt);
    else fprintf(stderr, "putranse: testHead failed: %s\n", pk::last_error().c_str());
}
void testTail(PK_REAL* con, PK_INT index, bool) {
    int raw, filt;
    if (rank_row(con, 0, 1, index, &raw, &filt) == PK_OK) accumulate(g_r, raw, filt);
    else fprintf(stderr, "putranse: testTail failed: %s\n", pk::last_error().c_str());
}
void validHead(PK_REAL* con, PK_INT index) {
    int raw, filt;
    if (rank_row(con, 1, 0, index, &raw, &filt) == PK_OK) { if (filt < 10) g_valid_l10 += 1; }
    else fprintf(stderr, "putranse: validHead failed: %s\n", pk::last_error().c_str());
}
void validTail(PK_REAL* con, PK_INT index) {
    int raw, filt;
    if (rank_row(con, 1, 1, index, &raw, &filt) == PK_OK) { if (filt < 10) g_valid_r10 += 1; }
    else fprintf(stderr, "putranse: validTail failed: %s\n", pk::last_error().c_str());
}

void getNegTest(void) {   // Test.h:573-586
    if (do_neg_test() != PK_OK) fprintf(stderr, "putranse: getNegTest failed: %s\n", pk::last_error().c_str());
}
void getTestBatch(PK_INT* ph, PK_INT* pt, PK_INT* pr, PK_INT* nh, PK_INT* nt, PK_INT* nr) {   // Test.h:588-599
    if (do_neg_test() != PK_OK) {
        fprintf(stderr, "putranse: getTestBatch failed: %s\n", pk::last_error().c_str());
        return;
    }
    const std::vector<pk::Tri>& q = pk::G().graph.test;
    for (size_t i = 0; i < q.size(); ++i) {
        ph[i] = q[i].h; pt[i] = q[i].t; pr[i] = q[i].r;
        nh[i] = g_neg_test[i * 3]; nr[i] = g_neg_test[i * 3 + 1]; nt[i] = g_neg_test[i * 3 + 2];
    }
}

void test_link_prediction(bool) {  // Test.h:398-454 (no table printing)
    const float n = (float)pk::G().graph.test.size();
    Acc l = g_l, r = g_r;
    for (Acc* a : {&l, &r}) {
        a->rank /= n; a->reci /= n; a->tot10 /= n; a->tot3 /= n; a->tot1 /= n;
        a->f_rank /= n; a->f_reci /= n; a->f_tot10 /= n; a->f_tot3 /= n; a->f_tot1 /= n;
    }
    g_mrr = (l.f_reci + r.f_reci) / 2;
    g_mr = (l.f_rank + r.f_rank) / 2;
    g_hit10 = (l.f_tot10 + r.f_tot10) / 2;
    g_hit3 = (l.f_tot3 + r.f_tot3) / 2;
    g_hit1 = (l.f_tot1 + r.f_tot1) / 2;
}
PK_REAL getTestLinkMRR(bool) { return g_mrr; }
PK_REAL getTestLinkMR(bool) { return g_mr; }
PK_REAL getTestLinkHit10(bool) { return g_hit10; }
PK_REAL getTestLinkHit3(bool) { return g_hit3; }
PK_REAL getTestLinkHit1(bool) { return g_hit1; }
PK_REAL getValidHit10(void) {  // Valid.h:242-257
    const float n = (float)pk::G().graph.valid.size();
    return (g_valid_l10 / n + g_valid_r10 / n) / 2;
}

}  // extern "C"
