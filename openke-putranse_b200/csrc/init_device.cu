// Initial embedding tables of many universes ON THE DEVICE, bit-identical to the reference's model
// constructors after torch.manual_seed(seed) (same contract as pk_torch_init_tables in
// torch_init_host.cpp, which documents how torch's CPU generator is replayed).  One thread block per
// universe owns one MT19937 state in shared memory; the 624-word twist is done cooperatively in the
// three dependency-free segments of the recurrence, outputs are tempered in parallel and written
// coalesced.  This removes the host RNG time and the H2D copy of the tables from the end-to-end path.
#include <algorithm>
#include <vector>

#include "common.hpp"

namespace {

constexpr int INIT_THREADS = 256;
constexpr int MT_N = 624, MT_M = 397;

struct InitParams {
    const int64_t* seeds;    // [n]
    const int64_t* rows;     // [n*T]
    const int64_t* row_off;  // [n*T]
    const float* bounds;     // [n*T] already rounded to float
    float* out[8];
    int dims[8];
    int T, fused;
};

__device__ __forceinline__ uint32_t twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000U) | (nxt & 0x7fffffffU);
    return far ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
}

__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= (y >> 18);
    return y;
}

__global__ void __launch_bounds__(INIT_THREADS) k_init_tables(const __grid_constant__ InitParams P) {
    __shared__ uint32_t s[MT_N];
    __shared__ long long start[9];   // first draw of table t's uniform fill; start[T] = end
    const int u = blockIdx.x, tid = threadIdx.x, T = P.T;
    if (tid == 0) {
        uint32_t x = (uint32_t)((uint64_t)P.seeds[u] & 0xffffffffULL);
        s[0] = x;
        for (int j = 1; j < MT_N; ++j) { x = 1812433253U * (x ^ (x >> 30)) + (uint32_t)j; s[j] = x; }
        long long skip = 0;
        for (int t = 0; t < T; ++t) {   // nn.Embedding's default normal_(): n draws, +16 when n % 16 != 0
            const long long sz = P.rows[(long long)u * T + t] * P.dims[t];
            skip += sz + (sz % 16 ? 16 : 0);
        }
        start[0] = skip;
        for (int t = 0; t < T; ++t) start[t + 1] = start[t] + P.rows[(long long)u * T + t] * P.dims[t];
    }
    __syncthreads();
    const long long total = start[T], first = start[0];
    for (long long base = 0; base < total; base += MT_N) {
        // twist: k in [0,227) reads only old words; [227,454) reads new words of the first segment;
        // [454,623) of the second; word 623 reads new words 0 and 396
        uint32_t v[3];
        int cnt = 0;
        for (int k = tid; k < MT_N - MT_M; k += INIT_THREADS) v[cnt++] = twist(s[k], s[k + 1], s[k + MT_M]);
        __syncthreads();
        cnt = 0;
        for (int k = tid; k < MT_N - MT_M; k += INIT_THREADS) s[k] = v[cnt++];
        __syncthreads();
        cnt = 0;
        for (int k = MT_N - MT_M + tid; k < 2 * (MT_N - MT_M); k += INIT_THREADS) v[cnt++] = twist(s[k], s[k + 1], s[k + MT_M - MT_N]);
        __syncthreads();
        cnt = 0;
        for (int k = MT_N - MT_M + tid; k < 2 * (MT_N - MT_M); k += INIT_THREADS) s[k] = v[cnt++];
        __syncthreads();
        cnt = 0;
        for (int k = 2 * (MT_N - MT_M) + tid; k < MT_N - 1; k += INIT_THREADS) v[cnt++] = twist(s[k], s[k + 1], s[k + MT_M - MT_N]);
        __syncthreads();
        cnt = 0;
        for (int k = 2 * (MT_N - MT_M) + tid; k < MT_N - 1; k += INIT_THREADS) s[k] = v[cnt++];
        __syncthreads();
        if (tid == 0) s[MT_N - 1] = twist(s[MT_N - 1], s[0], s[MT_M - 1]);
        __syncthreads();
        if (base + MT_N <= first) continue;   // still inside the discarded normal_() draws
        for (int j = tid; j < MT_N; j += INIT_THREADS) {
            const long long g = base + j;
            if (g < first || g >= total) continue;
            int t = 0;
            while (t + 1 < T && g >= start[t + 1]) ++t;
            const float to = P.bounds[(long long)u * T + t], from = -to, span = to - from;
            const float x = (float)(temper(s[j]) & 0xffffffU) * (1.0f / 16777216.0f);
            const float val = P.fused ? __fmaf_rn(x, span, from) : __fadd_rn(__fmul_rn(x, span), from);
            P.out[t][P.row_off[(long long)u * T + t] * P.dims[t] + (g - start[t])] = val;
        }
        __syncthreads();
    }
}

// Kernel arguments travel through a ring of (pinned host, device) buffer pairs: launches of different launch slots are
// in flight together and must not share one scratch, and a copy from pageable memory beyond 64 KB makes the launching
// thread wait for the stream.  A pair is reused only after the launch that read it has finished (event).
struct ArgRing {
    static constexpr int N = 8;
    void* h[N] = {};
    void* d[N] = {};
    size_t cap[N] = {};
    cudaEvent_t ev[N] = {};
    int next = 0;
};
thread_local ArgRing g_args;

}  // namespace

extern "C" int pk_init_tables_device(int n, const int64_t* seeds, int n_tables, const int64_t* rows, const int32_t* dims,
                                     float* const* d_out, const int64_t* row_off, const double* bounds, int fused, void* stream) {
    pk::launch_counter() = 0;
    if (n < 0 || n_tables < 1 || n_tables > 8 || !seeds || !rows || !dims || !d_out || !row_off || !bounds)
        return pk::fail(PK_ERR_ARG, "pk_init_tables_device: bad argument");
    if (n == 0) return PK_OK;
    const size_t nt = (size_t)n * n_tables;
    for (size_t i = 0; i < nt; ++i)
        if (rows[i] * dims[i % n_tables] < 16)
            return pk::fail(PK_ERR_UNSUPPORTED, "pk_init_tables_device: a table has fewer than 16 elements (torch takes another path there)");
    // host arguments -> one device buffer: seeds | rows | row_off | bounds(float)
    const size_t bytes = 8 * (size_t)n + 16 * nt + 4 * nt;
    const int slot = g_args.next;
    g_args.next = (g_args.next + 1) % ArgRing::N;
    if (!g_args.ev[slot]) PK_CUDA(cudaEventCreateWithFlags(&g_args.ev[slot], cudaEventDisableTiming));
    else PK_CUDA(cudaEventSynchronize(g_args.ev[slot]));
    if (g_args.cap[slot] < bytes) {   // all pairs at once: an allocation waits for the launches in flight
        const size_t want = std::max<size_t>(bytes * 2, 256 << 10);
        for (int r = 0; r < ArgRing::N; ++r) {
            if (g_args.cap[r] >= want) continue;
            if (g_args.ev[r]) PK_CUDA(cudaEventSynchronize(g_args.ev[r]));
            if (g_args.h[r]) cudaFreeHost(g_args.h[r]);
            if (g_args.d[r]) cudaFree(g_args.d[r]);
            g_args.h[r] = g_args.d[r] = nullptr;
            g_args.cap[r] = 0;
            PK_CUDA(cudaHostAlloc(&g_args.h[r], want, cudaHostAllocDefault));
            PK_CUDA(cudaMalloc(&g_args.d[r], want));
            g_args.cap[r] = want;
        }
    }
    int64_t* hs = reinterpret_cast<int64_t*>(g_args.h[slot]);
    int64_t* hr = hs + n;
    int64_t* ho = hr + nt;
    float* hb = reinterpret_cast<float*>(ho + nt);
    for (int i = 0; i < n; ++i) hs[i] = seeds[i];
    for (size_t i = 0; i < nt; ++i) { hr[i] = rows[i]; ho[i] = row_off[i]; hb[i] = (float)bounds[i]; }
    cudaStream_t st = (cudaStream_t)stream;
    PK_CUDA(cudaMemcpyAsync(g_args.d[slot], g_args.h[slot], bytes, cudaMemcpyHostToDevice, st));
    InitParams P;
    unsigned char* d = static_cast<unsigned char*>(g_args.d[slot]);
    P.seeds = reinterpret_cast<const int64_t*>(d);
    P.rows = P.seeds + n;
    P.row_off = P.rows + nt;
    P.bounds = reinterpret_cast<const float*>(P.row_off + nt);
    for (int t = 0; t < 8; ++t) { P.out[t] = t < n_tables ? d_out[t] : nullptr; P.dims[t] = t < n_tables ? dims[t] : 0; }
    P.T = n_tables;
    P.fused = fused;
    k_init_tables<<<n, INIT_THREADS, 0, st>>>(P);
    PK_LAUNCHED("k_init_tables");
    PK_CUDA(cudaEventRecord(g_args.ev[slot], st));
    return PK_OK;
}
