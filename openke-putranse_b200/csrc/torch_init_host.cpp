// Initial embedding tables of many universes on host threads, bit-identical to what the
// reference's model constructors leave behind after torch.manual_seed(seed):
//   nn.Embedding(rows, dim)            -> weight.normal_()        (torch's default init; discarded)
//   nn.init.xavier_uniform_(weight)    -> weight.uniform_(-a, a)  (reference TransE.py:17-22 and twins)
// with every table of a space drawing from ONE mt19937 stream in creation order.  The torch CPU
// generator is a plain MT19937 seeded with the low 32 bits of the seed; normal_() on n >= 16
// contiguous floats consumes n draws (+16 when n % 16 != 0: the tail block is redrawn), and
// uniform_() maps one 32-bit draw per element: (draw & (2^24-1)) * 2^-24 * (to - from) + from in
// float arithmetic.  The Python layer verifies this replay against torch itself before trusting it
// (openke/config/Parallel_Universe_Config.py: _native_init_ok) and falls back to torch otherwise.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "common.hpp"

namespace {

struct Mt19937 {
    uint32_t s[624];
    int idx;
    explicit Mt19937(uint64_t seed) {
        s[0] = (uint32_t)(seed & 0xffffffffULL);
        for (int j = 1; j < 624; ++j) s[j] = 1812433253U * (s[j - 1] ^ (s[j - 1] >> 30)) + (uint32_t)j;
        idx = 624;
    }
    void refill() {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (s[k] & 0x80000000U) | (s[(k + 1) % 624] & 0x7fffffffU);
            s[k] = s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
        }
        idx = 0;
    }
    uint32_t next() {
        if (idx >= 624) refill();
        uint32_t y = s[idx++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    }
    void skip(int64_t n) {
        while (n > 0) {
            if (idx >= 624) refill();
            const int64_t take = std::min<int64_t>(n, 624 - idx);
            idx += (int)take;
            n -= take;
        }
    }
};

}  // namespace

extern "C" int pk_torch_init_tables(int n, const int64_t* seeds, int n_tables, const int64_t* rows, const int32_t* dims,
                                    float* const* out, const int64_t* row_off, const double* bounds, int fused, int nthreads) {
    if (n < 0 || n_tables < 1 || n_tables > 8 || !seeds || !rows || !dims || !out || !row_off || !bounds)
        return pk::fail(PK_ERR_ARG, "pk_torch_init_tables: bad argument");
    for (int64_t i = 0; i < (int64_t)n * n_tables; ++i)
        if (rows[i] * dims[i % n_tables] < 16)
            return pk::fail(PK_ERR_UNSUPPORTED, "pk_torch_init_tables: a table has fewer than 16 elements (torch takes another path there)");
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = std::min(nthreads, std::max(n, 1));
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            Mt19937 mt((uint64_t)seeds[i]);
            for (int t = 0; t < n_tables; ++t) {
                const int64_t sz = rows[(int64_t)i * n_tables + t] * dims[t];
                mt.skip(sz + (sz % 16 ? 16 : 0));
            }
            for (int t = 0; t < n_tables; ++t) {
                const int64_t sz = rows[(int64_t)i * n_tables + t] * dims[t];
                float* p = out[t] + row_off[(int64_t)i * n_tables + t] * dims[t];
                const float to = (float)bounds[(int64_t)i * n_tables + t], from = (float)(-bounds[(int64_t)i * n_tables + t]);
                const float span = to - from;
                if (fused) {
                    for (int64_t e = 0; e < sz; ++e) p[e] = std::fmaf((float)(mt.next() & 0xffffffU) * (1.0f / 16777216.0f), span, from);
                } else {
                    for (int64_t e = 0; e < sz; ++e) {
                        volatile float prod = (float)(mt.next() & 0xffffffU) * (1.0f / 16777216.0f) * span;
                        p[e] = prod + from;
                    }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return PK_OK;
}
