// The reference draws every universe's hyper-parameters from Python's `random` module after random.seed(seed)
// (openke/config/Parallel_Universe_Config.py:157-161,211-212,232,237-240).  Per universe that is a 624-word Mersenne
// Twister seeding plus five draws — 8-17 us of interpreter time on the launching thread, 1-2 ms per chunk of 100
// universes with the GPU waiting.  This is CPython's generator restated (Modules/_randommodule.c: init_by_array seeding
// of an int, genrand_uint32; Lib/random.py: _randbelow_with_getrandbits, uniform; Objects/floatobject.c: round(x, n) as
// a correctly rounded decimal conversion and back), bit-identical for the value ranges checked below; anything else is
// refused so that the caller falls back to the interpreter (tests/test_host.py::test_native_hyper_draws_equal_python_random).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.hpp"

namespace {

struct PyRandom {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t* key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= len) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    void seed_int(int64_t a) {   // random.seed(int): the absolute value as 32-bit words, least significant first
        uint64_t n = a < 0 ? (uint64_t)0 - (uint64_t)a : (uint64_t)a;
        uint32_t key[2] = {(uint32_t)n, (uint32_t)(n >> 32)};
        init_by_array(key, key[1] ? 2 : 1);
    }
    uint32_t next32() {
        if (idx >= 624) {
            int k;
            for (k = 0; k < 624 - 397; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
                mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; k < 623; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
                mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double random() {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    int64_t randbelow(uint32_t n) {   // _randbelow_with_getrandbits, n in [1, 2^32)
        int k = 0;
        for (uint32_t v = n; v; v >>= 1) ++k;
        uint32_t r;
        do r = next32() >> (32 - k); while (r >= n);
        return (int64_t)r;
    }
};

double py_round(double x, int ndigits) {   // float.__round__(ndigits): shortest way to say "correctly rounded, ties to even"
    char buf[400];
    snprintf(buf, sizeof buf, "%.*f", ndigits, x);
    return strtod(buf, nullptr);
}

bool width_ok(int64_t lo, int64_t hi) { return hi > lo && hi - lo < (int64_t)0xffffffffLL; }

}  // namespace

extern "C" int pk_python_hyper_draws(int n, const int64_t* seeds, int64_t tc_lo, int64_t tc_hi, double bal_lo, double bal_hi,
                                     int64_t margin_lo, int64_t margin_hi, int64_t epochs_lo, int64_t epochs_hi, int draw_epochs,
                                     double lr_lo, double lr_hi, int lr_digits, int64_t* tc, double* balance, int64_t* margin,
                                     int64_t* epochs, double* lr) {
    if (n < 0 || !seeds || !tc || !balance || !margin || !epochs || !lr) return pk::fail(PK_ERR_ARG, "pk_python_hyper_draws: null argument");
    if (!width_ok(tc_lo, tc_hi) || !width_ok(margin_lo, margin_hi) || (draw_epochs && !width_ok(epochs_lo, epochs_hi)) ||
        lr_digits < 0 || lr_digits > 22 || !std::isfinite(bal_lo) || !std::isfinite(bal_hi) || !std::isfinite(lr_lo) || !std::isfinite(lr_hi))
        return pk::fail(PK_ERR_UNSUPPORTED, "pk_python_hyper_draws: range outside what the native generator covers (use random.Random)");
    PyRandom g;
    for (int i = 0; i < n; ++i) {
        g.seed_int(seeds[i]);
        tc[i] = tc_lo + g.randbelow((uint32_t)(tc_hi - tc_lo));
        balance[i] = py_round(bal_lo + (bal_hi - bal_lo) * g.random(), 2);
        margin[i] = margin_lo + g.randbelow((uint32_t)(margin_hi - margin_lo));
        epochs[i] = draw_epochs ? epochs_lo + g.randbelow((uint32_t)(epochs_hi - epochs_lo)) : epochs_lo;
        lr[i] = py_round(lr_lo + (lr_hi - lr_lo) * g.random(), lr_digits);
    }
    return PK_OK;
}
