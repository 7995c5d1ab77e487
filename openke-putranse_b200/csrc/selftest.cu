// Device self-test of the arithmetic helpers in kge_device.cuh: the refined-reciprocal division and
// the rsqrt-based square root must round exactly like the IEEE operations the reference's PyTorch
// kernels use (x / n, sqrt(x)); the train kernels rely on that for fp32 parity.
#include "common.hpp"
#include "kge_device.cuh"

namespace {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// operand magnitudes 2^-40 .. 2^40 (the train kernels see norms, gradients and Adagrad sums in that band)
__device__ __forceinline__ float operand(uint32_t r) {
    const uint32_t mant = r & 0x7fffffu;
    const uint32_t expo = 127u - 40u + (hash32(r ^ 0x9e3779b9u) % 81u);
    return __uint_as_float((expo << 23) | mant);
}

__global__ void k_selftest_arith(uint64_t n, uint32_t seed, unsigned long long* bad) {
    unsigned long long bd = 0, bs = 0, bz = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r0 = hash32((uint32_t)i * 2u + seed), r1 = hash32((uint32_t)i * 2u + 1u + seed * 31u);
        float a = operand(r0), b = operand(r1);
        if (r0 & 0x80000000u) a = -a;
        const float q0 = __fdiv_rn(a, b), q1 = pkd::div_nr(a, b, pkd::rcp_nr(b));
        bd += __float_as_uint(q0) != __float_as_uint(q1);
        const float s0 = __fsqrt_rn(b), s1 = pkd::sqrt0(b);
        bs += __float_as_uint(s0) != __float_as_uint(s1);
    }
    // exact zeros
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        bz += pkd::div_nr(0.f, 3.f, pkd::rcp_nr(3.f)) != 0.f;
        bz += pkd::div_nr(0.f, 1e-12f, pkd::rcp_nr(1e-12f)) != 0.f;
        bz += pkd::div_nr(0.f, 1e-10f, pkd::rcp_nr(1e-10f)) != 0.f;
        bz += pkd::sqrt0(0.f) != 0.f;
    }
    if (bd) atomicAdd(bad + 0, bd);
    if (bs) atomicAdd(bad + 1, bs);
    if (bz) atomicAdd(bad + 2, bz);
}

}  // namespace

extern "C" int pk_selftest_arith(int64_t n, uint32_t seed, int64_t* mismatches3) {
    pk::launch_counter() = 0;
    if (n < 0 || !mismatches3) return pk::fail(PK_ERR_ARG, "pk_selftest_arith: bad argument");
    unsigned long long* d = nullptr;
    PK_CUDA(cudaMalloc(&d, 24));
    PK_CUDA(cudaMemset(d, 0, 24));
    k_selftest_arith<<<148 * 4, 256>>>((uint64_t)n, seed, d);
    PK_LAUNCHED("k_selftest_arith");
    unsigned long long h[3];
    PK_CUDA(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
    cudaFree(d);
    for (int i = 0; i < 3; ++i) mismatches3[i] = (int64_t)h[i];
    return PK_OK;
}
