// Error/launch bookkeeping shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/putranse.h"

namespace pk {

std::string& last_error();  // thread-local
int& launch_counter();      // thread-local: kernel launches issued by the current pk_* call

inline int fail(int code, const std::string& msg) {
    last_error() = msg;
    return code;
}

inline int cuda_fail(cudaError_t e, const char* what) {
    return fail(PK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace pk

#define PK_CUDA(expr)                                            \
    do {                                                         \
        cudaError_t e__ = (expr);                                \
        if (e__ != cudaSuccess) return pk::cuda_fail(e__, #expr); \
    } while (0)

#define PK_LAUNCHED(what)                                            \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return pk::cuda_fail(e__, what);     \
        ++pk::launch_counter();                                      \
    } while (0)
