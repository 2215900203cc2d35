// Shared host-side helpers of libpifpaf_b200 (error slot, launch counter, CUDA checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "pifpaf_b200.h"

namespace pifpaf {

// thread-local last-error slot (pifpaf_last_error)
std::string& last_error_slot();
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PIFPAF_CUDA_TRY(expr)                                                          \
    do {                                                                               \
        cudaError_t err__ = (expr);                                                    \
        if (err__ != cudaSuccess) {                                                    \
            ::pifpaf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                                __FILE__, __LINE__);                                   \
            return PIFPAF_E_CUDA;                                                      \
        }                                                                              \
    } while (0)

#define PIFPAF_CHECK_ARG(cond, msg)                           \
    do {                                                      \
        if (!(cond)) {                                        \
            ::pifpaf::set_error("bad argument: %s", msg);     \
            return PIFPAF_E_BADARG;                           \
        }                                                     \
    } while (0)

#define PIFPAF_LAUNCH_CHECK()                                                         \
    do {                                                                              \
        ::pifpaf::count_launch();                                                     \
        cudaError_t err__ = cudaGetLastError();                                       \
        if (err__ != cudaSuccess) {                                                   \
            ::pifpaf::set_error("kernel launch failed: %s (%s:%d)",                   \
                                cudaGetErrorString(err__), __FILE__, __LINE__);       \
            return PIFPAF_E_CUDA;                                                     \
        }                                                                             \
    } while (0)

}  // namespace pifpaf
