// Shared host-side helpers of libmccnn_b200 (error string, launch accounting, argument checks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/mccnn_b200.h"

namespace mccnn {

void set_error(const char *fmt, ...);
void clear_error();
void count_launch(int n = 1);

inline int dpitch(int D) { return (D + 3) & ~3; }

// Every launch goes through this: records the launch for mccnn_launch_count() and turns a failed
// launch into MCCNN_ERR_CUDA with the CUDA message.
int check_launch(const char *what);

}  // namespace mccnn

#define MCCNN_REQUIRE(cond, ...)                       \
    do {                                               \
        if (!(cond)) {                                 \
            mccnn::set_error(__VA_ARGS__);             \
            return MCCNN_ERR_ARG;                      \
        }                                              \
    } while (0)

#define MCCNN_LAUNCHED(what)                           \
    do {                                               \
        int rc__ = mccnn::check_launch(what);          \
        if (rc__ != MCCNN_OK) return rc__;             \
    } while (0)

#define MCCNN_CUDA(call)                                                       \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            mccnn::set_error("%s: %s", #call, cudaGetErrorString(e__));        \
            return MCCNN_ERR_CUDA;                                             \
        }                                                                      \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
