// Error plumbing shared by every translation unit of libdomainrag_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/domainrag_b200.h"

namespace drag {

// Status codes (DRAG_OK, DRAG_ERR_*) come from the public header.
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define DRAG_CUDA(expr)                                                                        \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            return ::drag::fail(DRAG_ERR_CUDA, std::string(#expr) + ": " +             \
                                                           cudaGetErrorString(_e));            \
        }                                                                                      \
    } while (0)

#define DRAG_REQUIRE(cond, msg)                                                                \
    do {                                                                                       \
        if (!(cond)) return ::drag::fail(DRAG_ERR_INVALID, std::string(msg));          \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

int device_sm_count();

// Every kernel launch of the library bumps this counter (bench.py reports it as `gpu_launches`; drag_launch_count).
extern long long g_launch_count;
static inline void count_launch() { ++g_launch_count; }

}  // namespace drag
