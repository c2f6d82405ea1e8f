// common.h -- shared host helpers of libhyperpose_b200 (error string, CUDA checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

namespace hpb {
void set_error(const char* fmt, ...);
const char* get_error();
}

#define HP_CUDA_TRY(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            hpb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return HP_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)
