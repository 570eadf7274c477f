// Host-side helpers shared by the translation units of libfab_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include "fab_b200.h"

int fab_fail(int code, const std::string& msg);               // records the thread's last error
int fab_cuda_fail(cudaError_t e, const char* what);
bool fab_flow_ok(const fab_flow_desc* f);
bool fab_target_ok(const fab_target_desc* t);
inline bool fab_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define FAB_CK_LAUNCH(what)                                         \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return fab_cuda_fail(_e, what);      \
    } while (0)
