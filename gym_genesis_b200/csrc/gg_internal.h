// Internal declarations shared by the translation units of libgg_b200.so (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

struct GgPinCache {                  // verified pinned-host address ranges (gg_internal_pinned)
    static constexpr int kN = 16;
    uintptr_t base[kN]; size_t size[kN]; int hits[kN]; int next;
};
bool gg_internal_pinned(GgPinCache* pc, const void* p);

// Makes `device` current for the duration of one API call and restores the caller's device afterwards (torch reads its current
// device from the CUDA runtime: a simulator on cuda:1 must not change it behind the caller's back).
#include <cuda_runtime.h>
struct GgDeviceGuard {
    int prev = -1; bool changed = false; cudaError_t err = cudaSuccess;
    explicit GgDeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) { err = cudaSetDevice(device); changed = err == cudaSuccess; }
    }
    ~GgDeviceGuard() { if (changed) cudaSetDevice(prev); }
    GgDeviceGuard(const GgDeviceGuard&) = delete;
    GgDeviceGuard& operator=(const GgDeviceGuard&) = delete;
};
