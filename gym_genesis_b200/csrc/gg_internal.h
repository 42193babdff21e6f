// Internal declarations shared by the translation units of libgg_b200.so (not part of the C ABI).
#pragma once
#include <stddef.h>
#include <stdint.h>

struct GgPinCache {                  // verified pinned-host address ranges (gg_internal_pinned)
    static constexpr int kN = 16;
    uintptr_t base[kN]; size_t size[kN]; int hits[kN]; int next;
};
bool gg_internal_pinned(GgPinCache* pc, const void* p);
