// recent_calls.h -- host-side memory of the last forward / re-blend calls of a thread (plain C++, no CUDA).
//
// gstar_raster_reblend must know the layout of the call it re-blends (instance capacity, hit-log slots, R) without
// reading the device header back, so every forward / re-blend leaves an entry here, keyed by the address of its (aligned)
// image buffer.  Least-recently-USED replacement: an entry that was just refreshed -- the caller's caching allocator hands
// the same address to call after call -- or just looked up -- the source of several re-blends -- is never the next to go.
// (A FIFO cursor is wrong here: a refreshed entry can sit exactly under the cursor and be evicted by the very next call.)
#pragma once
#include <stddef.h>
#include <stdint.h>

struct RecentCalls {
    struct Entry {
        const char* img = nullptr;
        size_t cap = 0, log_slots = 0;
        uint32_t R = 0;
        int P = 0, W = 0, H = 0;
        unsigned long long stamp = 0;  // last use (remember or find); 0 = never used
    };
    static constexpr int N = 16;
    Entry e[N];
    unsigned long long clock = 0;

    void remember(const char* img, size_t cap, size_t log_slots, uint32_t R, int P, int W, int H)
    {
        int slot = -1;
        for (int i = 0; i < N; i++)
            if (e[i].stamp != 0 && e[i].img == img) slot = i;
        if (slot < 0) {
            slot = 0;
            for (int i = 1; i < N; i++)
                if (e[i].stamp < e[slot].stamp) slot = i;
        }
        e[slot].img = img; e[slot].cap = cap; e[slot].log_slots = log_slots;
        e[slot].R = R; e[slot].P = P; e[slot].W = W; e[slot].H = H;
        e[slot].stamp = ++clock;
    }
    const Entry* find(const char* img)
    {
        for (int i = 0; i < N; i++)
            if (e[i].stamp != 0 && e[i].img == img) {
                e[i].stamp = ++clock;
                return &e[i];
            }
        return nullptr;
    }
};
