// TEST INFRASTRUCTURE ONLY.  CPU lane emulator for tile_match_gym_b200/csrc/tmg_device.cuh.
//
// Lets the test-suite run the *device* code of the product on the CPU, one group of L lanes at a time,
// each lane a ucontext fiber that yields at every warp collective.  It exists to debug and fuzz kernel
// logic in a container without a GPU; it is NOT a CPU fallback: the product package never builds,
// loads or links it, and the GPU parity tests do not use it.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(x) __attribute__((aligned(x)))

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 v; v.x = x; v.y = y; return v; }

namespace emu {
struct Dim { unsigned x, y, z; };
struct Lane {
    ucontext_t ctx;
    Dim tid;
    int lane_in_warp;
    bool done;
    char* stack;
};
extern Lane* cur;          // fiber that is running
extern Dim block_idx, block_dim, grid_dim;
extern int group_lanes;    // L of the running group
extern int group_shift;    // first lane-in-warp of the running group
extern int site_id;        // source line of the collective the running lane is about to enter
const uint64_t* gather(uint64_t v, unsigned mask);  // every lane's value, indexed by lane-in-warp
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::block_idx)
#define blockDim (emu::block_dim)
#define gridDim (emu::grid_dim)

static inline int __ffs(int x) { return x == 0 ? 0 : __builtin_ctz((unsigned)x) + 1; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u)); }
static inline long long clock64() { return 0; }
static inline void __threadfence() {}
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T atomicAdd(T* a, T v) { T o = *a; *a += v; return o; }
using std::max;
using std::min;

static inline void __syncwarp(unsigned mask) { emu::gather(0, mask); }
static inline int __any_sync(unsigned mask, int pred) {
    const uint64_t* v = emu::gather(pred ? 1 : 0, mask);
    for (int i = 0; i < 32; ++i) if (((mask >> i) & 1u) && v[i]) return 1;
    return 0;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t* v = emu::gather(pred ? 1 : 0, mask);
    unsigned out = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) out |= (unsigned)(v[i] & 1) << i;
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T val, int src, int width = 32) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    const int me = emu::cur->lane_in_warp;
    const int base = me & ~(width - 1);
    return (T)(uint32_t)v[base + (src & (width - 1))];
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T val, int d, int width = 32) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    const int me = emu::cur->lane_in_warp;
    const int base = me & ~(width - 1);
    const int idx = (me - base) + d;
    return (T)(uint32_t)v[idx < width ? base + idx : me];
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T val, int d, int width = 32) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    const int me = emu::cur->lane_in_warp;
    const int base = me & ~(width - 1);
    const int idx = (me - base) - d;
    return (T)(uint32_t)v[idx >= 0 ? base + idx : me];
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T val, int lanemask, int width = 32) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    const int me = emu::cur->lane_in_warp;
    return (T)(uint32_t)v[(me ^ lanemask) & 31];
}
static inline int __reduce_add_sync(unsigned mask, int val) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    int s = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) s += (int)(uint32_t)v[i];
    return s;
}
static inline int __reduce_max_sync(unsigned mask, int val) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    int s = INT32_MIN;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) s = std::max(s, (int)(uint32_t)v[i]);
    return s;
}
static inline int __reduce_min_sync(unsigned mask, int val) {
    const uint64_t* v = emu::gather((uint64_t)(uint32_t)val, mask);
    int s = INT32_MAX;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) s = std::min(s, (int)(uint32_t)v[i]);
    return s;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned val) {
    const uint64_t* v = emu::gather((uint64_t)val, mask);
    unsigned s = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) s |= (unsigned)v[i];
    return s;
}
