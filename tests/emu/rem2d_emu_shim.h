// rem2d_emu_shim.h — TEST INFRASTRUCTURE. Lets gym_rem2d_b200/csrc/rem2d_device.cuh compile with g++ (-DREM2D_EMU) so that the
// warp-cooperative device code can be executed on the CPU: every lane of a warp is a host thread, __syncwarp / __shfl_sync /
// __ballot_sync are barriers and exchanges among the threads named by the mask. With -ffp-contract=off the float32 arithmetic is
// operation for operation what nvcc -fmad=false emits, so the emulated kernel must equal the oracle bit for bit — which checks the
// lane-group logic (schedules, strided loops, leader sections, group barriers) in a container without a GPU. Not a product path.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <sched.h>

#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __grid_constant__

namespace emu {
struct Barrier {                       // sense-reversing spin barrier for a fixed set of lanes
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 0;
};
struct Warp {
    Barrier bar[64];                   // one per distinct lane mask seen (groups of a warp + the full warp); keyed lazily
    std::atomic<uint32_t> bar_mask[64];
    std::atomic<int> n_bar{0};
    uint64_t slot[32];                 // shuffle / vote exchange
    std::atomic<long> n_syncs{0};
    Warp() { for (auto& m : bar_mask) m.store(0); }
};
struct Ctx { Warp* warp; int lane; int local_sense[64]; };
extern thread_local Ctx* ctx;

inline int barrier_index(Warp* w, uint32_t mask) {
    for (;;) {
        int n = w->n_bar.load(std::memory_order_acquire);
        for (int i = 0; i < n; ++i) if (w->bar_mask[i].load(std::memory_order_acquire) == mask) return i;
        // register a new mask (rare): spin-lock on n_bar via CAS to n+1 after filling slot n
        uint32_t expect = 0;
        if (n < 64 && w->bar_mask[n].compare_exchange_strong(expect, mask)) {
            w->bar[n].n = __builtin_popcount(mask);
            w->n_bar.store(n + 1, std::memory_order_release);
            return n;
        }
        sched_yield();
    }
}
inline void sync(uint32_t mask) {
    Ctx* c = ctx;
    if ((mask & (mask - 1)) == 0) return;                 // a single lane
    int i = barrier_index(c->warp, mask);
    Barrier& b = c->warp->bar[i];
    int s = c->local_sense[i] ^= 1;
    if (b.count.fetch_add(1, std::memory_order_acq_rel) + 1 == b.n) {
        b.count.store(0, std::memory_order_relaxed);
        c->warp->n_syncs.fetch_add(1, std::memory_order_relaxed);
        b.sense.store(s, std::memory_order_release);
    } else {
        int spins = 0;
        while (b.sense.load(std::memory_order_acquire) != s) { if (++spins > 64) { sched_yield(); spins = 0; } }
    }
}
template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
}  // namespace emu

inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::sync(mask); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) {
    emu::Ctx* c = emu::ctx;
    c->warp->slot[c->lane] = emu::to_bits(v);
    emu::sync(mask);
    T r = emu::from_bits<T>(c->warp->slot[src & 31]);
    emu::sync(mask);
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x) { return __shfl_sync(mask, v, emu::ctx->lane ^ x); }
template <class T> inline T __shfl_down_sync(unsigned mask, T v, int d) {
    int src = emu::ctx->lane + d;
    return __shfl_sync(mask, v, src < 32 && ((mask >> src) & 1) ? src : emu::ctx->lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Ctx* c = emu::ctx;
    c->warp->slot[c->lane] = pred ? 1 : 0;
    emu::sync(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1) && c->warp->slot[l]) r |= 1u << l;
    emu::sync(mask);
    return r;
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { sched_yield(); }
