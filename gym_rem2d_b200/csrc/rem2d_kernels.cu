// rem2d_kernels.cu — the three kernels of the hot path (reset, step, episode with its bulk and tail modes). Compiled once per kernel (-DREM2D_KERNEL_ID=0..3) so the
// translation units build in parallel; every capacity class runs the same code with its own Layout (kernel parameter).
#include <cstdlib>
#include "rem2d_classes.h"

using namespace rem2d;

#if REM2D_KERNEL_ID == 0
// Build the world of every creature of a class (static creature -> lane mapping, used by rem2d_step).
__global__ void __launch_bounds__(32) reset_kernel(const __grid_constant__ Layout L, float* state, const int* __restrict__ lane_creature,
                                                   DevPop p) {
    const int lane = threadIdx.x, batch = blockIdx.x;
    Sim sim;
    sim.L = L; sim.set_mode(false);
    sim.g = state + (size_t)batch * L.words * 32 + lane;
    sim.build_world(p, lane_creature[batch * 32 + lane]);
}
void rem2d_launch_reset(const Layout& L, int grid, cudaStream_t st, float* state, const int* lane_creature, DevPop p) {
    reset_kernel<<<grid, 32, 0, st>>>(L, state, lane_creature, p);
}
#endif

#if REM2D_KERNEL_ID == 1
// One warp per batch of 32 creatures; each lane advances its creature by up to n_ticks ticks.
__global__ void __launch_bounds__(32, 1) step_kernel(const __grid_constant__ Layout L, float* state, int n_ticks,
                                                     const Terrain* __restrict__ ter, const Consts* __restrict__ k,
                                                     unsigned long long* counters) {
    extern __shared__ float hot[];
    const int lane = threadIdx.x, batch = blockIdx.x;
    Sim sim;
    sim.L = L; sim.set_mode(false);
    sim.g = state + (size_t)batch * L.words * 32 + lane;
    sim.h = hot + lane;
    sim.ter = ter; sim.k = k;
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) sim.cnt.c[i] = 0u;
    sim.nb = sim.Si(S_NB); sim.nj = sim.nb - 1;
    if (sim.nb > 0) {
        for (int t = 0; t < n_ticks; ++t) {
            if (!sim.Si(S_ALIVE)) break;
            sim.tick();
        }
    }
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) {
        unsigned long long v = sim.cnt.c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&counters[i], v);
    }
}
void rem2d_launch_step(const Layout& L, int grid, cudaStream_t st, float* state, int n_ticks, const Terrain* ter, const Consts* k,
                       unsigned long long* counters) {
    step_kernel<<<grid, 32, L.hot_words * 128, st>>>(L, state, n_ticks, ter, k, counters);
}
#endif

#if REM2D_KERNEL_ID == 2
// Whole episodes, two modes of ONE kernel (one code image: co-resident warps of both modes share the instruction cache).
//
// mode 0, bulk: one lane per creature with dynamic lane refill: every lane pulls the next creature of its class from a
// queue (big creatures first), builds its world in the lane's column of the warp's state block, ticks it until the
// episode ends, writes fitness / ticks and pulls the next one. Lanes of a warp are therefore always busy until the
// queue drains, instead of idling until the longest-lived creature of a fixed batch dies; and the cold state of the
// few hundred resident warps stays L2-resident. Long-lived creatures are parked for mode 1.
//
// mode 1, tail: ONE WARP PER CREATURE for the long-lived creatures that bound the makespan. Lane 0 runs the scalar
// parts of the tick on the creature's parked column; all 32 lanes share the 180 velocity iterations as a bit-identical
// dependency wavefront (Sim::wavefront_velocity), which cuts the per-tick latency of a large creature several times.
// mode 2, warp per creature from tick 0: like mode 1, but the warp builds the world itself. For SMALL populations (every
// creature gets its own resident warp): the run time is then one creature lifetime at the low tail-mode tick latency
// instead of one at the bulk latency (pop 1024: ~5x sooner), at 1/32 of the bulk mode's lane efficiency.
// Dynamic shared memory: hot_words * 128 B (bulk) or thot_rows * 128 B + nb version counters (modes 1, 2).
__global__ void __launch_bounds__(32, 1) episode_kernel(const __grid_constant__ Layout L, int mode, float* slots,
                                                        const int* __restrict__ order, int n_order, int* queue, DevPop p,
                                                        const Terrain* __restrict__ ter, const Consts* __restrict__ k, int max_ticks,
                                                        double* fitness, int* ticks, int* alive, int* status,
                                                        unsigned long long* counters, ParkPolicy park, float* park_state,
                                                        int* park_creature, int* park_count, int first_slot) {
    extern __shared__ float hot[];
    const int lane = threadIdx.x;
    const bool tail = mode != 0;
    Sim sim;
    sim.L = L; sim.set_mode(tail);
    sim.ter = ter; sim.k = k;
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) sim.cnt.c[i] = 0u;
    int my = -1, park_at = park.ticks, loop_iter = 0;
    const int tail_slot = first_slot + blockIdx.x;
    bool exhausted = false;
    int* ver = (int*)(hot + L.thot_rows * 32);       // tail mode only
    if (mode == 2) {
        // whole episode of creature order[blockIdx.x] by this warp: lane 0 builds the world in column blockIdx.x of `slots`
        my = order[blockIdx.x];
        sim.g = slots + (size_t)(blockIdx.x >> 5) * L.words * 32 + (blockIdx.x & 31);
        sim.h = hot;
        if (lane == 0) sim.build_world(p, my);
        __syncwarp();
        sim.nb = sim.Si(S_NB); sim.nj = sim.nb - 1;
        exhausted = true;
    } else if (tail) {
        // the slot was allocated by a bulk warp that may still be running: wait until its column has been published
        const int slot = first_slot + blockIdx.x;
        if (lane == 0) {
            int v;
            while ((v = atomicAdd(&park_creature[slot], 0)) == 0) __nanosleep(500);
            my = v - 1;
            __threadfence();
        }
        my = __shfl_sync(0xffffffffu, my, 0);
        if (park.tail_trace && lane == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            park.tail_trace[slot * 4 + 1] = (unsigned)(t / 1000ull);
        }
        sim.g = park_state + (size_t)(slot >> 5) * L.words * 32 + (slot & 31);
        sim.h = hot;
        sim.nb = sim.Si(S_NB); sim.nj = sim.nb - 1;
        exhausted = true;
    } else {
        sim.g = slots + (size_t)blockIdx.x * L.words * 32 + lane;
        sim.h = hot + lane;
    }
    Cnt snapshot = sim.cnt;
    for (;;) {
        if (my < 0 && !exhausted) {
            int idx = atomicAdd(queue, 1);
            if (idx < n_order) {
                my = order[idx]; sim.build_world(p, my); snapshot = sim.cnt;
                park_at = idx >= park.late_from ? park.late_ticks : park.ticks;
            } else exhausted = true;
        }
        const unsigned live = __ballot_sync(0xffffffffu, my >= 0);
        if (park.trace && lane == 0 && (loop_iter & 3) == 0 && (loop_iter >> 2) < REM2D_TRACE_SAMPLES) {
            unsigned long long t; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned int* tr = park.trace + ((size_t)blockIdx.x * REM2D_TRACE_SAMPLES + (loop_iter >> 2)) * 2;
            tr[0] = (unsigned)(t / 1000ull);
            tr[1] = (unsigned)__popc(live) | ((unsigned)(loop_iter & 0xffff) << 8) | (smid << 24);
        }
        ++loop_iter;
        if (!live) break;
        // drain: no refill any more and only a few lanes of this warp still work -> hand them to the tail mode
        const bool drain = !tail && park.drain_lanes > 0 && __popc(live) <= park.drain_lanes && __any_sync(0xffffffffu, exhausted);
        const bool active = tail ? lane == 0 : my >= 0;
        int nt = 0, solved = 0;
        if (active) solved = sim.tick_pre(nt) ? 1 : 0;
        bool wave = false;
        if (tail) {
            solved = __shfl_sync(0xffffffffu, solved, 0);
            nt = __shfl_sync(0xffffffffu, nt, 0);
            wave = solved && sim.nj + nt <= 64;
            __syncwarp();
        }
        if (wave) sim.wavefront_velocity(nt, ver, lane);
        else if (active && solved) sim.solve_velocity(nt);
        if (tail) __syncwarp();
        if (active) {
            sim.tick_post(solved != 0, nt);
            const int t = sim.Si(S_TICKS), st = sim.Si(S_STATUS);
            if (!sim.Si(S_ALIVE) || t >= max_ticks || st) {
                fitness[my] = sim.Sd(S_FIT_LO); ticks[my] = t; alive[my] = sim.Si(S_ALIVE); status[my] = st;
                // a creature that outgrew a capacity of this class is re-run by the host in the next class up:
                // its partial work must not be counted
                if (st) sim.cnt = snapshot;
                my = -1;
            } else if (!tail && park.ticks > 0 &&
                       (t >= park_at || drain ||
                        (park.lead_from > 0 && t >= park.lead_from && k->terminate &&
                         (double)sim.B(BF_CX, 0) - sim.Sd(S_WOD_LO) >= (double)park.lead)) &&
                       *(volatile int*)park_count < park.cap) {
                // long-lived creature: park its state; the latency-oriented tail mode (one warp per creature) finishes it.
                // (the counter never exceeds the cap: the host hands every counted slot to a tail launch)
                int slot = -1, seen = *(volatile int*)park_count;
                while (seen < park.cap) {
                    const int prev = atomicCAS(park_count, seen, seen + 1);
                    if (prev == seen) { slot = seen; break; }
                    seen = prev;
                }
                if (slot >= 0) {
                    float* dst = park_state + (size_t)(slot >> 5) * L.words * 32 + (slot & 31);
                    for (int w = 0; w < L.words; ++w) dst[w * 32] = sim.g[w * 32];
                    if (park.tail_trace) {
                        unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                        park.tail_trace[slot * 4] = (unsigned)(tt / 1000ull);
                    }
                    __threadfence();                                 // the column is visible before the slot is published
                    atomicExch(&park_creature[slot], my + 1);        // 0 = allocated but not yet published
                    my = -1;
                }
            }
        }
        if (tail) my = __shfl_sync(0xffffffffu, my, 0);
    }
    if (mode == 1 && park.tail_trace && lane == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        park.tail_trace[tail_slot * 4 + 2] = (unsigned)(t / 1000ull);
        park.tail_trace[tail_slot * 4 + 3] = (unsigned)loop_iter - 1u;
    }
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) {
        unsigned long long v = sim.cnt.c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&counters[i], v);
    }
}
void rem2d_launch_episode(const Layout& L, int grid, cudaStream_t st, float* slots, const int* order, int n_order, int* queue, DevPop p,
                          const Terrain* ter, const Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                          unsigned long long* counters, ParkPolicy park, float* park_state, int* park_creature, int* park_count) {
    episode_kernel<<<grid, 32, L.hot_words * 128, st>>>(L, 0, slots, order, n_order, queue, p, ter, k, max_ticks, fitness, ticks, alive,
                                                        status, counters, park, park_state, park_creature, park_count, 0);
}
void rem2d_launch_warp_mode(const Layout& L, int n, cudaStream_t st, float* slots, const int* order, DevPop p, const Terrain* ter,
                            const Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                            unsigned long long* counters) {
    ParkPolicy none = {0, 0, 0, 0, 0, 0, 0.0f, nullptr, nullptr};
    episode_kernel<<<n, 32, (L.thot_rows * 32 + L.nb) * 4, st>>>(L, 2, slots, order, n, nullptr, p, ter, k, max_ticks, fitness, ticks,
                                                                 alive, status, counters, none, nullptr, nullptr, nullptr, 0);
}
void rem2d_launch_tail(const Layout& L, int grid, cudaStream_t st, float* park_state, int* park_creature, int first_slot, int n_parked,
                       const Terrain* ter, const Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                       unsigned long long* counters, unsigned int* tail_trace) {
    (void)n_parked;      // grid == number of parked creatures handed over
    ParkPolicy none = {0, 0, 0, 0, 0, 0, 0.0f, nullptr, tail_trace};
    episode_kernel<<<grid, 32, (L.thot_rows * 32 + L.nb) * 4, st>>>(L, 1, nullptr, nullptr, 0, nullptr, DevPop(), ter, k, max_ticks, fitness,
                                                               ticks, alive, status, counters, none, park_state, park_creature, nullptr,
                                                               first_slot);
}
#endif

// Attributes: every translation unit sets those of its own kernel; rem2d_set_kernel_attributes (kernel 0's unit) calls all.
// All kernels that can be resident together should agree on the shared-memory carve-out of the SM: a small-smem kernel
// (tail) would otherwise pin its SMs in a large-L1 configuration and lock the big episode CTAs of other classes out
// (measured: 1.4x slower whole run when a tail kernel was resident next to the episode kernels).
cudaError_t rem2d_attr_step(int max_hot_words, int carve);
cudaError_t rem2d_attr_episode(int max_hot_words, int carve);
#if REM2D_KERNEL_ID == 0
cudaError_t rem2d_set_kernel_attributes(int max_hot_words, int carve) {
    cudaError_t e = rem2d_attr_step(max_hot_words, carve);
    if (e != cudaSuccess) return e;
    return rem2d_attr_episode(max_hot_words, carve);
}
#elif REM2D_KERNEL_ID == 1
cudaError_t rem2d_attr_step(int max_hot_words, int carve) {
    cudaError_t e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_hot_words * 128);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
}
#elif REM2D_KERNEL_ID == 2
cudaError_t rem2d_attr_episode(int max_hot_words, int carve) {
    cudaError_t e = cudaFuncSetAttribute(episode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_hot_words * 128);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(episode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
}
#endif
