// rem2d_class.cu — the three kernels of ONE capacity class (compiled once per class with -DREM2D_CLASS_ID=k).
#include "rem2d_classes.h"

using namespace rem2d;

// Build the world of every creature of this class (static creature -> lane mapping, used by rem2d_step).
template <int NB, int NC, int NT>
__global__ void __launch_bounds__(32) reset_kernel(float* state, const int* __restrict__ lane_creature, DevPop p) {
    using SimT = Sim<NB, NC, NT>;
    const int lane = threadIdx.x, batch = blockIdx.x;
    SimT sim;
    sim.g = state + (size_t)batch * SimT::WORDS * 32 + lane;
    sim.build_world(p, lane_creature[batch * 32 + lane]);
}

// Whole episodes with dynamic lane refill: every lane pulls the next creature of its class from a queue
// (big creatures first), builds its world in the lane's column of the warp's state block, ticks it until the
// episode ends, writes fitness / ticks and pulls the next one. Lanes of a warp are therefore always busy until
// the queue drains, instead of idling until the longest-lived creature of a fixed batch dies; and the cold
// state of the few hundred resident warps stays L2-resident.
template <int NB, int NC, int NT>
__global__ void __launch_bounds__(32, 1) episode_kernel(float* slots, const int* __restrict__ order, int n_order, int* queue,
                                                     DevPop p, const Terrain* __restrict__ ter, const Consts* __restrict__ k,
                                                     int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                                                     unsigned long long* counters, int park_ticks, int park_cap,
                                                     float* park_state, int* park_creature, int* park_count) {
    using SimT = Sim<NB, NC, NT>;
    extern __shared__ float hot[];
    const int lane = threadIdx.x;
    SimT sim;
    sim.g = slots + (size_t)blockIdx.x * SimT::WORDS * 32 + lane;
    sim.h = hot + lane;
    sim.ter = ter; sim.k = k;
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) sim.cnt.c[i] = 0u;
    int my = -1;
    bool exhausted = false;
    Cnt snapshot = sim.cnt;
    for (;;) {
        if (my < 0 && !exhausted) {
            int idx = atomicAdd(queue, 1);
            if (idx < n_order) { my = order[idx]; sim.build_world(p, my); snapshot = sim.cnt; }
            else exhausted = true;
        }
        if (!__any_sync(0xffffffffu, my >= 0)) break;
        if (my >= 0) {
            sim.tick();
            const int t = sim.Si(S_TICKS), st = sim.Si(S_STATUS);
            if (!sim.Si(S_ALIVE) || t >= max_ticks || st) {
                fitness[my] = sim.Sd(S_FIT_LO); ticks[my] = t; alive[my] = sim.Si(S_ALIVE); status[my] = st;
                // a creature that outgrew a capacity of this class is re-run by the host in the next class up:
                // its partial work must not be counted
                if (st) sim.cnt = snapshot;
                my = -1;
            } else if (park_ticks > 0 && t >= park_ticks && *(volatile int*)park_count < park_cap) {
                // long-lived creature: park its state; the latency-oriented tail kernel (one warp per creature) finishes it.
                // Only a bounded number of creatures is parked (the tail kernel trades throughput for latency): in an evolved
                // population where most creatures live long, the rest simply continue here.
                // (the counter never exceeds the cap: the host hands every counted slot to a tail kernel)
                int slot = -1, seen = *(volatile int*)park_count;
                while (seen < park_cap) {
                    const int prev = atomicCAS(park_count, seen, seen + 1);
                    if (prev == seen) { slot = seen; break; }
                    seen = prev;
                }
                if (slot < 0) continue;
                float* dst = park_state + (size_t)(slot >> 5) * SimT::WORDS * 32 + (slot & 31);
                for (int w = 0; w < SimT::WORDS; ++w) dst[w * 32] = sim.g[w * 32];
                __threadfence();                                 // the column is visible before the slot is published
                atomicExch(&park_creature[slot], my + 1);        // 0 = allocated but not yet published
                my = -1;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) {
        unsigned long long v = sim.cnt.c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&counters[i], v);
    }
}

// One warp per batch of 32 creatures; each lane advances its creature by up to n_ticks ticks.
template <int NB, int NC, int NT>
__global__ void __launch_bounds__(32, 1) step_kernel(float* state, int n_ticks, const Terrain* __restrict__ ter,
                                                  const Consts* __restrict__ k, unsigned long long* counters) {
    using SimT = Sim<NB, NC, NT>;
    extern __shared__ float hot[];
    const int lane = threadIdx.x, batch = blockIdx.x;
    SimT sim;
    sim.g = state + (size_t)batch * SimT::WORDS * 32 + lane;
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


// Tail kernel: ONE WARP PER CREATURE for the few long-lived creatures that bound the makespan. Lane 0 runs the scalar
// parts of the tick on the creature's parked column; all 32 lanes share the 180 velocity iterations as a bit-identical
// dependency wavefront (Sim::wavefront_velocity), which cuts the per-tick latency of a large creature several times.
template <int NB, int NC, int NT>
__global__ void __launch_bounds__(32, 1) tail_kernel(float* park_state, int* park_creature, int first_slot, int n_parked,
                                                     const Terrain* __restrict__ ter, const Consts* __restrict__ k, int max_ticks,
                                                     double* fitness, int* ticks, int* alive, int* status,
                                                     unsigned long long* counters) {
    using SimT = Sim<NB, NC, NT, 1>;
    __shared__ float hot[SimT::HOT_WORDS];
    __shared__ int ver[NB];
    const int lane = threadIdx.x, slot = first_slot + blockIdx.x;
    if ((int)blockIdx.x >= n_parked) return;
    // the slot was allocated by an episode kernel that may still be running: wait until its column has been published
    int my = -1;
    if (lane == 0) {
        int v;
        while ((v = atomicAdd(&park_creature[slot], 0)) == 0) __nanosleep(500);
        my = v - 1;
        __threadfence();
    }
    my = __shfl_sync(0xffffffffu, my, 0);
    SimT sim;
    sim.g = park_state + (size_t)(slot >> 5) * SimT::WORDS * 32 + (slot & 31);
    sim.h = hot;
    sim.ter = ter; sim.k = k;
#pragma unroll
    for (int i = 0; i < REM2D_N_COUNTERS; ++i) sim.cnt.c[i] = 0u;
    sim.nb = sim.Si(S_NB); sim.nj = sim.nb - 1;
    for (;;) {
        int nt = 0, solved = 0;
        if (lane == 0) solved = sim.tick_pre(nt) ? 1 : 0;
        solved = __shfl_sync(0xffffffffu, solved, 0);
        nt = __shfl_sync(0xffffffffu, nt, 0);
        __syncwarp();
        if (solved) {
            if (sim.nj + nt <= 64) sim.wavefront_velocity(nt, ver, lane);
            else if (lane == 0) sim.solve_velocity(nt);
        }
        __syncwarp();
        int done = 0;
        if (lane == 0) {
            sim.tick_post(solved != 0, nt);
            const int t = sim.Si(S_TICKS), st = sim.Si(S_STATUS);
            if (!sim.Si(S_ALIVE) || t >= max_ticks || st) {
                fitness[my] = sim.Sd(S_FIT_LO); ticks[my] = t; alive[my] = sim.Si(S_ALIVE); status[my] = st;
                done = 1;
            }
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done) break;
    }
    if (lane == 0) {
        const int st = sim.Si(S_STATUS);
        if (!st)
            for (int i = 0; i < REM2D_N_COUNTERS; ++i)
                if (sim.cnt.c[i]) atomicAdd(&counters[i], (unsigned long long)sim.cnt.c[i]);
    }
}

#define X(i, NB_, NC_, NT_) \
    constexpr int kNB_##i = NB_, kNC_##i = NC_, kNT_##i = NT_;
REM2D_CLASSES(X)
#undef X
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
constexpr int kNB = CAT(kNB_, REM2D_CLASS_ID), kNC = CAT(kNC_, REM2D_CLASS_ID), kNT = CAT(kNT_, REM2D_CLASS_ID);
using SimK = Sim<kNB, kNC, kNT>;

static cudaError_t set_attributes() {
    cudaError_t e = cudaFuncSetAttribute(step_kernel<kNB, kNC, kNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SimK::HOT_WORDS * 128);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(episode_kernel<kNB, kNC, kNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SimK::HOT_WORDS * 128);
    if (e != cudaSuccess) return e;
    // All kernels that can be resident together should agree on the shared-memory carve-out of the SM: a small-smem kernel
    // (tail) would otherwise pin its SMs in a large-L1 configuration and lock the big episode CTAs of other classes out
    // (measured: 1.4x slower whole run when a tail kernel was resident next to the episode kernels).
    e = cudaFuncSetAttribute(step_kernel<kNB, kNC, kNT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(episode_kernel<kNB, kNC, kNT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tail_kernel<kNB, kNC, kNT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
static void launch_reset(int grid, cudaStream_t st, float* state, const int* lane_creature, DevPop p) {
    reset_kernel<kNB, kNC, kNT><<<grid, 32, 0, st>>>(state, lane_creature, p);
}
static void launch_step(int grid, cudaStream_t st, float* state, int n_ticks, const Terrain* ter, const Consts* k,
                        unsigned long long* counters) {
    step_kernel<kNB, kNC, kNT><<<grid, 32, SimK::HOT_WORDS * 128, st>>>(state, n_ticks, ter, k, counters);
}
static void launch_episode(int grid, cudaStream_t st, float* slots, const int* order, int n_order, int* queue, DevPop p,
                           const Terrain* ter, const Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                           unsigned long long* counters, int park_ticks, int park_cap, float* park_state, int* park_creature,
                           int* park_count) {
    episode_kernel<kNB, kNC, kNT><<<grid, 32, SimK::HOT_WORDS * 128, st>>>(slots, order, n_order, queue, p, ter, k, max_ticks,
                                                                            fitness, ticks, alive, status, counters, park_ticks, park_cap,
                                                                            park_state, park_creature, park_count);
}
static void launch_tail(int grid, cudaStream_t st, float* park_state, int* park_creature, int first_slot, int n_parked, const Terrain* ter,
                        const Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                        unsigned long long* counters) {
    tail_kernel<kNB, kNC, kNT><<<grid, 32, 0, st>>>(park_state, park_creature, first_slot, n_parked, ter, k, max_ticks, fitness, ticks, alive,
                                                    status, counters);
}
extern const ClassOps CAT(rem2d_class_ops_, REM2D_CLASS_ID) = {
    kNB, kNC, kNT, SimK::NJ, SimK::OFF_BODY, SimK::OFF_JOINT, SimK::OFF_CONT, SimK::OFF_EDGE, SimK::WORDS, SimK::HOT_WORDS,
    set_attributes, launch_reset, launch_step, launch_episode, launch_tail };
