// rem2d_classes.h — capacity classes and the launch interface of the three kernels (rem2d_kernels.cu; one translation
// unit per kernel so that they compile in parallel). All classes run the SAME code: the class only selects a Layout.
#pragma once
#include <cuda_runtime.h>
#include "rem2d_device.cuh"

// NB bodies, NC contact-pool slots (fat-AABB overlaps), NT touching contacts staged in shared memory
// (further touching contacts, up to NC, spill to the cold block: correct but slower).
// hot words/lane = 5*NB + 17*(NB-1) + 21*NT; one warp needs 128 B per word (NB=22: 593 words = 74 KB -> 3 warps/SM).
#define REM2D_CLASSES(X) \
    X(0, 1, 10, 3)       \
    X(1, 2, 16, 4)       \
    X(2, 4, 28, 4)       \
    X(3, 8, 48, 6)       \
    X(4, 12, 64, 6)      \
    X(5, 16, 80, 6)      \
    X(6, 22, 104, 6)     \
    X(7, 32, 144, 8)     \
    X(8, 44, 192, 10)
#define N_CLASSES 9

// When the bulk (lane-per-creature) episode kernel hands a creature over to the warp-per-creature tail mode (launches of the same kernel).
struct ParkPolicy {
    int ticks;        // park a creature that is still alive after this many ticks (0: never park)
    int cap;          // at most this many creatures of the class are parked
    int late_ticks;   // threshold for creatures pulled from position >= late_from of the class queue (the late starters
    int late_from;    //   bound the makespan: they move to the low-latency kernel sooner)
    int drain_lanes;  // once the queue is empty, a warp with <= this many live lanes parks them all and exits
    int lead_from;    // lifetime prediction: from this tick on (0: off) a creature whose lead over the wall of death is at
    float lead;       //   least `lead` (world units; lead / wod_speed = ticks it would survive standing still) is parked
    // diagnostics (REM2D_TRACE=1): every 4th tick lane 0 of each warp records {globaltimer us, live lanes | tick << 8 |
    // smid << 24}; REM2D_TRACE_SAMPLES entries per warp. Null in production.
    unsigned int* trace;
    unsigned int* tail_trace;   // per park slot: {us parked, us tail warp started, us finished, ticks run by the tail warp}
};
#define REM2D_TRACE_SAMPLES 1024

// Launchers (rem2d_kernels.cu). `carve` = cudaFuncAttributePreferredSharedMemoryCarveout for all kernels.
cudaError_t rem2d_set_kernel_attributes(int max_hot_words, int carve);
void rem2d_launch_reset(const rem2d::Layout& L, int grid, cudaStream_t st, float* state, const int* lane_creature, rem2d::DevPop p);
void rem2d_launch_step(const rem2d::Layout& L, int grid, cudaStream_t st, float* state, int n_ticks, const rem2d::Terrain* ter,
                       const rem2d::Consts* k, unsigned long long* counters);
void rem2d_launch_episode(const rem2d::Layout& L, int grid, cudaStream_t st, float* slots, const int* order, int n_order, int* queue,
                          rem2d::DevPop p, const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness,
                          int* ticks, int* alive, int* status, unsigned long long* counters, ParkPolicy park, float* park_state,
                          int* park_creature, int* park_count);
void rem2d_launch_warp_mode(const rem2d::Layout& L, int n, cudaStream_t st, float* slots, const int* order, rem2d::DevPop p,
                            const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness, int* ticks, int* alive,
                            int* status, unsigned long long* counters);
void rem2d_launch_tail(const rem2d::Layout& L, int grid, cudaStream_t st, float* park_state, int* park_creature, int first_slot,
                       int n_parked, const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness, int* ticks,
                       int* alive, int* status, unsigned long long* counters, unsigned int* tail_trace);

// Per-class view used by the host code.
struct ClassOps {
    rem2d::Layout L;
    int nb, nc, nt, nj, off_body, off_joint, off_cont, off_edge, words, hot_words;
    explicit ClassOps(int NB, int NC, int NT) : L(rem2d::make_layout(NB, NC, NT)) {
        nb = L.nb; nc = L.nc; nt = L.nt; nj = L.nj; off_body = rem2d::S_COUNT; off_joint = L.off_joint; off_cont = L.off_cont;
        off_edge = L.off_edge; words = L.words; hot_words = L.hot_words;
    }
    template <class... A> void reset(A... a) const { rem2d_launch_reset(L, a...); }
    template <class... A> void step(A... a) const { rem2d_launch_step(L, a...); }
    template <class... A> void episode(A... a) const { rem2d_launch_episode(L, a...); }
    template <class... A> void tail(A... a) const { rem2d_launch_tail(L, a...); }
    template <class... A> void warp_mode(A... a) const { rem2d_launch_warp_mode(L, a...); }
};
