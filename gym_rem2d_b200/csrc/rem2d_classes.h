// rem2d_classes.h — capacity classes and the launch interface of the three kernels (rem2d_kernels.cu; one translation
// unit per kernel so that they compile in parallel). All classes run the SAME code: the class only selects a Layout and
// a default group size.
#pragma once
#include <cuda_runtime.h>
#include "rem2d_device.cuh"

// NB bodies, NC contact-pool slots (fat-AABB overlaps), NT touching contacts staged in shared memory (further touching
// contacts, up to NC, spill to the cold block: correct but slower), GS = largest log2(lanes per creature) the automatic choice
// may give the class (rem2d_cuda.cu: choose_groups_and_grids; 1-2 body creatures have nothing to share). Hot words per creature
// = 5*NB + 17*(NB-1) + 21*NT (+ schedule rows per warp when lanes are shared); shared memory, not registers, bounds the resident
// creatures per SM, so lanes are the free resource: a group of 2^gs lanes shares one creature's solver sweeps (static modulo
// schedule, rem2d_device.cuh) and its per-body / per-joint / per-contact loops.
#define REM2D_CLASSES(X) \
    X(0, 1, 10, 3, 0)    \
    X(1, 2, 16, 4, 0)    \
    X(2, 4, 28, 4, 1)    \
    X(3, 8, 48, 6, 3)    \
    X(4, 12, 64, 6, 5)   \
    X(5, 16, 80, 6, 5)   \
    X(6, 22, 104, 6, 5)  \
    X(7, 32, 144, 8, 5)  \
    X(8, 44, 192, 10, 5)
#define N_CLASSES 9

// When the queue-mode episode kernel hands a creature over to the tail mode (launches of the same kernel).
struct ParkPolicy {
    int ticks;        // park a creature that is still alive after this many ticks (0: never park)
    int cap;          // at most this many creatures of the class are parked
    float lead_x;     // > 0: a creature whose root has reached x >= lead_x (after lead_from ticks) is parked at once: the wall of death
    int lead_from;    //   needs ticks/wod_speed ticks to get there, so the creature would reach the tick threshold anyway - it just
                      //   moves to the low-latency launches sooner (the long-lived creatures bound the makespan)
    int late_from;    // creatures pulled from this position of the class queue on (= after the first round) are late starters:
    int late_ticks;   //   they park after late_ticks ticks - they bound the makespan, so they move to the low-latency launches sooner
    // diagnostics (trace option): every 4th tick lane 0 of each warp records {globaltimer us, live creatures | tick << 8 |
    // smid << 24}; REM2D_TRACE_SAMPLES entries per warp. Null in production.
    unsigned int* trace;
    unsigned int* tail_trace;   // per park slot: {us parked, us tail warp started, us finished, ticks run by the tail warp}
};
#define REM2D_TRACE_SAMPLES 1024

// Launchers (rem2d_kernels.cu). `carve` = cudaFuncAttributePreferredSharedMemoryCarveout for all kernels.
cudaError_t rem2d_set_kernel_attributes(int max_hot_bytes, int carve);
int rem2d_episode_blocks_per_sm(int image, int dyn_smem_bytes);
void rem2d_launch_reset(const rem2d::Layout& L, int gs, int n_batches, cudaStream_t st, float* state, const int* lane_creature, rem2d::DevPop p);
void rem2d_launch_step(const rem2d::Layout& L, int gs, int n_batches, cudaStream_t st, float* state, int n_ticks, const rem2d::Terrain* ter,
                       const rem2d::Consts* k, unsigned long long* counters);
void rem2d_launch_episode(const rem2d::Layout& L, int image, int gs, int grid, cudaStream_t st, float* slots, const int* order, int n_order, int* queue,
                          rem2d::DevPop p, const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness,
                          int* ticks, int* alive, int* status, unsigned long long* counters, ParkPolicy park, float* park_state,
                          int* park_creature, int* park_count, int refill);
void rem2d_launch_tail(const rem2d::Layout& L, int image, int gs, cudaStream_t st, float* park_state, int* park_creature, int first_slot,
                       int n_parked, const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness, int* ticks,
                       int* alive, int* status, unsigned long long* counters, unsigned int* tail_trace);

// Per-class view used by the host code.
struct ClassOps {
    rem2d::Layout L;
    int nb, nc, nt, nj, off_body, off_joint, off_cont, off_edge, words, hot_words, gs;
    explicit ClassOps(int NB, int NC, int NT, int GS) : L(rem2d::make_layout(NB, NC, NT)) {
        nb = L.nb; nc = L.nc; nt = L.nt; nj = L.nj; off_body = rem2d::S_COUNT; off_joint = L.off_joint; off_cont = L.off_cont;
        off_edge = L.off_edge; words = L.words; hot_words = L.hot_words; gs = GS;
    }
    int hot_bytes(int gshift) const { return rem2d::make_hot_layout(L, gshift).rows * 128; }
    template <class... A> void reset(A... a) const { rem2d_launch_reset(L, a...); }
    template <class... A> void step(A... a) const { rem2d_launch_step(L, a...); }
    template <class... A> void episode(A... a) const { rem2d_launch_episode(L, a...); }
    template <class... A> void tail(A... a) const { rem2d_launch_tail(L, a...); }
};
