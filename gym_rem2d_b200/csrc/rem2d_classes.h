// rem2d_classes.h — capacity classes and the per-class launch table (one translation unit per class, so the
// template instantiations compile in parallel).
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

struct ClassOps {
    int nb, nc, nt, nj, off_body, off_joint, off_cont, off_edge, words, hot_words;
    cudaError_t (*set_attributes)();
    void (*reset)(int grid, cudaStream_t st, float* state, const int* lane_creature, rem2d::DevPop p);
    void (*step)(int grid, cudaStream_t st, float* state, int n_ticks, const rem2d::Terrain* ter, const rem2d::Consts* k,
                 unsigned long long* counters);
    void (*episode)(int grid, cudaStream_t st, float* slots, const int* order, int n_order, int* queue, rem2d::DevPop p,
                    const rem2d::Terrain* ter, const rem2d::Consts* k, int max_ticks, double* fitness, int* ticks, int* alive,
                    int* status, unsigned long long* counters, int park_ticks, int park_cap, float* park_state,
                    int* park_creature, int* park_count);
    void (*tail)(int grid, cudaStream_t st, float* park_state, int* park_creature, int first_slot, int n_parked, const rem2d::Terrain* ter,
                 const rem2d::Consts* k, int max_ticks, double* fitness, int* ticks, int* alive, int* status,
                 unsigned long long* counters);
};
#define X(i, NB, NC, NT) extern const ClassOps rem2d_class_ops_##i;
REM2D_CLASSES(X)
#undef X
