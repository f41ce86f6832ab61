// rem2d_host_util.h — host-side helpers shared by the CUDA library (rem2d_cuda.cu) and the CPU emulation harness of the device
// code (tests/emu, test infrastructure).
#pragma once
#include <stdint.h>
#include <math.h>

// joint order inside the creature's island: DFS of b2World::Solve from the newest body, each body's joint
// list newest first. Pure topology, so it is computed once on the host.
static inline void island_joint_order(int nb, const int16_t* parent, uint8_t* order) {
    const int nj = nb - 1;
    if (nj <= 0) return;
    char bodyFlag[64] = {0}, jointFlag[64] = {0};      // nb <= 44 (largest capacity class)
    int stack[64], sp = 0, n = 0;
    stack[sp++] = nb - 1;
    bodyFlag[nb - 1] = 1;
    while (sp > 0) {
        const int b = stack[--sp];
        for (int j = nj - 1; j >= 0; --j) {
            if (parent[j] != b && j + 1 != b) continue;
            if (jointFlag[j]) continue;
            const int other = parent[j] == b ? j + 1 : parent[j];
            order[n++] = (uint8_t)j;
            jointFlag[j] = 1;
            if (bodyFlag[other]) continue;
            stack[sp++] = other;
            bodyFlag[other] = 1;
        }
    }
}


// Terrain table from the reference's 200-vertex height field (Modular2DEnv.py:294-306): edgeShape vertices are Python doubles
// rounded to float32; fat AABB = b2EdgeShape::ComputeAABB (radius = polygonRadius) + aabbExtension, single float ops.
static inline void fill_terrain(rem2d::Terrain* t, const double* y, int n, double step) {
    memset(t, 0, sizeof(rem2d::Terrain));
    t->n_edges = n - 1;
    t->step = (float)step;
    for (int i = 0; i < n - 1; ++i) {
        float x1 = (float)((double)i * step), y1 = (float)y[i], x2 = (float)((double)(i + 1) * step), y2 = (float)y[i + 1];
        t->v1x[i] = x1; t->v1y[i] = y1; t->v2x[i] = x2; t->v2y[i] = y2;
        volatile float lox = (x1 < x2 ? x1 : x2), loy = (y1 < y2 ? y1 : y2), hix = (x1 > x2 ? x1 : x2), hiy = (y1 > y2 ? y1 : y2);
        volatile float a;
        a = lox - RB_POLY_RADIUS; t->flx[i] = a - RB_AABB_EXT;
        a = loy - RB_POLY_RADIUS; t->fly[i] = a - RB_AABB_EXT;
        a = hix + RB_POLY_RADIUS; t->fhx[i] = a + RB_AABB_EXT;
        a = hiy + RB_POLY_RADIUS; t->fhy[i] = a + RB_AABB_EXT;
    }
}
static inline rem2d::Consts make_consts(const rem2d_config* cfg) {
    rem2d::Consts k;
    k.dt = cfg->dt; k.gravity_y = cfg->gravity_y;
    k.friction = sqrtf(cfg->terrain_friction * cfg->module_friction);       // b2MixFriction
    k.vel_iters = cfg->velocity_iterations; k.pos_iters = cfg->position_iterations;
    k.continuous = cfg->continuous; k.allow_sleep = cfg->allow_sleep; k.terminate = cfg->terminate;
    k.evaluation_steps = cfg->evaluation_steps;
    k.p_gain = cfg->p_gain; k.wod_speed = cfg->wod_speed; k.env_length = cfg->env_length;
    return k;
}
