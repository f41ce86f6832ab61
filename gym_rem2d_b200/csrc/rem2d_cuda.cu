// rem2d_cuda.cu — host side + C-ABI (include/rem2d.h) of the CUDA build, sm_100a only (kernels: rem2d_kernels.cu).
//
// Sorts creatures into capacity classes (by body count), keeps one lane-interleaved state block per 32 creatures in
// HBM, and runs every class on its own stream so small and large classes overlap on the 148 SMs: rem2d_step on a
// static creature -> lane mapping, rem2d_run_episodes / rem2d_evaluate on persistent warps that pull creatures from a
// per-class queue (bulk mode), hand long-lived creatures to warp-per-creature launches (tail mode), or give every
// creature its own warp when the population is small. There is no CPU physics fallback: every tick is computed by
// rem2d::Sim::tick on the device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <string>
#include <vector>

#include "rem2d_classes.h"
#include "rem2d_host_util.h"

using namespace rem2d;

// ------------------------------------------------------------------ capacity classes (rem2d_classes.h)
static const ClassOps g_classes_tab[N_CLASSES] = {
#define X(i, NB, NC, NT, GS) ClassOps(NB, NC, NT, GS),
    REM2D_CLASSES(X)
#undef X
};
#define g_classes(k) (g_classes_tab[k])

// fitness / ticks / alive / status of every creature of a class -> creature-indexed outputs
__global__ void gather_kernel(const float* state, const int* __restrict__ lane_creature, int n_lanes, int words,
                              double* fitness, int* ticks, int* alive, int* status) {
    int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= n_lanes) return;
    int c = lane_creature[gl];
    if (c < 0) return;
    const float* g = state + (size_t)(gl >> 5) * words * 32 + (gl & 31);
    fitness[c] = __hiloint2double(__float_as_int(g[S_FIT_HI * 32]), __float_as_int(g[S_FIT_LO * 32]));
    ticks[c] = __float_as_int(g[S_TICKS * 32]);
    alive[c] = __float_as_int(g[S_ALIVE * 32]);
    status[c] = __float_as_int(g[S_STATUS * 32]);
}

// root x / wall of death / alive of every creature of a class -> creature-indexed outputs (rem2d_read_roots)
__global__ void roots_kernel(const float* state, const int* __restrict__ lane_creature, int n_lanes, int words,
                             float* root_x, double* wod, int* alive) {
    int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= n_lanes) return;
    int c = lane_creature[gl];
    if (c < 0) return;
    const float* g = state + (size_t)(gl >> 5) * words * 32 + (gl & 31);
    root_x[c] = g[(S_COUNT + BF_CX) * 32];
    wod[c] = __hiloint2double(__float_as_int(g[S_WOD_HI * 32]), __float_as_int(g[S_WOD_LO * 32]));
    alive[c] = __float_as_int(g[S_ALIVE * 32]);
}

// ---- survivor compaction between tick phases of the evaluate path -------------------------------------------------
// plan: every lane whose episode ended (dead, tick budget reached, or a capacity overflow) writes its result; every
// survivor gets a dense destination slot (order within the class is irrelevant for results).
__global__ void compact_plan_kernel(const float* state, const int* __restrict__ lane_creature, int n_lanes, int words, int max_ticks,
                                    double* fitness, int* ticks, int* alive, int* status, int* dst_slot, int* n_alive) {
    int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= n_lanes) return;
    int c = lane_creature[gl];
    int dst = -1;
    if (c >= 0) {
        const float* g = state + (size_t)(gl >> 5) * words * 32 + (gl & 31);
        int a = __float_as_int(g[S_ALIVE * 32]), t = __float_as_int(g[S_TICKS * 32]), st = __float_as_int(g[S_STATUS * 32]);
        if (!a || st || t >= max_ticks) {
            fitness[c] = __hiloint2double(__float_as_int(g[S_FIT_HI * 32]), __float_as_int(g[S_FIT_LO * 32]));
            ticks[c] = t; alive[c] = a; status[c] = st;
        } else dst = atomicAdd(n_alive, 1);
    }
    dst_slot[gl] = dst;
}
// copy: one CTA per source lane moves the creature's column into its slot of the other (ping-pong) buffer
__global__ void compact_copy_kernel(const float* src, float* dst, const int* __restrict__ dst_slot, const int* __restrict__ src_creature,
                                    int* dst_creature, int words) {
    int gl = blockIdx.x;
    int d = dst_slot[gl];
    if (d < 0) return;
    const float* s = src + (size_t)(gl >> 5) * words * 32 + (gl & 31);
    float* o = dst + (size_t)(d >> 5) * words * 32 + (d & 31);
    for (int w = threadIdx.x; w < words; w += blockDim.x) o[w * 32] = s[w * 32];
    if (threadIdx.x == 0) dst_creature[d] = src_creature[gl];
}
// pad: the unused lanes of the last destination batch become empty lanes
__global__ void compact_pad_kernel(float* dst, int* dst_creature, int words, const int* n_alive) {
    int n = *n_alive;
    if (n == 0 || (n & 31) == 0) return;
    int d = n + threadIdx.x;
    if ((d >> 5) != ((n - 1) >> 5)) return;
    float* o = dst + (size_t)(d >> 5) * words * 32 + (d & 31);
    o[S_NB * 32] = __int_as_float(0);
    o[S_ALIVE * 32] = __int_as_float(0);
    dst_creature[d] = -1;
}

// FP32 issue-rate microbenchmark: 8 independent multiply / add chains per thread. The step kernel is built with
// -fmad=false (parity), so its ceiling is the non-fused FMUL/FADD issue rate: 1 FLOP per lane per cycle.
__global__ void fp32_issue_kernel(float* out, int iters, float b, float c) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * b; a1 = a1 * b; a2 = a2 * b; a3 = a3 * b; a4 = a4 * b; a5 = a5 * b; a6 = a6 * b; a7 = a7 * b;
        a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------ handle
// Device buffers are grow-only and reused across uploads: re-allocating ~1 GB of state blocks on every rem2d_evaluate
// call cost ~0.4 s per generation.
#define N_COUNTER_WORDS (REM2D_N_COUNTERS + 12 * 16)      // work counters + phase cycles of diagnostic builds (REM2D_PHASE_TIMING)
struct Buf { void* p = nullptr; size_t cap = 0; };
static cudaError_t ensure(Buf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return cudaSuccess;
    // a buffer that has to grow will grow again (an evolving population shifts towards the large classes generation by
    // generation): 50 % headroom then, 12.5 % on the first allocation
    const bool regrow = b.p != nullptr;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + (regrow ? bytes / 2 : bytes / 8) + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

struct ClassState {
    std::vector<int> lane_creature;     // host copy: [n_batches*32], -1 = padding lane
    int n_batches = 0;
    int n_members = 0;                  // creatures of this class (lane_creature[0..n_members) are real)
    int episode_grid = 0;               // resident warps of the persistent episode kernel
    int grid2 = 0, gs2 = 0;             // second, wider-grouped launch for the creatures that do not fit in the first round
    cudaStream_t stream2 = nullptr;
    cudaEvent_t done2 = nullptr;
    ParkPolicy park{};
    double work = 0.0;                  // sum of the per-creature cost estimates (grid sizing)
    float* d_state = nullptr;
    int* d_lane_creature = nullptr;
    int* d_queue = nullptr;
    Buf b_state, b_state2, b_lc, b_lcw0, b_lcw1, b_dst, b_small, b_trace, b_ttrace;   // backing storage (grow-only)
    Buf b_redo_order, b_redo_slots;     // promotion re-runs (grow-only)
    // phased evaluation with survivor compaction (ping-pong buffers)
    float* d_state2 = nullptr;
    int* d_lc_work[2] = {nullptr, nullptr};   // lane -> creature maps of the compacted phases (d_lane_creature stays the static map)
    const int* lc_cur = nullptr;
    int lc_next = 0;
    int done_ticks = 0;
    int* d_dst_slot = nullptr;
    int* d_n_alive = nullptr;
    int* h_n_alive = nullptr;            // pinned
    int cur_lanes = 0;                  // lanes (padded to 32) of the current phase
    int phase = 0;
    bool active = false, pending = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t t_begin = nullptr, t_end = nullptr;     // per-class timeline of the last rem2d_run_episodes (diagnostics)
};

// Tuning / diagnostic options (rem2d_set_option; the REM2D_* environment variables of the same names are read ONCE, at
// rem2d_create, as initial values - tools/sweep_groups.py, tools/ea_pop_sweep.py and the tests use them). Defaults are the measured optimum.
struct Options {
    int warp_mode_max = -1;      // largest population that runs one warp per creature from tick 0 (-1: 48 per SM)
    int park_ticks = -1;         // park threshold of the queue mode (-1: automatic, 0: never park)
    double park_cap = -1.0;      // fraction of a class that may be parked (-1: automatic)
    int priority_mode = 1;       // rem2d_set_priority: 0 = expected lifetime >= 130 ticks first, 1 = and among those the longest first
    int overflow_wave = 1;       // queue the CTAs a class could not seat in its first wave behind the first launches (see choose_groups_and_grids)
    int image = -1;              // kernel image of the episode launches (-1: automatic, 0: ~200 registers, 1: 128 registers)
    int park_late_ticks = -1;    // park threshold of creatures pulled after the first round (-1: same as park_ticks)
    int park_lead = 0;           // park a creature as soon as its root is where the wall of death will be at the park threshold.
                                 // Measured and rejected as default: more creatures are parked, and early, while the GPU is still full -
                                 // the tail launches wait 100-300 ms for resources (950-1200 ms against 800 ms)
    double smem_budget_kb = 227.0, small_weight = 1.0;
    double wide_weight = 1.0;    // weight of the wide class next to one-lane classes in a throughput-bound mix (first wave)
    int min_class = 0;
    int group_shift = -1;        // log2 lanes per creature in the queue / step kernels (-1: per class default)
    int class_gs[N_CLASSES] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};   // per capacity class ("class_gs_<k>", -1: default)
    int tail_group_shift = 5;    // log2 lanes per creature of the tail launches (few creatures are parked by default: a warp each)
    int second_group_shift = -1; // log2 lanes per creature of a second launch for the creatures a large class cannot seat in its first
                                 // round (-1: off = refill the first launch's lanes). Measured and rejected as default: the GPU is
                                 // still throughput-bound when the first round ends, wide groups only add issue load (+25 % run time)
    int trace = 0;
    int phased = 0;              // tick phases with survivor compaction instead of the persistent queue kernel
};

struct rem2d_handle {
    rem2d_config cfg;
    Options opt;
    int cur_gs[N_CLASSES] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     // lanes per creature (log2) of the queue mode, per class
    cudaStream_t user_stream = nullptr;
    Terrain* d_ter = nullptr;
    Consts* d_consts = nullptr;
    unsigned long long* d_counters = nullptr;
    // Tail-mode launches are made on demand, each on a stream of this pool that is idle at that moment: a tail launch lives as
    // long as its longest creature (hundreds of ms), so a second kernel queued behind it on the same stream would wait for
    // it (and block its hardware queue for other streams). The pool grows when no stream is idle.
    std::vector<cudaStream_t> tail_pool;
    std::vector<char> tail_used;         // streams that received work during the current rem2d_run_episodes
    cudaStream_t poll_stream = nullptr;
    cudaEvent_t pool_done = nullptr;
    int* h_poll = nullptr;               // pinned [N_CLASSES]
    bool have_terrain = false, have_pop = false;
    bool state_valid = false;           // per-creature state blocks hold a consistent snapshot (reset/step path)
    bool results_valid = false;         // d_fitness/d_ticks/... were written by the episode kernel
    int n_sms = 0;
    int max_warps_per_sm[2] = {8, 16};  // resident warps of the two episode kernel images per SM by registers (occupancy API)
    double first_wave_frac = 1.0;       // share of the shared memory the first-wave grids are sized for
    int prio_high = 0;                  // greatest stream priority of the device (tail pool)
    bool overflow_wave = false;         // the second launches are overflow waves of the first (same width, both refill)
    int image = 0;                      // kernel image of the current population (0: 200 registers, 1: 128 registers)
    int n_edges = 0;
    // population
    int n_creatures = 0, n_bodies = 0, n_joints = 0;
    std::vector<int32_t> body_off;
    std::vector<float> priority;         // rem2d_set_priority: expected lifetimes for the next upload (empty: none)
    std::vector<int> creature_class, creature_lane;     // lane index within the class (batch*32+lane)
    Buf d_pop_mem[16];
    Buf b_results, b_roots;
    DevPop dpop{};
    ClassState cls[N_CLASSES];
    double* d_fitness = nullptr; int *d_ticks = nullptr, *d_alive = nullptr, *d_status = nullptr;
    std::vector<double> h_fitness; std::vector<int> h_ticks, h_alive, h_status;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_fork = nullptr;
    float last_ms = 0.0f;
    int64_t launches = 0;
    std::string err;
};

static thread_local std::string g_create_err;



#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return REM2D_E_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

// group shift of class k for the QUEUE mode of the uploaded population (choose_groups_and_grids)
static int class_gs(const rem2d_handle* h, int k) { return h->cur_gs[k]; }
// group shift of the stepping kernels (rem2d_reset / rem2d_step: static creature -> column mapping, no residency limit): the
// option if given, else as wide as keeps all warps of the population resident at once (a single creature gets a whole warp)
static int step_gs(const rem2d_handle* h, int k) {
    const int o = h->opt.group_shift >= 0 ? h->opt.group_shift : h->opt.class_gs[k];
    if (o >= 0) return std::min(5, o);
    int total = 0;
    for (int q = 0; q < N_CLASSES; ++q) total += h->cls[q].n_batches;
    int gs = 0;
    while (gs < 5 && (total << (gs + 1)) <= h->n_sms * 8) ++gs;
    return g_classes(k).nb <= 2 ? 0 : gs;
}
static bool set_option(Options& o, const char* name, double v) {
    const std::string n(name);
    if (n == "warp_mode_max") o.warp_mode_max = (int)v;
    else if (n == "park_ticks") o.park_ticks = (int)v;
    else if (n == "park_cap") o.park_cap = v;
    else if (n == "park_late_ticks") o.park_late_ticks = (int)v;
    else if (n == "park_lead") o.park_lead = (int)v;
    else if (n == "smem_budget_kb") o.smem_budget_kb = v;
    else if (n == "small_weight") o.small_weight = v;
    else if (n == "wide_weight") o.wide_weight = v;
    else if (n == "min_class") o.min_class = std::max(0, std::min(N_CLASSES - 1, (int)v));
    else if (n == "group_shift") o.group_shift = (int)v;
    else if (n == "tail_group_shift") o.tail_group_shift = std::max(0, std::min(5, (int)v));
    else if (n == "second_group_shift") o.second_group_shift = std::max(-1, std::min(5, (int)v));
    else if (n == "trace") o.trace = (int)v;
    else if (n == "phased") o.phased = (int)v;
    else if (n == "image") o.image = std::max(-1, std::min(1, (int)v));
    else if (n == "overflow_wave") o.overflow_wave = (int)v != 0;
    else if (n == "priority_mode") o.priority_mode = (int)v;
    else if (n.rfind("class_gs_", 0) == 0 && n.size() == 10 && n[9] >= '0' && n[9] < '0' + N_CLASSES) o.class_gs[n[9] - '0'] = (int)v;
    else return false;
    return true;
}
static void options_from_env(Options& o) {
    static const char* names[] = {"warp_mode_max", "park_ticks", "park_cap", "park_late_ticks", "park_lead", "smem_budget_kb", "small_weight", "min_class",
                                  "group_shift", "tail_group_shift", "second_group_shift", "trace", "image", "overflow_wave", "wide_weight", "priority_mode"};
    for (const char* n : names) {
        std::string env = "REM2D_";
        for (const char* q = n; *q; ++q) env += (char)toupper(*q);
        if (const char* e = getenv(env.c_str())) set_option(o, n, atof(e));
    }
    if (const char* e = getenv("REM2D_EPISODE_MODE")) o.phased = !strcmp(e, "phased");
    if (const char* e = getenv("REM2D_CLASS_GS"))          // "0,0,1,2,2,2,2,3,3": group shift per capacity class
        for (int k = 0; k < N_CLASSES && *e; ++k) { o.class_gs[k] = atoi(e); while (*e && *e != ',') ++e; if (*e == ',') ++e; }
}

static void free_population(rem2d_handle* h) {          // logical reset; the buffers stay allocated for reuse
    for (auto& c : h->cls) { c.n_batches = 0; c.n_members = 0; c.lane_creature.clear(); }
    h->have_pop = false; h->state_valid = false; h->results_valid = false;
}
static void release_buffers(rem2d_handle* h) {
    for (auto& b : h->d_pop_mem) { if (b.p) cudaFree(b.p); b = Buf(); }
    for (auto& c : h->cls) {
        Buf* bufs[] = {&c.b_state, &c.b_state2, &c.b_lc, &c.b_lcw0, &c.b_lcw1, &c.b_dst, &c.b_small, &c.b_trace, &c.b_ttrace,
                       &c.b_redo_order, &c.b_redo_slots};
        for (Buf* b : bufs) { if (b->p) cudaFree(b->p); *b = Buf(); }
        if (c.h_n_alive) cudaFreeHost(c.h_n_alive);
        c.h_n_alive = nullptr;
    }
    if (h->b_results.p) cudaFree(h->b_results.p);
    h->b_results = Buf();
    if (h->b_roots.p) cudaFree(h->b_roots.p);
    h->b_roots = Buf();
}

extern "C" {

void rem2d_default_config(rem2d_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->dt = (float)(1.0 / 50);        // Modular2DEnv.py:26,634
    cfg->velocity_iterations = 180;     // Modular2DEnv.py:634
    cfg->position_iterations = 60;
    cfg->gravity_y = -10.0f;
    cfg->module_friction = (float)0.1;  // simple_module.py:289
    cfg->terrain_friction = 2.5f;       // Modular2DEnv.py:62
    cfg->p_gain = 1.9;                  // Modular2DEnv.py:601
    cfg->wod_speed = 0.04;              // Modular2DEnv.py:52
    cfg->env_length = 100.0;            // REM2D_main.py:350
    cfg->evaluation_steps = 10000;      // REM2D_main.py:350
    cfg->continuous = 1;
    cfg->allow_sleep = 1;
    cfg->terminate = 1;
    cfg->device = 0;
    cfg->stream = nullptr;
    cfg->sincos_mode = 0;
}
int rem2d_abi_version(void) { return REM2D_ABI_VERSION; }
const char* rem2d_backend(void) { return "cuda-sm_100a"; }

int rem2d_create(const rem2d_config* cfg, rem2d_handle** out) {
    if (!cfg || !out) { g_create_err = "rem2d_create: NULL argument"; return REM2D_E_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("rem2d_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return REM2D_E_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "rem2d_create: bad device ordinal"; return REM2D_E_INVALID; }
    if (cfg->sincos_mode != 0) { g_create_err = "rem2d_create: sincos_mode != 0 is an oracle-only option"; return REM2D_E_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return REM2D_E_CUDA; }
    rem2d_handle* h = new rem2d_handle();
    h->cfg = *cfg;
    h->user_stream = (cudaStream_t)cfg->stream;
    options_from_env(h->opt);
    // every failure path releases what was created so far (rem2d_destroy copes with a partially built handle)
    auto fail = [&](const char* what, cudaError_t err) { g_create_err = std::string(what) + ": " + cudaGetErrorString(err); rem2d_destroy(h); return REM2D_E_CUDA; };
    if ((e = cudaMalloc(&h->d_ter, sizeof(Terrain))) != cudaSuccess) return fail("cudaMalloc terrain", e);
    if ((e = cudaMalloc(&h->d_consts, sizeof(Consts))) != cudaSuccess) return fail("cudaMalloc consts", e);
    if ((e = cudaMalloc(&h->d_counters, sizeof(unsigned long long) * N_COUNTER_WORDS)) != cudaSuccess) return fail("cudaMalloc counters", e);
    if ((e = cudaMemset(h->d_counters, 0, sizeof(unsigned long long) * N_COUNTER_WORDS)) != cudaSuccess) return fail("cudaMemset counters", e);
    Consts k = make_consts(cfg);
    if ((e = cudaMemcpy(h->d_consts, &k, sizeof(k), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy consts", e);
    for (auto& c : h->cls) {
        if ((e = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&c.done2, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaEventCreate(&c.t_begin)) != cudaSuccess || (e = cudaEventCreate(&c.t_end)) != cudaSuccess) return fail("cudaEventCreate", e);
    }
    if ((e = cudaEventCreate(&h->ev_start)) != cudaSuccess || (e = cudaEventCreate(&h->ev_stop)) != cudaSuccess) return fail("cudaEventCreate", e);
    {   // tail launches and the polling copies run at the device's greatest stream priority: their CTAs / copies go ahead of
        // the queued overflow-wave CTAs of the episode launches
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) h->prio_high = greatest;
        else cudaGetLastError();
    }
    h->tail_pool.assign(64, nullptr);
    for (auto& st : h->tail_pool)
        if ((e = cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, h->prio_high)) != cudaSuccess) return fail("cudaStreamCreate (tail pool)", e);
    if ((e = cudaStreamCreateWithPriority(&h->poll_stream, cudaStreamNonBlocking, h->prio_high)) != cudaSuccess) return fail("cudaStreamCreate (poll)", e);
    if ((e = cudaEventCreateWithFlags(&h->pool_done, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaMallocHost(&h->h_poll, sizeof(int) * 16)) != cudaSuccess) return fail("cudaMallocHost", e);
    if ((e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("cudaEventCreate", e);
    cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    {
        int carve = cudaSharedmemCarveoutMaxShared, max_hot = 0;
        if (const char* ev = getenv("REM2D_CARVEOUT")) carve = atoi(ev);      // experiment: percent of the unified L1/shared array
        for (int q = 0; q < N_CLASSES; ++q)
            for (int gs = 0; gs <= 5; ++gs) max_hot = std::max(max_hot, g_classes(q).hot_bytes(gs));
        if ((e = rem2d_set_kernel_attributes(max_hot, carve)) != cudaSuccess) return fail("cudaFuncSetAttribute", e);
        for (int im = 0; im < 2; ++im) h->max_warps_per_sm[im] = std::max(1, rem2d_episode_blocks_per_sm(im, 1024));
    }
    *out = h;
    return REM2D_OK;
}

int rem2d_set_option(rem2d_handle* h, const char* name, double value) {
    if (!h || !name) return REM2D_E_INVALID;
    if (!set_option(h->opt, name, value)) { h->err = std::string("set_option: unknown option ") + name; return REM2D_E_INVALID; }
    return REM2D_OK;
}

int rem2d_destroy(rem2d_handle* h) {
    if (!h) return REM2D_E_INVALID;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    free_population(h);
    release_buffers(h);
    for (auto& c : h->cls) {
        if (c.stream) cudaStreamDestroy(c.stream);
        if (c.stream2) cudaStreamDestroy(c.stream2);
        if (c.done) cudaEventDestroy(c.done);
        if (c.done2) cudaEventDestroy(c.done2);
        if (c.t_begin) cudaEventDestroy(c.t_begin);
        if (c.t_end) cudaEventDestroy(c.t_end);
    }
    for (auto& st : h->tail_pool) if (st) cudaStreamDestroy(st);
    if (h->poll_stream) cudaStreamDestroy(h->poll_stream);
    if (h->pool_done) cudaEventDestroy(h->pool_done);
    if (h->h_poll) cudaFreeHost(h->h_poll);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->d_ter) cudaFree(h->d_ter);
    if (h->d_consts) cudaFree(h->d_consts);
    if (h->d_counters) cudaFree(h->d_counters);
    cudaGetLastError();
    delete h;
    return REM2D_OK;
}

const char* rem2d_last_error(rem2d_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int rem2d_set_terrain(rem2d_handle* h, const double* y, int32_t n, double step) {
    if (!h) return REM2D_E_INVALID;
    if (!y || n < 2 || n > RB_MAX_EDGES) { h->err = "set_terrain: need 2..200 vertices"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    Terrain* t = new Terrain();
    fill_terrain(t, y, n, step);
    cudaError_t e = cudaMemcpy(h->d_ter, t, sizeof(Terrain), cudaMemcpyHostToDevice);
    delete t;
    if (e != cudaSuccess) { h->err = std::string("set_terrain: ") + cudaGetErrorString(e); return REM2D_E_CUDA; }
    h->n_edges = n - 1;
    h->have_terrain = true;
    return REM2D_OK;
}

}  // extern "C"

template <typename T>
static cudaError_t upload_array(rem2d_handle* h, int slot, const T* src, size_t n, const T** dst) {
    cudaError_t e = ensure(h->d_pop_mem[slot], (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    void* d = h->d_pop_mem[slot].p;
    if (n) e = cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, h->user_stream);
    *dst = (const T*)d;
    return e;
}

static int launch_reset(rem2d_handle* h);

// Resident warps of the persistent episode kernels for the current group shifts (h->cur_gs). All classes run concurrently, so
// the shared memory of the SMs (227 KB each) and the resident-warp limit set by the register file are divided among them in
// proportion to their work; a class never gets more warps than it has creatures, and what it cannot use is handed to the
// others. Without this the largest class would occupy every SM until its last creature dies and the remaining classes would
// run after it. Returns true iff EVERY creature of every class has a group from the start (one round, under-filled GPU).
static bool size_grids(rem2d_handle* h, double warp_frac = 0.97) {
    const double smem_kb = h->opt.smem_budget_kb, small_weight = h->opt.small_weight;
    double budget = (double)h->n_sms * smem_kb * 1024.0 * 0.98 * h->first_wave_frac;
    double work[N_CLASSES], smem[N_CLASSES];
    bool fixed[N_CLASSES];
    for (int k = 0; k < N_CLASSES; ++k) {
        const int gs = h->cur_gs[k];
        work[k] = h->cls[k].work; fixed[k] = h->cls[k].n_members == 0; smem[k] = g_classes(k).hot_bytes(gs) + 1024.0;
        if (g_classes(k).nb <= 8) work[k] *= small_weight;
        // a group of G lanes is one resident creature: G times the warps for the same residency, and the creature ticks
        // p(G) times faster (measured tick latencies, profiles/r2_timeline_*.txt)
        static const double speedup[6] = {1.0, 1.15, 1.5, 1.9, 2.2, 2.5};
        work[k] *= (double)(1 << gs) / speedup[gs];
        if (gs > 0 && h->first_wave_frac < 1.0) work[k] *= h->opt.wide_weight;
        h->cls[k].episode_grid = 0;
    }
    // warps_k = W * work_k with W such that sum_k warps_k * smem_k = budget: every class then needs about the same
    // number of sequential creature-lifetimes per group times its own tick latency, i.e. the classes finish together.
    // The register file bounds the resident warps as well (h->max_warps_per_sm, from the occupancy API): a launch whose
    // CTAs do not fit waits for the CTAs of the classes launched before it to EXIT, which serialises the classes
    // (measured: the small classes started 400 ms late when 9 warps per SM fit and the grids asked for 9.7).
    double warp_budget = (double)h->n_sms * h->max_warps_per_sm[h->image] * warp_frac;
    bool all_fit = true;
    for (int round = 0; round <= N_CLASSES; ++round) {
        double denom = 0.0, wsum = 0.0;
        for (int k = 0; k < N_CLASSES; ++k) if (!fixed[k]) { denom += work[k] * smem[k]; wsum += work[k]; }
        if (denom <= 0.0) break;
        const double W = std::max(0.0, std::min(budget / denom, warp_budget / wsum));
        bool changed = false;
        for (int k = 0; k < N_CLASSES; ++k) {
            if (fixed[k]) continue;
            const int per = 32 >> h->cur_gs[k], need = (h->cls[k].n_members + per - 1) / per;    // one group per creature
            if ((int)(W * work[k]) >= need) {        // the class fits entirely: fix it and give the rest back
                h->cls[k].episode_grid = need;
                budget -= need * smem[k];
                warp_budget -= need;
                fixed[k] = true; changed = true;
            }
        }
        if (!changed) {
            for (int k = 0; k < N_CLASSES; ++k)
                if (!fixed[k]) { h->cls[k].episode_grid = std::max(1, (int)(W * work[k])); all_fit = false; }
            break;
        }
    }
    for (int k = 0; k < N_CLASSES; ++k)
        if (h->cls[k].n_members && h->cls[k].episode_grid < 1) h->cls[k].episode_grid = 1;
    return all_fit;
}

// Lanes per creature of every class for the uploaded population, then the grids. Throughput and latency pull in opposite
// directions: one lane per creature issues the fewest instructions per creature-tick, a group of G lanes ticks a creature
// up to ~2x faster but spends more issue slots on it (idle lanes in the sweeps, leader sections). Measured on B200
// (tools/sweep_groups.py): with 65536 creatures the GPU is throughput-bound and G = 1 everywhere is fastest; when the
// population leaves the GPU under-filled (every creature resident from the start) the run time is one creature lifetime
// times the tick latency, and wider groups win. So: start from G = 1 and widen the groups of the large classes (up to the
// class table's value) as long as every creature of every class still has a group from the start. Explicit options
// ("group_shift", "class_gs_<k>") override the choice.
static void choose_groups_and_grids(rem2d_handle* h) {
    bool forced[N_CLASSES];
    for (int k = 0; k < N_CLASSES; ++k) {
        const int o = h->opt.group_shift >= 0 ? h->opt.group_shift : h->opt.class_gs[k];
        forced[k] = o >= 0;
        h->cur_gs[k] = forced[k] ? std::min(5, o) : 0;
    }
    // under-filled GPU <=> everything fits with one lane per creature in the 128-register image: then widen (below)
    h->image = 1;
    h->first_wave_frac = 1.0;
    bool all_fit = size_grids(h);
    if (!all_fit) {
        // Throughput-bound: one lane per creature - except the largest class (33-44 bodies): its 4.6 KB of solver state per
        // creature make a 32-creature CTA 148 KB of shared memory, ONE resident warp per SM at a 5 ms tick. 8 lanes per creature
        // (4 creatures, 19 KB per CTA) pack the same residency into warps that tick twice as fast and fit beside the other
        // classes' CTAs. Measured on EA-configured populations (max_size 40; tools/ea_pop_sweep.py, profiles/r2_ea_pop_sweep.txt):
        // 1.5-1.6x on the whole evaluation (1946 -> 1271 ms, 3311 -> 2112, 5925 -> 3676); widening the 23-32 body class as well
        // gains nothing, the smaller classes lose (the bench population, <= 21 bodies, is unaffected).
        for (int k = 0; k < N_CLASSES; ++k)
            if (!forced[k]) h->cur_gs[k] = g_classes(k).nb > 32 ? 3 : 0;
        h->image = h->opt.image >= 0 ? h->opt.image : 0;
        size_grids(h);
        // Mixed widths: the cost model's tick latencies are those of one-lane warps on moderately loaded SMs; next to
        // hundreds of 8-lane warps the one-lane classes tick 2x slower than modelled (4.8 ms against 1.55 ms for the wide
        // class) and a first wave that asks for all of the shared memory does not even fit (736 of 822 CTAs resident,
        // the launches behind them held back). With an overflow wave it is better to seat 75 % statically and let the
        // queued CTAs take the rest as it frees up: 3560 -> 2610 ms, 1930 -> 1600, 1155 -> 1130 ms on the three
        // EA-configured populations; all-one-lane populations (the bench) lose 4 % with it and keep the full first wave.
        bool mixed = false;
        for (int k = 0; k < N_CLASSES; ++k) mixed |= h->cls[k].n_members > 0 && h->cur_gs[k] > 0;
        if (mixed && h->opt.overflow_wave && h->opt.second_group_shift < 0 && h->opt.smem_budget_kb > 226.0) {
            h->first_wave_frac = 0.75;
            size_grids(h);
        }
    }
    for (int k = 0; k < N_CLASSES; ++k) { h->cls[k].grid2 = 0; h->cls[k].gs2 = 0; }
    h->overflow_wave = false;
    if (!all_fit) {
        // Throughput-bound population: one lane per creature, lanes refilled from the class queue. OPTION "second_group_shift"
        // (off by default): the creatures of the large classes that do not get a lane in the first round are run by a second
        // launch with wide groups (2-4x lower tick latency) whose CTAs become resident as the first launch's warps exit,
        // instead of one more lap at the bulk tick latency. Measured on the bench population: the GPU is still
        // throughput-bound when the first round ends (the small classes run until ~550 ms), so the wide groups only add issue
        // load: 1050-1200 ms against 815-880 ms with refill.
        const int gs2 = h->opt.second_group_shift;
        for (int k = 0; k < N_CLASSES && gs2 >= 0; ++k) {
            ClassState& cs = h->cls[k];
            const int per1 = 32 >> h->cur_gs[k];
            const int left = cs.n_members - cs.episode_grid * per1;
            if (left <= 0 || g_classes(k).nb < 12 || forced[k]) continue;
            cs.gs2 = std::max(gs2, h->cur_gs[k]);
            cs.grid2 = (left + (32 >> cs.gs2) - 1) / (32 >> cs.gs2);
        }
        // OVERFLOW WAVE (default): the first-wave grids above come from a static cost model; when it is off for a population -
        // measured on an EA-configured one: the 33-44 body class done after 1.9 s, the others at 3.1-3.6 s, with the freed 18 MB
        // of shared memory idle in between (profiles/r2_timeline_ea_gen2_before.txt) - nothing could use the resources of a
        // class that has finished. So every class also queues, BEHIND all first launches, the CTAs it could not seat: the
        // work distributor places them as earlier CTAs exit, they pull from the same class queue, and a CTA that finds its
        // queue empty exits at once. The hardware does the balancing.
        if (gs2 < 0 && h->opt.overflow_wave) {
            h->overflow_wave = true;
            for (int k = 0; k < N_CLASSES; ++k) {
                ClassState& cs = h->cls[k];
                const int per = 32 >> h->cur_gs[k];
                const int left = cs.n_members - cs.episode_grid * per;
                if (!cs.n_members || left <= 0) continue;
                cs.gs2 = h->cur_gs[k];
                cs.grid2 = (left + per - 1) / per;
            }
        }
        return;
    }
    // widest uniform group width (capped per class) with which every creature has a group from the start AND the warps fill
    // at most 60 % of the resident-warp limit: wide groups are issue-bound before the SMs are full. Measured optimum
    // (tools/small_pop.py): 16384 creatures -> 4 lanes, 8192 -> 8, 4096 -> 8-16, <= 1024 -> 32
    for (int gs = 5; gs >= 1; --gs) {
        for (int k = 0; k < N_CLASSES; ++k)
            if (!forced[k]) h->cur_gs[k] = std::min(gs, g_classes(k).gs);
        if (size_grids(h, 0.6)) { size_grids(h); return; }
    }
    for (int k = 0; k < N_CLASSES; ++k)
        if (!forced[k]) h->cur_gs[k] = 0;
    size_grids(h);
}

static int upload_impl(rem2d_handle* h, const rem2d_population* pop, bool do_reset) {
    if (!h || !pop) return REM2D_E_INVALID;
    if (pop->n_creatures < 0 || pop->n_joints != pop->n_bodies - pop->n_creatures) { h->err = "upload: inconsistent counts"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    free_population(h);
    const int n = pop->n_creatures;
    h->n_creatures = n; h->n_bodies = pop->n_bodies; h->n_joints = pop->n_joints;
    h->body_off.assign(pop->body_off, pop->body_off + n + 1);
    // classify + validate
    h->creature_class.assign(n, -1);
    h->creature_lane.assign(n, -1);
    std::vector<std::vector<int>> members(N_CLASSES);
    const int min_class = h->opt.min_class;      // experiment: force small creatures into a larger class
    for (int c = 0; c < n; ++c) {
        int nb = pop->body_off[c + 1] - pop->body_off[c];
        if (nb < 1) { h->err = "upload: creature without a root body"; return REM2D_E_INVALID; }
        int k = -1;
        for (int q = min_class; q < N_CLASSES; ++q) if (nb <= g_classes(q).nb) { k = q; break; }
        if (k < 0) { char buf[128]; snprintf(buf, sizeof(buf), "upload: creature %d has %d bodies (> %d supported)", c, nb, g_classes(N_CLASSES - 1).nb); h->err = buf; return REM2D_E_CAPACITY; }
        int j0 = pop->body_off[c] - c;
        for (int j = 0; j < nb - 1; ++j)
            if (pop->joint_parent[j0 + j] < 0 || pop->joint_parent[j0 + j] > j) { h->err = "upload: joint parent must precede its child"; return REM2D_E_INVALID; }
        h->creature_class[c] = k;
        members[k].push_back(c);
    }
    std::vector<uint8_t> order((size_t)std::max(pop->n_joints, 1), 0);
    for (int c = 0; c < n; ++c) {
        int nb = pop->body_off[c + 1] - pop->body_off[c], j0 = pop->body_off[c] - c;
        island_joint_order(nb, pop->joint_parent + j0, order.data() + j0);
    }
    size_t nbod = (size_t)pop->n_bodies, nj = (size_t)pop->n_joints;
    CK(upload_array(h, 0, pop->body_off, (size_t)n + 1, &h->dpop.body_off));
    CK(upload_array(h, 1, pop->shape, nbod, &h->dpop.shape));
    CK(upload_array(h, 2, pop->hx, nbod, &h->dpop.hx));
    CK(upload_array(h, 3, pop->hy, nbod, &h->dpop.hy));
    CK(upload_array(h, 4, pop->x0, nbod, &h->dpop.x0));
    CK(upload_array(h, 5, pop->y0, nbod, &h->dpop.y0));
    CK(upload_array(h, 6, pop->a0, nbod, &h->dpop.a0));
    CK(upload_array(h, 7, pop->joint_parent, nj, &h->dpop.joint_parent));
    CK(upload_array(h, 8, pop->anchor_a, nj * 2, &h->dpop.anchor_a));
    CK(upload_array(h, 9, pop->anchor_b, nj * 2, &h->dpop.anchor_b));
    CK(upload_array(h, 10, pop->lower, nj, &h->dpop.lower));
    CK(upload_array(h, 11, pop->upper, nj, &h->dpop.upper));
    CK(upload_array(h, 12, pop->max_torque, nj, &h->dpop.max_torque));
    CK(upload_array(h, 13, pop->ctrl, nbod * 5, &h->dpop.ctrl));
    CK(upload_array(h, 14, (const uint8_t*)order.data(), nj, &h->dpop.joint_order));
    // the pageable source buffers (order, caller arrays) must stay valid until the copies ran
    CK(cudaStreamSynchronize(h->user_stream));
    for (int k = 0; k < N_CLASSES; ++k) {
        auto& m = members[k];
        if (m.empty()) continue;
        // big creatures first: batches of similar size limit lane divergence, and the costly batches start early; creatures
        // the caller expects to be long-lived (rem2d_set_priority: >= 130 ticks, i.e. they outrun the wall of death) go to
        // the very front of the class queue, the longest-lived first in buckets of 32 ticks (longest-processing-time-first).
        // Measured with the parents' lifetimes as the hint (profiles/r2_ea_pop_sweep.txt): selected populations, where most
        // creatures are long-lived, evaluate 7-11 % faster with the ordering than with the threshold alone (1570 -> 1480,
        // 2575 -> 2380, 2130 -> 1900 ms); a random population with a perfect hint gains 20 % either way (800 -> 640 ms).
        const bool prio = (int)h->priority.size() == n;
        // "long-lived" = expected to be alive when the wall of death has passed the start pad (root at x = 5: tick 125 at the
        // reference's 0.04 per tick, + 4 %)
        const float long_lived = (h->cfg.terminate && h->cfg.wod_speed > 0.0) ? (float)(5.2 / h->cfg.wod_speed) : 130.0f;
        std::stable_sort(m.begin(), m.end(), [&](int a, int b) {
            int la = prio && h->priority[a] >= long_lived, lb = prio && h->priority[b] >= long_lived;
            if (prio && h->opt.priority_mode == 1) {      // ... and among those the longest expected lifetime first, in buckets of 32 ticks
                la = la ? 1 + ((int)std::min(h->priority[a], 8192.0f) >> 5) : 0;
                lb = lb ? 1 + ((int)std::min(h->priority[b], 8192.0f) >> 5) : 0;
            }
            if (la != lb) return la > lb;
            return (pop->body_off[a + 1] - pop->body_off[a]) > (pop->body_off[b + 1] - pop->body_off[b]);
        });
        ClassState& cs = h->cls[k];
        cs.n_batches = (int)((m.size() + 31) / 32);
        cs.n_members = (int)m.size();
        cs.episode_grid = cs.n_batches;     // sized below once every class is known
        cs.lane_creature.assign((size_t)cs.n_batches * 32, -1);
        for (size_t i = 0; i < m.size(); ++i) { cs.lane_creature[i] = m[i]; h->creature_lane[m[i]] = (int)i; }
        const size_t lane_bytes = cs.lane_creature.size() * sizeof(int);
        const size_t state_bytes = (size_t)(cs.n_batches + 2) * g_classes(k).words * 32 * sizeof(float);   // (+2: second launch's columns)
        CK(ensure(cs.b_lc, lane_bytes)); cs.d_lane_creature = (int*)cs.b_lc.p;
        CK(cudaMemcpy(cs.d_lane_creature, cs.lane_creature.data(), cs.lane_creature.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(ensure(cs.b_state, state_bytes)); cs.d_state = (float*)cs.b_state.p;
        CK(ensure(cs.b_state2, state_bytes)); cs.d_state2 = (float*)cs.b_state2.p;
        CK(ensure(cs.b_lcw0, lane_bytes)); cs.d_lc_work[0] = (int*)cs.b_lcw0.p;
        CK(ensure(cs.b_lcw1, lane_bytes)); cs.d_lc_work[1] = (int*)cs.b_lcw1.p;
        CK(ensure(cs.b_dst, lane_bytes)); cs.d_dst_slot = (int*)cs.b_dst.p;
        CK(ensure(cs.b_small, 64)); cs.d_n_alive = (int*)cs.b_small.p; cs.d_queue = (int*)cs.b_small.p + 4;
        if (!cs.h_n_alive) CK(cudaMallocHost(&cs.h_n_alive, sizeof(int)));
    }
    {   // per-creature cost ~ tick latency of its size: measured ~0.3 ms + 0.085 ms per body for a resident warp; lone bodies
        // fall asleep after landing and cost almost nothing
        for (int k = 0; k < N_CLASSES; ++k) {
            h->cls[k].work = 0.0;
            for (int c : members[k]) { int nbc = pop->body_off[c + 1] - pop->body_off[c]; h->cls[k].work += nbc == 1 ? 1.0 : 3.5 + nbc; }
        }
        choose_groups_and_grids(h);
    }
    {
        const size_t nn = (size_t)std::max(n, 1);
        CK(ensure(h->b_results, nn * (sizeof(double) + 3 * sizeof(int))));
        h->d_fitness = (double*)h->b_results.p;
        h->d_ticks = (int*)(h->d_fitness + nn);
        h->d_alive = h->d_ticks + nn;
        h->d_status = h->d_alive + nn;
    }
    h->h_fitness.assign(n, 0.0); h->h_ticks.assign(n, 0); h->h_alive.assign(n, 0); h->h_status.assign(n, 0);
    h->have_pop = true;
    h->priority.clear();                 // the hint applies to one upload
    return do_reset ? launch_reset(h) : REM2D_OK;
}

extern "C" int rem2d_upload(rem2d_handle* h, const rem2d_population* pop) { return upload_impl(h, pop, true); }

// fork the class streams off the user stream / join them back
static int fork_streams(rem2d_handle* h) {
    CK(cudaEventRecord(h->ev_fork, h->user_stream));
    for (auto& c : h->cls) if (c.n_batches) { CK(cudaStreamWaitEvent(c.stream, h->ev_fork, 0)); CK(cudaStreamWaitEvent(c.stream2, h->ev_fork, 0)); }
    return REM2D_OK;
}
static int join_streams(rem2d_handle* h) {
    for (auto& c : h->cls) if (c.n_batches) {
        CK(cudaEventRecord(c.done, c.stream)); CK(cudaStreamWaitEvent(h->user_stream, c.done, 0));
        CK(cudaEventRecord(c.done2, c.stream2)); CK(cudaStreamWaitEvent(h->user_stream, c.done2, 0));
    }
    return REM2D_OK;
}

static int join_streams_all(rem2d_handle* h) {
    for (auto& c : h->cls) if (c.n_batches) { CK(cudaEventRecord(c.done, c.stream)); CK(cudaStreamWaitEvent(h->user_stream, c.done, 0)); }
    return REM2D_OK;
}

static int launch_reset(rem2d_handle* h) {
    if (!h->have_terrain) { h->err = "reset: no terrain set"; return REM2D_E_INVALID; }
    int rc = fork_streams(h);
    if (rc) return rc;
    for (int k = N_CLASSES - 1; k >= 0; --k) {
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        g_classes(k).reset(step_gs(h, k), cs.n_batches, cs.stream, cs.d_state, cs.d_lane_creature, h->dpop);
        h->launches++;
    }
    CK(cudaGetLastError());
    rc = join_streams(h);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * N_COUNTER_WORDS, h->user_stream));
    h->state_valid = true; h->results_valid = false;
    return REM2D_OK;
}

// Promotion: creatures that exceeded a capacity of their class (contact pool) are re-run from tick 0 on the refill
// episode kernel of the next larger class until they fit. Rare (very fine terrains, pile-ups).
static int promote_overflowed(rem2d_handle* h, int max_ticks) {
    const int n = h->n_creatures;
    std::vector<int> cls_of(h->creature_class);
    for (int round = 0; round < N_CLASSES; ++round) {
        CK(cudaMemcpyAsync(h->h_status.data(), h->d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, h->user_stream));
        CK(cudaStreamSynchronize(h->user_stream));
        std::vector<std::vector<int>> redo(N_CLASSES);
        bool any = false;
        for (int c = 0; c < n; ++c)
            if (h->h_status[c]) {
                if (cls_of[c] + 1 >= N_CLASSES) {
                    char buf[160];
                    snprintf(buf, sizeof(buf), "creature %d exceeds the capacities of the largest class (status %d)", c, h->h_status[c]);
                    h->err = buf;
                    return REM2D_E_CAPACITY;
                }
                cls_of[c] += 1;
                redo[cls_of[c]].push_back(c);
                any = true;
            }
        if (!any) break;
        for (int k = 0; k < N_CLASSES; ++k) {
            if (redo[k].empty()) continue;
            ClassState& cs = h->cls[k];
            const int gs = class_gs(h, k), per = 32 >> gs;
            const int grid = (int)((redo[k].size() + per - 1) / per), batches = (int)((redo[k].size() + 31) / 32);
            CK(ensure(cs.b_redo_order, sizeof(int) * (redo[k].size() + 1)));
            CK(ensure(cs.b_redo_slots, (size_t)batches * g_classes(k).words * 32 * sizeof(float)));
            int* d_order = (int*)cs.b_redo_order.p;
            int* d_queue = d_order + redo[k].size();
            CK(cudaMemcpyAsync(d_order, redo[k].data(), sizeof(int) * redo[k].size(), cudaMemcpyHostToDevice, h->user_stream));
            CK(cudaMemsetAsync(d_queue, 0, sizeof(int), h->user_stream));
            g_classes(k).episode(h->image, gs, grid, h->user_stream, (float*)cs.b_redo_slots.p, d_order, (int)redo[k].size(), d_queue, h->dpop, h->d_ter,
                                 h->d_consts, max_ticks, h->d_fitness, h->d_ticks, h->d_alive, h->d_status, h->d_counters,
                                 ParkPolicy{0, 0, 0.0f, 0, 0, 0, nullptr, nullptr}, nullptr, nullptr, nullptr, 1);
            h->launches++;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(h->user_stream));      // redo[] (pageable source of the copy) goes out of scope
        }
    }
    return REM2D_OK;
}

// Whole episodes in tick phases with survivor compaction (default evaluate path).
// Every class advances independently on its own stream: step_kernel(phase ticks) -> plan (harvest finished creatures,
// assign dense slots to survivors) -> copy columns into the other buffer -> the host reads the survivor count and
// launches the next, smaller phase. Almost every creature of a random population dies when the wall of death reaches
// the start pad (tick ~126), so the first phase is 128 ticks; afterwards the few survivors are repacked every 32 ticks
// instead of keeping their original warps alive at 1-3 live lanes out of 32.
static int launch_phased(rem2d_handle* h, int max_ticks) {
    if (!h->have_terrain) { h->err = "run_episodes: no terrain set"; return REM2D_E_INVALID; }
    CK(cudaEventRecord(h->ev_start, h->user_stream));
    int rc = launch_reset(h);           // builds every world into the static batches (also zeroes the counters)
    if (rc) return rc;
    rc = fork_streams(h);
    if (rc) return rc;
    int n_active = 0;
    for (int k = N_CLASSES - 1; k >= 0; --k) {
        ClassState& cs = h->cls[k];
        cs.active = cs.n_batches > 0; cs.pending = false; cs.phase = 0; cs.cur_lanes = cs.n_batches * 32;
        cs.lc_cur = cs.d_lane_creature; cs.lc_next = 0; cs.done_ticks = 0;
        if (cs.active) ++n_active;
    }
    const int first_phase = 128, next_phase = 32;
    auto enqueue = [&](int k) -> int {
        ClassState& cs = h->cls[k];
        const int words = g_classes(k).words;
        const int batches = cs.cur_lanes / 32;
        int ticks = cs.phase == 0 ? first_phase : next_phase;
        if (ticks > max_ticks - cs.done_ticks) ticks = max_ticks - cs.done_ticks;
        cs.done_ticks += ticks;
        int* lc_dst = cs.d_lc_work[cs.lc_next];
        g_classes(k).step(step_gs(h, k), batches, cs.stream, cs.d_state, ticks, h->d_ter, h->d_consts, h->d_counters);
        CK(cudaMemsetAsync(cs.d_n_alive, 0, sizeof(int), cs.stream));
        compact_plan_kernel<<<(cs.cur_lanes + 127) / 128, 128, 0, cs.stream>>>(cs.d_state, cs.lc_cur, cs.cur_lanes, words, max_ticks,
                                                                              h->d_fitness, h->d_ticks, h->d_alive, h->d_status,
                                                                              cs.d_dst_slot, cs.d_n_alive);
        compact_copy_kernel<<<cs.cur_lanes, 128, 0, cs.stream>>>(cs.d_state, cs.d_state2, cs.d_dst_slot, cs.lc_cur, lc_dst, words);
        compact_pad_kernel<<<1, 32, 0, cs.stream>>>(cs.d_state2, lc_dst, words, cs.d_n_alive);
        CK(cudaMemcpyAsync(cs.h_n_alive, cs.d_n_alive, sizeof(int), cudaMemcpyDeviceToHost, cs.stream));
        CK(cudaGetLastError());
        h->launches += 4;
        cs.pending = true;
        return REM2D_OK;
    };
    for (int k = N_CLASSES - 1; k >= 0; --k)
        if (h->cls[k].active) { rc = enqueue(k); if (rc) return rc; }
    while (n_active > 0) {
        bool progressed = false;
        for (int k = N_CLASSES - 1; k >= 0; --k) {
            ClassState& cs = h->cls[k];
            if (!cs.active || !cs.pending) continue;
            cudaError_t q = cudaStreamQuery(cs.stream);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) { h->err = std::string("phase: ") + cudaGetErrorString(q); return REM2D_E_CUDA; }
            progressed = true;
            cs.pending = false;
            std::swap(cs.d_state, cs.d_state2);
            cs.lc_cur = cs.d_lc_work[cs.lc_next];
            cs.lc_next ^= 1;
            int alive = *cs.h_n_alive;
            cs.phase++;
            // creatures that reached the tick budget were harvested by the plan kernel, so alive == 0 ends the class
            if (alive == 0) { cs.active = false; --n_active; continue; }
            cs.cur_lanes = ((alive + 31) / 32) * 32;
            rc = enqueue(k);
            if (rc) return rc;
        }
        if (!progressed) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    rc = join_streams_all(h);
    if (rc) return rc;
    rc = promote_overflowed(h, max_ticks);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_stop, h->user_stream));
    h->state_valid = false; h->results_valid = true;
    return REM2D_OK;
}

// Whole episodes for the uploaded population on the persistent episode kernels (one per class, concurrent).
static int launch_episodes(rem2d_handle* h, int max_ticks) {
    if (!h->have_terrain) { h->err = "run_episodes: no terrain set"; return REM2D_E_INVALID; }
    // creatures still alive after park_ticks are finished by tail-mode launches (REM2D_PARK_TICKS=0 disables parking)
    // Measured on B200 (tools/sweep_groups.py, 65536 L-system creatures): parking at 200-256 ticks (~0.5-1.4 % of the creatures)
    // is the best trade: a tail warp finishes a creature 3-4x sooner than its bulk lane would, but it occupies a whole warp, so
    // earlier thresholds (thousands of parked creatures) slow the bulk down more than they shorten the critical path; later
    // ones leave the longest-lived creatures on the slow path. Drain / late-start parking never paid off. The number of
    // parked creatures per class is bounded (a quarter of the class, at most 16 per SM): in an evolved population where most creatures live
    // long, the rest simply stay on their lanes.
    int park_ticks = 200;            // measured (tools/sweep_groups.py, 65536 creatures): 256 -> 820 ms, 200 -> 790 ms, 160 -> 930 ms
    double cap_frac = 0.25;
    {   // Under-filled GPU (every creature of every class has a group from the start, e.g. 16384 creatures): tail warps find
        // free SM resources, so parking earlier pays (16384 creatures: 450 -> 413 ms with 160 ticks and a quarter of a class).
        bool single_round = true;
        for (int k = 0; k < N_CLASSES; ++k) {
            const int per = 32 >> class_gs(h, k);
            if (h->cls[k].n_batches && h->cls[k].episode_grid * per < h->cls[k].n_members) single_round = false;
        }
        bool second = false;
        for (int k = 0; k < N_CLASSES; ++k) second |= h->cls[k].grid2 > 0 && !h->overflow_wave;
        if (single_round || second) { park_ticks = 160; cap_frac = 0.25; }
    }
    if (h->opt.park_ticks >= 0) park_ticks = h->opt.park_ticks;
    if (h->opt.park_cap >= 0.0) cap_frac = h->opt.park_cap;
    CK(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * N_COUNTER_WORDS, h->user_stream));
    CK(cudaEventRecord(h->ev_start, h->user_stream));
    int rc = fork_streams(h);
    if (rc) return rc;
    // Small populations: one warp per creature from tick 0 (mode 2). All creatures then advance at the low tail-mode tick
    // latency and the run takes about one creature lifetime. Measured (tools/small_pop.py, L-system creatures): 2.4x faster
    // than the bulk mode at 128 creatures, 1.9x at 1024, 1.2x at 6144, break-even near 10^4, 0.6x at 16384 (the bulk mode has
    // 32x the lane efficiency).
    const int warp_mode_max = h->opt.warp_mode_max >= 0 ? h->opt.warp_mode_max : h->n_sms * 48;
    const bool warp_mode = h->n_creatures <= warp_mode_max;
    if (warp_mode) park_ticks = 0;
    const bool trace = h->opt.trace != 0;
    // Launch order = decreasing shared memory per CTA (first-fit decreasing): the work distributor places the CTAs of the
    // launches in order, and a CTA that finds no SM with enough free shared memory holds back every launch behind it.
    // Measured (profiles/r2_timeline_ea_gen2_before.txt): with the 33-44 body class at 8 lanes per creature (22 KB per CTA)
    // launched first, its 822 CTAs left < 109 KB free on every SM, the 23-32 body class (109 KB per CTA) could not start, and
    // all seven smaller classes waited behind it until the first launch had finished (1.3 s of a 3.8 s evaluation).
    int launch_order[N_CLASSES];
    for (int k = 0; k < N_CLASSES; ++k) launch_order[k] = k;
    std::stable_sort(launch_order, launch_order + N_CLASSES, [&](int a, int b) {
        const size_t sa = g_classes(a).hot_bytes(warp_mode ? 5 : class_gs(h, a)), sb = g_classes(b).hot_bytes(warp_mode ? 5 : class_gs(h, b));
        return sa != sb ? sa > sb : a > b;
    });
    for (int oi = 0; oi < N_CLASSES; ++oi) {
        const int k = launch_order[oi];
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        CK(cudaMemsetAsync(cs.d_queue, 0, sizeof(int), cs.stream));
        CK(cudaEventRecord(cs.t_begin, cs.stream));
        if (warp_mode) {      // queue mode with a whole warp per creature and one warp for every creature
            g_classes(k).episode(1, 5, cs.n_members, cs.stream, cs.d_state, cs.d_lane_creature, cs.n_members, cs.d_queue, h->dpop, h->d_ter,
                                 h->d_consts, max_ticks, h->d_fitness, h->d_ticks, h->d_alive, h->d_status, h->d_counters,
                                 ParkPolicy{0, 0, 0.0f, 0, 0, 0, nullptr, nullptr}, nullptr, nullptr, nullptr, 1);
            CK(cudaEventRecord(cs.t_end, cs.stream));
            h->launches++;
            continue;
        }
        CK(cudaMemsetAsync(cs.d_n_alive, 0, sizeof(int), cs.stream));
        CK(cudaMemsetAsync(cs.d_lc_work[0], 0, cs.lane_creature.size() * sizeof(int), cs.stream));     // "unpublished" markers
        ParkPolicy park;
        park.ticks = park_ticks < max_ticks ? park_ticks : 0;
        park.cap = std::min(cs.n_members, std::max(32, std::min((int)(h->n_sms * 64 * cap_frac), (int)(cs.n_members * cap_frac))));
        // late starters (pulled when a lane of the first round frees up): park right after the wall of death has passed the
        // start pad (tick 126), so that the long-lived ones among them continue at the tail launches' tick latency
        // a root that is already where the wall of death will be at the park threshold survives until then unless it walks back
        park.lead_x = (h->opt.park_lead && h->cfg.terminate) ? (float)(park.ticks * h->cfg.wod_speed) : 0.0f;
        park.lead_from = 32;
        park.late_from = cs.episode_grid * (32 >> class_gs(h, k));
        park.late_ticks = h->opt.park_late_ticks >= 0 ? std::min(h->opt.park_late_ticks, park.ticks) : park.ticks;
        park.trace = nullptr; park.tail_trace = nullptr;
        if (trace) {
            const size_t tb = (size_t)std::max(cs.n_members, 1) * 4 * sizeof(unsigned int);
            CK(ensure(cs.b_ttrace, tb));
            CK(cudaMemsetAsync(cs.b_ttrace.p, 0, tb, cs.stream));
            park.tail_trace = (unsigned int*)cs.b_ttrace.p;
            const size_t bytes = (size_t)cs.episode_grid * REM2D_TRACE_SAMPLES * 2 * sizeof(unsigned int);
            CK(ensure(cs.b_trace, bytes));
            CK(cudaMemsetAsync(cs.b_trace.p, 0, bytes, cs.stream));
            park.trace = (unsigned int*)cs.b_trace.p;
        }
        if (cs.grid2 > 0) { CK(cudaEventRecord(cs.done2, cs.stream)); CK(cudaStreamWaitEvent(cs.stream2, cs.done2, 0)); }   // queue / park counters are reset
        cs.park = park;
        g_classes(k).episode(h->image, class_gs(h, k), cs.episode_grid, cs.stream, cs.d_state, cs.d_lane_creature, cs.n_members, cs.d_queue, h->dpop,
                             h->d_ter, h->d_consts, max_ticks, h->d_fitness, h->d_ticks, h->d_alive, h->d_status, h->d_counters,
                             park, cs.d_state2, cs.d_lc_work[0], cs.d_n_alive, (cs.grid2 > 0 && !h->overflow_wave) ? 0 : 1);
        CK(cudaMemcpyAsync(cs.h_n_alive, cs.d_n_alive, sizeof(int), cudaMemcpyDeviceToHost, cs.stream));
        CK(cudaEventRecord(cs.t_end, cs.stream));
        h->launches++;
    }
    // second launches (after ALL first launches, so that their CTAs queue behind them): the creatures a class could not seat
    // in its first round, in wide groups, on the columns behind the first launch's
    for (int oi = 0; oi < N_CLASSES && !warp_mode; ++oi) {
        const int k = launch_order[oi];
        ClassState& cs = h->cls[k];
        if (!cs.n_batches || cs.grid2 <= 0) continue;
        const int per1 = 32 >> class_gs(h, k);
        const size_t first_block = ((size_t)cs.episode_grid * per1 + 31) / 32;
        ParkPolicy park = cs.park;
        park.trace = nullptr;
        g_classes(k).episode(h->image, cs.gs2, cs.grid2, cs.stream2, cs.d_state + first_block * g_classes(k).words * 32, cs.d_lane_creature, cs.n_members,
                             cs.d_queue, h->dpop, h->d_ter, h->d_consts, max_ticks, h->d_fitness, h->d_ticks, h->d_alive, h->d_status,
                             h->d_counters, park, cs.d_state2, cs.d_lc_work[0], cs.d_n_alive, 1);
        h->launches++;
    }
    CK(cudaGetLastError());
    if (park_ticks > 0 && park_ticks < max_ticks) {
        // Tail: while the episode kernels run, poll their park counters and hand newly parked creatures to the
        // warp-per-creature tail mode right away (pool of streams), so the sequential ticks of the longest-lived
        // creatures overlap the bulk instead of extending the run; the last launch of a class happens when its episode
        // kernel has finished.
        bool running[N_CLASSES];
        int launched[N_CLASSES], waited[N_CLASSES] = {};
        int n_running = 0;
        size_t rr = 0;
        h->tail_used.assign(h->tail_pool.size(), 0);
        auto idle_stream = [&]() -> cudaStream_t {
            for (size_t i = 0; i < h->tail_pool.size(); ++i) {
                const size_t s = (rr + i) % h->tail_pool.size();
                if (!h->tail_used[s] || cudaStreamQuery(h->tail_pool[s]) == cudaSuccess) { rr = s + 1; h->tail_used[s] = 1; return h->tail_pool[s]; }
            }
            cudaStream_t st = nullptr;
            // highest priority: tail CTAs are placed before the queued overflow-wave CTAs of the episode launches
            if (h->tail_pool.size() < 4096 && cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, h->prio_high) == cudaSuccess) {
                h->tail_pool.push_back(st); h->tail_used.push_back(1);
                return st;
            }
            cudaGetLastError();
            const size_t s = rr++ % h->tail_pool.size();      // cannot grow: queue behind a busy stream
            h->tail_used[s] = 1;
            return h->tail_pool[s];
        };
        for (int k = 0; k < N_CLASSES; ++k) { running[k] = h->cls[k].n_batches > 0; launched[k] = 0; n_running += running[k] ? 1 : 0; }
        while (n_running > 0) {
            bool finished_now[N_CLASSES];
            for (int k = 0; k < N_CLASSES; ++k) {
                finished_now[k] = false;
                if (!running[k]) continue;
                cudaError_t q = cudaStreamQuery(h->cls[k].stream);       // BEFORE reading the counter: no slot can be missed
                if (q == cudaSuccess && h->cls[k].grid2 > 0) q = cudaStreamQuery(h->cls[k].stream2);
                if (q == cudaSuccess) finished_now[k] = true;
                else if (q != cudaErrorNotReady) { h->err = std::string("episode kernel: ") + cudaGetErrorString(q); return REM2D_E_CUDA; }
                CK(cudaMemcpyAsync(&h->h_poll[k], h->cls[k].d_n_alive, sizeof(int), cudaMemcpyDeviceToHost, h->poll_stream));
            }
            CK(cudaStreamSynchronize(h->poll_stream));
            for (int k = N_CLASSES - 1; k >= 0; --k) {
                if (!running[k]) continue;
                ClassState& cs = h->cls[k];
                const int cnt = h->h_poll[k];
                // early launches in chunks (a handful of creatures or whatever is there when the class is done)
                if (cnt > launched[k]) ++waited[k]; else waited[k] = 0;
                if (cnt > launched[k] && (finished_now[k] || cnt - launched[k] >= 4 || waited[k] >= 3)) {
                    waited[k] = 0;
                    // a handful of parked creatures (random populations): a warp each; dozens at once (evolved populations,
                    // where many creatures outlive the threshold): 8 lanes each - 4x fewer warps at nearly the same tick latency
                    const int tail_gs = (cnt - launched[k] >= 32) ? std::min(3, h->opt.tail_group_shift) : h->opt.tail_group_shift;
                    g_classes(k).tail(h->image, tail_gs, idle_stream(), cs.d_state2, cs.d_lc_work[0], launched[k], cnt - launched[k],
                                      h->d_ter, h->d_consts, max_ticks, h->d_fitness, h->d_ticks, h->d_alive, h->d_status, h->d_counters,
                                      trace ? (unsigned int*)cs.b_ttrace.p : nullptr);
                    CK(cudaGetLastError());
                    h->launches++;
                    launched[k] = cnt;
                }
                if (finished_now[k]) { running[k] = false; --n_running; }
            }
            if (n_running > 0) std::this_thread::sleep_for(std::chrono::microseconds(500));
        }
        for (size_t s = 0; s < h->tail_pool.size(); ++s) {
            if (!h->tail_used[s]) continue;
            CK(cudaEventRecord(h->pool_done, h->tail_pool[s]));
            CK(cudaStreamWaitEvent(h->user_stream, h->pool_done, 0));
        }
    }
    rc = join_streams(h);
    if (rc) return rc;
    rc = promote_overflowed(h, max_ticks);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_stop, h->user_stream));
    h->state_valid = false; h->results_valid = true;
    return REM2D_OK;
}

extern "C" int rem2d_reset(rem2d_handle* h) {
    if (!h) return REM2D_E_INVALID;
    if (!h->have_pop) { h->err = "reset: no population uploaded"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    return launch_reset(h);
}

static int gather(rem2d_handle* h) {
    if (!h->state_valid && !h->results_valid) { h->err = "no results: call rem2d_reset/rem2d_step or rem2d_run_episodes first"; return REM2D_E_INVALID; }
    for (int k = 0; k < N_CLASSES && h->state_valid; ++k) {
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        int n_lanes = cs.n_batches * 32;
        gather_kernel<<<(n_lanes + 127) / 128, 128, 0, h->user_stream>>>(cs.d_state, cs.d_lane_creature, n_lanes, g_classes(k).words,
                                                                        h->d_fitness, h->d_ticks, h->d_alive, h->d_status);
        h->launches++;
    }
    CK(cudaGetLastError());
    int n = h->n_creatures;
    CK(cudaMemcpyAsync(h->h_fitness.data(), h->d_fitness, sizeof(double) * n, cudaMemcpyDeviceToHost, h->user_stream));
    CK(cudaMemcpyAsync(h->h_ticks.data(), h->d_ticks, sizeof(int) * n, cudaMemcpyDeviceToHost, h->user_stream));
    CK(cudaMemcpyAsync(h->h_alive.data(), h->d_alive, sizeof(int) * n, cudaMemcpyDeviceToHost, h->user_stream));
    CK(cudaMemcpyAsync(h->h_status.data(), h->d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, h->user_stream));
    CK(cudaStreamSynchronize(h->user_stream));
    for (int c = 0; c < n; ++c)
        if (h->h_status[c]) {
            char buf[200];
            snprintf(buf, sizeof(buf), "creature %d exceeded a capacity of its class (status %d: 1 contact pool, 4 TOI island)", c, h->h_status[c]);
            h->err = buf;
            return REM2D_E_CAPACITY;
        }
    return REM2D_OK;
}

extern "C" {

int rem2d_step(rem2d_handle* h, int32_t n_ticks) {
    if (!h) return REM2D_E_INVALID;
    if (!h->have_pop || n_ticks < 0) { h->err = "step: no population / negative tick count"; return REM2D_E_INVALID; }
    if (!h->state_valid) { h->err = "step: per-creature state was consumed by rem2d_run_episodes/rem2d_evaluate; call rem2d_reset first"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    CK(cudaEventRecord(h->ev_start, h->user_stream));
    int rc = fork_streams(h);
    if (rc) return rc;
    for (int k = N_CLASSES - 1; k >= 0; --k) {        // most expensive class first
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        g_classes(k).step(step_gs(h, k), cs.n_batches, cs.stream, cs.d_state, n_ticks, h->d_ter, h->d_consts, h->d_counters);
        h->launches++;
    }
    CK(cudaGetLastError());
    rc = join_streams(h);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev_stop, h->user_stream));
    CK(cudaEventSynchronize(h->ev_stop));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev_start, h->ev_stop));
    return REM2D_OK;
}

int rem2d_fitness(rem2d_handle* h, double* out) {
    if (!h || !out) return REM2D_E_INVALID;
    if (!h->have_pop) { h->err = "fitness: no population uploaded"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    int rc = gather(h);
    if (rc) return rc;
    memcpy(out, h->h_fitness.data(), sizeof(double) * h->n_creatures);
    return REM2D_OK;
}

int rem2d_get_counters(rem2d_handle* h, uint64_t* out) {
    if (!h || !out) return REM2D_E_INVALID;
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->user_stream));
    CK(cudaMemcpy(out, h->d_counters, sizeof(unsigned long long) * REM2D_N_COUNTERS, cudaMemcpyDeviceToHost));
    return REM2D_OK;
}
// Diagnostics (libraries built with -DREM2D_PHASE_TIMING): warp-cycles per tick phase of the last evaluation,
// out[(mode * 6 + gs) * 16 + phase], mode 0 = queue launches, 1 = tail launches; zeros in production builds.
int rem2d_debug_phases(rem2d_handle* h, uint64_t* out) {
    if (!h || !out) return REM2D_E_INVALID;
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->user_stream));
    CK(cudaMemcpy(out, h->d_counters + REM2D_N_COUNTERS, sizeof(unsigned long long) * 12 * 16, cudaMemcpyDeviceToHost));
    return REM2D_OK;
}

int rem2d_run_episodes(rem2d_handle* h, int32_t max_ticks) {
    if (!h) return REM2D_E_INVALID;
    if (!h->have_pop || max_ticks < 0) { h->err = "run_episodes: no population / negative tick count"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    // Default: persistent kernel with per-lane refill. REM2D_EPISODE_MODE=phased selects tick phases with survivor
    // compaction instead (measured slower at pop 65536: the run is bound by the sequential ticks of the longest-lived
    // large creature, and phases add a launch/sync per 32 ticks to exactly that critical path).
    int rc = h->opt.phased ? launch_phased(h, max_ticks) : launch_episodes(h, max_ticks);
    if (rc) return rc;
    CK(cudaEventSynchronize(h->ev_stop));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev_start, h->ev_stop));
    return REM2D_OK;
}

int rem2d_ticks(rem2d_handle* h, int32_t* out) {
    if (!h || !out) return REM2D_E_INVALID;
    if (!h->have_pop) { h->err = "ticks: no population uploaded"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    int rc = gather(h);
    if (rc) return rc;
    memcpy(out, h->h_ticks.data(), sizeof(int) * h->n_creatures);
    return REM2D_OK;
}

int rem2d_evaluate(rem2d_handle* h, const rem2d_population* pop, int32_t max_ticks, double* fitness_out, int32_t* ticks_out) {
    int rc = upload_impl(h, pop, false);      // the episode kernel builds the worlds itself
    if (rc) return rc;
    rc = rem2d_run_episodes(h, max_ticks);
    if (rc) return rc;
    rc = gather(h);
    if (rc) return rc;
    if (fitness_out) memcpy(fitness_out, h->h_fitness.data(), sizeof(double) * h->n_creatures);
    if (ticks_out) memcpy(ticks_out, h->h_ticks.data(), sizeof(int) * h->n_creatures);
    return REM2D_OK;
}

// Measured non-fused FP32 issue peak of this device in GFLOP/s (best of 5), for the roofline denominator.
int rem2d_measure_fp32_peak(rem2d_handle* h, double* gflops) {
    if (!h || !gflops) return REM2D_E_INVALID;
    cudaSetDevice(h->cfg.device);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    float* d = nullptr;
    CK(cudaMalloc(&d, sizeof(float) * blocks * threads));
    double best = 0.0;
    for (int r = 0; r < 6; ++r) {
        CK(cudaEventRecord(h->ev_start, h->user_stream));
        fp32_issue_kernel<<<blocks, threads, 0, h->user_stream>>>(d, iters, 1.0000001f, 1e-7f);
        CK(cudaEventRecord(h->ev_stop, h->user_stream));
        CK(cudaEventSynchronize(h->ev_stop));
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, h->ev_start, h->ev_stop));
        double gf = (double)blocks * threads * iters * 16.0 / (ms * 1e-3) / 1e9;
        if (r > 0 && gf > best) best = gf;
    }
    cudaFree(d);
    *gflops = best;
    return REM2D_OK;
}

// Diagnostics: per capacity class, [nb, members, resident warps, kernel begin ms, kernel end ms] of the last
// rem2d_run_episodes (refill mode), times relative to its start. out has room for 5 * 16 floats; returns #classes.
int rem2d_debug_class_timeline(rem2d_handle* h, float* out) {
    if (!h || !out) return REM2D_E_INVALID;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (int k = 0; k < N_CLASSES; ++k) {
        ClassState& cs = h->cls[k];
        float b = -1.0f, e = -1.0f;
        if (cs.n_batches) { cudaEventElapsedTime(&b, h->ev_start, cs.t_begin); cudaEventElapsedTime(&e, h->ev_start, cs.t_end); }
        out[5 * k] = (float)g_classes(k).nb; out[5 * k + 1] = (float)cs.n_members; out[5 * k + 2] = (float)cs.episode_grid;
        out[5 * k + 3] = b; out[5 * k + 4] = e;
    }
    cudaGetLastError();
    return N_CLASSES;
}

// Diagnostics: the REM2D_TRACE=1 samples of class k of the last rem2d_run_episodes: [warps][REM2D_TRACE_SAMPLES][2] uint32.
// Returns the number of warps (0: none), negative on error; copies at most max_words words.
int rem2d_debug_trace(rem2d_handle* h, int k, unsigned int* out, int64_t max_words) {
    if (!h || !out || k < 0 || k >= N_CLASSES) return REM2D_E_INVALID;
    ClassState& cs = h->cls[k];
    if (!cs.b_trace.p || !cs.n_batches) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    size_t words = (size_t)cs.episode_grid * REM2D_TRACE_SAMPLES * 2;
    if ((int64_t)words > max_words) words = (size_t)max_words;
    CK(cudaMemcpy(out, cs.b_trace.p, words * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return cs.episode_grid;
}

// Diagnostics: per park slot {us parked, us tail start, us tail end, tail ticks} of class k (REM2D_TRACE=1). Returns #slots.
int rem2d_debug_tail_trace(rem2d_handle* h, int k, unsigned int* out, int64_t max_words) {
    if (!h || !out || k < 0 || k >= N_CLASSES) return REM2D_E_INVALID;
    ClassState& cs = h->cls[k];
    if (!cs.b_ttrace.p || !cs.n_batches) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    int parked = 0;
    CK(cudaMemcpy(&parked, cs.d_n_alive, sizeof(int), cudaMemcpyDeviceToHost));
    size_t words = (size_t)parked * 4;
    if ((int64_t)words > max_words) words = (size_t)max_words;
    CK(cudaMemcpy(out, cs.b_ttrace.p, words * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return parked;
}

int rem2d_set_priority(rem2d_handle* h, const float* expected_ticks, int32_t n) {
    if (!h || n < 0) return REM2D_E_INVALID;
    if (!expected_ticks || n == 0) h->priority.clear();
    else h->priority.assign(expected_ticks, expected_ticks + n);
    return REM2D_OK;
}

int rem2d_read_roots(rem2d_handle* h, float* root_x, double* wod, int32_t* alive) {
    if (!h) return REM2D_E_INVALID;
    if (!h->have_pop) { h->err = "read_roots: no population uploaded"; return REM2D_E_INVALID; }
    if (!h->state_valid) { h->err = "read_roots: per-creature state was consumed by rem2d_run_episodes; call rem2d_reset first"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    const size_t n = (size_t)std::max(h->n_creatures, 1);
    CK(ensure(h->b_roots, n * (sizeof(double) + sizeof(float) + sizeof(int))));
    double* d_wod = (double*)h->b_roots.p;
    float* d_x = (float*)(d_wod + n);
    int* d_alive = (int*)(d_x + n);
    for (int k = 0; k < N_CLASSES; ++k) {
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        const int n_lanes = cs.n_batches * 32;
        roots_kernel<<<(n_lanes + 127) / 128, 128, 0, h->user_stream>>>(cs.d_state, cs.d_lane_creature, n_lanes, g_classes(k).words, d_x, d_wod, d_alive);
        h->launches++;
    }
    CK(cudaGetLastError());
    if (root_x) CK(cudaMemcpyAsync(root_x, d_x, sizeof(float) * h->n_creatures, cudaMemcpyDeviceToHost, h->user_stream));
    if (wod) CK(cudaMemcpyAsync(wod, d_wod, sizeof(double) * h->n_creatures, cudaMemcpyDeviceToHost, h->user_stream));
    if (alive) CK(cudaMemcpyAsync(alive, d_alive, sizeof(int) * h->n_creatures, cudaMemcpyDeviceToHost, h->user_stream));
    CK(cudaStreamSynchronize(h->user_stream));
    return REM2D_OK;
}

float rem2d_last_step_ms(rem2d_handle* h) { return h ? h->last_ms : 0.0f; }
int64_t rem2d_launch_count(rem2d_handle* h) { return h ? h->launches : 0; }

// Test/diagnostic path: copies the whole state of every class to the host and decodes it there.
int rem2d_read_state(rem2d_handle* h, rem2d_state_view* out) {
    if (!h || !out) return REM2D_E_INVALID;
    if (!h->have_pop) { h->err = "read_state: no population uploaded"; return REM2D_E_INVALID; }
    if (!h->state_valid) { h->err = "read_state: per-creature state was consumed by rem2d_run_episodes; call rem2d_reset first"; return REM2D_E_INVALID; }
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->user_stream));
    for (int k = 0; k < N_CLASSES; ++k) {
        ClassState& cs = h->cls[k];
        if (!cs.n_batches) continue;
        const ClassOps& ci = g_classes(k);
        std::vector<float> st((size_t)cs.n_batches * ci.words * 32);
        CK(cudaMemcpy(st.data(), cs.d_state, st.size() * sizeof(float), cudaMemcpyDeviceToHost));
        auto asint = [](float f) { int i; memcpy(&i, &f, 4); return i; };
        for (size_t gl = 0; gl < cs.lane_creature.size(); ++gl) {
            int c = cs.lane_creature[gl];
            if (c < 0) continue;
            const float* g = st.data() + (gl >> 5) * (size_t)ci.words * 32 + (gl & 31);
            auto S = [&](int f) { return g[f * 32]; };
            auto B = [&](int f, int i) { return g[(ci.off_body + i * BF_COUNT + f) * 32]; };
            auto J = [&](int f, int j) { return g[(ci.off_joint + j * JF_COUNT + f) * 32]; };
            auto C = [&](int f, int q) { return g[(ci.off_cont + q * CF_COUNT + f) * 32]; };
            int b0 = h->body_off[c], nb = h->body_off[c + 1] - b0, j0 = b0 - c;
            int awake = 0;
            for (int i = 0; i < nb; ++i) {
                if (out->pose) { out->pose[3 * (b0 + i)] = B(BF_CX, i); out->pose[3 * (b0 + i) + 1] = B(BF_CY, i); out->pose[3 * (b0 + i) + 2] = B(BF_A, i); }
                if (out->vel) { out->vel[3 * (b0 + i)] = B(BF_VX, i); out->vel[3 * (b0 + i) + 1] = B(BF_VY, i); out->vel[3 * (b0 + i) + 2] = B(BF_W, i); }
                awake |= asint(B(BF_FLAGS, i)) & BFL_AWAKE;
            }
            for (int j = 0; j < nb - 1; ++j) {
                if (out->joint_impulse) {
                    float* o = &out->joint_impulse[4 * (j0 + j)];
                    o[0] = J(JF_IMPX, j); o[1] = J(JF_IMPY, j); o[2] = J(JF_IMPZ, j); o[3] = J(JF_MIMP, j);
                }
                if (out->limit_state) out->limit_state[j0 + j] = asint(J(JF_LIMIT, j));
                if (out->motor_speed) out->motor_speed[j0 + j] = J(JF_MSPEED, j);
            }
            if (out->alive) out->alive[c] = asint(S(S_ALIVE));
            if (out->ticks) out->ticks[c] = asint(S(S_TICKS));
            if (out->awake) out->awake[c] = awake ? 1 : 0;
            if (out->wod) { int lo = asint(S(S_WOD_LO)), hi = asint(S(S_WOD_HI)); uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); out->wod[c] = d; }
            int nc = asint(S(S_NC)), ntouch = 0;
            for (int q = 0; q < nc; ++q) ntouch += ((asint(C(CF_KEY, q)) >> 16) & CK_TOUCHING) ? 1 : 0;
            if (out->n_contacts) out->n_contacts[c] = nc;
            if (out->n_touching) out->n_touching[c] = ntouch;
            if (out->touching_pairs && out->max_pairs > 0) {
                int32_t* tp = &out->touching_pairs[(size_t)c * out->max_pairs * 2];
                float* ti = out->touching_impulse ? &out->touching_impulse[(size_t)c * out->max_pairs * 4] : nullptr;
                for (int q = 0; q < out->max_pairs * 2; ++q) tp[q] = -1;
                if (ti) for (int q = 0; q < out->max_pairs * 4; ++q) ti[q] = 0.0f;
                std::vector<std::pair<int, int>> pairs;    // (body<<8 | edge, pool index)
                for (int q = 0; q < nc; ++q) {
                    int key = asint(C(CF_KEY, q));
                    if (!((key >> 16) & CK_TOUCHING)) continue;
                    pairs.push_back({((key & 0xff) << 8) | ((key >> 8) & 0xff), q});
                }
                std::sort(pairs.begin(), pairs.end());
                for (size_t q = 0; q < pairs.size() && (int)q < out->max_pairs; ++q) {
                    int pq = pairs[q].second;
                    int key = asint(C(CF_KEY, pq));
                    int count = (key >> (16 + CK_COUNT_SHIFT)) & 3;
                    tp[2 * q] = key & 0xff; tp[2 * q + 1] = (key >> 8) & 0xff;
                    if (ti) {
                        ti[4 * q] = C(CF_P0N, pq); ti[4 * q + 1] = count > 1 ? C(CF_P1N, pq) : 0.0f;
                        ti[4 * q + 2] = C(CF_P0T, pq); ti[4 * q + 3] = count > 1 ? C(CF_P1T, pq) : 0.0f;
                    }
                }
            }
        }
    }
    return REM2D_OK;
}

}  // extern "C"
