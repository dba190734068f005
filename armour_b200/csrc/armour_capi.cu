// C ABI of libarmour_b200.so (see include/armour_b200.h).  One translation unit: the kernels are
// included below so they share the __constant__ block.  No CPU fallback anywhere: every compute entry
// point launches CUDA kernels on the context's stream and reports CUDA errors as ARMOUR_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/armour_b200.h"
#include "bezier.cuh"
#include "device_constants.cuh"
// The reach-set kernel is compiled twice (see k1_pz.cuh):
//   k1lat: one unit per CTA built by 3 groups of 128 threads (task list, MG mode) -> lowest latency for one planning problem
//   k1thr: 2 CTAs per SM of 8 units each in lock step, 64 threads per unit -> highest throughput for batches.  Instruction
//          fetch bounds this kernel: the units of a CTA run the same operation at the same time (without the lock step the
//          instruction cache thrashes: 2.3x slower), and two lock-step domains per SM overlap each other's barriers.  A unit
//          keeps only its control block in shared memory; arena, scratch and joint reachable set live in its global scratch,
//          behind the L1 cache that the unclaimed shared memory becomes (profiles/r3_k1thr_experiments.md: 445 -> 360 us
//          per problem; round-1 sweep with one CTA per SM: 8 groups 535 us, 10: 519, 12: 503, 14: 497)
#ifdef K1_PROFILE
constexpr int K1_PROF_SITES = 512;
__device__ long long g_k1prof[128 * K1_PROF_SITES * 2];
__device__ int g_k1spill[8];  // arena blocks in global memory, all arena blocks, scratch pools in global memory
#endif
// Latency configuration: MG mode, K1LAT_GROUPS groups of K1LAT_NT threads build ONE unit together (one CTA per SM)
#ifndef K1LAT_NT
#define K1LAT_NT 128
#endif
#ifndef K1LAT_CTAS
#define K1LAT_CTAS 1
#endif
#ifndef K1LAT_GROUPS
#define K1LAT_GROUPS 3
#endif
#ifndef K1LAT_MG
#define K1LAT_MG 1
#endif
#ifndef K1LAT_TAB_16THS
#define K1LAT_TAB_16THS 11  // share (in sixteenths) of a group's dynamic shared memory given to the scratch pool
#endif
#define K1_TAB_16THS K1LAT_TAB_16THS
#ifndef K1LAT_DYN_CAP
#define K1LAT_DYN_CAP -1  // shared memory per group beyond the fixed region: no cap
#endif
#define K1_DYN_CAP K1LAT_DYN_CAP
#define K1_NS k1lat
#define K1_NT K1LAT_NT
#define K1_CTAS K1LAT_CTAS
#define K1_GROUPS K1LAT_GROUPS
#define K1_MG K1LAT_MG
#include "k1_reachsets.cuh"
#undef K1_NS
#undef K1_NT
#undef K1_CTAS
#undef K1_GROUPS
#undef K1_MG
#undef K1_TAB_16THS
#define K1_MG 0
#ifndef K1THR_NT
#define K1THR_NT 64
#endif
#ifndef K1THR_GROUPS
#define K1THR_GROUPS 8
#endif
#ifndef K1THR_CTAS
#define K1THR_CTAS 2
#endif
#ifndef K1THR_TAB_EIGHTHS
#define K1THR_TAB_EIGHTHS 4  // share (in eighths) of a group's dynamic shared memory given to the scratch pool
#endif
#ifndef K1THR_DYN_CAP
#define K1THR_DYN_CAP 0  // no shared-memory arena / scratch per group: the global scratch behind a large L1 is faster
#endif
#ifndef K1THR_JRS_GLOBAL
#define K1THR_JRS_GLOBAL 1  // the joint-reachable-set region of a group sits in its global scratch too (only the control block in shared memory)
#endif
#undef K1_JRS_GLOBAL
#define K1_JRS_GLOBAL K1THR_JRS_GLOBAL
#undef K1_DYN_CAP
#define K1_DYN_CAP K1THR_DYN_CAP
#undef K1_TAB_EIGHTHS
#define K1_TAB_EIGHTHS K1THR_TAB_EIGHTHS
#define K1_TAB_16THS (2 * K1THR_TAB_EIGHTHS)
#define K1_NS k1thr
#define K1_NT K1THR_NT
#define K1_CTAS K1THR_CTAS
#define K1_GROUPS K1THR_GROUPS
#include "k1_reachsets.cuh"
#undef K1_NS
#undef K1_NT
#undef K1_CTAS
#undef K1_GROUPS
#include "k3_constraints.cuh"
#include "k4_solver.cuh"
#include "armtd.cuh"
#include "layout.h"

using namespace armour;

struct armour_ctx {
    armour_config cfg;
    RobotConstants rc;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    Batch B;                 // device pointers + dimensions of the current batch
    int built_nprob = 0;     // problems with valid reach sets
    size_t hp_capacity = 0;  // doubles allocated for B.hp_cand
    int hp_nprob = 0;        // problems the candidate buffers were sized for
    int* d_unit_flag = nullptr;  // [max_problems * T] completion stamps of the reach-set units (latency path)
    size_t obs_capacity = 0;
    double* d_in = nullptr;  // [3][max_problems][NF] q0, qd0, qdd0
    double* d_obs = nullptr;
    double* d_k = nullptr;   // staging for the host-pointer evaluation calls
    double* d_g = nullptr;
    double* d_jac = nullptr;
    size_t g_capacity = 0, jac_capacity = 0;
    double* d_jnz = nullptr;  // packed non-zeros of the Jacobian (structured evaluation)
    size_t jnz_capacity = 0;
    int* d_verdict = nullptr;  // [2][max_problems]
    // batched device solver (allocated on first use)
    double* d_solver = nullptr;   // state arrays + second g buffer + linearised rows
    int* d_solver_i = nullptr;    // have_best, status, iters, evals, running counter
    double* d_solver_io = nullptr;  // q_des, k_opt staging of the host-pointer call
    size_t solver_doubles = 0, solver_ints = 0;  // elements allocated for d_solver / d_solver_i
    k1lat::K1Scratch k1_lat;  // scratch of the latency configuration of k_reachsets
    k1thr::K1Scratch k1_thr;  // scratch of the throughput configuration (allocated on first use)
    bool k1_thr_ready = false;
    long long launches = 0;
    std::string last_error;
    std::vector<double> h_torque_radius;  // host mirror of problem 0..built_nprob-1 (lazy)
    bool h_torque_valid = false;
    std::vector<double> h_q0, h_qd0, h_qdd0;
    // ARMTD comparison planner (armour_armtd_*): imported joint reachable set, KPA-layout outputs, trajectory parameters
    bool armtd = false;
    double* d_jrs = nullptr;   // [ARMTD_T_PADDED][NF][6]
    double* d_ga = nullptr;    // [m_armtd]
    double* d_ja = nullptr;    // [m_armtd][NF]
    size_t ga_capacity = 0, ja_capacity = 0;
    ArmtdParams armtd_par;
};

namespace {

// FP64 FMA throughput probe: 8 independent accumulator chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = double(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fma_rn(x[i], a, 0.5);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

int fail(armour_ctx* c, int code, const std::string& msg) {
    if (c) c->last_error = msg;
    return code;
}
#define CU(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(ctx, ARMOUR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

int ensure_eval_buffers(armour_ctx* ctx, int nprob, bool need_g, bool need_jac) {
    const size_t m = size_t(ctx->B.m());
    if (need_g && ctx->g_capacity < nprob * m) {
        if (ctx->d_g) cudaFree(ctx->d_g);
        ctx->d_g = nullptr;
        CU(dalloc(&ctx->d_g, nprob * m));
        ctx->g_capacity = nprob * m;
    }
    if (need_jac && ctx->jac_capacity < nprob * m * NF) {
        if (ctx->d_jac) cudaFree(ctx->d_jac);
        ctx->d_jac = nullptr;
        CU(dalloc(&ctx->d_jac, nprob * m * NF));
        ctx->jac_capacity = nprob * m * NF;
    }
    return ARMOUR_OK;
}

int ensure_obstacle_buffers(armour_ctx* ctx, int nprob, int nobs) {
    const size_t need_obs = size_t(nprob) * nobs * 12;
    if (ctx->obs_capacity < need_obs) {
        if (ctx->d_obs) cudaFree(ctx->d_obs);
        ctx->d_obs = nullptr;
        CU(dalloc(&ctx->d_obs, need_obs));
        ctx->obs_capacity = need_obs;
    }
    Batch& B = ctx->B;
    B.O = nobs;
    B.nprob = nprob;
    const size_t need_hp = size_t(nprob) * (B.T / TB) * B.hp_chunk();
    if (ctx->hp_capacity < need_hp || ctx->hp_nprob < nprob) {
        // (the candidate lists of the current batch go with the old buffers: whatever was built must be rebuilt)
        if (B.hp_cand) cudaFree(B.hp_cand);
        if (B.hp_cnt) cudaFree(B.hp_cnt);
        if (B.hp_slow) cudaFree(B.hp_slow);
        B.hp_cand = nullptr;
        B.hp_cnt = nullptr;
        B.hp_slow = nullptr;
        ctx->hp_capacity = 0;
        ctx->hp_nprob = 0;
        ctx->built_nprob = 0;
        const size_t np = std::max(nprob, ctx->hp_nprob);
        CU(dalloc(&B.hp_cand, need_hp));
        CU(dalloc(&B.hp_cnt, need_hp / (HP_CAP * 4)));
        CU(dalloc(&B.hp_slow, np));
        CU(cudaMemsetAsync(B.hp_slow, 0, np * sizeof(int), ctx->stream));
        ctx->hp_capacity = need_hp;
        ctx->hp_nprob = nprob;
    }
    B.obstacles = ctx->d_obs;
    return ARMOUR_OK;
}

// Pick the kernel configuration by batch size: up to two waves of the latency configuration's CTAs are
// faster there; beyond that the lock-step throughput configuration wins.
int launch_build(armour_ctx* ctx, const Batch& B, int* nl, bool* latency) {
    const long long nunits = (long long)B.nprob * B.T;
    const bool thr = ctx->cfg.max_problems > 1 && nunits > 2LL * ctx->k1_lat.grid;
    *latency = !thr;
    if (thr) {
        if (!ctx->k1_thr_ready) {
            CU(k1thr::k1_scratch_create(&ctx->k1_thr, ctx->cfg, ctx->rc, ctx->stream));
            ctx->k1_thr_ready = true;
        }
        CU(k1thr::launch_reachsets(B, ctx->k1_thr, ctx->stream, nl));
    } else {
        CU(k1lat::launch_reachsets(B, ctx->k1_lat, ctx->stream, nl, ctx->d_unit_flag));
    }
    return ARMOUR_OK;
}

// The robot / planner constants live in ONE __constant__ block per device, shared by every context of the process.
// Before a context launches anything it makes sure the block holds ITS constants: contexts with identical
// configurations share the block freely; switching between contexts that differ (gripper model, threshold, k_range,
// uncertainties ...) waits for the device to drain and uploads again, so no kernel ever runs on another context's
// constants.  (Correct, not fast: interleave differing contexts on one device sparingly.)
std::mutex g_const_mutex;
struct ConstOwner {
    bool valid = false;
    RobotConstants rc;
};
ConstOwner g_const_owner[64];

int ensure_constants(armour_ctx* ctx) {
    const int dev = ctx->cfg.device;
    if (dev < 0 || dev >= 64) return fail(ctx, ARMOUR_ERR_ARG, "device ordinal");
    std::lock_guard<std::mutex> lock(g_const_mutex);
    ConstOwner& o = g_const_owner[dev];
    if (o.valid && std::memcmp(&o.rc, &ctx->rc, sizeof(RobotConstants)) == 0) return ARMOUR_OK;
    CU(cudaDeviceSynchronize());  // kernels of the previous owner may still read the block
    CU(upload_constants(ctx->rc, ctx->stream));
    o.rc = ctx->rc;
    o.valid = true;
    return ARMOUR_OK;
}

int check_batch(armour_ctx* ctx, int nprob, int nobs) {
    if (!ctx) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->cfg.max_problems)
        return fail(ctx, ARMOUR_ERR_STATE, "nprob outside [1, max_problems]");
    if (nobs < 0) return fail(ctx, ARMOUR_ERR_ARG, "negative obstacle count");
    if (nobs > ctx->cfg.max_obstacles)
        return fail(ctx, ARMOUR_ERR_OBSTACLES, "number of obstacles larger than max_obstacles");
    return ARMOUR_OK;
}

int fetch_torque_radius(armour_ctx* ctx) {
    if (ctx->h_torque_valid) return ARMOUR_OK;
    const size_t n = size_t(ctx->built_nprob) * NF * ctx->B.T;
    ctx->h_torque_radius.resize(n);
    CU(cudaMemcpyAsync(ctx->h_torque_radius.data(), ctx->B.torque_radius, n * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->h_torque_valid = true;
    return ARMOUR_OK;
}

}  // namespace

extern "C" {

int armour_abi_version(void) { return ARMOUR_ABI_VERSION; }

const char* armour_status_string(int s) {
    switch (s) {
        case ARMOUR_OK: return "ok";
        case ARMOUR_ERR_ARG: return "invalid argument";
        case ARMOUR_ERR_CUDA: return "CUDA error";
        case ARMOUR_ERR_OBSTACLES: return "too many obstacles";
        case ARMOUR_ERR_CAPACITY: return "monomial table capacity exceeded";
        case ARMOUR_ERR_STATE: return "invalid call order or batch size";
        case ARMOUR_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

int armour_config_default(armour_config* cfg) {
    if (!cfg) return ARMOUR_ERR_ARG;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = int(sizeof(armour_config));
    cfg->device = 0;
    cfg->robot_model = 0;
    cfg->num_time_steps = 128;
    cfg->max_obstacles = 40;
    cfg->max_problems = 1;
    cfg->cap_link_monomials = 32;
    cfg->cap_torque_monomials = 64;
    cfg->cap_work_monomials = 768;
    cfg->simplify_threshold = 5e-4;
    for (int i = 0; i < NF; i++) cfg->k_range[i] = M_PI / 48;
    cfg->mass_uncertainty = -1;
    cfg->inertia_uncertainty = -1;
    return ARMOUR_OK;
}

int armour_ctx_create(const armour_config* cfg, armour_ctx** out) {
    if (!cfg || !out) return ARMOUR_ERR_ARG;
    *out = nullptr;
    if (cfg->struct_size != int(sizeof(armour_config))) return ARMOUR_ERR_ARG;
    if (cfg->num_time_steps < TB || cfg->num_time_steps > 128 || cfg->num_time_steps % TB != 0) return ARMOUR_ERR_ARG;
    if (cfg->max_problems < 1 || cfg->max_obstacles < 0) return ARMOUR_ERR_ARG;
    if (cfg->robot_model < 0 || cfg->robot_model > 1) return ARMOUR_ERR_ARG;
    if (cfg->cap_link_monomials < 1 || cfg->cap_torque_monomials < 1 || cfg->cap_work_monomials < 64)
        return ARMOUR_ERR_ARG;
    // the stored tables are fetched with 16-byte bulk copies (keys are 2 bytes): capacities in multiples of 8
    if (cfg->cap_link_monomials % 8 != 0 || cfg->cap_torque_monomials % 8 != 0) return ARMOUR_ERR_ARG;
    if (cfg->max_problems > 65535) return ARMOUR_ERR_ARG;  // the problem index is gridDim.y of the constraint kernels
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device >= ndev) return ARMOUR_ERR_CUDA;
    armour_ctx* ctx = new (std::nothrow) armour_ctx();
    if (!ctx) return ARMOUR_ERR_NOMEM;
    ctx->cfg = *cfg;
    ctx->rc = make_robot_constants(cfg->robot_model);
    ctx->rc.num_time_steps = cfg->num_time_steps;
    ctx->rc.simplify_threshold = cfg->simplify_threshold;
    for (int i = 0; i < NF; i++) ctx->rc.k_range[i] = cfg->k_range[i];
    if (cfg->mass_uncertainty >= 0) ctx->rc.mass_uncertainty = cfg->mass_uncertainty;
    if (cfg->inertia_uncertainty >= 0) ctx->rc.inertia_uncertainty = cfg->inertia_uncertainty;

    auto bail = [&](const char* what, cudaError_t e) {
        std::fprintf(stderr, "armour_ctx_create: %s: %s\n", what, cudaGetErrorString(e));
        armour_ctx_destroy(ctx);
        return ARMOUR_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail("cudaStreamCreate", e);
    ctx->stream = ctx->own_stream;
    if (ensure_constants(ctx) != ARMOUR_OK) return bail("upload_constants", cudaGetLastError());

    Batch& B = ctx->B;
    std::memset(&B, 0, sizeof(B));
    B.T = cfg->num_time_steps;
    B.NJ = ctx->rc.num_joints;
    B.capL = cfg->cap_link_monomials;
    B.capU = cfg->cap_torque_monomials;
    const size_t P = size_t(cfg->max_problems), T = size_t(B.T), NJ = size_t(B.NJ);
#define ALLOC(ptr, n) \
    if ((e = dalloc(&ptr, (n))) != cudaSuccess) return bail(#ptr, e)
    ALLOC(ctx->d_in, 3 * P * NF);
    ALLOC(B.link_n, P * T * NJ);
    ALLOC(B.link_c, P * T * NJ * 3);
    ALLOC(B.link_key, P * T * NJ * B.capL);
    ALLOC(B.link_g, P * T * NJ * B.capL * 3);
    ALLOC(B.u_n, P * T * NF);
    ALLOC(B.u_c, P * T * NF);
    ALLOC(B.u_r, P * T * NF);
    ALLOC(B.u_key, P * T * NF * B.capU);
    ALLOC(B.u_g, P * T * NF * B.capU);
    ALLOC(B.torque_radius, P * NF * T);
    ALLOC(B.link_gens, P * T * NJ * 18);
    ALLOC(B.link_r, P * T * NJ * 3);
    ALLOC(B.link_sliced, P * T * NJ * 3);
    ALLOC(B.status, P);
    ALLOC(ctx->d_unit_flag, P * T);
    ALLOC(ctx->d_k, P * NF);
    ALLOC(ctx->d_verdict, 2 * P);
#undef ALLOC
    B.q0 = ctx->d_in;
    B.qd0 = ctx->d_in + P * NF;
    B.qdd0 = ctx->d_in + 2 * P * NF;
    if ((e = cudaMemsetAsync(B.status, 0, P * sizeof(int), ctx->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMemsetAsync(ctx->d_unit_flag, 0, P * T * sizeof(int), ctx->stream)) != cudaSuccess) return bail("memset", e);
    // no table may ever be read with an uninitialised count (a failed build leaves its tables untouched)
    if ((e = cudaMemsetAsync(B.link_n, 0, P * T * NJ * sizeof(int), ctx->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMemsetAsync(B.u_n, 0, P * T * NF * sizeof(int), ctx->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = k1lat::k1_scratch_create(&ctx->k1_lat, ctx->cfg, ctx->rc, ctx->stream)) != cudaSuccess) return bail("k1 scratch", e);
    if ((long long)cfg->max_problems * B.T > 2LL * ctx->k1_lat.grid) {  // this context can see batches: set up the throughput kernel too
        if ((e = k1thr::k1_scratch_create(&ctx->k1_thr, ctx->cfg, ctx->rc, ctx->stream)) != cudaSuccess) return bail("k1 throughput scratch", e);
        ctx->k1_thr_ready = true;
        cudaFuncAttributes fa2;
        if ((e = cudaFuncGetAttributes(&fa2, k1thr::k_reachsets)) != cudaSuccess) return bail("load k_reachsets (throughput)", e);
    }
    {   // load the kernels now (CUDA loads modules lazily at first use), so that the first build is not charged for it
        cudaFuncAttributes fa;
        if ((e = cudaFuncGetAttributes(&fa, k1lat::k_reachsets)) != cudaSuccess) return bail("load k_reachsets", e);
        if ((e = cudaFuncGetAttributes(&fa, k_hyperplanes)) != cudaSuccess) return bail("load k_hyperplanes", e);
        if ((e = cudaFuncGetAttributes(&fa, k_constraints)) != cudaSuccess) return bail("load k_constraints", e);
        if ((e = cudaFuncSetAttribute(k_constraints, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TB * (MAXJ * K3_LTAB_BYTES + NF * K3_UTAB_BYTES) + K3_CAND_ARENA)) != cudaSuccess)
            return bail("k_constraints shared memory", e);
        if ((e = cudaFuncGetAttributes(&fa, k_constraints_slow)) != cudaSuccess) return bail("load k_constraints_slow", e);
        if ((e = cudaFuncGetAttributes(&fa, k_verdict)) != cudaSuccess) return bail("load k_verdict", e);
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail("sync", e);
    *out = ctx;
    return ARMOUR_OK;
}

int armour_ctx_destroy(armour_ctx* ctx) {
    if (!ctx) return ARMOUR_OK;
    cudaSetDevice(ctx->cfg.device);
    Batch& B = ctx->B;
    void* ptrs[] = {ctx->d_solver, ctx->d_solver_i, ctx->d_solver_io, ctx->d_unit_flag, ctx->d_jnz, ctx->d_jrs, ctx->d_ga, ctx->d_ja,
                    ctx->d_in, ctx->d_obs, ctx->d_k, ctx->d_g, ctx->d_jac, ctx->d_verdict, B.link_n, B.link_c,
                    B.link_key, B.link_g, B.u_n, B.u_c, B.u_r, B.u_key, B.u_g, B.torque_radius, B.link_gens, B.link_r, B.hp_cand, B.hp_cnt, B.hp_slow,
                    B.link_sliced, B.status};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    k1lat::k1_scratch_destroy(&ctx->k1_lat);
    k1thr::k1_scratch_destroy(&ctx->k1_thr);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return ARMOUR_OK;
}

int armour_ctx_set_stream(armour_ctx* ctx, void* s) {
    if (!ctx) return ARMOUR_ERR_ARG;
    ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
    return ARMOUR_OK;
}
int armour_ctx_synchronize(armour_ctx* ctx) {
    if (!ctx) return ARMOUR_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}
const char* armour_last_error(const armour_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
long long armour_kernel_launches(const armour_ctx* ctx) { return ctx ? ctx->launches : 0; }
int armour_num_joints(const armour_ctx* ctx) { return ctx ? ctx->B.NJ : ARMOUR_ERR_ARG; }
int armour_num_time_steps(const armour_ctx* ctx) { return ctx ? ctx->B.T : ARMOUR_ERR_ARG; }
int armour_num_constraints(const armour_ctx* ctx, int nobs) {
    if (!ctx || nobs < 0) return ARMOUR_ERR_ARG;
    return NF * ctx->B.T + ctx->B.NJ * ctx->B.T * nobs + 4 * NF;
}

int armour_ctx_reserve(armour_ctx* ctx, int nprob, int nobs) {
    int rc = check_batch(ctx, nprob, nobs);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    const int O = ctx->B.O, np = ctx->B.nprob;
    rc = ensure_obstacle_buffers(ctx, nprob, nobs);
    ctx->B.O = O;  // reserving does not change the current batch
    ctx->B.nprob = np;
    if (rc) return rc;
    ctx->B.O = nobs;
    rc = ensure_eval_buffers(ctx, nprob, true, true);
    ctx->B.O = O;
    return rc;
}

// ---- build -----------------------------------------------------------------------------------------
int armour_batch_reachsets_build_device(armour_ctx* ctx, int nprob, const double* d_q0, const double* d_qd0,
                                        const double* d_qdd0, const double* d_obstacles, int nobs) {
    int rc = check_batch(ctx, nprob, nobs);
    if (rc) return rc;
    if (!d_q0 || !d_qd0 || !d_qdd0 || (nobs > 0 && !d_obstacles)) return fail(ctx, ARMOUR_ERR_ARG, "null input");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    rc = ensure_obstacle_buffers(ctx, nprob, nobs);
    if (rc) return rc;
    const size_t P = size_t(ctx->cfg.max_problems);
    ctx->B.epoch++;
    int nl = 0;
    bool latency = false;
    if (nobs > 0) {
        // the kernels read the caller's buffers; k_hyperplanes copies them into the context's own (for the evaluations that
        // follow) — no copies in front of the build.  In the latency configuration k_hyperplanes is a programmatic dependent
        // of k_reachsets and starts on the intervals that are complete.
        Batch Bb = ctx->B;
        Bb.q0 = d_q0;
        Bb.qd0 = d_qd0;
        Bb.qdd0 = d_qdd0;
        Bb.obstacles = d_obstacles;
        { int rc2 = launch_build(ctx, Bb, &nl, &latency); if (rc2) return rc2; }
        HpStage stage{ctx->d_in, ctx->d_in + P * NF, ctx->d_in + 2 * P * NF, ctx->d_obs};
        CU(launch_hyperplanes(Bb, ctx->stream, latency ? ctx->d_unit_flag : nullptr, stage));
        nl++;
    } else {
        const size_t nb = size_t(nprob) * NF * sizeof(double);
        CU(cudaMemcpyAsync(ctx->d_in, d_q0, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_in + P * NF, d_qd0, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_in + 2 * P * NF, d_qdd0, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        { int rc2 = launch_build(ctx, ctx->B, &nl, &latency); if (rc2) return rc2; }
    }
    ctx->launches += nl;
    ctx->built_nprob = nprob;
    ctx->h_torque_valid = false;
    ctx->h_q0.clear();
    return ARMOUR_OK;
}

int armour_batch_reachsets_build(armour_ctx* ctx, int nprob, const double* q0, const double* qd0, const double* qdd0,
                                 const double* obstacles, int nobs) {
    int rc = check_batch(ctx, nprob, nobs);
    if (rc) return rc;
    if (!q0 || !qd0 || !qdd0 || (nobs > 0 && !obstacles)) return fail(ctx, ARMOUR_ERR_ARG, "null input");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    rc = ensure_obstacle_buffers(ctx, nprob, nobs);
    if (rc) return rc;
    const size_t P = size_t(ctx->cfg.max_problems);
    const size_t nb = size_t(nprob) * NF * sizeof(double);
    CU(cudaMemcpyAsync(ctx->d_in, q0, nb, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_in + P * NF, qd0, nb, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_in + 2 * P * NF, qdd0, nb, cudaMemcpyHostToDevice, ctx->stream));
    if (nobs > 0)
        CU(cudaMemcpyAsync(ctx->d_obs, obstacles, size_t(nprob) * nobs * 12 * sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
    ctx->B.epoch++;
    int nl = 0;
    bool latency = false;
    { int rc2 = launch_build(ctx, ctx->B, &nl, &latency); if (rc2) return rc2; }
    ctx->launches += nl;
    CU(launch_hyperplanes(ctx->B, ctx->stream, (latency && nobs > 0) ? ctx->d_unit_flag : nullptr));
    ctx->launches += (nobs > 0);
    ctx->built_nprob = nprob;
    ctx->h_torque_valid = false;
    ctx->h_q0.assign(q0, q0 + size_t(nprob) * NF);
    ctx->h_qd0.assign(qd0, qd0 + size_t(nprob) * NF);
    ctx->h_qdd0.assign(qdd0, qdd0 + size_t(nprob) * NF);
    // surface capacity overflows of the build (never truncate silently)
    std::vector<int> st(nprob);
    CU(cudaMemcpyAsync(st.data(), ctx->B.status, nprob * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < nprob; p++)
        if ((st[p] >> 3) == ctx->B.epoch && (st[p] & 7) != 0) {
            // which check failed first, and where (diagnostics of the build kernel: stats[4..6])
            int site[3] = {0, 0, 0};
            const int* dstats = latency ? ctx->k1_lat.stats : ctx->k1_thr.stats;
            std::string where;
            if (dstats && cudaMemcpy(site, dstats + 4, sizeof(site), cudaMemcpyDeviceToHost) == cudaSuccess && site[2] == ctx->B.epoch) {
                const int line = site[0] >> 3;
                where = "; first failing check: code " + std::to_string(site[0] & 7) + " at " +
                        (line >= 10000 ? "k1_reachsets.cuh:" + std::to_string(line - 10000) : "k1_pz.cuh:" + std::to_string(line)) +
                        ", problem " + std::to_string(site[1] / ctx->B.T) + " interval " + std::to_string(site[1] % ctx->B.T);
            }
            return fail(ctx, ARMOUR_ERR_CAPACITY,
                        "reach-set build overflowed a monomial table (problem " + std::to_string(p) + ", code " +
                            std::to_string(st[p] & 7) + ")" + where);
        }
    return ARMOUR_OK;
}

int armour_reachsets_build(armour_ctx* ctx, const double* q0, const double* qd0, const double* qdd0,
                           const double* obstacles, int nobs) {
    return armour_batch_reachsets_build(ctx, 1, q0, qd0, qdd0, obstacles, nobs);
}

int armour_batch_get_build_status(armour_ctx* ctx, int nprob, int* out) {
    if (!ctx || !out || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    CU(cudaMemcpyAsync(out, ctx->B.status, nprob * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < nprob; p++)  // device words are stamped with the build ordinal: report this build's code only
        out[p] = ((out[p] >> 3) == ctx->B.epoch && (out[p] & 7) != 0) ? ARMOUR_ERR_CAPACITY : ARMOUR_OK;
    return ARMOUR_OK;
}

int armour_batch_get_monomial_counts(armour_ctx* ctx, int nprob, int* link_n, int* u_n) {
    if (!ctx || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    const size_t T = ctx->B.T, NJ = ctx->B.NJ;
    if (link_n)
        CU(cudaMemcpyAsync(link_n, ctx->B.link_n, nprob * T * NJ * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (u_n) CU(cudaMemcpyAsync(u_n, ctx->B.u_n, nprob * T * NF * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

int armour_batch_get_candidate_counts(armour_ctx* ctx, int nprob, unsigned char* out) {
    if (!ctx || !out || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    const size_t rows = size_t(ctx->B.NJ) * ctx->B.T * ctx->B.O;
    if (rows) CU(cudaMemcpyAsync(out, ctx->B.hp_cnt, nprob * rows, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

int armour_chunk_intervals(void) { return TB; }

#ifdef K1_PROFILE
// developer builds only: per-(interval, operation site) cycle counts of the last single-problem builds
extern "C" int armour_debug_k1_profile(long long* out, int reset) {
    cudaDeviceSynchronize();
    if (out) {
        int sp[8];
        if (cudaMemcpyFromSymbol(sp, g_k1spill, sizeof(sp)) == cudaSuccess)
            std::printf("k1 profile: %d of %d arena blocks in global memory, %d scratch pools in global memory, %d cross products on "
                        "the table path (largest term count %d)\n", sp[0], sp[1], sp[2], sp[3], sp[4]);
    }
    if (out && cudaMemcpyFromSymbol(out, g_k1prof, sizeof(long long) * 128 * K1_PROF_SITES * 2) != cudaSuccess) return ARMOUR_ERR_CUDA;
    if (reset) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_k1prof) != cudaSuccess) return ARMOUR_ERR_CUDA;
        if (cudaMemset(p, 0, sizeof(long long) * 128 * K1_PROF_SITES * 2) != cudaSuccess) return ARMOUR_ERR_CUDA;
    }
    return ARMOUR_OK;
}
#endif

int armour_measure_fp64_peak(armour_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return ARMOUR_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device));
    double* d_out = nullptr;
    CU(dalloc(&d_out, 1));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8, block = 256;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, ctx->stream));
        k_fp64_peak<<<grid, block, 0, ctx->stream>>>(d_out, iters, 1.0000001);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * double(iters) * double(grid) * block;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return ARMOUR_OK;
}

// ---- evaluate --------------------------------------------------------------------------------------
int armour_batch_eval_device(armour_ctx* ctx, int nprob, const double* d_k, double* d_g, double* d_values) {
    if (!ctx || !d_k) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "evaluate before build");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    Batch B = ctx->B;
    B.nprob = nprob;
    CU(launch_constraints(B, d_k, d_g, d_values, ctx->stream));
    ctx->launches += 1 + (B.O > 0);
    return ARMOUR_OK;
}

int armour_batch_eval(armour_ctx* ctx, int nprob, const double* k, double* g, double* values) {
    if (!ctx || !k) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "evaluate before build");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    int rc = ensure_eval_buffers(ctx, nprob, g != nullptr, values != nullptr);
    if (rc) return rc;
    const size_t m = size_t(ctx->B.m());
    CU(cudaMemcpyAsync(ctx->d_k, k, size_t(nprob) * NF * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = armour_batch_eval_device(ctx, nprob, ctx->d_k, g ? ctx->d_g : nullptr, values ? ctx->d_jac : nullptr);
    if (rc) return rc;
    if (g) CU(cudaMemcpyAsync(g, ctx->d_g, nprob * m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (values)
        CU(cudaMemcpyAsync(values, ctx->d_jac, nprob * m * NF * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

// ---- structured Jacobian (offered beside the dense one; the reference declares its Jacobian dense) ------------------
long long armour_jacobian_nnz(const armour_ctx* ctx, int nobs) {
    if (!ctx || nobs < 0) return ARMOUR_ERR_ARG;
    return jac_nnz(ctx->B.T, ctx->B.NJ, nobs);
}

int armour_jacobian_structure(const armour_ctx* ctx, int nobs, int* iRow, int* jCol) {
    if (!ctx || nobs < 0 || !iRow || !jCol) return ARMOUR_ERR_ARG;
    const int T = ctx->B.T, NJ = ctx->B.NJ;
    size_t n = 0;
    for (int r = 0; r < NF * T; r++)
        for (int c = 0; c < NF; c++, n++) {
            iRow[n] = r;
            jCol[n] = c;
        }
    for (int l = 0; l < NJ; l++)
        for (int r = 0; r < T * nobs; r++)
            for (int c = 0; c < jac_link_width(l); c++, n++) {
                iRow[n] = NF * T + l * T * nobs + r;
                jCol[n] = c;
            }
    for (int r = 0; r < 4 * NF; r++, n++) {
        iRow[n] = NF * T + NJ * T * nobs + r;
        jCol[n] = r % NF;
    }
    return ARMOUR_OK;
}

int armour_batch_eval_structured_device(armour_ctx* ctx, int nprob, const double* d_k, double* d_g, double* d_values_nnz) {
    if (!ctx || !d_k || !d_values_nnz) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "evaluate before build");
    CU(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_eval_buffers(ctx, nprob, false, true);
    if (rc) return rc;
    rc = armour_batch_eval_device(ctx, nprob, d_k, d_g, ctx->d_jac);
    if (rc) return rc;
    Batch B = ctx->B;
    B.nprob = nprob;
    CU(launch_pack_jacobian(B, ctx->d_jac, d_values_nnz, ctx->stream));
    ctx->launches += 1;
    return ARMOUR_OK;
}

int armour_batch_eval_structured(armour_ctx* ctx, int nprob, const double* k, double* g, double* values_nnz) {
    if (!ctx || !k || !values_nnz) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "evaluate before build");
    CU(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_eval_buffers(ctx, nprob, g != nullptr, true);
    if (rc) return rc;
    const size_t m = size_t(ctx->B.m());
    const size_t nnz = size_t(jac_nnz(ctx->B.T, ctx->B.NJ, ctx->B.O));
    if (ctx->jnz_capacity < nprob * nnz) {
        if (ctx->d_jnz) cudaFree(ctx->d_jnz);
        ctx->d_jnz = nullptr;
        ctx->jnz_capacity = 0;
        CU(dalloc(&ctx->d_jnz, nprob * nnz));
        ctx->jnz_capacity = nprob * nnz;
    }
    CU(cudaMemcpyAsync(ctx->d_k, k, size_t(nprob) * NF * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = armour_batch_eval_structured_device(ctx, nprob, ctx->d_k, g ? ctx->d_g : nullptr, ctx->d_jnz);
    if (rc) return rc;
    if (g) CU(cudaMemcpyAsync(g, ctx->d_g, nprob * m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(values_nnz, ctx->d_jnz, nprob * nnz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

int armour_eval_jac_g_structured(armour_ctx* ctx, const double* k, double* values_nnz) {
    return armour_batch_eval_structured(ctx, 1, k, nullptr, values_nnz);
}

int armour_eval_g(armour_ctx* ctx, const double* k, double* g) { return armour_batch_eval(ctx, 1, k, g, nullptr); }
int armour_eval_jac_g(armour_ctx* ctx, const double* k, double* values) {
    return armour_batch_eval(ctx, 1, k, nullptr, values);
}
int armour_eval_g_jac(armour_ctx* ctx, const double* k, double* g, double* values) {
    return armour_batch_eval(ctx, 1, k, g, values);
}

int armour_batch_verdict_device(armour_ctx* ctx, int nprob, const double* d_g, int* d_feasible, int* d_first) {
    if (!ctx || !d_g || !d_feasible || !d_first) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "verdict before build");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    Batch B = ctx->B;
    B.nprob = nprob;
    CU(launch_verdict(B, d_g, d_feasible, d_first, ctx->stream));
    ctx->launches += 1;
    return ARMOUR_OK;
}

// ---- batched device solver -------------------------------------------------------------
void armour_solver_options_default(armour_solver_options* opt) {
    if (!opt) return;
    opt->max_iter = 60;
    opt->tol = 1e-4;
    opt->torque_tol = 1e-2;
    opt->collision_tol = 1e-4;
    opt->qp_sweeps = 200;
    opt->qp_update_budget = 16384;
}

int armour_batch_solve_device(armour_ctx* ctx, int nprob, const double* d_q_des, const armour_solver_options* opt_in,
                              double* d_k_opt, int* d_feasible, int* d_first, int* d_iters) {
    if (!ctx || !d_q_des || !d_k_opt || !d_feasible || !d_first) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "solve before build");
    armour_solver_options opt;
    armour_solver_options_default(&opt);
    if (opt_in) opt = *opt_in;
    if (opt.max_iter < 1 || !(opt.tol > 0) || opt.qp_sweeps < 1) return fail(ctx, ARMOUR_ERR_ARG, "solver options");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    Batch B = ctx->B;
    B.nprob = nprob;
    const size_t m = size_t(B.m());
    int rc = ensure_eval_buffers(ctx, nprob, true, true);
    if (rc) return rc;
    // scratch: state arrays + second g buffer + linearised rows.  Its size depends on nprob AND on m (obstacle count of
    // the current batch): capacity is tracked in elements and the layout below always uses this call's nprob.
    const size_t P = size_t(nprob);
    const size_t need_d = P * (3 * NF + 4) + P * m + P * SOLVER_ROWCAP * SOLVER_ROWW;
    const size_t need_i = 8 * P + 1;
    if (ctx->solver_doubles < need_d) {
        if (ctx->d_solver) cudaFree(ctx->d_solver);
        ctx->d_solver = nullptr;
        ctx->solver_doubles = 0;
        CU(dalloc(&ctx->d_solver, need_d));
        ctx->solver_doubles = need_d;
    }
    if (ctx->solver_ints < need_i) {
        if (ctx->d_solver_i) cudaFree(ctx->d_solver_i);
        ctx->d_solver_i = nullptr;
        ctx->solver_ints = 0;
        CU(dalloc(&ctx->d_solver_i, need_i));
        ctx->solver_ints = need_i;
    }
    SolverState S;
    double* w = ctx->d_solver;
    S.x = w; w += P * NF;
    S.xt = w; w += P * NF;
    S.best = w; w += P * NF;
    S.f = w; w += P;
    S.viol = w; w += P;
    S.fbest = w; w += P;
    S.delta = w; w += P;
    double* d_gt = w; w += P * m;
    S.rows = w;
    S.have_best = ctx->d_solver_i;
    S.status = S.have_best + P;
    S.iters = S.status + P;
    S.evals = S.iters + P;
    S.dbg = S.evals + P;
    int* d_running = S.dbg + 3 * P;
    int* d_list = d_running + 1;  // [P] problems still running
    S.q_des = d_q_des;
    S.tol = opt.tol;
    S.torque_tol = opt.torque_tol;
    S.collision_tol = opt.collision_tol;
    S.max_iter = opt.max_iter;
    S.qp_sweeps = opt.qp_sweeps;
    S.qp_update_budget = opt.qp_update_budget;
    cudaStream_t st = ctx->stream;
    const size_t step_smem = solver_step_smem(B.m());  // ballot masks of the row gathering: 64 B per 256 rows
    if (step_smem > 200 * 1024) return fail(ctx, ARMOUR_ERR_ARG, "too many constraint rows for the batched solver");
    if (step_smem > 40 * 1024)  // beyond the default limit (a few hundred obstacles): opt in, per device (cheap, idempotent)
        CU(cudaFuncSetAttribute(k_solver_step, cudaFuncAttributeMaxDynamicSharedMemorySize, int(step_smem)));
    // x = 0 (armtd_NLP::get_starting_point), g(0)
    CU(cudaMemsetAsync(S.x, 0, size_t(nprob) * NF * sizeof(double), st));
    CU(cudaMemsetAsync(S.xt, 0, size_t(nprob) * NF * sizeof(double), st));
    CU(launch_constraints(B, S.x, ctx->d_g, nullptr, st));
    k_solver_start<<<nprob, SOLVER_THREADS, 0, st>>>(B, S, ctx->d_g);
    CU(cudaGetLastError());
    ctx->launches += 2 + (B.O > 0);
    // the list of running problems is rebuilt every `compact_every` iterations (a host synchronisation each time; between two
    // rebuilds the solver kernels skip finished problems themselves, only the constraint kernel still evaluates them)
    static const int compact_every = [] {
        const char* v = std::getenv("ARMOUR_SOLVER_COMPACT_EVERY");  // developer knob
        const int n = v ? std::atoi(v) : 0;
        return n >= 1 ? n : 4;
    }();
    Batch Ba = B;      // launches over the problems still running (all of them at first)
    int nactive = nprob;
    for (int it = 0; it < opt.max_iter && nactive > 0; it++) {
        Ba.nprob = nactive;
        CU(launch_constraints(Ba, S.x, ctx->d_g, ctx->d_jac, st));
        k_solver_step<<<nactive, SOLVER_THREADS, step_smem, st>>>(Ba, S, ctx->d_g, ctx->d_jac, it);
        CU(launch_constraints(Ba, S.xt, d_gt, nullptr, st));
        k_solver_accept<<<nactive, SOLVER_THREADS, 0, st>>>(Ba, S, d_gt, it == opt.max_iter - 1 ? 1 : 0);
        CU(cudaGetLastError());  // (covers k_solver_step too: launch errors are sticky until read)
        ctx->launches += 4 + 2 * (B.O > 0);
        if ((it + 1) % compact_every == 0 && it + 1 < opt.max_iter) {  // every fourth iteration: who is still running?
            CU(cudaMemsetAsync(d_running, 0, sizeof(int), st));
            k_solver_compact<<<(nprob + 255) / 256, 256, 0, st>>>(S, nprob, d_running, d_list);
            CU(cudaMemcpyAsync(&nactive, d_running, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            ctx->launches += 1;
            Ba.plist = d_list;
        }
    }
    k_solver_final<<<(nprob + 255) / 256, 256, 0, st>>>(S, nprob, d_k_opt, nullptr);
    CU(launch_constraints(B, d_k_opt, ctx->d_g, nullptr, st));
    CU(launch_verdict(B, ctx->d_g, d_feasible, d_first, st));
    if (d_iters) CU(cudaMemcpyAsync(d_iters, S.iters, size_t(nprob) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (std::getenv("ARMOUR_SOLVER_DEBUG")) {  // developer aid: row / active-set iteration / drop counts of the last step per problem
        std::vector<int> dbg(size_t(nprob) * 3);
        CU(cudaMemcpyAsync(dbg.data(), S.dbg, dbg.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        long long r = 0, sw = 0, mv = 0;
        int rmax = 0, mvmax = 0;
        for (int p = 0; p < nprob; p++) {
            r += dbg[p * 3];
            sw += dbg[p * 3 + 1];
            mv += dbg[p * 3 + 2];
            rmax = std::max(rmax, dbg[p * 3]);
            mvmax = std::max(mvmax, dbg[p * 3 + 2]);
        }
        std::printf("solver debug: last step per problem: rows mean %.0f max %d, active-set iterations mean %.1f, dropped rows mean %.1f max %d\n",
                    double(r) / nprob, rmax, double(sw) / nprob, double(mv) / nprob, mvmax);
    }
    ctx->launches += 3;
    CU(cudaGetLastError());
    return ARMOUR_OK;
}

int armour_batch_solve(armour_ctx* ctx, int nprob, const double* q_des, const armour_solver_options* opt, double* k_opt,
                       int* feasible, int* first_violation, int* iterations) {
    if (!ctx || !q_des || !k_opt || !feasible) return ARMOUR_ERR_ARG;
    if (nprob < 1 || nprob > ctx->built_nprob) return fail(ctx, ARMOUR_ERR_STATE, "solve before build");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    if (!ctx->d_solver_io) CU(dalloc(&ctx->d_solver_io, size_t(ctx->cfg.max_problems) * (2 * NF + 2)));
    const size_t P = size_t(ctx->cfg.max_problems);
    double* d_q = ctx->d_solver_io;
    double* d_ko = d_q + P * NF;
    int* d_i = reinterpret_cast<int*>(d_ko + P * NF);  // feasible, first, iterations: 3 * P ints in 2 * P doubles
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(d_q, q_des, size_t(nprob) * NF * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = armour_batch_solve_device(ctx, nprob, d_q, opt, d_ko, d_i, d_i + P, d_i + 2 * P);
    if (rc) return rc;
    CU(cudaMemcpyAsync(k_opt, d_ko, size_t(nprob) * NF * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(feasible, d_i, size_t(nprob) * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (first_violation) CU(cudaMemcpyAsync(first_violation, d_i + P, size_t(nprob) * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iterations) CU(cudaMemcpyAsync(iterations, d_i + 2 * P, size_t(nprob) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ARMOUR_OK;
}

// ---- getters ---------------------------------------------------------------------------------------
int armour_batch_get_torque_radius(armour_ctx* ctx, int nprob, double* out) {
    if (!ctx || !out || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    int rc = fetch_torque_radius(ctx);
    if (rc) return rc;
    std::memcpy(out, ctx->h_torque_radius.data(), size_t(nprob) * NF * ctx->B.T * sizeof(double));
    return ARMOUR_OK;
}
int armour_get_torque_radius(armour_ctx* ctx, double* out) { return armour_batch_get_torque_radius(ctx, 1, out); }

int armour_batch_get_link_independent_generators(armour_ctx* ctx, int nprob, double* out) {
    if (!ctx || !out || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    CU(cudaMemcpyAsync(out, ctx->B.link_gens, size_t(nprob) * ctx->B.T * ctx->B.NJ * 18 * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}
int armour_get_link_independent_generators(armour_ctx* ctx, double* out) {
    return armour_batch_get_link_independent_generators(ctx, 1, out);
}

int armour_get_link_sliced_center(armour_ctx* ctx, double* out) {
    if (!ctx || !out || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    CU(cudaMemcpyAsync(out, ctx->B.link_sliced, size_t(ctx->B.T) * ctx->B.NJ * 3 * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

int armour_batch_get_bounds(armour_ctx* ctx, int nprob, double* g_l, double* g_u) {  // KPR/NLPclass.cu:116-165
    if (!ctx || !g_l || !g_u || nprob < 1 || nprob > ctx->built_nprob) return ARMOUR_ERR_ARG;
    int rc = fetch_torque_radius(ctx);
    if (rc) return rc;
    const RobotConstants& R = ctx->rc;
    const int T = ctx->B.T, NJ = ctx->B.NJ, O = ctx->B.O, m = ctx->B.m();
    for (int p = 0; p < nprob; p++) {
        double* gl = g_l + size_t(p) * m;
        double* gu = g_u + size_t(p) * m;
        const double* tr = ctx->h_torque_radius.data() + size_t(p) * NF * T;
        for (int t = 0; t < T; t++)
            for (int j = 0; j < NF; j++) {
                gl[t * NF + j] = -R.torque_limits[j] + tr[j * T + t];
                gu[t * NF + j] = R.torque_limits[j] - tr[j * T + t];
            }
        int off = NF * T;
        for (int i = off; i < off + T * NJ * O; i++) {
            gl[i] = -1e19;
            gu[i] = 0;
        }
        off += T * NJ * O;
        for (int rep = 0; rep < 2; rep++, off += NF)
            for (int i = 0; i < NF; i++) {
                gl[off + i] = R.state_limits_lb[i] + R.qe;
                gu[off + i] = R.state_limits_ub[i] - R.qe;
            }
        for (int rep = 0; rep < 2; rep++, off += NF)
            for (int i = 0; i < NF; i++) {
                gl[off + i] = -R.speed_limits[i] + R.qde;
                gu[off + i] = R.speed_limits[i] - R.qde;
            }
    }
    return ARMOUR_OK;
}
int armour_get_bounds(armour_ctx* ctx, double* g_l, double* g_u) { return armour_batch_get_bounds(ctx, 1, g_l, g_u); }

int armour_verdict(armour_ctx* ctx, const double* g, int* feasible, int* first_violation) {  // KPR/NLPclass.cu:449-537
    if (!ctx || !g || !feasible || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    int rc = fetch_torque_radius(ctx);
    if (rc) return rc;
    const RobotConstants& R = ctx->rc;
    const int T = ctx->B.T, NJ = ctx->B.NJ, O = ctx->B.O;
    const double* tr = ctx->h_torque_radius.data();
    auto done = [&](int ok, int row) {
        *feasible = ok;
        if (first_violation) *first_violation = row;
        return ARMOUR_OK;
    };
    for (int t = 0; t < T; t++)
        for (int j = 0; j < NF; j++) {
            const double v = g[t * NF + j];
            if (v < -R.torque_limits[j] + tr[j * T + t] - R.torque_violation_threshold ||
                v > R.torque_limits[j] - tr[j * T + t] + R.torque_violation_threshold)
                return done(0, t * NF + j);
        }
    int off = NF * T;
    for (int i = 0; i < NJ * T * O; i++)
        if (g[off + i] > R.collision_violation_threshold) return done(0, off + i);
    off += NJ * T * O;
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = 0; i < NF; i++)
            if (g[off + i] < R.state_limits_lb[i] + R.qe || g[off + i] > R.state_limits_ub[i] - R.qe)
                return done(0, off + i);
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = 0; i < NF; i++)
            if (g[off + i] < -R.speed_limits[i] + R.qde || g[off + i] > R.speed_limits[i] - R.qde)
                return done(0, off + i);
    return done(1, -1);
}

int armour_cost(armour_ctx* ctx, const double* q_des, const double* k, double* obj, double* grad) {
    if (!ctx || !q_des || !k || ctx->built_nprob < 1 || ctx->h_q0.empty()) return ARMOUR_ERR_ARG;
    const RobotConstants& R = ctx->rc;
    auto wrap = [](double a) {  // KPR/NLPclass.cu:6-15
        while (a < -M_PI) a += 2 * M_PI;
        while (a > M_PI) a -= 2 * M_PI;
        return a;
    };
    const double tp = R.t_plan, D = R.duration;
    double qp[NF];
    for (int i = 0; i < NF; i++)
        qp[i] = bez_q(ctx->h_q0[i], ctx->h_qd0[i] * D, ctx->h_qdd0[i] * D * D, R.k_range[i] * k[i], tp);
    if (obj) {  // KPR/NLPclass.cu:222-233
        double v = pw2(wrap(q_des[0] - qp[0])) + pw2(wrap(q_des[2] - qp[2])) + pw2(wrap(q_des[4] - qp[4])) +
                   pw2(wrap(q_des[6] - qp[6])) + pw2(q_des[1] - qp[1]) + pw2(q_des[3] - qp[3]) + pw2(q_des[5] - qp[5]);
        *obj = v * R.cost_scale;
    }
    if (grad) {  // KPR/NLPclass.cu:252-264
        for (int i = 0; i < NF; i++) {
            const double dk = pw3(tp) * (6 * pw2(tp) - 15 * tp + 10) * R.k_range[i];
            grad[i] = (i % 2 == 0) ? (2 * wrap(qp[i] - q_des[i]) * dk) : (2 * (qp[i] - q_des[i]) * dk);
            grad[i] *= R.cost_scale;
        }
    }
    return ARMOUR_OK;
}

// ---- reach-set tables ------------------------------------------------------------------------------
int armour_export_reachsets(armour_ctx* ctx, int prob, armour_reachset_tables* out) {
    if (!ctx || !out || prob < 0 || prob >= ctx->built_nprob) return ARMOUR_ERR_ARG;
    const Batch& B = ctx->B;
    const size_t T = B.T, NJ = B.NJ;
    if (out->cap_link < B.capL || out->cap_u < B.capU) return fail(ctx, ARMOUR_ERR_ARG, "export caps too small");
    std::vector<uint16_t> lk(T * NJ * B.capL), uk(T * NF * B.capU);
    std::vector<double> lg(T * NJ * B.capL * 3), ug(T * NF * B.capU);
    cudaStream_t st = ctx->stream;
#define D2H(dst, src, n) CU(cudaMemcpyAsync(dst, src, (n), cudaMemcpyDeviceToHost, st))
    D2H(out->link_n, B.link_n + prob * T * NJ, T * NJ * sizeof(int));
    D2H(out->link_center, B.link_c + prob * T * NJ * 3, T * NJ * 3 * sizeof(double));
    D2H(lk.data(), B.link_key + prob * T * NJ * B.capL, lk.size() * sizeof(uint16_t));
    D2H(lg.data(), B.link_g + prob * T * NJ * B.capL * 3, lg.size() * sizeof(double));
    D2H(out->u_n, B.u_n + prob * T * NF, T * NF * sizeof(int));
    D2H(out->u_center, B.u_c + prob * T * NF, T * NF * sizeof(double));
    D2H(uk.data(), B.u_key + prob * T * NF * B.capU, uk.size() * sizeof(uint16_t));
    D2H(ug.data(), B.u_g + prob * T * NF * B.capU, ug.size() * sizeof(double));
    if (out->u_radius) D2H(out->u_radius, B.u_r + prob * T * NF, T * NF * sizeof(double));
    if (out->torque_radius) D2H(out->torque_radius, B.torque_radius + prob * NF * T, NF * T * sizeof(double));
    if (out->link_gens) D2H(out->link_gens, B.link_gens + prob * T * NJ * 18, T * NJ * 18 * sizeof(double));
#undef D2H
    CU(cudaStreamSynchronize(st));
    // entries past the monomial count are padding: export them as zero
    for (size_t i = 0; i < T * NJ; i++)
        for (int mI = 0; mI < out->cap_link; mI++) {
            const bool valid = mI < out->link_n[i] && mI < B.capL;
            out->link_key[i * out->cap_link + mI] = valid ? lk[i * B.capL + mI] : 0;
            for (int e = 0; e < 3; e++)
                out->link_coeff[(i * out->cap_link + mI) * 3 + e] = valid ? lg[(i * B.capL + mI) * 3 + e] : 0.0;
        }
    for (size_t i = 0; i < T * NF; i++)
        for (int mI = 0; mI < out->cap_u; mI++) {
            const bool valid = mI < out->u_n[i] && mI < B.capU;
            out->u_key[i * out->cap_u + mI] = valid ? uk[i * B.capU + mI] : 0;
            out->u_coeff[i * out->cap_u + mI] = valid ? ug[i * B.capU + mI] : 0.0;
        }
    return ARMOUR_OK;
}

int armour_import_reachsets(armour_ctx* ctx, int prob, int nprob_total, const armour_reachset_tables* in,
                            const double* q0, const double* qd0, const double* qdd0, const double* obstacles, int nobs) {
    int rc = check_batch(ctx, nprob_total, nobs);
    if (rc) return rc;
    if (!in || !in->u_radius || !in->torque_radius || !in->link_gens || !q0 || !qd0 || !qdd0 || prob < 0 ||
        prob >= nprob_total)
        return ARMOUR_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    if (prob == 0 || ctx->B.O != nobs || ctx->B.nprob != nprob_total) {
        rc = ensure_obstacle_buffers(ctx, nprob_total, nobs);
        if (rc) return rc;
    }
    if (prob == 0) ctx->B.epoch++;  // imported tables: no build failure can be pending
    const Batch& B = ctx->B;
    const size_t T = B.T, NJ = B.NJ, P = size_t(ctx->cfg.max_problems);
    std::vector<uint16_t> lk(T * NJ * B.capL, 0), uk(T * NF * B.capU, 0);
    std::vector<double> lg(T * NJ * B.capL * 3, 0.0), ug(T * NF * B.capU, 0.0);
    for (size_t i = 0; i < T * NJ; i++) {
        if (in->link_n[i] > B.capL) return fail(ctx, ARMOUR_ERR_CAPACITY, "imported link table exceeds cap_link_monomials");
        for (int mI = 0; mI < in->link_n[i]; mI++) {
            lk[i * B.capL + mI] = uint16_t(in->link_key[i * in->cap_link + mI]);
            for (int e = 0; e < 3; e++) lg[(i * B.capL + mI) * 3 + e] = in->link_coeff[(i * in->cap_link + mI) * 3 + e];
        }
    }
    for (size_t i = 0; i < T * NF; i++) {
        if (in->u_n[i] > B.capU) return fail(ctx, ARMOUR_ERR_CAPACITY, "imported torque table exceeds cap_torque_monomials");
        for (int mI = 0; mI < in->u_n[i]; mI++) {
            uk[i * B.capU + mI] = uint16_t(in->u_key[i * in->cap_u + mI]);
            ug[i * B.capU + mI] = in->u_coeff[i * in->cap_u + mI];
        }
    }
    cudaStream_t st = ctx->stream;
#define H2D(dst, src, n) CU(cudaMemcpyAsync(dst, src, (n), cudaMemcpyHostToDevice, st))
    H2D(B.link_n + prob * T * NJ, in->link_n, T * NJ * sizeof(int));
    H2D(B.link_c + prob * T * NJ * 3, in->link_center, T * NJ * 3 * sizeof(double));
    H2D(B.link_key + prob * T * NJ * B.capL, lk.data(), lk.size() * sizeof(uint16_t));
    H2D(B.link_g + prob * T * NJ * B.capL * 3, lg.data(), lg.size() * sizeof(double));
    H2D(B.u_n + prob * T * NF, in->u_n, T * NF * sizeof(int));
    H2D(B.u_c + prob * T * NF, in->u_center, T * NF * sizeof(double));
    H2D(B.u_key + prob * T * NF * B.capU, uk.data(), uk.size() * sizeof(uint16_t));
    H2D(B.u_g + prob * T * NF * B.capU, ug.data(), ug.size() * sizeof(double));
    H2D(B.torque_radius + prob * NF * T, in->torque_radius, NF * T * sizeof(double));
    H2D(B.link_gens + prob * T * NJ * 18, in->link_gens, T * NJ * 18 * sizeof(double));
    std::vector<double> lr(T * NJ * 3);
    for (size_t i = 0; i < T * NJ; i++)
        for (int e = 0; e < 3; e++) lr[i * 3 + e] = in->link_gens[i * 18 + e + (3 + e) * 3];
    H2D(B.link_r + prob * T * NJ * 3, lr.data(), lr.size() * sizeof(double));
    H2D(B.u_r + prob * T * NF, in->u_radius, T * NF * sizeof(double));
    H2D(ctx->d_in + prob * NF, q0, NF * sizeof(double));
    H2D(ctx->d_in + P * NF + prob * NF, qd0, NF * sizeof(double));
    H2D(ctx->d_in + 2 * P * NF + prob * NF, qdd0, NF * sizeof(double));
    if (nobs > 0) H2D(ctx->d_obs + size_t(prob) * nobs * 12, obstacles, size_t(nobs) * 12 * sizeof(double));
#undef H2D
    CU(cudaStreamSynchronize(st));  // the staging vectors die at return
    if (prob == nprob_total - 1) {
        if (nobs > 0) CU(cudaMemsetAsync(ctx->B.hp_slow, 0, size_t(nprob_total) * sizeof(int), ctx->stream));
        CU(launch_hyperplanes(ctx->B, ctx->stream));
        ctx->launches += (nobs > 0);
        ctx->built_nprob = nprob_total;
        ctx->h_torque_valid = false;
    }
    if (prob == 0) {
        ctx->h_q0.assign(q0, q0 + NF);
        ctx->h_qd0.assign(qd0, qd0 + NF);
        ctx->h_qdd0.assign(qdd0, qdd0 + NF);
    }
    return ARMOUR_OK;
}


// ---- ARMTD comparison planner (csrc/armtd.cuh) ---------------------------------------------------------------
int armour_armtd_ctx_create(const armour_config* cfg_in, armour_ctx** out) {
    if (!out) return ARMOUR_ERR_ARG;
    armour_config cfg;
    if (cfg_in) {
        cfg = *cfg_in;
    } else {
        int rc = armour_config_default(&cfg);
        if (rc) return rc;
    }
    cfg.num_time_steps = ARMTD_T_PADDED;  // 100 intervals of KPA + 4 padding intervals (chunks of 8)
    cfg.max_problems = 1;
    cfg.robot_model = 0;  // KPA/Parameters.h:6: KinovaWithoutGripperInfo.h
    armour_ctx* ctx = nullptr;
    int rc = armour_ctx_create(&cfg, &ctx);
    if (rc) return rc;
    cudaError_t e = dalloc(&ctx->d_jrs, size_t(ARMTD_T_PADDED) * NF * 6);
    if (e != cudaSuccess) {
        armour_ctx_destroy(ctx);
        return ARMOUR_ERR_NOMEM;
    }
    ctx->armtd = true;
    ctx->B.jrs_ext = ctx->d_jrs;
    *out = ctx;
    return ARMOUR_OK;
}

int armour_armtd_num_constraints(const armour_ctx* ctx) {
    if (!ctx || !ctx->armtd) return ARMOUR_ERR_ARG;
    return ctx->B.NJ * ARMTD_T * ctx->B.O + 4 * NF;
}

int armour_armtd_build(armour_ctx* ctx, const double* q0, const double* qd0, const double* jrs, const double* k_range,
                       const double* obstacles, int nobs) {
    if (!ctx || !q0 || !qd0 || !jrs || !k_range) return ARMOUR_ERR_ARG;
    if (!ctx->armtd) return fail(ctx, ARMOUR_ERR_STATE, "not an ARMTD context (armour_armtd_ctx_create)");
    // ConstantAccelerationCurve::makePolyZono, KPA/Trajectory.cu:29-62: the cos / sin models of joint i over interval t
    std::vector<double> ext(size_t(ARMTD_T_PADDED) * NF * 6);
    auto at = [&](int a, int i, int t) { return jrs[(size_t(a) * NF + i) * ARMTD_T + t]; };
    for (int t = 0; t < ARMTD_T_PADDED; t++) {
        const int ts = t < ARMTD_T ? t : ARMTD_T - 1;  // padding intervals repeat the last one
        for (int i = 0; i < NF; i++) {
            const double cos_q0 = std::cos(q0[i]), sin_q0 = std::sin(q0[i]);
            double* o = &ext[(size_t(t) * NF + i) * 6];
            o[0] = cos_q0 * at(0, i, ts) - sin_q0 * at(3, i, ts);
            o[1] = cos_q0 * at(1, i, ts) - sin_q0 * at(4, i, ts);
            double e = std::fabs(cos_q0) * at(2, i, ts) + std::fabs(sin_q0) * at(5, i, ts);
            e *= 4.0;
            o[2] = e;
            o[3] = cos_q0 * at(3, i, ts) + sin_q0 * at(0, i, ts);
            o[4] = cos_q0 * at(4, i, ts) + sin_q0 * at(1, i, ts);
            e = std::fabs(cos_q0) * at(5, i, ts) + std::fabs(sin_q0) * at(2, i, ts);
            e *= 4.0;
            o[5] = e;
        }
    }
    CU(cudaSetDevice(ctx->cfg.device));
    CU(cudaMemcpyAsync(ctx->d_jrs, ext.data(), ext.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // ext dies at return
    for (int i = 0; i < NF; i++) {
        ctx->armtd_par.q0[i] = q0[i];
        ctx->armtd_par.qd0[i] = qd0[i];
        ctx->armtd_par.k_range[i] = k_range[i];
    }
    const double zero[NF] = {0, 0, 0, 0, 0, 0, 0};
    return armour_batch_reachsets_build(ctx, 1, q0, qd0, zero, obstacles, nobs);
}

int armour_armtd_eval(armour_ctx* ctx, const double* k, double* g, double* values) {
    if (!ctx || !k || (!g && !values)) return ARMOUR_ERR_ARG;
    if (!ctx->armtd || ctx->built_nprob < 1) return fail(ctx, ARMOUR_ERR_STATE, "evaluate before armour_armtd_build");
    CU(cudaSetDevice(ctx->cfg.device));
    { int rc_c = ensure_constants(ctx); if (rc_c) return rc_c; }
    int rc = ensure_eval_buffers(ctx, 1, true, values != nullptr);
    if (rc) return rc;
    const size_t ma = size_t(armour_armtd_num_constraints(ctx));
    if (ctx->ga_capacity < ma) {
        if (ctx->d_ga) cudaFree(ctx->d_ga);
        ctx->d_ga = nullptr;
        ctx->ga_capacity = 0;
        CU(dalloc(&ctx->d_ga, ma));
        ctx->ga_capacity = ma;
    }
    if (values && ctx->ja_capacity < ma * NF) {
        if (ctx->d_ja) cudaFree(ctx->d_ja);
        ctx->d_ja = nullptr;
        ctx->ja_capacity = 0;
        CU(dalloc(&ctx->d_ja, ma * NF));
        ctx->ja_capacity = ma * NF;
    }
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_k, k, NF * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = armour_batch_eval_device(ctx, 1, ctx->d_k, ctx->d_g, values ? ctx->d_jac : nullptr);
    if (rc) return rc;
    const int rows = ctx->B.NJ * ARMTD_T * ctx->B.O;
    const int blocks = rows > 0 ? (rows + 255) / 256 : 1;
    k_armtd_assemble<<<blocks, 256, 0, st>>>(ctx->armtd_par, ctx->B.NJ, ctx->B.O, ctx->d_k, ctx->d_g, values ? ctx->d_jac : nullptr,
                                             ctx->d_ga, values ? ctx->d_ja : nullptr);
    CU(cudaGetLastError());
    ctx->launches++;
    if (g) CU(cudaMemcpyAsync(g, ctx->d_ga, ma * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (values) CU(cudaMemcpyAsync(values, ctx->d_ja, ma * NF * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ARMOUR_OK;
}

int armour_armtd_get_bounds(armour_ctx* ctx, double* g_l, double* g_u) {  // KPA/NLPclass.cu:75-142
    if (!ctx || !g_l || !g_u || !ctx->armtd || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    const RobotConstants& R = ctx->rc;
    int off = ctx->B.NJ * ARMTD_T * ctx->B.O;
    for (int i = 0; i < off; i++) {
        g_l[i] = -1e19;
        g_u[i] = 0;
    }
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = 0; i < NF; i++) {
            g_l[off + i] = R.state_limits_lb[i] + R.qe;
            g_u[off + i] = R.state_limits_ub[i] - R.qe;
        }
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = 0; i < NF; i++) {
            g_l[off + i] = -R.speed_limits[i] + R.qde;
            g_u[off + i] = R.speed_limits[i] - R.qde;
        }
    return ARMOUR_OK;
}

int armour_armtd_verdict(armour_ctx* ctx, const double* g, int* feasible, int* first_violation) {  // KPA/NLPclass.cu:366-455
    if (!ctx || !g || !feasible || !ctx->armtd || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    const RobotConstants& R = ctx->rc;
    const int O = ctx->B.O;
    auto done = [&](int ok, int row) {
        *feasible = ok;
        if (first_violation) *first_violation = row;
        return ARMOUR_OK;
    };
    // the reference's loop runs over NUM_FACTORS - 1 links (KPA/NLPclass.cu:400): the rows of the last link are not checked
    for (int i = 0; i < NF - 1; i++)
        for (int j = 0; j < ARMTD_T; j++)
            for (int h = 0; h < O; h++)
                if (g[(i * ARMTD_T + j) * O + h] > R.collision_violation_threshold) return done(0, (i * ARMTD_T + j) * O + h);
    int off = ctx->B.NJ * ARMTD_T * O;
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = off; i < off + NF; i++)
            if (g[i] < R.state_limits_lb[i - off] + R.qe || g[i] > R.state_limits_ub[i - off] - R.qe) return done(0, i);
    for (int rep = 0; rep < 2; rep++, off += NF)
        for (int i = off; i < off + NF; i++)
            if (g[i] < -R.speed_limits[i - off] + R.qde || g[i] > R.speed_limits[i - off] - R.qde) return done(0, i);
    return done(1, -1);
}

int armour_armtd_cost(armour_ctx* ctx, const double* q_des, const double* k, double* obj, double* grad) {  // KPA/NLPclass.cu:178-243
    if (!ctx || !q_des || !k || !ctx->armtd || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    const ArmtdParams& A = ctx->armtd_par;
    auto wrap = [](double a) {
        while (a < -M_PI) a += 2 * M_PI;
        while (a > M_PI) a -= 2 * M_PI;
        return a;
    };
    double qp[NF];
    for (int i = 0; i < NF; i++) qp[i] = A.q0[i] + A.qd0[i] * 0.5 + A.k_range[i] * k[i] * 0.125;
    if (obj) {
        const double v = pw2(wrap(q_des[0] - qp[0])) + pw2(wrap(q_des[2] - qp[2])) + pw2(wrap(q_des[4] - qp[4])) +
                         pw2(wrap(q_des[6] - qp[6])) + pw2(q_des[1] - qp[1]) + pw2(q_des[3] - qp[3]) + pw2(q_des[5] - qp[5]);
        *obj = v * ctx->rc.cost_scale;
    }
    if (grad)
        for (int i = 0; i < NF; i++) {
            const double dk = A.k_range[i] * 0.125;
            grad[i] = (i % 2 == 0) ? (2 * wrap(qp[i] - q_des[i]) * dk) : (2 * (qp[i] - q_des[i]) * dk);
            grad[i] *= ctx->rc.cost_scale;
        }
    return ARMOUR_OK;
}

// armour_joint_position_center.out / armour_joint_position_radius.out of KPA/armtd_main.cu:232-256: [100][NJ][3] and [100][NJ][18]
int armour_armtd_get_link_sliced_center(armour_ctx* ctx, double* out) {
    if (!ctx || !out || !ctx->armtd || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    CU(cudaMemcpyAsync(out, ctx->B.link_sliced, size_t(ARMTD_T) * ctx->B.NJ * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}
int armour_armtd_get_link_independent_generators(armour_ctx* ctx, double* out) {
    if (!ctx || !out || !ctx->armtd || ctx->built_nprob < 1) return ARMOUR_ERR_ARG;
    CU(cudaMemcpyAsync(out, ctx->B.link_gens, size_t(ARMTD_T) * ctx->B.NJ * 18 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return ARMOUR_OK;
}

}  // extern "C"
