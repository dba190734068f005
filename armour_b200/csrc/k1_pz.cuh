// K1 building blocks: a sparse polynomial zonotope (PZ) algebra executed by one CTA on tables that live
// in shared memory.
//
// What it replaces: the reference's host-side PZsparse class (KPR/PZsparse.h:50-183, KPR/PZsparse.cu),
// whose every operation ends in simplify() = std::sort over heap-allocated monomials + merge + prune
// (KPR/PZsparse.cu:284-350).  Here a PZ is a block of doubles in a CTA-private arena
//     [centre: sz][radius lane 0: sz][radius lane 1: sz][keys: n x u64][coeff: n x sz]
// (sz = 1, 3 or 9 = scalar, 3-vector, 3x3 column-major) and simplify() is a pass through an
// open-addressing hash table keyed by the reference's 63-bit degree hash (KPR/PZsparse.h:23-40):
//   * keys are inserted with atomicMax ("larger key keeps the slot, the smaller one moves on") — the
//     ordered-hashing rule makes the final slot layout a function of the key SET only, never of the
//     thread schedule, so every result of this file is bitwise reproducible run to run;
//   * coefficients of equal keys are summed in a fixed order (one __syncthreads between the term
//     groups that can collide; two-term sums use a commutative atomic add);
//   * a finalize pass applies the reference's prune rule (Frobenius norm of the whole coefficient
//     <= threshold -> |coeff| moves into the radius) and compacts the survivors in slot order.
// Two radius lanes are carried so that the nominal and the interval-parameter Newton-Euler passes
// (KPR/Dynamics.h:41-47) are one pass: their polynomial parts are bit-identical (only radii differ).
// Coefficient / centre arithmetic is round-to-nearest without FMA contraction (-fmad=false), like
// the host reference; radii are accumulated with round-up intrinsics so the outer bound stays sound.
#include <cmath>
#include <cstdint>

#include "robot_constants.h"

#ifndef ARMOUR_EMU
#include <cuda_runtime.h>
#define K1_DI __device__ __forceinline__
#define K1_OP __device__ __noinline__
#else
#define K1_DI inline
#define K1_OP inline
#endif

// This file, k1_interval.cuh and k1_reachsets.cuh are compiled once per kernel configuration (armour_capi.cu
// includes k1_reachsets.cuh twice: a latency configuration and a throughput configuration), each time into
// its own namespace K1_NS; hence no include guards.
#ifndef K1_NS
#define K1_NS k1
#endif

namespace armour {
namespace K1_NS {

typedef unsigned long long u64;

#ifndef K1_NT
#define K1_NT 256
#endif
#ifndef K1_CTAS
#define K1_CTAS 2
#endif
#ifndef K1_GROUPS
#define K1_GROUPS 1
#endif
#ifndef K1_MG
#define K1_MG 0
#endif
constexpr int NT = K1_NT;    // threads working on one (problem, interval) unit
constexpr int GROUPS = K1_GROUPS;  // units built side by side by one CTA (CTA = GROUPS * NT threads), op by op in
                                   // step, so that the instruction stream of an operation is fetched once for all
// MG ("multi-group") mode: the GROUPS groups of a CTA work on the SAME unit.  The unit is cut into tasks (a few
// PZ operations each) whose operands travel through a per-CTA mailbox in global memory; every group claims the
// next task of a fixed priority list, waits for its inputs, runs it in its own arena and publishes the results
// (k1_reachsets.cuh, build_unit_mg).  Without MG the groups build different units in lock step.
constexpr bool MG = K1_MG != 0;
#ifndef K1_JRS_GLOBAL
#define K1_JRS_GLOBAL 0
#endif
constexpr int CTAS_PER_SM = K1_CTAS;  // resident CTAs per SM (shared memory is split evenly)
constexpr int NW = NT / 32;  // warps per CTA
static_assert(NT % 32 == 0 && NT >= 32, "a unit is built by whole warps");
static_assert(GROUPS <= 15, "one named barrier per group besides barrier 0 (16 per CTA)");
static_assert(GROUPS == 1 || NT >= 64, "several groups of ONE warp per CTA: not finished (profiles/r3_k1thr_experiments.md)");
constexpr int RED_STRIDE = 12;
constexpr int MASK_WORDS = 256;  // a merge operation handles up to 32 * MASK_WORDS candidate monomials

// failure codes written to Batch::status (a build never truncates silently)
enum { FAIL_ARENA = 1, FAIL_TABLE = 2, FAIL_SCRATCH = 3, FAIL_LINK_CAP = 4, FAIL_TORQUE_CAP = 5 };

constexpr u64 KEY_K_ONLY = 1ull << 14;       // max_hash_dependent_k_only        (KPR/PZsparse.h:37)
constexpr u64 KEY_K_LINKS = 1ull << 35;      // max_hash_dependent_k_links_only  (KPR/PZsparse.h:39)
constexpr u64 KEY_K_MASK = KEY_K_ONLY - 1;   // dependent_k_mask                 (KPR/PZsparse.h:40)
K1_DI u64 key_k(int i) { return 1ull << (2 * i); }
K1_DI u64 key_qde(int i) { return 1ull << (14 + i); }
K1_DI u64 key_qdae(int i) { return 1ull << (21 + i); }
K1_DI u64 key_qddae(int i) { return 1ull << (28 + i); }
K1_DI u64 key_cosqe(int i) { return 1ull << (35 + 2 * i); }
K1_DI u64 key_sinqe(int i) { return 1ull << (49 + 2 * i); }

// ---- PZ handles -----------------------------------------------------------------------------------
// A PZ is referred to by an 8-byte handle (virtual arena offset in words, monomial count) that is passed BY
// VALUE and lives in registers; the element size (1, 3, 9) is a template parameter of the operation.
// Nothing in this file takes the address of a handle or of per-thread state, so the kernel has no stack
// frame: all CTA-uniform state sits in the shared-memory control block K1S below.  (Round-1 profile: the
// previous by-reference plumbing cost 3 KB of local memory per thread and 1.5 G local loads per 64 problems.)
struct PZ8 {
    int off;
    int n;
};
struct PZH {  // resolved view of a block (registers only; built by view<SZ>())
    double* p;
    int n;
    int sz;
};
K1_DI double* pz_c(const PZH& h) { return h.p; }
K1_DI double* pz_r(const PZH& h, int lane) { return h.p + (1 + lane) * h.sz; }
K1_DI u64* pz_keys(const PZH& h) { return reinterpret_cast<u64*>(h.p + 3 * h.sz); }
K1_DI double* pz_coef(const PZH& h) { return h.p + 3 * h.sz + h.n; }
K1_DI int pz_words(int n, int sz) { return 3 * sz + n * (1 + sz); }

struct Tab {
    u64* keys;
    double* acc;
    int cap, shift;  // home slot = (key * golden) >> shift
};

// Per-CTA control block at the start of dynamic shared memory.  Written by thread 0 at kernel start (the
// immutable part) or by all threads with identical values (fail); read by everybody.
struct K1S {
    double* gbase;   // global continuation of the virtual arena: [spill space GW words | F/N scratch FW words]
    char* tab_g;     // global-memory pool for the rare hash table that does not fit (all zero between operations)
    double thr;      // simplify threshold (KPR/PZsparse.cu:321)
    double thr2;     // largest double x with sqrt(x) <= thr: the prune test on the sum of squares, without the root
    int AW;          // words of the shared-memory part of the virtual arena (JRS region + working arena)
    int GW, FW;
    int tab_s_bytes, tab_g_bytes;
    int fail;        // first failure code of the current unit (0 = none)
    int fail_line;   // source line of the check that set it (k1_pz.cuh; 10000 + line for k1_reachsets.cuh)
    int n_tab_global;  // statistics
    int unit;
    int cnt[NW];     // survivor counts per warp
    int jrs_n[40];   // monomial counts of the joint reachable set blocks: R at [0..8], qd/qda/qdda at [16+8g+i]
    PZ8 Fg[MAXJ], Ng[MAXJ];  // F_i / N_i of the forward Newton-Euler pass (blocks in the F/N scratch)
    double red[NW * RED_STRIDE];   // partial sums (prune amounts)
    double red2[NW * RED_STRIDE];  // partial sums (|coefficient| sums)
    double misc[16];               // [0..7) u_nom radius, [8..15) disturbance radius
    int nsurv;                     // survivor counter of op_cross (list path)
    int flip;                      // which survivor-mask buffer the next merge operation uses
    unsigned mask[2][MASK_WORDS];  // survivor bits of a merge operation, by merged position (double-buffered)
};
constexpr int K1S_BYTES = (int(sizeof(K1S)) + 15) & ~15;

// CTA-shared exchange block of the MG mode (lives in the CTA header)
constexpr int MB_SLOTS = 20 * (MAXJ + 1);
constexpr int MG_MAX_TASKS = 16 * (MAXJ + 1);
struct K1X {
    int group_bytes;          // (first two words: the layout of the plain header)
    int unit;
    int seq;                  // sequence number of the unit this CTA is building: the value of a signalled event
    int next;                 // next unclaimed task
    int ntasks;
    int bump;                 // words of the mailbox handed out
    int mbox_words;
    int pad0;
    double* mbox;             // this CTA's mailbox
    double misc[16];          // [0..7) u_nom radius, [8..15) disturbance radius (written by the torque tasks)
    int ev[MB_SLOTS];         // ev[slot] == seq: the block (or plain event) `slot` of this unit is published
    PZ8 mb[MB_SLOTS];         // mailbox offset (words; -1 = failed) and monomial count of a published block
    unsigned short tasks[MG_MAX_TASKS];  // kind << 8 | joint, in claim order
};
constexpr int K1_HDR_BYTES = MG ? ((int(sizeof(K1X)) + 15) & ~15) : 16;

// Dynamic shared memory of a CTA: [CTA header: bytes per group, ...][group 0][group 1]...; every group has
// its own control block, arena and table pool, and its own named barrier.
#ifndef ARMOUR_EMU
K1_DI unsigned char* smem_cta() {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    return k1_smem;
}
K1_DI int k1_tid() { return GROUPS == 1 ? int(threadIdx.x) : int(threadIdx.x) % NT; }
K1_DI int k1_group() { return GROUPS == 1 ? 0 : int(threadIdx.x) / NT; }
K1_DI unsigned char* smem_base() {
    if (GROUPS == 1) return smem_cta() + K1_HDR_BYTES;
    return smem_cta() + K1_HDR_BYTES + size_t(k1_group()) * size_t(*reinterpret_cast<const int*>(smem_cta()));
}
// barrier over the NT threads of one unit
K1_DI void k1_sync() {
    if (GROUPS == 1) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + k1_group()), "r"(NT) : "memory");
    }
}
// barrier over the whole CTA: keeps the groups in step between operations
K1_DI void k1_sync_cta() {
#ifndef K1_NO_LOCKSTEP
    if (GROUPS > 1 && !MG) __syncthreads();
#endif
}
K1_DI K1X& k1x() { return *reinterpret_cast<K1X*>(smem_cta()); }
#else
inline unsigned char* smem_cta() { return reinterpret_cast<unsigned char*>(emu::S().dyn_smem); }
inline int k1_tid() { return int(threadIdx.x); }
inline int k1_group() { return 0; }
inline unsigned char* smem_base() { return smem_cta() + K1_HDR_BYTES; }
inline void k1_sync() { __syncthreads(); }
inline void k1_sync_cta() {}
inline K1X& k1x() { return *reinterpret_cast<K1X*>(smem_cta()); }
#endif
K1_DI K1S& k1s() { return *reinterpret_cast<K1S*>(smem_base()); }
// K1_JRS_GLOBAL: the fixed joint-reachable-set region [0, JRS words) of the virtual arena lives at the front of the group's
// global scratch instead of shared memory (throughput configuration: everything but the control block goes through L1).
// The virtual offsets do not change: gbase points behind the region, so vptr() resolves both halves to gbase - AW + off.
#if K1_JRS_GLOBAL
K1_DI double* arena0() { return k1s().gbase - k1s().AW; }
#else
K1_DI double* arena0() { return reinterpret_cast<double*>(smem_base() + K1S_BYTES); }
#endif
K1_DI char* tab_s0() { return reinterpret_cast<char*>(arena0() + k1s().AW); }

K1_DI void set_fail_at(int code, int line) {
    K1S& S = k1s();
    if (!S.fail) {  // every thread of a site writes the same values
        S.fail = code;
        S.fail_line = line;
    }
}
// the source line of the capacity check that failed first is kept for the error message (armour_last_error)
#define set_fail(code) set_fail_at(code, __LINE__)

// The arena is one virtual offset space over two segments: [0, AW) in shared memory (the fixed JRS region
// first), then the CTA's global scratch: [AW, AW + GW) is spill space for the few long intervals whose live
// set outgrows shared memory (L2-resident), [AW + GW, AW + GW + FW) holds the F_i / N_i blocks.  A block never
// straddles the shared / global boundary.
K1_DI double* vptr(int off) {
    const K1S& S = k1s();
    return off < S.AW ? arena0() + off : S.gbase + (off - S.AW);
}
template <int SZ>
K1_DI PZH view(PZ8 h) {
    PZH v;
    v.p = vptr(h.off);
    v.n = h.n;
    v.sz = SZ;
    return v;
}
template <int SZ>
K1_DI int end_of(PZ8 h) { return h.off + pz_words(h.n, SZ); }  // arena top after an operation that produced h
K1_DI int arena_place(int off, int w) {  // first offset >= off where a block of w words may start
    const int AW = k1s().AW;
    return (off < AW && off + w > AW) ? AW : off;
}
// allocate a block for n monomials at the arena top `top` (uniform); h.n = 0 and a failure code on overflow
template <int SZ>
K1_DI PZ8 pz_alloc(int top, int n, bool* ok) {
    const K1S& S = k1s();
    const int w = pz_words(n, SZ);
    const int at = arena_place(top, w);
    PZ8 h;
    if (S.fail || at + w > S.AW + S.GW) {
        set_fail(FAIL_ARENA);
        h.off = 0;
        h.n = 0;
        *ok = false;
        return h;
    }
#ifdef K1_PROFILE
    if (k1_tid() == 0 && at >= S.AW) atomicAdd(&g_k1spill[0], 1);
    if (k1_tid() == 0) atomicAdd(&g_k1spill[1], 1);
#endif
    h.off = at;
    h.n = n;
    *ok = true;
    return h;
}

// ---- small helpers --------------------------------------------------------------------------------
K1_DI double frob(const double* v, int n) {  // Frobenius norm, entries in storage order (oracle frob_norm)
    double s = 0;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    return sqrt(s);
}
template <int N>
K1_DI double frobN(const double* v) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += v[i] * v[i];
    return sqrt(s);
}
template <int N>
K1_DI double sumsqN(const double* v) {  // the argument of the square root in frobN, same order of operations
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; i++) s += v[i] * v[i];
    return s;
}
// The prune test of simplify() (KPR/PZsparse.cu:321) is sqrt(ss) <= thr.  The IEEE square root is correctly
// rounded and monotone, so the test equals ss <= thr2 with thr2 = max{x : sqrt(x) <= thr} (threshold_sq below,
// computed once per kernel): same decisions bit for bit, one comparison instead of a square root.
K1_DI bool norm_le(double ss, double thr2) { return ss <= thr2; }
K1_DI double threshold_sq(double thr) {
    double x = thr * thr;
    while (sqrt(x) > thr) x = nextafter(x, 0.0);
    for (;;) {
        const double y = nextafter(x, 1e300);
        if (!(sqrt(y) <= thr)) break;
        x = y;
    }
    return x;
}
// C(3 x P) = A(3x3) * B(3 x P), column-major, inner index ascending, no FMA (Eigen-like: KPR/PZsparse.cu:864-994)
template <int P, bool TRANS>
K1_DI void matmul3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int j = 0; j < P; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double acc = (TRANS ? A[i * 3] : A[i]) * B[j * 3];
#pragma unroll
            for (int k = 1; k < 3; k++) acc += (TRANS ? A[k + i * 3] : A[i + k * 3]) * B[k + j * 3];
            C[i + j * 3] = acc;
        }
}
// same with every product and sum rounded up (all inputs non-negative): sound radius propagation
template <int P, bool TRANS>
K1_DI void matmul3_up(const double* A, const double* B, double* C) {
#pragma unroll
    for (int j = 0; j < P; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double acc = __dmul_ru(TRANS ? A[i * 3] : A[i], B[j * 3]);
#pragma unroll
            for (int k = 1; k < 3; k++)
                acc = __dadd_ru(acc, __dmul_ru(TRANS ? A[k + i * 3] : A[i + k * 3], B[k + j * 3]));
            C[i + j * 3] = acc;
        }
}

// (not inlined: the kernel is instruction-fetch bound, and this sequence used to be a quarter of its code)
K1_OP double warp_sum_up(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_ru(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- hash table -----------------------------------------------------------------------------------
K1_DI int tab_home(const Tab& t, u64 key) { return int((key * 0x9E3779B97F4A7C15ull) >> t.shift); }

K1_DI void tab_insert(const Tab& t, u64 key) {
    int h = tab_home(t, key);
    u64 carry = key;
    for (;;) {
        const u64 old = atomicMax(&t.keys[h], carry);
        if (old == 0 || old == carry) return;
        if (old < carry) carry = old;  // the slot now holds our (larger) key; move the displaced one on
        h = (h + 1) & (t.cap - 1);
    }
}
K1_DI int tab_find(const Tab& t, u64 key) {
    int h = tab_home(t, key);
    while (t.keys[h] != key) h = (h + 1) & (t.cap - 1);
    return h;
}

// Pick a table for `nterms` candidate keys with `na` accumulators per slot.  Shared memory when it
// fits (load <= 2/3, or <= 0.8 at half the size), else the CTA's global pool.
K1_DI bool tab_select(int nterms, int na, Tab& t) {
    K1S& S = k1s();
    int cap = 64, lg = 6;
    while (cap * 2 < nterms * 3) {
        cap <<= 1;
        lg++;
    }
    const int slot_bytes = 8 * (1 + na);
    char* base = nullptr;
    if (cap * slot_bytes <= S.tab_s_bytes) {
        base = tab_s0();
    } else if (cap > 64 && nterms * 5 <= (cap / 2) * 4 && (cap / 2) * slot_bytes <= S.tab_s_bytes) {
        cap >>= 1;
        lg--;
        base = tab_s0();
    } else if (cap * slot_bytes <= S.tab_g_bytes) {
        base = S.tab_g;
        if (k1_tid() == 0) S.n_tab_global++;
#ifdef K1_PROFILE
        if (k1_tid() == 0) atomicAdd(&g_k1spill[2], 1);
#endif
    } else {
        set_fail(FAIL_TABLE);
        return false;
    }
    t.keys = reinterpret_cast<u64*>(base);
    t.acc = reinterpret_cast<double*>(base + size_t(cap) * 8);
    t.cap = cap;
    t.shift = 64 - lg;
    return true;
}

// Finalize a table: `fin(acc[NA] -> out[SZ], rad[SZ])` decides keep / prune per occupied slot.
// Pass 1 rewrites kept slots with their output coefficients, clears pruned ones, counts survivors per
// warp segment and reduces the pruned amounts; pass 2 compacts the survivors (slot order) into a new
// arena block at `top` and leaves the table all-zero.  On return every thread holds the handle and
// rad_total[SZ] (block-wide pruned amounts, rounded up); the caller fills centre / radii and must
// __syncthreads() before the block is read.
template <int NA, int SZ, class Fin>
K1_DI PZ8 tab_finalize(int top, const Tab& t, Fin fin, double* rad_total, bool* ok_out) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    const int seg = (t.cap / NW) < 32 ? 32 : (t.cap / NW);
    const int s0 = warp * seg;
    const int s1 = (s0 + seg) < t.cap ? (s0 + seg) : t.cap;
    double rad[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = 0.0;
    int count = 0;
    for (int s = s0 + lane; s < s1; s += 32) {
        const u64 key = t.keys[s];
        bool keep = false;
        if (key != 0) {
            double a[NA], out[SZ];
#pragma unroll
            for (int e = 0; e < NA; e++) a[e] = t.acc[size_t(s) * NA + e];
            keep = fin(a, out, rad);
            if (keep) {
#pragma unroll
                for (int e = 0; e < SZ; e++) t.acc[size_t(s) * NA + e] = out[e];
#pragma unroll
                for (int e = SZ; e < NA; e++) t.acc[size_t(s) * NA + e] = 0.0;
            } else {
                t.keys[s] = 0;
#pragma unroll
                for (int e = 0; e < NA; e++) t.acc[size_t(s) * NA + e] = 0.0;
            }
        }
        count += __popc(__ballot_sync(0xffffffffu, keep));
    }
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = warp_sum_up(rad[e]);
    if (lane == 0) {
        S.cnt[warp] = count;
#pragma unroll
        for (int e = 0; e < SZ; e++) S.red[warp * RED_STRIDE + e] = rad[e];
    }
    k1_sync();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int v = S.cnt[w];
        if (w < warp) before += v;
        total += v;
    }
#pragma unroll
    for (int e = 0; e < SZ; e++) {
        double v = S.red[e];
#pragma unroll
        for (int w = 1; w < NW; w++) v = __dadd_ru(v, S.red[w * RED_STRIDE + e]);
        rad_total[e] = v;
    }
    bool ok;
    const PZ8 h8 = pz_alloc<SZ>(top, total, &ok);
    const PZH h = view<SZ>(h8);
    u64* ok_keys = pz_keys(h);
    double* ok_coef = pz_coef(h);
    int run = before;
    for (int s = s0 + lane; s < s1; s += 32) {
        const u64 key = t.keys[s];
        const bool keep = key != 0;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = run + __popc(b & ((1u << lane) - 1u));
            if (ok) ok_keys[pos] = key;
#pragma unroll
            for (int e = 0; e < SZ; e++) {
                if (ok) ok_coef[size_t(pos) * SZ + e] = t.acc[size_t(s) * NA + e];
                t.acc[size_t(s) * NA + e] = 0.0;
            }
            t.keys[s] = 0;
        }
        run += __popc(b);
    }
    *ok_out = ok;
    return h8;
}

// block-wide sum over the monomials of |coeff| per component (rounded up), NOT including |centre|.
// Writes warp partials to S.red2; the caller syncs, then reads entries with abs_collect_entry.
template <int SZ>
K1_OP void abs_sum_partial_impl(const double* cf, int n) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    double s[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) s[e] = 0.0;
#pragma unroll 1
    for (int m = tid; m < n; m += NT)
#pragma unroll
        for (int e = 0; e < SZ; e++) s[e] = __dadd_ru(s[e], fabs(cf[size_t(m) * SZ + e]));
#pragma unroll
    for (int e = 0; e < SZ; e++) s[e] = warp_sum_up(s[e]);
    if (lane == 0)
#pragma unroll
        for (int e = 0; e < SZ; e++) S.red2[warp * RED_STRIDE + e] = s[e];
}
template <int SZ>
K1_DI void abs_sum_partial(const PZH& h) { abs_sum_partial_impl<SZ>(pz_coef(h), h.n); }
// |centre| + sum |coeff| of ONE component, from the partials of abs_sum_partial
K1_DI double abs_collect_entry(const PZH& h, int comp) {
    const K1S& S = k1s();
    double v = fabs(pz_c(h)[comp]);
#pragma unroll 1
    for (int w = 0; w < NW; w++) v = __dadd_ru(v, S.red2[w * RED_STRIDE + comp]);
    return v;
}
// the same for a short operand: the calling thread walks the monomials itself
K1_DI double abs_serial_entry(const PZH& h, int comp) {
    const double* cf = pz_coef(h);
    double v = fabs(pz_c(h)[comp]);
#pragma unroll 1
    for (int m = 0; m < h.n; m++) v = __dadd_ru(v, fabs(cf[size_t(m) * h.sz + comp]));
    return v;
}

// ---- arena management -----------------------------------------------------------------------------
// Move w words from virtual offset src to dst (dst < src; the ranges may overlap): chunks are read by
// all threads, then written after a barrier.
K1_OP void move_words(int dst_off, int src_off, int w) {
    const int tid = k1_tid();
    double* dst = vptr(dst_off);
    const double* src = vptr(src_off);
    for (int base = 0; base < w; base += NT * 4) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = base + q * NT + tid;
            v[q] = (i < w) ? src[i] : 0.0;
        }
        k1_sync();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = base + q * NT + tid;
            if (i < w) dst[i] = v[q];
        }
    }
    k1_sync();
}
// Slide block h down to the cursor (blocks are kept in ascending address order); advances the cursor.
template <int SZ>
K1_DI void keep(int& cursor, PZ8& h) {
    if (k1s().fail) return;
    const int w = pz_words(h.n, SZ);
    const int at = arena_place(cursor, w);
    if (at != h.off) {
        move_words(at, h.off, w);
        h.off = at;
    }
    cursor = at + w;
}

// copy a block to the CTA's F / N scratch (kept until the backward pass); gtop is the scratch cursor
template <int SZ>
K1_DI PZ8 spill_global(int& gtop, PZ8 h8) {
    const K1S& S = k1s();
    const int w = pz_words(h8.n, SZ);
    PZ8 g = h8;
    if (S.fail || gtop + w > S.FW) {
        set_fail(FAIL_SCRATCH);
        g.n = 0;
        return g;
    }
    g.off = S.AW + S.GW + gtop;
    gtop += w;
    const double* src = vptr(h8.off);
    double* dst = vptr(g.off);
    for (int i = k1_tid(); i < w; i += NT) dst[i] = src[i];
    k1_sync();
    return g;
}

template <int SZ>
K1_DI PZ8 pz_zero(int top) {  // PZ with no monomials, zero centre and radii
    bool ok;
    const PZ8 h = pz_alloc<SZ>(top, 0, &ok);
    if (ok && k1_tid() < 3 * SZ) vptr(h.off)[k1_tid()] = 0.0;
    k1_sync();
    return h;
}

// ---- merge machinery ------------------------------------------------------------------------------
// Every PZ keeps its monomials SORTED by key (like the reference after simplify(), KPR/PZsparse.cu:284-350).
// Sums and products with a short operand are then merges of a few sorted lists: each candidate monomial
// finds its place in the merged order with binary searches (no hashing, no atomics on keys), the first
// candidate of a run of equal keys ("owner") adds up the run in a fixed order, applies the prune rule and
// parks the survivor at its merged position in a dense scratch array; a bit mask of survivors gives the
// output positions by a warp scan.
struct Dense {
    u64* keys;
    double* coef;
    unsigned* mask;   // survivor bits, all zero on entry
    int flip;         // which mask buffer this operation uses (the other one is zeroed for the next operation)
    int M;
    bool global;      // scratch lives in the global pool: give it back zeroed
};
K1_DI bool dense_select(int M, int sz, Dense& d) {
    K1S& S = k1s();
    // More candidates than the survivor masks of the control block hold (32 * MASK_WORDS; thresholds far below the default
    // get there): the mask then lives behind the scratch itself, in the global pool, whose all-zero invariant gives the
    // "all zero on entry" the mask needs; dense_emit clears it again.  d.flip = 2 marks that case.
    const bool big = M > 32 * MASK_WORDS;
    const size_t data_bytes = size_t(M) * 8 * (1 + sz);
    const size_t bytes = data_bytes + (big ? size_t((M + 63) >> 6) * 8 : 0);
    char* base;
    d.global = false;
    if (!big && bytes <= size_t(S.tab_s_bytes)) {
        base = tab_s0();
    } else if (bytes <= size_t(S.tab_g_bytes)) {
        base = S.tab_g;
        d.global = true;
        if (k1_tid() == 0) S.n_tab_global++;
#ifdef K1_PROFILE
        if (k1_tid() == 0) atomicAdd(&g_k1spill[2], 1);
#endif
    } else {
        set_fail(FAIL_TABLE);
        return false;
    }
    d.keys = reinterpret_cast<u64*>(base);
    d.coef = reinterpret_cast<double*>(base + size_t(M) * 8);
    const int f = big ? 2 : S.flip;
    d.mask = big ? reinterpret_cast<unsigned*>(base + data_bytes) : S.mask[f];
    d.flip = f;
    d.M = M;
    return true;
}
// number of keys[0..n) that are < target (keys ascending); branch-free so that several searches interleave
K1_DI int lower_bound(const u64* keys, int n, u64 target) {
    int lo = 0;
    int step = 1;
    while (step < n) step <<= 1;
    for (; step > 0; step >>= 1) {
        const int probe = lo + step;
        if (probe <= n && keys[probe - 1] < target) lo = probe;
    }
    return lo;
}
// reduce the per-thread pruned amounts to S.red (skipped when nothing was pruned in the warp)
template <int SZ>
K1_DI void rad_publish(const double* rad, bool pruned_any) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    double r[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) r[e] = rad[e];
    if (__any_sync(0xffffffffu, pruned_any)) {
#pragma unroll
        for (int e = 0; e < SZ; e++) r[e] = warp_sum_up(r[e]);
    }
    if (lane == 0)
#pragma unroll
        for (int e = 0; e < SZ; e++) S.red[warp * RED_STRIDE + e] = r[e];
}
template <int SZ>
K1_DI void rad_collect(double* rad_total) {
    const K1S& S = k1s();
#pragma unroll
    for (int e = 0; e < SZ; e++) {
        double v = S.red[e];
#pragma unroll
        for (int w = 1; w < NW; w++) v = __dadd_ru(v, S.red[w * RED_STRIDE + e]);
        rad_total[e] = v;
    }
}
// Second half of a merge operation (after the barrier that follows the scatter): count the survivors,
// allocate the output block at `top` and copy them in merged (= key) order.  The caller writes the
// centre / radii and ends with __syncthreads().  Kept out of line and with run-time loops: one copy of this
// code per element size serves every merge operation (code size is what bounds this kernel).
template <int SZ, bool BIG>
K1_OP PZ8 dense_emit_impl(int top, u64* keys, int M, int flip, int global) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    double* coef = reinterpret_cast<double*>(keys + M);
    unsigned* big_mask = reinterpret_cast<unsigned*>(coef + size_t(M) * SZ);  // (BIG: see dense_select)
    const unsigned* mask = BIG ? big_mask : S.mask[flip];
    const int nwords = (M + 31) >> 5;
    const int rounds = (nwords + 31) >> 5;
    int total = 0;
#pragma unroll 1
    for (int j = 0; j < rounds; j++) {
        const int wi = j * 32 + lane;
        total += __popc(wi < nwords ? mask[wi] : 0u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    // prepare the mask buffer of the next merge operation; its last readers finished before the barrier above
    if (!BIG) {
        for (int i = tid; i < MASK_WORDS; i += NT) S.mask[flip ^ 1][i] = 0u;
        if (tid == 0) S.flip = flip ^ 1;
    }
    bool ok;
    const PZ8 h8 = pz_alloc<SZ>(top, total, &ok);
    const PZH h = view<SZ>(h8);
    u64* out_keys = pz_keys(h);
    double* out_coef = pz_coef(h);
    int running = 0;
#pragma unroll 1
    for (int j = 0; j < rounds; j++) {
        const int wi = j * 32 + lane;
        const unsigned w = wi < nwords ? mask[wi] : 0u;
        const int v = __popc(w);
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int excl = incl - v + running;
        running += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll 1
        for (int cc = warp; cc < 32; cc += NW) {
            const int c = j * 32 + cc;
            if (c >= nwords) break;  // uniform per warp
            const int base = __shfl_sync(0xffffffffu, excl, cc);
            const unsigned word = __shfl_sync(0xffffffffu, w, cc);
            const int p = c * 32 + lane;
            if (ok && ((word >> lane) & 1u)) {
                const int rank = base + __popc(word & ((1u << lane) - 1u));
                out_keys[rank] = keys[p];
#pragma unroll
                for (int e = 0; e < SZ; e++) out_coef[size_t(rank) * SZ + e] = coef[size_t(p) * SZ + e];
            }
            if (global && p < M) {  // the global pool is handed back all-zero
                keys[p] = 0;
#pragma unroll
                for (int e = 0; e < SZ; e++) coef[size_t(p) * SZ + e] = 0.0;
            }
        }
    }
    if (BIG) {  // the mask behind the scratch goes back all-zero too (every warp is done reading it)
        k1_sync();
        for (int i = tid; i < nwords; i += NT) big_mask[i] = 0u;
    }
    return h8;
}
template <int SZ>
K1_DI PZ8 dense_emit(int top, const Dense& d, bool* ok_out) {
    const PZ8 h8 = d.flip == 2 ? dense_emit_impl<SZ, true>(top, d.keys, d.M, d.flip, 1)
                               : dense_emit_impl<SZ, false>(top, d.keys, d.M, d.flip, d.global ? 1 : 0);
    *ok_out = !k1s().fail;
    return h8;
}
// park a surviving candidate at merged position pos
template <int SZ>
K1_DI void dense_put(const Dense& d, int pos, u64 key, const double* v) {
    d.keys[pos] = key;
#pragma unroll
    for (int e = 0; e < SZ; e++) d.coef[size_t(pos) * SZ + e] = v[e];
    atomicOr(&d.mask[pos >> 5], 1u << (pos & 31));
}

// Sort the monomials of a block by key, in place (used after the hash-table product, whose survivors come out
// in slot order): every thread ranks its monomials against all keys, then stores them at their ranks.
template <int SZ>
K1_DI void sort_block(PZ8 h8, bool ok) {
    constexpr int R = 4;  // up to R * NT monomials
    const int tid = k1_tid();
    const PZH h = view<SZ>(h8);
    const int n = ok ? h.n : 0;
    u64* keys = pz_keys(h);
    double* cf = pz_coef(h);
    if (n > R * NT) {
        // more survivors than the register path holds (thresholds far below the default): rank every monomial, scatter into
        // the global table pool, copy back, hand the pool back all-zero
        K1S& S = k1s();
        if (size_t(n) * 8 * (1 + SZ) > size_t(S.tab_g_bytes)) {
            set_fail(FAIL_TABLE);
            return;
        }
        u64* sk = reinterpret_cast<u64*>(S.tab_g);
        double* sv = reinterpret_cast<double*>(S.tab_g + size_t(n) * 8);
        for (int i = tid; i < n; i += NT) {
            const u64 ki = keys[i];
            int c = 0;
            for (int q = 0; q < n; q++) c += (keys[q] < ki);
            sk[c] = ki;
            for (int e = 0; e < SZ; e++) sv[size_t(c) * SZ + e] = cf[size_t(i) * SZ + e];
        }
        k1_sync();
        for (int i = tid; i < n; i += NT) {
            keys[i] = sk[i];
            sk[i] = 0;
            for (int e = 0; e < SZ; e++) {
                cf[size_t(i) * SZ + e] = sv[size_t(i) * SZ + e];
                sv[size_t(i) * SZ + e] = 0.0;
            }
        }
        k1_sync();
        return;
    }
    u64 k[R];
    double v[R][SZ];
    int rank[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int i = tid + r * NT;
        rank[r] = -1;
        if (i < n) {
            k[r] = keys[i];
#pragma unroll
            for (int e = 0; e < SZ; e++) v[r][e] = cf[size_t(i) * SZ + e];
            int c = 0;
#pragma unroll 2
            for (int q = 0; q < n; q++) c += (keys[q] < k[r]);
            rank[r] = c;
        }
    }
    k1_sync();
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (rank[r] >= 0) {
            keys[rank[r]] = k[r];
#pragma unroll
            for (int e = 0; e < SZ; e++) cf[size_t(rank[r]) * SZ + e] = v[r][e];
        }
    }
}

// ---- linear operations (operator+, operator-, addOneDimPZ, element extraction, double*PZ) -------
// out = src0 (+) src1 with, per source: optional scalar extraction (comp_in >= 0 -> placed at comp_out of
// the output) and a scale factor.  Mirrors KPR/PZsparse.cu:743-834 (+,-), :996-1030 (double*PZ, no
// simplify), :678-697 (operator()), :1068-1085 (addOneDimPZ): one simplify() at the end.
struct LinSrc {
    PZH h;
    int comp_in;   // -1: whole coefficient (sz_in == SZ); else scalar element of the source
    int comp_out;  // destination component when comp_in >= 0 or the source is a scalar placed into a vector
    double scale;  // coefficient = scale * c  (1.0 for plain add, -1.0 for subtraction)
};

template <int SZ>
K1_DI void lin_value(const LinSrc& s, const double* v, double* out) {  // v: source coefficient / centre
#pragma unroll
    for (int e = 0; e < SZ; e++) out[e] = 0.0;
    if (s.comp_in < 0 && s.h.sz == SZ) {
#pragma unroll
        for (int e = 0; e < SZ; e++) out[e] = s.scale * v[e];
    } else {
        const double x = s.scale * v[s.comp_in < 0 ? 0 : s.comp_in];
#pragma unroll
        for (int e = 0; e < SZ; e++)
            if (e == s.comp_out) out[e] = x;
    }
}
template <int SZ>
K1_DI void lin_radius(const LinSrc& s, int lane, double* out) {  // radius * |scale| (rounded up)
    const double* rv = pz_r(s.h, lane);
    const double sc = fabs(s.scale);
#pragma unroll
    for (int q = 0; q < SZ; q++) out[q] = 0.0;
    if (s.comp_in < 0 && s.h.sz == SZ) {
#pragma unroll
        for (int q = 0; q < SZ; q++) out[q] = __dmul_ru(sc, rv[q]);
    } else {
        const double x = __dmul_ru(sc, rv[s.comp_in < 0 ? 0 : s.comp_in]);
#pragma unroll
        for (int q = 0; q < SZ; q++)
            if (q == s.comp_out) out[q] = x;
    }
}

// sources are described by scalars passed by value: (handle, element size, comp_in, comp_out, scale)
// A merge of two sorted lists: candidate positions = own index + rank in the other list (source 0 first on a
// tie); the owner of a key adds s0's value and s1's value (at most two terms: order-free).
template <int SZ>
K1_OP PZ8 op_lin2(int top, PZ8 a8, int a_sz, int a_in, int a_out, double a_scale, PZ8 b8, int b_sz, int b_in, int b_out,
                  double b_scale) {
    const PZ8 dummy = {0, 0};
    K1S& S = k1s();
    if (S.fail) return dummy;
    const int tid = k1_tid();
    LinSrc s0, s1;
    s0.h.p = vptr(a8.off); s0.h.n = a8.n; s0.h.sz = a_sz; s0.comp_in = a_in; s0.comp_out = a_out; s0.scale = a_scale;
    s1.h.p = vptr(b8.off); s1.h.n = b8.n; s1.h.sz = b_sz; s1.comp_in = b_in; s1.comp_out = b_out; s1.scale = b_scale;
    const int na = s0.h.n, nb = s1.h.n;
    Dense d;
    if (!dense_select(na + nb, SZ, d)) return dummy;
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    double rad[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = 0.0;
    bool pruned_any = false;
    for (int it = tid; it < na + nb; it += NT) {
        const bool fromA = it < na;
        const int idx = fromA ? it : it - na;
        const LinSrc& src = fromA ? s0 : s1;
        const LinSrc& oth = fromA ? s1 : s0;
        const u64* ks = pz_keys(src.h);
        const u64* ko = pz_keys(oth.h);
        const double* cs = pz_coef(src.h) + size_t(idx) * src.h.sz;
        const u64 key = ks[idx];
        const int lb = lower_bound(ko, oth.h.n, key);
        const bool hit = lb < oth.h.n && ko[lb] == key;
        const double* co = pz_coef(oth.h) + size_t(hit ? lb : 0) * oth.h.sz;
        // an extracted element that is exactly zero takes no part (KPR/PZsparse.cu:678-697 keeps it, simplify drops it)
        const bool absent = src.comp_in >= 0 && cs[src.comp_in] == 0.0;
        const bool hit_valid = hit && !(oth.comp_in >= 0 && co[oth.comp_in] == 0.0);
        const int pos = idx + lb + ((!fromA && hit) ? 1 : 0);
        const bool owner = !absent && (fromA || !hit_valid);
        if (owner) {
            double v[SZ], v2[SZ];
            lin_value<SZ>(src, cs, v);
            if (fromA && hit_valid) {
                lin_value<SZ>(oth, co, v2);
#pragma unroll
                for (int e = 0; e < SZ; e++) v[e] = v[e] + v2[e];
            }
            if (norm_le(sumsqN<SZ>(v), thr)) {
#pragma unroll
                for (int e = 0; e < SZ; e++) rad[e] = __dadd_ru(rad[e], fabs(v[e]));
                pruned_any = true;
            } else {
                dense_put<SZ>(d, pos, key, v);
            }
        }
    }
    rad_publish<SZ>(rad, pruned_any);
    k1_sync();
    bool ok;
    const PZ8 h8 = dense_emit<SZ>(top, d, &ok);
    if (ok && tid < SZ) {
        double radt[SZ];
        rad_collect<SZ>(radt);
        const PZH h = view<SZ>(h8);
        const int e = tid;
        double c0[SZ], c1[SZ];
        lin_value<SZ>(s0, pz_c(s0.h), c0);
        lin_value<SZ>(s1, pz_c(s1.h), c1);
        pz_c(h)[e] = c0[e] + c1[e];
#pragma unroll
        for (int lane = 0; lane < 2; lane++) {
            double r0[SZ], r1[SZ];
            lin_radius<SZ>(s0, lane, r0);
            lin_radius<SZ>(s1, lane, r1);
            pz_r(h, lane)[e] = __dadd_ru(__dadd_ru(r0[e], r1[e]), radt[e]);
        }
    }
    k1_sync();
    return h8;
}

template <int SZ>
K1_DI PZ8 op_add(int top, PZ8 a, PZ8 b) {  // KPR/PZsparse.cu:743-764
    return op_lin2<SZ>(top, a, SZ, -1, 0, 1.0, b, SZ, -1, 0, 1.0);
}
// a.addOneDimPZ(s, comp, 0) for a 3-vector a and a scalar s (KPR/PZsparse.cu:1068-1085)
K1_DI PZ8 op_add_one_dim(int top, PZ8 a, PZ8 s, int comp) { return op_lin2<3>(top, a, 3, -1, 0, 1.0, s, 1, 0, comp, 1.0); }

// ---- element-wise ("map") operations: the key set does not change, so no table is needed ---------
// Generic driver: fn(m, out[SZ], rad[SZ]) -> keep.  Two passes (count, then recompute + write) keep the
// monomials in their input order.
template <int SZ, class Fn>
K1_DI PZ8 map_op(int top, int n_in, const u64* keys_in, Fn fn, double* rad_total, bool* ok_out) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    const int chunk = ((n_in + NW - 1) / NW + 31) & ~31;  // per-warp contiguous chunk, multiple of 32
    const int m0 = warp * chunk;
    const int m1 = (m0 + chunk) < n_in ? (m0 + chunk) : n_in;
    double rad[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = 0.0;
    int count = 0;
    for (int mb = m0; mb < m1; mb += 32) {
        const int m = mb + lane;
        bool keep = false;
        if (m < m1) {
            double out[SZ];
            keep = fn(m, out, rad);
        }
        count += __popc(__ballot_sync(0xffffffffu, keep));
    }
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = warp_sum_up(rad[e]);
    if (lane == 0) {
        S.cnt[warp] = count;
#pragma unroll
        for (int e = 0; e < SZ; e++) S.red[warp * RED_STRIDE + e] = rad[e];
    }
    k1_sync();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const int v = S.cnt[w];
        if (w < warp) before += v;
        total += v;
    }
#pragma unroll
    for (int e = 0; e < SZ; e++) {
        double v = S.red[e];
#pragma unroll
        for (int w = 1; w < NW; w++) v = __dadd_ru(v, S.red[w * RED_STRIDE + e]);
        rad_total[e] = v;
    }
    bool ok;
    const PZ8 h8 = pz_alloc<SZ>(top, total, &ok);
    *ok_out = ok;
    if (!ok) return h8;
    const PZH h = view<SZ>(h8);
    u64* okk = pz_keys(h);
    double* oc = pz_coef(h);
    int run = before;
    for (int mb = m0; mb < m1; mb += 32) {
        const int m = mb + lane;
        bool keep = false;
        double out[SZ], dump[SZ];
#pragma unroll
        for (int e = 0; e < SZ; e++) dump[e] = 0.0;
        if (m < m1) keep = fn(m, out, dump);
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = run + __popc(b & ((1u << lane) - 1u));
            okk[pos] = keys_in[m];
#pragma unroll
            for (int e = 0; e < SZ; e++) oc[size_t(pos) * SZ + e] = out[e];
        }
        run += __popc(b);
    }
    return h8;  // caller writes the header, then __syncthreads()
}

// prune rule for one merged coefficient (KPR/PZsparse.cu:321-336)
template <int SZ>
K1_DI bool prune_or_keep(const double* v, double thr, double* rad) {
    if (norm_le(sumsqN<SZ>(v), thr)) {
#pragma unroll
        for (int e = 0; e < SZ; e++) rad[e] = __dadd_ru(rad[e], fabs(v[e]));
        return false;
    }
    return true;
}

// cross(PZ a, const b) and cross(const a, PZ b): three scaled scalar subtractions (each simplified on its
// own: a component with |c| <= thr moves to that component's radius) and a stack (KPR/PZsparse.cu:1118-1132,
// 1153-1167, 1087-1116).  left_const == false: result = x (PZ) cross v (const)  [cross_pz_mat]
//                         left_const == true : result = v (const) cross x (PZ)  [cross_mat_pz]
K1_DI void cross_const_coef(bool left_const, const double* v, const double* g, double thr, double* out, double* rad,
                            bool* any) {
    double c[3];
    if (!left_const) {  // r_e = v[(e+2)%3] * g[(e+1)%3] - v[(e+1)%3] * g[(e+2)%3]
        c[0] = v[2] * g[1] - v[1] * g[2];
        c[1] = v[0] * g[2] - v[2] * g[0];
        c[2] = v[1] * g[0] - v[0] * g[1];
    } else {  // r_e = v[(e+1)%3] * g[(e+2)%3] - v[(e+2)%3] * g[(e+1)%3]
        c[0] = v[1] * g[2] - v[2] * g[1];
        c[1] = v[2] * g[0] - v[0] * g[2];
        c[2] = v[0] * g[1] - v[1] * g[0];
    }
    bool a = false;
#pragma unroll
    for (int e = 0; e < 3; e++) {
        if (norm_le(c[e] * c[e], thr)) {
            rad[e] = __dadd_ru(rad[e], fabs(c[e]));
            out[e] = 0.0;
        } else {
            out[e] = c[e];
            a = true;
        }
    }
    *any = a;
}
K1_OP PZ8 op_cross_const(int top, PZ8 x8, const double* v, bool left_const) {
    const PZ8 dummy = {0, 0};
    K1S& S = k1s();
    if (S.fail) return dummy;
    const int tid = k1_tid();
    const PZH x = view<3>(x8);
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    const double* cf = pz_coef(x);
    const double v0 = v[0], v1 = v[1], v2 = v[2];
    double rad[3];
    bool ok;
    const PZ8 h8 = map_op<3>(
        top, x.n, pz_keys(x),
        [=](int m, double* out, double* r) {
            const double vv[3] = {v0, v1, v2};
            bool any;
            cross_const_coef(left_const, vv, cf + size_t(m) * 3, thr, out, r, &any);
            if (!any) return false;
            return prune_or_keep<3>(out, thr, r);  // stack3's own simplify (3-vector norm)
        },
        rad, &ok);
    if (ok && tid < 3) {
        const PZH h = view<3>(h8);
        const int e = tid, e1 = (e + 1) % 3, e2 = (e + 2) % 3;
        const double* xc = pz_c(x);
        const double vv[3] = {v0, v1, v2};
        // centre: scale() multiplies centre * s (KPR/PZsparse.cu:1004), then the subtraction
        pz_c(h)[e] = !left_const ? (xc[e1] * vv[e2] - xc[e2] * vv[e1]) : (xc[e2] * vv[e1] - xc[e1] * vv[e2]);
#pragma unroll
        for (int lane = 0; lane < 2; lane++) {
            const double* xr = pz_r(x, lane);
            const double a = !left_const ? __dmul_ru(xr[e1], fabs(vv[e2])) : __dmul_ru(xr[e2], fabs(vv[e1]));
            const double b = !left_const ? __dmul_ru(xr[e2], fabs(vv[e1])) : __dmul_ru(xr[e1], fabs(vv[e2]));
            pz_r(h, lane)[e] = __dadd_ru(__dadd_ru(a, b), rad[e]);
        }
    }
    k1_sync();
    return h8;
}

// PZ(constant, radius = pct*|constant|) * x  for a scalar constant (mass) or a 3x3 constant (inertia), and
// x * PZ(constant 3-vector) for a 3x3 x.  The constant operand has no monomials, so the product keeps the
// key set of x (KPR/PZsparse.cu:864-994 with an empty polynomial on one side); lane 1 carries the
// uncertain-parameter radius (KPR/PZsparse.cu:93-98, KPR/Dynamics.cu:30-40).
//   kind 0: scalar s * vec3 x        kind 1: mat3 M * vec3 x        kind 2: mat3 x * const vec3 P
template <int KIND>
K1_OP PZ8 op_const_mul(int top, const double* K, double pct_lane1, PZ8 x8) {
    const PZ8 dummy = {0, 0};
    K1S& S = k1s();
    if (S.fail) return dummy;
    const int tid = k1_tid();
    constexpr int XS = (KIND == 2) ? 9 : 3;
    const PZH x = view<XS>(x8);
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    const double* cf = pz_coef(x);
    double kk[9];
#pragma unroll
    for (int i = 0; i < 9; i++) kk[i] = (KIND == 0) ? ((i == 0) ? K[0] : 0.0) : ((KIND == 1 || i < 3) ? K[i] : 0.0);
    abs_sum_partial<XS>(x);  // collected after map_op's internal barrier
    double rad[3];
    bool ok;
    const PZ8 h8 = map_op<3>(
        top, x.n, pz_keys(x),
        [=](int m, double* out, double* r) {
            const double* g = cf + size_t(m) * XS;
            if (KIND == 0) {
#pragma unroll
                for (int e = 0; e < 3; e++) out[e] = kk[0] * g[e];
            } else if (KIND == 1) {
                matmul3<1, false>(kk, g, out);
            } else {
                matmul3<1, false>(g, kk, out);
            }
            return prune_or_keep<3>(out, thr, r);
        },
        rad, &ok);
    if (ok && tid < 3) {
        const PZH h = view<3>(h8);
        const int e = tid;
        const double* xc = pz_c(x);
        // row e of the constant operand (scalar: the scalar itself) and |.| of it
        if (KIND == 0) {
            pz_c(h)[e] = kk[0] * xc[e];
        } else if (KIND == 1) {
            double acc = K[e] * xc[0];
#pragma unroll 1
            for (int k = 1; k < 3; k++) acc += K[e + 3 * k] * xc[k];
            pz_c(h)[e] = acc;
        } else {
            double acc = xc[e] * K[0];
#pragma unroll 1
            for (int k = 1; k < 3; k++) acc += xc[e + 3 * k] * K[k];
            pz_c(h)[e] = acc;
        }
#pragma unroll 1
        for (int lane = 0; lane < 2; lane++) {
            const double* xr = pz_r(x, lane);
            const double pct = lane ? pct_lane1 : 0.0;
            double ra2 = 0.0, ra3 = 0.0, rr = 0.0;
            if (KIND == 0) {
                const double aK = fabs(K[0]), rK = __dmul_ru(pct, aK);
                ra2 = __dmul_ru(aK, xr[e]);
                ra3 = __dmul_ru(rK, abs_collect_entry(x, e));
                rr = __dmul_ru(rK, xr[e]);
            } else if (KIND == 1) {
#pragma unroll 1
                for (int k = 0; k < 3; k++) {
                    const double aK = fabs(K[e + 3 * k]), rK = __dmul_ru(pct, aK);
                    const double p2 = __dmul_ru(aK, xr[k]), p3 = __dmul_ru(rK, abs_collect_entry(x, k)), pr = __dmul_ru(rK, xr[k]);
                    ra2 = (k == 0) ? p2 : __dadd_ru(ra2, p2);
                    ra3 = (k == 0) ? p3 : __dadd_ru(ra3, p3);
                    rr = (k == 0) ? pr : __dadd_ru(rr, pr);
                }
            } else {  // x (3x3 PZ) * constant vector: the constant has no radius
#pragma unroll 1
                for (int k = 0; k < 3; k++) {
                    const double p3 = __dmul_ru(xr[e + 3 * k], fabs(K[k]));
                    ra3 = (k == 0) ? p3 : __dadd_ru(ra3, p3);
                }
            }
            pz_r(h, lane)[e] = __dadd_ru(__dadd_ru(rr, __dadd_ru(ra2, ra3)), rad[e]);
        }
    }
    k1_sync();
    return h8;
}

// ---- PZ * PZ with a 3x3 left operand (KPR/PZsparse.cu:864-994) -----------------------------------
// P = 1: 3x3 * 3x1, P = 3: 3x3 * 3x3.  TRANS: the left operand is used transposed (R_t = R.transpose(),
// KPR/Trajectory.cu:143).  On this path one operand is always short (a joint rotation with <= 3 monomials, or
// the link box): the "outer" operand O.  The product is the merge of nO + 2 sorted lists
//     list 0      : inner monomial j times the centre of the outer operand      keys kI[j]
//     list 1 + o  : inner monomial j times outer monomial o                      keys kI[j] + kO[o]
//     list nO + 1 : centre of the inner operand times outer monomial o           keys kO[o]
// (degree hashes add without carry, KPR/PZsparse.cu:938-940; adding a constant keeps a list sorted).  The
// owner of a key adds the contributions in the fixed order: left-polynomial x right-centre, left-centre x
// right-polynomial, then the pair products for o = 0, 1, 2.
constexpr int MUL_MAX_OUTER = 3;
template <int P, bool TRANS>
K1_OP PZ8 op_mul33(int top, PZ8 L8, PZ8 R8) {
    constexpr int SZ = 3 * P;
    const PZ8 dummy = {0, 0};
    K1S& S = k1s();
    if (S.fail) return dummy;
    const int tid = k1_tid();
    const PZH L = view<9>(L8), R = view<SZ>(R8);
    const int nL = L.n, nR = R.n;
    const bool outerL = nL <= nR;
    const int nO = outerL ? nL : nR, nI = outerL ? nR : nL;
    if (nO > MUL_MAX_OUTER) {
        set_fail(FAIL_TABLE);
        return dummy;
    }
    const int M = nI * (nO + 1) + nO;
    Dense d;
    if (!dense_select(M, SZ, d)) return dummy;
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    const u64* kL = pz_keys(L);
    const u64* kR = pz_keys(R);
    const double* cL = pz_coef(L);
    const double* cR = pz_coef(R);
    const u64* kO = outerL ? kL : kR;
    const u64* kI = outerL ? kR : kL;
    u64 ko[MUL_MAX_OUTER];
#pragma unroll
    for (int o = 0; o < MUL_MAX_OUTER; o++) ko[o] = (o < nO) ? kO[o] : ~0ull;
    double cenL[9], cenR[SZ];
#pragma unroll
    for (int q = 0; q < 9; q++) cenL[q] = pz_c(L)[q];
#pragma unroll
    for (int q = 0; q < SZ; q++) cenR[q] = pz_c(R)[q];
    double rad[SZ];
#pragma unroll
    for (int e = 0; e < SZ; e++) rad[e] = 0.0;
    bool pruned_any = false;
    for (int it = tid; it < M; it += NT) {
        // which list, which element
        int a, j;
        if (it < nI * (nO + 1)) {
            a = it / nI;
            j = it - a * nI;
        } else {
            a = nO + 1;
            j = it - nI * (nO + 1);  // outer monomial index
        }
        // (select chains instead of indexing ko[] with a run-time value keep it in registers)
        const u64 ko_a = (a == 1) ? ko[0] : ((a == 2) ? ko[1] : ko[2]);
        const u64 ko_j = (j == 0) ? ko[0] : ((j == 1) ? ko[1] : ko[2]);
        const u64 key = (a == 0) ? kI[j] : ((a <= nO) ? kI[j] + ko_a : ko_j);
        // rank and hit of `key` in every list: list b holds kI[.] + (b ? ko[b-1] : 0), so the rank of `key` there
        // is the rank of key - ko[b-1] in kI.  The four binary searches run over the same array with the same step
        // sequence: one loop, four independent probes per step (the latency of one search instead of four); the
        // search in the candidate's own list is known without looking (rank j, hit).
        int lb[MUL_MAX_OUTER + 1];
        bool hit[MUL_MAX_OUTER + 1];
        {
            u64 tg[MUL_MAX_OUTER + 1];
            bool on[MUL_MAX_OUTER + 1];
            tg[0] = key;
            on[0] = (a != 0);
#pragma unroll
            for (int o = 0; o < MUL_MAX_OUTER; o++) {
                on[o + 1] = o < nO && key >= ko[o] && a != o + 1;
                tg[o + 1] = key - ((o < nO) ? ko[o] : 0ull);
            }
#pragma unroll
            for (int b = 0; b <= MUL_MAX_OUTER; b++) lb[b] = 0;
            int step = 1;
            while (step < nI) step <<= 1;
            for (; step > 0; step >>= 1) {
#pragma unroll
                for (int b = 0; b <= MUL_MAX_OUTER; b++) {
                    const int probe = lb[b] + step;
                    if (on[b] && probe <= nI && kI[probe - 1] < tg[b]) lb[b] = probe;
                }
            }
#pragma unroll
            for (int b = 0; b <= MUL_MAX_OUTER; b++) {
                hit[b] = on[b] && lb[b] < nI && kI[lb[b]] == tg[b];
                if (!on[b]) lb[b] = 0;
            }
            if (a <= nO) {  // own list: element j itself
#pragma unroll
                for (int b = 0; b <= MUL_MAX_OUTER; b++)
                    if (b == a) {
                        lb[b] = j;
                        hit[b] = true;
                    }
            }
        }
        int lbH = 0, hitH = -1;
#pragma unroll
        for (int o = 0; o < MUL_MAX_OUTER; o++) {
            if (o < nO) {
                lbH += (ko[o] < key);
                if (ko[o] == key) hitH = o;
            }
        }
        // merged position: ties are ordered by list index
        int pos = lbH, first = (hitH >= 0) ? nO + 1 : nO + 2;
#pragma unroll
        for (int b = MUL_MAX_OUTER; b >= 0; b--) {
            if (b <= nO) {
                pos += lb[b];
                if (hit[b]) first = b;
            }
        }
#pragma unroll
        for (int b = 0; b <= MUL_MAX_OUTER; b++)
            if (b <= nO && b < a && hit[b]) pos += 1;
        if (first != a) continue;   // not the owner of this key
        // sum of the run, fixed order
        double acc[SZ];
#pragma unroll
        for (int e = 0; e < SZ; e++) acc[e] = 0.0;
        double v[SZ];
        const bool inL = outerL ? (hitH >= 0) : hit[0];  // key among the monomials of L
        const bool inR = outerL ? hit[0] : (hitH >= 0);  // key among the monomials of R
        if (inL) {
            const int i = outerL ? hitH : lb[0];
            matmul3<P, TRANS>(cL + size_t(i) * 9, cenR, v);
#pragma unroll
            for (int e = 0; e < SZ; e++) acc[e] += v[e];
        }
        if (inR) {
            const int jj = outerL ? lb[0] : hitH;
            matmul3<P, TRANS>(cenL, cR + size_t(jj) * SZ, v);
#pragma unroll
            for (int e = 0; e < SZ; e++) acc[e] += v[e];
        }
#pragma unroll
        for (int o = 0; o < MUL_MAX_OUTER; o++) {
            if (o < nO && hit[o + 1]) {
                const int ji = lb[o + 1];
                if (outerL)
                    matmul3<P, TRANS>(cL + size_t(o) * 9, cR + size_t(ji) * SZ, v);
                else
                    matmul3<P, TRANS>(cL + size_t(ji) * 9, cR + size_t(o) * SZ, v);
#pragma unroll
                for (int e = 0; e < SZ; e++) acc[e] += v[e];
            }
        }
        if (norm_le(sumsqN<SZ>(acc), thr)) {
#pragma unroll
            for (int e = 0; e < SZ; e++) rad[e] = __dadd_ru(rad[e], fabs(acc[e]));
            pruned_any = true;
        } else {
            dense_put<SZ>(d, pos, key, acc);
        }
    }
    rad_publish<SZ>(rad, pruned_any);
    if (outerL) abs_sum_partial<SZ>(R); else abs_sum_partial<9>(L);
    k1_sync();
    bool ok;
    const PZ8 h8 = dense_emit<SZ>(top, d, &ok);
    if (ok && tid < SZ) {
        double radt[SZ];
        rad_collect<SZ>(radt);
        const PZH h = view<SZ>(h8);
        const int e = tid, i = e % 3, jc = e / 3;
        // entry (i, jc) of the product: sum over k of Lm(i, k) * Rm(k, jc), k ascending
        const double* pcL = pz_c(L);
        const double* pcR = pz_c(R);
        double acc = (TRANS ? pcL[i * 3] : pcL[i]) * pcR[jc * 3];
#pragma unroll 1
        for (int k = 1; k < 3; k++) acc += (TRANS ? pcL[k + i * 3] : pcL[i + k * 3]) * pcR[k + jc * 3];
        pz_c(h)[e] = acc;
        double aL[3], aR[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int li = TRANS ? (k + i * 3) : (i + k * 3), ri = k + jc * 3;
            aL[k] = outerL ? abs_serial_entry(L, li) : abs_collect_entry(L, li);
            aR[k] = outerL ? abs_collect_entry(R, ri) : abs_serial_entry(R, ri);
        }
#pragma unroll 1
        for (int lane = 0; lane < 2; lane++) {
            const double* rL = pz_r(L, lane);
            const double* rR = pz_r(R, lane);
            double ra2 = 0.0, ra3 = 0.0, rr = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int li = TRANS ? (k + i * 3) : (i + k * 3), ri = k + jc * 3;
                const double p2 = __dmul_ru(aL[k], rR[ri]), p3 = __dmul_ru(rL[li], aR[k]), pr = __dmul_ru(rL[li], rR[ri]);
                ra2 = (k == 0) ? p2 : __dadd_ru(ra2, p2);
                ra3 = (k == 0) ? p3 : __dadd_ru(ra3, p3);
                rr = (k == 0) ? pr : __dadd_ru(rr, pr);
            }
            pz_r(h, lane)[e] = __dadd_ru(__dadd_ru(rr, __dadd_ru(ra2, ra3)), radt[e]);
        }
    }
    k1_sync();
    return h8;
}

// ---- cross(PZ a, PZ b) (KPR/PZsparse.cu:1134-1151) ------------------------------------------------
// Six scalar PZ products a_x * b_y (each with its own simplify), three scalar subtractions (simplify)
// and a stack (simplify), fused: the six products share ONE key set {a_i + b_j}, so one table with six
// accumulators per slot reproduces all ten simplify() calls slot by slot.
//   P0 = a1*b2, P1 = a2*b1, P2 = a2*b0, P3 = a0*b2, P4 = a0*b1, P5 = a1*b0;  r_e = P_{2e} - P_{2e+1}
K1_DI void cross_six(const double* a, const double* b, double* v) {
    v[0] = a[1] * b[2];
    v[1] = a[2] * b[1];
    v[2] = a[2] * b[0];
    v[3] = a[0] * b[2];
    v[4] = a[0] * b[1];
    v[5] = a[1] * b[0];
}
// simplify() chain of one key of cross(PZ, PZ): the six scalar products, the three differences, the stack
K1_DI bool cross_fin(double thr, const double* a, double* out, double* r) {
    double p[6];
    bool have[6];
#pragma unroll
    for (int x = 0; x < 6; x++) {  // simplify() of each scalar product
        have[x] = !norm_le(a[x] * a[x], thr);
        p[x] = have[x] ? a[x] : 0.0;
        if (!have[x]) r[x >> 1] = __dadd_ru(r[x >> 1], fabs(a[x]));
    }
    bool any = false;
#pragma unroll
    for (int e = 0; e < 3; e++) {  // simplify() of P_{2e} - P_{2e+1}
        out[e] = 0.0;
        if (have[2 * e] || have[2 * e + 1]) {
            const double d = p[2 * e] + (-p[2 * e + 1]);
            if (norm_le(d * d, thr)) {
                r[e] = __dadd_ru(r[e], fabs(d));
            } else {
                out[e] = d;
                any = true;
            }
        }
    }
    if (!any) return false;
    return prune_or_keep<3>(out, thr, r);  // simplify() of the stack
}
// centre and radii of cross(A, B) written by threads 0..2 (rad = pruned amounts of the whole operation)
K1_DI void cross_header(PZ8 h8, const PZH& A, const PZH& B, bool outerA, const double* rad) {
    const int tid = k1_tid();
    if (tid < 3) {
        const PZH h = view<3>(h8);
        const int e = tid;
        const int e1 = (e + 1) % 3, e2 = (e + 2) % 3;
        const double aA1 = outerA ? abs_serial_entry(A, e1) : abs_collect_entry(A, e1);
        const double aA2 = outerA ? abs_serial_entry(A, e2) : abs_collect_entry(A, e2);
        const double aB1 = outerA ? abs_collect_entry(B, e1) : abs_serial_entry(B, e1);
        const double aB2 = outerA ? abs_collect_entry(B, e2) : abs_serial_entry(B, e2);
        const double* ca = pz_c(A);
        const double* cb = pz_c(B);
        // r_e = a_{e1} b_{e2} - a_{e2} b_{e1}
        pz_c(h)[e] = ca[e1] * cb[e2] - ca[e2] * cb[e1];
#pragma unroll 1
        for (int lane = 0; lane < 2; lane++) {
            const double* ra = pz_r(A, lane);
            const double* rb = pz_r(B, lane);
            // radius of a scalar product x*y: rx*ry + (|x|*ry + rx*|y|)   (KPR/PZsparse.cu:944-989)
            const double p0 = __dadd_ru(__dmul_ru(ra[e1], rb[e2]),
                                        __dadd_ru(__dmul_ru(aA1, rb[e2]), __dmul_ru(ra[e1], aB2)));
            const double p1 = __dadd_ru(__dmul_ru(ra[e2], rb[e1]),
                                        __dadd_ru(__dmul_ru(aA2, rb[e1]), __dmul_ru(ra[e2], aB1)));
            pz_r(h, lane)[e] = __dadd_ru(__dadd_ru(p0, p1), rad[e]);
        }
    }
}

// List path of cross(PZ, PZ): everything in shared memory, no accumulators in the table.  The hash table holds
// only the key set; linked lists group the contributing terms ("pairs") by slot; the owner thread of a slot
// adds its terms up in registers in the fixed order [A_i x centre(B)], [centre(A) x B_j], pairs by ascending
// outer index — the order of the table path below, so both paths give bit-identical results — and runs the
// simplify chain.  The few survivors are parked as (key, coef) records and written out ranked by key.
// Pool layout: keys[cap] u64 | list heads[cap] u32 | links[P] u16 | survivor records (4 words each).
// Returns false (nothing written) when the pool is too small for this operand pair: the caller takes the table path.
K1_DI bool cross_list_path(int top, const PZH& A, const PZH& B, PZ8* out_h) {
    K1S& S = k1s();
    const int tid = k1_tid(), lane = tid & 31, warp = tid >> 5;
    const int nA = A.n, nB = B.n;
    const bool outerA = nA <= nB;
    const int nI = outerA ? nB : nA;
    const long long Pll = (long long)nA + nB + (long long)nA * nB;
    if (Pll > 65535) return false;  // terms are listed by their 16-bit number
    const int P = int(Pll);
    // Scratch: the group's pool, extended downwards into whatever the arena has free above `top` (the operands lie
    // below `top`; the output block goes AT `top`, after the table is dead - OUT_RESERVE words are kept clear of the
    // table for it, and a larger output may run into the dead keys but never into the survivor records, see n_cap).
    constexpr int OUT_RESERVE = 9 + 4 * 64;
    char* base = tab_s0();
    size_t pool_bytes = size_t(S.tab_s_bytes);
    bool borrowed = false;
    {
        const int base_words = (top + OUT_RESERVE + 1) & ~1;
        if (base_words < S.AW) {
            base = reinterpret_cast<char*>(arena0() + base_words);
            pool_bytes += size_t(S.AW - base_words) * 8;
            borrowed = true;
        }
    }
    // the table holds keys only and equal keys share a slot: sized for a load of at most 0.94 if every term had
    // its own key (about half of them do), 2/3 when that still fits
    int cap = 64, lg = 6;
    while (cap * 2 < P * 3) {
        cap <<= 1;
        lg++;
    }
    if (size_t(cap) * 12 + size_t(P) * 2 + 2048 > pool_bytes && (cap >> 1) >= P + (P >> 4)) {
        cap >>= 1;
        lg--;
    }
    size_t fixed = (size_t(cap) * 12 + size_t(P) * 2 + 7) & ~size_t(7);
    bool in_global = false;
    if (fixed + 32 * 32 > pool_bytes) {
        // Shared memory is too small (the lock-step throughput configuration leaves a group ~10 KB): the same lists
        // in the group's global pool (L2-resident).  Still far less traffic than the accumulator table - one atomic
        // per term instead of six read-modify-writes, sums in registers.  The pool is all zero between operations.
        cap = 64, lg = 6;
        while (cap * 2 < P * 3) {
            cap <<= 1;
            lg++;
        }
        fixed = (size_t(cap) * 12 + size_t(P) * 2 + 7) & ~size_t(7);
        if (fixed + 32 * 128 > size_t(S.tab_g_bytes)) return false;
        base = S.tab_g;
        pool_bytes = size_t(S.tab_g_bytes);
        borrowed = false;
        in_global = true;
        if (tid == 0) S.n_tab_global++;
    }
    int surv_max = int((pool_bytes - fixed) / 32);
    if (borrowed) {
        const int n_cap = 64 + cap / 4 - 4;  // output of n monomials = 9 + 4n words <= OUT_RESERVE + the key array
        surv_max = surv_max < n_cap ? surv_max : n_cap;
    }
    if (in_global && surv_max > 2048) surv_max = 2048;
    Tab t;
    t.keys = reinterpret_cast<u64*>(base);
    t.acc = nullptr;
    t.cap = cap;
    t.shift = 64 - lg;
    unsigned* cnt = reinterpret_cast<unsigned*>(base + size_t(cap) * 8);
    unsigned short* list = reinterpret_cast<unsigned short*>(cnt + cap);
    double* surv = reinterpret_cast<double*>(base + fixed);
    {
        unsigned* z = reinterpret_cast<unsigned*>(base);
        if (!in_global)
            for (int i = tid; i < cap * 3; i += NT) z[i] = 0u;
        if (tid == 0) S.nsurv = 0;
    }
    k1_sync();
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    const u64* kA = pz_keys(A);
    const u64* kB = pz_keys(B);
    const double* cA = pz_coef(A);
    const double* cB = pz_coef(B);
    const u64* kO = outerA ? kA : kB;
    const u64* kI = outerA ? kB : kA;
    // the key set
    for (int q = tid; q < P; q += NT) {
        u64 key;
        if (q < nA) {
            key = kA[q];
        } else if (q < nA + nB) {
            key = kB[q - nA];
        } else {
            const int r = q - nA - nB, o = r / nI, j = r - o * nI;
            key = kO[o] + kI[j];
        }
        tab_insert(t, key);
    }
    if (outerA) abs_sum_partial<3>(B); else abs_sum_partial<3>(A);
    k1_sync();
    // terms per slot as linked lists: head[s] = 1 + the last term that arrived, next[q] = 1 + the one before it (0 ends
    // the list).  One pass with one atomic exchange per term; the arrival order does not matter, the owner visits
    // the terms in ascending term number.  Term numbers q: [0, nA) A_i x centre(B), [nA, nA + nB) centre(A) x B_j,
    // then outer o x inner j at nA + nB + o * nI + j (ascending q = the fixed order of the sums).
    for (int q = tid; q < P; q += NT) {
        u64 key;
        if (q < nA) {
            key = kA[q];
        } else if (q < nA + nB) {
            key = kB[q - nA];
        } else {
            const int r = q - nA - nB, o = r / nI, j = r - o * nI;
            key = kO[o] + kI[j];
        }
        list[q] = (unsigned short)atomicExch(&cnt[tab_find(t, key)], unsigned(q + 1));
    }
    k1_sync();
    // owner pass, slots distributed like tab_finalize (same order of the pruned-amount sums)
    double rad[3] = {0.0, 0.0, 0.0};
    {
        const double* ccA = pz_c(A);
        const double* ccB = pz_c(B);
        const double cenA[3] = {ccA[0], ccA[1], ccA[2]};
        const double cenB[3] = {ccB[0], ccB[1], ccB[2]};
        const int seg = (cap / NW) < 32 ? 32 : (cap / NW);
        const int s0 = warp * seg;
        const int s1 = (s0 + seg) < cap ? (s0 + seg) : cap;
        for (int s = s0 + lane; s < s1; s += 32) {
            const u64 key = t.keys[s];
            if (key == 0) continue;
            const unsigned head = cnt[s];
            double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            int last = -1;
            for (;;) {
                int best = 0x7fffffff;  // next term in ascending order (the terms of a slot are distinct)
                for (unsigned x = head; x != 0u; x = list[x - 1]) {
                    const int u = int(x) - 1;
                    if (u > last && u < best) best = u;
                }
                if (best == 0x7fffffff) break;
                last = best;
                double v[6];
                if (best < nA) {
                    cross_six(cA + size_t(best) * 3, cenB, v);
                } else if (best < nA + nB) {
                    cross_six(cenA, cB + size_t(best - nA) * 3, v);
                } else {
                    const int r = best - nA - nB, o = r / nI, j = r - o * nI;
                    if (outerA)
                        cross_six(cA + size_t(o) * 3, cB + size_t(j) * 3, v);
                    else
                        cross_six(cA + size_t(j) * 3, cB + size_t(o) * 3, v);
                }
#pragma unroll
                for (int x = 0; x < 6; x++) a[x] += v[x];
            }
            double out[3];
            if (cross_fin(thr, a, out, rad)) {
                const int i = atomicAdd(&S.nsurv, 1);
                if (i < surv_max) {
                    reinterpret_cast<u64*>(surv)[size_t(i) * 4] = key;
                    surv[size_t(i) * 4 + 1] = out[0];
                    surv[size_t(i) * 4 + 2] = out[1];
                    surv[size_t(i) * 4 + 3] = out[2];
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 3; e++) rad[e] = warp_sum_up(rad[e]);
    if (lane == 0)
#pragma unroll
        for (int e = 0; e < 3; e++) S.red[warp * RED_STRIDE + e] = rad[e];
    k1_sync();
    const int n = S.nsurv;
    if (n > surv_max) {  // uniform; nothing of the output exists yet
        if (in_global) {
            k1_sync();
            unsigned* z = reinterpret_cast<unsigned*>(base);
            const int words = int((fixed + size_t(surv_max) * 32) / 4);
            for (int i = tid; i < words; i += NT) z[i] = 0u;
            k1_sync();
        }
        return false;
    }
    double radt[3];
    rad_collect<3>(radt);
    bool ok;
    const PZ8 h8 = pz_alloc<3>(top, n, &ok);
    if (ok) {
        const PZH h = view<3>(h8);
        u64* okk = pz_keys(h);
        double* oc = pz_coef(h);
        for (int r = tid; r < n; r += NT) {  // survivors ranked by key (the records are in arrival order)
            const u64 key = reinterpret_cast<const u64*>(surv)[size_t(r) * 4];
            int rank = 0;
            for (int x = 0; x < n; x++) rank += (reinterpret_cast<const u64*>(surv)[size_t(x) * 4] < key);
            okk[rank] = key;
#pragma unroll
            for (int e = 0; e < 3; e++) oc[size_t(rank) * 3 + e] = surv[size_t(r) * 4 + 1 + e];
        }
        cross_header(h8, A, B, outerA, radt);
    }
    k1_sync();
    if (in_global) {  // hand the pool back all zero
        unsigned* z = reinterpret_cast<unsigned*>(base);
        const int words = int((fixed + size_t(n) * 32) / 4);
        for (int i = tid; i < words; i += NT) z[i] = 0u;
        k1_sync();
    }
    *out_h = h8;
    return true;
}

K1_OP PZ8 op_cross(int top, PZ8 A8, PZ8 B8) {
    const PZ8 dummy = {0, 0};
    K1S& S = k1s();
    if (S.fail) return dummy;
    const int tid = k1_tid();
    const PZH A = view<3>(A8), B = view<3>(B8);
    const int nA = A.n, nB = B.n;
    {
        PZ8 h8;
        if (cross_list_path(top, A, B, &h8)) return h8;
        if (S.fail) return dummy;
    }
#ifdef K1_PROFILE
    if (tid == 0) {
        atomicAdd(&g_k1spill[3], 1);
        atomicMax(&g_k1spill[4], nA + nB + nA * nB);
    }
#endif
    Tab t;
    if (!tab_select(nA + nB + nA * nB, 6, t)) return dummy;
    if (reinterpret_cast<char*>(t.keys) == tab_s0()) {  // merge operations leave the shared pool dirty
        u64* z = t.keys;
        const int words = t.cap * 7;
        for (int i = tid; i < words; i += NT) z[i] = 0;
        k1_sync();
    }
    const double thr = S.thr2;  // squared-threshold constant of norm_le
    const u64* kA = pz_keys(A);
    const u64* kB = pz_keys(B);
    const double* cA = pz_coef(A);
    const double* cB = pz_coef(B);
    const bool outerA = nA <= nB;
    const int nO = outerA ? nA : nB, nI = outerA ? nB : nA;
    const u64* kO = outerA ? kA : kB;
    const u64* kI = outerA ? kB : kA;
    for (int j = tid; j < nI; j += NT) {
        const u64 kj = kI[j];
        tab_insert(t, kj);
        for (int o = 0; o < nO; o++) tab_insert(t, kj + kO[o]);
    }
    for (int o = tid; o < nO; o += NT) tab_insert(t, kO[o]);
    if (outerA) abs_sum_partial<3>(B); else abs_sum_partial<3>(A);
    k1_sync();
    {
        const double* cc = pz_c(B);
        const double cen[3] = {cc[0], cc[1], cc[2]};
        for (int i = tid; i < nA; i += NT) {
            double v[6];
            cross_six(cA + size_t(i) * 3, cen, v);
            const int slot = tab_find(t, kA[i]);
#pragma unroll
            for (int e = 0; e < 6; e++) t.acc[size_t(slot) * 6 + e] += v[e];
        }
    }
    k1_sync();
    {
        const double* cc = pz_c(A);
        const double cen[3] = {cc[0], cc[1], cc[2]};
        for (int j = tid; j < nB; j += NT) {
            double v[6];
            cross_six(cen, cB + size_t(j) * 3, v);
            const int slot = tab_find(t, kB[j]);
#pragma unroll
            for (int e = 0; e < 6; e++) t.acc[size_t(slot) * 6 + e] += v[e];
        }
    }
    k1_sync();
    for (int o = 0; o < nO; o++) {
        const u64 ko = kO[o];
        const double* co = (outerA ? cA : cB) + size_t(o) * 3;
        const double ov[3] = {co[0], co[1], co[2]};
        for (int j = tid; j < nI; j += NT) {
            double v[6];
            if (outerA)
                cross_six(ov, cB + size_t(j) * 3, v);
            else
                cross_six(cA + size_t(j) * 3, ov, v);
            const int slot = tab_find(t, ko + kI[j]);
#pragma unroll
            for (int e = 0; e < 6; e++) t.acc[size_t(slot) * 6 + e] += v[e];
        }
        k1_sync();
    }
    double rad[3];
    bool ok;
    const PZ8 h8 = tab_finalize<6, 3>(
        top, t, [thr](const double* a, double* out, double* r) { return cross_fin(thr, a, out, r); }, rad, &ok);
    if (ok) cross_header(h8, A, B, outerA, rad);
    k1_sync();
    sort_block<3>(h8, ok);  // survivors leave the table in slot order
    k1_sync();
    return h8;
}

}  // namespace K1_NS
}  // namespace armour
