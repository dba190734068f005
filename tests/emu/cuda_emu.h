// TEST INFRASTRUCTURE — a minimal single-OS-thread CUDA execution emulator.
//
// Lets the library's kernels (written for nvcc / sm_100a) be compiled with g++ and executed on the
// CPU *inside tests only*, so kernel logic can be checked against the oracle on a box without a GPU.
// Every CUDA thread of a block is a ucontext fiber; __syncthreads / warp collectives yield to a
// round-robin scheduler.  Blocks run one after another.  This is not a product path: nothing under
// armour_b200/ includes it, and the shipped library has no CPU fallback.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "../../oracle/interval.h"  // directed-rounding primitives (test side may use the oracle)

#define ARMOUR_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr int cudaSuccess = 0;

namespace emu {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    dim3 tid;
};
struct Warp {
    unsigned long long xch[32];
    int arrive = 0;
    unsigned gen = 0;
};
struct State {
    std::vector<Fiber> fibers;
    ucontext_t main;
    int cur = -1;
    int nthreads = 0;
    int bar_count = 0;
    unsigned bar_gen = 0;
    Warp warps[32];
    std::function<void()> body;
    char* dyn_smem = nullptr;
};
inline State& S() {
    static State s;
    return s;
}
inline void yield() {
    State& s = S();
    swapcontext(&s.fibers[s.cur].ctx, &s.main);
}
inline void trampoline() {
    State& s = S();
    s.body();
    s.fibers[s.cur].done = true;
    swapcontext(&s.fibers[s.cur].ctx, &s.main);
}
inline int linear_tid() { return int(threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y); }

inline void syncthreads() {
    State& s = S();
    const unsigned g = s.bar_gen;
    if (++s.bar_count == s.nthreads) {
        s.bar_count = 0;
        s.bar_gen++;
    } else {
        while (s.bar_gen == g) yield();
    }
}
inline void warp_barrier(unsigned mask) {
    State& s = S();
    Warp& w = s.warps[linear_tid() / 32];
    const int expect = __builtin_popcount(mask);
    const unsigned g = w.gen;
    if (++w.arrive == expect) {
        w.arrive = 0;
        w.gen++;
    } else {
        while (w.gen == g) yield();
    }
}
template <class T>
inline T shfl_generic(unsigned mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    Warp& w = S().warps[linear_tid() / 32];
    const int lane = linear_tid() & 31;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.xch[lane] = bits;
    warp_barrier(mask);
    unsigned long long r = w.xch[src & 31];
    if (!((mask >> (src & 31)) & 1)) r = bits;
    warp_barrier(mask);
    T out;
    std::memcpy(&out, &r, sizeof(T));
    return out;
}

// run `body` as a kernel: grid x block fibers, `smem_bytes` of dynamic shared memory
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
    State& s = S();
    gridDim = grid;
    blockDim = block;
    const int nt = int(block.x * block.y * block.z);
    std::vector<char> smem(smem_bytes + 64);
    s.dyn_smem = smem.data();
    s.body = body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx = dim3(bx, by, bz);
                s.nthreads = nt;
                s.bar_count = 0;
                for (auto& w : s.warps) w.arrive = 0;
                s.fibers.assign(nt, Fiber());
                for (int i = 0; i < nt; i++) {
                    Fiber& f = s.fibers[i];
                    f.stack.resize(512 * 1024);
                    f.tid = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack.data();
                    f.ctx.uc_stack.ss_size = f.stack.size();
                    f.ctx.uc_link = &s.main;
                    makecontext(&f.ctx, trampoline, 0);
                }
                int alive = nt;
                while (alive > 0) {
                    alive = 0;
                    for (int i = 0; i < nt; i++) {
                        if (s.fibers[i].done) continue;
                        alive++;
                        s.cur = i;
                        threadIdx = s.fibers[i].tid;
                        swapcontext(&s.main, &s.fibers[i].ctx);
                    }
                }
            }
    s.dyn_smem = nullptr;
}

}  // namespace emu

// ---- CUDA built-ins ---------------------------------------------------------------------------------
inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int = 32) { return emu::shfl_generic(mask, v, src); }
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int = 32) {
    return emu::shfl_generic(mask, v, (emu::linear_tid() & 31) ^ lanemask);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int = 32) {
    const int lane = emu::linear_tid() & 31;
    return emu::shfl_generic(mask, v, lane >= int(d) ? lane - int(d) : lane);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int = 32) {
    const int lane = emu::linear_tid() & 31;
    return emu::shfl_generic(mask, v, lane + int(d) < 32 ? lane + int(d) : lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Warp& w = emu::S().warps[emu::linear_tid() / 32];
    const int lane = emu::linear_tid() & 31;
    w.xch[lane] = pred ? 1 : 0;
    emu::warp_barrier(mask);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask >> i) & 1) && w.xch[i]) r |= 1u << i;
    emu::warp_barrier(mask);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }

template <class T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T>
inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T>
inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T>
inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

inline double __dadd_ru(double a, double b) { return orc::rnd::add_up(a, b); }
inline double __dadd_rd(double a, double b) { return orc::rnd::add_dn(a, b); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_ru(double a, double b) { return orc::rnd::sub_up(a, b); }
inline double __dsub_rd(double a, double b) { return orc::rnd::sub_dn(a, b); }
inline double __dmul_ru(double a, double b) { return orc::rnd::mul_up(a, b); }
inline double __dmul_rd(double a, double b) { return orc::rnd::mul_dn(a, b); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rd(double a, double b) { return orc::rnd::div_dn(a, b); }
inline double __dsqrt_ru(double a) { return orc::rnd::sqrt_up(a); }
inline double __dsqrt_rd(double a) { return orc::rnd::sqrt_dn(a); }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
using std::max;
using std::min;
