// tests/cuda_emul/cuda_emul.h -- a tiny cooperative CUDA-thread emulator for the CPU.
//
// TEST INFRASTRUCTURE ONLY.  It lets the *unmodified* kernel sources under
// lsc_dr_planner_b200/csrc/*.cuh be compiled with g++ and stepped on the CPU (one CTA at a time,
// every CUDA thread a ucontext fiber, __syncthreads / __syncwarp / __shfl_* implemented with
// cooperative yields), so index and race bugs are caught without a GPU round trip.  The product
// never links this: liblscqp.so is built by nvcc only and fails loudly without a device.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __noinline__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct double2 { double x, y; };
struct float2 { float x, y; };
template <class T> inline T __ldg(const T* p) { return *p; }
namespace emu {
struct Thread {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};
struct State {
    std::vector<Thread> threads;
    ucontext_t sched;
    int cur = 0;
    int nthreads = 0;
    // block barrier
    int arrived = 0; unsigned long gen = 0;
    // warp barriers / shuffle slots
    std::vector<int> warr; std::vector<unsigned long> wgen;
    std::vector<double> slot_d;
    std::function<void()> body;
    std::vector<char> smem;
};
inline State& st() { static State s; return s; }
inline void yield() { State& s = st(); swapcontext(&s.threads[s.cur].ctx, &s.sched); }
inline void trampoline() {
    State& s = st();
    s.body();
    s.threads[s.cur].done = true;
    swapcontext(&s.threads[s.cur].ctx, &s.sched);
}
}  // namespace emu

extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
#ifdef CUDA_EMUL_IMPL
thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

inline void __syncthreads() {
    emu::State& s = emu::st();
    const unsigned long g = s.gen;
    if (++s.arrived == s.nthreads) { s.arrived = 0; s.gen++; return; }
    while (s.gen == g) emu::yield();
}
inline int __syncthreads_or(int pred);
inline void __syncwarp(unsigned = 0xffffffffu) {
    emu::State& s = emu::st();
    const int w = s.cur >> 5;
    const int wsize = std::min(32, s.nthreads - (w << 5));
    const unsigned long g = s.wgen[w];
    if (++s.warr[w] == wsize) { s.warr[w] = 0; s.wgen[w]++; return; }
    while (s.wgen[w] == g) emu::yield();
}
inline double __shfl_xor_sync(unsigned, double v, int mask) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = v;
    __syncwarp();
    const int src = (s.cur & ~31) | ((s.cur & 31) ^ mask);
    const double r = s.slot_d[src];
    __syncwarp();
    return r;
}
inline int __shfl_xor_sync(unsigned m, int v, int mask) { return (int) __shfl_xor_sync(m, (double) v, mask); }
inline double __shfl_sync(unsigned, double v, int src_lane) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = v;
    __syncwarp();
    const double r = s.slot_d[(s.cur & ~31) | (src_lane & 31)];
    __syncwarp();
    return r;
}
inline double __shfl_down_sync(unsigned, double v, int delta) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = v;
    __syncwarp();
    const int lane = s.cur & 31;
    const double r = lane + delta < 32 ? s.slot_d[s.cur + delta] : v;
    __syncwarp();
    return r;
}
inline int __shfl_up_sync(unsigned, int v, int delta) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = (double) v;
    __syncwarp();
    const int lane = s.cur & 31;
    const int r = lane >= delta ? (int) s.slot_d[s.cur - delta] : v;
    __syncwarp();
    return r;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = pred ? 1.0 : 0.0;
    __syncwarp();
    const int w0 = s.cur & ~31, wn = std::min(32, s.nthreads - w0);
    unsigned r = 0;
    for (int i = 0; i < wn; i++) if (s.slot_d[w0 + i] != 0.0) r |= 1u << i;
    __syncwarp();
    return r;
}
inline int __any_sync(unsigned, int pred) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = pred ? 1.0 : 0.0;
    __syncwarp();
    const int w0 = s.cur & ~31, wn = std::min(32, s.nthreads - w0);
    int r = 0;
    for (int i = 0; i < wn; i++) r |= s.slot_d[w0 + i] != 0.0;
    __syncwarp();
    return r;
}
inline int __syncthreads_or(int pred) {
    // two accumulators keyed by the parity of the barrier generation, each tagged with the generation it belongs to
    // (a thread that runs ahead into the next barrier must not disturb the value the others still have to read)
    emu::State& s = emu::st();
    static int acc[2] = {0, 0};
    static unsigned long tag[2] = {~0ul, ~0ul};
    const unsigned long g = s.gen;
    const int slot = (int) (g & 1);
    if (tag[slot] != g) { acc[slot] = 0; tag[slot] = g; }
    if (pred) acc[slot] = 1;
    __syncthreads();
    return acc[slot];
}
inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }   // fibers are cooperative
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline void __threadfence_system() {}
inline void __threadfence() {}
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __fdividef(float a, float b) { return a / b; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int) v); }
inline long long __double_as_longlong(double v) { long long r; std::memcpy(&r, &v, 8); return r; }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    emu::State& s = emu::st();
    s.slot_d[s.cur] = (double) v;
    __syncwarp();
    const int w0 = s.cur & ~31, wn = std::min(32, s.nthreads - w0);
    double r = s.slot_d[w0];
    for (int i = 1; i < wn; i++) r = std::min(r, s.slot_d[w0 + i]);
    __syncwarp();
    return (unsigned) r;
}
#undef __launch_bounds__
#define __launch_bounds__(...)
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fsqrt_rn(float a) { volatile float r = sqrtf(a); return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }

// the kernel's `extern __shared__ double sm[];`
#define __shared__
extern double* emu_dyn_smem;
#ifdef CUDA_EMUL_IMPL
double* emu_dyn_smem = nullptr;
#endif
#define EMU_SHARED_DECL(name) double* name = emu_dyn_smem

namespace emu {
// Run `kernel(args)` over grid x block threads, one CTA at a time.
template <class F>
void launch(unsigned grid, unsigned block, size_t smem_bytes, F&& kernel_call) {
    State& s = st();
    s.nthreads = (int) block;
    s.smem.assign(smem_bytes + 64, 0);
    // poison shared memory with NaNs so that reads of unwritten cells show up
    {
        double* d = reinterpret_cast<double*>(s.smem.data());
        for (size_t i = 0; i < smem_bytes / 8; i++) d[i] = NAN;
    }
    emu_dyn_smem = reinterpret_cast<double*>(s.smem.data());
    s.warr.assign((block + 31) / 32, 0); s.wgen.assign((block + 31) / 32, 0);
    s.slot_d.assign(block, 0.0);
    s.body = kernel_call;
    for (unsigned b = 0; b < grid; b++) {
        s.threads.clear(); s.threads.resize(block);
        s.arrived = 0;
        for (unsigned t = 0; t < block; t++) {
            Thread& th = s.threads[t];
            th.stack.resize(256 * 1024);
            getcontext(&th.ctx);
            th.ctx.uc_stack.ss_sp = th.stack.data();
            th.ctx.uc_stack.ss_size = th.stack.size();
            th.ctx.uc_link = &s.sched;
            makecontext(&th.ctx, (void (*)()) trampoline, 0);
        }
        int remaining = (int) block;
        while (remaining > 0) {
            for (unsigned t = 0; t < block; t++) {
                if (s.threads[t].done) continue;
                s.cur = (int) t;
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
                swapcontext(&s.sched, &s.threads[t].ctx);
                if (s.threads[t].done) remaining--;
            }
        }
    }
}
}  // namespace emu
