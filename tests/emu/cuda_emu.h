// TEST INFRASTRUCTURE ONLY.  A small cooperative SIMT emulator so the CUDA kernels under
// fastf_b200/csrc can be compiled with g++ (-DFASTF_EMU) and logic-tested in a container that has
// no GPU.  One CTA runs at a time; each CUDA thread is a fiber (hand-rolled x86-64 stack switch: no syscalls); warp collectives and
// __syncthreads are rendezvous points.  A round in which no fiber makes progress is reported as a
// deadlock (catches divergent collectives).  FASTF_EMU_SHUFFLE=<seed> permutes the fiber visiting
// order each round to shake out missing-barrier bugs.  Nothing here models performance.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void *cudaStream_t;
struct uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };

// ---- just enough of the CUDA runtime for the host orchestration (capi.cu) to run on host memory ----
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct EmuEvent { double t; };
typedef EmuEvent *cudaEvent_t;
#include <time.h>
inline double emu_now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("FASTF_EMU_DEVICES"); *n = e ? atoi(e) : 1; return cudaSuccess; }   // several "devices" = the same host
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = *total_b = (size_t)16 << 30; return cudaSuccess; }   // the emulated device has "16 GiB"
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent{0}; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new EmuEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static const
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

inline dim3 threadIdx, blockIdx, blockDim, gridDim;
using std::max;
using std::min;

#if !defined(__x86_64__)
#error "cuda_emu.h: the fiber switch below is x86-64 System V only"
#endif
// saves the callee-saved registers on the current stack, stores its stack pointer in *save_sp, continues on new_sp
extern "C" void fastf_emu_switch(void **save_sp, void *new_sp);
__asm__(".text\n.globl fastf_emu_switch\n.type fastf_emu_switch,@function\nfastf_emu_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
        ".size fastf_emu_switch, .-fastf_emu_switch\n");

namespace emu {
struct Fiber {
    void *sp = nullptr;   // saved stack pointer while the fiber is switched out
    char *stack = nullptr;
    bool done = true;
};
struct Rdv {
    uint64_t vals[32], out[32];
    uint32_t arrived = 0, gen = 0, toread = 0;
    bool reading = false;
};
struct State {
    std::vector<Fiber> fibers;
    void *sched_sp = nullptr;
    int cur = -1, live = 0, nthreads = 0;
    bool progress = false;
    uint32_t bar_arrived = 0, bar_gen = 0;
    std::map<uint64_t, Rdv> rdv;   // key = warp << 32 | mask
    std::function<void()> body;
    uint64_t shuffle_seed = 0;
    std::vector<unsigned char> dyn;   // dynamic shared memory of the running CTA
    uint64_t spin_rounds = 0;
    bool spun = false;
};
inline State g;
static const size_t STACK_BYTES = 256 * 1024;

inline void yield() { fastf_emu_switch(&g.fibers[g.cur].sp, g.sched_sp); }
// a thread polling shared memory for another warp: lets the others run; endless polling by everybody is reported
inline void spin_yield() { g.spun = true; yield(); }
inline unsigned char *dyn_smem() { return g.dyn.data(); }
inline void fiber_entry()
{
    g.body();
    Fiber &f = g.fibers[g.cur];
    f.done = true;
    g.live--;
    g.progress = true;
    if (g.live > 0 && g.bar_arrived == (uint32_t)g.live && g.bar_arrived) { g.bar_arrived = 0; g.bar_gen++; }
    fastf_emu_switch(&f.sp, g.sched_sp);
    abort();   // a finished fiber is never resumed
}
inline void set_tid(int i)
{
    threadIdx.x = i % blockDim.x;
    threadIdx.y = (i / blockDim.x) % blockDim.y;
    threadIdx.z = i / (blockDim.x * blockDim.y);
}
template <class F> void launch(dim3 grid, dim3 block, size_t smem_bytes, F &&body)
{
    g.dyn.assign(smem_bytes + 16, 0);
    const char *sh = getenv("FASTF_EMU_SHUFFLE");
    g.shuffle_seed = sh ? strtoull(sh, 0, 10) : 0;
    gridDim = grid; blockDim = block;
    int T = (int)(block.x * block.y * block.z);
    if ((int)g.fibers.size() < T) g.fibers.resize(T);
    g.body = body;
    std::vector<int> order(T);
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx = dim3(bx, by, bz);
        g.nthreads = T; g.live = T; g.bar_arrived = 0; g.bar_gen = 0; g.rdv.clear();
        for (int i = 0; i < T; i++) {
            Fiber &f = g.fibers[i];
            if (!f.stack) f.stack = (char *)malloc(STACK_BYTES);
            // initial frame: six callee-saved register slots, then the entry address `ret` jumps to (rsp % 16 == 8 on entry)
            void **top = (void **)(((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15);
            top -= 8;
            for (int r = 0; r < 6; r++) top[r] = nullptr;
            top[6] = (void *)fiber_entry;
            top[7] = nullptr;
            f.sp = (void *)top;
            f.done = false;
            order[i] = i;
        }
        uint64_t rs = g.shuffle_seed * 0x9e3779b97f4a7c15ull + bx + 1;
        while (g.live > 0) {
            g.progress = false;
            if (g.shuffle_seed) for (int i = T - 1; i > 0; i--) { rs = rs * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[i], order[(rs >> 33) % (i + 1)]); }
            for (int k = 0; k < T; k++) {
                int i = order[k];
                if (g.fibers[i].done) continue;
                g.cur = i;
                set_tid(i);
                fastf_emu_switch(&g.sched_sp, g.fibers[i].sp);
            }
            if (!g.progress && g.spun && g.live > 0) {
                // only pollers ran: fine as long as it does not go on forever
                g.spun = false;
                if (++g.spin_rounds > 200000000ull) { fprintf(stderr, "cuda_emu: LIVELOCK in block %u (threads only poll)\n", bx); abort(); }
                continue;
            }
            g.spun = false;
            if (g.progress) g.spin_rounds = 0;
            if (!g.progress && g.live > 0) {
                fprintf(stderr, "cuda_emu: DEADLOCK in block %u (%d live threads blocked; divergent barrier/collective?)\n", bx, g.live);
                abort();
            }
        }
    }
}
inline void syncthreads()
{
    uint32_t my = g.bar_gen;
    g.bar_arrived++;
    g.progress = true;
    if (g.bar_arrived == (uint32_t)g.live) { g.bar_arrived = 0; g.bar_gen++; return; }
    while (g.bar_gen == my) yield();
}
// every lane named in mask deposits v; returns the 32 deposited values (only lanes in mask are meaningful)
struct Gathered { uint64_t v[32]; };
inline Gathered exchange(uint32_t mask, uint64_t v)
{
    int tid = g.cur;
    uint32_t lane = (uint32_t)tid & 31u, warp = (uint32_t)tid >> 5;
    if (!(mask & (1u << lane))) { fprintf(stderr, "cuda_emu: lane %u not in its own collective mask %08x\n", lane, mask); abort(); }
    // lanes beyond the CTA size do not exist
    uint32_t exist = (g.nthreads - (int)warp * 32 >= 32) ? 0xffffffffu : ((1u << (g.nthreads - warp * 32)) - 1u);
    uint32_t need = mask & exist;
    Rdv &R = g.rdv[((uint64_t)warp << 32) | mask];
    while (R.reading) yield();
    R.vals[lane] = v;
    R.arrived |= 1u << lane;
    g.progress = true;
    uint32_t my = R.gen;
    if (R.arrived == need) {
        memcpy(R.out, R.vals, sizeof R.out);
        R.reading = true; R.toread = (uint32_t)__builtin_popcount(need); R.gen++; R.arrived = 0;
    } else {
        while (R.gen == my) yield();
    }
    Gathered G;
    memcpy(G.v, R.out, sizeof G.v);
    if (--R.toread == 0) R.reading = false;
    g.progress = true;
    return G;
}
}   // namespace emu

#define FASTF_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })

inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::exchange(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) {}

template <class T> inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32)
{
    emu::Gathered G = emu::exchange(mask, emu_bits(v));
    uint32_t lane = threadIdx.x & 31u;
    uint32_t s = (lane & ~(uint32_t)(width - 1)) | ((uint32_t)src & (uint32_t)(width - 1));
    if (!(mask & (1u << s))) return v;
    return emu_unbits<T>(G.v[s]);
}
template <class T> inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32)
{
    emu::Gathered G = emu::exchange(mask, emu_bits(v));
    uint32_t lane = threadIdx.x & 31u;
    uint32_t base = lane & ~(uint32_t)(width - 1);
    if ((lane - base) < delta) return v;
    return emu_unbits<T>(G.v[lane - delta]);
}
template <class T> inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32)
{
    emu::Gathered G = emu::exchange(mask, emu_bits(v));
    uint32_t lane = threadIdx.x & 31u;
    uint32_t base = lane & ~(uint32_t)(width - 1);
    if ((lane - base) + delta >= (uint32_t)width) return v;
    return emu_unbits<T>(G.v[lane + delta]);
}
template <class T> inline T __shfl_xor_sync(uint32_t mask, T v, int lanemask, int width = 32)
{
    emu::Gathered G = emu::exchange(mask, emu_bits(v));
    uint32_t lane = threadIdx.x & 31u;
    uint32_t s = lane ^ (uint32_t)lanemask;
    if ((s & ~(uint32_t)(width - 1)) != (lane & ~(uint32_t)(width - 1))) return v;
    return emu_unbits<T>(G.v[s]);
}
inline uint32_t __ballot_sync(uint32_t mask, int pred)
{
    emu::Gathered G = emu::exchange(mask, pred ? 1 : 0);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && G.v[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(uint32_t mask, int pred)
{
    emu::Gathered G = emu::exchange(mask, pred ? 1 : 0);
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && !G.v[i] && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads) return 0;
    return 1;
}
template <class T> inline uint32_t __match_any_sync(uint32_t mask, T v)
{
    emu::Gathered G = emu::exchange(mask, emu_bits(v));
    uint64_t mine = emu_bits(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads && G.v[i] == mine) r |= 1u << i;
    return r;
}
inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v)
{
    emu::Gathered G = emu::exchange(mask, v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads) r += (uint32_t)G.v[i];
    return r;
}
inline uint32_t __reduce_or_sync(uint32_t mask, uint32_t v)
{
    emu::Gathered G = emu::exchange(mask, v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads) r |= (uint32_t)G.v[i];
    return r;
}
inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v)
{
    emu::Gathered G = emu::exchange(mask, v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads) r = std::max(r, (uint32_t)G.v[i]);
    return r;
}
inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v)
{
    emu::Gathered G = emu::exchange(mask, v);
    uint32_t r = 0xffffffffu;
    for (int i = 0; i < 32; i++) if ((mask & (1u << i)) && (int)((threadIdx.x & ~31u) + i) < emu::g.nthreads) r = std::min(r, (uint32_t)G.v[i]);
    return r;
}

inline float __frcp_rn(float x) { return 1.0f / x; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __popcll(uint64_t v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((uint32_t)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline uint32_t __brev(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { uint64_t x = ((uint64_t)hi << 32) | lo; return (uint32_t)(x >> (s & 31)); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
