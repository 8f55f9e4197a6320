// libfastf_gpu.so -- implementation of include/fastf_gpu.h: host orchestration of the sm_100a kernels.
//
// Everything that touches read data runs on the device (kernels in the .cuh files next to this one).
// The host side here only (a) walks the BGZF BSIZE chain (~20 bytes per <= 64 KiB block), (b) moves
// bytes, (c) sizes buffers from device-side counters and (d) launches.  There is no CPU fallback:
// without a CUDA device fastf_ctx_create fails and nothing else can be called.
//
// Streams: `compute` runs the per-chunk chain  header? -> inflate -> parse -> counts -> gather;
// `copy` brings the next chunk's compressed bytes in (fastf_bam2db_feed) while the previous chunk is
// inflating; `mt` extends the MT19937 keep-bit stream while chunks are being parsed.  Per-chunk state
// is double buffered ("slots"), so the only host wait per chunk is for the previous chunk's 32-byte
// counter snapshot, which sizes the candidate array.
#include "../../include/fastf_gpu.h"
#include "common.cuh"
#include "bgzf_index.h"
#include "bgzf_inflate.cuh"
#include "bgzf_inflate_tps.cuh"
#include "bgzf_crc32.cuh"
#include "bam_parse.cuh"
#include "scan_mt_sample.cuh"
#include "radix_dedup.cuh"
#include "mt_jump.h"
#include "freq.cuh"
#include "bam_tags.cuh"
#include "bam_straddle.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>
#include <unordered_map>
#include <string>
#include <chrono>
#ifndef FASTF_EMU
#include <cuda.h>   // driver API: cuMemBatchDecompressAsync (Blackwell hardware decompression engine)
#endif
static double wall_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#define FASTF_ABI_VERSION 1

struct PoolEntry { void *p; size_t cap; };
struct fastf_ctx {
    int device;
    cudaStream_t compute, copy, mt, infl;
    char err[1024];
    u32 launches;   // kernels launched through this context (bench: gpu_launches)
    int n_sm;
    bool tps_attr_set;
    bool crc_attr_set;
    u64 taghist_chunk_blocks = 0, taghist_round0_mask = 0;   // fastf_taghist_test_hooks
    void *mtj_polys;   // device copy of x^(2^k) mod phi, k = 0..44 (uploaded on first use)
    void *mtj_scratch;
    // size-bucketed caches of device / pinned allocations: a job's buffers are recycled by the next job on the same
    // context, so steady-state calls do not pay cudaMalloc / cudaMallocHost (both synchronise the device)
    std::vector<PoolEntry> *dev_pool, *pin_pool;
    // Sizes the buffers of the last bam2db job ended with.  A buffer that grows in mid-job drains every stream (dev_reserve) and
    // leaves the device idle while the host queues the next chunk, so the next job starts its growing buffers at these sizes.
    size_t hint_cand = 0, hint_keepbits = 0, hint_ring = 0;
};

static int ctx_fail(fastf_ctx *ctx, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (ctx) vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
    va_end(ap);
    if (ctx && getenv("FASTF_VERBOSE")) fprintf(stderr, "fastf_gpu: %s\n", ctx->err);
    return 1;
}
#define CK(call)                                                                                                               \
    do {                                                                                                                       \
        cudaError_t e_ = (call);                                                                                               \
        if (e_ != cudaSuccess) return ctx_fail(ctx, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
    } while (0)
#define CKL(name)                                                                                                              \
    do {                                                                                                                       \
        ctx->launches++;                                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                                                   \
        if (e_ != cudaSuccess) return ctx_fail(ctx, "%s:%d: launch of %s -> %s", __FILE__, __LINE__, name, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(expr)                                                                                                              \
    do {                                                                                                                       \
        int r_ = (expr);                                                                                                       \
        if (r_) return r_;                                                                                                     \
    } while (0)

static const char *status_string(u32 st, char *buf, size_t n)
{
    static const char *names[] = {"bad-btype", "bad-stored", "bad-codelens", "bad-symbol", "bad-distance", "out-overflow", "size-mismatch", "in-overrun",
                                  "record-straddles-bgzf-block", "record-corrupt", "umi-too-long", "aux-corrupt", "bad-bam-header", "bgzf-crc32-mismatch", "tag-not-a-string"};
    buf[0] = 0;
    for (u32 b = 0; b < 15; b++)
        if (st & (1u << b)) { strncat(buf, names[b], n - strlen(buf) - 2); strncat(buf, " ", n - strlen(buf) - 1); }
    return buf;
}

// ---- growable device / pinned buffers, recycled through the context's pools ------------------------
static void *pool_take(std::vector<PoolEntry> &pool, size_t bytes)
{
    int best = -1;
    for (size_t i = 0; i < pool.size(); i++)
        if (pool[i].cap >= bytes && (best < 0 || pool[i].cap < pool[(size_t)best].cap)) best = (int)i;
    if (best < 0 || pool[(size_t)best].cap > 2 * bytes + (1u << 20)) return nullptr;   // do not burn a huge buffer on a small request
    void *p = pool[(size_t)best].p;
    pool.erase(pool.begin() + best);
    return p;
}
static size_t pool_cap_of(const std::vector<PoolEntry> &pool, size_t bytes)
{
    size_t best = 0;
    for (auto &e : pool)
        if (e.cap >= bytes && (!best || e.cap < best)) best = e.cap;
    return best;
}
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T> T *as() const { return (T *)p; }
};
static void dev_release(fastf_ctx *ctx, DevBuf &b)
{
    if (b.p) ctx->dev_pool->push_back(PoolEntry{b.p, b.cap});
    b.p = nullptr;
    b.cap = 0;
}
static int dev_reserve(fastf_ctx *ctx, DevBuf &b, size_t bytes, size_t keep_bytes = 0, cudaStream_t s = 0)
{
    if (bytes <= b.cap) return 0;
    size_t ncap = std::max(bytes, b.cap + b.cap / 2);
    ncap = (ncap + 255) & ~(size_t)255;
    size_t pc = pool_cap_of(*ctx->dev_pool, ncap);
    void *np = nullptr;
    if (pc && pc <= 2 * ncap + (1u << 20)) { np = pool_take(*ctx->dev_pool, ncap); ncap = pc; }
    if (!np) CK(cudaMalloc(&np, ncap));
    if (b.p) {
        if (keep_bytes) CK(cudaMemcpyAsync(np, b.p, keep_bytes, cudaMemcpyDeviceToDevice, s));
        // the old buffer goes back to the pool: nothing in flight on any of our streams may still touch it
        CK(cudaStreamSynchronize(ctx->compute));
        CK(cudaStreamSynchronize(ctx->copy));
        CK(cudaStreamSynchronize(ctx->infl));
        CK(cudaStreamSynchronize(ctx->mt));
        dev_release(ctx, b);
    }
    b.p = np;
    b.cap = ncap;
    return 0;
}
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T> T *as() const { return (T *)p; }
};
static void pin_release(fastf_ctx *ctx, PinBuf &b)
{
    if (b.p) ctx->pin_pool->push_back(PoolEntry{b.p, b.cap});
    b.p = nullptr;
    b.cap = 0;
}
static int pin_reserve(fastf_ctx *ctx, PinBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    size_t ncap = std::max(bytes, b.cap * 2);
    ncap = (ncap + 63) & ~(size_t)63;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->compute));
        CK(cudaStreamSynchronize(ctx->copy));
        CK(cudaStreamSynchronize(ctx->infl));
        CK(cudaStreamSynchronize(ctx->mt));
        pin_release(ctx, b);
    }
    size_t pc = pool_cap_of(*ctx->pin_pool, ncap);
    if (pc && pc <= 2 * ncap + (1u << 20)) { b.p = pool_take(*ctx->pin_pool, ncap); b.cap = pc; }
    if (!b.p) { CK(cudaMallocHost(&b.p, ncap)); b.cap = ncap; }
    return 0;
}
static void pools_trim(fastf_ctx *ctx)
{
    for (auto &e : *ctx->dev_pool) cudaFree(e.p);
    for (auto &e : *ctx->pin_pool) cudaFreeHost(e.p);
    ctx->dev_pool->clear();
    ctx->pin_pool->clear();
}

// Large results go to pageable host memory the caller owns (malloc'ed result arrays).  A plain cudaMemcpy into pageable memory is
// staged by the driver at a few GB/s; here the bytes cross PCIe into two pinned bounce buffers (from the context's pool) while
// the host copies the previous piece out, so the transfer runs at memcpy speed.  Synchronous: dst is complete on return.
// memcpy of a bounce piece into the caller's (fresh, never touched) pageable memory: page faults and a single core's copy rate
// would otherwise bound the D2H of a few hundred MB of results, so a few threads take a slice each
static void par_memcpy(void *dst, const void *src, size_t n)
{
    const size_t SLICE = (size_t)4 << 20;
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = std::min<size_t>(std::min<size_t>(hw > 1 ? hw / 2 : 1, 8), n / SLICE);
    if (nt < 2) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    for (size_t t = 1; t < nt; t++) {
        const size_t o = t * per;
        if (o < n) th.emplace_back([=] { memcpy((u8 *)dst + o, (const u8 *)src + o, std::min(per, n - o)); });
    }
    memcpy(dst, src, std::min(per, n));
    for (auto &x : th) x.join();
}
static int d2h_pageable(fastf_ctx *ctx, void *dst, const void *src_dev, size_t bytes, cudaStream_t s)
{
    if (bytes == 0) return 0;
    const size_t PIECE = (size_t)16 << 20;
    if (bytes <= ((size_t)1 << 20)) {
        CK(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return 0;
    }
    PinBuf bounce[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int rc = 0;
    for (int k = 0; k < 2 && !rc; k++) { rc = pin_reserve(ctx, bounce[k], PIECE); if (!rc && cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) rc = ctx_fail(ctx, "d2h: event creation failed"); }
    const size_t n_pieces = (bytes + PIECE - 1) / PIECE;
    for (size_t k = 0; k <= n_pieces && !rc; k++) {
        if (k < n_pieces) {
            const size_t o = k * PIECE, m = std::min(PIECE, bytes - o);
            if (cudaMemcpyAsync(bounce[k & 1].p, (const u8 *)src_dev + o, m, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaEventRecord(ev[k & 1], s) != cudaSuccess) rc = ctx_fail(ctx, "d2h: copy failed");
        }
        if (k > 0 && !rc) {
            const size_t o = (k - 1) * PIECE, m = std::min(PIECE, bytes - o);
            if (cudaEventSynchronize(ev[(k - 1) & 1]) != cudaSuccess) rc = ctx_fail(ctx, "d2h: copy failed");
            else par_memcpy((u8 *)dst + o, bounce[(k - 1) & 1].p, m);
        }
    }
    cudaStreamSynchronize(s);
    for (int k = 0; k < 2; k++) { if (ev[k]) cudaEventDestroy(ev[k]); pin_release(ctx, bounce[k]); }
    return rc;
}

struct Timer {   // CUDA-event stopwatch on one stream; accumulates into *acc at collect()
    cudaEvent_t a = nullptr, b = nullptr;
    bool armed = false;
    int init() { return cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess; }
    void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    void stop(cudaStream_t s) { cudaEventRecord(b, s); armed = true; }
    void collect(float *acc) { if (armed) { float ms = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); *acc += ms; armed = false; } }
};

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
extern "C" int fastf_abi_version(void) { return FASTF_ABI_VERSION; }
#ifndef FASTF_SRC_HASH
#define FASTF_SRC_HASH "unknown"
#endif
#define FASTF_STR2(x) #x
#define FASTF_STR(x) FASTF_STR2(x)
#if FASTF_TPS_PROF
// debug builds only (-DFASTF_TPS_PROF=1, scripts/inflate_ab.py): the kernel's role counters, see bgzf_inflate_tps.cuh
extern "C" int fastf_debug_tps_prof(unsigned long long *out16, int reset)
{
    if (out16 && cudaMemcpyFromSymbol(out16, g_fastf_tps_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_fastf_tps_prof, z, sizeof z) != cudaSuccess) return 1; }
    return 0;
}
#endif
extern "C" const char *fastf_build_info(void)
{
    return "streams=" FASTF_STR(FASTF_TPS_STREAMS) " lanes=" FASTF_STR(FASTF_TPS_LANES) " svc=" FASTF_STR(FASTF_TPS_SVC_WARPS) " lbits=" FASTF_STR(FASTF_TPS_LBITS) " dbits=" FASTF_STR(FASTF_TPS_DBITS)
           " ring=" FASTF_STR(FASTF_TPS_RING) " src=" FASTF_SRC_HASH;
}

static char g_create_err[512] = "";

extern "C" int fastf_ctx_create(int device, fastf_ctx **out)
{
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        snprintf(g_create_err, sizeof g_create_err, "no CUDA device (%s); libfastf_gpu has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return 1;
    }
    if (device < 0 || device >= n) { snprintf(g_create_err, sizeof g_create_err, "device %d out of range (0..%d)", device, n - 1); return 1; }
    if (cudaSetDevice(device) != cudaSuccess) { snprintf(g_create_err, sizeof g_create_err, "cudaSetDevice(%d) failed", device); return 1; }
    fastf_ctx *ctx = (fastf_ctx *)calloc(1, sizeof *ctx);
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->mt, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->infl, cudaStreamNonBlocking) != cudaSuccess) {
        snprintf(g_create_err, sizeof g_create_err, "cudaStreamCreate failed");
        free(ctx);
        return 1;
    }
#ifdef FASTF_EMU
    ctx->n_sm = 2;
#else
    if (cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->n_sm <= 0) ctx->n_sm = 148;
    // LZ77 match sources are short reads at random places of windows that do not fit the L2 together: a miss should fetch one
    // 32-byte sector, not a 64/128-byte neighbourhood (measured in profiles/: DRAM read traffic of the inflate kernel)
    if (const char *g = getenv("FASTF_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
#endif
    std::thread([] { fastf_mtj::tables(); }).detach();   // GF(2) jump tables (0.2 s of host work) while the first job streams
    ctx->dev_pool = new std::vector<PoolEntry>();
    ctx->pin_pool = new std::vector<PoolEntry>();
    *out = ctx;
    return 0;
}
extern "C" void fastf_ctx_destroy(fastf_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->mtj_polys) cudaFree(ctx->mtj_polys);
    if (ctx->mtj_scratch) cudaFree(ctx->mtj_scratch);
    pools_trim(ctx);
    delete ctx->dev_pool;
    delete ctx->pin_pool;
    cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->copy);
    cudaStreamDestroy(ctx->mt);
    cudaStreamDestroy(ctx->infl);
    free(ctx);
}
extern "C" const char *fastf_last_error(const fastf_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" uint32_t fastf_launch_count(const fastf_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *fastf_compute_stream(const fastf_ctx *ctx) { return ctx ? (void *)ctx->compute : nullptr; }
extern "C" int fastf_host_alloc(fastf_ctx *ctx, size_t bytes, void **out) { CK(cudaSetDevice(ctx->device)); CK(cudaMallocHost(out, bytes ? bytes : 1)); return 0; }
extern "C" void fastf_host_free(fastf_ctx *ctx, void *p) { (void)ctx; if (p) cudaFreeHost(p); }
extern "C" int fastf_device_alloc(fastf_ctx *ctx, size_t bytes, void **out) { CK(cudaSetDevice(ctx->device)); CK(cudaMalloc(out, bytes ? bytes : 1)); return 0; }
extern "C" void fastf_device_free(fastf_ctx *ctx, void *p) { (void)ctx; if (p) cudaFree(p); }
extern "C" int fastf_memcpy_h2d(fastf_ctx *ctx, void *d, const void *s, size_t n) { CK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, ctx->compute)); CK(cudaStreamSynchronize(ctx->compute)); return 0; }
static int d2h_pageable(fastf_ctx *ctx, void *dst, const void *src_dev, size_t bytes, cudaStream_t s);
extern "C" int fastf_memcpy_d2h(fastf_ctx *ctx, void *d, const void *s, size_t n) { CK(cudaSetDevice(ctx->device)); return d2h_pageable(ctx, d, s, n, ctx->compute); }
extern "C" int fastf_synchronize(fastf_ctx *ctx) { CK(cudaStreamSynchronize(ctx->copy)); CK(cudaStreamSynchronize(ctx->infl)); CK(cudaStreamSynchronize(ctx->mt)); CK(cudaStreamSynchronize(ctx->compute)); return 0; }
extern "C" void fastf_free(void *p) { free(p); }
extern "C" void fastf_ctx_trim(fastf_ctx *ctx) { if (ctx) { cudaSetDevice(ctx->device); cudaDeviceSynchronize(); pools_trim(ctx); } }

// ---------------------------------------------------------------------------------------------------
// host helpers that define the sampling contract
// ---------------------------------------------------------------------------------------------------
// keep <=> genrand_real1() < rate  with genrand_real1() = genrand_int32()*(1.0/4294967295.0)
// (reference src/mt19937ar.c:149-153) and the comparison done in double against the float rate
// (reference src/bam2db_ds.c:385-390).  The product is monotone in u, so the rule is u < T.
extern "C" uint64_t fastf_keep_threshold(float rate_depth)
{
    const double r = (double)rate_depth;
    if (r <= 0.0) return 0;     // every draw is >= rate: nothing is kept.  A NaN rate falls through: `x >= NaN` is false for every draw, the reference
                                // drops nothing, and the search below ends at 2^32 (keep everything)
    uint64_t lo = 0, hi = 4294967296ull;   // smallest u with u*(1/4294967295) >= r, or 2^32 if none
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        double x = (double)(uint32_t)mid * (1.0 / 4294967295.0);
        if (x >= r) hi = mid; else lo = mid + 1;
    }
    return lo;
}

namespace {
struct HostMT {   // Matsumoto-Nishimura mt19937ar, restated (reference src/mt19937ar.c:60-73,105-140); only for SampleInt's <= n_cells draws
    u32 mt[624];
    int mti;
    void init(u32 s)
    {
        mt[0] = s;
        for (mti = 1; mti < 624; mti++) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (u32)mti;
    }
    u32 next()
    {
        if (mti >= 624) {
            for (int k = 0; k < 624; k++) {
                u32 y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            mti = 0;
        }
        u32 y = mt[mti++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};
}   // namespace

// reference src/bam2db_ds.c:240-244 + SampleInt src/utils.c:29-75 (without replacement) + qsort(vsI)
extern "C" uint64_t fastf_sample_cells(uint64_t n_cells, float rate_cell, uint32_t seed, uint64_t *out, uint64_t *d0)
{
    const float prod = (float)n_cells * rate_cell;   // size_t * float -> float arithmetic
    if (!(prod >= 0.0f) || prod >= 18446744073709551616.0f) return UINT64_MAX;
    const uint64_t ns = (uint64_t)prod;
    if (ns > n_cells) return UINT64_MAX;             // the reference prints a message and exit(1)s
    if (ns == n_cells) {
        for (uint64_t i = 0; i < n_cells; i++) out[i] = i;
        if (d0) *d0 = 0;
        return ns;
    }
    HostMT mt;
    mt.init(seed);
    std::vector<uint64_t> pool(n_cells);
    for (uint64_t i = 0; i < n_cells; i++) pool[i] = i;
    uint64_t ntotal = n_cells;
    for (uint64_t i = 0; i < ns; i++) {
        uint64_t idx = (uint64_t)mt.next() % ntotal;
        out[i] = pool[idx];
        if (idx != ntotal - 1) pool[idx] = pool[ntotal - 1];
        ntotal--;
    }
    std::sort(out, out + ns);
    if (d0) *d0 = ns;
    return ns;
}

// The reference prints its histogram BST in pre-order (src/filter.c:139-148).  The BST built by inserting keys in read order
// (src/filter.c:105-124) is the Cartesian tree of `first` (first-occurrence ordinal) over the keys in ascending byte order:
// the earliest key is the root, smaller keys form its left subtree, larger ones its right subtree, recursively.
// order_out[k] = index (in ascending key order) of the k-th line of whitelist.txt.  O(n), iterative.
extern "C" int fastf_cartesian_preorder(const uint32_t *first, uint64_t n, uint64_t *order_out)
{
    if (n == 0) return 0;
    const int64_t NIL = -1;
    std::vector<int64_t> left(n, NIL), right(n, NIL), stack;
    stack.reserve(64);
    for (uint64_t i = 0; i < n; i++) {
        int64_t last = NIL;
        while (!stack.empty() && first[stack.back()] > first[i]) { last = stack.back(); stack.pop_back(); }
        left[i] = last;
        if (!stack.empty()) right[stack.back()] = (int64_t)i;
        stack.push_back((int64_t)i);
    }
    const int64_t root = stack.front();
    stack.clear();
    stack.push_back(root);
    uint64_t k = 0;
    while (!stack.empty()) {
        int64_t v = stack.back();
        stack.pop_back();
        order_out[k++] = (uint64_t)v;
        if (right[v] != NIL) stack.push_back(right[v]);
        if (left[v] != NIL) stack.push_back(left[v]);
    }
    return k == n ? 0 : 1;
}

extern "C" int64_t fastf_bgzf_index_host(const void *buf, size_t n, uint64_t *in_off, uint32_t *in_len, uint32_t *isize, uint64_t cap, size_t *consumed)
{
    std::vector<FastfBgzfBlock> blocks;
    size_t used = 0;
    int rc = fastf_bgzf_index((const uint8_t *)buf, n, 0, blocks, &used);
    if (consumed) *consumed = used;
    if (rc != FASTF_BGZF_OK && rc != FASTF_BGZF_NEED_MORE) return -1;
    if (blocks.size() > cap) return -2;
    for (size_t i = 0; i < blocks.size(); i++) { in_off[i] = blocks[i].in_off; in_len[i] = blocks[i].in_len; isize[i] = blocks[i].isize; }
    return (int64_t)blocks.size();
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
#define FASTF_INFLATE_HW_ENGINE 0x100u   // flag in the `inflate_lanes` argument (include/fastf_gpu.h)

// after a hardware-engine batch: actual byte counts -> status words
__global__ void __launch_bounds__(256) fastf_de_check_kernel(u32 *__restrict__ act_status, const u32 *__restrict__ isize, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) act_status[i] = (isize[i] != 0 && act_status[i] != isize[i]) ? (u32)FASTF_ST_SIZE_MISMATCH : 0u;
}

struct DeScratch {   // per-launch state of the inflate engines
#ifndef FASTF_EMU
    std::vector<CUmemDecompressParams> params;   // parameter array of a hardware-engine batch; must stay alive until the batch has run
#endif
    DevBuf counter;                               // work counter of the persistent thread-per-stream kernel
    DevBuf sorted;                                // its per-stream sorted-symbol lists (global scratch)
};
#define FASTF_INFLATE_TPS 1u   // inflate_lanes 1..4 select a shape of the thread-per-stream kernel; 8/16/32 the lock-step kernel
#define FASTF_INFLATE_DEFAULT 2u   // 0 = default: the thread-per-stream kernel

// CRC-32 of every inflated block against its BGZF trailer (htslib does this in bgzf_read_block); sets FASTF_ST_BAD_CRC in status[]
static int launch_crc(fastf_ctx *ctx, u32 lanes, const u8 *comp, u64 comp_total, const u64 *in_off, const u32 *in_len, const u8 *infl, const u64 *out_off, const u32 *isize, u32 nblocks,
                      u32 *status, cudaStream_t s)
{
    if (nblocks == 0 || (lanes & FASTF_INFLATE_NO_CRC)) return 0;
    u32 grid = (nblocks + FASTF_CRC_WARPS - 1) / FASTF_CRC_WARPS;
    if (grid > (u32)ctx->n_sm) grid = (u32)ctx->n_sm;   // one CTA per SM (the tables are built once per CTA); the rest is a grid-stride loop
#ifndef FASTF_EMU
    if (!ctx->crc_attr_set) { CK(cudaFuncSetAttribute(fastf_bgzf_crc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FastfCrcTables))); ctx->crc_attr_set = true; }
#endif
    FASTF_LAUNCH(fastf_bgzf_crc32_kernel, grid, FASTF_CRC_WARPS * 32, sizeof(FastfCrcTables), s, comp, comp_total, in_off, in_len, infl, out_off, isize, nblocks, status);
    CKL("bgzf_crc32");
    return 0;
}

// h_* = host copies of the block index (needed to build the engine's parameter array)
static int launch_inflate(fastf_ctx *ctx, u32 lanes, const u8 *comp, u64 comp_total, const u64 *in_off, const u32 *in_len, const u64 *out_off, const u32 *isize, u32 nblocks, u8 *out,
                          u32 *status, cudaStream_t s, DeScratch *de, const u64 *h_in_off, const u32 *h_in_len, const u64 *h_out_off, const u32 *h_isize)
{
    if (nblocks == 0) return 0;
    if (lanes & FASTF_INFLATE_HW_ENGINE) {
#ifdef FASTF_EMU
        return ctx_fail(ctx, "inflate: the hardware decompression engine does not exist in the emulator build");
#else
        // Blackwell decompression engine: one DEFLATE operation per BGZF block, submitted as one batch in stream order.
        // dstActBytes lands in the status array and is turned into status bits by a small kernel afterwards.
        de->params.clear();
        de->params.reserve(nblocks);
        for (u32 i = 0; i < nblocks; i++) {
            if (h_isize[i] == 0) continue;   // empty (EOF) blocks produce nothing
            CUmemDecompressParams p;
            memset(&p, 0, sizeof p);
            p.srcNumBytes = h_in_len[i];
            p.dstNumBytes = h_isize[i];
            p.dstActBytes = (cuuint32_t *)(status + i);
            p.src = comp + h_in_off[i];
            p.dst = out + h_out_off[i];
            p.algo = CU_MEM_DECOMPRESS_ALGORITHM_DEFLATE;
            de->params.push_back(p);
        }
        CK(cudaMemsetAsync(status, 0, (size_t)nblocks * sizeof(u32), s));
        if (!de->params.empty()) {
            // the driver entry point is resolved through the runtime: no link-time dependency on libcuda (absent on build hosts)
            typedef CUresult (*decompress_fn)(CUmemDecompressParams *, size_t, unsigned int, size_t *, CUstream);
            static decompress_fn fn = nullptr;
            if (!fn) {
                void *sym = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint("cuMemBatchDecompressAsync", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !sym)
                    return ctx_fail(ctx, "inflate: this driver does not export cuMemBatchDecompressAsync (hardware decompression engine unavailable)");
                fn = (decompress_fn)sym;
            }
            size_t erri = 0;
            CUresult r = fn(de->params.data(), de->params.size(), 0, &erri, (CUstream)s);
            if (r != CUDA_SUCCESS)
                return ctx_fail(ctx, "inflate: cuMemBatchDecompressAsync failed (CUresult %d) at operation %zu; is the hardware decompression engine available on this GPU?", (int)r, erri);
        }
        FASTF_LAUNCH(fastf_de_check_kernel, (nblocks + 255) / 256, 256, 0, s, status, isize, nblocks);
        CKL("de_check");
        return 0;
#endif
    }
    lanes &= 0xffu;   // the kernel shape; the flag bits (no-CRC, straddle) were for the callers
    if (lanes >= 1 && lanes <= 4) {
        // thread-per-stream kernel: persistent CTAs (one per SM), FASTF_TPS_STREAMS streams each; blocks are handed out by a global counter.
        // one shape is built: <FASTF_TPS_LANES decoding lanes per decoder warp, FASTF_TPS_SVC_WARPS service warps> (bgzf_inflate_tps.cuh); lanes 1..4 all select it
        const size_t smem = sizeof(FastfTpsStream) * FASTF_TPS_STREAMS + sizeof(FastfTpsShared);
        TRY(dev_reserve(ctx, de->counter, 64));
        CK(cudaMemsetAsync(de->counter.p, 0, sizeof(u32), s));
        FastfTpsArgs A;
        A.comp = comp; A.comp_total = comp_total; A.in_off = in_off; A.in_len = in_len; A.out_off = out_off; A.isize = isize; A.nblocks = nblocks; A.out = out; A.status = status;
        A.next_block = de->counter.as<u32>();
        u32 grid = (nblocks + FASTF_TPS_STREAMS - 1) / FASTF_TPS_STREAMS;
        if (grid > (u32)ctx->n_sm) grid = (u32)ctx->n_sm;
        TRY(dev_reserve(ctx, de->sorted, (size_t)grid * FASTF_TPS_STREAMS * FASTF_TPS_SORTED_U16 * sizeof(u16)));
        A.sorted = de->sorted.as<u16>();
#ifndef FASTF_EMU
        if (!ctx->tps_attr_set) {
            CK(cudaFuncSetAttribute(fastf_bgzf_inflate_tps_kernel<FASTF_TPS_LANES, FASTF_TPS_SVC_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->tps_attr_set = true;
        }
#endif
        FASTF_LAUNCH((fastf_bgzf_inflate_tps_kernel<FASTF_TPS_LANES, FASTF_TPS_SVC_WARPS>), grid, FASTF_TPS_THREADS, smem, s, A);
        CKL("bgzf_inflate_tps");
        return 0;
    }
    if (lanes == 8) {
        FASTF_LAUNCH(fastf_bgzf_inflate_kernel<8>, (nblocks + 3) / 4, 32, 0, s, comp, comp_total, in_off, in_len, out_off, isize, nblocks, out, status);
    } else if (lanes == 16) {
        FASTF_LAUNCH(fastf_bgzf_inflate_kernel<16>, (nblocks + 1) / 2, 32, 0, s, comp, comp_total, in_off, in_len, out_off, isize, nblocks, out, status);
    } else {
        FASTF_LAUNCH(fastf_bgzf_inflate_kernel<32>, nblocks, 32, 0, s, comp, comp_total, in_off, in_len, out_off, isize, nblocks, out, status);
    }
    CKL("bgzf_inflate");
    return 0;
}

// exclusive scan of every row of a [nrows][n] u32 matrix in place; totals[nrows]
static int launch_scan_rows(fastf_ctx *ctx, u32 *data, u64 n, u32 nrows, u32 *totals, cudaStream_t s)
{
    FASTF_LAUNCH(fastf_scan_rows_kernel, nrows, FASTF_SCAN_THREADS, 0, s, data, n, totals);
    CKL("scan_rows");
    return 0;
}

// ---- LSD radix sort over a chosen set of 8-bit digit windows --------------------------------------
struct SortScratch {
    DevBuf hist, totals, dbase;
};
// Windows: greedy cover of the bit positions set in `varying` (bits that differ between keys).
static int plan_windows(u64 varying, u32 *shifts)
{
    int n = 0;
    u32 b = 0;
    while (b < 64) {
        if ((varying >> b) & 1ull) { shifts[n++] = b; b += 8; } else b++;
    }
    return n;
}
// Sorts n keys (and optional u32 payload).  keys/alt (and vals/vals_alt) are ping-pong buffers of n elements;
// *sorted_in_alt tells where the result ended up.
static int sort_keys(fastf_ctx *ctx, SortScratch &S, u64 *keys, u64 *alt, u32 *vals, u32 *vals_alt, u64 n, const u32 *shifts, int npass, bool *sorted_in_alt, cudaStream_t s)
{
    *sorted_in_alt = false;
    if (n == 0 || npass == 0) return 0;
    if (n >= 0xffffffffull) return ctx_fail(ctx, "sort: %llu keys exceed the 2^32-1 limit of one device sort", (unsigned long long)n);
    const u32 ntiles = (u32)((n + FASTF_RS_TILE - 1) / FASTF_RS_TILE);
    TRY(dev_reserve(ctx, S.hist, (size_t)256 * ntiles * sizeof(u32)));
    TRY(dev_reserve(ctx, S.totals, 256 * sizeof(u32)));
    TRY(dev_reserve(ctx, S.dbase, 256 * sizeof(u32)));
    u64 *src = keys, *dst = alt;
    u32 *vsrc = vals, *vdst = vals_alt;
    for (int p = 0; p < npass; p++) {
        FASTF_LAUNCH(fastf_radix_hist_kernel, ntiles, FASTF_RS_THREADS, 0, s, (const u64 *)src, n, shifts[p], S.hist.as<u32>(), ntiles);
        CKL("radix_hist");
        TRY(launch_scan_rows(ctx, S.hist.as<u32>(), ntiles, 256, S.totals.as<u32>(), s));
        FASTF_LAUNCH(fastf_radix_digit_base_kernel, 1, 256, 0, s, (const u32 *)S.totals.as<u32>(), S.dbase.as<u32>());
        CKL("radix_digit_base");
        if (vals) {
            FASTF_LAUNCH(fastf_radix_scatter_kernel<true>, ntiles, FASTF_RS_THREADS, 0, s, (const u64 *)src, (const u32 *)vsrc, dst, vdst, n, shifts[p], (const u32 *)S.hist.as<u32>(),
                         (const u32 *)S.dbase.as<u32>(), ntiles);
        } else {
            FASTF_LAUNCH(fastf_radix_scatter_kernel<false>, ntiles, FASTF_RS_THREADS, 0, s, (const u64 *)src, (const u32 *)nullptr, dst, (u32 *)nullptr, n, shifts[p],
                         (const u32 *)S.hist.as<u32>(), (const u32 *)S.dbase.as<u32>(), ntiles);
        }
        CKL("radix_scatter");
        std::swap(src, dst);
        std::swap(vsrc, vdst);
    }
    *sorted_in_alt = (src == alt);
    return 0;
}
static void sort_scratch_release(fastf_ctx *ctx, SortScratch &S) { dev_release(ctx, S.hist); dev_release(ctx, S.totals); dev_release(ctx, S.dbase); }

// OR / AND of all keys -> which bit positions vary (device reduction, 16 bytes back)
__global__ void __launch_bounds__(256) fastf_key_bits_kernel(const u64 *__restrict__ keys, u64 n, u64 *__restrict__ or_and)
{
    u64 o = 0, a = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) { u64 k = keys[i]; o |= k; a &= k; }
    for (int d = 16; d; d >>= 1) { o |= __shfl_xor_sync(FASTF_FULL_MASK, o, d); a &= __shfl_xor_sync(FASTF_FULL_MASK, a, d); }
    if ((threadIdx.x & 31u) == 0) { atomicOr((unsigned long long *)&or_and[0], (unsigned long long)o); atomicAnd((unsigned long long *)&or_and[1], (unsigned long long)a); }
}

// ---- run-length / segmented count over sorted keys ------------------------------------------------
struct RleScratch {
    DevBuf tile_counts, tile_totals, grp_key, grp_first, grp_dstart, grp_val, count, out_gene, out_cell;
    PinBuf totals_host;
};
static void rle_scratch_release(fastf_ctx *ctx, RleScratch &R)
{
    dev_release(ctx, R.tile_counts); dev_release(ctx, R.tile_totals); dev_release(ctx, R.grp_key); dev_release(ctx, R.grp_first); dev_release(ctx, R.grp_dstart); dev_release(ctx, R.grp_val);
    dev_release(ctx, R.count); dev_release(ctx, R.out_gene); dev_release(ctx, R.out_cell);
    pin_release(ctx, R.totals_host);
}
// After this: R.grp_key/grp_first/grp_dstart(/grp_val)/count hold ngroups entries on device; with split_bits_gene > 0
// R.out_gene / R.out_cell hold the split group key.
static int rle_groups(fastf_ctx *ctx, RleScratch &R, const u64 *sorted, const u32 *vals, u64 n, u32 group_shift, u32 nn_bit, u32 split_bits_gene, u64 *ngroups_out, u64 *ndistinct_out,
                      cudaStream_t s)
{
    *ngroups_out = 0;
    if (ndistinct_out) *ndistinct_out = 0;
    if (n == 0) return 0;
    if (n >= 0xffffffffull) return ctx_fail(ctx, "rle: %llu keys exceed the 2^32-1 limit", (unsigned long long)n);
    const u32 ntiles = (u32)((n + FASTF_RLE_TILE - 1) / FASTF_RLE_TILE);
    TRY(dev_reserve(ctx, R.tile_counts, (size_t)2 * ntiles * sizeof(u32)));
    TRY(dev_reserve(ctx, R.tile_totals, 2 * sizeof(u32)));
    TRY(pin_reserve(ctx, R.totals_host, 2 * sizeof(u32)));
    FASTF_LAUNCH(fastf_rle_count_kernel, ntiles, FASTF_RLE_THREADS, 0, s, sorted, n, group_shift, nn_bit, R.tile_counts.as<u32>(), ntiles);
    CKL("rle_count");
    TRY(launch_scan_rows(ctx, R.tile_counts.as<u32>(), ntiles, 2, R.tile_totals.as<u32>(), s));
    CK(cudaMemcpyAsync(R.totals_host.p, R.tile_totals.p, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const u32 ngroups = R.totals_host.as<u32>()[0], ndistinct = R.totals_host.as<u32>()[1];
    TRY(dev_reserve(ctx, R.grp_key, (size_t)ngroups * sizeof(u64)));
    TRY(dev_reserve(ctx, R.grp_first, (size_t)ngroups * sizeof(u32)));
    TRY(dev_reserve(ctx, R.grp_dstart, (size_t)ngroups * sizeof(u32)));
    TRY(dev_reserve(ctx, R.count, (size_t)ngroups * sizeof(u32)));
    if (vals) TRY(dev_reserve(ctx, R.grp_val, (size_t)ngroups * sizeof(u32)));
    if (split_bits_gene) { TRY(dev_reserve(ctx, R.out_gene, (size_t)ngroups * sizeof(u32))); TRY(dev_reserve(ctx, R.out_cell, (size_t)ngroups * sizeof(u32))); }
    FASTF_LAUNCH(fastf_rle_emit_kernel, ntiles, FASTF_RLE_THREADS, 0, s, sorted, vals, n, group_shift, nn_bit, (const u32 *)R.tile_counts.as<u32>(), ntiles, R.grp_key.as<u64>(),
                 R.grp_first.as<u32>(), R.grp_dstart.as<u32>(), vals ? R.grp_val.as<u32>() : (u32 *)nullptr);
    CKL("rle_emit");
    if (ngroups) {
        FASTF_LAUNCH(fastf_rle_finish_kernel, (ngroups + 255) / 256, 256, 0, s, (const u32 *)R.grp_dstart.as<u32>(), ngroups, ndistinct, R.count.as<u32>(), (const u64 *)R.grp_key.as<u64>(),
                     split_bits_gene, split_bits_gene ? R.out_gene.as<u32>() : (u32 *)nullptr, split_bits_gene ? R.out_cell.as<u32>() : (u32 *)nullptr);
        CKL("rle_finish");
    }
    *ngroups_out = ngroups;
    if (ndistinct_out) *ndistinct_out = ndistinct;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// string tables -> device
// ---------------------------------------------------------------------------------------------------
struct DevTable {
    DevBuf slots, pool;
    FastfStrTableView view;
    u32 count = 0;
};
static int table_upload(fastf_ctx *ctx, DevTable &T, const char *keys, const u32 *off, u32 n)
{
    FastfStrTableHost H;
    H.init(n);
    for (u32 i = 0; i < n; i++) {
        // first insertion wins; the reference's hash_table_insert refuses duplicates (src/hashtable.c:70-95)
        H.insert(keys + off[i], off[i + 1] - off[i], i + 1);
    }
    H.finish();
    TRY(dev_reserve(ctx, T.slots, H.slots.size() * sizeof(FastfStrSlot)));
    TRY(dev_reserve(ctx, T.pool, H.pool.size()));
    CK(cudaMemcpy(T.slots.p, H.slots.data(), H.slots.size() * sizeof(FastfStrSlot), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(T.pool.p, H.pool.data(), H.pool.size(), cudaMemcpyHostToDevice));
    T.view.slots = T.slots.as<FastfStrSlot>();
    T.view.pool = T.pool.as<uint8_t>();
    T.view.mask = H.mask;
    memcpy(T.view.pw1, H.pw1, sizeof H.pw1);
    memcpy(T.view.pw2, H.pw2, sizeof H.pw2);
    T.count = n;
    return 0;
}
static u32 bits_for(u32 max_value)
{
    u32 b = 1;
    while (b < 32 && (max_value >> b)) b++;
    return b;
}

// ---------------------------------------------------------------------------------------------------
// chunked BGZF -> inflated bytes machinery shared by bam2db and freq
// ---------------------------------------------------------------------------------------------------
struct BlockIndexDev {   // per-chunk block index on device (one allocation, 8-byte fields first)
    DevBuf buf;
    PinBuf host;
    u32 cap_blocks = 0;
    u64 *in_off, *out_off, *stage_off, *dst_base;
    u32 *in_len, *isize, *nrec, *ncbv, *st_infl, *st_parse;
    u64 *h_in_off, *h_out_off, *h_stage_off;
    u32 *h_in_len, *h_isize;
};
static size_t index_bytes_dev(u32 nb) { return (size_t)nb * (4 * sizeof(u64) + 6 * sizeof(u32)); }
static size_t index_bytes_up(u32 nb) { return (size_t)nb * (3 * sizeof(u64) + 2 * sizeof(u32)); }
static int index_reserve(fastf_ctx *ctx, BlockIndexDev &I, u32 nb)
{
    if (nb <= I.cap_blocks) return 0;
    u32 cap = std::max(nb, I.cap_blocks * 2);
    cap = (cap + 63u) & ~63u;
    TRY(dev_reserve(ctx, I.buf, index_bytes_dev(cap)));
    TRY(pin_reserve(ctx, I.host, index_bytes_up(cap)));
    I.cap_blocks = cap;
    // upload region first (in_off, out_off, stage_off, in_len, isize), device-only region after
    u8 *d = I.buf.as<u8>();
    I.in_off = (u64 *)d; d += (size_t)cap * 8;
    I.out_off = (u64 *)d; d += (size_t)cap * 8;
    I.stage_off = (u64 *)d; d += (size_t)cap * 8;
    I.in_len = (u32 *)d; d += (size_t)cap * 4;
    I.isize = (u32 *)d; d += (size_t)cap * 4;
    I.dst_base = (u64 *)d; d += (size_t)cap * 8;
    I.nrec = (u32 *)d; d += (size_t)cap * 4;
    I.ncbv = (u32 *)d; d += (size_t)cap * 4;
    I.st_infl = (u32 *)d; d += (size_t)cap * 4;
    I.st_parse = (u32 *)d;
    u8 *h = I.host.as<u8>();
    I.h_in_off = (u64 *)h; h += (size_t)cap * 8;
    I.h_out_off = (u64 *)h; h += (size_t)cap * 8;
    I.h_stage_off = (u64 *)h; h += (size_t)cap * 8;
    I.h_in_len = (u32 *)h; h += (size_t)cap * 4;
    I.h_isize = (u32 *)h;
    return 0;
}
// The block index of a chunk (a few MB) is PULLED by a kernel out of the pinned host arrays instead of being pushed through the
// copy engine: there it queues behind the bulk H2D copies of the next chunks' compressed bytes (FIFO per direction) and the inflate
// that waits for it starts up to three copies late (measured: the first inflate of a host-fed job 90 ms after its bytes arrived).
__global__ void __launch_bounds__(256) fastf_pull_words_kernel(u32 *__restrict__ dst, const u32 *__restrict__ src_host, u64 n_words)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (u64)gridDim.x * blockDim.x) dst[i] = src_host[i];
}
static int index_upload(fastf_ctx *ctx, BlockIndexDev &I, cudaStream_t s)
{
    const size_t bytes = index_bytes_up(I.cap_blocks);
#ifdef FASTF_EMU
    CK(cudaMemcpyAsync(I.buf.p, I.host.p, bytes, cudaMemcpyHostToDevice, s));
#else
    void *src = nullptr;
    CK(cudaHostGetDevicePointer(&src, I.host.p, 0));
    FASTF_LAUNCH(fastf_pull_words_kernel, 2 * ctx->n_sm, 256, 0, s, I.buf.as<u32>(), (const u32 *)src, (u64)(bytes / 4));
    CKL("pull_index");
#endif
    return 0;
}
static void index_release(fastf_ctx *ctx, BlockIndexDev &I) { dev_release(ctx, I.buf); pin_release(ctx, I.host); I.cap_blocks = 0; }

// ---------------------------------------------------------------------------------------------------
// bam2db job
// ---------------------------------------------------------------------------------------------------
#define FASTF_DEFAULT_CHUNK (512ull << 20)
#define FASTF_MAX_BLOCKS_PER_CHUNK (1u << 22)

struct ChunkSlot {
    BlockIndexDev idx;
    DevBuf stage;         // per-block candidate staging
    DevBuf virt;          // FASTF_BAM_STRADDLE: record-start guesses and virtual blocks
    DevBuf infl;          // inflated bytes of the chunk (double buffered: chunk i+1 inflates while chunk i is parsed)
    DeScratch de;
    cudaEvent_t ev_infl = nullptr, ev_gather = nullptr;
    PinBuf snap;          // counters snapshot {n_records, n_candidates, status_or, chunk_candidates}
    cudaEvent_t ev_copy = nullptr, ev_done = nullptr;
    u32 nblocks = 0;
    bool pending = false; // parse launched, gather not yet
};

struct fastf_bam2db_job {
    fastf_ctx *ctx;
    fastf_bam2db_params prm;
    FastfKeyLayout L;
    DevTable cells, genes;
    u32 lanes;
    u64 chunk_bytes;
    ChunkSlot slot[2];
    u32 next_slot = 0;
    // compressed bytes of host-fed chunks: a ring of three, so that the copy of chunk i is issued before the host waits for
    // anything and overlaps the inflate of chunks i-2 and i-1
    struct CompRing { DevBuf buf; cudaEvent_t ev_copy = nullptr, ev_free = nullptr; bool used = false; } comp_ring[3];
    u32 comp_seq = 0;
    u64 ring_estimate = 0;
    // FASTF_FEED_TIMING=1: where a host-fed job spends its time (stderr at finish): H2D copies by CUDA events, host waits by wall clock
    bool feed_timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ft_copy_ev, ft_infl_ev;
    std::vector<cudaEvent_t> ft_parse_ev;
    std::vector<double> ft_host_submit;
    u64 ft_copy_bytes = 0;
    double ft_wait_s = 0, ft_index_s = 0, ft_feed_s = 0;
    // Blocks wait here until a chunk is full, ACROSS feed calls: the persistent inflate kernel keeps n_sm x FASTF_TPS_STREAMS blocks
    // in flight, so a launch over an exact multiple of that many blocks has no half-empty last round (measured: +11 % inflate
    // throughput over 2 GiB chunks cut at the feed boundaries).
    struct PendingChunk {
        std::vector<FastfBgzfBlock> blocks;
        u64 infl = 0;
        const u8 *comp_dev = nullptr;   // device feeds: the caller's buffer
        u64 comp_total = 0;
        CompRing *ring = nullptr;       // host feeds: where the compressed bytes are being staged
        u64 fill = 0;                   // bytes staged so far (multiple of 4)
    } pending;
    u64 chunk_blocks = FASTF_MAX_BLOCKS_PER_CHUNK;
    DevBuf counters;          // u64[4]: n_records, n_candidates, status_or, (unused)
    DevBuf hdr_off;           // u64: offset of the first alignment record inside the current chunk
    DevBuf cand;              // all candidates (CB-valid reads) in file order
    u64 cand_cap = 0;
    u64 n_records = 0, n_cand = 0;   // host copies after the last finalized chunk
    u32 status = 0;
    bool header_done = false;
    std::vector<u8> carry;    // partial BGZF block left over by fastf_bam2db_feed
    // MT19937 keep bits
    DevBuf mt_state, keepbits, mt_states, mt_scratch;
    u64 mt_pairs_done = 0;     // twist pairs generated since mt_origin
    u64 mt_origin = 0;         // stream index of bit 0 of keepbits (0 unless the job jumped ahead)
    bool mt_seeded = false;
    cudaEvent_t ev_mt = nullptr;
    // sampling / sort / count
    bool sampled_done = false;
    DevBuf tile_valid, tile_tot, sample_counters, kept, orand;
    PinBuf small_host;
    u64 n_sampled = 0, n_valid = 0;
    SortScratch sortS;
    RleScratch rleS;
    // stats
    u64 n_blocks = 0, comp_bytes = 0, infl_bytes = 0;
    u64 n_blocks_fed = 0, n_blocks_done = 0;   // blocks handed to run_blocks / blocks whose candidate counts have come back
    u32 launches0 = 0, n_chunks = 0;
    Timer t_infl[2], t_crc[2], t_parse[2], t_gather[2], t_mt[2], t_sample, t_sort, t_count;
    u32 mt_launches = 0;
    cudaEvent_t ev_first = nullptr, ev_last = nullptr;
    bool first_recorded = false;
    float ms_inflate = 0, ms_crc = 0, ms_parse = 0, ms_gather = 0, ms_mt = 0, ms_sample = 0, ms_sort = 0, ms_count = 0;
};

static u32 stage_cap_for(u32 isize) { return isize / 36u + 1u; }   // a record is >= 4 + 32 bytes

extern "C" void fastf_bam2db_job_free(fastf_bam2db_job *job)
{
    if (!job) return;
    fastf_ctx *ctx = job->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->compute);
    cudaStreamSynchronize(ctx->copy);
    cudaStreamSynchronize(ctx->infl);
    cudaStreamSynchronize(ctx->mt);
    for (int i = 0; i < 2; i++) {
        ChunkSlot &S = job->slot[i];
        index_release(ctx, S.idx); dev_release(ctx, S.stage); dev_release(ctx, S.virt); dev_release(ctx, S.infl); dev_release(ctx, S.de.counter); dev_release(ctx, S.de.sorted); pin_release(ctx, S.snap);
        if (S.ev_copy) cudaEventDestroy(S.ev_copy);
        if (S.ev_infl) cudaEventDestroy(S.ev_infl);
        if (S.ev_gather) cudaEventDestroy(S.ev_gather);
        if (S.ev_done) cudaEventDestroy(S.ev_done);
        job->t_infl[i].destroy(); job->t_crc[i].destroy(); job->t_parse[i].destroy(); job->t_gather[i].destroy();
    }
    ctx->hint_cand = std::max(ctx->hint_cand, job->cand.cap);
    ctx->hint_keepbits = std::max(ctx->hint_keepbits, job->keepbits.cap);
    for (auto &R : job->comp_ring) ctx->hint_ring = std::max(ctx->hint_ring, R.buf.cap);
    for (auto &R : job->comp_ring) {
        dev_release(ctx, R.buf);
        if (R.ev_copy) cudaEventDestroy(R.ev_copy);
        if (R.ev_free) cudaEventDestroy(R.ev_free);
    }
    job->t_mt[0].destroy(); job->t_mt[1].destroy(); job->t_sample.destroy(); job->t_sort.destroy(); job->t_count.destroy();
    if (job->ev_mt) cudaEventDestroy(job->ev_mt);
    if (job->ev_first) cudaEventDestroy(job->ev_first);
    if (job->ev_last) cudaEventDestroy(job->ev_last);
    dev_release(ctx, job->cells.slots); dev_release(ctx, job->cells.pool); dev_release(ctx, job->genes.slots); dev_release(ctx, job->genes.pool);
    dev_release(ctx, job->counters); dev_release(ctx, job->hdr_off); dev_release(ctx, job->cand); dev_release(ctx, job->mt_state); dev_release(ctx, job->keepbits); dev_release(ctx, job->mt_states); dev_release(ctx, job->mt_scratch);
    dev_release(ctx, job->tile_valid); dev_release(ctx, job->tile_tot); dev_release(ctx, job->sample_counters); dev_release(ctx, job->kept); dev_release(ctx, job->orand);
    pin_release(ctx, job->small_host);
    sort_scratch_release(ctx, job->sortS);
    rle_scratch_release(ctx, job->rleS);
    delete job;
}

extern "C" int fastf_bam2db_begin(fastf_ctx *ctx, const fastf_bam2db_params *p, fastf_bam2db_job **out)
{
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    if (!p || (p->n_cells && (!p->cell_keys || !p->cell_off)) || (p->n_genes && (!p->gene_keys || !p->gene_off))) return ctx_fail(ctx, "bam2db_begin: null table pointers");
    if (p->keep_threshold > 4294967296ull) return ctx_fail(ctx, "bam2db_begin: keep_threshold > 2^32");
    fastf_bam2db_job *job = new fastf_bam2db_job();
    { const char *e = getenv("FASTF_FEED_TIMING"); job->feed_timing = e && *e && *e != '0'; }
    job->ctx = ctx;
    job->prm = *p;
    job->launches0 = ctx->launches;
    {
        const u32 l = p->inflate_lanes & 0xffu;
        job->lanes = ((l == 8 || l == 16 || l == 32 || (l >= 1 && l <= 4)) ? l : FASTF_INFLATE_DEFAULT) | (p->inflate_lanes & (FASTF_INFLATE_HW_ENGINE | FASTF_INFLATE_NO_CRC | FASTF_BAM_STRADDLE));
    }
    job->chunk_bytes = p->chunk_inflated_bytes ? std::max<u64>(p->chunk_inflated_bytes, 1u << 20) : FASTF_DEFAULT_CHUNK;
    // the persistent thread-per-stream kernel keeps 64 streams per SM busy: give every launch several blocks per stream
    if (!p->chunk_inflated_bytes && (job->lanes & 0xffu) >= 1 && (job->lanes & 0xffu) <= 4) {
        u64 rounds = 2;   // full rounds of the persistent kernel per chunk (2 -> 37888 blocks, <= 2.4 GiB on 148 SMs)
        if (const char *e = getenv("FASTF_CHUNK_ROUNDS")) { const long v = atol(e); if (v >= 1 && v <= 16) rounds = (u64)v; }
        job->chunk_blocks = rounds * (u64)ctx->n_sm * FASTF_TPS_STREAMS;
        job->chunk_bytes = job->chunk_blocks * 65536ull;
    }
    if (job->lanes & FASTF_BAM_STRADDLE) {
        // records may run across block boundaries: keep the whole file in one chunk so that none is cut by a chunk boundary
        if (p->headerless) { delete job; return ctx_fail(ctx, "bam2db_begin: FASTF_BAM_STRADDLE needs the whole file in one job (a later shard does not know where its first record starts)"); }
        job->chunk_blocks = FASTF_MAX_BLOCKS_PER_CHUNK;
        job->chunk_bytes = ~0ull >> 2;
    }
    FastfKeyLayout &L = job->L;
    L.umi_max_bytes = p->umi_max_bytes ? p->umi_max_bytes : 3;
    if (L.umi_max_bytes > 4) { delete job; return ctx_fail(ctx, "bam2db_begin: umi_max_bytes must be 1..4 (UMIs up to 16 bases)"); }
    L.bits_umi = 1 + 8 * L.umi_max_bytes + 3;
    L.bits_gene = bits_for(p->n_genes);
    L.bits_cell = bits_for(p->n_cells);
    if (L.bits_cell + L.bits_gene + L.bits_umi > 63) { delete job; return ctx_fail(ctx, "bam2db_begin: key layout needs %u bits (> 63)", L.bits_cell + L.bits_gene + L.bits_umi); }
    int rc = 0;
    rc = rc || table_upload(ctx, job->cells, p->cell_keys, p->cell_off, p->n_cells);
    rc = rc || table_upload(ctx, job->genes, p->gene_keys, p->gene_off, p->n_genes);
    rc = rc || dev_reserve(ctx, job->counters, 4 * sizeof(u64));
    rc = rc || dev_reserve(ctx, job->hdr_off, sizeof(u64));
    rc = rc || dev_reserve(ctx, job->mt_state, 624 * sizeof(u32));
    rc = rc || dev_reserve(ctx, job->sample_counters, 2 * sizeof(u64));
    rc = rc || dev_reserve(ctx, job->orand, 2 * sizeof(u64));
    rc = rc || pin_reserve(ctx, job->small_host, 64);
    for (int i = 0; i < 2 && !rc; i++) {
        rc = rc || pin_reserve(ctx, job->slot[i].snap, 4 * sizeof(u64));
        rc = rc || cudaEventCreateWithFlags(&job->slot[i].ev_copy, cudaEventDisableTiming) != cudaSuccess;
        rc = rc || cudaEventCreateWithFlags(&job->slot[i].ev_done, cudaEventDisableTiming) != cudaSuccess;
        rc = rc || cudaEventCreateWithFlags(&job->slot[i].ev_infl, cudaEventDisableTiming) != cudaSuccess;
        rc = rc || cudaEventCreateWithFlags(&job->slot[i].ev_gather, cudaEventDisableTiming) != cudaSuccess;
        rc = rc || job->t_infl[i].init() || job->t_crc[i].init() || job->t_parse[i].init() || job->t_gather[i].init();
    }
    for (auto &R : job->comp_ring) {
        rc = rc || cudaEventCreateWithFlags(&R.ev_copy, cudaEventDisableTiming) != cudaSuccess;
        rc = rc || cudaEventCreateWithFlags(&R.ev_free, cudaEventDisableTiming) != cudaSuccess;
    }
    rc = rc || job->t_mt[0].init() || job->t_mt[1].init() || job->t_sample.init() || job->t_sort.init() || job->t_count.init();
    rc = rc || cudaEventCreateWithFlags(&job->ev_mt, cudaEventDisableTiming) != cudaSuccess;
    rc = rc || cudaEventCreate(&job->ev_first) != cudaSuccess || cudaEventCreate(&job->ev_last) != cudaSuccess;
    if (!rc) rc = cudaMemsetAsync(job->counters.p, 0, 4 * sizeof(u64), ctx->compute) != cudaSuccess;
    job->header_done = p->headerless != 0;

    if (rc) { if (!ctx->err[0]) ctx_fail(ctx, "bam2db_begin: resource setup failed"); fastf_bam2db_job_free(job); return 1; }
    *out = job;
    return 0;
}

// Extend the keep-bit stream so that it covers stream indices [mt_origin, n_draws).  Runs on the mt stream.
static int mt_extend(fastf_bam2db_job *job, u64 n_draws)
{
    fastf_ctx *ctx = job->ctx;
    if (n_draws <= job->mt_origin) return 0;
    const u64 pairs = (n_draws - job->mt_origin + 1247) / 1248;
    if (pairs <= job->mt_pairs_done) return 0;
    const size_t need = (size_t)pairs * 39 * sizeof(u32);
    if (need > job->keepbits.cap) {
        // grow geometrically; the copy keeps the bits produced so far
        size_t want = std::max(std::max(need + need / 2, (size_t)(64u << 20)), ctx->hint_keepbits);
        TRY(dev_reserve(ctx, job->keepbits, want, (size_t)job->mt_pairs_done * 39 * sizeof(u32), ctx->mt));
    }
    Timer &tm = job->t_mt[job->mt_launches++ & 1u];   // the launch two extensions back has long finished
    tm.collect(&job->ms_mt);
    tm.start(ctx->mt);
    FASTF_LAUNCH(fastf_mt19937_kernel, 1, FASTF_MT_THREADS, 0, ctx->mt, job->prm.seed, job->mt_state.as<u32>(), job->mt_seeded ? 0u : 1u, job->mt_pairs_done, pairs - job->mt_pairs_done,
                 job->prm.keep_threshold, (u32 *)nullptr, job->keepbits.as<u32>());
    CKL("mt19937");
    tm.stop(ctx->mt);
    job->mt_seeded = true;
    job->mt_pairs_done = pairs;
    return 0;
}

// Leave in `state` (624 words, device) the MT19937 window at stream index `origin`: seed, then apply x^(2^k) mod phi for every
// set bit k of origin (jump-ahead, mt_jump.h).  Returns 1 when the polynomial tables are unavailable.
static int mt_state_at(fastf_ctx *ctx, u32 seed, u64 origin, u32 *state, cudaStream_t s)
{
    const fastf_mtj::Tables &T = fastf_mtj::tables();
    if (!T.ok || (origin >> T.pow2.size()) != 0) return 1;
    if (!ctx->mtj_polys) {
        std::vector<uint64_t> flat(T.pow2.size() * FASTF_MT_POLY_WORDS);
        for (size_t k = 0; k < T.pow2.size(); k++) memcpy(flat.data() + k * FASTF_MT_POLY_WORDS, T.pow2[k].data(), FASTF_MT_POLY_WORDS * sizeof(uint64_t));
        CK(cudaMalloc(&ctx->mtj_polys, flat.size() * sizeof(uint64_t)));
        CK(cudaMalloc(&ctx->mtj_scratch, (size_t)(FASTF_MT_DEG + 624 + 64) * sizeof(u32)));
        CK(cudaMemcpy(ctx->mtj_polys, flat.data(), flat.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    // seed only (no pairs): leaves the window x_0 .. x_623 in state
    FASTF_LAUNCH(fastf_mt19937_kernel, 1, FASTF_MT_THREADS, 0, s, seed, state, 1u, (u64)0, (u64)0, (u64)0, (u32 *)nullptr, (u32 *)nullptr);
    CKL("mt19937_seed");
    for (u32 k = 0; k < T.pow2.size(); k++) {
        if (!((origin >> k) & 1ull)) continue;
        FASTF_LAUNCH(fastf_mt_jump_kernel, 1, FASTF_MTJ_THREADS, 0, s, state, (const u64 *)ctx->mtj_polys + (size_t)k * FASTF_MT_POLY_WORDS, (u32 *)ctx->mtj_scratch);
        CKL("mt_jump");
    }
    return 0;
}

// Restart the job's keep-bit stream at stream index `origin`; whatever was generated before is dropped.
static int mt_jump_to(fastf_bam2db_job *job, u64 origin)
{
    fastf_ctx *ctx = job->ctx;
    Timer &tm = job->t_mt[job->mt_launches++ & 1u];
    tm.collect(&job->ms_mt);
    tm.start(ctx->mt);
    const int rc = mt_state_at(ctx, job->prm.seed, origin, job->mt_state.as<u32>(), ctx->mt);
    tm.stop(ctx->mt);
    if (rc) return rc;
    job->mt_seeded = true;
    job->mt_origin = origin;
    job->mt_pairs_done = 0;
    return 0;
}

// Keep bits for stream indices [first, first + n) generated by FASTF_MT_SEGMENTS CTAs at once: every CTA jumps to the start of
// its own segment, then runs the normal twist.  Used once the number of draws is known (fastf_bam2db_sample); replaces whatever
// the job had generated speculatively.  Returns 1 when the jump tables are unavailable (caller falls back to one sequential CTA).
#define FASTF_MT_SEGMENTS 32
static int mt_generate_parallel(fastf_bam2db_job *job, u64 first, u64 n, cudaStream_t s)
{
    fastf_ctx *ctx = job->ctx;
    const fastf_mtj::Tables &T = fastf_mtj::tables();
    if (!T.ok) return 1;
    const u64 pairs = (n + 1247) / 1248;
    const u32 K = (u32)std::min<u64>(FASTF_MT_SEGMENTS, std::max<u64>(1, pairs / 64));
    const u64 ppc = (pairs + K - 1) / K;
    if (((first + (u64)K * ppc * 1248ull) >> T.pow2.size()) != 0) return 1;
    if (!ctx->mtj_polys) {
        std::vector<uint64_t> flat(T.pow2.size() * FASTF_MT_POLY_WORDS);
        for (size_t k = 0; k < T.pow2.size(); k++) memcpy(flat.data() + k * FASTF_MT_POLY_WORDS, T.pow2[k].data(), FASTF_MT_POLY_WORDS * sizeof(uint64_t));
        CK(cudaMalloc(&ctx->mtj_polys, flat.size() * sizeof(uint64_t)));
        CK(cudaMalloc(&ctx->mtj_scratch, (size_t)(FASTF_MT_DEG + 624 + 64) * sizeof(u32)));
        CK(cudaMemcpy(ctx->mtj_polys, flat.data(), flat.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    TRY(dev_reserve(ctx, job->mt_states, (size_t)K * 624 * sizeof(u32)));
    TRY(dev_reserve(ctx, job->mt_scratch, (size_t)K * FASTF_MTJ_SCRATCH * sizeof(u32)));
    TRY(dev_reserve(ctx, job->keepbits, (size_t)K * ppc * 39 * sizeof(u32)));
    Timer &tm = job->t_mt[job->mt_launches++ & 1u];
    tm.collect(&job->ms_mt);
    tm.start(s);
    FASTF_LAUNCH(fastf_mt_jump_batch_kernel, K, FASTF_MTJ_THREADS, 0, s, job->prm.seed, (const u64 *)ctx->mtj_polys, (u32)T.pow2.size(), first, ppc * 1248ull, job->mt_states.as<u32>(),
                 job->mt_scratch.as<u32>());
    CKL("mt_jump_batch");
    FASTF_LAUNCH(fastf_mt19937_kernel, K, FASTF_MT_THREADS, 0, s, job->prm.seed, job->mt_states.as<u32>(), 0u, (u64)0, ppc, job->prm.keep_threshold, (u32 *)nullptr, job->keepbits.as<u32>());
    CKL("mt19937");
    tm.stop(s);
    job->mt_seeded = true;
    job->mt_origin = first;
    job->mt_pairs_done = (u64)K * ppc;
    return 0;
}

// Wait for the slot's counters, size the candidate array, gather the slot's staged candidates.
static int finalize_slot(fastf_bam2db_job *job, u32 si)
{
    fastf_ctx *ctx = job->ctx;
    ChunkSlot &S = job->slot[si];
    if (!S.pending) return 0;
    const double ft_w0 = job->feed_timing ? wall_seconds() : 0;
    CK(cudaEventSynchronize(S.ev_done));
    if (job->feed_timing) job->ft_wait_s += wall_seconds() - ft_w0;
    const u64 *snap = S.snap.as<u64>();
    const u64 n_records = snap[0], n_cand = snap[1];
    job->n_blocks_done += S.nblocks;
    job->status |= (u32)snap[2];
    if (job->status) {
        char buf[256];
        if (job->status == FASTF_ST_UMI_TOO_LONG)
            return ctx_fail(ctx, "bam2db: umi-too-long: a UB tag holds more than %u bases; begin the job with a larger umi_max_bytes", 4u * job->L.umi_max_bytes);
        return ctx_fail(ctx, "bam2db: malformed input in chunk ending at block %llu: %s", (unsigned long long)job->n_blocks, status_string(job->status, buf, sizeof buf));
    }
    if (n_cand > job->cand_cap) {
        u64 want = std::max<u64>(n_cand + n_cand / 2, 1u << 20);
        want = std::max<u64>(want, ctx->hint_cand / sizeof(u64));
        // the blocks already handed to the job will bring candidates at the rate seen so far
        if (job->n_blocks_done) want = std::max<u64>(want, (u64)((double)n_cand * (double)job->n_blocks_fed / (double)job->n_blocks_done * 1.03) + (1u << 16));
        TRY(dev_reserve(ctx, job->cand, want * sizeof(u64), job->n_cand * sizeof(u64), ctx->compute));
        job->cand_cap = job->cand.cap / sizeof(u64);
    }
    job->t_gather[si].collect(&job->ms_gather);
    job->t_gather[si].start(ctx->compute);
    if (S.nblocks) {
        FASTF_LAUNCH(fastf_stage_gather_kernel, (S.nblocks + 7) / 8, 256, 0, ctx->compute, (const u64 *)S.stage.as<u64>(), (const u64 *)S.idx.stage_off, (const u32 *)S.idx.ncbv,
                     (const u64 *)S.idx.dst_base, S.nblocks, job->cand.as<u64>());
        CKL("stage_gather");
    }
    job->t_gather[si].stop(ctx->compute);
    CK(cudaEventRecord(S.ev_gather, ctx->compute));   // the slot's index arrays are free for the next upload (inflate stream) after this
    CK(cudaEventRecord(job->ev_last, ctx->compute));
    job->n_records = n_records;
    job->n_cand = n_cand;
    S.pending = false;

    return 0;
}

// FASTF_BAM_STRADDLE: record-start guesses per BGZF block -> virtual blocks [v_k, v_next) for the per-block kernels (bam_straddle.cuh).
// scratch holds guess u64[nb] | virt_off u64[nb] | virt_size u32[nb]; status_word collects impossible layouts.
static int launch_virtual_blocks(fastf_ctx *ctx, DevBuf &scratch, const u8 *infl, u64 infl_bytes, const u64 *blk_off, const u32 *blk_isize, u32 nb, const u64 *hdr_off, u32 *status_word,
                                 const u64 **virt_off, const u32 **virt_size, cudaStream_t s)
{
    TRY(dev_reserve(ctx, scratch, (size_t)std::max<u32>(nb, 1) * 20 + 64));
    u64 *guess = scratch.as<u64>(), *voff = guess + nb;
    u32 *vsize = (u32 *)(voff + nb);
    if (nb) {
        FASTF_LAUNCH(fastf_bam_guess_kernel, (nb + 7) / 8, 256, 0, s, infl, infl_bytes, blk_off, blk_isize, nb, hdr_off, guess);
        CKL("bam_guess");
        FASTF_LAUNCH(fastf_bam_virtual_blocks_kernel, (nb + 255) / 256, 256, 0, s, (const u64 *)guess, nb, infl_bytes, voff, vsize, status_word);
        CKL("bam_virtual_blocks");
    }
    *virt_off = voff;
    *virt_size = vsize;
    return 0;
}

// One chunk: blocks with payload offsets relative to `comp_dev` (the caller's device buffer, or the ring entry the host bytes were
// staged into by stage_host_bytes: their H2D copies are already queued on the copy stream).
static int run_chunk(fastf_bam2db_job *job, const FastfBgzfBlock *blocks, u32 nb, const u8 *comp_dev, u64 comp_total, fastf_bam2db_job::CompRing *ring)
{
    fastf_ctx *ctx = job->ctx;
    const u32 si = job->next_slot;
    ChunkSlot &S = job->slot[si];
    if (ring) CK(cudaEventRecord(ring->ev_copy, ctx->copy));
    // the slot was used two chunks ago: its gather must have been issued (finalize) before we reuse its buffers
    TRY(finalize_slot(job, si));
    TRY(index_reserve(ctx, S.idx, nb + 1));
    const bool straddle = (job->lanes & FASTF_BAM_STRADDLE) != 0;
    u64 out_total = 0, stage_total = 0;
    for (u32 i = 0; i < nb; i++) {
        S.idx.h_in_off[i] = blocks[i].in_off;
        S.idx.h_in_len[i] = blocks[i].in_len;
        S.idx.h_isize[i] = blocks[i].isize;
        S.idx.h_out_off[i] = out_total;
        S.idx.h_stage_off[i] = stage_total;
        out_total += blocks[i].isize;
        // straddle mode: a virtual block holds the records that START between this block's guess and the next one's
        stage_total += straddle ? stage_cap_for(blocks[i].isize + (i + 1 < nb ? blocks[i + 1].isize : 0)) + 1u : stage_cap_for(blocks[i].isize);
    }
    S.idx.h_stage_off[nb] = stage_total;   // the kernels read the slice capacity as stage_off[b + 1] - stage_off[b]
    if (ring) CK(cudaStreamWaitEvent(ctx->infl, ring->ev_copy, 0));
    // S.infl / S.stage / S.idx were last used by chunk i-2, whose parse and gather have completed (finalize_slot above)
    TRY(dev_reserve(ctx, S.infl, out_total + 64));
    TRY(dev_reserve(ctx, S.stage, stage_total * sizeof(u64) + 64));
    if (!job->first_recorded) { CK(cudaEventRecord(job->ev_first, ctx->infl)); job->first_recorded = true; }
    // inflate runs on its own stream so that chunk i+1 inflates (SM kernel or hardware engine) while chunk i is parsed;
    // the gather of the slot's previous chunk (compute stream) still reads the index arrays we are about to overwrite
    CK(cudaStreamWaitEvent(ctx->infl, S.ev_gather, 0));
    TRY(index_upload(ctx, S.idx, ctx->infl));
    job->t_infl[si].collect(&job->ms_inflate);
    job->t_infl[si].start(ctx->infl);
    cudaEvent_t fti0 = nullptr, fti1 = nullptr;
    if (job->feed_timing && ring) { CK(cudaEventCreate(&fti0)); CK(cudaEventCreate(&fti1)); CK(cudaEventRecord(fti0, ctx->infl)); job->ft_host_submit.push_back(wall_seconds()); }
    TRY(launch_inflate(ctx, job->lanes, comp_dev, comp_total, S.idx.in_off, S.idx.in_len, S.idx.out_off, S.idx.isize, nb, S.infl.as<u8>(), S.idx.st_infl, ctx->infl, &S.de, S.idx.h_in_off,
                       S.idx.h_in_len, S.idx.h_out_off, S.idx.h_isize));
    if (fti0) { CK(cudaEventRecord(fti1, ctx->infl)); job->ft_infl_ev.push_back({fti0, fti1}); }
    job->t_infl[si].stop(ctx->infl);
    CK(cudaEventRecord(S.ev_infl, ctx->infl));
    CK(cudaStreamWaitEvent(ctx->compute, S.ev_infl, 0));
    job->t_crc[si].collect(&job->ms_crc);
    job->t_crc[si].start(ctx->compute);
    TRY(launch_crc(ctx, job->lanes, comp_dev, comp_total, S.idx.in_off, S.idx.in_len, S.infl.as<u8>(), S.idx.out_off, S.idx.isize, nb, S.idx.st_infl, ctx->compute));
    job->t_crc[si].stop(ctx->compute);
    if (ring) { CK(cudaEventRecord(ring->ev_free, ctx->compute)); ring->used = true; }   // compressed bytes no longer needed
    job->t_parse[si].collect(&job->ms_parse);
    job->t_parse[si].start(ctx->compute);
    if (!job->header_done) {
        FASTF_LAUNCH(fastf_bam_header_kernel, 1, 32, 0, ctx->compute, (const u8 *)S.infl.as<u8>(), out_total, job->hdr_off.as<u64>(), (u32 *)(job->counters.as<u64>() + 2));
        CKL("bam_header");
        job->header_done = true;
    } else {
        CK(cudaMemsetAsync(job->hdr_off.p, 0, sizeof(u64), ctx->compute));
    }
    const u64 *p_off = S.idx.out_off;
    const u32 *p_size = S.idx.isize;
    if (straddle) TRY(launch_virtual_blocks(ctx, S.virt, (const u8 *)S.infl.as<u8>(), out_total, S.idx.out_off, S.idx.isize, nb, job->hdr_off.as<u64>(), (u32 *)(job->counters.as<u64>() + 2), &p_off, &p_size,
                                            ctx->compute));
    if (nb) {
        FASTF_LAUNCH(fastf_bam_parse_kernel, (nb + FASTF_PARSE_WARPS - 1) / FASTF_PARSE_WARPS, FASTF_PARSE_WARPS * 32, 0, ctx->compute, (const u8 *)S.infl.as<u8>(), (u64)((out_total + 15) & ~15ull),
                     p_off, p_size, nb, (const u64 *)job->hdr_off.as<u64>(), job->cells.view, job->genes.view, job->L, (const u64 *)S.idx.stage_off,
                     S.stage.as<u64>(), S.idx.nrec, S.idx.ncbv, S.idx.st_parse);
        CKL("bam_parse");
    }
    FASTF_LAUNCH(fastf_chunk_counts_kernel, 1, FASTF_SCAN_THREADS, 0, ctx->compute, (const u32 *)S.idx.nrec, (const u32 *)S.idx.ncbv, (const u32 *)S.idx.st_infl, (const u32 *)S.idx.st_parse, nb,
                 S.idx.dst_base, job->counters.as<u64>());
    CKL("chunk_counts");
    job->t_parse[si].stop(ctx->compute);
    CK(cudaMemcpyAsync(S.snap.p, job->counters.p, 4 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaEventRecord(S.ev_done, ctx->compute));
    if (job->feed_timing && ring) { cudaEvent_t e = nullptr; CK(cudaEventCreate(&e)); CK(cudaEventRecord(e, ctx->compute)); job->ft_parse_ev.push_back(e); }
    S.nblocks = nb;
    S.pending = true;
    job->n_blocks += nb;
    job->n_chunks++;
    job->infl_bytes += out_total;
    job->next_slot ^= 1u;
    // The previous chunk (slot si ^ 1) is gathered when its slot comes round again (top of the next run_chunk) or at the end of
    // the job: waiting for it here would keep the host from queueing more than one chunk ahead of the device.
    return 0;
}

static int submit_pending(fastf_bam2db_job *job)
{
    fastf_bam2db_job::PendingChunk &P = job->pending;
    if (P.blocks.empty()) return 0;
    const u8 *comp = P.ring ? P.ring->buf.as<u8>() : P.comp_dev;
    const u64 total = P.ring ? P.fill : P.comp_total;
    int rc = run_chunk(job, P.blocks.data(), (u32)P.blocks.size(), comp, total, P.ring);
    P.blocks.clear();
    P.infl = 0; P.comp_dev = nullptr; P.comp_total = 0; P.ring = nullptr; P.fill = 0;
    return rc;
}

// Queue the H2D copy of host bytes [lo, hi) behind what the pending chunk has staged so far; returns their offset in the ring entry.
static int stage_host_bytes(fastf_bam2db_job *job, const u8 *host_base, u64 lo, u64 hi, u64 *at)
{
    fastf_ctx *ctx = job->ctx;
    fastf_bam2db_job::PendingChunk &P = job->pending;
    if (!P.ring) {
        // The copy goes out before the host waits for anything.  The ring entry was last read by the chunk three back (inflate +
        // CRC), which the copy stream waits for on the device.
        P.ring = &job->comp_ring[job->comp_seq++ % 3u];
        P.fill = 0;
        if (P.ring->used) CK(cudaStreamWaitEvent(ctx->copy, P.ring->ev_free, 0));
    }
    const u64 bytes = hi - lo, padded = (bytes + 3) & ~3ull;
    // growing moves the buffer: the bytes staged so far travel along (dev_reserve drains the streams before it lets go of the old one),
    // so an entry starts at the size a whole chunk is expected to need
    u64 want = P.fill + padded + 16;
    if (want > P.ring->buf.cap) want = std::max<u64>(want, std::max<u64>(ctx->hint_ring, job->ring_estimate));
    TRY(dev_reserve(ctx, P.ring->buf, want, P.fill, ctx->copy));
    cudaEvent_t ft0 = nullptr, ft1 = nullptr;
    if (job->feed_timing) { CK(cudaEventCreate(&ft0)); CK(cudaEventCreate(&ft1)); CK(cudaEventRecord(ft0, ctx->copy)); }
    CK(cudaMemcpyAsync(P.ring->buf.as<u8>() + P.fill, host_base + lo, bytes, cudaMemcpyHostToDevice, ctx->copy));
    if (job->feed_timing) { CK(cudaEventRecord(ft1, ctx->copy)); job->ft_copy_ev.push_back({ft0, ft1}); job->ft_copy_bytes += bytes; }
    *at = P.fill;
    P.fill += padded;
    return 0;
}

// Append indexed blocks (payload offsets relative to comp_dev, or to host_base for a host feed) to the pending chunk and launch
// every chunk that fills up.  What is left waits for the next feed or for drain_chunks.
static int run_blocks(fastf_bam2db_job *job, const std::vector<FastfBgzfBlock> &blocks, const u8 *comp_dev, u64 comp_total, const u8 *host_base)
{
    fastf_bam2db_job::PendingChunk &P = job->pending;
    const size_t n = blocks.size();
    size_t i = 0;
    job->n_blocks_fed += n;
    // compressed bytes a full chunk of this feed will stage (ring entries are sized once, see stage_host_bytes): never more than the feed holds
    if (host_base && n) {
        u64 isz = 0;
        for (size_t k = 0; k < n; k++) isz += blocks[k].isize;
        const double span = (double)(blocks[n - 1].in_off + blocks[n - 1].in_len + 8 - blocks[0].in_off);
        const double per_chunk = std::min<double>((double)job->chunk_blocks, (double)job->chunk_bytes / std::max<double>((double)isz / (double)n, 1.0));
        job->ring_estimate = (u64)std::min<double>(span, span / (double)n * 1.03 * per_chunk) + (1u << 20);
    }
    while (i < n) {
        // a chunk reads its compressed bytes from one buffer: blocks of another device buffer (or of the other kind of feed) start a new one
        if (!P.blocks.empty() && (host_base ? P.ring == nullptr : (P.ring != nullptr || P.comp_dev != comp_dev))) TRY(submit_pending(job));
        size_t j = i;
        u64 infl = P.infl;
        const u64 byte0 = blocks[i].in_off & ~3ull;
        while (j < n && P.blocks.size() + (j - i) < job->chunk_blocks && (P.blocks.empty() && j == i ? true : infl + blocks[j].isize <= job->chunk_bytes) &&
               (!host_base || j == i || P.fill + (blocks[j].in_off + blocks[j].in_len + 8 - byte0) <= job->chunk_bytes + (1u << 20))) {
            infl += blocks[j].isize;
            j++;
        }
        if (j == i) { TRY(submit_pending(job)); continue; }   // the pending chunk is full
        if (host_base) {
            // copy [first payload rounded down to 4, end of the last block's CRC32 / ISIZE trailer) and rebase the offsets
            const u64 hi = blocks[j - 1].in_off + blocks[j - 1].in_len + 8;
            u64 at = 0;
            TRY(stage_host_bytes(job, host_base, byte0, hi, &at));
            for (size_t k = i; k < j; k++) { FastfBgzfBlock b = blocks[k]; b.in_off = at + (b.in_off - byte0); P.blocks.push_back(b); }
        } else {
            P.comp_dev = comp_dev;
            P.comp_total = comp_total;
            P.blocks.insert(P.blocks.end(), blocks.begin() + i, blocks.begin() + j);
        }
        P.infl = infl;
        i = j;
        if (P.blocks.size() >= job->chunk_blocks || i < n) TRY(submit_pending(job));   // full, or the next block did not fit
    }
    return 0;
}

extern "C" int fastf_bam2db_feed(fastf_bam2db_job *job, const void *host_bytes, size_t n)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    if (job->sampled_done) return ctx_fail(ctx, "bam2db_feed: job already sampled");
    struct FeedClock { fastf_bam2db_job *j; double t0; ~FeedClock() { if (j->feed_timing) j->ft_feed_s += wall_seconds() - t0; } } feed_clock{job, job->feed_timing ? wall_seconds() : 0};
    const u8 *p = (const u8 *)host_bytes;
    job->comp_bytes += n;
    std::vector<FastfBgzfBlock> blocks;
    // 1. complete a partial block left over from the previous call
    while (!job->carry.empty() && n) {
        std::vector<u8> &c = job->carry;
        size_t want = 18;
        if (c.size() >= 18) {
            blocks.clear();
            size_t used = 0;
            int rc = fastf_bgzf_index(c.data(), c.size(), 0, blocks, &used);
            if (rc != FASTF_BGZF_OK && rc != FASTF_BGZF_NEED_MORE) return ctx_fail(ctx, "bam2db_feed: not a BGZF block (index error %d)", rc);
            if (!blocks.empty()) {
                // run the completed block(s) out of the carry buffer (pageable copy; rare path).  The copy must finish
                // before the carry buffer is edited, so drain the copy stream here.
                TRY(run_blocks(job, blocks, nullptr, 0, c.data()));
                CK(cudaStreamSynchronize(ctx->copy));
                c.erase(c.begin(), c.begin() + used);
                continue;
            }
            // header visible: total block size = BSIZE+1 (read it the same way the indexer does)
            u32 xlen = (u32)c[10] | ((u32)c[11] << 8);
            want = 12 + (size_t)xlen;
            if (c.size() >= want) {
                u32 bsize = 0;
                for (u32 x = 0; x + 4 <= xlen;) {
                    const u8 *sf = c.data() + 12 + x;
                    u32 slen = (u32)sf[2] | ((u32)sf[3] << 8);
                    if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = ((u32)sf[4] | ((u32)sf[5] << 8)) + 1;
                    x += 4 + slen;
                }
                if (!bsize) return ctx_fail(ctx, "bam2db_feed: gzip member without a BGZF BC field");
                want = bsize;
            }
        }
        size_t take = std::min(n, want > c.size() ? want - c.size() : (size_t)1);
        c.insert(c.end(), p, p + take);
        p += take;
        n -= take;
    }
    if (!job->carry.empty()) {
        // n == 0: maybe the carry became a whole block exactly
        blocks.clear();
        size_t used = 0;
        int rc = fastf_bgzf_index(job->carry.data(), job->carry.size(), 0, blocks, &used);
        if (rc != FASTF_BGZF_OK && rc != FASTF_BGZF_NEED_MORE) return ctx_fail(ctx, "bam2db_feed: not a BGZF block (index error %d)", rc);
        if (!blocks.empty()) {
            TRY(run_blocks(job, blocks, nullptr, 0, job->carry.data()));
            CK(cudaStreamSynchronize(ctx->copy));
            job->carry.erase(job->carry.begin(), job->carry.begin() + used);
        }
        return 0;
    }
    if (!n) return 0;
    // 2. whole blocks straight out of the caller's buffer
    // (one chunk's worth of block headers at a time: the device starts on chunk i while the host walks the headers of chunk i+1)
    size_t used = 0;
    for (;;) {
        blocks.clear();
        size_t step = 0;
        int rc = fastf_bgzf_index(p + used, n - used, used, blocks, &step, job->chunk_bytes, (size_t)job->chunk_blocks);
        if (rc != FASTF_BGZF_OK && rc != FASTF_BGZF_NEED_MORE && rc != FASTF_BGZF_LIMIT) return ctx_fail(ctx, "bam2db_feed: not a BGZF stream at byte %zu (index error %d)", used + step, rc);
        used += step;
        if (!blocks.empty()) TRY(run_blocks(job, blocks, nullptr, 0, p));
        if (rc != FASTF_BGZF_LIMIT) break;
    }
    // 3. keep the tail.  The caller may reuse its buffer after we return: wait for the copies.
    if (used < n) job->carry.assign(p + used, p + n);
    CK(cudaStreamSynchronize(ctx->copy));
    return 0;
}

extern "C" int fastf_bam2db_feed_device(fastf_bam2db_job *job, const void *dev_bytes, size_t nbytes, const uint64_t *in_off, const uint32_t *in_len, const uint32_t *isize, uint64_t nblocks)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    if (job->sampled_done) return ctx_fail(ctx, "bam2db_feed_device: job already sampled");
    if (((uintptr_t)dev_bytes & 3u) != 0) return ctx_fail(ctx, "bam2db_feed_device: dev_bytes must be 4-byte aligned");
    std::vector<FastfBgzfBlock> blocks(nblocks);
    for (u64 i = 0; i < nblocks; i++) {
        if (in_off[i] + in_len[i] + 8 > nbytes || isize[i] > 65536) return ctx_fail(ctx, "bam2db_feed_device: block %llu outside the buffer", (unsigned long long)i);
        blocks[i].in_off = in_off[i]; blocks[i].in_len = in_len[i]; blocks[i].isize = isize[i]; blocks[i].crc32 = 0;
    }
    job->comp_bytes += nbytes;
    return run_blocks(job, blocks, (const u8 *)dev_bytes, nbytes & ~(u64)3, nullptr);
}

static int drain_chunks(fastf_bam2db_job *job)
{
    fastf_ctx *ctx = job->ctx;
    if (!job->carry.empty()) return ctx_fail(ctx, "bam2db: input ends inside a BGZF block (%zu trailing bytes)", job->carry.size());
    TRY(submit_pending(job));
    TRY(finalize_slot(job, job->next_slot));        // older one first (file order of the gathers does not matter, bases are absolute)
    TRY(finalize_slot(job, job->next_slot ^ 1u));
    return 0;
}

extern "C" int fastf_bam2db_counts(fastf_bam2db_job *job, uint64_t *n_records, uint64_t *n_cb_valid)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    TRY(drain_chunks(job));
    if (n_records) *n_records = job->n_records;
    if (n_cb_valid) *n_cb_valid = job->n_cand;
    return 0;
}

extern "C" int fastf_bam2db_sample(fastf_bam2db_job *job, uint64_t ordinal_base)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    if (job->sampled_done) return ctx_fail(ctx, "bam2db_sample: already sampled");
    TRY(drain_chunks(job));
    const u64 n = job->n_cand;
    const u64 first_draw = job->prm.d0 + ordinal_base;
    job->n_sampled = job->n_valid = 0;
    if (n >= 0xffffffffull) return ctx_fail(ctx, "bam2db_sample: %llu CB-valid reads exceed the 2^32-1 limit of one device; shard over more GPUs", (unsigned long long)n);
    if (n) {
        // the number of draws is known now: generate exactly the keep bits [first_draw, first_draw + n) with 32 CTAs that each
        // jump (GF(2) jump-ahead) to their own segment of the reference's single stream.  Fallback without the jump tables: one
        // CTA generates the stream sequentially from the seed.
        if (mt_generate_parallel(job, first_draw, n, ctx->mt)) TRY(mt_extend(job, first_draw + n));
        CK(cudaEventRecord(job->ev_mt, ctx->mt));
        CK(cudaStreamWaitEvent(ctx->compute, job->ev_mt, 0));
        const u32 ntiles = (u32)((n + FASTF_SAMPLE_TILE - 1) / FASTF_SAMPLE_TILE);
        TRY(dev_reserve(ctx, job->tile_valid, (size_t)ntiles * sizeof(u32)));
        TRY(dev_reserve(ctx, job->tile_tot, sizeof(u32)));
        CK(cudaMemsetAsync(job->sample_counters.p, 0, 2 * sizeof(u64), ctx->compute));
        job->t_sample.start(ctx->compute);
        FASTF_LAUNCH(fastf_sample_count_kernel, ntiles, FASTF_SAMPLE_THREADS, 0, ctx->compute, (const u64 *)job->cand.as<u64>(), n, (const u32 *)job->keepbits.as<u32>(), first_draw - job->mt_origin,
                     job->tile_valid.as<u32>(), job->sample_counters.as<u64>());
        CKL("sample_count");
        TRY(launch_scan_rows(ctx, job->tile_valid.as<u32>(), ntiles, 1, job->tile_tot.as<u32>(), ctx->compute));
        CK(cudaMemcpyAsync(job->small_host.p, job->sample_counters.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));
        job->n_sampled = job->small_host.as<u64>()[0];
        job->n_valid = job->small_host.as<u64>()[1];
        TRY(dev_reserve(ctx, job->kept, std::max<u64>(job->n_valid, 1) * sizeof(u64)));
        FASTF_LAUNCH(fastf_sample_scatter_kernel, ntiles, FASTF_SAMPLE_THREADS, 0, ctx->compute, (const u64 *)job->cand.as<u64>(), n, (const u32 *)job->keepbits.as<u32>(), first_draw - job->mt_origin,
                     (const u32 *)job->tile_valid.as<u32>(), job->kept.as<u64>());
        CKL("sample_scatter");
        job->t_sample.stop(ctx->compute);
        CK(cudaEventRecord(job->ev_last, ctx->compute));
    }
    job->sampled_done = true;
    return 0;
}

extern "C" int fastf_bam2db_kept_device(fastf_bam2db_job *job, uint64_t **dev_keys, uint64_t *n)
{
    fastf_ctx *ctx = job->ctx;
    if (!job->sampled_done) return ctx_fail(ctx, "bam2db_kept_device: call fastf_bam2db_sample first");
    *dev_keys = job->kept.as<u64>();
    *n = job->n_valid;
    return 0;
}

extern "C" int fastf_bam2db_sample_counts(fastf_bam2db_job *job, uint64_t *sampled, uint64_t *valid)
{
    fastf_ctx *ctx = job->ctx;
    if (!job->sampled_done) return ctx_fail(ctx, "bam2db_sample_counts: call fastf_bam2db_sample first");
    if (sampled) *sampled = job->n_sampled;
    if (valid) *valid = job->n_valid;
    return 0;
}

extern "C" int fastf_bam2db_key_layout(fastf_bam2db_job *job, uint32_t *bits_cell, uint32_t *bits_gene, uint32_t *bits_umi)
{
    *bits_cell = job->L.bits_cell; *bits_gene = job->L.bits_gene; *bits_umi = job->L.bits_umi;
    return 0;
}

// sorted-unaware front half shared by finish and the device-level entry points: figure out which bits vary
static int varying_bits(fastf_ctx *ctx, DevBuf &orand, PinBuf &host, const u64 *keys, u64 n, u64 *varying, cudaStream_t s)
{
    u64 init[2] = {0ull, ~0ull};
    CK(cudaMemcpyAsync(orand.p, init, sizeof init, cudaMemcpyHostToDevice, s));
    u32 grid = (u32)std::min<u64>((n + 255) / 256, 148 * 8);
    FASTF_LAUNCH(fastf_key_bits_kernel, grid ? grid : 1, 256, 0, s, keys, n, orand.as<u64>());
    CKL("key_bits");
    CK(cudaMemcpyAsync(host.p, orand.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *varying = host.as<u64>()[0] ^ host.as<u64>()[1];
    return 0;
}

static int coo_to_host(fastf_ctx *ctx, RleScratch &R, u64 nnz, u32 **m_gene, u32 **m_cell, u32 **m_count, cudaStream_t s)
{
    *m_gene = (u32 *)malloc(std::max<u64>(nnz, 1) * sizeof(u32));
    *m_cell = (u32 *)malloc(std::max<u64>(nnz, 1) * sizeof(u32));
    *m_count = (u32 *)malloc(std::max<u64>(nnz, 1) * sizeof(u32));
    if (!*m_gene || !*m_cell || !*m_count) return ctx_fail(ctx, "out of host memory for %llu COO rows", (unsigned long long)nnz);
    if (nnz) {
        TRY(d2h_pageable(ctx, *m_gene, R.out_gene.p, nnz * sizeof(u32), s));
        TRY(d2h_pageable(ctx, *m_cell, R.out_cell.p, nnz * sizeof(u32), s));
        TRY(d2h_pageable(ctx, *m_count, R.count.p, nnz * sizeof(u32), s));
        CK(cudaStreamSynchronize(s));
    }
    return 0;
}

// counters, sizes and the per-stage CUDA-event clocks of a job (no result arrays)
static int fill_stats(fastf_bam2db_job *job, fastf_bam2db_result *res)
{
    fastf_ctx *ctx = job->ctx;
    const FastfKeyLayout &L = job->L;
    res->total = job->n_records;
    res->cb_valid = job->n_cand;
    res->sampled = job->n_sampled;
    res->valid = job->n_valid;
    res->bits_cell = L.bits_cell; res->bits_gene = L.bits_gene; res->bits_umi = L.bits_umi; res->umi_max_bytes = L.umi_max_bytes;
    res->n_blocks = job->n_blocks; res->compressed_bytes = job->comp_bytes; res->inflated_bytes = job->infl_bytes;
    res->status = job->status;
    for (int i = 0; i < 2; i++) { job->t_infl[i].collect(&job->ms_inflate); job->t_crc[i].collect(&job->ms_crc); job->t_parse[i].collect(&job->ms_parse); job->t_gather[i].collect(&job->ms_gather); }
    job->t_mt[0].collect(&job->ms_mt); job->t_mt[1].collect(&job->ms_mt); job->t_sample.collect(&job->ms_sample); job->t_sort.collect(&job->ms_sort); job->t_count.collect(&job->ms_count);
    res->ms_inflate = job->ms_inflate; res->ms_crc = job->ms_crc; res->ms_parse = job->ms_parse; res->ms_gather = job->ms_gather; res->ms_mt = job->ms_mt;
    res->ms_sample = job->ms_sample; res->ms_sort = job->ms_sort; res->ms_count = job->ms_count;
    if (job->first_recorded) {
        CK(cudaEventRecord(job->ev_last, ctx->compute));
        CK(cudaEventSynchronize(job->ev_last));
        CK(cudaEventElapsedTime(&res->ms_device_total, job->ev_first, job->ev_last));
    }
    res->n_launches = ctx->launches - job->launches0;
    res->n_chunks = job->n_chunks;
    return 0;
}

extern "C" int fastf_bam2db_stats(fastf_bam2db_job *job, fastf_bam2db_result *res)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    CK(cudaStreamSynchronize(ctx->compute));
    CK(cudaStreamSynchronize(ctx->mt));
    return fill_stats(job, res);
}

extern "C" int fastf_bam2db_finish(fastf_bam2db_job *job, fastf_bam2db_result *res)
{
    fastf_ctx *ctx = job->ctx;
    CK(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    if (!job->sampled_done) TRY(fastf_bam2db_sample(job, 0));
    if (job->feed_timing && !job->ft_copy_ev.empty()) {
        CK(cudaStreamSynchronize(ctx->copy));
        double ms = 0, span = 0;
        for (auto &e : job->ft_copy_ev) { float t = 0; cudaEventElapsedTime(&t, e.first, e.second); ms += t; }
        { float t = 0; cudaEventElapsedTime(&t, job->ft_copy_ev.front().first, job->ft_copy_ev.back().second); span = t; }
        fprintf(stderr, "[fastf feed timing] H2D %zu copies, %.2f GB in %.1f ms of copy time = %.1f GB/s (first start to last end %.1f ms); host: %.1f ms inside feed calls, of which %.1f ms waiting for chunks to finish\n",
                job->ft_copy_ev.size(), job->ft_copy_bytes / 1e9, ms, job->ft_copy_bytes / 1e6 / std::max(ms, 1e-3), span, 1e3 * job->ft_feed_s, 1e3 * job->ft_wait_s);
        CK(cudaDeviceSynchronize());
        {
            // per chunk, ms since the first copy started: inflate start / end, parse end, and when the host submitted it
            cudaEvent_t o = job->ft_copy_ev.front().first;
            fprintf(stderr, "[fastf feed timing] chunk: host-submit | inflate start..end | parse end   (ms since the first H2D started; copies: start..end)\n");
            for (size_t k = 0; k < job->ft_infl_ev.size(); k++) {
                float a = 0, b = 0, c = 0;
                cudaEventElapsedTime(&a, o, job->ft_infl_ev[k].first); cudaEventElapsedTime(&b, o, job->ft_infl_ev[k].second);
                if (k < job->ft_parse_ev.size()) cudaEventElapsedTime(&c, o, job->ft_parse_ev[k]);
                fprintf(stderr, "[fastf feed timing]   %2zu: host %.1f | %.1f..%.1f | %.1f\n", k, 1e3 * (job->ft_host_submit[k] - job->ft_host_submit[0]), a, b, c);
            }
            for (size_t k = 0; k < job->ft_copy_ev.size(); k++) {
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, o, job->ft_copy_ev[k].first); cudaEventElapsedTime(&b, o, job->ft_copy_ev[k].second);
                fprintf(stderr, "[fastf feed timing]   copy %2zu: %.1f..%.1f\n", k, a, b);
            }
        }
        for (auto &e : job->ft_copy_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto &e : job->ft_infl_ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto &e : job->ft_parse_ev) cudaEventDestroy(e);
        job->ft_copy_ev.clear(); job->ft_infl_ev.clear(); job->ft_parse_ev.clear();
    }
    const u64 n = job->n_valid;
    const FastfKeyLayout &L = job->L;
    if (job->prm.want_rows) {
        res->row_keys = (u64 *)malloc(std::max<u64>(n, 1) * sizeof(u64));
        if (!res->row_keys) return ctx_fail(ctx, "out of host memory for %llu rows", (unsigned long long)n);
        if (n) TRY(d2h_pageable(ctx, res->row_keys, job->kept.p, n * sizeof(u64), ctx->compute));
        res->n_rows = n;
    }
    u64 nnz = 0;
    if (n) {
        u64 varying = 0;
        job->t_sort.start(ctx->compute);
        TRY(varying_bits(ctx, job->orand, job->small_host, job->kept.as<u64>(), n, &varying, ctx->compute));
        u32 shifts[8];
        const int npass = plan_windows(varying, shifts);
        bool in_alt = false;
        // the candidate array is dead after sampling and at least as large as `kept`: reuse it as the ping-pong buffer
        TRY(sort_keys(ctx, job->sortS, job->kept.as<u64>(), job->cand.as<u64>(), nullptr, nullptr, n, shifts, npass, &in_alt, ctx->compute));
        job->t_sort.stop(ctx->compute);
        const u64 *sorted = in_alt ? job->cand.as<u64>() : job->kept.as<u64>();
        job->t_count.start(ctx->compute);
        TRY(rle_groups(ctx, job->rleS, sorted, nullptr, n, L.bits_umi, L.bits_umi - 1, L.bits_gene, &nnz, nullptr, ctx->compute));
        job->t_count.stop(ctx->compute);
        CK(cudaEventRecord(job->ev_last, ctx->compute));
    }
    TRY(coo_to_host(ctx, job->rleS, nnz, &res->m_gene, &res->m_cell, &res->m_count, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    CK(cudaStreamSynchronize(ctx->mt));
    TRY(fill_stats(job, res));
    res->nnz = nnz;
    return 0;
}

extern "C" void fastf_bam2db_result_free(fastf_bam2db_result *res)
{
    if (!res) return;
    free(res->m_gene); free(res->m_cell); free(res->m_count); free(res->row_keys);
    res->m_gene = res->m_cell = res->m_count = nullptr;
    res->row_keys = nullptr;
}

// ---------------------------------------------------------------------------------------------------
// device-level building blocks
// ---------------------------------------------------------------------------------------------------
extern "C" int fastf_sort_u64_device(fastf_ctx *ctx, uint64_t *dev_keys, uint32_t *dev_vals, uint64_t n, uint32_t key_bits)
{
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    SortScratch S;
    DevBuf alt, valt;
    int rc = dev_reserve(ctx, alt, n * sizeof(u64));
    if (!rc && dev_vals) rc = dev_reserve(ctx, valt, n * sizeof(u32));
    u32 shifts[8];
    int npass = 0;
    for (u32 b = 0; b < key_bits && npass < 8; b += 8) shifts[npass++] = b;
    bool in_alt = false;
    if (!rc) rc = sort_keys(ctx, S, dev_keys, alt.as<u64>(), dev_vals, valt.as<u32>(), n, shifts, npass, &in_alt, ctx->compute);
    if (!rc && in_alt) {
        rc = cudaMemcpyAsync(dev_keys, alt.p, n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->compute) != cudaSuccess;
        if (!rc && dev_vals) rc = cudaMemcpyAsync(dev_vals, valt.p, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->compute) != cudaSuccess;
    }
    if (cudaStreamSynchronize(ctx->compute) != cudaSuccess && !rc) rc = ctx_fail(ctx, "sort_u64_device: stream error");
    sort_scratch_release(ctx, S);
    dev_release(ctx, alt);
    dev_release(ctx, valt);
    return rc;
}

// same, results left on the device in caller-provided arrays of capacity >= n (multi-GPU driver: the pieces travel over NCCL)
extern "C" int fastf_dedup_count_device_out(fastf_ctx *ctx, const uint64_t *dev_sorted_keys, uint64_t n, uint32_t bits_gene, uint32_t bits_umi, uint64_t *nnz, uint32_t *dev_gene, uint32_t *dev_cell,
                                            uint32_t *dev_count)
{
    CK(cudaSetDevice(ctx->device));
    RleScratch R;
    u64 ng = 0;
    int rc = rle_groups(ctx, R, dev_sorted_keys, nullptr, n, bits_umi, bits_umi - 1, bits_gene, &ng, nullptr, ctx->compute);
    if (!rc && ng) {
        rc = cudaMemcpyAsync(dev_gene, R.out_gene.p, ng * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->compute) != cudaSuccess ||
             cudaMemcpyAsync(dev_cell, R.out_cell.p, ng * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->compute) != cudaSuccess ||
             cudaMemcpyAsync(dev_count, R.count.p, ng * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->compute) != cudaSuccess;
        if (rc) ctx_fail(ctx, "dedup_count_device_out: copy failed");
    }
    if (cudaStreamSynchronize(ctx->compute) != cudaSuccess && !rc) rc = ctx_fail(ctx, "dedup_count_device_out: stream error");
    *nnz = ng;
    rle_scratch_release(ctx, R);
    return rc;
}

extern "C" int fastf_dedup_count_device(fastf_ctx *ctx, const uint64_t *dev_sorted_keys, uint64_t n, uint32_t bits_gene, uint32_t bits_umi, uint64_t *nnz, uint32_t **m_gene, uint32_t **m_cell,
                                        uint32_t **m_count)
{
    CK(cudaSetDevice(ctx->device));
    RleScratch R;
    u64 ng = 0;
    int rc = rle_groups(ctx, R, dev_sorted_keys, nullptr, n, bits_umi, bits_umi - 1, bits_gene, &ng, nullptr, ctx->compute);
    if (!rc) rc = coo_to_host(ctx, R, ng, m_gene, m_cell, m_count, ctx->compute);
    *nnz = ng;
    rle_scratch_release(ctx, R);
    return rc;
}

// Destination of a cell for the multi-GPU exchange: all keys of one (cell, gene) group must meet on one rank.  The "hash" is
// order preserving -- an equal-width range partition of the 1-based cell index, which is itself the position of the barcode in a
// file-ordered random sample, so depth is spread evenly -- and therefore the ranks' (cell, gene)-sorted COO pieces concatenate
// in rank order without a merge.
static inline __host__ __device__ u32 fastf_cell_dest(u32 cell, u32 n_cells, u32 nparts) { return (u32)(((u64)(cell - 1u) * nparts) / (n_cells ? n_cells : 1u)); }

__global__ void __launch_bounds__(256) fastf_tag_dest_kernel(u64 *__restrict__ keys, u64 n, u32 cell_shift, u32 key_bits, u32 n_cells, u32 nparts)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 k = keys[i];
    u32 d = fastf_cell_dest((u32)(k >> cell_shift), n_cells, nparts);
    keys[i] = k | ((u64)(d < nparts ? d : nparts - 1u) << key_bits);
}
// heads of runs of equal keys -> compacted, with the destination tag stripped; part boundaries by binary search
__global__ void __launch_bounds__(256) fastf_part_bounds_kernel(const u64 *__restrict__ keys, u64 n, u32 key_bits, u32 nparts, u64 *__restrict__ bounds)
{
    u32 p = threadIdx.x;
    if (p > nparts) return;
    u64 lo = 0, hi = n;
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if ((keys[mid] >> key_bits) < (u64)p) lo = mid + 1; else hi = mid; }
    bounds[p] = lo;
}
__global__ void __launch_bounds__(256) fastf_strip_tag_kernel(u64 *__restrict__ keys, u64 n, u32 key_bits)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] &= (1ull << key_bits) - 1ull;
}

extern "C" int fastf_unique_partition_device(fastf_ctx *ctx, uint64_t *dev_keys, uint64_t n, uint32_t key_bits, uint32_t bits_gene, uint32_t bits_umi, uint32_t n_cells, uint32_t nparts,
                                             uint64_t *dev_out_keys, uint64_t *part_counts)
{
    CK(cudaSetDevice(ctx->device));
    for (u32 p = 0; p < nparts; p++) part_counts[p] = 0;
    if (n == 0) return 0;
    if (nparts == 0 || nparts > 256 || key_bits + (nparts > 1 ? bits_for(nparts - 1) : 0) > 64)
        return ctx_fail(ctx, "unique_partition: %u key bits leave no room for the destination tag of %u parts", key_bits, nparts);
    cudaStream_t s = ctx->compute;
    SortScratch S;
    RleScratch R;
    DevBuf alt, orand, bounds;
    PinBuf host;
    int rc = 0;
    auto body = [&]() -> int {
        TRY(dev_reserve(ctx, alt, n * sizeof(u64)));
        TRY(dev_reserve(ctx, orand, 2 * sizeof(u64)));
        TRY(dev_reserve(ctx, bounds, 257 * sizeof(u64)));
        TRY(pin_reserve(ctx, host, 257 * sizeof(u64)));
        FASTF_LAUNCH(fastf_tag_dest_kernel, (u32)((n + 255) / 256), 256, 0, s, dev_keys, n, bits_gene + bits_umi, key_bits, n_cells, nparts);
        CKL("tag_dest");
        u64 varying = 0;
        TRY(varying_bits(ctx, orand, host, dev_keys, n, &varying, s));
        u32 shifts[8];
        const int npass = plan_windows(varying, shifts);
        bool in_alt = false;
        TRY(sort_keys(ctx, S, dev_keys, alt.as<u64>(), nullptr, nullptr, n, shifts, npass, &in_alt, s));
        const u64 *sorted = in_alt ? alt.as<u64>() : dev_keys;
        // unique: every key is its own group (group_shift 0); grp_key = the distinct keys in sorted order
        u64 nuniq = 0;
        TRY(rle_groups(ctx, R, sorted, nullptr, n, 0, 64, 0, &nuniq, nullptr, s));
        FASTF_LAUNCH(fastf_part_bounds_kernel, 1, 256, 0, s, (const u64 *)R.grp_key.as<u64>(), nuniq, key_bits, nparts, bounds.as<u64>());
        CKL("part_bounds");
        CK(cudaMemcpyAsync(dev_out_keys, R.grp_key.p, nuniq * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        FASTF_LAUNCH(fastf_strip_tag_kernel, (u32)((nuniq + 255) / 256), 256, 0, s, dev_out_keys, nuniq, key_bits);
        CKL("strip_tag");
        CK(cudaMemcpyAsync(host.p, bounds.p, (nparts + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (u32 p = 0; p < nparts; p++) part_counts[p] = host.as<u64>()[p + 1] - host.as<u64>()[p];
        return 0;
    };
    rc = body();
    cudaStreamSynchronize(s);
    sort_scratch_release(ctx, S);
    rle_scratch_release(ctx, R);
    dev_release(ctx, alt); dev_release(ctx, orand); dev_release(ctx, bounds);
    pin_release(ctx, host);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer wrappers around single kernels (tests, smoke)
// ---------------------------------------------------------------------------------------------------
struct InflatedFile {
    DevBuf comp, infl;
    DeScratch de;
    BlockIndexDev idx;
    u64 n_blocks = 0, infl_bytes = 0;
    u32 status = 0;
};
static void inflated_release(fastf_ctx *ctx, InflatedFile &F) { dev_release(ctx, F.comp); dev_release(ctx, F.infl); dev_release(ctx, F.de.counter); dev_release(ctx, F.de.sorted); index_release(ctx, F.idx); }

// Inflate a whole BGZF image (host bytes, or device bytes + host index) into F.infl in one launch.
// out_prefix: bytes kept free (and preserved across calls) in front of the inflated blocks -- streamed text carries the tail of the previous chunk there.
static int inflate_whole(fastf_ctx *ctx, InflatedFile &F, const void *host_bytes, size_t n, const u8 *dev_bytes, const std::vector<FastfBgzfBlock> *pre, u32 lanes, float *ms, cudaStream_t s, u64 out_prefix = 0)
{
    std::vector<FastfBgzfBlock> local;
    const std::vector<FastfBgzfBlock> *blocks = pre;
    if (!pre) {
        size_t used = 0;
        int rc = fastf_bgzf_index((const u8 *)host_bytes, n, 0, local, &used);
        if (rc != FASTF_BGZF_OK) return ctx_fail(ctx, "inflate: not a whole BGZF stream (index error %d at byte %zu of %zu)", rc, used, n);
        blocks = &local;
    }
    const size_t nb = blocks->size();
    if (nb >= 0xffffffffull) return ctx_fail(ctx, "inflate: too many blocks");
    TRY(index_reserve(ctx, F.idx, (u32)std::max<size_t>(nb, 1)));
    // host bytes: only the span of these blocks travels (a pre-indexed subset = one chunk of a larger file), offsets are rebased
    const u64 lo = (!dev_bytes && nb) ? ((*blocks)[0].in_off & ~3ull) : 0;
    const u64 hi = (!dev_bytes && nb) ? std::min<u64>((*blocks)[nb - 1].in_off + (*blocks)[nb - 1].in_len + 8, n) : (dev_bytes ? 0 : n);
    u64 total = 0;
    for (size_t i = 0; i < nb; i++) {
        F.idx.h_in_off[i] = (*blocks)[i].in_off - lo; F.idx.h_in_len[i] = (*blocks)[i].in_len; F.idx.h_isize[i] = (*blocks)[i].isize;
        F.idx.h_out_off[i] = out_prefix + total; F.idx.h_stage_off[i] = 0;
        total += (*blocks)[i].isize;
    }
    const u8 *comp = dev_bytes;
    u64 comp_total = n & ~(u64)3;
    const u64 span = hi > lo ? hi - lo : 0;
    if (!dev_bytes) {
        const u64 padded = (span + 3) & ~3ull;
        TRY(dev_reserve(ctx, F.comp, padded + 16));
        comp = F.comp.as<u8>();
        comp_total = padded;
    }
    TRY(dev_reserve(ctx, F.infl, out_prefix + total + 64, out_prefix, s));
    TRY(index_upload(ctx, F.idx, s));
    {
        const u32 l = lanes & 0xffu;
        lanes = ((l == 8 || l == 16 || l == 32 || (l >= 1 && l <= 4)) ? l : FASTF_INFLATE_DEFAULT) | (lanes & (FASTF_INFLATE_HW_ENGINE | FASTF_INFLATE_NO_CRC | FASTF_BAM_STRADDLE));
    }
    // One launch over all blocks.  (Measured on freq, 117 k blocks: sending the host bytes in groups of two kernel rounds on the copy
    // stream while the previous group inflates is SLOWER end to end, 491 vs 514 M reads/s, and four launches instead of one cost the
    // device-resident path 7 %: every launch pays for building 128 tables per SM before its decoders start, and for its tail.
    // FASTF_INFLATE_GROUP=<blocks> re-enables the grouping for experiments.)  The inflate clock is the sum of the launches.
    size_t group = std::max<size_t>(nb, 1);
    if (const char *e = getenv("FASTF_INFLATE_GROUP")) { const long v = atol(e); group = v > 0 ? (size_t)v : std::max<size_t>(nb, 1); }   // A/B knob: 0 = one launch
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ev_copy = nullptr;
    if (!dev_bytes) CK(cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming));
    int rc_l = 0;
    for (size_t g0 = 0; g0 < nb && !rc_l; g0 += group) {
        const size_t g1 = std::min(nb, g0 + group);
        if (!dev_bytes) {
            // bytes of this group: from its first payload (the very first group: from lo) to the end of its last block's trailer
            const u64 b0 = g0 == 0 ? 0 : (F.idx.h_in_off[g0] & ~3ull);
            const u64 b1 = std::min<u64>(F.idx.h_in_off[g1 - 1] + F.idx.h_in_len[g1 - 1] + 8, span);
            if (b1 > b0) CK(cudaMemcpyAsync(F.comp.as<u8>() + b0, (const u8 *)host_bytes + lo + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy));
            CK(cudaEventRecord(ev_copy, ctx->copy));
            CK(cudaStreamWaitEvent(s, ev_copy, 0));
        }
        if (ms) { cudaEvent_t a = nullptr, b = nullptr; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); ev.push_back(a); ev.push_back(b); CK(cudaEventRecord(a, s)); }
        rc_l = launch_inflate(ctx, lanes, comp, comp_total, F.idx.in_off + g0, F.idx.in_len + g0, F.idx.out_off + g0, F.idx.isize + g0, (u32)(g1 - g0), F.infl.as<u8>(), F.idx.st_infl + g0, s, &F.de,
                              F.idx.h_in_off + g0, F.idx.h_in_len + g0, F.idx.h_out_off + g0, F.idx.h_isize + g0);
        if (ms) CK(cudaEventRecord(ev.back(), s));
        if (!rc_l) rc_l = launch_crc(ctx, lanes, comp, dev_bytes ? (u64)n : comp_total, F.idx.in_off + g0, F.idx.in_len + g0, F.infl.as<u8>(), F.idx.out_off + g0, F.idx.isize + g0, (u32)(g1 - g0),
                                     F.idx.st_infl + g0, s);
    }
    if (rc_l) { for (auto e : ev) cudaEventDestroy(e); if (ev_copy) cudaEventDestroy(ev_copy); return rc_l; }
    // OR of the per-block status words
    std::vector<u32> st(nb);
    if (nb) CK(cudaMemcpyAsync(st.data(), F.idx.st_infl, nb * sizeof(u32), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (!dev_bytes) CK(cudaStreamSynchronize(ctx->copy));
    if (ms) {
        *ms = 0;
        for (size_t i = 0; i + 1 < ev.size(); i += 2) { float t = 0; cudaEventElapsedTime(&t, ev[i], ev[i + 1]); *ms += t; }
    }
    for (auto e : ev) cudaEventDestroy(e);
    if (ev_copy) cudaEventDestroy(ev_copy);
    F.status = 0;
    for (size_t i = 0; i < nb; i++) F.status |= st[i];
    F.n_blocks = nb;
    F.infl_bytes = total;
    if (F.status) {
        char buf[256];
        return ctx_fail(ctx, "inflate: malformed deflate data: %s", status_string(F.status, buf, sizeof buf));
    }
    return 0;
}

extern "C" int fastf_inflate_host(fastf_ctx *ctx, const void *bgzf_bytes, size_t n, int lanes, void **out, size_t *out_n, float *ms)
{
    CK(cudaSetDevice(ctx->device));
    *out = nullptr;
    *out_n = 0;
    InflatedFile F;
    int rc = inflate_whole(ctx, F, bgzf_bytes, n, nullptr, nullptr, (u32)lanes, ms, ctx->compute);
    if (!rc) {
        *out = malloc(F.infl_bytes ? F.infl_bytes : 1);
        if (!*out) rc = ctx_fail(ctx, "inflate_host: out of host memory");
        if (!rc && F.infl_bytes) rc = fastf_memcpy_d2h(ctx, *out, F.infl.p, F.infl_bytes);
        *out_n = F.infl_bytes;
    }
    inflated_release(ctx, F);
    return rc;
}

static int mt_host_common(fastf_ctx *ctx, uint32_t seed, uint64_t n, uint64_t threshold, uint32_t *out_words, uint32_t *out_bits)
{
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    const u64 pairs = (n + 1247) / 1248;
    DevBuf state, out;
    int rc = dev_reserve(ctx, state, 624 * sizeof(u32));
    const size_t out_bytes = out_words ? (size_t)pairs * 1248 * sizeof(u32) : (size_t)pairs * 39 * sizeof(u32);
    if (!rc) rc = dev_reserve(ctx, out, out_bytes);
    if (!rc) {
        FASTF_LAUNCH(fastf_mt19937_kernel, 1, FASTF_MT_THREADS, 0, ctx->compute, seed, state.as<u32>(), 1u, (u64)0, pairs, threshold, out_words ? out.as<u32>() : (u32 *)nullptr,
                     out_words ? (u32 *)nullptr : out.as<u32>());
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = ctx_fail(ctx, "mt19937 launch failed");
    }
    if (!rc) rc = out_words ? fastf_memcpy_d2h(ctx, out_words, out.p, (size_t)n * sizeof(u32)) : fastf_memcpy_d2h(ctx, out_bits, out.p, (size_t)((n + 31) / 32) * sizeof(u32));
    dev_release(ctx, state);
    dev_release(ctx, out);
    return rc;
}
extern "C" int fastf_mt19937_host_from(fastf_ctx *ctx, uint32_t seed, uint64_t first, uint64_t n, uint32_t *out_words)
{
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    const u64 pairs = (n + 1247) / 1248;
    DevBuf state, out;
    int rc = dev_reserve(ctx, state, 624 * sizeof(u32));
    if (!rc) rc = dev_reserve(ctx, out, (size_t)pairs * 1248 * sizeof(u32));
    if (!rc) rc = mt_state_at(ctx, seed, first, state.as<u32>(), ctx->compute);
    if (rc == 1 && !ctx->err[0]) ctx_fail(ctx, "mt19937_host_from: jump tables unavailable");
    if (!rc) {
        FASTF_LAUNCH(fastf_mt19937_kernel, 1, FASTF_MT_THREADS, 0, ctx->compute, seed, state.as<u32>(), 0u, (u64)0, pairs, (u64)0, out.as<u32>(), (u32 *)nullptr);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = ctx_fail(ctx, "mt19937 launch failed");
    }
    if (!rc) rc = fastf_memcpy_d2h(ctx, out_words, out.p, (size_t)n * sizeof(u32));
    dev_release(ctx, state);
    dev_release(ctx, out);
    return rc;
}
extern "C" int fastf_mt19937_host(fastf_ctx *ctx, uint32_t seed, uint64_t n, uint32_t *out_words) { return mt_host_common(ctx, seed, n, 0, out_words, nullptr); }
extern "C" int fastf_mt19937_keepbits_host(fastf_ctx *ctx, uint32_t seed, uint64_t n, uint64_t threshold, uint32_t *out_bits) { return mt_host_common(ctx, seed, n, threshold, nullptr, out_bits); }

extern "C" int fastf_sort_u64_host(fastf_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t n, uint32_t key_bits)
{
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    DevBuf dk, dv;
    int rc = dev_reserve(ctx, dk, n * sizeof(u64));
    if (!rc && vals) rc = dev_reserve(ctx, dv, n * sizeof(u32));
    if (!rc) rc = fastf_memcpy_h2d(ctx, dk.p, keys, n * sizeof(u64));
    if (!rc && vals) rc = fastf_memcpy_h2d(ctx, dv.p, vals, n * sizeof(u32));
    if (!rc) rc = fastf_sort_u64_device(ctx, dk.as<u64>(), vals ? dv.as<u32>() : nullptr, n, key_bits);
    if (!rc) rc = fastf_memcpy_d2h(ctx, keys, dk.p, n * sizeof(u64));
    if (!rc && vals) rc = fastf_memcpy_d2h(ctx, vals, dv.p, n * sizeof(u32));
    dev_release(ctx, dk);
    dev_release(ctx, dv);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// freq
// ---------------------------------------------------------------------------------------------------
// The text (inflated FASTQ) streams through HBM chunk by chunk; what stays resident is one u64 key per read.  A chunk's text sits
// at buf + FASTF_FREQ_CARRY with the last bytes of the text before it copied in front, so that a sequence line that crosses the
// chunk boundary is seen whole by exactly one chunk (fastf_freq_keys_kernel).
#define FASTF_FREQ_CARRY 64u   // >= 2 * FASTF_FREQ_EXC_STRIDE, multiple of 16
struct FreqStream {
    DevBuf tiles, tot, keys, exc_cnt, exc_ord, exc_bytes;
    PinBuf host;
    Timer t_keys;
    u64 nl_total = 0;            // newlines of the text so far
    u64 bytes_total = 0;
    u64 keys_used = 0;           // records with a key slot so far
    u32 exc_cap = 0, n_exc = 0;
    u8 tail[FASTF_FREQ_CARRY];   // host copy of the last bytes of the text so far
    u32 tail_len = 0;
    u32 key_len = 0;
};
static void freq_stream_release(fastf_ctx *ctx, FreqStream &Q)
{
    for (DevBuf *b : {&Q.tiles, &Q.tot, &Q.keys, &Q.exc_cnt, &Q.exc_ord, &Q.exc_bytes}) dev_release(ctx, *b);
    pin_release(ctx, Q.host);
    Q.t_keys.destroy();
}
static int freq_stream_init(fastf_ctx *ctx, FreqStream &Q, u32 key_len, cudaStream_t s)
{
    Q.key_len = key_len;
    if (Q.t_keys.init()) return ctx_fail(ctx, "freq: event creation failed");
    TRY(pin_reserve(ctx, Q.host, 256));
    TRY(dev_reserve(ctx, Q.tot, sizeof(u32)));
    TRY(dev_reserve(ctx, Q.exc_cnt, sizeof(u32)));
    CK(cudaMemsetAsync(Q.exc_cnt.p, 0, sizeof(u32), s));
    return 0;
}
// one chunk: buf holds chunk_bytes of text at buf + FASTF_FREQ_CARRY (the bytes in front are free)
static int freq_stream_chunk(fastf_ctx *ctx, FreqStream &Q, u8 *buf, u64 chunk_bytes, bool last, fastf_freq_result *res, cudaStream_t s)
{
    const u32 carry = Q.tail_len;
    if (carry) CK(cudaMemcpyAsync(buf + FASTF_FREQ_CARRY - carry, Q.tail, carry, cudaMemcpyHostToDevice, s));
    const u64 lead = FASTF_FREQ_CARRY - carry;           // first byte of the text inside buf
    const u8 *text = buf + (lead & ~15ull);              // 16-byte aligned for the vector loads
    const u64 skip = lead & 15ull;
    const u64 n = skip + carry + chunk_bytes;
    const u64 ntiles64 = (n + FASTF_NL_TILE - 1) / FASTF_NL_TILE;
    if (ntiles64 >= 0xffffffffull) return ctx_fail(ctx, "freq: chunk too large");
    const u32 ntiles = (u32)std::max<u64>(ntiles64, 1);
    TRY(dev_reserve(ctx, Q.tiles, (size_t)ntiles * sizeof(u32)));
    Q.t_keys.collect(&res->ms_keys);
    Q.t_keys.start(s);
    FASTF_LAUNCH(fastf_nl_count_kernel, ntiles, FASTF_NL_THREADS, 0, s, text, n, Q.tiles.as<u32>(), skip);
    CKL("nl_count");
    TRY(launch_scan_rows(ctx, Q.tiles.as<u32>(), ntiles, 1, Q.tot.as<u32>(), s));
    CK(cudaMemcpyAsync(Q.host.p, Q.tot.p, sizeof(u32), cudaMemcpyDeviceToHost, s));
    const u32 new_tail = (u32)std::min<u64>(FASTF_FREQ_CARRY, carry + chunk_bytes);
    if (new_tail) CK(cudaMemcpyAsync(Q.host.as<u8>() + 64, text + n - new_tail, new_tail, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const u64 nl_here = Q.host.as<u32>()[0];             // newlines of carry + chunk (a chunk holds < 2^32 bytes)
    u64 nl_carry = 0;
    for (u32 i = 0; i < carry; i++) nl_carry += Q.tail[i] == '\n';
    const u64 nl_base = Q.nl_total - nl_carry;           // global index of the first newline of this text
    Q.nl_total = nl_base + nl_here;
    Q.bytes_total += chunk_bytes;
    // key slots: record r exists as soon as newline 4r does
    const u64 n_keys = Q.nl_total ? (Q.nl_total - 1) / 4 + 1 : 0;
    if (n_keys >= 0xffffffffull) return ctx_fail(ctx, "freq: more than 2^32-1 reads in one pass");
    TRY(dev_reserve(ctx, Q.keys, std::max<u64>(n_keys, 1) * sizeof(u64), Q.keys_used * sizeof(u64), s));
    if (Q.exc_cap == 0) Q.exc_cap = 1u << 16;
    const u32 exc_before = Q.n_exc;
    for (int attempt = 0; attempt < 2; attempt++) {
        TRY(dev_reserve(ctx, Q.exc_ord, (size_t)Q.exc_cap * sizeof(u32), (size_t)exc_before * sizeof(u32), s));
        TRY(dev_reserve(ctx, Q.exc_bytes, (size_t)Q.exc_cap * FASTF_FREQ_EXC_STRIDE, (size_t)exc_before * FASTF_FREQ_EXC_STRIDE, s));
        CK(cudaMemcpyAsync(Q.exc_cnt.p, &exc_before, sizeof(u32), cudaMemcpyHostToDevice, s));
        FASTF_LAUNCH(fastf_freq_keys_kernel, ntiles, FASTF_NL_THREADS, 0, s, text, n, (const u32 *)Q.tiles.as<u32>(), Q.key_len, Q.keys.as<u64>(), n_keys, Q.exc_cnt.as<u32>(), Q.exc_cap,
                     Q.exc_ord.as<u32>(), Q.exc_bytes.as<u8>(), skip, nl_base, skip + carry, (u32)last);
        CKL("freq_keys");
        CK(cudaMemcpyAsync(Q.host.p, Q.exc_cnt.p, sizeof(u32), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        Q.n_exc = Q.host.as<u32>()[0];
        if (Q.n_exc <= Q.exc_cap) break;
        Q.exc_cap = std::max(Q.n_exc, Q.exc_cap * 2);   // rerun this chunk with room for every exceptional read
    }
    Q.t_keys.stop(s);
    Q.keys_used = n_keys;
    memcpy(Q.tail, Q.host.as<u8>() + 64, new_tail);
    Q.tail_len = new_tail;
    return 0;
}

// after the last chunk: compact the good keys with their read ordinals, sort, run-length encode, copy back
static int freq_stream_finish(fastf_ctx *ctx, FreqStream &Q, fastf_freq_result *res, cudaStream_t s)
{
    DevBuf ckeys, cidx, kalt, valt, orand, tiles, tot;
    PinBuf host;
    SortScratch S;
    RleScratch R;
    Timer t_sort, t_rle;
    auto cleanup = [&]() {
        for (DevBuf *b : {&ckeys, &cidx, &kalt, &valt, &orand, &tiles, &tot}) dev_release(ctx, *b);
        pin_release(ctx, host);
        sort_scratch_release(ctx, S);
        rle_scratch_release(ctx, R);
        t_sort.destroy(); t_rle.destroy();
    };
    auto body = [&]() -> int {
        if (t_sort.init() || t_rle.init()) return ctx_fail(ctx, "freq: event creation failed");
        TRY(pin_reserve(ctx, host, 64));
        TRY(dev_reserve(ctx, orand, 2 * sizeof(u64)));
        TRY(dev_reserve(ctx, tot, sizeof(u32)));
        Q.t_keys.collect(&res->ms_keys);
        const u64 n = Q.bytes_total, n_newlines = Q.nl_total;
        const u8 last = Q.tail_len ? Q.tail[Q.tail_len - 1] : (u8)'\n';
        const u64 n_lines = n_newlines + ((n && last != '\n') ? 1 : 0);
        // get_fastq reads four lines per record; a record exists as soon as its id line does (reference src/filter.c:22-34)
        res->n_lines = n_lines;
        res->last_byte_is_newline = (u8)(n == 0 || last == '\n');
        res->n_reads = (n_lines + 3) / 4;
        // records whose sequence line starts after newline 4r: r = 0 .. n_keys_dev-1 where newline 4r exists; a trailing record whose
        // id line is not newline-terminated has an empty key (NUL immediately): the host handles it
        const u64 n_keys_dev = Q.keys_used;
        const u32 n_exc = Q.n_exc;
        const u64 n_good = n_keys_dev - n_exc;
        u64 ngroups = 0;
        if (n_good) {
            const u32 ctiles = (u32)((n_keys_dev + FASTF_CP_TILE - 1) / FASTF_CP_TILE);
            TRY(dev_reserve(ctx, tiles, (size_t)ctiles * sizeof(u32)));
            TRY(dev_reserve(ctx, ckeys, n_good * sizeof(u64)));
            TRY(dev_reserve(ctx, cidx, n_good * sizeof(u32)));
            TRY(dev_reserve(ctx, kalt, n_good * sizeof(u64)));
            TRY(dev_reserve(ctx, valt, n_good * sizeof(u32)));
            t_sort.start(s);
            FASTF_LAUNCH(fastf_compact_count_kernel, ctiles, FASTF_CP_THREADS, 0, s, (const u64 *)Q.keys.as<u64>(), n_keys_dev, tiles.as<u32>());
            CKL("compact_count");
            TRY(launch_scan_rows(ctx, tiles.as<u32>(), ctiles, 1, tot.as<u32>(), s));
            FASTF_LAUNCH(fastf_compact_scatter_kernel, ctiles, FASTF_CP_THREADS, 0, s, (const u64 *)Q.keys.as<u64>(), n_keys_dev, (const u32 *)tiles.as<u32>(), ckeys.as<u64>(), cidx.as<u32>());
            CKL("compact_scatter");
            dev_release(ctx, Q.keys);   // the compacted copy is what is sorted
            u64 varying = 0;
            TRY(varying_bits(ctx, orand, host, ckeys.as<u64>(), n_good, &varying, s));
            u32 shifts[8];
            const int npass = plan_windows(varying, shifts);
            bool in_alt = false;
            TRY(sort_keys(ctx, S, ckeys.as<u64>(), kalt.as<u64>(), cidx.as<u32>(), valt.as<u32>(), n_good, shifts, npass, &in_alt, s));
            t_sort.stop(s);
            t_rle.start(s);
            u64 nd = 0;
            TRY(rle_groups(ctx, R, in_alt ? kalt.as<u64>() : ckeys.as<u64>(), in_alt ? valt.as<u32>() : cidx.as<u32>(), n_good, 0, 64, 0, &ngroups, &nd, s));
            t_rle.stop(s);
        }
        // ---- results to host ----
        res->n_keys = ngroups;
        res->key = (u64 *)malloc(std::max<u64>(ngroups, 1) * sizeof(u64));
        res->count = (u32 *)malloc(std::max<u64>(ngroups, 1) * sizeof(u32));
        res->first = (u32 *)malloc(std::max<u64>(ngroups, 1) * sizeof(u32));
        res->n_exceptions = n_exc;
        res->exc_stride = FASTF_FREQ_EXC_STRIDE;
        res->exc_ordinal = (u32 *)malloc(std::max<u64>(n_exc, 1) * sizeof(u32));
        res->exc_bytes = (u8 *)malloc(std::max<u64>(n_exc, 1) * FASTF_FREQ_EXC_STRIDE);
        if (!res->key || !res->count || !res->first || !res->exc_ordinal || !res->exc_bytes) return ctx_fail(ctx, "freq: out of host memory");
        if (ngroups) {
            // with group_shift 0 every distinct key is a group and counts all its copies: count = next first - first
            TRY(d2h_pageable(ctx, res->key, R.grp_key.p, ngroups * sizeof(u64), s));
            TRY(d2h_pageable(ctx, res->first, R.grp_val.p, ngroups * sizeof(u32), s));
            TRY(d2h_pageable(ctx, res->count, R.grp_first.p, ngroups * sizeof(u32), s));
        }
        if (n_exc) {
            CK(cudaMemcpyAsync(res->exc_ordinal, Q.exc_ord.p, (size_t)n_exc * sizeof(u32), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(res->exc_bytes, Q.exc_bytes.p, (size_t)n_exc * FASTF_FREQ_EXC_STRIDE, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
        // grp_first[g] = index of the group's first element in the sorted array -> multiplicity by differencing
        for (u64 g = 0; g < ngroups; g++) {
            u64 nxt = (g + 1 < ngroups) ? res->count[g + 1] : n_good;
            res->count[g] = (u32)(nxt - res->count[g]);
        }
        t_sort.collect(&res->ms_sort);
        t_rle.collect(&res->ms_rle);
        return 0;
    };
    int rc = body();
    cudaStreamSynchronize(s);
    cleanup();
    return rc;
}

static bool looks_like_gzip(const u8 *p, size_t n) { return n >= 2 && p[0] == 0x1f && p[1] == 0x8b; }

// blocks per streamed chunk of the tag / freq jobs: two rounds of the persistent inflate kernel (tests shrink it to force many chunks)
static size_t stream_chunk_blocks(const fastf_ctx *ctx)
{
    if (ctx->taghist_chunk_blocks) return (size_t)ctx->taghist_chunk_blocks;
    if (const char *e = getenv("FASTF_STREAM_CHUNK_BLOCKS")) { const long v = atol(e); if (v > 0) return (size_t)v; }   // tests: many chunks on small inputs
    return 2ull * (size_t)ctx->n_sm * FASTF_TPS_STREAMS;
}

static int freq_common(fastf_ctx *ctx, const void *host_bytes, size_t n, const u8 *dev_bytes, const std::vector<FastfBgzfBlock> *pre, uint32_t key_len, uint32_t lanes, fastf_freq_result *res)
{
    CK(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    if (key_len == 0 || key_len > FASTF_FREQ_MAX_KEY) return ctx_fail(ctx, "freq: len_cellbarcode + len_umi must be 1..%d (got %u)", FASTF_FREQ_MAX_KEY, key_len);
    const u32 l0 = ctx->launches;
    cudaStream_t s = ctx->compute;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    InflatedFile F;
    FreqStream Q;
    DevBuf plain;
    std::vector<FastfBgzfBlock> all, part;
    auto body = [&]() -> int {
        TRY(freq_stream_init(ctx, Q, key_len, s));
        cudaEventRecord(e0, s);
        if (dev_bytes || looks_like_gzip((const u8 *)host_bytes, n)) {
            const std::vector<FastfBgzfBlock> *blocks = pre;
            if (!dev_bytes) {
                // plain (non-BGZF) gzip cannot be inflated block-parallel: refuse loudly rather than fall back to the CPU
                const u8 *h = (const u8 *)host_bytes;
                if (n < 18 || !(h[3] & 4)) return ctx_fail(ctx, "freq: input is single-member gzip, not BGZF; recompress with bgzip (no CPU fallback; the reference reads it through zlib's gzopen)");
                size_t used = 0;
                const int irc = fastf_bgzf_index(h, n, 0, all, &used);
                if (irc != FASTF_BGZF_OK) return ctx_fail(ctx, "inflate: not a whole BGZF stream (index error %d at byte %zu of %zu)", irc, used, n);
                blocks = &all;
            }
            // chunks of whole BGZF blocks: inflate -> newline scan -> keys; only the keys stay resident
            const size_t nb = blocks->size(), per = stream_chunk_blocks(ctx);
            res->n_blocks = nb;
            for (size_t b0 = 0; b0 < nb || b0 == 0; b0 += per) {
                const size_t b1 = std::min(nb, b0 + per);
                part.assign(blocks->begin() + (ptrdiff_t)b0, blocks->begin() + (ptrdiff_t)b1);
                float ms_infl = 0;
                TRY(inflate_whole(ctx, F, host_bytes, n, dev_bytes, &part, lanes, &ms_infl, s, FASTF_FREQ_CARRY));
                res->ms_inflate += ms_infl;
                res->status |= F.status;
                TRY(freq_stream_chunk(ctx, Q, F.infl.as<u8>(), F.infl_bytes, b1 >= nb, res, s));
                if (nb == 0) break;
            }
        } else {
            // plain text: pieces of the host buffer
            size_t PIECE = (size_t)256 << 20;
            if (const char *e = getenv("FASTF_STREAM_CHUNK_BLOCKS")) { const long v = atol(e); if (v > 0) PIECE = (size_t)v * 1000; }   // tests: pieces of a few KB
            TRY(dev_reserve(ctx, plain, std::min<size_t>(n, PIECE) + FASTF_FREQ_CARRY + 64));
            for (size_t o = 0; o < n || o == 0; o += PIECE) {
                const size_t m = std::min(PIECE, n - o);
                if (m) CK(cudaMemcpyAsync(plain.as<u8>() + FASTF_FREQ_CARRY, (const u8 *)host_bytes + o, m, cudaMemcpyHostToDevice, s));
                TRY(freq_stream_chunk(ctx, Q, plain.as<u8>(), m, o + m >= n, res, s));
                if (n == 0) break;
            }
        }
        res->compressed_bytes = n;
        res->inflated_bytes = Q.bytes_total;
        return freq_stream_finish(ctx, Q, res, s);
    };
    int rc = body();
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&res->ms_device_total, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    inflated_release(ctx, F);
    freq_stream_release(ctx, Q);
    dev_release(ctx, plain);
    res->n_launches = ctx->launches - l0;
    if (rc) fastf_freq_result_free(res);
    return rc;
}

extern "C" int fastf_freq_gpu(fastf_ctx *ctx, const void *host_bytes, size_t n, uint32_t key_len, uint32_t inflate_lanes, fastf_freq_result *res)
{
    return freq_common(ctx, host_bytes, n, nullptr, nullptr, key_len, inflate_lanes, res);
}
extern "C" int fastf_freq_gpu_device(fastf_ctx *ctx, const void *dev_bytes, size_t nbytes, const uint64_t *in_off, const uint32_t *in_len, const uint32_t *isize, uint64_t nblocks, uint32_t key_len,
                                     uint32_t inflate_lanes, fastf_freq_result *res)
{
    if (((uintptr_t)dev_bytes & 3u) != 0) return ctx_fail(ctx, "freq_gpu_device: dev_bytes must be 4-byte aligned");
    std::vector<FastfBgzfBlock> blocks(nblocks);
    for (u64 i = 0; i < nblocks; i++) {
        if (in_off[i] + in_len[i] + 8 > nbytes || isize[i] > 65536) return ctx_fail(ctx, "freq_gpu_device: block %llu outside the buffer", (unsigned long long)i);
        blocks[i].in_off = in_off[i]; blocks[i].in_len = in_len[i]; blocks[i].isize = isize[i]; blocks[i].crc32 = 0;
    }
    return freq_common(ctx, nullptr, nbytes, (const u8 *)dev_bytes, &blocks, key_len, inflate_lanes, res);
}
// ---------------------------------------------------------------------------------------------------
// crb / extract: histogram of one aux tag (or of the pair of two) over all records of a BAM image
// ---------------------------------------------------------------------------------------------------
extern "C" void fastf_taghist_result_free(fastf_taghist_result *res)
{
    if (!res) return;
    free(res->first); free(res->count); free(res->ivalue); free(res->a_off); free(res->a_len); free(res->b_len); free(res->strings);
    res->first = nullptr; res->count = nullptr; res->ivalue = nullptr; res->a_off = nullptr; res->a_len = nullptr; res->b_len = nullptr; res->strings = nullptr;
}

// One chunk's groups, merged on the host across chunks (a value's count adds up, its first occurrence is the smallest ordinal)
struct TagAgg { u64 count; u64 first; };

extern "C" int fastf_taghist_gpu(fastf_ctx *ctx, const void *host_bytes, size_t n, const char *tag_a, uint32_t mode, const char *tag_b, uint32_t inflate_lanes, fastf_taghist_result *res)
{
    CK(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    if (!tag_a || !tag_a[0] || !tag_a[1]) return ctx_fail(ctx, "taghist: a tag is two characters");
    if (tag_b && (!tag_b[0] || !tag_b[1])) return ctx_fail(ctx, "taghist: a tag is two characters");
    if (mode > FASTF_TAG_MODE_INT || (mode == FASTF_TAG_MODE_INT && tag_b)) return ctx_fail(ctx, "taghist: mode 0 = string (optionally a pair), 1 = integer");
    if (!looks_like_gzip((const u8 *)host_bytes, n)) return ctx_fail(ctx, "taghist: not a BGZF stream");
    res->mode = mode;
    const u32 l0 = ctx->launches;
    cudaStream_t s = ctx->compute;
    InflatedFile F;
    DevBuf hdr_off, counters, stage_off, stage, keys, loc_a, loc_b, vals, kalt, valt, orand, coll, rep_a, rep_b, blob_off, blob, virt;
    const bool straddle = (inflate_lanes & FASTF_BAM_STRADDLE) != 0;   // records may cross BGZF block boundaries (bam_straddle.cuh): one chunk
    PinBuf host;
    SortScratch S;
    RleScratch R;
    Timer t_tags, t_sort, t_rle;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<FastfBgzfBlock> all, part;
    std::vector<u64> h_stage_off, h_rep_a, h_rep_b, h_blob_off, h_key;
    std::vector<u32> h_start, h_first;
    std::vector<char> h_blob;
    // groups merged across chunks: strings keyed by "A \0 B" (values hold no NUL), integers by value
    std::unordered_map<std::string, TagAgg> smap;
    std::unordered_map<int32_t, TagAgg> imap;
    auto cleanup = [&]() {
        for (DevBuf *b : {&hdr_off, &counters, &stage_off, &stage, &keys, &loc_a, &loc_b, &vals, &kalt, &valt, &orand, &coll, &rep_a, &rep_b, &blob_off, &blob, &virt}) dev_release(ctx, *b);
        pin_release(ctx, host);
        sort_scratch_release(ctx, S);
        rle_scratch_release(ctx, R);
        t_tags.destroy(); t_sort.destroy(); t_rle.destroy();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        inflated_release(ctx, F);
    };
    auto body = [&]() -> int {
        if (t_tags.init() || t_sort.init() || t_rle.init() || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return ctx_fail(ctx, "taghist: event creation failed");
        {
            size_t used = 0;
            int rc = fastf_bgzf_index((const u8 *)host_bytes, n, 0, all, &used);
            if (rc != FASTF_BGZF_OK) return ctx_fail(ctx, "taghist: not a whole BGZF stream (index error %d at byte %zu of %zu)", rc, used, n);
        }
        CK(cudaEventRecord(e0, s));
        res->n_blocks = all.size(); res->compressed_bytes = n;
        TRY(pin_reserve(ctx, host, 64));
        TRY(dev_reserve(ctx, hdr_off, sizeof(u64)));
        TRY(dev_reserve(ctx, counters, 4 * sizeof(u64)));
        TRY(dev_reserve(ctx, orand, 2 * sizeof(u64)));
        TRY(dev_reserve(ctx, coll, sizeof(u32)));
        FastfTagQuery Q;
        Q.a0 = (u8)tag_a[0]; Q.a1 = (u8)tag_a[1];
        Q.b0 = tag_b ? (u8)tag_b[0] : 0u; Q.b1 = tag_b ? (u8)tag_b[1] : 0u;
        Q.mode = mode;
        // The file streams through HBM in chunks of whole blocks (two full rounds of the persistent inflate kernel each); every chunk
        // is grouped on the device, the per-chunk groups are merged here.
        size_t chunk_blocks = std::max<size_t>(1, ctx->taghist_chunk_blocks ? ctx->taghist_chunk_blocks : 2ull * (size_t)ctx->n_sm * FASTF_TPS_STREAMS);
        if (straddle) chunk_blocks = std::max<size_t>(all.size(), 1);
        else if (!ctx->taghist_chunk_blocks) {
            // a file whose inflated bytes, staging planes (worst case 2/3 of them) and key arrays fit HBM comfortably goes through in ONE
            // chunk: no host-side merge at all
            u64 infl_total = 0;
            for (auto &b : all) infl_total += b.isize;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (double)infl_total * 2.6 + (double)n < 0.7 * (double)free_b && all.size() < 0xffffffffull) chunk_blocks = all.size();
        }
        u64 hit_base = 0;
        bool first_chunk = true, single = false;
        u64 single_groups = 0, single_hits = 0;
        res->hash_rounds = 1;
        smap.reserve(1u << 16);
        for (size_t c0 = 0; c0 < all.size() || first_chunk; c0 += chunk_blocks) {
            const size_t c1 = std::min(all.size(), c0 + chunk_blocks);
            part.assign(all.begin() + c0, all.begin() + c1);
            float ms_infl = 0;
            TRY(inflate_whole(ctx, F, host_bytes, n, nullptr, &part, inflate_lanes, &ms_infl, s));
            res->ms_inflate += ms_infl;
            res->inflated_bytes += F.infl_bytes;
            const u32 nb = (u32)F.n_blocks;
            // per-block staging slices (a record is >= 36 bytes)
            h_stage_off.resize((size_t)nb + 1);
            u64 plane = 0;
            for (u32 i = 0; i < nb; i++) {
                h_stage_off[i] = plane;
                plane += straddle ? stage_cap_for(F.idx.h_isize[i] + (i + 1 < nb ? F.idx.h_isize[i + 1] : 0)) + 1u : stage_cap_for(F.idx.h_isize[i]);
            }
            h_stage_off[nb] = plane;   // the kernel reads the slice capacity as stage_off[b + 1] - stage_off[b]
            TRY(dev_reserve(ctx, stage_off, h_stage_off.size() * sizeof(u64)));
            CK(cudaMemcpyAsync(stage_off.p, h_stage_off.data(), h_stage_off.size() * sizeof(u64), cudaMemcpyHostToDevice, s));
            TRY(dev_reserve(ctx, stage, std::max<u64>(plane, 1) * 3 * sizeof(u64)));
            u64 n_hits = 0, ngroups = 0, n_rec = 0;
            for (u32 round = 0;; round++) {
                if (round == 4) return ctx_fail(ctx, "taghist: 64-bit hash collisions in four rounds with different seeds");
                Q.seed = 0x9e3779b97f4a7c15ull * round;
                Q.key_mask = (round == 0 && ctx->taghist_round0_mask) ? ctx->taghist_round0_mask : ~0ull;
                if (round + 1 > res->hash_rounds) res->hash_rounds = round + 1;
                CK(cudaMemsetAsync(counters.p, 0, 4 * sizeof(u64), s));
                CK(cudaMemsetAsync(hdr_off.p, 0, sizeof(u64), s));
                t_tags.collect(&res->ms_tags);
                t_tags.start(s);
                if (first_chunk) {
                    FASTF_LAUNCH(fastf_bam_header_kernel, 1, 32, 0, s, (const u8 *)F.infl.as<u8>(), F.infl_bytes, hdr_off.as<u64>(), (u32 *)(counters.as<u64>() + 2));
                    CKL("bam_header");
                }
                const u64 *p_off = F.idx.out_off;
                const u32 *p_size = F.idx.isize;
                if (straddle) TRY(launch_virtual_blocks(ctx, virt, (const u8 *)F.infl.as<u8>(), F.infl_bytes, F.idx.out_off, F.idx.isize, nb, hdr_off.as<u64>(), (u32 *)(counters.as<u64>() + 2), &p_off, &p_size, s));
                if (nb) {
                    FASTF_LAUNCH(fastf_bam_tags_kernel, (nb + FASTF_PARSE_WARPS - 1) / FASTF_PARSE_WARPS, FASTF_PARSE_WARPS * 32, 0, s, (const u8 *)F.infl.as<u8>(), (u64)((F.infl_bytes + 15) & ~15ull),
                                 p_off, p_size, nb, (const u64 *)hdr_off.as<u64>(), Q, (const u64 *)stage_off.as<u64>(), stage.as<u64>(), plane, F.idx.nrec,
                                 F.idx.ncbv, F.idx.st_parse);
                    CKL("bam_tags");
                }
                FASTF_LAUNCH(fastf_chunk_counts_kernel, 1, FASTF_SCAN_THREADS, 0, s, (const u32 *)F.idx.nrec, (const u32 *)F.idx.ncbv, (const u32 *)F.idx.st_infl, (const u32 *)F.idx.st_parse, nb,
                             F.idx.dst_base, counters.as<u64>());
                CKL("chunk_counts");
                t_tags.stop(s);
                CK(cudaMemcpyAsync(host.p, counters.p, 4 * sizeof(u64), cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                n_rec = host.as<u64>()[0];
                n_hits = host.as<u64>()[1];
                res->status |= (u32)host.as<u64>()[2];
                if (res->status) {
                    char buf[256];
                    if (res->status & FASTF_ST_TAG_TYPE)
                        return ctx_fail(ctx, "taghist: tag-not-a-string: a record carries %c%c%s as a non-string value%s (the reference passes bam_aux2Z()'s NULL to strcmp/strcpy there)", tag_a[0],
                                        tag_a[1], tag_b ? " or its partner tag" : "", tag_b ? ", or lacks the partner tag" : "");
                    return ctx_fail(ctx, "taghist: malformed input: %s", status_string(res->status, buf, sizeof buf));
                }
                if (hit_base + n_hits >= 0xffffffffull) return ctx_fail(ctx, "taghist: more than 2^32-1 tagged records");
                ngroups = 0;
                if (!n_hits) break;
                TRY(dev_reserve(ctx, keys, n_hits * sizeof(u64)));
                TRY(dev_reserve(ctx, loc_a, n_hits * sizeof(u64)));
                TRY(dev_reserve(ctx, loc_b, n_hits * sizeof(u64)));
                TRY(dev_reserve(ctx, kalt, n_hits * sizeof(u64)));
                TRY(dev_reserve(ctx, vals, n_hits * sizeof(u32)));
                TRY(dev_reserve(ctx, valt, n_hits * sizeof(u32)));
                t_sort.collect(&res->ms_sort);
                t_sort.start(s);
                u64 *planes[3] = {keys.as<u64>(), loc_a.as<u64>(), loc_b.as<u64>()};
                for (int k = 0; k < 3; k++) {
                    FASTF_LAUNCH(fastf_stage_gather_kernel, (nb + 7) / 8, 256, 0, s, (const u64 *)(stage.as<u64>() + (u64)k * plane), (const u64 *)stage_off.as<u64>(), (const u32 *)F.idx.ncbv,
                                 (const u64 *)F.idx.dst_base, nb, planes[k]);
                    CKL("stage_gather");
                }
                FASTF_LAUNCH(fastf_iota_kernel, (u32)((n_hits + 255) / 256), 256, 0, s, vals.as<u32>(), n_hits);
                CKL("iota");
                u64 varying = 0;
                TRY(varying_bits(ctx, orand, host, keys.as<u64>(), n_hits, &varying, s));
                u32 shifts[8];
                const int npass = plan_windows(varying, shifts);
                bool in_alt = false;
                TRY(sort_keys(ctx, S, keys.as<u64>(), kalt.as<u64>(), vals.as<u32>(), valt.as<u32>(), n_hits, shifts, npass, &in_alt, s));
                t_sort.stop(s);
                const u64 *sorted = in_alt ? kalt.as<u64>() : keys.as<u64>();
                const u32 *perm = in_alt ? valt.as<u32>() : vals.as<u32>();
                t_rle.collect(&res->ms_rle);
                t_rle.start(s);
                u64 nd = 0;
                TRY(rle_groups(ctx, R, sorted, perm, n_hits, 0, 64, 0, &ngroups, &nd, s));
                u32 collided = 0;
                if (mode == FASTF_TAG_MODE_STRING) {
                    CK(cudaMemsetAsync(coll.p, 0, sizeof(u32), s));
                    FASTF_LAUNCH(fastf_taghist_verify_kernel, (u32)((n_hits + 255) / 256), 256, 0, s, (const u8 *)F.infl.as<u8>(), sorted, perm, (const u64 *)loc_a.as<u64>(),
                                 (const u64 *)loc_b.as<u64>(), n_hits, coll.as<u32>());
                    CKL("taghist_verify");
                    CK(cudaMemcpyAsync(host.p, coll.p, sizeof(u32), cudaMemcpyDeviceToHost, s));
                    CK(cudaStreamSynchronize(s));
                    collided = host.as<u32>()[0];
                }
                t_rle.stop(s);
                if (!collided) break;   // every group holds one value: done.  Otherwise hash again with another seed.
            }
            // ---- this chunk's groups to the host, merged into the maps ----
            res->n_records += n_rec;
            res->n_hits += n_hits;
            if (ngroups) {
                h_start.resize(ngroups); h_first.resize(ngroups); h_key.resize(ngroups);
                CK(cudaMemcpyAsync(h_first.data(), R.grp_val.p, ngroups * sizeof(u32), cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(h_start.data(), R.grp_first.p, ngroups * sizeof(u32), cudaMemcpyDeviceToHost, s));
                CK(cudaMemcpyAsync(h_key.data(), R.grp_key.p, ngroups * sizeof(u64), cudaMemcpyDeviceToHost, s));
                if (mode == FASTF_TAG_MODE_STRING) {
                    TRY(dev_reserve(ctx, rep_a, ngroups * sizeof(u64)));
                    TRY(dev_reserve(ctx, rep_b, ngroups * sizeof(u64)));
                    FASTF_LAUNCH(fastf_taghist_reps_kernel, (u32)((ngroups + 255) / 256), 256, 0, s, (const u32 *)R.grp_val.as<u32>(), (const u64 *)loc_a.as<u64>(), (const u64 *)loc_b.as<u64>(),
                                 (u32)ngroups, rep_a.as<u64>(), rep_b.as<u64>());
                    CKL("taghist_reps");
                    h_rep_a.resize(ngroups); h_rep_b.resize(ngroups); h_blob_off.resize(ngroups);
                    CK(cudaMemcpyAsync(h_rep_a.data(), rep_a.p, ngroups * sizeof(u64), cudaMemcpyDeviceToHost, s));
                    CK(cudaMemcpyAsync(h_rep_b.data(), rep_b.p, ngroups * sizeof(u64), cudaMemcpyDeviceToHost, s));
                }
                CK(cudaStreamSynchronize(s));
                if (mode == FASTF_TAG_MODE_STRING) {
                    u64 total = 0;
                    for (u64 g = 0; g < ngroups; g++) { h_blob_off[g] = total; total += (h_rep_a[g] & 0xffffu) + (h_rep_b[g] & 0xffffu); }
                    h_blob.resize(std::max<u64>(total, 1));
                    TRY(dev_reserve(ctx, blob_off, ngroups * sizeof(u64)));
                    TRY(dev_reserve(ctx, blob, std::max<u64>(total, 1)));
                    CK(cudaMemcpyAsync(blob_off.p, h_blob_off.data(), ngroups * sizeof(u64), cudaMemcpyHostToDevice, s));
                    FASTF_LAUNCH(fastf_taghist_strings_kernel, (u32)((ngroups * 32 + 255) / 256), 256, 0, s, (const u8 *)F.infl.as<u8>(), (const u64 *)rep_a.as<u64>(), (const u64 *)rep_b.as<u64>(),
                                 (const u64 *)blob_off.as<u64>(), (u32)ngroups, blob.as<u8>());
                    CKL("taghist_strings");
                    if (total) CK(cudaMemcpyAsync(h_blob.data(), blob.p, total, cudaMemcpyDeviceToHost, s));
                    CK(cudaStreamSynchronize(s));
                }
                if (c0 == 0 && c1 == all.size()) {
                    // the whole file was one chunk: its groups are the result, no merge
                    single = true;
                    single_groups = ngroups;
                    single_hits = n_hits;
                    break;
                }
                std::string k;
                for (u64 g = 0; g < ngroups; g++) {
                    const u64 cnt = (g + 1 < ngroups ? h_start[g + 1] : (u32)n_hits) - h_start[g], fst = hit_base + h_first[g];
                    TagAgg *a;
                    if (mode == FASTF_TAG_MODE_INT) a = &imap.emplace((int32_t)(u32)h_key[g], TagAgg{0, ~0ull}).first->second;
                    else {
                        const u32 la = (u32)(h_rep_a[g] & 0xffffu), lb = (u32)(h_rep_b[g] & 0xffffu);
                        k.assign(h_blob.data() + h_blob_off[g], la);
                        k.push_back('\0');
                        k.append(h_blob.data() + h_blob_off[g] + la, lb);
                        a = &smap.emplace(k, TagAgg{0, ~0ull}).first->second;
                    }
                    a->count += cnt;
                    if (fst < a->first) a->first = fst;
                }
            }
            hit_base += n_hits;
            first_chunk = false;
            if (all.empty()) break;
        }
        // ---- merged groups -> result arrays ----
        const u64 ngroups = single ? single_groups : (mode == FASTF_TAG_MODE_INT ? imap.size() : smap.size());
        res->n_groups = ngroups;
        const u64 ng1 = std::max<u64>(ngroups, 1);
        u64 sbytes = 0;
        for (auto &kv : smap) sbytes += kv.first.size() - 1;
        if (single && mode == FASTF_TAG_MODE_STRING) for (u64 g = 0; g < ngroups; g++) sbytes += (h_rep_a[g] & 0xffffu) + (h_rep_b[g] & 0xffffu);
        res->first = (u32 *)malloc(ng1 * sizeof(u32));
        res->count = (u32 *)malloc(ng1 * sizeof(u32));
        res->ivalue = (int32_t *)malloc(ng1 * sizeof(int32_t));
        res->a_off = (u64 *)malloc(ng1 * sizeof(u64));
        res->a_len = (u32 *)malloc(ng1 * sizeof(u32));
        res->b_len = (u32 *)malloc(ng1 * sizeof(u32));
        res->strings = (char *)malloc(std::max<u64>(sbytes, 1));
        res->strings_bytes = sbytes;
        if (!res->first || !res->count || !res->ivalue || !res->a_off || !res->a_len || !res->b_len || !res->strings) return ctx_fail(ctx, "taghist: out of host memory");
        u64 g = 0, at = 0;
        if (single) {
            for (; g < ngroups; g++) {
                res->count[g] = (u32)((g + 1 < ngroups ? h_start[g + 1] : (u32)single_hits) - h_start[g]);
                res->first[g] = h_first[g];
                res->ivalue[g] = (int32_t)(u32)h_key[g];
                res->a_off[g] = mode == FASTF_TAG_MODE_STRING ? h_blob_off[g] : 0;
                res->a_len[g] = mode == FASTF_TAG_MODE_STRING ? (u32)(h_rep_a[g] & 0xffffu) : 0;
                res->b_len[g] = mode == FASTF_TAG_MODE_STRING ? (u32)(h_rep_b[g] & 0xffffu) : 0;
            }
            if (mode == FASTF_TAG_MODE_STRING && sbytes) memcpy(res->strings, h_blob.data(), sbytes);
        }
        for (auto &kv : imap) { res->ivalue[g] = kv.first; res->first[g] = (u32)kv.second.first; res->count[g] = (u32)kv.second.count; res->a_off[g] = 0; res->a_len[g] = 0; res->b_len[g] = 0; g++; }
        for (auto &kv : smap) {
            const std::string &k = kv.first;
            const size_t la = k.find('\0'), lb = k.size() - la - 1;
            res->ivalue[g] = 0; res->first[g] = (u32)kv.second.first; res->count[g] = (u32)kv.second.count;
            res->a_off[g] = at; res->a_len[g] = (u32)la; res->b_len[g] = (u32)lb;
            memcpy(res->strings + at, k.data(), la);
            memcpy(res->strings + at + la, k.data() + la + 1, lb);
            at += la + lb;
            g++;
        }
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        t_tags.collect(&res->ms_tags); t_sort.collect(&res->ms_sort); t_rle.collect(&res->ms_rle);
        cudaEventElapsedTime(&res->ms_device_total, e0, e1);
        return 0;
    };
    int rc = body();
    cleanup();
    res->n_launches = ctx->launches - l0;
    if (rc) fastf_taghist_result_free(res);
    return rc;
}

/* test hooks: a smaller streaming chunk (so that tiny fixtures exercise the cross-chunk merge), and a mask ANDed onto the hash keys of
 * the FIRST round only, which forces collisions there: the byte-for-byte verification must catch them and the second round must win */
extern "C" void fastf_taghist_test_hooks(fastf_ctx *ctx, uint64_t chunk_blocks, uint64_t round0_key_mask)
{
    ctx->taghist_chunk_blocks = chunk_blocks;
    ctx->taghist_round0_mask = round0_key_mask;
}

extern "C" void fastf_freq_result_free(fastf_freq_result *res)
{
    if (!res) return;
    free(res->key); free(res->count); free(res->first); free(res->exc_ordinal); free(res->exc_bytes);
    res->key = nullptr; res->count = nullptr; res->first = nullptr; res->exc_ordinal = nullptr; res->exc_bytes = nullptr;
}
