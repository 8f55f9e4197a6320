// Multi-GPU bam2db inside the library: ONE host process drives G devices of a node through the C-ABI of
// include/fastf_gpu.h -- one feeder thread and one context per device -- and exchanges the locally deduplicated keys with an NCCL
// all-to-all (grouped ncclSend / ncclRecv over NVLink) partitioned by cell, so that every (cell, gene) group is counted on one GPU.
//
//   1. the BGZF blocks of the file are sharded contiguously, device by device (records do not straddle blocks in htslib files);
//      the BAM header lies in device 0's shard, later shards run "headerless";
//   2. every device inflates + parses its shard; the per-shard counts {records, CB-valid reads} give each shard the global ordinal
//      of its first CB-valid read = its position in the reference's single MT19937 draw sequence (reference src/bam2db_ds.c:385);
//   3. depth sampling on device at stream index d0 + base + local ordinal; the kept rows (table `umi`, read order) are read back
//      shard by shard when the caller wants them;
//   4. local sort + unique, grouped by destination = order-preserving range partition of the cell index
//      (fastf_unique_partition_device); ONE all-to-all of u64 keys;
//   5. local sort + run-length dedup / count per device; the COO pieces concatenate in device order (cells are range partitioned
//      and every piece is (cell, gene)-sorted), which is the single-GPU result and the reference's `mtx` table.
//
// This translation unit uses nothing but the public C-ABI, the CUDA runtime and NCCL.  NCCL is bound at run time (dlopen of
// libnccl.so.2: a process that already holds an NCCL -- e.g. one that imported torch -- keeps its own copy).  The emulator build
// (tests, no GPU) replaces the collective by host copies; everything else is the same code.
#include "../../include/fastf_gpu.h"
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#ifndef FASTF_EMU   // the emulator build needs no CUDA call here: its "device" memory is host memory
#include <cuda_runtime.h>
#include <dlfcn.h>
#endif

namespace {

#ifndef FASTF_EMU
// the few NCCL entry points of the exchange, declared as in nccl.h 2.x (stable C API)
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;      // ncclSuccess = 0
typedef int ncclDataType_t;    // ncclUint64 = 5
struct Nccl {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};
Nccl &nccl()
{
    static Nccl N;
    static bool tried = false;
    if (tried) return N;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { N.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return N; }
#define FASTF_NCCL_SYM(field, name) *(void **)(&N.field) = dlsym(h, name); if (!N.field) { N.why = std::string("libnccl lacks ") + name; return N; }
    FASTF_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    FASTF_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    FASTF_NCCL_SYM(GroupStart, "ncclGroupStart")
    FASTF_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    FASTF_NCCL_SYM(Send, "ncclSend")
    FASTF_NCCL_SYM(Recv, "ncclRecv")
    FASTF_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FASTF_NCCL_SYM
    N.ok = true;
    return N;
}
#endif

struct Shard {
    int device = 0;
    fastf_ctx *ctx = nullptr;
    fastf_bam2db_job *job = nullptr;
    size_t byte_lo = 0, byte_hi = 0;
    uint64_t n_rec = 0, n_cbv = 0, sampled = 0, valid = 0, n_kept = 0;
    uint64_t *kept = nullptr;              // device, owned by the job
    uint64_t *send = nullptr, *recv = nullptr;   // device
    uint32_t *coo[3] = {nullptr, nullptr, nullptr};
    std::vector<uint64_t> part;            // keys for each destination
    uint64_t n_send = 0, n_recv = 0, nnz = 0;
    std::vector<uint64_t> rows;            // kept rows of this shard, read order (want_rows)
    fastf_bam2db_result st;                // stage clocks / byte counts of this shard
    int rc = 0;
    std::string err;
};

void fail(Shard &s, const char *what)
{
    s.rc = 1;
    const char *e = s.ctx ? fastf_last_error(s.ctx) : fastf_last_error(nullptr);
    s.err = std::string(what) + ": " + (e ? e : "?");
}

// every shard does `fn` on its own host thread (the C-ABI calls block on their device)
template <class F> bool for_all(std::vector<Shard> &S, F fn)
{
#ifdef FASTF_EMU
    for (size_t g = 0; g < S.size(); g++) if (!S[g].rc) fn(S[g], g);   // the SIMT emulator runs one kernel at a time
#else
    std::vector<std::thread> th;
    for (size_t g = 1; g < S.size(); g++) th.emplace_back([&, g] { if (!S[g].rc) fn(S[g], g); });
    if (!S[0].rc) fn(S[0], 0);
    for (auto &t : th) t.join();
#endif
    for (auto &s : S) if (s.rc) return false;
    return true;
}

}   // namespace

static char g_sharded_err[768] = "";
extern "C" const char *fastf_sharded_last_error(void) { return g_sharded_err; }

extern "C" int fastf_bam2db_run_sharded(int n_devices, const int *devices, const fastf_bam2db_params *params, const void *bgzf_bytes, size_t n_bytes, fastf_bam2db_result *res)
{
    g_sharded_err[0] = 0;
    memset(res, 0, sizeof *res);
    if (n_devices < 1 || n_devices > 64) { snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: %d devices", n_devices); return 1; }
    const int G = n_devices;
    const uint8_t *bytes = (const uint8_t *)bgzf_bytes;
    std::vector<Shard> S((size_t)G);
    int rc = 1;
#ifndef FASTF_EMU
    std::vector<ncclComm_t> comms;
#endif
    auto cleanup = [&] {
        for (auto &s : S) {
            if (s.ctx) {
                if (s.send) fastf_device_free(s.ctx, s.send);
                if (s.recv) fastf_device_free(s.ctx, s.recv);
                for (auto &c : s.coo) if (c) fastf_device_free(s.ctx, c);
            }
            if (s.job) fastf_bam2db_job_free(s.job);
            if (s.ctx) fastf_ctx_destroy(s.ctx);
            s = Shard();
        }
#ifndef FASTF_EMU
        for (auto c : comms) if (c) nccl().CommDestroy(c);
        comms.clear();
#endif
    };
    auto first_error = [&] {
        for (auto &s : S) if (s.rc) { snprintf(g_sharded_err, sizeof g_sharded_err, "device %d: %s", s.device, s.err.c_str()); return; }
    };

    // ---- 1. block index (host) and contiguous block shards ----
    std::vector<uint64_t> in_off;
    std::vector<uint32_t> in_len, isz;
    {
        size_t cap = std::max<size_t>(1024, n_bytes / 8192 + 16), used = 0;   // BGZF blocks of BAM files hold ~20 KB; the index grows on demand
        int64_t nb;
        for (;;) {
            in_off.resize(cap); in_len.resize(cap); isz.resize(cap);
            nb = fastf_bgzf_index_host(bytes, n_bytes, in_off.data(), in_len.data(), isz.data(), cap, &used);
            if (nb != -2) break;
            cap *= 4;
        }
        if (nb < 0 || used != n_bytes) { snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: not a whole BGZF file (index stopped at byte %zu of %zu)", used, n_bytes); return 1; }
        in_off.resize((size_t)nb); in_len.resize((size_t)nb); isz.resize((size_t)nb);
    }
    const size_t nb = in_off.size();
    // a block starts where the previous one's trailer (CRC32 + ISIZE) ends, whatever its extra field holds
    auto block_start = [&](size_t i) -> size_t { return i == 0 ? 0 : (i < nb ? (size_t)(in_off[i - 1] + in_len[i - 1] + 8) : n_bytes); };
    for (int g = 0; g < G; g++) {
        const size_t base = nb / (size_t)G, rem = nb % (size_t)G;
        const size_t lo = base * (size_t)g + std::min<size_t>((size_t)g, rem), hi = lo + base + ((size_t)g < rem ? 1 : 0);
        S[(size_t)g].device = devices ? devices[g] : g;
        S[(size_t)g].byte_lo = block_start(lo);
        S[(size_t)g].byte_hi = block_start(hi);
    }

    // ---- 2. contexts, jobs, feed (one host thread per device) ----
    const size_t PIECE = (size_t)128 << 20;
    for_all(S, [&](Shard &s, size_t g) {
        if (fastf_ctx_create(s.device, &s.ctx)) { fail(s, "ctx_create"); return; }
        fastf_bam2db_params p = *params;
        p.headerless = g != 0;
        p.want_rows = 0;   // rows are read back from the device below, before the exchange reorders them
        if (fastf_bam2db_begin(s.ctx, &p, &s.job)) { fail(s, "begin"); return; }
        void *pin[2] = {nullptr, nullptr};
        if (fastf_host_alloc(s.ctx, PIECE, &pin[0]) || fastf_host_alloc(s.ctx, PIECE, &pin[1])) { fail(s, "host_alloc"); return; }
        int which = 0;
        for (size_t at = s.byte_lo; at < s.byte_hi && !s.rc; at += PIECE, which ^= 1) {
            const size_t n = std::min(PIECE, s.byte_hi - at);
            memcpy(pin[which], bytes + at, n);   // page cache / caller's buffer -> pinned ring
            if (fastf_bam2db_feed(s.job, pin[which], n)) fail(s, "feed");
        }
        if (!s.rc && fastf_bam2db_counts(s.job, &s.n_rec, &s.n_cbv)) fail(s, "counts");
        fastf_host_free(s.ctx, pin[0]);
        fastf_host_free(s.ctx, pin[1]);
    });
    bool ok = true;
    for (auto &s : S) ok = ok && !s.rc;
    if (!ok) { first_error(); cleanup(); return 1; }

    // ---- 3. global draw ordinals, sampling, local unique + partition ----
    uint32_t bits_cell = 0, bits_gene = 0, bits_umi = 0;
    {
        uint64_t base = 0;
        std::vector<uint64_t> bases((size_t)G);
        for (int g = 0; g < G; g++) { bases[(size_t)g] = base; base += S[(size_t)g].n_cbv; }
        ok = for_all(S, [&](Shard &s, size_t g) {
            if (fastf_bam2db_sample(s.job, bases[g])) { fail(s, "sample"); return; }
            if (fastf_bam2db_sample_counts(s.job, &s.sampled, &s.valid) || fastf_bam2db_kept_device(s.job, &s.kept, &s.n_kept)) { fail(s, "kept"); return; }
            uint32_t bc, bg, bu;
            if (fastf_bam2db_key_layout(s.job, &bc, &bg, &bu)) { fail(s, "key_layout"); return; }
            if (g == 0) { bits_cell = bc; bits_gene = bg; bits_umi = bu; }
            if (params->want_rows && s.n_kept) {
                s.rows.resize((size_t)s.n_kept);
                if (fastf_memcpy_d2h(s.ctx, s.rows.data(), s.kept, (size_t)s.n_kept * 8)) { fail(s, "rows d2h"); return; }
            }
            void *d = nullptr;
            if (fastf_device_alloc(s.ctx, (size_t)std::max<uint64_t>(s.n_kept, 1) * 8, &d)) { fail(s, "device_alloc"); return; }
            s.send = (uint64_t *)d;
            s.part.assign((size_t)G, 0);
            if (fastf_unique_partition_device(s.ctx, s.kept, s.n_kept, bc + bg + bu, bg, bu, params->n_cells, (uint32_t)G, s.send, s.part.data())) { fail(s, "unique_partition"); return; }
            s.n_send = 0;
            for (auto c : s.part) s.n_send += c;
        });
    }
    if (!ok) { first_error(); cleanup(); return 1; }
    const uint32_t key_bits = bits_cell + bits_gene + bits_umi;

    // ---- 4. the all-to-all ----
    for (int g = 0; g < G; g++) {
        Shard &s = S[(size_t)g];
        s.n_recv = 0;
        for (int p = 0; p < G; p++) s.n_recv += S[(size_t)p].part[(size_t)g];
    }
    ok = for_all(S, [&](Shard &s, size_t) {
        void *d = nullptr;
        if (fastf_device_alloc(s.ctx, (size_t)std::max<uint64_t>(s.n_recv, 1) * 8, &d)) { fail(s, "device_alloc"); return; }
        s.recv = (uint64_t *)d;
        for (auto &c : s.coo) {
            if (fastf_device_alloc(s.ctx, (size_t)std::max<uint64_t>(s.n_recv, 1) * 4, &d)) { fail(s, "device_alloc"); return; }
            c = (uint32_t *)d;
        }
    });
    if (!ok) { first_error(); cleanup(); return 1; }
#ifdef FASTF_EMU
    for (int g = 0; g < G; g++) {   // "device" memory is host memory here
        uint64_t ro = 0;
        for (int p = 0; p < G; p++) {
            uint64_t so = 0;
            for (int q = 0; q < g; q++) so += S[(size_t)p].part[(size_t)q];
            const uint64_t c = S[(size_t)p].part[(size_t)g];
            memcpy(S[(size_t)g].recv + ro, S[(size_t)p].send + so, (size_t)c * 8);
            ro += c;
        }
    }
#else
    if (G > 1) {
        Nccl &N = nccl();
        if (!N.ok) { snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: NCCL unavailable (%s)", N.why.c_str()); cleanup(); return 1; }
        comms.assign((size_t)G, nullptr);
        std::vector<int> devs((size_t)G);
        for (int g = 0; g < G; g++) devs[(size_t)g] = S[(size_t)g].device;
        ncclResult_t r = N.CommInitAll(comms.data(), G, devs.data());
        if (r) { snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: ncclCommInitAll: %s", N.GetErrorString(r)); cleanup(); return 1; }
        r = N.GroupStart();
        for (int g = 0; g < G && !r; g++) {
            Shard &s = S[(size_t)g];
            cudaStream_t st = (cudaStream_t)fastf_compute_stream(s.ctx);
            uint64_t so = 0, ro = 0;
            for (int p = 0; p < G && !r; p++) {
                const uint64_t cs = s.part[(size_t)p], cr = S[(size_t)p].part[(size_t)g];
                if (cs) r = N.Send(s.send + so, (size_t)cs, 5 /* ncclUint64 */, p, comms[(size_t)g], st);
                if (cr && !r) r = N.Recv(s.recv + ro, (size_t)cr, 5, p, comms[(size_t)g], st);
                so += cs; ro += cr;
            }
        }
        ncclResult_t r2 = N.GroupEnd();
        if (r || r2) { snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: NCCL all-to-all: %s", N.GetErrorString(r ? r : r2)); cleanup(); return 1; }
        for (auto &s : S) { cudaSetDevice(s.device); if (fastf_synchronize(s.ctx)) { fail(s, "synchronize after the exchange"); first_error(); cleanup(); return 1; } }
    } else if (S[0].n_recv) {
        cudaSetDevice(S[0].device);
        if (cudaMemcpy(S[0].recv, S[0].send, (size_t)S[0].n_recv * 8, cudaMemcpyDeviceToDevice) != cudaSuccess) {
            snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: device copy failed");
            cleanup();
            return 1;
        }
    }
#endif

    // ---- 5. local sort + dedup / count; COO pieces back to the host ----
    std::vector<std::vector<uint32_t>> piece[3];
    for (auto &v : piece) v.resize((size_t)G);
    ok = for_all(S, [&](Shard &s, size_t g) {
        if (s.n_recv && fastf_sort_u64_device(s.ctx, s.recv, nullptr, s.n_recv, key_bits)) { fail(s, "sort"); return; }
        if (fastf_dedup_count_device_out(s.ctx, s.recv, s.n_recv, bits_gene, bits_umi, &s.nnz, s.coo[0], s.coo[1], s.coo[2])) { fail(s, "dedup_count"); return; }
        for (int c = 0; c < 3; c++) {
            piece[c][g].resize((size_t)s.nnz);
            if (s.nnz && fastf_memcpy_d2h(s.ctx, piece[c][g].data(), s.coo[c], (size_t)s.nnz * 4)) { fail(s, "coo d2h"); return; }
        }
        memset(&s.st, 0, sizeof s.st);
        if (fastf_bam2db_stats(s.job, &s.st)) { fail(s, "stats"); return; }
    });
    if (!ok) { first_error(); cleanup(); return 1; }

    // ---- 6. the result, laid out like fastf_bam2db_finish's ----
    {
        uint64_t nnz = 0, n_rows = 0;
        for (auto &s : S) {
            res->total += s.n_rec; res->cb_valid += s.n_cbv; res->sampled += s.sampled; res->valid += s.valid;
            nnz += s.nnz; n_rows += s.rows.size();
            res->n_blocks += s.st.n_blocks; res->compressed_bytes += s.st.compressed_bytes; res->inflated_bytes += s.st.inflated_bytes;
            res->status |= s.st.status; res->n_launches += s.st.n_launches; res->n_chunks += s.st.n_chunks;
            res->ms_inflate = std::max(res->ms_inflate, s.st.ms_inflate); res->ms_parse = std::max(res->ms_parse, s.st.ms_parse); res->ms_crc = std::max(res->ms_crc, s.st.ms_crc);
            res->ms_mt = std::max(res->ms_mt, s.st.ms_mt); res->ms_sample = std::max(res->ms_sample, s.st.ms_sample); res->ms_device_total = std::max(res->ms_device_total, s.st.ms_device_total);
        }
        res->nnz = nnz;
        res->bits_cell = bits_cell; res->bits_gene = bits_gene; res->bits_umi = bits_umi; res->umi_max_bytes = (bits_umi - 4) / 8;
        res->m_gene = (uint32_t *)malloc(std::max<uint64_t>(nnz, 1) * 4);
        res->m_cell = (uint32_t *)malloc(std::max<uint64_t>(nnz, 1) * 4);
        res->m_count = (uint32_t *)malloc(std::max<uint64_t>(nnz, 1) * 4);
        res->row_keys = params->want_rows ? (uint64_t *)malloc(std::max<uint64_t>(n_rows, 1) * 8) : nullptr;
        if (!res->m_gene || !res->m_cell || !res->m_count || (params->want_rows && !res->row_keys)) {
            snprintf(g_sharded_err, sizeof g_sharded_err, "run_sharded: out of host memory");
            fastf_bam2db_result_free(res);
            cleanup();
            return 1;
        }
        uint64_t o = 0, ro = 0;
        uint32_t *dst[3] = {res->m_gene, res->m_cell, res->m_count};
        for (int g = 0; g < G; g++) {
            for (int c = 0; c < 3; c++) memcpy(dst[c] + o, piece[c][(size_t)g].data(), (size_t)S[(size_t)g].nnz * 4);
            o += S[(size_t)g].nnz;
            if (params->want_rows) { memcpy(res->row_keys + ro, S[(size_t)g].rows.data(), S[(size_t)g].rows.size() * 8); ro += S[(size_t)g].rows.size(); }
        }
        res->n_rows = params->want_rows ? n_rows : 0;
        rc = 0;
    }
    // exchanged keys (diagnostic) ride in ms_sort's neighbour: callers read fastf_sharded_exchanged()
    {
        uint64_t x = 0;
        for (auto &s : S) x += s.n_send;
        extern std::atomic<uint64_t> g_fastf_sharded_exchanged;
        g_fastf_sharded_exchanged = x;
    }
    cleanup();
    return rc;
}

std::atomic<uint64_t> g_fastf_sharded_exchanged{0};
extern "C" uint64_t fastf_sharded_exchanged(void) { return g_fastf_sharded_exchanged.load(); }
